// hevcb_sort.cu -- order in which the parser threads take the NALs: stable radix sort (CUB) of a 32-bit shape key, so
// that the 32 NALs of a warp are of the same type and begin alike (same first payload bytes) and walk the same branches.
// Kept in its own translation unit (hevcb_parse.cu is built with different optimiser flags).
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <stdint.h>

#include "hevcb_internal.h"

namespace {
__global__ void iota_kernel(int32_t* v, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { v[i] = (int32_t)i; }
}
} // namespace

// keys[n] (device) -> perm[n] (device): perm[i] = index of the NAL with the i-th smallest key, ties in stream order
int hevcb_sort_perm(hevcb_ctx* ctx, const uint32_t* d_keys, int64_t n, int32_t* d_perm, cudaStream_t stream)
{
    if (n <= 0) { return HEVCB_OK; }
    if (n > 0x7FFFFFFFll) {
        HEVCB_SET_ERR(ctx, "hevcb_parse: more than 2^31 NALs in one call");
        return HEVCB_E_ARG;
    }
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, d_keys, (uint32_t*)nullptr, (const int32_t*)nullptr, d_perm, (int)n, 0, 32, stream);
    const size_t keys_out = ((size_t)n * 4 + 255) & ~(size_t)255, idx_in = keys_out;
    int rc = hevcb_reserve(ctx, &ctx->parse_sort, keys_out + idx_in + temp_bytes + 256);
    if (rc != HEVCB_OK) { return rc; }
    uint8_t* b = reinterpret_cast<uint8_t*>(ctx->parse_sort.p);
    uint32_t* d_keys_out = reinterpret_cast<uint32_t*>(b);
    int32_t* d_idx = reinterpret_cast<int32_t*>(b + keys_out);
    void* d_temp = b + keys_out + idx_in;
    iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_idx, n);
    HEVCB_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_keys, d_keys_out, (const int32_t*)d_idx, d_perm, (int)n, 0, 32, stream));
    ctx->launches += 4; // iota + the sort's passes (histogram, scan, onesweep x N are counted as three)
    return HEVCB_OK;
}
