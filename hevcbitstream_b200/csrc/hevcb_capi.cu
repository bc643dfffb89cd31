// hevcb_capi.cu -- C ABI of libhevcb200 (include/hevcb.h): context lifecycle and the host-buffer
// wrappers around the device entry points.  No CPU implementation exists behind any entry point.
#include <cuda_runtime.h>
#include <new>
#include <stdint.h>
#include <stdlib.h>

#include "hevcb_internal.h"

char g_hevcb_create_error[512] = {0};

extern "C" {

HEVCB_API int hevcb_version(void) { return HEVCB_VERSION_MAJOR * 1000 + HEVCB_VERSION_MINOR; }

HEVCB_API int hevcb_create(int device, hevcb_ctx** out)
{
    if (!out) { return HEVCB_E_ARG; }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        snprintf(g_hevcb_create_error, sizeof(g_hevcb_create_error),
                 "no usable CUDA device (%s); libhevcb200 has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return HEVCB_E_NODEVICE;
    }
    if (device < 0 || device >= count) {
        snprintf(g_hevcb_create_error, sizeof(g_hevcb_create_error), "device %d out of range (0..%d)", device, count - 1);
        return HEVCB_E_ARG;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        snprintf(g_hevcb_create_error, sizeof(g_hevcb_create_error), "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        return HEVCB_E_NODEVICE;
    }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        snprintf(g_hevcb_create_error, sizeof(g_hevcb_create_error), "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return HEVCB_E_NODEVICE;
    }
    if (prop.major < 10) {
        snprintf(g_hevcb_create_error, sizeof(g_hevcb_create_error),
                 "device %d is sm_%d%d; libhevcb200 is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return HEVCB_E_NODEVICE;
    }
    hevcb_ctx* ctx = new (std::nothrow) hevcb_ctx();
    if (!ctx) { return HEVCB_E_NOMEM; }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        snprintf(g_hevcb_create_error, sizeof(g_hevcb_create_error), "cudaStreamCreate: %s", cudaGetErrorString(e));
        delete ctx;
        return HEVCB_E_CUDA;
    }
    ctx->pinned_bytes = 4096;
    e = cudaMallocHost(&ctx->pinned, ctx->pinned_bytes);
    if (e != cudaSuccess) {
        snprintf(g_hevcb_create_error, sizeof(g_hevcb_create_error), "cudaMallocHost: %s", cudaGetErrorString(e));
        cudaStreamDestroy(ctx->stream);
        delete ctx;
        return HEVCB_E_CUDA;
    }
    if (const char* e2 = getenv("HEVCB_SCAN_DEBUG")) { ctx->scan_debug_flags = atoll(e2); } // kernel experiment switches
    *out = ctx;
    return HEVCB_OK;
}

HEVCB_API void hevcb_destroy(hevcb_ctx* ctx)
{
    if (!ctx) { return; }
    cudaSetDevice(ctx->device);
    hevcb_devbuf* bufs[] = {&ctx->scan_scratch, &ctx->h_in, &ctx->h_rbsp, &ctx->h_a0, &ctx->h_a1, &ctx->h_a2, &ctx->h_a3, &ctx->h_misc};
    for (hevcb_devbuf* b : bufs) {
        if (b->p) { cudaFree(b->p); }
    }
    if (ctx->pinned) { cudaFreeHost(ctx->pinned); }
    if (ctx->stream) { cudaStreamDestroy(ctx->stream); }
    delete ctx;
}

HEVCB_API const char* hevcb_last_error(const hevcb_ctx* ctx) { return ctx ? ctx->err : g_hevcb_create_error; }
HEVCB_API int64_t hevcb_launch_count(const hevcb_ctx* ctx) { return ctx ? ctx->launches : 0; }
HEVCB_API int hevcb_sm_count(const hevcb_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

HEVCB_API int hevcb_scan_strip_device(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, int64_t* d_nal_start, int64_t* d_nal_end,
                                      int64_t cap_nals, uint8_t* d_rbsp, int64_t* d_rbsp_off, int64_t* d_rbsp_end,
                                      hevcb_scan_summary* d_summary, void* stream)
{
    if (!ctx) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    return hevcb_launch_scan_strip(ctx, d_buf, size, d_nal_start, d_nal_end, cap_nals, d_rbsp, d_rbsp_off, d_rbsp_end, d_summary,
                                   (cudaStream_t)stream);
}

HEVCB_API int hevcb_scan_strip_host(hevcb_ctx* ctx, const uint8_t* buf, int64_t size, int64_t* nal_start, int64_t* nal_end,
                                    int64_t cap_nals, uint8_t* rbsp, int64_t* rbsp_off, int64_t* rbsp_end, hevcb_scan_summary* summary)
{
    if (!ctx || !summary || size < 0 || cap_nals < 0 || !nal_start || !nal_end || (size > 0 && !buf)) {
        HEVCB_SET_ERR(ctx, "hevcb_scan_strip_host: invalid argument");
        return HEVCB_E_ARG;
    }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t in_bytes = ((size_t)size + 15u) & ~(size_t)15u;
    const size_t arr_bytes = (size_t)(cap_nals > 0 ? cap_nals : 1) * sizeof(int64_t);
    int rc;
    if ((rc = hevcb_reserve(ctx, &ctx->h_in, in_bytes + 16)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a0, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a1, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a2, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a3, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_misc, 256)) != HEVCB_OK) { return rc; }
    if (rbsp) {
        if ((rc = hevcb_reserve(ctx, &ctx->h_rbsp, in_bytes + 16)) != HEVCB_OK) { return rc; }
    }
    if (size > 0) { HEVCB_CUDA(ctx, cudaMemcpyAsync(ctx->h_in.p, buf, (size_t)size, cudaMemcpyHostToDevice, st)); }
    hevcb_scan_summary* d_sum = reinterpret_cast<hevcb_scan_summary*>(ctx->h_misc.p);
    rc = hevcb_launch_scan_strip(ctx, reinterpret_cast<const uint8_t*>(ctx->h_in.p), size, reinterpret_cast<int64_t*>(ctx->h_a0.p),
                                 reinterpret_cast<int64_t*>(ctx->h_a1.p), cap_nals, rbsp ? reinterpret_cast<uint8_t*>(ctx->h_rbsp.p) : nullptr,
                                 reinterpret_cast<int64_t*>(ctx->h_a2.p), reinterpret_cast<int64_t*>(ctx->h_a3.p), d_sum, st);
    if (rc != HEVCB_OK) { return rc; }
    hevcb_scan_summary* p_sum = reinterpret_cast<hevcb_scan_summary*>(ctx->pinned);
    HEVCB_CUDA(ctx, cudaMemcpyAsync(p_sum, d_sum, sizeof(hevcb_scan_summary), cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    *summary = *p_sum;
    int64_t n = summary->n_nals < cap_nals ? summary->n_nals : cap_nals;
    if (n > 0) {
        HEVCB_CUDA(ctx, cudaMemcpyAsync(nal_start, ctx->h_a0.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(nal_end, ctx->h_a1.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        if (rbsp_off) { HEVCB_CUDA(ctx, cudaMemcpyAsync(rbsp_off, ctx->h_a2.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st)); }
        if (rbsp_end) { HEVCB_CUDA(ctx, cudaMemcpyAsync(rbsp_end, ctx->h_a3.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st)); }
    }
    if (rbsp && summary->rbsp_bytes > 0) {
        HEVCB_CUDA(ctx, cudaMemcpyAsync(rbsp, ctx->h_rbsp.p, (size_t)summary->rbsp_bytes, cudaMemcpyDeviceToHost, st));
    }
    HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    if (summary->overflow) {
        HEVCB_SET_ERR(ctx, "hevcb_scan_strip_host: %lld NALs exceed cap_nals %lld", (long long)summary->n_nals, (long long)cap_nals);
        return HEVCB_E_CAPACITY;
    }
    return HEVCB_OK;
}

} // extern "C"
