// micro-benchmark: what a read-once / write-once byte stream can reach on this GPU with the load / store schemes the
// scan + strip kernel could be built from.  Prints GB/s of INPUT per variant; a copy moves 2x that.
//   read_ldg        grid of small CTAs, 16-byte loads, U rows per thread in flight, XOR-reduced (read ceiling)
//   copy_ldg        the same + 16-byte streaming stores (copy ceiling in our own code)
//   copy_tma_ring   persistent CTAs, cp.async.bulk tiles into an S-stage shared-memory ring, workers LDS -> STG
//   copy_tma_bulk   the same, stores by cp.async.bulk shared -> global
//   chain_copy      register-resident tiles + decoupled look-back (single pass chained scan): every tile's output
//                   offset depends on the tiles before it, the SWAR zero-pair test runs on every chunk
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o stream_ceiling stream_ceiling.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ldg_stream(const uint4* p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// ---------------------------------------------------------------------------------------------- read / copy, plain
template <int U, bool kStore>
__global__ void __launch_bounds__(256) k_ldg(const uint4* __restrict__ in, uint4* __restrict__ out, size_t nvec, unsigned* sink)
{
    // tile = 256 threads x U vectors, consecutive rows of 4 KiB
    const size_t tile = (size_t)blockIdx.x * (256 * U);
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; u++) {
        const size_t i = tile + (size_t)u * 256 + threadIdx.x;
        v[u] = (i < nvec) ? ldg_stream(in + i) : make_uint4(0, 0, 0, 0);
    }
    if (kStore) {
#pragma unroll
        for (int u = 0; u < U; u++) {
            const size_t i = tile + (size_t)u * 256 + threadIdx.x;
            if (i < nvec) { __stcs(out + i, v[u]); }
        }
    } else {
        unsigned x = 0;
#pragma unroll
        for (int u = 0; u < U; u++) { x ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w; }
        if (x == 0x12345678u) { *sink = x; }
    }
}

// ---------------------------------------------------------------------------------------------- TMA ring
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_s2g(void* dst, const void* src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// mode 0: workers read the stage (LDS.128) and store with STG.128; mode 1: bulk store shared -> global; mode 2: read only (LDS + xor)
template <int kTile, int kStages, int kMode>
__global__ void __launch_bounds__(288) k_tma_ring(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long n_tiles, unsigned* sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + (size_t)kTile * kStages);
    unsigned long long* empty = full + kStages;
    const int tid = threadIdx.x, warp = tid >> 5;
    const long long G = gridDim.x, first = blockIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kMode == 1 ? 1 : 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == 8) { // producer warp
        if ((tid & 31) == 0) {
            long long it = 0;
            for (long long t = first; t < n_tiles; t += G, it++) {
                const int s = (int)(it % kStages);
                const uint32_t round = (uint32_t)(it / kStages);
                if (round >= 1) { while (!mbar_try_wait(&empty[s], (round - 1) & 1u)) {} }
                mbar_expect_tx(&full[s], kTile);
                tma_g2s(smem + (size_t)s * kTile, in + (size_t)t * kTile, kTile, &full[s]);
            }
        }
        return;
    }
    // consumers (8 warps)
    int s = 0; uint32_t ph = 0; unsigned x = 0;
    for (long long t = first; t < n_tiles; t += G) {
        while (!mbar_try_wait(&full[s], ph & 1u)) {}
        const uint8_t* st = smem + (size_t)s * kTile;
        if (kMode == 1) {
            if (tid == 0) {
                tma_s2g(out + (size_t)t * kTile, st, kTile);
                tma_commit();
                tma_wait_read<0>(); // the stage has been read: it may be reloaded
                mbar_arrive(&empty[s]);
            }
        } else {
#pragma unroll 4
            for (int i = tid * 16; i < kTile; i += 256 * 16) {
                const uint4 v = *reinterpret_cast<const uint4*>(st + i);
                if (kMode == 0) { __stcs(reinterpret_cast<uint4*>(out + (size_t)t * kTile + i), v); }
                else { x ^= v.x ^ v.y ^ v.z ^ v.w; }
            }
            __syncwarp();
            if ((tid & 31) == 0) { mbar_arrive(&empty[s]); }
        }
        if (s + 1 == kStages) { s = 0; ph ^= 1u; } else { s++; }
    }
    if (kMode == 2 && x == 0x12345678u) { *sink = x; }
}

// ---------------------------------------------------------------------------------------------- chained scan copy
// state word: bits 63..62 status (1 aggregate, 2 inclusive prefix), low bits value
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

__device__ __forceinline__ uint32_t zero_pair_own(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wn)
{
    // pairs (j, j+1) of zero bytes starting inside the chunk: byte j of (w | w>>8) is zero iff both are zero
    const uint32_t m0 = w0 | __funnelshift_r(w0, w1, 8), m1 = w1 | __funnelshift_r(w1, w2, 8), m2 = w2 | __funnelshift_r(w2, w3, 8),
                   m3 = w3 | __funnelshift_r(w3, wn, 8);
    const uint32_t c = 0x01010101u;
    return (((m0 - c) & ~m0) | ((m1 - c) & ~m1) | ((m2 - c) & ~m2) | ((m3 - c) & ~m3)) & 0x80808080u;
}

// kRows rows of 512 B per warp, 8 warps: tile = kRows * 4 KiB.  kDyn: tile index from an atomic counter.
template <int kRows, int kCtasPerSm>
__global__ void __launch_bounds__(256, kCtasPerSm) k_chain_copy(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long n_tiles,
                                                                unsigned long long* __restrict__ state, unsigned* __restrict__ counter, int shift)
{
    constexpr int kTile = kRows * 4096;
    __shared__ unsigned long long s_tile;
    __shared__ uint32_t s_wagg[8];
    __shared__ unsigned long long s_excl;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_tile = atomicAdd(counter, 1u); }
    __syncthreads();
    const long long t = (long long)s_tile;
    if (t >= n_tiles) { return; }
    const uint8_t* base = in + (size_t)t * kTile + (size_t)warp * (kRows * 512) + lane * 16;
    uint4 v[kRows];
#pragma unroll
    for (int r = 0; r < kRows; r++) { v[r] = ldg_stream(reinterpret_cast<const uint4*>(base + r * 512)); }
    // analysis: zero pairs starting in the lane's chunk (needs the first word of the next chunk)
    uint32_t slow = 0;
#pragma unroll
    for (int r = 0; r < kRows; r++) {
        uint32_t wn = __shfl_down_sync(0xFFFFFFFFu, v[r].x, 1);
        const uint32_t nx = (r + 1 < kRows) ? v[r + 1].x : 0x01010101u;
        const uint32_t n0 = __shfl_sync(0xFFFFFFFFu, nx, 0);
        if (lane == 31) { wn = n0; }
        slow |= (zero_pair_own(v[r].x, v[r].y, v[r].z, v[r].w, wn) != 0u ? 1u : 0u) << r;
    }
    const uint32_t anyslow = __reduce_or_sync(0xFFFFFFFFu, slow);
    uint32_t wdel = __popc(anyslow) & (uint32_t)shift; // "removed bytes" of the warp (0 unless shift is set: then misaligned output)
    if (lane == 0) { s_wagg[warp] = wdel; }
    __syncthreads();
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) { const uint32_t a = s_wagg[w]; if (w < warp) { before += a; } total += a; }
    if (warp == 0) {
        unsigned long long excl = 0;
        if (t == 0) {
            if (lane == 0) { st_relaxed(&state[0], (2ull << 62) | (unsigned long long)total); }
        } else {
            if (lane == 0) { st_relaxed(&state[t], (1ull << 62) | (unsigned long long)total); }
            long long look = t - 1;
            for (;;) {
                const long long idx = look - lane;
                unsigned long long sv = (idx >= 0) ? ld_relaxed(&state[idx]) : (2ull << 62);
                // all 32 must be published
                while (__any_sync(0xFFFFFFFFu, (sv >> 62) == 0ull)) {
                    if ((sv >> 62) == 0ull) { sv = ld_relaxed(&state[idx]); }
                }
                const uint32_t pm = __ballot_sync(0xFFFFFFFFu, (sv >> 62) == 2ull);
                const int firstp = pm ? (__ffs((int)pm) - 1) : 32;
                unsigned long long contrib = (lane <= firstp) ? (sv & ((1ull << 62) - 1)) : 0ull;
#pragma unroll
                for (int d = 16; d >= 1; d >>= 1) { contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, d); }
                excl += contrib;
                if (pm) { break; }
                look -= 32;
            }
            if (lane == 0) { st_relaxed(&state[t], (2ull << 62) | (excl + total)); }
        }
        if (lane == 0) { s_excl = excl; }
    }
    __syncthreads();
    const unsigned long long excl = s_excl + before;
    // write: destination = position - removed bytes before; aligned 16-byte stores built from the lane's and its left neighbour's vector
    const size_t srcpos = (size_t)t * kTile + (size_t)warp * (kRows * 512);
    const size_t dstpos = srcpos - excl;
    const uint32_t mis = (uint32_t)(excl & 15u); // output vector q (aligned) starts at source byte q*16 + mis of ... shifted right
    uint8_t* o = out + dstpos;
    if (mis == 0) {
#pragma unroll
        for (int r = 0; r < kRows; r++) { __stcs(reinterpret_cast<uint4*>(o + r * 512 + lane * 16), v[r]); }
    } else {
        // aligned destination vector = bytes [mis, mis + 16) of (previous chunk : own chunk)
        const uint32_t q = mis >> 2, sh = (mis & 3u) * 8u;
        uint8_t* oa = reinterpret_cast<uint8_t*>(reinterpret_cast<uintptr_t>(o) & ~(uintptr_t)15);
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            // previous chunk: lane-1 of the same row, or lane 31 of the previous row
            uint4 p;
            p.x = __shfl_up_sync(0xFFFFFFFFu, v[r].x, 1); p.y = __shfl_up_sync(0xFFFFFFFFu, v[r].y, 1);
            p.z = __shfl_up_sync(0xFFFFFFFFu, v[r].z, 1); p.w = __shfl_up_sync(0xFFFFFFFFu, v[r].w, 1);
            if (r > 0) {
                const uint4 pr = v[r - 1];
                const uint32_t a = __shfl_sync(0xFFFFFFFFu, pr.x, 31), b = __shfl_sync(0xFFFFFFFFu, pr.y, 31), c = __shfl_sync(0xFFFFFFFFu, pr.z, 31),
                               d = __shfl_sync(0xFFFFFFFFu, pr.w, 31);
                if (lane == 0) { p = make_uint4(a, b, c, d); }
            }
            const uint32_t W[8] = {p.x, p.y, p.z, p.w, v[r].x, v[r].y, v[r].z, v[r].w};
            uint32_t xw[5];
#pragma unroll
            for (int e = 0; e < 5; e++) { xw[e] = (q == 0u) ? W[e] : (q == 1u) ? W[e + 1] : (q == 2u) ? W[e + 2] : W[e + 3]; }
            uint4 o4;
            o4.x = __funnelshift_r(xw[0], xw[1], sh); o4.y = __funnelshift_r(xw[1], xw[2], sh);
            o4.z = __funnelshift_r(xw[2], xw[3], sh); o4.w = __funnelshift_r(xw[3], xw[4], sh);
            if (!(r == 0 && lane == 0)) { __stcs(reinterpret_cast<uint4*>(oa + r * 512 + lane * 16), o4); }
        }
    }
}


// ---------------------------------------------------------------------------------------------- persistent chained copy
// Persistent CTAs (all resident), tile t = blockIdx.x + i * gridDim.x.  The loads of the CTA's next tile are issued before the
// look-back of the current one (second register set).  Look-back window = 32 lanes x kW 16-byte states per round trip.
__device__ __forceinline__ ulonglong2 ld_state16(const ulonglong2* p)
{
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state16(ulonglong2* p, ulonglong2 v) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory"); }

template <int kRows, int kW, int kCtasPerSm>
__global__ void __launch_bounds__(256, kCtasPerSm) k_chain_persist(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, long long n_tiles,
                                                                   ulonglong2* __restrict__ state, int shift)
{
    constexpr int kTile = kRows * 4096;
    __shared__ uint32_t s_wagg[2][8];
    __shared__ unsigned long long s_excl[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long G = gridDim.x;
    uint4 v[kRows], nx[kRows];
    long long t = blockIdx.x;
    if (t < n_tiles) {
        const uint8_t* base = in + (size_t)t * kTile + (size_t)warp * (kRows * 512) + lane * 16;
#pragma unroll
        for (int r = 0; r < kRows; r++) { v[r] = ldg_stream(reinterpret_cast<const uint4*>(base + r * 512)); }
    }
    for (int it = 0; t < n_tiles; t += G, it++) {
        const long long tn = t + G;
        if (tn < n_tiles) {
            const uint8_t* base = in + (size_t)tn * kTile + (size_t)warp * (kRows * 512) + lane * 16;
#pragma unroll
            for (int r = 0; r < kRows; r++) { nx[r] = ldg_stream(reinterpret_cast<const uint4*>(base + r * 512)); }
        }
        uint32_t slow = 0;
#pragma unroll
        for (int r = 0; r < kRows; r++) {
            uint32_t wn = __shfl_down_sync(0xFFFFFFFFu, v[r].x, 1);
            const uint32_t nxw = (r + 1 < kRows) ? v[r + 1].x : 0x01010101u;
            const uint32_t n0 = __shfl_sync(0xFFFFFFFFu, nxw, 0);
            if (lane == 31) { wn = n0; }
            slow |= (zero_pair_own(v[r].x, v[r].y, v[r].z, v[r].w, wn) != 0u ? 1u : 0u) << r;
        }
        const uint32_t anyslow = __reduce_or_sync(0xFFFFFFFFu, slow);
        const uint32_t wdel = __popc(anyslow) & (uint32_t)shift;
        const int pb = it & 1;
        if (lane == 0) { s_wagg[pb][warp] = wdel; }
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) { const uint32_t a = s_wagg[pb][w]; if (w < warp) { before += a; } total += a; }
        if (warp == 0) {
            unsigned long long excl = 0;
            if (t == 0) {
                if (lane == 0) { st_state16(&state[0], make_ulonglong2((2ull << 62) | (unsigned long long)total, 0ull)); }
            } else {
                if (lane == 0) { st_state16(&state[t], make_ulonglong2((1ull << 62) | (unsigned long long)total, 0ull)); }
                long long look = t - 1;
                for (;;) {
                    // lane l looks at states look - (l * kW + j): distance d = l * kW + j, descending stream order
                    ulonglong2 sv[kW];
#pragma unroll
                    for (int j = 0; j < kW; j++) {
                        const long long idx = look - ((long long)lane * kW + j);
                        sv[j] = (idx >= 0) ? ld_state16(&state[idx]) : make_ulonglong2(2ull << 62, 0ull);
                    }
                    uint32_t dp_w, dm_w;
                    for (;;) {
                        // nearest prefix and nearest unpublished state of the window; only what lies in front of the nearest prefix matters
                        uint32_t dp = 0xFFFFu, dm = 0xFFFFu;
#pragma unroll
                        for (int j = kW - 1; j >= 0; j--) {
                            const uint32_t stt = (uint32_t)(sv[j].x >> 62);
                            if (stt == 2u) { dp = (uint32_t)(lane * kW + j); }
                            if (stt == 0u) { dm = (uint32_t)(lane * kW + j); }
                        }
                        dp_w = __reduce_min_sync(0xFFFFFFFFu, dp);
                        dm_w = __reduce_min_sync(0xFFFFFFFFu, dm);
                        if (dm_w > dp_w || dm_w == 0xFFFFu) { break; }
#pragma unroll
                        for (int j = 0; j < kW; j++) {
                            const uint32_t d = (uint32_t)(lane * kW + j);
                            if ((sv[j].x >> 62) == 0ull && d < dp_w) { sv[j] = ld_state16(&state[look - (long long)d]); }
                        }
                    }
                    unsigned long long contrib = 0;
#pragma unroll
                    for (int j = 0; j < kW; j++) {
                        if ((uint32_t)(lane * kW + j) <= dp_w) { contrib += sv[j].x & ((1ull << 62) - 1); }
                    }
#pragma unroll
                    for (int d = 16; d >= 1; d >>= 1) { contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, d); }
                    excl += contrib;
                    if (dp_w != 0xFFFFu) { break; }
                    look -= 32 * kW;
                }
                if (lane == 0) { st_state16(&state[t], make_ulonglong2((2ull << 62) | (excl + total), 0ull)); }
            }
            if (lane == 0) { s_excl[pb] = excl; }
        }
        __syncthreads();
        const unsigned long long excl = s_excl[pb] + before;
        const size_t srcpos = (size_t)t * kTile + (size_t)warp * (kRows * 512);
        const size_t dstpos = srcpos - excl;
        const uint32_t mis = (uint32_t)(excl & 15u);
        uint8_t* o = out + dstpos;
        if (mis == 0) {
#pragma unroll
            for (int r = 0; r < kRows; r++) { __stcs(reinterpret_cast<uint4*>(o + r * 512 + lane * 16), v[r]); }
        } else {
            const uint32_t q = mis >> 2, sh = (mis & 3u) * 8u;
            uint8_t* oa = reinterpret_cast<uint8_t*>(reinterpret_cast<uintptr_t>(o) & ~(uintptr_t)15);
#pragma unroll
            for (int r = 0; r < kRows; r++) {
                uint4 p;
                p.x = __shfl_up_sync(0xFFFFFFFFu, v[r].x, 1); p.y = __shfl_up_sync(0xFFFFFFFFu, v[r].y, 1);
                p.z = __shfl_up_sync(0xFFFFFFFFu, v[r].z, 1); p.w = __shfl_up_sync(0xFFFFFFFFu, v[r].w, 1);
                if (r > 0) {
                    const uint4 pr = v[r - 1];
                    const uint32_t a = __shfl_sync(0xFFFFFFFFu, pr.x, 31), b = __shfl_sync(0xFFFFFFFFu, pr.y, 31), c = __shfl_sync(0xFFFFFFFFu, pr.z, 31),
                                   d = __shfl_sync(0xFFFFFFFFu, pr.w, 31);
                    if (lane == 0) { p = make_uint4(a, b, c, d); }
                }
                const uint32_t W[8] = {p.x, p.y, p.z, p.w, v[r].x, v[r].y, v[r].z, v[r].w};
                uint32_t xw[5];
#pragma unroll
                for (int e = 0; e < 5; e++) { xw[e] = (q == 0u) ? W[e] : (q == 1u) ? W[e + 1] : (q == 2u) ? W[e + 2] : W[e + 3]; }
                uint4 o4;
                o4.x = __funnelshift_r(xw[0], xw[1], sh); o4.y = __funnelshift_r(xw[1], xw[2], sh);
                o4.z = __funnelshift_r(xw[2], xw[3], sh); o4.w = __funnelshift_r(xw[3], xw[4], sh);
                if (!(r == 0 && lane == 0)) { __stcs(reinterpret_cast<uint4*>(oa + r * 512 + lane * 16), o4); }
            }
        }
#pragma unroll
        for (int r = 0; r < kRows; r++) { v[r] = nx[r]; }
    }
}

// ---------------------------------------------------------------------------------------------- host
static float time_it(void (*fn)(void*), void* arg, int iters)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    fn(arg); fn(arg);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) { fn(arg); }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / iters;
}

struct Args {
    uint8_t* in; uint8_t* out; size_t bytes; unsigned* sink; unsigned long long* state; unsigned* counter; int sms;
};
static Args g;

template <int U, bool S> static void run_ldg(void*)
{
    const size_t nvec = g.bytes / 16;
    const unsigned grid = (unsigned)((nvec + 256 * U - 1) / (256 * U));
    k_ldg<U, S><<<grid, 256>>>((const uint4*)g.in, (uint4*)g.out, nvec, g.sink);
}
template <int T, int S, int M, int C> static void run_ring(void*)
{
    const size_t smem = (size_t)T * S + 16 * S + 64;
    static bool set = false;
    if (!set) { CK(cudaFuncSetAttribute(k_tma_ring<T, S, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); set = true; }
    k_tma_ring<T, S, M><<<g.sms * C, 288, smem>>>(g.in, g.out, (long long)(g.bytes / T), g.sink);
}
template <int R, int C> static void run_chain0(void*)
{
    const long long n_tiles = (long long)(g.bytes / (R * 4096));
    CK(cudaMemsetAsync(g.state, 0, (size_t)n_tiles * 8 + 64));
    CK(cudaMemsetAsync(g.counter, 0, 4));
    k_chain_copy<R, C><<<(unsigned)n_tiles, 256>>>(g.in, g.out, n_tiles, g.state, g.counter, 0);
}
template <int R, int C> static void run_chain1(void*)
{
    const long long n_tiles = (long long)(g.bytes / (R * 4096));
    CK(cudaMemsetAsync(g.state, 0, (size_t)n_tiles * 8 + 64));
    CK(cudaMemsetAsync(g.counter, 0, 4));
    k_chain_copy<R, C><<<(unsigned)n_tiles, 256>>>(g.in, g.out, n_tiles, g.state, g.counter, 0xFF);
}

template <int R, int W, int C, int SHIFT> static void run_persist(void*)
{
    const long long n_tiles = (long long)(g.bytes / (R * 4096));
    CK(cudaMemsetAsync(g.state, 0, (size_t)n_tiles * 16 + 64));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_chain_persist<R, W, C>, 256, 0));
    if (nb > C) { nb = C; }
    const uint8_t* in = g.in; uint8_t* out = g.out; long long nt = n_tiles; ulonglong2* st = (ulonglong2*)g.state; int sh = SHIFT;
    void* args[] = {(void*)&in, (void*)&out, (void*)&nt, (void*)&st, (void*)&sh};
    CK(cudaLaunchCooperativeKernel((const void*)k_chain_persist<R, W, C>, dim3(g.sms * nb), dim3(256), args, 0, 0));
}

static void report(const char* name, float ms)
{
    printf("%-44s %8.3f ms  %8.1f GB/s input\n", name, ms, (double)g.bytes / ms / 1e6);
    fflush(stdout);
}

int main(int argc, char** argv)
{
    const double gib = argc > 1 ? atof(argv[1]) : 2.0;
    g.bytes = (size_t)(gib * (1 << 30)) & ~(size_t)((1 << 20) - 1);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    g.sms = prop.multiProcessorCount;
    printf("device %s, %d SMs, %.2f GiB\n", prop.name, g.sms, gib);
    CK(cudaMalloc(&g.in, g.bytes + 4096));
    CK(cudaMalloc(&g.out, g.bytes + 4096));
    CK(cudaMalloc(&g.sink, 64));
    CK(cudaMalloc(&g.state, (g.bytes / 4096) * 16 + 4096));
    CK(cudaMalloc(&g.counter, 64));
    {   // pseudo-random bytes with a few zero pairs, like an escaped payload
        std::vector<uint8_t> h(64 << 20);
        uint64_t x = 88172645463325252ull;
        for (size_t i = 0; i < h.size(); i += 8) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; memcpy(&h[i], &x, 8); }
        for (size_t off = 0; off < g.bytes; off += h.size()) { CK(cudaMemcpy(g.in + off, h.data(), std::min(h.size(), g.bytes - off), cudaMemcpyHostToDevice)); }
    }
    CK(cudaMemset(g.out, 0, g.bytes));
    const int it = 10;
    {
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaMemcpy(g.out, g.in, g.bytes, cudaMemcpyDeviceToDevice)); CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int i = 0; i < it; i++) { CK(cudaMemcpyAsync(g.out, g.in, g.bytes, cudaMemcpyDeviceToDevice)); }
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        report("cudaMemcpy D2D", ms / it);
    }
    report("read_ldg U=4", time_it(run_ldg<4, false>, nullptr, it));
    report("read_ldg U=8", time_it(run_ldg<8, false>, nullptr, it));
    report("copy_ldg U=2", time_it(run_ldg<2, true>, nullptr, it));
    report("copy_ldg U=4", time_it(run_ldg<4, true>, nullptr, it));
    report("copy_ldg U=8", time_it(run_ldg<8, true>, nullptr, it));
    report("tma_ring read  16K x4 stages, 2 CTA/SM", time_it(run_ring<16384, 4, 2, 2>, nullptr, it));
    report("tma_ring read  32K x2 stages, 2 CTA/SM", time_it(run_ring<32768, 2, 2, 2>, nullptr, it));
    report("tma_ring read  32K x3 stages, 2 CTA/SM", time_it(run_ring<32768, 3, 2, 2>, nullptr, it));
    report("tma_ring read  32K x6 stages, 1 CTA/SM", time_it(run_ring<32768, 6, 2, 1>, nullptr, it));
    report("tma_ring read   8K x8 stages, 2 CTA/SM", time_it(run_ring<8192, 8, 2, 2>, nullptr, it));
    report("tma_ring copy  16K x4 stages, 2 CTA/SM", time_it(run_ring<16384, 4, 0, 2>, nullptr, it));
    report("tma_ring copy  32K x2 stages, 2 CTA/SM", time_it(run_ring<32768, 2, 0, 2>, nullptr, it));
    report("tma_ring copy  32K x3 stages, 2 CTA/SM", time_it(run_ring<32768, 3, 0, 2>, nullptr, it));
    report("tma_ring copy  32K x6 stages, 1 CTA/SM", time_it(run_ring<32768, 6, 0, 1>, nullptr, it));
    report("tma_ring copy   8K x8 stages, 2 CTA/SM", time_it(run_ring<8192, 8, 0, 2>, nullptr, it));
    report("tma_ring bulk  16K x4 stages, 2 CTA/SM", time_it(run_ring<16384, 4, 1, 2>, nullptr, it));
    report("tma_ring bulk  32K x3 stages, 2 CTA/SM", time_it(run_ring<32768, 3, 1, 2>, nullptr, it));
    report("tma_ring bulk  32K x6 stages, 1 CTA/SM", time_it(run_ring<32768, 6, 1, 1>, nullptr, it));
    report("chain_copy aligned   4 rows (16K) occ 4", time_it(run_chain0<4, 4>, nullptr, it));
    report("chain_copy aligned   4 rows (16K) occ 6", time_it(run_chain0<4, 6>, nullptr, it));
    report("chain_copy aligned   8 rows (32K) occ 3", time_it(run_chain0<8, 3>, nullptr, it));
    report("chain_copy aligned   8 rows (32K) occ 4", time_it(run_chain0<8, 4>, nullptr, it));
    report("chain_copy aligned   2 rows (8K)  occ 8", time_it(run_chain0<2, 8>, nullptr, it));
    report("chain_copy shifted   4 rows (16K) occ 4", time_it(run_chain1<4, 4>, nullptr, it));
    report("chain_copy shifted   8 rows (32K) occ 3", time_it(run_chain1<8, 3>, nullptr, it));
    report("persist  4 rows W=1 occ 4 aligned", time_it(run_persist<4, 1, 4, 0>, nullptr, it));
    report("persist  4 rows W=2 occ 4 aligned", time_it(run_persist<4, 2, 4, 0>, nullptr, it));
    report("persist  4 rows W=4 occ 4 aligned", time_it(run_persist<4, 4, 4, 0>, nullptr, it));
    report("persist  8 rows W=1 occ 3 aligned", time_it(run_persist<8, 1, 3, 0>, nullptr, it));
    report("persist  8 rows W=2 occ 3 aligned", time_it(run_persist<8, 2, 3, 0>, nullptr, it));
    report("persist  4 rows W=8 occ 4 aligned", time_it(run_persist<4, 8, 4, 0>, nullptr, it));
    report("persist  4 rows W=8 occ 6 aligned", time_it(run_persist<4, 8, 6, 0>, nullptr, it));
    report("persist  4 rows W=8 occ 8 aligned", time_it(run_persist<4, 8, 8, 0>, nullptr, it));
    report("persist  8 rows W=4 occ 3 aligned", time_it(run_persist<8, 4, 3, 0>, nullptr, it));
    report("persist  8 rows W=8 occ 3 aligned", time_it(run_persist<8, 8, 3, 0>, nullptr, it));
    report("persist  8 rows W=4 occ 4 aligned", time_it(run_persist<8, 4, 4, 0>, nullptr, it));
    report("persist  2 rows W=8 occ 8 aligned", time_it(run_persist<2, 8, 8, 0>, nullptr, it));
    report("persist  4 rows W=8 occ 4 shifted", time_it(run_persist<4, 8, 4, 255>, nullptr, it));
    report("persist  4 rows W=8 occ 6 shifted", time_it(run_persist<4, 8, 6, 255>, nullptr, it));
    report("persist  8 rows W=4 occ 3 shifted", time_it(run_persist<8, 4, 3, 255>, nullptr, it));
    {
        run_persist<4, 8, 4, 0>(nullptr);
        CK(cudaDeviceSynchronize());
        std::vector<uint8_t> a(1 << 20), b(1 << 20);
        CK(cudaMemcpy(a.data(), g.in + (g.bytes - (1 << 20)), 1 << 20, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), g.out + (g.bytes - (1 << 20)), 1 << 20, cudaMemcpyDeviceToHost));
        printf("persist aligned output check: %s\n", memcmp(a.data(), b.data(), 1 << 20) == 0 ? "ok" : "MISMATCH");
    }
    // check: aligned chain copy reproduces the input
    {
        run_chain0<4, 4>(nullptr);
        CK(cudaDeviceSynchronize());
        std::vector<uint8_t> a(1 << 20), b(1 << 20);
        CK(cudaMemcpy(a.data(), g.in + (g.bytes - (1 << 20)), 1 << 20, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), g.out + (g.bytes - (1 << 20)), 1 << 20, cudaMemcpyDeviceToHost));
        printf("chain_copy aligned output check: %s\n", memcmp(a.data(), b.data(), 1 << 20) == 0 ? "ok" : "MISMATCH");
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
