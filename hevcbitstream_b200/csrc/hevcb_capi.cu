// hevcb_capi.cu -- C ABI of libhevcb200 (include/hevcb.h): context lifecycle and the host-buffer
// wrappers around the device entry points.  No CPU implementation exists behind any entry point.
#include <cuda_runtime.h>
#include <new>
#include <stdint.h>
#include <stdlib.h>

#include "hevcb_internal.h"
#include "../../include/hevcb_layout.h"

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
    cudaDeviceSetLimit(cudaLimitStackSize, 4096); // the syntax walker keeps derived RPS tables in local memory
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
    cudaEventCreateWithFlags(&ctx->ev_stats, cudaEventDisableTiming);
    if (const char* e2 = getenv("HEVCB_SCAN_DEBUG")) { ctx->scan_debug_flags = atoll(e2); } // kernel experiment switches
    if (const char* e3 = getenv("HEVCB_HOST_CHUNK")) { const long long v = atoll(e3); if (v >= 4096) { ctx->host_chunk = v; } }
    *out = ctx;
    return HEVCB_OK;
}

HEVCB_API void hevcb_destroy(hevcb_ctx* ctx)
{
    if (!ctx) { return; }
    cudaSetDevice(ctx->device);
    hevcb_devbuf* bufs[] = {&ctx->scan_scratch, &ctx->scan_timing, &ctx->h_in, &ctx->h_rbsp, &ctx->h_a0, &ctx->h_a1, &ctx->h_a2, &ctx->h_a3, &ctx->h_misc,
                            &ctx->insert_scratch, &ctx->rewrite_slots, &ctx->rewrite_scratch, &ctx->rewrite_staging, &ctx->wstruct, &ctx->parse_scratch, &ctx->parse_ps, &ctx->parse_sort, &ctx->h_p[0], &ctx->h_p[1], &ctx->h_p[2], &ctx->h_p[3], &ctx->h_p[4],
                            &ctx->h_p[5], &ctx->h_p[6], &ctx->h_p[7], &ctx->h_p[8], &ctx->h_p[9]};
    for (hevcb_devbuf* b : bufs) {
        if (b->p) { cudaFree(b->p); }
    }
    hevcb_devbuf* pbufs[] = {&ctx->p_in[0], &ctx->p_in[1], &ctx->p_img[0], &ctx->p_img[1], &ctx->p_arr[0], &ctx->p_arr[1], &ctx->p_sum};
    for (hevcb_devbuf* b : pbufs) {
        if (b->p) { cudaFree(b->p); }
    }
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_in[i]) { cudaEventDestroy(ctx->ev_in[i]); }
        if (ctx->ev_k[i]) { cudaEventDestroy(ctx->ev_k[i]); }
        if (ctx->ev_fix[i]) { cudaEventDestroy(ctx->ev_fix[i]); }
        if (ctx->ev_out[i]) { cudaEventDestroy(ctx->ev_out[i]); }
    }
    if (ctx->ev_stats) { cudaEventDestroy(ctx->ev_stats); }
    if (ctx->pinned_sums) { cudaFreeHost(ctx->pinned_sums); }
    if (ctx->s_in) { cudaStreamDestroy(ctx->s_in); }
    if (ctx->s_out) { cudaStreamDestroy(ctx->s_out); }
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

namespace {
// local -> whole-stream coordinates for the NALs a shard owns (pipelined host path)
__global__ void globalize_offsets_kernel(int64_t* ns, int64_t* ne, int64_t* ro, int64_t* re, int64_t first, int64_t n, int64_t byte_base,
                                         int64_t rbsp_base)
{
    const int64_t i = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        ns[i] += byte_base;
        ne[i] += byte_base;
        ro[i] += rbsp_base;
        if (re[i] >= 0) { re[i] += rbsp_base; }
    }
}
} // namespace

// Large host buffers: the stream is cut into shards (hevcb_plan_shards) that are copied in, scanned and copied out on three
// streams with two buffer slots, so that the PCIe link carries input and output at the same time; the shard records are
// stitched on the host exactly as they are between GPUs.
static int scan_strip_host_pipelined(hevcb_ctx* ctx, const uint8_t* buf, int64_t size, int64_t* nal_start, int64_t* nal_end, int64_t cap_nals,
                                     uint8_t* rbsp, int64_t* rbsp_off, int64_t* rbsp_end, hevcb_scan_summary* summary)
{
    int64_t chunk = ctx->host_chunk;
    while ((size + chunk - 1) / chunk > HEVCB_MAX_SHARDS) { chunk *= 2; }
    const int K = (int)((size + chunk - 1) / chunk);
    int64_t bounds[HEVCB_MAX_SHARDS + 1];
    int rc = hevcb_plan_shards(buf, size, K, bounds);
    if (rc != HEVCB_OK) { return rc; }
    int64_t max_own = 0;
    for (int r = 0; r < K; r++) { if (bounds[r + 1] - bounds[r] > max_own) { max_own = bounds[r + 1] - bounds[r]; } }
    int64_t capS = max_own / 3 + 8;
    if (capS > cap_nals + 8) { capS = cap_nals + 8; }
    if (capS < 8) { capS = 8; }
    if (!ctx->s_in) {
        HEVCB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        HEVCB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            HEVCB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
            HEVCB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_k[i], cudaEventDisableTiming));
            HEVCB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fix[i], cudaEventDisableTiming));
            HEVCB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming));
        }
        HEVCB_CUDA(ctx, cudaMallocHost(&ctx->pinned_sums, sizeof(hevcb_shard_summary) * HEVCB_MAX_SHARDS));
    }
    for (int i = 0; i < 2; i++) {
        if ((rc = hevcb_reserve(ctx, &ctx->p_in[i], (size_t)max_own + 64)) != HEVCB_OK) { return rc; }
        if (rbsp && (rc = hevcb_reserve(ctx, &ctx->p_img[i], (size_t)max_own + 64)) != HEVCB_OK) { return rc; }
        if ((rc = hevcb_reserve(ctx, &ctx->p_arr[i], (size_t)capS * 8 * 4)) != HEVCB_OK) { return rc; }
    }
    if ((rc = hevcb_reserve(ctx, &ctx->p_sum, sizeof(hevcb_shard_summary) * 2)) != HEVCB_OK) { return rc; }
    hevcb_shard_summary* h_sums = reinterpret_cast<hevcb_shard_summary*>(ctx->pinned_sums);
    memset(h_sums, 0, sizeof(hevcb_shard_summary) * K);
    cudaStream_t s_k = ctx->stream;
    bool slot_used[2] = {false, false};
    auto shard_geom = [&](int r, int64_t& lo, int64_t& own, int64_t& halo, int& first, int& last) {
        lo = bounds[r];
        own = bounds[r + 1] - bounds[r];
        first = (own > 0 && lo == 0) ? 1 : 0;
        last = (own > 0 && bounds[r + 1] == size) ? 1 : 0;
        halo = last ? 0 : (size - bounds[r + 1] < 16 ? size - bounds[r + 1] : 16);
    };
    // Shards without bytes (a run of bytes < 2 can swallow a whole nominal chunk, hevcb_plan_shards) take no slot and no
    // launch: the pipeline walks the non-empty shards, slot = position in that list, and their zeroed records reach the stitch.
    int live[HEVCB_MAX_SHARDS];
    int n_live = 0;
    for (int r = 0; r < K; r++) { if (bounds[r + 1] - bounds[r] > 0) { live[n_live++] = r; } }
    auto copy_in = [&](int j) -> int {
        int64_t lo, own, halo; int first, last;
        shard_geom(live[j], lo, own, halo, first, last);
        const int sl = j & 1;
        if (slot_used[sl]) { HEVCB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, ctx->ev_k[sl], 0)); } // the scan of shard r - 2 has read this slot
        HEVCB_CUDA(ctx, cudaMemcpyAsync(ctx->p_in[sl].p, buf + lo, (size_t)(own + halo), cudaMemcpyHostToDevice, ctx->s_in));
        HEVCB_CUDA(ctx, cudaEventRecord(ctx->ev_in[sl], ctx->s_in));
        return HEVCB_OK;
    };
    int64_t byte_base = 0, rbsp_base = 0, nal_base = 0;
    bool overflow = false;
    if (n_live > 0 && (rc = copy_in(0)) != HEVCB_OK) { return rc; }
    for (int j = 0; j < n_live; j++) {
        const int r = live[j];
        int64_t lo, own, halo; int first, last;
        shard_geom(r, lo, own, halo, first, last);
        const int sl = j & 1;
        int64_t* arr = reinterpret_cast<int64_t*>(ctx->p_arr[sl].p);
        int64_t *d_ns = arr, *d_ne = arr + capS, *d_ro = arr + 2 * capS, *d_re = arr + 3 * capS;
        hevcb_shard_summary* d_sum = reinterpret_cast<hevcb_shard_summary*>(ctx->p_sum.p) + sl;
        HEVCB_CUDA(ctx, cudaStreamWaitEvent(s_k, ctx->ev_in[sl], 0));
        if (slot_used[sl]) { HEVCB_CUDA(ctx, cudaStreamWaitEvent(s_k, ctx->ev_out[sl], 0)); } // outputs of shard r - 2 have left the slot
        rc = hevcb_launch_scan_strip_shard(ctx, reinterpret_cast<const uint8_t*>(ctx->p_in[sl].p), own, halo, first, last, d_ns, d_ne, capS,
                                           rbsp ? reinterpret_cast<uint8_t*>(ctx->p_img[sl].p) : nullptr, d_ro, d_re, d_sum, s_k);
        if (rc != HEVCB_OK) { return rc; }
        HEVCB_CUDA(ctx, cudaMemcpyAsync(&h_sums[r], d_sum, sizeof(hevcb_shard_summary), cudaMemcpyDeviceToHost, s_k));
        HEVCB_CUDA(ctx, cudaEventRecord(ctx->ev_k[sl], s_k));
        slot_used[sl] = true;
        // the next shard's input travels while this one is scanned (its slot is free once the scan of shard r - 1 is done)
        if (j + 1 < n_live && (rc = copy_in(j + 1)) != HEVCB_OK) { return rc; }
        HEVCB_CUDA(ctx, cudaEventSynchronize(ctx->ev_k[sl])); // record of shard r: where its outputs go
        const hevcb_shard_summary& S = h_sums[r];
        if (S.overflow) { overflow = true; }
        const int64_t fl = first ? 0 : 1;
        int64_t n_local = S.n_nals < capS ? S.n_nals : capS;
        const int64_t cnt = n_local > fl ? n_local - fl : 0;
        if (cnt > 0) {
            globalize_offsets_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, s_k>>>(d_ns, d_ne, d_ro, d_re, fl, n_local, byte_base, rbsp_base);
            ctx->launches++;
        }
        HEVCB_CUDA(ctx, cudaEventRecord(ctx->ev_fix[sl], s_k));
        HEVCB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, ctx->ev_fix[sl], 0));
        if (rbsp && S.rbsp_bytes > 0) {
            HEVCB_CUDA(ctx, cudaMemcpyAsync(rbsp + rbsp_base, ctx->p_img[sl].p, (size_t)S.rbsp_bytes, cudaMemcpyDeviceToHost, ctx->s_out));
        }
        int64_t fit = cnt;
        if (nal_base + fit > cap_nals) { fit = cap_nals > nal_base ? cap_nals - nal_base : 0; overflow = true; }
        if (fit > 0) {
            HEVCB_CUDA(ctx, cudaMemcpyAsync(nal_start + nal_base, d_ns + fl, (size_t)fit * 8, cudaMemcpyDeviceToHost, ctx->s_out));
            HEVCB_CUDA(ctx, cudaMemcpyAsync(nal_end + nal_base, d_ne + fl, (size_t)fit * 8, cudaMemcpyDeviceToHost, ctx->s_out));
            if (rbsp_off) { HEVCB_CUDA(ctx, cudaMemcpyAsync(rbsp_off + nal_base, d_ro + fl, (size_t)fit * 8, cudaMemcpyDeviceToHost, ctx->s_out)); }
            if (rbsp_end) { HEVCB_CUDA(ctx, cudaMemcpyAsync(rbsp_end + nal_base, d_re + fl, (size_t)fit * 8, cudaMemcpyDeviceToHost, ctx->s_out)); }
        }
        HEVCB_CUDA(ctx, cudaEventRecord(ctx->ev_out[sl], ctx->s_out));
        byte_base += own;
        rbsp_base += S.rbsp_bytes;
        nal_base += cnt;
    }
    HEVCB_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
    HEVCB_CUDA(ctx, cudaStreamSynchronize(s_k));
    // join the shards (same code as between GPUs) and patch the entries that depend on a neighbour
    hevcb_stitch_result* res = new (std::nothrow) hevcb_stitch_result();
    if (!res) { return HEVCB_E_NOMEM; }
    rc = hevcb_stitch(h_sums, K, res);
    if (rc != HEVCB_OK) { delete res; HEVCB_SET_ERR(ctx, "hevcb_scan_strip_host: inconsistent shard records"); return rc; }
    for (int i = 0; i < res->n_patches; i++) {
        const hevcb_stitch_patch& p = res->patches[i];
        const int64_t g = res->nal_base[p.shard] + p.index - res->first_local[p.shard];
        if (g < 0 || g >= cap_nals) { if (g >= cap_nals) { overflow = true; } continue; }
        if (p.set_start) {
            nal_start[g] = res->byte_base[p.shard] + p.nal_start;
            if (rbsp_off) { rbsp_off[g] = res->rbsp_base[p.shard] + p.rbsp_off; }
        }
        nal_end[g] = res->byte_base[p.shard] + p.nal_end;
        if (rbsp_end) { rbsp_end[g] = p.rbsp_end < 0 ? -1 : res->rbsp_base[p.shard] + p.rbsp_end; }
    }
    *summary = res->global;
    if (summary->n_nals > cap_nals || overflow) { summary->overflow = 1; }
    delete res;
    if (summary->overflow) {
        HEVCB_SET_ERR(ctx, "hevcb_scan_strip_host: %lld NALs exceed cap_nals %lld", (long long)summary->n_nals, (long long)cap_nals);
        return HEVCB_E_CAPACITY;
    }
    return HEVCB_OK;
}

HEVCB_API int hevcb_scan_strip_host(hevcb_ctx* ctx, const uint8_t* buf, int64_t size, int64_t* nal_start, int64_t* nal_end,
                                    int64_t cap_nals, uint8_t* rbsp, int64_t* rbsp_off, int64_t* rbsp_end, hevcb_scan_summary* summary)
{
    if (!ctx || !summary || size < 0 || cap_nals < 0 || !nal_start || !nal_end || (size > 0 && !buf)) {
        HEVCB_SET_ERR(ctx, "hevcb_scan_strip_host: invalid argument");
        return HEVCB_E_ARG;
    }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (size >= 2 * ctx->host_chunk) { return scan_strip_host_pipelined(ctx, buf, size, nal_start, nal_end, cap_nals, rbsp, rbsp_off, rbsp_end, summary); }
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


HEVCB_API int hevcb_scan_strip_shard_device(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t own, int64_t halo, int is_first, int is_last,
                                            int64_t* d_nal_start, int64_t* d_nal_end, int64_t cap_nals, uint8_t* d_rbsp, int64_t* d_rbsp_off,
                                            int64_t* d_rbsp_end, hevcb_shard_summary* d_summary, void* stream)
{
    if (!ctx) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    return hevcb_launch_scan_strip_shard(ctx, d_buf, own, halo, is_first, is_last, d_nal_start, d_nal_end, cap_nals, d_rbsp, d_rbsp_off,
                                         d_rbsp_end, d_summary, (cudaStream_t)stream);
}

namespace {
struct PatchPack {
    int n;
    hevcb_stitch_patch p[8];
};
__global__ void apply_patches_kernel(PatchPack pk, int64_t* ns, int64_t* ne, int64_t* ro, int64_t* re, int64_t cap)
{
    const int i = threadIdx.x;
    if (i < pk.n && pk.p[i].index >= 0 && pk.p[i].index < cap) {
        const hevcb_stitch_patch& p = pk.p[i];
        if (p.set_start) { ns[p.index] = p.nal_start; ro[p.index] = p.rbsp_off; }
        ne[p.index] = p.nal_end;
        re[p.index] = p.rbsp_end;
    }
}
} // namespace

HEVCB_API int hevcb_apply_patches_device(hevcb_ctx* ctx, const hevcb_stitch_result* res, int shard, int64_t* d_nal_start, int64_t* d_nal_end,
                                         int64_t* d_rbsp_off, int64_t* d_rbsp_end, int64_t cap_nals, void* stream)
{
    if (!ctx || !res || !d_nal_start || !d_nal_end || !d_rbsp_off || !d_rbsp_end) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    PatchPack pk;
    pk.n = 0;
    for (int i = 0; i < res->n_patches; i++) {
        if (res->patches[i].shard == shard && pk.n < 8) { pk.p[pk.n++] = res->patches[i]; }
    }
    if (pk.n == 0) { return HEVCB_OK; }
    apply_patches_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pk, d_nal_start, d_nal_end, d_rbsp_off, d_rbsp_end, cap_nals);
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}

HEVCB_API int hevcb_reframe_device(hevcb_ctx* ctx, const uint8_t* d_buf, const int64_t* d_nal_start, const int64_t* d_nal_end, int64_t n_nals,
                                   int start_code_len, int len_size, uint8_t* d_out, int64_t out_cap, int64_t* d_out_off,
                                   hevcb_insert_summary* d_summary, void* stream)
{
    if (!ctx) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    return hevcb_launch_frame(ctx, d_buf, d_nal_start, d_nal_end, n_nals, start_code_len, len_size, d_out, out_cap, d_out_off, d_summary, (cudaStream_t)stream);
}

HEVCB_API int hevcb_lenpref_index_device(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, int len_size, const int64_t* d_sample_off, int64_t n_samples,
                                         int64_t* d_nal_start, int64_t* d_nal_end, int64_t cap_nals, int64_t* d_total, void* stream)
{
    if (!ctx) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    return hevcb_launch_lenpref_index(ctx, d_buf, size, len_size, d_sample_off, n_samples, d_nal_start, d_nal_end, cap_nals, d_total, (cudaStream_t)stream);
}

HEVCB_API int hevcb_rewrite_device(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, const int64_t* d_nal_start, const int64_t* d_nal_end,
                                   const uint8_t* d_rbsp, const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n_nals,
                                   const hevcb_parse_buffers* parsed, const hevcb_edit_set* edits, uint8_t* d_out, int64_t out_cap,
                                   int64_t* d_out_start, int64_t* d_out_end, hevcb_rewrite_summary* d_summary, void* stream)
{
    if (!ctx) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    return hevcb_launch_rewrite(ctx, d_buf, size, d_nal_start, d_nal_end, d_rbsp, d_rbsp_off, d_rbsp_end, n_nals, parsed, edits, d_out, out_cap,
                                d_out_start, d_out_end, d_summary, (cudaStream_t)stream);
}

HEVCB_API int hevcb_write_nal_host(hevcb_ctx* ctx, int nal_unit_type, int nal_layer_id, int nal_temporal_id_plus1, const void* vps, const void* sps,
                                   const void* pps, const void* sh, uint8_t* nal_out, int64_t size, int64_t* nal_bytes)
{
    if (!ctx || !vps || !sps || !pps || !sh || !nal_out || !nal_bytes || size < 0) {
        HEVCB_SET_ERR(ctx, "hevcb_write_nal_host: invalid argument");
        return HEVCB_E_ARG;
    }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    *nal_bytes = -1;
    const bool slice = (nal_unit_type >= 0 && nal_unit_type <= 9) || (nal_unit_type >= 16 && nal_unit_type <= 21);
    if (!slice && nal_unit_type != 32 && nal_unit_type != 33 && nal_unit_type != 34) { return HEVCB_OK; } // default: return -1 (hevc_stream.c:1313)
    cudaStream_t st = ctx->stream;
    const int64_t rcap = size * 3 / 4; // hevc_stream.c:1266
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_vps = 0, o_sps = o_vps + up(sizeof(hevc_vps_t)), o_pps = o_sps + up(sizeof(hevc_sps_t)), o_sh = o_pps + up(sizeof(hevc_pps_t));
    const size_t o_ctx = o_sh + up(sizeof(hevc_slice_header_t)), o_res = o_ctx + up(hevcb_write_struct_scratch_bytes()), o_off = o_res + 256;
    const size_t o_rbsp = o_off + 256, o_nal = o_rbsp + up((size_t)rcap + 64), total = o_nal + up((size_t)rcap * 3 / 2 + 128);
    int rc = hevcb_reserve(ctx, &ctx->wstruct, total);
    if (rc != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_misc, 256)) != HEVCB_OK) { return rc; }
    uint8_t* b = reinterpret_cast<uint8_t*>(ctx->wstruct.p);
    HEVCB_CUDA(ctx, cudaMemcpyAsync(b + o_vps, vps, sizeof(hevc_vps_t), cudaMemcpyHostToDevice, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(b + o_sps, sps, sizeof(hevc_sps_t), cudaMemcpyHostToDevice, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(b + o_pps, pps, sizeof(hevc_pps_t), cudaMemcpyHostToDevice, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(b + o_sh, sh, sizeof(hevc_slice_header_t), cudaMemcpyHostToDevice, st));
    HEVCB_CUDA(ctx, cudaMemsetAsync(b + o_rbsp, 0, (size_t)rcap + 64, st)); // the reference writes into a calloc'ed buffer
    const int32_t hdr = (nal_unit_type & 0xFF) | ((nal_layer_id & 0xFF) << 8) | ((nal_temporal_id_plus1 & 0xFF) << 16);
    int64_t* d_res = reinterpret_cast<int64_t*>(b + o_res);
    rc = hevcb_launch_write_struct(ctx, hdr, reinterpret_cast<const int32_t*>(b + o_vps), reinterpret_cast<const int32_t*>(b + o_sps),
                                   reinterpret_cast<const int32_t*>(b + o_pps), reinterpret_cast<const int32_t*>(b + o_sh), b + o_ctx, b + o_rbsp, rcap, d_res,
                                   st);
    if (rc != HEVCB_OK) { return rc; }
    int64_t* p_res = reinterpret_cast<int64_t*>(reinterpret_cast<uint8_t*>(ctx->pinned) + 1024);
    HEVCB_CUDA(ctx, cudaMemcpyAsync(p_res, d_res, 16, cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    if (!p_res[1]) { return HEVCB_OK; } // overrun: -1 (hevc_stream.c:1317)
    // rbsp_to_nal of what was written (hevc_stream.c:1324-1326)
    int64_t* d_off = reinterpret_cast<int64_t*>(b + o_off);
    const int64_t seg[2] = {0, p_res[0]};
    HEVCB_CUDA(ctx, cudaMemcpyAsync(d_off, seg, 16, cudaMemcpyHostToDevice, st));
    hevcb_insert_summary* d_sum = reinterpret_cast<hevcb_insert_summary*>(ctx->h_misc.p);
    rc = hevcb_launch_insert(ctx, b + o_rbsp, d_off, d_off + 1, 1, 0, b + o_nal, (int64_t)((size_t)rcap * 3 / 2 + 64), d_off + 2, d_sum, st);
    if (rc != HEVCB_OK) { return rc; }
    hevcb_insert_summary* p_sum = reinterpret_cast<hevcb_insert_summary*>(ctx->pinned);
    HEVCB_CUDA(ctx, cudaMemcpyAsync(p_sum, d_sum, sizeof(hevcb_insert_summary), cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    if (p_sum->out_bytes > 0) {
        HEVCB_CUDA(ctx, cudaMemcpyAsync(nal_out, b + o_nal, (size_t)p_sum->out_bytes, cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    *nal_bytes = p_sum->out_bytes;
    return HEVCB_OK;
}

HEVCB_API int hevcb_insert_device(hevcb_ctx* ctx, const uint8_t* d_rbsp, const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n_nals,
                                  int start_code_len, uint8_t* d_out, int64_t out_cap, int64_t* d_out_off, hevcb_insert_summary* d_summary,
                                  void* stream)
{
    if (!ctx) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    return hevcb_launch_insert(ctx, d_rbsp, d_rbsp_off, d_rbsp_end, n_nals, start_code_len, d_out, out_cap, d_out_off, d_summary,
                               (cudaStream_t)stream);
}

HEVCB_API int hevcb_insert_host(hevcb_ctx* ctx, const uint8_t* rbsp, int64_t rbsp_bytes, const int64_t* rbsp_off, const int64_t* rbsp_end,
                                int64_t n_nals, int start_code_len, uint8_t* out, int64_t out_cap, int64_t* out_off, hevcb_insert_summary* summary)
{
    if (!ctx || !summary || !out_off || rbsp_bytes < 0 || n_nals < 0 || out_cap < 0 || (n_nals > 0 && (!rbsp_off || !rbsp_end)) ||
        (rbsp_bytes > 0 && !rbsp) || (out_cap > 0 && !out)) {
        HEVCB_SET_ERR(ctx, "hevcb_insert_host: invalid argument");
        return HEVCB_E_ARG;
    }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t in_bytes = ((size_t)rbsp_bytes + 15u) & ~(size_t)15u;
    const size_t arr_bytes = (size_t)(n_nals + 1) * sizeof(int64_t);
    int rc;
    if ((rc = hevcb_reserve(ctx, &ctx->h_rbsp, in_bytes + 16)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_in, (size_t)out_cap + 16)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a0, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a2, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a3, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_misc, 256)) != HEVCB_OK) { return rc; }
    if (rbsp_bytes > 0) { HEVCB_CUDA(ctx, cudaMemcpyAsync(ctx->h_rbsp.p, rbsp, (size_t)rbsp_bytes, cudaMemcpyHostToDevice, st)); }
    if (n_nals > 0) {
        HEVCB_CUDA(ctx, cudaMemcpyAsync(ctx->h_a2.p, rbsp_off, (size_t)n_nals * 8, cudaMemcpyHostToDevice, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(ctx->h_a3.p, rbsp_end, (size_t)n_nals * 8, cudaMemcpyHostToDevice, st));
    }
    hevcb_insert_summary* d_sum = reinterpret_cast<hevcb_insert_summary*>(ctx->h_misc.p);
    rc = hevcb_launch_insert(ctx, reinterpret_cast<const uint8_t*>(ctx->h_rbsp.p), reinterpret_cast<const int64_t*>(ctx->h_a2.p),
                             reinterpret_cast<const int64_t*>(ctx->h_a3.p), n_nals, start_code_len, reinterpret_cast<uint8_t*>(ctx->h_in.p), out_cap,
                             reinterpret_cast<int64_t*>(ctx->h_a0.p), d_sum, st);
    if (rc != HEVCB_OK) { return rc; }
    hevcb_insert_summary* p_sum = reinterpret_cast<hevcb_insert_summary*>(ctx->pinned);
    HEVCB_CUDA(ctx, cudaMemcpyAsync(p_sum, d_sum, sizeof(hevcb_insert_summary), cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(out_off, ctx->h_a0.p, arr_bytes, cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    *summary = *p_sum;
    if (summary->overflow) {
        HEVCB_SET_ERR(ctx, "hevcb_insert_host: %lld output bytes exceed out_cap %lld", (long long)summary->out_bytes, (long long)out_cap);
        return HEVCB_E_CAPACITY;
    }
    if (summary->out_bytes > 0) {
        HEVCB_CUDA(ctx, cudaMemcpyAsync(out, ctx->h_in.p, (size_t)summary->out_bytes, cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    return HEVCB_OK;
}

HEVCB_API int hevcb_parse_device(hevcb_ctx* ctx, const uint8_t* d_buf, const int64_t* d_nal_start, const int64_t* d_nal_end, const uint8_t* d_rbsp,
                                 const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n_nals, const hevcb_parse_buffers* out,
                                 hevcb_parse_summary* d_summary, void* stream)
{
    if (!ctx) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    return hevcb_launch_parse(ctx, d_buf, d_nal_start, d_nal_end, d_rbsp, d_rbsp_off, d_rbsp_end, n_nals, out, d_summary, nullptr, (cudaStream_t)stream);
}

HEVCB_API int hevcb_parse_shard_device(hevcb_ctx* ctx, const uint8_t* d_buf, const int64_t* d_nal_start, const int64_t* d_nal_end, const uint8_t* d_rbsp,
                                       const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n_nals, const hevcb_parse_buffers* out,
                                       hevcb_parse_summary* d_summary, const hevcb_parse_chain* chain, void* stream)
{
    if (!ctx) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    return hevcb_launch_parse(ctx, d_buf, d_nal_start, d_nal_end, d_rbsp, d_rbsp_off, d_rbsp_end, n_nals, out, d_summary, chain, (cudaStream_t)stream);
}

HEVCB_API int hevcb_index_host(hevcb_ctx* ctx, const uint8_t* buf, int64_t size, hevcb_stream_index* idx)
{
    return hevcb_index_host_chain(ctx, buf, size, idx, nullptr);
}

HEVCB_API int hevcb_index_host_chain(hevcb_ctx* ctx, const uint8_t* buf, int64_t size, hevcb_stream_index* idx, const hevcb_parse_chain* chain)
{
    if (!ctx || !idx || size < 0 || (size > 0 && !buf) || idx->cap_nals < 0 || !idx->nal_start || !idx->nal_end || !idx->rbsp_off || !idx->rbsp_end ||
        !idx->p.rc || !idx->p.nal_hdr || !idx->p.kind || !idx->p.ubflag || !idx->p.hdr_end || !idx->p.cols || !idx->p.pair_off) {
        HEVCB_SET_ERR(ctx, "hevcb_index_host: invalid argument");
        return HEVCB_E_ARG;
    }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int64_t cap = idx->cap_nals;
    const size_t in_bytes = ((size_t)size + 15u) & ~(size_t)15u;
    const size_t arr_bytes = (size_t)(cap > 0 ? cap : 1) * sizeof(int64_t);
    int rc;
    if ((rc = hevcb_reserve(ctx, &ctx->h_in, in_bytes + 16)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_rbsp, in_bytes + 16)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a0, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a1, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a2, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a3, arr_bytes)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_misc, 256)) != HEVCB_OK) { return rc; }
    if (size > 0) { HEVCB_CUDA(ctx, cudaMemcpyAsync(ctx->h_in.p, buf, (size_t)size, cudaMemcpyHostToDevice, st)); }
    hevcb_scan_summary* d_sum = reinterpret_cast<hevcb_scan_summary*>(ctx->h_misc.p);
    hevcb_parse_summary* d_psum = reinterpret_cast<hevcb_parse_summary*>(reinterpret_cast<uint8_t*>(ctx->h_misc.p) + 128);
    const uint8_t* d_in = reinterpret_cast<const uint8_t*>(ctx->h_in.p);
    uint8_t* d_rbsp = reinterpret_cast<uint8_t*>(ctx->h_rbsp.p);
    int64_t* d_ns = reinterpret_cast<int64_t*>(ctx->h_a0.p);
    int64_t* d_ne = reinterpret_cast<int64_t*>(ctx->h_a1.p);
    int64_t* d_ro = reinterpret_cast<int64_t*>(ctx->h_a2.p);
    int64_t* d_re = reinterpret_cast<int64_t*>(ctx->h_a3.p);
    rc = hevcb_launch_scan_strip(ctx, d_in, size, d_ns, d_ne, cap, d_rbsp, d_ro, d_re, d_sum, st);
    if (rc != HEVCB_OK) { return rc; }
    hevcb_scan_summary* p_sum = reinterpret_cast<hevcb_scan_summary*>(ctx->pinned);
    hevcb_parse_summary* p_psum = reinterpret_cast<hevcb_parse_summary*>(reinterpret_cast<uint8_t*>(ctx->pinned) + 128);
    HEVCB_CUDA(ctx, cudaMemcpyAsync(p_sum, d_sum, sizeof(hevcb_scan_summary), cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    idx->scan = *p_sum;
    memset(&idx->parse, 0, sizeof(idx->parse));
    if (idx->scan.overflow || idx->scan.n_nals > cap) {
        HEVCB_SET_ERR(ctx, "hevcb_index_host: %lld NALs exceed cap_nals %lld", (long long)idx->scan.n_nals, (long long)cap);
        return HEVCB_E_CAPACITY;
    }
    const int64_t n = idx->scan.n_nals;
    // device staging of the parse outputs
    const size_t sizes[10] = {(size_t)n * 4, (size_t)n * 4, (size_t)n, (size_t)n, (size_t)n * 4, (size_t)n * 32, (size_t)(n + 1) * 8,
                              (size_t)idx->p.cap_pairs * 4, (size_t)idx->p.cap_pairs * 4, idx->p.pair_pos ? (size_t)idx->p.cap_pairs * 4 : 0};
    for (int i = 0; i < 10; i++) {
        if ((rc = hevcb_reserve(ctx, &ctx->h_p[i], sizes[i] + 16)) != HEVCB_OK) { return rc; }
    }
    hevcb_parse_buffers d;
    d.rc = reinterpret_cast<int32_t*>(ctx->h_p[0].p);
    d.nal_hdr = reinterpret_cast<int32_t*>(ctx->h_p[1].p);
    d.kind = reinterpret_cast<uint8_t*>(ctx->h_p[2].p);
    d.ubflag = reinterpret_cast<uint8_t*>(ctx->h_p[3].p);
    d.hdr_end = reinterpret_cast<int32_t*>(ctx->h_p[4].p);
    d.cols = reinterpret_cast<int32_t*>(ctx->h_p[5].p);
    d.pair_off = reinterpret_cast<int64_t*>(ctx->h_p[6].p);
    d.pair_field = reinterpret_cast<uint32_t*>(ctx->h_p[7].p);
    d.pair_value = reinterpret_cast<int32_t*>(ctx->h_p[8].p);
    d.cap_pairs = idx->p.cap_pairs;
    d.pair_pos = idx->p.pair_pos ? reinterpret_cast<uint32_t*>(ctx->h_p[9].p) : nullptr; // host array given: the trace variant
    d.flags = idx->p.flags; d.pad = 0;
    rc = hevcb_launch_parse(ctx, d_in, d_ns, d_ne, d_rbsp, d_ro, d_re, n, &d, d_psum, chain, st);
    if (rc != HEVCB_OK) { return rc; }
    HEVCB_CUDA(ctx, cudaMemcpyAsync(p_psum, d_psum, sizeof(hevcb_parse_summary), cudaMemcpyDeviceToHost, st));
    if (n > 0) {
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->nal_start, d_ns, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->nal_end, d_ne, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->rbsp_off, d_ro, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->rbsp_end, d_re, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->p.rc, d.rc, sizes[0], cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->p.nal_hdr, d.nal_hdr, sizes[1], cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->p.kind, d.kind, sizes[2], cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->p.ubflag, d.ubflag, sizes[3], cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->p.hdr_end, d.hdr_end, sizes[4], cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->p.cols, d.cols, sizes[5], cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->p.pair_off, d.pair_off, sizes[6], cudaMemcpyDeviceToHost, st));
    }
    if (idx->rbsp && idx->scan.rbsp_bytes > 0) {
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->rbsp, d_rbsp, (size_t)idx->scan.rbsp_bytes, cudaMemcpyDeviceToHost, st));
    }
    HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    idx->parse = *p_psum;
    const int64_t np = idx->parse.n_pairs < idx->p.cap_pairs ? idx->parse.n_pairs : idx->p.cap_pairs;
    if (np > 0 && idx->p.pair_field && idx->p.pair_value) {
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->p.pair_field, d.pair_field, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->p.pair_value, d.pair_value, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
        if (idx->p.pair_pos) { HEVCB_CUDA(ctx, cudaMemcpyAsync(idx->p.pair_pos, d.pair_pos, (size_t)np * 4, cudaMemcpyDeviceToHost, st)); }
        HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (idx->parse.overflow) {
        HEVCB_SET_ERR(ctx, "hevcb_index_host: %lld syntax elements exceed cap_pairs %lld", (long long)idx->parse.n_pairs, (long long)idx->p.cap_pairs);
        return HEVCB_E_CAPACITY;
    }
    return HEVCB_OK;
}

// Header parse of RBSPs the caller already holds (no start codes, no emulation prevention bytes): n segments of one host buffer.
HEVCB_API int hevcb_parse_rbsp_host(hevcb_ctx* ctx, const uint8_t* rbsp, int64_t rbsp_bytes, const int64_t* rbsp_off, const int64_t* rbsp_end,
                                    int64_t n, const hevcb_parse_buffers* out, hevcb_parse_summary* summary, const hevcb_parse_chain* chain)
{
    if (!ctx || !out || !summary || n < 0 || rbsp_bytes < 0 || (n > 0 && (!rbsp_off || !rbsp_end)) || (rbsp_bytes > 0 && !rbsp) || !out->rc ||
        !out->nal_hdr || !out->kind || !out->ubflag || !out->hdr_end || !out->cols || !out->pair_off) {
        HEVCB_SET_ERR(ctx, "hevcb_parse_rbsp_host: invalid argument");
        return HEVCB_E_ARG;
    }
    for (int64_t k = 0; k < n; k++) {
        if (rbsp_end[k] >= 0 && (rbsp_off[k] < 0 || rbsp_end[k] < rbsp_off[k] || rbsp_end[k] > rbsp_bytes)) {
            HEVCB_SET_ERR(ctx, "hevcb_parse_rbsp_host: segment %lld lies outside the buffer", (long long)k);
            return HEVCB_E_ARG;
        }
    }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    memset(summary, 0, sizeof(*summary));
    int rc;
    if (n == 0) {
        if ((rc = hevcb_reserve(ctx, &ctx->h_misc, 256)) != HEVCB_OK) { return rc; }
        return hevcb_launch_parse(ctx, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, out, reinterpret_cast<hevcb_parse_summary*>(ctx->h_misc.p), chain, st);
    }
    const size_t img = ((size_t)rbsp_bytes + 31u) & ~(size_t)15u;
    if ((rc = hevcb_reserve(ctx, &ctx->h_rbsp, img + 16)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a2, (size_t)n * 8)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_a3, (size_t)n * 8)) != HEVCB_OK) { return rc; }
    if ((rc = hevcb_reserve(ctx, &ctx->h_misc, 256)) != HEVCB_OK) { return rc; }
    uint8_t* d_rbsp = reinterpret_cast<uint8_t*>(ctx->h_rbsp.p);
    int64_t* d_ro = reinterpret_cast<int64_t*>(ctx->h_a2.p);
    int64_t* d_re = reinterpret_cast<int64_t*>(ctx->h_a3.p);
    HEVCB_CUDA(ctx, cudaMemsetAsync(d_rbsp + (img - 32), 0, 32, st)); // the bit reader's aligned 8-byte loads reach past the last byte
    if (rbsp_bytes > 0) { HEVCB_CUDA(ctx, cudaMemcpyAsync(d_rbsp, rbsp, (size_t)rbsp_bytes, cudaMemcpyHostToDevice, st)); }
    HEVCB_CUDA(ctx, cudaMemcpyAsync(d_ro, rbsp_off, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(d_re, rbsp_end, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    const size_t sizes[10] = {(size_t)n * 4, (size_t)n * 4, (size_t)n, (size_t)n, (size_t)n * 4, (size_t)n * 32, (size_t)(n + 1) * 8,
                              (size_t)out->cap_pairs * 4, (size_t)out->cap_pairs * 4, out->pair_pos ? (size_t)out->cap_pairs * 4 : 0};
    for (int i = 0; i < 10; i++) {
        if ((rc = hevcb_reserve(ctx, &ctx->h_p[i], sizes[i] + 16)) != HEVCB_OK) { return rc; }
    }
    hevcb_parse_buffers d;
    d.rc = reinterpret_cast<int32_t*>(ctx->h_p[0].p);
    d.nal_hdr = reinterpret_cast<int32_t*>(ctx->h_p[1].p);
    d.kind = reinterpret_cast<uint8_t*>(ctx->h_p[2].p);
    d.ubflag = reinterpret_cast<uint8_t*>(ctx->h_p[3].p);
    d.hdr_end = reinterpret_cast<int32_t*>(ctx->h_p[4].p);
    d.cols = reinterpret_cast<int32_t*>(ctx->h_p[5].p);
    d.pair_off = reinterpret_cast<int64_t*>(ctx->h_p[6].p);
    d.pair_field = reinterpret_cast<uint32_t*>(ctx->h_p[7].p);
    d.pair_value = reinterpret_cast<int32_t*>(ctx->h_p[8].p);
    d.cap_pairs = out->cap_pairs;
    d.pair_pos = out->pair_pos ? reinterpret_cast<uint32_t*>(ctx->h_p[9].p) : nullptr;
    d.flags = out->flags; d.pad = 0;
    // the RBSP doubles as "the NAL bytes": rc[k] is then the RBSP size (buf_size 1 keeps the trailing-00-00-03 rule, which is about
    // bytes the caller stripped, away from it)
    hevcb_parse_chain ch;
    ch.sps_in = chain ? chain->sps_in : nullptr; ch.pps_in = chain ? chain->pps_in : nullptr;
    ch.sps_out = chain ? chain->sps_out : nullptr; ch.pps_out = chain ? chain->pps_out : nullptr;
    ch.buf_size = 1;
    hevcb_parse_summary* d_psum = reinterpret_cast<hevcb_parse_summary*>(reinterpret_cast<uint8_t*>(ctx->h_misc.p) + 128);
    rc = hevcb_launch_parse(ctx, d_rbsp, d_ro, d_re, d_rbsp, d_ro, d_re, n, &d, d_psum, &ch, st);
    if (rc != HEVCB_OK) { return rc; }
    hevcb_parse_summary* p_psum = reinterpret_cast<hevcb_parse_summary*>(reinterpret_cast<uint8_t*>(ctx->pinned) + 128);
    HEVCB_CUDA(ctx, cudaMemcpyAsync(p_psum, d_psum, sizeof(hevcb_parse_summary), cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(out->rc, d.rc, sizes[0], cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(out->nal_hdr, d.nal_hdr, sizes[1], cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(out->kind, d.kind, sizes[2], cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(out->ubflag, d.ubflag, sizes[3], cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(out->hdr_end, d.hdr_end, sizes[4], cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(out->cols, d.cols, sizes[5], cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(out->pair_off, d.pair_off, sizes[6], cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    *summary = *p_psum;
    const int64_t np = summary->n_pairs < out->cap_pairs ? summary->n_pairs : out->cap_pairs;
    if (np > 0 && out->pair_field && out->pair_value) {
        HEVCB_CUDA(ctx, cudaMemcpyAsync(out->pair_field, d.pair_field, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
        HEVCB_CUDA(ctx, cudaMemcpyAsync(out->pair_value, d.pair_value, (size_t)np * 4, cudaMemcpyDeviceToHost, st));
        if (out->pair_pos) { HEVCB_CUDA(ctx, cudaMemcpyAsync(out->pair_pos, d.pair_pos, (size_t)np * 4, cudaMemcpyDeviceToHost, st)); }
        HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    }
    if (summary->overflow) {
        HEVCB_SET_ERR(ctx, "hevcb_parse_rbsp_host: %lld syntax elements exceed cap_pairs %lld", (long long)summary->n_pairs, (long long)out->cap_pairs);
        return HEVCB_E_CAPACITY;
    }
    return HEVCB_OK;
}

// Host-side formatting of device results into the reference's struct layout: zero the struct the NAL wrote and scatter
// its (field, value) pairs.  No parsing happens here.
HEVCB_API int hevcb_materialize(const hevcb_stream_index* idx, int64_t k, void* nal, void* vps, void* sps, void* pps, void* sh)
{
    if (!idx || k < 0 || k >= idx->scan.n_nals) { return HEVCB_E_ARG; }
    const int32_t hdr = idx->p.nal_hdr[k];
    if (hdr != -1 && nal) {
        hevc_nal_t* nn = reinterpret_cast<hevc_nal_t*>(nal);
        nn->nal_unit_type = hdr & 0xFF;
        nn->nal_layer_id = (hdr >> 8) & 0xFF;
        nn->nal_temporal_id_plus1 = (hdr >> 16) & 0xFF;
    }
    int32_t* dst = nullptr;
    size_t words = 0;
    switch (idx->p.kind[k]) {
        case HEVCB_KIND_VPS: dst = reinterpret_cast<int32_t*>(vps); words = sizeof(hevc_vps_t) / 4; break;
        case HEVCB_KIND_SPS: dst = reinterpret_cast<int32_t*>(sps); words = sizeof(hevc_sps_t) / 4; break;
        case HEVCB_KIND_PPS: dst = reinterpret_cast<int32_t*>(pps); words = sizeof(hevc_pps_t) / 4; break;
        case HEVCB_KIND_SLICE: dst = reinterpret_cast<int32_t*>(sh); words = sizeof(hevc_slice_header_t) / 4; break;
        default: break;
    }
    if (dst) {
        memset(dst, 0, words * 4);
        const int64_t a = idx->p.pair_off[k], b = idx->p.pair_off[k + 1];
        for (int64_t i = a; i < b && i < idx->p.cap_pairs; i++) {
            uint32_t f = idx->p.pair_field[i];
            if (f & HEVCB_TRACE_SPECIAL) { continue; } // trace variant: a printed line that is not a struct member
            f &= ~HEVCB_TRACE_SILENT;
            if (f < words) { dst[f] = idx->p.pair_value[i]; }
        }
    }
    return idx->p.rc[k];
}

} // extern "C"
