// hevcb_internal.h -- context object and helpers shared by the translation units of libhevcb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hevcb.h"

struct hevcb_devbuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct hevcb_ctx {
    int device = 0;
    int sm_count = 0;
    int64_t launches = 0;
    char err[512] = {0};
    // scan scratch: [0,64) counters, then one 16-byte state word per tile
    hevcb_devbuf scan_scratch;
    hevcb_devbuf scan_timing;             // HEVCB_SCAN_TIMING: event stamps of the scan pipeline (measurement aid)
    int scan_timing_ctas = 0;
    hevcb_devbuf rewrite_slots;           // rewrite: per-NAL header slots of the first write pass
    hevcb_devbuf insert_scratch;          // insert: per-NAL output sizes + block sums
    int fused_smem_set = 0;               // insert: the single-pass kernel's shared-memory attribute is set
    hevcb_devbuf rewrite_scratch, rewrite_staging; // rewrite: part arrays, written headers
    hevcb_devbuf wstruct;                 // hevcb_write_nal_host: uploaded structs, contexts, RBSP, NAL
    // what the last hevcb_parse_* left on the device (consumed by hevcb_rewrite_*)
    struct {
        int64_t n = -1;
        int spec = 0; // the parse ran with HEVCB_PARSE_SPEC
        const void *cls = nullptr, *sps_ord = nullptr, *pps_ord = nullptr, *cnt = nullptr, *perm = nullptr;
        void *sps_tab = nullptr, *pps_tab = nullptr, *sps_scratch = nullptr;
    } last_parse;
    hevcb_devbuf parse_scratch, parse_ps; // parser: per-NAL scratch arrays, parameter-set context tables
    hevcb_devbuf parse_sort;              // parser: radix sort workspace
    hevcb_devbuf h_p[10];                 // staging of the parse outputs for the *_host entry points
    // staging used by the *_host entry points
    hevcb_devbuf h_in, h_rbsp, h_a0, h_a1, h_a2, h_a3, h_misc;
    void* pinned = nullptr; // small pinned block for summaries
    size_t pinned_bytes = 0;
    cudaStream_t stream = nullptr; // stream used by *_host entry points
    // pipelined host path (hevcb_scan_strip_host on large buffers): copy-in / kernel / copy-out streams, two slots
    cudaStream_t s_in = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_fix[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    hevcb_devbuf p_in[2], p_img[2], p_arr[2], p_sum;
    void* pinned_sums = nullptr;   // HEVCB_MAX_SHARDS shard records
    int64_t host_chunk = 64ll << 20; // bytes per shard of the pipelined host path (HEVCB_HOST_CHUNK); 64 MiB measured best
    int scan_blocks_per_sm = 0;
    cudaEvent_t ev_stats = nullptr; // scan kernel: heavy-tile count of the last launch has arrived in the pinned block
    bool stats_pending = false;
    double last_heavy_frac = 0.0, last_flagged_frac = 0.0;
    long long scan_debug_flags = 0; // experiment switches of the scan kernel (HEVCB_SCAN_DEBUG); 0 in production
};

extern char g_hevcb_create_error[512];

#define HEVCB_SET_ERR(ctx, ...)                                         \
    do {                                                                \
        if (ctx) { snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); } \
    } while (0)

#define HEVCB_CUDA(ctx, call)                                                                          \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            HEVCB_SET_ERR(ctx, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return HEVCB_E_CUDA;                                                                       \
        }                                                                                              \
    } while (0)

// grow-only device buffer
static inline int hevcb_reserve(hevcb_ctx* ctx, hevcb_devbuf* b, size_t bytes)
{
    if (b->bytes >= bytes && b->p) { return HEVCB_OK; }
    if (b->p) { cudaFree(b->p); b->p = nullptr; b->bytes = 0; }
    size_t want = bytes + (bytes >> 3) + 256;
    cudaError_t e = cudaMalloc(&b->p, want);
    if (e != cudaSuccess) {
        HEVCB_SET_ERR(ctx, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        b->p = nullptr;
        return HEVCB_E_NOMEM;
    }
    b->bytes = want;
    return HEVCB_OK;
}

// kernels / launchers implemented in the .cu files
int hevcb_launch_scan_strip(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, int64_t* d_nal_start, int64_t* d_nal_end,
                            int64_t cap_nals, uint8_t* d_rbsp, int64_t* d_rbsp_off, int64_t* d_rbsp_end,
                            hevcb_scan_summary* d_summary, cudaStream_t stream);

int hevcb_launch_parse(hevcb_ctx* ctx, const uint8_t* d_buf, const int64_t* d_nal_start, const int64_t* d_nal_end, const uint8_t* d_rbsp,
                       const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n, const hevcb_parse_buffers* out,
                       hevcb_parse_summary* d_summary, const hevcb_parse_chain* chain, cudaStream_t stream);

int hevcb_launch_insert(hevcb_ctx* ctx, const uint8_t* d_rbsp, const int64_t* d_off, const int64_t* d_end, int64_t n, int sc_len, uint8_t* d_out,
                        int64_t out_cap, int64_t* d_out_off, hevcb_insert_summary* d_summary, cudaStream_t stream);

int hevcb_launch_scan_strip_shard(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t own, int64_t halo, int is_first, int is_last,
                                  int64_t* d_nal_start, int64_t* d_nal_end, int64_t cap_nals, uint8_t* d_rbsp, int64_t* d_rbsp_off,
                                  int64_t* d_rbsp_end, hevcb_shard_summary* d_summary, cudaStream_t stream);

int hevcb_launch_frame(hevcb_ctx* ctx, const uint8_t* d_buf, const int64_t* d_nal_start, const int64_t* d_nal_end, int64_t n, int sc_len, int len_size,
                       uint8_t* d_out, int64_t out_cap, int64_t* d_out_off, hevcb_insert_summary* d_summary, cudaStream_t stream);
int hevcb_launch_lenpref_index(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, int len_size, const int64_t* d_sample_off, int64_t n_samples,
                               int64_t* d_nal_start, int64_t* d_nal_end, int64_t cap_nals, int64_t* d_total, cudaStream_t stream);

int hevcb_launch_assemble3(hevcb_ctx* ctx, const uint8_t* raw_base, const int64_t* raw_off, const int64_t* raw_end, const uint8_t* a_base,
                           const int64_t* a_off, const int64_t* a_end, const uint8_t* b_base, const int64_t* b_off, const int64_t* b_end, int64_t n,
                           uint8_t* d_out, int64_t out_cap, int64_t* d_out_off, hevcb_insert_summary* d_summary, cudaStream_t stream);

int hevcb_launch_rewrite(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, const int64_t* d_nal_start, const int64_t* d_nal_end,
                         const uint8_t* d_rbsp, const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n, const hevcb_parse_buffers* parsed,
                         const hevcb_edit_set* edits, uint8_t* d_out, int64_t out_cap, int64_t* d_out_start, int64_t* d_out_end,
                         hevcb_rewrite_summary* d_summary, cudaStream_t stream);

int hevcb_launch_write_struct(hevcb_ctx* ctx, int32_t nal_hdr, const int32_t* d_vps, const int32_t* d_sps, const int32_t* d_pps, const int32_t* d_sh,
                              void* d_ctx_scratch, uint8_t* d_out, int64_t cap, int64_t* d_result, cudaStream_t stream);
size_t hevcb_write_struct_scratch_bytes();

int hevcb_sort_perm(hevcb_ctx* ctx, const uint32_t* d_keys, int64_t n, int32_t* d_perm, cudaStream_t stream);
