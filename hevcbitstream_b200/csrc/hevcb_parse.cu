// hevcb_parse.cu -- batched HEVC header parser for sm_100a: read_hevc_nal_unit (hevc_stream.c:155-241) for every NAL
// of a stream at once, working on the EPB-free image produced by hevcb_scan_strip_*.
//
// Dependency-ordered passes (a slice uses "the most recent SPS / PPS NAL before it", not an id lookup):
//   1. classify      one thread per NAL: 2-byte NAL header -> class (VPS / SPS / PPS / slice / unsupported / strip error)
//   2. ordinal scan  inclusive counts of SPS and PPS NALs -> for every NAL the ordinal of the last SPS / PPS before it
//   3. PS pass       VPS / SPS / PPS NALs (rare): full parse, count of syntax elements, and the compact per-ordinal
//                    context (hevcb_sps_ctx incl. derived RPS tables, hevcb_pps_ctx) kept in device memory
//   4. slice pass    slice segment headers: parse against the context of their SPS / PPS ordinal, count elements,
//                    write rc, header end and the SoA columns
//   5. offset scan   exclusive scan of the per-NAL element counts -> pair_off
//   6. emit          every parsed NAL walks its syntax once more and writes its (field, value) pairs at pair_off[k]
// The walker itself (bit reader with 64-bit window and clz exp-Golomb, all syntax structures) is hevcb_syntax.h.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <stdint.h>

#include "hevcb_internal.h"
#include "hevcb_syntax.h"

namespace {

constexpr int kCls_None = 0, kCls_Vps = 1, kCls_Sps = 2, kCls_Pps = 3, kCls_Slice = 4, kCls_Other = 5, kCls_Aux = 6, kCls_StripErr = 255;

// ---- generic 3-kernel scan over a per-element functor -------------------------------------------------------
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename T>
__device__ __forceinline__ T block_incl_scan(T v, T* warp_sums, T& block_total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) { v += o; }
    }
    if (lane == 31) { warp_sums[warp] = v; }
    __syncthreads();
    if (warp == 0) {
        T w = (lane < kScanThreads / 32) ? warp_sums[lane] : T(0);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T o = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= d) { w += o; }
        }
        if (lane < kScanThreads / 32) { warp_sums[lane] = w; }
    }
    __syncthreads();
    const T base = (warp > 0) ? warp_sums[warp - 1] : T(0);
    block_total = warp_sums[kScanThreads / 32 - 1];
    return v + base;
}

// value extractors
struct ClsIsSps { const uint8_t* cls; __device__ long long operator()(int64_t i) const { return cls[i] == kCls_Sps ? 1 : 0; } };
struct ClsIsPps { const uint8_t* cls; __device__ long long operator()(int64_t i) const { return cls[i] == kCls_Pps ? 1 : 0; } };
struct CntVal { const int32_t* cnt; __device__ long long operator()(int64_t i) const { return (long long)cnt[i]; } };
// header bytes that do not fit the per-NAL slot of the first write pass (they are written again, compactly, by the second)
constexpr int kHdrSlot = 128;
struct CntBig {
    const int32_t* cnt; const int64_t* slot_off;
    __device__ long long operator()(int64_t i) const { return (long long)cnt[i] > slot_off[i + 1] - slot_off[i] ? (long long)cnt[i] : 0ll; }
};
// room of a NAL's slot: kHdrSlot bytes for a slice header, the writer's whole capacity for a parameter set (their walk is the
// longest single thread of the pass, so it should not be repeated)
struct SlotCap {
    const uint8_t* cls; const int64_t* nal_start; const int64_t* nal_end; int64_t n; int slice_slot;
    __device__ long long operator()(int64_t i) const
    {
        if (i >= n) { return 0; }
        const int c = cls[i];
        if (c == kCls_Slice) { return slice_slot; }
        if (c == kCls_Vps || c == kCls_Sps || c == kCls_Pps) {
            const long long nsz = nal_end[i] - nal_start[i];
            return ((((nsz * 2 + 64) * 3) / 4) + 31) & ~15ll;
        }
        return 0;
    }
};

template <class F>
__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(F f, int64_t n, long long* block_sums)
{
    __shared__ long long ws[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    long long s = 0;
    for (int j = 0; j < kScanItems; j++) {
        const int64_t i = base + (int64_t)j * kScanThreads + threadIdx.x;
        if (i < n) { s += f(i); }
    }
    long long tot;
    block_incl_scan<long long>(s, ws, tot);
    if (threadIdx.x == 0) { block_sums[blockIdx.x] = tot; }
}

// single block: exclusive scan of block_sums in place; total written to block_sums[nb]
__global__ void __launch_bounds__(kScanThreads) scan_blocksums_kernel(long long* block_sums, int64_t nb)
{
    __shared__ long long ws[kScanThreads / 32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) { carry_s = 0; }
    __syncthreads();
    for (int64_t base = 0; base < nb; base += kScanThreads) {
        const int64_t i = base + threadIdx.x;
        const long long v = (i < nb) ? block_sums[i] : 0;
        long long tot;
        const long long inc = block_incl_scan<long long>(v, ws, tot);
        const long long carry = carry_s;
        __syncthreads();
        if (i < nb) { block_sums[i] = carry + inc - v; }
        if (threadIdx.x == 0) { carry_s = carry + tot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { block_sums[nb] = carry_s; }
}

// out[i] = (inclusive ? sum_{j<=i} : sum_{j<i}) f(j); blocked arrangement per thread for coalescing via striding
template <class F, typename TOut, bool kInclusive>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(F f, int64_t n, const long long* block_sums, TOut* out)
{
    __shared__ long long ws[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    long long carry = block_sums[blockIdx.x];
    for (int j = 0; j < kScanItems; j++) {
        const int64_t i = base + (int64_t)j * kScanThreads + threadIdx.x;
        const long long v = (i < n) ? f(i) : 0;
        long long tot;
        const long long inc = block_incl_scan<long long>(v, ws, tot);
        if (i < n) { out[i] = (TOut)(carry + (kInclusive ? inc : inc - v)); }
        carry += tot;
        __syncthreads();
    }
}

template <class F, typename TOut, bool kInclusive>
int run_scan(hevcb_ctx* ctx, F f, int64_t n, TOut* out, long long* block_sums /* nb + 1 */, cudaStream_t stream)
{
    const int64_t nb = (n + kScanTile - 1) / kScanTile;
    if (nb == 0) { return HEVCB_OK; }
    scan_reduce_kernel<F><<<(unsigned)nb, kScanThreads, 0, stream>>>(f, n, block_sums);
    scan_blocksums_kernel<<<1, kScanThreads, 0, stream>>>(block_sums, nb);
    scan_apply_kernel<F, TOut, kInclusive><<<(unsigned)nb, kScanThreads, 0, stream>>>(f, n, block_sums, out);
    ctx->launches += 3;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}

// ---- pass 1: classify ----------------------------------------------------------------------------------------
__global__ void classify_kernel(const uint8_t* __restrict__ rbsp, const int64_t* __restrict__ rbsp_off, const int64_t* __restrict__ rbsp_end,
                                int64_t n, uint8_t* __restrict__ cls, int32_t* __restrict__ nal_hdr, int32_t* __restrict__ rc,
                                uint8_t* __restrict__ kind, uint8_t* __restrict__ ubflag, int32_t* __restrict__ cnt, int32_t* __restrict__ hdr_end,
                                uint32_t* __restrict__ sortkey, uint32_t flags)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) { return; }
    const int64_t off = rbsp_off[k], end = rbsp_end[k];
    sortkey[k] = 0xFFFFFFFFu; // NALs the parser does not walk go last
    cnt[k] = 0;
    kind[k] = HEVCB_KIND_NONE;
    ubflag[k] = 0;
    hdr_end[k] = 0;
    rc[k] = -1;
    if (end < 0) { // nal_to_rbsp failed: read_hevc_nal_unit returns -1 before touching h->nal (hevc_stream.c:165-168)
        cls[k] = (uint8_t)kCls_StripErr;
        nal_hdr[k] = -1;
        return;
    }
    const int64_t size = end - off;
    const uint32_t b0 = size > 0 ? rbsp[off] : 0u, b1 = size > 1 ? rbsp[off + 1] : 0u;
    const uint32_t type = (b0 >> 1) & 0x3Fu;
    const uint32_t layer = ((b0 & 1u) << 5) | (b1 >> 3);
    const uint32_t tid = b1 & 7u;
    nal_hdr[k] = (int32_t)(type | (layer << 8) | (tid << 16));
    uint8_t c = (uint8_t)kCls_Other;
    if (hevcb_is_slice_type((int)type)) { c = (uint8_t)kCls_Slice; }
    else if (type == 32u) { c = (uint8_t)kCls_Vps; }
    else if (type == 33u) { c = (uint8_t)kCls_Sps; }
    else if (type == 34u) { c = (uint8_t)kCls_Pps; }
    else if ((flags & HEVCB_PARSE_AUX) && type >= 35u && type <= 40u) { c = (uint8_t)kCls_Aux; } // extension mode
    cls[k] = c;
    if (c != (uint8_t)kCls_Other) {
        // shape key: NAL type, then the first payload bytes (first_slice_segment_in_pic_flag, ids, slice_type, ... live there):
        // NALs with equal keys take the same branches for most of the header
        const uint32_t b2 = size > 2 ? rbsp[off + 2] : 0u, b3 = size > 3 ? rbsp[off + 3] : 0u, b4 = size > 4 ? rbsp[off + 4] : 0u;
        sortkey[k] = (type << 24) | (b2 << 16) | (b3 << 8) | b4;
    }
}

// consumed bytes reported by read_hevc_nal_unit: nal_size, minus one when a trailing 00 00 03 was dropped (h264_nal.c:170-173,197)
__device__ __forceinline__ int32_t consumed_bytes(const uint8_t* __restrict__ buf, int64_t start, int64_t end, int64_t buf_size)
{
    const int64_t size = end - start;
    if (end > buf_size) { return (int32_t)size; } // ends in a later shard: corrected by the caller (hevcb_stitch_patch.ends_003)
    if (size >= 3 && buf[end - 1] == 3 && buf[end - 2] == 0 && buf[end - 3] == 0) { return (int32_t)(size - 1); }
    return (int32_t)size;
}

// ---- passes 3/4/6: parse (count or emit) ------------------------------------------------------------------------
struct ParseArgs {
    const uint8_t* buf;
    int64_t buf_size;
    const int64_t* nal_start;
    const int64_t* nal_end;
    const uint8_t* rbsp;
    const int64_t* rbsp_off;
    const int64_t* rbsp_end;
    int64_t n;
    const uint8_t* cls;
    const int32_t* sps_ord; // inclusive count of SPS NALs up to and including k
    const int32_t* pps_ord;
    const int32_t* perm;    // thread i takes NAL perm[i]: NALs of the same shape side by side (hevcb_sort.cu)
    hevcb_sps_ctx* sps_tab;  // [n_sps + 1]; entry 0 = the zeroed state before any SPS (calloc in hevc_new)
    hevcb_pps_ctx* pps_tab;  // [n_pps + 1]
    hevcb_sps_ctx* sps_scratch; // emit pass: where re-parsed SPS contexts go (same shape as sps_tab)
    hevcb_pps_ctx* pps_scratch;
    int32_t* rc;
    uint8_t* kind;
    uint8_t* ubflag;
    int32_t* cnt;
    int32_t* hdr_end;
    int32_t* cols; // 8 columns x n
    const int64_t* pair_off;
    uint32_t* pair_field;
    int32_t* pair_value;
    uint32_t* pair_pos; // trace mode (read_debug variant): bit position of every record
    int64_t cap_pairs;
    bool spec;          // HEVCB_PARSE_SPEC: the spec-correct walk (slices resolve their parameter sets by id over the whole tables)
};

// kTrace: the read_debug variant of the walk (hevcb_sink_t<true>): the lists hold what read_debug_hevc_nal_unit prints; NALs of
// unsupported types then contribute their four NAL header lines (they are taken with the slices).
// (eight blocks per SM: the walk is bound by the latency of scattered header reads, so occupancy pays more than the 30 registers it
// costs: 1 M distinct headers 9.2 -> 8.05 ms)
template <bool kEmit, bool kSlices, bool kTrace>
__global__ void __launch_bounds__(128, 8) parse_kernel(ParseArgs a)
{
    const int64_t ti = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= a.n) { return; }
    const int64_t k = a.perm[a.n - 1 - ti]; // shape order, last shape first: the long parameter-set walks start with the first blocks
    const int c = a.cls[k];
    const bool is_ps = (c == kCls_Vps || c == kCls_Sps || c == kCls_Pps);
    const bool is_slice = (c == kCls_Slice) || (c == kCls_Aux) || (kTrace && c == kCls_Other); // no dependencies among these: one pass
    if (!kEmit) {
        if (kSlices ? !is_slice : !is_ps) { return; }
    } else {
        if (!is_ps && !is_slice) { return; }
    }
    const int64_t off = a.rbsp_off[k];
    const int64_t size = a.rbsp_end[k] - off;
    // ordinals: entry 0 of the tables is the all-zero state, SPS / PPS number j lives at index j (1-based)
    const int sps_count = a.sps_ord[k], pps_count = a.pps_ord[k];
    const hevcb_sps_ctx* sps_in = &a.sps_tab[c == kCls_Sps ? sps_count - 1 : sps_count];
    const hevcb_pps_ctx* pps_in = &a.pps_tab[c == kCls_Pps ? pps_count - 1 : pps_count];
    hevcb_sps_ctx* sps_out = nullptr;
    hevcb_pps_ctx* pps_out = nullptr;
    if (c == kCls_Sps) { sps_out = kEmit ? &a.sps_scratch[sps_count] : &a.sps_tab[sps_count]; }
    if (c == kCls_Pps) { pps_out = kEmit ? &a.pps_scratch[pps_count] : &a.pps_tab[pps_count]; }
    hevcb_nal_result r;
    typedef hevcb_sink_t<kTrace> SinkT;
    const hevcb_ps_lookup lk{a.sps_tab, a.pps_tab, sps_count, pps_count}; // spec mode, slices: every SPS / PPS NAL in front of this one
    const hevcb_ps_lookup* lkp = (a.spec && is_slice) ? &lk : nullptr;
    if (!kEmit) {
        SinkT sink{nullptr, nullptr, 0, nullptr};
        hevcb_parse_nal(a.rbsp + off, size, sink, sps_in, pps_in, sps_out, pps_out, r, c == kCls_Aux, a.spec, lkp);
        a.cnt[k] = (int32_t)sink.n;
        a.kind[k] = (uint8_t)r.kind;
        a.ubflag[k] = (uint8_t)(r.flags & 0xFFu);
        a.rc[k] = r.ok ? consumed_bytes(a.buf, a.nal_start[k], a.nal_end[k], a.buf_size) : -1;
        a.hdr_end[k] = kSlices ? r.hdr_end : (int32_t)r.end_bits;
        if (kSlices) {
            a.cols[0 * a.n + k] = r.cols.slice_type;
            a.cols[1 * a.n + k] = r.cols.slice_qp_delta;
            a.cols[2 * a.n + k] = r.cols.slice_pic_order_cnt_lsb;
            a.cols[3 * a.n + k] = r.cols.first_slice_segment_in_pic_flag;
            a.cols[4 * a.n + k] = r.cols.slice_segment_address;
            a.cols[5 * a.n + k] = r.cols.dependent_slice_segment_flag;
            a.cols[6 * a.n + k] = r.cols.num_entry_point_offsets;
            a.cols[7 * a.n + k] = r.cols.short_term_ref_pic_set_idx;
        }
    } else {
        const int64_t po = a.pair_off[k];
        const int64_t pn = (int64_t)a.cnt[k];
        if (pn == 0 || po + pn > a.cap_pairs) { return; }
        SinkT sink{a.pair_field + po, a.pair_value + po, 0, kTrace ? a.pair_pos + po : nullptr};
        hevcb_parse_nal(a.rbsp + off, size, sink, sps_in, pps_in, sps_out, pps_out, r, c == kCls_Aux, a.spec, lkp);
        if (sink.n != (uint32_t)pn) { a.ubflag[k] |= 0x80u; a.cols[k] = (int32_t)r.end_bits; a.cols[a.n + k] = (int32_t)sink.n; } // self-check
    }
}

struct ParseSummaryDev {
    int64_t n_nals, n_ok, n_pairs, n_vps, n_sps, n_pps, n_slices;
    int32_t overflow, pad;
};

__global__ void parse_summary_kernel(const uint8_t* __restrict__ cls, const int32_t* __restrict__ rc, int64_t n, const long long* pair_total,
                                     int64_t cap_pairs, hevcb_parse_summary* out)
{
    __shared__ unsigned long long acc[5];
    if (threadIdx.x < 5) { acc[threadIdx.x] = 0; }
    __syncthreads();
    unsigned long long ok = 0, v = 0, s = 0, p = 0, sl = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        ok += rc[k] >= 0;
        const int c = cls[k];
        v += c == kCls_Vps; s += c == kCls_Sps; p += c == kCls_Pps; sl += c == kCls_Slice;
    }
    atomicAdd(&acc[0], ok); atomicAdd(&acc[1], v); atomicAdd(&acc[2], s); atomicAdd(&acc[3], p); atomicAdd(&acc[4], sl);
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicAdd((unsigned long long*)&out->n_ok, acc[0]);
        atomicAdd((unsigned long long*)&out->n_vps, acc[1]);
        atomicAdd((unsigned long long*)&out->n_sps, acc[2]);
        atomicAdd((unsigned long long*)&out->n_pps, acc[3]);
        atomicAdd((unsigned long long*)&out->n_slices, acc[4]);
        if (blockIdx.x == 0) {
            out->n_nals = n;
            out->n_pairs = *pair_total;
            out->overflow = (*pair_total > cap_pairs) ? 1 : 0;
        }
    }
}

// ---- header rewrite: write pass (count / emit), part composition, epilogue ----------------------------------------
struct RewriteArgs {
    const int64_t* nal_start;
    const int64_t* nal_end;
    const int64_t* rbsp_off;
    const int64_t* rbsp_end;
    int64_t n, size;
    const uint8_t* cls;
    const int32_t* sps_ord;
    const int32_t* pps_ord;
    const hevcb_sps_ctx* sps_tab;
    const hevcb_pps_ctx* pps_tab;
    hevcb_sps_ctx* sps_scratch;
    const int32_t* rc;
    const int32_t* nal_hdr;
    const uint8_t* ubflag;
    const int32_t* hdr_end;
    const int64_t* pair_off;
    const uint32_t* pair_field;
    const int32_t* pair_value;
    int64_t cap_pairs;
    const int32_t* perm; // thread order of the parse (same-shape NALs side by side)
    int32_t* wlen;       // [n] bytes of the written header part (0: the NAL is copied through)
    const int64_t* woff; // exclusive scan of the header lengths beyond kHdrSlot
    uint8_t* staging;    // headers longer than kHdrSlot, compact (second write pass)
    uint8_t* slots;      // the first write pass keeps the first bytes of every header in the NAL's slot
    const int64_t* slot_off; // [n + 2] exclusive scan of the slot sizes (SlotCap)
    int64_t *raw_off, *raw_end, *a_off, *a_end, *b_off, *b_end; // [n + 1]
    hevcb_edit_set edits;
    bool spec; // the parse ran with HEVCB_PARSE_SPEC: the write walk follows the same rules (slices resolve their parameter sets by id)
};

template <bool kEmit>
__global__ void __launch_bounds__(128) write_kernel(RewriteArgs a)
{
    const int64_t ti = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ti >= a.n) { return; }
    // shape order, last shape first: the parameter sets (highest keys) are the longest single-thread walks of the pass and
    // start with the first blocks instead of forming its tail
    const int64_t k = a.perm[a.n - 1 - ti];
    const int c = a.cls[k];
    const bool is_slice = (c == kCls_Slice);
    if (!kEmit) { a.wlen[k] = 0; }
    if (!(is_slice || c == kCls_Vps || c == kCls_Sps || c == kCls_Pps)) { return; }
    const int64_t room = a.slot_off[k + 1] - a.slot_off[k];
    if (kEmit && (int64_t)a.wlen[k] <= room) { return; } // not rewritten, or complete in its slot
    const int64_t po = a.pair_off[k];
    const int64_t pn = a.pair_off[k + 1] - po;
    if (!kEmit) {
        const int64_t rs = a.rbsp_end[k] - a.rbsp_off[k];
        if (a.rc[k] < 0 || a.ubflag[k] != 0 || po + pn > a.cap_pairs || (is_slice && (int64_t)a.hdr_end[k] > rs)) { return; }
    }
    const int sps_count = a.sps_ord[k], pps_count = a.pps_ord[k];
    const hevcb_sps_ctx* sps_in = &a.sps_tab[c == kCls_Sps ? sps_count - 1 : sps_count];
    const hevcb_pps_ctx* pps_in = &a.pps_tab[c == kCls_Pps ? pps_count - 1 : pps_count];
    hevcb_sps_ctx* scr = nullptr;
    if (c == kCls_Sps) { // working copy for the derived RPS tables, zeroed like the struct the reference starts from
        scr = &a.sps_scratch[sps_count];
        int32_t* z = reinterpret_cast<int32_t*>(scr);
        for (size_t i = 0; i < sizeof(hevcb_sps_ctx) / 4; i++) { z[i] = 0; }
    }
    const int64_t nsz = a.nal_end[k] - a.nal_start[k];
    const int64_t wcap = ((is_slice ? (int64_t)16384 : nsz * 2 + 64) * 3) / 4; // write_hevc_nal_unit: rbsp_size = size * 3 / 4 (hevc_stream.c:1266)
    const int kind = is_slice ? HEVCB_KIND_SLICE : (c == kCls_Vps ? HEVCB_KIND_VPS : (c == kCls_Sps ? HEVCB_KIND_SPS : HEVCB_KIND_PPS));
    hevcb_replay rp{a.pair_field + po, a.pair_value + po, (uint32_t)pn, 0u, kind, &a.edits};
    hevcb_bitwriter bw;
    if (kEmit) { bw.init(a.staging + a.woff[k], (int64_t)a.wlen[k]); } else { bw.init(a.slots + a.slot_off[k], wcap, room); }
    hevcb_write_result wr;
    const hevcb_ps_lookup lk{a.sps_tab, a.pps_tab, sps_count, pps_count};
    hevcb_write_nal(rp, bw, a.nal_hdr[k], sps_in, pps_in, scr, wr, a.spec, (a.spec && is_slice) ? &lk : nullptr);
    if (!kEmit) {
        const int64_t len = is_slice ? (int64_t)wr.hdr_bytes : wr.bytes;
        a.wlen[k] = (wr.ok && wr.bytes > 0 && len > 0 && len < (1 << 30)) ? (int32_t)len : 0;
    }
}

__global__ void compose_parts_kernel(RewriteArgs a)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k > a.n) { return; }
    const int64_t prev_end = k > 0 ? a.nal_end[k - 1] : 0;
    int64_t ro = prev_end, re, ao = 0, ae = 0, bo = 0, be = 0;
    if (k == a.n) {
        re = a.size; // what follows the last NAL
    } else if (a.wlen[k] > 0) {
        re = a.nal_start[k];
        // absolute addresses (the assembly is given a null base): the slot of the first write pass, or the compact staging
        ao = (int64_t)(uintptr_t)((int64_t)a.wlen[k] <= a.slot_off[k + 1] - a.slot_off[k] ? a.slots + a.slot_off[k] : a.staging + a.woff[k]);
        ae = ao + a.wlen[k];
        if (a.cls[k] == kCls_Slice) { bo = a.rbsp_off[k] + a.hdr_end[k]; be = a.rbsp_end[k]; }
    } else {
        re = a.nal_end[k]; // copied through together with what precedes it
    }
    if (re < ro) { re = ro; }
    a.raw_off[k] = ro; a.raw_end[k] = re; a.a_off[k] = ao; a.a_end[k] = ae; a.b_off[k] = bo; a.b_end[k] = be;
}

__global__ void rewrite_finish_kernel(RewriteArgs a, const int64_t* __restrict__ out_off, const hevcb_insert_summary* __restrict__ isum,
                                      int64_t* __restrict__ out_start, int64_t* __restrict__ out_end, unsigned long long* n_rewritten,
                                      hevcb_rewrite_summary* summary, int final_pass)
{
    if (final_pass) {
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            summary->n_nals = a.n;
            summary->n_rewritten = (int64_t)*n_rewritten;
            summary->out_bytes = isum->out_bytes;
            summary->n_inserted = isum->n_inserted;
            summary->overflow = isum->overflow;
            summary->pad = 0;
        }
        return;
    }
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = k < a.n;
    int rew = 0;
    if (live) {
        const int64_t prev_end = k > 0 ? a.nal_end[k - 1] : 0;
        const int64_t gap = a.nal_start[k] > prev_end ? a.nal_start[k] - prev_end : 0;
        out_start[k] = out_off[k] + gap;
        out_end[k] = out_off[k + 1];
        rew = a.wlen[k] > 0 ? 1 : 0;
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, rew != 0);
    if ((threadIdx.x & 31) == 0 && m) { atomicAdd(n_rewritten, (unsigned long long)__popc(m)); }
}

// ---- write_hevc_nal_unit from caller-owned structs (one NAL, one thread): the compatibility path ------------------------
struct WriteStructArgs {
    const int32_t* vps;
    const int32_t* sps;
    const int32_t* pps;
    const int32_t* sh;
    int32_t nal_hdr;
    hevcb_sps_ctx* sps_ctx; // derived from the SPS struct the way the writer walks it
    hevcb_sps_ctx* sps_scr;
    uint8_t* out;
    int64_t cap;            // RBSP room: size * 3 / 4 of the caller's buffer (hevc_stream.c:1266)
    int64_t* result;        // [0] RBSP bytes, [1] ok
};

__global__ void write_struct_kernel(WriteStructArgs a)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) { return; }
    int32_t* z = reinterpret_cast<int32_t*>(a.sps_ctx);
    for (size_t i = 0; i < sizeof(hevcb_sps_ctx) / 4; i++) { z[i] = 0; }
    z = reinterpret_cast<int32_t*>(a.sps_scr);
    for (size_t i = 0; i < sizeof(hevcb_sps_ctx) / 4; i++) { z[i] = 0; }
    hevcb_bits nob;
    nob.init(nullptr, 0);
    hevcb_sink nos{nullptr, nullptr, 0};
    hevcb_pps_ctx pc;
    { // what a slice needs from the current SPS / PPS: walk the structs like the writer does, writing nowhere
        hevcb_replay rp{nullptr, nullptr, 0u, 0u, HEVCB_KIND_SPS, nullptr, a.sps};
        hevcb_bitwriter bw;
        bw.init(nullptr, (int64_t)1 << 40);
        hevcb_walker<hevcb_sink, true> w(nob, nos, &bw, &rp);
        w.seq_parameter_set(*a.sps_ctx);
        hevcb_replay rq{nullptr, nullptr, 0u, 0u, HEVCB_KIND_PPS, nullptr, a.pps};
        hevcb_bitwriter bq;
        bq.init(nullptr, (int64_t)1 << 40);
        hevcb_walker<hevcb_sink, true> wq(nob, nos, &bq, &rq);
        wq.pic_parameter_set(pc);
    }
    const int t = a.nal_hdr & 0xFF;
    const int32_t* dense = hevcb_is_slice_type(t) ? a.sh : (t == 32 ? a.vps : (t == 33 ? a.sps : a.pps));
    const int kind = hevcb_is_slice_type(t) ? HEVCB_KIND_SLICE : (t == 32 ? HEVCB_KIND_VPS : (t == 33 ? HEVCB_KIND_SPS : HEVCB_KIND_PPS));
    hevcb_replay rp{nullptr, nullptr, 0u, 0u, kind, nullptr, dense};
    hevcb_bitwriter bw;
    bw.init(a.out, a.cap);
    hevcb_write_result wr;
    hevcb_write_nal(rp, bw, a.nal_hdr, a.sps_ctx, &pc, a.sps_scr, wr);
    a.result[0] = wr.bytes;
    a.result[1] = wr.ok;
}

} // namespace

int hevcb_launch_write_struct(hevcb_ctx* ctx, int32_t nal_hdr, const int32_t* d_vps, const int32_t* d_sps, const int32_t* d_pps, const int32_t* d_sh,
                              void* d_ctx_scratch, uint8_t* d_out, int64_t cap, int64_t* d_result, cudaStream_t stream)
{
    WriteStructArgs a;
    a.vps = d_vps; a.sps = d_sps; a.pps = d_pps; a.sh = d_sh; a.nal_hdr = nal_hdr;
    a.sps_ctx = reinterpret_cast<hevcb_sps_ctx*>(d_ctx_scratch);
    a.sps_scr = a.sps_ctx + 1;
    a.out = d_out; a.cap = cap; a.result = d_result;
    write_struct_kernel<<<1, 32, 0, stream>>>(a);
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}

size_t hevcb_write_struct_scratch_bytes() { return 2 * sizeof(hevcb_sps_ctx); }

int hevcb_launch_rewrite(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, const int64_t* d_nal_start, const int64_t* d_nal_end,
                         const uint8_t* d_rbsp, const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n, const hevcb_parse_buffers* parsed,
                         const hevcb_edit_set* edits, uint8_t* d_out, int64_t out_cap, int64_t* d_out_start, int64_t* d_out_end,
                         hevcb_rewrite_summary* d_summary, cudaStream_t stream)
{
    if (n < 0 || size < 0 || !parsed || !d_summary || !d_out || out_cap < 0 || (edits && (edits->n < 0 || edits->n > HEVCB_MAX_EDITS)) ||
        (n > 0 && (!d_buf || !d_nal_start || !d_nal_end || !d_rbsp || !d_rbsp_off || !d_rbsp_end || !d_out_start || !d_out_end))) {
        HEVCB_SET_ERR(ctx, "hevcb_rewrite: invalid argument");
        return HEVCB_E_ARG;
    }
    if (ctx->last_parse.n != n) {
        HEVCB_SET_ERR(ctx, "hevcb_rewrite: must follow hevcb_parse_device of the same %lld NALs on this context", (long long)n);
        return HEVCB_E_ARG;
    }
    const int64_t m = n + 1; // + the bytes that follow the last NAL
    const int64_t nb = (m + kScanTile - 1) / kScanTile;
    size_t need = 0;
    auto take = [&](size_t bytes) { size_t o = need; need += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_wlen = take((size_t)m * 4), o_woff = take((size_t)(m + 1) * 8), o_parts = take((size_t)m * 8 * 6), o_ooff = take((size_t)(m + 1) * 8);
    const size_t o_soff = take((size_t)(m + 2) * 8);
    const size_t o_bs = take((size_t)(nb + 3) * 8), o_isum = take(sizeof(hevcb_insert_summary)), o_cnt = take(64);
    int rcx = hevcb_reserve(ctx, &ctx->rewrite_scratch, need);
    if (rcx != HEVCB_OK) { return rcx; }
    uint8_t* base = reinterpret_cast<uint8_t*>(ctx->rewrite_scratch.p);
    RewriteArgs a;
    a.nal_start = d_nal_start; a.nal_end = d_nal_end; a.rbsp_off = d_rbsp_off; a.rbsp_end = d_rbsp_end; a.n = n; a.size = size;
    a.cls = reinterpret_cast<const uint8_t*>(ctx->last_parse.cls);
    a.sps_ord = reinterpret_cast<const int32_t*>(ctx->last_parse.sps_ord);
    a.pps_ord = reinterpret_cast<const int32_t*>(ctx->last_parse.pps_ord);
    a.sps_tab = reinterpret_cast<const hevcb_sps_ctx*>(ctx->last_parse.sps_tab);
    a.pps_tab = reinterpret_cast<const hevcb_pps_ctx*>(ctx->last_parse.pps_tab);
    a.sps_scratch = reinterpret_cast<hevcb_sps_ctx*>(ctx->last_parse.sps_scratch);
    a.perm = reinterpret_cast<const int32_t*>(ctx->last_parse.perm);
    a.rc = parsed->rc; a.nal_hdr = parsed->nal_hdr; a.ubflag = parsed->ubflag; a.hdr_end = parsed->hdr_end;
    a.pair_off = parsed->pair_off; a.pair_field = parsed->pair_field; a.pair_value = parsed->pair_value; a.cap_pairs = parsed->cap_pairs;
    a.wlen = reinterpret_cast<int32_t*>(base + o_wlen);
    int64_t* woff = reinterpret_cast<int64_t*>(base + o_woff);
    a.woff = woff;
    a.staging = nullptr;
    a.slots = nullptr;
    a.slot_off = reinterpret_cast<const int64_t*>(base + o_soff);
    int64_t* parts = reinterpret_cast<int64_t*>(base + o_parts);
    a.raw_off = parts; a.raw_end = parts + m; a.a_off = parts + 2 * m; a.a_end = parts + 3 * m; a.b_off = parts + 4 * m; a.b_end = parts + 5 * m;
    if (edits) { a.edits = *edits; } else { a.edits.n = 0; }
    a.spec = ctx->last_parse.spec != 0;
    int64_t* out_off = reinterpret_cast<int64_t*>(base + o_ooff);
    long long* bsums = reinterpret_cast<long long*>(base + o_bs);
    hevcb_insert_summary* isum = reinterpret_cast<hevcb_insert_summary*>(base + o_isum);
    unsigned long long* n_rew = reinterpret_cast<unsigned long long*>(base + o_cnt);
    HEVCB_CUDA(ctx, cudaMemsetAsync(n_rew, 0, 8, stream));
    HEVCB_CUDA(ctx, cudaMemsetAsync(a.wlen, 0, (size_t)m * 4, stream));

    const unsigned g128 = (unsigned)((m + 127) / 128);
    long long h_total = 0;
    {
        // slot sizes -> offsets; the slot buffer's size is known after the scan
        long long h_slots = 0;
        const int64_t ns = m + 1, nbs = (ns + kScanTile - 1) / kScanTile;
        int slice_slot = kHdrSlot; // HEVCB_HDR_SLOT: test switch (small slots send the slice headers through the second pass)
        if (const char* e = getenv("HEVCB_HDR_SLOT")) { const int v = atoi(e); if (v >= 0 && v <= 4096) { slice_slot = v & ~15; } }
        int rcs = run_scan<SlotCap, int64_t, false>(ctx, SlotCap{a.cls, d_nal_start, d_nal_end, n, slice_slot}, ns, reinterpret_cast<int64_t*>(base + o_soff), bsums, stream);
        if (rcs != HEVCB_OK) { return rcs; }
        HEVCB_CUDA(ctx, cudaMemcpyAsync(&h_slots, bsums + nbs, 8, cudaMemcpyDeviceToHost, stream));
        HEVCB_CUDA(ctx, cudaStreamSynchronize(stream));
        if ((rcx = hevcb_reserve(ctx, &ctx->rewrite_slots, (size_t)h_slots + 64)) != HEVCB_OK) { return rcx; }
        a.slots = reinterpret_cast<uint8_t*>(ctx->rewrite_slots.p);
    }
    if (n > 0) {
        write_kernel<false><<<g128, 128, 0, stream>>>(a);
        ctx->launches++;
        HEVCB_CUDA(ctx, cudaGetLastError());
        int rcs = run_scan<CntBig, int64_t, false>(ctx, CntBig{a.wlen, a.slot_off}, m, woff, bsums, stream);
        if (rcs != HEVCB_OK) { return rcs; }
        HEVCB_CUDA(ctx, cudaMemcpyAsync(&h_total, bsums + nb, 8, cudaMemcpyDeviceToHost, stream));
        HEVCB_CUDA(ctx, cudaStreamSynchronize(stream)); // the staging size depends on the total header bytes
    }
    if ((rcx = hevcb_reserve(ctx, &ctx->rewrite_staging, (size_t)h_total + 64)) != HEVCB_OK) { return rcx; }
    a.staging = reinterpret_cast<uint8_t*>(ctx->rewrite_staging.p);
    if (n > 0) {
        write_kernel<true><<<g128, 128, 0, stream>>>(a);
        ctx->launches++;
    }
    compose_parts_kernel<<<g128, 128, 0, stream>>>(a);
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    int rca = hevcb_launch_assemble3(ctx, d_buf, a.raw_off, a.raw_end, nullptr, a.a_off, a.a_end, d_rbsp, a.b_off, a.b_end, m, d_out, out_cap, out_off,
                                     isum, stream);
    if (rca != HEVCB_OK) { return rca; }
    if (n > 0) {
        rewrite_finish_kernel<<<g128, 128, 0, stream>>>(a, out_off, isum, d_out_start, d_out_end, n_rew, d_summary, 0);
        ctx->launches++;
    }
    rewrite_finish_kernel<<<1, 32, 0, stream>>>(a, out_off, isum, d_out_start, d_out_end, n_rew, d_summary, 1);
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}

int hevcb_launch_parse(hevcb_ctx* ctx, const uint8_t* d_buf, const int64_t* d_nal_start, const int64_t* d_nal_end, const uint8_t* d_rbsp,
                       const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n, const hevcb_parse_buffers* out,
                       hevcb_parse_summary* d_summary, const hevcb_parse_chain* chain, cudaStream_t stream)
{
    if (n < 0 || !out || !d_summary || (n > 0 && (!d_buf || !d_nal_start || !d_nal_end || !d_rbsp || !d_rbsp_off || !d_rbsp_end))) {
        HEVCB_SET_ERR(ctx, "hevcb_parse: invalid argument");
        return HEVCB_E_ARG;
    }
    if (n > 0 && (!out->rc || !out->nal_hdr || !out->kind || !out->ubflag || !out->hdr_end || !out->cols || !out->pair_off ||
                  (out->cap_pairs > 0 && (!out->pair_field || !out->pair_value)))) {
        HEVCB_SET_ERR(ctx, "hevcb_parse: output buffers missing");
        return HEVCB_E_ARG;
    }
    if (chain && (out->flags & HEVCB_PARSE_SPEC)) {
        HEVCB_SET_ERR(ctx, "hevcb_parse: the spec-correct mode is not available for shard parses (a slice may refer to any earlier parameter set)");
        return HEVCB_E_ARG;
    }
    HEVCB_CUDA(ctx, cudaMemsetAsync(d_summary, 0, sizeof(hevcb_parse_summary), stream));
    ctx->last_parse.n = -1;
    if (n == 0) {
        ctx->last_parse.n = 0;
        if (chain) { // nothing in the shard changes the state
            if (chain->sps_out) { if (chain->sps_in) { memcpy(chain->sps_out, chain->sps_in, sizeof(hevcb_sps_ctx)); } else { memset(chain->sps_out, 0, sizeof(hevcb_sps_ctx)); } }
            if (chain->pps_out) { if (chain->pps_in) { memcpy(chain->pps_out, chain->pps_in, sizeof(hevcb_pps_ctx)); } else { memset(chain->pps_out, 0, sizeof(hevcb_pps_ctx)); } }
        }
        return HEVCB_OK;
    }
    const int64_t nb = (n + kScanTile - 1) / kScanTile;
    // scratch: cls (n), cnt (n x4), sps_ord (n x4), pps_ord (n x4), block sums (nb + 1) x3, pair total
    size_t need = 0;
    auto take = [&](size_t bytes) { size_t o = need; need += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_cls = take((size_t)n), o_cnt = take((size_t)n * 4), o_so = take((size_t)n * 4), o_po = take((size_t)n * 4);
    const size_t o_key = take((size_t)n * 4), o_perm = take((size_t)n * 4);
    const size_t o_bs = take((size_t)(nb + 1) * 8), o_tot = take(64);
    int rcx = hevcb_reserve(ctx, &ctx->parse_scratch, need);
    if (rcx != HEVCB_OK) { return rcx; }
    uint8_t* base = reinterpret_cast<uint8_t*>(ctx->parse_scratch.p);
    uint8_t* cls = base + o_cls;
    int32_t* cnt = reinterpret_cast<int32_t*>(base + o_cnt);
    int32_t* sps_ord = reinterpret_cast<int32_t*>(base + o_so);
    int32_t* pps_ord = reinterpret_cast<int32_t*>(base + o_po);
    long long* bsums = reinterpret_cast<long long*>(base + o_bs);
    long long* pair_total = reinterpret_cast<long long*>(base + o_tot);
    uint32_t* sortkey = reinterpret_cast<uint32_t*>(base + o_key);
    int32_t* perm = reinterpret_cast<int32_t*>(base + o_perm);

    const unsigned g128 = (unsigned)((n + 127) / 128);
    classify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_rbsp, d_rbsp_off, d_rbsp_end, n, cls, out->nal_hdr, out->rc, out->kind,
                                                                    out->ubflag, cnt, out->hdr_end, sortkey, out->flags);
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    int rcp = hevcb_sort_perm(ctx, sortkey, n, perm, stream);
    if (rcp != HEVCB_OK) { return rcp; }
    int rcs = run_scan<ClsIsSps, int32_t, true>(ctx, ClsIsSps{cls}, n, sps_ord, bsums, stream);
    if (rcs != HEVCB_OK) { return rcs; }
    long long h_counts[2] = {0, 0};
    HEVCB_CUDA(ctx, cudaMemcpyAsync(&h_counts[0], bsums + nb, 8, cudaMemcpyDeviceToHost, stream));
    HEVCB_CUDA(ctx, cudaStreamSynchronize(stream)); // the table sizes below depend on the counts
    rcs = run_scan<ClsIsPps, int32_t, true>(ctx, ClsIsPps{cls}, n, pps_ord, bsums, stream);
    if (rcs != HEVCB_OK) { return rcs; }
    HEVCB_CUDA(ctx, cudaMemcpyAsync(&h_counts[1], bsums + nb, 8, cudaMemcpyDeviceToHost, stream));
    HEVCB_CUDA(ctx, cudaStreamSynchronize(stream));
    const size_t n_sps = (size_t)h_counts[0], n_pps = (size_t)h_counts[1];
    // context tables: entry 0 is the zeroed initial state; a second copy receives the emit pass' re-parse
    const size_t sps_bytes = (n_sps + 1) * sizeof(hevcb_sps_ctx), pps_bytes = (n_pps + 1) * sizeof(hevcb_pps_ctx);
    if ((rcx = hevcb_reserve(ctx, &ctx->parse_ps, 2 * sps_bytes + 2 * pps_bytes + 1024)) != HEVCB_OK) { return rcx; }
    uint8_t* pb = reinterpret_cast<uint8_t*>(ctx->parse_ps.p);
    HEVCB_CUDA(ctx, cudaMemsetAsync(pb, 0, 2 * sps_bytes + 2 * pps_bytes, stream));
    if (chain && chain->sps_in) { HEVCB_CUDA(ctx, cudaMemcpyAsync(pb, chain->sps_in, sizeof(hevcb_sps_ctx), cudaMemcpyHostToDevice, stream)); }
    if (chain && chain->pps_in) { HEVCB_CUDA(ctx, cudaMemcpyAsync(pb + 2 * sps_bytes, chain->pps_in, sizeof(hevcb_pps_ctx), cudaMemcpyHostToDevice, stream)); }
    ParseArgs a;
    a.buf_size = (chain && chain->buf_size > 0) ? chain->buf_size : 0x7FFFFFFFFFFFFFFFll;
    a.buf = d_buf; a.nal_start = d_nal_start; a.nal_end = d_nal_end; a.rbsp = d_rbsp; a.rbsp_off = d_rbsp_off; a.rbsp_end = d_rbsp_end;
    a.n = n; a.cls = cls; a.sps_ord = sps_ord; a.pps_ord = pps_ord; a.perm = perm;
    a.sps_tab = reinterpret_cast<hevcb_sps_ctx*>(pb);
    a.sps_scratch = reinterpret_cast<hevcb_sps_ctx*>(pb + sps_bytes);
    a.pps_tab = reinterpret_cast<hevcb_pps_ctx*>(pb + 2 * sps_bytes);
    a.pps_scratch = reinterpret_cast<hevcb_pps_ctx*>(pb + 2 * sps_bytes + pps_bytes);
    a.rc = out->rc; a.kind = out->kind; a.ubflag = out->ubflag; a.cnt = cnt; a.hdr_end = out->hdr_end; a.cols = out->cols;
    a.pair_off = out->pair_off; a.pair_field = out->pair_field; a.pair_value = out->pair_value; a.cap_pairs = out->cap_pairs;
    a.spec = (out->flags & HEVCB_PARSE_SPEC) != 0u;

    ctx->last_parse.n = n;
    ctx->last_parse.spec = a.spec ? 1 : 0;
    ctx->last_parse.cls = cls; ctx->last_parse.sps_ord = sps_ord; ctx->last_parse.pps_ord = pps_ord; ctx->last_parse.cnt = cnt;
    ctx->last_parse.perm = perm;
    ctx->last_parse.sps_tab = a.sps_tab; ctx->last_parse.pps_tab = a.pps_tab; ctx->last_parse.sps_scratch = a.sps_scratch;
    const bool trace = out->pair_pos != nullptr;
    a.pair_pos = out->pair_pos;
    if (trace) {
        parse_kernel<false, false, true><<<g128, 128, 0, stream>>>(a);
        parse_kernel<false, true, true><<<g128, 128, 0, stream>>>(a);
    } else {
        parse_kernel<false, false, false><<<g128, 128, 0, stream>>>(a); // parameter sets first
        parse_kernel<false, true, false><<<g128, 128, 0, stream>>>(a);  // then the slices that depend on them
    }
    ctx->launches += 2;
    HEVCB_CUDA(ctx, cudaGetLastError());
    rcs = run_scan<CntVal, int64_t, false>(ctx, CntVal{cnt}, n, out->pair_off, bsums, stream);
    if (rcs != HEVCB_OK) { return rcs; }
    HEVCB_CUDA(ctx, cudaMemcpyAsync(pair_total, bsums + nb, 8, cudaMemcpyDeviceToDevice, stream));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(out->pair_off + n, bsums + nb, 8, cudaMemcpyDeviceToDevice, stream));
    if (trace) { parse_kernel<true, false, true><<<g128, 128, 0, stream>>>(a); } else { parse_kernel<true, false, false><<<g128, 128, 0, stream>>>(a); }
    parse_summary_kernel<<<148, 256, 0, stream>>>(cls, out->rc, n, pair_total, out->cap_pairs, d_summary);
    ctx->launches += 2;
    HEVCB_CUDA(ctx, cudaGetLastError());
    if (chain && (chain->sps_out || chain->pps_out)) { // state after the shard's last SPS / PPS (entry 0 = the entering state)
        if (chain->sps_out) { HEVCB_CUDA(ctx, cudaMemcpyAsync(chain->sps_out, a.sps_tab + n_sps, sizeof(hevcb_sps_ctx), cudaMemcpyDeviceToHost, stream)); }
        if (chain->pps_out) { HEVCB_CUDA(ctx, cudaMemcpyAsync(chain->pps_out, a.pps_tab + n_pps, sizeof(hevcb_pps_ctx), cudaMemcpyDeviceToHost, stream)); }
        HEVCB_CUDA(ctx, cudaStreamSynchronize(stream));
    }
    return HEVCB_OK;
}

// ---- the device bit reader / writer on their own: a script of bs_* calls executed by hevcb_bits / hevcb_bitwriter ------------------
namespace {
__global__ void bs_read_kernel(const uint8_t* __restrict__ bytes, int64_t size, const hevcb_bs_op* __restrict__ ops, int n_ops, int32_t* __restrict__ values,
                               int64_t* __restrict__ bitpos, int32_t* __restrict__ overrun)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) { return; }
    hevcb_bits b;
    b.init(bytes, size);
    for (int i = 0; i < n_ops; i++) {
        int32_t v = 0;
        switch (ops[i].kind) {
            case HEVCB_BS_U: v = (int32_t)b.read_u(ops[i].n); break;
            case HEVCB_BS_U1: v = (int32_t)b.read_u1(); break;
            case HEVCB_BS_U8: v = (int32_t)b.read_u8(); break;
            case HEVCB_BS_UE: v = (int32_t)b.read_ue(); break;
            case HEVCB_BS_SE: v = b.read_se(); break;
            case HEVCB_BS_SKIP: b.skip(ops[i].n); break;
            default: break;
        }
        values[i] = v;
        bitpos[i] = b.pos;
        overrun[i] = b.overrun() ? 1 : 0;
    }
}
__global__ void bs_write_kernel(const hevcb_bs_op* __restrict__ ops, int n_ops, uint8_t* __restrict__ out, int64_t cap, int64_t* __restrict__ result)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) { return; }
    hevcb_bitwriter w;
    w.init(out, cap);
    for (int i = 0; i < n_ops; i++) {
        switch (ops[i].kind) {
            case HEVCB_BS_U: w.write_u(ops[i].n, (uint32_t)ops[i].value); break;
            case HEVCB_BS_U1: w.write_u(1, (uint32_t)ops[i].value); break;
            case HEVCB_BS_U8: w.write_u(8, (uint32_t)ops[i].value); break;
            case HEVCB_BS_UE: w.write_ue((uint32_t)ops[i].value); break;
            case HEVCB_BS_SE: w.write_se(ops[i].value); break;
            default: break;
        }
    }
    // a pending partial byte is flushed the way the reference leaves it in memory: its bits at the top of the byte, the rest 0
    if (w.nacc > 0 && (w.pos >> 3) < w.room) { out[w.pos >> 3] = (uint8_t)(w.acc << (8 - w.nacc)); }
    result[0] = w.pos;
    result[1] = w.overrun() ? 1 : 0;
}
} // namespace

extern "C" HEVCB_API int hevcb_bs_read_host(hevcb_ctx* ctx, const uint8_t* bytes, int64_t size, const hevcb_bs_op* ops, int n_ops, int32_t* values,
                                            int64_t* bitpos, int32_t* overrun)
{
    if (!ctx || size < 0 || n_ops < 0 || (size > 0 && !bytes) || (n_ops > 0 && (!ops || !values || !bitpos || !overrun))) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n_ops == 0) { return HEVCB_OK; }
    cudaStream_t st = ctx->stream;
    const size_t o_ops = ((size_t)size + 64 + 255) & ~(size_t)255, o_val = o_ops + (((size_t)n_ops * sizeof(hevcb_bs_op) + 255) & ~(size_t)255);
    const size_t o_pos = o_val + (((size_t)n_ops * 4 + 255) & ~(size_t)255), o_ovr = o_pos + (((size_t)n_ops * 8 + 255) & ~(size_t)255);
    const size_t total = o_ovr + (size_t)n_ops * 4 + 256;
    int rc = hevcb_reserve(ctx, &ctx->wstruct, total);
    if (rc != HEVCB_OK) { return rc; }
    uint8_t* base = reinterpret_cast<uint8_t*>(ctx->wstruct.p);
    HEVCB_CUDA(ctx, cudaMemsetAsync(base, 0, o_ops, st)); // the reader's aligned 8-byte loads reach past the last byte
    if (size > 0) { HEVCB_CUDA(ctx, cudaMemcpyAsync(base, bytes, (size_t)size, cudaMemcpyHostToDevice, st)); }
    HEVCB_CUDA(ctx, cudaMemcpyAsync(base + o_ops, ops, (size_t)n_ops * sizeof(hevcb_bs_op), cudaMemcpyHostToDevice, st));
    bs_read_kernel<<<1, 32, 0, st>>>(base, size, reinterpret_cast<const hevcb_bs_op*>(base + o_ops), n_ops, reinterpret_cast<int32_t*>(base + o_val),
                                     reinterpret_cast<int64_t*>(base + o_pos), reinterpret_cast<int32_t*>(base + o_ovr));
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    HEVCB_CUDA(ctx, cudaMemcpyAsync(values, base + o_val, (size_t)n_ops * 4, cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(bitpos, base + o_pos, (size_t)n_ops * 8, cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaMemcpyAsync(overrun, base + o_ovr, (size_t)n_ops * 4, cudaMemcpyDeviceToHost, st));
    HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    return HEVCB_OK;
}

extern "C" HEVCB_API int hevcb_bs_write_host(hevcb_ctx* ctx, const hevcb_bs_op* ops, int n_ops, uint8_t* out, int64_t cap, int64_t* bits_written, int32_t* overrun)
{
    if (!ctx || n_ops < 0 || cap < 0 || (n_ops > 0 && !ops) || (cap > 0 && !out) || !bits_written || !overrun) { return HEVCB_E_ARG; }
    HEVCB_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t o_ops = ((size_t)cap + 64 + 255) & ~(size_t)255, o_res = o_ops + (((size_t)n_ops * sizeof(hevcb_bs_op) + 255) & ~(size_t)255);
    int rc = hevcb_reserve(ctx, &ctx->wstruct, o_res + 256);
    if (rc != HEVCB_OK) { return rc; }
    uint8_t* base = reinterpret_cast<uint8_t*>(ctx->wstruct.p);
    HEVCB_CUDA(ctx, cudaMemsetAsync(base, 0, o_ops, st));
    if (n_ops > 0) { HEVCB_CUDA(ctx, cudaMemcpyAsync(base + o_ops, ops, (size_t)n_ops * sizeof(hevcb_bs_op), cudaMemcpyHostToDevice, st)); }
    bs_write_kernel<<<1, 32, 0, st>>>(reinterpret_cast<const hevcb_bs_op*>(base + o_ops), n_ops, base, cap, reinterpret_cast<int64_t*>(base + o_res));
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    int64_t res[2] = {0, 0};
    HEVCB_CUDA(ctx, cudaMemcpyAsync(res, base + o_res, 16, cudaMemcpyDeviceToHost, st));
    if (cap > 0) { HEVCB_CUDA(ctx, cudaMemcpyAsync(out, base, (size_t)cap, cudaMemcpyDeviceToHost, st)); }
    HEVCB_CUDA(ctx, cudaStreamSynchronize(st));
    *bits_written = res[0];
    *overrun = (int32_t)res[1];
    return HEVCB_OK;
}

extern "C" HEVCB_API int hevcb_ps_context_bytes(int64_t* sps_bytes, int64_t* pps_bytes)
{
    if (sps_bytes) { *sps_bytes = (int64_t)sizeof(hevcb_sps_ctx); }
    if (pps_bytes) { *pps_bytes = (int64_t)sizeof(hevcb_pps_ctx); }
    return HEVCB_OK;
}
