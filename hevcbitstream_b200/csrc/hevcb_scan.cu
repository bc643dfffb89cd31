// hevcb_scan.cu -- fused Annex-B start-code scan + emulation-prevention strip for sm_100a.
//
// Replaces the reference's per-NAL loop  find_nal_unit (h264_nal.c:38-76) -> nal_to_rbsp
// (h264_nal.c:147-200, called at hevc_stream.c:165)  by one pass over the byte range (a whole stream, or one shard of a
// byte-range partition, see ScanGeom):
//
//   * a persistent cooperative grid (2 CTAs / SM) is split into ANALYSER and WRITER CTAs plus one SCANNER warp; the
//     dependency analyser -> scanner -> writer is one-directional (see the comment in front of the kernel);
//   * 32 KiB tiles (+128 B leading / 16 B trailing halo) are staged into shared memory with TMA bulk copies
//     (cp.async.bulk + mbarrier), two stages per CTA; a tile is loaded by one analyser and, a few microseconds later,
//     by one writer -- out of L2, the analysers stay within a window of the writers' progress;
//   * every lane owns 16 bytes: an exact "two adjacent zero bytes?" SWAR test sends the common case down a fast path;
//     the chunks that fail it are ranked in stream order and their exact predicate bit masks (hevcb_chunk_analyze) are
//     built 32 at a time, every lane busy;
//   * counts (start codes, kept bytes) go through redux, the ordered (last-event-kind, error) carry through warp ballots,
//     lane -> row -> warp -> tile; across tiles one warp scans the 16-byte tile aggregates in stream order;
//   * the EPB-free image is written as aligned 16-byte vectors, funnel-shifted by the tile-uniform misalignment; NAL
//     offsets are written by the lanes that own the events (hevcb_scan_emit_kernel for tiles handed over as event
//     records, the writer itself for tiles with removed bytes or very many events); tiles with removed bytes are
//     compacted in shared memory first, so the image never sees byte-granular stores.
//
// HBM traffic: input read once (second load from L2), image written once, 32 B of metadata per NAL.  Tensor cores unused:
// nothing here is a contraction.
//
// HEVCB_SCAN_DEBUG (context creation) is a bit mask of measurement switches used to attribute time to the parts of the
// kernel (results are wrong with any of them set): 1 no scanner / fake prefixes, 4 no image write, 32 writers off,
// 64 analysis off (bits 8..15: rows to flag), 128 writers do not wait for the scanner, 512 every tile takes the clean
// path, 1024 no event records, 2048 / 65536 the row-by-row analysis / writer paths of interior tiles, bits 12..15 rows up to
// which the analyser stays row-wise (+1).  HEVCB_SCAN_ANALYSERS / HEVCB_SCAN_WINDOW override the role split and the L2 window.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "hevcb_internal.h"
#include "hevcb_scan_core.h"

namespace {

#ifndef HEVCB_SCAN_WORKERS
#define HEVCB_SCAN_WORKERS 8
#endif
#ifndef HEVCB_SCAN_CTAS
#define HEVCB_SCAN_CTAS 2
#endif
constexpr int kWorkers = HEVCB_SCAN_WORKERS;   // warps that analyse / write tiles
constexpr int kSyncThreads = (kWorkers + 1) * 32; // workers + the control warp (publishes aggregates, fetches prefixes, issues TMA)
constexpr int kThreads = (kWorkers + 2) * 32;     // + the scanner warp (only active in CTA 0)
constexpr int kWorkerThreads = kWorkers * 32;
constexpr int kRowBytes = 512;  // one warp-row: 32 lanes x 16 B
#ifndef HEVCB_SCAN_ROWS
#define HEVCB_SCAN_ROWS 8
#endif
constexpr int kRowsPerWarp = HEVCB_SCAN_ROWS; // rows a warp walks in order (its carries stay in registers); 4 or 8
constexpr int kRows = kWorkers * kRowsPerWarp;
constexpr int kTileBytes = kRows * kRowBytes; // 32 KiB
constexpr int kLead = 128;                    // leading halo: a whole 128-byte line so that every bulk copy starts line-aligned
constexpr int kStageBytes = kLead + kTileBytes + 16;
#ifndef HEVCB_SCAN_STAGES
#define HEVCB_SCAN_STAGES 2
#endif
constexpr int kStages = HEVCB_SCAN_STAGES;       // tiles per CTA in shared memory (one being worked on, the others in flight)
#ifndef HEVCB_SCAN_AUNROLL
#define HEVCB_SCAN_AUNROLL 1
#endif
constexpr int kAnalyserUnroll = HEVCB_SCAN_AUNROLL; // the analyser's loop over groups of four rows stays rolled (two roles share the instruction cache)
#ifndef HEVCB_SCAN_PERLANE
#define HEVCB_SCAN_PERLANE 10
#endif
constexpr int kScanPerLane = HEVCB_SCAN_PERLANE; // tile aggregates per lane and batch of the scanner warp (320 tiles per batch)

// byte range handled by one launch (see hevcb_chunk_analyze): a whole stream or one shard of a byte-range partition
struct ScanGeom {
    int64_t size;       // bytes present (owned + following halo)
    int64_t own;        // owned bytes: tiles, image and reported positions stop here
    int64_t evl;        // events / error positions honoured below this
    uint32_t init_n;    // 1: a NAL is considered open at position 0 (local index 0: the piece of a NAL begun in an earlier shard)
    uint32_t init_kind; // carry entering the range
    long long window;   // analysers load at most this many tiles ahead of the writers' progress
};

struct WarpAgg {
    uint32_t n;     // start codes in the warp's rows
    uint32_t k;     // kept bytes
    uint32_t kind;  // ordered carry summary of the warp's rows
    uint32_t err;
    uint32_t del;   // some row has removed bytes or is partially valid
    uint32_t rows;  // bit i: row i of the warp needs exact treatment
    uint32_t pad[2];
};

// what a writer CTA needs to know about a tile before it touches it
struct TilePrefix {
    unsigned long long n;    // start codes before the tile
    unsigned long long k;    // kept bytes before the tile
    unsigned long long mask; // bit r: row r of the tile needs exact treatment (two adjacent zero bytes nearby, or a stream edge)
    uint32_t kind;           // ordered carry entering the tile
    uint32_t err;
};

// dynamic shared memory layout
struct __align__(16) SmemLayout {
    uint8_t stage[kStages][kStageBytes];
    unsigned long long mbar[kStages];
    unsigned long long done[kStages]; // analyser: the workers are through with the stage (one arrival per warp)
    uint32_t evcount[kStages];        // analyser: event records written for the tile in the stage
    WarpAgg wagg[kStages][kWorkers];  // per-warp aggregates of the tile in a stage (writer: index i & 1)
    TilePrefix pref[2];        // writer: prefix + row mask of this CTA's i-th tile (index i & 1)
    uint8_t slowmap[kWorkers][kRowsPerWarp * 32]; // the warp's chunks that need exact analysis, in stream order
    alignas(16) uint16_t delmask[kWorkers][kRowsPerWarp * 32]; // writer: removed bytes of every chunk of the warp's rows (tiles with removed bytes)
};

// ---- tile state for the decoupled look-back: one 16-byte word, read/written with single 128-bit accesses
constexpr unsigned long long kStatusAgg = 1ull, kStatusPrefix = 2ull;
__device__ __forceinline__ ulonglong2 pack_state(unsigned long long status, unsigned long long n, unsigned long long k,
                                                 uint32_t kind, uint32_t err)
{
    ulonglong2 s;
    s.x = (status << 62) | ((unsigned long long)kind << 60) | ((unsigned long long)(err & 1u) << 59) | (n & ((1ull << 40) - 1));
    s.y = k;
    return s;
}
// aggregate of one tile: x = status | kind | err | start codes (14 bits) | kept bytes (16 bits), y = row mask
constexpr uint32_t kEvCap = 128;      // event records an analyser may leave per tile (four per lane of the emit pass)
constexpr uint32_t kEvByWriter = 0xFFu; // "records" value: the writer CTA emits this tile's NAL boundaries itself
__device__ __forceinline__ ulonglong2 pack_agg(uint32_t n, uint32_t k, uint32_t kind, uint32_t err, unsigned long long mask, uint32_t records = kEvByWriter)
{
    ulonglong2 s;
    s.x = (kStatusAgg << 62) | ((unsigned long long)kind << 60) | ((unsigned long long)(err & 1u) << 59) | ((unsigned long long)(records & 0xFFu) << 40) |
          ((unsigned long long)(k & 0xFFFFFu) << 20) |
          (unsigned long long)(n & 0xFFFFFu);
    s.y = mask;
    return s;
}
__device__ __forceinline__ uint32_t agg_n(const ulonglong2& s) { return (uint32_t)(s.x & 0xFFFFFull); }
__device__ __forceinline__ uint32_t agg_k(const ulonglong2& s) { return (uint32_t)((s.x >> 20) & 0xFFFFFull); }
__device__ __forceinline__ uint32_t agg_kind(const ulonglong2& s) { return (uint32_t)(s.x >> 60) & 3u; }
__device__ __forceinline__ uint32_t agg_err(const ulonglong2& s) { return (uint32_t)(s.x >> 59) & 1u; }
__device__ __forceinline__ uint32_t agg_records(const ulonglong2& s) { return (uint32_t)(s.x >> 40) & 0xFFu; }
__device__ __forceinline__ ulonglong2 ld_state(const ulonglong2* p)
{
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(ulonglong2* p, ulonglong2 v)
{
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
}

// ---- TMA / mbarrier helpers (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// named barriers: 0 is __syncthreads; kBarWork = the worker warps only; kBarE / kBarP = workers + control warp (writer role)
constexpr int kBarWork = 1, kBarE = 3, kBarP = 4;
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// issue the bulk copies for tile `t` into stage buffer `st` (one elected thread)
__device__ __forceinline__ void issue_tile_load(uint8_t* st, unsigned long long* bar, const uint8_t* buf, int64_t size, long long t)
{
    const int64_t t0 = (int64_t)t * kTileBytes;
    const int64_t size16 = (size + 15) & ~(int64_t)15; // reads stay inside the 16-byte block of the last byte
    int64_t lo = t0 - kLead;
    uint32_t dst_off = 0;
    if (lo < 0) { lo = 0; dst_off = kLead; }
    int64_t hi = t0 + kTileBytes + 16;
    if (hi > size16) { hi = size16; }
    uint32_t bytes = (uint32_t)(hi - lo);
    mbar_expect_tx(bar, bytes);
    tma_bulk_g2s(st + dst_off, buf + lo, bytes, bar);
}

struct DevSink {
    int64_t* ns;
    int64_t* ne;
    int64_t* ro;
    int64_t* re;
    int64_t cap;
    long long* first_empty;
    __device__ __forceinline__ void open(int64_t k, int64_t start, int64_t off)
    {
        if (k < cap) { ns[k] = start; ro[k] = off; }
    }
    __device__ __forceinline__ void close(int64_t k, int64_t end, int64_t rend, bool empty)
    {
        if (k < cap) { ne[k] = end; re[k] = rend; }
        if (empty) { atomicMin(first_empty, (long long)k); }
    }
};

// scratch header (device): [1] first zero-length NAL index
struct ScanHeader {
    unsigned long long tiles_written; // writer progress (analysers stay within a window of it so that a tile's second load hits L2)
    long long first_empty;
    ulonglong2 final_state; // inclusive prefix over all tiles, written by the scanner warp
    unsigned long long heavy_tiles;  // tiles the writers had to analyse themselves (feeds the role split of the next launch)
    unsigned long long flagged_rows; // rows the analysers had to analyse exactly
    unsigned long long pad[2];
};

// ordered-carry resolution inside a warp: lane l receives the (kind, err) state produced by lanes < l.
// has/sc3/er describe each lane's own segment summary.  Returns kind PASS when no lower lane has an event.
__device__ __forceinline__ void warp_carry_in(uint32_t Eb, uint32_t Sb, uint32_t Rb, int lane, uint32_t& kind, uint32_t& err)
{
    const uint32_t below = (1u << lane) - 1u;
    const uint32_t lower = Eb & below;
    if (lower) {
        const int p = 31 - __clz((int)lower);
        kind = ((Sb >> p) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
        err = (Rb & below & ~((1u << p) - 1u)) != 0u;
    } else {
        kind = HEVCB_KIND_PASS;
        err = (Rb & below) != 0u;
    }
}
// summary of the whole warp's segments
__device__ __forceinline__ void warp_carry_total(uint32_t Eb, uint32_t Sb, uint32_t Rb, uint32_t& kind, uint32_t& err)
{
    if (Eb) {
        const int p = 31 - __clz((int)Eb);
        kind = ((Sb >> p) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
        err = (Rb >> p) != 0u;
    } else {
        kind = HEVCB_KIND_PASS;
        err = Rb != 0u;
    }
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) { v += o; }
    }
    return v;
}

// exact "does [g0-2, g0+17] contain two adjacent zero bytes" test: byte j of (w | w>>8) is zero iff b[j] and
// b[j+1] are both zero, and (m - 0x01..) & ~m & 0x80.. is non-zero iff some byte of m is zero.
__device__ __forceinline__ uint32_t zero_pair_any(uint32_t wp, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wn)
{
    const uint32_t mp = (wp | __funnelshift_r(wp, w0, 8)) | 0x0000FFFFu; // only pairs starting at g0-2, g0-1
    const uint32_t m0 = w0 | __funnelshift_r(w0, w1, 8);
    const uint32_t m1 = w1 | __funnelshift_r(w1, w2, 8);
    const uint32_t m2 = w2 | __funnelshift_r(w2, w3, 8);
    const uint32_t m3 = w3 | __funnelshift_r(w3, wn, 8);
    const uint32_t c = 0x01010101u, h = 0x80808080u;
    return (((mp - c) & ~mp) | ((m0 - c) & ~m0) | ((m1 - c) & ~m1) | ((m2 - c) & ~m2) | ((m3 - c) & ~m3)) & h;
}

// cold path, kept out of line so that the hot loop stays small in the instruction cache
__device__ __noinline__ uint3 analyze_cold(uint32_t wp, uint4 v, uint32_t wn, int64_t g0, int64_t size, int64_t own, int64_t evl)
{
    const hevcb_chunk_masks m = hevcb_chunk_analyze(wp, v.x, v.y, v.z, v.w, wn, g0, size, own, evl);
    return make_uint3(m.ev | (m.sc << 16), m.del | (m.err << 16), m.valid | (m.scb << 16));
}

__device__ __noinline__ uint3 analyze_interior_cold(uint32_t wp, uint4 v, uint32_t wn)
{
    const hevcb_chunk_masks m = hevcb_chunk_analyze_interior(wp, v.x, v.y, v.z, v.w, wn);
    return make_uint3(m.ev | (m.sc << 16), m.del | (m.err << 16), m.valid | (m.scb << 16));
}

__device__ __noinline__ void emit_cold(uint32_t evsc, uint32_t deler, uint32_t misc, int64_t g0, int64_t nbase, int64_t kbase, uint32_t ck,
                                       uint32_t ce, DevSink sink)
{
    hevcb_chunk_masks m;
    m.ev = evsc & 0xFFFFu;
    m.sc = evsc >> 16;
    m.del = deler & 0xFFFFu;
    m.err = deler >> 16;
    m.valid = misc & 0xFFFFu;
    m.scb = misc >> 16;
    hevcb_chunk_emit(m, g0, nbase, kbase, ck, ce, sink);
}

// copy `nv` 16-byte vectors from shared memory (16-byte aligned `src16`, byte offset Q*4 + sh/8 into it) to the
// 16-byte aligned global destination; Q selects the word offset at compile time, sh is the byte shift in bits.
template <int Q>
__device__ __forceinline__ void copy_vectors(uint8_t* __restrict__ dst16, const uint8_t* __restrict__ src16, uint32_t nv, uint32_t sh, int tid,
                                             uint32_t nthreads = kWorkerThreads)
{
    for (uint32_t vi = tid; vi < nv; vi += nthreads) {
        const uint4 lo = *reinterpret_cast<const uint4*>(src16 + (vi << 4));
        const uint4 hi = *reinterpret_cast<const uint4*>(src16 + (vi << 4) + 16);
        const uint32_t W[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        uint4 o4;
        o4.x = __funnelshift_r(W[Q], W[Q + 1], sh);
        o4.y = __funnelshift_r(W[Q + 1], W[Q + 2], sh);
        o4.z = __funnelshift_r(W[Q + 2], W[Q + 3], sh);
        o4.w = __funnelshift_r(W[Q + 3], W[Q + 4], sh);
        __stcs(reinterpret_cast<uint4*>(dst16 + (vi << 4)), o4);
    }
}

#ifndef HEVCB_SPIN_PAUSE
#define HEVCB_SPIN_PAUSE __nanosleep(32)
#endif

// one warp copies a clean 512-byte row (16-byte aligned in shared memory) to an arbitrarily aligned global address
__device__ __forceinline__ void copy_row_clean(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int lane)
{
    const uint32_t head = (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
    if ((uint32_t)lane < head) { dst[lane] = src[lane]; }
    const uint32_t nv = (kRowBytes - head) >> 4; // 31 or 32 vectors
    if ((uint32_t)lane < nv) {
        const uint32_t so = head + ((uint32_t)lane << 4);
        uint4 o4;
        if (head == 0u) {
            o4 = *reinterpret_cast<const uint4*>(src + so);
        } else {
            const uint32_t a = so & ~15u, q = (so & 15u) >> 2, sh = (so & 3u) * 8u;
            const uint4 lo = *reinterpret_cast<const uint4*>(src + a);
            const uint4 hi = *reinterpret_cast<const uint4*>(src + a + 16); // at most 16 bytes past the row: still inside the stage
            const uint32_t W[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            uint32_t x[5];
#pragma unroll
            for (int e = 0; e < 5; e++) { x[e] = (q == 0u) ? W[e] : (q == 1u) ? W[e + 1] : (q == 2u) ? W[e + 2] : W[(e + 3) & 7]; }
            o4.x = __funnelshift_r(x[0], x[1], sh);
            o4.y = __funnelshift_r(x[1], x[2], sh);
            o4.z = __funnelshift_r(x[2], x[3], sh);
            o4.w = __funnelshift_r(x[3], x[4], sh);
        }
        __stcs(reinterpret_cast<uint4*>(dst + so), o4);
    }
    const uint32_t done = head + (nv << 4);
    if ((uint32_t)lane < kRowBytes - done) { dst[done + lane] = src[done + lane]; }
}

// Scanner warp (one per grid): the chained scan over the tile aggregates.  Batch by batch (320 tiles, 10 per lane, in
// stream order) it waits for the aggregates the analyser CTAs publish, combines them with shuffles / ballots and
// publishes every tile's EXCLUSIVE prefix (start codes, kept bytes, ordered carry) into tile_excl[].  One reader per
// aggregate and one 16-byte poll per tile replace an all-to-all look-back, whose polling traffic on a few cache lines
// was measured to cost ~14k cycles per wave of 296 tiles.
__device__ __forceinline__ void scanner_warp(const ulonglong2* __restrict__ tile_state, ulonglong2* __restrict__ tile_excl,
                                             long long n_tiles, ScanHeader* __restrict__ hdr, int lane, uint32_t init_n, uint32_t init_kind)
{
    unsigned long long runN = init_n, runK = 0;
    uint32_t runKind = init_kind, runErr = 0;
    for (long long base = 0; base < n_tiles; base += 32 * kScanPerLane) {
        const long long first = base + (long long)lane * kScanPerLane;
        ulonglong2 sv[kScanPerLane];
#pragma unroll
        for (int j = 0; j < kScanPerLane; j++) {
            const long long idx = first + j;
            if (idx < n_tiles) { sv[j] = ld_state(&tile_state[idx]); }
            else { sv[j] = pack_agg(0, 0, HEVCB_KIND_PASS, 0, 0ull); } // past the end: identity
        }
        for (;;) { // re-poll, one batch per round trip, the aggregates that are not published yet
            bool missing = false;
#pragma unroll
            for (int j = 0; j < kScanPerLane; j++) { missing = missing || ((sv[j].x >> 62) == 0ull); }
            if (!__any_sync(0xFFFFFFFFu, missing)) { break; }
#pragma unroll
            for (int j = 0; j < kScanPerLane; j++) {
                if ((sv[j].x >> 62) == 0ull) { sv[j] = ld_state(&tile_state[first + j]); }
            }
        }
        __threadfence(); // the aggregates observed above happen before the prefixes published below (writers re-read them)
        // lane totals
        uint32_t ln = 0, lk = 0, lkind = HEVCB_KIND_PASS, lerr = 0;
#pragma unroll
        for (int j = 0; j < kScanPerLane; j++) {
            ln += agg_n(sv[j]);
            lk += agg_k(sv[j]);
            hevcb_carry_combine(lkind, lerr, agg_kind(sv[j]), agg_err(sv[j]));
        }
        // exclusive scan over lanes (ascending lane = stream order)
        const uint32_t nin = warp_incl_scan(ln, lane), kin = warp_incl_scan(lk, lane);
        const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, lkind != HEVCB_KIND_PASS);
        const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lkind == HEVCB_KIND_SC3);
        const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, lerr != 0u);
        uint32_t ck, ce;
        warp_carry_in(Eb, Sb, Rb, lane, ck, ce);
        uint32_t cKind = runKind, cErr = runErr; // carry entering this lane's first tile
        hevcb_carry_combine(cKind, cErr, ck, ce);
        unsigned long long cN = runN + (nin - ln), cK = runK + (kin - lk);
#pragma unroll
        for (int j = 0; j < kScanPerLane; j++) {
            const long long idx = first + j;
            if (idx < n_tiles) { st_state(&tile_excl[idx], pack_state(kStatusPrefix, cN, cK, cKind, cErr)); }
            cN += agg_n(sv[j]);
            cK += agg_k(sv[j]);
            hevcb_carry_combine(cKind, cErr, agg_kind(sv[j]), agg_err(sv[j]));
        }
        // running state after the batch = lane 31's state after its last tile
        runN = __shfl_sync(0xFFFFFFFFu, cN, 31);
        runK = __shfl_sync(0xFFFFFFFFu, cK, 31);
        runKind = __shfl_sync(0xFFFFFFFFu, cKind, 31);
        runErr = __shfl_sync(0xFFFFFFFFu, cErr, 31);
    }
    if (lane == 0) { hdr->final_state = pack_state(kStatusPrefix, runN, runK, runKind, runErr); }
}

// per-row analysis shared by both roles: exact masks of one 512-byte row (lanes without two adjacent zero bytes nearby take
// the default), the warp's ballots, and the row's contribution to the warp aggregate
struct RowMasks {
    uint32_t evsc, deler, misc;
    uint32_t Xb, Db; // some lane has an event / error position; some lane removes bytes or is partially owned
};
__device__ __forceinline__ uint3 default_masks(int64_t g0, const ScanGeom& geom)
{
    const int64_t rem = geom.own - g0;
    return make_uint3(0u, 0u, rem >= 16 ? 0xFFFFu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u)));
}
__device__ __forceinline__ RowMasks summarize_row(const uint3 m3, uint32_t& wN, uint32_t& wK, uint32_t& wKind, uint32_t& wErr);
__device__ __forceinline__ RowMasks analyze_row(uint32_t wp, const uint4 v, uint32_t wn, bool slow, int64_t g0, const ScanGeom& geom,
                                                uint32_t& wN, uint32_t& wK, uint32_t& wKind, uint32_t& wErr)
{
    uint3 m3 = default_masks(g0, geom);
    if (slow) { m3 = analyze_cold(wp, v, wn, g0, geom.size, geom.own, geom.evl); }
    return summarize_row(m3, wN, wK, wKind, wErr);
}
__device__ __forceinline__ RowMasks summarize_row(const uint3 m3, uint32_t& wN, uint32_t& wK, uint32_t& wKind, uint32_t& wErr)
{
    const uint32_t ev = m3.x & 0xFFFFu, sc = m3.x >> 16, del = m3.y & 0xFFFFu, er = m3.y >> 16, valid = m3.z & 0xFFFFu;
    uint32_t lk = HEVCB_KIND_PASS, le = (er != 0u); // lane summary for the ordered carry
    if (ev != 0u) {
        const int tp = 31 - __clz((int)ev);
        lk = ((sc >> tp) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
        le = ((er >> tp) >> 1) != 0u;
    }
    const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, ev != 0u);
    const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lk == HEVCB_KIND_SC3);
    const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, le != 0u);
    RowMasks r;
    r.evsc = m3.x; r.deler = m3.y; r.misc = m3.z;
    r.Xb = __ballot_sync(0xFFFFFFFFu, (ev | er) != 0u);
    r.Db = __ballot_sync(0xFFFFFFFFu, (del != 0u) || (valid != 0xFFFFu));
    wN += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(sc));
    wK += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(valid & ~del));
    uint32_t rk, re;
    warp_carry_total(Eb, Sb, Rb, rk, re);
    hevcb_carry_combine(wKind, wErr, rk, re);
    return r;
}

// writer: the same for row r of a staged tile, operands read from shared memory
__device__ __forceinline__ RowMasks analyze_staged_row(const uint8_t* st, int r, int lane, int64_t t0, const ScanGeom& geom, uint32_t& wN,
                                                       uint32_t& wK, uint32_t& wKind, uint32_t& wErr)
{
    const uint8_t* rp = st + kLead + r * kRowBytes + lane * 16;
    const uint4 v = *reinterpret_cast<const uint4*>(rp);
    const uint32_t wp = *reinterpret_cast<const uint32_t*>(rp - 4);
    const uint32_t wn = *reinterpret_cast<const uint32_t*>(rp + 16);
    const bool slow = zero_pair_any(wp, v.x, v.y, v.z, v.w, wn) != 0u;
    const int64_t g0 = t0 + (int64_t)r * kRowBytes + lane * 16;
    return analyze_row(wp, v, wn, slow, g0, geom, wN, wK, wKind, wErr);
}

// copy of L bytes out of a stage (16-byte aligned source) by a group of `nthreads` threads (the CTA's workers for a whole tile, one
// warp for its own rows): aligned 16-byte vectors, funnel-shifted by the (copy-uniform) misalignment of the destination
__device__ __forceinline__ void copy_span(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t L, int tid, uint32_t nthreads)
{
    const uint32_t head0 = (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
    const uint32_t head = head0 < L ? head0 : L;
    if ((uint32_t)tid < head) { dst[tid] = src[tid]; }
    const uint32_t nv = (L - head) >> 4;
    const uint32_t sh = (head & 3u) * 8u; // source misalignment is uniform
    switch (head >> 2) {
        case 0: copy_vectors<0>(dst + head, src, nv, sh, tid, nthreads); break;
        case 1: copy_vectors<1>(dst + head, src, nv, sh, tid, nthreads); break;
        case 2: copy_vectors<2>(dst + head, src, nv, sh, tid, nthreads); break;
        default: copy_vectors<3>(dst + head, src, nv, sh, tid, nthreads); break;
    }
    const uint32_t done = head + (nv << 4);
    if ((uint32_t)tid < L - done) { dst[done + tid] = src[done + tid]; }
}
__device__ __noinline__ void copy_tile(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t L, int tid)
{
    copy_span(dst, src, L, tid, kWorkerThreads);
}

// boundary fix-ups of a staged tile: positions < 0 read as non-zero, positions >= size read as zero
__device__ __forceinline__ void fix_stage(uint8_t* st, long long t, int64_t t0, int64_t size, int tid)
{
    const int64_t valid_end = (int64_t)kLead + (size - t0); // smem offset of position `size`
    if (t == 0 || valid_end < kStageBytes) {
        if (t == 0 && tid < kLead) { st[tid] = 0xFF; }
        if (valid_end < kStageBytes) {
            for (int i = (int)valid_end + tid; i < kStageBytes; i += kWorkerThreads) { st[i] = 0; }
        }
        bar_sync(kBarWork, kWorkerThreads);
    }
}

// Analyser, interior tile: exact analysis of the warp's "slow" chunks (two adjacent zero bytes nearby) only.  The slow chunks of
// the warp's eight rows are ranked in stream order (ballots), their indices parked in shared memory, and the lanes then take
// them 32 at a time: one pass of the exact masks with every lane busy instead of one pass per flagged row with a few lanes
// busy.  The ordered carry, the counts and the event records follow from ballots over the ranked chunks (the combine is
// associative, chunks without a zero pair contribute nothing).  slow8 bit i: this lane's chunk of row i is slow.
__device__ __noinline__ void analyse_slow_chunks(const uint8_t* __restrict__ st, uint8_t* __restrict__ slowmap, int warp, int lane, uint32_t slow8,
                                                 int64_t t0, long long t, const ScanGeom& geom, uint32_t* evcount, uint4* __restrict__ tile_events,
                                                 uint32_t& wN, uint32_t& wDel, uint32_t& wKind, uint32_t& wErr, uint32_t& rows)
{
    const uint32_t below = (1u << lane) - 1u;
    uint32_t total = 0;
#pragma unroll
    for (int i = 0; i < kRowsPerWarp; i++) {
        const bool mine = ((slow8 >> i) & 1u) != 0u;
        const uint32_t sb = __ballot_sync(0xFFFFFFFFu, mine);
        if (mine) { slowmap[total + (uint32_t)__popc(sb & below)] = (uint8_t)(i * 32 + lane); }
        rows |= (sb != 0u ? 1u : 0u) << i;
        total += (uint32_t)__popc(sb);
    }
    __syncwarp();
    for (uint32_t base = 0; base < total; base += 32u) {
        const uint32_t slot = base + (uint32_t)lane;
        uint32_t evsc = 0u, deler = 0u, misc = 0xFFFFu, chunk = 0u;
        if (slot < total) {
            chunk = (uint32_t)(warp * kRowsPerWarp * 32) + slowmap[slot];
            const uint8_t* rp = st + kLead + chunk * 16u;
            const uint4 v = *reinterpret_cast<const uint4*>(rp);
            const uint32_t wp = *reinterpret_cast<const uint32_t*>(rp - 4);
            const uint32_t wn = *reinterpret_cast<const uint32_t*>(rp + 16);
            const hevcb_chunk_masks m = hevcb_chunk_analyze_interior(wp, v.x, v.y, v.z, v.w, wn); // interior tile: no position limits
            evsc = m.ev | (m.sc << 16); deler = m.del | (m.err << 16); misc = m.valid | (m.scb << 16);
        }
        const uint32_t ev = evsc & 0xFFFFu, sc = evsc >> 16, del = deler & 0xFFFFu, er = deler >> 16;
        uint32_t lk = HEVCB_KIND_PASS, le = (er != 0u);
        if (ev != 0u) {
            const int tp = 31 - __clz((int)ev);
            lk = ((sc >> tp) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
            le = ((er >> tp) >> 1) != 0u;
        }
        const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, ev != 0u);
        const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lk == HEVCB_KIND_SC3);
        const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, le != 0u);
        const uint32_t Xb = __ballot_sync(0xFFFFFFFFu, (ev | er) != 0u);
        wN += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(sc));
        wDel += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(del));
        uint32_t rk, re;
        warp_carry_total(Eb, Sb, Rb, rk, re);
        hevcb_carry_combine(wKind, wErr, rk, re);
        if (Xb != 0u) { // chunks with an event or an error position leave a record (one shared-memory atomic per batch)
            uint32_t first = 0u;
            if (lane == 0) { first = atomicAdd(evcount, (uint32_t)__popc(Xb)); }
            first = __shfl_sync(0xFFFFFFFFu, first, 0);
            const uint32_t rslot = first + (uint32_t)__popc(Xb & below);
            if (((ev | er) != 0u) && rslot < kEvCap) { tile_events[(size_t)t * kEvCap + rslot] = make_uint4(chunk, evsc, deler, misc); }
        }
    }
    __syncwarp(); // the map is rewritten for the warp's next tile
}

// ====================================================================================================================
// The grid is split into two roles.
//   ANALYSER CTAs walk the tiles (round-robin among themselves), build the exact predicates where two adjacent zero bytes
//   occur, and publish one 16-byte aggregate per tile: start codes, kept bytes, ordered carry, and a 64-bit mask of the
//   rows that need exact treatment.  They never wait for anything but their own TMA loads.
//   The SCANNER warp turns aggregates into exclusive prefixes, in stream order.
//   WRITER CTAs walk the same tiles some microseconds later (the second load of a tile is served by L2, so DRAM still sees
//   every input byte once): they wait for the tile's prefix, redo the exact analysis of the flagged rows only, emit the
//   NAL boundaries and write the EPB-free image.
// The dependency analyser -> scanner -> writer is one-directional: no CTA ever waits on a CTA of its own kind, which is
// what removes the wave-by-wave lock-step a single-role pipeline suffers from (its prefix round trip through L2 is as
// long as the work on a tile).
// ====================================================================================================================
__global__ void __launch_bounds__(kThreads, HEVCB_SCAN_CTAS) hevcb_scan_strip_kernel(
    const uint8_t* __restrict__ buf, const ScanGeom geom, long long n_tiles, long long n_analysers, ScanHeader* __restrict__ hdr,
    ulonglong2* __restrict__ tile_state, ulonglong2* __restrict__ tile_excl, uint4* __restrict__ tile_events,
    int64_t* __restrict__ nal_start, int64_t* __restrict__ nal_end, int64_t cap_nals, uint8_t* __restrict__ rbsp,
    int64_t* __restrict__ rbsp_off, int64_t* __restrict__ rbsp_end, long long debug_flags)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const unsigned dbg = (unsigned)debug_flags; // experiment switches, 0 in production
    const int64_t size = geom.size;
    const bool analyser = (long long)blockIdx.x < n_analysers;
    const long long G = analyser ? n_analysers : (long long)gridDim.x - n_analysers; // CTAs of this role
    const long long first_tile = analyser ? (long long)blockIdx.x : (long long)blockIdx.x - n_analysers;

    if (tid == kWorkerThreads) {
        for (int s = 0; s < kStages; s++) { mbar_init(&sm.mbar[s], 1); mbar_init(&sm.done[s], kWorkers); sm.evcount[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        for (int s = 0; s < kStages; s++) { // this CTA's first tiles
            const long long t = first_tile + (long long)s * G;
            if (t < n_tiles) { issue_tile_load(sm.stage[s], &sm.mbar[s], buf, size, t); }
        }
    }
    __syncthreads();

    if (warp == kWorkers + 1) { // scanner warp: one per grid, never joins the CTA's barriers
        if (blockIdx.x == 0 && !(dbg & 1u)) { scanner_warp(tile_state, tile_excl, n_tiles, hdr, lane, geom.init_n, geom.init_kind); }
        return;
    }

    if (!analyser && (dbg & 32u)) { return; } // experiment: writers off
    if (analyser) {
        // ============================================ ANALYSER ============================================
        if (warp == kWorkers) {
            // control warp: all workers are done with the stage -> publish the tile's aggregate, reload the stage
            int s = 0;
            uint32_t done_bits = 0;
            long long seen_written = 0;
            unsigned long long heavy_local = 0, flagged_local = 0;
            for (long long t = first_tile; t < n_tiles; t += G) {
                while (!mbar_try_wait(&sm.done[s], (done_bits >> s) & 1u)) {}
                done_bits ^= (1u << s);
                uint32_t tile_n = 0, tile_k = 0, ak = HEVCB_KIND_PASS, ae = 0, adel = 0;
                unsigned long long mask = 0ull;
#pragma unroll
                for (int w = 0; w < kWorkers; w++) {
                    const WarpAgg a = sm.wagg[s][w];
                    tile_n += a.n;
                    tile_k += a.k;
                    hevcb_carry_combine(ak, ae, a.kind, a.err);
                    mask |= (unsigned long long)(a.rows & 0xFFu) << (w * kRowsPerWarp);
                    adel |= a.del;
                }
                if (lane == 0) {
                    // A tile whose bytes are all kept and whose event chunks fit into the record list looks CLEAN to the
                    // writer (mask 0): its NAL boundaries are emitted from the records by hevcb_scan_emit_kernel.
                    const uint32_t nrec = sm.evcount[s];
                    sm.evcount[s] = 0;
                    const bool light = (adel == 0u) && (nrec <= kEvCap) && (tile_k == (uint32_t)kTileBytes) && !(dbg & 1024u);
                    st_state(&tile_state[t], light ? pack_agg(tile_n, tile_k, ak, ae, 0ull, nrec) : pack_agg(tile_n, tile_k, ak, ae, mask, kEvByWriter));
                    if (!light && mask != 0ull) { heavy_local++; }
                    flagged_local += (unsigned long long)__popcll(mask);
                    const long long nt = t + (long long)kStages * G; // the tile that reuses this stage
                    if (nt < n_tiles) {
                        // stay within `window` tiles of the writers: what they load a second time is then still in L2
                        while (nt > seen_written + geom.window) {
                            unsigned long long w;
                            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(&hdr->tiles_written) : "memory");
                            seen_written = (long long)w;
                            if (nt > seen_written + geom.window) { __nanosleep(200); }
                        }
                        fence_proxy_async();
                        issue_tile_load(sm.stage[s], &sm.mbar[s], buf, size, nt);
                    }
                }
                __syncwarp();
                s = (s + 1 == kStages) ? 0 : s + 1;
            }
            if (lane == 0 && heavy_local) { atomicAdd(&hdr->heavy_tiles, heavy_local); }
            if (lane == 0 && flagged_local) { atomicAdd(&hdr->flagged_rows, flagged_local); }
            return;
        }
        // workers: wait for the tile, analyse their rows, hand the warp aggregate to the control warp; they never wait for
        // it (the stage they move on to was loaded three tiles ago; a stage is only reloaded after its aggregate was read)
        uint32_t phase_bits = 0;
        int s = 0;
        for (long long t = first_tile; t < n_tiles; t += G) {
            const int64_t t0 = (int64_t)t * kTileBytes;
            uint8_t* st = sm.stage[s];
            while (!mbar_try_wait(&sm.mbar[s], (phase_bits >> s) & 1u)) {}
            phase_bits ^= (1u << s);
            fix_stage(st, t, t0, size, tid);
            const bool interior = geom.evl - t0 >= (int64_t)kTileBytes + 32; // every byte (and its halo) owned and below the event limit
            uint32_t wN = 0, wK = 0, wKind = HEVCB_KIND_PASS, wErr = 0, rows = 0, anydel = 0;
            if (dbg & 64u) { wK = kRowsPerWarp * kRowBytes; rows = (dbg >> 8) & 0xFFu; } // experiment: analysis off, rows flagged as told
            else if (interior && !(dbg & 2048u)) {
                // interior tile: SWAR zero-pair test of the warp's eight rows, four rows (independent instruction chains) at a
                // time; one vote sends the common "no two adjacent zero bytes anywhere" case on, otherwise the slow chunks are
                // analysed exactly, ranked in stream order (analyse_slow_chunks)
                uint32_t slow8 = 0;
#pragma unroll kAnalyserUnroll
                for (int i0 = 0; i0 < kRowsPerWarp; i0 += 4) {
                    const uint8_t* rp = st + kLead + (warp * kRowsPerWarp + i0) * kRowBytes + lane * 16;
                    uint4 v[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) { v[k] = *reinterpret_cast<const uint4*>(rp + k * kRowBytes); }
                    uint32_t wp[4], wn[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        wp[k] = *reinterpret_cast<const uint32_t*>(rp + k * kRowBytes - 4);
                        wn[k] = *reinterpret_cast<const uint32_t*>(rp + k * kRowBytes + 16);
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        slow8 |= (zero_pair_any(wp[k], v[k].x, v[k].y, v[k].z, v[k].w, wn[k]) != 0u ? 1u : 0u) << (i0 + k);
                    }
                }
                uint32_t wDel = 0;
                const uint32_t rowany = __reduce_or_sync(0xFFFFFFFFu, slow8); // bit i: row i has a slow chunk
                if (rowany != 0u) {
                    const uint32_t thr = (dbg & 0xF000u) ? ((dbg >> 12) & 15u) - 1u : 1u; // experiment switch; production: 1
                    if ((uint32_t)__popc(rowany) <= thr) {
                        // a single flagged row: row by row (measured: shorter dependency chain than ranking the chunks first)
                        uint32_t dummyK = 0;
                        for (uint32_t rm = rowany; rm != 0u; rm &= rm - 1u) {
                            const int i = __ffs((int)rm) - 1;
                            const RowMasks r = analyze_staged_row(st, warp * kRowsPerWarp + i, lane, t0, geom, wN, dummyK, wKind, wErr);
                            rows |= 1u << i;
                            wDel += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(r.deler & 0xFFFFu));
                            if (((r.evsc & 0xFFFFu) | (r.deler >> 16)) != 0u) {
                                const uint32_t slot = atomicAdd(&sm.evcount[s], 1u);
                                if (slot < kEvCap) {
                                    tile_events[(size_t)t * kEvCap + slot] = make_uint4((uint32_t)((warp * kRowsPerWarp + i) * 32 + lane), r.evsc, r.deler, r.misc);
                                }
                            }
                        }
                    } else {
                        analyse_slow_chunks(st, sm.slowmap[warp], warp, lane, slow8, t0, t, geom, &sm.evcount[s], tile_events, wN, wDel, wKind, wErr, rows);
                    }
                }
                wK = kRowsPerWarp * kRowBytes - wDel;
                anydel = wDel;
            }
            else
            // Rows are taken four at a time: the four loads, halo exchanges and zero-pair tests are independent
            // instruction chains, and one vote sends the common "no two adjacent zero bytes anywhere" case on.
            // (Loops over rows are kept rolled on purpose: two roles share the SM's instruction cache.)
#pragma unroll kAnalyserUnroll
            for (int i0 = 0; i0 < kRowsPerWarp; i0 += 4) {
                const int rbase = warp * kRowsPerWarp + i0;
                const uint8_t* rp = st + kLead + rbase * kRowBytes + lane * 16;
                uint4 v[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { v[k] = *reinterpret_cast<const uint4*>(rp + k * kRowBytes); }
                // the four bytes in front of and behind every chunk straight from shared memory (two narrow loads per chunk
                // instead of shuffles + divergent edge fix-ups for lanes 0 / 31)
                uint32_t wp[4], wn[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    wp[k] = *reinterpret_cast<const uint32_t*>(rp + k * kRowBytes - 4);
                    wn[k] = *reinterpret_cast<const uint32_t*>(rp + k * kRowBytes + 16);
                }
                uint32_t slowmask = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    slowmask |= (zero_pair_any(wp[k], v[k].x, v[k].y, v[k].z, v[k].w, wn[k]) != 0u ? 1u : 0u) << k;
                }
                if (!__any_sync(0xFFFFFFFFu, slowmask != 0u) && interior) { // four rows on the fast path: 2 KiB kept
                    wK += 4 * kRowBytes;
                    continue;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const bool slow = ((slowmask >> k) & 1u) != 0u;
                    if (!__any_sync(0xFFFFFFFFu, slow) && interior) { wK += kRowBytes; continue; }
                    const int64_t g0 = t0 + (int64_t)(rbase + k) * kRowBytes + lane * 16;
                    const RowMasks r = analyze_row(wp[k], v[k], wn[k], slow, g0, geom, wN, wK, wKind, wErr);
                    rows |= 1u << (i0 + k);
                    anydel |= (r.Db != 0u) ? 1u : 0u;
                    if (((r.evsc & 0xFFFFu) | (r.deler >> 16)) != 0u) { // chunk with an event or an error position: leave a record
                        const uint32_t slot = atomicAdd(&sm.evcount[s], 1u);
                        if (slot < kEvCap) {
                            tile_events[(size_t)t * kEvCap + slot] = make_uint4((uint32_t)((rbase + k) * 32 + lane), r.evsc, r.deler, r.misc);
                        }
                    }
                }
            }
            if (lane == 0) {
                WarpAgg a;
                a.n = wN; a.k = wK; a.kind = wKind; a.err = wErr; a.del = anydel; a.rows = rows;
                a.pad[0] = a.pad[1] = 0;
                sm.wagg[s][warp] = a;
                mbar_arrive(&sm.done[s]);
            }
            __syncwarp();
            s = (s + 1 == kStages) ? 0 : s + 1;
        }
        return;
    }

    // ================================================ WRITER ================================================
    if (warp == kWorkers) {
        // control warp: lane j polls the prefix and the aggregate (row mask) of this CTA's tile number q + j, kAhead tiles
        // ahead of the workers, so that the two L2 round trips never sit between two tiles.  Per tile: wait until lane 0 has
        // both words -> shared memory -> [E] the workers are done with the previous stage -> reload it -> [P] release the
        // workers into the tile -> every lane hands its words down by one lane.
        constexpr int kAhead = 4;
        if (first_tile >= n_tiles) { return; }
        long long tq = first_tile + (long long)lane * G; // tile polled by this lane
        ulonglong2 ex = make_ulonglong2(0ull, 0ull), ag = make_ulonglong2(0ull, 0ull);
        int s = 0;
        for (long long t = first_tile, it = 0; t < n_tiles; t += G, it++) {
            for (;;) {
                if (lane < kAhead && tq < n_tiles) {
                    if (dbg & 1u) {
                        ex = pack_state(kStatusPrefix, 0, (unsigned long long)tq * kTileBytes, HEVCB_KIND_Z3, 0);
                        ag = pack_agg(0, 0, 0, 0, ~0ull);
                    } else {
                        if (dbg & 128u) { ex = pack_state(kStatusPrefix, 0, (unsigned long long)tq * kTileBytes, HEVCB_KIND_Z3, 0); } // experiment: scanner off the path
                        else if ((ex.x >> 62) == 0ull) { ex = ld_state(&tile_excl[tq]); }
                        if ((ag.x >> 62) == 0ull) { ag = ld_state(&tile_state[tq]); }
                    }
                }
                const bool ready = (ex.x >> 62) != 0ull && (ag.x >> 62) != 0ull;
                if (__shfl_sync(0xFFFFFFFFu, ready ? 1 : 0, 0)) { break; }
                __nanosleep(20);
            }
            if (lane == 0) {
                TilePrefix tp;
                tp.n = ex.x & ((1ull << 40) - 1); tp.k = ex.y; tp.kind = (uint32_t)(ex.x >> 60) & 3u; tp.err = (uint32_t)(ex.x >> 59) & 1u;
                tp.mask = ag.y;
                sm.pref[it & 1] = tp;
            }
            __syncwarp();
            if (it > 0) {
                bar_sync(kBarE, kSyncThreads); // every worker finished reading the stage of the previous tile
                if (lane == 0) { atomicAdd(&hdr->tiles_written, 1ull); }
                const int ps = (s == 0) ? kStages - 1 : s - 1;
                const long long nt = (t - G) + (long long)kStages * G;
                if (lane == 0 && nt < n_tiles) {
                    fence_proxy_async();
                    issue_tile_load(sm.stage[ps], &sm.mbar[ps], buf, size, nt);
                }
            }
            bar_arrive(kBarP, kSyncThreads);
            // hand down: lane j takes over what lane j + 1 has polled so far; the last polling lane starts a new tile
            ex.x = __shfl_down_sync(0xFFFFFFFFu, ex.x, 1); ex.y = __shfl_down_sync(0xFFFFFFFFu, ex.y, 1);
            ag.x = __shfl_down_sync(0xFFFFFFFFu, ag.x, 1); ag.y = __shfl_down_sync(0xFFFFFFFFu, ag.y, 1);
            tq += G;
            if (lane >= kAhead - 1) { ex = make_ulonglong2(0ull, 0ull); ag = make_ulonglong2(0ull, 0ull); }
            s = (s + 1 == kStages) ? 0 : s + 1;
        }
        bar_sync(kBarE, kSyncThreads); // the last tile
        if (lane == 0) { atomicAdd(&hdr->tiles_written, 1ull); }
        return;
    }

    DevSink sink{nal_start, nal_end, rbsp_off, rbsp_end, cap_nals, &hdr->first_empty};
    uint32_t phase_bits = 0;
    int s = 0;
    for (long long t = first_tile, it = 0; t < n_tiles; t += G, it++) {
        const int64_t t0 = (int64_t)t * kTileBytes;
        uint8_t* st = sm.stage[s];
        bar_sync(kBarP, kSyncThreads); // the tile's prefix and row mask are in shared memory
        const TilePrefix pref = sm.pref[it & 1];
        while (!mbar_try_wait(&sm.mbar[s], (phase_bits >> s) & 1u)) {}
        phase_bits ^= (1u << s);
        const long long tileN = (long long)pref.n;
        const long long tileK = (long long)pref.k;

        if (pref.mask == 0ull) {
            // ---- clean interior tile: aligned 16-byte vectors, funnel-shifted by the (tile-uniform) misalignment
            if (rbsp != nullptr && !(dbg & 4u)) { copy_tile(rbsp + tileK, st + kLead, kTileBytes, tid); }
            bar_arrive(kBarE, kSyncThreads);
            s = (s + 1 == kStages) ? 0 : s + 1;
            continue;
        }

        // ---- tile with flagged rows: exact masks of those rows, warp aggregates, ordered emission, row-wise write-out
        fix_stage(st, t, t0, size, tid);
        if (geom.evl - t0 >= (int64_t)kTileBytes + 32 && !(dbg & 65536u)) {
            // Interior tile.  The warp ranks the slow chunks of its eight rows in stream order and takes them 32 at a time
            // (every lane busy, see analyse_slow_chunks); pass 1 parks the masks and builds the warp aggregate, pass 2 emits.
            uint8_t* const map = sm.slowmap[warp];
            const uint32_t below = (1u << lane) - 1u;
            uint32_t total = 0;
#pragma unroll 1
            for (int i0 = 0; i0 < kRowsPerWarp; i0 += 4) {
                const uint8_t* rp = st + kLead + (warp * kRowsPerWarp + i0) * kRowBytes + lane * 16;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint4 v = *reinterpret_cast<const uint4*>(rp + k * kRowBytes);
                    const uint32_t wp = *reinterpret_cast<const uint32_t*>(rp + k * kRowBytes - 4);
                    const uint32_t wn = *reinterpret_cast<const uint32_t*>(rp + k * kRowBytes + 16);
                    const bool mine = zero_pair_any(wp, v.x, v.y, v.z, v.w, wn) != 0u;
                    const uint32_t sb = __ballot_sync(0xFFFFFFFFu, mine);
                    if (mine) { map[total + (uint32_t)__popc(sb & below)] = (uint8_t)((i0 + k) * 32 + lane); }
                    total += (uint32_t)__popc(sb);
                }
            }
            __syncwarp();
            constexpr int kBatches = kRowsPerWarp;
            uint32_t p_evsc[kBatches], p_deler[kBatches], p_misc[kBatches];
            uint32_t wN = 0, wDel = 0, wKind = HEVCB_KIND_PASS, wErr = 0;
#pragma unroll 1
            for (int b = 0; b < kBatches; b++) {
                const uint32_t slot = (uint32_t)b * 32u + (uint32_t)lane;
                if ((uint32_t)b * 32u >= total) { break; }
                uint32_t evsc = 0u, deler = 0u, misc = 0xFFFFu;
                if (slot < total) {
                    const uint32_t chunk = (uint32_t)(warp * kRowsPerWarp * 32) + map[slot];
                    const uint8_t* rp = st + kLead + chunk * 16u;
                    const uint4 v = *reinterpret_cast<const uint4*>(rp);
                    const uint32_t wp = *reinterpret_cast<const uint32_t*>(rp - 4);
                    const uint32_t wn = *reinterpret_cast<const uint32_t*>(rp + 16);
                    const uint3 m3 = analyze_interior_cold(wp, v, wn); // interior tile: no position limits
                    evsc = m3.x; deler = m3.y; misc = m3.z;
                }
                p_evsc[b] = evsc; p_deler[b] = deler; p_misc[b] = misc;
                const uint32_t ev = evsc & 0xFFFFu, sc = evsc >> 16, del = deler & 0xFFFFu, er = deler >> 16;
                uint32_t lk = HEVCB_KIND_PASS, le = (er != 0u);
                if (ev != 0u) {
                    const int tp = 31 - __clz((int)ev);
                    lk = ((sc >> tp) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
                    le = ((er >> tp) >> 1) != 0u;
                }
                const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, ev != 0u);
                const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lk == HEVCB_KIND_SC3);
                const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, le != 0u);
                wN += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(sc));
                wDel += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(del));
                uint32_t rk, re;
                warp_carry_total(Eb, Sb, Rb, rk, re);
                hevcb_carry_combine(wKind, wErr, rk, re);
            }
            if (lane == 0) {
                WarpAgg a;
                a.n = wN; a.k = kRowsPerWarp * kRowBytes - wDel; a.kind = wKind; a.err = wErr; a.del = wDel; a.rows = 0;
                a.pad[0] = a.pad[1] = 0;
                sm.wagg[it & 1][warp] = a;
            }
            bar_sync(kBarWork, kWorkerThreads);
            uint32_t rN = 0, rK = 0, rKind = HEVCB_KIND_PASS, rErr = 0, tileDel = 0;
            {
                uint32_t tn = 0, tk = 0, ak = HEVCB_KIND_PASS, ae = 0;
#pragma unroll
                for (int w = 0; w < kWorkers; w++) {
                    const WarpAgg a = sm.wagg[it & 1][w];
                    if (w == warp) { rN = tn; rK = tk; rKind = ak; rErr = ae; }
                    tn += a.n;
                    tk += a.k;
                    hevcb_carry_combine(ak, ae, a.kind, a.err);
                    tileDel += a.del;
                }
            }
            const bool write_img = (rbsp != nullptr) && !(dbg & 4u);
            const bool compacting = write_img && tileDel != 0u && wDel != 0u; // this warp's rows lose bytes
            uint16_t* const dm = sm.delmask[warp];
            if (compacting) {
                for (int e = lane * 8; e < kRowsPerWarp * 32; e += 256) { *reinterpret_cast<uint4*>(dm + e) = make_uint4(0u, 0u, 0u, 0u); }
                __syncwarp();
            }
            // pass 2: ordered emission over the ranked chunks
            {
                uint32_t cKind = pref.kind, cErr = pref.err; // carry entering this warp = tile carry (+) warps before it
                hevcb_carry_combine(cKind, cErr, rKind, rErr);
                uint32_t nrun = rN, drun = (uint32_t)(warp * kRowsPerWarp * kRowBytes) - rK; // start codes / removed bytes before, within the tile
#pragma unroll 1
                for (int b = 0; b < kBatches; b++) {
                    const uint32_t slot = (uint32_t)b * 32u + (uint32_t)lane;
                    if ((uint32_t)b * 32u >= total) { break; }
                    const uint32_t evsc = p_evsc[b], deler = p_deler[b], misc = p_misc[b];
                    const uint32_t ev = evsc & 0xFFFFu, sc = evsc >> 16, del = deler & 0xFFFFu, er = deler >> 16;
                    const uint32_t local = (slot < total) ? (uint32_t)map[slot] : 0u;
                    if (!__any_sync(0xFFFFFFFFu, (ev | er) != 0u)) {
                        // nothing to emit and no carry change in this batch (EPB-dense payload): only the removed bytes count
                        if (tileDel != 0u) {
                            if (compacting && del != 0u) { dm[local] = (uint16_t)del; }
                            drun += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(del));
                        }
                        continue;
                    }
                    uint32_t lk = HEVCB_KIND_PASS, le = (er != 0u);
                    if (ev != 0u) {
                        const int tp = 31 - __clz((int)ev);
                        lk = ((sc >> tp) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
                        le = ((er >> tp) >> 1) != 0u;
                    }
                    const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, ev != 0u);
                    const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lk == HEVCB_KIND_SC3);
                    const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, le != 0u);
                    uint32_t ck, ce;
                    warp_carry_in(Eb, Sb, Rb, lane, ck, ce);
                    if (ck == HEVCB_KIND_PASS) { ck = cKind; ce |= cErr; } // inherit the carry entering the batch
                    const uint32_t c = (uint32_t)__popc(sc);
                    const uint32_t ninc = warp_incl_scan(c, lane);
                    uint32_t dinc = 0, d = 0;
                    if (tileDel != 0u) { // (uniform) tiles that keep every byte need no removed-byte ranks
                        d = (uint32_t)__popc(del);
                        dinc = warp_incl_scan(d, lane);
                        if (compacting && del != 0u) { dm[local] = (uint16_t)del; }
                    }
                    if ((ev | er) != 0u) {
                        const uint32_t chunk = (uint32_t)(warp * kRowsPerWarp * 32) + local;
                        emit_cold(evsc, deler, misc, t0 + (int64_t)chunk * 16, (int64_t)(tileN + nrun + (ninc - c)),
                                  (int64_t)(tileK + (long long)chunk * 16 - (long long)(drun + dinc - d)), ck, ce, sink);
                    }
                    uint32_t rk, re;
                    warp_carry_total(Eb, Sb, Rb, rk, re);
                    hevcb_carry_combine(cKind, cErr, rk, re);
                    nrun += __shfl_sync(0xFFFFFFFFu, ninc, 31);
                    if (tileDel != 0u) { drun += __shfl_sync(0xFFFFFFFFu, dinc, 31); }
                }
            }
            if (write_img) {
                if (tileDel == 0u) {
                    copy_tile(rbsp + tileK, st + kLead, kTileBytes, tid); // every byte kept: one shifted vector copy by all workers
                } else {
                    // tile with removed bytes: every warp closes the gaps of its own rows in place (shared memory), then copies
                    // its kept bytes out as one span
                    uint8_t* const wb = st + kLead + warp * (kRowsPerWarp * kRowBytes);
                    uint32_t out = kRowsPerWarp * kRowBytes;
                    if (compacting) {
                        __syncwarp();
                        out = 0;
#pragma unroll 1
                        for (int i = 0; i < kRowsPerWarp; i++) {
                            const uint32_t del = dm[i * 32 + lane];
                            if (out == (uint32_t)(i * kRowBytes) && !__any_sync(0xFFFFFFFFu, del != 0u)) { out += kRowBytes; continue; } // still in place
                            const uint4 v = *reinterpret_cast<const uint4*>(wb + i * kRowBytes + lane * 16);
                            const uint32_t keep = 0xFFFFu & ~del;
                            const uint32_t c = (uint32_t)__popc(keep);
                            const uint32_t inc = warp_incl_scan(c, lane);
                            __syncwarp(); // every lane holds its chunk before bytes move
                            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                            uint8_t* o = wb + out + (inc - c);
#pragma unroll
                            for (int j = 0; j < 16; j++) {
                                if ((keep >> j) & 1u) { *o++ = (uint8_t)(w[j >> 2] >> (8 * (j & 3))); }
                            }
                            __syncwarp();
                            out += __shfl_sync(0xFFFFFFFFu, inc, 31);
                        }
                    }
                    copy_span(rbsp + tileK + rK, wb, out, lane, 32u);
                }
            }
            bar_arrive(kBarE, kSyncThreads); // the control warp may now reload this stage; workers do not wait
            s = (s + 1 == kStages) ? 0 : s + 1;
            continue;
        }
        const uint32_t myrows = (uint32_t)(pref.mask >> (warp * kRowsPerWarp)) & 0xFFu;
        // pass 1: aggregates of this warp's flagged rows; their masks are parked (thread-local memory) for pass 2
        uint32_t wN = 0, wK = 0, wKind = HEVCB_KIND_PASS, wErr = 0, anyD = 0;
        uint32_t k_evsc[kRowsPerWarp], k_deler[kRowsPerWarp], k_misc[kRowsPerWarp], k_flags = 0;
#pragma unroll 1
        for (int i = 0; i < kRowsPerWarp; i++) {
            if (!((myrows >> i) & 1u)) { wK += kRowBytes; continue; }
            const RowMasks m = analyze_staged_row(st, warp * kRowsPerWarp + i, lane, t0, geom, wN, wK, wKind, wErr);
            anyD |= (m.Db != 0u) ? 1u : 0u;
            k_evsc[i] = m.evsc; k_deler[i] = m.deler; k_misc[i] = m.misc;
            k_flags |= ((m.Xb != 0u) ? 1u : 0u) << i;
            k_flags |= ((m.Db != 0u) ? 1u : 0u) << (8 + i);
        }
        if (lane == 0) {
            WarpAgg a;
            a.n = wN; a.k = wK; a.kind = wKind; a.err = wErr; a.del = anyD; a.rows = myrows;
            a.pad[0] = a.pad[1] = 0;
            sm.wagg[it & 1][warp] = a;
        }
        bar_sync(kBarWork, kWorkerThreads);
        // aggregate of the warps before this one, tile totals
        uint32_t rN = 0, rK = 0, rKind = HEVCB_KIND_PASS, rErr = 0, compact = 0, tile_k = 0;
        {
            uint32_t tn = 0, ak = HEVCB_KIND_PASS, ae = 0;
#pragma unroll
            for (int w = 0; w < kWorkers; w++) {
                const WarpAgg a = sm.wagg[it & 1][w];
                if (w == warp) { rN = tn; rK = tile_k; rKind = ak; rErr = ae; }
                tn += a.n;
                tile_k += a.k;
                hevcb_carry_combine(ak, ae, a.kind, a.err);
                compact |= a.del;
            }
        }
        const bool write_rows = (rbsp != nullptr) && !(dbg & 4u);
        const bool dirty_out = (compact != 0u) && write_rows;
        // pass 2: ordered emission of NAL boundaries (the masks of the few flagged rows are rebuilt instead of being kept in
        // registers across the barrier); tiles with removed bytes are also written out here, row by row
        {
            uint32_t cKind = pref.kind, cErr = pref.err; // carry entering this warp = tile carry (+) warps before it
            hevcb_carry_combine(cKind, cErr, rKind, rErr);
#pragma unroll 1
            for (int i = 0; i < kRowsPerWarp; i++) {
                const int r = warp * kRowsPerWarp + i;
                if (!((myrows >> i) & 1u)) {
                    if (dirty_out) { copy_row_clean(rbsp + tileK + rK, st + kLead + r * kRowBytes, lane); }
                    rK += kRowBytes;
                    continue;
                }
                const bool rowX = ((k_flags >> i) & 1u) != 0u, rowD = ((k_flags >> (8 + i)) & 1u) != 0u;
                if (!rowD) {
                    if (dirty_out) { copy_row_clean(rbsp + tileK + rK, st + kLead + r * kRowBytes, lane); }
                    if (!rowX) { rK += kRowBytes; continue; } // nothing to emit
                }
                const uint32_t evsc = k_evsc[i], deler = k_deler[i], misc = k_misc[i];
                const uint32_t ev = evsc & 0xFFFFu, sc = evsc >> 16, del = deler & 0xFFFFu, er = deler >> 16, valid = misc & 0xFFFFu;
                const uint32_t keep = valid & ~del;
                uint32_t klane, rowKept;
                if (rowD) {
                    const uint32_t c = (uint32_t)__popc(keep);
                    const uint32_t inc = warp_incl_scan(c, lane);
                    klane = inc - c;
                    rowKept = __shfl_sync(0xFFFFFFFFu, inc, 31);
                } else {
                    klane = (uint32_t)lane * 16u;
                    rowKept = kRowBytes;
                }
                klane += rK;
                if (rowX) {
                    uint32_t lk = HEVCB_KIND_PASS, le = (er != 0u);
                    if (ev != 0u) {
                        const int tp = 31 - __clz((int)ev);
                        lk = ((sc >> tp) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
                        le = ((er >> tp) >> 1) != 0u;
                    }
                    const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, ev != 0u);
                    const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lk == HEVCB_KIND_SC3);
                    const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, le != 0u);
                    uint32_t ck, ce;
                    warp_carry_in(Eb, Sb, Rb, lane, ck, ce);
                    if (ck == HEVCB_KIND_PASS) { ck = cKind; ce |= cErr; } // inherit the carry entering the row
                    const uint32_t c = (uint32_t)__popc(sc);
                    const uint32_t ninc = warp_incl_scan(c, lane);
                    if ((ev | er) != 0u) {
                        const int64_t g0 = t0 + (int64_t)r * kRowBytes + lane * 16;
                        emit_cold(evsc, deler, misc, g0, (int64_t)(tileN + rN + (ninc - c)), (int64_t)(tileK + klane), ck, ce, sink);
                    }
                    uint32_t rk, re;
                    warp_carry_total(Eb, Sb, Rb, rk, re);
                    hevcb_carry_combine(cKind, cErr, rk, re);
                    rN += __shfl_sync(0xFFFFFFFFu, ninc, 31);
                }
                if (rowD && dirty_out) { // row with removed / out-of-range bytes: byte-granular stores of the kept bytes
                    const uint4 v = *reinterpret_cast<const uint4*>(st + kLead + r * kRowBytes + lane * 16);
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                    uint8_t* o = rbsp + tileK + klane;
#pragma unroll
                    for (int j = 0; j < 16; j++) {
                        if ((keep >> j) & 1u) { *o++ = (uint8_t)(w[j >> 2] >> (8 * (j & 3))); }
                    }
                }
                rK += rowKept;
            }
        }
        // ---- tile without removed bytes: one shifted vector copy of the whole tile by all workers
        if (write_rows && !dirty_out) { copy_tile(rbsp + tileK, st + kLead, tile_k, tid); }
        bar_arrive(kBarE, kSyncThreads); // the control warp may now reload this stage; workers do not wait
        s = (s + 1 == kStages) ? 0 : s + 1;
    }
}

// Emit pass: NAL boundaries of the tiles whose event chunks the analysers left as records (one warp per tile, one lane per
// record).  Records are in arrival order: they are ranked by chunk index, then the lanes run the same ordered-carry logic
// a row of the writer runs.  Tiles of that kind keep all their bytes, so a chunk's image offset is prefix + chunk * 16.
__global__ void __launch_bounds__(256) hevcb_scan_emit_kernel(long long n_tiles, ScanHeader* __restrict__ hdr, const ulonglong2* __restrict__ tile_state,
                                                              const ulonglong2* __restrict__ tile_excl, const uint4* __restrict__ tile_events,
                                                              int64_t* __restrict__ nal_start, int64_t* __restrict__ nal_end, int64_t cap_nals,
                                                              int64_t* __restrict__ rbsp_off, int64_t* __restrict__ rbsp_end)
{
    __shared__ uint4 sorted[8][kEvCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long t = (long long)blockIdx.x * 8 + warp;
    if (t >= n_tiles) { return; }
    const ulonglong2 ag = tile_state[t];
    const uint32_t nrec = agg_records(ag);
    if (nrec == 0u || nrec == kEvByWriter) { return; }
    const ulonglong2 ex = tile_excl[t];
    const long long tileN = (long long)(ex.x & ((1ull << 40) - 1)), tileK = (long long)ex.y;
    const uint32_t pKind = (uint32_t)(ex.x >> 60) & 3u, pErr = (uint32_t)(ex.x >> 59) & 1u;
    // rank by chunk index (distinct per record); lane l holds records l, l + 32, l + 64, ...
    constexpr int kPerLane = (int)(kEvCap / 32);
    uint4 rec_[kPerLane];
    uint32_t rank_[kPerLane];
#pragma unroll
    for (int q = 0; q < kPerLane; q++) {
        rec_[q] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0xFFFFu);
        rank_[q] = 0;
        if ((uint32_t)(q * 32 + lane) < nrec) { rec_[q] = tile_events[(size_t)t * kEvCap + q * 32 + lane]; }
    }
#pragma unroll
    for (int p = 0; p < kPerLane; p++) {
        if ((uint32_t)(p * 32) < nrec) { // warp-uniform: slots beyond the record count hold nothing
#pragma unroll 4
            for (int i = 0; i < 32; i++) {
                const uint32_t other = __shfl_sync(0xFFFFFFFFu, rec_[p].x, i);
#pragma unroll
                for (int q = 0; q < kPerLane; q++) { rank_[q] += (other < rec_[q].x) ? 1u : 0u; }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < kPerLane; q++) {
        if ((uint32_t)(q * 32 + lane) < nrec) { sorted[warp][rank_[q]] = rec_[q]; }
    }
    __syncwarp();
    uint32_t cKind = pKind, cErr = pErr; // carry entering the round
    long long nbase = tileN;
    DevSink sink{nal_start, nal_end, rbsp_off, rbsp_end, cap_nals, &hdr->first_empty};
    for (uint32_t base = 0; base < nrec; base += 32) {
        uint4 rec = make_uint4(0u, 0u, 0u, 0xFFFFu);
        if (base + (uint32_t)lane < nrec) { rec = sorted[warp][base + lane]; }
        const uint32_t evsc = rec.y, deler = rec.z, misc = rec.w;
        const uint32_t ev = evsc & 0xFFFFu, sc = evsc >> 16, er = deler >> 16;
        uint32_t lk = HEVCB_KIND_PASS, le = (er != 0u);
        if (ev != 0u) {
            const int tp = 31 - __clz((int)ev);
            lk = ((sc >> tp) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
            le = ((er >> tp) >> 1) != 0u;
        }
        const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, ev != 0u);
        const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lk == HEVCB_KIND_SC3);
        const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, le != 0u);
        uint32_t ck, ce;
        warp_carry_in(Eb, Sb, Rb, lane, ck, ce);
        if (ck == HEVCB_KIND_PASS) { ck = cKind; ce |= cErr; } // inherit the carry entering the round
        const uint32_t c = (uint32_t)__popc(sc);
        const uint32_t ninc = warp_incl_scan(c, lane);
        if ((ev | er) != 0u) {
            const int64_t g0 = (int64_t)t * kTileBytes + (int64_t)rec.x * 16;
            emit_cold(evsc, deler, misc, g0, (int64_t)(nbase + (ninc - c)), (int64_t)(tileK + (long long)rec.x * 16), ck, ce, sink);
        }
        uint32_t rk, re;
        warp_carry_total(Eb, Sb, Rb, rk, re);
        hevcb_carry_combine(cKind, cErr, rk, re);
        nbase += __shfl_sync(0xFFFFFFFFu, ninc, 31);
    }
}

// single-thread epilogue: reference end-of-buffer rules over the last 8 bytes + summary
__global__ void hevcb_scan_finalize_kernel(const uint8_t* __restrict__ buf, int64_t size, long long n_tiles,
                                           const ScanHeader* __restrict__ hdr, const ulonglong2* __restrict__ tile_state,
                                           int64_t* nal_start, int64_t* nal_end, int64_t* rbsp_off, int64_t* rbsp_end,
                                           int64_t cap_nals, hevcb_scan_summary* summary)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) { return; }
    int64_t N = 0, K = 0;
    uint32_t kind = HEVCB_KIND_Z3, err = 0;
    if (n_tiles > 0) {
        const ulonglong2 sv = hdr->final_state;
        N = (int64_t)(sv.x & ((1ull << 40) - 1));
        K = (int64_t)sv.y;
        kind = (uint32_t)(sv.x >> 60) & 3u;
        err = (uint32_t)(sv.x >> 59) & 1u;
    }
    long long fe = hdr->first_empty;
    if (fe == 0x7FFFFFFFFFFFFFFFll) { fe = -1; }
    auto fetch = [buf](int64_t pos) -> uint32_t { return (uint32_t)buf[pos]; };
    hevcb_scan_summary_core core;
    hevcb_scan_finalize(size, N, kind, err, K, (int64_t)fe, fetch, nal_start, nal_end, rbsp_off, rbsp_end, cap_nals, &core);
    summary->n_nals = core.n_nals;
    summary->n_terminated = core.n_terminated;
    summary->last_rc = core.last_rc;
    summary->overflow = core.overflow;
    summary->last_start = core.last_start;
    summary->last_end = core.last_end;
    summary->rbsp_bytes = core.rbsp_bytes;
    summary->n_epb = core.n_epb;
}

__global__ void hevcb_scan_init_kernel(ScanHeader* hdr, uint32_t init_n, int64_t* nal_start, int64_t* rbsp_off, int64_t cap_nals)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        hdr->tiles_written = 0ull;
        hdr->heavy_tiles = 0ull;
        hdr->flagged_rows = 0ull;
        hdr->first_empty = 0x7FFFFFFFFFFFFFFFll;
        if (init_n && cap_nals > 0) { nal_start[0] = 0; rbsp_off[0] = 0; } // the NAL piece that enters the shard
    }
}

// epilogue of a shard pass: no end-of-buffer rules here (hevcb_stitch applies them once, to the last shard)
__global__ void hevcb_scan_shard_finalize_kernel(const uint8_t* __restrict__ buf, const ScanGeom geom, long long n_tiles,
                                                 const ScanHeader* __restrict__ hdr, int64_t* nal_start, int64_t* nal_end, int64_t* rbsp_off,
                                                 int64_t* rbsp_end, int64_t cap_nals, int is_first, int is_last, hevcb_shard_summary* out)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) { return; }
    int64_t N = geom.init_n, K = 0;
    uint32_t kind = geom.init_kind, err = 0;
    if (n_tiles > 0) {
        const ulonglong2 sv = hdr->final_state;
        N = (int64_t)(sv.x & ((1ull << 40) - 1));
        K = (int64_t)sv.y;
        kind = (uint32_t)(sv.x >> 60) & 3u;
        err = (uint32_t)(sv.x >> 59) & 1u;
    }
    long long fe = hdr->first_empty;
    if (fe == 0x7FFFFFFFFFFFFFFFll) { fe = -1; }
    auto fetch = [buf](int64_t pos) -> uint32_t { return (uint32_t)buf[pos]; };
    hevcb_shard_finalize(geom.own, N, kind, err, K, (int64_t)fe, fetch, nal_start, nal_end, rbsp_off, rbsp_end, cap_nals, is_first, is_last, out);
}

} // namespace


static int launch_scan_common(hevcb_ctx* ctx, const uint8_t* d_buf, const ScanGeom& geom, int64_t* d_nal_start, int64_t* d_nal_end,
                              int64_t cap_nals, uint8_t* d_rbsp, int64_t* d_rbsp_off, int64_t* d_rbsp_end, ScanHeader** hdr_out,
                              long long* n_tiles_out, cudaStream_t stream)
{
    if (((uintptr_t)d_buf & 15u) || ((uintptr_t)d_rbsp & 15u)) {
        HEVCB_SET_ERR(ctx, "hevcb_scan_strip: buf and rbsp must be 16-byte aligned");
        return HEVCB_E_ALIGN;
    }
    const long long n_tiles = (long long)((geom.own + kTileBytes - 1) / kTileBytes);
    const size_t n_states = (size_t)(n_tiles > 0 ? n_tiles : 1);
    const size_t need_states = sizeof(ScanHeader) + 2 * n_states * sizeof(ulonglong2);
    const size_t need = need_states + n_states * kEvCap * sizeof(uint4); // + the event records of the tiles (not cleared)
    int rc = hevcb_reserve(ctx, &ctx->scan_scratch, need);
    if (rc != HEVCB_OK) { return rc; }
    ScanHeader* hdr = reinterpret_cast<ScanHeader*>(ctx->scan_scratch.p);
    ulonglong2* states = reinterpret_cast<ulonglong2*>(reinterpret_cast<uint8_t*>(ctx->scan_scratch.p) + sizeof(ScanHeader));
    ulonglong2* excl = states + n_states;
    uint4* events = reinterpret_cast<uint4*>(excl + n_states);

    HEVCB_CUDA(ctx, cudaMemsetAsync(ctx->scan_scratch.p, 0, need_states, stream));
    hevcb_scan_init_kernel<<<1, 32, 0, stream>>>(hdr, geom.init_n, d_nal_start, d_rbsp_off, cap_nals);
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());

    if (n_tiles > 0) {
        const size_t smem = sizeof(SmemLayout);
        if (ctx->scan_blocks_per_sm == 0) {
            HEVCB_CUDA(ctx, cudaFuncSetAttribute(hevcb_scan_strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int nb = 0;
            HEVCB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, hevcb_scan_strip_kernel, kThreads, smem));
            if (nb < 1) { HEVCB_SET_ERR(ctx, "scan kernel does not fit on an SM"); return HEVCB_E_CUDA; }
            ctx->scan_blocks_per_sm = nb;
        }
        long long grid = (long long)ctx->sm_count * ctx->scan_blocks_per_sm;
        if (grid > 2 * n_tiles) { grid = 2 * n_tiles; }
        if (grid < 2) { grid = 2; }
        // analyser CTAs (the rest are writers): 60 % is the measured optimum when the writers mostly copy (NALs >= 4 KiB); when
        // the previous launch on this context found that the writers had to analyse most tiles themselves (tiny NALs,
        // EPB-dense payloads) they get the larger share.  Results do not depend on the split.
        if (ctx->ev_stats && ctx->stats_pending && cudaEventQuery(ctx->ev_stats) == cudaSuccess) {
            const unsigned long long* st = reinterpret_cast<const unsigned long long*>(ctx->pinned) + 64;
            ctx->last_heavy_frac = st[2] ? (double)st[0] / (double)st[2] : 0.0;
            ctx->last_flagged_frac = st[2] ? (double)st[1] / ((double)st[2] * kRows) : 0.0;
            ctx->stats_pending = false;
        }
        // 3/8 analysers when the writers analyse most tiles themselves (tiny NALs, EPB-dense payloads); 65 % when more than a fifth
        // of the rows needed exact analysis (NALs <= 2 KiB: the analysers are the bottleneck); 60 % otherwise (measured optima, 4 GiB)
        const double ff = ctx->last_flagged_frac;
        long long n_an = ctx->last_heavy_frac > 0.5 ? (grid * 3 + 4) / 8 : (ff > 0.2 ? (grid * 13 + 10) / 20 : (grid * 3 + 2) / 5);
        if (n_an < 1) { n_an = 1; }
        if (n_an > grid - 1) { n_an = grid - 1; }
        if (const char* e = getenv("HEVCB_SCAN_ANALYSERS")) { const long long v = atoll(e); if (v >= 1 && v < grid) { n_an = v; } }
        // cooperative launch: writers wait on the scanner, the scanner on the analysers: every CTA must be resident
        long long nt = n_tiles;
        long long dbg = ctx->scan_debug_flags;
        ScanGeom g = geom;
        g.window = 1100; // tiles (34 MiB): more than both roles keep in flight (2 stages x grid) plus a scanner batch (measured, bench.py on one box: 1200 tiles 2259 GB/s, 1100 2241, below 1050 the analysers are throttled)
        if (const char* e = getenv("HEVCB_SCAN_WINDOW")) { const long long v = atoll(e); if (v >= kStages * grid) { g.window = v; } }
        if (g.window < kStages * grid) { g.window = kStages * grid; }
        if (dbg & 32u) { g.window = 1ll << 40; } // experiment "writers off": nothing to wait for
        void* args[] = {(void*)&d_buf, (void*)&g, (void*)&nt, (void*)&n_an, (void*)&hdr, (void*)&states, (void*)&excl, (void*)&events, (void*)&d_nal_start, (void*)&d_nal_end,
                        (void*)&cap_nals, (void*)&d_rbsp, (void*)&d_rbsp_off, (void*)&d_rbsp_end, (void*)&dbg};
        HEVCB_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)hevcb_scan_strip_kernel, dim3((unsigned)grid), dim3(kThreads), args, smem, stream));
        hevcb_scan_emit_kernel<<<(unsigned)((n_tiles + 7) / 8), 256, 0, stream>>>(n_tiles, hdr, states, excl, events, d_nal_start, d_nal_end, cap_nals,
                                                                                 d_rbsp_off, d_rbsp_end);
        ctx->launches += 2;
        HEVCB_CUDA(ctx, cudaGetLastError());
        if (ctx->ev_stats && !ctx->stats_pending) { // heavy-tile count of this launch, read back without ever waiting for it
            unsigned long long* st = reinterpret_cast<unsigned long long*>(ctx->pinned) + 64;
            st[2] = (unsigned long long)n_tiles;
            HEVCB_CUDA(ctx, cudaMemcpyAsync(&st[0], &hdr->heavy_tiles, 16, cudaMemcpyDeviceToHost, stream)); // heavy_tiles, flagged_rows
            HEVCB_CUDA(ctx, cudaEventRecord(ctx->ev_stats, stream));
            ctx->stats_pending = true;
        }
    }
    *hdr_out = hdr;
    *n_tiles_out = n_tiles;
    return HEVCB_OK;
}

int hevcb_launch_scan_strip(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, int64_t* d_nal_start, int64_t* d_nal_end,
                            int64_t cap_nals, uint8_t* d_rbsp, int64_t* d_rbsp_off, int64_t* d_rbsp_end,
                            hevcb_scan_summary* d_summary, cudaStream_t stream)
{
    if (size < 0 || cap_nals < 0 || !d_nal_start || !d_nal_end || !d_rbsp_off || !d_rbsp_end || !d_summary || (size > 0 && !d_buf)) {
        HEVCB_SET_ERR(ctx, "hevcb_scan_strip: invalid argument");
        return HEVCB_E_ARG;
    }
    ScanGeom geom;
    geom.size = size; geom.own = size; geom.evl = size - HEVCB_TAIL_ZONE; geom.init_n = 0; geom.init_kind = HEVCB_KIND_Z3; geom.window = 0;
    ScanHeader* hdr = nullptr;
    long long n_tiles = 0;
    int rc = launch_scan_common(ctx, d_buf, geom, d_nal_start, d_nal_end, cap_nals, d_rbsp, d_rbsp_off, d_rbsp_end, &hdr, &n_tiles, stream);
    if (rc != HEVCB_OK) { return rc; }
    hevcb_scan_finalize_kernel<<<1, 32, 0, stream>>>(d_buf, size, n_tiles, hdr, nullptr, d_nal_start, d_nal_end, d_rbsp_off, d_rbsp_end,
                                                     cap_nals, d_summary);
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}

int hevcb_launch_scan_strip_shard(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t own, int64_t halo, int is_first, int is_last,
                                  int64_t* d_nal_start, int64_t* d_nal_end, int64_t cap_nals, uint8_t* d_rbsp, int64_t* d_rbsp_off,
                                  int64_t* d_rbsp_end, hevcb_shard_summary* d_summary, cudaStream_t stream)
{
    if (own < 0 || halo < 0 || halo > 16 || cap_nals < 1 || !d_nal_start || !d_nal_end || !d_rbsp_off || !d_rbsp_end || !d_summary ||
        (own > 0 && !d_buf) || (is_last && halo != 0) || (!is_last && own > 0 && halo < 3)) {
        HEVCB_SET_ERR(ctx, "hevcb_scan_strip_shard: invalid argument (inner shards need a halo of 3..16 bytes, the last shard none)");
        return HEVCB_E_ARG;
    }
    ScanGeom geom;
    geom.size = own + halo;
    geom.own = own;
    geom.evl = is_last ? own - HEVCB_TAIL_ZONE : own;
    geom.init_n = is_first ? 0u : 1u;
    geom.init_kind = is_first ? HEVCB_KIND_Z3 : HEVCB_KIND_SC3;
    geom.window = 0;
    ScanHeader* hdr = nullptr;
    long long n_tiles = 0;
    int rc = launch_scan_common(ctx, d_buf, geom, d_nal_start, d_nal_end, cap_nals, d_rbsp, d_rbsp_off, d_rbsp_end, &hdr, &n_tiles, stream);
    if (rc != HEVCB_OK) { return rc; }
    hevcb_scan_shard_finalize_kernel<<<1, 32, 0, stream>>>(d_buf, geom, n_tiles, hdr, d_nal_start, d_nal_end, d_rbsp_off, d_rbsp_end, cap_nals,
                                                           is_first, is_last, d_summary);
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}
