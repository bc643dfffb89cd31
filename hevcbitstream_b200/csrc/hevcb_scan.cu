// hevcb_scan.cu -- fused Annex-B start-code scan + emulation-prevention strip for sm_100a.
//
// Replaces the reference's per-NAL loop  find_nal_unit (h264_nal.c:38-76) -> nal_to_rbsp
// (h264_nal.c:147-200, called at hevc_stream.c:165)  by one pass over the byte range (a whole stream, or one shard of a
// byte-range partition, see ScanGeom):
//
//   * a persistent cooperative grid, one CTA per SM, every CTA a warp-specialised pipeline over a six-stage shared-memory
//     ring: PRODUCER (TMA bulk copies) -> ANALYSER warps -> aggregate published -> [SCANNER warp: prefixes, across CTAs] ->
//     WRITER warps -> stage free.  A tile is loaded ONCE and stays in shared memory until its prefix has arrived; nothing
//     inside a CTA waits in lock-step (see the comment in front of the kernel);
//   * 32 KiB tiles (+128 B leading / 16 B trailing halo) are staged with cp.async.bulk + mbarrier;
//   * every lane owns 16 bytes: an exact "two adjacent zero bytes?" SWAR test sends the common case down a fast path;
//     the chunks that fail it are ranked in stream order and their exact predicate bit masks (hevcb_chunk_analyze) are
//     built 32 at a time, every lane busy;
//   * counts (start codes, kept bytes) go through redux, the ordered (last-event-kind, error) carry through warp ballots,
//     lane -> row -> warp -> tile; across tiles one warp scans the 16-byte tile aggregates in stream order;
//   * the EPB-free image is written as aligned 16-byte vectors, funnel-shifted by the tile-uniform misalignment; NAL
//     offsets are written by the lanes that own the events (hevcb_scan_emit_kernel for tiles handed over as event
//     records, the writer warps for tiles with removed bytes or very many events); tiles with removed bytes are
//     compacted in shared memory first, so the image never sees byte-granular stores.
//
// HBM traffic: input read once, image written once, 32 B of metadata per NAL.  Tensor cores unused: nothing here is a
// contraction.
//
// HEVCB_SCAN_DEBUG (context creation) is a bit mask of measurement switches used to attribute time to the parts of the
// kernel (results are wrong with any of them set): 4 no image write, 64 analysis off (bits 8..15: rows to flag),
// 1024 no event records, 2048 / 65536 the row-by-row analysis / writer paths of interior tiles, bits 12..15 rows up to
// which the analyser stays row-wise (+1).
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
#define HEVCB_SCAN_CTAS 1
#endif
// warp roles inside a CTA: kAWarps analysers (kARows rows of a tile each), kWWarps writers (kWRows rows each), then one warp each
// for: analyser control (publishes tile aggregates), writer control (fetches tile prefixes), producer (TMA).  The last CTA of
// the grid is the scanner.  Analysis is the longer job per row (and a tile's aggregate waits for its slowest warp), so it gets
// twice the warps.
constexpr int kAWarps = 16, kARows = 4;
constexpr int kWWarps = 8, kWRows = 8;
constexpr int kAThreads = kAWarps * 32, kWThreads = kWWarps * 32;
constexpr int kWarpAC = kAWarps + kWWarps, kWarpWC = kWarpAC + 1, kWarpProd = kWarpAC + 2;
constexpr int kThreads = (kAWarps + kWWarps + 3) * 32;
constexpr int kRowBytes = 512;  // one warp-row: 32 lanes x 16 B
constexpr int kRows = kAWarps * kARows;
static_assert(kRows == kWWarps * kWRows && kRows == 64, "a tile is 64 rows: its row mask is one 64-bit word");
constexpr int kTileBytes = kRows * kRowBytes; // 32 KiB
constexpr int kLead = 128;                    // leading halo: a whole 128-byte line so that every bulk copy starts line-aligned
constexpr int kStageBytes = kLead + kTileBytes + 16;
#ifndef HEVCB_SCAN_STAGES
#define HEVCB_SCAN_STAGES 6
#endif
constexpr int kStages = HEVCB_SCAN_STAGES;       // ring of tiles per CTA: in flight / being analysed / waiting for their prefix / being written
#ifndef HEVCB_SCAN_REFINE
#define HEVCB_SCAN_REFINE 1 // analysers narrow the filter's superset of slow chunks to the exact set when it exceeds one batch
#endif
#ifndef HEVCB_SCAN_AUNROLL
#define HEVCB_SCAN_AUNROLL 1
#endif
constexpr int kAnalyserUnroll = HEVCB_SCAN_AUNROLL; // the analyser's loop over groups of four rows stays rolled (two roles share the instruction cache)
#ifndef HEVCB_SCAN_PERLANE
#define HEVCB_SCAN_PERLANE 4
#endif
constexpr int kScanPerLane = HEVCB_SCAN_PERLANE; // most tile aggregates per lane and batch of a scanner warp
#ifndef HEVCB_SCAN_BATCH_PER_LANE
#define HEVCB_SCAN_BATCH_PER_LANE 2
#endif

// byte range handled by one launch (see hevcb_chunk_analyze): a whole stream or one shard of a byte-range partition
struct ScanGeom {
    int64_t size;       // bytes present (owned + following halo)
    int64_t own;        // owned bytes: tiles, image and reported positions stop here
    int64_t evl;        // events / error positions honoured below this
    uint32_t init_n;    // 1: a NAL is considered open at position 0 (local index 0: the piece of a NAL begun in an earlier shard)
    uint32_t init_kind; // carry entering the range
    int scan_per_lane;  // scanner batch = 32 x this many tiles; a batch must fit into what the rings let the analysers run ahead
    long long prefetch; // tiles between a claimed tile and the tile it prefetches into L2 (0: off)
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
    unsigned long long full[kStages];  // the tile has landed (TMA transaction count)
    unsigned long long done[kStages];  // the analyser warps are through with the stage (one arrival per warp)
    unsigned long long ready[kStages]; // the tile's prefix and row mask are in pref[] (one arrival, writer control warp)
    unsigned long long freeb[kStages]; // the writer warps are through with the stage (one arrival per warp): it may be reloaded
    unsigned long long pub[kStages];   // the tile's aggregate is published and its row mask is in aggmask[] (one arrival, analyser control warp)
    long long tile[kStages];           // tile in the stage (claimed by the producer from the grid-wide counter); -1: no more tiles
    unsigned long long aggmask[kStages]; // rows of the tile that need exact treatment by the writers (0: clean or handed over as records)
    uint32_t evcount[kStages];         // event records the analysers wrote for the tile in the stage
    WarpAgg wagg[kStages][kAWarps];    // analysers: per-warp aggregates of the tile in a stage
    WarpAgg waggW[kStages][kWWarps];   // writers: the same for the tiles at the edges of the byte range (interior tiles: wagg)
    TilePrefix pref[kStages];          // prefix + row mask of the tile in a stage
    uint8_t slowmapA[kAWarps][kARows * 32]; // analyser warp: its chunks that need exact analysis, in stream order
    uint8_t slowmapW[kWWarps][kWRows * 32]; // writer warp: the same
    alignas(16) uint16_t delmask[kWWarps][kWRows * 32]; // writer: removed bytes of every chunk of the warp's rows (tiles with removed bytes)
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
__device__ __forceinline__ bool mbar_test_wait(unsigned long long* bar, uint32_t parity) // never blocks
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Blocking wait.  A failed try_wait comes back after ~20 cycles, so a bare retry loop issues an instruction every few cycles per
// waiting warp: on streams of short NALs (long waits in both roles) a third of everything the SM issued were such polls, taken
// from the warps that had work.  After HEVCB_WAIT_POLLS misses the warp sleeps between polls (the bare loop
// is the form the compiler gives a YIELD; short waits -- the clean tiles of large NALs -- keep their wake-up latency).
#ifndef HEVCB_WAIT_SLEEP_NS
#define HEVCB_WAIT_SLEEP_NS 64
#endif
#ifndef HEVCB_WAIT_POLLS
#define HEVCB_WAIT_POLLS 32 // bare polls (~20 cycles each) before the warp starts to sleep between them: short waits keep their latency
#endif
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity)
{
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (HEVCB_WAIT_SLEEP_NS > 0 && ++polls > HEVCB_WAIT_POLLS) { __nanosleep(HEVCB_WAIT_SLEEP_NS); }
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// asks L2 to fetch a byte range from HBM (no destination on the SM): the tile's later bulk copy into shared memory then hits L2
__device__ __forceinline__ void tma_prefetch_l2(const void* src_gmem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// named barriers: 0 is __syncthreads; kBarWork = the analyser warps, kBarW = the writer warps
constexpr int kBarWork = 1, kBarW = 2;
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

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

// a worker warp is through with a stage: one arrival per warp
__device__ __forceinline__ void release_stage(unsigned long long* bar, int lane)
{
    __syncwarp();
    if (lane == 0) { mbar_arrive(bar); }
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
    unsigned long long next_tile; // grid-wide tile counter: a CTA with a free stage claims the next tile
    long long first_empty;
    ulonglong2 final_state; // inclusive prefix over all tiles, written by the scanner warp
    unsigned long long heavy_tiles;  // tiles the writers had to analyse themselves (a statistic: no launch parameter depends on earlier calls)
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

// the same for the pairs that START inside the chunk [g0, g0 + 16): needs byte g0 + 16 only (low byte of wn)
__device__ __forceinline__ uint32_t zero_pair_own(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wn)
{
    const uint32_t m0 = w0 | __funnelshift_r(w0, w1, 8);
    const uint32_t m1 = w1 | __funnelshift_r(w1, w2, 8);
    const uint32_t m2 = w2 | __funnelshift_r(w2, w3, 8);
    const uint32_t m3 = w3 | __funnelshift_r(w3, wn, 8);
    const uint32_t c = 0x01010101u, h = 0x80808080u;
    return (((m0 - c) & ~m0) | ((m1 - c) & ~m1) | ((m2 - c) & ~m2) | ((m3 - c) & ~m3)) & h;
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
                                             uint32_t nthreads = kWThreads)
{
#pragma unroll 4
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

// Scanner CTA (the last CTA of the grid; it takes no tiles): the chained scan over the tile aggregates.  Batches of 32 x per_lane
// consecutive tiles go round-robin to kScanWarps warps.  A warp waits for the aggregates the analyser control warps publish (one
// reader and one 16-byte poll per aggregate: an all-to-all look-back over hundreds of CTAs costs more in polling traffic on a few
// L2 lines than the whole pass, tests/micro/stream_ceiling.cu), scans them with shuffles / ballots, takes the running prefix from
// the warp of the previous batch through SHARED memory (the only serial step: ~100 cycles per batch, while the L2 round trips of
// the batches overlap across the warps) and publishes every tile's EXCLUSIVE prefix (start codes, kept bytes, ordered carry).
constexpr int kScanWarps = 16;
constexpr int kTimingIters = 256, kTimingEvents = 8;
// running prefix handed from batch to batch: ONE 16-byte shared-memory word, so that taking it over and handing it on is one
// vector load / store without fences: x = batches folded in so far (24 bits) | start codes (40 bits), y = kind | err | kept bytes
typedef ulonglong2 ScanRun;
__device__ __forceinline__ ulonglong2 lds_v2(const volatile ScanRun* p)
{
    ulonglong2 v;
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(smem_u32((const void*)p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_v2(volatile ScanRun* p, ulonglong2 v)
{
    asm volatile("st.volatile.shared.v2.u64 [%0], {%1, %2};" ::"r"(smem_u32((const void*)p)), "l"(v.x), "l"(v.y) : "memory");
}
__device__ __forceinline__ ulonglong2 pack_run(unsigned long long seq, unsigned long long n, unsigned long long k, uint32_t kind, uint32_t err)
{
    return make_ulonglong2((seq << 40) | (n & ((1ull << 40) - 1)), ((unsigned long long)kind << 62) | ((unsigned long long)(err & 1u) << 61) | k);
}
__device__ __forceinline__ void scanner_warps(const ulonglong2* __restrict__ tile_state, ulonglong2* __restrict__ tile_excl, long long n_tiles,
                                              ScanHeader* __restrict__ hdr, volatile ScanRun* run, int warp, int lane, int per_lane,
                                              unsigned long long* __restrict__ tdbg)
{
#ifndef HEVCB_SCAN_TIMING_BUILD
#define SSTAMP(b, ev) do { } while (0)
#else
#define SSTAMP(b, ev)                                                                                            \
    do {                                                                                                         \
        if (tdbg != nullptr && lane == 0 && ((b) & 7) == 0 && ((b) >> 3) < kTimingIters) {                       \
            unsigned long long ts_;                                                                              \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_));                                              \
            tdbg[((size_t)blockIdx.x * kTimingIters + (size_t)((b) >> 3)) * kTimingEvents + (ev)] = ts_;         \
        }                                                                                                        \
    } while (0)
#endif
    const long long batch_tiles = 32ll * per_lane;
    const long long n_batches = (n_tiles + batch_tiles - 1) / batch_tiles;
    for (long long b = warp; b < n_batches; b += kScanWarps) {
        const long long first = b * batch_tiles + (long long)lane * per_lane;
        SSTAMP(b, 0);
        ulonglong2 sv[kScanPerLane];
#pragma unroll
        for (int j = 0; j < kScanPerLane; j++) {
            const long long idx = first + j;
            if (j < per_lane && idx < n_tiles) { sv[j] = ld_state(&tile_state[idx]); }
            else { sv[j] = pack_agg(0, 0, HEVCB_KIND_PASS, 0, 0ull); } // past the end: identity
        }
        for (;;) { // re-poll, one batch per round trip, the aggregates that are not published yet
            bool missing = false;
#pragma unroll
            for (int j = 0; j < kScanPerLane; j++) { missing = missing || ((sv[j].x >> 62) == 0ull); }
            if (!__any_sync(0xFFFFFFFFu, missing)) { break; }
            __nanosleep(64);
#pragma unroll
            for (int j = 0; j < kScanPerLane; j++) {
                if ((sv[j].x >> 62) == 0ull) { sv[j] = ld_state(&tile_state[first + j]); }
            }
        }
        SSTAMP(b, 1);
        // lane totals
        uint32_t ln = 0, lk = 0, lkind = HEVCB_KIND_PASS, lerr = 0;
#pragma unroll
        for (int j = 0; j < kScanPerLane; j++) {
            ln += agg_n(sv[j]);
            lk += agg_k(sv[j]);
            hevcb_carry_combine(lkind, lerr, agg_kind(sv[j]), agg_err(sv[j]));
        }
        // exclusive scan over lanes (ascending lane = stream order), batch totals
        const uint32_t nin = warp_incl_scan(ln, lane), kin = warp_incl_scan(lk, lane);
        const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, lkind != HEVCB_KIND_PASS);
        const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lkind == HEVCB_KIND_SC3);
        const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, lerr != 0u);
        uint32_t ck, ce, bk, be;
        warp_carry_in(Eb, Sb, Rb, lane, ck, ce);
        warp_carry_total(Eb, Sb, Rb, bk, be);
        const uint32_t totN = __shfl_sync(0xFFFFFFFFu, nin, 31), totK = __shfl_sync(0xFFFFFFFFu, kin, 31);
        // serial step: running prefix of the batches before this one, handed on to the next batch's warp (every lane reads the
        // same word: a broadcast load, no divergence)
        ulonglong2 rw;
        do { rw = lds_v2(run); } while ((rw.x >> 40) != (unsigned long long)b);
        SSTAMP(b, 2);
        const unsigned long long runN = rw.x & ((1ull << 40) - 1), runK = rw.y & ((1ull << 61) - 1);
        const uint32_t runKind = (uint32_t)(rw.y >> 62), runErr = (uint32_t)(rw.y >> 61) & 1u;
        if (lane == 0) {
            uint32_t ok = runKind, oe = runErr;
            hevcb_carry_combine(ok, oe, bk, be);
            sts_v2(run, pack_run((unsigned long long)(b + 1), runN + totN, runK + totK, ok, oe));
            if (b == n_batches - 1) { hdr->final_state = pack_state(kStatusPrefix, runN + totN, runK + totK, ok, oe); }
        }
        uint32_t cKind = runKind, cErr = runErr; // carry entering this lane's first tile
        hevcb_carry_combine(cKind, cErr, ck, ce);
        unsigned long long cN = runN + (nin - ln), cK = runK + (kin - lk);
#pragma unroll
        for (int j = 0; j < kScanPerLane; j++) {
            const long long idx = first + j;
            if (j < per_lane && idx < n_tiles) { st_state(&tile_excl[idx], pack_state(kStatusPrefix, cN, cK, cKind, cErr)); }
            cN += agg_n(sv[j]);
            cK += agg_k(sv[j]);
            hevcb_carry_combine(cKind, cErr, agg_kind(sv[j]), agg_err(sv[j]));
        }
        SSTAMP(b, 3);
    }
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
    copy_span(dst, src, L, tid, kWThreads);
}

// one warp copies L bytes from shared memory at ANY alignment to global memory at any alignment: aligned 16-byte stores, the source as
// two aligned 16-byte loads funnel-shifted by the (copy-uniform) misalignment between source and destination
template <int Q>
__device__ __forceinline__ void smem_vectors_out(uint4* __restrict__ da, const uint8_t* __restrict__ sa, uint32_t nv, uint32_t sh, int lane)
{
    for (uint32_t i = (uint32_t)lane; i < nv; i += 32u) {
        const uint4 lo = *reinterpret_cast<const uint4*>(sa + (i << 4));
        uint4 o4 = lo;
        if (Q != 0 || sh != 0u) {
            const uint4 hi = *reinterpret_cast<const uint4*>(sa + (i << 4) + 16);
            const uint32_t W[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            o4.x = __funnelshift_r(W[Q], W[Q + 1], sh); o4.y = __funnelshift_r(W[Q + 1], W[Q + 2], sh);
            o4.z = __funnelshift_r(W[Q + 2], W[Q + 3], sh); o4.w = __funnelshift_r(W[Q + 3], W[Q + 4], sh);
        }
        __stcs(da + i, o4);
    }
}
__device__ __noinline__ void copy_span_any(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, uint32_t L, int lane)
{
    if (L == 0u) { return; }
    uint32_t head = (16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u;
    if (head > L) { head = L; }
    if ((uint32_t)lane < head) { dst[lane] = src[lane]; }
    const uint32_t nv = (L - head) >> 4;
    const uint8_t* s0 = src + head;
    const uint32_t mis = smem_u32(s0) & 15u, sh = (mis & 3u) * 8u;
    const uint8_t* sa = s0 - mis; // (reads stay inside the stage: at most 16 bytes behind the span, the stage has a 16-byte trailing halo)
    uint4* da = reinterpret_cast<uint4*>(dst + head);
    switch (mis >> 2) {
        case 0: smem_vectors_out<0>(da, sa, nv, sh, lane); break;
        case 1: smem_vectors_out<1>(da, sa, nv, sh, lane); break;
        case 2: smem_vectors_out<2>(da, sa, nv, sh, lane); break;
        default: smem_vectors_out<3>(da, sa, nv, sh, lane); break;
    }
    const uint32_t done = head + (nv << 4);
    if ((uint32_t)lane < L - done) { dst[done + lane] = src[done + lane]; }
}

// boundary fix-ups of a staged tile: positions < 0 read as non-zero, positions >= size read as zero
__device__ __forceinline__ void fix_stage(uint8_t* st, long long t, int64_t t0, int64_t size, int tid)
{
    const int64_t valid_end = (int64_t)kLead + (size - t0); // smem offset of position `size`
    if (t == 0 || valid_end < kStageBytes) {
        if (t == 0 && tid < kLead) { st[tid] = 0xFF; }
        if (valid_end < kStageBytes) {
            for (int i = (int)valid_end + tid; i < kStageBytes; i += kAThreads) { st[i] = 0; }
        }
        bar_sync(kBarWork, kAThreads);
    }
}

// Analyser, interior tile: exact analysis of the warp's "slow" chunks (two adjacent zero bytes nearby) only.  The slow chunks of
// the warp's eight rows are ranked in stream order (ballots), their indices parked in shared memory, and the lanes then take
// them 32 at a time: one pass of the exact masks with every lane busy instead of one pass per flagged row with a few lanes
// busy.  The ordered carry, the counts and the event records follow from ballots over the ranked chunks (the combine is
// associative, chunks without a zero pair contribute nothing).  slow8 bit i: this lane's chunk of row i is slow.
// (Results come back BY VALUE: reference parameters of an out-of-line function would pin the caller's running aggregates to the
// stack for the whole tile loop -- a local-memory store and load on the critical path of every tile, and with 200 KB of the SM's
// L1 carved out as shared memory those mostly miss.)
struct SlowAgg {
    uint32_t n, del, kind, err, rows;
};
__device__ __noinline__ SlowAgg analyse_slow_chunks(const uint8_t* __restrict__ st, uint8_t* __restrict__ slowmap, int warp, int lane, uint32_t slow8,
                                                    long long t, uint32_t* evcount, uint4* __restrict__ tile_events)
{
    uint32_t wN = 0, wDel = 0, wKind = HEVCB_KIND_PASS, wErr = 0, rows = 0;
    const uint32_t below = (1u << lane) - 1u;
    // The filter's "slow" is a superset (a chunk, its predecessor and its successor for every pair).  When that is more than one
    // batch of exact masks (streams of short NALs, dense payloads), the exact condition is worth its price: a chunk needs the masks
    // iff two adjacent zero bytes start inside it or in the two bytes in front of it (zero_pair_any, the writers' test) -- about a
    // third of the superset.  (Operands re-read from the stage: the caller's hot loop keeps its registers to itself.)
    if (HEVCB_SCAN_REFINE && __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(slow8)) > 32u) {
#pragma unroll 1
        for (int i = 0; i < kARows; i++) {
            if ((slow8 >> i) & 1u) {
                const uint8_t* rp = st + kLead + (warp * kARows + i) * kRowBytes + lane * 16;
                const uint4 v = *reinterpret_cast<const uint4*>(rp);
                const uint32_t wp = *reinterpret_cast<const uint32_t*>(rp - 4);
                const uint32_t wn = *reinterpret_cast<const uint32_t*>(rp + 16);
                if (zero_pair_any(wp, v.x, v.y, v.z, v.w, wn) == 0u) { slow8 &= ~(1u << i); }
            }
        }
        __syncwarp();
    }
    uint32_t total = 0;
#pragma unroll
    for (int i = 0; i < kARows; i++) {
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
            chunk = (uint32_t)(warp * kARows * 32) + slowmap[slot];
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
        // chunks with an event or an error position leave a record (one shared-memory atomic per batch) -- unless the tile already
        // has more of them than the list takes: it then goes to the writers as a whole and its records are never read
        if (Xb != 0u && *reinterpret_cast<volatile uint32_t*>(evcount) <= kEvCap) {
            uint32_t first = 0u;
            if (lane == 0) { first = atomicAdd(evcount, (uint32_t)__popc(Xb)); }
            first = __shfl_sync(0xFFFFFFFFu, first, 0);
            const uint32_t rslot = first + (uint32_t)__popc(Xb & below);
            if (((ev | er) != 0u) && rslot < kEvCap) { tile_events[(size_t)t * kEvCap + rslot] = make_uint4(chunk, evsc, deler, misc); }
        }
    }
    __syncwarp(); // the map is rewritten for the warp's next tile
    SlowAgg r;
    r.n = wN; r.del = wDel; r.kind = wKind; r.err = wErr; r.rows = rows;
    return r;
}

// ====================================================================================================================
// One CTA per SM; every CTA is a pipeline of specialised warps over a ring of kStages tile buffers, CTA c takes the tiles
// c, c + G, c + 2G, ...:
//   PRODUCER (one thread)  waits until the writers have released a stage and reloads it with the CTA's next tile (TMA).
//   ANALYSER warps         build the exact predicates where two adjacent zero bytes occur and leave per-warp aggregates;
//   the analyser CONTROL warp publishes one 16-byte aggregate per tile: start codes, kept bytes, ordered carry, and either a
//                          64-bit mask of the rows that need exact treatment or the number of event records left for the emit pass.
//   The SCANNER warp (one per grid) turns aggregates into exclusive prefixes, in stream order.
//   the writer CONTROL warp polls the prefix of the CTA's next tiles (a few tiles ahead) and hands it to the
//   WRITER warps,          which write the EPB-free image out of the SAME stage (redoing the exact analysis of flagged rows
//                          only) and emit the NAL boundaries of heavy tiles.
// The tile is read from HBM once and never reloaded.  Inside a CTA the roles are coupled by mbarriers per stage only (full ->
// done -> ready -> free): the analysers run ahead of the writers by as many tiles as the scanner needs to deliver a prefix,
// so the prefix round trip through L2 (as long as the work on a tile) never stalls a warp that has work to do.
// ====================================================================================================================
__global__ void __launch_bounds__(kThreads, HEVCB_SCAN_CTAS) hevcb_scan_strip_kernel(
    const uint8_t* __restrict__ buf, const ScanGeom geom, long long n_tiles, ScanHeader* __restrict__ hdr,
    ulonglong2* __restrict__ tile_state, ulonglong2* __restrict__ tile_excl, uint4* __restrict__ tile_events,
    int64_t* __restrict__ nal_start, int64_t* __restrict__ nal_end, int64_t cap_nals, uint8_t* __restrict__ rbsp,
    int64_t* __restrict__ rbsp_off, int64_t* __restrict__ rbsp_end, long long debug_flags, unsigned long long* __restrict__ tdbg)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
    const int lane = threadIdx.x & 31;
    // measurement aid (HEVCB_SCAN_TIMING=1): globaltimer stamps of the pipeline events of the first kTimingIters tiles of every CTA
    // (compiled in with -DHEVCB_SCAN_TIMING_BUILD only, `make EXTRA=-DHEVCB_SCAN_TIMING_BUILD`: the stamps' addresses and predicates
    // cost the hot loops registers, and the writers kept them on the stack)
#ifndef HEVCB_SCAN_TIMING_BUILD
#define TSTAMP(iter, ev) do { } while (0)
#else
#define TSTAMP(iter, ev)                                                                                         \
    do {                                                                                                         \
        if (tdbg != nullptr && (iter) < kTimingIters) {                                                          \
            unsigned long long ts_;                                                                              \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts_));                                              \
            tdbg[((size_t)blockIdx.x * kTimingIters + (size_t)(iter)) * kTimingEvents + (ev)] = ts_;             \
        }                                                                                                        \
    } while (0)
#endif
    const int role_warp = threadIdx.x >> 5;
    const unsigned dbg = (unsigned)debug_flags; // experiment switches, 0 in production
    const int64_t size = geom.size;

    if (blockIdx.x == gridDim.x - 1) {
        // ============================================ SCANNER CTA ============================================
        volatile ScanRun* run = reinterpret_cast<volatile ScanRun*>(smem_raw);
        if (threadIdx.x == 0) { sts_v2(run, pack_run(0ull, geom.init_n, 0ull, geom.init_kind, 0u)); }
        __syncthreads();
        if (role_warp < kScanWarps) { scanner_warps(tile_state, tile_excl, n_tiles, hdr, run, role_warp, lane, geom.scan_per_lane, tdbg); }
        return;
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&sm.full[s], 1); mbar_init(&sm.done[s], kAWarps); mbar_init(&sm.ready[s], 1); mbar_init(&sm.freeb[s], kWWarps);
            mbar_init(&sm.pub[s], 1);
            sm.evcount[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
    }
    __syncthreads();

    if (role_warp == kWarpProd) {
        // ============================================ PRODUCER ============================================
        // Tiles are claimed from a grid-wide counter the moment a stage is free, so a CTA that falls behind simply takes fewer
        // tiles: aggregates are published in (nearly) tile order and a scanner batch never waits for a slow CTA's turn.
        if (lane == 0) {
            for (long long i = 0;; i++) {
                const int s = (int)(i % kStages);
                const uint32_t round = (uint32_t)(i / kStages);
                if (round >= 1u) { mbar_wait(&sm.freeb[s], (round - 1u) & 1u); } // the writers have released the stage
                const long long t = (long long)atomicAdd(&hdr->next_tile, 1ull);
                TSTAMP(i, 0);
                if (t >= n_tiles) { // end marker: travels through the roles like a tile
                    sm.tile[s] = -1;
                    mbar_arrive(&sm.full[s]);
                    break;
                }
                sm.tile[s] = t;
                fence_proxy_async();
                issue_tile_load(sm.stage[s], &sm.full[s], buf, size, t);
                // Every claimed tile also pulls the tile `prefetch` tiles further on from HBM into L2.  The ring of an SM is small
                // (six tiles) and a tile cannot leave it before the prefixes of all earlier tiles are known, so the ring must not
                // also have to cover HBM latency and its variance: with the prefetch the bulk copies are L2 hits.
                const long long tp = t + geom.prefetch;
                if (geom.prefetch > 0 && tp < n_tiles) {
                    const int64_t p0 = (int64_t)tp * kTileBytes;
                    const int64_t size16 = (size + 15) & ~(int64_t)15;
                    const int64_t p1 = p0 + kTileBytes < size16 ? p0 + kTileBytes : size16;
                    if (p1 > p0) { tma_prefetch_l2(buf + p0, (uint32_t)(p1 - p0)); }
                }
            }
        }
        return;
    }

    if (role_warp == kWarpAC) {
        // ============================================ ANALYSER CONTROL ============================================
        // all analyser warps are done with the stage -> publish the tile's aggregate
        int s = 0;
        uint32_t done_bits = 0;
        unsigned long long heavy_local = 0, flagged_local = 0;
        for (long long it_ac = 0;; it_ac++) {
            mbar_wait(&sm.done[s], (done_bits >> s) & 1u);
            done_bits ^= (1u << s);
            const long long t = sm.tile[s];
            if (t < 0) {
                if (lane == 0) { mbar_arrive(&sm.pub[s]); }
                break;
            }
            if (lane == 0) { TSTAMP(it_ac, 2); }
            uint32_t tile_n = 0, tile_k = 0, ak = HEVCB_KIND_PASS, ae = 0, adel = 0;
            unsigned long long mask = 0ull;
#pragma unroll
            for (int w = 0; w < kAWarps; w++) {
                const WarpAgg a = sm.wagg[s][w];
                tile_n += a.n;
                tile_k += a.k;
                hevcb_carry_combine(ak, ae, a.kind, a.err);
                mask |= (unsigned long long)(a.rows & ((1u << kARows) - 1u)) << (w * kARows);
                adel |= a.del;
            }
            if (lane == 0) {
                // A tile whose bytes are all kept and whose event chunks fit into the record list looks CLEAN to the
                // writers (mask 0): its NAL boundaries are emitted from the records by hevcb_scan_emit_kernel.
                const uint32_t nrec = sm.evcount[s];
                sm.evcount[s] = 0;
                const bool light = (adel == 0u) && (nrec <= kEvCap) && (tile_k == (uint32_t)kTileBytes) && !(dbg & 1024u);
                st_state(&tile_state[t], light ? pack_agg(tile_n, tile_k, ak, ae, 0ull, nrec) : pack_agg(tile_n, tile_k, ak, ae, mask, kEvByWriter));
                sm.aggmask[s] = light ? 0ull : mask;
                mbar_arrive(&sm.pub[s]);
                if (!light && mask != 0ull) { heavy_local++; }
                flagged_local += (unsigned long long)__popcll(mask);
            }
            __syncwarp();
            s = (s + 1 == kStages) ? 0 : s + 1;
        }
        if (lane == 0 && heavy_local) { atomicAdd(&hdr->heavy_tiles, heavy_local); }
        if (lane == 0 && flagged_local) { atomicAdd(&hdr->flagged_rows, flagged_local); }
        return;
    }

    if (role_warp == kWarpWC) {
        // ============================================ WRITER CONTROL ============================================
        // Lane j looks after the CTA's j-th next tile (kAhead tiles ahead of the writers): it waits until the tile's aggregate
        // is published (then the tile index and row mask are known), then polls the tile's prefix, so that the L2 round trips
        // never sit between two tiles.  Per tile: lane 0 has the prefix -> shared memory -> release the writers into the
        // tile -> every lane hands its state down by one lane.
        constexpr int kAhead = 4;
        long long my_i = lane;          // iteration (position in this CTA's tile sequence) this lane looks after
        long long tq = -2;              // its tile: -2 not known yet, -1 end marker
        unsigned long long mk = 0ull;   // its row mask
        ulonglong2 ex = make_ulonglong2(0ull, 0ull);
        for (long long it_wc = 0;; it_wc++) {
            const int s = (int)(it_wc % kStages);
            for (;;) {
                if (lane < kAhead) {
                    const int ms = (int)(my_i % kStages);
                    if (tq == -2 && mbar_test_wait(&sm.pub[ms], (uint32_t)(my_i / kStages) & 1u)) { tq = sm.tile[ms]; mk = sm.aggmask[ms]; }
                    if (tq >= 0 && (ex.x >> 62) == 0ull) { ex = ld_state(&tile_excl[tq]); }
                }
                const bool ready = (tq == -1) || (tq >= 0 && (ex.x >> 62) != 0ull);
                if (__shfl_sync(0xFFFFFFFFu, ready ? 1 : 0, 0)) { break; }
                __nanosleep(20);
            }
            const long long t0q = __shfl_sync(0xFFFFFFFFu, tq, 0);
            if (lane == 0) {
                // pref[s] is free: the stage's previous tile was written before the stage could be reloaded and analysed again
                if (tq >= 0) {
                    TilePrefix tp;
                    tp.n = ex.x & ((1ull << 40) - 1); tp.k = ex.y; tp.kind = (uint32_t)(ex.x >> 60) & 3u; tp.err = (uint32_t)(ex.x >> 59) & 1u;
                    tp.mask = mk;
                    sm.pref[s] = tp;
                    TSTAMP(it_wc, 3);
                }
                mbar_arrive(&sm.ready[s]);
            }
            __syncwarp();
            if (t0q < 0) { break; }
            // hand down: lane j takes over what lane j + 1 knows so far; the last looking lane starts on a new tile
            ex.x = __shfl_down_sync(0xFFFFFFFFu, ex.x, 1); ex.y = __shfl_down_sync(0xFFFFFFFFu, ex.y, 1);
            tq = __shfl_down_sync(0xFFFFFFFFu, tq, 1); mk = __shfl_down_sync(0xFFFFFFFFu, mk, 1);
            my_i += 1;
            if (lane >= kAhead - 1) { ex = make_ulonglong2(0ull, 0ull); tq = -2; mk = 0ull; }
        }
        return;
    }

    if (role_warp < kAWarps) {
        // ============================================ ANALYSER ============================================
        const int tid = threadIdx.x;
        // workers: wait for the tile, analyse their rows, hand the warp aggregate to the control warp; they never wait for
        // it (the stage they move on to was loaded three tiles ago; a stage is only reloaded after its aggregate was read)
        uint32_t phase_bits = 0;
        int s = 0;
        for (long long it_an = 0;; it_an++) {
            uint8_t* st = sm.stage[s];
            mbar_wait(&sm.full[s], (phase_bits >> s) & 1u);
            phase_bits ^= (1u << s);
            const long long t = sm.tile[s];
            if (t < 0) { // end marker: hand it on
                if (lane == 0) { mbar_arrive(&sm.done[s]); }
                break;
            }
            const int64_t t0 = (int64_t)t * kTileBytes;
            // Which rows a warp takes rotates from tile to tile: in a stream of equally sized NALs the start codes keep falling
            // into the same rows of every tile, and the exact analysis of such a row (about as long as the whole fast path)
            // would otherwise always land on the same warp and set the pace of the CTA.
            const int warp = (role_warp + (int)(it_an & (kAWarps - 1))) & (kAWarps - 1);
            if (tid == 0) { TSTAMP(it_an, 1); }
            fix_stage(st, t, t0, size, tid);
            const bool interior = geom.evl - t0 >= (int64_t)kTileBytes + 32; // every byte (and its halo) owned and below the event limit
            uint32_t wN = 0, wK = 0, wKind = HEVCB_KIND_PASS, wErr = 0, rows = 0, anydel = 0;
            if (dbg & 64u) { wK = kARows * kRowBytes; rows = (dbg >> 8) & 0xFFu; } // experiment: analysis off, rows flagged as told
            else if (interior && !(dbg & 2048u)) {
                // interior tile: which chunks have two adjacent zero bytes NEARBY (inside [g0 - 2, g0 + 18)) and need the exact masks?
                // Every lane tests the pairs that START in its own chunk (needs one byte of the next chunk: a shuffle), one ballot per
                // row collects them, and a chunk is slow when it, its predecessor or its successor starts a pair (a superset of the
                // exact condition; the next / previous chunk of the warp's first / last chunk come from two single-lane halo loads).
                // All rows are loaded up front: eight independent 16-byte shared-memory loads per lane in flight.
                const uint8_t* wbase = st + kLead + (warp * kARows) * kRowBytes;
                // (written in phases -- loads, shuffles, arithmetic, ballots -- because the warp-synchronous operations keep their
                // program order: interleaving them row by row would serialise the rows' dependency chains)
                uint4 v[kARows];
                uint32_t wn[kARows];
#pragma unroll
                for (int k = 0; k < kARows; k++) { v[k] = *reinterpret_cast<const uint4*>(wbase + k * kRowBytes + lane * 16); }
#pragma unroll
                for (int k = 0; k < kARows; k++) { // lane 31: the word behind its chunk is the first word of the next row
                    wn[k] = 0u;
                    if (lane == 31) { wn[k] = *reinterpret_cast<const uint32_t*>(wbase + (k + 1) * kRowBytes); }
                }
                uint32_t edge = 0; // lane 0: the four bytes in front of the warp's rows; lane 31: the four bytes behind them
                if (lane == 0) { edge = *reinterpret_cast<const uint32_t*>(wbase - 4); }
                if (lane == 31) { edge = wn[kARows - 1]; }
#pragma unroll
                for (int k = 0; k < kARows; k++) {
                    const uint32_t nx = __shfl_down_sync(0xFFFFFFFFu, v[k].x, 1);
                    if (lane != 31) { wn[k] = nx; }
                }
                bool own[kARows];
#pragma unroll
                for (int k = 0; k < kARows; k++) { own[k] = zero_pair_own(v[k].x, v[k].y, v[k].z, v[k].w, wn[k]) != 0u; }
                uint32_t P[kARows];
#pragma unroll
                for (int k = 0; k < kARows; k++) { P[k] = __ballot_sync(0xFFFFFFFFu, own[k]); }
                // pairs that start in the two bytes in front of the warp's rows (or in the last of them with the row's first byte),
                // and the pair made of the two bytes behind them
                const bool pf = (lane == 0) && (((edge >> 16) == 0u) || (((edge >> 24) == 0u) && ((v[0].x & 0xFFu) == 0u)));
                const bool nf = (lane == 31) && ((edge & 0xFFFFu) == 0u);
                const uint32_t prevflag = __ballot_sync(0xFFFFFFFFu, pf) & 1u, nextflag = (__ballot_sync(0xFFFFFFFFu, nf) >> 31) & 1u;
                uint32_t slow8 = 0, rowany = 0;
#pragma unroll
                for (int k = 0; k < kARows; k++) {
                    const uint32_t before = (k > 0) ? (P[(k + kARows - 1) % kARows] >> 31) : prevflag;
                    const uint32_t after = (k + 1 < kARows) ? (P[(k + 1) % kARows] & 1u) : nextflag;
                    const uint32_t S = P[k] | (P[k] << 1) | (P[k] >> 1) | before | (after << 31);
                    slow8 |= ((S >> lane) & 1u) << k;
                    rowany |= (S != 0u ? 1u : 0u) << k;
                }
                if (tid == 0) { TSTAMP(it_an, 7); }
                uint32_t wDel = 0; // rowany bit i: row i has a slow chunk
                if (rowany != 0u) {
                    const uint32_t thr = (dbg & 0xF000u) ? ((dbg >> 12) & 15u) - 1u : 1u; // experiment switch; production: 1
                    if ((uint32_t)__popc(rowany) <= thr) {
                        // A single flagged row (a stream of large NALs: the row a start code falls into): exact masks of the
                        // flagged lanes straight from the registers the filter loaded, one row = one lane per chunk, so the ordered
                        // carry is three ballots.  (Ranking the chunks first pays off when many rows are flagged.)
                        // (a rolled loop with register selects instead of an unrolled one: this path runs once per tile in one or
                        // two warps, its instructions are rarely in the instruction cache, so its size is what it costs)
#pragma unroll 1
                        for (uint32_t rm = rowany; rm != 0u; rm &= rm - 1u) {
                            const int k = __ffs((int)rm) - 1; // warp-uniform
                            uint4 vk = v[0];
                            uint32_t wnk = wn[0], pw = 0u;
#pragma unroll
                            for (int q = 1; q < kARows; q++) {
                                if (k == q) { vk = v[q]; wnk = wn[q]; pw = v[q - 1].w; }
                            }
                            uint32_t wp = __shfl_up_sync(0xFFFFFFFFu, vk.w, 1);
                            const uint32_t prow = __shfl_sync(0xFFFFFFFFu, pw, 31);
                            if (lane == 0) { wp = (k > 0) ? prow : edge; }
                            uint32_t ev = 0u, sc = 0u, del = 0u, er = 0u, scb = 0u;
                            if ((slow8 >> k) & 1u) {
                                const hevcb_chunk_masks m = hevcb_chunk_analyze_interior(wp, vk.x, vk.y, vk.z, vk.w, wnk);
                                ev = m.ev; sc = m.sc; del = m.del; er = m.err; scb = m.scb;
                            }
                            uint32_t lk = HEVCB_KIND_PASS, le = (er != 0u); // lane summary for the ordered carry
                            if (ev != 0u) {
                                const int tp = 31 - __clz((int)ev);
                                lk = ((sc >> tp) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
                                le = ((er >> tp) >> 1) != 0u;
                            }
                            const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, ev != 0u);
                            const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lk == HEVCB_KIND_SC3);
                            const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, le != 0u);
                            const uint32_t Db = __ballot_sync(0xFFFFFFFFu, del != 0u);
                            const uint32_t Cb = __ballot_sync(0xFFFFFFFFu, sc != 0u); // (Sb only tells whose LAST event is a start code)
                            if (Cb != 0u) { wN += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(sc)); }
                            if (Db != 0u) { wDel += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(del)); }
                            uint32_t rk, re;
                            warp_carry_total(Eb, Sb, Rb, rk, re);
                            hevcb_carry_combine(wKind, wErr, rk, re);
                            rows |= 1u << k;
                            if ((ev | er) != 0u) { // chunk with an event or an error position: leave a record for the emit pass
                                const uint32_t slot = atomicAdd(&sm.evcount[s], 1u);
                                if (slot < kEvCap) {
                                    tile_events[(size_t)t * kEvCap + slot] =
                                        make_uint4((uint32_t)((warp * kARows + k) * 32 + lane), ev | (sc << 16), del | (er << 16), 0xFFFFu | (scb << 16));
                                }
                            }
                        }
                    } else {
                        const SlowAgg r = analyse_slow_chunks(st, sm.slowmapA[role_warp], warp, lane, slow8, t, &sm.evcount[s], tile_events);
                        wN = r.n; wDel = r.del; wKind = r.kind; wErr = r.err; rows = r.rows;
                    }
                }
                wK = kARows * kRowBytes - wDel;
                anydel = wDel;
            }
            else
            // Rows are taken four at a time: the four loads, halo exchanges and zero-pair tests are independent
            // instruction chains, and one vote sends the common "no two adjacent zero bytes anywhere" case on.
            // (Loops over rows are kept rolled on purpose: two roles share the SM's instruction cache.)
#pragma unroll kAnalyserUnroll
            for (int i0 = 0; i0 < kARows; i0 += 4) {
                const int rbase = warp * kARows + i0;
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
                if (tid == 0) { TSTAMP(it_an, 6); }
                mbar_arrive(&sm.done[s]);
            }
            __syncwarp();
            s = (s + 1 == kStages) ? 0 : s + 1;
        }
        return;
    }

    // ================================================ WRITER ================================================
    const int tid = (int)threadIdx.x - kAThreads;
    const int warp = tid >> 5;
    DevSink sink{nal_start, nal_end, rbsp_off, rbsp_end, cap_nals, &hdr->first_empty};
    // (one loop-carried value: the stage and its phase are derived from the iteration count; with three of them the compiler kept
    // the loop state on the stack -- local memory in the hot loop, see SlowAgg)
    for (uint32_t it = 0;; it++) {
        const int s = (int)(it % (uint32_t)kStages);
        const uint32_t phase = (it / (uint32_t)kStages) & 1u;
        uint8_t* st = sm.stage[s];
        mbar_wait(&sm.ready[s], phase); // the tile's prefix and row mask are in shared memory
        const long long t = sm.tile[s];
        if (t < 0) { break; } // end marker
        const int64_t t0 = (int64_t)t * kTileBytes;
        const TilePrefix pref = sm.pref[s];
        mbar_wait(&sm.full[s], phase);
        if (tid == 0) { TSTAMP(it, 4); }
        const long long tileN = (long long)pref.n;
        const long long tileK = (long long)pref.k;

        if (pref.mask == 0ull) {
            // ---- clean interior tile: aligned 16-byte vectors, funnel-shifted by the (tile-uniform) misalignment
            if (rbsp != nullptr && !(dbg & 4u)) { copy_tile(rbsp + tileK, st + kLead, kTileBytes, tid); }
            if (tid == 0) { TSTAMP(it, 5); }
            release_stage(&sm.freeb[s], lane);
            continue;
        }

        // ---- tile with flagged rows: exact masks of those rows, warp aggregates, ordered emission, row-wise write-out
        if (geom.evl - t0 >= (int64_t)kTileBytes + 32 && !(dbg & (65536u | 2048u | 64u))) {
            // Interior tile.  The warp ranks the slow chunks of its eight rows in stream order and takes them 32 at a time
            // (every lane busy, see analyse_slow_chunks): exact masks and ordered emission in one pass.
            uint8_t* const map = sm.slowmapW[warp];
            const uint32_t below = (1u << lane) - 1u;
            uint32_t total = 0;
            // rows the analysers did not flag hold no slow chunk (their test is a superset of the one below): a tile with one removed
            // byte -- a stream with an emulation prevention byte every few NALs -- is ranked in the few flagged rows only
            const uint32_t flagged8 = (uint32_t)(pref.mask >> (warp * kWRows)) & ((1u << kWRows) - 1u);
#pragma unroll 1
            for (int i0 = 0; i0 < kWRows; i0 += 4) {
                if (((flagged8 >> i0) & 0xFu) == 0u) { continue; }
                const uint8_t* rp = st + kLead + (warp * kWRows + i0) * kRowBytes + lane * 16;
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
            constexpr int kBatches = kWRows;
            // The warp aggregates of the tile are the ones the ANALYSERS left in shared memory (one per group of kARows rows; the
            // stage is not reloaded, so they are not rewritten, before the writers release it): what lies in front of this warp's
            // rows -- start codes, kept bytes, ordered carry -- and how many bytes the tile and this warp's rows lose are known
            // before the warp looks at a single chunk.  The writers therefore build every exact mask once, emit straight from it,
            // and do not wait for each other (only tiles that lose bytes keep one barrier, in front of the in-place compaction).
            constexpr int kGroupsPerW = kWRows / kARows;
            static_assert(kAWarps <= 32 && kWRows % kARows == 0, "one lane per analyser aggregate");
            uint32_t rN, rK, rKind, rErr, tileDel, wDel;
            bool moving = false;
            {
                uint32_t gn = 0, gk = 0, gkind = HEVCB_KIND_PASS, gerr = 0, gdel = 0;
                if (lane < kAWarps) {
                    const WarpAgg a = sm.wagg[s][lane];
                    gn = a.n; gk = a.k; gkind = a.kind; gerr = a.err; gdel = a.del;
                }
                const int g0 = warp * kGroupsPerW;
                const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, gkind != HEVCB_KIND_PASS);
                const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, gkind == HEVCB_KIND_SC3);
                const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, gerr != 0u);
                warp_carry_in(Eb, Sb, Rb, g0, rKind, rErr);
                rN = __reduce_add_sync(0xFFFFFFFFu, lane < g0 ? gn : 0u);
                rK = __reduce_add_sync(0xFFFFFFFFu, lane < g0 ? gk : 0u);
                tileDel = __reduce_add_sync(0xFFFFFFFFu, gdel);
                wDel = 0;
                if (tileDel != 0u) {
                    static_assert(kGroupsPerW == 2, "a writer warp's rows are two analyser groups");
                    const uint32_t pairdel = gdel + __shfl_xor_sync(0xFFFFFFFFu, gdel, 1);
                    wDel = __shfl_sync(0xFFFFFFFFu, pairdel, g0);
                    moving = __any_sync(0xFFFFFFFFu, pairdel > 4u); // some writer warp closes gaps in place (the same answer in every warp)
                }
            }
            const bool write_img = (rbsp != nullptr) && !(dbg & 4u);
            const bool compacting = write_img && tileDel != 0u && wDel != 0u; // this warp's rows lose bytes
            uint16_t* const dm = sm.delmask[warp];
            if (compacting) {
                for (int e = lane * 8; e < kWRows * 32; e += 256) { *reinterpret_cast<uint4*>(dm + e) = make_uint4(0u, 0u, 0u, 0u); }
                __syncwarp();
            }
            // one pass over the ranked chunks: exact masks, ordered emission
            {
                uint32_t cKind = pref.kind, cErr = pref.err; // carry entering this warp = tile carry (+) warps before it
                hevcb_carry_combine(cKind, cErr, rKind, rErr);
                uint32_t nrun = rN, drun = (uint32_t)(warp * kWRows * kRowBytes) - rK; // start codes / removed bytes before, within the tile
#pragma unroll 1
                for (int b = 0; b < kBatches; b++) {
                    const uint32_t slot = (uint32_t)b * 32u + (uint32_t)lane;
                    if ((uint32_t)b * 32u >= total) { break; }
                    uint32_t evsc = 0u, deler = 0u, misc = 0xFFFFu, local = 0u;
                    if (slot < total) {
                        local = (uint32_t)map[slot];
                        const uint8_t* rp = st + kLead + ((uint32_t)(warp * kWRows * 32) + local) * 16u;
                        const uint4 v = *reinterpret_cast<const uint4*>(rp);
                        const uint32_t wp = *reinterpret_cast<const uint32_t*>(rp - 4);
                        const uint32_t wn = *reinterpret_cast<const uint32_t*>(rp + 16);
                        const uint3 m3 = analyze_interior_cold(wp, v, wn); // interior tile: no position limits
                        evsc = m3.x; deler = m3.y; misc = m3.z;
                    }
                    const uint32_t ev = evsc & 0xFFFFu, sc = evsc >> 16, del = deler & 0xFFFFu, er = deler >> 16;
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
                        const uint32_t chunk = (uint32_t)(warp * kWRows * 32) + local;
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
            // bytes move inside the stage from here on: every writer warp must be through with its neighbours' edge bytes
            if (write_img && moving) { bar_sync(kBarW, kWThreads); }
            if (write_img) {
                if (tileDel == 0u) {
                    copy_tile(rbsp + tileK, st + kLead, kTileBytes, tid); // every byte kept: one shifted vector copy by all workers
                } else {
                    // tile with removed bytes: every warp closes the gaps of its own rows in place (shared memory), then copies
                    // its kept bytes out as one span
                    uint8_t* const wb = st + kLead + warp * (kWRows * kRowBytes);
                    uint32_t out = kWRows * kRowBytes;
                    if (compacting && wDel <= 4u) {
                        // A few removed bytes (an emulation prevention byte in a slice header now and then: the common case of real
                        // streams): the kept bytes are a few runs; each goes out as one vector copy of its own, nothing moves in
                        // shared memory (closing the gaps byte by byte made this warp the straggler of its tile).
                        __syncwarp();
                        uint32_t cur = 0;
                        uint8_t* dptr = rbsp + tileK + rK;
#pragma unroll 1
                        for (int i = 0; i < kWRows; i++) {
                            const uint32_t del = dm[i * 32 + lane];
                            uint32_t lanes = __ballot_sync(0xFFFFFFFFu, del != 0u);
                            while (lanes) {
                                const int src = __ffs((int)lanes) - 1;
                                lanes &= lanes - 1u;
                                uint32_t m = __shfl_sync(0xFFFFFFFFu, del, src);
                                while (m) {
                                    const uint32_t pos = (uint32_t)((i * 32 + src) * 16 + (__ffs((int)m) - 1));
                                    m &= m - 1u;
                                    copy_span_any(dptr, wb + cur, pos - cur, lane);
                                    dptr += pos - cur;
                                    cur = pos + 1u;
                                }
                            }
                        }
                        copy_span_any(dptr, wb + cur, (uint32_t)(kWRows * kRowBytes) - cur, lane);
                        if (tid == 0) { TSTAMP(it, 5); }
                        release_stage(&sm.freeb[s], lane);
                        continue;
                    }
                    if (compacting) {
                        __syncwarp();
                        out = 0;
#pragma unroll 1
                        for (int i = 0; i < kWRows; i++) {
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
            if (tid == 0) { TSTAMP(it, 5); }
            release_stage(&sm.freeb[s], lane); // the control warp may now reload this stage; workers do not wait
            continue;
        }
        const uint32_t myrows = (uint32_t)(pref.mask >> (warp * kWRows)) & 0xFFu;
        // pass 1: aggregates of this warp's flagged rows; their masks are parked (thread-local memory) for pass 2
        uint32_t wN = 0, wK = 0, wKind = HEVCB_KIND_PASS, wErr = 0, anyD = 0;
        uint32_t k_evsc[kWRows], k_deler[kWRows], k_misc[kWRows], k_flags = 0;
#pragma unroll 1
        for (int i = 0; i < kWRows; i++) {
            if (!((myrows >> i) & 1u)) { wK += kRowBytes; continue; }
            const RowMasks m = analyze_staged_row(st, warp * kWRows + i, lane, t0, geom, wN, wK, wKind, wErr);
            anyD |= (m.Db != 0u) ? 1u : 0u;
            k_evsc[i] = m.evsc; k_deler[i] = m.deler; k_misc[i] = m.misc;
            k_flags |= ((m.Xb != 0u) ? 1u : 0u) << i;
            k_flags |= ((m.Db != 0u) ? 1u : 0u) << (8 + i);
        }
        if (lane == 0) {
            WarpAgg a;
            a.n = wN; a.k = wK; a.kind = wKind; a.err = wErr; a.del = anyD; a.rows = myrows;
            a.pad[0] = a.pad[1] = 0;
            sm.waggW[s][warp] = a;
        }
        bar_sync(kBarW, kWThreads);
        // aggregate of the warps before this one, tile totals
        uint32_t rN = 0, rK = 0, rKind = HEVCB_KIND_PASS, rErr = 0, compact = 0, tile_k = 0;
        {
            uint32_t tn = 0, ak = HEVCB_KIND_PASS, ae = 0;
#pragma unroll
            for (int w = 0; w < kWWarps; w++) {
                const WarpAgg a = sm.waggW[s][w];
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
            for (int i = 0; i < kWRows; i++) {
                const int r = warp * kWRows + i;
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
        if (tid == 0) { TSTAMP(it, 5); }
        release_stage(&sm.freeb[s], lane); // the control warp may now reload this stage; workers do not wait
    }
}

// Emit pass: NAL boundaries of the tiles whose event chunks the analysers left as records.  Records are in arrival order: they are
// ranked by chunk index, then walked with the same ordered-carry logic a row of the writer runs.  Tiles of that kind keep all their
// bytes, so a chunk's image offset is prefix + chunk * 16.
// A warp looks after kEmitTilesPerWarp consecutive tiles.  A tile with at most kEmitSerial records -- every tile of a stream of large NALs: a 32 KiB
// tile of 16 KiB NALs holds two start codes -- is walked by ONE THREAD (a sorting network over registers, then hevcb_chunk_emit
// record by record), the warp's tiles side by side.  (One warp per tile, ranking by 32 shuffles and emitting with ballots and scans,
// cost ~500 instructions per tile whatever the record count: 105 us for the 131 072 tiles of the 4 GiB headline stream, 7 % of the
// step.)  Tiles with more records are then taken one at a time by the whole warp, one lane per record.
constexpr int kEmitSerial = 4;
constexpr int kEmitTilesPerWarp = 4; // (not 32: the tiles with many records of a warp are taken one after the other, and the grid must still fill the GPU)
constexpr int kEmitTilesPerBlock = 8 * kEmitTilesPerWarp;
__device__ __forceinline__ void rec_cswap(uint4& a, uint4& b)
{
    if (b.x < a.x) { const uint4 t = a; a = b; b = t; }
}
__global__ void __launch_bounds__(256) hevcb_scan_emit_kernel(long long n_tiles, ScanHeader* __restrict__ hdr, const ulonglong2* __restrict__ tile_state,
                                                              const ulonglong2* __restrict__ tile_excl, const uint4* __restrict__ tile_events,
                                                              int64_t* __restrict__ nal_start, int64_t* __restrict__ nal_end, int64_t cap_nals,
                                                              int64_t* __restrict__ rbsp_off, int64_t* __restrict__ rbsp_end)
{
    __shared__ uint4 sorted[8][kEvCap];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long first_tile = ((long long)blockIdx.x * 8 + warp) * kEmitTilesPerWarp;
    if (first_tile >= n_tiles) { return; }
    DevSink sink{nal_start, nal_end, rbsp_off, rbsp_end, cap_nals, &hdr->first_empty};
    uint32_t my_nrec = 0;
    {
        const long long t = first_tile + lane;
        if (lane < kEmitTilesPerWarp && t < n_tiles) {
            my_nrec = agg_records(tile_state[t]);
            if (my_nrec == kEvByWriter) { my_nrec = 0u; }
        }
        if (my_nrec != 0u && my_nrec <= (uint32_t)kEmitSerial) {
            const ulonglong2 ex = tile_excl[t];
            long long nbase = (long long)(ex.x & ((1ull << 40) - 1));
            const long long tileK = (long long)ex.y;
            uint32_t cKind = (uint32_t)(ex.x >> 60) & 3u, cErr = (uint32_t)(ex.x >> 59) & 1u;
            uint4 rec[kEmitSerial];
#pragma unroll
            for (int q = 0; q < kEmitSerial; q++) {
                rec[q] = make_uint4(0xFFFFFFFFu, 0u, 0u, 0xFFFFu); // (sorts behind every record)
                if ((uint32_t)q < my_nrec) { rec[q] = tile_events[(size_t)t * kEvCap + q]; }
            }
            static_assert(kEmitSerial == 4, "the sorting network below is for four records");
            rec_cswap(rec[0], rec[1]); rec_cswap(rec[2], rec[3]); rec_cswap(rec[0], rec[2]); rec_cswap(rec[1], rec[3]); rec_cswap(rec[1], rec[2]);
#pragma unroll
            for (int q = 0; q < kEmitSerial; q++) {
                if ((uint32_t)q < my_nrec) {
                    const uint32_t evsc = rec[q].y, deler = rec[q].z, misc = rec[q].w;
                    const uint32_t ev = evsc & 0xFFFFu, sc = evsc >> 16, er = deler >> 16;
                    if ((ev | er) != 0u) {
                        const int64_t g0 = (int64_t)t * kTileBytes + (int64_t)rec[q].x * 16;
                        emit_cold(evsc, deler, misc, g0, (int64_t)nbase, (int64_t)(tileK + (long long)rec[q].x * 16), cKind, cErr, sink);
                    }
                    uint32_t lk = HEVCB_KIND_PASS, le = (er != 0u); // the chunk's summary for the ordered carry
                    if (ev != 0u) {
                        const int tp = 31 - __clz((int)ev);
                        lk = ((sc >> tp) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
                        le = ((er >> tp) >> 1) != 0u;
                    }
                    hevcb_carry_combine(cKind, cErr, lk, le);
                    nbase += (long long)__popc(sc);
                }
            }
        }
    }
    // tiles with more records: one at a time, one lane per record
    uint32_t todo = __ballot_sync(0xFFFFFFFFu, my_nrec > (uint32_t)kEmitSerial);
    while (todo != 0u) {
        const int src = __ffs((int)todo) - 1;
        todo &= todo - 1u;
        const long long t = first_tile + src;
        const uint32_t nrec = __shfl_sync(0xFFFFFFFFu, my_nrec, src);
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
                const int cnt = (int)(nrec - (uint32_t)(p * 32) < 32u ? nrec - (uint32_t)(p * 32) : 32u);
#pragma unroll 4
                for (int i = 0; i < cnt; i++) {
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
        __syncwarp(); // sorted[warp] is rewritten for the warp's next tile
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
        hdr->next_tile = 0ull;
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
        // one CTA per SM (the ring takes most of the SM's shared memory); cooperative launch: writers wait on the scanner, the
        // scanner on the analysers of every CTA, so every CTA must be resident
        long long workers = (long long)ctx->sm_count * ctx->scan_blocks_per_sm - 1; // one SM runs the scanner CTA
        if (workers > n_tiles) { workers = n_tiles; }
        if (workers < 1) { HEVCB_SET_ERR(ctx, "scan kernel needs at least two SMs"); return HEVCB_E_CUDA; }
        const long long grid = workers + 1;
        long long nt = n_tiles;
        long long dbg = ctx->scan_debug_flags;
        ScanGeom g = geom;
        // A scanner batch needs the aggregates of 32 x scan_per_lane consecutive tiles, and a tile is only written (its stage only
        // freed) once its batch is complete, so the rings must be able to hold the batch of the oldest unwritten tile entirely
        // (tiles are claimed in order, so the tiles the rings hold are a contiguous range of up to kStages x workers tiles that
        // starts at the oldest unwritten one: two batches fit whenever the buffer has more tiles than the rings hold.)  Small
        // batches keep the wait for a batch's last tile short: one tile per lane.
        long long per_lane = HEVCB_SCAN_BATCH_PER_LANE;
        while (per_lane > 1 && 64 * per_lane > (long long)kStages * workers && n_tiles > (long long)kStages * workers) { per_lane--; }
        if (64 * per_lane > (long long)kStages * workers && n_tiles > (long long)kStages * workers) {
            HEVCB_SET_ERR(ctx, "scan kernel: device has too few SMs for the tile ring");
            return HEVCB_E_CUDA;
        }
        if (const char* e = getenv("HEVCB_SCAN_PERLANE_RT")) { const long long v = atoll(e); if (v >= 1 && v <= kScanPerLane && (64 * v <= (long long)kStages * workers || n_tiles <= (long long)kStages * workers)) { per_lane = v; } }
        g.scan_per_lane = (int)per_lane;
        g.prefetch = 4 * workers; // ~19 MB ahead of the claims on a whole B200: far more than the rings hold, a fraction of L2 (126 MB)
        if (const char* e = getenv("HEVCB_SCAN_PREFETCH")) { g.prefetch = atoll(e); }
        // the first tiles have no claim that would prefetch them: one bulk prefetch over that range in front of the kernel would
        // only help the first microseconds, it is left out
        unsigned long long* tdbg = nullptr;
#ifdef HEVCB_SCAN_TIMING_BUILD
        if (getenv("HEVCB_SCAN_TIMING")) { // measurement aid: event stamps, dumped by tools/scan_timing.py through hevcb_scan_timing_dump
            if (hevcb_reserve(ctx, &ctx->scan_timing, (size_t)grid * kTimingIters * kTimingEvents * 8) != HEVCB_OK) { return HEVCB_E_NOMEM; }
            tdbg = reinterpret_cast<unsigned long long*>(ctx->scan_timing.p);
            HEVCB_CUDA(ctx, cudaMemsetAsync(tdbg, 0, (size_t)grid * kTimingIters * kTimingEvents * 8, stream));
            ctx->scan_timing_ctas = (int)grid;
        }
#endif
        void* args[] = {(void*)&d_buf, (void*)&g, (void*)&nt, (void*)&hdr, (void*)&states, (void*)&excl, (void*)&events, (void*)&d_nal_start, (void*)&d_nal_end,
                        (void*)&cap_nals, (void*)&d_rbsp, (void*)&d_rbsp_off, (void*)&d_rbsp_end, (void*)&dbg, (void*)&tdbg};
        HEVCB_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)hevcb_scan_strip_kernel, dim3((unsigned)grid), dim3(kThreads), args, smem, stream));
        hevcb_scan_emit_kernel<<<(unsigned)((n_tiles + kEmitTilesPerBlock - 1) / kEmitTilesPerBlock), 256, 0, stream>>>(n_tiles, hdr, states, excl, events, d_nal_start, d_nal_end, cap_nals,
                                                                                 d_rbsp_off, d_rbsp_end);
        ctx->launches += 2;
        HEVCB_CUDA(ctx, cudaGetLastError());
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
    geom.size = size; geom.own = size; geom.evl = size - HEVCB_TAIL_ZONE; geom.init_n = 0; geom.init_kind = HEVCB_KIND_Z3; geom.scan_per_lane = kScanPerLane; geom.prefetch = 0;
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
    geom.scan_per_lane = kScanPerLane;
    geom.prefetch = 0;
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

// measurement aid: copies the event stamps of the last launch made with HEVCB_SCAN_TIMING set (ctas x 256 tiles x 8 events, ns)
extern "C" HEVCB_API int64_t hevcb_scan_timing_dump(hevcb_ctx* ctx, unsigned long long* out, int64_t cap)
{
    if (!ctx || !ctx->scan_timing.p) { return 0; }
    const int64_t n = (int64_t)ctx->scan_timing_ctas * kTimingIters * kTimingEvents;
    if (out && cap >= n) { cudaMemcpy(out, ctx->scan_timing.p, (size_t)n * 8, cudaMemcpyDeviceToHost); }
    return n;
}
