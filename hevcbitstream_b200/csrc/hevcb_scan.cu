// hevcb_scan.cu -- fused Annex-B start-code scan + emulation-prevention strip for sm_100a.
//
// Replaces the reference's per-NAL loop  find_nal_unit (h264_nal.c:38-76) -> nal_to_rbsp
// (h264_nal.c:147-200, called at hevc_stream.c:165)  by ONE pass over the byte range:
//
//   * persistent CTAs (cooperative launch: all co-resident) walk 32 KiB tiles round-robin, tile t -> CTA t mod grid,
//     so that the predecessor tiles of a tile are always being processed at the same time by the neighbouring CTAs;
//   * each tile (+16 B halo on either side) is staged into shared memory with TMA bulk copies
//     (cp.async.bulk + mbarrier), double buffered so the next tile is in flight while this one is processed;
//   * every lane owns 16 bytes: a conservative "two adjacent zero bytes?" SWAR test sends the common
//     case down a fast path, exact predicate bit-masks (hevcb_chunk_analyze) are built otherwise;
//   * counts (start codes, kept bytes) and the ordered (last-event-kind, error) carry are combined with
//     warp ballots / redux at lane -> row -> tile level and across tiles with a single-pass decoupled
//     look-back over a 16-byte tile-state word;
//   * NAL offsets are written by the lanes that own the events (ordered compaction), the EPB-free image
//     is written as an aligned, funnel-shifted 16-byte-vector copy out of shared memory.
//
// HBM traffic: input read once, image written once, 32 B of metadata per NAL.  Tensor cores unused:
// nothing here is a contraction.
#include <cuda_runtime.h>
#include <stdint.h>

#include "hevcb_internal.h"
#include "hevcb_scan_core.h"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowBytes = 512;  // one warp-row: 32 lanes x 16 B
constexpr int kRowsPerWarp = 8; // rows a warp walks in order (its carries stay in registers)
constexpr int kRows = kWarps * kRowsPerWarp;
constexpr int kTileBytes = kRows * kRowBytes; // 32 KiB
constexpr int kLead = 16;                     // leading halo
constexpr int kStageBytes = kLead + kTileBytes + 16;
constexpr int kStages = 2;
constexpr int kOutBytes = kTileBytes + 32;
constexpr int kLookPerLane = 2; // predecessor tile states per lane; every warp takes a 64-tile slice: window = 512 tiles per round

struct WarpAgg {
    uint32_t n;     // start codes in the warp's rows
    uint32_t k;     // kept bytes
    uint32_t kind;  // ordered carry summary of the warp's rows
    uint32_t err;
    uint32_t del;   // some row has removed bytes or is partially valid
    uint32_t pad[3];
};

// ordered summary of one warp's slice of the look-back window (slice 0 is the nearest)
struct LookPart {
    unsigned long long n;  // start codes in the slice up to and including its nearest prefix tile
    unsigned long long k;  // kept bytes, same extent
    uint32_t kind;         // carry summary of that extent (PASS when it saw no event)
    uint32_t err;
    uint32_t has_prefix;   // the slice contains a tile whose inclusive prefix is published
    uint32_t pad;
};

// dynamic shared memory layout
struct __align__(16) SmemLayout {
    uint8_t stage[kStages][kStageBytes];
    uint8_t out[kOutBytes];
    unsigned long long mbar[kStages];
    WarpAgg wagg[kWarps];
    LookPart look[kWarps];
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
__device__ __forceinline__ ulonglong2 ld_state(const ulonglong2* p)
{
    ulonglong2 v;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(ulonglong2* p, ulonglong2 v)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y) : "memory");
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
    unsigned long long reserved;
    long long first_empty;
    unsigned long long pad[6];
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

__global__ void __launch_bounds__(kThreads, 2) hevcb_scan_strip_kernel(
    const uint8_t* __restrict__ buf, int64_t size, long long n_tiles, ScanHeader* __restrict__ hdr,
    ulonglong2* __restrict__ tile_state, int64_t* __restrict__ nal_start, int64_t* __restrict__ nal_end,
    int64_t cap_nals, uint8_t* __restrict__ rbsp, int64_t* __restrict__ rbsp_off, int64_t* __restrict__ rbsp_end)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    SmemLayout& sm = *reinterpret_cast<SmemLayout*>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&sm.mbar[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        for (int s = 0; s < kStages; s++) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < n_tiles) { issue_tile_load(sm.stage[s], &sm.mbar[s], buf, size, t); }
        }
    }
    __syncthreads();

    DevSink sink{nal_start, nal_end, rbsp_off, rbsp_end, cap_nals, &hdr->first_empty};
    uint32_t phase_bits = 0;
    int s = 0;
    for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const int64_t t0 = (int64_t)t * kTileBytes;
        uint8_t* st = sm.stage[s];
        while (!mbar_try_wait(&sm.mbar[s], (phase_bits >> s) & 1u)) {}
        phase_bits ^= (1u << s);

        // ---- boundary fix-ups: positions < 0 read as non-zero, positions >= size read as zero
        const int64_t valid_end = (int64_t)kLead + (size - t0); // smem offset of position `size`
        if (t == 0 || valid_end < kStageBytes) {
            if (t == 0 && tid < kLead) { st[tid] = 0xFF; }
            if (valid_end < kStageBytes) {
                for (int i = (int)valid_end + tid; i < kStageBytes; i += kThreads) { st[i] = 0; }
            }
            __syncthreads();
        }

        // ---- phase 1: per-lane masks; the warp walks its rows in order and keeps the carries in registers
        uint32_t m_evsc[kRowsPerWarp];  // ev | sc << 16
        uint32_t m_deler[kRowsPerWarp]; // del | err << 16
        uint32_t m_misc[kRowsPerWarp];  // valid | scb << 16
        uint32_t rowflags = 0;          // bit i: row has events/errors; bit 16+i: row has removed bytes / partial validity
        uint32_t wN = 0, wK = 0, wKind = HEVCB_KIND_PASS, wErr = 0;
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; i++) {
            const int r = warp * kRowsPerWarp + i;
            const int off = kLead + r * kRowBytes + lane * 16;
            const uint4 v = *reinterpret_cast<const uint4*>(st + off);
            uint32_t wp = __shfl_up_sync(0xFFFFFFFFu, v.w, 1);
            uint32_t wn = __shfl_down_sync(0xFFFFFFFFu, v.x, 1);
            if (lane == 0) { wp = *reinterpret_cast<const uint32_t*>(st + off - 4); }
            if (lane == 31) { wn = *reinterpret_cast<const uint32_t*>(st + off + 16); }
            const int64_t g0 = t0 + (int64_t)r * kRowBytes + lane * 16;
            hevcb_chunk_masks m;
            const bool slow = hevcb_maybe_zero_pair(wp, v.x, v.y, v.z, v.w, wn) != 0u;
            const int64_t rem = size - g0;
            if (slow) {
                m = hevcb_chunk_analyze(wp, v.x, v.y, v.z, v.w, wn, g0, size);
            } else {
                m.ev = m.sc = m.scb = m.del = m.err = 0u;
                m.valid = rem >= 16 ? 0xFFFFu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
            }
            m_evsc[i] = m.ev | (m.sc << 16);
            m_deler[i] = m.del | (m.err << 16);
            m_misc[i] = m.valid | (m.scb << 16);
            const uint32_t Ab = __ballot_sync(0xFFFFFFFFu, slow || rem < 16);
            if (Ab == 0u) { // whole row on the fast path: 512 kept bytes, nothing else
                wK += kRowBytes;
                continue;
            }
            uint32_t lk, le;
            hevcb_chunk_summary(m, lk, le);
            const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, m.ev != 0u);
            const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lk == HEVCB_KIND_SC3);
            const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, le != 0u);
            const uint32_t Xb = __ballot_sync(0xFFFFFFFFu, (m.ev | m.err) != 0u);
            const uint32_t Db = __ballot_sync(0xFFFFFFFFu, (m.del != 0u) || (m.valid != 0xFFFFu));
            wN += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(m.sc));
            wK += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(m.valid & ~m.del));
            uint32_t rk, re;
            warp_carry_total(Eb, Sb, Rb, rk, re);
            hevcb_carry_combine(wKind, wErr, rk, re);
            rowflags |= (Xb != 0u ? 1u : 0u) << i;
            rowflags |= (Db != 0u ? 1u : 0u) << (16 + i);
        }
        if (lane == 0) {
            WarpAgg a;
            a.n = wN; a.k = wK; a.kind = wKind; a.err = wErr; a.del = (rowflags >> 16) != 0u;
            a.pad[0] = a.pad[1] = a.pad[2] = 0;
            sm.wagg[warp] = a;
        }
        __syncthreads();

        // ---- cross-tile look-back: every warp examines a 64-tile slice of the window (512 tiles per round trip)
        uint32_t tile_n = 0, tile_k = 0, ak = HEVCB_KIND_PASS, ae = 0;
        bool compact = false;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            const WarpAgg a = sm.wagg[w];
            tile_n += a.n;
            tile_k += a.k;
            hevcb_carry_combine(ak, ae, a.kind, a.err);
            compact = compact || (a.del != 0u);
        }
        unsigned long long exN = 0, exK = 0;
        uint32_t exKind = HEVCB_KIND_Z3, exErr = 0;
        if (t > 0) {
            if (tid == 0) { st_state(&tile_state[t], pack_state(kStatusAgg, tile_n, tile_k, ak, ae)); }
            uint32_t kindF = HEVCB_KIND_PASS, errF = 0;
            long long base = t - 1;
            for (;;) {
                // lane-local ordered summary of its tiles (j = 0 is the nearer one)
                uint32_t aggN = 0, aggK = 0, lkind = HEVCB_KIND_PASS, lerr = 0;
                unsigned long long pN = 0, pK = 0;
                bool hasP = false;
                ulonglong2 sv[kLookPerLane];
                const long long first = base - (long long)((warp * 32 + lane) * kLookPerLane);
#pragma unroll
                for (int j = 0; j < kLookPerLane; j++) {
                    const long long idx = first - j;
                    if (idx >= 0) { sv[j] = ld_state(&tile_state[idx]); }
                    else { sv[j] = pack_state(kStatusPrefix, 0, 0, HEVCB_KIND_Z3, 0); } // before the stream
                }
#pragma unroll
                for (int j = 0; j < kLookPerLane; j++) {
                    const long long idx = first - j;
                    while ((sv[j].x >> 62) == 0ull) { sv[j] = ld_state(&tile_state[idx]); }
                    if (!hasP) {
                        const uint32_t skind = (uint32_t)(sv[j].x >> 60) & 3u;
                        const uint32_t serr = (uint32_t)(sv[j].x >> 59) & 1u;
                        if (lkind == HEVCB_KIND_PASS) { lerr |= serr; lkind = skind; }
                        if ((sv[j].x >> 62) == kStatusPrefix) {
                            hasP = true;
                            pN = sv[j].x & ((1ull << 40) - 1);
                            pK = sv[j].y;
                        } else {
                            aggN += (uint32_t)(sv[j].x & ((1ull << 40) - 1));
                            aggK += (uint32_t)sv[j].y;
                        }
                    }
                }
                // warp-level ordered summary of the slice
                const uint32_t pm = __ballot_sync(0xFFFFFFFFu, hasP);
                const int pl = pm ? (__ffs((int)pm) - 1) : 32;
                const bool act = lane <= pl;
                unsigned long long sN = __reduce_add_sync(0xFFFFFFFFu, act ? aggN : 0u);
                unsigned long long sK = __reduce_add_sync(0xFFFFFFFFu, act ? aggK : 0u);
                if (pl < 32) {
                    sN += __shfl_sync(0xFFFFFFFFu, pN, pl);
                    sK += __shfl_sync(0xFFFFFFFFu, pK, pl);
                }
                const uint32_t eb = __ballot_sync(0xFFFFFFFFu, act && lkind != HEVCB_KIND_PASS);
                const uint32_t rb = __ballot_sync(0xFFFFFFFFu, act && lerr != 0u);
                uint32_t skind = HEVCB_KIND_PASS, serr;
                if (eb) {
                    const int f = __ffs((int)eb) - 1; // nearest lane whose tiles saw an event
                    skind = __shfl_sync(0xFFFFFFFFu, lkind, f);
                    const uint32_t upto = (f == 31) ? 0xFFFFFFFFu : ((2u << f) - 1u);
                    serr = (rb & upto) != 0u;
                } else {
                    serr = rb != 0u;
                }
                if (lane == 0) {
                    LookPart lp;
                    lp.n = sN; lp.k = sK; lp.kind = skind; lp.err = serr; lp.has_prefix = (pl < 32) ? 1u : 0u; lp.pad = 0;
                    sm.look[warp] = lp;
                }
                __syncthreads();
                // every thread combines the slices nearest-first
                bool found = false;
#pragma unroll
                for (int w = 0; w < kWarps; w++) {
                    if (!found) {
                        const LookPart lp = sm.look[w];
                        exN += lp.n;
                        exK += lp.k;
                        if (kindF == HEVCB_KIND_PASS) { errF |= lp.err; kindF = lp.kind; }
                        found = lp.has_prefix != 0u;
                    }
                }
                if (found) { break; }
                base -= (long long)kThreads * kLookPerLane;
                __syncthreads(); // sm.look is rewritten by the next round
            }
            exKind = kindF;
            exErr = errF;
        }
        if (tid == 0) {
            uint32_t ik = exKind, ie = exErr;
            hevcb_carry_combine(ik, ie, ak, ae);
            st_state(&tile_state[t], pack_state(kStatusPrefix, exN + tile_n, exK + tile_k, ik, ie));
        }

        // ---- phase 2: ordered emission of NAL boundaries; compaction when bytes were removed
        const long long tileN = (long long)exN;
        const long long tileK = (long long)exK;
        // carry entering this warp = tile carry (+) aggregates of the warps before it
        uint32_t rN = 0, rK = 0, rKind = exKind, rErr = exErr;
        const uint32_t tileKept = tile_k;
#pragma unroll
        for (int w = 0; w < kWarps; w++) {
            if (w < warp) {
                const WarpAgg a = sm.wagg[w];
                rN += a.n; rK += a.k; hevcb_carry_combine(rKind, rErr, a.kind, a.err);
            }
        }
        const bool do_compact = compact && (rbsp != nullptr);
#pragma unroll
        for (int i = 0; i < kRowsPerWarp; i++) {
            const bool rowX = ((rowflags >> i) & 1u) != 0u;
            const bool rowD = ((rowflags >> (16 + i)) & 1u) != 0u;
            if (!rowX && !rowD && !do_compact) { rK += kRowBytes; continue; }
            const int r = warp * kRowsPerWarp + i;
            hevcb_chunk_masks m;
            m.ev = m_evsc[i] & 0xFFFFu;
            m.sc = m_evsc[i] >> 16;
            m.del = m_deler[i] & 0xFFFFu;
            m.err = m_deler[i] >> 16;
            m.valid = m_misc[i] & 0xFFFFu;
            m.scb = m_misc[i] >> 16;
            const uint32_t keep = m.valid & ~m.del;
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
                uint32_t lk, le;
                hevcb_chunk_summary(m, lk, le);
                const uint32_t Eb = __ballot_sync(0xFFFFFFFFu, m.ev != 0u);
                const uint32_t Sb = __ballot_sync(0xFFFFFFFFu, lk == HEVCB_KIND_SC3);
                const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, le != 0u);
                uint32_t ck, ce;
                warp_carry_in(Eb, Sb, Rb, lane, ck, ce);
                if (ck == HEVCB_KIND_PASS) { ck = rKind; ce |= rErr; } // inherit the carry entering the row
                const uint32_t c = (uint32_t)__popc(m.sc);
                const uint32_t ninc = warp_incl_scan(c, lane);
                if ((m.ev | m.err) != 0u) {
                    const int64_t g0 = t0 + (int64_t)r * kRowBytes + lane * 16;
                    hevcb_chunk_emit(m, g0, (int64_t)(tileN + rN + (ninc - c)), (int64_t)(tileK + klane), ck, ce, sink);
                }
                uint32_t rk, re;
                warp_carry_total(Eb, Sb, Rb, rk, re);
                hevcb_carry_combine(rKind, rErr, rk, re);
                rN += __shfl_sync(0xFFFFFFFFu, ninc, 31);
            }
            if (do_compact) {
                // byte-granular compaction of this lane's kept bytes into sm.out
                const int off = kLead + r * kRowBytes + lane * 16;
                const uint4 v = *reinterpret_cast<const uint4*>(st + off);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                uint32_t o = klane;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if ((keep >> j) & 1u) { sm.out[o++] = (uint8_t)(w[j >> 2] >> (8 * (j & 3))); }
                }
            }
            rK += rowKept;
        }

        // ---- copy-out: aligned 16-byte vectors, funnel-shifted by the (tile-uniform) misalignment
        if (rbsp != nullptr) {
            if (compact) { __syncthreads(); }
            const uint8_t* src = compact ? sm.out : (st + kLead);
            const uint32_t L = tileKept;
            uint8_t* dst = rbsp + tileK;
            const uint32_t head0 = (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
            const uint32_t head = head0 < L ? head0 : L;
            if ((uint32_t)tid < head) { dst[tid] = src[tid]; }
            const uint32_t nv = (L - head) >> 4;
            const uint32_t r16 = head & 15u; // source misalignment, uniform over the tile
            const uint32_t q = r16 >> 2;
            const uint32_t sh = (r16 & 3u) * 8u;
            for (uint32_t vi = tid; vi < nv; vi += kThreads) {
                const uint32_t so = head + (vi << 4);
                const uint32_t a = so & ~15u;
                const uint4 lo = *reinterpret_cast<const uint4*>(src + a);
                uint4 o4;
                if (r16 == 0u) {
                    o4 = lo;
                } else {
                    const uint4 hi = *reinterpret_cast<const uint4*>(src + a + 16);
                    const uint32_t W[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
                    uint32_t x[5];
#pragma unroll
                    for (int e = 0; e < 5; e++) {
                        // q is tile-uniform: select without dynamic register indexing
                        x[e] = (q == 0u) ? W[e] : (q == 1u) ? W[e + 1] : (q == 2u) ? W[e + 2] : W[(e + 3) & 7];
                    }
                    o4.x = __funnelshift_r(x[0], x[1], sh);
                    o4.y = __funnelshift_r(x[1], x[2], sh);
                    o4.z = __funnelshift_r(x[2], x[3], sh);
                    o4.w = __funnelshift_r(x[3], x[4], sh);
                }
                __stcs(reinterpret_cast<uint4*>(dst + so), o4);
            }
            const uint32_t done = head + (nv << 4);
            if ((uint32_t)tid < L - done) { dst[done + tid] = src[done + tid]; }
        }
        __syncthreads(); // every read of stage s (and of sm.out) is finished

        if (tid == 0) {
            const long long nt = t + (long long)kStages * gridDim.x; // the tile that reuses this stage
            if (nt < n_tiles) {
                fence_proxy_async();
                issue_tile_load(sm.stage[s], &sm.mbar[s], buf, size, nt);
            }
        }
        s = (s + 1 == kStages) ? 0 : s + 1;
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
        const ulonglong2 sv = tile_state[n_tiles - 1];
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

__global__ void hevcb_scan_init_kernel(ScanHeader* hdr)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        hdr->reserved = 0ull;
        hdr->first_empty = 0x7FFFFFFFFFFFFFFFll;
    }
}

} // namespace

int hevcb_launch_scan_strip(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, int64_t* d_nal_start, int64_t* d_nal_end,
                            int64_t cap_nals, uint8_t* d_rbsp, int64_t* d_rbsp_off, int64_t* d_rbsp_end,
                            hevcb_scan_summary* d_summary, cudaStream_t stream)
{
    if (size < 0 || cap_nals < 0 || !d_nal_start || !d_nal_end || !d_rbsp_off || !d_rbsp_end || !d_summary || (size > 0 && !d_buf)) {
        HEVCB_SET_ERR(ctx, "hevcb_scan_strip: invalid argument");
        return HEVCB_E_ARG;
    }
    if (((uintptr_t)d_buf & 15u) || ((uintptr_t)d_rbsp & 15u)) {
        HEVCB_SET_ERR(ctx, "hevcb_scan_strip: buf and rbsp must be 16-byte aligned");
        return HEVCB_E_ALIGN;
    }
    const long long n_tiles = (long long)((size + kTileBytes - 1) / kTileBytes);
    const size_t need = sizeof(ScanHeader) + (size_t)(n_tiles > 0 ? n_tiles : 1) * sizeof(ulonglong2);
    int rc = hevcb_reserve(ctx, &ctx->scan_scratch, need);
    if (rc != HEVCB_OK) { return rc; }
    ScanHeader* hdr = reinterpret_cast<ScanHeader*>(ctx->scan_scratch.p);
    ulonglong2* states = reinterpret_cast<ulonglong2*>(reinterpret_cast<uint8_t*>(ctx->scan_scratch.p) + sizeof(ScanHeader));

    HEVCB_CUDA(ctx, cudaMemsetAsync(ctx->scan_scratch.p, 0, need, stream));
    hevcb_scan_init_kernel<<<1, 32, 0, stream>>>(hdr);
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
        if (grid > n_tiles) { grid = n_tiles; }
        // cooperative launch: the chained look-back needs every CTA of the grid to be resident
        long long nt = n_tiles;
        void* args[] = {(void*)&d_buf, (void*)&size, (void*)&nt, (void*)&hdr, (void*)&states, (void*)&d_nal_start, (void*)&d_nal_end,
                        (void*)&cap_nals, (void*)&d_rbsp, (void*)&d_rbsp_off, (void*)&d_rbsp_end};
        HEVCB_CUDA(ctx, cudaLaunchCooperativeKernel((const void*)hevcb_scan_strip_kernel, dim3((unsigned)grid), dim3(kThreads), args, smem, stream));
        ctx->launches++;
    }
    hevcb_scan_finalize_kernel<<<1, 32, 0, stream>>>(d_buf, size, n_tiles, hdr, states, d_nal_start, d_nal_end, d_rbsp_off, d_rbsp_end,
                                                     cap_nals, d_summary);
    ctx->launches++;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}
