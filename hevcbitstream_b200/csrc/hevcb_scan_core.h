// hevcb_scan_core.h -- per-chunk predicates and the sequential tail rules of the Annex-B scan.
//
// Pure functions shared by the sm_100a kernels (hevcb_scan.cu) and by the CPU-side *test* build
// (tests/hostsim) that checks this logic against the oracle without a GPU.  The product library
// only ever calls them from device code.
//
// Reference semantics restated here (file:line in /root/reference):
//   find_nal_unit        h264_nal.c:38-76   start = after first 00 00 01, end = first 00 00 00 | 00 00 01
//   nal_to_rbsp          h264_nal.c:147-200 drop 03 after 00 00; -1 on 00 00 0{0,1,2} and on 00 00 03 xx>3
//
// Stream model used by the kernels ("events"):
//   an EVENT is a position c with b[c]==0 && b[c+1]==0 && b[c+2]<=1.  If b[c+2]==1 it is a start-code
//   event (SC3) and opens the NAL starting at c+3; every event that directly follows an SC3 event
//   closes that NAL (nal_end = c).  Events are only honoured for c < T = size-8; the last 8 bytes are
//   resolved by hevcb_scan_tail() with the reference's exact end-of-buffer rules.
//   Emulation prevention is context free: byte p is removed iff b[p]==3 && b[p-1]==0 && b[p-2]==0.
#pragma once
#include <stdint.h>

#include "../../include/hevcb.h"

#if defined(__CUDACC__)
#define HEVCB_HD __host__ __device__ __forceinline__
#else
#define HEVCB_HD static inline
#endif

#define HEVCB_TAIL_ZONE 8  // events at positions >= size - HEVCB_TAIL_ZONE are left to hevcb_scan_tail

// kinds carried between chunks / rows / tiles ("last event seen so far")
#define HEVCB_KIND_PASS 0u  // no event in the segment
#define HEVCB_KIND_SC3 1u   // last event opens a NAL
#define HEVCB_KIND_Z3 2u    // last event is 00 00 00 (or an unusable start code): outside any NAL

// exact per-byte "== 0" flags of a 32-bit word: 0x80 in every byte that is zero
HEVCB_HD uint32_t hevcb_zero_flags(uint32_t w)
{
    uint32_t t = (w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | w | 0x7F7F7F7Fu);
}
// gather the four 0x80 flags of a word into bits 0..3
HEVCB_HD uint32_t hevcb_gather4(uint32_t f) { return (((f >> 7) * 0x00204081u) >> 21) & 0xFu; }

// conservative zero-pair test: returns nonzero if the 20 bytes [g0-2, g0+18) MAY contain two adjacent
// zero bytes (never misses one).  wp = bytes g0-4..g0-1, wn = bytes g0+16..g0+19.
HEVCB_HD uint32_t hevcb_maybe_zero_pair(uint32_t wp, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wn)
{
    // (w - 0x01..) & ~w & 0x80..: superset of the zero bytes
    uint32_t zp = (wp - 0x01010101u) & ~wp & 0x80800000u; // only bytes g0-2, g0-1 matter
    uint32_t z0 = (w0 - 0x01010101u) & ~w0 & 0x80808080u;
    uint32_t z1 = (w1 - 0x01010101u) & ~w1 & 0x80808080u;
    uint32_t z2 = (w2 - 0x01010101u) & ~w2 & 0x80808080u;
    uint32_t z3 = (w3 - 0x01010101u) & ~w3 & 0x80808080u;
    uint32_t zn = (wn - 0x01010101u) & ~wn & 0x00008080u; // bytes g0+16, g0+17
    // flag of byte j+1 moved onto byte j: (hi:lo) >> 8
    uint32_t p = zp & ((zp >> 8) | (z0 << 24));
    p |= z0 & ((z0 >> 8) | (z1 << 24));
    p |= z1 & ((z1 >> 8) | (z2 << 24));
    p |= z2 & ((z2 >> 8) | (z3 << 24));
    p |= z3 & ((z3 >> 8) | (zn << 24));
    return p;
}

struct hevcb_chunk_masks {
    uint32_t ev;   // bit j: honoured event at g0+j            (j in 0..15)
    uint32_t sc;   // bit j: that event is a start code (SC3)
    uint32_t scb;  // bit t: honoured SC3 event at g0+t-3      (t in 0..2, i.e. positions -3,-2,-1)
    uint32_t del;  // bit j: byte g0+j is an emulation prevention byte (removed)
    uint32_t err;  // bit j: nal_to_rbsp error position, counts only if the byte is inside a NAL
    uint32_t valid;// bit j: g0+j < size
};

// Exact masks for the 16-byte chunk at global position g0.  wp = bytes g0-4..g0-1, wn = bytes
// g0+16..g0+19.  Bytes at positions < 0 must be presented as non-zero, bytes >= size as zero (the
// product's zero-padding rule).
//
// Three limits describe the byte range (a whole stream, or one shard of a byte-range partition):
//   size  bytes that exist as data (for a shard: the owned bytes + the halo that follows them)
//   own   bytes that are kept / reported by this pass (image, removal and error masks stop here)
//   evl   events and error positions are honoured below evl (whole stream and last shard: own - HEVCB_TAIL_ZONE, the rest
//         is resolved by hevcb_scan_tail; inner shard: own, the patterns read on into the halo)
// The three per-byte facts every predicate is built from, as 21-bit masks (bit (j+3) <-> position g0+j, j in [-3, 17]):
// LE: byte <= 3, B0 / B1: bit 0 / bit 1 of the byte.  Then  == 0: LE & ~B0 & ~B1,  == 1: LE & B0 & ~B1,  == 3: LE & B0 & B1.
struct hevcb_chunk_bits {
    uint32_t LE, B0, B1;
};
HEVCB_HD uint32_t hevcb_gather_bit0(uint32_t w) { return (((w & 0x01010101u) * 0x00204081u) >> 21) & 0xFu; }
HEVCB_HD hevcb_chunk_bits hevcb_chunk_classify(uint32_t wp, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wn)
{
    const uint32_t cfc = 0xFCFCFCFCu;
    hevcb_chunk_bits b;
    b.LE = (hevcb_gather4(hevcb_zero_flags(wp & cfc)) >> 1) | (hevcb_gather4(hevcb_zero_flags(w0 & cfc)) << 3) |
           (hevcb_gather4(hevcb_zero_flags(w1 & cfc)) << 7) | (hevcb_gather4(hevcb_zero_flags(w2 & cfc)) << 11) |
           (hevcb_gather4(hevcb_zero_flags(w3 & cfc)) << 15) | ((hevcb_gather4(hevcb_zero_flags(wn & cfc)) & 3u) << 19);
    b.B0 = (hevcb_gather_bit0(wp) >> 1) | (hevcb_gather_bit0(w0) << 3) | (hevcb_gather_bit0(w1) << 7) | (hevcb_gather_bit0(w2) << 11) |
           (hevcb_gather_bit0(w3) << 15) | ((hevcb_gather_bit0(wn) & 3u) << 19);
    b.B1 = (hevcb_gather_bit0(wp >> 1) >> 1) | (hevcb_gather_bit0(w0 >> 1) << 3) | (hevcb_gather_bit0(w1 >> 1) << 7) |
           (hevcb_gather_bit0(w2 >> 1) << 11) | (hevcb_gather_bit0(w3 >> 1) << 15) | ((hevcb_gather_bit0(wn >> 1) & 3u) << 19);
    return b;
}

// Exact masks of a chunk whose 16 bytes and the 3 bytes behind them are all owned, below the event limit and inside the data
// (every chunk of an interior tile): no position limits apply.
HEVCB_HD hevcb_chunk_masks hevcb_chunk_analyze_interior(uint32_t wp, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wn)
{
    hevcb_chunk_masks m;
    const hevcb_chunk_bits b = hevcb_chunk_classify(wp, w0, w1, w2, w3, wn);
    const uint32_t Z = b.LE & ~(b.B0 | b.B1), O = b.LE & b.B0 & ~b.B1, T3 = b.LE & b.B0 & b.B1;
    const uint32_t P = Z & (Z >> 1);                       // bit (j+3): b[j]==0 && b[j+1]==0
    const uint32_t EV = P & ((b.LE & ~b.B1) >> 2) & 0x7FFFFu; // third byte <= 1; events at j in [-3, 15]
    const uint32_t SC = P & (O >> 2) & 0x7FFFFu;
    const uint32_t PP = P << 2;                            // bit (j+3): b[j-2]==0 && b[j-1]==0
    const uint32_t DEL = T3 & PP;
    const uint32_t ERR1 = b.LE & ~T3 & PP & ~(EV << 2);    // third byte of an honoured event is not an error
    const uint32_t ERR2 = DEL & ((~b.LE) >> 1);            // EPB followed by > 3
    m.ev = (EV >> 3) & 0xFFFFu;
    m.sc = (SC >> 3) & 0xFFFFu;
    m.scb = SC & 7u;
    m.del = (DEL >> 3) & 0xFFFFu;
    m.err = ((ERR1 | ERR2) >> 3) & 0xFFFFu;
    m.valid = 0xFFFFu;
    return m;
}

HEVCB_HD hevcb_chunk_masks hevcb_chunk_analyze(uint32_t wp, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wn,
                                               int64_t g0, int64_t size, int64_t own, int64_t evl)
{
    if (own - g0 >= 16 && evl - g0 >= 16 && size - 1 - g0 >= 16) { return hevcb_chunk_analyze_interior(wp, w0, w1, w2, w3, wn); }
    hevcb_chunk_masks m;
    // 21-bit masks, bit (j+3) <-> position g0+j, j in [-3, 17]
    const hevcb_chunk_bits b = hevcb_chunk_classify(wp, w0, w1, w2, w3, wn);
    const uint32_t Z = b.LE & ~(b.B0 | b.B1), O = b.LE & b.B0 & ~b.B1;
    const uint32_t T3 = b.LE & b.B0 & b.B1 & 0x7FFF8u;    // positions 0..15
    const uint32_t LE3 = b.LE & 0xFFFF8u;                  // positions 0..16

    // position limits
    int64_t rem = own - g0;                      // positions j < rem are owned
    uint32_t valid = rem >= 16 ? 0xFFFFu : (rem <= 0 ? 0u : ((1u << (int)rem) - 1u));
    int64_t remT = evl - g0;                     // events honoured for j < remT (j may be -3..-1)
    uint32_t evlim = remT >= 16 ? 0x7FFFFu : (remT <= -3 ? 0u : ((1u << (int)(remT + 3)) - 1u));

    uint32_t P = Z & (Z >> 1);                   // bit (j+3): b[j]==0 && b[j+1]==0, j in [-3, 16]
    uint32_t EV = P & ((Z | O) >> 2) & evlim;    // events at j in [-3, 15]
    uint32_t SC = P & (O >> 2) & evlim;
    uint32_t PP = P << 2;                        // bit (j+3): b[j-2]==0 && b[j-1]==0
    uint32_t DEL = T3 & PP;                      // j in [0, 15]
    uint32_t LT3 = LE3 & ~T3;
    uint32_t ERR1 = LT3 & PP & ~(EV << 2);       // third byte of an honoured event is not an error
    uint32_t GT3n = (~LE3) >> 1;                 // bit (j+3): b[j+1] > 3
    int64_t remS = size - 1 - g0;                // positions j with j + 1 inside the data
    uint32_t v1 = remS >= 16 ? 0xFFFFu : (remS <= 0 ? 0u : ((1u << (int)remS) - 1u));
    uint32_t ERR2 = DEL & GT3n & (v1 << 3);      // EPB followed by > 3, only when that byte exists

    m.ev = (EV >> 3) & 0xFFFFu;
    m.sc = (SC >> 3) & 0xFFFFu;
    m.scb = SC & 7u;
    m.del = (DEL >> 3) & valid;
    // error positions >= T belong to hevcb_scan_tail (the NAL they fall into is only known there)
    uint32_t errlim = remT >= 16 ? 0xFFFFu : (remT <= 0 ? 0u : ((1u << (int)remT) - 1u));
    m.err = ((ERR1 | ERR2) >> 3) & valid & errlim;
    m.valid = valid;
    return m;
}

// whole stream: everything is owned, the last HEVCB_TAIL_ZONE bytes are left to hevcb_scan_tail
HEVCB_HD hevcb_chunk_masks hevcb_chunk_analyze(uint32_t wp, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t wn,
                                               int64_t g0, int64_t size)
{
    return hevcb_chunk_analyze(wp, w0, w1, w2, w3, wn, g0, size, size, size - HEVCB_TAIL_ZONE);
}

HEVCB_HD int hevcb_popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
HEVCB_HD int hevcb_ctz(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
HEVCB_HD int hevcb_top(uint32_t x) // index of highest set bit, x != 0
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}

// (kind, err) summary of one chunk for the ordered carry:
//   kind = PASS when the chunk has no event, else the kind of its last event;
//   err  = PASS: any error position in the chunk; else: any error position after the last event.
HEVCB_HD void hevcb_chunk_summary(const hevcb_chunk_masks& m, uint32_t& kind, uint32_t& err)
{
    if (m.ev == 0u) { kind = HEVCB_KIND_PASS; err = (m.err != 0u); return; }
    int t = hevcb_top(m.ev);
    kind = ((m.sc >> t) & 1u) ? HEVCB_KIND_SC3 : HEVCB_KIND_Z3;
    err = ((m.err >> t) >> 1) != 0u;
}

// ordered combine: state a followed by segment b
HEVCB_HD void hevcb_carry_combine(uint32_t& kind, uint32_t& err, uint32_t bkind, uint32_t berr)
{
    if (bkind != HEVCB_KIND_PASS) { kind = bkind; err = berr; }
    else { err |= berr; }
}

// Walk the events / error positions of one chunk in order and report NAL boundaries to `sink`:
//   sink.open(k, nal_start, rbsp_off)                  NAL k starts (k = running SC3 index)
//   sink.close(k, nal_end, rbsp_end /* -1 = nal_to_rbsp error */, empty)
// nbase = number of SC3 events before the chunk, kbase = kept bytes before the chunk,
// (kind_in, err_in) = ordered carry entering the chunk.
template <class Sink>
HEVCB_HD void hevcb_chunk_emit(const hevcb_chunk_masks& m, int64_t g0, int64_t nbase, int64_t kbase,
                               uint32_t kind_in, uint32_t err_in, Sink& sink)
{
    uint32_t todo = m.ev | m.err;
    bool open = (kind_in == HEVCB_KIND_SC3);
    uint32_t e = err_in;
    int64_t k = nbase;
    int prev_ev = -100;
    const uint32_t keep = m.valid & ~m.del;
    while (todo) {
        int j = hevcb_ctz(todo);
        todo &= todo - 1u;
        if ((m.ev >> j) & 1u) {
            int64_t pos = g0 + j;
            int64_t outpos = kbase + hevcb_popc(keep & ((1u << j) - 1u));
            if (open) {
                bool empty = (prev_ev >= 0) ? (prev_ev == j - 3) : (j < 3 && ((m.scb >> j) & 1u));
                sink.close(k - 1, pos, e ? (int64_t)-1 : outpos, empty);
            }
            if ((m.sc >> j) & 1u) { sink.open(k, pos + 3, outpos + 3); k++; open = true; e = 0u; }
            else { open = false; }
            prev_ev = j;
        } else if (open) {
            e = 1u;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Tail: sequential continuation of the reference loop over the last HEVCB_TAIL_ZONE bytes.
// ------------------------------------------------------------------------------------------------

struct hevcb_tail_in {
    int64_t size;
    int64_t n;          // NALs opened by honoured events (positions < T)
    uint32_t kind;      // kind of the last honoured event (HEVCB_KIND_Z3 when there was none)
    uint32_t err;       // nal_to_rbsp error seen so far inside the open NAL (kind == SC3 only)
    int64_t open_start; // nal_start[n-1] when kind == SC3
    int64_t prev_end;   // nal_end[n-1] when kind != SC3 and n > 0
    int64_t kept_total; // kept bytes over the whole buffer (= size - removed bytes)
};

struct hevcb_tail_nal {
    int64_t start, end, rbsp_off, rbsp_end; // rbsp_end = -1 on nal_to_rbsp error
};

struct hevcb_tail_out {
    int n_new;               // NALs completed or opened in the tail (records in nal[]), at most 4
    int closes_open;         // nal[0] completes the NAL that was open at T (its start/rbsp_off are already stored)
    hevcb_tail_nal nal[4];
    int32_t last_rc;         // terminating find_nal_unit return value (0 or -1)
    int32_t last_is_nal;     // 1 if the terminating call left an unterminated NAL (rc == -1) -> it is nal[n_new-1]
    int64_t last_start, last_end;
    int64_t first_empty;     // index (relative to nal[]) of a zero-length NAL that stops the loop, or -1
};

// byte fetch with the product's padding rule: positions >= size read as 0
template <typename Fetch>
HEVCB_HD uint32_t hevcb_tail_byte(const Fetch& f, int64_t pos, int64_t size)
{
    return (pos >= 0 && pos < size) ? (uint32_t)f(pos) : 0u;
}

// number of removed (EPB) bytes in [from, size)
template <typename Fetch>
HEVCB_HD int64_t hevcb_tail_count_del(const Fetch& f, int64_t from, int64_t size)
{
    int64_t c = 0;
    for (int64_t p = (from < 2 ? 2 : from); p < size; p++) {
        if (f(p) == 3 && f(p - 1) == 0 && f(p - 2) == 0) { c++; }
    }
    return c;
}

// nal_to_rbsp error scan over [from, end) of a NAL that starts at nal_start (h264_nal.c:153-168)
template <typename Fetch>
HEVCB_HD uint32_t hevcb_tail_err(const Fetch& f, int64_t nal_start, int64_t from, int64_t end)
{
    uint32_t e = 0;
    for (int64_t p = from; p < end; p++) {
        if (p - 2 >= nal_start && f(p - 1) == 0 && f(p - 2) == 0) {
            uint32_t b = f(p);
            if (b < 3) { e = 1; }
            if (b == 3 && p + 1 < end && f(p + 1) > 3) { e = 1; }
        }
    }
    return e;
}

// Continue the canonical `while (find_nal_unit(p, sz, &s, &e) > 0)` loop (hevc_analyze.c:135) from
// position T = max(0, size - 8) given the state produced by the honoured events.
template <typename Fetch>
HEVCB_HD void hevcb_scan_tail(const hevcb_tail_in& in, const Fetch& fetch, hevcb_tail_out& out)
{
    const int64_t size = in.size;
    const int64_t T = size > HEVCB_TAIL_ZONE ? size - HEVCB_TAIL_ZONE : 0;
    out.n_new = 0;
    out.closes_open = 0;
    out.first_empty = -1;
    out.last_rc = 0;
    out.last_is_nal = 0;
    out.last_start = 0;
    out.last_end = 0;

    bool open = (in.kind == HEVCB_KIND_SC3);
    int64_t s = open ? in.open_start : 0;        // start of the NAL being searched for its end
    int64_t p = (!open && in.n > 0) ? in.prev_end : 0; // origin of the current find_nal_unit call
    uint32_t err = open ? in.err : 0u;
    int64_t err_from = T;                        // bytes < T of the open NAL are covered by in.err
    bool first = true;

#define B(pos) hevcb_tail_byte(fetch, (pos), size)
    for (;;) {
        if (!open) {
            // ---- start search (h264_nal.c:46-62), origin p, continuing at i0 >= p
            // positions in [p, T-1) cannot match (they would have been honoured events)
            int64_t i = (first && T - 1 > p) ? T - 1 : p;
            bool found = false;
            for (;;) {
                bool sc3 = B(i) == 0 && B(i + 1) == 0 && B(i + 2) == 1;
                bool sc4 = B(i) == 0 && B(i + 1) == 0 && B(i + 2) == 0 && B(i + 3) == 1;
                if (sc3 || sc4) { found = true; break; }
                i++;
                if ((i - p) + 4 >= size - p) { break; } // did not find nal start
            }
            if (!found) {
                out.last_rc = 0; out.last_start = p; out.last_end = p;
                break;
            }
            if (!(B(i) == 0 && B(i + 1) == 0 && B(i + 2) == 1)) { i++; }
            s = i + 3;
            open = true;
            err = 0;
            err_from = s;
            // record the new NAL's start
            if (out.n_new < 4) {
                out.nal[out.n_new].start = s;
                out.nal[out.n_new].rbsp_off = in.kept_total - ((size - s) - hevcb_tail_count_del(fetch, s, size));
                out.nal[out.n_new].end = -1;
                out.nal[out.n_new].rbsp_end = -1;
            }
            out.n_new++;
            first = false;
            continue;
        }
        // ---- end search (h264_nal.c:64-72) for the NAL starting at s
        int64_t i = s;
        if (first && T > s) { i = T; }
        int64_t e = -1;
        for (;;) {
            bool endc = B(i) == 0 && B(i + 1) == 0 && B(i + 2) <= 1;
            // position s is always tested; later positions only after passing the i+3 < size check
            if (endc) { e = i; break; }
            i++;
            if (i + 3 >= size) { break; }
        }
        int idx;
        if (first) {
            // completing the NAL that was already open at T: it becomes nal[0] with start fields unused
            out.closes_open = 1;
            out.nal[0].start = s;
            out.nal[0].rbsp_off = -1;
            out.n_new = 1;
            idx = 0;
        } else {
            idx = out.n_new - 1;
        }
        first = false;
        if (e < 0) {
            // stream ended first: unterminated NAL [s, size), rc = -1
            int64_t from = err_from > s ? err_from : s;
            err |= hevcb_tail_err(fetch, s, from, size);
            if (idx < 4) {
                out.nal[idx].end = size;
                out.nal[idx].rbsp_end = err ? -1 : in.kept_total;
            }
            out.last_rc = -1; out.last_is_nal = 1; out.last_start = s; out.last_end = size;
            break;
        }
        {
            int64_t from = err_from > s ? err_from : s;
            err |= hevcb_tail_err(fetch, s, from, e);
            if (idx < 4) {
                out.nal[idx].end = e;
                out.nal[idx].rbsp_end = err ? -1 : in.kept_total - ((size - e) - hevcb_tail_count_del(fetch, e, size));
            }
        }
        if (e == s) {
            // zero-length NAL: find_nal_unit returns 0 and the loop stops
            out.first_empty = idx;
            out.last_rc = 0; out.last_start = s; out.last_end = s;
            break;
        }
        open = false;
        p = e;
    }
#undef B
}

// ------------------------------------------------------------------------------------------------
// Finalize: merge the honoured-event result with the tail and produce the summary.
// ------------------------------------------------------------------------------------------------

struct hevcb_scan_summary_core {
    int64_t n_nals;        // NAL units a reference reader visits: terminated ones + unterminated last NAL
    int64_t n_terminated;  // NALs for which find_nal_unit returned > 0
    int32_t last_rc;       // return value of the call that ended the loop: 0 or -1
    int32_t overflow;      // 1 when more NALs were found than the output arrays can hold
    int64_t last_start;    // *nal_start / *nal_end of that last call, as absolute offsets
    int64_t last_end;
    int64_t rbsp_bytes;    // bytes of the EPB-free image (= size - n_epb)
    int64_t n_epb;         // emulation prevention bytes removed over the whole buffer
};

template <typename Fetch>
HEVCB_HD void hevcb_scan_finalize(int64_t size, int64_t n_main, uint32_t kind, uint32_t err, int64_t kept_total,
                                  int64_t first_empty_main, const Fetch& fetch,
                                  int64_t* nal_start, int64_t* nal_end, int64_t* rbsp_off, int64_t* rbsp_end, int64_t cap,
                                  hevcb_scan_summary_core* out)
{
    out->rbsp_bytes = kept_total;
    out->n_epb = size - kept_total;
    out->overflow = 0;
    if (first_empty_main >= 0 && first_empty_main < n_main) {
        // the reference loop stopped at a zero-length NAL found among the honoured events
        out->n_nals = first_empty_main;
        out->n_terminated = first_empty_main;
        out->last_rc = 0;
        int64_t s = (first_empty_main < cap) ? nal_start[first_empty_main] : -1;
        out->last_start = s;
        out->last_end = s;
        if (first_empty_main > cap) { out->overflow = 1; }
        return;
    }
    if (n_main > cap) {
        out->overflow = 1;
        out->n_nals = n_main; out->n_terminated = n_main; out->last_rc = 0; out->last_start = -1; out->last_end = -1;
        return;
    }
    hevcb_tail_in in;
    in.size = size;
    in.n = n_main;
    in.kind = kind;
    in.err = err;
    in.open_start = (kind == HEVCB_KIND_SC3 && n_main > 0) ? nal_start[n_main - 1] : 0;
    in.prev_end = (kind != HEVCB_KIND_SC3 && n_main > 0) ? nal_end[n_main - 1] : 0;
    in.kept_total = kept_total;
    hevcb_tail_out t;
    hevcb_scan_tail(in, fetch, t);
    int64_t base = n_main - (t.closes_open ? 1 : 0);
    for (int i = 0; i < t.n_new && i < 4; i++) {
        int64_t idx = base + i;
        if (idx >= cap) { out->overflow = 1; continue; }
        if (!(t.closes_open && i == 0)) { nal_start[idx] = t.nal[i].start; rbsp_off[idx] = t.nal[i].rbsp_off; }
        nal_end[idx] = t.nal[i].end;
        rbsp_end[idx] = t.nal[i].rbsp_end;
    }
    int64_t total = base + t.n_new;
    if (t.first_empty >= 0) { total = base + t.first_empty; }
    out->n_nals = total;
    out->n_terminated = total - (t.last_is_nal ? 1 : 0);
    out->last_rc = t.last_rc;
    out->last_start = t.last_start;
    out->last_end = t.last_end;
}

// ------------------------------------------------------------------------------------------------
// Shard epilogue: the record one shard of a byte-range partition contributes to hevcb_stitch.  No end-of-buffer rules
// here; (N, kind, err, K) is the state after the honoured events of the shard.
// ------------------------------------------------------------------------------------------------
template <typename Fetch>
HEVCB_HD void hevcb_shard_finalize(int64_t own, int64_t N, uint32_t kind, uint32_t err, int64_t K, int64_t first_empty, const Fetch& fetch,
                                   const int64_t* nal_start, const int64_t* nal_end, const int64_t* rbsp_off, const int64_t* rbsp_end,
                                   int64_t cap, int is_first, int is_last, hevcb_shard_summary* out)
{
    int64_t fe = first_empty;
    if (fe < 0 || fe >= N) { fe = -1; }
    out->own = own;
    out->n_nals = N;
    out->first_empty = fe;
    out->first_empty_start = (fe >= 0 && fe < cap) ? nal_start[fe] : -1;
    out->rbsp_bytes = K;
    out->n_epb = own - K;
    out->is_first = is_first;
    out->is_last = is_last;
    out->open_at_end = (kind == HEVCB_KIND_SC3) ? 1 : 0;
    out->open_err = (int32_t)err;
    out->overflow = N > cap ? 1 : 0;
    out->pad = 0;
    const bool have0 = N > 0 && cap > 0;
    const bool closed0 = have0 && (N > 1 || kind != HEVCB_KIND_SC3);
    out->head_end = closed0 ? nal_end[0] : -1;
    out->head_rbsp_end = closed0 ? rbsp_end[0] : -1;
    for (int i = 0; i < 3; i++) {
        const int64_t pos = out->head_end - 3 + i;
        out->head_last3[i] = (closed0 && pos >= 0 && pos < own) ? (uint8_t)fetch(pos) : (uint8_t)0xFF;
    }
    const bool havel = N > 0 && N <= cap;
    out->last_nal_start = havel ? nal_start[N - 1] : -1;
    out->last_nal_end = (havel && kind != HEVCB_KIND_SC3) ? nal_end[N - 1] : -1;
    out->last_rbsp_off = havel ? rbsp_off[N - 1] : -1;
    const int64_t nt = own < 32 ? own : 32;
    out->tail_len = (int32_t)nt;
    for (int i = 0; i < 32; i++) { out->tail[i] = (i < nt) ? (uint8_t)fetch(own - nt + i) : (uint8_t)0; }
}
