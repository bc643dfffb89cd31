// hevcb_insert.cu -- batched emulation-prevention insertion for sm_100a: rbsp_to_nal (h264_nal.c:92-132) for every
// NAL of a batch, optionally framing the result as Annex-B (start code before every NAL).
//
// rbsp_to_nal is a 3-state machine (count of zero bytes since the last non-zero byte or the last insertion): before a
// byte <= 3 that arrives with count == 2 it emits 03 and restarts the count.  After a run of m zero bytes the state is
// 0 / 1 / 2 for m = 0 / odd / even, so the only thing that crosses a 16-byte chunk is the length of the zero run that
// precedes it; that is a "distance to the last non-zero byte", resolved inside a warp row with one ballot and carried
// between rows in a register.  One warp walks one NAL (512 bytes per step, 16 per lane); NALs are distributed over the
// warps of a persistent grid.  Two passes: count insertions per NAL -> exclusive scan of the output sizes -> write.
// Rows without insertions are written as aligned 16-byte vectors assembled with shuffles + funnel shifts, rows with
// insertions byte by byte.
//
// The same kernels assemble the output of the header rewrite (hevcb_rewrite_device): a NAL is then made of up to three
// parts -- bytes copied verbatim (what precedes the NAL in the input, or a NAL that is passed through), the rewritten
// header (escaped) and the original payload (escaped, the zero-run state continuing across the junction).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "hevcb_internal.h"

namespace {

constexpr int kInsThreads = 256;
constexpr int kInsWarps = kInsThreads / 32;

__device__ __forceinline__ uint32_t zero_flags(uint32_t w)
{
    const uint32_t t = (w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | w | 0x7F7F7F7Fu);
}
__device__ __forceinline__ uint32_t gather4(uint32_t f) { return (((f >> 7) * 0x00204081u) >> 21) & 0xFu; }
__device__ __forceinline__ uint32_t zero_mask16(const uint4& v)
{
    return gather4(zero_flags(v.x)) | (gather4(zero_flags(v.y)) << 4) | (gather4(zero_flags(v.z)) << 8) | (gather4(zero_flags(v.w)) << 12);
}
__device__ __forceinline__ uint32_t byte_of(const uint4& v, int j)
{
    const uint32_t w = (j < 4) ? v.x : (j < 8) ? v.y : (j < 12) ? v.z : v.w;
    return (w >> (8 * (j & 3))) & 0xFFu;
}
__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) { v += o; }
    }
    return v;
}

struct RowInfo {
    uint4 v;        // this lane's 16 bytes (aligned chunk of the image)
    uint32_t valid; // bit j: byte belongs to the NAL
    uint32_t ins;   // bit j: 03 is inserted before byte j
    bool fast;      // (warp uniform) interior row that cannot take an insertion: every byte valid, nothing inserted
};

// One 512-byte row of a part: loads, zero-run carry, insertion mask.  run_m = length of the zero run that ends right in
// front of the row's first byte of the part (warp uniform), updated to the run that reaches the end of the row.
// Rows whose loads are issued together (template parameter kInsAhead of the part walkers): one row per round trip leaves the
// warp latency bound on long parts (16 KiB NALs: 1073 -> 1317 GB/s with four), short parts keep the short loop.
__device__ __forceinline__ uint4 load_row_chunk(const uint8_t* __restrict__ img, int64_t row, int64_t off, int64_t end, int lane)
{
    const int64_t cpos = row + lane * 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (cpos < end && cpos + 16 > off) { v = *reinterpret_cast<const uint4*>(img + cpos); }
    return v;
}
template <typename I>
__device__ __forceinline__ RowInfo insert_row(const uint4 v, I row, I off, I end, int lane, uint32_t& run_m)
{
    RowInfo r;
    const I cpos = row + lane * 16;
    r.v = v;
    if (row >= off && row + 512 <= end) {
        // Interior row, the common case: no insertion is possible when the row holds no two adjacent zero bytes and the run
        // that enters it cannot complete one with the row's first byte.  One SWAR test + one vote instead of the exact masks.
        uint32_t nx = __shfl_down_sync(0xFFFFFFFFu, r.v.x, 1);
        if (lane == 31) { nx = 0xFFFFFFFFu; } // a pair that straddles the row end is seen by the next row through run_m
        const uint32_t m0 = r.v.x | __funnelshift_r(r.v.x, r.v.y, 8), m1 = r.v.y | __funnelshift_r(r.v.y, r.v.z, 8);
        const uint32_t m2 = r.v.z | __funnelshift_r(r.v.z, r.v.w, 8), m3 = r.v.w | __funnelshift_r(r.v.w, nx, 8);
        const uint32_t c = 0x01010101u, h = 0x80808080u;
        const uint32_t pair = (((m0 - c) & ~m0) | ((m1 - c) & ~m1) | ((m2 - c) & ~m2) | ((m3 - c) & ~m3)) & h; // exact: some byte of m is zero
        const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, r.v.x, 0) & 0xFFu, bl = __shfl_sync(0xFFFFFFFFu, r.v.w, 31) >> 24;
        const bool entering = (run_m >= 1u && b0 == 0u) || (run_m >= 2u && b0 <= 3u);
        if (!__any_sync(0xFFFFFFFFu, pair != 0u) && !entering) {
            r.valid = 0xFFFFu;
            r.ins = 0u;
            r.fast = true;
            run_m = (bl == 0u) ? 1u : 0u;
            return r;
        }
    }
    // validity of the 16 positions
    const I lo = off - cpos, hi = end - cpos; // valid j in [lo, hi)
    uint32_t valid = 0xFFFFu;
    if (lo > 0) { valid &= (lo >= 16) ? 0u : (0xFFFFu << (int)lo); }
    if (hi < 16) { valid &= (hi <= 0) ? 0u : ((1u << (int)hi) - 1u); }
    valid &= 0xFFFFu;
    r.valid = valid;
    const uint32_t lead = (off > row) ? (uint32_t)(off - row) : 0u; // bytes of the row in front of the part (first row only, < 16)
    const uint32_t Z = zero_mask16(r.v) & valid;
    const uint32_t R = valid & ~Z; // bytes that end a zero run
    const uint32_t tz = R ? (uint32_t)(15 - (31 - __clz((int)R))) : 16u; // bytes after the last run-ending byte of the chunk
    const uint32_t Rb = __ballot_sync(0xFFFFFFFFu, R != 0u);
    const uint32_t lower = Rb & ((1u << lane) - 1u);
    uint32_t m_in;
    {
        const int p = lower ? (31 - __clz((int)lower)) : 0;
        const uint32_t tzp = __shfl_sync(0xFFFFFFFFu, tz, p);
        m_in = lower ? (tzp + 16u * (uint32_t)(lane - p - 1)) : (run_m + 16u * (uint32_t)lane - (lane ? lead : 0u));
    }
    // next row's carry (bytes behind `end` are counted here; the caller takes them off after the last row)
    {
        const int p = Rb ? (31 - __clz((int)Rb)) : 0;
        const uint32_t tzp = __shfl_sync(0xFFFFFFFFu, tz, p);
        run_m = Rb ? (tzp + 16u * (uint32_t)(31 - p)) : (run_m + 512u - lead);
    }
    // state machine only where an insertion is possible: two zeros in a row somewhere, or a run reaching into the chunk
    uint32_t ins = 0;
    const bool need = ((Z & (Z >> 1)) != 0u) || (m_in >= 2u) || (m_in >= 1u && (Z & (valid & (0u - valid))) != 0u);
    if (need && valid) {
        uint32_t count = (m_in == 0u) ? 0u : ((m_in & 1u) ? 1u : 2u);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if ((valid >> j) & 1u) {
                const uint32_t bv = byte_of(r.v, j);
                if (count == 2u && bv <= 3u) { ins |= 1u << j; count = 0u; }
                count = (bv == 0u) ? count + 1u : 0u;
            }
        }
    }
    r.ins = ins;
    r.fast = false;
    return r;
}

// the parts of the NALs to assemble; a part is absent when its offset array is null or off >= end
struct AssembleParts {
    const uint8_t* raw_base; const int64_t* raw_off; const int64_t* raw_end; // copied verbatim
    const uint8_t* a_base; const int64_t* a_off; const int64_t* a_end;       // escaped
    const uint8_t* b_base; const int64_t* b_off; const int64_t* b_end;       // escaped, continues the state of A
    int sc_len;      // start code written in front of every NAL (0: none)
    int skip_neg_b;  // hevcb_insert semantics: b_end < 0 => the NAL emits nothing at all
    int len_size;    // 1 / 2 / 4: a big-endian length of the NAL's bytes is written in front of it (length-prefixed framing); 0: none
};

// one warp copies `len` bytes between arbitrarily aligned global addresses: 16-byte stores, the source taken as two aligned
// 16-byte loads funnel-shifted by the (copy-uniform) misalignment between source and destination
__device__ __forceinline__ void warp_copy_bytes(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int64_t len, int lane)
{
    if (len <= 0) { return; }
    int64_t head = (int64_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
    if (head > len) { head = len; }
    if (lane < head) { dst[lane] = src[lane]; }
    const int64_t nv = (len - head) >> 4;
    const uint8_t* s0 = src + head;
    const uint32_t mis = (uint32_t)((uintptr_t)s0 & 15u), q = mis >> 2, sh = (mis & 3u) * 8u;
    const uint4* sa = reinterpret_cast<const uint4*>(s0 - mis);
    uint4* da = reinterpret_cast<uint4*>(dst + head);
    for (int64_t i = lane; i < nv; i += 32) {
        const uint4 lo = __ldg(sa + i);
        uint4 o4 = lo;
        if (mis != 0u) {
            const uint4 hi = __ldg(sa + i + 1); // stays inside the 16-byte block of the copy's last source byte
            const uint32_t W[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
            uint32_t x[5];
#pragma unroll
            for (int e = 0; e < 5; e++) { x[e] = (q == 0u) ? W[e] : (q == 1u) ? W[e + 1] : (q == 2u) ? W[e + 2] : W[e + 3]; }
            o4.x = __funnelshift_r(x[0], x[1], sh); o4.y = __funnelshift_r(x[1], x[2], sh);
            o4.z = __funnelshift_r(x[2], x[3], sh); o4.w = __funnelshift_r(x[3], x[4], sh);
        }
        da[i] = o4;
    }
    const int64_t done = head + (nv << 4);
    if (lane < len - done) { dst[done + lane] = src[done + lane]; }
}

// insertions of one escaped part (count pass)
template <int kInsAhead>
__device__ __forceinline__ uint32_t count_part_n(const uint8_t* __restrict__ base, int64_t off, int64_t end, int lane, uint32_t& run_m)
{
    uint32_t total = 0;
    int64_t row = off & ~(int64_t)15;
    for (; row < end; row += 512 * kInsAhead) {
        uint4 v[kInsAhead];
#pragma unroll
        for (int k = 0; k < kInsAhead; k++) { v[k] = load_row_chunk(base, row + 512 * k, off, end, lane); }
#pragma unroll
        for (int k = 0; k < kInsAhead; k++) {
            if (row + 512 * k < end) {
                const RowInfo r = insert_row(v[k], row + 512 * k, off, end, lane, run_m);
                if (!r.fast) { total += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(r.ins)); }
            }
        }
    }
    {
        const int64_t row0 = off & ~(int64_t)15;
        const int64_t after = row0 + ((end - row0 + 511) / 512) * 512; // end of the last row
        run_m -= (uint32_t)(after - end); // bytes of the last row behind the part
    }
    return total;
}

__device__ __forceinline__ uint32_t count_part(const uint8_t* __restrict__ base, int64_t off, int64_t end, int lane, uint32_t& run_m)
{
    if (end <= off) { return 0u; }
    const int64_t len = end - off;
    if (len > 2048) { return count_part_n<4>(base, off, end, lane, run_m); }
    if (len > 512) { return count_part_n<2>(base, off, end, lane, run_m); }
    return count_part_n<1>(base, off, end, lane, run_m);
}

// pass 1: output size of every NAL (start code + verbatim bytes + escaped bytes)
// ---- long escaped parts are cut into pieces that are walked by different warps --------------------------------------------
// rbsp_to_nal's state is 0 behind every non-zero byte, so a part B of at least kSplitMin bytes is cut at positions whose
// preceding byte is non-zero (searched forward from every nominal cut; a cut that finds none within kSplitSearch bytes is
// dropped): the pieces are independent sub-problems.  The NAL's own entry keeps the start code, the verbatim bytes and part
// A; the pieces of B go to a side list ("extras") that a second pair of launches walks, one warp per piece.
constexpr int64_t kSplitSeg = 32 << 10, kSplitMin = 64 << 10, kSplitSearch = 4096;
constexpr int kSplitMaxPieces = 4096;
struct SplitList {
    uint8_t* flag;        // per NAL: 1 = part B is walked through the side list
    int64_t* e_off;       // per piece: extent of the piece in part B's source
    int64_t* e_end;
    int64_t* e_k;         // NAL the piece belongs to
    int32_t* e_j;         // index of the piece inside its NAL
    int32_t* e_np;        // pieces of that NAL
    int64_t* e_size;      // output bytes of the piece (count pass)
    int64_t* e_prefix;    // output offset of the piece inside its NAL
    unsigned long long* n_extras;
    int64_t cap;
};
__device__ __forceinline__ int64_t split_pieces(int64_t boff, int64_t bend, int64_t& seg)
{
    seg = kSplitSeg;
    if (bend < 0 || bend - boff < kSplitMin) { return 1; }
    const int64_t len = bend - boff;
    int64_t np = len / seg;
    if (np > kSplitMaxPieces) {
        seg = ((len / kSplitMaxPieces) + 511) & ~(int64_t)511;
        np = len / seg;
    }
    return np;
}
__global__ void __launch_bounds__(256) split_kernel(const AssembleParts P, int64_t n, SplitList L)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !P.b_off) { return; }
    const int64_t boff = P.b_off[k], bend = P.b_end[k];
    int64_t seg;
    const int64_t np = split_pieces(boff, bend, seg);
    if (np < 2) { return; }
    const int64_t base = (int64_t)atomicAdd(L.n_extras, (unsigned long long)np);
    if (base + np > L.cap) { // no room in the side list: the NAL stays with its own warp; the slots it was handed are marked unused
        for (int64_t j = 0; base + j < L.cap && j < np; j++) { L.e_k[base + j] = -1; }
        return;
    }
    L.flag[k] = 1;
    int64_t cut = boff;
    for (int64_t j = 0; j < np; j++) {
        int64_t next = bend;
        if (j + 1 < np) {
            int64_t q = boff + (j + 1) * seg;
            const int64_t lim = q + kSplitSearch;
            while (q < lim && P.b_base[q - 1] == 0) { q++; }
            next = (q < lim) ? q : cut; // no usable cut: this piece is empty, the next one starts where this one did
        }
        L.e_off[base + j] = cut;
        L.e_end[base + j] = next;
        L.e_k[base + j] = k;
        L.e_j[base + j] = (int32_t)j;
        L.e_np[base + j] = (int32_t)np;
        cut = next;
    }
}
// count pass over the pieces: output bytes of every piece (piece 0 continues the state part A leaves behind)
__global__ void __launch_bounds__(kInsThreads) extras_count_kernel(const AssembleParts P, SplitList L, unsigned long long* __restrict__ n_ins_total)
{
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * kInsWarps + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * kInsWarps;
    int64_t ne = (int64_t)*L.n_extras;
    if (ne > L.cap) { ne = L.cap; }
    unsigned long long local_ins = 0;
    for (int64_t e = wid; e < ne; e += nw) {
        const int64_t k = L.e_k[e];
        if (k < 0) { continue; } // slot of a NAL that found no room in the list
        uint32_t run_m = 0;
        if (L.e_j[e] == 0 && P.a_off) { (void)count_part(P.a_base, P.a_off[k], P.a_end[k], lane, run_m); }
        const int64_t off = L.e_off[e], end = L.e_end[e];
        const uint32_t ins = count_part(P.b_base, off, end, lane, run_m);
        if (lane == 0) { L.e_size[e] = (end > off ? end - off : 0) + (int64_t)ins; }
        local_ins += ins;
    }
    if (lane == 0 && local_ins) { atomicAdd(n_ins_total, local_ins); }
}
// output offsets of the pieces inside their NAL; the NAL's size becomes head + all pieces
__global__ void __launch_bounds__(256) extras_prefix_kernel(SplitList L, int64_t* __restrict__ out_size)
{
    int64_t ne = (int64_t)*L.n_extras;
    if (ne > L.cap) { ne = L.cap; }
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = L.e_k[e];
        if (k < 0 || L.e_j[e] != 0) { continue; }
        int64_t run = out_size[k]; // start code + verbatim bytes + part A, written by the NAL's own warp
        const int np = L.e_np[e];
        for (int j = 0; j < np; j++) {
            L.e_prefix[e + j] = run;
            run += L.e_size[e + j];
        }
        out_size[k] = run;
    }
}

// ---- short NALs: one LANE per NAL -----------------------------------------------------------------------------------------
// A warp per NAL leaves 28 of 32 lanes idle on a 64-byte NAL (measured 58 GB/s).  A NAL of at most kLaneMax escaped bytes and no
// other part is walked by ONE lane with the byte state machine itself (rbsp_to_nal as written, h264_nal.c:92-132), the 32 lanes of
// a warp taking 32 consecutive NALs: their bytes are neighbours in memory, so the lanes' 16-byte loads share cache lines.  The
// longer NALs of the group then take the warp-cooperative walk, one after the other.
constexpr int64_t kLaneMax = 112; // (32 x kLaneOut bytes of staging per warp must fit the 48 KB of static shared memory of a block)
constexpr int kLaneOut = (int)kLaneMax + (int)kLaneMax / 2 + 4; // most output bytes of such a NAL: one 03 per two bytes + start code
__device__ __forceinline__ bool lane_sized(const AssembleParts& P, int64_t boff, int64_t bend)
{
    return P.raw_off == nullptr && P.a_off == nullptr && P.len_size == 0 && bend > boff && bend - boff <= kLaneMax;
}
// kWrite: emit start code + escaped bytes at out[o ...); returns the insertions
template <bool kWrite>
__device__ __forceinline__ uint32_t lane_rbsp_to_nal(const uint8_t* __restrict__ base, int64_t boff, int64_t bend, int sc_len, uint8_t* __restrict__ out, int64_t o)
{
    uint32_t ins = 0, count = 0;
    if (kWrite) {
        for (int i = 0; i < sc_len; i++) { out[o++] = (i == sc_len - 1) ? 1 : 0; }
    }
    for (int64_t c = boff & ~(int64_t)15; c < bend; c += 16) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + c));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const int64_t p = c + j;
            if (p >= boff && p < bend) {
                const uint32_t bv = (w[j >> 2] >> (8 * (j & 3))) & 0xFFu;
                if (count == 2u && bv <= 3u) {
                    ins++;
                    count = 0u;
                    if (kWrite) { out[o++] = 3; }
                }
                if (kWrite) { out[o++] = (uint8_t)bv; }
                count = (bv == 0u) ? count + 1u : 0u;
            }
        }
    }
    return ins;
}

__global__ void __launch_bounds__(kInsThreads) insert_count_kernel(const AssembleParts P, int64_t n, int64_t* __restrict__ out_size,
                                                                   unsigned long long* __restrict__ n_ins_total, const uint8_t* __restrict__ split_flag)
{
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * kInsWarps + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * kInsWarps;
    unsigned long long local_ins = 0;
    for (int64_t g = wid; g * 32 < n; g += nw) {
      // the group's short NALs, one per lane
      const int64_t kl = g * 32 + lane;
      bool later = false; // this lane's NAL takes the warp-cooperative walk below
      if (kl < n) {
          const int64_t lb = P.b_off ? P.b_off[kl] : 0, le = P.b_off ? P.b_end[kl] : 0;
          if (lane_sized(P, lb, le)) {
              const uint32_t ins = lane_rbsp_to_nal<false>(P.b_base, lb, le, P.sc_len, nullptr, 0);
              out_size[kl] = (int64_t)P.sc_len + (le - lb) + (int64_t)ins;
              if (ins) { atomicAdd(n_ins_total, (unsigned long long)ins); }
          } else { later = true; }
      }
      uint32_t big = __ballot_sync(0xFFFFFFFFu, later);
      while (big) {
        const int64_t k = g * 32 + (__ffs((int)big) - 1);
        big &= big - 1u;
        const int64_t boff = P.b_off ? P.b_off[k] : 0, bend = P.b_off ? P.b_end[k] : 0;
        if (P.skip_neg_b && (bend < 0 || bend < boff)) {
            if (lane == 0) { out_size[k] = 0; }
            continue;
        }
        const int64_t roff = P.raw_off ? P.raw_off[k] : 0, rend = P.raw_off ? P.raw_end[k] : 0;
        const int64_t aoff = P.a_off ? P.a_off[k] : 0, aend = P.a_off ? P.a_end[k] : 0;
        uint32_t run_m = 0;
        uint32_t total = count_part(P.a_base, aoff, aend, lane, run_m);
        const bool split = (bend - boff >= kSplitMin) && split_flag[k] != 0; // part B is counted piece by piece (extras_count_kernel)
        if (!split) { total += count_part(P.b_base, boff, bend, lane, run_m); }
        if (lane == 0) {
            out_size[k] = (int64_t)P.sc_len + (int64_t)P.len_size + (rend > roff ? rend - roff : 0) + (aend > aoff ? aend - aoff : 0) +
                          ((bend > boff && !split) ? bend - boff : 0) + (int64_t)total;
        }
        local_ins += total;
      }
    }
    if (lane == 0 && local_ins) { atomicAdd(n_ins_total, local_ins); }
}

// writes one analysed row at dst; returns the bytes written (warp uniform)
__device__ __forceinline__ uint32_t write_row(const RowInfo& r, uint8_t* __restrict__ dst, int lane)
{
    const bool clean = r.fast || (__ballot_sync(0xFFFFFFFFu, r.ins != 0u || r.valid != 0xFFFFu) == 0u);
    uint32_t cnt = 16u, inc = ((uint32_t)lane + 1u) * 16u, row_total = 512u;
    if (!clean) { // output offsets of the lanes: only rows with insertions or partial chunks need the scan
        cnt = (uint32_t)__popc(r.valid) + (uint32_t)__popc(r.ins);
        inc = warp_incl_scan_u32(cnt, lane);
        row_total = __shfl_sync(0xFFFFFFFFu, inc, 31);
    }
    if (clean) {
        // full row without insertions: destination vector d (16-byte aligned) = source bytes [16 d + head - 16 ...]
        const uint32_t head = (uint32_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
        if (head != 0u) { // the bytes in front of the first / behind the last aligned vector, one per lane
            // lanes 0..15 look at lane 0's chunk (head bytes), lanes 16..31 at lane 31's chunk (tail bytes)
            const uint32_t w0 = __shfl_sync(0xFFFFFFFFu, r.v.x, lane < 16 ? 0 : 31), w1 = __shfl_sync(0xFFFFFFFFu, r.v.y, lane < 16 ? 0 : 31);
            const uint32_t w2 = __shfl_sync(0xFFFFFFFFu, r.v.z, lane < 16 ? 0 : 31), w3 = __shfl_sync(0xFFFFFFFFu, r.v.w, lane < 16 ? 0 : 31);
            const uint32_t j = (uint32_t)lane & 15u;
            const uint32_t w = (j < 4u) ? w0 : (j < 8u) ? w1 : (j < 12u) ? w2 : w3;
            const uint8_t bv = (uint8_t)(w >> (8u * (j & 3u)));
            if (lane < 16) { if (j < head) { dst[j] = bv; } }
            else { if (j >= head) { dst[496u + j] = bv; } }
        }
        // lane l assembles destination bytes [head + 16 l, head + 16 l + 16) from its chunk and the next lane's
        uint4 nx;
        nx.x = __shfl_down_sync(0xFFFFFFFFu, r.v.x, 1);
        nx.y = __shfl_down_sync(0xFFFFFFFFu, r.v.y, 1);
        nx.z = __shfl_down_sync(0xFFFFFFFFu, r.v.z, 1);
        nx.w = __shfl_down_sync(0xFFFFFFFFu, r.v.w, 1);
        if (head == 0u) {
            *reinterpret_cast<uint4*>(dst + lane * 16) = r.v;
        } else {
            const uint32_t W[8] = {r.v.x, r.v.y, r.v.z, r.v.w, nx.x, nx.y, nx.z, nx.w};
            const uint32_t q = head >> 2, sh = (head & 3u) * 8u;
            uint32_t x[5];
#pragma unroll
            for (int e = 0; e < 5; e++) { x[e] = (q == 0u) ? W[e] : (q == 1u) ? W[e + 1] : (q == 2u) ? W[e + 2] : W[(e + 3) & 7]; }
            uint4 o4;
            o4.x = __funnelshift_r(x[0], x[1], sh);
            o4.y = __funnelshift_r(x[1], x[2], sh);
            o4.z = __funnelshift_r(x[2], x[3], sh);
            o4.w = __funnelshift_r(x[3], x[4], sh);
            if (lane < 31) { *reinterpret_cast<uint4*>(dst + head + lane * 16) = o4; } // lane 31's 16 - head bytes were stored above
        }
    } else {
        uint8_t* p = dst + (inc - cnt);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if ((r.valid >> j) & 1u) {
                if ((r.ins >> j) & 1u) { *p++ = 3; }
                *p++ = (uint8_t)byte_of(r.v, j);
            }
        }
    }
    return row_total;
}

// writes one escaped part at out + o; returns the bytes written (warp uniform)
template <int kInsAhead>
__device__ __forceinline__ int64_t write_part_n(const uint8_t* __restrict__ base, int64_t off, int64_t end, uint8_t* __restrict__ out, int64_t o, int lane,
                                                uint32_t& run_m)
{
    const int64_t o0 = o;
    const int64_t row0 = off & ~(int64_t)15;
    for (int64_t rowq = row0; rowq < end; rowq += 512 * kInsAhead) {
        uint4 vq[kInsAhead];
#pragma unroll
        for (int k = 0; k < kInsAhead; k++) { vq[k] = load_row_chunk(base, rowq + 512 * k, off, end, lane); }
#pragma unroll
        for (int k = 0; k < kInsAhead; k++) {
            const int64_t row = rowq + 512 * k;
            if (row >= end) { break; }
            const RowInfo r = insert_row(vq[k], row, off, end, lane, run_m);
            o += write_row(r, out + o, lane);
        }
    }
    run_m -= (uint32_t)(row0 + ((end - row0 + 511) / 512) * 512 - end);
    return o - o0;
}

__device__ __forceinline__ int64_t write_part(const uint8_t* __restrict__ base, int64_t off, int64_t end, uint8_t* __restrict__ out, int64_t o, int lane,
                                              uint32_t& run_m)
{
    if (end <= off) { return 0; }
    const int64_t len = end - off;
    if (len > 2048) { return write_part_n<4>(base, off, end, out, o, lane, run_m); }
    if (len > 512) { return write_part_n<2>(base, off, end, out, o, lane, run_m); }
    return write_part_n<1>(base, off, end, out, o, lane, run_m);
}

// pass 2: write start code, verbatim bytes and escaped bytes of every NAL at out_off[k]
__global__ void __launch_bounds__(kInsThreads) insert_write_kernel(const AssembleParts P, int64_t n, const int64_t* __restrict__ out_off,
                                                                   uint8_t* __restrict__ out, int64_t out_cap, const uint8_t* __restrict__ split_flag)
{
    __shared__ __align__(16) uint8_t lane_stage[kInsWarps][32 * kLaneOut + 32]; // groups of short NALs: their output, staged (see below)
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * kInsWarps + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * kInsWarps;
    for (int64_t g = wid; g * 32 < n; g += nw) {
      const int64_t kl = g * 32 + lane;
      bool later = false;
      const int64_t lb = (kl < n && P.b_off) ? P.b_off[kl] : 0, le = (kl < n && P.b_off) ? P.b_end[kl] : 0;
      const bool mine = kl < n && lane_sized(P, lb, le); // (see insert_count_kernel)
      const uint32_t in_group = __ballot_sync(0xFFFFFFFFu, kl < n), small_mask = __ballot_sync(0xFFFFFFFFu, mine);
      const int64_t g_end = (g * 32 + 32 < n) ? g * 32 + 32 : n;
      const int64_t o_first = out_off[g * 32], o_last = out_off[g_end];
      if (small_mask == in_group && o_last <= out_cap) {
          // Every NAL of the group is short: their outputs are one contiguous range.  The lanes write their bytes into shared memory
          // (byte-granular stores to global memory cost one L2 sector operation per byte: measured 5.6 ms per GiB of 64-byte NALs),
          // the warp then copies the range out in 16-byte vectors.  The staging keeps the range's alignment modulo 16.
          uint8_t* const sw = lane_stage[threadIdx.x >> 5];
          const uint32_t phase = (uint32_t)((uintptr_t)(out + o_first) & 15u);
          if (mine) { (void)lane_rbsp_to_nal<true>(P.b_base, lb, le, P.sc_len, sw + phase, out_off[kl] - o_first); }
          __syncwarp();
          const int64_t len = o_last - o_first;
          uint8_t* const dst = out + o_first;
          int64_t head = (int64_t)((16u - phase) & 15u);
          if (head > len) { head = len; }
          if (lane < head) { dst[lane] = sw[phase + lane]; }
          const int64_t nv = (len - head) >> 4;
          for (int64_t i = lane; i < nv; i += 32) {
              *reinterpret_cast<uint4*>(dst + head + (i << 4)) = *reinterpret_cast<const uint4*>(sw + phase + head + (i << 4)); // (phase + head is 0 or 16)
          }
          const int64_t done = head + (nv << 4);
          if (lane < len - done) { dst[done + lane] = sw[phase + done + lane]; }
          __syncwarp(); // the staging is rewritten by the warp's next group
          continue;
      }
      if (kl < n) {
          if (mine) {
              if (out_off[kl + 1] <= out_cap) { (void)lane_rbsp_to_nal<true>(P.b_base, lb, le, P.sc_len, out, out_off[kl]); }
          } else { later = true; }
      }
      uint32_t big = __ballot_sync(0xFFFFFFFFu, later);
      while (big) {
        const int64_t k = g * 32 + (__ffs((int)big) - 1);
        big &= big - 1u;
        const int64_t boff = P.b_off ? P.b_off[k] : 0, bend = P.b_off ? P.b_end[k] : 0;
        if (P.skip_neg_b && (bend < 0 || bend < boff)) { continue; }
        int64_t o = out_off[k];
        if (out_off[k + 1] > out_cap) { continue; } // capacity overflow is reported by the summary
        if (lane < P.sc_len) { out[o + lane] = (lane == P.sc_len - 1) ? 1 : 0; }
        o += P.sc_len;
        if (P.len_size) { // big-endian length of what follows
            const uint64_t nbytes = (uint64_t)(out_off[k + 1] - out_off[k] - P.len_size);
            if (lane < P.len_size) { out[o + lane] = (uint8_t)(nbytes >> (8 * (P.len_size - 1 - lane))); }
            o += P.len_size;
        }
        if (P.raw_off) {
            const int64_t roff = P.raw_off[k], rend = P.raw_end[k];
            warp_copy_bytes(out + o, P.raw_base + roff, rend - roff, lane);
            if (rend > roff) { o += rend - roff; }
        }
        uint32_t run_m = 0;
        if (P.a_off) { o += write_part(P.a_base, P.a_off[k], P.a_end[k], out, o, lane, run_m); }
        if (bend - boff < kSplitMin || !split_flag[k]) { o += write_part(P.b_base, boff, bend, out, o, lane, run_m); }
      }
    }
}

// pass 2 over the pieces of the long parts
__global__ void __launch_bounds__(kInsThreads) extras_write_kernel(const AssembleParts P, SplitList L, const int64_t* __restrict__ out_off,
                                                                   uint8_t* __restrict__ out, int64_t out_cap)
{
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * kInsWarps + (threadIdx.x >> 5);
    const int64_t nw = (int64_t)gridDim.x * kInsWarps;
    int64_t ne = (int64_t)*L.n_extras;
    if (ne > L.cap) { ne = L.cap; }
    for (int64_t e = wid; e < ne; e += nw) {
        const int64_t k = L.e_k[e];
        if (k < 0) { continue; }
        if (out_off[k + 1] > out_cap) { continue; } // capacity overflow is reported by the summary
        uint32_t run_m = 0;
        if (L.e_j[e] == 0 && P.a_off) { (void)count_part(P.a_base, P.a_off[k], P.a_end[k], lane, run_m); }
        (void)write_part(P.b_base, L.e_off[e], L.e_end[e], out, out_off[k] + L.e_prefix[e], lane, run_m);
    }
}

__global__ void insert_summary_kernel(const int64_t* out_off, int64_t n, int64_t out_cap, const unsigned long long* n_ins, hevcb_insert_summary* s)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        s->n_nals = n;
        s->out_bytes = out_off[n];
        s->n_inserted = (int64_t)*n_ins;
        s->overflow = out_off[n] > out_cap ? 1 : 0;
        s->pad = 0;
    }
}

// ---- exclusive scan of int64 sizes (3 kernels) -----------------------------------------------------------------
constexpr int kSThreads = 512, kSItems = 8, kSTile = kSThreads * kSItems;
__device__ __forceinline__ long long block_incl_scan_ll(long long v, long long* ws, long long& tot)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const long long o = __shfl_up_sync(0xFFFFFFFFu, v, d);
        if (lane >= d) { v += o; }
    }
    if (lane == 31) { ws[warp] = v; }
    __syncthreads();
    if (warp == 0) {
        long long w = (lane < kSThreads / 32) ? ws[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= d) { w += o; }
        }
        if (lane < kSThreads / 32) { ws[lane] = w; }
    }
    __syncthreads();
    const long long base = warp > 0 ? ws[warp - 1] : 0;
    tot = ws[kSThreads / 32 - 1];
    return v + base;
}
__global__ void __launch_bounds__(kSThreads) sizes_reduce_kernel(const int64_t* in, int64_t n, long long* bs)
{
    __shared__ long long ws[kSThreads / 32];
    long long s = 0;
    for (int j = 0; j < kSItems; j++) {
        const int64_t i = (int64_t)blockIdx.x * kSTile + (int64_t)j * kSThreads + threadIdx.x;
        if (i < n) { s += in[i]; }
    }
    long long tot;
    block_incl_scan_ll(s, ws, tot);
    if (threadIdx.x == 0) { bs[blockIdx.x] = tot; }
}
__global__ void __launch_bounds__(kSThreads) sizes_blocksums_kernel(long long* bs, int64_t nb)
{
    __shared__ long long ws[kSThreads / 32];
    __shared__ long long carry_s;
    if (threadIdx.x == 0) { carry_s = 0; }
    __syncthreads();
    for (int64_t base = 0; base < nb; base += kSThreads) {
        const int64_t i = base + threadIdx.x;
        const long long v = i < nb ? bs[i] : 0;
        long long tot;
        const long long inc = block_incl_scan_ll(v, ws, tot);
        const long long carry = carry_s;
        __syncthreads();
        if (i < nb) { bs[i] = carry + inc - v; }
        if (threadIdx.x == 0) { carry_s = carry + tot; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { bs[nb] = carry_s; }
}
__global__ void __launch_bounds__(kSThreads) sizes_apply_kernel(const int64_t* in, int64_t n, const long long* bs, int64_t nb, int64_t* out)
{
    __shared__ long long ws[kSThreads / 32];
    long long carry = bs[blockIdx.x];
    for (int j = 0; j < kSItems; j++) {
        const int64_t i = (int64_t)blockIdx.x * kSTile + (int64_t)j * kSThreads + threadIdx.x;
        const long long v = i < n ? in[i] : 0;
        long long tot;
        const long long inc = block_incl_scan_ll(v, ws, tot);
        if (i < n) { out[i] = carry + inc - v; }
        carry += tot;
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { out[n] = bs[nb]; }
}


// ====================================================================================================================
// Single-pass assembly: ONE read of every source byte (SURVEY 8d scores the insertion as N_rbsp + N_nal).
//
// The two-pass kernels above read every escaped byte twice (count, then write).  Here the work is cut into ITEMS -- pieces of at
// most kPieceCap bytes of one part of one NAL -- and a CTA takes kFWarps consecutive items (a TILE, claimed with a ticket so
// that every earlier tile is already running): each warp pulls its piece into shared memory with one TMA bulk copy, counts its
// insertions there, the tile's size goes through a decoupled look-back over 8-byte tile states (aggregate / inclusive prefix),
// and the warps write their pieces out of the SAME shared memory once the tile's output offset is known.  Pieces need nothing
// from each other: rbsp_to_nal's state at a cut is a function of the zero run in front of it, which the piece's warp counts by
// looking back from the cut (usually one byte).  NAL k's out_off is written by its first item.
// Launches: plan -> 3-kernel scan of (items, plain bytes) per NAL -> item descriptors -> fused kernel -> summary.
// ====================================================================================================================
constexpr int kFWarps = 8, kFThreads = kFWarps * 32;
constexpr int kPieceCap = 4224;                      // bytes of a piece (a multiple of 16; 8.25 rows: a 16 KiB NAL is four pieces at any alignment)
constexpr int kPieceRows = (kPieceCap + 511) / 512;  // 9
constexpr int kPieceSmem = kPieceCap + 32;           // + slack: the vector copies read up to 31 bytes behind a piece
#ifndef HEVCB_FUSED_LOOK
#define HEVCB_FUSED_LOOK 2
#endif
constexpr int kLook = HEVCB_FUSED_LOOK;              // a scanner batch is 32 x kLook tiles (it waits for the slowest of them)
constexpr int kItemShift = 37;                       // packed per-NAL value: items << 37 | plain bytes (start code + prefix + parts, no insertions)
constexpr unsigned long long kBytesMask = (1ull << kItemShift) - 1ull;

struct FusedHeader {
    unsigned long long ticket;
    unsigned long long n_ins;
    unsigned long long pad[2];
};

__device__ __forceinline__ int64_t piece_count(int64_t off, int64_t end)
{
    if (end <= off) { return 0; }
    const int64_t L = end - (off & ~(int64_t)15);
    return (L + kPieceCap - 1) / kPieceCap;
}
// piece j of np of the part [off, end): source window [s0, s0 + plen) with s0 16-byte aligned; the piece is its intersection with the part
__device__ __forceinline__ void piece_bounds(int64_t off, int64_t end, int64_t np, int64_t j, int64_t& s0, int64_t& p0, int64_t& p1)
{
    const int64_t a0 = off & ~(int64_t)15, L = end - a0;
    const int64_t plen = (((L + np - 1) / np) + 15) & ~(int64_t)15;
    s0 = a0 + j * plen;
    p0 = s0 > off ? s0 : off;
    p1 = s0 + plen < end ? s0 + plen : end;
    if (p1 < p0) { p1 = p0; }
}

__global__ void __launch_bounds__(256) fused_plan_kernel(const AssembleParts P, int64_t n, int64_t* __restrict__ packed)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) { return; }
    const int64_t boff = P.b_off ? P.b_off[k] : 0, bend = P.b_off ? P.b_end[k] : 0;
    unsigned long long items = 1, bytes = 0;
    if (!(P.skip_neg_b && (bend < 0 || bend < boff))) {
        const int64_t roff = P.raw_off ? P.raw_off[k] : 0, rend = P.raw_off ? P.raw_end[k] : 0;
        const int64_t aoff = P.a_off ? P.a_off[k] : 0, aend = P.a_off ? P.a_end[k] : 0;
        const int64_t ni = piece_count(roff, rend) + piece_count(aoff, aend) + piece_count(boff, bend);
        items = (unsigned long long)(ni > 0 ? ni : 1);
        bytes = (unsigned long long)((int64_t)P.sc_len + (int64_t)P.len_size + (rend > roff ? rend - roff : 0) + (aend > aoff ? aend - aoff : 0) + (bend > boff ? bend - boff : 0));
    }
    packed[k] = (int64_t)((items << kItemShift) | (bytes & kBytesMask));
}

// What a warp of the fused kernel needs to know about its item, written by fused_fill_kernel so that the fused kernel starts
// with ONE load instead of a chain of dependent ones (item -> NAL -> extents -> the bytes in front of the piece).
struct __align__(16) ItemDesc {
    int64_t s0;       // source window start (16-byte aligned offset into the part's base)
    int32_t lead;     // the piece starts at s0 + lead
    int32_t len;      // bytes of the piece
    int32_t k;        // NAL
    uint32_t flags;   // bits 0..1: kItem*, bit 2: first item of its NAL, bit 3: the NAL produces nothing (hevcb_insert: nal_to_rbsp failed)
    uint32_t run_m;   // escaped pieces: zero bytes that end right in front of the piece (rbsp_to_nal's state at the cut)
    uint32_t raw_len; // bytes of the NAL's verbatim part (the length prefix of re-framed NALs)
};
enum { kItemNone = 0, kItemRaw = 1, kItemEscA = 2, kItemEscB = 3 };

// zero bytes that end right in front of position p of base[], not looking below lo; all = the run reaches lo
__device__ __forceinline__ uint32_t zeros_back_scalar(const uint8_t* __restrict__ base, int64_t lo, int64_t p, bool& all)
{
    uint32_t m = 0;
    while (p > lo && base[p - 1] == 0) { p--; m++; }
    all = (p <= lo);
    return m;
}

__device__ __forceinline__ void fill_item(const AssembleParts& P, int64_t k, int64_t j, ItemDesc* __restrict__ d)
{
    ItemDesc it;
    it.s0 = 0; it.lead = 0; it.len = 0; it.k = (int32_t)k; it.flags = (j == 0 ? 4u : 0u); it.run_m = 0; it.raw_len = 0;
    const int64_t boff = P.b_off ? P.b_off[k] : 0, bend = P.b_off ? P.b_end[k] : 0;
    if (P.skip_neg_b && (bend < 0 || bend < boff)) {
        it.flags |= 8u;
    } else {
        const int64_t roff = P.raw_off ? P.raw_off[k] : 0, rend = P.raw_off ? P.raw_end[k] : 0;
        const int64_t aoff = P.a_off ? P.a_off[k] : 0, aend = P.a_off ? P.a_end[k] : 0;
        const int64_t nr = piece_count(roff, rend), na = piece_count(aoff, aend), nb = piece_count(boff, bend);
        it.raw_len = (uint32_t)(rend > roff ? rend - roff : 0);
        int64_t s0 = 0, p0 = 0, p1 = 0;
        uint32_t kind = kItemNone;
        bool all;
        if (j < nr) {
            kind = kItemRaw;
            piece_bounds(roff, rend, nr, j, s0, p0, p1);
        } else if (j < nr + na) {
            kind = kItemEscA;
            piece_bounds(aoff, aend, na, j - nr, s0, p0, p1);
            it.run_m = zeros_back_scalar(P.a_base, aoff, p0, all);
        } else if (j < nr + na + nb) {
            kind = kItemEscB;
            piece_bounds(boff, bend, nb, j - nr - na, s0, p0, p1);
            it.run_m = zeros_back_scalar(P.b_base, boff, p0, all);
            if (all && aend > aoff) { bool all2; it.run_m += zeros_back_scalar(P.a_base, aoff, aend, all2); } // the run continues into the header part
        } // else: a NAL without bytes, its only item writes the prefix
        if (p1 <= p0) { kind = kItemNone; }
        it.s0 = s0; it.lead = (int32_t)(p0 - s0); it.len = (int32_t)(p1 - p0);
        it.flags |= kind;
    }
    *d = it;
}

// item descriptors (items of one NAL are consecutive): one thread per item, the NAL found by bisection of the scanned item counts
__global__ void __launch_bounds__(256) fused_fill_kernel(const AssembleParts P, const int64_t* __restrict__ first, int64_t n, ItemDesc* __restrict__ items,
                                                         int64_t cap_items)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_items = (int64_t)((unsigned long long)first[n] >> kItemShift);
    if (i >= n_items || n_items > cap_items) { return; }
    int64_t lo = 0, hi = n; // the NAL k with first[k] <= i < first[k + 1]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)((unsigned long long)first[mid] >> kItemShift) <= i) { lo = mid; } else { hi = mid; }
    }
    fill_item(P, lo, i - (int64_t)((unsigned long long)first[lo] >> kItemShift), &items[i]);
}

__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one warp copies len bytes from shared memory (any alignment) to global memory (any alignment): aligned 16-byte stores, the source as
// two aligned 16-byte loads funnel-shifted by the (copy-uniform) misalignment between source and destination (Q = its word part)
template <int Q>
__device__ __forceinline__ void smem_vectors_out(uint4* __restrict__ da, const uint8_t* __restrict__ sa, int64_t nv, uint32_t sh, int lane)
{
#pragma unroll 2
    for (int64_t i = lane; i < nv; i += 32) {
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
__device__ __forceinline__ void warp_copy_from_smem(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int64_t len, int lane)
{
    if (len <= 0) { return; }
    int64_t head = (int64_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u);
    if (head > len) { head = len; }
    if (lane < head) { dst[lane] = src[lane]; }
    const int64_t nv = (len - head) >> 4;
    const uint8_t* s0 = src + head;
    const uint32_t mis = smem_addr(s0) & 15u, sh = (mis & 3u) * 8u;
    const uint8_t* sa = s0 - mis;
    uint4* da = reinterpret_cast<uint4*>(dst + head);
    switch (mis >> 2) {
        case 0: smem_vectors_out<0>(da, sa, nv, sh, lane); break;
        case 1: smem_vectors_out<1>(da, sa, nv, sh, lane); break;
        case 2: smem_vectors_out<2>(da, sa, nv, sh, lane); break;
        default: smem_vectors_out<3>(da, sa, nv, sh, lane); break;
    }
    const int64_t done = head + (nv << 4);
    if (lane < len - done) { dst[done + lane] = src[done + lane]; }
}

// Scanner CTA of the fused kernel (the CTA that draws ticket 0): turns the tile aggregates into exclusive prefixes, in tile order.
// Batches of 32 x kLook consecutive tiles go round-robin to the CTA's warps; a warp polls its batch's aggregates (one reader per state),
// scans them with shuffles, takes the running total from the warp of the previous batch through ONE shared-memory word (the only serial
// step) and publishes every tile's exclusive prefix.  (A look-back by the tiles themselves piles up: with ~900 tiles in flight most
// predecessors hold only an aggregate, every tile walks hundreds of states back, and the chain of inclusive prefixes advances one window
// per L2 round trip -- measured at 40 % of a tile's lifetime.)
__device__ __forceinline__ void fused_scanner(const unsigned long long* __restrict__ tile_state, unsigned long long* __restrict__ tile_excl, int64_t n_tiles,
                                              volatile ulonglong2* run, int warp, int lane, int64_t* __restrict__ total_out)
{
    const int64_t batch_tiles = 32 * kLook, n_batches = (n_tiles + batch_tiles - 1) / batch_tiles;
    for (int64_t b = warp; b < n_batches; b += kFWarps) {
        const int64_t first = b * batch_tiles + (int64_t)lane * kLook;
        unsigned long long w[kLook];
#pragma unroll
        for (int q = 0; q < kLook; q++) { w[q] = (first + q < n_tiles) ? ld_relaxed_u64(&tile_state[first + q]) : (1ull << 62); }
        for (;;) {
            bool missing = false;
#pragma unroll
            for (int q = 0; q < kLook; q++) { missing = missing || ((w[q] >> 62) == 0ull); }
            if (!__any_sync(0xFFFFFFFFu, missing)) { break; }
            __nanosleep(64);
#pragma unroll
            for (int q = 0; q < kLook; q++) { if ((w[q] >> 62) == 0ull) { w[q] = ld_relaxed_u64(&tile_state[first + q]); } }
        }
        long long mine = 0;
#pragma unroll
        for (int q = 0; q < kLook; q++) { mine += (long long)(w[q] & ((1ull << 62) - 1ull)); }
        long long inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) { inc += o; }
        }
        const long long tot = __shfl_sync(0xFFFFFFFFu, inc, 31);
        unsigned long long seq, base;
        do { // the running total of the batches before this one (every lane reads the same word)
            asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(seq), "=l"(base) : "r"(smem_addr((const void*)run)) : "memory");
        } while (seq != (unsigned long long)b);
        if (lane == 0) {
            asm volatile("st.volatile.shared.v2.u64 [%0], {%1, %2};" ::"r"(smem_addr((const void*)run)), "l"((unsigned long long)(b + 1)), "l"(base + (unsigned long long)tot) : "memory");
            if (b == n_batches - 1) { *total_out = (int64_t)(base + (unsigned long long)tot); }
        }
        long long e = (long long)base + inc - mine;
#pragma unroll
        for (int q = 0; q < kLook; q++) {
            if (first + q < n_tiles) { st_relaxed_u64(&tile_excl[first + q], (1ull << 62) | (unsigned long long)e); }
            e += (long long)(w[q] & ((1ull << 62) - 1ull));
        }
    }
}

struct __align__(16) FusedSmem {
    uint8_t piece[kFWarps][kPieceSmem];
    unsigned long long bar[kFWarps];
    uint32_t run_in[kFWarps][kPieceRows + 1]; // zero run entering every row of an escaped piece (count pass -> write pass)
    long long size[kFWarps];
    long long excl;
    long long ticket;
    long long n_items, n_tiles; // (kept here rather than in registers: see the note on the kernel's register budget)
    uint4 meta[kFWarps];        // per warp: what the write phase needs of the item (NAL, raw length, flags), parked across the prefix wait
};

__global__ void __launch_bounds__(kFThreads, 6) fused_assemble_kernel(const AssembleParts P, int64_t n, const int64_t* __restrict__ first,
                                                                   const ItemDesc* __restrict__ items, int64_t cap_items,
                                                                   unsigned long long* __restrict__ tile_state, unsigned long long* __restrict__ tile_excl,
                                                                   FusedHeader* __restrict__ hdr, int64_t* __restrict__ out_off, uint8_t* __restrict__ out, int64_t out_cap, int stagger_ns, int persistent)
{
    extern __shared__ __align__(128) uint8_t fsm_raw[];
    FusedSmem& sm = *reinterpret_cast<FusedSmem*>(fsm_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const int64_t n_items = (int64_t)((unsigned long long)first[n] >> kItemShift);
        if (n_items > cap_items) { return; } // only when the sources alone exceed out_cap: the summary kernel reports the plain sizes
        const int64_t n_tiles = (n_items + kFWarps - 1) / kFWarps;
        if (!persistent && (int64_t)blockIdx.x > n_tiles) { return; } // one tile per CTA: n_tiles workers and the scanner
        if (threadIdx.x == 0) { sm.n_items = n_items; sm.n_tiles = n_tiles; } // (read after the first __syncthreads of the loop)
    }
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(&sm.bar[warp])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    uint32_t phase = 0; // parity of this warp's mbarrier (flips with every piece the warp loads)
    // The CTAs of an SM start staggered: the in-order prefix couples the tiles in flight, and CTAs that all draw their tickets at the same
    // moment keep loading, counting and writing in lock-step (read bursts, then write bursts) instead of overlapping the phases.
    if (persistent && stagger_ns > 0 && gridDim.x > 1) { __nanosleep((unsigned)(blockIdx.x * 6ull / gridDim.x) * (unsigned)stagger_ns); }
  // persistent CTAs: every round draws a ticket; ticket 0 makes its CTA the scanner, ticket i the worker of tile i - 1 (so every earlier
  // tile is running or done, and so is the scanner)
  for (;;) {
    if (threadIdx.x == 0) {
        sm.ticket = (long long)atomicAdd(&hdr->ticket, 1ull);
        if (sm.ticket == 0) { reinterpret_cast<volatile unsigned long long*>(sm.piece[0])[0] = 0ull; reinterpret_cast<volatile unsigned long long*>(sm.piece[0])[1] = 0ull; }
    }
    __syncthreads();
    if (sm.ticket == 0) {
        fused_scanner(tile_state, tile_excl, sm.n_tiles, reinterpret_cast<volatile ulonglong2*>(sm.piece[0]), warp, lane, out_off + n);
        return;
    }
    if (sm.ticket - 1 >= sm.n_tiles) { return; }
    const int64_t item = (sm.ticket - 1) * kFWarps + warp;
    const bool active = item < sm.n_items;
    // ---- what this warp's item is.  (Everything that lives across the count, the wait for the prefix and the write is 32 bits wide
    // and the tile index stays in shared memory: at 42 registers per thread -- six CTAs per SM -- the 64-bit positions of the
    // descriptor used to live on the stack, and with the SM's L1 carved out as shared memory those loads go to L2.)
    enum { kNone = 0, kRaw = 1, kEsc = 2 };
    int kind = kNone;
    bool first_item = false, nal_skipped = false;
    int k = 0, q0 = 0, q1 = 0; // NAL; the piece inside its shared-memory window [q0, q1)
    uint32_t run_m = 0, nbytes_raw = 0;
    uint8_t* const sp = sm.piece[warp];
    if (active) {
        const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(&items[item])), d1 = __ldg(reinterpret_cast<const uint4*>(&items[item]) + 1);
        q0 = (int32_t)d0.z;
        q1 = q0 + (int32_t)d0.w;
        k = (int32_t)d1.x;
        const uint32_t fl = d1.y;
        run_m = d1.z;
        nbytes_raw = d1.w;
        first_item = (fl & 4u) != 0u;
        nal_skipped = (fl & 8u) != 0u;
        const uint32_t ik = fl & 3u;
        kind = (ik == kItemNone) ? kNone : (ik == kItemRaw ? kRaw : kEsc);
        if (kind != kNone && q1 <= q0) { kind = kNone; }
        if (kind != kNone) {
            const uint32_t bytes = (uint32_t)((q1 + 15) & ~15);
            if (lane == 0) {
                const int64_t s0 = (int64_t)(((unsigned long long)d0.y << 32) | d0.x);
                const uint8_t* base = (ik == kItemRaw) ? P.raw_base : (ik == kItemEscA ? P.a_base : P.b_base);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&sm.bar[warp])), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(sp)), "l"(base + s0),
                             "r"(bytes), "r"(smem_addr(&sm.bar[warp]))
                             : "memory");
            }
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_addr(&sm.bar[warp])), "r"(phase) : "memory");
            }
            phase ^= 1u;
        }
    }
    // ---- count
    const bool writes_prefix = active && first_item && !nal_skipped;
    int size = 0;
    uint32_t fastmask = 0, n_ins = 0;
    const int rows = (kind != kNone) ? ((q1 + 511) >> 9) : 0;
    if (kind == kEsc) {
        int r = 0;
        while (r < rows) {
            const int row = r << 9, cpos = row + lane * 16;
            if (row >= q0 && row + 2048 <= q1) {
                // Four interior rows at once (the common case): the zero-pair test of insert_row over 2 KiB with ONE vote.  A pair that
                // straddles two rows of the group is seen through the next row's first word; the pair across the group's end is left to
                // run_m, as in insert_row.
                uint4 v[4];
                uint32_t nx[4];
#pragma unroll
                for (int g = 0; g < 4; g++) { v[g] = *reinterpret_cast<const uint4*>(sp + cpos + (g << 9)); }
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    nx[g] = __shfl_down_sync(0xFFFFFFFFu, v[g].x, 1);
                    const uint32_t nrow = __shfl_sync(0xFFFFFFFFu, v[g < 3 ? g + 1 : 3].x, 0); // first word of the next row of the group
                    if (lane == 31) { nx[g] = (g < 3) ? nrow : 0xFFFFFFFFu; }
                }
                uint32_t pair = 0;
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    const uint32_t m0 = v[g].x | __funnelshift_r(v[g].x, v[g].y, 8), m1 = v[g].y | __funnelshift_r(v[g].y, v[g].z, 8);
                    const uint32_t m2 = v[g].z | __funnelshift_r(v[g].z, v[g].w, 8), m3 = v[g].w | __funnelshift_r(v[g].w, nx[g], 8);
                    const uint32_t c = 0x01010101u;
                    pair |= ((m0 - c) & ~m0) | ((m1 - c) & ~m1) | ((m2 - c) & ~m2) | ((m3 - c) & ~m3);
                }
                const uint32_t b0 = __shfl_sync(0xFFFFFFFFu, v[0].x, 0) & 0xFFu, bl = __shfl_sync(0xFFFFFFFFu, v[3].w, 31) >> 24;
                const bool entering = (run_m >= 1u && b0 == 0u) || (run_m >= 2u && b0 <= 3u);
                if (!__any_sync(0xFFFFFFFFu, (pair & 0x80808080u) != 0u) && !entering) {
                    fastmask |= 0xFu << r;
                    run_m = (bl == 0u) ? 1u : 0u;
                    r += 4;
                    continue;
                }
            }
            uint4 v = make_uint4(0, 0, 0, 0);
            if (cpos < q1 && cpos + 16 > q0) { v = *reinterpret_cast<const uint4*>(sp + cpos); }
            if (lane == 0) { sm.run_in[warp][r] = run_m; }
            const RowInfo ri = insert_row<int>(v, row, q0, q1, lane, run_m);
            if (ri.fast) { fastmask |= 1u << r; }
            else { n_ins += __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(ri.ins)); }
            r++;
        }
        size = (q1 - q0) + (int)n_ins;
    } else if (kind == kRaw) {
        size = q1 - q0;
    }
    if (writes_prefix) { size += P.sc_len + P.len_size; }
    if (lane == 0) {
        sm.size[warp] = size;
        sm.meta[warp] = make_uint4((uint32_t)k, nbytes_raw, (first_item ? 1u : 0u) | (writes_prefix ? 2u : 0u), 0u);
    }
    __syncthreads();
    // ---- the tile's output offset: the aggregate goes to the scanner CTA, which answers with the exclusive prefix
    if (threadIdx.x == 0) {
        long long total = 0;
#pragma unroll
        for (int w = 0; w < kFWarps; w++) { total += sm.size[w]; }
        const int64_t t = sm.ticket - 1;
        st_relaxed_u64(&tile_state[t], (1ull << 62) | (unsigned long long)total);
        unsigned long long e;
        while (((e = ld_relaxed_u64(&tile_excl[t])) >> 62) == 0ull) { __nanosleep(32); }
        sm.excl = (long long)(e & ((1ull << 62) - 1ull));
    }
    __syncthreads();
    long long o = sm.excl;
    for (int w = 0; w < warp; w++) { o += sm.size[w]; }
    const uint4 meta = sm.meta[warp]; // (an inactive warp parked zeros: no NAL offset, no prefix, kind none, size 0)
    if ((meta.z & 1u) && lane == 0) { out_off[(int32_t)meta.x] = o; }
    if (n_ins && lane == 0) { atomicAdd(&hdr->n_ins, (unsigned long long)n_ins); }
    if (kind != kNone || (meta.z & 2u)) {
      if (o + sm.size[warp] <= out_cap) { // (capacity overflow is reported by the summary)
        if (meta.z & 2u) {
            if (lane < P.sc_len) { out[o + lane] = (lane == P.sc_len - 1) ? 1 : 0; }
            o += P.sc_len;
            if (P.len_size) { // big-endian length of what follows (verbatim NALs only: the launcher keeps escaped parts off this path)
                if (lane < P.len_size) { out[o + lane] = (uint8_t)((uint64_t)meta.y >> (8 * (P.len_size - 1 - lane))); }
                o += P.len_size;
            }
        }
        __syncwarp();
        if (kind == kRaw) {
            warp_copy_from_smem(out + o, sp + q0, (int64_t)(q1 - q0), lane);
        } else if (kind == kEsc) {
            int r = 0;
            while (r < rows) {
                if ((fastmask >> r) & 1u) { // a run of rows that take no insertion: one shifted vector copy
                    const int nrun = __ffs((int)~(fastmask >> r)) - 1; // (bit 31 of fastmask is never set: at most kPieceRows rows)
                    warp_copy_from_smem(out + o, sp + (r << 9), (int64_t)nrun << 9, lane);
                    o += (long long)nrun << 9;
                    r += nrun;
                    continue;
                }
                const int row = r << 9, cpos = row + lane * 16;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (cpos < q1 && cpos + 16 > q0) { v = *reinterpret_cast<const uint4*>(sp + cpos); }
                uint32_t rm = sm.run_in[warp][r];
                const RowInfo ri = insert_row<int>(v, row, q0, q1, lane, rm);
                o += write_row(ri, out + o, lane);
                r++;
            }
        }
      }
    }
    if (!persistent) { return; }
    __syncthreads(); // the pieces, sizes and the ticket word are rewritten by the next round
  }
}

__global__ void __launch_bounds__(256) fused_summary_kernel(const int64_t* __restrict__ first, int64_t n, int64_t cap_items, int64_t out_cap,
                                                            const FusedHeader* __restrict__ hdr, int64_t* __restrict__ out_off, hevcb_insert_summary* s)
{
    const int64_t n_items = (int64_t)((unsigned long long)first[n] >> kItemShift);
    const bool plain = n_items > cap_items; // more pieces than the table holds: only when the sources alone exceed out_cap
    if (plain) { // out_off = the sizes without insertions (a lower bound); the caller grows the buffer and runs again
        for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= n; k += (int64_t)gridDim.x * blockDim.x) {
            out_off[k] = (int64_t)((unsigned long long)first[k] & kBytesMask);
        }
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int64_t total = plain ? (int64_t)((unsigned long long)first[n] & kBytesMask) : out_off[n];
        s->n_nals = n;
        s->out_bytes = total;
        s->n_inserted = (int64_t)hdr->n_ins;
        s->overflow = (plain || total > out_cap) ? 1 : 0;
        s->pad = 0;
    }
}

} // namespace

static int launch_assemble_two_pass(hevcb_ctx* ctx, const AssembleParts& P, int64_t n, uint8_t* d_out, int64_t out_cap, int64_t* d_out_off,
                                    hevcb_insert_summary* d_summary, cudaStream_t stream)
{
    const int64_t nb = (n + kSTile - 1) / kSTile;
    const int64_t n1 = n > 0 ? n : 1;
    const int64_t capE = out_cap / kSplitSeg + 1024; // pieces of long parts (a piece is at least kSplitSeg bytes of output)
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_sizes = 0, o_bs = up((size_t)n1 * 8), o_misc = o_bs + up((size_t)(nb + 2) * 8), o_flag = o_misc + 256;
    const size_t o_e64 = o_flag + up((size_t)n1), o_e32 = o_e64 + up((size_t)capE * 8) * 5, need = o_e32 + up((size_t)capE * 4) * 2;
    int rc = hevcb_reserve(ctx, &ctx->insert_scratch, need);
    if (rc != HEVCB_OK) { return rc; }
    uint8_t* sb = reinterpret_cast<uint8_t*>(ctx->insert_scratch.p);
    int64_t* sizes = reinterpret_cast<int64_t*>(sb + o_sizes);
    long long* bs = reinterpret_cast<long long*>(sb + o_bs);
    unsigned long long* n_ins = reinterpret_cast<unsigned long long*>(sb + o_misc);
    SplitList L;
    L.n_extras = n_ins + 1;
    L.flag = sb + o_flag;
    L.e_off = reinterpret_cast<int64_t*>(sb + o_e64);
    L.e_end = reinterpret_cast<int64_t*>(sb + o_e64 + up((size_t)capE * 8));
    L.e_k = reinterpret_cast<int64_t*>(sb + o_e64 + up((size_t)capE * 8) * 2);
    L.e_size = reinterpret_cast<int64_t*>(sb + o_e64 + up((size_t)capE * 8) * 3);
    L.e_prefix = reinterpret_cast<int64_t*>(sb + o_e64 + up((size_t)capE * 8) * 4);
    L.e_j = reinterpret_cast<int32_t*>(sb + o_e32);
    L.e_np = reinterpret_cast<int32_t*>(sb + o_e32 + up((size_t)capE * 4));
    L.cap = capE;
    HEVCB_CUDA(ctx, cudaMemsetAsync(n_ins, 0, 16, stream));
    if (n == 0) {
        HEVCB_CUDA(ctx, cudaMemsetAsync(d_out_off, 0, 8, stream));
        HEVCB_CUDA(ctx, cudaMemsetAsync(d_summary, 0, sizeof(hevcb_insert_summary), stream));
        return HEVCB_OK;
    }
    HEVCB_CUDA(ctx, cudaMemsetAsync(L.flag, 0, (size_t)n, stream));
    long long grid = (long long)ctx->sm_count * 8;
    long long egrid = grid; // the side list's length is only known on the device: a full grid, idle when the list is empty
    const long long max_grid = ((n + 31) / 32 + kInsWarps - 1) / kInsWarps; // a warp takes groups of 32 consecutive NALs
    if (grid > max_grid) { grid = max_grid; }
    const long long max_egrid = (capE + kInsWarps - 1) / kInsWarps;
    if (egrid > max_egrid) { egrid = max_egrid; }
    split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(P, n, L);
    insert_count_kernel<<<(unsigned)grid, kInsThreads, 0, stream>>>(P, n, sizes, n_ins, L.flag);
    extras_count_kernel<<<(unsigned)egrid, kInsThreads, 0, stream>>>(P, L, n_ins);
    extras_prefix_kernel<<<(unsigned)((capE + 255) / 256 < 1024 ? (capE + 255) / 256 : 1024), 256, 0, stream>>>(L, sizes);
    sizes_reduce_kernel<<<(unsigned)nb, kSThreads, 0, stream>>>(sizes, n, bs);
    sizes_blocksums_kernel<<<1, kSThreads, 0, stream>>>(bs, nb);
    sizes_apply_kernel<<<(unsigned)nb, kSThreads, 0, stream>>>(sizes, n, bs, nb, d_out_off);
    insert_write_kernel<<<(unsigned)grid, kInsThreads, 0, stream>>>(P, n, d_out_off, d_out, out_cap, L.flag);
    extras_write_kernel<<<(unsigned)egrid, kInsThreads, 0, stream>>>(P, L, d_out_off, d_out, out_cap);
    insert_summary_kernel<<<1, 32, 0, stream>>>(d_out_off, n, out_cap, n_ins, d_summary);
    ctx->launches += 10;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}

// single-pass path (see the comment in front of fused_assemble_kernel)
static int launch_assemble_fused(hevcb_ctx* ctx, const AssembleParts& P, int64_t n, uint8_t* d_out, int64_t out_cap, int64_t* d_out_off,
                                 hevcb_insert_summary* d_summary, cudaStream_t stream)
{
    const int64_t nb = (n + kSTile - 1) / kSTile;
    // every part of every NAL takes at most len / kPieceCap + 2 items, a NAL without bytes one: enough whenever the sources fit into out_cap
    const int64_t n_parts = (P.raw_off ? 1 : 0) + (P.a_off ? 1 : 0) + (P.b_off ? 1 : 0);
    const int64_t cap_items = (2 * n_parts + 1) * n + out_cap / kPieceCap + 64;
    const int64_t cap_tiles = (cap_items + kFWarps - 1) / kFWarps;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_hdr = 0, o_state = 256, o_packed = o_state + up((size_t)cap_tiles * 8) * 2, o_first = o_packed + up((size_t)(n + 1) * 8);
    const size_t o_bs = o_first + up((size_t)(n + 2) * 8), o_items = o_bs + up((size_t)(nb + 2) * 8), need = o_items + up((size_t)cap_items * sizeof(ItemDesc));
    int rc = hevcb_reserve(ctx, &ctx->insert_scratch, need);
    if (rc != HEVCB_OK) { return rc; }
    uint8_t* sb = reinterpret_cast<uint8_t*>(ctx->insert_scratch.p);
    FusedHeader* hdr = reinterpret_cast<FusedHeader*>(sb + o_hdr);
    unsigned long long* states = reinterpret_cast<unsigned long long*>(sb + o_state);
    unsigned long long* excl = reinterpret_cast<unsigned long long*>(sb + o_state + up((size_t)cap_tiles * 8));
    int64_t* packed = reinterpret_cast<int64_t*>(sb + o_packed);
    int64_t* first = reinterpret_cast<int64_t*>(sb + o_first);
    long long* bs = reinterpret_cast<long long*>(sb + o_bs);
    ItemDesc* items = reinterpret_cast<ItemDesc*>(sb + o_items);
    HEVCB_CUDA(ctx, cudaMemsetAsync(sb, 0, o_packed, stream)); // header + tile states
    if (!ctx->fused_smem_set) {
        HEVCB_CUDA(ctx, cudaFuncSetAttribute(fused_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem)));
        ctx->fused_smem_set = 1;
    }
    const unsigned gn = (unsigned)((n + 255) / 256);
    long long fgrid = (long long)ctx->sm_count * 6; // persistent: six resident CTAs per SM (shared memory), tiles drawn by ticket
    if (fgrid > cap_tiles + 1) { fgrid = cap_tiles + 1; }
    int stagger_ns = 0, persistent = 1;
    if (const char* e = getenv("HEVCB_FUSED_PERSISTENT")) { persistent = atoi(e); }
    if (!persistent) { fgrid = cap_tiles + 1; }
    if (const char* e = getenv("HEVCB_FUSED_STAGGER")) { stagger_ns = atoi(e); }
    fused_plan_kernel<<<gn, 256, 0, stream>>>(P, n, packed);
    sizes_reduce_kernel<<<(unsigned)nb, kSThreads, 0, stream>>>(packed, n, bs);
    sizes_blocksums_kernel<<<1, kSThreads, 0, stream>>>(bs, nb);
    sizes_apply_kernel<<<(unsigned)nb, kSThreads, 0, stream>>>(packed, n, bs, nb, first);
    fused_fill_kernel<<<(unsigned)((cap_items + 255) / 256), 256, 0, stream>>>(P, first, n, items, cap_items);
    fused_assemble_kernel<<<(unsigned)fgrid, kFThreads, sizeof(FusedSmem), stream>>>(P, n, first, items, cap_items, states, excl, hdr, d_out_off, d_out, out_cap, stagger_ns, persistent);
    fused_summary_kernel<<<gn < 1024u ? gn : 1024u, 256, 0, stream>>>(first, n, cap_items, out_cap, hdr, d_out_off, d_summary);
    ctx->launches += 7;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}

// Which path: the single-pass kernel holds one piece per warp, so it pays off when NALs are not tiny (one warp-item per NAL part);
// batches of very short NALs (average below kFusedMinAvg output bytes) stay with the two-pass kernels.  HEVCB_INSERT_PATH=fused|twopass
// forces a path (tests run every case through both).
constexpr int64_t kFusedMinAvg = 4096;
static int launch_assemble(hevcb_ctx* ctx, const AssembleParts& P, int64_t n, uint8_t* d_out, int64_t out_cap, int64_t* d_out_off,
                           hevcb_insert_summary* d_summary, cudaStream_t stream)
{
    bool can_fuse = n > 0 && n < (1ll << 26) && out_cap < (1ll << 36) && !(P.len_size != 0 && (P.a_off || P.b_off)) && (((uintptr_t)P.raw_base & 15u) == 0);
    bool fused = can_fuse && out_cap / n >= kFusedMinAvg;
    if (const char* e = getenv("HEVCB_INSERT_PATH")) {
        if (e[0] == 'f') { fused = can_fuse; }
        else if (e[0] == 't') { fused = false; }
    }
    return fused ? launch_assemble_fused(ctx, P, n, d_out, out_cap, d_out_off, d_summary, stream)
                 : launch_assemble_two_pass(ctx, P, n, d_out, out_cap, d_out_off, d_summary, stream);
}

int hevcb_launch_insert(hevcb_ctx* ctx, const uint8_t* d_rbsp, const int64_t* d_off, const int64_t* d_end, int64_t n, int sc_len, uint8_t* d_out,
                        int64_t out_cap, int64_t* d_out_off, hevcb_insert_summary* d_summary, cudaStream_t stream)
{
    if (n < 0 || (sc_len != 0 && sc_len != 3 && sc_len != 4) || !d_out_off || !d_summary || (n > 0 && (!d_rbsp || !d_off || !d_end || !d_out))) {
        HEVCB_SET_ERR(ctx, "hevcb_insert: invalid argument");
        return HEVCB_E_ARG;
    }
    if ((uintptr_t)d_rbsp & 15u) {
        HEVCB_SET_ERR(ctx, "hevcb_insert: rbsp must be 16-byte aligned");
        return HEVCB_E_ALIGN;
    }
    AssembleParts P;
    P.raw_base = nullptr; P.raw_off = nullptr; P.raw_end = nullptr;
    P.a_base = nullptr; P.a_off = nullptr; P.a_end = nullptr;
    P.b_base = d_rbsp; P.b_off = d_off; P.b_end = d_end;
    P.sc_len = sc_len;
    P.skip_neg_b = 1;
    P.len_size = 0;
    return launch_assemble(ctx, P, n, d_out, out_cap, d_out_off, d_summary, stream);
}

// three-part assembly used by the header rewrite (hevcb_parse.cu): verbatim bytes, escaped header, escaped payload
int hevcb_launch_assemble3(hevcb_ctx* ctx, const uint8_t* raw_base, const int64_t* raw_off, const int64_t* raw_end, const uint8_t* a_base,
                           const int64_t* a_off, const int64_t* a_end, const uint8_t* b_base, const int64_t* b_off, const int64_t* b_end, int64_t n,
                           uint8_t* d_out, int64_t out_cap, int64_t* d_out_off, hevcb_insert_summary* d_summary, cudaStream_t stream)
{
    if (((uintptr_t)a_base & 15u) || ((uintptr_t)b_base & 15u)) {
        HEVCB_SET_ERR(ctx, "hevcb_assemble: escaped sources must be 16-byte aligned");
        return HEVCB_E_ALIGN;
    }
    AssembleParts P;
    P.raw_base = raw_base; P.raw_off = raw_off; P.raw_end = raw_end;
    P.a_base = a_base; P.a_off = a_off; P.a_end = a_end;
    P.b_base = b_base; P.b_off = b_off; P.b_end = b_end;
    P.sc_len = 0;
    P.skip_neg_b = 0;
    P.len_size = 0;
    return launch_assemble(ctx, P, n, d_out, out_cap, d_out_off, d_summary, stream);
}

// ---- length-prefixed framing (ISO/IEC 14496-15 sample format: hvcC / MP4 / Matroska) <-> Annex-B -----------------------------
// A NAL unit is the same bytes in both framings (emulation prevention included); only what separates the units differs: a
// start code in front (Annex-B) or a big-endian byte count of 1, 2 or 4 bytes (lengthSizeMinusOne + 1).  Both directions are
// "copy the NAL bytes verbatim behind a freshly written prefix": the verbatim part of the assembly above.
int hevcb_launch_frame(hevcb_ctx* ctx, const uint8_t* d_buf, const int64_t* d_nal_start, const int64_t* d_nal_end, int64_t n, int sc_len, int len_size,
                       uint8_t* d_out, int64_t out_cap, int64_t* d_out_off, hevcb_insert_summary* d_summary, cudaStream_t stream)
{
    if (n < 0 || !d_out_off || !d_summary || (n > 0 && (!d_buf || !d_nal_start || !d_nal_end || !d_out)) || (sc_len != 0 && sc_len != 3 && sc_len != 4) ||
        (len_size != 0 && len_size != 1 && len_size != 2 && len_size != 4) || (sc_len != 0 && len_size != 0)) {
        HEVCB_SET_ERR(ctx, "hevcb frame conversion: invalid argument");
        return HEVCB_E_ARG;
    }
    AssembleParts P;
    P.raw_base = d_buf; P.raw_off = d_nal_start; P.raw_end = d_nal_end;
    P.a_base = nullptr; P.a_off = nullptr; P.a_end = nullptr;
    P.b_base = nullptr; P.b_off = nullptr; P.b_end = nullptr;
    P.sc_len = sc_len;
    P.skip_neg_b = 0;
    P.len_size = len_size;
    return launch_assemble(ctx, P, n, d_out, out_cap, d_out_off, d_summary, stream);
}

namespace {
// Walks the length fields of one sample (a run of length-prefixed NAL units that ends at a known position: the sample table of the
// container knows it).  The fields chain -- every length tells where the next one is -- so a sample is walked by one thread; the
// samples are independent.  count pass: NALs per sample (negative: the chain runs past the sample's end); fill pass: extents.
template <bool kFill>
__global__ void __launch_bounds__(256) lenpref_walk_kernel(const uint8_t* __restrict__ buf, const int64_t* __restrict__ sample_off, int64_t n_samples, int64_t size,
                                                           int len_size, int64_t* __restrict__ count, const int64_t* __restrict__ first, int64_t* __restrict__ nal_start,
                                                           int64_t* __restrict__ nal_end, int64_t cap_nals, unsigned long long* __restrict__ n_bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) { return; }
    int64_t p = sample_off ? sample_off[i] : 0;
    int64_t end = sample_off ? sample_off[i + 1] : size;
    if (p < 0) { p = 0; }
    if (end > size) { end = size; }
    int64_t k = kFill ? first[i] : 0, n = 0;
    bool bad = false;
    while (p < end) {
        if (p + len_size > end) { bad = true; break; }
        uint64_t len = 0;
        for (int b = 0; b < len_size; b++) { len = (len << 8) | buf[p + b]; }
        p += len_size;
        if ((int64_t)len > end - p) { bad = true; break; }
        if (kFill && k < cap_nals) { nal_start[k] = p; nal_end[k] = p + (int64_t)len; }
        k++; n++;
        p += (int64_t)len;
    }
    if (!kFill) {
        count[i] = n;
        if (bad) { atomicAdd(n_bad, 1ull); }
    }
}
} // namespace

// length-prefixed samples -> NAL extents (d_nal_start / d_nal_end, positions of the NAL bytes inside d_buf); d_total: [0] NALs found,
// [1] samples whose length chain is broken (the NALs in front of the break are kept)
int hevcb_launch_lenpref_index(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, int len_size, const int64_t* d_sample_off, int64_t n_samples,
                               int64_t* d_nal_start, int64_t* d_nal_end, int64_t cap_nals, int64_t* d_total, cudaStream_t stream)
{
    if (size < 0 || n_samples < 0 || (len_size != 1 && len_size != 2 && len_size != 4) || !d_total || (size > 0 && !d_buf) ||
        (cap_nals > 0 && (!d_nal_start || !d_nal_end)) || cap_nals < 0) {
        HEVCB_SET_ERR(ctx, "hevcb_lenpref_index: invalid argument");
        return HEVCB_E_ARG;
    }
    const int64_t ns = d_sample_off ? n_samples : 1; // no sample table: the whole buffer is one sample (a serial walk)
    const int64_t nb = (ns + kSTile - 1) / kSTile;
    auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const size_t o_cnt = 0, o_first = up((size_t)(ns + 1) * 8), o_bs = o_first + up((size_t)(ns + 2) * 8), need = o_bs + up((size_t)(nb + 2) * 8) + 256;
    int rc = hevcb_reserve(ctx, &ctx->insert_scratch, need);
    if (rc != HEVCB_OK) { return rc; }
    uint8_t* sb = reinterpret_cast<uint8_t*>(ctx->insert_scratch.p);
    int64_t* cnt = reinterpret_cast<int64_t*>(sb + o_cnt);
    int64_t* first = reinterpret_cast<int64_t*>(sb + o_first);
    long long* bs = reinterpret_cast<long long*>(sb + o_bs);
    unsigned long long* n_bad = reinterpret_cast<unsigned long long*>(d_total + 1);
    HEVCB_CUDA(ctx, cudaMemsetAsync(d_total, 0, 16, stream));
    if (ns == 0) { return HEVCB_OK; }
    const unsigned g = (unsigned)((ns + 255) / 256);
    lenpref_walk_kernel<false><<<g, 256, 0, stream>>>(d_buf, d_sample_off, ns, size, len_size, cnt, nullptr, nullptr, nullptr, 0, n_bad);
    sizes_reduce_kernel<<<(unsigned)nb, kSThreads, 0, stream>>>(cnt, ns, bs);
    sizes_blocksums_kernel<<<1, kSThreads, 0, stream>>>(bs, nb);
    sizes_apply_kernel<<<(unsigned)nb, kSThreads, 0, stream>>>(cnt, ns, bs, nb, first);
    lenpref_walk_kernel<true><<<g, 256, 0, stream>>>(d_buf, d_sample_off, ns, size, len_size, nullptr, first, d_nal_start, d_nal_end, cap_nals, nullptr);
    HEVCB_CUDA(ctx, cudaMemcpyAsync(d_total, first + ns, 8, cudaMemcpyDeviceToDevice, stream));
    ctx->launches += 5;
    HEVCB_CUDA(ctx, cudaGetLastError());
    return HEVCB_OK;
}
