// hevcb_syntax.h -- bit reader and HEVC header syntax walker used by the batched parser kernels.
//
// Host/device code: compiled by nvcc into the sm_100a parser kernels (hevcb_parse.cu) and by g++ into the TEST-ONLY
// host build (tests/hostsim) that pins this logic against the reference without a GPU.  The product library only
// calls it from device code.
//
// Behaviour restated (reference file:line):
//   bit reader        bs.h:117-221   reads past the end yield 0 bits but still advance; ue() prefix loop stops at 32
//                                    zeros or when the consumed bit makes the cursor reach the end; overrun = byte
//                                    cursor strictly beyond the end
//   NAL dispatch      hevc_stream.c:155-241            VPS  :243-300     SPS :303-401 (+ :404-416)   PPS :419-521
//   slice header      hevc_stream.c:782-941 (+ :944-966 list modification, :969-1029 pred weight table)
//   PTL :652-755   scaling list :758-779   st_ref_pic_set :1032-1085   VUI :1088-1157   HRD :1160-1218
//   derived RPS state hevc_stream.in.c:26-113 (file-static tables in the reference; per-SPS tables here)
// The reference's deviations from the HEVC specification are reproduced on purpose.
//
// Every parsed syntax element is reported to a `Sink` as (field index, value), the field index being the offset in
// ints inside the struct the NAL writes (include/hevcb_layout.h).  Zero-filling that struct and scattering the pairs
// reproduces what read_hevc_nal_unit leaves in hevc_stream_t.
#pragma once
#include <stdint.h>

#include "../../include/hevcb.h"
#include "../../include/hevcb_layout.h"

#if defined(__CUDACC__)
#define HEVCB_SHD __host__ __device__
#else
#define HEVCB_SHD
#endif

// ------------------------------------------------------------------------------------------------
// bit reader: 64-bit MSB-first window over the RBSP bytes, refilled on demand
// ------------------------------------------------------------------------------------------------
struct hevcb_bits {
    const uint8_t* base; // first RBSP byte
    int64_t size;        // RBSP size in bytes
    int64_t pos;         // bit position
    int64_t wbyte;       // byte index of the first byte held in `win`
    uint64_t win;        // bytes [wbyte, wbyte + 8), big endian; bytes at or beyond `size` are 0

    HEVCB_SHD void init(const uint8_t* b, int64_t n)
    {
        base = b;
        size = n;
        pos = 0;
        wbyte = -16;
        win = 0;
    }
    // Loads bytes [byte, byte + 8) as a big-endian window, bytes at or beyond `size` read as 0.  Two aligned 8-byte loads
    // instead of eight byte loads: the buffer must be readable up to the next 8-byte boundary behind its last byte (every
    // image buffer of this library is; hevcb.h).
    HEVCB_SHD void refill(int64_t byte)
    {
        uint64_t v = 0;
        const int64_t valid = size - byte; // bytes of the window that exist
        if (byte >= 0 && valid > 0) {
            const uintptr_t p = (uintptr_t)(base + byte);
            const uintptr_t a = p & ~(uintptr_t)7;
            const uint32_t o = (uint32_t)(p - a) * 8u;
            const uint64_t lo = *reinterpret_cast<const uint64_t*>(a);
            uint64_t le = lo >> o;
            if (o != 0u && (int64_t)(8u - o / 8u) < valid) { le |= *reinterpret_cast<const uint64_t*>(a + 8) << (64u - o); }
#if defined(__CUDA_ARCH__)
            const uint32_t l32 = (uint32_t)le, h32 = (uint32_t)(le >> 32);
            v = ((uint64_t)__byte_perm(l32, 0u, 0x0123) << 32) | (uint64_t)__byte_perm(h32, 0u, 0x0123);
#else
            v = __builtin_bswap64(le);
#endif
            if (valid < 8) { v &= ~0ull << (8 * (8 - (int)valid)); }
        }
        win = v;
        wbyte = byte;
    }
    // next 32 bits at the cursor, without consuming them
    HEVCB_SHD uint32_t peek32()
    {
        const int64_t byte = pos >> 3;
        if (byte < wbyte || byte > wbyte + 3) { refill(byte); }
        const uint32_t sh = (uint32_t)(pos - (wbyte << 3)); // 0..31
        return (uint32_t)((win << sh) >> 32);
    }
    HEVCB_SHD uint32_t read_u(int n) // bs_read_u, n <= 32 (bs.h:160)
    {
        if (n <= 0) { return 0u; }
        if (n > 32) { pos += n - 32; n = 32; } // wider than 32 only on corrupt input (undefined shifts in the reference)
        const uint32_t v = peek32() >> (32 - n);
        pos += n;
        return v;
    }
    HEVCB_SHD void skip(int n) // bs_skip_u (bs.h:171): f(n, v) elements of the read variant, n may exceed 32
    {
        if (n > 0) { pos += n; }
    }
    HEVCB_SHD uint32_t read_u1() { return read_u(1); }
    HEVCB_SHD uint32_t read_u8() { return read_u(8); } // the FAST_U8 path reads the same bits (bs.h:182)
    HEVCB_SHD uint32_t read_ue() // bs.h:195-207
    {
        const int64_t endbits = size << 3;
        int64_t j_eof = endbits - 1 - pos; // first consumed bit that makes bs_eof() true
        if (j_eof < 0) { j_eof = 0; }
        const uint32_t w = peek32();
        uint32_t i = 32u;
        if (w != 0u) {
#if defined(__CUDA_ARCH__)
            i = (uint32_t)__clz((int)w);
#else
            i = (uint32_t)__builtin_clz(w);
#endif
        }
        if ((int64_t)i > j_eof) { i = (uint32_t)j_eof; }
        pos += (int64_t)i + 1; // the zeros and the bit that ended the loop
        uint32_t r = read_u((int)i);
        r += (i < 32u) ? ((1u << i) - 1u) : 0u; // 1 << 32 evaluates to 1 in the x86 reference
        return r;
    }
    HEVCB_SHD int32_t read_se() // bs.h:209-221
    {
        const int32_t r = (int32_t)read_ue();
        // (r + 1) / 2 in 32-bit arithmetic that wraps, as the x86 reference computes it for r = INT_MAX
        return (r & 1) ? (int32_t)((uint32_t)r + 1u) / 2 : -(r / 2);
    }
    HEVCB_SHD bool byte_aligned() const { return (pos & 7) == 0; }
    HEVCB_SHD int64_t byte_pos() const { return pos >> 3; }
    HEVCB_SHD bool overrun() const { return (pos >> 3) > size; } // bs.h:119
};

// ------------------------------------------------------------------------------------------------
// sinks
// ------------------------------------------------------------------------------------------------
// One sink type for both passes of the parser (count, then emit), so that both run the very same instantiation of
// the walker: with `field` null it only counts.
//
// kTraceT = true is the read_debug variant (read_debug_hevc_nal_unit, hevc_stream.c:2343-3436; format process.pl:90-113): the
// list then holds one record per element the reference PRINTS, in print order, each with the bit position of the reader
// before the read: struct members as (field index, value), the f(n, v) elements and the NAL header under
// HEVCB_TRACE_SPECIAL | id.  Values the reader stores without reading bits are kept too, marked HEVCB_TRACE_SILENT, so that
// the record list still materialises the struct.
enum hevcb_trace_id { // f(n, v) elements and the other lines of the dump that are not struct members
    HEVCB_TI_FORBIDDEN_ZERO_BIT = 1, HEVCB_TI_NAL_UNIT_TYPE, HEVCB_TI_NAL_LAYER_ID, HEVCB_TI_NAL_TEMPORAL_ID_PLUS1,
    HEVCB_TI_VPS_RESERVED_FFFF, HEVCB_TI_GENERAL_ZERO_34, HEVCB_TI_GENERAL_ZERO_43, HEVCB_TI_GENERAL_ZERO_BIT, HEVCB_TI_RESERVED_ZERO_XX,
    HEVCB_TI_SUB_LAYER_ZERO_34, HEVCB_TI_SUB_LAYER_ZERO_43, HEVCB_TI_SUB_LAYER_ZERO_BIT, HEVCB_TI_RBSP_STOP_ONE, HEVCB_TI_RBSP_ALIGN_ZERO,
    HEVCB_TI_ALIGN_ONE, HEVCB_TI_ALIGN_ZERO, HEVCB_TI_SLICE_RESERVED_FLAG, HEVCB_TI_SH_EXTENSION_DATA_BYTE,
    HEVCB_TI_OPEN_LINE, // (= 19) position prefix only, no newline: the "// ERROR" site of ref_pic_list_modification_flag_l1 (hevc_stream.c:3147)
    HEVCB_TI_FF_BYTE, // (= 20) filler data, extension mode
    HEVCB_TI_COUNT
};
template <bool kTraceT>
struct hevcb_sink_t {
    static constexpr bool kTrace = kTraceT;
    uint32_t* field;
    int32_t* value;
    uint32_t n;
    uint32_t* bitpos; // trace only
    HEVCB_SHD void put(uint32_t f, int32_t v)
    {
        if (field) {
            field[n] = kTraceT ? (f | HEVCB_TRACE_SILENT) : f;
            value[n] = v;
            if (kTraceT) { bitpos[n] = 0u; }
        }
        n++;
    }
    HEVCB_SHD void trace(int64_t pos, uint32_t code, int32_t v)
    {
        if (field) {
            field[n] = code;
            value[n] = v;
            bitpos[n] = pos > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)pos;
        }
        n++;
    }
};
typedef hevcb_sink_t<false> hevcb_sink;
typedef hevcb_sink_t<true> hevcb_trace_sink;
typedef hevcb_sink hevcb_count_sink;
typedef hevcb_sink hevcb_emit_sink;

// ------------------------------------------------------------------------------------------------
// parameter-set context kept on the device for the slices that follow
// ------------------------------------------------------------------------------------------------
#define HEVCB_RPS_SLOTS 33 // sets 0..31 of an SPS + the slice-local set at index num_short_term_ref_pic_sets

struct hevcb_rps_entry { // derived variables of one short-term RPS (hevc_stream.in.c:26-32)
    int32_t num_neg, num_pos, num_delta;
    uint32_t used_s0, used_s1; // UsedByCurrPicS0/S1 as bit masks
    int32_t dpoc_s0[32];
    int32_t dpoc_s1[32];
};

struct hevcb_sps_ctx {
    int32_t log2_min_cb_minus3, log2_diff_max_min_cb, pic_width, pic_height;
    int32_t separate_colour_plane_flag, chroma_format_idc, log2_max_poc_lsb_minus4;
    int32_t num_short_term_ref_pic_sets, long_term_ref_pics_present_flag, num_long_term_ref_pics_sps;
    int32_t sps_temporal_mvp_enabled_flag, sample_adaptive_offset_enabled_flag;
    uint32_t used_by_curr_pic_lt_sps_mask;
    int32_t seq_parameter_set_id; // spec mode: the id slices resolve the SPS by
    int32_t pad[2];
    hevcb_rps_entry rps[HEVCB_RPS_SLOTS];
};

struct hevcb_pps_ctx {
    int32_t seq_parameter_set_id, dependent_slice_segments_enabled_flag, output_flag_present_flag, num_extra_slice_header_bits;
    int32_t cabac_init_present_flag, num_ref_idx_l0_default_active_minus1, num_ref_idx_l1_default_active_minus1;
    int32_t pps_slice_chroma_qp_offsets_present_flag, weighted_pred_flag, weighted_bipred_flag, tiles_enabled_flag;
    int32_t entropy_coding_sync_enabled_flag, pps_loop_filter_across_slices_enabled_flag, deblocking_filter_override_enabled_flag;
    int32_t lists_modification_present_flag, slice_segment_header_extension_present_flag, chroma_qp_offset_list_enabled_flag;
    int32_t pic_parameter_set_id;                // spec mode: the id slices resolve the PPS by
    int32_t pps_deblocking_filter_disabled_flag; // spec mode: inherited by slice_deblocking_filter_disabled_flag
    int32_t pad[1];
};

// Spec-correct mode (HEVCB_PARSE_SPEC).  The default walk reproduces the reference, deviations from the HEVC syntax
// included; the switch turns exactly these into what the standard says (oracle: the reference's own template
// with the same fixes, oracle/make_spec_ref.py):
//   A-1  the SPS ends with rbsp_trailing_bits        A-2/3  a slice resolves its PPS by pic_parameter_set_id and that PPS's SPS by
//   seq_parameter_set_id (the most recent NAL with that id in front of the slice), derived RPS tables per SPS
//   A-4  ref_pic_list_modification_flag_l1 / list_entry_l1 are read      A-5  use_delta_flag inferred 1
//   A-6  PPS deblocking offsets present when the filter is NOT disabled   A-7  slice_deblocking_filter_disabled_flag inherits the PPS's
//   A-8  HRD: cpb_cnt_minus1 present when low_delay_hrd_flag is 0, fixed_pic_rate_within_cvs_flag inferred 1, cpb_cnt_minus1 + 1 entries
//        per sub-layer; VPS: cprms_present_flag[0] inferred 1
// The tables of a parse: entry j (1-based) = the j-th SPS / PPS NAL of the stream, entry 0 = the zeroed state before any.
struct hevcb_ps_lookup {
    const hevcb_sps_ctx* sps_tab;
    const hevcb_pps_ctx* pps_tab;
    int sps_count, pps_count; // SPS / PPS NALs in front of the slice
};

// per-slice columns written besides the (field, value) pairs
struct hevcb_slice_cols {
    int32_t slice_type, slice_qp_delta, slice_pic_order_cnt_lsb, first_slice_segment_in_pic_flag;
    int32_t slice_segment_address, dependent_slice_segment_flag, num_entry_point_offsets, short_term_ref_pic_set_idx;
};

HEVCB_SHD inline int hevcb_ceil_log2(int64_t n) // ceil(log2(n)) as the reference computes it in double; n <= 0 -> 0 bits
{
    if (n <= 1) { return 0; }
    int b = 0;
    uint64_t v = (uint64_t)(n - 1);
    while (v) { b++; v >>= 1; }
    return b;
}

#define HF(type, member) HEVCB_FIELD(type, member)

// ------------------------------------------------------------------------------------------------
// write side: bit writer (bs.h:224-331) and the source of the values a writer walks over
// ------------------------------------------------------------------------------------------------
struct hevcb_bitwriter {
    uint8_t* dst;  // null: count only
    int64_t cap;   // bytes that may be stored (bits beyond are dropped like bs_write_u1 past the end, bs.h:228)
    int64_t room;  // bytes of dst that exist (<= cap): a count pass may keep the first bytes in a small slot
    int64_t pos;   // bits written
    uint64_t acc;  // pending bits of the current byte (nacc < 8 between calls)
    int nacc;

    HEVCB_SHD void init(uint8_t* d, int64_t c) { dst = d; cap = c; room = c; pos = 0; acc = 0; nacc = 0; }
    HEVCB_SHD void init(uint8_t* d, int64_t c, int64_t r) { dst = d; cap = c; room = r < c ? r : c; pos = 0; acc = 0; nacc = 0; }
    HEVCB_SHD void put_bits(int n, uint32_t v) // 0 <= n <= 32: low n bits of v, MSB first
    {
        if (n <= 0) { return; }
        const uint64_t m = (n >= 32) ? 0xFFFFFFFFull : ((1ull << n) - 1ull);
        acc = (acc << n) | ((uint64_t)v & m);
        nacc += n;
        pos += n;
        while (nacc >= 8) {
            const int64_t byte = ((pos - nacc) >> 3);
            const uint8_t x = (uint8_t)(acc >> (nacc - 8));
            if (dst && byte < room) { dst[byte] = x; }
            nacc -= 8;
        }
        acc &= 0xFFull;
    }
    // bs_write_u (bs.h:240): bit i is (v >> (n - i - 1)) & 1; shift counts >= 32 wrap on the x86 reference
    HEVCB_SHD void write_u(int n, uint32_t v)
    {
        if (n <= 32) { put_bits(n, v); return; }
        for (int i = 0; i < n; i++) { put_bits(1, (v >> ((n - i - 1) & 31)) & 1u); }
    }
    HEVCB_SHD void write_ue(uint32_t v) // bs.h:264-319
    {
        if (v == 0u) { put_bits(1, 1u); return; }
        const uint32_t vv = v + 1u;
        int len = 1; // len_table[0] == 1: v == 0xFFFFFFFF wraps to 0
        if (vv != 0u) {
#if defined(__CUDA_ARCH__)
            len = 32 - __clz((int)vv);
#else
            len = 32 - __builtin_clz(vv);
#endif
        }
        write_u(2 * len - 1, vv);
    }
    HEVCB_SHD void write_se(int32_t v) // bs.h:321-331
    {
        if (v <= 0) { write_ue((uint32_t)(-v * 2)); } else { write_ue((uint32_t)(v * 2 - 1)); }
    }
    HEVCB_SHD bool byte_aligned() const { return (pos & 7) == 0; }
    HEVCB_SHD int64_t bytes() const { return pos >> 3; } // bs_pos: whole bytes only, a trailing partial byte is not counted
    HEVCB_SHD bool overrun() const { return (pos >> 3) > cap; } // bs.h:119: the byte cursor strictly beyond the end
};

// edits applied while rewriting: hevcb_edit_set (include/hevcb.h)
// The values of one parsed NAL, i.e. the struct read_hevc_nal_unit filled (zero + scatter of the pairs), consulted the
// way write_hevc_nal_unit reads that struct: field by field, last write wins, absent = 0.  The pairs are in syntax
// order, so the common case is a hit at the cursor.
struct hevcb_replay {
    const uint32_t* field;
    const int32_t* value;
    uint32_t n, cur;
    int32_t kind;
    const hevcb_edit_set* edits;
    const int32_t* dense; // alternative source: the struct itself as an int array (write_hevc_nal_unit from caller-owned structs)
    uint32_t run_f = 0xFFFFFFFFu, run_e = 0; // last run of a multiply-stored field: its field and the index of its last pair

    HEVCB_SHD int32_t load(uint32_t f, bool multi)
    {
        int32_t v = 0;
        if (dense) {
            v = dense[f];
        } else if (!multi && cur < n && field[cur] == f) {
            v = value[cur];
            cur++;
        } else if (multi && cur < n && field[cur] == f) {
            // a field the reader stored several times in a row: the struct holds the last value of the run; every call
            // consumes one pair of the run and returns that last value
            uint32_t e = run_e;
            if (run_f != f || cur > run_e) { // the end of the run is looked up once per run, not once per call (runs of 64 coefficients)
                e = cur;
                while (e + 1 < n && field[e + 1] == f) { e++; }
                run_f = f;
                run_e = e;
            }
            v = value[e];
            cur++;
        } else {
            for (int64_t i = (int64_t)n - 1; i >= 0; i--) {
                if (field[i] == f) {
                    v = value[i];
                    if (!multi) { cur = (uint32_t)i + 1u; }
                    break;
                }
            }
        }
        if (edits) {
            for (int i = 0; i < edits->n; i++) {
                const hevcb_edit_rule& e = edits->e[i];
                if (e.kind == kind && e.field == f) {
                    v = (e.op == HEVCB_EDIT_ADD) ? v + e.arg : (e.op == HEVCB_EDIT_SET) ? e.arg : (v ^ e.arg);
                }
            }
        }
        return v;
    }
    HEVCB_SHD void skip_if(uint32_t f) // a value the reader stored without reading bits
    {
        if (cur < n && field[cur] == f) { cur++; }
    }
};

// ------------------------------------------------------------------------------------------------
// the walker
// ------------------------------------------------------------------------------------------------
// kWrite = false: the read variant (read_hevc_*): bits -> (field, value) pairs in `s`.
// kWrite = true : the write variant (write_hevc_*, hevc_stream.c:1249-2312): the same walk with every value taken from
//                 `rp` (the struct the reader filled) and written to `bw`.  The two variants of the reference are
//                 generated from one template and differ only in what is listed at fx(), level8() and ovr() below.
template <class Sink, bool kWrite = false>
struct hevcb_walker {
    hevcb_bits& b;
    Sink& s;
    uint32_t flags; // bit0: an index ran past the reference's array bounds (undefined behaviour there)
    hevcb_bitwriter* bw;
    hevcb_replay* rp;
    bool spec = false;                    // spec-correct mode (HEVCB_PARSE_SPEC): the fixes listed at hevcb_ps_lookup below
    const hevcb_ps_lookup* lookup = nullptr; // spec mode: the parameter sets a slice may refer to, resolved by id

    HEVCB_SHD hevcb_walker(hevcb_bits& bb, Sink& ss) : b(bb), s(ss), flags(0), bw(nullptr), rp(nullptr) {}
    HEVCB_SHD hevcb_walker(hevcb_bits& bb, Sink& ss, hevcb_bitwriter* w, hevcb_replay* r) : b(bb), s(ss), flags(0), bw(w), rp(r) {}

    // Loops whose trip count comes out of the bitstream stop once the reader (writer) has run past its buffer: every further
    // element is 0 and the NAL fails anyway (rc -1), but a count of 2^31 read out of garbage -- a slice parsed against the wrong
    // parameter set, say -- would otherwise keep one GPU thread busy for minutes.  (The reference walks on and crashes there.)
    HEVCB_SHD bool runaway(int i) const
    {
        if constexpr (kWrite) { return i > 64 && bw->overrun(); } else { return i > 64 && b.overrun(); }
    }
    // one parsed element: (field, value) pair, or a trace record with the position the read started at
    HEVCB_SHD void rep(uint32_t f, int32_t v, int64_t p0)
    {
        if constexpr (Sink::kTrace) { s.trace(p0, f, v); } else { (void)p0; s.put(f, v); }
    }
    // value(x, u(n)) / u1 / u8 / ue / se : read + report, or load + write
    HEVCB_SHD int32_t u(uint32_t f, int n)
    {
        if constexpr (kWrite) { const int32_t v = rp->load(f, false); bw->write_u(n, (uint32_t)v); return v; }
        else { const int64_t p0 = b.pos; const int32_t v = (int32_t)b.read_u(n); rep(f, v, p0); return v; }
    }
    HEVCB_SHD int32_t u1(uint32_t f) { return u(f, 1); }
    HEVCB_SHD int32_t u8(uint32_t f) { return u(f, 8); }
    HEVCB_SHD int32_t ue(uint32_t f)
    {
        if constexpr (kWrite) { const int32_t v = rp->load(f, false); bw->write_ue((uint32_t)v); return v; }
        else { const int64_t p0 = b.pos; const int32_t v = (int32_t)b.read_ue(); rep(f, v, p0); return v; }
    }
    HEVCB_SHD int32_t se(uint32_t f)
    {
        if constexpr (kWrite) { const int32_t v = rp->load(f, false); bw->write_se(v); return v; }
        else { const int64_t p0 = b.pos; const int32_t v = b.read_se(); rep(f, v, p0); return v; }
    }
    // a field the reader stores several times (only the last value survives in the struct the writer reads)
    HEVCB_SHD int32_t se_multi(uint32_t f)
    {
        if constexpr (kWrite) { const int32_t v = rp->load(f, true); bw->write_se(v); return v; }
        else { return se(f); }
    }
    // array element: reported only inside the reference's bounds (the dump prints it either way)
    HEVCB_SHD int32_t au(uint32_t f, int idx, int bound, int n) { return raw_u(f + (uint32_t)idx, idx >= 0 && idx < bound, n); }
    HEVCB_SHD int32_t aue(uint32_t f, int idx, int bound)
    {
        const bool in = idx >= 0 && idx < bound;
        if constexpr (kWrite) { int32_t v = 0; if (in) { v = rp->load(f + (uint32_t)idx, false); } else { flags |= 1u; } bw->write_ue((uint32_t)v); return v; }
        else {
            const int64_t p0 = b.pos;
            const int32_t v = (int32_t)b.read_ue();
            if (in) { rep(f + (uint32_t)idx, v, p0); } else { flags |= 1u; }
            return v;
        }
    }
    HEVCB_SHD int32_t ase(uint32_t f, int idx, int bound) { return raw_se(f + (uint32_t)idx, idx >= 0 && idx < bound); }
    // element at an already computed field index; `in` = inside the reference's array
    HEVCB_SHD int32_t raw_u(uint32_t f, bool in, int n)
    {
        if constexpr (kWrite) { int32_t v = 0; if (in) { v = rp->load(f, false); } else { flags |= 1u; } bw->write_u(n, (uint32_t)v); return v; }
        else { const int64_t p0 = b.pos; const int32_t v = (int32_t)b.read_u(n); if (in) { rep(f, v, p0); } else { flags |= 1u; } return v; }
    }
    HEVCB_SHD int32_t raw_se(uint32_t f, bool in)
    {
        if constexpr (kWrite) { int32_t v = 0; if (in) { v = rp->load(f, false); } else { flags |= 1u; } bw->write_se(v); return v; }
        else { const int64_t p0 = b.pos; const int32_t v = b.read_se(); if (in) { rep(f, v, p0); } else { flags |= 1u; } return v; }
    }
    HEVCB_SHD void put_idx(uint32_t f, int idx, int bound, int32_t v, int64_t p0)
    {
        if (idx >= 0 && idx < bound) { rep(f + (uint32_t)idx, v, p0); } else { flags |= 1u; }
    }
    // f(n, v): bits the reader skips (bs_skip_u) and the writer emits as the constant v; the read_debug variant reads and
    // prints them under the element's name (id)
    HEVCB_SHD void fx(int n, uint32_t v, int id)
    {
        if constexpr (kWrite) { (void)id; bw->write_u(n, v); }
        else if constexpr (Sink::kTrace) {
            const int64_t p0 = b.pos;
            uint32_t r;
            if (n <= 32) { r = b.read_u(n); }
            else { r = 0u; for (int i = 0; i < n; i++) { r |= b.read_u1() << ((n - i - 1) & 31); } } // bs_read_u(b, 34 | 43): shift counts wrap on x86
            s.trace(p0, HEVCB_TRACE_SPECIAL | (uint32_t)id, (int32_t)r);
        }
        else { (void)id; b.skip(n); }
    }
    // a value the reader stores without reading bits (init_slice_hevc, defaults): nothing is written
    HEVCB_SHD void syn(uint32_t f, int32_t v)
    {
        if constexpr (kWrite) { rp->skip_if(f); } else { s.put(f, v); }
    }
    // num_ref_idx_l{0,1}_active_minus1: write_hevc_slice_header first overwrites the struct member with the PPS default
    // (hevc_stream.c:1898-1899) and then writes THAT value where the reader parses the override (:1965-1967)
    HEVCB_SHD int32_t ovr(uint32_t f, int32_t dflt)
    {
        if constexpr (kWrite) { rp->skip_if(f); bw->write_ue((uint32_t)dflt); return dflt; }
        else { return ue(f); }
    }
    // sub_layer_level_idc: read as u(8) by read_hevc_*, with bs_read_u1 / bs_write_u1 by the generated read_debug and write
    // variants (hevc_stream.c:751 vs :2939 and :1845)
    HEVCB_SHD int32_t level8(uint32_t f, int idx, int bound) { return au(f, idx, bound, (kWrite || Sink::kTrace) ? 1 : 8); }
    HEVCB_SHD bool aligned() const
    {
        if constexpr (kWrite) { return bw->byte_aligned(); } else { return b.byte_aligned(); }
    }

    // 7.3.2.11 / 7.3.2.12: a one bit, then zero bits up to the byte boundary (hevc_stream.c:630-649 / :1724-1743)
    HEVCB_SHD void trailing_bits(bool byte_alignment = false)
    {
        fx(1, 1u, byte_alignment ? HEVCB_TI_ALIGN_ONE : HEVCB_TI_RBSP_STOP_ONE);
        while (!aligned()) { fx(1, 0u, byte_alignment ? HEVCB_TI_ALIGN_ZERO : HEVCB_TI_RBSP_ALIGN_ZERO); }
    }

    // ---- 7.3.3 profile_tier_level (hevc_stream.c:652-755) --------------------------------------
    HEVCB_SHD void profile_tier_level(uint32_t base, int max_sub_layers_minus1)
    {
        typedef hevc_profile_tier_level_t P;
        u(base + HF(P, general_profile_space), 2);
        u1(base + HF(P, general_tier_flag));
        const int idc = u(base + HF(P, general_profile_idc), 5);
        uint32_t compat = 0;
        for (int i = 0; i < 32; i++) { compat |= (uint32_t)u1(base + HF(P, general_profile_compatibility_flag) + i) << i; }
        u1(base + HF(P, general_progressive_source_flag));
        u1(base + HF(P, general_interlaced_source_flag));
        u1(base + HF(P, general_non_packed_constraint_flag));
        u1(base + HF(P, general_frame_only_constraint_flag));
        if (idc == 4 || ((compat >> 4) & 1) || idc == 5 || ((compat >> 5) & 1) || idc == 6 || ((compat >> 6) & 1) || idc == 7 || ((compat >> 7) & 1)) {
            u1(base + HF(P, general_max_12bit_constraint_flag));
            u1(base + HF(P, general_max_10bit_constraint_flag));
            u1(base + HF(P, general_max_8bit_constraint_flag));
            u1(base + HF(P, general_max_422chroma_constraint_flag));
            u1(base + HF(P, general_max_420chroma_constraint_flag));
            u1(base + HF(P, general_max_monochrome_constraint_flag));
            u1(base + HF(P, general_intra_constraint_flag));
            u1(base + HF(P, general_one_picture_only_constraint_flag));
            u1(base + HF(P, general_lower_bit_rate_constraint_flag));
            fx(34, 0u, HEVCB_TI_GENERAL_ZERO_34);
        } else {
            fx(43, 0u, HEVCB_TI_GENERAL_ZERO_43);
        }
        if ((idc >= 1 && idc <= 5) || ((compat >> 1) & 1) || ((compat >> 2) & 1) || ((compat >> 3) & 1) || ((compat >> 4) & 1) || ((compat >> 5) & 1)) {
            u1(base + HF(P, general_inbld_flag));
        } else {
            fx(1, 0u, HEVCB_TI_GENERAL_ZERO_BIT);
        }
        u8(base + HF(P, general_level_idc));
        uint32_t prof_present = 0, level_present = 0;
        for (int i = 0; i < max_sub_layers_minus1; i++) {
            prof_present |= (uint32_t)(au(base + HF(P, sub_layer_profile_present_flag), i, HEVCB_MAX_SUBLAYERS, 1) & 1) << (i & 31);
            level_present |= (uint32_t)(au(base + HF(P, sub_layer_level_present_flag), i, HEVCB_MAX_SUBLAYERS, 1) & 1) << (i & 31);
        }
        if (max_sub_layers_minus1 > 0) {
            for (int i = max_sub_layers_minus1; i < 8; i++) { fx(2, 0u, HEVCB_TI_RESERVED_ZERO_XX); }
        }
        for (int i = 0; i < max_sub_layers_minus1; i++) {
            if ((prof_present >> (i & 31)) & 1u) {
                au(base + HF(P, sub_layer_profile_space), i, HEVCB_MAX_SUBLAYERS, 2);
                au(base + HF(P, sub_layer_tier_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                const int sidc = au(base + HF(P, sub_layer_profile_idc), i, HEVCB_MAX_SUBLAYERS, 5);
                uint32_t sc = 0;
                for (int j = 0; j < 32; j++) {
                    const int32_t v = raw_u(base + HF(P, sub_layer_profile_compatibility_flag) + (uint32_t)(i * 32 + j), i < HEVCB_MAX_SUBLAYERS, 1);
                    sc |= (uint32_t)(v & 1) << j;
                }
                au(base + HF(P, sub_layer_progressive_source_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                au(base + HF(P, sub_layer_interlaced_source_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                au(base + HF(P, sub_layer_non_packed_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                au(base + HF(P, sub_layer_frame_only_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                if (sidc == 4 || ((sc >> 4) & 1) || sidc == 5 || ((sc >> 5) & 1) || sidc == 6 || ((sc >> 6) & 1) || sidc == 7 || ((sc >> 7) & 1)) {
                    au(base + HF(P, sub_layer_max_12bit_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                    au(base + HF(P, sub_layer_max_10bit_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                    au(base + HF(P, sub_layer_max_8bit_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                    au(base + HF(P, sub_layer_max_422chroma_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                    au(base + HF(P, sub_layer_max_420chroma_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                    au(base + HF(P, sub_layer_max_monochrome_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                    au(base + HF(P, sub_layer_intra_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                    au(base + HF(P, sub_layer_one_picture_only_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                    au(base + HF(P, sub_layer_lower_bit_rate_constraint_flag), i, HEVCB_MAX_SUBLAYERS, 1);
                    fx(34, 0u, HEVCB_TI_SUB_LAYER_ZERO_34);
                } else {
                    fx(43, 0u, HEVCB_TI_SUB_LAYER_ZERO_43);
                }
                // the reference tests the ADDRESS of sub_layer_profile_compatibility_flag[1] (always true): App. A-10
                au(base + HF(P, sub_layer_inbld_flag), i, HEVCB_MAX_SUBLAYERS, 1);
            }
            if ((level_present >> (i & 31)) & 1u) { level8(base + HF(P, sub_layer_level_idc), i, HEVCB_MAX_SUBLAYERS); }
        }
    }

    // ---- E.2.3 sub-layer HRD (hevc_stream.c:1207-1218): CpbCnt + 1 entries (App. A-8) -----------
    HEVCB_SHD void sub_layer_hrd(uint32_t base, int cpb_cnt, int sub_pic)
    {
        typedef hevc_sub_layer_hrd_t H;
        for (int i = 0; i < cpb_cnt + (spec ? 0 : 1); i++) { if (runaway(i)) { break; } // the reference walks one entry more than cpb_cnt_minus1 + 1 (App. A-8)
            aue(base + HF(H, bit_rate_value_minus1), i, HEVCB_MAX_CPB_CNT);
            aue(base + HF(H, cpb_size_value_minus1), i, HEVCB_MAX_CPB_CNT);
            if (sub_pic) {
                aue(base + HF(H, cpb_size_du_value_minus1), i, HEVCB_MAX_CPB_CNT);
                aue(base + HF(H, bit_rate_du_value_minus1), i, HEVCB_MAX_CPB_CNT);
            }
            au(base + HF(H, cbr_flag), i, HEVCB_MAX_CPB_CNT, 1);
        }
    }

    // ---- E.2.2 HRD parameters (hevc_stream.c:1160-1204) ------------------------------------------
    HEVCB_SHD void hrd_parameters(uint32_t base, int common_inf_present, int max_sub_layers_minus1)
    {
        typedef hevc_hrd_t H;
        int nal = 0, vcl = 0, sub_pic = 0; // the struct is zeroed by the enclosing memset: absent flags read as 0
        if (common_inf_present) {
            nal = u1(base + HF(H, nal_hrd_parameters_present_flag));
            vcl = u1(base + HF(H, vcl_hrd_parameters_present_flag));
            if (nal || vcl) {
                sub_pic = u1(base + HF(H, sub_pic_hrd_params_present_flag));
                if (sub_pic) {
                    u8(base + HF(H, tick_divisor_minus2));
                    u(base + HF(H, du_cpb_removal_delay_increment_length_minus1), 5);
                    u1(base + HF(H, sub_pic_cpb_params_in_pic_timing_sei_flag));
                    u(base + HF(H, dpb_output_delay_du_length_minus1), 5);
                }
                u(base + HF(H, bit_rate_scale), 4);
                u(base + HF(H, cpb_size_scale), 4);
                if (sub_pic) { u(base + HF(H, cpb_size_du_scale), 4); }
                u(base + HF(H, initial_cpb_removal_delay_length_minus1), 5);
                u(base + HF(H, au_cpb_removal_delay_length_minus1), 5);
                u(base + HF(H, dpb_output_delay_length_minus1), 5);
            }
        }
        for (int i = 0; i <= max_sub_layers_minus1; i++) {
            const int general = au(base + HF(H, fixed_pic_rate_general_flag), i, HEVCB_MAX_SUBLAYERS, 1);
            int within = 0, low_delay = 0, cpb_cnt_minus1 = 0;
            if (!general) { within = au(base + HF(H, fixed_pic_rate_within_cvs_flag), i, HEVCB_MAX_SUBLAYERS, 1); }
            else if (spec) { within = 1; if (i < HEVCB_MAX_SUBLAYERS) { syn(base + HF(H, fixed_pic_rate_within_cvs_flag) + (uint32_t)i, 1); } } // inferred
            if (within) { aue(base + HF(H, elemental_duration_in_tc_minus1), i, HEVCB_MAX_SUBLAYERS); }
            else { low_delay = au(base + HF(H, low_delay_hrd_flag), i, HEVCB_MAX_SUBLAYERS, 1); }
            if (spec ? !low_delay : low_delay) { cpb_cnt_minus1 = aue(base + HF(H, cpb_cnt_minus1), i, HEVCB_MAX_SUBLAYERS); } // the reference reads it when SET (App. A-8)
            const uint32_t sl = (uint32_t)(sizeof(hevc_sub_layer_hrd_t) / sizeof(int));
            if (nal) { if (i < HEVCB_MAX_SUBLAYERS) { sub_layer_hrd(base + HF(H, sub_layer_hrd_nal) + (uint32_t)i * sl, cpb_cnt_minus1 + 1, sub_pic); } }
            if (vcl) { if (i < HEVCB_MAX_SUBLAYERS) { sub_layer_hrd(base + HF(H, sub_layer_hrd_vcl) + (uint32_t)i * sl, cpb_cnt_minus1 + 1, sub_pic); } }
        }
    }

    // ---- 7.3.4 scaling list data (hevc_stream.c:758-779): every delta coef lands in [sizeId][matrixId] (App. A-9)
    HEVCB_SHD void scaling_list_data(uint32_t base)
    {
        typedef hevc_scaling_list_data_t L;
        for (int size_id = 0; size_id < 4; size_id++) {
            for (int matrix_id = 0; matrix_id < 6; matrix_id += (size_id == 3) ? 3 : 1) {
                const int mode = u1(base + HF(L, scaling_list_pred_mode_flag) + (uint32_t)(size_id * 6 + matrix_id));
                if (!mode) {
                    ue(base + HF(L, scaling_list_pred_matrix_id_delta) + (uint32_t)(size_id * 6 + matrix_id));
                } else {
                    int coef_num = 1 << (4 + (size_id << 1));
                    if (coef_num > 64) { coef_num = 64; }
                    if (size_id > 1) { se(base + HF(L, scaling_list_dc_coef_minus8) + (uint32_t)((size_id - 2) * 6 + matrix_id)); }
                    for (int i = 0; i < coef_num; i++) { se_multi(base + HF(L, scaling_list_delta_coef) + (uint32_t)(size_id * 64 + matrix_id)); }
                }
            }
        }
    }

    // ---- 7.3.7 st_ref_pic_set + derived variables (hevc_stream.c:1032-1085, hevc_stream.in.c:61-113) ----------
    // `tbl` = derived entries of the enclosing SPS (read for RefRpsIdx), `cur` = entry being produced (index idx).
    HEVCB_SHD void st_ref_pic_set(uint32_t base, int idx, int num_sets, const hevcb_rps_entry* tbl, hevcb_rps_entry& cur)
    {
        typedef hevc_st_ref_pic_set_t R;
        int inter = 0;
        if (idx != 0) { inter = u1(base + HF(R, inter_ref_pic_set_prediction_flag)); }
        if (inter) {
            int delta_idx_minus1 = 0;
            if (idx == num_sets) { delta_idx_minus1 = ue(base + HF(R, delta_idx_minus1)); }
            const int sign = u1(base + HF(R, delta_rps_sign));
            const int abs_minus1 = ue(base + HF(R, abs_delta_rps_minus1));
            int ref_idx = idx - (delta_idx_minus1 + 1);
            if (ref_idx < 0 || ref_idx >= HEVCB_RPS_SLOTS) { flags |= 1u; ref_idx = 0; } // out of bounds in the reference
            const hevcb_rps_entry& ref = tbl[ref_idx];
            uint64_t used = 0, use_delta = 0; // use_delta_flag stays 0 when used_by_curr_pic_flag is 1 (App. A-5)
            for (int j = 0; j <= ref.num_delta; j++) { if (runaway(j)) { break; }
                const int ub = au(base + HF(R, used_by_curr_pic_flag), j, HEVCB_MAX_PICS, 1);
                if (j < 64) { used |= (uint64_t)(ub & 1) << j; }
                if (!ub) {
                    const int ud = au(base + HF(R, use_delta_flag), j, HEVCB_MAX_PICS, 1);
                    if (j < 64) { use_delta |= (uint64_t)(ud & 1) << j; }
                } else if (spec) { // inferred 1
                    if (j < HEVCB_MAX_PICS) { syn(base + HF(R, use_delta_flag) + (uint32_t)j, 1); }
                    if (j < 64) { use_delta |= 1ull << j; }
                }
            }
            // updateNumDeltaPocs, inter case
            const int delta_rps = (1 - 2 * sign) * (abs_minus1 + 1);
            int i = 0;
            uint32_t us0 = 0, us1 = 0;
            int32_t d0[32], d1[32];
            for (int j = ref.num_pos - 1; j >= 0; j--) {
                const int jj = j & 31;
                const int dpoc = ref.dpoc_s1[jj] + delta_rps;
                const int k = ref.num_neg + j;
                if (dpoc < 0 && k < 64 && ((use_delta >> k) & 1)) { if (i < 32) { d0[i] = dpoc; us0 |= (uint32_t)((used >> k) & 1) << i; } i++; }
            }
            if (delta_rps < 0 && ref.num_delta < 64 && ((use_delta >> ref.num_delta) & 1)) {
                if (i < 32) { d0[i] = delta_rps; us0 |= (uint32_t)((used >> ref.num_delta) & 1) << i; }
                i++;
            }
            for (int j = 0; j < ref.num_neg; j++) {
                const int dpoc = ref.dpoc_s0[j & 31] + delta_rps;
                if (dpoc < 0 && j < 64 && ((use_delta >> j) & 1)) { if (i < 32) { d0[i] = dpoc; us0 |= (uint32_t)((used >> j) & 1) << i; } i++; }
            }
            const int nneg = i;
            i = 0;
            for (int j = ref.num_neg - 1; j >= 0; j--) {
                const int dpoc = ref.dpoc_s0[j & 31] + delta_rps;
                if (dpoc > 0 && j < 64 && ((use_delta >> j) & 1)) { if (i < 32) { d1[i] = dpoc; us1 |= (uint32_t)((used >> j) & 1) << i; } i++; }
            }
            if (delta_rps > 0 && ref.num_delta < 64 && ((use_delta >> ref.num_delta) & 1)) {
                if (i < 32) { d1[i] = delta_rps; us1 |= (uint32_t)((used >> ref.num_delta) & 1) << i; }
                i++;
            }
            for (int j = 0; j < ref.num_pos; j++) {
                const int dpoc = ref.dpoc_s1[j & 31] + delta_rps;
                const int k = ref.num_neg + j;
                if (dpoc > 0 && k < 64 && ((use_delta >> k) & 1)) { if (i < 32) { d1[i] = dpoc; us1 |= (uint32_t)((used >> k) & 1) << i; } i++; }
            }
            const int npos = i;
            if (nneg > 32 || npos > 32) { flags |= 1u; }
            // entries the reference does not overwrite keep their previous values: carry them over from `cur`
            for (int j = 0; j < 32; j++) {
                if (j < nneg) { cur.dpoc_s0[j] = d0[j]; cur.used_s0 = (cur.used_s0 & ~(1u << j)) | (us0 & (1u << j)); }
                if (j < npos) { cur.dpoc_s1[j] = d1[j]; cur.used_s1 = (cur.used_s1 & ~(1u << j)) | (us1 & (1u << j)); }
            }
            cur.num_neg = nneg;
            cur.num_pos = npos;
        } else {
            const int nneg = ue(base + HF(R, num_negative_pics));
            const int npos = ue(base + HF(R, num_positive_pics));
            int32_t acc = 0;
            for (int i = 0; i < nneg; i++) { if (runaway(i)) { break; }
                const int d = aue(base + HF(R, delta_poc_s0_minus1), i, HEVCB_MAX_PICS);
                const int ub = au(base + HF(R, used_by_curr_pic_s0_flag), i, HEVCB_MAX_PICS, 1);
                acc = (i == 0) ? -(d + 1) : acc - (d + 1);
                if (i < 32) { cur.dpoc_s0[i] = acc; cur.used_s0 = (cur.used_s0 & ~(1u << i)) | ((uint32_t)(ub & 1) << i); }
            }
            acc = 0;
            for (int i = 0; i < npos; i++) { if (runaway(i)) { break; }
                const int d = aue(base + HF(R, delta_poc_s1_minus1), i, HEVCB_MAX_PICS);
                const int ub = au(base + HF(R, used_by_curr_pic_s1_flag), i, HEVCB_MAX_PICS, 1);
                acc = (i == 0) ? (d + 1) : acc + (d + 1);
                if (i < 32) { cur.dpoc_s1[i] = acc; cur.used_s1 = (cur.used_s1 & ~(1u << i)) | ((uint32_t)(ub & 1) << i); }
            }
            cur.num_neg = nneg;
            cur.num_pos = npos;
        }
        cur.num_delta = cur.num_neg + cur.num_pos;
    }

    // ---- E.2.1 VUI (hevc_stream.c:1088-1157) -----------------------------------------------------
    HEVCB_SHD void vui_parameters(uint32_t base, int sps_max_sub_layers_minus1)
    {
        typedef hevc_vui_t V;
        if (u1(base + HF(V, aspect_ratio_info_present_flag))) {
            if (u8(base + HF(V, aspect_ratio_idc)) == 255) { // SAR_Extended
                u(base + HF(V, sar_width), 16);
                u(base + HF(V, sar_height), 16);
            }
        }
        if (u1(base + HF(V, overscan_info_present_flag))) { u1(base + HF(V, overscan_appropriate_flag)); }
        if (u1(base + HF(V, video_signal_type_present_flag))) {
            u(base + HF(V, video_format), 3);
            u1(base + HF(V, video_full_range_flag));
            if (u1(base + HF(V, colour_description_present_flag))) {
                u8(base + HF(V, colour_primaries));
                u8(base + HF(V, transfer_characteristics));
                u8(base + HF(V, matrix_coefficients));
            }
        }
        if (u1(base + HF(V, chroma_loc_info_present_flag))) {
            ue(base + HF(V, chroma_sample_loc_type_top_field));
            ue(base + HF(V, chroma_sample_loc_type_bottom_field));
        }
        u1(base + HF(V, neutral_chroma_indication_flag));
        u1(base + HF(V, field_seq_flag));
        u1(base + HF(V, frame_field_info_present_flag));
        if (u1(base + HF(V, default_display_window_flag))) {
            ue(base + HF(V, def_disp_win_left_offset));
            ue(base + HF(V, def_disp_win_right_offset));
            ue(base + HF(V, def_disp_win_top_offset));
            ue(base + HF(V, def_disp_win_bottom_offset));
        }
        if (u1(base + HF(V, vui_timing_info_present_flag))) {
            u(base + HF(V, vui_num_units_in_tick), 32);
            u(base + HF(V, vui_time_scale), 32);
            if (u1(base + HF(V, vui_poc_proportional_to_timing_flag))) { ue(base + HF(V, vui_num_ticks_poc_diff_one_minus1)); }
            if (u1(base + HF(V, vui_hrd_parameters_present_flag))) { hrd_parameters(base + HF(V, hrd), 1, sps_max_sub_layers_minus1); }
        }
        if (u1(base + HF(V, bitstream_restriction_flag))) {
            u1(base + HF(V, tiles_fixed_structure_flag));
            u1(base + HF(V, motion_vectors_over_pic_boundaries_flag));
            u1(base + HF(V, restricted_ref_pic_lists_flag));
            ue(base + HF(V, min_spatial_segmentation_idc));
            ue(base + HF(V, max_bytes_per_pic_denom));
            ue(base + HF(V, max_bits_per_min_cu_denom));
            ue(base + HF(V, log2_max_mv_length_horizontal));
            ue(base + HF(V, log2_max_mv_length_vertical));
        }
    }

    // ---- 7.3.2.1 VPS (hevc_stream.c:243-300) -----------------------------------------------------
    HEVCB_SHD void video_parameter_set()
    {
        typedef hevc_vps_t V;
        u(HF(V, vps_video_parameter_set_id), 4);
        u1(HF(V, vps_base_layer_internal_flag));
        u1(HF(V, vps_base_layer_available_flag));
        u(HF(V, vps_max_layers_minus1), 6);
        const int msl = u(HF(V, vps_max_sub_layers_minus1), 3);
        u1(HF(V, vps_temporal_id_nesting_flag));
        fx(16, 0xFFFFu, HEVCB_TI_VPS_RESERVED_FFFF);
        profile_tier_level(HF(V, ptl), msl);
        const int ordering = u1(HF(V, vps_sub_layer_ordering_info_present_flag));
        for (int i = (ordering ? 0 : msl); i <= msl; i++) {
            aue(HF(V, vps_max_dec_pic_buffering_minus1), i, HEVCB_MAX_SUBLAYERS);
            aue(HF(V, vps_max_num_reorder_pics), i, HEVCB_MAX_SUBLAYERS);
            aue(HF(V, vps_max_latency_increase_plus1), i, HEVCB_MAX_SUBLAYERS);
        }
        const int max_layer_id = u(HF(V, vps_max_layer_id), 6);
        const int num_layer_sets_minus1 = ue(HF(V, vps_num_layer_sets_minus1));
        for (int i = 1; i <= num_layer_sets_minus1; i++) {
            for (int j = 0; j <= max_layer_id; j++) { if (runaway(j)) { break; }
                raw_u(HF(V, layer_id_included_flag) + (uint32_t)(i * HEVCB_MAX_SUBLAYERS + j), i < HEVCB_MAX_SUBLAYERS && j < HEVCB_MAX_SUBLAYERS, 1);
            }
        }
        if (u1(HF(V, vps_timing_info_present_flag))) {
            u(HF(V, vps_num_units_in_tick), 32);
            u(HF(V, vps_time_scale), 32);
            if (u1(HF(V, vps_poc_proportional_to_timing_flag))) { ue(HF(V, vps_num_ticks_poc_diff_one_minus1)); }
            const int num_hrd = ue(HF(V, vps_num_hrd_parameters));
            for (int i = 0; i < num_hrd; i++) { if (runaway(i)) { break; }
                aue(HF(V, hrd_layer_set_idx), i, HEVCB_MAX_HRD_PARAM);
                int cprms = 0; // cprms_present_flag[0] is neither read nor inferred (App. A-8)
                if (i > 0) { cprms = au(HF(V, cprms_present_flag), i, HEVCB_MAX_HRD_PARAM, 1); }
                else if (spec) { cprms = 1; syn(HF(V, cprms_present_flag), 1); } // inferred
                const uint32_t hs = (uint32_t)(sizeof(hevc_hrd_t) / sizeof(int));
                if (i < HEVCB_MAX_HRD_PARAM) { hrd_parameters(HF(V, hrd) + (uint32_t)i * hs, cprms, msl); }
                else if constexpr (kWrite) { flags |= 1u; }
                else { flags |= 1u; hevcb_sink dummy{nullptr, nullptr, 0}; hevcb_walker<hevcb_count_sink> w(b, dummy); w.hrd_parameters(0, cprms, msl); }
            }
        }
        u1(HF(V, vps_extension_flag));
        trailing_bits();
    }

    // ---- 7.3.2.2 SPS (hevc_stream.c:303-416): no rbsp_trailing_bits (App. A-1) -------------------
    HEVCB_SHD void seq_parameter_set(hevcb_sps_ctx& c)
    {
        typedef hevc_sps_t S;
        u(HF(S, sps_video_parameter_set_id), 4);
        const int msl = u(HF(S, sps_max_sub_layers_minus1), 3);
        u1(HF(S, sps_temporal_id_nesting_flag));
        profile_tier_level(HF(S, ptl), msl);
        c.seq_parameter_set_id = ue(HF(S, sps_seq_parameter_set_id));
        c.chroma_format_idc = ue(HF(S, chroma_format_idc));
        c.separate_colour_plane_flag = 0;
        if (c.chroma_format_idc == 3) { c.separate_colour_plane_flag = u1(HF(S, separate_colour_plane_flag)); }
        c.pic_width = ue(HF(S, pic_width_in_luma_samples));
        c.pic_height = ue(HF(S, pic_height_in_luma_samples));
        if (u1(HF(S, conformance_window_flag))) {
            ue(HF(S, conf_win_left_offset));
            ue(HF(S, conf_win_right_offset));
            ue(HF(S, conf_win_top_offset));
            ue(HF(S, conf_win_bottom_offset));
        }
        ue(HF(S, bit_depth_luma_minus8));
        ue(HF(S, bit_depth_chroma_minus8));
        c.log2_max_poc_lsb_minus4 = ue(HF(S, log2_max_pic_order_cnt_lsb_minus4));
        const int ordering = u1(HF(S, sps_sub_layer_ordering_info_present_flag));
        for (int i = (ordering ? 0 : msl); i <= msl; i++) {
            aue(HF(S, sps_max_dec_pic_buffering_minus1), i, HEVCB_MAX_SUBLAYERS);
            aue(HF(S, sps_max_num_reorder_pics), i, HEVCB_MAX_SUBLAYERS);
            aue(HF(S, sps_max_latency_increase_plus1), i, HEVCB_MAX_SUBLAYERS);
        }
        c.log2_min_cb_minus3 = ue(HF(S, log2_min_luma_coding_block_size_minus3));
        c.log2_diff_max_min_cb = ue(HF(S, log2_diff_max_min_luma_coding_block_size));
        ue(HF(S, log2_min_luma_transform_block_size_minus2));
        ue(HF(S, log2_diff_max_min_luma_transform_block_size));
        ue(HF(S, max_transform_hierarchy_depth_inter));
        ue(HF(S, max_transform_hierarchy_depth_intra));
        if (u1(HF(S, scaling_list_enabled_flag))) {
            if (u1(HF(S, sps_scaling_list_data_present_flag))) { scaling_list_data(HF(S, scaling_list_data)); }
        }
        u1(HF(S, amp_enabled_flag));
        c.sample_adaptive_offset_enabled_flag = u1(HF(S, sample_adaptive_offset_enabled_flag));
        if (u1(HF(S, pcm_enabled_flag))) {
            u(HF(S, pcm_sample_bit_depth_luma_minus1), 4);
            u(HF(S, pcm_sample_bit_depth_chroma_minus1), 4);
            ue(HF(S, log2_min_pcm_luma_coding_block_size_minus3));
            ue(HF(S, log2_diff_max_min_pcm_luma_coding_block_size));
            u1(HF(S, pcm_loop_filter_disabled_flag));
        }
        c.num_short_term_ref_pic_sets = ue(HF(S, num_short_term_ref_pic_sets));
        const uint32_t rs = (uint32_t)(sizeof(hevc_st_ref_pic_set_t) / sizeof(int));
        for (int i = 0; i < c.num_short_term_ref_pic_sets; i++) { if (runaway(i)) { break; }
            if (i < HEVCB_MAX_PICS) {
                st_ref_pic_set(HF(S, st_ref_pic_set) + (uint32_t)i * rs, i, c.num_short_term_ref_pic_sets, c.rps, c.rps[i]);
            } else if constexpr (kWrite) {
                flags |= 1u;
            } else { // beyond the reference's array: keep the bit cursor moving, report nothing
                flags |= 1u;
                hevcb_sink dummy{nullptr, nullptr, 0};
                hevcb_walker<hevcb_count_sink> w(b, dummy);
                hevcb_rps_entry scratch = c.rps[HEVCB_RPS_SLOTS - 1];
                w.st_ref_pic_set(0, i, c.num_short_term_ref_pic_sets, c.rps, scratch);
            }
        }
        c.long_term_ref_pics_present_flag = u1(HF(S, long_term_ref_pics_present_flag));
        c.num_long_term_ref_pics_sps = 0;
        c.used_by_curr_pic_lt_sps_mask = 0;
        if (c.long_term_ref_pics_present_flag) {
            c.num_long_term_ref_pics_sps = ue(HF(S, num_long_term_ref_pics_sps));
            for (int i = 0; i < c.num_long_term_ref_pics_sps; i++) { if (runaway(i)) { break; }
                au(HF(S, lt_ref_pic_poc_lsb_sps), i, HEVCB_MAX_PICS, c.log2_max_poc_lsb_minus4 + 4);
                const int f = au(HF(S, used_by_curr_pic_lt_sps_flag), i, HEVCB_MAX_PICS, 1);
                if (i < 32) { c.used_by_curr_pic_lt_sps_mask |= (uint32_t)(f & 1) << i; }
            }
        }
        c.sps_temporal_mvp_enabled_flag = u1(HF(S, sps_temporal_mvp_enabled_flag));
        u1(HF(S, strong_intra_smoothing_enabled_flag));
        if (u1(HF(S, vui_parameters_present_flag))) { vui_parameters(HF(S, vui), msl); }
        int range_ext = 0;
        if (u1(HF(S, sps_extension_present_flag))) {
            range_ext = u1(HF(S, sps_range_extension_flag));
            u1(HF(S, sps_multilayer_extension_flag));
            u1(HF(S, sps_3d_extension_flag));
            u(HF(S, sps_extension_5bits), 5);
        }
        if (range_ext) {
            const uint32_t e = HF(S, sps_range_ext);
            typedef hevc_sps_range_ext_t E;
            u1(e + HF(E, transform_skip_rotation_enabled_flag));
            u1(e + HF(E, transform_skip_context_enabled_flag));
            u1(e + HF(E, implicit_rdpcm_enabled_flag));
            u1(e + HF(E, explicit_rdpcm_enabled_flag));
            u1(e + HF(E, extended_precision_processing_flag));
            u1(e + HF(E, intra_smoothing_disabled_flag));
            u1(e + HF(E, high_precision_offsets_enabled_flag));
            u1(e + HF(E, persistent_rice_adaptation_enabled_flag));
            u1(e + HF(E, cabac_bypass_alignment_enabled_flag));
        }
        if (spec) { trailing_bits(); } // the reference's SPS has none (App. A-1)
    }

    // ---- 7.3.2.3 PPS (hevc_stream.c:419-521) ------------------------------------------------------
    HEVCB_SHD void pic_parameter_set(hevcb_pps_ctx& c)
    {
        typedef hevc_pps_t P;
        c.pic_parameter_set_id = ue(HF(P, pic_parameter_set_id));
        c.seq_parameter_set_id = ue(HF(P, seq_parameter_set_id));
        c.dependent_slice_segments_enabled_flag = u1(HF(P, dependent_slice_segments_enabled_flag));
        c.output_flag_present_flag = u1(HF(P, output_flag_present_flag));
        c.num_extra_slice_header_bits = u(HF(P, num_extra_slice_header_bits), 3);
        u1(HF(P, sign_data_hiding_enabled_flag));
        c.cabac_init_present_flag = u1(HF(P, cabac_init_present_flag));
        c.num_ref_idx_l0_default_active_minus1 = ue(HF(P, num_ref_idx_l0_default_active_minus1));
        c.num_ref_idx_l1_default_active_minus1 = ue(HF(P, num_ref_idx_l1_default_active_minus1));
        se(HF(P, init_qp_minus26));
        u1(HF(P, constrained_intra_pred_flag));
        const int transform_skip = u1(HF(P, transform_skip_enabled_flag));
        if (u1(HF(P, cu_qp_delta_enabled_flag))) { ue(HF(P, diff_cu_qp_delta_depth)); }
        se(HF(P, pps_cb_qp_offset));
        se(HF(P, pps_cr_qp_offset));
        c.pps_slice_chroma_qp_offsets_present_flag = u1(HF(P, pps_slice_chroma_qp_offsets_present_flag));
        c.weighted_pred_flag = u1(HF(P, weighted_pred_flag));
        c.weighted_bipred_flag = u1(HF(P, weighted_bipred_flag));
        u1(HF(P, transquant_bypass_enabled_flag));
        c.tiles_enabled_flag = u1(HF(P, tiles_enabled_flag));
        c.entropy_coding_sync_enabled_flag = u1(HF(P, entropy_coding_sync_enabled_flag));
        if (c.tiles_enabled_flag) {
            const int cols = ue(HF(P, num_tile_columns_minus1));
            const int rows = ue(HF(P, num_tile_rows_minus1));
            if (!u1(HF(P, uniform_spacing_flag))) {
                for (int i = 0; i < cols; i++) { if (runaway(i)) { break; } aue(HF(P, column_width_minus1), i, HEVCB_MAX_PICS); }
                for (int i = 0; i < rows; i++) { if (runaway(i)) { break; } aue(HF(P, row_height_minus1), i, HEVCB_MAX_PICS); }
            }
            u1(HF(P, loop_filter_across_tiles_enabled_flag));
        }
        c.pps_loop_filter_across_slices_enabled_flag = u1(HF(P, pps_loop_filter_across_slices_enabled_flag));
        c.deblocking_filter_override_enabled_flag = 0;
        c.pps_deblocking_filter_disabled_flag = 0;
        if (u1(HF(P, deblocking_filter_control_present_flag))) {
            c.deblocking_filter_override_enabled_flag = u1(HF(P, deblocking_filter_override_enabled_flag));
            c.pps_deblocking_filter_disabled_flag = u1(HF(P, pps_deblocking_filter_disabled_flag));
            if (spec ? !c.pps_deblocking_filter_disabled_flag : c.pps_deblocking_filter_disabled_flag) { // the reference reads the offsets when the filter is DISABLED (App. A-6)
                se(HF(P, pps_beta_offset_div2));
                se(HF(P, pps_tc_offset_div2));
            }
        }
        if (u1(HF(P, pps_scaling_list_data_present_flag))) { scaling_list_data(HF(P, scaling_list_data)); }
        c.lists_modification_present_flag = u1(HF(P, lists_modification_present_flag));
        ue(HF(P, log2_parallel_merge_level_minus2));
        c.slice_segment_header_extension_present_flag = u1(HF(P, slice_segment_header_extension_present_flag));
        int range_ext = 0;
        if (u1(HF(P, pps_extension_present_flag))) {
            range_ext = u1(HF(P, pps_range_extension_flag));
            u1(HF(P, pps_multilayer_extension_flag));
            u1(HF(P, pps_3d_extension_flag));
            u1(HF(P, pps_extension_5bits)); // ONE bit (App. A-6)
        }
        c.chroma_qp_offset_list_enabled_flag = 0;
        if (range_ext) {
            const uint32_t e = HF(P, pps_range_ext);
            typedef hevc_pps_range_ext_t E;
            if (transform_skip) { ue(e + HF(E, log2_max_transform_skip_block_size_minus2)); }
            u1(e + HF(E, cross_component_prediction_enabled_flag));
            c.chroma_qp_offset_list_enabled_flag = u1(e + HF(E, chroma_qp_offset_list_enabled_flag));
            if (c.chroma_qp_offset_list_enabled_flag) {
                ue(e + HF(E, diff_cu_chroma_qp_offset_depth));
                const int len_minus1 = ue(e + HF(E, chroma_qp_offset_list_len_minus1));
                for (int i = 0; i <= len_minus1; i++) { if (runaway(i)) { break; }
                    ase(e + HF(E, cb_qp_offset_list), i, HEVCB_MAX_PICS);
                    ase(e + HF(E, cr_qp_offset_list), i, HEVCB_MAX_PICS);
                }
            }
            ue(e + HF(E, log2_sao_offset_scale_luma));
            ue(e + HF(E, log2_sao_offset_scale_chroma));
        }
        trailing_bits();
    }

    // ---- extension mode: the NAL types the reference defines readers for but never dispatches (hevc_stream.in.c:499-573; its
    // switch returns -1 for them, which stays the default).  Fields are reported under kind HEVCB_KIND_AUX with the numbers of
    // hevcb.h (HEVCB_AUX_*); SEI payload bytes stay in the image, a message is (type, size, offset of its payload in the RBSP).
    HEVCB_SHD void access_unit_delimiter() // 7.3.2.5, hevc_stream.in.c:549-553
    {
        u(HEVCB_AUX_AUD_PIC_TYPE, 3);
        trailing_bits();
    }
    HEVCB_SHD void filler_data() // 7.3.2.8, hevc_stream.in.c:566-573: bs_next_bits(b, 8) == 0xFF (zero bits behind the end)
    {
        int32_t n = 0;
        while ((b.peek32() >> 24) == 0xFFu) { fx(8, 0xFFu, HEVCB_TI_FF_BYTE); n++; }
        syn(HEVCB_AUX_FD_FF_BYTES, n);
        trailing_bits();
    }
    HEVCB_SHD int32_t ff_coded_number() // _read_ff_coded_number, h264_stream.c:88-98
    {
        int32_t n1 = 0;
        uint32_t n2;
        do { n2 = b.read_u8(); n1 += (int32_t)n2; } while (n2 == 0xFFu);
        return n1;
    }
    HEVCB_SHD bool more_rbsp_data(int64_t last_one_bit) const // h264_stream.c:62-84
    {
        if (b.byte_pos() >= b.size) { return false; }
        return last_one_bit != b.pos; // the next bit is 0, or it is a 1 that is not the last 1 of the RBSP
    }
    HEVCB_SHD void sei_rbsp() // 7.3.2.4 / 7.3.5 (hevc_stream.in.c:499-546 under HAVE_SEI; payload: read_sei_payload, h264_sei.c:75-92)
    {
        int64_t last_one = -1; // bit position of the last 1 bit of the RBSP
        for (int64_t i = b.size - 1; i >= 0; i--) {
            const uint32_t x = b.base[i];
            if (x) {
                int tz = 0;
                while (!((x >> tz) & 1u)) { tz++; }
                last_one = i * 8 + (7 - tz);
                break;
            }
        }
        int32_t count = 0;
        do {
            const int32_t type = ff_coded_number();
            const int32_t size = ff_coded_number();
            syn(HEVCB_AUX_SEI_TYPE, type);
            syn(HEVCB_AUX_SEI_SIZE, size);
            syn(HEVCB_AUX_SEI_OFFSET, (int32_t)(b.pos >> 3));
            if (size > 0) { b.pos += 8 * (int64_t)size; } // read_sei_payload: payloadSize bytes, kept in the image
            count++;
        } while (more_rbsp_data(last_one) && !b.overrun() && count < (1 << 20));
        trailing_bits();
    }

    // getNumPicTotalCurr (hevc_stream.in.c:35-59)
    HEVCB_SHD static int num_pic_total_curr(const hevcb_sps_ctx& sps, const hevcb_rps_entry& e, int num_lt_sps, int num_lt_pics,
                                            const int* lt_idx_sps, uint32_t used_lt_mask)
    {
        int n = 0;
        for (int i = 0; i < e.num_neg && i < 32; i++) { n += (e.used_s0 >> i) & 1u; }
        for (int i = 0; i < e.num_pos && i < 32; i++) { n += (e.used_s1 >> i) & 1u; }
        const long long total = (long long)num_lt_sps + (long long)num_lt_pics;
        for (int i = 0; i < total && i < 32; i++) {
            int used;
            if (i < num_lt_sps) { const int k = lt_idx_sps[i]; used = (k >= 0 && k < 32) ? ((sps.used_by_curr_pic_lt_sps_mask >> k) & 1u) : 0; }
            else { used = (used_lt_mask >> i) & 1u; }
            n += used;
        }
        // entries 32 and up (only on corrupt input; the counts come out of the bitstream): an SPS candidate counts as index 0, a
        // slice-local one as unused -- in closed form, a count of 2^31 must not become a loop
        if (total > 32 && num_lt_sps > 32) {
            const long long m = ((long long)num_lt_sps < total ? (long long)num_lt_sps : total) - 32;
            n += (int)(m * (long long)(sps.used_by_curr_pic_lt_sps_mask & 1u));
        }
        return n;
    }

    // ---- 7.3.6 slice segment header (hevc_stream.c:782-941) --------------------------------------
    // Returns nothing; `cols` receives the per-slice columns.  `sps`/`pps` = the most recent SPS / PPS NAL that
    // precedes the slice in stream order (pointer indexing with id 0, not a table lookup).
    HEVCB_SHD void slice_segment_header(int nal_unit_type, const hevcb_sps_ctx& sps_recent, const hevcb_pps_ctx& pps_recent, hevcb_slice_cols& cols)
    {
        typedef hevc_slice_header_t H;
        syn(HF(H, collocated_from_l0_flag), 1); // init_slice_hevc (hevc_stream.c:18-23)
        const int first = u1(HF(H, first_slice_segment_in_pic_flag));
        if (nal_unit_type >= 16 && nal_unit_type <= 23) { u1(HF(H, no_output_of_prior_pics_flag)); }
        const int pps_id = ue(HF(H, pic_parameter_set_id));
        // The reference parses against the most recent PPS / SPS NAL whatever the ids say (it indexes its single structs).
        // Spec mode: the most recent PPS NAL with this id in front of the slice, then the most recent SPS NAL with that PPS's
        // seq_parameter_set_id (entry 0, the zeroed state, when there is none -- what the reference's calloc'ed tables hold).
        const hevcb_sps_ctx* sps_p = &sps_recent;
        const hevcb_pps_ctx* pps_p = &pps_recent;
        if (spec && lookup != nullptr) {
            int j = lookup->pps_count;
            while (j > 0 && lookup->pps_tab[j].pic_parameter_set_id != pps_id) { j--; }
            pps_p = &lookup->pps_tab[j];
            int q = lookup->sps_count;
            while (q > 0 && lookup->sps_tab[q].seq_parameter_set_id != pps_p->seq_parameter_set_id) { q--; }
            sps_p = &lookup->sps_tab[q];
            if (pps_id < 0 || pps_id > 255 || pps_p->seq_parameter_set_id < 0 || pps_p->seq_parameter_set_id > 31) { flags |= 1u; } // past the reference's tables
        }
        const hevcb_sps_ctx& sps = *sps_p;
        const hevcb_pps_ctx& pps = *pps_p;
        if (!spec && (pps_id != 0 || pps.seq_parameter_set_id != 0)) { flags |= 1u; } // the reference indexes past its single PPS / SPS
        int num_ref_idx_l0 = pps.num_ref_idx_l0_default_active_minus1, num_ref_idx_l1 = pps.num_ref_idx_l1_default_active_minus1;
        syn(HF(H, num_ref_idx_l0_active_minus1), num_ref_idx_l0);
        syn(HF(H, num_ref_idx_l1_active_minus1), num_ref_idx_l1);
        int dependent = 0;
        cols.first_slice_segment_in_pic_flag = first;
        cols.slice_segment_address = 0;
        cols.slice_type = 0;
        cols.slice_qp_delta = 0;
        cols.slice_pic_order_cnt_lsb = 0;
        cols.num_entry_point_offsets = 0;
        cols.short_term_ref_pic_set_idx = 0;
        if (!first) {
            if (pps.dependent_slice_segments_enabled_flag) { dependent = u1(HF(H, dependent_slice_segment_flag)); }
            // getSliceSegmentAddressBitLength (hevc_stream.in.c:115-123)
            int ctb_log2 = sps.log2_min_cb_minus3 + 3 + sps.log2_diff_max_min_cb;
            if (ctb_log2 < 0 || ctb_log2 > 30) { flags |= 1u; ctb_log2 = ctb_log2 < 0 ? 0 : 30; }
            const int64_t ctb = (int64_t)1 << ctb_log2;
            const int64_t wc = ((int64_t)sps.pic_width + ctb - 1) >> ctb_log2;
            const int64_t hc = ((int64_t)sps.pic_height + ctb - 1) >> ctb_log2;
            const int32_t pic_size = (int32_t)(wc * hc);
            cols.slice_segment_address = u(HF(H, slice_segment_address), hevcb_ceil_log2(pic_size));
        }
        cols.dependent_slice_segment_flag = dependent;
        if (!dependent) {
            for (int i = 0; i < pps.num_extra_slice_header_bits; i++) { fx(1, 1u, HEVCB_TI_SLICE_RESERVED_FLAG); } // the writer emits 1
            const int slice_type = ue(HF(H, slice_type));
            cols.slice_type = slice_type;
            if (pps.output_flag_present_flag) { u1(HF(H, pic_output_flag)); }
            if (sps.separate_colour_plane_flag == 1) { u(HF(H, colour_plane_id), 2); }
            int slice_temporal_mvp = 0, sao_luma = 0, sao_chroma = 0;
            int short_term_sps_flag = 0, st_idx = 0, num_lt_sps = 0, num_lt_pics = 0;
            int lt_idx_sps[32];
            uint32_t used_lt_mask = 0;
            hevcb_rps_entry local; // derived variables of the slice-local RPS (index num_short_term_ref_pic_sets)
            bool have_local = false;
            for (int i = 0; i < 32; i++) { lt_idx_sps[i] = 0; }
            if (nal_unit_type != 19 && nal_unit_type != 20) {
                cols.slice_pic_order_cnt_lsb = u(HF(H, slice_pic_order_cnt_lsb), sps.log2_max_poc_lsb_minus4 + 4);
                short_term_sps_flag = u1(HF(H, short_term_ref_pic_set_sps_flag));
                const int nsets = sps.num_short_term_ref_pic_sets;
                if (!short_term_sps_flag) {
                    const int slot = (nsets >= 0 && nsets < HEVCB_RPS_SLOTS) ? nsets : HEVCB_RPS_SLOTS - 1;
                    if (slot != nsets) { flags |= 1u; }
                    local = sps.rps[slot]; // entries the parse does not overwrite keep the table's previous content
                    st_ref_pic_set(HF(H, st_ref_pic_set), nsets, nsets, sps.rps, local);
                    have_local = true;
                } else if (nsets > 1) {
                    st_idx = u(HF(H, short_term_ref_pic_set_idx), hevcb_ceil_log2(nsets));
                    cols.short_term_ref_pic_set_idx = st_idx;
                }
                if (sps.long_term_ref_pics_present_flag) {
                    if (sps.num_long_term_ref_pics_sps > 0) { num_lt_sps = ue(HF(H, num_long_term_sps)); }
                    num_lt_pics = ue(HF(H, num_long_term_pics));
                    const int64_t tot = (int64_t)num_lt_sps + (int64_t)num_lt_pics;
                    for (int64_t i = 0; i < tot; i++) {
                        const int ii = (int)(i < 0x7fffffff ? i : 0x7fffffff);
                        if (i < num_lt_sps) {
                            if (sps.num_long_term_ref_pics_sps > 1) {
                                const int v = au(HF(H, lt_idx_sps), ii, HEVCB_MAX_PICS, hevcb_ceil_log2(sps.num_long_term_ref_pics_sps));
                                if (ii < 32) { lt_idx_sps[ii] = v; }
                            }
                        } else {
                            au(HF(H, poc_lsb_lt), ii, HEVCB_MAX_PICS, sps.log2_max_poc_lsb_minus4 + 4);
                            const int f = au(HF(H, used_by_curr_pic_lt_flag), ii, HEVCB_MAX_PICS, 1);
                            if (ii < 32) { used_lt_mask |= (uint32_t)(f & 1) << ii; }
                        }
                        if (au(HF(H, delta_poc_msb_present_flag), ii, HEVCB_MAX_PICS, 1)) { aue(HF(H, delta_poc_msb_cycle_lt), ii, HEVCB_MAX_PICS); }
                        if (b.overrun() && i > 64) { break; } // runaway count on a truncated NAL: the result is -1 either way
                    }
                }
                if (sps.sps_temporal_mvp_enabled_flag) { slice_temporal_mvp = u1(HF(H, slice_temporal_mvp_enabled_flag)); }
            }
            if (sps.sample_adaptive_offset_enabled_flag) {
                sao_luma = u1(HF(H, slice_sao_luma_flag));
                const int cat = (sps.separate_colour_plane_flag == 0) ? sps.chroma_format_idc : 0;
                if (cat != 0) { sao_chroma = u1(HF(H, slice_sao_chroma_flag)); }
            }
            if (slice_type == 1 || slice_type == 0) { // P or B (HEVC_SLICE_TYPE_P = 1, _B = 0)
                if (u1(HF(H, num_ref_idx_active_override_flag))) {
                    num_ref_idx_l0 = ovr(HF(H, num_ref_idx_l0_active_minus1), num_ref_idx_l0);
                    if (slice_type == 0) { num_ref_idx_l1 = ovr(HF(H, num_ref_idx_l1_active_minus1), num_ref_idx_l1); }
                }
                // the RPS that getNumPicTotalCurr consults: slice-local entry or the SPS entry short_term_ref_pic_set_idx
                const int cur_idx = short_term_sps_flag ? st_idx : sps.num_short_term_ref_pic_sets;
                const int slot = (cur_idx >= 0 && cur_idx < HEVCB_RPS_SLOTS) ? cur_idx : HEVCB_RPS_SLOTS - 1;
                const hevcb_rps_entry& e = (have_local && !short_term_sps_flag) ? local : sps.rps[slot];
                if (pps.lists_modification_present_flag) {
                    const int total = num_pic_total_curr(sps, e, num_lt_sps, num_lt_pics, lt_idx_sps, used_lt_mask);
                    if (total > 1) { // 7.3.6.2 (hevc_stream.c:944-966); flag_l1 is never read (App. A-4)
                        typedef hevc_ref_pics_lists_mod_t M;
                        const uint32_t m = HF(H, rpld);
                        if (u1(m + HF(M, ref_pic_list_modification_flag_l0))) {
                            for (int i = 0; i <= num_ref_idx_l0; i++) {
                                au(m + HF(M, list_entry_l0), i, HEVCB_MAX_PICS, hevcb_ceil_log2(total));
                                if (b.overrun() && i > 64) { break; }
                            }
                        }
                        if (spec) {
                            if (slice_type == 0 && u1(m + HF(M, ref_pic_list_modification_flag_l1))) {
                                for (int i = 0; i <= num_ref_idx_l1; i++) {
                                    au(m + HF(M, list_entry_l1), i, HEVCB_MAX_PICS, hevcb_ceil_log2(total));
                                    if (b.overrun() && i > 64) { break; }
                                }
                            }
                        } else if constexpr (!kWrite && Sink::kTrace) { // B slices: the position prefix of the element that is never read
                            if (slice_type == 0) { s.trace(b.pos, HEVCB_TRACE_SPECIAL | (uint32_t)HEVCB_TI_OPEN_LINE, 0); }
                        }
                    }
                }
                if (slice_type == 0) { u1(HF(H, mvd_l1_zero_flag)); }
                if (pps.cabac_init_present_flag) { u1(HF(H, cabac_init_flag)); }
                if (slice_temporal_mvp) {
                    int collocated_from_l0 = 1;
                    if (slice_type == 0) { collocated_from_l0 = u1(HF(H, collocated_from_l0_flag)); }
                    if ((collocated_from_l0 && num_ref_idx_l0 > 0) || (!collocated_from_l0 && num_ref_idx_l1 > 0)) { ue(HF(H, collocated_ref_idx)); }
                }
                if ((pps.weighted_pred_flag && slice_type == 1) || (pps.weighted_bipred_flag && slice_type == 0)) {
                    pred_weight_table(HF(H, pwt), sps, slice_type, num_ref_idx_l0, num_ref_idx_l1);
                }
                ue(HF(H, five_minus_max_num_merge_cand));
            }
            cols.slice_qp_delta = se(HF(H, slice_qp_delta));
            if (pps.pps_slice_chroma_qp_offsets_present_flag) {
                se(HF(H, slice_cb_qp_offset));
                se(HF(H, slice_cr_qp_offset));
            }
            if (pps.chroma_qp_offset_list_enabled_flag) { u1(HF(H, cu_chroma_qp_offset_enabled_flag)); }
            int override_flag = 0, slice_deblocking_disabled = 0;
            if (pps.deblocking_filter_override_enabled_flag) { override_flag = u1(HF(H, deblocking_filter_override_flag)); }
            if (spec) { slice_deblocking_disabled = pps.pps_deblocking_filter_disabled_flag; syn(HF(H, slice_deblocking_filter_disabled_flag), slice_deblocking_disabled); } // inherited
            if (override_flag) {
                slice_deblocking_disabled = u1(HF(H, slice_deblocking_filter_disabled_flag));
                if (!slice_deblocking_disabled) {
                    se(HF(H, slice_beta_offset_div2));
                    se(HF(H, slice_tc_offset_div2));
                }
            }
            if (pps.pps_loop_filter_across_slices_enabled_flag && (sao_luma || sao_chroma || !slice_deblocking_disabled)) {
                u1(HF(H, slice_loop_filter_across_slices_enabled_flag));
            }
        }
        if (pps.tiles_enabled_flag || pps.entropy_coding_sync_enabled_flag) {
            const int n = ue(HF(H, num_entry_point_offsets));
            cols.num_entry_point_offsets = n;
            if (n > 0) {
                const int len_minus1 = ue(HF(H, offset_len_minus1));
                for (int i = 0; i < n; i++) { if (runaway(i)) { break; }
                    // u(offset_len_minus1 + 1): widths beyond 32 only occur on corrupt input
                    const int w = len_minus1 + 1;
                    if constexpr (kWrite) {
                        raw_u(HF(H, entry_point_offset_minus1) + (uint32_t)i, i < HEVCB_MAX_PICS, w);
                    } else {
                        int32_t v;
                        const int64_t p0 = b.pos;
                        if (w <= 32) { v = (int32_t)b.read_u(w); } else { b.skip(w - 32); v = (int32_t)b.read_u(32); flags |= 1u; }
                        put_idx(HF(H, entry_point_offset_minus1), i, HEVCB_MAX_PICS, v, p0);
                    }
                    if (b.overrun() && i > 64) { break; }
                }
            }
        }
        if (pps.slice_segment_header_extension_present_flag) {
            const int len = ue(HF(H, slice_segment_header_extension_length));
            for (int i = 0; i < len; i++) { if (runaway(i)) { break; }
                fx(8, 0u, HEVCB_TI_SH_EXTENSION_DATA_BYTE);
                if (b.overrun()) { break; }
            }
        }
        trailing_bits(true); // byte_alignment(): same bit pattern handling (hevc_stream.c:641-649)
    }

    // ---- 7.3.6.3 pred_weight_table (hevc_stream.c:969-1029) ---------------------------------------
    HEVCB_SHD void pred_weight_table(uint32_t base, const hevcb_sps_ctx& sps, int slice_type, int n0, int n1)
    {
        typedef hevc_pred_weight_table_t W;
        ue(base + HF(W, luma_log2_weight_denom));
        const int cat = (sps.separate_colour_plane_flag == 0) ? sps.chroma_format_idc : 0;
        if (cat != 0) { se(base + HF(W, delta_chroma_log2_weight_denom)); }
        one_list(base, cat, n0, HF(W, luma_weight_l0_flag), HF(W, chroma_weight_l0_flag), HF(W, delta_luma_weight_l0), HF(W, luma_offset_l0),
                 HF(W, delta_chroma_weight_l0), HF(W, delta_chroma_offset_l0));
        if (slice_type == 0) {
            one_list(base, cat, n1, HF(W, luma_weight_l1_flag), HF(W, chroma_weight_l1_flag), HF(W, delta_luma_weight_l1), HF(W, luma_offset_l1),
                     HF(W, delta_chroma_weight_l1), HF(W, delta_chroma_offset_l1));
        }
    }
    HEVCB_SHD void one_list(uint32_t base, int cat, int n, uint32_t f_lw, uint32_t f_cw, uint32_t f_dl, uint32_t f_lo, uint32_t f_dcw, uint32_t f_dco)
    {
        uint64_t lw = 0, cw = 0; // chroma flags stay 0 (struct zeroed) when ChromaArrayType == 0
        for (int i = 0; i <= n; i++) { if (runaway(i)) { break; }
            const int f = au(base + f_lw, i, HEVCB_MAX_PICS, 1);
            if (i < 64) { lw |= (uint64_t)(f & 1) << i; }
            if (b.overrun() && i > 64) { break; }
        }
        if (cat != 0) {
            for (int i = 0; i <= n; i++) { if (runaway(i)) { break; }
                const int f = au(base + f_cw, i, HEVCB_MAX_PICS, 1);
                if (i < 64) { cw |= (uint64_t)(f & 1) << i; }
                if (b.overrun() && i > 64) { break; }
            }
        }
        for (int i = 0; i <= n; i++) { if (runaway(i)) { break; }
            if (i < 64 && ((lw >> i) & 1)) {
                ase(base + f_dl, i, HEVCB_MAX_PICS);
                ase(base + f_lo, i, HEVCB_MAX_PICS);
            }
            if (i < 64 && ((cw >> i) & 1)) {
                for (int j = 0; j < 2; j++) {
                    raw_se(base + f_dcw + (uint32_t)(i * 2 + j), i < HEVCB_MAX_PICS);
                    raw_se(base + f_dco + (uint32_t)(i * 2 + j), i < HEVCB_MAX_PICS);
                }
            }
            if (b.overrun() && i > 64) { break; }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// one NAL: the dispatcher of read_hevc_nal_unit (hevc_stream.c:155-241) after nal_to_rbsp succeeded
// ------------------------------------------------------------------------------------------------
struct hevcb_nal_result {
    int32_t nal_unit_type, nal_layer_id, nal_temporal_id_plus1;
    int32_t kind;       // HEVCB_KIND_*
    int32_t ok;         // 1: the reference returns nal_size; 0: it returns -1 (unsupported type or overrun)
    int32_t hdr_end;    // slices: RBSP byte offset of the cursor after byte_alignment (slice data starts one byte later)
    uint32_t flags;
    int64_t end_bits;   // bit position of the reader when the NAL's syntax ended
    hevcb_slice_cols cols;
};

HEVCB_SHD inline bool hevcb_is_slice_type(int t) { return (t >= 0 && t <= 9) || (t >= 16 && t <= 21); }

// Parses one RBSP.  `sps_in`/`pps_in`: context for slices; `sps_out`/`pps_out`: filled when the NAL is an SPS / PPS
// (may be null when the caller does not need them).
template <class Sink>
HEVCB_SHD inline void hevcb_parse_nal(const uint8_t* rbsp, int64_t rbsp_size, Sink& sink, const hevcb_sps_ctx* sps_in, const hevcb_pps_ctx* pps_in,
                                      hevcb_sps_ctx* sps_out, hevcb_pps_ctx* pps_out, hevcb_nal_result& r, bool aux = false, bool spec = false,
                                      const hevcb_ps_lookup* lookup = nullptr)
{
    hevcb_bits b;
    b.init(rbsp, rbsp_size);
    const uint32_t fzb = b.read_u(1); // forbidden_zero_bit is not validated (App. A-12)
    r.nal_unit_type = (int32_t)b.read_u(6);
    r.nal_layer_id = (int32_t)b.read_u(6);
    r.nal_temporal_id_plus1 = (int32_t)b.read_u(3);
    if constexpr (Sink::kTrace) { // the first four lines of every NAL's dump (hevc_stream.c:2363-2366)
        sink.trace(0, HEVCB_TRACE_SPECIAL | (uint32_t)HEVCB_TI_FORBIDDEN_ZERO_BIT, (int32_t)fzb);
        sink.trace(1, HEVCB_TRACE_SPECIAL | (uint32_t)HEVCB_TI_NAL_UNIT_TYPE, r.nal_unit_type);
        sink.trace(7, HEVCB_TRACE_SPECIAL | (uint32_t)HEVCB_TI_NAL_LAYER_ID, r.nal_layer_id);
        sink.trace(13, HEVCB_TRACE_SPECIAL | (uint32_t)HEVCB_TI_NAL_TEMPORAL_ID_PLUS1, r.nal_temporal_id_plus1);
    }
    r.kind = HEVCB_KIND_NONE;
    r.ok = 0;
    r.hdr_end = 0;
    r.flags = 0;
    r.end_bits = 0;
    r.cols = hevcb_slice_cols{0, 0, 0, 0, 0, 0, 0, 0};
    hevcb_walker<Sink> w(b, sink);
    w.spec = spec;
    w.lookup = lookup;
    const int t = r.nal_unit_type;
    if (hevcb_is_slice_type(t)) {
        r.kind = HEVCB_KIND_SLICE;
        w.slice_segment_header(t, *sps_in, *pps_in, r.cols);
        r.hdr_end = (int32_t)b.byte_pos();
        w.trailing_bits(); // read_hevc_rbsp_slice_trailing_bits skips 8 bits of slice data (App. A-11)
    } else if (t == 32) {
        r.kind = HEVCB_KIND_VPS;
        w.video_parameter_set();
    } else if (t == 33) {
        r.kind = HEVCB_KIND_SPS;
        w.seq_parameter_set(*sps_out);
    } else if (t == 34) {
        r.kind = HEVCB_KIND_PPS;
        w.pic_parameter_set(*pps_out);
    } else if (aux && t >= 35 && t <= 40) { // extension mode only
        r.kind = HEVCB_KIND_AUX;
        if (t == 35) { w.access_unit_delimiter(); }
        else if (t == 38) { w.filler_data(); }
        else if (t >= 39) { w.sei_rbsp(); }
        // 36 / 37: end of sequence / end of bitstream have no payload (hevc_stream.in.c:556-563)
    } else {
        r.flags = w.flags;
        return; // default: return -1 with h->nal already filled (hevc_stream.c:220-221)
    }
    r.flags = w.flags;
    r.end_bits = b.pos;
    r.ok = b.overrun() ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------
// one NAL, write side: write_hevc_nal_unit (hevc_stream.c:1249-1335) up to (not including) rbsp_to_nal
// ------------------------------------------------------------------------------------------------
struct hevcb_write_result {
    int32_t ok;        // 1: the reference's writer returns > 0 for this NAL (supported type, no overrun)
    int32_t hdr_bytes; // slices: bytes up to and including byte_alignment(); the writer then appends rbsp_trailing_bits (0x80)
    int64_t bytes;     // RBSP bytes produced (bs_pos: a trailing partial byte is dropped, as happens for an SPS, App. A-1)
    uint32_t flags;
};

// Writes the RBSP of one NAL from the values its parse produced (`rp`) into `bw` (count only when bw.dst is null).
// nal_hdr = nal_unit_type | nal_layer_id << 8 | nal_temporal_id_plus1 << 16 (what the parse left in h->nal).
// `sps_scratch`: working copy for the derived RPS tables an SPS builds while it is walked.
HEVCB_SHD inline void hevcb_write_nal(hevcb_replay& rp, hevcb_bitwriter& bw, int32_t nal_hdr, const hevcb_sps_ctx* sps_in, const hevcb_pps_ctx* pps_in,
                                      hevcb_sps_ctx* sps_scratch, hevcb_write_result& r, bool spec = false, const hevcb_ps_lookup* lookup = nullptr)
{
    hevcb_bits nob;
    nob.init(nullptr, 0);
    hevcb_sink nos{nullptr, nullptr, 0};
    const int t = nal_hdr & 0xFF;
    r.ok = 0;
    r.hdr_bytes = 0;
    r.bytes = 0;
    r.flags = 0;
    bw.write_u(1, 0u); // forbidden_zero_bit
    bw.write_u(6, (uint32_t)t);
    bw.write_u(6, (uint32_t)((nal_hdr >> 8) & 0xFF));
    bw.write_u(3, (uint32_t)((nal_hdr >> 16) & 0xFF));
    hevcb_walker<hevcb_sink, true> w(nob, nos, &bw, &rp);
    w.spec = spec;
    w.lookup = lookup;
    if (hevcb_is_slice_type(t)) {
        hevcb_slice_cols cols;
        w.slice_segment_header(t, *sps_in, *pps_in, cols);
        r.hdr_bytes = (int32_t)bw.bytes();
        w.trailing_bits(); // write_hevc_rbsp_slice_trailing_bits: no slice data is written
    } else if (t == 32) {
        w.video_parameter_set();
    } else if (t == 33) {
        w.seq_parameter_set(*sps_scratch);
    } else if (t == 34) {
        hevcb_pps_ctx pc;
        w.pic_parameter_set(pc);
    } else {
        return;
    }
    r.flags = w.flags;
    r.bytes = bw.bytes();
    r.ok = (bw.overrun() || (w.flags & 1u)) ? 0 : 1;
}
