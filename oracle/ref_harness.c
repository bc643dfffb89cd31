/*
 * oracle/ref_harness.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin driver around the UNMODIFIED reference sources (leslie-wang/hevcbitstream).
 * It is compiled together with the reference's own .c files where they lie under
 * /root/reference (see oracle/Makefile); nothing from the reference is copied into
 * this repository.  The resulting oracle/_ref/libhevcref.so is git-ignored and only
 * ever loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * hevc_stream.c is pulled into THIS translation unit with #include so that the
 * harness can reset the reference's file-static RPS tables (hevc_stream.in.c:26-32)
 * between independent streams; every other reference file is compiled separately.
 *
 * Exposed (all prefixed ref_):
 *   ref_scan_all      the canonical `while (find_nal_unit(p, sz, &s, &e) > 0)` loop
 *                     (hevc_analyze.c:135-176) returning absolute offsets
 *   ref_strip_all     nal_to_rbsp per NAL (h264_nal.c:147)
 *   ref_insert_all    rbsp_to_nal per NAL (h264_nal.c:92)
 *   ref_parse_all     read_hevc_nal_unit per NAL (hevc_stream.c:155) + state digests/dumps
 *   ref_write_*       write_hevc_nal_unit from caller-filled structs (hevc_stream.c:1249)
 *   ref_rewrite_all   SURVEY 3.4 parse -> edit -> write -> splice -> rbsp_to_nal composition
 *   ref_gen_stream    synthetic Annex-B generator built on the reference's own writer
 *   ref_time_*        timing loops used as the CPU baseline
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "hevc_stream.c" /* the committed, generated reference parser (read variant :155, write :1249) */

int peek_hevc_nal_unit(hevc_stream_t* h, uint8_t* buf, int size); /* hevc_nal.c:97 (undeclared in headers) */

#define REF_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* helpers                                                                                      */
/* ------------------------------------------------------------------------------------------ */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Reset the reference's process-global derived RPS tables so that independent streams parsed in
 * one process do not see each other's state. */
#ifdef REF_SPEC /* spec-correct build (oracle/make_spec_ref.py): the derived RPS tables exist once per SPS id; ids written by the generator */
static int g_sps_ids[32], g_n_sps_ids, g_pps_ids[64], g_n_pps_ids;
static void note_id(int* ids, int* n, int id)
{
    for (int i = 0; i < *n; i++) { if (ids[i] == id) { return; } }
    ids[(*n)++] = id;
}
#endif
REF_API void ref_reset_static_state(void)
{
#ifdef REF_SPEC
    memset(NumDeltaPocs_, 0, sizeof(NumDeltaPocs_));
    memset(NumNegativePics_, 0, sizeof(NumNegativePics_));
    memset(NumPositivePics_, 0, sizeof(NumPositivePics_));
    memset(DeltaPocS0_, 0, sizeof(DeltaPocS0_));
    memset(UsedByCurrPicS0_, 0, sizeof(UsedByCurrPicS0_));
    memset(DeltaPocS1_, 0, sizeof(DeltaPocS1_));
    memset(UsedByCurrPicS1_, 0, sizeof(UsedByCurrPicS1_));
    spec_rps_cur = 0;
    return;
#endif
    memset(NumDeltaPocs, 0, sizeof(NumDeltaPocs));
    memset(NumNegativePics, 0, sizeof(NumNegativePics));
    memset(NumPositivePics, 0, sizeof(NumPositivePics));
    memset(DeltaPocS0, 0, sizeof(DeltaPocS0));
    memset(UsedByCurrPicS0, 0, sizeof(UsedByCurrPicS0));
    memset(DeltaPocS1, 0, sizeof(DeltaPocS1));
    memset(UsedByCurrPicS1, 0, sizeof(UsedByCurrPicS1));
}

/* order-sensitive 64-bit digest of an int array: h = h*M + (uint32)x + 1 (mod 2^64). The same
 * function is exported by oracle/liboracle.so (oracle_hash_ints) for the product side of tests. */
static uint64_t hash_ints(uint64_t h, const int* p, size_t n)
{
    const uint64_t M = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < n; i++) { h = h * M + (uint64_t)(uint32_t)p[i] + 1ull; }
    return h;
}
static uint64_t hash_bytes(uint64_t h, const uint8_t* p, size_t n)
{
    const uint64_t M = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < n; i++) { h = h * M + (uint64_t)p[i] + 1ull; }
    return h;
}

REF_API uint64_t ref_hash_ints(uint64_t seed, const int* p, int64_t n) { return hash_ints(seed, p, (size_t)n); }
REF_API uint64_t ref_hash_bytes(uint64_t seed, const uint8_t* p, int64_t n) { return hash_bytes(seed, p, (size_t)n); }

REF_API int ref_sizeof(int what)
{
    switch (what) {
        case 0: return (int)sizeof(hevc_vps_t);
        case 1: return (int)sizeof(hevc_sps_t);
        case 2: return (int)sizeof(hevc_pps_t);
        case 3: return (int)sizeof(hevc_slice_header_t);
        case 4: return (int)sizeof(hevc_nal_t);
        case 5: return (int)sizeof(hevc_stream_t);
        case 6: return (int)sizeof(bs_t);
        case 7: return (int)sizeof(hevc_hrd_t);
        case 8: return (int)sizeof(hevc_vui_t);
        case 9: return (int)sizeof(hevc_profile_tier_level_t);
        case 10: return (int)sizeof(hevc_st_ref_pic_set_t);
        case 11: return (int)sizeof(hevc_scaling_list_data_t);
        case 12: return (int)sizeof(hevc_pred_weight_table_t);
        default: return -1;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* byte layer                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* One find_nal_unit call on an arbitrarily large buffer.  The reference API takes `int size`; a
 * call that returns through its normal path (> 0, or 0 with *nal_start > 0 = zero-length NAL) never
 * evaluated a size-dependent exit, so a truncated window gives the same answer as the full buffer.
 * Only the size-dependent exits (0 "no start", -1 "no end") need the true remaining size. */
#define REF_WINDOW (1 << 30)
static int find_nal_big(uint8_t* p, int64_t rem, int64_t* s, int64_t* e)
{
    int64_t skipped = 0;
    for (;;) {
        int w = (rem - skipped > REF_WINDOW) ? REF_WINDOW : (int)(rem - skipped);
        int ns = 0, ne = 0;
        int r = find_nal_unit(p + skipped, w, &ns, &ne);
        int truncated = (rem - skipped > (int64_t)w);
        if (r > 0 || (r == 0 && ns > 0) || !truncated) {
            if (r == 0 && ns == 0 && ne == 0) { *s = 0; *e = 0; return 0; }
            *s = skipped + ns; *e = skipped + ne; return r;
        }
        if (r == 0) { skipped += w - 8; continue; } /* no start code in this window: slide (keeps i==0 test harmless) */
        /* r == -1 in a truncated window: a NAL longer than the window; not supported by this harness */
        fprintf(stderr, "ref_harness: NAL longer than %d bytes\n", REF_WINDOW);
        abort();
    }
}

/* buf must be followed by >= 8 readable bytes (the reference reads up to buf[size+2]); callers pad
 * with zeros, which is the padding semantics the product documents. */
REF_API int64_t ref_scan_all(uint8_t* buf, int64_t size, int64_t* starts, int64_t* ends, int64_t cap,
                             int32_t* last_rc, int64_t* last_start, int64_t* last_end)
{
    uint8_t* p = buf;
    int64_t sz = size;
    int64_t n = 0, s = 0, e = 0;
    int r;
    while ((r = find_nal_big(p, sz, &s, &e)) > 0) {
        if (n < cap) { starts[n] = (p - buf) + s; ends[n] = (p - buf) + e; }
        n++;
        p += e;
        sz -= e;
    }
    *last_rc = r;
    *last_start = (p - buf) + s;
    *last_end = (p - buf) + e;
    return n;
}

REF_API int ref_find_nal_unit(uint8_t* buf, int size, int* nal_start, int* nal_end)
{
    return find_nal_unit(buf, size, nal_start, nal_end);
}
REF_API int ref_nal_to_rbsp(const uint8_t* nal, int* nal_size, uint8_t* rbsp, int* rbsp_size)
{
    return nal_to_rbsp(nal, nal_size, rbsp, rbsp_size);
}
REF_API int ref_rbsp_to_nal(const uint8_t* rbsp, const int* rbsp_size, uint8_t* nal, int* nal_size)
{
    return rbsp_to_nal(rbsp, rbsp_size, nal, nal_size);
}

/* nal_to_rbsp for every NAL; RBSPs are appended densely to rbsp_out. rc[k] = return value (-1 on
 * error, nothing appended), nal_size[k] = bytes consumed as reported by the reference. */
REF_API int64_t ref_strip_all(const uint8_t* buf, const int64_t* starts, const int64_t* ends, int64_t n,
                              uint8_t* rbsp_out, int64_t rbsp_cap, int64_t* rbsp_off, int32_t* rc, int32_t* nal_size)
{
    int64_t o = 0;
    for (int64_t k = 0; k < n; k++) {
        int ns = (int)(ends[k] - starts[k]);
        int rs = ns;
        rbsp_off[k] = o;
        if (o + ns > rbsp_cap) { return -1; }
        int r = nal_to_rbsp(buf + starts[k], &ns, rbsp_out + o, &rs);
        rc[k] = r;
        nal_size[k] = ns;
        if (r >= 0) { o += r; }
    }
    rbsp_off[n] = o;
    return o;
}

/* rbsp_to_nal for every RBSP (rbsp_off has n+1 entries); NALs appended densely, optionally each
 * preceded by a start code of sc_len bytes (0, 3 or 4). nal_off has n+1 entries. */
REF_API int64_t ref_insert_all(const uint8_t* rbsp, const int64_t* rbsp_off, const int64_t* rbsp_end, int64_t n,
                               int sc_len, uint8_t* out, int64_t out_cap, int64_t* nal_off)
{
    int64_t o = 0;
    for (int64_t k = 0; k < n; k++) {
        int rs = (int)(rbsp_end[k] - rbsp_off[k]);
        if (o + sc_len + (int64_t)rs * 3 / 2 + 8 > out_cap) { return -1; }
        for (int i = 0; i < sc_len; i++) { out[o++] = (i == sc_len - 1) ? 1 : 0; }
        nal_off[k] = o;
        int ns = 0;
        int r = rbsp_to_nal(rbsp + rbsp_off[k], &rs, out + o, &ns);
        if (r < 0) { return -1; }
        o += r;
    }
    nal_off[n] = o;
    return o;
}

/* ------------------------------------------------------------------------------------------ */
/* syntax layer: parse                                                                          */
/* ------------------------------------------------------------------------------------------ */

static int is_slice_nut(int t) { return (t >= 0 && t <= 9) || (t >= 16 && t <= 21); }

typedef struct {
    int32_t rc;            /* read_hevc_nal_unit return value */
    int32_t strip_rc;      /* nal_to_rbsp return value on the same bytes */
    int32_t nal_unit_type; /* h->nal after the call */
    int32_t nal_layer_id;
    int32_t nal_temporal_id_plus1;
    int32_t slice_data_size; /* h->slice_data->rbsp_size after a slice NAL, else 0 */
    uint64_t state_hash;   /* digest of the struct the NAL wrote (sh / vps / sps / pps), 0 otherwise */
    uint64_t slice_data_hash;
} ref_nal_record; /* 40 bytes */

/* Parse every NAL in order with ONE hevc_stream_t, exactly like hevc_analyze.c:148-149 but with the
 * non-debug reader.  Optional dumps (may be NULL): sh_dump[n][sizeof(sh)/4] gets h->sh after every
 * slice NAL that passed the strip; vps/sps/pps dumps get the struct after each such PS NAL, in order
 * of appearance, up to the given capacities (counts returned through *n_vps etc.). */
REF_API int64_t ref_parse_all(uint8_t* buf, const int64_t* starts, const int64_t* ends, int64_t n,
                              ref_nal_record* rec, int32_t* sh_dump,
                              int32_t* vps_dump, int32_t vps_cap, int32_t* n_vps,
                              int32_t* sps_dump, int32_t sps_cap, int32_t* n_sps,
                              int32_t* pps_dump, int32_t pps_cap, int32_t* n_pps)
{
    hevc_stream_t* h = hevc_new();
    ref_reset_static_state();
    int cv = 0, cs = 0, cp = 0;
    int64_t ok = 0;
    for (int64_t k = 0; k < n; k++) {
        int size = (int)(ends[k] - starts[k]);
        uint8_t* nal = buf + starts[k];
        ref_nal_record* r = &rec[k];
        memset(r, 0, sizeof(*r));
        {
            int ns = size, rs = size;
            uint8_t* tmp = (uint8_t*)malloc(size > 0 ? size : 1);
            r->strip_rc = nal_to_rbsp(nal, &ns, tmp, &rs);
            free(tmp);
        }
        r->rc = read_hevc_nal_unit(h, nal, size);
        r->nal_unit_type = h->nal->nal_unit_type;
        r->nal_layer_id = h->nal->nal_layer_id;
        r->nal_temporal_id_plus1 = h->nal->nal_temporal_id_plus1;
        if (r->rc >= 0) { ok++; }
        if (r->strip_rc < 0) { continue; }
        int t = h->nal->nal_unit_type;
        if (is_slice_nut(t)) {
            r->state_hash = hash_ints(0, (const int*)h->sh, sizeof(hevc_slice_header_t) / 4);
            r->slice_data_size = h->slice_data->rbsp_size;
            if (h->slice_data->rbsp_buf && h->slice_data->rbsp_size > 0) {
                r->slice_data_hash = hash_bytes(0, h->slice_data->rbsp_buf, (size_t)h->slice_data->rbsp_size);
            }
            if (sh_dump) { memcpy(sh_dump + k * (int64_t)(sizeof(hevc_slice_header_t) / 4), h->sh, sizeof(hevc_slice_header_t)); }
        } else if (t == HEVC_NAL_UNIT_TYPE_VPS_NUT) {
            r->state_hash = hash_ints(0, (const int*)h->vps, sizeof(hevc_vps_t) / 4);
            if (vps_dump && cv < vps_cap) { memcpy(vps_dump + (int64_t)cv * (sizeof(hevc_vps_t) / 4), h->vps, sizeof(hevc_vps_t)); }
            cv++;
        } else if (t == HEVC_NAL_UNIT_TYPE_SPS_NUT) {
            r->state_hash = hash_ints(0, (const int*)h->sps, sizeof(hevc_sps_t) / 4);
            if (sps_dump && cs < sps_cap) { memcpy(sps_dump + (int64_t)cs * (sizeof(hevc_sps_t) / 4), h->sps, sizeof(hevc_sps_t)); }
            cs++;
        } else if (t == HEVC_NAL_UNIT_TYPE_PPS_NUT) {
            r->state_hash = hash_ints(0, (const int*)h->pps, sizeof(hevc_pps_t) / 4);
            if (pps_dump && cp < pps_cap) { memcpy(pps_dump + (int64_t)cp * (sizeof(hevc_pps_t) / 4), h->pps, sizeof(hevc_pps_t)); }
            cp++;
        }
    }
    if (n_vps) { *n_vps = cv; }
    if (n_sps) { *n_sps = cs; }
    if (n_pps) { *n_pps = cp; }
    if (h->slice_data && h->slice_data->rbsp_buf) { free(h->slice_data->rbsp_buf); h->slice_data->rbsp_buf = NULL; }
    hevc_free(h);
    return ok;
}

/* read_debug_hevc_nal_unit dump (hevc_stream.c:2343) of every NAL to `path` with the hevc_analyze
 * framing lines (hevc_analyze.c:139-149).  stdout is redirected to the file for the duration. */
REF_API int ref_analyze_to_file(uint8_t* buf, int64_t size, const char* path, int verbose)
{
    FILE* f = fopen(path, "wt");
    if (!f) { return -1; }
    fflush(stdout);
    FILE* saved_dbg = h264_dbgfile;
    h264_dbgfile = f;
    /* field lines always go to stdout in the reference; route stdout into the same file */
    int saved_fd = dup(fileno(stdout));
    dup2(fileno(f), fileno(stdout));

    hevc_stream_t* h = hevc_new();
    ref_reset_static_state();
    uint8_t* p = buf;
    int64_t sz = size, s = 0, e = 0;
    int r;
    while ((r = find_nal_big(p, sz, &s, &e)) > 0 || r == -1) {
        if (verbose > 0) {
            fprintf(h264_dbgfile, "!! Found NAL at offset %lld (0x%04llX), size %lld (0x%04llX) \n",
                    (long long)((p - buf) + s), (long long)((p - buf) + s), (long long)(e - s), (long long)(e - s));
            fflush(h264_dbgfile);
        }
        read_debug_hevc_nal_unit(h, p + s, (int)(e - s));
        fflush(stdout);
        if (r == -1) { break; }
        p += e;
        sz -= e;
    }
    fflush(stdout);
    dup2(saved_fd, fileno(stdout));
    close(saved_fd);
    h264_dbgfile = saved_dbg;
    fclose(f);
    if (h->slice_data && h->slice_data->rbsp_buf) { free(h->slice_data->rbsp_buf); h->slice_data->rbsp_buf = NULL; }
    hevc_free(h);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* syntax layer: write                                                                          */
/* ------------------------------------------------------------------------------------------ */

/* A persistent stream object for struct-driven writer tests: the caller copies struct images in,
 * then asks the reference to write a NAL of the given type. */
static hevc_stream_t* g_wh = NULL;
REF_API void ref_writer_reset(void)
{
    if (g_wh) {
        if (g_wh->slice_data && g_wh->slice_data->rbsp_buf) { free(g_wh->slice_data->rbsp_buf); g_wh->slice_data->rbsp_buf = NULL; }
        hevc_free(g_wh);
    }
    g_wh = hevc_new();
    ref_reset_static_state();
}
REF_API void* ref_writer_struct(int what)
{
    if (!g_wh) { ref_writer_reset(); }
    switch (what) {
        case 0: return g_wh->vps;
        case 1: return g_wh->sps;
        case 2: return g_wh->pps;
        case 3: return g_wh->sh;
        case 4: return g_wh->nal;
        default: return NULL;
    }
}
/* write_hevc_nal_unit (hevc_stream.c:1249) with h->nal preset by the caller */
REF_API int ref_writer_write(uint8_t* out, int cap)
{
    if (!g_wh) { ref_writer_reset(); }
    return write_hevc_nal_unit(g_wh, out, cap);
}
/* read back with the same stream object (SURVEY App. A-1 resync) */
REF_API int ref_writer_read(uint8_t* nal, int size)
{
    if (!g_wh) { ref_writer_reset(); }
    return read_hevc_nal_unit(g_wh, nal, size);
}

/* ------------------------------------------------------------------------------------------ */
/* parse -> edit -> write -> splice -> rbsp_to_nal                                  */
/* ------------------------------------------------------------------------------------------ */

/* For every NAL of the input stream produce the rewritten NAL:
 *   slices : nal_to_rbsp; read_hevc_nal_unit; header end = RBSP offset after byte_alignment;
 *            sh->slice_qp_delta += qp_delta_add; write_hevc_nal_unit into scratch; scratch RBSP
 *            minus its final 0x80 byte + original RBSP from the old header end; rbsp_to_nal.
 *   SPS    : read; if vui present, vui.video_full_range_flag ^= vui_flip; write_hevc_nal_unit;
 *            then re-read the written SPS with the same h (App. A-1).
 *   others (VPS/PPS)   : read, write_hevc_nal_unit.
 *   unsupported / failing NALs are copied through unchanged.
 * Output is a fresh Annex-B stream: each NAL is preceded by the same start-code bytes it had in the
 * input (everything between the previous NAL's end and this NAL's start is copied verbatim).
 */
static int slice_header_rbsp_len(uint8_t* rbsp, int rbsp_size, hevc_stream_t* h)
{
    /* re-run the reference reader on the RBSP to find where byte_alignment leaves the cursor */
    bs_t b;
    bs_init(&b, rbsp, rbsp_size);
    bs_skip_u(&b, 16); /* forbidden_zero_bit + 3 header fields (hevc_stream.c:176-179) */
    read_hevc_slice_header(h, &b);
    return (int)(b.p - b.start);
}

REF_API int64_t ref_rewrite_all(uint8_t* buf, int64_t size, const int64_t* starts, const int64_t* ends, int64_t n,
                                int qp_delta_add, int vui_flip, uint8_t* out, int64_t out_cap,
                                int64_t* out_starts, int64_t* out_ends)
{
    hevc_stream_t* h = hevc_new();
    ref_reset_static_state();
    int64_t o = 0, prev_end = 0;
    for (int64_t k = 0; k < n; k++) {
        int nsz = (int)(ends[k] - starts[k]);
        uint8_t* nal = buf + starts[k];
        int64_t gap = starts[k] - prev_end;
        if (o + gap + (int64_t)nsz * 2 + 64 > out_cap) { hevc_free(h); return -1; }
        memcpy(out + o, buf + prev_end, (size_t)gap);
        o += gap;
        prev_end = ends[k];
        out_starts[k] = o;

        uint8_t* rbsp = (uint8_t*)calloc(1, nsz + 16);
        int ns = nsz, rs = nsz;
        int src = nal_to_rbsp(nal, &ns, rbsp, &rs);
        int rc = (src < 0) ? -1 : read_hevc_nal_unit(h, nal, nsz);
        int t = h->nal->nal_unit_type;
        int done = 0;
        if (rc >= 0 && is_slice_nut(t)) {
            int hdr_len = slice_header_rbsp_len(rbsp, rs, h);
            if (hdr_len <= rs) {
                h->sh->slice_qp_delta += qp_delta_add;
                int cap = 16384; /* header-only scratch: the reference writer emits header + 0x80 only */
                uint8_t* tmp = (uint8_t*)calloc(1, cap);
                int wn = write_hevc_nal_unit(h, tmp, cap);
                if (wn > 0) {
                    uint8_t* wr = (uint8_t*)calloc(1, cap);
                    int wns = wn, wrs = cap;
                    int wr_rc = nal_to_rbsp(tmp, &wns, wr, &wrs);
                    if (wr_rc > 0) {
                        int newhdr = wr_rc - 1; /* drop the writer's final 0x80 */
                        int total = newhdr + (rs - hdr_len);
                        uint8_t* nr = (uint8_t*)malloc(total + 16);
                        memcpy(nr, wr, newhdr);
                        memcpy(nr + newhdr, rbsp + hdr_len, rs - hdr_len);
                        int outn = 0;
                        int r2 = rbsp_to_nal(nr, &total, out + o, &outn);
                        if (r2 >= 0) { o += r2; done = 1; }
                        free(nr);
                    }
                    free(wr);
                }
                free(tmp);
            }
        } else if (rc >= 0 && (t == HEVC_NAL_UNIT_TYPE_VPS_NUT || t == HEVC_NAL_UNIT_TYPE_SPS_NUT || t == HEVC_NAL_UNIT_TYPE_PPS_NUT)) {
            if (t == HEVC_NAL_UNIT_TYPE_SPS_NUT && vui_flip && h->sps->vui_parameters_present_flag && h->sps->vui.video_signal_type_present_flag) {
                h->sps->vui.video_full_range_flag ^= 1;
            }
            int cap = nsz * 2 + 64;
            int wn = write_hevc_nal_unit(h, out + o, cap);
            if (wn > 0) {
                if (t == HEVC_NAL_UNIT_TYPE_SPS_NUT) { read_hevc_nal_unit(h, out + o, wn); }
                o += wn;
                done = 1;
            }
        }
        if (!done) { memcpy(out + o, nal, nsz); o += nsz; }
        out_ends[k] = o;
        free(rbsp);
    }
    /* trailing bytes after the last NAL */
    if (size > prev_end) {
        if (o + (size - prev_end) > out_cap) { hevc_free(h); return -1; }
        memcpy(out + o, buf + prev_end, (size_t)(size - prev_end));
        o += size - prev_end;
    }
    if (h->slice_data && h->slice_data->rbsp_buf) { free(h->slice_data->rbsp_buf); h->slice_data->rbsp_buf = NULL; }
    hevc_free(h);
    return o;
}

/* ------------------------------------------------------------------------------------------ */
/* generator: synthetic Annex-B streams written by the reference's own writer                   */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    uint64_t seed;
    int32_t profile;       /* 0 = BASELINE config 1 (Main 1920x1080, IDR every idr_period, P otherwise)
                              1 = rich: multi-slice, tiles/WPP entry points, long-term refs, slice-local and
                                  inter-predicted RPS, pred-weight tables, list modification, VUI+HRD, re-sent PS */
    int32_t idr_period;
    int64_t n_slices;
    int32_t payload_min;   /* random payload bytes appended after each slice header */
    int32_t payload_max;
    int32_t zero_heavy_pct;/* percentage of slices whose payload is drawn from {0,1,2,3,rand} (EPB-dense) */
    int32_t extra_zero_pct;/* percentage of NALs followed by 1-3 trailing_zero_8bits */
    int32_t ps_period;     /* rich: re-send (fresh random) SPS/PPS every ps_period slices; 0 = never */
    int32_t unsupported_pct;/* rich: percentage of extra NALs of types the reference rejects (AUD/SEI/...) */
} ref_gen_params;

static uint64_t g_rng = 88172645463325252ull;
static inline uint64_t rnd64(void)
{
    g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17;
    return g_rng;
}
static inline int rr(int lo, int hi) { return lo + (int)(rnd64() % (uint64_t)(hi - lo + 1)); }
static inline int pct(int p) { return (int)(rnd64() % 100) < p; }

static void gen_ptl(hevc_profile_tier_level_t* ptl, int rich, int max_sub_layers_minus1)
{
    memset(ptl, 0, sizeof(*ptl));
    ptl->general_profile_idc = rich ? rr(1, 5) : 1;
    ptl->general_tier_flag = rich ? rr(0, 1) : 0;
    for (int i = 0; i < 32; i++) { ptl->general_profile_compatibility_flag[i] = rich ? pct(15) : 0; }
    ptl->general_profile_compatibility_flag[1] = 1;
    ptl->general_profile_compatibility_flag[2] = 1;
    ptl->general_progressive_source_flag = 1;
    ptl->general_frame_only_constraint_flag = 1;
    if (rich) {
        ptl->general_max_12bit_constraint_flag = rr(0, 1);
        ptl->general_max_10bit_constraint_flag = rr(0, 1);
        ptl->general_max_8bit_constraint_flag = rr(0, 1);
        ptl->general_max_422chroma_constraint_flag = rr(0, 1);
        ptl->general_max_420chroma_constraint_flag = rr(0, 1);
        ptl->general_intra_constraint_flag = rr(0, 1);
        ptl->general_lower_bit_rate_constraint_flag = rr(0, 1);
        ptl->general_inbld_flag = rr(0, 1);
    }
    ptl->general_level_idc = rich ? rr(30, 186) : 120;
    for (int i = 0; i < max_sub_layers_minus1; i++) {
        ptl->sub_layer_profile_present_flag[i] = rr(0, 1);
        ptl->sub_layer_level_present_flag[i] = 0; /* u8-vs-u1 writer/reader mismatch in the committed file */
        ptl->sub_layer_profile_space[i] = rr(0, 3);
        ptl->sub_layer_tier_flag[i] = rr(0, 1);
        ptl->sub_layer_profile_idc[i] = rr(0, 8);
        for (int j = 0; j < 32; j++) { ptl->sub_layer_profile_compatibility_flag[i][j] = pct(20); }
        ptl->sub_layer_progressive_source_flag[i] = rr(0, 1);
        ptl->sub_layer_interlaced_source_flag[i] = rr(0, 1);
        ptl->sub_layer_non_packed_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_frame_only_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_max_12bit_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_max_10bit_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_max_8bit_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_max_422chroma_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_max_420chroma_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_max_monochrome_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_intra_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_one_picture_only_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_lower_bit_rate_constraint_flag[i] = rr(0, 1);
        ptl->sub_layer_inbld_flag[i] = rr(0, 1);
    }
}

static void gen_sub_layer_hrd(hevc_sub_layer_hrd_t* s)
{
    for (int i = 0; i < MAX_CPB_CNT; i++) {
        s->bit_rate_value_minus1[i] = rr(0, 60000);
        s->cpb_size_value_minus1[i] = rr(0, 60000);
        s->cpb_size_du_value_minus1[i] = rr(0, 1000);
        s->bit_rate_du_value_minus1[i] = rr(0, 1000);
        s->cbr_flag[i] = rr(0, 1);
    }
}

static void gen_hrd(hevc_hrd_t* hrd, int max_sub_layers_minus1)
{
    memset(hrd, 0, sizeof(*hrd));
    hrd->nal_hrd_parameters_present_flag = rr(0, 1);
    hrd->vcl_hrd_parameters_present_flag = rr(0, 1);
    hrd->sub_pic_hrd_params_present_flag = rr(0, 1);
    hrd->tick_divisor_minus2 = rr(0, 255);
    hrd->du_cpb_removal_delay_increment_length_minus1 = rr(0, 31);
    hrd->sub_pic_cpb_params_in_pic_timing_sei_flag = rr(0, 1);
    hrd->dpb_output_delay_du_length_minus1 = rr(0, 31);
    hrd->bit_rate_scale = rr(0, 15);
    hrd->cpb_size_scale = rr(0, 15);
    hrd->cpb_size_du_scale = rr(0, 15);
    hrd->initial_cpb_removal_delay_length_minus1 = rr(0, 31);
    hrd->au_cpb_removal_delay_length_minus1 = rr(0, 31);
    hrd->dpb_output_delay_length_minus1 = rr(0, 31);
    for (int i = 0; i <= max_sub_layers_minus1; i++) {
        hrd->fixed_pic_rate_general_flag[i] = rr(0, 1);
        /* keep struct values consistent with what a reader reconstructs (fields not coded stay 0) */
        hrd->fixed_pic_rate_within_cvs_flag[i] = hrd->fixed_pic_rate_general_flag[i] ? 0 : rr(0, 1);
        if (hrd->fixed_pic_rate_within_cvs_flag[i]) {
            hrd->elemental_duration_in_tc_minus1[i] = rr(0, 2047);
            hrd->low_delay_hrd_flag[i] = 0;
        } else {
            hrd->low_delay_hrd_flag[i] = rr(0, 1);
        }
        hrd->cpb_cnt_minus1[i] = hrd->low_delay_hrd_flag[i] ? rr(0, 4) : 0;
#ifdef REF_SPEC /* the spec's conditions: within_cvs inferred 1 under general, cpb_cnt_minus1 present when low_delay is 0 */
        if (hrd->fixed_pic_rate_general_flag[i]) {
            hrd->fixed_pic_rate_within_cvs_flag[i] = 1;
            hrd->elemental_duration_in_tc_minus1[i] = rr(0, 2047);
            hrd->low_delay_hrd_flag[i] = 0;
        }
        hrd->cpb_cnt_minus1[i] = hrd->low_delay_hrd_flag[i] ? 0 : rr(0, 4);
#endif
        gen_sub_layer_hrd(&hrd->sub_layer_hrd_nal[i]);
        gen_sub_layer_hrd(&hrd->sub_layer_hrd_vcl[i]);
    }
}

static void gen_scaling_list(hevc_scaling_list_data_t* sld)
{
    memset(sld, 0, sizeof(*sld));
    for (int s = 0; s < 4; s++) {
        for (int m = 0; m < 6; m++) {
            sld->scaling_list_pred_mode_flag[s][m] = rr(0, 1);
            sld->scaling_list_pred_matrix_id_delta[s][m] = rr(0, m);
            if (s >= 2) { sld->scaling_list_dc_coef_minus8[s - 2][m] = rr(-7, 247); }
            sld->scaling_list_delta_coef[s][m] = rr(-128, 127);
        }
    }
}

static void gen_vps(hevc_vps_t* vps, int rich)
{
    memset(vps, 0, sizeof(*vps));
    vps->vps_video_parameter_set_id = rich ? rr(0, 15) : 0;
    vps->vps_base_layer_internal_flag = 1;
    vps->vps_base_layer_available_flag = 1;
    vps->vps_max_layers_minus1 = 0;
    vps->vps_max_sub_layers_minus1 = rich ? rr(0, 2) : 0;
    vps->vps_temporal_id_nesting_flag = 1;
    gen_ptl(&vps->ptl, rich, vps->vps_max_sub_layers_minus1);
    vps->vps_sub_layer_ordering_info_present_flag = rich ? rr(0, 1) : 1;
    for (int i = 0; i < 8; i++) {
        vps->vps_max_dec_pic_buffering_minus1[i] = rich ? rr(0, 15) : 4;
        vps->vps_max_num_reorder_pics[i] = rich ? rr(0, 4) : 0;
        vps->vps_max_latency_increase_plus1[i] = rich ? rr(0, 100) : 0;
    }
    vps->vps_max_layer_id = rich ? rr(0, 3) : 0;
    vps->vps_num_layer_sets_minus1 = rich ? rr(0, 2) : 0;
    for (int i = 0; i < 4; i++) { for (int j = 0; j < 8; j++) { vps->layer_id_included_flag[i][j] = rr(0, 1); } }
    if (rich && pct(60)) {
        vps->vps_timing_info_present_flag = 1;
        vps->vps_num_units_in_tick = rr(1, 100000);
        vps->vps_time_scale = (int)(rnd64() & 0x7fffffff);
        vps->vps_poc_proportional_to_timing_flag = rr(0, 1);
        vps->vps_num_ticks_poc_diff_one_minus1 = rr(0, 1000);
        vps->vps_num_hrd_parameters = rr(0, 2);
        for (int i = 0; i < vps->vps_num_hrd_parameters; i++) {
            vps->hrd_layer_set_idx[i] = rr(0, 2);
            vps->cprms_present_flag[i] = (i > 0) ? rr(0, 1) : 0; /* [0] is never coded (App. A-8) */
#ifdef REF_SPEC
            if (i == 0) { vps->cprms_present_flag[i] = 1; } /* inferred 1 */
#endif
            gen_hrd(&vps->hrd[i], vps->vps_max_sub_layers_minus1);
            if (!vps->cprms_present_flag[i]) {
                /* common info not coded: a reader keeps zeros, which drive the sub-layer loops */
                hevc_hrd_t* hr = &vps->hrd[i];
                hr->nal_hrd_parameters_present_flag = 0; hr->vcl_hrd_parameters_present_flag = 0;
                hr->sub_pic_hrd_params_present_flag = 0;
            }
        }
    }
    vps->vps_extension_flag = 0;
}

/* st_ref_pic_set idx `idx` of `num`; consistent with reader reconstruction (App. A-5) */
static void gen_st_rps(hevc_st_ref_pic_set_t* r, int idx, int num)
{
    memset(r, 0, sizeof(*r));
    int inter = (idx != 0) && pct(40);
    if (inter) {
        int ref;
        r->inter_ref_pic_set_prediction_flag = 1;
        if (idx == num) { r->delta_idx_minus1 = rr(0, idx - 1); }
        ref = idx - (r->delta_idx_minus1 + 1);
        if (NumDeltaPocs[ref] > 20) { inter = 0; memset(r, 0, sizeof(*r)); }
        else {
            r->delta_rps_sign = rr(0, 1);
            r->abs_delta_rps_minus1 = rr(0, 3);
            for (int j = 0; j <= NumDeltaPocs[ref]; j++) {
                r->used_by_curr_pic_flag[j] = rr(0, 1);
                r->use_delta_flag[j] = r->used_by_curr_pic_flag[j] ? 0 : rr(0, 1);
#ifdef REF_SPEC /* inferred 1 when not present */
                if (r->used_by_curr_pic_flag[j]) { r->use_delta_flag[j] = 1; }
#endif
            }
        }
    }
    if (!inter) {
        r->num_negative_pics = rr(0, 4);
        r->num_positive_pics = rr(0, 3);
        for (int i = 0; i < r->num_negative_pics; i++) { r->delta_poc_s0_minus1[i] = rr(0, 3); r->used_by_curr_pic_s0_flag[i] = rr(0, 1); }
        for (int i = 0; i < r->num_positive_pics; i++) { r->delta_poc_s1_minus1[i] = rr(0, 3); r->used_by_curr_pic_s1_flag[i] = rr(0, 1); }
    }
}

static void gen_vui(hevc_sps_t* sps)
{
    hevc_vui_t* v = &sps->vui;
    memset(v, 0, sizeof(*v));
    v->aspect_ratio_info_present_flag = rr(0, 1);
    v->aspect_ratio_idc = pct(30) ? 255 : rr(0, 16);
    v->sar_width = rr(1, 65535);
    v->sar_height = rr(1, 65535);
    v->overscan_info_present_flag = rr(0, 1);
    v->overscan_appropriate_flag = rr(0, 1);
    v->video_signal_type_present_flag = rr(0, 1);
    v->video_format = rr(0, 5);
    v->video_full_range_flag = rr(0, 1);
    v->colour_description_present_flag = rr(0, 1);
    v->colour_primaries = rr(0, 12);
    v->transfer_characteristics = rr(0, 18);
    v->matrix_coefficients = rr(0, 11);
    v->chroma_loc_info_present_flag = rr(0, 1);
    v->chroma_sample_loc_type_top_field = rr(0, 5);
    v->chroma_sample_loc_type_bottom_field = rr(0, 5);
    v->neutral_chroma_indication_flag = rr(0, 1);
    v->field_seq_flag = rr(0, 1);
    v->frame_field_info_present_flag = rr(0, 1);
    v->default_display_window_flag = rr(0, 1);
    v->def_disp_win_left_offset = rr(0, 64);
    v->def_disp_win_right_offset = rr(0, 64);
    v->def_disp_win_top_offset = rr(0, 64);
    v->def_disp_win_bottom_offset = rr(0, 64);
    v->vui_timing_info_present_flag = rr(0, 1);
    v->vui_num_units_in_tick = rr(1, 100000);
    v->vui_time_scale = (int)(rnd64() & 0x7fffffff);
    v->vui_poc_proportional_to_timing_flag = rr(0, 1);
    v->vui_num_ticks_poc_diff_one_minus1 = rr(0, 1000);
    v->vui_hrd_parameters_present_flag = rr(0, 1);
    gen_hrd(&v->hrd, sps->sps_max_sub_layers_minus1);
    v->bitstream_restriction_flag = rr(0, 1);
    v->tiles_fixed_structure_flag = rr(0, 1);
    v->motion_vectors_over_pic_boundaries_flag = rr(0, 1);
    v->restricted_ref_pic_lists_flag = rr(0, 1);
    v->min_spatial_segmentation_idc = rr(0, 4095);
    v->max_bytes_per_pic_denom = rr(0, 16);
    v->max_bits_per_min_cu_denom = rr(0, 16);
    v->log2_max_mv_length_horizontal = rr(0, 15);
    v->log2_max_mv_length_vertical = rr(0, 15);
}

static void gen_sps(hevc_stream_t* h, int rich)
{
    hevc_sps_t* sps = h->sps;
    memset(sps, 0, sizeof(*sps));
    sps->sps_video_parameter_set_id = 0;
    sps->sps_max_sub_layers_minus1 = rich ? rr(0, 2) : 0;
    sps->sps_temporal_id_nesting_flag = 1;
    gen_ptl(&sps->ptl, rich, sps->sps_max_sub_layers_minus1);
    sps->sps_seq_parameter_set_id = 0; /* must stay 0: slices index h->sps by pointer arithmetic */
#ifdef REF_SPEC /* id-indexed tables: any id */
    sps->sps_seq_parameter_set_id = rich ? rr(0, 3) : 0;
    note_id(g_sps_ids, &g_n_sps_ids, sps->sps_seq_parameter_set_id);
    spec_rps_cur = sps->sps_seq_parameter_set_id;
#endif
    sps->chroma_format_idc = rich ? rr(0, 3) : 1;
    sps->separate_colour_plane_flag = (sps->chroma_format_idc == 3) ? rr(0, 1) : 0;
    if (rich) {
        static const int ws[] = {1920, 1280, 3840, 416, 832, 64, 8192};
        static const int hs[] = {1080, 720, 2160, 240, 480, 64, 4320};
        int i = rr(0, 6);
        sps->pic_width_in_luma_samples = ws[i];
        sps->pic_height_in_luma_samples = hs[i];
    } else {
        sps->pic_width_in_luma_samples = 1920;
        sps->pic_height_in_luma_samples = 1080;
    }
    sps->conformance_window_flag = rich ? rr(0, 1) : 1;
    sps->conf_win_bottom_offset = 4;
    if (rich) { sps->conf_win_left_offset = rr(0, 8); sps->conf_win_right_offset = rr(0, 8); sps->conf_win_top_offset = rr(0, 8); }
    sps->bit_depth_luma_minus8 = rich ? rr(0, 4) : 0;
    sps->bit_depth_chroma_minus8 = rich ? rr(0, 4) : 0;
    sps->log2_max_pic_order_cnt_lsb_minus4 = rich ? rr(0, 12) : 4;
    sps->sps_sub_layer_ordering_info_present_flag = rich ? rr(0, 1) : 1;
    for (int i = 0; i < 8; i++) {
        sps->sps_max_dec_pic_buffering_minus1[i] = rich ? rr(0, 15) : 4;
        sps->sps_max_num_reorder_pics[i] = rich ? rr(0, 4) : 0;
        sps->sps_max_latency_increase_plus1[i] = rich ? rr(0, 100) : 0;
    }
    sps->log2_min_luma_coding_block_size_minus3 = rich ? rr(0, 1) : 0;
    sps->log2_diff_max_min_luma_coding_block_size = rich ? rr(1, 3 - sps->log2_min_luma_coding_block_size_minus3) : 3;
    sps->log2_min_luma_transform_block_size_minus2 = 0;
    sps->log2_diff_max_min_luma_transform_block_size = 3;
    sps->max_transform_hierarchy_depth_inter = rich ? rr(0, 4) : 2;
    sps->max_transform_hierarchy_depth_intra = rich ? rr(0, 4) : 2;
    if (rich && pct(40)) {
        sps->scaling_list_enabled_flag = 1;
        sps->sps_scaling_list_data_present_flag = rr(0, 1);
        gen_scaling_list(&sps->scaling_list_data);
    }
    sps->amp_enabled_flag = rich ? rr(0, 1) : 1;
    sps->sample_adaptive_offset_enabled_flag = rich ? rr(0, 1) : 1;
    if (rich && pct(30)) {
        sps->pcm_enabled_flag = 1;
        sps->pcm_sample_bit_depth_luma_minus1 = rr(0, 15);
        sps->pcm_sample_bit_depth_chroma_minus1 = rr(0, 15);
        sps->log2_min_pcm_luma_coding_block_size_minus3 = rr(0, 2);
        sps->log2_diff_max_min_pcm_luma_coding_block_size = rr(0, 2);
        sps->pcm_loop_filter_disabled_flag = rr(0, 1);
    }
    /* the writer consults the static RPS tables while writing each set, so fill + write order matters:
     * sets are generated here in index order against the tables as left by the sets already generated;
     * we therefore update the tables the same way the writer will (updateNumDeltaPocs). */
    sps->num_short_term_ref_pic_sets = rich ? rr(0, 10) : 1;
    for (int i = 0; i < sps->num_short_term_ref_pic_sets; i++) {
        hevc_st_ref_pic_set_t* r = &sps->st_ref_pic_set[i];
        if (rich) { gen_st_rps(r, i, sps->num_short_term_ref_pic_sets); }
        else { memset(r, 0, sizeof(*r)); r->num_negative_pics = 1; r->delta_poc_s0_minus1[0] = 0; r->used_by_curr_pic_s0_flag[0] = 1; }
        /* mirror of hevc_stream.in.c:1038-1060 side effects so that later sets see the right tables */
        if (!r->inter_ref_pic_set_prediction_flag) {
            for (int j = 0; j < r->num_negative_pics; j++) {
                UsedByCurrPicS0[i][j] = r->used_by_curr_pic_s0_flag[j];
                DeltaPocS0[i][j] = (j == 0 ? 0 : DeltaPocS0[i][j - 1]) - (r->delta_poc_s0_minus1[j] + 1);
            }
            for (int j = 0; j < r->num_positive_pics; j++) {
                UsedByCurrPicS1[i][j] = r->used_by_curr_pic_s1_flag[j];
                DeltaPocS1[i][j] = (j == 0 ? 0 : DeltaPocS1[i][j - 1]) + (r->delta_poc_s1_minus1[j] + 1);
            }
        }
        updateNumDeltaPocs(r, i);
    }
    if (rich && pct(50)) {
        sps->long_term_ref_pics_present_flag = 1;
        sps->num_long_term_ref_pics_sps = rr(0, 6);
        for (int i = 0; i < sps->num_long_term_ref_pics_sps; i++) {
            sps->lt_ref_pic_poc_lsb_sps[i] = rr(0, (1 << (sps->log2_max_pic_order_cnt_lsb_minus4 + 4)) - 1);
            sps->used_by_curr_pic_lt_sps_flag[i] = rr(0, 1);
        }
    }
    sps->sps_temporal_mvp_enabled_flag = rich ? rr(0, 1) : 1;
    sps->strong_intra_smoothing_enabled_flag = rich ? rr(0, 1) : 1;
    if (rich && pct(60)) { sps->vui_parameters_present_flag = 1; gen_vui(sps); }
    if (rich && pct(30)) {
        sps->sps_extension_present_flag = 1;
        sps->sps_range_extension_flag = rr(0, 1);
        sps->sps_multilayer_extension_flag = 0;
        sps->sps_3d_extension_flag = 0;
        sps->sps_extension_5bits = 0;
        sps->sps_range_ext.transform_skip_rotation_enabled_flag = rr(0, 1);
        sps->sps_range_ext.transform_skip_context_enabled_flag = rr(0, 1);
        sps->sps_range_ext.implicit_rdpcm_enabled_flag = rr(0, 1);
        sps->sps_range_ext.explicit_rdpcm_enabled_flag = rr(0, 1);
        sps->sps_range_ext.extended_precision_processing_flag = rr(0, 1);
        sps->sps_range_ext.intra_smoothing_disabled_flag = rr(0, 1);
        sps->sps_range_ext.high_precision_offsets_enabled_flag = rr(0, 1);
        sps->sps_range_ext.persistent_rice_adaptation_enabled_flag = rr(0, 1);
        sps->sps_range_ext.cabac_bypass_alignment_enabled_flag = rr(0, 1);
    }
}

static void gen_pps(hevc_stream_t* h, int rich)
{
    hevc_pps_t* pps = h->pps;
    memset(pps, 0, sizeof(*pps));
    pps->pic_parameter_set_id = 0; /* must stay 0 */
    pps->seq_parameter_set_id = 0;
#ifdef REF_SPEC
    pps->pic_parameter_set_id = rich ? rr(0, 5) : 0;
    pps->seq_parameter_set_id = g_sps_ids[rr(0, g_n_sps_ids - 1)];
    note_id(g_pps_ids, &g_n_pps_ids, pps->pic_parameter_set_id);
#endif
    pps->init_qp_minus26 = rich ? rr(-20, 20) : 0;
    pps->cu_qp_delta_enabled_flag = rich ? rr(0, 1) : 1;
    pps->diff_cu_qp_delta_depth = rich ? rr(0, 3) : 0;
    pps->pps_loop_filter_across_slices_enabled_flag = rich ? rr(0, 1) : 1;
    if (!rich) { return; }
    pps->dependent_slice_segments_enabled_flag = rr(0, 1);
    pps->output_flag_present_flag = rr(0, 1);
    pps->num_extra_slice_header_bits = pct(30) ? rr(1, 3) : 0;
    pps->sign_data_hiding_enabled_flag = rr(0, 1);
    pps->cabac_init_present_flag = rr(0, 1);
    pps->num_ref_idx_l0_default_active_minus1 = rr(0, 5);
    pps->num_ref_idx_l1_default_active_minus1 = rr(0, 5);
    pps->constrained_intra_pred_flag = rr(0, 1);
    pps->transform_skip_enabled_flag = rr(0, 1);
    pps->pps_cb_qp_offset = rr(-12, 12);
    pps->pps_cr_qp_offset = rr(-12, 12);
    pps->pps_slice_chroma_qp_offsets_present_flag = rr(0, 1);
    pps->weighted_pred_flag = rr(0, 1);
    pps->weighted_bipred_flag = rr(0, 1);
    pps->transquant_bypass_enabled_flag = rr(0, 1);
    pps->tiles_enabled_flag = pct(40);
    pps->entropy_coding_sync_enabled_flag = pct(30);
    pps->num_tile_columns_minus1 = rr(0, 5);
    pps->num_tile_rows_minus1 = rr(0, 5);
    pps->uniform_spacing_flag = rr(0, 1);
    for (int i = 0; i < 8; i++) { pps->column_width_minus1[i] = rr(0, 10); pps->row_height_minus1[i] = rr(0, 10); }
    pps->loop_filter_across_tiles_enabled_flag = rr(0, 1);
    pps->deblocking_filter_control_present_flag = rr(0, 1);
    if (pps->deblocking_filter_control_present_flag) {
        pps->deblocking_filter_override_enabled_flag = rr(0, 1);
        pps->pps_deblocking_filter_disabled_flag = rr(0, 1);
        if (pps->pps_deblocking_filter_disabled_flag) { pps->pps_beta_offset_div2 = rr(-6, 6); pps->pps_tc_offset_div2 = rr(-6, 6); }
    }
    if (pct(30)) { pps->pps_scaling_list_data_present_flag = 1; gen_scaling_list(&pps->scaling_list_data); }
    pps->lists_modification_present_flag = rr(0, 1);
    pps->log2_parallel_merge_level_minus2 = rr(0, 4);
    pps->slice_segment_header_extension_present_flag = pct(25);
    if (pct(35)) {
        pps->pps_extension_present_flag = 1;
        pps->pps_range_extension_flag = rr(0, 1);
        pps->pps_multilayer_extension_flag = rr(0, 1);
        pps->pps_3d_extension_flag = rr(0, 1);
        pps->pps_extension_5bits = rr(0, 1); /* coded as ONE bit (App. A-6) */
        if (pps->pps_range_extension_flag) {
            hevc_pps_range_ext_t* e = &pps->pps_range_ext;
            e->log2_max_transform_skip_block_size_minus2 = pps->transform_skip_enabled_flag ? rr(0, 3) : 0;
            e->cross_component_prediction_enabled_flag = rr(0, 1);
            e->chroma_qp_offset_list_enabled_flag = rr(0, 1);
            if (e->chroma_qp_offset_list_enabled_flag) {
                e->diff_cu_chroma_qp_offset_depth = rr(0, 3);
                e->chroma_qp_offset_list_len_minus1 = rr(0, 5);
                for (int i = 0; i <= e->chroma_qp_offset_list_len_minus1; i++) { e->cb_qp_offset_list[i] = rr(-12, 12); e->cr_qp_offset_list[i] = rr(-12, 12); }
            }
            e->log2_sao_offset_scale_luma = rr(0, 4);
            e->log2_sao_offset_scale_chroma = rr(0, 4);
        }
    }
}

static int ilog2_ceil(int n) { int b = 0; while ((1 << b) < n) { b++; } return b; }

/* fill h->sh consistently with h->pps / h->sps AS THE READER SEES THEM (after resync) */
static void gen_slice(hevc_stream_t* h, int rich, int nut, int force_type)
{
    hevc_slice_header_t* sh = h->sh;
    hevc_pps_t* pps = h->pps;
    hevc_sps_t* sps = h->sps;
    memset(sh, 0, sizeof(*sh));
    sh->collocated_from_l0_flag = 1;
    sh->first_slice_segment_in_pic_flag = rich ? pct(60) : 1;
    sh->no_output_of_prior_pics_flag = rich ? rr(0, 1) : 0;
    sh->pic_parameter_set_id = 0;
#ifdef REF_SPEC /* the slice is generated against the parameter sets its ids select (the tables hold what a reader has seen) */
    sh->pic_parameter_set_id = g_pps_ids[rr(0, g_n_pps_ids - 1)];
    pps = h->pps_table[sh->pic_parameter_set_id];
    sps = h->sps_table[pps->seq_parameter_set_id];
    spec_rps_cur = pps->seq_parameter_set_id & 31;
#endif
    sh->num_ref_idx_l0_active_minus1 = pps->num_ref_idx_l0_default_active_minus1;
    sh->num_ref_idx_l1_active_minus1 = pps->num_ref_idx_l1_default_active_minus1;
    if (!sh->first_slice_segment_in_pic_flag) {
        if (pps->dependent_slice_segments_enabled_flag) { sh->dependent_slice_segment_flag = pct(30); }
        int bits = getSliceSegmentAddressBitLength(sps);
        sh->slice_segment_address = (bits > 0) ? (int)(rnd64() % (1ull << (bits > 30 ? 30 : bits))) : 0;
    }
    if (sh->dependent_slice_segment_flag) { goto tail; }
    sh->slice_type = (force_type >= 0) ? force_type : rr(0, 2);
    /* IDR pictures carry I slices only; a P/B-typed IDR slice makes the reference consult derived RPS state left
     * behind by an earlier slice (process-global tables), which no stateless parser can reproduce */
    if (nut == HEVC_NAL_UNIT_TYPE_IDR_W_RADL || nut == HEVC_NAL_UNIT_TYPE_IDR_N_LP) { sh->slice_type = HEVC_SLICE_TYPE_I; }
    sh->pic_output_flag = rr(0, 1);
    sh->colour_plane_id = rr(0, 2);
    if (nut != HEVC_NAL_UNIT_TYPE_IDR_W_RADL && nut != HEVC_NAL_UNIT_TYPE_IDR_N_LP) {
        int pocbits = sps->log2_max_pic_order_cnt_lsb_minus4 + 4;
        sh->slice_pic_order_cnt_lsb = (int)(rnd64() % (1ull << pocbits));
        int nsets = sps->num_short_term_ref_pic_sets;
        sh->short_term_ref_pic_set_sps_flag = (nsets > 0 && rich) ? pct(50) : (nsets > 0);
        if (!sh->short_term_ref_pic_set_sps_flag) {
            if (rich) { gen_st_rps(&sh->st_ref_pic_set, nsets, nsets); }
            else { sh->st_ref_pic_set.num_negative_pics = 1; sh->st_ref_pic_set.used_by_curr_pic_s0_flag[0] = 1; }
        } else if (nsets > 1) {
            sh->short_term_ref_pic_set_idx = rr(0, nsets - 1);
        }
        if (sps->long_term_ref_pics_present_flag) {
            if (sps->num_long_term_ref_pics_sps > 0) { sh->num_long_term_sps = rr(0, 3); }
            sh->num_long_term_pics = rr(0, 3);
            for (int i = 0; i < sh->num_long_term_sps + sh->num_long_term_pics; i++) {
                if (i < sh->num_long_term_sps) {
                    if (sps->num_long_term_ref_pics_sps > 1) { sh->lt_idx_sps[i] = rr(0, sps->num_long_term_ref_pics_sps - 1); }
                } else {
                    sh->poc_lsb_lt[i] = (int)(rnd64() % (1ull << pocbits));
                    sh->used_by_curr_pic_lt_flag[i] = rr(0, 1);
                }
                sh->delta_poc_msb_present_flag[i] = rr(0, 1);
                if (sh->delta_poc_msb_present_flag[i]) { sh->delta_poc_msb_cycle_lt[i] = rr(0, 40); }
            }
        }
        if (sps->sps_temporal_mvp_enabled_flag) { sh->slice_temporal_mvp_enabled_flag = rr(0, 1); }
    }
    if (sps->sample_adaptive_offset_enabled_flag) {
        sh->slice_sao_luma_flag = rr(0, 1);
        int cat = sps->separate_colour_plane_flag == 0 ? sps->chroma_format_idc : 0;
        if (cat != 0) { sh->slice_sao_chroma_flag = rr(0, 1); }
    }
    if (sh->slice_type == HEVC_SLICE_TYPE_P || sh->slice_type == HEVC_SLICE_TYPE_B) {
        sh->num_ref_idx_active_override_flag = rich ? rr(0, 1) : 0;
        if (sh->num_ref_idx_active_override_flag) {
            sh->num_ref_idx_l0_active_minus1 = rr(0, 14);
            if (sh->slice_type == HEVC_SLICE_TYPE_B) { sh->num_ref_idx_l1_active_minus1 = rr(0, 14); }
        }
        /* list modification: entries are u(ceil(log2(NumPicTotalCurr))); values only need to fit */
        sh->rpld.ref_pic_list_modification_flag_l0 = rr(0, 1);
        for (int i = 0; i < 32; i++) { sh->rpld.list_entry_l0[i] = rr(0, 1); sh->rpld.list_entry_l1[i] = 0; }
        sh->rpld.ref_pic_list_modification_flag_l1 = 0; /* never coded (App. A-4) */
        sh->mvd_l1_zero_flag = rr(0, 1);
        sh->cabac_init_flag = rr(0, 1);
        if (sh->slice_temporal_mvp_enabled_flag) {
            if (sh->slice_type == HEVC_SLICE_TYPE_B) { sh->collocated_from_l0_flag = rr(0, 1); }
            if ((sh->collocated_from_l0_flag && sh->num_ref_idx_l0_active_minus1 > 0) ||
                (!sh->collocated_from_l0_flag && sh->num_ref_idx_l1_active_minus1 > 0)) {
                sh->collocated_ref_idx = rr(0, 3);
            }
        }
        {
            hevc_pred_weight_table_t* w = &sh->pwt;
            int cat = sps->separate_colour_plane_flag == 0 ? sps->chroma_format_idc : 0;
            w->luma_log2_weight_denom = rr(0, 7);
            if (cat != 0) { w->delta_chroma_log2_weight_denom = rr(-3, 3); }
            for (int i = 0; i <= sh->num_ref_idx_l0_active_minus1; i++) {
                w->luma_weight_l0_flag[i] = rr(0, 1);
                w->chroma_weight_l0_flag[i] = (cat != 0) ? rr(0, 1) : 0;
                if (w->luma_weight_l0_flag[i]) { w->delta_luma_weight_l0[i] = rr(-128, 127); w->luma_offset_l0[i] = rr(-128, 127); }
                if (w->chroma_weight_l0_flag[i]) {
                    for (int j = 0; j < 2; j++) { w->delta_chroma_weight_l0[i][j] = rr(-128, 127); w->delta_chroma_offset_l0[i][j] = rr(-512, 511); }
                }
            }
            if (sh->slice_type == HEVC_SLICE_TYPE_B) {
                for (int i = 0; i <= sh->num_ref_idx_l1_active_minus1; i++) {
                    w->luma_weight_l1_flag[i] = rr(0, 1);
                    w->chroma_weight_l1_flag[i] = (cat != 0) ? rr(0, 1) : 0;
                    if (w->luma_weight_l1_flag[i]) { w->delta_luma_weight_l1[i] = rr(-128, 127); w->luma_offset_l1[i] = rr(-128, 127); }
                    if (w->chroma_weight_l1_flag[i]) {
                        for (int j = 0; j < 2; j++) { w->delta_chroma_weight_l1[i][j] = rr(-128, 127); w->delta_chroma_offset_l1[i][j] = rr(-512, 511); }
                    }
                }
            }
        }
        sh->five_minus_max_num_merge_cand = rr(0, 4);
    }
    sh->slice_qp_delta = rich ? rr(-26, 25) : -3;
    sh->slice_cb_qp_offset = rr(-12, 12);
    sh->slice_cr_qp_offset = rr(-12, 12);
    sh->cu_chroma_qp_offset_enabled_flag = rr(0, 1);
    if (pps->deblocking_filter_override_enabled_flag) { sh->deblocking_filter_override_flag = rr(0, 1); }
#ifdef REF_SPEC /* inherited from the PPS unless overridden (the writer's conditions read the struct) */
    sh->slice_deblocking_filter_disabled_flag = pps->pps_deblocking_filter_disabled_flag;
#endif
    if (sh->deblocking_filter_override_flag) {
        sh->slice_deblocking_filter_disabled_flag = rr(0, 1);
        if (!sh->slice_deblocking_filter_disabled_flag) { sh->slice_beta_offset_div2 = rr(-6, 6); sh->slice_tc_offset_div2 = rr(-6, 6); }
    }
    sh->slice_loop_filter_across_slices_enabled_flag = rr(0, 1);
tail:
    if (pps->tiles_enabled_flag || pps->entropy_coding_sync_enabled_flag) {
        sh->num_entry_point_offsets = pct(50) ? rr(1, 32) : 0;
        if (sh->num_entry_point_offsets > 0) {
            sh->offset_len_minus1 = rr(0, 31);
            for (int i = 0; i < sh->num_entry_point_offsets; i++) {
                uint64_t m = (sh->offset_len_minus1 == 31) ? 0xffffffffull : ((1ull << (sh->offset_len_minus1 + 1)) - 1);
                sh->entry_point_offset_minus1[i] = (int)(uint32_t)(rnd64() & m);
            }
        }
    }
    if (pps->slice_segment_header_extension_present_flag) { sh->slice_segment_header_extension_length = rr(0, 6); }
}

typedef struct { uint8_t* p; int64_t n, cap; int fail; } outbuf;
static void ob_put(outbuf* o, const uint8_t* src, int64_t n)
{
    if (o->n + n > o->cap) { o->fail = 1; return; }
    memcpy(o->p + o->n, src, (size_t)n);
    o->n += n;
}
static void ob_startcode(outbuf* o, int len4, int extra_zero_pct)
{
    static const uint8_t sc[8] = {0, 0, 0, 0, 0, 0, 0, 1};
    int zeros = len4 ? 3 : 2;
    if (extra_zero_pct > 0 && pct(extra_zero_pct)) { zeros += rr(1, 3); }
    ob_put(o, sc + (7 - zeros), zeros + 1);
}

/* write one PS NAL (type in h->nal) with the reference writer; SPS is re-read (App. A-1) */
static void emit_ps(hevc_stream_t* h, int nut, outbuf* o, int extra_zero_pct)
{
    uint8_t tmp[65536];
    h->nal->forbidden_zero_bit = 0;
    h->nal->nal_unit_type = nut;
    h->nal->nal_layer_id = 0;
    h->nal->nal_temporal_id_plus1 = 1;
    int n = write_hevc_nal_unit(h, tmp, (int)sizeof(tmp));
    if (n <= 0) { o->fail = 1; return; }
    if (nut == HEVC_NAL_UNIT_TYPE_SPS_NUT || nut == HEVC_NAL_UNIT_TYPE_PPS_NUT) {
        /* make the in-memory state equal to what any reader reconstructs from the bytes; a failing read
         * (the SPS writer drops its last partial byte, App. A-1) leaves the same partial state a reader gets */
        (void)read_hevc_nal_unit(h, tmp, n);
    }
    ob_startcode(o, 1, extra_zero_pct);
    ob_put(o, tmp, n);
}

REF_API int64_t ref_gen_stream(const ref_gen_params* gp, uint8_t* out, int64_t cap)
{
    hevc_stream_t* h = hevc_new();
    outbuf o = {out, 0, cap, 0};
    int rich = gp->profile == 1;
    g_rng = gp->seed ? gp->seed : 88172645463325252ull;
    ref_reset_static_state();
#ifdef REF_SPEC
    g_n_sps_ids = g_n_pps_ids = 0;
#endif

    gen_vps(h->vps, rich);
    emit_ps(h, HEVC_NAL_UNIT_TYPE_VPS_NUT, &o, gp->extra_zero_pct);
    gen_sps(h, rich);
    emit_ps(h, HEVC_NAL_UNIT_TYPE_SPS_NUT, &o, gp->extra_zero_pct);
    gen_pps(h, rich);
    emit_ps(h, HEVC_NAL_UNIT_TYPE_PPS_NUT, &o, gp->extra_zero_pct);

    int maxpay = gp->payload_max > 0 ? gp->payload_max : 0;
    int cap_nal = maxpay * 2 + 16384;
    int cap_hdr = 8192; /* header-only scratch for the reference writer (it callocs and copies `size` bytes per call) */
    uint8_t* tmp = (uint8_t*)malloc(cap_hdr);
    uint8_t* rb = (uint8_t*)malloc(cap_nal);
    uint8_t* fin = (uint8_t*)malloc(cap_nal);

    for (int64_t s = 0; s < gp->n_slices && !o.fail; s++) {
        if (rich && gp->ps_period > 0 && s > 0 && (s % gp->ps_period) == 0) {
            int which = rr(0, 3);
            if (which == 0) { gen_vps(h->vps, rich); emit_ps(h, HEVC_NAL_UNIT_TYPE_VPS_NUT, &o, gp->extra_zero_pct); }
            if (which <= 1) { gen_sps(h, rich); emit_ps(h, HEVC_NAL_UNIT_TYPE_SPS_NUT, &o, gp->extra_zero_pct); }
            gen_pps(h, rich); emit_ps(h, HEVC_NAL_UNIT_TYPE_PPS_NUT, &o, gp->extra_zero_pct);
        }
        if (rich && gp->unsupported_pct > 0 && pct(gp->unsupported_pct)) {
            /* a NAL type the reference dispatcher rejects (hevc_stream.c:220-221): AUD / SEI / reserved */
            static const int ts[] = {35, 39, 40, 36, 38, 10, 22, 41, 63};
            uint8_t u[16];
            int t = ts[rr(0, 8)];
            int len = rr(2, 12);
            u[0] = (uint8_t)(t << 1); u[1] = 1;
            for (int i = 2; i < len; i++) { u[i] = (uint8_t)rr(4, 255); }
            u[len - 1] = 0x80;
            ob_startcode(&o, 0, gp->extra_zero_pct);
            ob_put(&o, u, len);
        }
        int nut, ftype;
        if (!rich) {
            int idr = gp->idr_period > 0 ? (s % gp->idr_period) == 0 : (s == 0);
            nut = idr ? HEVC_NAL_UNIT_TYPE_IDR_W_RADL : HEVC_NAL_UNIT_TYPE_TRAIL_R;
            ftype = idr ? HEVC_SLICE_TYPE_I : HEVC_SLICE_TYPE_P;
        } else {
            static const int nuts[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 16, 17, 18, 19, 20, 21, 1, 1, 1, 0};
            nut = nuts[rr(0, 19)];
            ftype = -1;
        }
        h->nal->forbidden_zero_bit = 0;
        h->nal->nal_unit_type = nut;
        h->nal->nal_layer_id = 0;
        h->nal->nal_temporal_id_plus1 = rich ? rr(1, 3) : 1;
        gen_slice(h, rich, nut, ftype);
        int n = write_hevc_nal_unit(h, tmp, cap_hdr);
        if (n <= 0) { o.fail = 1; break; }
        int ns = n, rs = cap_nal;
        int r = nal_to_rbsp(tmp, &ns, rb, &rs);
        if (r <= 0) { o.fail = 1; break; }
        int hdr = r - 1; /* drop the writer's final 0x80 */
        int pay = gp->payload_min + (gp->payload_max > gp->payload_min ? (int)(rnd64() % (uint64_t)(gp->payload_max - gp->payload_min + 1)) : 0);
        if (pay < 1) { pay = 1; } /* >= 1 payload byte after every slice header */
        int zh = gp->zero_heavy_pct > 0 && pct(gp->zero_heavy_pct);
        uint8_t* q = rb + hdr;
        if (zh) {
            for (int i = 0; i < pay; i++) { uint64_t v = rnd64(); int sel = (int)(v & 7); q[i] = (sel < 3) ? 0 : (sel < 6 ? (uint8_t)(sel - 2) : (uint8_t)(v >> 8)); }
        } else {
            int i = 0;
            for (; i + 8 <= pay; i += 8) { uint64_t v = rnd64(); memcpy(q + i, &v, 8); }
            for (; i < pay; i++) { q[i] = (uint8_t)rnd64(); }
        }
        q[pay] = 0x80;
        int total = hdr + pay + 1;
        int fn = 0;
        int r2 = rbsp_to_nal(rb, &total, fin, &fn);
        if (r2 <= 0) { o.fail = 1; break; }
        ob_startcode(&o, rich ? rr(0, 1) : 0, gp->extra_zero_pct);
        ob_put(&o, fin, r2);
    }
    free(tmp); free(rb); free(fin);
    if (h->slice_data && h->slice_data->rbsp_buf) { free(h->slice_data->rbsp_buf); h->slice_data->rbsp_buf = NULL; }
    hevc_free(h);
    return o.fail ? -1 : o.n;
}

/* ------------------------------------------------------------------------------------------ */
/* CPU baseline timing loops: best-of-`reps` seconds                    */
/* ------------------------------------------------------------------------------------------ */

/* mode 0: find_nal_unit loop only; 1: + nal_to_rbsp per NAL; 2: + read_hevc_nal_unit per NAL.
 * Returns best wall time in seconds; *n_nals = NALs visited. */
REF_API double ref_time_loop(uint8_t* buf, int64_t size, int mode, int reps, int64_t* n_nals)
{
    double best = 1e30;
    uint8_t* scratch = (uint8_t*)malloc(size > 0 ? (size_t)size : 1);
    for (int it = 0; it < reps; it++) {
        hevc_stream_t* h = hevc_new();
        ref_reset_static_state();
        uint8_t* p = buf;
        int64_t sz = size, s = 0, e = 0, n = 0;
        int r;
        double t0 = now_s();
        while ((r = find_nal_big(p, sz, &s, &e)) > 0 || r == -1) {
            int len = (int)(e - s);
            if (mode == 1) { int ns = len, rs = len; (void)nal_to_rbsp(p + s, &ns, scratch, &rs); }
            else if (mode == 2) { (void)read_hevc_nal_unit(h, p + s, len); }
            n++;
            if (r == -1) { break; }
            p += e;
            sz -= e;
        }
        double t1 = now_s();
        if (t1 - t0 < best) { best = t1 - t0; }
        *n_nals = n;
        if (h->slice_data && h->slice_data->rbsp_buf) { free(h->slice_data->rbsp_buf); h->slice_data->rbsp_buf = NULL; }
        hevc_free(h);
    }
    free(scratch);
    return best;
}
