/*
 * hevcb.h -- C ABI of libhevcb200: batched, B200-native (sm_100a) entry points for the bitstream hot
 * path of leslie-wang/hevcbitstream.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Each entry point replaces a per-NAL loop over the reference's functions (file:line in the reference):
 *
 *   hevcb_scan_strip_*   while (find_nal_unit(p, sz, &s, &e) > 0) ...      h264_nal.c:38-76, loop hevc_analyze.c:135-176
 *                        + nal_to_rbsp(nal, &nal_size, rbsp, &rbsp_size)   h264_nal.c:147-200 (called at hevc_stream.c:165)
 *   hevcb_insert_*       rbsp_to_nal(rbsp, &rbsp_size, nal, &nal_size)     h264_nal.c:92-132   (called at hevc_stream.c:1326)
 *   hevcb_parse_*        read_hevc_nal_unit(h, buf, size)                  hevc_stream.c:155-241 (+ :243-1218, bs.h:126-221)
 *   hevcb_materialize    the hevc_stream_t state read_hevc_nal_unit leaves behind (hevc_stream.h:556-569)
 *   hevcb_rewrite_*      read -> edit -> write_hevc_nal_unit -> rbsp_to_nal   hevc_stream.c:1249-1335
 *
 * There is NO CPU implementation behind these calls: every one of them fails with HEVCB_E_NODEVICE
 * when no CUDA device is usable.  `_device` variants take device pointers and a cudaStream_t (passed
 * as void*); `_host` variants take host pointers and include the host<->device copies.
 *
 * Semantics shared by all entry points
 *   - offsets are absolute byte offsets into the input buffer, 64-bit;
 *   - a buffer behaves as if it were followed by zero bytes (the reference reads up to buf[size+2];
 *     its behaviour there is defined here as "zero padded");
 *   - per-NAL status codes are the reference's: >= 0 success, -1 failure.
 */
#ifndef HEVCB_H
#define HEVCB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEVCB_API __attribute__((visibility("default")))

#define HEVCB_VERSION_MAJOR 0
#define HEVCB_VERSION_MINOR 1

/* error codes (negative, distinct from the reference's -1 per-NAL status) */
#define HEVCB_OK 0
#define HEVCB_E_NODEVICE (-100) /* no usable CUDA device / driver: there is no CPU fallback */
#define HEVCB_E_CUDA (-101)     /* a CUDA runtime call failed; see hevcb_last_error() */
#define HEVCB_E_ARG (-102)      /* invalid argument */
#define HEVCB_E_ALIGN (-103)    /* device buffers must be 16-byte aligned */
#define HEVCB_E_CAPACITY (-104) /* more NALs / bytes than the output arrays can hold */
#define HEVCB_E_NOMEM (-105)

typedef struct hevcb_ctx hevcb_ctx;

/* Result of one scan (+strip) pass.  Mirrors what the canonical loop
 *     while (find_nal_unit(p, sz, &s, &e) > 0) { ...; p += e; sz -= e; }
 * leaves behind: the NALs it visited and the outputs of the call that ended it. */
typedef struct hevcb_scan_summary {
    int64_t n_nals;       /* NAL units a reference reader visits: terminated ones + the unterminated last NAL */
    int64_t n_terminated; /* NALs for which find_nal_unit returned > 0 */
    int32_t last_rc;      /* return value of the call that ended the loop: 0 (no start / zero-length NAL) or -1 */
    int32_t overflow;     /* 1 when n_nals exceeds cap_nals (arrays hold the first cap_nals entries) */
    int64_t last_start;   /* *nal_start / *nal_end of that last call, as absolute offsets */
    int64_t last_end;
    int64_t rbsp_bytes;   /* bytes of the EPB-free image written to `rbsp` (= size - n_epb) */
    int64_t n_epb;        /* emulation prevention bytes (00 00 03) removed over the whole buffer */
} hevcb_scan_summary;

/* ---- lifecycle -------------------------------------------------------------------------------- */

/* Creates a context bound to CUDA device `device`.  Returns HEVCB_E_NODEVICE if CUDA is unusable. */
HEVCB_API int hevcb_create(int device, hevcb_ctx** out);
HEVCB_API void hevcb_destroy(hevcb_ctx* ctx);
/* Human-readable description of the last error on this context (or of the last failed hevcb_create
 * when ctx is NULL). */
HEVCB_API const char* hevcb_last_error(const hevcb_ctx* ctx);
HEVCB_API int hevcb_version(void);
/* Number of kernels this context has launched so far (bench.py reports it as gpu_launches). */
HEVCB_API int64_t hevcb_launch_count(const hevcb_ctx* ctx);
/* SM count of the bound device (grid sizing is a multiple of it). */
HEVCB_API int hevcb_sm_count(const hevcb_ctx* ctx);

/* ---- scan + EPB strip --------------------------------------------------------------------------
 *
 * One pass over `size` bytes of Annex-B data:
 *   nal_start[k], nal_end[k]   what find_nal_unit reports for the k-th NAL of the canonical loop
 *                              (h264_nal.c:38-76); the unterminated last NAL (rc -1) has nal_end == size.
 *   rbsp                       (optional) the input with every emulation prevention byte removed
 *                              (byte p is removed iff b[p]==3 && b[p-1]==0 && b[p-2]==0).  The RBSP that
 *                              nal_to_rbsp (h264_nal.c:147-200) produces for NAL k is
 *                              rbsp[rbsp_off[k] .. rbsp_end[k]); start codes stay in between.
 *   rbsp_off[k], rbsp_end[k]   extent of NAL k's RBSP inside `rbsp`; rbsp_end[k] == -1 when nal_to_rbsp
 *                              returns -1 for that NAL (00 00 0{0,1,2} inside, or 00 00 03 followed by > 3).
 * `rbsp` may be NULL (scan only; rbsp_off / rbsp_end are still produced).  `rbsp` needs `size` bytes, rounded up to a multiple
 * of 16 (the parser and the insert kernels read it with aligned vector loads).
 * All arrays hold cap_nals entries.  `summary` is written on the device for the _device variant (read it
 * after synchronising the stream).  Device pointers `buf` and `rbsp` must be 16-byte aligned.
 */
HEVCB_API int hevcb_scan_strip_device(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size,
                                      int64_t* d_nal_start, int64_t* d_nal_end, int64_t cap_nals,
                                      uint8_t* d_rbsp, int64_t* d_rbsp_off, int64_t* d_rbsp_end,
                                      hevcb_scan_summary* d_summary, void* stream);

/* Same with host buffers: copies `buf` to the device, runs the pass, copies the results back.
 * Any of rbsp / rbsp_off / rbsp_end may be NULL.  Returns HEVCB_E_CAPACITY if summary->overflow. */
HEVCB_API int hevcb_scan_strip_host(hevcb_ctx* ctx, const uint8_t* buf, int64_t size,
                                    int64_t* nal_start, int64_t* nal_end, int64_t cap_nals,
                                    uint8_t* rbsp, int64_t* rbsp_off, int64_t* rbsp_end,
                                    hevcb_scan_summary* summary);


/* ---- byte-range sharding of one stream (multi-GPU) -----------------------------------------------
 *
 * One long Annex-B stream is cut into contiguous shards, one per GPU.  Every GPU runs the scan + strip pass
 * over its own bytes only; the pieces are joined from one small record per shard (all-gathered by the caller, e.g. with
 * ncclAllGather) by hevcb_stitch, which is pure host arithmetic over those records.  No payload moves for the scan.
 *
 * Cut points: hevcb_plan_shards moves every nominal cut forward to the next position p whose preceding byte is >= 2.
 * No start code or emulation-prevention pattern can straddle such a cut backwards, so a shard needs no bytes from
 * before its start -- only a HALO of the 3..16 bytes that follow it (patterns that begin in its last two bytes).
 *
 * Inside a shard that is not the first, local NAL index 0 is the piece of the NAL that was open at the cut
 * (nal_start[0] = rbsp_off[0] = 0); whether such a NAL really exists is only known after stitching.  All offsets in the
 * per-shard arrays are local to the shard (and to its image).
 */
typedef struct hevcb_shard_summary {
    int64_t own;            /* owned bytes */
    int64_t n_nals;         /* local NAL indices in use (including the entering piece when !is_first) */
    int64_t first_empty;    /* local index of the first zero-length NAL (the reference loop stops there), -1: none */
    int64_t first_empty_start; /* its nal_start */
    int64_t rbsp_bytes;     /* bytes of the shard's EPB-free image */
    int64_t n_epb;
    int64_t head_end;       /* nal_end / rbsp_end of local NAL 0 when it was closed inside the shard, else -1 / -1 */
    int64_t head_rbsp_end;
    int64_t last_nal_start; /* nal_start / rbsp_off of local NAL n_nals-1; nal_end of it, or -1 while it is open */
    int64_t last_rbsp_off;
    int64_t last_nal_end;
    int32_t is_first, is_last;
    int32_t open_at_end;    /* local NAL n_nals-1 is still open where the shard ends */
    int32_t open_err;       /* a nal_to_rbsp error was already seen inside it */
    int32_t overflow;       /* n_nals > cap_nals */
    int32_t tail_len;       /* the last min(32, own) bytes of the shard: the last shard's end-of-stream rules are */
    uint8_t tail[32];       /* applied to them by hevcb_stitch (h264_nal.c:46-72 at the end of the buffer) */
    uint8_t head_last3[3];  /* the three bytes in front of head_end (0xFF where they lie before the shard) */
    uint8_t pad;
} hevcb_shard_summary;

/* Scan + strip of one shard.  d_buf holds own + halo bytes (halo: the bytes that follow the shard in the stream, 3..16;
 * 0 for the last shard).  Events are honoured for positions < own (last shard: < own - 8, the rest is resolved by
 * hevcb_stitch).  cap_nals >= 1. */
HEVCB_API int hevcb_scan_strip_shard_device(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t own, int64_t halo, int is_first, int is_last,
                                            int64_t* d_nal_start, int64_t* d_nal_end, int64_t cap_nals, uint8_t* d_rbsp,
                                            int64_t* d_rbsp_off, int64_t* d_rbsp_end, hevcb_shard_summary* d_summary, void* stream);

/* Host helper: cut points for n_shards shards of buf[0..size).  bounds has n_shards + 1 entries, bounds[0] = 0,
 * bounds[n_shards] = size; trailing shards may be empty.  The last non-empty shard is at least 64 bytes long (or the
 * stream is one shard). */
HEVCB_API int hevcb_plan_shards(const uint8_t* buf, int64_t size, int n_shards, int64_t* bounds);

#define HEVCB_MAX_SHARDS 64

typedef struct hevcb_stitch_patch { /* overwrite entry `index` of shard `shard`'s arrays (local coordinates) */
    int32_t shard;
    int32_t set_start; /* 1: nal_start / rbsp_off are to be written too (NALs found by the end-of-stream rules) */
    int64_t index;
    int64_t nal_start, rbsp_off;
    int64_t nal_end;   /* may exceed the shard's own size: the NAL ends in a later shard */
    int64_t rbsp_end;  /* -1, or the end in the shard's image extended by the continuation bytes (cont_bytes) */
    int32_t ends_003;  /* the NAL's last three bytes are 00 00 03: read_hevc_nal_unit reports one byte less (h264_nal.c:170) */
    int32_t pad;
} hevcb_stitch_patch;

typedef struct hevcb_stitch_result {
    hevcb_scan_summary global;               /* what hevcb_scan_strip_* reports for the whole stream (global offsets) */
    int32_t n_shards;
    int32_t n_patches;
    int64_t byte_base[HEVCB_MAX_SHARDS];     /* global offset of the shard's byte 0 */
    int64_t rbsp_base[HEVCB_MAX_SHARDS];     /* global offset of the shard's image in the whole-stream image */
    int64_t first_local[HEVCB_MAX_SHARDS];   /* local index of the first NAL the shard owns (0, or 1 when local NAL 0 is a piece) */
    int64_t n_owned[HEVCB_MAX_SHARDS];       /* NALs owned (those that start in the shard and that the reference loop visits) */
    int64_t nal_base[HEVCB_MAX_SHARDS];      /* global index of the first owned NAL */
    /* continuation of the shard's last owned NAL in later shards: image bytes [0, cont_last_bytes) of shard
     * cont_last_shard plus the whole images of the shards in between; cont_bytes is their sum (0: none) */
    int32_t cont_last_shard[HEVCB_MAX_SHARDS];
    int64_t cont_last_bytes[HEVCB_MAX_SHARDS];
    int64_t cont_bytes[HEVCB_MAX_SHARDS];
    hevcb_stitch_patch patches[HEVCB_MAX_SHARDS + 8];
} hevcb_stitch_result;

/* Joins the shard records (in stream order).  Pure host code, identical on every rank. */
HEVCB_API int hevcb_stitch(const hevcb_shard_summary* shards, int n_shards, hevcb_stitch_result* out);

/* Writes the patches that concern shard `shard` into its device arrays (one small kernel, no copies). */
HEVCB_API int hevcb_apply_patches_device(hevcb_ctx* ctx, const hevcb_stitch_result* res, int shard, int64_t* d_nal_start, int64_t* d_nal_end,
                                         int64_t* d_rbsp_off, int64_t* d_rbsp_end, int64_t cap_nals, void* stream);

/* The join and the patches in one small kernel, for a distributed step without a device->host round trip: d_records = the
 * n_shards records as the all_gather left them in device memory (stream order), d_result = device memory for the join's result (copy
 * it back when the global numbers are wanted; n_patches = -1 reports what hevcb_stitch reports as HEVCB_E_ARG).  The arrays are this
 * rank's (shard `shard`).  For C callers with an ncclComm_t: ncclAllGather(d_summary, d_records, sizeof(hevcb_shard_summary), ncclChar,
 * comm, stream) between hevcb_scan_strip_shard_device and this call is the whole exchange. */
HEVCB_API int hevcb_stitch_apply_device(hevcb_ctx* ctx, const hevcb_shard_summary* d_records, int n_shards, int shard, int64_t* d_nal_start,
                                        int64_t* d_nal_end, int64_t* d_rbsp_off, int64_t* d_rbsp_end, int64_t cap_nals,
                                        hevcb_stitch_result* d_result, void* stream);

/* ---- batched EPB insertion (rbsp_to_nal) ---------------------------------------------------------
 *
 * rbsp_to_nal (h264_nal.c:92-132) for n RBSP segments at once.  Segment k is rbsp[rbsp_off[k] .. rbsp_end[k]) of one
 * device buffer (for instance the image + extents hevcb_scan_strip_* produced, or the payloads a writer laid out);
 * segments with rbsp_end[k] < 0 (nal_to_rbsp failed) produce nothing.  The output is one contiguous buffer:
 *     out[out_off[k] .. out_off[k] + start_code_len)            00 00 01 / 00 00 00 01 (start_code_len 3 / 4; 0: none)
 *     out[out_off[k] + start_code_len .. out_off[k+1])          the bytes rbsp_to_nal writes for segment k
 * out_off has n + 1 entries; out_off[n] is the total size.  The reference's capacity check is commented out
 * (h264_nal.c:101-107: it writes past a short buffer); here summary.overflow = 1 when out_cap is too small: nothing is
 * written for the NALs that do not fit, out_off is still complete so the caller can re-run with out_off[n] bytes.
 * `rbsp` must be 16-byte aligned and readable up to the next 16-byte boundary after the last segment; the segments
 * may start anywhere.
 */
typedef struct hevcb_insert_summary {
    int64_t n_nals;
    int64_t out_bytes;  /* = out_off[n] */
    int64_t n_inserted; /* emulation prevention bytes inserted */
    int32_t overflow;
    int32_t pad;
} hevcb_insert_summary;

HEVCB_API int hevcb_insert_device(hevcb_ctx* ctx, const uint8_t* d_rbsp, const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n_nals,
                                  int start_code_len, uint8_t* d_out, int64_t out_cap, int64_t* d_out_off,
                                  hevcb_insert_summary* d_summary, void* stream);

/* Host buffers: `rbsp` holds rbsp_bytes bytes; copies in, runs the pass, copies `out` (up to out_cap) and out_off back.
 * Returns HEVCB_E_CAPACITY if summary->overflow. */
HEVCB_API int hevcb_insert_host(hevcb_ctx* ctx, const uint8_t* rbsp, int64_t rbsp_bytes, const int64_t* rbsp_off, const int64_t* rbsp_end,
                                int64_t n_nals, int start_code_len, uint8_t* out, int64_t out_cap, int64_t* out_off,
                                hevcb_insert_summary* summary);

/* ---- length-prefixed framing <-> Annex-B -----------------------------------------------------------
 *
 * The step either side of the path in real pipelines: containers (MP4 / hvcC, Matroska) store every NAL unit behind a
 * big-endian byte count of 1, 2 or 4 bytes (lengthSizeMinusOne + 1, ISO/IEC 14496-15) instead of a start code.  The NAL bytes
 * themselves are identical in both framings (emulation prevention included), so a conversion copies them verbatim behind a
 * freshly written prefix, on the device, from extents that are already there:
 *   Annex-B -> length-prefixed   extents = nal_start / nal_end of hevcb_scan_strip_*;        start_code_len 0, len_size 1 | 2 | 4
 *   length-prefixed -> Annex-B   extents = hevcb_lenpref_index_device (below);               start_code_len 3 | 4, len_size 0
 * out[out_off[k] ..) = prefix, then buf[nal_start[k] .. nal_end[k]); out_off has n + 1 entries.  A NAL longer than the length
 * field can express is written with the low bytes of its length (the caller chose len_size).  `buf` must be readable up to the
 * next 16-byte boundary behind its last byte. */
HEVCB_API int hevcb_reframe_device(hevcb_ctx* ctx, const uint8_t* d_buf, const int64_t* d_nal_start, const int64_t* d_nal_end, int64_t n_nals,
                                   int start_code_len, int len_size, uint8_t* d_out, int64_t out_cap, int64_t* d_out_off,
                                   hevcb_insert_summary* d_summary, void* stream);
/* Finds the NAL units of length-prefixed data.  The length fields chain (each tells where the next one is), so the walk is
 * sequential inside a sample; samples are independent and walked in parallel: d_sample_off[n_samples + 1] = the sample
 * boundaries the container's sample table provides (NULL: the whole buffer is one sample).  d_total[0] = NALs found (extents
 * beyond cap_nals are not stored), d_total[1] = samples whose chain runs past their end (their NALs up to the break are kept). */
HEVCB_API int hevcb_lenpref_index_device(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, int len_size, const int64_t* d_sample_off, int64_t n_samples,
                                         int64_t* d_nal_start, int64_t* d_nal_end, int64_t cap_nals, int64_t* d_total, void* stream);

/* ---- batched header parse ----------------------------------------------------------------------
 *
 * read_hevc_nal_unit (hevc_stream.c:155-241) for every NAL of a stream at once, on the results of
 * hevcb_scan_strip_* (the RBSP image and the per-NAL extents).  Per NAL k:
 *   rc[k]        the reference's return value: bytes consumed (nal size, minus one if a trailing 00 00 03 was
 *                dropped) or -1 (nal_to_rbsp error, unsupported NAL type, or the reader ran past the end)
 *   nal_hdr[k]   nal_unit_type | nal_layer_id << 8 | nal_temporal_id_plus1 << 16, what the call leaves in h->nal;
 *                -1 when nal_to_rbsp failed (the reference returns before touching h->nal)
 *   kind[k]      which struct the NAL wrote: HEVCB_KIND_{NONE,VPS,SPS,PPS,SLICE} (hevcb_layout.h)
 *   pairs        (pair_field, pair_value)[pair_off[k] .. pair_off[k+1]): every syntax element the reference stores,
 *                in parse order; pair_field is the offset in ints inside hevc_vps_t / hevc_sps_t / hevc_pps_t /
 *                hevc_slice_header_t.  Zero-filling the struct and scattering the pairs reproduces, bit for bit,
 *                what read_hevc_nal_unit leaves in hevc_stream_t (hevcb_materialize does exactly that).
 *   hdr_end[k]   slices: RBSP byte offset of the cursor after byte_alignment(); the reference's slice_data copy
 *                (hevc_stream.c:605-613) is rbsp[rbsp_off[k] + hdr_end[k] + 1 .. rbsp_end[k])
 *   cols         SoA columns for slices, laid out [8][n]: slice_type, slice_qp_delta, slice_pic_order_cnt_lsb,
 *                first_slice_segment_in_pic_flag, slice_segment_address, dependent_slice_segment_flag,
 *                num_entry_point_offsets, short_term_ref_pic_set_idx
 *   ubflag[k]    non-zero when the NAL drives the reference out of its array bounds (undefined behaviour there);
 *                such elements are skipped, every other field still matches
 * Dependencies follow the reference, not the HEVC spec: a slice is parsed against the most recent SPS NAL and PPS
 * NAL that precede it in the stream, resolved on the device with an ordinal scan.
 */
typedef struct hevcb_parse_buffers { /* device pointers for hevcb_parse_device, host pointers inside hevcb_stream_index */
    int32_t* rc;          /* [n] */
    int32_t* nal_hdr;     /* [n] */
    uint8_t* kind;        /* [n] */
    uint8_t* ubflag;      /* [n] */
    int32_t* hdr_end;     /* [n] */
    int32_t* cols;        /* [8 * n] */
    int64_t* pair_off;    /* [n + 1] */
    uint32_t* pair_field; /* [cap_pairs] */
    int32_t* pair_value;  /* [cap_pairs] */
    int64_t cap_pairs;
    uint32_t* pair_pos;   /* NULL, or [cap_pairs]: selects the TRACE variant of the parse, see below */
    uint32_t flags;       /* HEVCB_PARSE_*; 0 = the reference's behaviour */
    uint32_t pad;
} hevcb_parse_buffers;

/* Extension mode (flags & HEVCB_PARSE_AUX).  The reference defines readers for access unit delimiters, end of
 * sequence / bitstream, filler data and SEI (hevc_stream.in.c:499-573) but never dispatches them: read_hevc_nal_unit returns -1
 * for types 35..40, which stays the default here.  With the flag those NALs are parsed: kind[k] = HEVCB_KIND_AUX, rc[k] = bytes
 * consumed (or -1 when the reader ran past the end), and the pair list holds, per NAL type:
 *   35 AUD     (HEVCB_AUX_AUD_PIC_TYPE, primary_pic_type)
 *   36 / 37    nothing (no payload)
 *   38 filler  (HEVCB_AUX_FD_FF_BYTES, number of ff_byte)
 *   39 / 40    per sei_message (ff-coded payloadType / payloadSize, h264_stream.c:88-98; raw payload, h264_sei.c:75-92), repeated while
 *              more_rbsp_data (h264_stream.c:62-84): (HEVCB_AUX_SEI_TYPE, t) (HEVCB_AUX_SEI_SIZE, n) (HEVCB_AUX_SEI_OFFSET, o): the
 *              payload is rbsp[rbsp_off[k] + o .. + n); the bytes stay in the image */
#define HEVCB_PARSE_AUX 1u
/* Spec-correct mode (flags & HEVCB_PARSE_SPEC).  The default walk reproduces the reference including where it departs
 * from the HEVC syntax: that is what parity is measured against, and it mis-parses most real encoder output.
 * With the flag the walk follows the standard in exactly these places (the oracle is the reference's own template with the same
 * fixes, regenerated by oracle/make_spec_ref.py): the SPS ends with rbsp_trailing_bits; a slice resolves its PPS by
 * pic_parameter_set_id and that PPS's SPS by seq_parameter_set_id (the most recent NAL with the id in front of the slice; derived
 * RPS tables per SPS); ref_pic_list_modification_flag_l1 / list_entry_l1 are read; use_delta_flag, fixed_pic_rate_within_cvs_flag
 * and cprms_present_flag[0] take their inferred values; the PPS deblocking offsets are present when the filter is not disabled and
 * slices inherit pps_deblocking_filter_disabled_flag; cpb_cnt_minus1 is present when low_delay_hrd_flag is 0 and a sub-layer has
 * cpb_cnt_minus1 + 1 entries.  Structures stay the reference's (pairs index the same structs).  hevcb_rewrite_device writes the results
 * of such a parse by the same rules (the SPS with its trailing bits, slices against the parameter sets their ids select).  Not
 * available for shard parses (hevcb_parse_shard_device): HEVCB_E_ARG. */
#define HEVCB_PARSE_SPEC 2u
#define HEVCB_AUX_AUD_PIC_TYPE 0u
#define HEVCB_AUX_FD_FF_BYTES 8u
#define HEVCB_AUX_SEI_TYPE 16u
#define HEVCB_AUX_SEI_SIZE 17u
#define HEVCB_AUX_SEI_OFFSET 18u

/* Trace variant (pair_pos != NULL): read_debug_hevc_nal_unit (hevc_stream.c:2343-3436) instead of read_hevc_nal_unit.  The list
 * of NAL k then holds one record per line the reference prints for that NAL ("%ld.%d: <expr>: %d \n", process.pl:90-113), in
 * print order: pair_pos = bit position of the reader before the element (byte = pos >> 3, bits_left = 8 - (pos & 7)),
 * pair_value = the value printed, pair_field = the struct member's field index, or HEVCB_TRACE_SPECIAL | id for the lines that
 * are not struct members (NAL header, f(n, v) elements; hevcb_trace_name resolves both to the text the reference prints).
 * Values the reader stores without reading bits carry HEVCB_TRACE_SILENT (not printed) so that hevcb_materialize still rebuilds
 * the struct from the list.  NALs of unsupported types contribute their four NAL header lines (rc stays -1).  The generated
 * read_debug variant reads sub_layer_level_idc with ONE bit where read_hevc_nal_unit reads eight (hevc_stream.c:2939 vs :751);
 * the trace variant follows it, so for such streams its state differs from the plain parse exactly as the reference's does.
 * hevcb_rewrite_device needs the results of a plain parse. */
#define HEVCB_TRACE_SPECIAL 0x80000000u
#define HEVCB_TRACE_SILENT 0x40000000u
#define HEVCB_TRACE_OPEN_LINE 19 /* id: position prefix only, no text, no newline (hevc_stream.c:3147) */
/* Text of a trace record as the reference prints it ("sps->sps_max_dec_pic_buffering_minus1 [ i ]", "rbsp_stop_one_bit", ...).
 * kind: HEVCB_KIND_* of the NAL.  Returns the length, 0 for records that print no text, -1 for an unknown code. */
HEVCB_API int hevcb_trace_name(int kind, uint32_t code, char* out, int cap);

typedef struct hevcb_parse_summary {
    int64_t n_nals;
    int64_t n_ok;     /* NALs for which read_hevc_nal_unit returns >= 0 */
    int64_t n_pairs;  /* total syntax elements; > cap_pairs means the pair arrays are incomplete (overflow = 1) */
    int64_t n_vps, n_sps, n_pps, n_slices;
    int32_t overflow;
    int32_t pad;
} hevcb_parse_summary;

HEVCB_API int hevcb_parse_device(hevcb_ctx* ctx, const uint8_t* d_buf, const int64_t* d_nal_start, const int64_t* d_nal_end,
                                 const uint8_t* d_rbsp, const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n_nals,
                                 const hevcb_parse_buffers* out, hevcb_parse_summary* d_summary, void* stream);

/* Parsing a shard of a stream (byte-range sharding): slices at the start of a shard depend on the last SPS / PPS
 * NAL of an EARLIER shard.  That state travels as two opaque blobs (hevcb_ps_context_bytes: the handful of SPS / PPS fields
 * slices need + the derived RPS tables that are file-static in the reference, hevc_stream.in.c:26-32; ~9 KB + 80 B) which
 * the ranks exchange with one all_gather.  All pointers of the chain are HOST pointers. */
typedef struct hevcb_parse_chain {
    const void* sps_in; /* state entering the shard; NULL = the zeroed state of hevc_new() */
    const void* pps_in;
    void* sps_out;      /* state after the shard's last SPS / PPS NAL (= the entering state when it has none); may be NULL */
    void* pps_out;
    int64_t buf_size;   /* bytes readable at d_buf (owned + halo).  A NAL that ends beyond it gets rc = its size; the caller
                           corrects it with hevcb_stitch_patch.ends_003 */
} hevcb_parse_chain;
HEVCB_API int hevcb_ps_context_bytes(int64_t* sps_bytes, int64_t* pps_bytes);
HEVCB_API int hevcb_parse_shard_device(hevcb_ctx* ctx, const uint8_t* d_buf, const int64_t* d_nal_start, const int64_t* d_nal_end,
                                       const uint8_t* d_rbsp, const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n_nals,
                                       const hevcb_parse_buffers* out, hevcb_parse_summary* d_summary, const hevcb_parse_chain* chain,
                                       void* stream);

/* ---- header rewrite ------------------------------------------------------------------------------
 *
 * The reference's edit loop: read_hevc_nal_unit -> change fields of h->sh / h->sps ... ->
 * write_hevc_nal_unit (hevc_stream.c:1249-1335) -> rbsp_to_nal, for every NAL of a stream at once:
 *   slices        new NAL = rbsp_to_nal( written header without the writer's final 0x80 byte ++ the original RBSP
 *                 from the old header end (hdr_end[k]) on ); the writer emits no slice data itself
 *   VPS/SPS/PPS   new NAL = what write_hevc_nal_unit produces from the parsed struct (+ edits)
 *   anything else (unsupported types, NALs the reader or the writer fails on) is copied through unchanged
 * Everything between two NALs (start codes, zero bytes) and after the last one is copied verbatim.
 * An edit names a field by (kind, field index in ints inside the kind's struct, see hevcb_layout.h / hevcb_field_index)
 * and must not change which syntax elements are present.
 */
#define HEVCB_EDIT_ADD 0
#define HEVCB_EDIT_SET 1
#define HEVCB_EDIT_XOR 2
#define HEVCB_MAX_EDITS 8
typedef struct hevcb_edit_rule {
    int32_t kind;   /* HEVCB_KIND_{VPS,SPS,PPS,SLICE} */
    uint32_t field;
    int32_t op;     /* HEVCB_EDIT_* */
    int32_t arg;
} hevcb_edit_rule;
typedef struct hevcb_edit_set {
    int32_t n;
    hevcb_edit_rule e[HEVCB_MAX_EDITS];
} hevcb_edit_set;

typedef struct hevcb_rewrite_summary {
    int64_t n_nals;
    int64_t n_rewritten; /* NALs that went through the writer (the rest was copied through) */
    int64_t out_bytes;
    int64_t n_inserted;  /* emulation prevention bytes inserted */
    int32_t overflow;    /* out_cap too small: out_start / out_end are complete, bytes of NALs that do not fit are missing */
    int32_t pad;
} hevcb_rewrite_summary;

/* Must follow hevcb_scan_strip_device + hevcb_parse_device on the same context, stream and arrays (it uses the
 * parameter-set tables the parse left on the device).  out_start[k] / out_end[k]: extent of NAL k in `out`. */
HEVCB_API int hevcb_rewrite_device(hevcb_ctx* ctx, const uint8_t* d_buf, int64_t size, const int64_t* d_nal_start, const int64_t* d_nal_end,
                                   const uint8_t* d_rbsp, const int64_t* d_rbsp_off, const int64_t* d_rbsp_end, int64_t n_nals,
                                   const hevcb_parse_buffers* parsed, const hevcb_edit_set* edits, uint8_t* d_out, int64_t out_cap,
                                   int64_t* d_out_start, int64_t* d_out_end, hevcb_rewrite_summary* d_summary, void* stream);

/* write_hevc_nal_unit (hevc_stream.c:1249-1335) for ONE NAL from caller-owned structs (host pointers to hevc_vps_t / hevc_sps_t /
 * hevc_pps_t / hevc_slice_header_t, hevcb_layout.h): the struct selected by nal_unit_type is written, a slice against the given
 * SPS and PPS; then rbsp_to_nal.  `size` is the caller's buffer size as in the reference (the writer gets size * 3 / 4 bytes of
 * RBSP room).  *nal_bytes: NAL size, or -1 where the reference returns -1 (unsupported type, overrun).  The compatibility layer's
 * write_hevc_nal_unit is this call; batches should use hevcb_rewrite_device. */
HEVCB_API int hevcb_write_nal_host(hevcb_ctx* ctx, int nal_unit_type, int nal_layer_id, int nal_temporal_id_plus1, const void* vps, const void* sps,
                                   const void* pps, const void* sh, uint8_t* nal_out, int64_t size, int64_t* nal_bytes);

/* Field index (offset in ints) of a member of the struct of `kind`, by name: "slice_qp_delta", "vui.video_full_range_flag",
 * "pwt.luma_offset_l0[3]", "st_ref_pic_set[2].delta_poc_s0_minus1[0]".  -1 when the path does not exist. */
HEVCB_API int64_t hevcb_field_index(int kind, const char* path);

/* Host-side index of a whole stream: the outputs of hevcb_scan_strip + hevcb_parse in caller-allocated host arrays. */
typedef struct hevcb_stream_index {
    int64_t cap_nals;   /* capacity of the per-NAL arrays */
    int64_t* nal_start; /* [cap_nals] */
    int64_t* nal_end;
    int64_t* rbsp_off;
    int64_t* rbsp_end;
    uint8_t* rbsp;      /* optional, `size` bytes: the EPB-free image */
    hevcb_parse_buffers p; /* host arrays sized for cap_nals / p.cap_pairs */
    hevcb_scan_summary scan;
    hevcb_parse_summary parse;
} hevcb_stream_index;

/* Annex-B bytes in host memory -> complete index (scan + strip + parse on the device, results copied back). */
HEVCB_API int hevcb_index_host(hevcb_ctx* ctx, const uint8_t* buf, int64_t size, hevcb_stream_index* idx);

/* Same, continuing from / handing on the parameter-set state of an earlier call (hevcb_parse_chain, all host pointers):
 * what lets a caller feed a stream piece by piece, like the reference's one-NAL-at-a-time read_hevc_nal_unit. */
HEVCB_API int hevcb_index_host_chain(hevcb_ctx* ctx, const uint8_t* buf, int64_t size, hevcb_stream_index* idx, const hevcb_parse_chain* chain);

/* ---- the bit layer on its own --------------------------------------------------------------------
 * bs_t's read / write calls (bs.h:126-331) as a script executed by the DEVICE bit reader (64-bit window, clz exp-Golomb) and
 * bit writer the parser / writer kernels are built on: the hook that pins them directly against the reference's known answers
 * (SURVEY Appendix B), independent of any syntax walk.  After op i: values[i] (reads), bitpos[i] = bits consumed so far (byte
 * = pos >> 3, bits_left = 8 - (pos & 7)), overrun[i] = bs_overrun().  The writer returns the bytes as the reference leaves them
 * in a zeroed buffer, the bit count and bs_overrun(); like bs_write_u1 it drops what does not fit into `cap` bytes. */
#define HEVCB_BS_U 0    /* u(n) / f(n) */
#define HEVCB_BS_U1 1
#define HEVCB_BS_U8 2
#define HEVCB_BS_UE 3
#define HEVCB_BS_SE 4
#define HEVCB_BS_SKIP 5 /* bs_skip_u(n); read side only */
typedef struct hevcb_bs_op {
    int32_t kind;
    int32_t n;     /* bit count for U / SKIP */
    int32_t value; /* value to write (write side) */
    int32_t pad;
} hevcb_bs_op;
HEVCB_API int hevcb_bs_read_host(hevcb_ctx* ctx, const uint8_t* bytes, int64_t size, const hevcb_bs_op* ops, int n_ops, int32_t* values,
                                 int64_t* bitpos, int32_t* overrun);
HEVCB_API int hevcb_bs_write_host(hevcb_ctx* ctx, const hevcb_bs_op* ops, int n_ops, uint8_t* out, int64_t cap, int64_t* bits_written, int32_t* overrun);

/* Header parse of RBSPs the caller already holds (NAL payloads without start codes and without emulation prevention bytes, e.g.
 * what nal_to_rbsp returned): n segments rbsp[rbsp_off[k] .. rbsp_end[k]) of one host buffer (a segment with rbsp_end[k] < 0 is
 * treated as a failed nal_to_rbsp).  `out` holds HOST arrays sized for n / out->cap_pairs; rc[k] is the RBSP size or -1.  An empty
 * segment is parsed like the reference parses a zero-length NAL: every read yields 0 bits (what hevc_analyze does with the
 * zero-length "last NAL" of a window, hevc_analyze.c:190-205). */
HEVCB_API int hevcb_parse_rbsp_host(hevcb_ctx* ctx, const uint8_t* rbsp, int64_t rbsp_bytes, const int64_t* rbsp_off, const int64_t* rbsp_end,
                                    int64_t n_nals, const hevcb_parse_buffers* out, hevcb_parse_summary* summary, const hevcb_parse_chain* chain);

/* Applies NAL k of an index to caller-owned structs the way read_hevc_nal_unit updates hevc_stream_t: h->nal always
 * (unless nal_to_rbsp failed), and the struct selected by kind[k] is zeroed and refilled.  The struct pointers are the
 * reference-layout types of hevcb_layout.h passed as void* (any of them may be NULL).  Calling it for k = 0, 1, 2, ...
 * reproduces the evolving state of the reference's single-stream reader.  Returns rc[k]. */
HEVCB_API int hevcb_materialize(const hevcb_stream_index* idx, int64_t k, void* nal, void* vps, void* sps, void* pps, void* sh);

#ifdef __cplusplus
}
#endif

#endif /* HEVCB_H */
