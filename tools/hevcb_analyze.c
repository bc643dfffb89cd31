/*
 * hevcb_analyze.c -- a reader in the shape of the reference's CLI (hevc_analyze.c:124-210), written against the reference's
 * own API names and built on the compatibility layer (include/hevcb_compat.h): find_nal_unit loop over the file, one
 * read_hevc_nal_unit per NAL, "!! Found NAL at offset ..." lines in the reference's format (-v), and one summary line per NAL
 * taken from the structs the call filled (instead of the per-field dump of read_debug_hevc_nal_unit).
 *
 *   gcc -O2 -Iinclude tools/hevcb_analyze.c -Lhevcbitstream_b200 -lhevcb200_compat -lhevcb200 -Wl,-rpath,$PWD/hevcbitstream_b200 -o hevcb_analyze
 *   ./hevcb_analyze [-v] stream.h265
 *
 * The whole file is read into memory (the reference refills a 32 MiB window; SURVEY 3.1 describes what that does to NALs
 * that straddle a refill).
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hevcb_compat.h"

static void summary(const hevc_stream_t* h, int rc)
{
    const int t = h->nal->nal_unit_type;
    if (rc < 0) { printf("nal_unit_type %d : not parsed (rc %d)\n", t, rc); return; }
    if ((t >= 0 && t <= 9) || (t >= 16 && t <= 21)) {
        printf("nal_unit_type %d : slice first_slice_segment_in_pic_flag %d slice_type %d slice_pic_order_cnt_lsb %d slice_qp_delta %d slice_data %d bytes\n", t,
               h->sh->first_slice_segment_in_pic_flag, h->sh->slice_type, h->sh->slice_pic_order_cnt_lsb, h->sh->slice_qp_delta, h->slice_data->rbsp_size);
    } else if (t == 32) {
        printf("nal_unit_type %d : VPS id %d max_sub_layers_minus1 %d general_profile_idc %d general_level_idc %d\n", t, h->vps->vps_video_parameter_set_id,
               h->vps->vps_max_sub_layers_minus1, h->vps->ptl.general_profile_idc, h->vps->ptl.general_level_idc);
    } else if (t == 33) {
        printf("nal_unit_type %d : SPS id %d %dx%d chroma_format_idc %d num_short_term_ref_pic_sets %d vui %d\n", t, h->sps->sps_seq_parameter_set_id,
               h->sps->pic_width_in_luma_samples, h->sps->pic_height_in_luma_samples, h->sps->chroma_format_idc, h->sps->num_short_term_ref_pic_sets,
               h->sps->vui_parameters_present_flag);
    } else if (t == 34) {
        printf("nal_unit_type %d : PPS id %d sps %d init_qp_minus26 %d tiles %d entropy_coding_sync %d\n", t, h->pps->pic_parameter_set_id,
               h->pps->seq_parameter_set_id, h->pps->init_qp_minus26, h->pps->tiles_enabled_flag, h->pps->entropy_coding_sync_enabled_flag);
    } else {
        printf("nal_unit_type %d\n", t);
    }
}

int main(int argc, char** argv)
{
    int verbose = 0, a = 1;
    if (argc > 1 && strcmp(argv[1], "-v") == 0) { verbose = 1; a = 2; }
    if (a >= argc) { fprintf(stderr, "usage: %s [-v] file\n", argv[0]); return 2; }
    FILE* f = fopen(argv[a], "rb");
    if (!f) { fprintf(stderr, "!! Error: could not open file: %s \n", strerror(errno)); return EXIT_FAILURE; }
    fseek(f, 0, SEEK_END);
    long fsz = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t* buf = (uint8_t*)calloc(1, (size_t)fsz + 16);
    if (fread(buf, 1, (size_t)fsz, f) != (size_t)fsz) { fprintf(stderr, "!! Error: read failed\n"); return EXIT_FAILURE; }
    fclose(f);
    hevc_stream_t* h = hevc_new();
    if (!h) { return EXIT_FAILURE; } /* no usable B200: the library has said why */

    uint8_t* p = buf;
    long sz = fsz;
    int nal_start = 0, nal_end = 0, r;
    while ((r = find_nal_unit(p, (int)sz, &nal_start, &nal_end)) > 0 || r == -1) { /* -1: the unterminated last NAL */
        if (verbose) {
            printf("!! Found NAL at offset %lld (0x%04llX), size %lld (0x%04llX) \n", (long long)((p - buf) + nal_start), (long long)((p - buf) + nal_start),
                   (long long)(nal_end - nal_start), (long long)(nal_end - nal_start));
        }
        summary(h, read_hevc_nal_unit(h, p + nal_start, nal_end - nal_start));
        if (r == -1) { break; }
        p += nal_end;
        sz -= nal_end;
    }
    hevc_free(h);
    free(buf);
    return 0;
}
