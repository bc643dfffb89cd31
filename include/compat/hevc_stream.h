/* hevc_stream.h -- drop-in for the reference's public HEVC header (hevc_stream.h:21-659): the parsed-syntax structs (field for
 * field: include/hevcb_layout.h, sizes checked at compile time), hevc_stream_t, the NAL / slice type codes and the entry points,
 * served by libhevcb200_compat on top of the batched CUDA kernels (include/hevcb_compat.h). */
#ifndef _HEVC_STREAM_H
#define _HEVC_STREAM_H 1

#include <assert.h>
#include <stdint.h>
#include <stdio.h>

#include "bs.h"
#include "h264_sei.h"
#include "../hevcb_compat.h"

#ifdef __cplusplus
extern "C" {
#endif

/* array bounds (hevc_stream.h:21-35) */
#define MAX_NUM_SUBLAYERS 32
#define MAX_NUM_HRD_PARAM 10
#define MAX_CPB_CNT 32
#define MAX_NUM_NEGATIVE_PICS 32
#define MAX_NUM_POSITIVE_PICS 32
#define MAX_NUM_REF_PICS_L0 32
#define MAX_NUM_REF_PICS_L1 32
#define MAX_NUM_SHORT_TERM_REF_PICS 32
#define MAX_NUM_LONG_TERM_REF_PICS 32
#define MAX_NUM_PALLETTE_PREDICTOR 32
#define MAX_NUM_CHROMA_QP_OFFSET_LST 32
#define MAX_NUM_ENTRY_POINT_OFFSET 32
#define MAX_NUM_TILE_COLUMN 32
#define MAX_NUM_TILE_ROW 32

/* Table 7-1: NAL unit type codes */
#define HEVC_NAL_UNIT_TYPE_TRAIL_N 0
#define HEVC_NAL_UNIT_TYPE_TRAIL_R 1
#define HEVC_NAL_UNIT_TYPE_TSA_N 2
#define HEVC_NAL_UNIT_TYPE_TSA_R 3
#define HEVC_NAL_UNIT_TYPE_STSA_N 4
#define HEVC_NAL_UNIT_TYPE_STSA_R 5
#define HEVC_NAL_UNIT_TYPE_RADL_N 6
#define HEVC_NAL_UNIT_TYPE_RADL_R 7
#define HEVC_NAL_UNIT_TYPE_RASL_N 8
#define HEVC_NAL_UNIT_TYPE_RASL_R 9
#define HEVC_NAL_UNIT_TYPE_RSV_VCL_N10 10
#define HEVC_NAL_UNIT_TYPE_RSV_VCL_R11 11
#define HEVC_NAL_UNIT_TYPE_RSV_VCL_N12 12
#define HEVC_NAL_UNIT_TYPE_RSV_VCL_R13 13
#define HEVC_NAL_UNIT_TYPE_RSV_VCL_N14 14
#define HEVC_NAL_UNIT_TYPE_RSV_VCL_R15 15
#define HEVC_NAL_UNIT_TYPE_BLA_W_LP 16
#define HEVC_NAL_UNIT_TYPE_BLA_W_RADL 17
#define HEVC_NAL_UNIT_TYPE_BLA_N_LP 18
#define HEVC_NAL_UNIT_TYPE_IDR_W_RADL 19
#define HEVC_NAL_UNIT_TYPE_IDR_N_LP 20
#define HEVC_NAL_UNIT_TYPE_CRA_NUT 21
#define HEVC_NAL_UNIT_TYPE_RSV_IRAP_VCL22 22
#define HEVC_NAL_UNIT_TYPE_RSV_IRAP_VCL23 23
#define HEVC_NAL_UNIT_TYPE_RSV_VCL24 24
#define HEVC_NAL_UNIT_TYPE_RSV_VCL25 25
#define HEVC_NAL_UNIT_TYPE_RSV_VCL26 26
#define HEVC_NAL_UNIT_TYPE_RSV_VCL27 27
#define HEVC_NAL_UNIT_TYPE_RSV_VCL28 28
#define HEVC_NAL_UNIT_TYPE_RSV_VCL29 29
#define HEVC_NAL_UNIT_TYPE_RSV_VCL30 30
#define HEVC_NAL_UNIT_TYPE_RSV_VCL31 31
#define HEVC_NAL_UNIT_TYPE_VPS_NUT 32
#define HEVC_NAL_UNIT_TYPE_SPS_NUT 33
#define HEVC_NAL_UNIT_TYPE_PPS_NUT 34
#define HEVC_NAL_UNIT_TYPE_AUD_NUT 35
#define HEVC_NAL_UNIT_TYPE_EOS_NUT 36
#define HEVC_NAL_UNIT_TYPE_EOB_NUT 37
#define HEVC_NAL_UNIT_TYPE_FD_NUT 38
#define HEVC_NAL_UNIT_TYPE_PREFIX_SEI_NUT 39
#define HEVC_NAL_UNIT_TYPE_SUFFIX_SEI_NUT 40
#define MAX_HEVC_VAL_UNIT_TYPE 40

/* Table 7-7: slice_type */
#define HEVC_SLICE_TYPE_B 0
#define HEVC_SLICE_TYPE_P 1
#define HEVC_SLICE_TYPE_I 2

#define HEVC_PROFILE_BASELINE 66
#define HEVC_PROFILE_MAIN 77
#define HEVC_PROFILE_EXTENDED 88
#define HEVC_PROFILE_HIGH 100

/* file handle for debug output */
extern FILE* h264_dbgfile;

static inline long decimal_to_binary(int n) /* the digits of n in base 2, read as a decimal number (write_debug's value format) */
{
    long digits = 0, place = 1;
    while (n != 0) {
        digits += (long)(n % 2) * place;
        n /= 2;
        place *= 10;
    }
    return digits;
}

#ifdef __cplusplus
}
#endif

#endif
