/* h264_sei.h -- drop-in for the reference's SEI container header (h264_sei.h:37-66): sei_t and its helpers.  In the reference the
 * HEVC SEI dispatch is compiled out (HAVE_SEI is never defined, hevc_stream.in.c:179-183): SEI NALs make read_hevc_nal_unit
 * return -1, which libhevcb200 reproduces; these helpers only exist so that callers of the header link. */
#ifndef _H264_SEI_H
#define _H264_SEI_H 1

#include <stdbool.h>
#include <stdint.h>

#include "bs.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct
{
    int payloadType;
    int payloadSize;
    union
    {
        uint8_t* data;
    };
} sei_t;

sei_t* sei_new();
void sei_free(sei_t* s);
void read_sei_end_bits(bs_t* b);
void read_sei_payload(sei_t* s, bs_t* b);
void write_sei_payload(sei_t* s, bs_t* b);
void read_debug_sei_payload(sei_t* s, bs_t* b);

/* D.1 SEI payload types */
#define SEI_TYPE_BUFFERING_PERIOD 0
#define SEI_TYPE_PIC_TIMING 1
#define SEI_TYPE_PAN_SCAN_RECT 2
#define SEI_TYPE_FILLER_PAYLOAD 3
#define SEI_TYPE_USER_DATA_REGISTERED_ITU_T_T35 4
#define SEI_TYPE_USER_DATA_UNREGISTERED 5
#define SEI_TYPE_RECOVERY_POINT 6
#define SEI_TYPE_DEC_REF_PIC_MARKING_REPETITION 7
#define SEI_TYPE_SPARE_PIC 8
#define SEI_TYPE_SCENE_INFO 9
#define SEI_TYPE_SUB_SEQ_INFO 10
#define SEI_TYPE_SUB_SEQ_LAYER_CHARACTERISTICS 11
#define SEI_TYPE_SUB_SEQ_CHARACTERISTICS 12
#define SEI_TYPE_FULL_FRAME_FREEZE 13
#define SEI_TYPE_FULL_FRAME_FREEZE_RELEASE 14
#define SEI_TYPE_FULL_FRAME_SNAPSHOT 15
#define SEI_TYPE_PROGRESSIVE_REFINEMENT_SEGMENT_START 16
#define SEI_TYPE_PROGRESSIVE_REFINEMENT_SEGMENT_END 17
#define SEI_TYPE_MOTION_CONSTRAINED_SLICE_GROUP_SET 18
#define SEI_TYPE_FILM_GRAIN_CHARACTERISTICS 19
#define SEI_TYPE_DEBLOCKING_FILTER_DISPLAY_PREFERENCE 20
#define SEI_TYPE_STEREO_VIDEO_INFO 21

#ifdef __cplusplus
}
#endif

#endif
