/* h264_stream.h -- drop-in for the byte-layer header of the reference (h264_stream.h:36-71), served by libhevcb200_compat:
 * the batched CUDA scan / strip / insert kernels answer find_nal_unit, nal_to_rbsp and rbsp_to_nal (include/hevcb_compat.h). */
#ifndef _H264_STREAM_H
#define _H264_STREAM_H 1

#include <assert.h>
#include <stdint.h>
#include <stdio.h>

#include "bs.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Table E-1: sample aspect ratio indicator */
#define SAR_Unspecified 0
#define SAR_1_1 1
#define SAR_12_11 2
#define SAR_10_11 3
#define SAR_16_11 4
#define SAR_40_33 5
#define SAR_24_11 6
#define SAR_20_11 7
#define SAR_32_11 8
#define SAR_80_33 9
#define SAR_18_11 10
#define SAR_15_11 11
#define SAR_64_33 12
#define SAR_160_99 13
#define SAR_Extended 255

extern int find_nal_unit(uint8_t* buf, int size, int* nal_start, int* nal_end);
extern int rbsp_to_nal(const uint8_t* rbsp_buf, const int* rbsp_size, uint8_t* nal_buf, int* nal_size);
extern int nal_to_rbsp(const uint8_t* nal_buf, int* nal_size, uint8_t* rbsp_buf, int* rbsp_size);

extern int more_rbsp_data(bs_t* bs);
extern int more_rbsp_trailing_data(bs_t* b);
extern int _read_ff_coded_number(bs_t* b);
extern void _write_ff_coded_number(bs_t* b, int n);
extern void debug_bytes(uint8_t* buf, int len);
void read_rbsp_trailing_bits(bs_t* b);
int intlog2(int x);
int is_slice_type(int slice_type, int cmp_type);

/* file handle for debug output (NULL: stdout) */
extern FILE* h264_dbgfile;

#ifdef __cplusplus
}
#endif

#endif
