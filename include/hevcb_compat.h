/*
 * hevcb_compat.h -- the reference's own per-NAL API (same names, signatures, return conventions and struct layouts),
 * served by the B200 library.  A program written against leslie-wang/hevcbitstream's headers
 *     find_nal_unit / nal_to_rbsp / rbsp_to_nal          h264_stream.h:54-57  (h264_nal.c:38-200)
 *     hevc_new / hevc_free / peek_hevc_nal_unit           hevc_stream.h:571-572, hevc_nal.c:34,64,97
 *     read_hevc_nal_unit / write_hevc_nal_unit            hevc_stream.c:155-241 / :1249-1335
 *     read_debug_hevc_nal_unit, debug_bytes, h264_dbgfile hevc_stream.c:2343, h264_stream.c:117,33
 *     more_rbsp_data, ff-coded numbers, sei_new/free ...  h264_stream.c:42-137, h264_sei.c (declared in include/compat/*.h)
 * links against libhevcb200_compat.so + libhevcb200.so instead of libhevcbitstream and behaves the same.  include/compat/
 * holds headers under the reference's own names (bs.h, h264_stream.h, h264_sei.h, hevc_stream.h): the reference's
 * hevc_analyze.c compiles against them unmodified and prints the same bytes (tests/test_compat_gpu.py).  Every call is
 * executed by the CUDA kernels of libhevcb200 (there is no CPU implementation: without a usable B200 hevc_new() returns
 * NULL and the byte-layer calls return -1 after printing the reason to stderr).
 *
 * The per-NAL calls are a compatibility path, not the fast path: one call = one trip to the GPU.  Two things keep a
 * legacy reader loop fast nevertheless:
 *   - find_nal_unit(p, n, ...) scans the WHOLE remaining buffer on its first call and answers the following calls of the
 *     canonical loop `while (find_nal_unit(p, sz, &s, &e) > 0) { ...; p += e; sz -= e; }` from that result;
 *   - new code should call the batched entry points of hevcb.h.
 * Threads: the reference's parser is not re-entrant (h264_dbgfile, file-static tables) and neither is this one; its byte-layer
 * functions are pure, here they share one context and the scan cache, so every exported function takes one process-wide lock:
 * concurrent callers are serialised, never racing.  The cached scan is validated against the buffer's size and a digest of its first
 * and last 64 bytes; a caller that rewrites the MIDDLE of a buffer between two calls of one find_nal_unit loop must restart the loop.
 * read_hevc_nal_unit parses slices against the parameter sets it has READ through this handle, not against edits made to h->sps /
 * h->pps in between (write_hevc_nal_unit does use the caller's structs).
 */
#ifndef HEVCB_COMPAT_H
#define HEVCB_COMPAT_H

#include <stdint.h>
#include <stdio.h>

#include "hevcb_layout.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int rbsp_size;
    uint8_t* rbsp_buf;
} hevc_slice_data_rbsp_t; /* hevc_stream.h:532-536 */

typedef struct {
    int primary_pic_type;
} hevc_aud_t; /* hevc_stream.h:544-547 */

typedef struct { /* hevc_stream.h:556-569 */
    hevc_nal_t* nal;
    hevc_vps_t* vps;
    hevc_sps_t* sps;
    hevc_pps_t* pps;
    hevc_aud_t* aud;
    hevc_slice_header_t* sh;
    hevc_slice_data_rbsp_t* slice_data;
    hevc_sps_t* sps_table[32];
    hevc_pps_t* pps_table[256];
} hevc_stream_t;

#define HEVCB_COMPAT_API __attribute__((visibility("default")))

HEVCB_COMPAT_API hevc_stream_t* hevc_new(void);
HEVCB_COMPAT_API void hevc_free(hevc_stream_t* h);
HEVCB_COMPAT_API int find_nal_unit(uint8_t* buf, int size, int* nal_start, int* nal_end);
HEVCB_COMPAT_API int nal_to_rbsp(const uint8_t* nal_buf, int* nal_size, uint8_t* rbsp_buf, int* rbsp_size);
HEVCB_COMPAT_API int rbsp_to_nal(const uint8_t* rbsp_buf, const int* rbsp_size, uint8_t* nal_buf, int* nal_size);
HEVCB_COMPAT_API int read_hevc_nal_unit(hevc_stream_t* h, uint8_t* buf, int size);
HEVCB_COMPAT_API int write_hevc_nal_unit(hevc_stream_t* h, uint8_t* buf, int size);
HEVCB_COMPAT_API int peek_hevc_nal_unit(hevc_stream_t* h, uint8_t* buf, int size);
/* hevc_stream.c:2343-3436: read_hevc_nal_unit + one "%ld.%d: <expr>: %d \n" line per syntax element on stdout (process.pl:90-113).
 * The walk runs on the device in its read_debug variant (hevcb.h: trace variant of the parse), the lines are formatted here. */
HEVCB_COMPAT_API int read_debug_hevc_nal_unit(hevc_stream_t* h, uint8_t* buf, int size);
/* h264_stream.c:117-126: hex dump, 16 bytes per line, to h264_dbgfile (stdout when NULL) */
HEVCB_COMPAT_API void debug_bytes(uint8_t* buf, int len);
HEVCB_COMPAT_API extern FILE* h264_dbgfile; /* h264_stream.c:33 */

#ifdef __cplusplus
}
#endif

#endif /* HEVCB_COMPAT_H */
