/*
 * hevcb_analyze.c -- hevc_analyze (hevc_analyze.c:64-219 of the reference) on the BATCHED API: the whole file goes through one
 * hevcb_index_host call in its trace variant (scan + EPB strip + read_debug walk of every NAL on the GPU, include/hevcb.h), and
 * the host only formats the records into the reference's stdout grammar:
 *
 *     !! Found NAL at offset %lld (0x%04llX), size %lld (0x%04llX) \n        (verbose > 0; to the -o file when given)
 *     <hex dump of up to 16 bytes, starting 4 bytes before the END OF THE PREVIOUS NAL>       (same)
 *     %ld.%d: <expr>: %d \n    per syntax element, always to stdout
 *
 * Same options as the reference: -o file, -v level, -p (a no-op there too), -h.  For files up to the reference's 32 MiB window the
 * output is byte-identical to hevc_analyze's (tests/test_compat_gpu.py); beyond it the reference mis-reports the NAL that
 * straddles every window refill and shifts all later offsets, which this tool does not imitate.  The four bytes the
 * reference dumps in front of the FIRST NAL lie before its malloc'ed buffer (the upper half of the chunk size: zeros); they are
 * printed as 00 here.
 *
 *   gcc -O2 -Iinclude tools/hevcb_analyze.c -Lhevcbitstream_b200 -lhevcb200 -Wl,-rpath,$PWD/hevcbitstream_b200 -o hevcb_analyze
 *
 * (The reference's own hevc_analyze.c also builds unmodified against include/compat/ + libhevcb200_compat: one GPU call per NAL.)
 */
#include <errno.h>
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "hevcb.h"

static void usage(void)
{
    fprintf(stderr, "hevcb_analyze (hevc_analyze 0.2.0 output format, libhevcb200)\n");
    fprintf(stderr, "Analyze H.265 bitstreams in Annex B format\nUsage: \nhevcb_analyze [options] <input bitstream>\noptions:\n");
    fprintf(stderr, "\t-o output_file, receives the '!! Found NAL' lines and hex dumps\n\t-v verbose_level, print more info\n");
    fprintf(stderr, "\t-p accepted and ignored, as in the reference\n\t-h print this message and exit\n\n");
}

/* debug_bytes (h264_stream.c:117-126) over [from, from + len) of the file; positions before the file print as 00 */
static void dump_bytes(FILE* f, const uint8_t* buf, long long from, int len)
{
    for (int i = 0; i < len; i++) {
        const long long pos = from + i;
        fprintf(f, "%02X ", pos < 0 ? 0 : buf[pos]);
        if ((i + 1) % 16 == 0) { fputc('\n', f); }
    }
    fputc('\n', f);
}

int main(int argc, char** argv)
{
    static struct option long_options[] = {{"probe", no_argument, NULL, 'p'}, {"output", required_argument, NULL, 'o'}, {"help", no_argument, NULL, 'h'},
                                           {"verbose", required_argument, NULL, 'v'}, {NULL, 0, NULL, 0}};
    int verbose = 1, c;
    FILE* dbg = NULL;
    if (argc < 2) { usage(); return EXIT_FAILURE; }
    while ((c = getopt_long(argc, argv, "o:phv:", long_options, NULL)) != -1) {
        switch (c) {
            case 'o': if (!dbg) { dbg = fopen(optarg, "wt"); } break;
            case 'p': verbose = 0; break;
            case 'v': verbose = atoi(optarg); break;
            default: usage(); return 1;
        }
    }
    if (optind >= argc) { usage(); return EXIT_FAILURE; }
    FILE* in = fopen(argv[optind], "rb");
    if (!in) { fprintf(stderr, "!! Error: could not open file: %s \n", strerror(errno)); return EXIT_FAILURE; }
    if (!dbg) { dbg = stdout; }
    fseek(in, 0, SEEK_END);
    const long long size = ftell(in);
    fseek(in, 0, SEEK_SET);
    uint8_t* buf = (uint8_t*)calloc(1, (size_t)size + 64);
    if (size > 0 && fread(buf, 1, (size_t)size, in) != (size_t)size) { fprintf(stderr, "!! Error: read failed: %s \n", strerror(errno)); return EXIT_FAILURE; }
    fclose(in);

    hevcb_ctx* ctx = NULL;
    if (hevcb_create(0, &ctx) != HEVCB_OK) { fprintf(stderr, "!! libhevcb200: %s (there is no CPU fallback)\n", hevcb_last_error(NULL)); return EXIT_FAILURE; }
    hevcb_stream_index idx;
    int64_t sps_bytes = 0, pps_bytes = 0;
    hevcb_ps_context_bytes(&sps_bytes, &pps_bytes);
    hevcb_parse_chain chain; /* the parameter-set state after the stream: needed when the scan ends on a zero-length NAL (below) */
    memset(&chain, 0, sizeof(chain));
    chain.sps_out = calloc(1, (size_t)sps_bytes);
    chain.pps_out = calloc(1, (size_t)pps_bytes);
    int64_t cap = size / 64 + 1024, cap_pairs = 0;
    int rc = HEVCB_E_CAPACITY;
    for (int attempt = 0; attempt < 4 && rc == HEVCB_E_CAPACITY; attempt++) {
        memset(&idx, 0, sizeof(idx));
        if (cap_pairs == 0) { cap_pairs = 128 * cap + 65536; }
        idx.cap_nals = cap;
        idx.nal_start = (int64_t*)malloc((size_t)cap * 8); idx.nal_end = (int64_t*)malloc((size_t)cap * 8);
        idx.rbsp_off = (int64_t*)malloc((size_t)cap * 8); idx.rbsp_end = (int64_t*)malloc((size_t)cap * 8);
        idx.p.rc = (int32_t*)malloc((size_t)cap * 4); idx.p.nal_hdr = (int32_t*)malloc((size_t)cap * 4);
        idx.p.kind = (uint8_t*)malloc((size_t)cap); idx.p.ubflag = (uint8_t*)malloc((size_t)cap);
        idx.p.hdr_end = (int32_t*)malloc((size_t)cap * 4); idx.p.cols = (int32_t*)malloc((size_t)cap * 32);
        idx.p.pair_off = (int64_t*)malloc((size_t)(cap + 1) * 8);
        idx.p.pair_field = (uint32_t*)malloc((size_t)cap_pairs * 4); idx.p.pair_value = (int32_t*)malloc((size_t)cap_pairs * 4);
        idx.p.pair_pos = (uint32_t*)malloc((size_t)cap_pairs * 4); /* non-NULL: the trace (read_debug) variant of the parse */
        idx.p.cap_pairs = cap_pairs;
        rc = hevcb_index_host_chain(ctx, buf, size, &idx, &chain);
        if (rc == HEVCB_E_CAPACITY) { /* the summaries hold the true counts */
            if (idx.scan.n_nals > cap) { cap = idx.scan.n_nals + 8; cap_pairs = 0; } else { cap_pairs = idx.parse.n_pairs + 8; }
            free(idx.nal_start); free(idx.nal_end); free(idx.rbsp_off); free(idx.rbsp_end); free(idx.p.rc); free(idx.p.nal_hdr); free(idx.p.kind);
            free(idx.p.ubflag); free(idx.p.hdr_end); free(idx.p.cols); free(idx.p.pair_off); free(idx.p.pair_field); free(idx.p.pair_value); free(idx.p.pair_pos);
        }
    }
    if (rc != HEVCB_OK) { fprintf(stderr, "!! libhevcb200: %s\n", hevcb_last_error(ctx)); return EXIT_FAILURE; }

    const int64_t n = idx.scan.n_nals;
    if (n == 0) {
        fprintf(stderr, "!! Did not find any NALs between offset %lld (0x%04llX), size %lld (0x%04llX), discarding \n", 0ll, 0ll, size, size);
    }
    char name[160];
    long long prev_end = 0;
    for (int64_t k = 0; k < n; k++) {
        const long long start = idx.nal_start[k], len = idx.nal_end[k] - idx.nal_start[k];
        if (verbose > 0) {
            fprintf(dbg, "!! Found NAL at offset %lld (0x%04llX), size %lld (0x%04llX) \n", start, start, len, len);
            dump_bytes(dbg, buf, prev_end - 4, len + 4 >= 16 ? 16 : (int)(len + 4));
        }
        for (int64_t i = idx.p.pair_off[k]; i < idx.p.pair_off[k + 1]; i++) {
            const uint32_t code = idx.p.pair_field[i], pos = idx.p.pair_pos[i];
            if (!(code & HEVCB_TRACE_SPECIAL) && (code & HEVCB_TRACE_SILENT)) { continue; }
            if (code == (HEVCB_TRACE_SPECIAL | (uint32_t)HEVCB_TRACE_OPEN_LINE)) { printf("%ld.%d: ", (long)(pos >> 3), 8 - (int)(pos & 7u)); continue; }
            if (hevcb_trace_name(idx.p.kind[k], code, name, (int)sizeof(name)) < 0) { snprintf(name, sizeof(name), "?%08x", code); }
            printf("%ld.%d: %s: %d \n", (long)(pos >> 3), 8 - (int)(pos & 7u), name, idx.p.pair_value[i]);
        }
        prev_end = idx.nal_end[k];
    }
    if (n > 0 && idx.scan.last_rc == 0) {
        /* The call that ended the reference's loop returned 0 (no further start code, or a zero-length NAL): hevc_analyze then
         * dumps "the last NAL" with the offsets that call left behind, i.e. a NAL of size 0 (hevc_analyze.c:190-205).  Its
         * read_debug walk reads zero bits throughout (type 0, a slice header against the current SPS / PPS). */
        const long long start = idx.scan.last_start, len = idx.scan.last_end - idx.scan.last_start;
        if (verbose > 0) {
            fprintf(dbg, "!! Found NAL at offset %lld (0x%04llX), size %lld (0x%04llX) \n", start, start, len, len);
            dump_bytes(dbg, buf, prev_end - 4, len + 4 >= 16 ? 16 : (int)(len + 4));
        }
        int64_t off0 = 0, end0 = 0, poff[2];
        int32_t rc0, hdr0, he0, cols0[8];
        uint8_t kind0, ub0;
        hevcb_parse_buffers pb;
        hevcb_parse_summary ps;
        hevcb_parse_chain ch2;
        memset(&ch2, 0, sizeof(ch2));
        ch2.sps_in = chain.sps_out; ch2.pps_in = chain.pps_out;
        pb.rc = &rc0; pb.nal_hdr = &hdr0; pb.kind = &kind0; pb.ubflag = &ub0; pb.hdr_end = &he0; pb.cols = cols0; pb.pair_off = poff;
        pb.cap_pairs = 1 << 16;
        pb.pair_field = (uint32_t*)malloc((size_t)pb.cap_pairs * 4); pb.pair_value = (int32_t*)malloc((size_t)pb.cap_pairs * 4);
        pb.pair_pos = (uint32_t*)malloc((size_t)pb.cap_pairs * 4);
        if (hevcb_parse_rbsp_host(ctx, NULL, 0, &off0, &end0, 1, &pb, &ps, &ch2) != HEVCB_OK) { fprintf(stderr, "!! libhevcb200: %s\n", hevcb_last_error(ctx)); return EXIT_FAILURE; }
        for (int64_t i = poff[0]; i < poff[1]; i++) {
            const uint32_t code = pb.pair_field[i], pos = pb.pair_pos[i];
            if (!(code & HEVCB_TRACE_SPECIAL) && (code & HEVCB_TRACE_SILENT)) { continue; }
            if (code == (HEVCB_TRACE_SPECIAL | (uint32_t)HEVCB_TRACE_OPEN_LINE)) { printf("%ld.%d: ", (long)(pos >> 3), 8 - (int)(pos & 7u)); continue; }
            if (hevcb_trace_name(kind0, code, name, (int)sizeof(name)) < 0) { snprintf(name, sizeof(name), "?%08x", code); }
            printf("%ld.%d: %s: %d \n", (long)(pos >> 3), 8 - (int)(pos & 7u), name, pb.pair_value[i]);
        }
    }
    hevcb_destroy(ctx);
    if (dbg != stdout) { fclose(dbg); }
    return 0;
}
