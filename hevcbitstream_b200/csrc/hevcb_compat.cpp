// hevcb_compat.cpp -- the reference's per-NAL API on top of the batched C ABI (include/hevcb_compat.h).  Host code only:
// buffers are handed to libhevcb200 (CUDA), results are unpacked into the reference's structs.
#include <map>
#include <mutex>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/hevcb.h"
#include "../../include/compat/bs.h"
#include "../../include/compat/h264_sei.h"
#include "../../include/hevcb_compat.h"

namespace {

hevcb_ctx* g_ctx = nullptr; // one context per process (device HEVCB_COMPAT_DEVICE, default 0)
// The reference's byte-layer functions are pure; here they share one context, its scratch buffers and the scan cache below.  Every
// exported function takes this lock, so concurrent callers are serialised instead of racing (two interleaved find_nal_unit loops
// still evict each other's cached scan: correct, but every call then scans again).
std::recursive_mutex g_mu;
#define HEVCB_COMPAT_LOCK std::lock_guard<std::recursive_mutex> compat_lock_(g_mu)

hevcb_ctx* context()
{
    if (!g_ctx) {
        const char* e = getenv("HEVCB_COMPAT_DEVICE");
        const int dev = e ? atoi(e) : 0;
        if (hevcb_create(dev, &g_ctx) != HEVCB_OK) {
            fprintf(stderr, "!! libhevcb200: %s (there is no CPU fallback)\n", hevcb_last_error(nullptr));
            g_ctx = nullptr;
        }
    }
    return g_ctx;
}

// result of the last whole-buffer scan, used to answer the calls of the canonical find_nal_unit loop
struct ScanCache {
    const uint8_t* base = nullptr;
    int64_t size = 0;
    std::vector<int64_t> ns, ne;
    hevcb_scan_summary sum;
    int64_t next = 0; // index of the NAL the next call of the loop asks for
    uint64_t probe = 0; // digest of the buffer's first and last bytes at scan time: a buffer rewritten in place is scanned again
} g_scan;

uint64_t probe_bytes(const uint8_t* buf, int64_t size)
{
    uint64_t h = 1469598103934665603ull ^ (uint64_t)size;
    const int64_t n = size < 64 ? size : 64;
    for (int64_t i = 0; i < n; i++) { h = (h ^ buf[i]) * 1099511628211ull; }
    for (int64_t i = size - n; i < size; i++) { h = (h ^ buf[i]) * 1099511628211ull; }
    return h;
}

bool scan_buffer(const uint8_t* buf, int64_t size)
{
    hevcb_ctx* ctx = context();
    if (!ctx) { return false; }
    int64_t cap = size / 64 + 1024;
    for (int attempt = 0; attempt < 2; attempt++) {
        g_scan.ns.assign((size_t)cap, 0);
        g_scan.ne.assign((size_t)cap, 0);
        const int rc = hevcb_scan_strip_host(ctx, buf, size, g_scan.ns.data(), g_scan.ne.data(), cap, nullptr, nullptr, nullptr, &g_scan.sum);
        if (rc == HEVCB_OK) {
            g_scan.base = buf;
            g_scan.size = size;
            g_scan.next = 0;
            g_scan.probe = probe_bytes(buf, size);
            return true;
        }
        if (rc != HEVCB_E_CAPACITY) { break; }
        cap = g_scan.sum.n_nals + 8; // the summary reports the true count
    }
    fprintf(stderr, "!! libhevcb200: %s\n", hevcb_last_error(ctx));
    g_scan.base = nullptr;
    return false;
}

// parameter-set state that read_hevc_nal_unit carries from call to call, per hevc_stream_t
struct StreamState {
    std::vector<uint8_t> sps, pps;
};
std::map<hevc_stream_t*, StreamState> g_streams;

// index buffers reused by read_hevc_nal_unit
struct IndexBuffers {
    std::vector<int64_t> a[4];
    std::vector<uint8_t> rbsp, kind, ubflag, stream;
    std::vector<int32_t> rc, hdr, hdr_end, cols, pval;
    std::vector<uint32_t> pfield, ppos;
    std::vector<int64_t> pair_off;
} g_ib;

} // namespace

extern "C" {

hevc_stream_t* hevc_new(void) // hevc_nal.c:34-55
{
    HEVCB_COMPAT_LOCK;
    if (!context()) { return nullptr; }
    hevc_stream_t* h = (hevc_stream_t*)calloc(1, sizeof(hevc_stream_t));
    h->nal = (hevc_nal_t*)calloc(1, sizeof(hevc_nal_t));
    for (int i = 0; i < 32; i++) { h->sps_table[i] = (hevc_sps_t*)calloc(1, sizeof(hevc_sps_t)); }
    for (int i = 0; i < 256; i++) { h->pps_table[i] = (hevc_pps_t*)calloc(1, sizeof(hevc_pps_t)); }
    h->vps = (hevc_vps_t*)calloc(1, sizeof(hevc_vps_t));
    h->sps = (hevc_sps_t*)calloc(1, sizeof(hevc_sps_t));
    h->pps = (hevc_pps_t*)calloc(1, sizeof(hevc_pps_t));
    h->aud = (hevc_aud_t*)calloc(1, sizeof(hevc_aud_t));
    h->sh = (hevc_slice_header_t*)calloc(1, sizeof(hevc_slice_header_t));
    h->slice_data = (hevc_slice_data_rbsp_t*)calloc(1, sizeof(hevc_slice_data_rbsp_t));
    int64_t sb = 0, pb = 0;
    hevcb_ps_context_bytes(&sb, &pb);
    StreamState& st = g_streams[h];
    st.sps.assign((size_t)sb, 0);
    st.pps.assign((size_t)pb, 0);
    return h;
}

void hevc_free(hevc_stream_t* h) // hevc_nal.c:64-90 (which leaks slice_data->rbsp_buf; freed here)
{
    if (!h) { return; }
    HEVCB_COMPAT_LOCK;
    g_streams.erase(h);
    free(h->nal);
    for (int i = 0; i < 32; i++) { free(h->sps_table[i]); }
    for (int i = 0; i < 256; i++) { free(h->pps_table[i]); }
    if (h->slice_data) { free(h->slice_data->rbsp_buf); }
    free(h->slice_data);
    free(h->sh);
    free(h->aud);
    free(h->pps);
    free(h->sps);
    free(h->vps);
    free(h);
}

int find_nal_unit(uint8_t* buf, int size, int* nal_start, int* nal_end) // h264_nal.c:38-76
{
    HEVCB_COMPAT_LOCK;
    *nal_start = 0;
    *nal_end = 0;
    if (size < 0) { return 0; }
    // does this call continue the loop over the cached buffer?  (p advanced by the previous call's nal_end)
    bool hit = false;
    if (g_scan.base && buf >= g_scan.base && buf + size == g_scan.base + g_scan.size) {
        const int64_t off = buf - g_scan.base;
        const int64_t k = g_scan.next;
        const int64_t expect = (k == 0) ? 0 : ((k - 1 < (int64_t)g_scan.ne.size() && k - 1 < g_scan.sum.n_terminated) ? g_scan.ne[(size_t)(k - 1)] : -1);
        hit = (off == expect) && g_scan.probe == probe_bytes(g_scan.base, g_scan.size);
    }
    if (!hit) {
        if (!scan_buffer(buf, size)) { return -1; }
    }
    const int64_t off = buf - g_scan.base;
    const int64_t k = g_scan.next;
    if (k < g_scan.sum.n_terminated) {
        *nal_start = (int)(g_scan.ns[(size_t)k] - off);
        *nal_end = (int)(g_scan.ne[(size_t)k] - off);
        g_scan.next = k + 1;
        return *nal_end - *nal_start;
    }
    // the call that ends the loop: 0 (no further start code, or a zero-length NAL) or -1 (unterminated last NAL)
    *nal_start = (int)(g_scan.sum.last_start - off);
    *nal_end = (int)(g_scan.sum.last_end - off);
    g_scan.base = nullptr; // a later call starts a new scan
    return g_scan.sum.last_rc;
}

int nal_to_rbsp(const uint8_t* nal_buf, int* nal_size, uint8_t* rbsp_buf, int* rbsp_size) // h264_nal.c:147-200
{
    HEVCB_COMPAT_LOCK;
    hevcb_ctx* ctx = context();
    if (!ctx || *nal_size < 0) { return -1; }
    const int64_t n = *nal_size;
    if (n == 0) { *rbsp_size = 0; return 0; } // nothing to convert (the loop of h264_nal.c:153 does not run)
    if (n <= 2 && nal_buf[0] == 0 && nal_buf[n - 1] == 0) {
        // {00} and {00 00}: the reference's pattern tests need three bytes (i + 2 < nal_size, h264_nal.c:156) and copy these through;
        // behind a start code they would not be a NAL at all (the scanner reads 00 00 01 00 [00] as a zero-length unit)
        if (n > *rbsp_size) { return -1; }
        memset(rbsp_buf, 0, (size_t)n);
        *rbsp_size = (int)n;
        return (int)n;
    }
    std::vector<uint8_t>& s = g_ib.stream;
    s.assign((size_t)n + 3 + 16, 0);
    s[2] = 1; // 00 00 01 in front: the NAL becomes the single unit of a stream
    if (n > 0) { memcpy(s.data() + 3, nal_buf, (size_t)n); }
    int64_t ns[4], ne[4], ro[4], re[4];
    g_ib.rbsp.assign((size_t)n + 3 + 32, 0);
    hevcb_scan_summary sum;
    const int rc = hevcb_scan_strip_host(ctx, s.data(), n + 3, ns, ne, 4, g_ib.rbsp.data(), ro, re, &sum);
    if (rc != HEVCB_OK) { return -1; } // more than a handful of units inside: start codes in the payload
    // 00 00 0{0,1} inside the NAL end it early for the scanner: nal_to_rbsp's -1 (h264_nal.c:156)
    if (sum.n_nals != 1 || ns[0] != 3 || ne[0] != n + 3 || re[0] < 0) { return -1; }
    const int64_t len = re[0] - ro[0];
    if (len > *rbsp_size) { return -1; } // output overflow (h264_nal.c:179)
    memcpy(rbsp_buf, g_ib.rbsp.data() + ro[0], (size_t)len);
    // a trailing 00 00 03 is not consumed (h264_nal.c:170-174)
    const bool drop = n >= 3 && nal_buf[n - 1] == 3 && nal_buf[n - 2] == 0 && nal_buf[n - 3] == 0;
    *nal_size = (int)(drop ? n - 1 : n);
    *rbsp_size = (int)len;
    return (int)len;
}

int rbsp_to_nal(const uint8_t* rbsp_buf, const int* rbsp_size, uint8_t* nal_buf, int* nal_size) // h264_nal.c:92-132
{
    HEVCB_COMPAT_LOCK;
    hevcb_ctx* ctx = context();
    if (!ctx || *rbsp_size < 0) { return -1; }
    const int64_t n = *rbsp_size;
    int64_t off = 0, end = n, out_off[2];
    std::vector<uint8_t>& o = g_ib.rbsp;
    o.assign((size_t)(n + n / 2 + 64), 0);
    hevcb_insert_summary sum;
    const int rc = hevcb_insert_host(ctx, rbsp_buf, n, &off, &end, 1, 0, o.data(), (int64_t)o.size(), out_off, &sum);
    if (rc != HEVCB_OK) {
        fprintf(stderr, "!! libhevcb200: %s\n", hevcb_last_error(ctx));
        return -1;
    }
    memcpy(nal_buf, o.data(), (size_t)sum.out_bytes); // like the reference, the caller provides the room (checks commented out, :101-107)
    *nal_size = (int)sum.out_bytes;
    return (int)sum.out_bytes;
}

int peek_hevc_nal_unit(hevc_stream_t* h, uint8_t* buf, int size) // hevc_nal.c:97-115
{
    const unsigned b0 = size > 0 ? buf[0] : 0u, b1 = size > 1 ? buf[1] : 0u; // reads past the end yield zero bits (bs.h)
    h->nal->nal_unit_type = (int)((b0 >> 1) & 0x3Fu);
    h->nal->nal_layer_id = (int)(((b0 & 1u) << 5) | (b1 >> 3));
    h->nal->nal_temporal_id_plus1 = (int)(b1 & 7u);
    if (h->nal->nal_unit_type <= 0 || h->nal->nal_unit_type > 40) { return -1; }
    return h->nal->nal_unit_type;
}

} // extern "C"

namespace {
// read_hevc_nal_unit (hevc_stream.c:155-241) and, with `debug`, read_debug_hevc_nal_unit (:2343-3436): the same call with the
// parse run in its trace variant and one line printed per record
int read_nal(hevc_stream_t* h, uint8_t* buf, int size, bool debug)
{
    HEVCB_COMPAT_LOCK;
    hevcb_ctx* ctx = context();
    std::map<hevc_stream_t*, StreamState>::iterator it = g_streams.find(h);
    if (!ctx || it == g_streams.end() || size < 0) { return -1; }
    StreamState& st = it->second;
    const int64_t n = size;
    std::vector<uint8_t>& s = g_ib.stream;
    s.assign((size_t)n + 3 + 16, 0);
    s[2] = 1;
    if (n > 0) { memcpy(s.data() + 3, buf, (size_t)n); }
    const int64_t cap = 4, cap_pairs = 1 << 20;
    for (int i = 0; i < 4; i++) { g_ib.a[i].assign((size_t)cap, 0); }
    g_ib.rbsp.assign((size_t)n + 3 + 32, 0);
    g_ib.rc.assign((size_t)cap, 0); g_ib.hdr.assign((size_t)cap, 0); g_ib.kind.assign((size_t)cap, 0); g_ib.ubflag.assign((size_t)cap, 0);
    g_ib.hdr_end.assign((size_t)cap, 0); g_ib.cols.assign((size_t)cap * 8, 0); g_ib.pair_off.assign((size_t)cap + 1, 0);
    g_ib.pfield.resize((size_t)cap_pairs); g_ib.pval.resize((size_t)cap_pairs);
    if (debug) { g_ib.ppos.resize((size_t)cap_pairs); }
    hevcb_stream_index idx;
    memset(&idx, 0, sizeof(idx));
    idx.cap_nals = cap;
    idx.nal_start = g_ib.a[0].data(); idx.nal_end = g_ib.a[1].data(); idx.rbsp_off = g_ib.a[2].data(); idx.rbsp_end = g_ib.a[3].data();
    idx.rbsp = g_ib.rbsp.data();
    idx.p.rc = g_ib.rc.data(); idx.p.nal_hdr = g_ib.hdr.data(); idx.p.kind = g_ib.kind.data(); idx.p.ubflag = g_ib.ubflag.data();
    idx.p.hdr_end = g_ib.hdr_end.data(); idx.p.cols = g_ib.cols.data(); idx.p.pair_off = g_ib.pair_off.data();
    idx.p.pair_field = g_ib.pfield.data(); idx.p.pair_value = g_ib.pval.data(); idx.p.cap_pairs = cap_pairs;
    idx.p.pair_pos = debug ? g_ib.ppos.data() : nullptr;
    std::vector<uint8_t> sps_out(st.sps.size()), pps_out(st.pps.size());
    hevcb_parse_chain chain;
    chain.sps_in = st.sps.data(); chain.pps_in = st.pps.data(); chain.sps_out = sps_out.data(); chain.pps_out = pps_out.data(); chain.buf_size = 0;
    if (n == 0) {
        // a zero-length NAL (hevc_analyze hands one over when its loop ends on return code 0, hevc_analyze.c:190-205): nal_to_rbsp
        // yields an empty RBSP and every read returns 0 bits; no start code can frame that for the scanner, so it is parsed as given
        idx.rbsp_off[0] = 0; idx.rbsp_end[0] = 0; idx.nal_start[0] = 0; idx.nal_end[0] = 0;
        if (hevcb_parse_rbsp_host(ctx, nullptr, 0, idx.rbsp_off, idx.rbsp_end, 1, &idx.p, &idx.parse, &chain) != HEVCB_OK) { return -1; }
        idx.scan.n_nals = 1;
    } else {
        const int rc = hevcb_index_host_chain(ctx, s.data(), n + 3, &idx, &chain);
        if (rc != HEVCB_OK) { return -1; }
        // nal_to_rbsp fails on the NAL (start codes inside, 00 00 02, ...): -1 before h->nal is touched (hevc_stream.c:165-167)
        if (idx.scan.n_nals != 1 || idx.nal_start[0] != 3 || idx.nal_end[0] != n + 3 || idx.rbsp_end[0] < 0) { return -1; }
    }
    st.sps.swap(sps_out);
    st.pps.swap(pps_out);
    if (debug) { // "%ld.%d: <expr>: %d \n" per element, always to stdout (the dbgfile redirect only exists in h264_stream.c)
        char name[160];
        for (int64_t i = idx.p.pair_off[0]; i < idx.p.pair_off[1]; i++) {
            const uint32_t code = idx.p.pair_field[i], pos = idx.p.pair_pos[i];
            if (!(code & HEVCB_TRACE_SPECIAL) && (code & HEVCB_TRACE_SILENT)) { continue; }
            if (code == (HEVCB_TRACE_SPECIAL | (uint32_t)HEVCB_TRACE_OPEN_LINE)) { printf("%ld.%d: ", (long)(pos >> 3), 8 - (int)(pos & 7u)); continue; }
            if (hevcb_trace_name(idx.p.kind[0], code, name, (int)sizeof(name)) < 0) { snprintf(name, sizeof(name), "?%08x", code); }
            printf("%ld.%d: %s: %d \n", (long)(pos >> 3), 8 - (int)(pos & 7u), name, idx.p.pair_value[i]);
        }
    }
    const int ret = hevcb_materialize(&idx, 0, h->nal, h->vps, h->sps, h->pps, h->sh);
    switch (idx.p.kind[0]) {
        case HEVCB_KIND_SPS: { // hevc_stream.c: memcpy(h->sps_table[sps->sps_seq_parameter_set_id], h->sps, ...)
            const int id = h->sps->sps_seq_parameter_set_id;
            if (id >= 0 && id < 32) { memcpy(h->sps_table[id], h->sps, sizeof(hevc_sps_t)); }
            break;
        }
        case HEVCB_KIND_PPS: {
            const int id = h->pps->pic_parameter_set_id;
            if (id >= 0 && id < 256) { memcpy(h->pps_table[id], h->pps, sizeof(hevc_pps_t)); }
            break;
        }
        case HEVCB_KIND_SLICE: { // slice data copy (hevc_stream.c:605-613): from one byte behind the aligned header end
            const int64_t from = idx.rbsp_off[0] + idx.p.hdr_end[0] + 1, len = idx.rbsp_end[0] - from;
            free(h->slice_data->rbsp_buf);
            h->slice_data->rbsp_buf = nullptr;
            h->slice_data->rbsp_size = (int)len;
            if (len > 0) {
                h->slice_data->rbsp_buf = (uint8_t*)malloc((size_t)len);
                memcpy(h->slice_data->rbsp_buf, idx.rbsp + from, (size_t)len);
            }
            break;
        }
        default: break;
    }
    return ret;
}

} // namespace

#pragma GCC visibility push(default) // everything below is API surface of the reference (include/compat/*.h declares it without attributes)
extern "C" {

FILE* h264_dbgfile = nullptr; // h264_stream.c:33

int read_hevc_nal_unit(hevc_stream_t* h, uint8_t* buf, int size) { return read_nal(h, buf, size, false); }       // hevc_stream.c:155
int read_debug_hevc_nal_unit(hevc_stream_t* h, uint8_t* buf, int size) { return read_nal(h, buf, size, true); } // hevc_stream.c:2343

// ---- host-side helpers of the reference's API surface (plain formatting / bs_t arithmetic, nothing to accelerate) ----------------

void debug_bytes(uint8_t* buf, int len) // h264_stream.c:117-126: every byte as "%02X ", a newline after every 16th and at the end
{
    FILE* f = h264_dbgfile ? h264_dbgfile : stdout;
    for (int i = 0; i < len; i++) {
        fprintf(f, "%02X ", buf[i]);
        if ((i + 1) % 16 == 0) { fputc('\n', f); }
    }
    fputc('\n', f);
}

int intlog2(int x) // h264_stream.c:42-52: ceil(log2(x)), 0 for x <= 0
{
    if (x <= 1) { return 0; }
    int bits = 0;
    for (unsigned v = (unsigned)x - 1u; v; v >>= 1) { bits++; }
    return bits;
}

int is_slice_type(int slice_type, int cmp_type) // h264_stream.c:54-60: types 5..9 alias 0..4
{
    return (slice_type >= 5 ? slice_type - 5 : slice_type) == (cmp_type >= 5 ? cmp_type - 5 : cmp_type);
}

int more_rbsp_data(bs_t* bs) // h264_stream.c:62-84: data follows unless the next 1 bit is the last 1 bit of the buffer
{
    if (bs_eof(bs)) { return 0; }
    if (bs_peek_u1(bs) == 0) { return 1; }
    bs_t t;
    bs_clone(&t, bs);
    bs_skip_u1(&t);
    while (!bs_eof(&t)) {
        if (bs_read_u1(&t)) { return 1; }
    }
    return 0;
}

int more_rbsp_trailing_data(bs_t* b) { return !bs_eof(b); } // h264_stream.c:86

int _read_ff_coded_number(bs_t* b) // h264_stream.c:88-98: bytes are summed up to and including the first one below 0xff
{
    int sum = 0;
    for (;;) {
        const int byte = (int)bs_read_u8(b);
        sum += byte;
        if (byte != 0xff) { return sum; }
    }
}

void _write_ff_coded_number(bs_t* b, int n) // h264_stream.c:100-115
{
    for (; n > 0xff; n -= 0xff) { bs_write_u8(b, 0xff); }
    bs_write_u8(b, (uint32_t)n);
}

void read_rbsp_trailing_bits(bs_t* b) // h264_stream.c:129-137
{
    bs_skip_u(b, 1);
    while (!bs_byte_aligned(b)) { bs_skip_u(b, 1); }
}

sei_t* sei_new() { return (sei_t*)calloc(1, sizeof(sei_t)); } // h264_sei.c:37-43
void sei_free(sei_t* s) // h264_sei.c:45-52
{
    if (!s) { return; }
    free(s->data);
    free(s);
}
void read_sei_end_bits(bs_t* b) // h264_sei.c:54-67
{
    if (!bs_byte_aligned(b)) {
        if (!bs_read_u1(b)) { fprintf(stderr, "WARNING: bit_equal_to_one is 0!!!!\n"); }
        while (!bs_byte_aligned(b)) {
            if (bs_read_u1(b)) { fprintf(stderr, "WARNING: bit_equal_to_zero is 1!!!!\n"); }
        }
    }
    read_rbsp_trailing_bits(b);
}
void read_sei_payload(sei_t* s, bs_t* b) // h264_sei.c:75-92: payloadSize raw bytes
{
    s->data = (uint8_t*)calloc(1, (size_t)(s->payloadSize > 0 ? s->payloadSize : 1));
    for (int i = 0; i < s->payloadSize; i++) { s->data[i] = (uint8_t)bs_read_u8(b); }
}
void write_sei_payload(sei_t* s, bs_t* b) // h264_sei.c:99-116
{
    for (int i = 0; i < s->payloadSize; i++) { bs_write_u8(b, s->data[i]); }
}
void read_debug_sei_payload(sei_t* s, bs_t* b) // h264_sei.c:123-140
{
    // As generated, only the position prefix sits inside the reference's loop; the read and the value line follow it once, with
    // i == payloadSize (a write one past the array there).  The visible output is reproduced, the byte is stored when it fits.
    s->data = (uint8_t*)calloc(1, (size_t)(s->payloadSize > 0 ? s->payloadSize : 0) + 1);
    for (int i = 0; i < s->payloadSize; i++) { printf("%ld.%d: ", (long)(b->p - b->start), b->bits_left); }
    const uint32_t v = bs_read_u8(b);
    s->data[s->payloadSize > 0 ? s->payloadSize : 0] = (uint8_t)v;
    printf("s->data[i]: %d \n", (int)(uint8_t)v);
}

int write_hevc_nal_unit(hevc_stream_t* h, uint8_t* buf, int size) // hevc_stream.c:1249-1335
{
    HEVCB_COMPAT_LOCK;
    hevcb_ctx* ctx = context();
    if (!ctx || !h || size < 0) { return -1; }
    int64_t n = -1;
    const int rc = hevcb_write_nal_host(ctx, h->nal->nal_unit_type, h->nal->nal_layer_id, h->nal->nal_temporal_id_plus1, h->vps, h->sps, h->pps, h->sh, buf,
                                        size, &n);
    if (rc != HEVCB_OK) {
        fprintf(stderr, "!! libhevcb200: %s\n", hevcb_last_error(ctx));
        return -1;
    }
    return (int)n;
}

} // extern "C"
#pragma GCC visibility pop
