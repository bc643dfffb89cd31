// tests/hostsim/hostsim_scan.cpp -- TEST-ONLY CPU emulation of the scan+strip kernel's *logic*.
//
// Compiles the host/device header hevcb_scan_core.h with g++ and walks a buffer chunk by chunk in
// stream order with the same analyze / emit / carry / finalize functions the sm_100a kernels use.  It
// lets `pytest -m "not gpu"` pin those functions against the oracle where no GPU exists.  It is never
// linked into the product library and the product never calls it.
#include <cstdint>
#include <cstring>
#include <vector>
#include "hevcb_scan_core.h"

namespace {
struct Sink {
    int64_t *ns, *ne, *ro, *re, cap;
    int64_t first_empty;
    void open(int64_t k, int64_t start, int64_t off) { if (k < cap) { ns[k] = start; ro[k] = off; } }
    void close(int64_t k, int64_t end, int64_t rend, bool empty)
    {
        if (k < cap) { ne[k] = end; re[k] = rend; }
        if (empty && (first_empty < 0 || k < first_empty)) { first_empty = k; }
    }
};
inline uint32_t ld32(const std::vector<uint8_t>& v, int64_t off)
{
    uint32_t w;
    memcpy(&w, v.data() + off, 4);
    return w;
}
} // namespace

extern "C" int64_t hostsim_scan_strip(const uint8_t* buf, int64_t size, int64_t* nal_start, int64_t* nal_end,
                                      int64_t* rbsp_off, int64_t* rbsp_end, int64_t cap, uint8_t* rbsp_out,
                                      hevcb_scan_summary_core* summary, int use_fast_test)
{
    // padded image: 16 bytes of 0xFF in front (positions < 0 are non-zero), zeros behind (padding rule)
    const int64_t LEAD = 16;
    int64_t nchunks = (size + 15) / 16;
    std::vector<uint8_t> img((size_t)(LEAD + nchunks * 16 + 32), 0);
    memset(img.data(), 0xFF, LEAD);
    if (size > 0) { memcpy(img.data() + LEAD, buf, (size_t)size); }

    Sink sink{nal_start, nal_end, rbsp_off, rbsp_end, cap, -1};
    int64_t N = 0, K = 0;
    uint32_t kind = HEVCB_KIND_Z3, err = 0;
    for (int64_t c = 0; c < nchunks; c++) {
        int64_t g0 = c * 16;
        int64_t o = LEAD + g0;
        uint32_t wp = ld32(img, o - 4), w0 = ld32(img, o), w1 = ld32(img, o + 4), w2 = ld32(img, o + 8),
                 w3 = ld32(img, o + 12), wn = ld32(img, o + 16);
        hevcb_chunk_masks m;
        if (use_fast_test && !hevcb_maybe_zero_pair(wp, w0, w1, w2, w3, wn)) {
            // fast path of the kernel: no events, no removals, no errors
            int64_t rem = size - g0;
            m.ev = m.sc = m.scb = m.del = m.err = 0;
            m.valid = rem >= 16 ? 0xFFFFu : ((1u << (int)rem) - 1u);
        } else {
            m = hevcb_chunk_analyze(wp, w0, w1, w2, w3, wn, g0, size);
        }
        hevcb_chunk_emit(m, g0, N, K, kind, err, sink);
        uint32_t keep = m.valid & ~m.del;
        if (rbsp_out) {
            for (int j = 0; j < 16; j++) {
                if ((keep >> j) & 1u) { rbsp_out[K + hevcb_popc(keep & ((1u << j) - 1u))] = buf[g0 + j]; }
            }
        }
        uint32_t ck, ce;
        hevcb_chunk_summary(m, ck, ce);
        hevcb_carry_combine(kind, err, ck, ce);
        N += hevcb_popc(m.sc);
        K += hevcb_popc(keep);
    }
    auto fetch = [&](int64_t pos) -> uint32_t { return buf[pos]; };
    hevcb_scan_finalize(size, N, kind, err, K, sink.first_empty, fetch, nal_start, nal_end, rbsp_off, rbsp_end, cap, summary);
    return summary->n_nals;
}

// one shard of a byte-range partition: buf holds own + halo bytes; mirrors hevcb_scan_strip_shard_device
extern "C" int64_t hostsim_scan_strip_shard(const uint8_t* buf, int64_t own, int64_t halo, int is_first, int is_last, int64_t* nal_start,
                                            int64_t* nal_end, int64_t* rbsp_off, int64_t* rbsp_end, int64_t cap, uint8_t* rbsp_out,
                                            hevcb_shard_summary* summary)
{
    const int64_t LEAD = 16;
    const int64_t size = own + halo;
    const int64_t evl = is_last ? own - HEVCB_TAIL_ZONE : own;
    int64_t nchunks = (own + 15) / 16;
    std::vector<uint8_t> img((size_t)(LEAD + nchunks * 16 + 48), 0);
    memset(img.data(), 0xFF, LEAD);
    if (size > 0) { memcpy(img.data() + LEAD, buf, (size_t)size); }

    Sink sink{nal_start, nal_end, rbsp_off, rbsp_end, cap, -1};
    int64_t N = is_first ? 0 : 1, K = 0;
    uint32_t kind = is_first ? HEVCB_KIND_Z3 : HEVCB_KIND_SC3, err = 0;
    if (!is_first && cap > 0) { nal_start[0] = 0; rbsp_off[0] = 0; }
    for (int64_t c = 0; c < nchunks; c++) {
        int64_t g0 = c * 16;
        int64_t o = LEAD + g0;
        uint32_t wp = ld32(img, o - 4), w0 = ld32(img, o), w1 = ld32(img, o + 4), w2 = ld32(img, o + 8),
                 w3 = ld32(img, o + 12), wn = ld32(img, o + 16);
        hevcb_chunk_masks m = hevcb_chunk_analyze(wp, w0, w1, w2, w3, wn, g0, size, own, evl);
        hevcb_chunk_emit(m, g0, N, K, kind, err, sink);
        uint32_t keep = m.valid & ~m.del;
        if (rbsp_out) {
            for (int j = 0; j < 16; j++) {
                if ((keep >> j) & 1u) { rbsp_out[K + hevcb_popc(keep & ((1u << j) - 1u))] = buf[g0 + j]; }
            }
        }
        uint32_t ck, ce;
        hevcb_chunk_summary(m, ck, ce);
        hevcb_carry_combine(kind, err, ck, ce);
        N += hevcb_popc(m.sc);
        K += hevcb_popc(keep);
    }
    auto fetch = [&](int64_t pos) -> uint32_t { return buf[pos]; };
    hevcb_shard_finalize(own, N, kind, err, K, sink.first_empty, fetch, nal_start, nal_end, rbsp_off, rbsp_end, cap, is_first, is_last, summary);
    return N;
}
