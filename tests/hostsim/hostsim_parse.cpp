// tests/hostsim/hostsim_parse.cpp -- TEST-ONLY CPU emulation of the batched header parser's *logic*.
//
// Runs hevcb_syntax.h (the host/device walker the sm_100a parser kernels use) NAL by NAL in stream order with the
// same dependency rule the kernels implement ("a slice sees the most recent SPS / PPS NAL before it"), materialises
// every parsed NAL into zero-filled structs by scattering the (field, value) pairs, and returns digests that
// tests/ compare with the reference's.  Never linked into the product library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "hevcb_syntax.h"

namespace {
uint64_t hash_ints(uint64_t h, const int32_t* p, size_t n)
{
    const uint64_t M = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < n; i++) { h = h * M + (uint64_t)(uint32_t)p[i] + 1ull; }
    return h;
}
uint64_t hash_bytes(uint64_t h, const uint8_t* p, size_t n)
{
    const uint64_t M = 0x9E3779B97F4A7C15ull;
    for (size_t i = 0; i < n; i++) { h = h * M + (uint64_t)p[i] + 1ull; }
    return h;
}
// nal_to_rbsp restated for the test harness only (h264_nal.c:147-200)
int64_t strip(const uint8_t* nal, int64_t n, std::vector<uint8_t>& out, int64_t* consumed)
{
    out.clear();
    int count = 0;
    int64_t i;
    for (i = 0; i < n; i++) {
        if (count == 2 && nal[i] < 3) { return -1; }
        if (count == 2 && nal[i] == 3) {
            if (i < n - 1 && nal[i + 1] > 3) { return -1; }
            if (i == n - 1) { break; }
            i++;
            count = 0;
        }
        out.push_back(nal[i]);
        count = nal[i] == 0 ? count + 1 : 0;
    }
    *consumed = i;
    return (int64_t)out.size();
}
} // namespace

struct sim_record {
    int32_t rc, strip_rc, nal_unit_type, nal_layer_id, nal_temporal_id_plus1, slice_data_size;
    uint64_t state_hash, slice_data_hash;
};

static int64_t parse_all_impl(const uint8_t* buf, const int64_t* starts, const int64_t* ends, int64_t n, sim_record* rec,
                              int32_t* sh_dump, int64_t* n_pairs_total, uint32_t* flags_out, bool spec)
{
    // spec mode (HEVCB_PARSE_SPEC): the tables of every SPS / PPS NAL seen so far, entry 0 = the zeroed state (as the kernels keep them)
    std::vector<hevcb_sps_ctx> sps_tab(1);
    std::vector<hevcb_pps_ctx> pps_tab(1);
    memset(&sps_tab[0], 0, sizeof(hevcb_sps_ctx));
    memset(&pps_tab[0], 0, sizeof(hevcb_pps_ctx));
    std::vector<uint8_t> rbsp;
    std::vector<uint32_t> fld(1 << 16);
    std::vector<int32_t> val(1 << 16);
    hevcb_sps_ctx* sps = new hevcb_sps_ctx();
    hevcb_pps_ctx* pps = new hevcb_pps_ctx();
    memset(sps, 0, sizeof(*sps));
    memset(pps, 0, sizeof(*pps));
    hevc_vps_t* vps_s = new hevc_vps_t();
    hevc_sps_t* sps_s = new hevc_sps_t();
    hevc_pps_t* pps_s = new hevc_pps_t();
    hevc_slice_header_t* sh_s = new hevc_slice_header_t();
    int32_t nal_type = 0, nal_layer = 0, nal_tid = 0; // h->nal persists across failing strips
    int64_t ok = 0, pairs = 0;
    uint32_t allflags = 0;
    for (int64_t k = 0; k < n; k++) {
        sim_record& r = rec[k];
        memset(&r, 0, sizeof(r));
        int64_t consumed = 0;
        const int64_t size = ends[k] - starts[k];
        const int64_t rs = strip(buf + starts[k], size, rbsp, &consumed);
        r.strip_rc = (int32_t)rs;
        if (rs < 0) {
            r.rc = -1;
            r.nal_unit_type = nal_type; r.nal_layer_id = nal_layer; r.nal_temporal_id_plus1 = nal_tid;
            continue;
        }
        rbsp.resize(rbsp.size() + 16, 0);
        // count pass, then emit pass (as the kernels do)
        hevcb_nal_result res;
        hevcb_sink cs{nullptr, nullptr, 0};
        hevcb_sps_ctx* sps_new = new hevcb_sps_ctx();
        hevcb_pps_ctx pps_new;
        memset(sps_new, 0, sizeof(*sps_new));
        memset(&pps_new, 0, sizeof(pps_new));
        hevcb_ps_lookup lk{sps_tab.data(), pps_tab.data(), (int)sps_tab.size() - 1, (int)pps_tab.size() - 1};
        hevcb_parse_nal(rbsp.data(), rs, cs, sps, pps, sps_new, &pps_new, res, false, spec, spec ? &lk : nullptr);
        if (cs.n > fld.size()) { fld.resize(cs.n); val.resize(cs.n); }
        hevcb_sink es{fld.data(), val.data(), 0};
        memset(sps_new, 0, sizeof(*sps_new));
        memset(&pps_new, 0, sizeof(pps_new));
        hevcb_nal_result res2;
        hevcb_parse_nal(rbsp.data(), rs, es, sps, pps, sps_new, &pps_new, res2, false, spec, spec ? &lk : nullptr);
        if (es.n != cs.n || res2.ok != res.ok) { return -2; }
        pairs += es.n;
        allflags |= res.flags;
        nal_type = res.nal_unit_type; nal_layer = res.nal_layer_id; nal_tid = res.nal_temporal_id_plus1;
        r.nal_unit_type = nal_type; r.nal_layer_id = nal_layer; r.nal_temporal_id_plus1 = nal_tid;
        r.rc = res.ok ? (int32_t)consumed : -1;
        if (r.rc >= 0) { ok++; }
        int32_t* dst = nullptr;
        size_t words = 0;
        if (res.kind == HEVCB_KIND_SLICE) { dst = (int32_t*)sh_s; words = sizeof(*sh_s) / 4; }
        else if (res.kind == HEVCB_KIND_VPS) { dst = (int32_t*)vps_s; words = sizeof(*vps_s) / 4; }
        else if (res.kind == HEVCB_KIND_SPS) { dst = (int32_t*)sps_s; words = sizeof(*sps_s) / 4; *sps = *sps_new; sps_tab.push_back(*sps_new); }
        else if (res.kind == HEVCB_KIND_PPS) { dst = (int32_t*)pps_s; words = sizeof(*pps_s) / 4; *pps = pps_new; pps_tab.push_back(pps_new); }
        delete sps_new;
        if (dst) {
            memset(dst, 0, words * 4);
            for (uint32_t i = 0; i < es.n; i++) {
                if (fld[i] >= words) { return -3; }
                dst[fld[i]] = val[i];
            }
            r.state_hash = hash_ints(0, dst, words);
            if (res.kind == HEVCB_KIND_SLICE) {
                if (sh_dump) { memcpy(sh_dump + k * (int64_t)words, dst, words * 4); }
                const int64_t sd_off = (int64_t)res.hdr_end + 1;
                r.slice_data_size = (int32_t)(rs - sd_off);
                if (r.slice_data_size > 0) { r.slice_data_hash = hash_bytes(0, rbsp.data() + sd_off, (size_t)r.slice_data_size); }
            }
        }
    }
    *n_pairs_total = pairs;
    *flags_out = allflags;
    delete sps; delete pps; delete vps_s; delete sps_s; delete pps_s; delete sh_s;
    return ok;
}

extern "C" int64_t hostsim_parse_all(const uint8_t* buf, const int64_t* starts, const int64_t* ends, int64_t n, sim_record* rec,
                                     int32_t* sh_dump, int64_t* n_pairs_total, uint32_t* flags_out)
{
    return parse_all_impl(buf, starts, ends, n, rec, sh_dump, n_pairs_total, flags_out, false);
}
// the same walk in spec-correct mode (oracle: oracle/_ref/libhevcref_spec.so)
extern "C" int64_t hostsim_parse_all_spec(const uint8_t* buf, const int64_t* starts, const int64_t* ends, int64_t n, sim_record* rec,
                                          int32_t* sh_dump, int64_t* n_pairs_total, uint32_t* flags_out)
{
    return parse_all_impl(buf, starts, ends, n, rec, sh_dump, n_pairs_total, flags_out, true);
}

// ------------------------------------------------------------------------------------------------
// rewrite: parse -> edit -> write walker -> splice -> EPB insertion, NAL by NAL (mirrors SURVEY 3.4 / ref_rewrite_all)
// ------------------------------------------------------------------------------------------------
namespace {
void insert_epb(const std::vector<uint8_t>& rbsp, std::vector<uint8_t>& out) // rbsp_to_nal restated (h264_nal.c:92-132), test only
{
    int count = 0;
    for (size_t i = 0; i < rbsp.size(); i++) {
        if (count == 2 && rbsp[i] <= 3) { out.push_back(3); count = 0; }
        out.push_back(rbsp[i]);
        count = rbsp[i] == 0 ? count + 1 : 0;
    }
}
} // namespace

extern "C" int64_t hostsim_rewrite_all(const uint8_t* buf, int64_t size, const int64_t* starts, const int64_t* ends, int64_t n,
                                       const hevcb_edit_set* edits, uint8_t* out, int64_t out_cap, int64_t* out_starts, int64_t* out_ends,
                                       int64_t* n_rewritten)
{
    std::vector<uint8_t> rbsp, o, nr, w;
    std::vector<uint32_t> fld(1 << 16);
    std::vector<int32_t> val(1 << 16);
    hevcb_sps_ctx* sps = new hevcb_sps_ctx();
    hevcb_sps_ctx* sps_new = new hevcb_sps_ctx();
    hevcb_sps_ctx* sps_scr = new hevcb_sps_ctx();
    hevcb_pps_ctx* pps = new hevcb_pps_ctx();
    memset(sps, 0, sizeof(*sps));
    memset(pps, 0, sizeof(*pps));
    int64_t prev_end = 0, done_count = 0;
    for (int64_t k = 0; k < n; k++) {
        o.insert(o.end(), buf + prev_end, buf + starts[k]);
        prev_end = ends[k];
        out_starts[k] = (int64_t)o.size();
        const int64_t nsz = ends[k] - starts[k];
        int64_t consumed = 0;
        const int64_t rs = strip(buf + starts[k], nsz, rbsp, &consumed);
        bool done = false;
        if (rs >= 0) {
            rbsp.resize(rbsp.size() + 16, 0);
            hevcb_nal_result res;
            hevcb_sink cs{nullptr, nullptr, 0};
            hevcb_pps_ctx pps_new;
            memset(sps_new, 0, sizeof(*sps_new));
            memset(&pps_new, 0, sizeof(pps_new));
            hevcb_parse_nal(rbsp.data(), rs, cs, sps, pps, sps_new, &pps_new, res);
            if (cs.n > fld.size()) { fld.resize(cs.n); val.resize(cs.n); }
            hevcb_sink es{fld.data(), val.data(), 0};
            memset(sps_new, 0, sizeof(*sps_new));
            memset(&pps_new, 0, sizeof(pps_new));
            hevcb_parse_nal(rbsp.data(), rs, es, sps, pps, sps_new, &pps_new, res);
            if (res.kind == HEVCB_KIND_SPS) { *sps = *sps_new; }
            if (res.kind == HEVCB_KIND_PPS) { *pps = pps_new; }
            const int32_t hdr = res.nal_unit_type | (res.nal_layer_id << 8) | (res.nal_temporal_id_plus1 << 16);
            if (res.ok && res.kind != HEVCB_KIND_NONE && !(res.kind == HEVCB_KIND_SLICE && res.hdr_end > rs)) {
                // the writer's scratch: write_hevc_nal_unit gets size*3/4 bytes of RBSP room (hevc_stream.c:1266)
                const int64_t wcap_nal = (res.kind == HEVCB_KIND_SLICE) ? 16384 : nsz * 2 + 64;
                const int64_t wcap = wcap_nal * 3 / 4;
                hevcb_write_result wr;
                for (int pass = 0; pass < 2; pass++) { // count, then emit (as the kernels do)
                    hevcb_replay rp{fld.data(), val.data(), es.n, 0, res.kind, edits};
                    hevcb_bitwriter bw;
                    if (pass == 1) { w.assign((size_t)wr.bytes + 8, 0); }
                    bw.init(pass == 0 ? nullptr : w.data(), wcap);
                    memset(sps_scr, 0, sizeof(*sps_scr));
                    hevcb_write_nal(rp, bw, hdr, sps, pps, sps_scr, wr);
                    if (!wr.ok) { break; }
                }
                if (wr.ok && wr.bytes > 0) {
                    nr.clear();
                    if (res.kind == HEVCB_KIND_SLICE) {
                        nr.insert(nr.end(), w.begin(), w.begin() + wr.hdr_bytes);
                        nr.insert(nr.end(), rbsp.begin() + res.hdr_end, rbsp.begin() + rs);
                    } else {
                        nr.insert(nr.end(), w.begin(), w.begin() + wr.bytes);
                    }
                    insert_epb(nr, o);
                    done = true;
                    done_count++;
                }
            }
        }
        if (!done) { o.insert(o.end(), buf + starts[k], buf + ends[k]); }
        out_ends[k] = (int64_t)o.size();
    }
    if (size > prev_end) { o.insert(o.end(), buf + prev_end, buf + size); }
    delete sps; delete sps_new; delete sps_scr; delete pps;
    *n_rewritten = done_count;
    if ((int64_t)o.size() > out_cap) { return -1; }
    memcpy(out, o.data(), o.size());
    return (int64_t)o.size();
}

// ------------------------------------------------------------------------------------------------
// trace: the read_debug variant of the walker + hevcb_trace_name -> the text hevc_analyze prints
// ------------------------------------------------------------------------------------------------
#include <cstdio>
#include <string>

#include "hevcb_fields.cu" // host code only: field tables and hevcb_trace_name

extern "C" int64_t hostsim_trace_all(const uint8_t* buf, const int64_t* starts, const int64_t* ends, int64_t n, int verbose, char* out,
                                     int64_t out_cap)
{
    std::string text;
    std::vector<uint8_t> rbsp;
    std::vector<uint32_t> fld(1 << 16), pos(1 << 16);
    std::vector<int32_t> val(1 << 16);
    hevcb_sps_ctx* sps = new hevcb_sps_ctx();
    hevcb_sps_ctx* sps_new = new hevcb_sps_ctx();
    hevcb_pps_ctx* pps = new hevcb_pps_ctx();
    memset(sps, 0, sizeof(*sps));
    memset(pps, 0, sizeof(*pps));
    char line[512], name[160];
    for (int64_t k = 0; k < n; k++) {
        const int64_t size = ends[k] - starts[k];
        if (verbose > 0) {
            snprintf(line, sizeof(line), "!! Found NAL at offset %lld (0x%04llX), size %lld (0x%04llX) \n", (long long)starts[k], (long long)starts[k],
                     (long long)size, (long long)size);
            text += line;
        }
        int64_t consumed = 0;
        const int64_t rs = strip(buf + starts[k], size, rbsp, &consumed);
        if (rs < 0) { continue; }
        rbsp.resize(rbsp.size() + 16, 0);
        hevcb_nal_result res;
        hevcb_trace_sink cs{nullptr, nullptr, 0, nullptr};
        hevcb_pps_ctx pps_new;
        memset(sps_new, 0, sizeof(*sps_new));
        memset(&pps_new, 0, sizeof(pps_new));
        hevcb_parse_nal(rbsp.data(), rs, cs, sps, pps, sps_new, &pps_new, res);
        if (cs.n > fld.size()) { fld.resize(cs.n); val.resize(cs.n); pos.resize(cs.n); }
        hevcb_trace_sink es{fld.data(), val.data(), 0, pos.data()};
        memset(sps_new, 0, sizeof(*sps_new));
        memset(&pps_new, 0, sizeof(pps_new));
        hevcb_parse_nal(rbsp.data(), rs, es, sps, pps, sps_new, &pps_new, res);
        if (es.n != cs.n) { return -2; }
        if (res.kind == HEVCB_KIND_SPS) { *sps = *sps_new; }
        if (res.kind == HEVCB_KIND_PPS) { *pps = pps_new; }
        for (uint32_t i = 0; i < es.n; i++) {
            if (!(fld[i] & HEVCB_TRACE_SPECIAL) && (fld[i] & HEVCB_TRACE_SILENT)) { continue; }
            const int len = hevcb_trace_name(res.kind, fld[i], name, (int)sizeof(name));
            if (len < 0) { return -3; }
            if (fld[i] == (HEVCB_TRACE_SPECIAL | (uint32_t)HEVCB_TRACE_OPEN_LINE)) {
                snprintf(line, sizeof(line), "%ld.%d: ", (long)(pos[i] >> 3), 8 - (int)(pos[i] & 7u));
            } else {
                snprintf(line, sizeof(line), "%ld.%d: %s: %d \n", (long)(pos[i] >> 3), 8 - (int)(pos[i] & 7u), name, val[i]);
            }
            text += line;
        }
    }
    delete sps; delete sps_new; delete pps;
    if ((int64_t)text.size() > out_cap) { return -1; }
    memcpy(out, text.data(), text.size());
    return (int64_t)text.size();
}
