// hevcb_fields.cu -- name -> field index lookup over the struct layouts of include/hevcb_layout.h, and the reverse: the text the
// reference's read_debug variant prints for an element (host code only).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <stdio.h>

#include "../../include/hevcb.h"
#include "../../include/hevcb_layout.h"

namespace {
struct FieldDesc {
    const char* name;
    uint32_t off;       // offset in ints inside the enclosing struct
    uint32_t n0, n1;    // array bounds (0 = not an array dimension)
    const FieldDesc* sub;
    uint32_t sub_count;
    uint32_t words;     // ints per element
};

#define FI(name) {#name, (uint32_t)(offsetof(CUR, name) / sizeof(int)), 0, 0, nullptr, 0, 1},
#define FA(name, n) {#name, (uint32_t)(offsetof(CUR, name) / sizeof(int)), (uint32_t)(n), 0, nullptr, 0, 1},
#define FB(name, n, m) {#name, (uint32_t)(offsetof(CUR, name) / sizeof(int)), (uint32_t)(n), (uint32_t)(m), nullptr, 0, 1},
#define FS(type, name) {#name, (uint32_t)(offsetof(CUR, name) / sizeof(int)), 0, 0, tbl_##type, cnt_##type, (uint32_t)(sizeof(type) / sizeof(int))},
#define FT(type, name, n) {#name, (uint32_t)(offsetof(CUR, name) / sizeof(int)), (uint32_t)(n), 0, tbl_##type, cnt_##type, (uint32_t)(sizeof(type) / sizeof(int))},
#define TABLE(type, FIELDS)                                                  \
    const FieldDesc tbl_##type[] = {FIELDS(FI, FA, FB, FS, FT)};             \
    const uint32_t cnt_##type = (uint32_t)(sizeof(tbl_##type) / sizeof(FieldDesc));

#define CUR hevc_sub_layer_hrd_t
TABLE(hevc_sub_layer_hrd_t, HEVCB_FIELDS_SUB_LAYER_HRD)
#undef CUR
#define CUR hevc_hrd_t
TABLE(hevc_hrd_t, HEVCB_FIELDS_HRD)
#undef CUR
#define CUR hevc_profile_tier_level_t
TABLE(hevc_profile_tier_level_t, HEVCB_FIELDS_PTL)
#undef CUR
#define CUR hevc_scaling_list_data_t
TABLE(hevc_scaling_list_data_t, HEVCB_FIELDS_SCALING_LIST)
#undef CUR
#define CUR hevc_vps_t
TABLE(hevc_vps_t, HEVCB_FIELDS_VPS)
#undef CUR
#define CUR hevc_st_ref_pic_set_t
TABLE(hevc_st_ref_pic_set_t, HEVCB_FIELDS_ST_RPS)
#undef CUR
#define CUR hevc_vui_t
TABLE(hevc_vui_t, HEVCB_FIELDS_VUI)
#undef CUR
#define CUR hevc_sps_range_ext_t
TABLE(hevc_sps_range_ext_t, HEVCB_FIELDS_SPS_RANGE_EXT)
#undef CUR
#define CUR hevc_sps_scc_ext_t
TABLE(hevc_sps_scc_ext_t, HEVCB_FIELDS_SPS_SCC_EXT)
#undef CUR
#define CUR hevc_sps_t
TABLE(hevc_sps_t, HEVCB_FIELDS_SPS)
#undef CUR
#define CUR hevc_pps_range_ext_t
TABLE(hevc_pps_range_ext_t, HEVCB_FIELDS_PPS_RANGE_EXT)
#undef CUR
#define CUR hevc_pps_t
TABLE(hevc_pps_t, HEVCB_FIELDS_PPS)
#undef CUR
#define CUR hevc_ref_pics_lists_mod_t
TABLE(hevc_ref_pics_lists_mod_t, HEVCB_FIELDS_RPLM)
#undef CUR
#define CUR hevc_pred_weight_table_t
TABLE(hevc_pred_weight_table_t, HEVCB_FIELDS_PWT)
#undef CUR
#define CUR hevc_slice_header_t
TABLE(hevc_slice_header_t, HEVCB_FIELDS_SLICE_HEADER)
#undef CUR

int64_t resolve(const FieldDesc* tbl, uint32_t cnt, const char* path)
{
    size_t len = 0;
    while (path[len] && path[len] != '.' && path[len] != '[') { len++; }
    for (uint32_t i = 0; i < cnt; i++) {
        const FieldDesc& f = tbl[i];
        if (strlen(f.name) != len || strncmp(f.name, path, len) != 0) { continue; }
        int64_t off = f.off;
        const char* p = path + len;
        const uint32_t dims[2] = {f.n0, f.n1};
        int nd = 0;
        int64_t idx[2] = {0, 0};
        while (*p == '[') {
            char* e = nullptr;
            const long v = strtol(p + 1, &e, 10);
            if (!e || *e != ']' || nd >= 2 || dims[nd] == 0 || v < 0 || (uint32_t)v >= dims[nd]) { return -1; }
            idx[nd++] = v;
            p = e + 1;
        }
        if (f.n1) { off += (idx[0] * f.n1 + idx[1]) * f.words; }
        else { off += idx[0] * f.words; }
        if (*p == '.') {
            if (!f.sub) { return -1; }
            const int64_t r = resolve(f.sub, f.sub_count, p + 1);
            return r < 0 ? -1 : off + r;
        }
        return *p ? -1 : off;
    }
    return -1;
}
} // namespace

extern "C" HEVCB_API int64_t hevcb_field_index(int kind, const char* path)
{
    if (!path) { return -1; }
    switch (kind) {
        case HEVCB_KIND_VPS: return resolve(tbl_hevc_vps_t, cnt_hevc_vps_t, path);
        case HEVCB_KIND_SPS: return resolve(tbl_hevc_sps_t, cnt_hevc_sps_t, path);
        case HEVCB_KIND_PPS: return resolve(tbl_hevc_pps_t, cnt_hevc_pps_t, path);
        case HEVCB_KIND_SLICE: return resolve(tbl_hevc_slice_header_t, cnt_hevc_slice_header_t, path);
        default: return -1;
    }
}

// ---- field index -> printed text -------------------------------------------------------------------------------------------
// read_debug_hevc_* prints the C lvalue text of the template (process.pl:90-113): "<struct pointer>-><member>"
// followed, for array members, by the index written SYMBOLICALLY the way the template spells it.  The spelling is regular per
// struct (st->style) with the handful of exceptions listed in kSuffix.
namespace {
struct TypeStyle {
    const FieldDesc* tbl;
    const char* prefix;
    const char* idx1; // suffix of a 1-D array member
    const char* idx2; // suffix of a 2-D array member
};
const TypeStyle kStyles[] = {
    {tbl_hevc_vps_t, "vps->", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_sps_t, "sps->", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_pps_t, "pps->", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_slice_header_t, "sh->", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_profile_tier_level_t, "ptl->", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_hrd_t, "hrd->", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_sub_layer_hrd_t, "sub_layer_hrd->", "[i]", "[i][j]"},
    {tbl_hevc_scaling_list_data_t, "sld->", "[ sizeId ]", "[ sizeId ][ matrixId ]"},
    {tbl_hevc_st_ref_pic_set_t, "st_ref_pic_set->", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_vui_t, "vui->", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_sps_range_ext_t, "sps_range_ext->", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_pps_range_ext_t, "pps_range_ext->", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_ref_pics_lists_mod_t, "sh->rpld.", "[ i ]", "[ i ][ j ]"},
    {tbl_hevc_pred_weight_table_t, "pwt->", "[i]", "[i][j]"},
};
struct SuffixException {
    const FieldDesc* tbl;
    const char* member;
    const char* suffix;
};
const SuffixException kSuffix[] = {
    {tbl_hevc_sps_t, "sps_max_dec_pic_buffering_minus1", " [ i ]"},
    {tbl_hevc_sps_t, "sps_max_num_reorder_pics", " [ i ]"},
    {tbl_hevc_sps_t, "sps_max_latency_increase_plus1", " [ i ]"},
    {tbl_hevc_st_ref_pic_set_t, "used_by_curr_pic_flag", "[ j ]"},
    {tbl_hevc_st_ref_pic_set_t, "use_delta_flag", "[ j ]"},
    {tbl_hevc_scaling_list_data_t, "scaling_list_dc_coef_minus8", "[ sizeId - 2 ][ matrixId ]"},
};
const char* const kSpecialNames[] = {
    nullptr, "forbidden_zero_bit", "nal->nal_unit_type", "nal->nal_layer_id", "nal->nal_temporal_id_plus1", "vps_reserved_0xffff_16bits",
    "general_reserved_zero_34bits", "general_reserved_zero_43bits", "general_reserved_zero_bit", "reserved_zero_xxbits",
    "sub_layer_reserved_zero_34bits", "sub_layer_reserved_zero_43bits", "sub_layer_reserved_zero_bit", "rbsp_stop_one_bit",
    "rbsp_alignment_zero_bit", "alignment_bit_equal_to_one", "alignment_bit_equal_to_zero", "slice_reserved_flag",
    "slice_segment_header_extension_data_byte", "" /* HEVCB_TRACE_OPEN_LINE */, "ff_byte",
};

// innermost member that holds field index f of the struct described by (tbl, cnt)
bool locate(const FieldDesc* tbl, uint32_t cnt, uint32_t f, const FieldDesc** owner_tbl, const FieldDesc** member)
{
    for (uint32_t i = 0; i < cnt; i++) {
        const FieldDesc& m = tbl[i];
        const uint32_t elems = (m.n0 ? m.n0 : 1u) * (m.n1 ? m.n1 : 1u);
        if (f < m.off || f >= m.off + elems * m.words) { continue; }
        if (m.sub) { return locate(m.sub, m.sub_count, (f - m.off) % m.words, owner_tbl, member); }
        *owner_tbl = tbl;
        *member = &m;
        return true;
    }
    return false;
}
} // namespace

extern "C" HEVCB_API int hevcb_trace_name(int kind, uint32_t code, char* out, int cap)
{
    if (!out || cap <= 0) { return -1; }
    out[0] = 0;
    if (code & HEVCB_TRACE_SPECIAL) {
        const uint32_t id = code & 0xFFFFu;
        if (id == 0 || id >= sizeof(kSpecialNames) / sizeof(kSpecialNames[0])) { return -1; }
        return snprintf(out, (size_t)cap, "%s", kSpecialNames[id]);
    }
    if (code & HEVCB_TRACE_SILENT) { return 0; }
    const FieldDesc* tbl = nullptr;
    uint32_t cnt = 0;
    switch (kind) {
        case HEVCB_KIND_VPS: tbl = tbl_hevc_vps_t; cnt = cnt_hevc_vps_t; break;
        case HEVCB_KIND_SPS: tbl = tbl_hevc_sps_t; cnt = cnt_hevc_sps_t; break;
        case HEVCB_KIND_PPS: tbl = tbl_hevc_pps_t; cnt = cnt_hevc_pps_t; break;
        case HEVCB_KIND_SLICE: tbl = tbl_hevc_slice_header_t; cnt = cnt_hevc_slice_header_t; break;
        case HEVCB_KIND_AUX: // extension mode: only the access unit delimiter has a printed struct member (hevc_stream.c:2707)
            return code == HEVCB_AUX_AUD_PIC_TYPE ? snprintf(out, (size_t)cap, "h->aud->primary_pic_type") : -1;
        default: return -1;
    }
    const FieldDesc *owner = nullptr, *m = nullptr;
    if (!locate(tbl, cnt, code, &owner, &m)) { return -1; }
    const TypeStyle* st = nullptr;
    for (const TypeStyle& s : kStyles) { if (s.tbl == owner) { st = &s; break; } }
    if (!st) { return -1; }
    const char* suffix = m->n1 ? st->idx2 : (m->n0 ? st->idx1 : "");
    for (const SuffixException& e : kSuffix) { if (e.tbl == owner && strcmp(e.member, m->name) == 0) { suffix = e.suffix; break; } }
    return snprintf(out, (size_t)cap, "%s%s%s", st->prefix, m->name, suffix);
}
