// hevcb_fields.cu -- name -> field index lookup over the struct layouts of include/hevcb_layout.h (host code only).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

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
