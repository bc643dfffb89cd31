/*
 * hevcb_layout.h -- field lists of the reference's parsed-syntax structs (hevc_stream.h:41-569 in the reference).
 *
 * Every struct of the reference is a flat sequence of `int`s (scalars, 1-D / 2-D int arrays, nested structs).  The
 * lists below restate that sequence once, as X-macros; they generate
 *   - the C structs of the drop-in ABI (same member names, order and array bounds => same layout),
 *   - the "field index" (offset in ints) every batched parser result refers to,
 *   - the printable names used by the hevc_analyze-style dump.
 * Layout is verified at compile time against the sizes measured on the reference (VPS 428136,
 * SPS 76256, PPS 1968, slice header 4024, HRD 41652, PTL 6988, VUI 41808, st_ref_pic_set 792, scaling list 1264,
 * pred weight table 2056 bytes) and at test time by comparing materialised structs with the reference's.
 *
 * Macro arguments:  I(name) int scalar | A(name, n) int[n] | B(name, n, m) int[n][m] | S(type, name) nested struct |
 *                   T(type, name, n) nested struct array
 */
#ifndef HEVCB_LAYOUT_H
#define HEVCB_LAYOUT_H

#include <stddef.h>
#include <stdint.h>

/* array bounds (hevc_stream.h:21-35) */
#define HEVCB_MAX_SUBLAYERS 32
#define HEVCB_MAX_HRD_PARAM 10
#define HEVCB_MAX_CPB_CNT 32
#define HEVCB_MAX_PICS 32 /* negative / positive / L0 / L1 / short-term / long-term / tiles / entry points / qp offset list */

#define HEVCB_FIELDS_SUB_LAYER_HRD(I, A, B, S, T) \
    A(bit_rate_value_minus1, HEVCB_MAX_CPB_CNT)   \
    A(cpb_size_value_minus1, HEVCB_MAX_CPB_CNT)   \
    A(cpb_size_du_value_minus1, HEVCB_MAX_CPB_CNT) \
    A(bit_rate_du_value_minus1, HEVCB_MAX_CPB_CNT) \
    A(cbr_flag, HEVCB_MAX_CPB_CNT)

#define HEVCB_FIELDS_HRD(I, A, B, S, T)               \
    I(nal_hrd_parameters_present_flag)                \
    I(vcl_hrd_parameters_present_flag)                \
    I(sub_pic_hrd_params_present_flag)                \
    I(tick_divisor_minus2)                            \
    I(du_cpb_removal_delay_increment_length_minus1)   \
    I(sub_pic_cpb_params_in_pic_timing_sei_flag)      \
    I(dpb_output_delay_du_length_minus1)              \
    I(bit_rate_scale)                                 \
    I(cpb_size_scale)                                 \
    I(cpb_size_du_scale)                              \
    I(initial_cpb_removal_delay_length_minus1)        \
    I(au_cpb_removal_delay_length_minus1)             \
    I(dpb_output_delay_length_minus1)                 \
    A(fixed_pic_rate_general_flag, HEVCB_MAX_SUBLAYERS) \
    A(fixed_pic_rate_within_cvs_flag, HEVCB_MAX_SUBLAYERS) \
    A(elemental_duration_in_tc_minus1, HEVCB_MAX_SUBLAYERS) \
    A(low_delay_hrd_flag, HEVCB_MAX_SUBLAYERS)        \
    A(cpb_cnt_minus1, HEVCB_MAX_SUBLAYERS)            \
    T(hevc_sub_layer_hrd_t, sub_layer_hrd_nal, HEVCB_MAX_SUBLAYERS) \
    T(hevc_sub_layer_hrd_t, sub_layer_hrd_vcl, HEVCB_MAX_SUBLAYERS)

#define HEVCB_FIELDS_PTL(I, A, B, S, T)                 \
    I(general_profile_space)                            \
    I(general_tier_flag)                                \
    I(general_profile_idc)                              \
    A(general_profile_compatibility_flag, 32)           \
    I(general_progressive_source_flag)                  \
    I(general_interlaced_source_flag)                   \
    I(general_non_packed_constraint_flag)               \
    I(general_frame_only_constraint_flag)               \
    I(general_max_12bit_constraint_flag)                \
    I(general_max_10bit_constraint_flag)                \
    I(general_max_8bit_constraint_flag)                 \
    I(general_max_422chroma_constraint_flag)            \
    I(general_max_420chroma_constraint_flag)            \
    I(general_max_monochrome_constraint_flag)           \
    I(general_intra_constraint_flag)                    \
    I(general_one_picture_only_constraint_flag)         \
    I(general_lower_bit_rate_constraint_flag)           \
    I(general_max_14bit_constraint_flag)                \
    I(general_inbld_flag)                               \
    I(general_level_idc)                                \
    A(sub_layer_profile_present_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_level_present_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_profile_space, HEVCB_MAX_SUBLAYERS)     \
    A(sub_layer_tier_flag, HEVCB_MAX_SUBLAYERS)         \
    A(sub_layer_profile_idc, HEVCB_MAX_SUBLAYERS)       \
    B(sub_layer_profile_compatibility_flag, HEVCB_MAX_SUBLAYERS, 32) \
    A(sub_layer_progressive_source_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_interlaced_source_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_non_packed_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_frame_only_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_max_12bit_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_max_10bit_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_max_8bit_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_max_422chroma_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_max_420chroma_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_max_monochrome_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_intra_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_one_picture_only_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_lower_bit_rate_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_max_14bit_constraint_flag, HEVCB_MAX_SUBLAYERS) \
    A(sub_layer_inbld_flag, HEVCB_MAX_SUBLAYERS)        \
    A(sub_layer_level_idc, HEVCB_MAX_SUBLAYERS)

#define HEVCB_FIELDS_SCALING_LIST(I, A, B, S, T) \
    B(scaling_list_pred_mode_flag, 4, 6)          \
    B(scaling_list_pred_matrix_id_delta, 4, 6)    \
    B(scaling_list_dc_coef_minus8, 2, 6)          \
    B(scaling_list_delta_coef, 4, 64)

#define HEVCB_FIELDS_VPS(I, A, B, S, T)                       \
    I(vps_video_parameter_set_id)                             \
    I(vps_base_layer_internal_flag)                           \
    I(vps_base_layer_available_flag)                          \
    I(vps_max_layers_minus1)                                  \
    I(vps_max_sub_layers_minus1)                              \
    I(vps_temporal_id_nesting_flag)                           \
    S(hevc_profile_tier_level_t, ptl)                         \
    I(vps_sub_layer_ordering_info_present_flag)               \
    A(vps_max_dec_pic_buffering_minus1, HEVCB_MAX_SUBLAYERS)  \
    A(vps_max_num_reorder_pics, HEVCB_MAX_SUBLAYERS)          \
    A(vps_max_latency_increase_plus1, HEVCB_MAX_SUBLAYERS)    \
    I(vps_max_layer_id)                                       \
    I(vps_num_layer_sets_minus1)                              \
    B(layer_id_included_flag, HEVCB_MAX_SUBLAYERS, HEVCB_MAX_SUBLAYERS) \
    I(vps_timing_info_present_flag)                           \
    I(vps_num_units_in_tick)                                  \
    I(vps_time_scale)                                         \
    I(vps_poc_proportional_to_timing_flag)                    \
    I(vps_num_ticks_poc_diff_one_minus1)                      \
    I(vps_num_hrd_parameters)                                 \
    A(hrd_layer_set_idx, HEVCB_MAX_HRD_PARAM)                 \
    A(cprms_present_flag, HEVCB_MAX_HRD_PARAM)                \
    T(hevc_hrd_t, hrd, HEVCB_MAX_HRD_PARAM)                   \
    I(vps_extension_flag)                                     \
    I(vps_extension_data_flag)

#define HEVCB_FIELDS_ST_RPS(I, A, B, S, T)         \
    I(inter_ref_pic_set_prediction_flag)           \
    I(delta_idx_minus1)                            \
    I(delta_rps_sign)                              \
    I(abs_delta_rps_minus1)                        \
    A(used_by_curr_pic_flag, HEVCB_MAX_PICS)       \
    A(use_delta_flag, HEVCB_MAX_PICS)              \
    I(num_negative_pics)                           \
    I(num_positive_pics)                           \
    A(delta_poc_s0_minus1, HEVCB_MAX_PICS)         \
    A(used_by_curr_pic_s0_flag, HEVCB_MAX_PICS)    \
    A(delta_poc_s1_minus1, HEVCB_MAX_PICS)         \
    A(used_by_curr_pic_s1_flag, HEVCB_MAX_PICS)

#define HEVCB_FIELDS_VUI(I, A, B, S, T)        \
    I(aspect_ratio_info_present_flag)          \
    I(aspect_ratio_idc)                        \
    I(sar_width)                               \
    I(sar_height)                              \
    I(overscan_info_present_flag)              \
    I(overscan_appropriate_flag)               \
    I(video_signal_type_present_flag)          \
    I(video_format)                            \
    I(video_full_range_flag)                   \
    I(colour_description_present_flag)         \
    I(colour_primaries)                        \
    I(transfer_characteristics)                \
    I(matrix_coefficients)                     \
    I(chroma_loc_info_present_flag)            \
    I(chroma_sample_loc_type_top_field)        \
    I(chroma_sample_loc_type_bottom_field)     \
    I(neutral_chroma_indication_flag)          \
    I(field_seq_flag)                          \
    I(frame_field_info_present_flag)           \
    I(default_display_window_flag)             \
    I(def_disp_win_left_offset)                \
    I(def_disp_win_right_offset)               \
    I(def_disp_win_top_offset)                 \
    I(def_disp_win_bottom_offset)              \
    I(vui_timing_info_present_flag)            \
    I(vui_num_units_in_tick)                   \
    I(vui_time_scale)                          \
    I(vui_poc_proportional_to_timing_flag)     \
    I(vui_num_ticks_poc_diff_one_minus1)       \
    I(vui_hrd_parameters_present_flag)         \
    S(hevc_hrd_t, hrd)                         \
    I(bitstream_restriction_flag)              \
    I(tiles_fixed_structure_flag)              \
    I(motion_vectors_over_pic_boundaries_flag) \
    I(restricted_ref_pic_lists_flag)           \
    I(min_spatial_segmentation_idc)            \
    I(max_bytes_per_pic_denom)                 \
    I(max_bits_per_min_cu_denom)               \
    I(log2_max_mv_length_horizontal)           \
    I(log2_max_mv_length_vertical)

#define HEVCB_FIELDS_SPS_RANGE_EXT(I, A, B, S, T)  \
    I(transform_skip_rotation_enabled_flag)        \
    I(transform_skip_context_enabled_flag)         \
    I(implicit_rdpcm_enabled_flag)                 \
    I(explicit_rdpcm_enabled_flag)                 \
    I(extended_precision_processing_flag)          \
    I(intra_smoothing_disabled_flag)               \
    I(high_precision_offsets_enabled_flag)         \
    I(persistent_rice_adaptation_enabled_flag)     \
    I(cabac_bypass_alignment_enabled_flag)

#define HEVCB_FIELDS_SPS_SCC_EXT(I, A, B, S, T)           \
    I(sps_curr_pic_ref_enabled_flag)                      \
    I(palette_mode_enabled_flag)                          \
    I(palette_max_size)                                   \
    I(delta_palette_max_predictor_size)                   \
    I(sps_palette_predictor_initializer_present_flag)     \
    I(sps_num_palette_predictor_initializer_minus1)       \
    B(sps_palette_predictor_initializers, 3, HEVCB_MAX_PICS) \
    I(motion_vector_resolution_control_idc)               \
    I(intra_boundary_filtering_disabled_flag)

#define HEVCB_FIELDS_SPS(I, A, B, S, T)                        \
    I(sps_video_parameter_set_id)                              \
    I(sps_max_sub_layers_minus1)                               \
    I(sps_temporal_id_nesting_flag)                            \
    S(hevc_profile_tier_level_t, ptl)                          \
    I(sps_seq_parameter_set_id)                                \
    I(chroma_format_idc)                                       \
    I(separate_colour_plane_flag)                              \
    I(pic_width_in_luma_samples)                               \
    I(pic_height_in_luma_samples)                              \
    I(conformance_window_flag)                                 \
    I(conf_win_left_offset)                                    \
    I(conf_win_right_offset)                                   \
    I(conf_win_top_offset)                                     \
    I(conf_win_bottom_offset)                                  \
    I(bit_depth_luma_minus8)                                   \
    I(bit_depth_chroma_minus8)                                 \
    I(log2_max_pic_order_cnt_lsb_minus4)                       \
    I(sps_sub_layer_ordering_info_present_flag)                \
    A(sps_max_dec_pic_buffering_minus1, HEVCB_MAX_SUBLAYERS)   \
    A(sps_max_num_reorder_pics, HEVCB_MAX_SUBLAYERS)           \
    A(sps_max_latency_increase_plus1, HEVCB_MAX_SUBLAYERS)     \
    I(log2_min_luma_coding_block_size_minus3)                  \
    I(log2_diff_max_min_luma_coding_block_size)                \
    I(log2_min_luma_transform_block_size_minus2)               \
    I(log2_diff_max_min_luma_transform_block_size)             \
    I(max_transform_hierarchy_depth_inter)                     \
    I(max_transform_hierarchy_depth_intra)                     \
    I(scaling_list_enabled_flag)                               \
    I(sps_scaling_list_data_present_flag)                      \
    S(hevc_scaling_list_data_t, scaling_list_data)             \
    I(amp_enabled_flag)                                        \
    I(sample_adaptive_offset_enabled_flag)                     \
    I(pcm_enabled_flag)                                        \
    I(pcm_sample_bit_depth_luma_minus1)                        \
    I(pcm_sample_bit_depth_chroma_minus1)                      \
    I(log2_min_pcm_luma_coding_block_size_minus3)              \
    I(log2_diff_max_min_pcm_luma_coding_block_size)            \
    I(pcm_loop_filter_disabled_flag)                           \
    I(num_short_term_ref_pic_sets)                             \
    T(hevc_st_ref_pic_set_t, st_ref_pic_set, HEVCB_MAX_PICS)   \
    I(long_term_ref_pics_present_flag)                         \
    I(num_long_term_ref_pics_sps)                              \
    A(lt_ref_pic_poc_lsb_sps, HEVCB_MAX_PICS)                  \
    A(used_by_curr_pic_lt_sps_flag, HEVCB_MAX_PICS)            \
    I(sps_temporal_mvp_enabled_flag)                           \
    I(strong_intra_smoothing_enabled_flag)                     \
    I(vui_parameters_present_flag)                             \
    S(hevc_vui_t, vui)                                         \
    I(sps_extension_present_flag)                              \
    I(sps_range_extension_flag)                                \
    I(sps_multilayer_extension_flag)                           \
    I(sps_3d_extension_flag)                                   \
    I(sps_extension_5bits)                                     \
    S(hevc_sps_range_ext_t, sps_range_ext)

#define HEVCB_FIELDS_PPS_RANGE_EXT(I, A, B, S, T)       \
    I(log2_max_transform_skip_block_size_minus2)        \
    I(cross_component_prediction_enabled_flag)          \
    I(chroma_qp_offset_list_enabled_flag)               \
    I(diff_cu_chroma_qp_offset_depth)                   \
    I(chroma_qp_offset_list_len_minus1)                 \
    A(cb_qp_offset_list, HEVCB_MAX_PICS)                \
    A(cr_qp_offset_list, HEVCB_MAX_PICS)                \
    I(log2_sao_offset_scale_luma)                       \
    I(log2_sao_offset_scale_chroma)

#define HEVCB_FIELDS_PPS(I, A, B, S, T)                  \
    I(pic_parameter_set_id)                              \
    I(seq_parameter_set_id)                              \
    I(dependent_slice_segments_enabled_flag)             \
    I(output_flag_present_flag)                          \
    I(num_extra_slice_header_bits)                       \
    I(sign_data_hiding_enabled_flag)                     \
    I(cabac_init_present_flag)                           \
    I(num_ref_idx_l0_default_active_minus1)              \
    I(num_ref_idx_l1_default_active_minus1)              \
    I(init_qp_minus26)                                   \
    I(constrained_intra_pred_flag)                       \
    I(transform_skip_enabled_flag)                       \
    I(cu_qp_delta_enabled_flag)                          \
    I(diff_cu_qp_delta_depth)                            \
    I(pps_cb_qp_offset)                                  \
    I(pps_cr_qp_offset)                                  \
    I(pps_slice_chroma_qp_offsets_present_flag)          \
    I(weighted_pred_flag)                                \
    I(weighted_bipred_flag)                              \
    I(transquant_bypass_enabled_flag)                    \
    I(tiles_enabled_flag)                                \
    I(entropy_coding_sync_enabled_flag)                  \
    I(num_tile_columns_minus1)                           \
    I(num_tile_rows_minus1)                              \
    I(uniform_spacing_flag)                              \
    A(column_width_minus1, HEVCB_MAX_PICS)               \
    A(row_height_minus1, HEVCB_MAX_PICS)                 \
    I(loop_filter_across_tiles_enabled_flag)             \
    I(pps_loop_filter_across_slices_enabled_flag)        \
    I(deblocking_filter_control_present_flag)            \
    I(deblocking_filter_override_enabled_flag)           \
    I(pps_deblocking_filter_disabled_flag)               \
    I(pps_beta_offset_div2)                              \
    I(pps_tc_offset_div2)                                \
    I(pps_scaling_list_data_present_flag)                \
    S(hevc_scaling_list_data_t, scaling_list_data)       \
    I(lists_modification_present_flag)                   \
    I(log2_parallel_merge_level_minus2)                  \
    I(slice_segment_header_extension_present_flag)       \
    I(pps_extension_present_flag)                        \
    I(pps_range_extension_flag)                          \
    I(pps_multilayer_extension_flag)                     \
    I(pps_3d_extension_flag)                             \
    I(pps_extension_5bits)                               \
    S(hevc_pps_range_ext_t, pps_range_ext)

#define HEVCB_FIELDS_RPLM(I, A, B, S, T)          \
    I(ref_pic_list_modification_flag_l0)          \
    A(list_entry_l0, HEVCB_MAX_PICS)              \
    I(ref_pic_list_modification_flag_l1)          \
    A(list_entry_l1, HEVCB_MAX_PICS)

#define HEVCB_FIELDS_PWT(I, A, B, S, T)             \
    I(luma_log2_weight_denom)                       \
    I(delta_chroma_log2_weight_denom)               \
    A(luma_weight_l0_flag, HEVCB_MAX_PICS)          \
    A(chroma_weight_l0_flag, HEVCB_MAX_PICS)        \
    A(delta_luma_weight_l0, HEVCB_MAX_PICS)         \
    A(luma_offset_l0, HEVCB_MAX_PICS)               \
    B(delta_chroma_weight_l0, HEVCB_MAX_PICS, 2)    \
    B(delta_chroma_offset_l0, HEVCB_MAX_PICS, 2)    \
    A(luma_weight_l1_flag, HEVCB_MAX_PICS)          \
    A(chroma_weight_l1_flag, HEVCB_MAX_PICS)        \
    A(delta_luma_weight_l1, HEVCB_MAX_PICS)         \
    A(luma_offset_l1, HEVCB_MAX_PICS)               \
    B(delta_chroma_weight_l1, HEVCB_MAX_PICS, 2)    \
    B(delta_chroma_offset_l1, HEVCB_MAX_PICS, 2)

#define HEVCB_FIELDS_SLICE_HEADER(I, A, B, S, T)        \
    I(first_slice_segment_in_pic_flag)                  \
    I(no_output_of_prior_pics_flag)                     \
    I(pic_parameter_set_id)                             \
    I(dependent_slice_segment_flag)                     \
    I(slice_segment_address)                            \
    I(slice_type)                                       \
    I(pic_output_flag)                                  \
    I(colour_plane_id)                                  \
    I(slice_pic_order_cnt_lsb)                          \
    I(short_term_ref_pic_set_sps_flag)                  \
    S(hevc_st_ref_pic_set_t, st_ref_pic_set)            \
    I(short_term_ref_pic_set_idx)                       \
    I(num_long_term_sps)                                \
    I(num_long_term_pics)                               \
    A(lt_idx_sps, HEVCB_MAX_PICS)                       \
    A(poc_lsb_lt, HEVCB_MAX_PICS)                       \
    A(used_by_curr_pic_lt_flag, HEVCB_MAX_PICS)         \
    A(delta_poc_msb_present_flag, HEVCB_MAX_PICS)       \
    A(delta_poc_msb_cycle_lt, HEVCB_MAX_PICS)           \
    I(slice_temporal_mvp_enabled_flag)                  \
    I(slice_sao_luma_flag)                              \
    I(slice_sao_chroma_flag)                            \
    I(num_ref_idx_active_override_flag)                 \
    I(num_ref_idx_l0_active_minus1)                     \
    I(num_ref_idx_l1_active_minus1)                     \
    S(hevc_ref_pics_lists_mod_t, rpld)                  \
    I(mvd_l1_zero_flag)                                 \
    I(cabac_init_flag)                                  \
    I(collocated_from_l0_flag)                          \
    I(collocated_ref_idx)                               \
    S(hevc_pred_weight_table_t, pwt)                    \
    I(five_minus_max_num_merge_cand)                    \
    I(slice_qp_delta)                                   \
    I(slice_cb_qp_offset)                               \
    I(slice_cr_qp_offset)                               \
    I(cu_chroma_qp_offset_enabled_flag)                 \
    I(deblocking_filter_override_flag)                  \
    I(slice_deblocking_filter_disabled_flag)            \
    I(slice_beta_offset_div2)                           \
    I(slice_tc_offset_div2)                             \
    I(slice_loop_filter_across_slices_enabled_flag)     \
    I(num_entry_point_offsets)                          \
    I(offset_len_minus1)                                \
    A(entry_point_offset_minus1, HEVCB_MAX_PICS)        \
    I(slice_segment_header_extension_length)

#define HEVCB_FIELDS_NAL(I, A, B, S, T) \
    I(forbidden_zero_bit)               \
    I(nal_unit_type)                    \
    I(nal_layer_id)                     \
    I(nal_temporal_id_plus1)

/* struct generators */
#define HEVCB_MEMBER_I(name) int name;
#define HEVCB_MEMBER_A(name, n) int name[n];
#define HEVCB_MEMBER_B(name, n, m) int name[n][m];
#define HEVCB_MEMBER_S(type, name) type name;
#define HEVCB_MEMBER_T(type, name, n) type name[n];
#define HEVCB_DEFINE_STRUCT(type, FIELDS) \
    typedef struct { FIELDS(HEVCB_MEMBER_I, HEVCB_MEMBER_A, HEVCB_MEMBER_B, HEVCB_MEMBER_S, HEVCB_MEMBER_T) } type;

HEVCB_DEFINE_STRUCT(hevc_sub_layer_hrd_t, HEVCB_FIELDS_SUB_LAYER_HRD)
HEVCB_DEFINE_STRUCT(hevc_hrd_t, HEVCB_FIELDS_HRD)
HEVCB_DEFINE_STRUCT(hevc_profile_tier_level_t, HEVCB_FIELDS_PTL)
HEVCB_DEFINE_STRUCT(hevc_scaling_list_data_t, HEVCB_FIELDS_SCALING_LIST)
HEVCB_DEFINE_STRUCT(hevc_vps_t, HEVCB_FIELDS_VPS)
HEVCB_DEFINE_STRUCT(hevc_st_ref_pic_set_t, HEVCB_FIELDS_ST_RPS)
HEVCB_DEFINE_STRUCT(hevc_vui_t, HEVCB_FIELDS_VUI)
HEVCB_DEFINE_STRUCT(hevc_sps_range_ext_t, HEVCB_FIELDS_SPS_RANGE_EXT)
HEVCB_DEFINE_STRUCT(hevc_sps_scc_ext_t, HEVCB_FIELDS_SPS_SCC_EXT)
HEVCB_DEFINE_STRUCT(hevc_sps_t, HEVCB_FIELDS_SPS)
HEVCB_DEFINE_STRUCT(hevc_pps_range_ext_t, HEVCB_FIELDS_PPS_RANGE_EXT)
HEVCB_DEFINE_STRUCT(hevc_pps_t, HEVCB_FIELDS_PPS)
HEVCB_DEFINE_STRUCT(hevc_ref_pics_lists_mod_t, HEVCB_FIELDS_RPLM)
HEVCB_DEFINE_STRUCT(hevc_pred_weight_table_t, HEVCB_FIELDS_PWT)
HEVCB_DEFINE_STRUCT(hevc_slice_header_t, HEVCB_FIELDS_SLICE_HEADER)
HEVCB_DEFINE_STRUCT(hevc_nal_t, HEVCB_FIELDS_NAL)

/* layout checks against the reference (sizes in bytes on x86-64) */
#define HEVCB_SIZE_CHECK(type, bytes) typedef char hevcb_size_check_##type[(sizeof(type) == (bytes)) ? 1 : -1];
HEVCB_SIZE_CHECK(hevc_sub_layer_hrd_t, 640)
HEVCB_SIZE_CHECK(hevc_hrd_t, 41652)
HEVCB_SIZE_CHECK(hevc_profile_tier_level_t, 6988)
HEVCB_SIZE_CHECK(hevc_scaling_list_data_t, 1264)
HEVCB_SIZE_CHECK(hevc_vps_t, 428136)
HEVCB_SIZE_CHECK(hevc_st_ref_pic_set_t, 792)
HEVCB_SIZE_CHECK(hevc_vui_t, 41808)
HEVCB_SIZE_CHECK(hevc_sps_t, 76256)
HEVCB_SIZE_CHECK(hevc_pps_t, 1968)
HEVCB_SIZE_CHECK(hevc_pred_weight_table_t, 2056)
HEVCB_SIZE_CHECK(hevc_slice_header_t, 4024)
HEVCB_SIZE_CHECK(hevc_nal_t, 16)

/* field index = offset in ints inside the struct the NAL writes */
#define HEVCB_FIELD(type, member) ((uint32_t)(offsetof(type, member) / sizeof(int)))

/* which struct a parsed NAL writes (hevcb_parse results, `kind` column) */
#define HEVCB_KIND_NONE 0
#define HEVCB_KIND_VPS 1
#define HEVCB_KIND_SPS 2
#define HEVCB_KIND_PPS 3
#define HEVCB_KIND_SLICE 4
#define HEVCB_KIND_AUX 5 /* extension mode only: AUD / EOS / EOB / filler / SEI (hevcb.h, HEVCB_PARSE_AUX) */

#endif /* HEVCB_LAYOUT_H */
