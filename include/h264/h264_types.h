/*
 * h264/h264_types.h — enums and plain-data structures of the libh264 API.
 *
 * Binary- and source-compatible with the reference's include/h264/h264_types.h
 * (field names, order and types are the ABI: Parrot-Developers/libh264
 * include/h264/h264_types.h:31-764; sizes checked against SURVEY.md Appendix B by
 * tests/test_host_abi.py).  Clause numbers refer to Rec. ITU-T H.264.
 */
#ifndef H264B200_TYPES_H
#define H264B200_TYPES_H

/* 7.4.1 nal_unit_type */
enum h264_nalu_type {
	H264_NALU_TYPE_UNKNOWN = 0, H264_NALU_TYPE_SLICE = 1, H264_NALU_TYPE_SLICE_DPA = 2,
	H264_NALU_TYPE_SLICE_DPB = 3, H264_NALU_TYPE_SLICE_DPC = 4, H264_NALU_TYPE_SLICE_IDR = 5,
	H264_NALU_TYPE_SEI = 6, H264_NALU_TYPE_SPS = 7, H264_NALU_TYPE_PPS = 8,
	H264_NALU_TYPE_AUD = 9, H264_NALU_TYPE_END_OF_SEQ = 10, H264_NALU_TYPE_END_OF_STREAM = 11,
	H264_NALU_TYPE_FILLER = 12,
};

/* 7.4.3 slice_type modulo 5; UNKNOWN for non-VCL NAL units */
enum h264_slice_type {
	H264_SLICE_TYPE_UNKNOWN = -1, H264_SLICE_TYPE_P = 0, H264_SLICE_TYPE_B = 1,
	H264_SLICE_TYPE_I = 2, H264_SLICE_TYPE_SP = 3, H264_SLICE_TYPE_SI = 4,
};
#define H264_SLICE_TYPE(_val) ((_val) % 5)

/* 7.4.5 macroblock types (values are this API's, not the standard's) */
enum h264_mb_type {
	H264_MB_TYPE_UNKNOWN = 0,
	H264_MB_TYPE_I_NxN, H264_MB_TYPE_I_16x16, H264_MB_TYPE_I_PCM, H264_MB_TYPE_SI,
	H264_MB_TYPE_P_16x16, H264_MB_TYPE_P_16x8, H264_MB_TYPE_P_8x16, H264_MB_TYPE_P_8x8,
	H264_MB_TYPE_P_8x8ref0, H264_MB_TYPE_P_SKIP,
	H264_MB_TYPE_B_Direct_16x16, H264_MB_TYPE_B_16x16, H264_MB_TYPE_B_16x8, H264_MB_TYPE_B_8x16,
	H264_MB_TYPE_B_8x8, H264_MB_TYPE_B_SKIP,
};

/* A.2 profile_idc */
enum h264_profile {
	H264_PROFILE_CAVLC_444 = 44, H264_PROFILE_BASELINE = 66, H264_PROFILE_MAIN = 77,
	H264_PROFILE_EXTENDED = 88, H264_PROFILE_HIGH = 100, H264_PROFILE_HIGH_10 = 110,
	H264_PROFILE_HIGH_422 = 122, H264_PROFILE_HIGH_444 = 244,
};

/* 6.2 chroma_format_idc */
enum h264_color_format {
	H264_COLOR_FORMAT_MONO = 0, H264_COLOR_FORMAT_YUV420 = 1, H264_COLOR_FORMAT_YUV422 = 2,
	H264_COLOR_FORMAT_YUV444 = 3,
};

/* E.2.1 aspect_ratio_idc (17..254 reserved) */
enum h264_aspect_ratio {
	H264_ASPECT_RATIO_UNSPECIFIED = 0, H264_ASPECT_RATIO_1_1, H264_ASPECT_RATIO_12_11,
	H264_ASPECT_RATIO_10_11, H264_ASPECT_RATIO_16_11, H264_ASPECT_RATIO_40_33,
	H264_ASPECT_RATIO_24_11, H264_ASPECT_RATIO_20_11, H264_ASPECT_RATIO_32_11,
	H264_ASPECT_RATIO_80_33, H264_ASPECT_RATIO_18_11, H264_ASPECT_RATIO_15_11,
	H264_ASPECT_RATIO_64_33, H264_ASPECT_RATIO_160_99, H264_ASPECT_RATIO_4_3,
	H264_ASPECT_RATIO_3_2, H264_ASPECT_RATIO_2_1,
	H264_ASPECT_RATIO_EXTENDED_SAR = 255,
};

/* D.1 payloadType */
enum h264_sei_type {
	H264_SEI_TYPE_BUFFERING_PERIOD = 0, H264_SEI_TYPE_PIC_TIMING, H264_SEI_TYPE_PAN_SCAN_RECT,
	H264_SEI_TYPE_FILLER_PAYLOAD, H264_SEI_TYPE_USER_DATA_REGISTERED,
	H264_SEI_TYPE_USER_DATA_UNREGISTERED, H264_SEI_TYPE_RECOVERY_POINT,
	H264_SEI_TYPE_DEC_REF_PIC_MARKING_REPETITION, H264_SEI_TYPE_SPARE_PIC,
	H264_SEI_TYPE_SCENE_INFO, H264_SEI_TYPE_SUB_SEQ_INFO,
	H264_SEI_TYPE_SUB_SEQ_LAYER_CHARACTERISTICS, H264_SEI_TYPE_SUB_SEQ_CHARACTERISTICS,
	H264_SEI_TYPE_FULL_FRAME_FREEZE, H264_SEI_TYPE_FULL_FRAME_FREEZE_RELEASE,
	H264_SEI_TYPE_FULL_FRAME_SNAPSHOT, H264_SEI_TYPE_PROGRESSIVE_REFINEMENT_SEGMENT_START,
	H264_SEI_TYPE_PROGRESSIVE_REFINEMENT_SEGMENT_END,
	H264_SEI_TYPE_MOTION_CONSTRAINED_SLICE_GROUP_SET, H264_SEI_TYPE_FILM_GRAIN_CHARACTERISTICS,
	H264_SEI_TYPE_DEBLOCKING_FILTER_DISPLAY_PREFERENCE, H264_SEI_TYPE_STEREO_VIDEO_INFO,
	H264_SEI_TYPE_POST_FILTER_HINT, H264_SEI_TYPE_TONE_MAPPING_INFO,
	H264_SEI_TYPE_SCALABILITY_INFO, H264_SEI_TYPE_SUB_PIC_SCALABLE_LAYER,
	H264_SEI_TYPE_NON_REQUIRED_LAYER_REP, H264_SEI_TYPE_PRIORITY_LAYER_INFO,
	H264_SEI_TYPE_LAYERS_NOT_PRESENT, H264_SEI_TYPE_LAYER_DEPENDENCY_CHANGE,
	H264_SEI_TYPE_SCALABLE_NESTING, H264_SEI_TYPE_BASE_LAYER_TEMPORAL_HRD,
	H264_SEI_TYPE_QUALITY_LAYER_INTEGRITY_CHECK, H264_SEI_TYPE_REDUNDANT_PIC_PROPERTY,
	H264_SEI_TYPE_TL0_DEP_REP_INDEX, H264_SEI_TYPE_TL_SWITCHING_POINT,
	H264_SEI_TYPE_PARALLEL_DECODING_INFO, H264_SEI_TYPE_MVC_SCALABLE_NESTING,
	H264_SEI_TYPE_VIEW_SCALABILITY_INFO, H264_SEI_TYPE_MULTIVIEW_SCENE_INFO,
	H264_SEI_TYPE_MULTIVIEW_ACQUISITION_INFO, H264_SEI_TYPE_NON_REQUIRED_VIEW_COMPONENT,
	H264_SEI_TYPE_VIEW_DEPENDENCY_CHANGE, H264_SEI_TYPE_OPERATION_POINTS_NOT_PRESENT,
	H264_SEI_TYPE_BASE_VIEW_TEMPORAL_HRD, H264_SEI_TYPE_FRAME_PACKING_ARRANGEMENT,
	H264_SEI_TYPE_MULTIVIEW_VIEW_POSITION, H264_SEI_TYPE_DISPLAY_ORIENTATION,
	H264_SEI_TYPE_MVCD_SCALABLE_NESTING, H264_SEI_TYPE_MVCD_VIEW_SCALABILITY_INFO,
	H264_SEI_TYPE_DEPTH_REPRESENTATION_INFO,
	H264_SEI_TYPE_THREE_DIMENSIONAL_REFERENCE_DISPLAYS_INFO, H264_SEI_TYPE_DEPTH_TIMING,
	H264_SEI_TYPE_DEPTH_SAMPLING_INFO,
	H264_SEI_TYPE_CONSTRAINED_DEPTH_PARAMETER_SET_IDENTIFIER = 54,
};

/* 7.3.1 */
struct h264_nalu_header {
	uint32_t forbidden_zero_bit, nal_ref_idc, nal_unit_type;
};

/* 7.3.2.1.1.1 scaling lists of an SPS or PPS; the _optimized_* members record
 * how the list was terminated so that a re-write is byte identical */
struct h264_scaling_matrix {
	int scaling_list_present_flag[12];
	int32_t scaling_list_4x4[6][16];
	int32_t scaling_list_8x8[6][64];
	int use_default_4x4[6], use_default_8x8[6];
	int _optimized_4x4[6], _optimized_8x8[6];
};

/* E.1.2 */
struct h264_hrd {
	uint32_t cpb_cnt_minus1, bit_rate_scale, cpb_size_scale;
	struct {
		uint32_t bit_rate_value_minus1, cpb_size_value_minus1;
		int cbr_flag;
	} cpb[32];
	uint32_t initial_cpb_removal_delay_length_minus1, cpb_removal_delay_length_minus1;
	uint32_t dpb_output_delay_length_minus1, time_offset_length;
};

/* E.1.1 */
struct h264_vui {
	int aspect_ratio_info_present_flag;
	uint32_t aspect_ratio_idc, sar_width, sar_height;
	int overscan_info_present_flag, overscan_appropriate_flag, video_signal_type_present_flag;
	uint32_t video_format;
	int video_full_range_flag, colour_description_present_flag;
	uint32_t colour_primaries, transfer_characteristics, matrix_coefficients;
	int chroma_loc_info_present_flag;
	uint32_t chroma_sample_loc_type_top_field, chroma_sample_loc_type_bottom_field;
	int timing_info_present_flag;
	uint32_t num_units_in_tick, time_scale;
	int fixed_frame_rate_flag, nal_hrd_parameters_present_flag;
	struct h264_hrd nal_hrd;
	int vcl_hrd_parameters_present_flag;
	struct h264_hrd vcl_hrd;
	int low_delay_hrd_flag, pic_struct_present_flag, bitstream_restriction_flag;
	int motion_vectors_over_pic_boundaries_flag;
	uint32_t max_bytes_per_pic_denom, max_bits_per_mb_denom;
	uint32_t log2_max_mv_length_horizontal, log2_max_mv_length_vertical;
	uint32_t max_num_reorder_frames, max_dec_frame_buffering;
};

/* 7.3.2.1 */
struct h264_sps {
	uint32_t profile_idc;
	int constraint_set0_flag, constraint_set1_flag, constraint_set2_flag;
	int constraint_set3_flag, constraint_set4_flag, constraint_set5_flag;
	uint32_t reserved_zero_2bits, level_idc, seq_parameter_set_id, chroma_format_idc;
	int separate_colour_plane_flag;
	uint32_t bit_depth_luma_minus8, bit_depth_chroma_minus8;
	int qpprime_y_zero_transform_bypass_flag, seq_scaling_matrix_present_flag;
	struct h264_scaling_matrix seq_scaling_matrix;
	uint32_t log2_max_frame_num_minus4, pic_order_cnt_type, log2_max_pic_order_cnt_lsb_minus4;
	int delta_pic_order_always_zero_flag;
	int32_t offset_for_non_ref_pic, offset_for_top_to_bottom_field;
	uint32_t num_ref_frames_in_pic_order_cnt_cycle; /* 0..255 */
	int32_t offset_for_ref_frame[256];
	uint32_t max_num_ref_frames;
	int gaps_in_frame_num_value_allowed_flag;
	uint32_t pic_width_in_mbs_minus1, pic_height_in_map_units_minus1;
	int frame_mbs_only_flag, mb_adaptive_frame_field_flag, direct_8x8_inference_flag;
	int frame_cropping_flag;
	uint32_t frame_crop_left_offset, frame_crop_right_offset;
	uint32_t frame_crop_top_offset, frame_crop_bottom_offset;
	int vui_parameters_present_flag;
	struct h264_vui vui;
};

/* 7.3.2.2 */
struct h264_pps {
	uint32_t pic_parameter_set_id, seq_parameter_set_id;
	int entropy_coding_mode_flag, bottom_field_pic_order_in_frame_present_flag;
	uint32_t num_slice_groups_minus1; /* 0..7 */
	uint32_t slice_group_map_type;
	uint32_t run_length_minus1[8], top_left[8], bottom_right[8];
	int slice_group_change_direction_flag;
	uint32_t slice_group_change_rate_minus1, pic_size_in_map_units_minus1;
	uint32_t slice_group_id[256];
	uint32_t num_ref_idx_l0_default_active_minus1, num_ref_idx_l1_default_active_minus1;
	int weighted_pred_flag;
	uint32_t weighted_bipred_idc;
	int32_t pic_init_qp_minus26, pic_init_qs_minus26, chroma_qp_index_offset;
	int deblocking_filter_control_present_flag, constrained_intra_pred_flag;
	int redundant_pic_cnt_present_flag;
	int _more_rbsp_data_present;
	int transform_8x8_mode_flag, pic_scaling_matrix_present_flag;
	struct h264_scaling_matrix pic_scaling_matrix;
	int32_t second_chroma_qp_index_offset;
};

/* 7.3.2.4 */
struct h264_aud {
	uint32_t primary_pic_type;
};

/* 7.3.3.1 / H.7.3.3.1.1 */
struct h264_rplm_item {
	uint32_t modification_of_pic_nums_idc;
	union {
		uint32_t abs_diff_pic_num_minus1, long_term_pic_num, abs_diff_view_idx_minus1;
	};
};
struct h264_rplm {
	int ref_pic_list_modification_flag_l0;
	struct h264_rplm_item pic_num_l0[32];
	int ref_pic_list_modification_flag_l1;
	struct h264_rplm_item pic_num_l1[32];
};

/* 7.3.3.2 */
struct h264_pwt_item {
	int luma_weight_flag;
	int32_t luma_weight, luma_offset;
	int chroma_weight_flag;
	int32_t chroma_weight[2], chroma_offset[2];
};
struct h264_pwt {
	uint32_t luma_log2_weight_denom, chroma_log2_weight_denom;
	struct h264_pwt_item l0[32], l1[32];
};

/* 7.3.3.3 */
struct h264_drpm_item {
	uint32_t memory_management_control_operation, difference_of_pic_nums_minus1;
	uint32_t long_term_pic_num, long_term_frame_idx, max_long_term_frame_idx_plus1;
};
struct h264_drpm {
	int no_output_of_prior_pics_flag, long_term_reference_flag;
	int adaptive_ref_pic_marking_mode_flag;
	struct h264_drpm_item mm[64];
};

/* 7.3.3 */
struct h264_slice_header {
	uint32_t first_mb_in_slice, slice_type, pic_parameter_set_id, colour_plane_id, frame_num;
	int field_pic_flag, bottom_field_flag;
	uint32_t idr_pic_id, pic_order_cnt_lsb;
	int32_t delta_pic_order_cnt_bottom, delta_pic_order_cnt[2];
	uint32_t redundant_pic_cnt;
	int direct_spatial_mv_pred_flag, num_ref_idx_active_override_flag;
	uint32_t num_ref_idx_l0_active_minus1, num_ref_idx_l1_active_minus1; /* 0..31 */
	struct h264_rplm rplm;
	struct h264_pwt pwt;
	struct h264_drpm drpm;
	uint32_t cabac_init_idc;
	int32_t slice_qp_delta;
	int sp_for_switch_flag;
	int32_t slice_qs_delta;
	uint32_t disable_deblocking_filter_idc;
	int32_t slice_alpha_c0_offset_div2, slice_beta_offset_div2;
	uint32_t slice_group_change_cycle;
};

/* D.1.1 */
struct h264_sei_buffering_period {
	uint32_t seq_parameter_set_id;
	struct {
		uint32_t initial_cpb_removal_delay, initial_cpb_removal_delay_offset;
	} nal_hrd_cpb[32];
	struct {
		uint32_t initial_cpb_removal_delay, initial_cpb_removal_delay_offset;
	} vcl_hrd_cpb[32];
};

/* D.1.2 */
struct h264_sei_pic_timing {
	uint32_t cpb_removal_delay, dpb_output_delay, pic_struct;
	struct {
		int clock_timestamp_flag;
		uint32_t ct_type;
		int nuit_field_based_flag;
		uint32_t counting_type;
		int full_timestamp_flag, discontinuity_flag, cnt_dropped_flag;
		uint32_t n_frames, seconds_value, minutes_value, hours_value;
		int seconds_flag, minutes_flag, hours_flag;
		int32_t time_offset;
	} clk_ts[3];
};

/* D.1.3 */
struct h264_sei_pan_scan_rect {
	uint32_t pan_scan_rect_id;
	int pan_scan_rect_cancel_flag;
	uint32_t pan_scan_cnt_minus1; /* 0..2 */
	struct {
		int32_t left_offset, right_offset, top_offset, bottom_offset;
	} pan_scan_rect[4];
	uint32_t pan_scan_rect_repetition_period;
};

/* D.1.4 */
struct h264_sei_filler_payload {
	const uint8_t *buf;
	size_t len;
};

/* D.1.5 */
struct h264_sei_user_data_registered {
	uint32_t country_code, country_code_extension_byte;
	const uint8_t *buf;
	size_t len;
};

/* D.1.6 */
struct h264_sei_user_data_unregistered {
	uint8_t uuid[16];
	const uint8_t *buf;
	size_t len;
};

/* D.1.7 */
struct h264_sei_recovery_point {
	uint32_t recovery_frame_cnt;
	int exact_match_flag, broken_link_flag;
	uint32_t changing_slice_group_idc;
};

struct h264_sei {
	enum h264_sei_type type;
	union {
		struct h264_sei_buffering_period buffering_period;
		struct h264_sei_pic_timing pic_timing;
		struct h264_sei_pan_scan_rect pan_scan_rect;
		struct h264_sei_filler_payload filler_payload;
		struct h264_sei_user_data_registered user_data_registered;
		struct h264_sei_user_data_unregistered user_data_unregistered;
		struct h264_sei_recovery_point recovery_point;
	};
	/* de-escaped payload bytes, owned by the context */
	struct {
		uint8_t *buf;
		size_t len;
	} raw;
};

/* variables derived from an SPS (7.4.2.1.1, 6.2) */
struct h264_sps_derived {
	uint32_t ChromaArrayType, SubWidthC, SubHeightC, MbWidthC, MbHeightC;
	uint32_t BitDepthLuma, QpBdOffsetLuma, BitDepthChroma, QpBdOffsetChroma, RawMbBits;
	uint32_t MaxFrameNum, MaxPicOrderCntLsb;
	uint32_t PicWidthInMbs, PicWidthInSamplesLuma, PicWidthInSamplesChroma;
	uint32_t PicHeightInMapUnits, PicSizeInMapUnits, FrameHeightInMbs;
	uint32_t CropUnitX, CropUnitY;
	uint32_t Width, Height;
};

/* summary of an SPS + PPS pair */
struct h264_info {
	uint32_t width, height;     /* picture size in pixels */
	uint8_t bit_depth_luma;
	uint32_t sar_width, sar_height; /* 1:1 when unknown */
	uint32_t crop_left, crop_top, crop_width, crop_height;
	int full_range;
	int colour_description_present;
	uint32_t colour_primaries, transfer_characteristics, matrix_coefficients;
	uint32_t num_units_in_tick, time_scale; /* 0 when unknown */
	float framerate;
	uint32_t framerate_num, framerate_den;
	uint32_t nal_hrd_bitrate, nal_hrd_cpb_size, vcl_hrd_bitrate, vcl_hrd_cpb_size;
};

H264_API const char *h264_nalu_type_str(enum h264_nalu_type val);
H264_API const char *h264_slice_type_str(enum h264_slice_type val);
H264_API const char *h264_mb_type_str(enum h264_mb_type val);
H264_API int h264_mb_type_is_intra(enum h264_mb_type val);
H264_API int h264_mb_type_is_inter(enum h264_mb_type val);
H264_API const char *h264_profile_str(enum h264_profile val);
H264_API const char *h264_color_format_str(enum h264_color_format val);
H264_API char *h264_aspect_ratio_str_alloc(enum h264_aspect_ratio val, uint32_t sar_width,
					   uint32_t sar_height);
H264_API const char *h264_sei_type_str(enum h264_sei_type val);

#endif /* H264B200_TYPES_H */
