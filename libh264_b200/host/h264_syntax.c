/*
 * h264_syntax.c — header syntax of Rec. ITU-T H.264 clause 7.3 / Annex D / Annex E,
 * written ONCE and walked in either direction: every field goes through f_u / f_ue /
 * f_se / ..., which read into or write from the same struct member depending on
 * io->mode.  (The reference instantiates an X-macro template three times,
 * src/h264_syntax.h + src/h264_syntax_ops.h; this is a run-time dispatch instead.)
 *
 * Behaviour kept identical to the reference, field for field, including its range
 * guards (-EIO), what it zero-initialises, and the order of context updates and
 * callbacks in h264_syntax_nalu (src/h264_syntax.h:1446-1604).
 */
#include "h264_priv.h"

#define TRY(expr)                                                                              \
	do {                                                                                   \
		int _r = (expr);                                                               \
		if (_r < 0)                                                                    \
			return _r;                                                             \
	} while (0)
#define GUARD(cond)                                                                            \
	do {                                                                                   \
		if (cond)                                                                      \
			return -EIO;                                                           \
	} while (0)
#define READING (io->mode == H264_IO_READ)

/* ---- field primitives ------------------------------------------------------------- */

static int f_u(struct h264_io *io, uint32_t *v, uint32_t n)
{
	if (READING) {
		uint32_t t = 0;
		int r = h264_bs_read_bits(io->bs, &t, n);
		if (r < 0)
			return r;
		*v = t;
		return 0;
	}
	int r = h264_bs_write_bits(io->bs, *v, n);
	return r < 0 ? r : 0;
}

static int f_flag(struct h264_io *io, int *v)
{
	uint32_t t = (uint32_t)*v;
	TRY(f_u(io, &t, 1));
	*v = (int)t;
	return 0;
}

static int f_u8(struct h264_io *io, uint8_t *v)
{
	uint32_t t = *v;
	TRY(f_u(io, &t, 8));
	*v = (uint8_t)t;
	return 0;
}

static int f_i(struct h264_io *io, int32_t *v, uint32_t n)
{
	if (READING) {
		int32_t t = 0;
		int r = h264_bs_read_bits_i(io->bs, &t, n);
		if (r < 0)
			return r;
		*v = t;
		return 0;
	}
	int r = h264_bs_write_bits_i(io->bs, *v, n);
	return r < 0 ? r : 0;
}

static int f_ue(struct h264_io *io, uint32_t *v)
{
	if (READING) {
		uint32_t t = 0;
		int r = h264_bs_read_bits_ue(io->bs, &t);
		if (r < 0)
			return r;
		*v = t;
		return 0;
	}
	int r = h264_bs_write_bits_ue(io->bs, *v);
	return r < 0 ? r : 0;
}

static int f_se(struct h264_io *io, int32_t *v)
{
	if (READING) {
		int32_t t = 0;
		int r = h264_bs_read_bits_se(io->bs, &t);
		if (r < 0)
			return r;
		*v = t;
		return 0;
	}
	int r = h264_bs_write_bits_se(io->bs, *v);
	return r < 0 ? r : 0;
}

static int f_trailing(struct h264_io *io)
{
	return READING ? h264_bs_read_rbsp_trailing_bits(io->bs)
		       : h264_bs_write_rbsp_trailing_bits(io->bs);
}

/* ---- 7.3.1 -------------------------------------------------------------------------- */

int h264_syntax_nalu_header(struct h264_io *io, struct h264_nalu_header *nh)
{
	TRY(f_u(io, &nh->forbidden_zero_bit, 1));
	GUARD(nh->forbidden_zero_bit != 0);
	TRY(f_u(io, &nh->nal_ref_idc, 2));
	TRY(f_u(io, &nh->nal_unit_type, 5));
	return 0;
}

/* ---- E.1.2, E.1.1 ------------------------------------------------------------------- */

static int syn_hrd(struct h264_io *io, struct h264_hrd *h)
{
	TRY(f_ue(io, &h->cpb_cnt_minus1));
	TRY(f_u(io, &h->bit_rate_scale, 4));
	TRY(f_u(io, &h->cpb_size_scale, 4));
	GUARD(h->cpb_cnt_minus1 > COUNT_OF(h->cpb)); /* the reference's (off-by-one) guard */
	for (uint32_t i = 0; i <= h->cpb_cnt_minus1 && i < COUNT_OF(h->cpb); i++) {
		TRY(f_ue(io, &h->cpb[i].bit_rate_value_minus1));
		TRY(f_ue(io, &h->cpb[i].cpb_size_value_minus1));
		TRY(f_flag(io, &h->cpb[i].cbr_flag));
	}
	TRY(f_u(io, &h->initial_cpb_removal_delay_length_minus1, 5));
	TRY(f_u(io, &h->cpb_removal_delay_length_minus1, 5));
	TRY(f_u(io, &h->dpb_output_delay_length_minus1, 5));
	TRY(f_u(io, &h->time_offset_length, 5));
	return 0;
}

static int syn_vui(struct h264_io *io, struct h264_vui *v)
{
	TRY(f_flag(io, &v->aspect_ratio_info_present_flag));
	if (v->aspect_ratio_info_present_flag) {
		TRY(f_u(io, &v->aspect_ratio_idc, 8));
		if (v->aspect_ratio_idc == H264_ASPECT_RATIO_EXTENDED_SAR) {
			TRY(f_u(io, &v->sar_width, 16));
			TRY(f_u(io, &v->sar_height, 16));
		}
	}
	TRY(f_flag(io, &v->overscan_info_present_flag));
	if (v->overscan_info_present_flag)
		TRY(f_flag(io, &v->overscan_appropriate_flag));
	TRY(f_flag(io, &v->video_signal_type_present_flag));
	if (v->video_signal_type_present_flag) {
		TRY(f_u(io, &v->video_format, 3));
		TRY(f_flag(io, &v->video_full_range_flag));
		TRY(f_flag(io, &v->colour_description_present_flag));
		if (v->colour_description_present_flag) {
			TRY(f_u(io, &v->colour_primaries, 8));
			TRY(f_u(io, &v->transfer_characteristics, 8));
			TRY(f_u(io, &v->matrix_coefficients, 8));
		}
	}
	TRY(f_flag(io, &v->chroma_loc_info_present_flag));
	if (v->chroma_loc_info_present_flag) {
		TRY(f_ue(io, &v->chroma_sample_loc_type_top_field));
		TRY(f_ue(io, &v->chroma_sample_loc_type_bottom_field));
	}
	TRY(f_flag(io, &v->timing_info_present_flag));
	if (v->timing_info_present_flag) {
		TRY(f_u(io, &v->num_units_in_tick, 32));
		TRY(f_u(io, &v->time_scale, 32));
		TRY(f_flag(io, &v->fixed_frame_rate_flag));
	}
	TRY(f_flag(io, &v->nal_hrd_parameters_present_flag));
	if (v->nal_hrd_parameters_present_flag)
		TRY(syn_hrd(io, &v->nal_hrd));
	TRY(f_flag(io, &v->vcl_hrd_parameters_present_flag));
	if (v->vcl_hrd_parameters_present_flag)
		TRY(syn_hrd(io, &v->vcl_hrd));
	if (v->nal_hrd_parameters_present_flag || v->vcl_hrd_parameters_present_flag)
		TRY(f_flag(io, &v->low_delay_hrd_flag));
	TRY(f_flag(io, &v->pic_struct_present_flag));
	TRY(f_flag(io, &v->bitstream_restriction_flag));
	if (v->bitstream_restriction_flag) {
		TRY(f_flag(io, &v->motion_vectors_over_pic_boundaries_flag));
		TRY(f_ue(io, &v->max_bytes_per_pic_denom));
		TRY(f_ue(io, &v->max_bits_per_mb_denom));
		TRY(f_ue(io, &v->log2_max_mv_length_horizontal));
		TRY(f_ue(io, &v->log2_max_mv_length_vertical));
		TRY(f_ue(io, &v->max_num_reorder_frames));
		TRY(f_ue(io, &v->max_dec_frame_buffering));
	}
	return 0;
}

/* ---- 7.3.2.1.1.1 scaling lists -------------------------------------------------------- */

static int syn_scaling_list(struct h264_io *io, int32_t *list, uint32_t size, int *use_default,
			    int *optimized)
{
	int32_t last = 8, next = 8, delta = 0;
	if (READING) {
		for (uint32_t i = 0; i < size; i++) {
			if (next != 0) {
				TRY(f_se(io, &delta));
				next = (last + delta + 256) % 256;
				*use_default = i == 0 && next == 0;
				*optimized = next == 0;
			}
			list[i] = next == 0 ? last : next;
			last = list[i];
		}
		return 0;
	}
	/* writing: when the list was read as "terminated early", terminate it again at
	 * the start of its constant tail so that the bytes come out identical */
	uint32_t tail = 0;
	if (*optimized) {
		for (uint32_t i = size - 1; i >= 1 && list[i] == list[i - 1]; i--)
			tail++;
		if (tail == size - 1 && list[0] == last)
			tail++;
	}
	for (uint32_t i = 0; i < size && next != 0; i++) {
		next = i < size - tail ? list[i] : 0;
		delta = (int8_t)((next - last) % 256);
		TRY(f_se(io, &delta));
		last = list[i];
	}
	return 0;
}

static int syn_scaling_matrix(struct h264_io *io, struct h264_scaling_matrix *m, uint32_t count)
{
	for (uint32_t i = 0; i < count; i++) {
		TRY(f_flag(io, &m->scaling_list_present_flag[i]));
		if (!m->scaling_list_present_flag[i])
			continue;
		if (i < 6)
			TRY(syn_scaling_list(io, m->scaling_list_4x4[i], 16, &m->use_default_4x4[i],
					     &m->_optimized_4x4[i]));
		else
			TRY(syn_scaling_list(io, m->scaling_list_8x8[i - 6], 64,
					     &m->use_default_8x8[i - 6], &m->_optimized_8x8[i - 6]));
	}
	return 0;
}

/* ---- 7.3.2.1 SPS ---------------------------------------------------------------------- */

static int profile_has_chroma_info(uint32_t p)
{
	static const uint8_t list[] = {100, 110, 122, 244, 44, 83, 86, 118, 128, 138, 139, 134, 135};
	for (size_t i = 0; i < sizeof(list); i++)
		if (p == list[i])
			return 1;
	return 0;
}

int h264_syntax_sps(struct h264_io *io, struct h264_sps *s)
{
	TRY(f_u(io, &s->profile_idc, 8));
	TRY(f_flag(io, &s->constraint_set0_flag));
	TRY(f_flag(io, &s->constraint_set1_flag));
	TRY(f_flag(io, &s->constraint_set2_flag));
	TRY(f_flag(io, &s->constraint_set3_flag));
	TRY(f_flag(io, &s->constraint_set4_flag));
	TRY(f_flag(io, &s->constraint_set5_flag));
	TRY(f_u(io, &s->reserved_zero_2bits, 2));
	TRY(f_u(io, &s->level_idc, 8));
	TRY(f_ue(io, &s->seq_parameter_set_id));
	if (profile_has_chroma_info(s->profile_idc)) {
		TRY(f_ue(io, &s->chroma_format_idc));
		if (s->chroma_format_idc == 3)
			TRY(f_flag(io, &s->separate_colour_plane_flag));
		TRY(f_ue(io, &s->bit_depth_luma_minus8));
		GUARD(s->bit_depth_luma_minus8 > 6);
		TRY(f_ue(io, &s->bit_depth_chroma_minus8));
		GUARD(s->bit_depth_chroma_minus8 > 6);
		TRY(f_flag(io, &s->qpprime_y_zero_transform_bypass_flag));
		TRY(f_flag(io, &s->seq_scaling_matrix_present_flag));
		if (s->seq_scaling_matrix_present_flag)
			TRY(syn_scaling_matrix(io, &s->seq_scaling_matrix,
					       s->chroma_format_idc != 3 ? 8 : 12));
	}
	TRY(f_ue(io, &s->log2_max_frame_num_minus4));
	TRY(f_ue(io, &s->pic_order_cnt_type));
	if (s->pic_order_cnt_type == 0) {
		TRY(f_ue(io, &s->log2_max_pic_order_cnt_lsb_minus4));
	} else if (s->pic_order_cnt_type == 1) {
		TRY(f_flag(io, &s->delta_pic_order_always_zero_flag));
		TRY(f_se(io, &s->offset_for_non_ref_pic));
		TRY(f_se(io, &s->offset_for_top_to_bottom_field));
		TRY(f_ue(io, &s->num_ref_frames_in_pic_order_cnt_cycle));
		GUARD(s->num_ref_frames_in_pic_order_cnt_cycle >= COUNT_OF(s->offset_for_ref_frame));
		for (uint32_t i = 0; i < s->num_ref_frames_in_pic_order_cnt_cycle; i++)
			TRY(f_se(io, &s->offset_for_ref_frame[i]));
	}
	TRY(f_ue(io, &s->max_num_ref_frames));
	TRY(f_flag(io, &s->gaps_in_frame_num_value_allowed_flag));
	TRY(f_ue(io, &s->pic_width_in_mbs_minus1));
	TRY(f_ue(io, &s->pic_height_in_map_units_minus1));
	TRY(f_flag(io, &s->frame_mbs_only_flag));
	if (!s->frame_mbs_only_flag)
		TRY(f_flag(io, &s->mb_adaptive_frame_field_flag));
	TRY(f_flag(io, &s->direct_8x8_inference_flag));
	TRY(f_flag(io, &s->frame_cropping_flag));
	if (s->frame_cropping_flag) {
		TRY(f_ue(io, &s->frame_crop_left_offset));
		TRY(f_ue(io, &s->frame_crop_right_offset));
		TRY(f_ue(io, &s->frame_crop_top_offset));
		TRY(f_ue(io, &s->frame_crop_bottom_offset));
	}
	TRY(f_flag(io, &s->vui_parameters_present_flag));
	if (s->vui_parameters_present_flag)
		TRY(syn_vui(io, &s->vui));
	return f_trailing(io);
}

/* ---- 7.3.2.2 PPS (after the two ids) --------------------------------------------------- */

int h264_syntax_pps_body(struct h264_io *io, const struct h264_sps *sps, struct h264_pps *p)
{
	TRY(f_flag(io, &p->entropy_coding_mode_flag));
	TRY(f_flag(io, &p->bottom_field_pic_order_in_frame_present_flag));
	TRY(f_ue(io, &p->num_slice_groups_minus1));
	if (p->num_slice_groups_minus1 > 0) {
		TRY(f_ue(io, &p->slice_group_map_type));
		switch (p->slice_group_map_type) {
		case 0:
			GUARD(p->num_slice_groups_minus1 > COUNT_OF(p->run_length_minus1));
			for (uint32_t i = 0; i <= p->num_slice_groups_minus1 &&
					     i < COUNT_OF(p->run_length_minus1);
			     i++)
				TRY(f_ue(io, &p->run_length_minus1[i]));
			break;
		case 1:
			break;
		case 2:
			GUARD(p->num_slice_groups_minus1 >= COUNT_OF(p->top_left));
			for (uint32_t i = 0; i < p->num_slice_groups_minus1; i++) {
				TRY(f_ue(io, &p->top_left[i]));
				TRY(f_ue(io, &p->bottom_right[i]));
			}
			break;
		case 3:
		case 4:
		case 5:
			TRY(f_flag(io, &p->slice_group_change_direction_flag));
			TRY(f_ue(io, &p->slice_group_change_rate_minus1));
			break;
		case 6: {
			TRY(f_ue(io, &p->pic_size_in_map_units_minus1));
			const uint32_t bits = h264_ceil_log2(p->num_slice_groups_minus1 + 1);
			GUARD(p->pic_size_in_map_units_minus1 > COUNT_OF(p->slice_group_id));
			for (uint32_t i = 0; i <= p->pic_size_in_map_units_minus1 &&
					     i < COUNT_OF(p->slice_group_id);
			     i++)
				TRY(f_u(io, &p->slice_group_id[i], bits));
			break;
		}
		default:
			return -EIO;
		}
	}
	TRY(f_ue(io, &p->num_ref_idx_l0_default_active_minus1));
	TRY(f_ue(io, &p->num_ref_idx_l1_default_active_minus1));
	TRY(f_flag(io, &p->weighted_pred_flag));
	TRY(f_u(io, &p->weighted_bipred_idc, 2));
	TRY(f_se(io, &p->pic_init_qp_minus26));
	TRY(f_se(io, &p->pic_init_qs_minus26));
	TRY(f_se(io, &p->chroma_qp_index_offset));
	TRY(f_flag(io, &p->deblocking_filter_control_present_flag));
	TRY(f_flag(io, &p->constrained_intra_pred_flag));
	TRY(f_flag(io, &p->redundant_pic_cnt_present_flag));
	if (READING && h264_bs_more_rbsp_data(io->bs))
		p->_more_rbsp_data_present = 1;
	if (p->_more_rbsp_data_present) {
		TRY(f_flag(io, &p->transform_8x8_mode_flag));
		TRY(f_flag(io, &p->pic_scaling_matrix_present_flag));
		if (p->pic_scaling_matrix_present_flag) {
			uint32_t n = 6;
			if (p->transform_8x8_mode_flag)
				n += sps->chroma_format_idc != 3 ? 2 : 6;
			TRY(syn_scaling_matrix(io, &p->pic_scaling_matrix, n));
		}
		TRY(f_se(io, &p->second_chroma_qp_index_offset));
	}
	return f_trailing(io);
}

static int syn_pps(struct h264_io *io, struct h264_pps *p)
{
	TRY(f_ue(io, &p->pic_parameter_set_id));
	TRY(f_ue(io, &p->seq_parameter_set_id));
	int r = h264_ctx_activate_sps(io->ctx, p->seq_parameter_set_id);
	if (r < 0)
		return r;
	return h264_syntax_pps_body(io, io->ctx->sps, p);
}

/* ---- Annex D SEI payloads -------------------------------------------------------------- */

static int sei_hrd_delays(struct h264_io *io, const struct h264_hrd *h, uint32_t *delay0,
			  size_t stride_words, size_t cap)
{
	const uint32_t n = h->initial_cpb_removal_delay_length_minus1 + 1;
	GUARD(h->cpb_cnt_minus1 > cap);
	for (uint32_t i = 0; i <= h->cpb_cnt_minus1 && i < cap; i++) {
		TRY(f_u(io, delay0 + i * stride_words, n));
		TRY(f_u(io, delay0 + i * stride_words + 1, n));
	}
	return 0;
}

static int sei_buffering_period(struct h264_io *io, struct h264_sei_buffering_period *s)
{
	TRY(f_ue(io, &s->seq_parameter_set_id));
	int r = h264_ctx_activate_sps(io->ctx, s->seq_parameter_set_id);
	if (r < 0)
		return r;
	const struct h264_vui *v = &io->ctx->sps->vui;
	if (v->nal_hrd_parameters_present_flag)
		TRY(sei_hrd_delays(io, &v->nal_hrd, &s->nal_hrd_cpb[0].initial_cpb_removal_delay, 2,
				   COUNT_OF(s->nal_hrd_cpb)));
	if (v->vcl_hrd_parameters_present_flag)
		TRY(sei_hrd_delays(io, &v->vcl_hrd, &s->vcl_hrd_cpb[0].initial_cpb_removal_delay, 2,
				   COUNT_OF(s->vcl_hrd_cpb)));
	return 0;
}

static int sei_pic_timing(struct h264_io *io, struct h264_sei_pic_timing *s)
{
	static const uint8_t clock_ts_count[16] = {1, 1, 1, 2, 2, 3, 3, 2, 3};
	const struct h264_sps *sps = io->ctx->sps;
	GUARD(sps == NULL);
	const struct h264_vui *v = &sps->vui;
	const struct h264_hrd *h = v->nal_hrd_parameters_present_flag   ? &v->nal_hrd
				   : v->vcl_hrd_parameters_present_flag ? &v->vcl_hrd
									: NULL;
	if (h != NULL) {
		TRY(f_u(io, &s->cpb_removal_delay, h->cpb_removal_delay_length_minus1 + 1));
		TRY(f_u(io, &s->dpb_output_delay, h->dpb_output_delay_length_minus1 + 1));
	}
	if (!v->pic_struct_present_flag)
		return 0;
	TRY(f_u(io, &s->pic_struct, 4));
	for (uint32_t i = 0; i < clock_ts_count[s->pic_struct & 15]; i++) {
		TRY(f_flag(io, &s->clk_ts[i].clock_timestamp_flag));
		if (!s->clk_ts[i].clock_timestamp_flag)
			continue;
		TRY(f_u(io, &s->clk_ts[i].ct_type, 2));
		TRY(f_flag(io, &s->clk_ts[i].nuit_field_based_flag));
		TRY(f_u(io, &s->clk_ts[i].counting_type, 5));
		TRY(f_flag(io, &s->clk_ts[i].full_timestamp_flag));
		TRY(f_flag(io, &s->clk_ts[i].discontinuity_flag));
		TRY(f_flag(io, &s->clk_ts[i].cnt_dropped_flag));
		TRY(f_u(io, &s->clk_ts[i].n_frames, 8));
		if (s->clk_ts[i].full_timestamp_flag) {
			TRY(f_u(io, &s->clk_ts[i].seconds_value, 6));
			TRY(f_u(io, &s->clk_ts[i].minutes_value, 6));
			TRY(f_u(io, &s->clk_ts[i].hours_value, 5));
		} else {
			/* seconds, then minutes, then hours, each only if the previous is there */
			TRY(f_flag(io, &s->clk_ts[i].seconds_flag));
			if (s->clk_ts[i].seconds_flag) {
				TRY(f_u(io, &s->clk_ts[i].seconds_value, 6));
				TRY(f_flag(io, &s->clk_ts[i].minutes_flag));
				if (s->clk_ts[i].minutes_flag) {
					TRY(f_u(io, &s->clk_ts[i].minutes_value, 6));
					TRY(f_flag(io, &s->clk_ts[i].hours_flag));
					if (s->clk_ts[i].hours_flag)
						TRY(f_u(io, &s->clk_ts[i].hours_value, 5));
				}
			}
		}
		const uint32_t n = h != NULL ? h->time_offset_length : 24;
		if (n > 0)
			TRY(f_i(io, &s->clk_ts[i].time_offset, n));
	}
	return 0;
}

static int sei_pan_scan_rect(struct h264_io *io, struct h264_sei_pan_scan_rect *s)
{
	TRY(f_ue(io, &s->pan_scan_rect_id));
	TRY(f_flag(io, &s->pan_scan_rect_cancel_flag));
	if (s->pan_scan_rect_cancel_flag)
		return 0;
	TRY(f_ue(io, &s->pan_scan_cnt_minus1));
	GUARD(s->pan_scan_cnt_minus1 > COUNT_OF(s->pan_scan_rect));
	for (uint32_t i = 0; i <= s->pan_scan_cnt_minus1 && i < COUNT_OF(s->pan_scan_rect); i++) {
		TRY(f_se(io, &s->pan_scan_rect[i].left_offset));
		TRY(f_se(io, &s->pan_scan_rect[i].right_offset));
		TRY(f_se(io, &s->pan_scan_rect[i].top_offset));
		TRY(f_se(io, &s->pan_scan_rect[i].bottom_offset));
	}
	return f_ue(io, &s->pan_scan_rect_repetition_period);
}

/* the rest of the payload as opaque bytes: a pointer into the payload when reading */
static int sei_bytes(struct h264_io *io, const uint8_t **buf, size_t *len)
{
	if (READING) {
		GUARD(!h264_bs_byte_aligned(io->bs));
		*buf = io->bs->cdata + io->bs->off;
		*len = io->bs->len - io->bs->off;
		return 0;
	}
	GUARD(*len != 0 && *buf == NULL);
	for (size_t i = 0; i < *len; i++) {
		uint32_t b = (*buf)[i];
		TRY(f_u(io, &b, 8));
	}
	return 0;
}

/* one SEI message body; io->bs is a bitstream over the payload alone (no escaping) */
int h264_syntax_sei_payload(struct h264_io *io, struct h264_sei *sei)
{
	const struct h264_ctx_cbs *cbs = io->cbs;
	const uint8_t *raw = sei->raw.buf;
	const size_t rawlen = sei->raw.len;
#define NOTIFY(cb, member)                                                                     \
	do {                                                                                   \
		if (READING && cbs != NULL && cbs->cb != NULL)                                 \
			cbs->cb(io->ctx, raw, rawlen, &sei->member, io->userdata);             \
	} while (0)
	switch (sei->type) {
	case H264_SEI_TYPE_BUFFERING_PERIOD:
		TRY(sei_buffering_period(io, &sei->buffering_period));
		NOTIFY(sei_buffering_period, buffering_period);
		break;
	case H264_SEI_TYPE_PIC_TIMING:
		TRY(sei_pic_timing(io, &sei->pic_timing));
		NOTIFY(sei_pic_timing, pic_timing);
		break;
	case H264_SEI_TYPE_PAN_SCAN_RECT:
		TRY(sei_pan_scan_rect(io, &sei->pan_scan_rect));
		NOTIFY(sei_pan_scan_rect, pan_scan_rect);
		break;
	case H264_SEI_TYPE_FILLER_PAYLOAD:
		TRY(sei_bytes(io, &sei->filler_payload.buf, &sei->filler_payload.len));
		NOTIFY(sei_filler_payload, filler_payload);
		break;
	case H264_SEI_TYPE_USER_DATA_REGISTERED:
		TRY(f_u(io, &sei->user_data_registered.country_code, 8));
		if (sei->user_data_registered.country_code == 0xff)
			TRY(f_u(io, &sei->user_data_registered.country_code_extension_byte, 8));
		TRY(sei_bytes(io, &sei->user_data_registered.buf, &sei->user_data_registered.len));
		NOTIFY(sei_user_data_registered, user_data_registered);
		break;
	case H264_SEI_TYPE_USER_DATA_UNREGISTERED:
		for (int i = 0; i < 16; i++)
			TRY(f_u8(io, &sei->user_data_unregistered.uuid[i]));
		TRY(sei_bytes(io, &sei->user_data_unregistered.buf,
			      &sei->user_data_unregistered.len));
		NOTIFY(sei_user_data_unregistered, user_data_unregistered);
		break;
	case H264_SEI_TYPE_RECOVERY_POINT:
		TRY(f_ue(io, &sei->recovery_point.recovery_frame_cnt));
		TRY(f_flag(io, &sei->recovery_point.exact_match_flag));
		TRY(f_flag(io, &sei->recovery_point.broken_link_flag));
		TRY(f_u(io, &sei->recovery_point.changing_slice_group_idc, 2));
		NOTIFY(sei_recovery_point, recovery_point);
		break;
	default:
		return 0; /* other payload types stay raw */
	}
#undef NOTIFY
	if (READING) {
		uint32_t bit;
		while (!h264_bs_byte_aligned(io->bs))
			TRY(f_u(io, &bit, 1)); /* tolerant of a wrong alignment pattern */
	} else if (!h264_bs_byte_aligned(io->bs)) {
		TRY(f_trailing(io));
	}
	return 0;
}

/* 7.3.2.3: sequence of (type, size, payload) */
static int syn_sei(struct h264_io *io)
{
	struct h264_ctx *ctx = io->ctx;
	if (!READING) {
		GUARD(ctx->sei_count == 0);
		for (uint32_t k = 0; k < ctx->sei_count; k++) {
			const struct h264_sei *sei = &ctx->sei_tab[k];
			GUARD(sei->raw.buf == NULL || sei->raw.len == 0);
			int r = h264_bs_write_bits_ff_coded(io->bs, (uint32_t)sei->type);
			if (r < 0)
				return r;
			r = h264_bs_write_bits_ff_coded(io->bs, (uint32_t)sei->raw.len);
			if (r < 0)
				return r;
			for (size_t i = 0; i < sei->raw.len; i++) {
				uint32_t b = sei->raw.buf[i];
				TRY(f_u(io, &b, 8));
			}
		}
		return f_trailing(io);
	}
	do {
		uint32_t type = 0, size = 0;
		int r = h264_bs_read_bits_ff_coded(io->bs, &type);
		if (r < 0)
			return r;
		r = h264_bs_read_bits_ff_coded(io->bs, &size);
		if (r < 0)
			return r;
		struct h264_sei *sei = NULL;
		r = h264_ctx_new_sei(ctx, &sei);
		if (r < 0)
			return r;
		sei->type = (enum h264_sei_type)type;
		sei->raw.buf = malloc(size ? size : 1);
		if (sei->raw.buf == NULL)
			return -ENOMEM;
		sei->raw.len = size;
		for (uint32_t i = 0; i < size; i++)
			TRY(f_u8(io, &sei->raw.buf[i]));
		if (io->cbs != NULL && io->cbs->sei != NULL)
			io->cbs->sei(ctx, sei->type, sei->raw.buf, sei->raw.len, io->userdata);
		struct h264_bitstream pbs;
		h264_bs_cinit(&pbs, sei->raw.buf, sei->raw.len, 0);
		struct h264_io pio = *io;
		pio.bs = &pbs;
		TRY(h264_syntax_sei_payload(&pio, sei));
	} while (h264_bs_more_rbsp_data(io->bs));
	return f_trailing(io);
}

/* ---- 7.3.2.4, 7.3.2.7 ------------------------------------------------------------------ */

static int syn_aud(struct h264_io *io, struct h264_aud *aud)
{
	TRY(f_u(io, &aud->primary_pic_type, 3));
	return f_trailing(io);
}

static int syn_filler(struct h264_io *io, size_t *len)
{
	if (READING) {
		uint32_t b = 0;
		*len = 0;
		while (h264_bs_next_bits(io->bs, &b, 8) == 8 && b == 0xff) {
			TRY(f_u(io, &b, 8));
			(*len)++;
		}
	} else {
		for (size_t i = 0; i < *len; i++) {
			uint32_t b = 0xff;
			TRY(f_u(io, &b, 8));
		}
	}
	return f_trailing(io);
}

/* ---- 7.3.3.1 / 7.3.3.2 / 7.3.3.3 ------------------------------------------------------- */

static int syn_rplm_list(struct h264_io *io, struct h264_rplm_item *it, uint32_t cap)
{
	uint32_t i = 0, op;
	do {
		GUARD(i >= cap);
		TRY(f_ue(io, &it[i].modification_of_pic_nums_idc));
		op = it[i].modification_of_pic_nums_idc;
		if (op == 0 || op == 1)
			TRY(f_ue(io, &it[i].abs_diff_pic_num_minus1));
		else if (op == 2)
			TRY(f_ue(io, &it[i].long_term_pic_num));
		else if (op == 4 || op == 5)
			TRY(f_ue(io, &it[i].abs_diff_view_idx_minus1));
		i++;
	} while (op != 3);
	return 0;
}

static int syn_pwt_entry(struct h264_io *io, struct h264_pwt_item *w, uint32_t chroma_array_type)
{
	TRY(f_flag(io, &w->luma_weight_flag));
	if (w->luma_weight_flag) {
		TRY(f_se(io, &w->luma_weight));
		TRY(f_se(io, &w->luma_offset));
	}
	if (chroma_array_type != 0) {
		TRY(f_flag(io, &w->chroma_weight_flag));
		if (w->chroma_weight_flag) {
			for (int c = 0; c < 2; c++) {
				TRY(f_se(io, &w->chroma_weight[c]));
				TRY(f_se(io, &w->chroma_offset[c]));
			}
		}
	}
	return 0;
}

static int syn_pwt(struct h264_io *io, struct h264_slice_header *sh, enum h264_slice_type type)
{
	const struct h264_sps *sps = io->ctx->sps;
	const uint32_t cat = sps->separate_colour_plane_flag ? 0 : sps->chroma_format_idc;
	struct h264_pwt *t = &sh->pwt;
	TRY(f_ue(io, &t->luma_log2_weight_denom));
	if (cat != 0)
		TRY(f_ue(io, &t->chroma_log2_weight_denom));
	GUARD(sh->num_ref_idx_l0_active_minus1 > COUNT_OF(t->l0));
	for (uint32_t i = 0; i <= sh->num_ref_idx_l0_active_minus1 && i < COUNT_OF(t->l0); i++)
		TRY(syn_pwt_entry(io, &t->l0[i], cat));
	if (type != H264_SLICE_TYPE_B)
		return 0;
	GUARD(sh->num_ref_idx_l1_active_minus1 > COUNT_OF(t->l1));
	for (uint32_t i = 0; i <= sh->num_ref_idx_l1_active_minus1 && i < COUNT_OF(t->l1); i++)
		TRY(syn_pwt_entry(io, &t->l1[i], cat));
	return 0;
}

static int syn_drpm(struct h264_io *io, struct h264_drpm *d)
{
	if (io->ctx->nalu_type == H264_NALU_TYPE_SLICE_IDR) {
		TRY(f_flag(io, &d->no_output_of_prior_pics_flag));
		return f_flag(io, &d->long_term_reference_flag);
	}
	TRY(f_flag(io, &d->adaptive_ref_pic_marking_mode_flag));
	if (!d->adaptive_ref_pic_marking_mode_flag)
		return 0;
	uint32_t i = 0, op;
	do {
		GUARD(i >= COUNT_OF(d->mm));
		struct h264_drpm_item *m = &d->mm[i++];
		TRY(f_ue(io, &m->memory_management_control_operation));
		op = m->memory_management_control_operation;
		if (op == 1 || op == 3)
			TRY(f_ue(io, &m->difference_of_pic_nums_minus1));
		if (op == 2)
			TRY(f_ue(io, &m->long_term_pic_num));
		if (op == 3 || op == 6)
			TRY(f_ue(io, &m->long_term_frame_idx));
		if (op == 4)
			TRY(f_ue(io, &m->max_long_term_frame_idx_plus1));
	} while (op != 0);
	return 0;
}

/* ---- 7.3.3 slice header ------------------------------------------------------------------ */

int h264_syntax_slice_header(struct h264_io *io, struct h264_slice_header *sh)
{
	struct h264_ctx *ctx = io->ctx;
	const int idr = ctx->nalu_type == H264_NALU_TYPE_SLICE_IDR;
	ctx->sh_bits = 0;
	TRY(f_ue(io, &sh->first_mb_in_slice));
	TRY(f_ue(io, &sh->slice_type));
	const enum h264_slice_type type = H264_SLICE_TYPE(sh->slice_type);
	TRY(f_ue(io, &sh->pic_parameter_set_id));
	int r = h264_ctx_activate_pps(ctx, sh->pic_parameter_set_id);
	if (r < 0)
		return r;
	const struct h264_sps *sps = ctx->sps;
	const struct h264_pps *pps = ctx->pps;
	if (READING) {
		sh->num_ref_idx_l0_active_minus1 = pps->num_ref_idx_l0_default_active_minus1;
		sh->num_ref_idx_l1_active_minus1 = pps->num_ref_idx_l1_default_active_minus1;
	}
	if (sps->separate_colour_plane_flag)
		TRY(f_u(io, &sh->colour_plane_id, 2));
	TRY(f_u(io, &sh->frame_num, sps->log2_max_frame_num_minus4 + 4));
	if (!sps->frame_mbs_only_flag) {
		TRY(f_flag(io, &sh->field_pic_flag));
		if (sh->field_pic_flag)
			TRY(f_flag(io, &sh->bottom_field_flag));
	}
	if (idr)
		TRY(f_ue(io, &sh->idr_pic_id));
	const int bottom_delta = pps->bottom_field_pic_order_in_frame_present_flag && !sh->field_pic_flag;
	if (sps->pic_order_cnt_type == 0) {
		TRY(f_u(io, &sh->pic_order_cnt_lsb, sps->log2_max_pic_order_cnt_lsb_minus4 + 4));
		if (bottom_delta)
			TRY(f_se(io, &sh->delta_pic_order_cnt_bottom));
	}
	if (sps->pic_order_cnt_type == 1 && !sps->delta_pic_order_always_zero_flag) {
		TRY(f_se(io, &sh->delta_pic_order_cnt[0]));
		if (bottom_delta)
			TRY(f_se(io, &sh->delta_pic_order_cnt[1]));
	}
	if (pps->redundant_pic_cnt_present_flag)
		TRY(f_ue(io, &sh->redundant_pic_cnt));
	if (type == H264_SLICE_TYPE_B)
		TRY(f_flag(io, &sh->direct_spatial_mv_pred_flag));
	if (type == H264_SLICE_TYPE_P || type == H264_SLICE_TYPE_SP || type == H264_SLICE_TYPE_B) {
		TRY(f_flag(io, &sh->num_ref_idx_active_override_flag));
		if (sh->num_ref_idx_active_override_flag) {
			TRY(f_ue(io, &sh->num_ref_idx_l0_active_minus1));
			if (type == H264_SLICE_TYPE_B)
				TRY(f_ue(io, &sh->num_ref_idx_l1_active_minus1));
		}
	}
	if (type != H264_SLICE_TYPE_I && type != H264_SLICE_TYPE_SI) {
		TRY(f_flag(io, &sh->rplm.ref_pic_list_modification_flag_l0));
		if (sh->rplm.ref_pic_list_modification_flag_l0)
			TRY(syn_rplm_list(io, sh->rplm.pic_num_l0, COUNT_OF(sh->rplm.pic_num_l0)));
	}
	if (type == H264_SLICE_TYPE_B) {
		TRY(f_flag(io, &sh->rplm.ref_pic_list_modification_flag_l1));
		if (sh->rplm.ref_pic_list_modification_flag_l1)
			TRY(syn_rplm_list(io, sh->rplm.pic_num_l1, COUNT_OF(sh->rplm.pic_num_l1)));
	}
	if ((pps->weighted_pred_flag && (type == H264_SLICE_TYPE_P || type == H264_SLICE_TYPE_SP)) ||
	    (pps->weighted_bipred_idc == 1 && type == H264_SLICE_TYPE_B))
		TRY(syn_pwt(io, sh, type));
	if (ctx->nalu_hdr.nal_ref_idc != 0)
		TRY(syn_drpm(io, &sh->drpm));
	if (pps->entropy_coding_mode_flag && type != H264_SLICE_TYPE_I && type != H264_SLICE_TYPE_SI)
		TRY(f_ue(io, &sh->cabac_init_idc));
	TRY(f_se(io, &sh->slice_qp_delta));
	if (type == H264_SLICE_TYPE_SP || type == H264_SLICE_TYPE_SI) {
		if (type == H264_SLICE_TYPE_SP)
			TRY(f_flag(io, &sh->sp_for_switch_flag));
		TRY(f_se(io, &sh->slice_qs_delta));
	}
	if (pps->deblocking_filter_control_present_flag) {
		TRY(f_ue(io, &sh->disable_deblocking_filter_idc));
		if (sh->disable_deblocking_filter_idc != 1) {
			TRY(f_se(io, &sh->slice_alpha_c0_offset_div2));
			TRY(f_se(io, &sh->slice_beta_offset_div2));
		}
	}
	if (pps->num_slice_groups_minus1 > 0 && pps->slice_group_map_type >= 3 &&
	    pps->slice_group_map_type <= 5) {
		const uint32_t units =
			(sps->pic_width_in_mbs_minus1 + 1) * (sps->pic_height_in_map_units_minus1 + 1);
		TRY(f_u(io, &sh->slice_group_change_cycle,
			h264_ceil_log2(units / (pps->slice_group_change_rate_minus1 + 1) + 1)));
	}
	/* raw bit position right after the header (NAL header byte and escapes included) */
	ctx->sh_bits = READING ? io->bs->off * 8 - io->bs->cachebits
			       : io->bs->off * 8 + io->bs->cachebits;
	return 0;
}

/* ---- slice data hand-off (7.3.4 is parsed on the GPU) ------------------------------------ */

static int syn_slice_data(struct h264_io *io, const uint8_t *nal, size_t nal_len)
{
	struct h264_ctx *ctx = io->ctx;
	struct h264_bitstream *bs = io->bs;
	if (READING) {
		/* where slice_data() starts, exactly as the reference records it
		 * (src/h264_syntax_slice_data.h:803-806) */
		ctx->rawdata.partialbits = bs->cachebits;
		ctx->rawdata.partial = (uint8_t)(bs->cache & ((1u << bs->cachebits) - 1));
		ctx->rawdata.buf = bs->cdata + bs->off;
		ctx->rawdata.len = bs->len - bs->off;
		if (io->reader != NULL && (io->reader->flags & H264_READER_FLAGS_SLICE_DATA))
			return h264_reader_slice_data(io->reader, ctx, nal, nal_len);
		return 0;
	}
	/* writing: the bits left over from the header byte, then the raw (still escaped)
	 * slice data bytes as they were read (src/h264_syntax_slice_data.h:813-830) */
	if (ctx->rawdata.partialbits == 0 && ctx->rawdata.len == 0)
		return 0;
	if (ctx->rawdata.partialbits) {
		int r = h264_bs_write_bits(bs, ctx->rawdata.partial, ctx->rawdata.partialbits);
		if (r < 0)
			return r;
	}
	GUARD(!h264_bs_byte_aligned(bs));
	GUARD(ctx->rawdata.len != 0 && ctx->rawdata.buf == NULL);
	return h264_bs_write_raw_bytes(bs, ctx->rawdata.buf, ctx->rawdata.len);
}

/* ---- 7.3.1 NAL unit: dispatch, context updates, callbacks -------------------------------- */

#define CALL(cb, ...)                                                                          \
	do {                                                                                   \
		if (io->cbs != NULL && io->cbs->cb != NULL)                                    \
			io->cbs->cb(ctx, __VA_ARGS__, io->userdata);                           \
	} while (0)

int h264_syntax_nalu(struct h264_io *io)
{
	struct h264_ctx *ctx = io->ctx;
	const uint8_t *buf = NULL;
	size_t len = 0;
	if (READING) {
		buf = io->bs->cdata + io->bs->off;
		len = io->bs->len;
		TRY(h264_ctx_clear_nalu(ctx));
	}
	TRY(h264_syntax_nalu_header(io, &ctx->nalu_hdr));
	ctx->nalu_type = (enum h264_nalu_type)ctx->nalu_hdr.nal_unit_type;
	CALL(nalu_begin, ctx->nalu_type, buf, len, &ctx->nalu_hdr);

	switch (ctx->nalu_type) {
	case H264_NALU_TYPE_SLICE:
	case H264_NALU_TYPE_SLICE_IDR: {
		if (READING) {
			struct h264_slice_header sh;
			memset(&sh, 0, sizeof(sh));
			TRY(h264_syntax_slice_header(io, &sh));
			const size_t bits = ctx->sh_bits;
			TRY(h264_ctx_set_slice_header(ctx, &sh));
			ctx->sh_bits = bits;
		} else {
			TRY(h264_syntax_slice_header(io, &ctx->sh));
		}
		TRY(syn_slice_data(io, buf, len));
		CALL(slice, buf, len, &ctx->sh);
		break;
	}
	case H264_NALU_TYPE_SEI:
		TRY(syn_sei(io));
		break;
	case H264_NALU_TYPE_SPS: {
		GUARD(ctx->nalu_hdr.nal_ref_idc == 0);
		if (READING) {
			struct h264_sps *sps = calloc(1, sizeof(*sps));
			if (sps == NULL)
				return -ENOMEM;
			sps->chroma_format_idc = 1; /* inferred when absent, 7.4.2.1.1 */
			int r = h264_syntax_sps(io, sps);
			if (r >= 0)
				r = h264_ctx_set_sps(ctx, sps);
			free(sps);
			if (r < 0)
				return r;
		} else {
			GUARD(ctx->sps == NULL);
			TRY(h264_syntax_sps(io, ctx->sps));
		}
		CALL(sps, buf, len, ctx->sps);
		break;
	}
	case H264_NALU_TYPE_PPS: {
		GUARD(ctx->nalu_hdr.nal_ref_idc == 0);
		if (READING) {
			struct h264_pps *pps = calloc(1, sizeof(*pps));
			if (pps == NULL)
				return -ENOMEM;
			int r = syn_pps(io, pps);
			if (r >= 0)
				r = h264_ctx_set_pps(ctx, pps);
			free(pps);
			if (r < 0)
				return r;
		} else {
			GUARD(ctx->pps == NULL);
			TRY(syn_pps(io, ctx->pps));
		}
		CALL(pps, buf, len, ctx->pps);
		break;
	}
	case H264_NALU_TYPE_AUD:
		GUARD(ctx->nalu_hdr.nal_ref_idc != 0);
		TRY(syn_aud(io, &ctx->aud));
		CALL(aud, buf, len, &ctx->aud);
		break;
	case H264_NALU_TYPE_FILLER:
		GUARD(ctx->nalu_hdr.nal_ref_idc != 0);
		TRY(syn_filler(io, &ctx->filler_len));
		break;
	default: /* data partitions, end of sequence/stream, extensions: not interpreted */
		ctx->nalu_unknown = 1;
		break;
	}

	if (READING) {
		/* 7.4.1.2.4: this NAL unit opens a new access unit */
		const unsigned t = (unsigned)ctx->nalu_type;
		if ((ctx->prev_vcl || ctx->prev_filler) &&
		    (t == H264_NALU_TYPE_AUD || t == H264_NALU_TYPE_SPS || t == H264_NALU_TYPE_PPS ||
		     t == H264_NALU_TYPE_SEI || (t >= 14 && t <= 18) || ctx->first_vcl)) {
			if (io->cbs != NULL && io->cbs->au_end != NULL)
				io->cbs->au_end(ctx, io->userdata);
		}
		ctx->prev_vcl = t == H264_NALU_TYPE_SLICE || t == H264_NALU_TYPE_SLICE_IDR;
		ctx->prev_filler = t == H264_NALU_TYPE_FILLER;
	}
	CALL(nalu_end, ctx->nalu_type, buf, len, &ctx->nalu_hdr);
	return 0;
}
