/*
 * h264_ctx.c — parsing context: parameter-set tables, active SPS/PPS, derived
 * variables, SEI table, access-unit change detection, stream info.
 * Reference behaviour: src/h264_ctx.c:56-702, src/h264.c:36-118 (same public API,
 * include/h264/h264_ctx.h:155-227); the state layout is this library's own.
 */
#include "h264_priv.h"

/* ---- derived variables ----------------------------------------------------------- */

/* 6.2 + 7.4.2.1.1 */
int h264_get_sps_derived(const struct h264_sps *sps, struct h264_sps_derived *d)
{
	/* SubWidthC, SubHeightC, MbWidthC, MbHeightC per ChromaArrayType */
	static const uint8_t chroma[4][4] = {{0, 0, 0, 0}, {2, 2, 8, 8}, {2, 1, 8, 16}, {1, 1, 16, 16}};
	if (sps == NULL || d == NULL)
		return -EINVAL;
	memset(d, 0, sizeof(*d));
	const uint32_t fields = 2 - (sps->frame_mbs_only_flag ? 1 : 0);
	d->ChromaArrayType = sps->separate_colour_plane_flag ? 0 : sps->chroma_format_idc;
	if (d->ChromaArrayType < 4) {
		d->SubWidthC = chroma[d->ChromaArrayType][0];
		d->SubHeightC = chroma[d->ChromaArrayType][1];
		d->MbWidthC = chroma[d->ChromaArrayType][2];
		d->MbHeightC = chroma[d->ChromaArrayType][3];
	}
	d->BitDepthLuma = sps->bit_depth_luma_minus8 + 8;
	d->QpBdOffsetLuma = 6 * sps->bit_depth_luma_minus8;
	d->BitDepthChroma = sps->bit_depth_chroma_minus8 + 8;
	d->QpBdOffsetChroma = 6 * sps->bit_depth_chroma_minus8;
	d->RawMbBits = 256 * d->BitDepthLuma + 2 * d->MbWidthC * d->MbHeightC * d->BitDepthChroma;
	d->MaxFrameNum = 1u << (sps->log2_max_frame_num_minus4 + 4);
	d->MaxPicOrderCntLsb = 1u << (sps->log2_max_pic_order_cnt_lsb_minus4 + 4);
	d->PicWidthInMbs = sps->pic_width_in_mbs_minus1 + 1;
	d->PicWidthInSamplesLuma = d->PicWidthInMbs * 16;
	d->PicWidthInSamplesChroma = d->PicWidthInMbs * d->MbWidthC;
	d->PicHeightInMapUnits = sps->pic_height_in_map_units_minus1 + 1;
	d->PicSizeInMapUnits = d->PicWidthInMbs * d->PicHeightInMapUnits;
	d->FrameHeightInMbs = fields * d->PicHeightInMapUnits;
	d->CropUnitX = d->ChromaArrayType == 0 ? 1 : d->SubWidthC;
	d->CropUnitY = (d->ChromaArrayType == 0 ? 1 : d->SubHeightC) * fields;
	d->Width = d->PicWidthInSamplesLuma -
		   d->CropUnitX * (sps->frame_crop_left_offset + sps->frame_crop_right_offset);
	d->Height = d->FrameHeightInMbs * 16 -
		    d->CropUnitY * (sps->frame_crop_top_offset + sps->frame_crop_bottom_offset);
	return 0;
}

/* 7.4.3: what follows from the active SPS + PPS + current slice header */
static void derive_slice_vars(struct h264_ctx *ctx)
{
	const struct h264_sps *sps = ctx->sps;
	const struct h264_pps *pps = ctx->pps;
	const struct h264_slice_header *sh = &ctx->sh;
	if (sps == NULL || pps == NULL)
		return;
	const uint32_t field = sh->field_pic_flag ? 1 : 0;
	ctx->MbaffFrameFlag = sps->mb_adaptive_frame_field_flag && !field;
	ctx->PicHeightInMbs = ctx->spsd.FrameHeightInMbs / (1 + field);
	ctx->PicSizeInMbs = ctx->spsd.PicWidthInMbs * ctx->PicHeightInMbs;
	ctx->MaxPicNum = ctx->spsd.MaxFrameNum << field;
	ctx->CurrPicNum = field ? 2 * sh->frame_num + 1 : sh->frame_num;
	ctx->SliceQPLuma = pps->pic_init_qp_minus26 + 26 + sh->slice_qp_delta;
	ctx->QSLuma = pps->pic_init_qs_minus26 + 26 + sh->slice_qs_delta;
	const uint64_t units = (uint64_t)sh->slice_group_change_cycle * ctx->SliceGroupChangeRate;
	ctx->MapUnitsInSliceGroup0 =
		units < ctx->spsd.PicSizeInMapUnits ? (uint32_t)units : ctx->spsd.PicSizeInMapUnits;
}

static void derive_after_sps(struct h264_ctx *ctx)
{
	if (ctx->sps != NULL)
		h264_get_sps_derived(ctx->sps, &ctx->spsd);
	derive_slice_vars(ctx);
}

/* ---- lifecycle --------------------------------------------------------------------- */

int h264_ctx_new(struct h264_ctx **ret_obj)
{
	if (ret_obj == NULL)
		return -EINVAL;
	*ret_obj = calloc(1, sizeof(struct h264_ctx));
	return *ret_obj == NULL ? -ENOMEM : 0;
}

void h264_ctx_drop_sei(struct h264_ctx *ctx)
{
	for (uint32_t i = 0; i < ctx->sei_count; i++)
		free(ctx->sei_tab[i].raw.buf);
	free(ctx->sei_tab);
	ctx->sei_tab = NULL;
	ctx->sei_count = 0;
}

static void clear_slice(struct h264_ctx *ctx)
{
	ctx->slice_type = 0;
	memset(&ctx->sh, 0, sizeof(ctx->sh));
	memset(&ctx->rawdata, 0, sizeof(ctx->rawdata));
	derive_slice_vars(ctx);
}

/* between NAL units: everything about the previous one goes, the parameter sets and
 * the "previous NAL was VCL / filler" memory stay (src/h264_ctx.c:267-281) */
int h264_ctx_clear_nalu(struct h264_ctx *ctx)
{
	if (ctx == NULL)
		return -EINVAL;
	ctx->nalu_type = H264_NALU_TYPE_UNKNOWN;
	memset(&ctx->nalu_hdr, 0, sizeof(ctx->nalu_hdr));
	ctx->nalu_unknown = 0;
	ctx->first_vcl = 0;
	memset(&ctx->aud, 0, sizeof(ctx->aud));
	h264_ctx_drop_sei(ctx);
	clear_slice(ctx);
	return 0;
}

int h264_ctx_clear(struct h264_ctx *ctx)
{
	if (ctx == NULL)
		return -EINVAL;
	h264_ctx_drop_sei(ctx);
	for (size_t i = 0; i < H264_SPS_MAX; i++)
		free(ctx->sps_tab[i]);
	for (size_t i = 0; i < H264_PPS_MAX; i++)
		free(ctx->pps_tab[i]);
	memset(ctx, 0, sizeof(*ctx));
	return 0;
}

int h264_ctx_destroy(struct h264_ctx *ctx)
{
	if (ctx != NULL) {
		h264_ctx_clear(ctx);
		free(ctx);
	}
	return 0;
}

/* ---- setters / getters --------------------------------------------------------------- */

int h264_ctx_set_nalu_header(struct h264_ctx *ctx, const struct h264_nalu_header *nh)
{
	if (ctx == NULL || nh == NULL)
		return -EINVAL;
	ctx->nalu_hdr = *nh;
	ctx->nalu_type = (enum h264_nalu_type)nh->nal_unit_type;
	return 0;
}

int h264_ctx_is_nalu_unknown(struct h264_ctx *ctx)
{
	return ctx != NULL && ctx->nalu_unknown;
}

int h264_ctx_set_aud(struct h264_ctx *ctx, const struct h264_aud *aud)
{
	if (ctx == NULL || aud == NULL)
		return -EINVAL;
	ctx->aud = *aud;
	return 0;
}

int h264_ctx_set_sps(struct h264_ctx *ctx, const struct h264_sps *sps)
{
	if (ctx == NULL || sps == NULL || sps->seq_parameter_set_id >= H264_SPS_MAX)
		return -EINVAL;
	struct h264_sps **slot = &ctx->sps_tab[sps->seq_parameter_set_id];
	if (*slot == NULL && (*slot = calloc(1, sizeof(**slot))) == NULL)
		return -ENOMEM;
	**slot = *sps;
	ctx->sps = *slot;
	derive_after_sps(ctx);
	return 0;
}

int h264_ctx_set_pps(struct h264_ctx *ctx, const struct h264_pps *pps)
{
	if (ctx == NULL || pps == NULL || pps->pic_parameter_set_id >= H264_PPS_MAX)
		return -EINVAL;
	struct h264_pps **slot = &ctx->pps_tab[pps->pic_parameter_set_id];
	if (*slot == NULL && (*slot = calloc(1, sizeof(**slot))) == NULL)
		return -ENOMEM;
	**slot = *pps;
	ctx->pps = *slot;
	ctx->SliceGroupChangeRate = pps->slice_group_change_rate_minus1 + 1;
	derive_slice_vars(ctx);
	return 0;
}

int h264_ctx_set_filler(struct h264_ctx *ctx, size_t len)
{
	if (ctx == NULL)
		return -EINVAL;
	ctx->filler_len = len;
	return 0;
}

int h264_ctx_activate_sps(struct h264_ctx *ctx, uint32_t id)
{
	if (ctx == NULL || id >= H264_SPS_MAX || ctx->sps_tab[id] == NULL)
		return -EINVAL;
	ctx->sps = ctx->sps_tab[id];
	derive_after_sps(ctx);
	return 0;
}

int h264_ctx_activate_pps(struct h264_ctx *ctx, uint32_t id)
{
	if (ctx == NULL || id >= H264_PPS_MAX || ctx->pps_tab[id] == NULL)
		return -EINVAL;
	ctx->pps = ctx->pps_tab[id];
	ctx->SliceGroupChangeRate = ctx->pps->slice_group_change_rate_minus1 + 1;
	return h264_ctx_activate_sps(ctx, ctx->pps->seq_parameter_set_id);
}

const struct h264_sps *h264_ctx_get_sps(struct h264_ctx *ctx)
{
	return ctx == NULL ? NULL : ctx->sps;
}

const struct h264_pps *h264_ctx_get_pps(struct h264_ctx *ctx)
{
	return ctx == NULL ? NULL : ctx->pps;
}

/* ---- SEI table -------------------------------------------------------------------------- */

int h264_ctx_new_sei(struct h264_ctx *ctx, struct h264_sei **out)
{
	if (out == NULL)
		return -EINVAL;
	*out = NULL;
	if (ctx == NULL)
		return -EINVAL;
	struct h264_sei *t = realloc(ctx->sei_tab, (ctx->sei_count + 1) * sizeof(*t));
	if (t == NULL)
		return -ENOMEM;
	ctx->sei_tab = t;
	*out = &t[ctx->sei_count++];
	memset(*out, 0, sizeof(**out));
	return 0;
}

/* the opaque-bytes members of a decoded SEI point into its own raw payload */
int h264_sei_fix_pointers(struct h264_sei *sei)
{
	if (sei == NULL || sei->raw.buf == NULL || sei->raw.len == 0)
		return -EINVAL;
	const uint8_t **buf = NULL;
	size_t *len = NULL, skip = 0;
	switch (sei->type) {
	case H264_SEI_TYPE_FILLER_PAYLOAD:
		buf = &sei->filler_payload.buf;
		len = &sei->filler_payload.len;
		break;
	case H264_SEI_TYPE_USER_DATA_REGISTERED:
		skip = sei->user_data_registered.country_code == 0xff ? 2 : 1;
		buf = &sei->user_data_registered.buf;
		len = &sei->user_data_registered.len;
		break;
	case H264_SEI_TYPE_USER_DATA_UNREGISTERED:
		skip = 16;
		buf = &sei->user_data_unregistered.buf;
		len = &sei->user_data_unregistered.len;
		break;
	default:
		return 0;
	}
	if (sei->raw.len < skip)
		return -EINVAL;
	*buf = sei->raw.buf + skip;
	*len = sei->raw.len - skip;
	return 0;
}

/* writer side: append a SEI message, serialising its payload now (no escaping) */
int h264_ctx_add_sei(struct h264_ctx *ctx, const struct h264_sei *sei)
{
	if (ctx == NULL || sei == NULL)
		return -EINVAL;
	struct h264_bitstream bs;
	struct h264_sei *copy = NULL;
	h264_bs_init(&bs, NULL, 0, 0);
	int res = h264_ctx_new_sei(ctx, &copy);
	if (res < 0)
		goto fail;
	*copy = *sei;
	copy->raw.buf = NULL;
	copy->raw.len = 0;
	struct h264_io io = {H264_IO_WRITE, &bs, ctx, NULL, NULL, NULL};
	res = h264_syntax_sei_payload(&io, copy);
	if (res >= 0)
		res = h264_bs_acquire_buf(&bs, &copy->raw.buf, &copy->raw.len);
	if (res >= 0)
		res = h264_sei_fix_pointers(copy);
	if (res >= 0) {
		h264_bs_clear(&bs);
		return 0;
	}
fail:
	if (copy != NULL) {
		free(copy->raw.buf);
		ctx->sei_count--;
	}
	h264_bs_clear(&bs);
	return res;
}

int h264_ctx_get_sei_count(struct h264_ctx *ctx)
{
	return ctx == NULL ? -EINVAL : (int)ctx->sei_count;
}

/* D.2.2: clockTimestamp of the first clock timestamp set, in time_scale units */
uint64_t h264_ctx_sei_pic_timing_to_ts(struct h264_ctx *ctx, const struct h264_sei_pic_timing *sei)
{
	if (ctx == NULL || sei == NULL || ctx->sps == NULL)
		return 0;
	const struct h264_vui *v = &ctx->sps->vui;
	if (v->time_scale == 0 || v->num_units_in_tick == 0)
		return 0;
	const uint64_t secs = ((uint64_t)sei->clk_ts[0].hours_value * 60 + sei->clk_ts[0].minutes_value) * 60 +
			      sei->clk_ts[0].seconds_value;
	uint64_t ts = secs * v->time_scale +
		      (uint64_t)sei->clk_ts[0].n_frames *
			      ((uint64_t)v->num_units_in_tick * (1 + (uint64_t)sei->clk_ts[0].nuit_field_based_flag));
	const int32_t off = sei->clk_ts[0].time_offset;
	if (off < 0 && (uint64_t)(-(int64_t)off) > ts)
		return 0;
	return ts + (uint64_t)(int64_t)off;
}

uint64_t h264_ctx_sei_pic_timing_to_us(struct h264_ctx *ctx, const struct h264_sei_pic_timing *sei)
{
	if (ctx == NULL || sei == NULL || ctx->sps == NULL || ctx->sps->vui.time_scale == 0)
		return 0;
	const uint64_t scale = ctx->sps->vui.time_scale;
	return (h264_ctx_sei_pic_timing_to_ts(ctx, sei) * 1000000 + scale / 2) / scale;
}

/* ---- slice header + first VCL NAL unit of a primary coded picture (7.4.1.2.4) ------------ */

static int starts_new_picture(const struct h264_ctx *ctx)
{
	const struct h264_sps *sps = ctx->sps;
	const struct h264_slice_header *a = &ctx->sh, *b = &ctx->prev_slice_hdr;
	const struct h264_nalu_header *na = &ctx->nalu_hdr, *nb = &ctx->prev_slice_nalu_hdr;
	const int idr_a = na->nal_unit_type == H264_NALU_TYPE_SLICE_IDR;
	const int idr_b = nb->nal_unit_type == H264_NALU_TYPE_SLICE_IDR;
	const int fields = sps != NULL && !sps->frame_mbs_only_flag;
	const uint32_t poc = sps != NULL ? sps->pic_order_cnt_type : 2;
	if (!ctx->prev_vcl && !ctx->prev_filler)
		return 1;
	return a->frame_num != b->frame_num || a->pic_parameter_set_id != b->pic_parameter_set_id ||
	       (fields && a->field_pic_flag != b->field_pic_flag) ||
	       (fields && a->field_pic_flag && b->field_pic_flag &&
		a->bottom_field_flag != b->bottom_field_flag) ||
	       (!na->nal_ref_idc != !nb->nal_ref_idc) ||
	       (poc == 0 && (a->pic_order_cnt_lsb != b->pic_order_cnt_lsb ||
			     a->delta_pic_order_cnt_bottom != b->delta_pic_order_cnt_bottom)) ||
	       (poc == 1 && (a->delta_pic_order_cnt[0] != b->delta_pic_order_cnt[0] ||
			     a->delta_pic_order_cnt[1] != b->delta_pic_order_cnt[1])) ||
	       idr_a != idr_b || (idr_a && idr_b && a->idr_pic_id != b->idr_pic_id);
}

int h264_ctx_set_slice_header(struct h264_ctx *ctx, const struct h264_slice_header *sh)
{
	if (ctx == NULL || sh == NULL)
		return -EINVAL;
	clear_slice(ctx);
	ctx->slice_type = H264_SLICE_TYPE(sh->slice_type);
	ctx->sh = *sh;
	derive_slice_vars(ctx);
	ctx->first_vcl = starts_new_picture(ctx);
	ctx->prev_slice_nalu_hdr = ctx->nalu_hdr;
	ctx->prev_slice_hdr = ctx->sh;
	return 0;
}

/* ---- stream info (src/h264_ctx.c:576-690) ----------------------------------------------- */

static const uint16_t sar_table[17][2] = {{1, 1},   {1, 1},   {12, 11}, {10, 11}, {16, 11}, {40, 33},
					  {24, 11}, {20, 11}, {32, 11}, {80, 33}, {18, 11}, {15, 11},
					  {64, 33}, {160, 99}, {4, 3},  {3, 2},   {2, 1}};

static void fill_info(const struct h264_sps *sps, const struct h264_sps_derived *d, struct h264_info *info)
{
	memset(info, 0, sizeof(*info));
	info->width = d->PicWidthInSamplesLuma;
	info->height = d->FrameHeightInMbs * 16;
	info->bit_depth_luma = (uint8_t)d->BitDepthLuma;
	info->crop_width = info->width;
	info->crop_height = info->height;
	if (sps->frame_cropping_flag) {
		info->crop_left = sps->frame_crop_left_offset * d->CropUnitX;
		info->crop_width = info->width - sps->frame_crop_right_offset * d->CropUnitX;
		info->crop_top = sps->frame_crop_top_offset * d->CropUnitY;
		info->crop_height = info->height - sps->frame_crop_bottom_offset * d->CropUnitY;
	}
	info->sar_width = info->sar_height = 1;
	if (!sps->vui_parameters_present_flag)
		return;
	const struct h264_vui *v = &sps->vui;
	if (v->aspect_ratio_info_present_flag) {
		if (v->aspect_ratio_idc == H264_ASPECT_RATIO_EXTENDED_SAR) {
			info->sar_width = v->sar_width;
			info->sar_height = v->sar_height;
		} else if (v->aspect_ratio_idc <= 16) {
			info->sar_width = sar_table[v->aspect_ratio_idc][0];
			info->sar_height = sar_table[v->aspect_ratio_idc][1];
		}
	}
	info->full_range = v->video_full_range_flag;
	info->colour_description_present = v->colour_description_present_flag ? 1 : 0;
	info->colour_primaries = v->colour_description_present_flag ? v->colour_primaries : 2;
	info->transfer_characteristics = v->colour_description_present_flag ? v->transfer_characteristics : 2;
	info->matrix_coefficients = v->colour_description_present_flag ? v->matrix_coefficients : 2;
	if (v->timing_info_present_flag) {
		info->num_units_in_tick = v->num_units_in_tick;
		info->time_scale = v->time_scale;
		info->framerate = v->num_units_in_tick ? (float)v->time_scale / 2.f / v->num_units_in_tick : 0.f;
		info->framerate_num = v->time_scale;
		info->framerate_den = v->num_units_in_tick;
		if (info->framerate_num % 2 == 0)
			info->framerate_num /= 2;
		else
			info->framerate_den *= 2;
	}
	if (v->nal_hrd_parameters_present_flag) {
		info->nal_hrd_bitrate = (v->nal_hrd.cpb[0].bit_rate_value_minus1 + 1) << (6 + v->nal_hrd.bit_rate_scale);
		info->nal_hrd_cpb_size = (v->nal_hrd.cpb[0].cpb_size_value_minus1 + 1) << (4 + v->nal_hrd.cpb_size_scale);
	}
	if (v->vcl_hrd_parameters_present_flag) {
		info->vcl_hrd_bitrate = (v->vcl_hrd.cpb[0].bit_rate_value_minus1 + 1) << (6 + v->vcl_hrd.bit_rate_scale);
		info->vcl_hrd_cpb_size = (v->vcl_hrd.cpb[0].cpb_size_value_minus1 + 1) << (4 + v->vcl_hrd.cpb_size_scale);
	}
}

int h264_ctx_get_info(struct h264_ctx *ctx, struct h264_info *info)
{
	if (ctx == NULL || info == NULL)
		return -EINVAL;
	if (ctx->sps == NULL || ctx->pps == NULL)
		return -EAGAIN;
	fill_info(ctx->sps, &ctx->spsd, info);
	return 0;
}

int h264_get_info(const uint8_t *sps, size_t sps_len, const uint8_t *pps, size_t pps_len,
		  struct h264_info *info)
{
	if (sps == NULL || pps == NULL || info == NULL)
		return -EINVAL;
	struct h264_sps *s = calloc(1, sizeof(*s));
	struct h264_pps *p = calloc(1, sizeof(*p));
	struct h264_sps_derived d;
	int res = s != NULL && p != NULL ? 0 : -ENOMEM;
	if (res >= 0)
		res = h264_parse_sps(sps, sps_len, s);
	if (res >= 0)
		res = h264_parse_pps(pps, pps_len, s, p);
	if (res >= 0)
		res = h264_get_sps_derived(s, &d);
	if (res >= 0)
		fill_info(s, &d, info);
	free(s);
	free(p);
	return res;
}

int h264_sar_to_aspect_ratio_idc(unsigned int sar_width, unsigned int sar_height)
{
	for (unsigned int i = 1; i < COUNT_OF(sar_table); i++)
		if (sar_table[i][0] == sar_width && sar_table[i][1] == sar_height)
			return (int)i;
	return H264_ASPECT_RATIO_EXTENDED_SAR;
}
