/*
 * h264_fmo.c — row A12: the macroblock -> slice group map of the current slice (8.2.2), from the
 * active PPS, the slice header and the derived variables of the context.  The reference builds
 * it inside its slice_data walk (h264_gen_slice_group_map + h264_mb_to_slice_group,
 * src/h264_fmo.c:217-291) and steps through it with h264_next_mb_addr (:308-319); here the
 * finished map goes to the slice kernel with the slice's parameter block.
 */
#include "h264_priv.h"
#include "fmo_map.h"

int h264_ctx_get_slice_group_map(const struct h264_ctx *ctx, uint8_t *map, size_t cap)
{
	if (ctx == NULL || ctx->sps == NULL || ctx->pps == NULL)
		return -EINVAL;
	const struct h264_pps *pps = ctx->pps;
	const uint32_t n = ctx->PicSizeInMbs;
	if (map == NULL || cap < n)
		return -ENOBUFS;
	if (pps->num_slice_groups_minus1 == 0) {
		memset(map, 0, n);
		return (int)n;
	}
	if (pps->num_slice_groups_minus1 > 7)
		return -EIO;
	struct fmo_desc d;
	memset(&d, 0, sizeof(d));
	d.num_slice_groups_minus1 = pps->num_slice_groups_minus1;
	d.map_type = pps->slice_group_map_type;
	d.run_length_minus1 = pps->run_length_minus1;
	d.top_left = pps->top_left;
	d.bottom_right = pps->bottom_right;
	d.change_direction_flag = pps->slice_group_change_direction_flag;
	d.map_units_in_slice_group0 = ctx->MapUnitsInSliceGroup0;
	d.slice_group_id = pps->slice_group_id;
	d.n_slice_group_id = pps->pic_size_in_map_units_minus1 + 1 < COUNT_OF(pps->slice_group_id)
				     ? pps->pic_size_in_map_units_minus1 + 1
				     : (uint32_t)COUNT_OF(pps->slice_group_id);
	d.pic_width_in_mbs = ctx->spsd.PicWidthInMbs;
	d.pic_height_in_map_units = ctx->spsd.PicHeightInMapUnits;
	d.frame_mbs_only_flag = ctx->sps->frame_mbs_only_flag;
	d.field_pic_flag = ctx->sh.field_pic_flag;
	d.mbaff_frame_flag = ctx->MbaffFrameFlag;
	d.pic_size_in_mbs = n;
	const size_t units = (size_t)d.pic_width_in_mbs * d.pic_height_in_map_units;
	uint8_t *u = malloc(units ? units : 1);
	if (u == NULL)
		return -ENOMEM;
	int r = fmo_map_units(&d, u);
	if (r == 0)
		fmo_mb_map(&d, u, map);
	free(u);
	return r < 0 ? -EIO : (int)n;
}
