/*
 * fmo_map.h — H.264 8.2.2: macroblock to slice group map (flexible macroblock ordering).
 * Plain C, header only: used by the host library (row A12: what the reference computes in
 * h264_gen_slice_group_map + h264_mb_to_slice_group, src/h264_fmo.c:30-291, before it walks a
 * slice's macroblocks with h264_next_mb_addr, :308-319) and by the synthetic stream generator.
 * The slice kernels get the finished macroblock -> group map, one byte per macroblock.
 */
#ifndef FMO_MAP_H
#define FMO_MAP_H

#include <stdint.h>
#include <string.h>

struct fmo_desc {
	uint32_t num_slice_groups_minus1; /* 1..7 */
	uint32_t map_type;                /* 0..6 */
	const uint32_t *run_length_minus1; /* type 0: [num_slice_groups_minus1 + 1] */
	const uint32_t *top_left;          /* type 2: [num_slice_groups_minus1] */
	const uint32_t *bottom_right;
	int change_direction_flag;         /* types 3..5 */
	uint32_t map_units_in_slice_group0; /* types 3..5: Min(cycle * rate, PicSizeInMapUnits) */
	const uint32_t *slice_group_id;    /* type 6: [n_slice_group_id] */
	uint32_t n_slice_group_id;
	uint32_t pic_width_in_mbs, pic_height_in_map_units;
	/* 8.2.2.8: how map units relate to macroblocks */
	int frame_mbs_only_flag, field_pic_flag, mbaff_frame_flag;
	uint32_t pic_size_in_mbs;
};

/* 8.2.2.1 - 8.2.2.7: map unit -> slice group; units = pic_width_in_mbs * pic_height_in_map_units */
static inline int fmo_map_units(const struct fmo_desc *d, uint8_t *map)
{
	const uint32_t W = d->pic_width_in_mbs, H = d->pic_height_in_map_units, n = W * H;
	const uint32_t groups = d->num_slice_groups_minus1 + 1;
	const int dir = d->change_direction_flag ? 1 : 0;
	const uint32_t upper_left = dir ? n - d->map_units_in_slice_group0 : d->map_units_in_slice_group0;
	uint32_t i, k;
	if (n == 0 || groups < 2 || groups > 8)
		return -1;
	switch (d->map_type) {
	case 0: /* interleaved: runs of each group in turn */
		for (i = 0; i < n;) {
			uint32_t g;
			for (g = 0; g < groups && i < n; g++) {
				const uint32_t run = d->run_length_minus1[g] + 1;
				for (k = 0; k < run && i + k < n; k++)
					map[i + k] = (uint8_t)g;
				i += run;
			}
		}
		break;
	case 1: /* dispersed */
		for (i = 0; i < n; i++)
			map[i] = (uint8_t)(((i % W) + (((i / W) * groups) / 2)) % groups);
		break;
	case 2: { /* foreground rectangles, lowest group on top, left-over = last group */
		int g;
		memset(map, (int)d->num_slice_groups_minus1, n);
		for (g = (int)d->num_slice_groups_minus1 - 1; g >= 0; g--) {
			const uint32_t y0 = d->top_left[g] / W, x0 = d->top_left[g] % W;
			const uint32_t y1 = d->bottom_right[g] / W, x1 = d->bottom_right[g] % W;
			uint32_t x, y;
			for (y = y0; y <= y1 && y < H; y++)
				for (x = x0; x <= x1 && x < W; x++)
					map[y * W + x] = (uint8_t)g;
		}
		break;
	}
	case 3: { /* box-out: group 0 grows in a spiral from the centre */
		int32_t x = (int32_t)((W - (uint32_t)dir) / 2), y = (int32_t)((H - (uint32_t)dir) / 2);
		int32_t left = x, top = y, right = x, bottom = y;
		int32_t xdir = dir - 1, ydir = dir;
		uint32_t vacant;
		memset(map, 1, n);
		for (k = 0; k < d->map_units_in_slice_group0; k += vacant) {
			uint8_t *u = &map[(uint32_t)y * W + (uint32_t)x];
			vacant = *u == 1;
			if (vacant)
				*u = 0;
			if (xdir == -1 && x == left) {
				left = left > 0 ? left - 1 : 0;
				x = left;
				xdir = 0;
				ydir = 2 * dir - 1;
			} else if (xdir == 1 && x == right) {
				right = right + 1 < (int32_t)W ? right + 1 : (int32_t)W - 1;
				x = right;
				xdir = 0;
				ydir = 1 - 2 * dir;
			} else if (ydir == -1 && y == top) {
				top = top > 0 ? top - 1 : 0;
				y = top;
				xdir = 1 - 2 * dir;
				ydir = 0;
			} else if (ydir == 1 && y == bottom) {
				bottom = bottom + 1 < (int32_t)H ? bottom + 1 : (int32_t)H - 1;
				y = bottom;
				xdir = 2 * dir - 1;
				ydir = 0;
			} else {
				x += xdir;
				y += ydir;
			}
		}
		break;
	}
	case 4: /* raster scan */
		for (i = 0; i < n; i++)
			map[i] = (uint8_t)(i < upper_left ? dir : 1 - dir);
		break;
	case 5: /* wipe: column by column */
		k = 0;
		for (i = 0; i < W; i++) {
			uint32_t j;
			for (j = 0; j < H; j++)
				map[j * W + i] = (uint8_t)(k++ < upper_left ? dir : 1 - dir);
		}
		break;
	case 6: /* explicit */
		for (i = 0; i < n; i++)
			map[i] = (uint8_t)(i < d->n_slice_group_id ? d->slice_group_id[i] : 0);
		break;
	default:
		return -1;
	}
	return 0;
}

/* 8.2.2.8: macroblock -> slice group from the map unit map */
static inline void fmo_mb_map(const struct fmo_desc *d, const uint8_t *units, uint8_t *mb_map)
{
	const uint32_t W = d->pic_width_in_mbs;
	uint32_t i;
	for (i = 0; i < d->pic_size_in_mbs; i++) {
		if (d->frame_mbs_only_flag || d->field_pic_flag)
			mb_map[i] = units[i];
		else if (d->mbaff_frame_flag)
			mb_map[i] = units[i / 2];
		else
			mb_map[i] = units[(i / (2 * W)) * W + (i % W)];
	}
}

/* 8.2.2 (the reference's h264_next_mb_addr, src/h264_fmo.c:308-319) */
static inline uint32_t fmo_next_mb_addr(const uint8_t *mb_map, uint32_t pic_size_in_mbs, uint32_t mb)
{
	uint32_t i = mb + 1;
	while (i < pic_size_in_mbs && mb_map[i] != mb_map[mb])
		i++;
	return i;
}

#endif /* FMO_MAP_H */
