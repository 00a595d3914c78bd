/*
 * synth_video.c — synthetic H.264 elementary stream generator (host C): SPS, PPS
 * and slices whose slice_data() is seeded-random but VALID CAVLC syntax (random
 * macroblock types, prediction data, coded block patterns and residual blocks,
 * entropy-coded with the CAVLC tables).  There is no encoder, ffmpeg or sample
 * stream in the build environment, so this is what BASELINE.json's configs 1/3/4
 * ("synthetic 1080p Baseline CAVLC stream", many-slice streams) are made of.
 *
 * Workload generation only — not part of the parse path.  Its validity is checked
 * by the reference itself (tests/test_cavlc.py: the reference reader consumes
 * every macroblock and ends each slice exactly at its stop bit).
 *
 * Syntax written follows Rec. ITU-T H.264 7.3.2.1/7.3.2.2/7.3.3/7.3.4/7.3.5 in
 * the subset the reader under test supports (frame pictures, one slice group).
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define CAVLC_TAB static const
#include "cavlc_luts.h"
#include "fmo_map.h"
#include "mbaff_nb.h"

struct synth_video_cfg {
	uint32_t width_mbs, height_mbs;
	uint32_t frames;
	uint32_t slices_per_frame;
	uint32_t profile_idc;       /* 66, 77, 100 ... */
	uint32_t chroma_format_idc; /* 0..3 (only sent when profile_idc >= 100) */
	uint32_t transform_8x8;     /* PPS transform_8x8_mode_flag */
	uint32_t idr_period;        /* IDR every n frames */
	uint32_t b_frames;          /* 1: odd non-IDR frames are B slices */
	uint32_t num_ref_frames;    /* num_ref_idx_lX_default_active = this (1..4) */
	uint32_t entropy_cabac;     /* 1: slice data comes from synth_cabac.cpp (CABAC) */
	uint32_t pct_skip;          /* % of P/B macroblocks skipped */
	uint32_t pct_intra_in_inter;/* % of coded P/B macroblocks that are intra */
	uint32_t pct_pcm;           /* % of intra macroblocks that are I_PCM (x0.1) */
	uint32_t coef_density;      /* 0..100: how many residual blocks carry coefficients */
	uint64_t seed;
	uint32_t cabac_twin;        /* CAVLC only: restrict to syntax CABAC can carry too (no
				       P_8x8ref0, 8x8-transform blocks with a set cbp bit are
				       never empty), for the CAVLC <-> CABAC twin test */
	uint32_t fmo;               /* CAVLC only: slice groups (8.2.2): number of groups (2..8, 0/1 =
				       none) | slice_group_map_type << 4 (0..6; 6 needs W * H <= 256);
				       every group of a frame is cut into slices_per_frame slices
				       | 0x100: MBAFF frames (frame_mbs_only_flag = 0,
				       mb_adaptive_frame_field_flag = 1, even height, no slice groups):
				       macroblock pairs, a random mb_field_decoding_flag per pair
				       | 0x200: field pictures (frame_mbs_only_flag = 0, every picture a
				       field of height_mbs / 2 macroblock rows, field_pic_flag = 1) */
};

/* CABAC slice data (synth_cabac.cpp): RBSP bytes of slice_data() incl. the stop bit */
struct synth_cabac_slice_cfg {
	uint32_t width_mbs, height_mbs, first_mb, count;
	uint32_t slice_type, chroma_array_type, transform_8x8, direct_8x8_inference;
	uint32_t num_ref_l0, num_ref_l1, cabac_init_idc;
	int32_t slice_qp;
	uint32_t pct_skip, pct_intra_in_inter, pct_pcm, coef_density;
};
uint64_t synth_cabac_slice(const struct synth_cabac_slice_cfg *cfg, uint64_t *rng_state, uint8_t *out,
			   uint64_t cap);

/* ---- random ---------------------------------------------------------------- */
struct rng {
	uint64_t s;
};
static inline uint32_t rnd(struct rng *r)
{
	uint64_t x = r->s;
	x ^= x << 13;
	x ^= x >> 7;
	x ^= x << 17;
	r->s = x;
	return (uint32_t)(x >> 24);
}
static inline uint32_t rnd_n(struct rng *r, uint32_t n)
{
	return n ? rnd(r) % n : 0;
}
static inline int rnd_pct(struct rng *r, uint32_t pct)
{
	return rnd_n(r, 100) < pct;
}

/* ---- bit writer with emulation prevention ------------------------------------ */
struct bw {
	uint8_t *buf;
	size_t cap, len;
	uint32_t acc;
	int nacc;  /* bits pending in acc (0..7) */
	int zeros; /* trailing zero bytes of the current NAL's output */
	int overflow;
	uint32_t *mark; /* pending raw-bit-offset mark inside the byte being assembled */
};

static void bw_raw(struct bw *w, uint8_t c)
{
	if (w->len < w->cap)
		w->buf[w->len] = c;
	else
		w->overflow = 1;
	w->len++;
}

static void bw_byte(struct bw *w, uint8_t c)
{
	if (w->zeros == 2 && c <= 3) {
		bw_raw(w, 3);
		w->zeros = 0;
		if (w->mark) /* the marked bits moved one byte further in the raw stream */
			*w->mark += 8;
	}
	w->mark = NULL;
	bw_raw(w, c);
	w->zeros = c == 0 ? w->zeros + 1 : 0;
}

static void bw_bits(struct bw *w, uint32_t v, int n)
{
	for (int i = n - 1; i >= 0; i--) {
		w->acc = (w->acc << 1) | ((v >> i) & 1);
		if (++w->nacc == 8) {
			bw_byte(w, (uint8_t)w->acc);
			w->acc = 0;
			w->nacc = 0;
		}
	}
}

static void bw_ue(struct bw *w, uint32_t v)
{
	uint32_t x = v + 1;
	int n = 0;
	while ((x >> n) > 1)
		n++;
	bw_bits(w, 0, n);
	bw_bits(w, x, n + 1);
}

static void bw_se(struct bw *w, int32_t v)
{
	bw_ue(w, v <= 0 ? (uint32_t)(-2 * v) : (uint32_t)(2 * v - 1));
}

static void bw_te(struct bw *w, uint32_t v, uint32_t max)
{
	if (max == 1)
		bw_bits(w, !v, 1);
	else
		bw_ue(w, v);
}

static void bw_trailing(struct bw *w)
{
	bw_bits(w, 1, 1);
	while (w->nacc)
		bw_bits(w, 0, 1);
}

static void bw_start_nal(struct bw *w, uint8_t hdr)
{
	bw_raw(w, 0);
	bw_raw(w, 0);
	bw_raw(w, 0);
	bw_raw(w, 1);
	w->zeros = 0;
	w->acc = 0;
	w->nacc = 0;
	bw_byte(w, hdr);
}

/* ---- CAVLC encode tables, inverted from the decode LUTs ----------------------- */
struct enc {
	uint8_t len;
	uint16_t code;
};
static struct enc e_coeff_token[5][128]; /* index t1 << 5 | tc */
static struct enc e_tz4x4[16][16], e_tz2x2[4][4], e_tz2x4[8][8], e_run[8][16];
static uint8_t e_cbp_chroma[2][48], e_cbp_nochroma[2][16]; /* [intra=0/inter=1][cbp] -> codeNum */
static int tables_ready;

static void invert(const uint16_t *lut, int K, struct enc *out, int nvals)
{
	for (int i = 0; i < nvals; i++)
		out[i].len = 0;
	for (int idx = 0; idx < (16 << K); idx++) {
		uint16_t e = lut[idx];
		if (!e)
			continue;
		int len = e & 0xff, val = e >> 8, lz = idx >> K;
		if (val >= nvals || out[val].len)
			continue;
		if (lz >= len) { /* all-zero codeword */
			out[val].len = (uint8_t)len;
			out[val].code = 0;
			continue;
		}
		int rest = len - lz - 1;
		uint32_t restbits = ((uint32_t)idx & ((1u << K) - 1)) >> (K - rest);
		out[val].len = (uint8_t)len;
		out[val].code = (uint16_t)((1u << rest) | restbits);
	}
}

static void build_tables(void)
{
	if (tables_ready)
		return;
	for (int t = 0; t < 5; t++)
		invert(cavlc_coeff_token[t], CAVLC_COEFF_TOKEN_K, e_coeff_token[t], 128);
	for (int t = 0; t < 16; t++)
		invert(cavlc_total_zeros_4x4[t], CAVLC_TOTAL_ZEROS_4X4_K, e_tz4x4[t], 16);
	for (int t = 0; t < 4; t++)
		invert(cavlc_total_zeros_2x2[t], CAVLC_TOTAL_ZEROS_2X2_K, e_tz2x2[t], 4);
	for (int t = 0; t < 8; t++)
		invert(cavlc_total_zeros_2x4[t], CAVLC_TOTAL_ZEROS_2X4_K, e_tz2x4[t], 8);
	for (int t = 0; t < 8; t++)
		invert(cavlc_run_before[t], CAVLC_RUN_BEFORE_K, e_run[t], 16);
	for (int c = 0; c < 48; c++) {
		e_cbp_chroma[0][cavlc_cbp_chroma[c][0]] = (uint8_t)c;
		e_cbp_chroma[1][cavlc_cbp_chroma[c][1]] = (uint8_t)c;
	}
	for (int c = 0; c < 16; c++) {
		e_cbp_nochroma[0][cavlc_cbp_nochroma[c][0]] = (uint8_t)c;
		e_cbp_nochroma[1][cavlc_cbp_nochroma[c][1]] = (uint8_t)c;
	}
	tables_ready = 1;
}

/* ---- per-slice macroblock state (what nC derivation needs) -------------------- */
struct mbstate {
	uint8_t nz[48];
	uint8_t avail;
};

struct gen {
	struct synth_video_cfg cfg;
	struct rng r;
	struct bw w;
	struct mbstate *mbs; /* PicSizeInMbs entries, reset per slice */
	uint32_t first_mb, cur_mb;
	uint32_t cat;        /* ChromaArrayType */
	uint32_t mbw_c, mbh_c;
	int slice_type;      /* 0 P, 1 B, 2 I */
	int cur_t8;          /* current macroblock uses the 8x8 transform */
	uint8_t *cabac_buf;  /* scratch for one slice's CABAC RBSP */
	uint64_t cabac_cap;
	uint64_t n_mbs;
	/* optional: the parameter block of every slice (include/h264gpu_slice.h layout) */
	struct slice_params_out *params;
	uint64_t params_cap, params_n;
	/* MBAFF / field pictures */
	int mbaff, cur_field, paff;
	uint32_t pic_h; /* PicHeightInMbs */
	uint8_t *pair_field;  /* mb_field_decoding_flag of every macroblock pair, reset per slice */
	int ref_present;      /* ref_idx is sent for this macroblock */
	uint32_t ref_range;   /* number of values te() of ref_idx covers */
	/* slice groups */
	uint32_t fmo_groups, fmo_type, fmo_rate, fmo_cycle_bits, fmo_cycle;
	int fmo_dir;
	uint8_t *fmo_units, *fmo_map; /* map unit / macroblock -> slice group of the current picture */
	uint32_t fmo_run[8], fmo_tl[8], fmo_br[8], fmo_ids[256];
};

/* the macroblock -> slice group map of the current picture (types 3..5 move with fmo_cycle) */
static void fmo_picture_map(struct gen *g)
{
	struct fmo_desc d;
	const uint32_t W = g->cfg.width_mbs, H = g->cfg.height_mbs, N = W * H;
	memset(&d, 0, sizeof(d));
	d.num_slice_groups_minus1 = g->fmo_groups - 1;
	d.map_type = g->fmo_type;
	d.run_length_minus1 = g->fmo_run;
	d.top_left = g->fmo_tl;
	d.bottom_right = g->fmo_br;
	d.change_direction_flag = g->fmo_dir;
	const uint64_t units = (uint64_t)g->fmo_cycle * g->fmo_rate;
	d.map_units_in_slice_group0 = units < N ? (uint32_t)units : N;
	d.slice_group_id = g->fmo_ids;
	d.n_slice_group_id = N <= 256 ? N : 256;
	d.pic_width_in_mbs = W;
	d.pic_height_in_map_units = H;
	d.frame_mbs_only_flag = 1;
	d.pic_size_in_mbs = N;
	fmo_map_units(&d, g->fmo_units);
	fmo_mb_map(&d, g->fmo_units, g->fmo_map);
}

static uint32_t gen_next_mb(const struct gen *g, uint32_t a)
{
	return g->fmo_groups >= 2 ? fmo_next_mb_addr(g->fmo_map, g->cfg.width_mbs * g->cfg.height_mbs, a) : a + 1;
}

/* mirror of struct h264gpu_slice_params (include/h264gpu_slice.h), kept local so the
 * generator has no dependency on the GPU library */
struct slice_params_out {
	uint64_t nal_off;
	uint32_t nal_len, data_bit_off, first_mb_in_slice, mb_out_off, mb_out_cap, row_state_off;
	uint16_t pic_width_in_mbs, pic_height_in_mbs;
	uint8_t slice_type, chroma_array_type, bit_depth_luma, bit_depth_chroma;
	uint8_t transform_8x8_mode_flag, direct_8x8_inference_flag;
	uint8_t num_ref_idx_l0_active_minus1, num_ref_idx_l1_active_minus1;
	uint8_t field_pic_flag, mbaff_frame_flag, entropy_coding_mode_flag, num_slice_groups_minus1;
	uint8_t cabac_init_idc;
	int8_t slice_qp;
	uint8_t reserved[2];
};

/* neighbouring blocks, frame/non-MBAFF (H.264 6.4.11.4, 6.4.11.5) */
static const uint8_t luma_xy[16][2] = {{0, 0}, {4, 0}, {0, 4}, {4, 4}, {8, 0}, {12, 0}, {8, 4}, {12, 4},
				       {0, 8}, {4, 8}, {0, 12}, {4, 12}, {8, 8}, {12, 8}, {8, 12}, {12, 12}};
static const uint8_t luma_idx[4][4] = {{0, 2, 8, 10}, {1, 3, 9, 11}, {4, 6, 12, 14}, {5, 7, 13, 15}};

static int mb_avail_a(struct gen *g)
{
	uint32_t W = g->cfg.width_mbs;
	return g->cur_mb >= g->first_mb + 1 && g->cur_mb % W != 0 && g->mbs[g->cur_mb - 1].avail;
}
static int mb_avail_b(struct gen *g)
{
	uint32_t W = g->cfg.width_mbs;
	return g->cur_mb >= g->first_mb + W && g->mbs[g->cur_mb - W].avail;
}

/* MBAFF frames: 6.4.10 (pairs to the left / above inside the slice) + 6.4.12.2 (mbaff_nb.h) */
static uint32_t calc_nc_mbaff(struct gen *g, int comp, int blk, int chroma_ac)
{
	const uint32_t W = g->cfg.width_mbs, cur = g->cur_mb, half = cur / 2;
	const int bottom = (int)(cur & 1);
	const int a_avail = half >= g->first_mb + 1 && half % W != 0 && g->mbs[2 * (half - 1)].avail;
	const int b_avail = half >= g->first_mb + W && g->mbs[2 * (half - W)].avail;
	const int a_field = a_avail ? g->pair_field[half - 1] : 0, b_field = b_avail ? g->pair_field[half - W] : 0;
	int x, y, maxH, wb, hb;
	if (!chroma_ac) {
		x = luma_xy[blk][0] / 4;
		y = luma_xy[blk][1] / 4;
		maxH = 16;
		wb = hb = 4;
	} else {
		x = blk & 1;
		y = blk >> 1;
		maxH = (int)g->mbh_c;
		wb = (int)g->mbw_c / 4;
		hb = (int)g->mbh_c / 4;
	}
#define BLK(xx, yy) (chroma_ac ? 2 * (yy) + (xx) : luma_idx[(xx)][(yy)])
	int availA = 1, availB = 1;
	uint32_t nA = 0, nB = 0;
	if (x > 0) {
		nA = g->mbs[cur].nz[comp * 16 + BLK(x - 1, y)];
	} else {
		int yM = 0;
		const int w = mbaff_left(g->cur_field, bottom, a_avail, a_field, 4 * y, maxH, &yM);
		if (w == MBAFF_NB_NONE)
			availA = 0;
		else
			nA = g->mbs[2 * (half - 1) + (w == MBAFF_NB_A_BOT)].nz[comp * 16 + BLK(wb - 1, yM / 4)];
	}
	if (y > 0) {
		nB = g->mbs[cur].nz[comp * 16 + BLK(x, y - 1)];
	} else {
		const int w = mbaff_up(g->cur_field, bottom, b_avail, b_field);
		if (w == MBAFF_NB_NONE)
			availB = 0;
		else if (w == MBAFF_NB_CUR_TOP)
			nB = g->mbs[cur - 1].nz[comp * 16 + BLK(x, hb - 1)];
		else
			nB = g->mbs[2 * (half - W) + (w == MBAFF_NB_B_BOT)].nz[comp * 16 + BLK(x, hb - 1)];
	}
#undef BLK
	if (availA && availB)
		return (nA + nB + 1) >> 1;
	return availA ? nA : availB ? nB : 0;
}

static uint32_t calc_nc(struct gen *g, int comp, int blk, int chroma_ac)
{
	if (g->mbaff)
		return calc_nc_mbaff(g, comp, blk, chroma_ac);
	uint32_t W = g->cfg.width_mbs;
	int availA, availB;
	uint32_t nA = 0, nB = 0;
	if (!chroma_ac) {
		int x = luma_xy[blk][0], y = luma_xy[blk][1];
		if (x > 0) {
			availA = 1;
			nA = g->mbs[g->cur_mb].nz[comp * 16 + luma_idx[(x - 4) / 4][y / 4]];
		} else if ((availA = mb_avail_a(g))) {
			nA = g->mbs[g->cur_mb - 1].nz[comp * 16 + luma_idx[3][y / 4]];
		}
		if (y > 0) {
			availB = 1;
			nB = g->mbs[g->cur_mb].nz[comp * 16 + luma_idx[x / 4][(y - 4) / 4]];
		} else if ((availB = mb_avail_b(g))) {
			nB = g->mbs[g->cur_mb - W].nz[comp * 16 + luma_idx[x / 4][3]];
		}
	} else {
		int x = (blk & 1) * 4, y = (blk >> 1) * 4;
		int wb = (int)g->mbw_c / 4, hb = (int)g->mbh_c / 4;
		if (x > 0) {
			availA = 1;
			nA = g->mbs[g->cur_mb].nz[comp * 16 + blk - 1];
		} else if ((availA = mb_avail_a(g))) {
			nA = g->mbs[g->cur_mb - 1].nz[comp * 16 + 2 * (y / 4) + (wb - 1)];
		}
		if (y > 0) {
			availB = 1;
			nB = g->mbs[g->cur_mb].nz[comp * 16 + blk - 2];
		} else if ((availB = mb_avail_b(g))) {
			nB = g->mbs[g->cur_mb - W].nz[comp * 16 + 2 * (hb - 1) + x / 4];
		}
	}
	if (availA && availB)
		return (nA + nB + 1) >> 1;
	return availA ? nA : availB ? nB : 0;
}

/* ---- residual block ------------------------------------------------------------ */

/* random coefficients: n slots, returns count of non-zeros */
static int gen_coeffs(struct gen *g, int16_t *c, int n)
{
	memset(c, 0, sizeof(int16_t) * (size_t)n);
	if (!rnd_pct(&g->r, g->cfg.coef_density))
		return 0;
	int tc;
	uint32_t sel = rnd_n(&g->r, 100);
	if (sel < 45)
		tc = 1 + (int)rnd_n(&g->r, 3);
	else if (sel < 80)
		tc = 1 + (int)rnd_n(&g->r, (uint32_t)(n < 8 ? n : 8));
	else
		tc = 1 + (int)rnd_n(&g->r, (uint32_t)n);
	if (tc > n)
		tc = n;
	int placed = 0;
	/* low-frequency biased positions */
	while (placed < tc) {
		int span = (rnd_n(&g->r, 4) == 0) ? n : (tc + 3 < n ? tc + 3 : n);
		int p = (int)rnd_n(&g->r, (uint32_t)span);
		if (c[p])
			continue;
		int mag;
		uint32_t m = rnd_n(&g->r, 1000);
		if (m < 600)
			mag = 1;
		else if (m < 850)
			mag = 2 + (int)rnd_n(&g->r, 3);
		else if (m < 980)
			mag = 5 + (int)rnd_n(&g->r, 30);
		else
			mag = 35 + (int)rnd_n(&g->r, 900);
		c[p] = (int16_t)((rnd(&g->r) & 1) ? -mag : mag);
		placed++;
	}
	return tc;
}

/*
 * CAVLC residual_block (7.3.5.3.2 / 9.2): coeff_token table by `table`
 * (0..2 VLC by nC, 3 = 6-bit FLC, 4 = chroma DC 4:2:0, 5 = chroma DC 4:2:2).
 * Returns total_coeff.
 */
static int put_residual_block(struct gen *g, const int16_t *c, int n, int table)
{
	struct bw *w = &g->w;
	int16_t lev[64];
	int run[64];
	int tc = 0, zeros_before_last = 0;
	/* scan from the highest-frequency coefficient downwards */
	int last = -1;
	for (int i = n - 1; i >= 0; i--) {
		if (c[i]) {
			if (last < 0)
				last = i;
			lev[tc] = c[i];
			/* zeros between this coefficient and the previous (lower) one */
			int z = 0;
			for (int j = i - 1; j >= 0 && !c[j]; j--)
				z++;
			run[tc] = z;
			tc++;
		}
	}
	int t1 = 0;
	while (t1 < tc && t1 < 3 && (lev[t1] == 1 || lev[t1] == -1))
		t1++;
	/* coeff_token */
	if (table == 3) {
		uint32_t code = 0;
		for (code = 0; code < 64; code++)
			if (cavlc_coeff_token_flc[code] == (0x80 | t1 << 5 | tc))
				break;
		bw_bits(w, code, 6);
	} else {
		int t = table <= 2 ? table : table - 1; /* 4 -> LUT 3 (4:2:0), 5 -> LUT 4 (4:2:2) */
		struct enc e = e_coeff_token[t][t1 << 5 | tc];
		bw_bits(w, e.code, e.len);
	}
	if (tc == 0)
		return 0;
	int suffix_len = (tc > 10 && t1 < 3) ? 1 : 0;
	for (int i = 0; i < tc; i++) {
		if (i < t1) {
			bw_bits(w, lev[i] < 0, 1);
			continue;
		}
		int level = lev[i];
		int code = level > 0 ? 2 * level - 2 : -2 * level - 1;
		if (i == t1 && t1 < 3)
			code -= 2;
		if (suffix_len == 0) {
			if (code < 14) {
				bw_bits(w, 1, code + 1);
			} else if (code < 30) {
				bw_bits(w, 1, 15);
				bw_bits(w, (uint32_t)(code - 14), 4);
			} else {
				bw_bits(w, 1, 16);
				bw_bits(w, (uint32_t)(code - 30), 12);
			}
		} else {
			int prefix = code >> suffix_len;
			if (prefix < 15) {
				bw_bits(w, 1, prefix + 1);
				bw_bits(w, (uint32_t)code & ((1u << suffix_len) - 1), suffix_len);
			} else {
				bw_bits(w, 1, 16);
				bw_bits(w, (uint32_t)(code - (15 << suffix_len)), 12);
			}
		}
		if (suffix_len == 0)
			suffix_len = 1;
		int a = level < 0 ? -level : level;
		if (a > (3 << (suffix_len - 1)) && suffix_len < 6)
			suffix_len++;
	}
	/* total_zeros: zeros below the highest coefficient */
	int total_zeros = last + 1 - tc;
	if (tc < n) {
		struct enc e;
		if (n == 4)
			e = e_tz2x2[tc][total_zeros];
		else if (n == 8)
			e = e_tz2x4[tc][total_zeros];
		else
			e = e_tz4x4[tc][total_zeros];
		bw_bits(w, e.code, e.len);
	}
	int zeros_left = total_zeros;
	for (int i = 0; i < tc - 1 && zeros_left > 0; i++) {
		struct enc e = e_run[zeros_left > 6 ? 7 : zeros_left][run[i]];
		bw_bits(w, e.code, e.len);
		zeros_left -= run[i];
	}
	(void)zeros_before_last;
	return tc;
}

static int table_for_nc(uint32_t nc)
{
	return nc < 2 ? 0 : nc < 4 ? 1 : nc < 8 ? 2 : 3;
}

/* keep |level| small enough that level_prefix never needs to exceed 15 */
static void clamp_levels(int16_t *c, int n)
{
	for (int i = 0; i < n; i++) {
		if (c[i] > 1000)
			c[i] = 1000;
		if (c[i] < -1000)
			c[i] = -1000;
	}
}

static void put_block(struct gen *g, int comp, int blk, int n, int chroma_ac)
{
	int16_t c[16];
	int nz = gen_coeffs(g, c, n);
	/* twin mode: CABAC has no coded_block_flag for an 8x8 block, so it must not be empty */
	if (g->cfg.cabac_twin && g->cur_t8 && !chroma_ac && comp == 0 && (blk & 3) == 0 && nz == 0)
		c[rnd_n(&g->r, (uint32_t)n)] = (rnd(&g->r) & 1) ? -1 : 1;
	clamp_levels(c, n);
	uint32_t nc = calc_nc(g, comp, blk, chroma_ac);
	int tc = put_residual_block(g, c, n, table_for_nc(nc));
	g->mbs[g->cur_mb].nz[comp * 16 + blk] = (uint8_t)tc;
}

static void put_residual_luma(struct gen *g, int comp, int i16, uint32_t cbp_luma)
{
	if (i16)
		put_block(g, comp, 0, 16, 0); /* DC: stored at blkIdx 0 like the reference */
	for (int i8 = 0; i8 < 4; i8++)
		for (int i4 = 0; i4 < 4; i4++)
			if (cbp_luma & (1u << i8))
				put_block(g, comp, i8 * 4 + i4, i16 ? 15 : 16, 0);
}

static void put_residual(struct gen *g, int i16, uint32_t cbp_luma, uint32_t cbp_chroma)
{
	put_residual_luma(g, 0, i16, cbp_luma);
	if (g->cat == 1 || g->cat == 2) {
		int nc8 = g->cat == 1 ? 1 : 2;
		for (int ic = 0; ic < 2; ic++) {
			if (cbp_chroma & 3) {
				int16_t c[8];
				gen_coeffs(g, c, 4 * nc8);
				clamp_levels(c, 4 * nc8);
				int tc = put_residual_block(g, c, 4 * nc8, g->cat == 1 ? 4 : 5);
				g->mbs[g->cur_mb].nz[(1 + ic) * 16 + 0] = (uint8_t)tc;
			}
		}
		for (int ic = 0; ic < 2; ic++)
			for (int b = 0; b < 4 * nc8; b++)
				if (cbp_chroma & 2)
					put_block(g, 1 + ic, b, 15, 1);
	} else if (g->cat == 3) {
		put_residual_luma(g, 1, i16, cbp_luma);
		put_residual_luma(g, 2, i16, cbp_luma);
	}
}

/* ---- macroblock layer ---------------------------------------------------------- */

static void put_cbp(struct gen *g, int intra, uint32_t cbp)
{
	if (g->cat == 1 || g->cat == 2)
		bw_ue(&g->w, e_cbp_chroma[intra ? 0 : 1][cbp]);
	else
		bw_ue(&g->w, e_cbp_nochroma[intra ? 0 : 1][cbp & 15]);
}

static uint32_t rand_cbp(struct gen *g)
{
	uint32_t luma = rnd_n(&g->r, 100) < 30 ? 0 : (rnd_n(&g->r, 100) < 40 ? 15 : rnd_n(&g->r, 16));
	uint32_t chroma = (g->cat == 1 || g->cat == 2) ? rnd_n(&g->r, 3) : 0;
	return luma | chroma << 4;
}

static int32_t rand_mvd(struct gen *g)
{
	uint32_t s = rnd_n(&g->r, 100);
	int32_t m = s < 50 ? 0 : s < 85 ? (int32_t)rnd_n(&g->r, 8) : (int32_t)rnd_n(&g->r, 300);
	return (rnd(&g->r) & 1) ? -m : m;
}

static void put_intra_mb(struct gen *g, uint32_t type_base)
{
	struct bw *w = &g->w;
	uint32_t sel = rnd_n(&g->r, 1000);
	if (sel < g->cfg.pct_pcm) {
		/* I_PCM */
		bw_ue(w, type_base + 25);
		while (w->nacc)
			bw_bits(w, 0, 1);
		for (int i = 0; i < 256; i++)
			bw_bits(w, rnd_n(&g->r, 256), 8);
		for (uint32_t i = 0; i < 2 * g->mbw_c * g->mbh_c; i++)
			bw_bits(w, rnd_n(&g->r, 256), 8);
		memset(g->mbs[g->cur_mb].nz, 16, 48);
		return;
	}
	if (sel < 500) {
		/* I_NxN */
		bw_ue(w, type_base + 0);
		int t8 = 0;
		if (g->cfg.transform_8x8) {
			t8 = (int)(rnd(&g->r) & 1);
			bw_bits(w, (uint32_t)t8, 1);
		}
		g->cur_t8 = t8;
		for (int i = 0; i < (t8 ? 4 : 16); i++) {
			if (rnd(&g->r) & 1) {
				bw_bits(w, 1, 1);
			} else {
				bw_bits(w, 0, 1);
				bw_bits(w, rnd_n(&g->r, 8), 3);
			}
		}
		if (g->cat == 1 || g->cat == 2)
			bw_ue(w, rnd_n(&g->r, 4));
		uint32_t cbp = rand_cbp(g);
		put_cbp(g, 1, cbp);
		if (cbp) {
			bw_se(w, (int32_t)rnd_n(&g->r, 7) - 3);
			put_residual(g, 0, cbp & 15, cbp >> 4);
		}
		return;
	}
	/* I_16x16: mb_type 1..24 = 1 + predmode + 4*cbp_chroma + 12*(cbp_luma==15) */
	uint32_t pm = rnd_n(&g->r, 4);
	uint32_t cc = (g->cat == 1 || g->cat == 2) ? rnd_n(&g->r, 3) : 0;
	uint32_t cl = (rnd(&g->r) & 1) ? 15 : 0;
	bw_ue(w, type_base + 1 + pm + 4 * cc + (cl ? 12 : 0));
	if (g->cat == 1 || g->cat == 2)
		bw_ue(w, rnd_n(&g->r, 4));
	bw_se(w, (int32_t)rnd_n(&g->r, 7) - 3);
	put_residual(g, 1, cl, cc);
}

static void put_inter_tail(struct gen *g, int allow_t8)
{
	struct bw *w = &g->w;
	uint32_t cbp = rand_cbp(g);
	put_cbp(g, 0, cbp);
	if ((cbp & 15) && g->cfg.transform_8x8 && allow_t8) {
		g->cur_t8 = (int)(rnd(&g->r) & 1);
		bw_bits(w, (uint32_t)g->cur_t8, 1);
	}
	if (cbp) {
		bw_se(w, (int32_t)rnd_n(&g->r, 7) - 3);
		put_residual(g, 0, cbp & 15, cbp >> 4);
	}
}

static void put_p_mb(struct gen *g)
{
	struct bw *w = &g->w;
	uint32_t sel = rnd_n(&g->r, 100);
	if (sel < 45) {
		bw_ue(w, 0); /* P_L0_16x16 */
		if (g->ref_present)
			bw_te(w, rnd_n(&g->r, g->ref_range), g->ref_range - 1);
		bw_se(w, rand_mvd(g));
		bw_se(w, rand_mvd(g));
		put_inter_tail(g, 1);
	} else if (sel < 70) {
		bw_ue(w, 1 + (rnd(&g->r) & 1)); /* 16x8 / 8x16 */
		if (g->ref_present) {
			bw_te(w, rnd_n(&g->r, g->ref_range), g->ref_range - 1);
			bw_te(w, rnd_n(&g->r, g->ref_range), g->ref_range - 1);
		}
		for (int i = 0; i < 4; i++)
			bw_se(w, rand_mvd(g));
		put_inter_tail(g, 1);
	} else {
		int ref0 = rnd_n(&g->r, 5) == 0 && !g->cfg.cabac_twin;
		bw_ue(w, ref0 ? 4 : 3); /* P_8x8ref0 / P_8x8 */
		uint32_t sub[4];
		int no_small = 1;
		for (int i = 0; i < 4; i++) {
			sub[i] = rnd_n(&g->r, 4);
			bw_ue(w, sub[i]);
			if (sub[i] != 0)
				no_small = 0;
		}
		if (g->ref_present && !ref0)
			for (int i = 0; i < 4; i++)
				bw_te(w, rnd_n(&g->r, g->ref_range), g->ref_range - 1);
		static const int nsub[4] = {1, 2, 2, 4};
		for (int i = 0; i < 4; i++)
			for (int s = 0; s < nsub[sub[i]]; s++) {
				bw_se(w, rand_mvd(g));
				bw_se(w, rand_mvd(g));
			}
		put_inter_tail(g, no_small);
	}
}

/* B sub_mb_type -> {NumSubMbPart, pred: 0 L0, 1 L1, 2 Bi, 3 direct} */
static const uint8_t b_sub[13][2] = {{4, 3}, {1, 0}, {1, 1}, {1, 2}, {2, 0}, {2, 0}, {2, 1},
				     {2, 1}, {2, 2}, {2, 2}, {4, 0}, {4, 1}, {4, 2}};
/* B mb_type 4..21 -> pred of the two partitions */
static const uint8_t b_part[18][2] = {{0, 0}, {0, 0}, {1, 1}, {1, 1}, {0, 1}, {0, 1}, {1, 0}, {1, 0}, {0, 2},
				      {0, 2}, {1, 2}, {1, 2}, {2, 0}, {2, 0}, {2, 1}, {2, 1}, {2, 2}, {2, 2}};

static void put_b_mb(struct gen *g)
{
	struct bw *w = &g->w;
	uint32_t sel = rnd_n(&g->r, 100);
	if (sel < 15) {
		bw_ue(w, 0); /* B_Direct_16x16 */
		put_inter_tail(g, 1); /* direct_8x8_inference_flag = 1 in our SPS */
	} else if (sel < 45) {
		uint32_t ty = 1 + rnd_n(&g->r, 3); /* L0, L1, Bi 16x16 */
		bw_ue(w, ty);
		int use0 = ty != 2, use1 = ty != 1;
		if (g->ref_present && use0)
			bw_te(w, rnd_n(&g->r, g->ref_range), g->ref_range - 1);
		if (g->ref_present && use1)
			bw_te(w, rnd_n(&g->r, g->ref_range), g->ref_range - 1);
		if (use0) {
			bw_se(w, rand_mvd(g));
			bw_se(w, rand_mvd(g));
		}
		if (use1) {
			bw_se(w, rand_mvd(g));
			bw_se(w, rand_mvd(g));
		}
		put_inter_tail(g, 1);
	} else if (sel < 75) {
		uint32_t ty = 4 + rnd_n(&g->r, 18);
		bw_ue(w, ty);
		const uint8_t *pp = b_part[ty - 4];
		if (g->ref_present)
			for (int i = 0; i < 2; i++)
				if (pp[i] != 1)
					bw_te(w, rnd_n(&g->r, g->ref_range), g->ref_range - 1);
		if (g->ref_present)
			for (int i = 0; i < 2; i++)
				if (pp[i] != 0)
					bw_te(w, rnd_n(&g->r, g->ref_range), g->ref_range - 1);
		for (int i = 0; i < 2; i++)
			if (pp[i] != 1) {
				bw_se(w, rand_mvd(g));
				bw_se(w, rand_mvd(g));
			}
		for (int i = 0; i < 2; i++)
			if (pp[i] != 0) {
				bw_se(w, rand_mvd(g));
				bw_se(w, rand_mvd(g));
			}
		put_inter_tail(g, 1);
	} else {
		bw_ue(w, 22); /* B_8x8 */
		uint32_t sub[4];
		int no_small = 1;
		for (int i = 0; i < 4; i++) {
			sub[i] = rnd_n(&g->r, 13);
			bw_ue(w, sub[i]);
			if (sub[i] != 0 && b_sub[sub[i]][0] > 1)
				no_small = 0;
		}
		if (g->ref_present)
			for (int i = 0; i < 4; i++)
				if (b_sub[sub[i]][1] != 3 && b_sub[sub[i]][1] != 1)
					bw_te(w, rnd_n(&g->r, g->ref_range), g->ref_range - 1);
		if (g->ref_present)
			for (int i = 0; i < 4; i++)
				if (b_sub[sub[i]][1] != 3 && b_sub[sub[i]][1] != 0)
					bw_te(w, rnd_n(&g->r, g->ref_range), g->ref_range - 1);
		for (int i = 0; i < 4; i++)
			if (b_sub[sub[i]][1] != 3 && b_sub[sub[i]][1] != 1)
				for (int s = 0; s < b_sub[sub[i]][0]; s++) {
					bw_se(w, rand_mvd(g));
					bw_se(w, rand_mvd(g));
				}
		for (int i = 0; i < 4; i++)
			if (b_sub[sub[i]][1] != 3 && b_sub[sub[i]][1] != 0)
				for (int s = 0; s < b_sub[sub[i]][0]; s++) {
					bw_se(w, rand_mvd(g));
					bw_se(w, rand_mvd(g));
				}
		put_inter_tail(g, no_small);
	}
}

/* ---- slice ----------------------------------------------------------------------- */

static void put_slice(struct gen *g, uint32_t frame, int idr, int type, uint32_t first, uint32_t count)
{
	struct bw *w = &g->w;
	uint32_t W = g->cfg.width_mbs, H = g->cfg.height_mbs;
	const size_t nal_off = w->len + 4;
	const uint64_t mb0 = g->n_mbs;
	bw_start_nal(w, idr ? 0x65 : (type == 1 ? 0x01 : 0x41));
	bw_ue(w, first);
	bw_ue(w, (uint32_t)type + ((frame & 1) ? 5 : 0)); /* both slice_type codings */
	bw_ue(w, 0);
	bw_bits(w, frame & 0xff, 8); /* frame_num, log2_max_frame_num = 8 */
	if (g->mbaff)
		bw_bits(w, 0, 1); /* field_pic_flag: a frame, so MbaffFrameFlag = 1 */
	if (g->paff) {
		bw_bits(w, 1, 1);         /* field_pic_flag */
		bw_bits(w, frame & 1, 1); /* bottom_field_flag */
	}
	if (idr)
		bw_ue(w, frame & 0xffff);
	bw_bits(w, (2 * frame) & 0xff, 8); /* pic_order_cnt_lsb */
	if (type == 1)
		bw_bits(w, 1, 1); /* direct_spatial_mv_pred_flag */
	if (type != 2)
		bw_bits(w, 0, 1); /* num_ref_idx_active_override_flag */
	if (type != 2)
		bw_bits(w, 0, 1); /* ref_pic_list_modification_flag_l0 */
	if (type == 1)
		bw_bits(w, 0, 1); /* ..._l1 */
	if (idr) {
		bw_bits(w, 0, 2); /* no_output_of_prior_pics, long_term_reference */
	} else if (type != 1) {
		bw_bits(w, 0, 1); /* adaptive_ref_pic_marking_mode_flag (nal_ref_idc != 0) */
	}
	const uint32_t cabac_init_idc = (g->cfg.entropy_cabac && type != 2) ? rnd_n(&g->r, 3) : 0;
	if (g->cfg.entropy_cabac && type != 2)
		bw_ue(w, cabac_init_idc);
	const int32_t qp_delta = (int32_t)rnd_n(&g->r, 9) - 4;
	bw_se(w, qp_delta); /* slice_qp_delta */
	bw_ue(w, 0);        /* disable_deblocking_filter_idc */
	bw_se(w, 0);
	bw_se(w, 0);
	if (g->fmo_groups >= 2 && g->fmo_type >= 3 && g->fmo_type <= 5 && g->fmo_cycle_bits)
		bw_bits(w, g->fmo_cycle, (int)g->fmo_cycle_bits); /* slice_group_change_cycle */
	/* raw bit offset of slice_data() in the NAL (what the reference keeps in
	 * ctx->slice.hdr_len); fixed up by bw_byte if an EPB lands before this byte */
	uint32_t data_bit_off = (uint32_t)((w->len - nal_off) * 8 + (size_t)w->nacc);
	w->mark = w->nacc ? &data_bit_off : NULL;

	/* slice data */
	g->slice_type = type;
	g->first_mb = first;
	memset(g->mbs, 0, sizeof(struct mbstate) * W * H);
	uint32_t pending_skip = 0;
	if (g->cfg.entropy_cabac) {
		struct synth_cabac_slice_cfg cc;
		memset(&cc, 0, sizeof(cc));
		cc.width_mbs = W;
		cc.height_mbs = H;
		cc.first_mb = first;
		cc.count = count;
		cc.slice_type = (uint32_t)type;
		cc.chroma_array_type = g->cat;
		cc.transform_8x8 = g->cfg.transform_8x8 ? 1 : 0;
		cc.direct_8x8_inference = 1;
		cc.num_ref_l0 = cc.num_ref_l1 = g->cfg.num_ref_frames;
		cc.cabac_init_idc = cabac_init_idc;
		cc.slice_qp = 26 + qp_delta;
		cc.pct_skip = g->cfg.pct_skip;
		cc.pct_intra_in_inter = g->cfg.pct_intra_in_inter;
		cc.pct_pcm = g->cfg.pct_pcm;
		cc.coef_density = g->cfg.coef_density;
		while (w->nacc)
			bw_bits(w, 1, 1); /* cabac_alignment_one_bit */
		w->mark = NULL;
		const uint64_t saved = g->r.s;
		uint64_t need = synth_cabac_slice(&cc, &g->r.s, g->cabac_buf, g->cabac_cap);
		if (need > g->cabac_cap) { /* grow and redo with the same random state */
			free(g->cabac_buf);
			g->cabac_cap = need + (need >> 2) + 4096;
			g->cabac_buf = malloc(g->cabac_cap);
			g->r.s = saved;
			need = synth_cabac_slice(&cc, &g->r.s, g->cabac_buf, g->cabac_cap);
		}
		for (uint64_t i = 0; i < need && i < g->cabac_cap; i++)
			bw_byte(w, g->cabac_buf[i]);
		g->n_mbs += count;
		count = 0; /* the CAVLC loop below does nothing */
	}
	g->ref_present = g->cfg.num_ref_frames > 1;
	g->ref_range = g->cfg.num_ref_frames;
	g->cur_field = 0;
	if (g->mbaff) {
		/* macroblock pairs first .. first + count - 1: addresses 2 * first .. 2 * (first + count) - 1 */
		memset(g->pair_field, 0, (size_t)W * H / 2);
		int top_skipped = 0;
		for (uint32_t a = 2 * first; a < 2 * (first + count); a++) {
			const int bottom = (int)(a & 1);
			g->cur_mb = a;
			g->cur_t8 = 0;
			g->mbs[a].avail = 1;
			g->n_mbs++;
			if (!bottom) {
				g->pair_field[a / 2] = (uint8_t)(rnd(&g->r) & 1);
				top_skipped = 0;
			}
			if (type != 2 && rnd_pct(&g->r, g->cfg.pct_skip)) {
				pending_skip++;
				if (!bottom)
					top_skipped = 1;
				else if (top_skipped)
					g->pair_field[a / 2] = 0; /* inferred by the reader; its counts are all 0 */
				continue;
			}
			if (type != 2) {
				bw_ue(w, pending_skip);
				pending_skip = 0;
			}
			/* mb_field_decoding_flag: with the top macroblock, or with the bottom one when the
			 * top was skipped (7.3.4) */
			g->cur_field = g->pair_field[a / 2];
			if (!bottom || top_skipped)
				bw_bits(w, (uint32_t)g->cur_field, 1);
			g->ref_present = g->cfg.num_ref_frames > 1 || g->cur_field;
			g->ref_range = g->cur_field ? 2 * g->cfg.num_ref_frames : g->cfg.num_ref_frames;
			if (type == 2)
				put_intra_mb(g, 0);
			else if (rnd_pct(&g->r, g->cfg.pct_intra_in_inter))
				put_intra_mb(g, type == 0 ? 5 : 23);
			else if (type == 0)
				put_p_mb(g);
			else
				put_b_mb(g);
		}
		count = 0;
	}
	for (uint32_t a = first, i = 0; i < count; i++, a = gen_next_mb(g, a)) {
		g->cur_mb = a;
		g->cur_t8 = 0;
		g->mbs[a].avail = 1;
		g->n_mbs++;
		if (type != 2 && rnd_pct(&g->r, g->cfg.pct_skip)) {
			pending_skip++;
			continue;
		}
		if (type != 2) {
			bw_ue(w, pending_skip);
			pending_skip = 0;
		}
		if (type == 2)
			put_intra_mb(g, 0);
		else if (rnd_pct(&g->r, g->cfg.pct_intra_in_inter))
			put_intra_mb(g, type == 0 ? 5 : 23);
		else if (type == 0)
			put_p_mb(g);
		else
			put_b_mb(g);
	}
	if (type != 2 && pending_skip)
		bw_ue(w, pending_skip);
	if (!g->cfg.entropy_cabac)
		bw_trailing(w); /* CABAC: the flush wrote the stop bit, the slice is byte aligned */
	w->mark = NULL;
	if (g->params && g->params_n < g->params_cap) {
		struct slice_params_out *p = &g->params[g->params_n];
		memset(p, 0, sizeof(*p));
		p->nal_off = nal_off;
		p->nal_len = (uint32_t)(w->len - nal_off);
		p->data_bit_off = data_bit_off;
		p->first_mb_in_slice = first;
		p->mb_out_off = (uint32_t)mb0;
		p->mb_out_cap = (uint32_t)(g->n_mbs - mb0);
		p->pic_width_in_mbs = (uint16_t)W;
		p->pic_height_in_mbs = (uint16_t)g->pic_h;
		p->field_pic_flag = (uint8_t)g->paff;
		p->slice_type = (uint8_t)type;
		p->chroma_array_type = (uint8_t)g->cat;
		p->bit_depth_luma = 8;
		p->bit_depth_chroma = 8;
		p->transform_8x8_mode_flag = (uint8_t)(g->cfg.transform_8x8 ? 1 : 0);
		p->direct_8x8_inference_flag = 1;
		p->num_ref_idx_l0_active_minus1 = (uint8_t)(g->cfg.num_ref_frames - 1);
		p->num_ref_idx_l1_active_minus1 = (uint8_t)(g->cfg.num_ref_frames - 1);
		p->entropy_coding_mode_flag = (uint8_t)(g->cfg.entropy_cabac ? 1 : 0);
		p->cabac_init_idc = (uint8_t)cabac_init_idc;
		p->slice_qp = (int8_t)(26 + qp_delta);
		p->num_slice_groups_minus1 = (uint8_t)(g->fmo_groups >= 2 ? g->fmo_groups - 1 : 0);
		p->mbaff_frame_flag = (uint8_t)g->mbaff;
	}
	g->params_n++;
}

/*
 * Generate the stream.  Returns bytes needed (== written when <= cap).
 * slice_mb_total (optional) gets the number of macroblocks in all slices.
 */
uint64_t synth_video(const struct synth_video_cfg *cfg, uint8_t *out, uint64_t cap,
		     uint64_t *total_mbs, uint64_t *total_slices, void *params_out,
		     uint64_t params_cap)
{
	struct gen g;
	build_tables();
	memset(&g, 0, sizeof(g));
	g.params = params_out;
	g.params_cap = params_out ? params_cap : 0;
	g.cfg = *cfg;
	if (g.cfg.num_ref_frames < 1)
		g.cfg.num_ref_frames = 1;
	if (g.cfg.num_ref_frames > 4)
		g.cfg.num_ref_frames = 4;
	if (g.cfg.slices_per_frame < 1)
		g.cfg.slices_per_frame = 1;
	if (g.cfg.idr_period < 1)
		g.cfg.idr_period = 1;
	g.r.s = cfg->seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
	if (!g.r.s)
		g.r.s = 1;
	g.w.buf = out;
	g.w.cap = cap;
	uint32_t W = cfg->width_mbs, H = cfg->height_mbs;
	int high = cfg->profile_idc >= 100;
	g.cat = high ? cfg->chroma_format_idc : 1;
	g.mbw_c = g.cat == 0 ? 0 : (g.cat == 3 ? 16 : 8);
	g.mbh_c = g.cat == 0 ? 0 : (g.cat == 1 ? 8 : 16);
	g.mbs = calloc((size_t)W * H, sizeof(struct mbstate));
	struct bw *w = &g.w;

	/* SPS */
	bw_start_nal(w, 0x67);
	bw_bits(w, cfg->profile_idc, 8);
	bw_bits(w, 0, 8);
	bw_bits(w, 40, 8);
	bw_ue(w, 0);
	if (high) {
		bw_ue(w, cfg->chroma_format_idc);
		if (cfg->chroma_format_idc == 3)
			bw_bits(w, 0, 1);
		bw_ue(w, 0);
		bw_ue(w, 0);
		bw_bits(w, 0, 1);
		bw_bits(w, 0, 1);
	}
	bw_ue(w, 4); /* log2_max_frame_num_minus4 */
	bw_ue(w, 0); /* pic_order_cnt_type */
	bw_ue(w, 4); /* log2_max_pic_order_cnt_lsb_minus4 */
	bw_ue(w, 4); /* max_num_ref_frames */
	bw_bits(w, 0, 1);
	g.mbaff = !cfg->entropy_cabac && (cfg->fmo & 0x100) && (H % 2) == 0;
	bw_ue(w, W - 1);
	g.paff = !g.mbaff && (cfg->fmo & 0x200) && (H % 2) == 0;
	g.pic_h = g.paff ? H / 2 : H;
	if (g.mbaff || g.paff) {
		g.pair_field = calloc((size_t)W * H / 2 + 1, 1);
		bw_ue(w, H / 2 - 1); /* pic_height_in_map_units_minus1: pairs / field rows */
		bw_bits(w, 0, 1);    /* frame_mbs_only_flag */
		bw_bits(w, g.mbaff ? 1 : 0, 1); /* mb_adaptive_frame_field_flag */
	} else {
		bw_ue(w, H - 1);
		bw_bits(w, 1, 1); /* frame_mbs_only_flag */
	}
	bw_bits(w, 1, 1); /* direct_8x8_inference_flag */
	bw_bits(w, 0, 1); /* frame_cropping_flag */
	bw_bits(w, 0, 1); /* vui_parameters_present_flag */
	bw_trailing(w);

	/* PPS */
	bw_start_nal(w, 0x68);
	bw_ue(w, 0);
	bw_ue(w, 0);
	bw_bits(w, cfg->entropy_cabac ? 1 : 0, 1);
	bw_bits(w, 0, 1);
	g.fmo_groups = cfg->entropy_cabac || g.mbaff ? 0 : (cfg->fmo & 15);
	g.fmo_type = (cfg->fmo >> 4) & 15;
	if (g.fmo_groups > 8 || g.fmo_type > 6 || (g.fmo_type == 6 && W * H > 256))
		g.fmo_groups = 0;
	if (g.fmo_groups >= 2 && g.fmo_type >= 3 && g.fmo_type <= 5)
		g.fmo_groups = 2; /* the evolving map types have two groups */
	if (g.fmo_groups < 2) {
		g.fmo_groups = 0;
		bw_ue(w, 0); /* num_slice_groups_minus1 */
	} else {
		const uint32_t N = W * H;
		g.fmo_units = malloc(N);
		g.fmo_map = malloc(N);
		bw_ue(w, g.fmo_groups - 1);
		bw_ue(w, g.fmo_type);
		if (g.fmo_type == 0) {
			for (uint32_t i = 0; i < g.fmo_groups; i++) {
				g.fmo_run[i] = W / 2 + 3 * i;
				bw_ue(w, g.fmo_run[i]);
			}
		} else if (g.fmo_type == 2) {
			for (uint32_t i = 0; i + 1 < g.fmo_groups; i++) {
				const uint32_t x0 = i % W, y0 = i % H;
				const uint32_t x1 = W / 2 + i < W ? W / 2 + i : W - 1, y1 = H / 2 + i < H ? H / 2 + i : H - 1;
				g.fmo_tl[i] = y0 * W + x0;
				g.fmo_br[i] = y1 * W + x1;
				bw_ue(w, g.fmo_tl[i]);
				bw_ue(w, g.fmo_br[i]);
			}
		} else if (g.fmo_type >= 3 && g.fmo_type <= 5) {
			g.fmo_dir = (int)(cfg->seed & 1);
			g.fmo_rate = W / 2 + 1;
			bw_bits(w, (uint32_t)g.fmo_dir, 1);
			bw_ue(w, g.fmo_rate - 1);
			uint32_t v = N / g.fmo_rate + 1, b = 0;
			while (b < 32 && (1ull << b) < v)
				b++;
			g.fmo_cycle_bits = b;
		} else if (g.fmo_type == 6) {
			uint32_t b = 0;
			while ((1u << b) < g.fmo_groups)
				b++;
			bw_ue(w, N - 1);
			for (uint32_t i = 0; i < N; i++) {
				g.fmo_ids[i] = (uint32_t)((((i + 1) * 2654435761u) >> 16) % g.fmo_groups);
				bw_bits(w, g.fmo_ids[i], (int)b);
			}
		}
	}
	bw_ue(w, g.cfg.num_ref_frames - 1);
	bw_ue(w, g.cfg.num_ref_frames - 1);
	bw_bits(w, 0, 1);
	bw_bits(w, 0, 2);
	bw_se(w, 0);
	bw_se(w, 0);
	bw_se(w, 0);
	bw_bits(w, 1, 1); /* deblocking_filter_control_present_flag */
	bw_bits(w, 0, 1);
	bw_bits(w, 0, 1);
	if (cfg->transform_8x8) {
		bw_bits(w, 1, 1);
		bw_bits(w, 0, 1);
		bw_se(w, 0);
	}
	bw_trailing(w);

	uint64_t nslices = 0;
	uint32_t S = g.cfg.slices_per_frame, N = W * g.pic_h;
	if (S > N)
		S = N;
	for (uint32_t f = 0; f < cfg->frames; f++) {
		int idr = (f % g.cfg.idr_period) == 0;
		int type = idr ? 2 : ((cfg->b_frames && (f & 1)) ? 1 : 0);
		if (g.fmo_groups >= 2) {
			/* every slice group of the picture in turn, each cut into S slices of consecutive
			 * macroblocks of the group */
			if (g.fmo_cycle_bits)
				g.fmo_cycle = f % ((N + g.fmo_rate - 1) / g.fmo_rate + 1);
			fmo_picture_map(&g);
			for (uint32_t grp = 0; grp < g.fmo_groups; grp++) {
				uint32_t n_grp = 0, a0 = N;
				for (uint32_t a = 0; a < N; a++)
					if (g.fmo_map[a] == grp) {
						if (n_grp++ == 0)
							a0 = a;
					}
				uint32_t a = a0, done = 0;
				for (uint32_t s = 0; s < S && n_grp > 0; s++) {
					const uint32_t upto = (uint32_t)((uint64_t)n_grp * (s + 1) / S);
					if (upto == done)
						continue;
					put_slice(&g, f, idr, type, a, upto - done);
					nslices++;
					for (; done < upto; done++)
						a = gen_next_mb(&g, a);
				}
			}
			continue;
		}
		for (uint32_t s = 0; s < S; s++) {
			const uint32_t units = g.mbaff ? N / 2 : N; /* first_mb_in_slice counts pairs in MBAFF frames */
			uint32_t first = (uint32_t)((uint64_t)units * s / S);
			uint32_t next = (uint32_t)((uint64_t)units * (s + 1) / S);
			if (next == first)
				continue;
			put_slice(&g, f, idr, type, first, next - first);
			nslices++;
		}
	}
	free(g.mbs);
	free(g.pair_field);
	free(g.fmo_units);
	free(g.fmo_map);
	free(g.cabac_buf);
	if (total_mbs)
		*total_mbs = g.n_mbs;
	if (total_slices)
		*total_slices = nslices;
	return g.w.len;
}
