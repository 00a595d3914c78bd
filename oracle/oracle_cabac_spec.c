/*
 * oracle_cabac_spec.c — an INDEPENDENT CABAC slice-data decoder written from Rec. ITU-T H.264
 * clauses 7.3.4 / 7.3.5 (syntax), 9.3.1 (initialisation), 9.3.2 (binarisations), 9.3.3.1
 * (ctxIdx derivation) and 9.3.3.2 (arithmetic decoding engine).
 *
 * TEST INFRASTRUCTURE ONLY.  The reference (Parrot-Developers/libh264) has no CABAC decoder (it
 * returns before CABAC slice data, src/h264_syntax_slice_data.h:715-717), so the GPU kernel K5
 * (libh264_b200/csrc/cabac_parse.cuh) and the synthetic stream generator share one syntax walker
 * (cabac_syntax.h): a wrong ctxIdxInc there is invisible to a generator-vs-kernel test.  This
 * file shares NOTHING with that walker or its engine: own bit reader, own arithmetic decoder,
 * own binarisations and context selection, table-driven where the standard is.  The only shared
 * inputs are constants of the standard (the (m, n) initialisation tables, rangeTabLPS and the
 * state transitions: cabac_tables.h, generated from the compiled reference's copies) and the
 * checksum definition of include/h264gpu_slice.h.
 *
 * Scope = the kernel's: I / P / B slices, frame macroblocks (no MBAFF, no field pictures), one
 * slice group, ChromaArrayType 0 / 1 / 2, 8x8 transform, I_PCM.
 * Parity: pinned by tests/test_cabac.py against (a) the REFERENCE's parse of the CAVLC twin of
 * every transcoded stream and (b) the reference writer's CABAC concealment slices.
 */
#include <errno.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "h264gpu_slice.h"

#define CABAC_TAB static const
#include "cabac_tables.h"

/* enum h264_mb_type (include/h264/h264_types.h:70-90) */
enum { T_UNKNOWN = 0, T_I_NxN, T_I_16x16, T_I_PCM, T_SI, T_P_16x16, T_P_16x8, T_P_8x16, T_P_8x8, T_P_8x8ref0,
       T_P_SKIP, T_B_Direct_16x16, T_B_16x16, T_B_16x8, T_B_8x16, T_B_8x8, T_B_SKIP };
enum { SL_P = 0, SL_B = 1, SL_I = 2 };
enum { PRED_L0 = 0, PRED_L1 = 1, PRED_BI = 2, PRED_DIRECT = 3 };

/* ---- 7.2 / 7.4.1.1: RBSP from the escaped NAL --------------------------------------------- */
struct rbsp {
	uint8_t *b;
	uint64_t n;      /* bytes */
	uint64_t *raw;   /* raw[i] = offset in the NAL of RBSP byte i (for end_bit) */
};

static int unescape(const uint8_t *nal, uint32_t len, struct rbsp *r)
{
	r->b = malloc((size_t)len + 8);
	r->raw = malloc(((size_t)len + 8) * sizeof(uint64_t));
	if (!r->b || !r->raw)
		return -ENOMEM;
	uint32_t zeros = 0;
	r->n = 0;
	for (uint32_t i = 0; i < len; i++) {
		if (zeros >= 2 && nal[i] == 3) {
			zeros = 0;
			continue; /* emulation_prevention_three_byte */
		}
		zeros = nal[i] == 0 ? zeros + 1 : 0;
		r->raw[r->n] = i;
		r->b[r->n++] = nal[i];
	}
	r->raw[r->n] = len;
	memset(r->b + r->n, 0, 8);
	return 0;
}

/* ---- 9.3.3.2 arithmetic decoding engine ---------------------------------------------------- */
struct dec {
	const struct rbsp *r;
	uint64_t bitpos;   /* next RBSP bit */
	uint32_t range, offset;
	uint8_t state[1024], mps[1024]; /* pStateIdx, valMPS */
	int overrun;
};

static uint32_t rd_bit(struct dec *d)
{
	const uint64_t byte = d->bitpos >> 3;
	if (byte >= d->r->n) {
		d->overrun = 1;
		d->bitpos++;
		return 0;
	}
	const uint32_t v = (d->r->b[byte] >> (7 - (d->bitpos & 7))) & 1;
	d->bitpos++;
	return v;
}

static uint32_t rd_bits(struct dec *d, int n)
{
	uint32_t v = 0;
	while (n-- > 0)
		v = (v << 1) | rd_bit(d);
	return v;
}

/* 9.3.1.1 */
static void init_contexts(struct dec *d, int slice_type, uint32_t cabac_init_idc, int qp)
{
	const int8_t(*mn)[2] = cabac_init_mn[slice_type == SL_I ? 0 : 1 + (cabac_init_idc % 3)];
	/* The standard clips SliceQPY to 0..51 here; the reference's h264_bac_state_init clips to 1..51
	 * (src/h264_bac.c:220), which differs only for SliceQPY = 0.  The reference's writer is what
	 * pins this decoder (its CABAC concealment slices must decode), so its clip is used; the K5
	 * kernel does the same.  Found by this decoder: with the standard's clip the reference's own
	 * slice with SliceQPY = 0 (tests/test_cabac.py, seed 416) fails after 3 macroblocks. */
	if (qp < 1)
		qp = 1;
	if (qp > 51)
		qp = 51;
	for (int i = 0; i < 1024; i++) {
		int pre = ((mn[i][0] * qp) >> 4) + mn[i][1];
		if (pre < 1)
			pre = 1;
		if (pre > 126)
			pre = 126;
		if (pre <= 63) {
			d->state[i] = (uint8_t)(63 - pre);
			d->mps[i] = 0;
		} else {
			d->state[i] = (uint8_t)(pre - 64);
			d->mps[i] = 1;
		}
	}
}

/* 9.3.1.2 */
static void init_engine(struct dec *d)
{
	d->range = 510;
	d->offset = rd_bits(d, 9);
}

static void renorm(struct dec *d)
{
	while (d->range < 256) {
		d->range <<= 1;
		d->offset = (d->offset << 1) | rd_bit(d);
	}
}

/* 9.3.3.2.1 DecodeDecision */
static uint32_t dd(struct dec *d, uint32_t ctx)
{
	const uint32_t s = d->state[ctx];
	const uint32_t lps = cabac_range_lps[s][(d->range >> 6) & 3];
	uint32_t bin;
	d->range -= lps;
	if (d->offset >= d->range) {
		bin = !d->mps[ctx];
		d->offset -= d->range;
		d->range = lps;
		if (s == 0)
			d->mps[ctx] = (uint8_t)!d->mps[ctx];
		d->state[ctx] = cabac_trans_lps[s];
	} else {
		bin = d->mps[ctx];
		d->state[ctx] = cabac_trans_mps[s];
	}
	renorm(d);
#ifdef ORACLE_CABAC_TRACE
	if (getenv("ORACLE_CABAC_BINS"))
		fprintf(stderr, "S %u %u\n", ctx, bin);
#endif
	return bin;
}

/* 9.3.3.2.3 DecodeBypass */
static uint32_t db(struct dec *d)
{
	d->offset = (d->offset << 1) | rd_bit(d);
	if (d->offset >= d->range) {
		d->offset -= d->range;
		return 1;
	}
	return 0;
}

/* 9.3.3.2.2.3 DecodeTerminate */
static uint32_t dt(struct dec *d)
{
	d->range -= 2;
	if (d->offset >= d->range)
		return 1; /* no renormalisation; the bits read so far end with the encoder's last flush bit */
	renorm(d);
	return 0;
}

/* ---- per-macroblock information the context selection of later macroblocks needs ---------- */
struct mbinfo {
	uint8_t avail;       /* decoded in this slice */
	uint8_t skip;
	uint8_t mb_type;     /* T_* */
	uint8_t intra, pcm, i16, inxn;
	uint8_t direct16;    /* B_Skip or B_Direct_16x16 */
	uint8_t cbp_luma, cbp_chroma;
	uint8_t chroma_pred;
	uint8_t t8;
	uint8_t cbf_luma[16]; /* per 4x4 block (z order): coded_block_flag (AC or 4x4 or its 8x8 block) */
	uint8_t cbf_dc[3];    /* Intra16x16 DC, Cb DC, Cr DC */
	uint8_t cbf_cac[2][8];
	int8_t ref[2][4];     /* per 8x8 quadrant; -1 = not used (predFlagLX = 0) */
	uint8_t direct8[4];   /* quadrant predicted in direct mode (B_Direct_8x8, B_Direct_16x16, B_Skip) */
	uint16_t amvd[2][16][2]; /* |mvd| per 4x4 block (raster index 4 * y + x), list, comp */
};

struct sl {
	struct dec d;
	const struct h264gpu_slice_params *sp;
	struct mbinfo *mb; /* PicSizeInMbs entries */
	uint32_t W, cat, slice_type;
	uint32_t cur;
	int prev_qp_delta_nonzero; /* of the previous macroblock in decoding order (9.3.3.1.1.5) */
	uint64_t hash;
	struct mbinfo *c;  /* current */
	const struct mbinfo *A, *B; /* neighbours or NULL */
};

static void hadd(struct sl *s, uint32_t field, uint32_t idx, int64_t v)
{
	if (v != 0)
		s->hash += h264gpu_mb_hash_term(field, idx, v);
#ifdef ORACLE_CABAC_TRACE
	if (v != 0 && getenv("ORACLE_CABAC_TRACE") && s->cur == (uint32_t)atoi(getenv("ORACLE_CABAC_TRACE")))
		fprintf(stderr, "mb %u f %u i %u v %lld\n", s->cur, field, idx, (long long)v);
#endif
}

/* 6.4.3 inverse 4x4 luma block scan: z-order index -> (x, y) in blocks */
static const uint8_t blk_x[16] = {0, 1, 0, 1, 2, 3, 2, 3, 0, 1, 0, 1, 2, 3, 2, 3};
static const uint8_t blk_y[16] = {0, 0, 1, 1, 0, 0, 1, 1, 2, 2, 3, 3, 2, 2, 3, 3};
static const uint8_t blk_of_xy[4][4] = {{0, 1, 4, 5}, {2, 3, 6, 7}, {8, 9, 12, 13}, {10, 11, 14, 15}}; /* [y][x] */

/* ---- binarisations + ctxIdx (9.3.2, 9.3.3.1) ---------------------------------------------- */

/* mb_skip_flag: 9.3.3.1.1.1 */
static uint32_t rd_mb_skip(struct sl *s)
{
	const uint32_t inc = (s->A && !s->A->skip) + (s->B && !s->B->skip);
	return dd(&s->d, (s->slice_type == SL_B ? 24 : 11) + inc);
}

/* the I-slice mb_type binarisation (Table 9-36) with ctxIdxOffset 3 (prefix = 0) or as the
 * suffix of P / B slices with ctxIdxOffset 17 (Table 9-39: ctxIdx by binIdx, b3 dependent) */
static uint32_t rd_mb_type_intra(struct sl *s, int suffix)
{
	struct dec *d = &s->d;
	if (!suffix) {
		/* 9.3.3.1.1.3: condTermFlagN = 0 when N is unavailable or I_NxN (SI: not handled) */
		const uint32_t inc = (s->A && s->A->mb_type != T_I_NxN) + (s->B && s->B->mb_type != T_I_NxN);
		if (!dd(d, 3 + inc))
			return 0; /* I_NxN */
		if (dt(d))
			return 25; /* I_PCM */
		uint32_t t = 1;
		t += 12 * dd(d, 3 + 3);           /* binIdx 2: luma coded */
		const uint32_t b3 = dd(d, 3 + 4); /* binIdx 3: chroma != 0 */
		if (b3)
			t += 4 + 4 * dd(d, 3 + 5);  /* binIdx 4 (b3 != 0): chroma == 2 */
		t += 2 * dd(d, 3 + 6);            /* pred mode high bit: binIdx 4 (b3 == 0) or 5 (b3 != 0) */
		t += dd(d, 3 + 7);                /* low bit: binIdx 5 or 6 */
		return t;
	}
	const uint32_t o = s->slice_type == SL_P ? 17 : 32; /* Table 9-34: suffix of P / SP, of B slices */
	if (!dd(d, o))
		return 0;
	if (dt(d))
		return 25;
	uint32_t t = 1;
	t += 12 * dd(d, o + 1);
	const uint32_t b3 = dd(d, o + 2);
	if (b3)
		t += 4 + 4 * dd(d, o + 2); /* binIdx 4, b3 != 0: ctxIdxInc 2 */
	t += 2 * dd(d, o + 3);
	t += dd(d, o + 3);
	return t;
}

/* raw mb_type in the numbering of Tables 7-11 / 7-13 / 7-14 (inter types first, then 5 + / 23 + intra) */
static uint32_t rd_mb_type(struct sl *s)
{
	struct dec *d = &s->d;
	if (s->slice_type == SL_I)
		return rd_mb_type_intra(s, 0);
	if (s->slice_type == SL_P) {
		/* Table 9-37: 0 0 0 P_L0_16x16, 0 1 1 P_L0_L0_16x8, 0 1 0 P_L0_L0_8x16, 0 0 1 P_8x8, prefix 1 = intra */
		if (dd(d, 14))
			return 5 + rd_mb_type_intra(s, 1);
		if (!dd(d, 15))
			return dd(d, 16) ? 3 : 0;
		return dd(d, 17) ? 1 : 2;
	}
	/* B, Table 9-37 (second part), ctxIdxOffset 27: binIdx 0: 0..2, 1: 3, 2: 5 - b1 ... */
	const uint32_t inc = (s->A && !s->A->direct16) + (s->B && !s->B->direct16);
	if (!dd(d, 27 + inc))
		return 0; /* B_Direct_16x16 */
	if (!dd(d, 27 + 3))
		return 1 + dd(d, 27 + 5); /* B_L0_16x16, B_L1_16x16 */
	uint32_t bits = dd(d, 27 + 4) << 3;
	bits |= dd(d, 27 + 5) << 2;
	bits |= dd(d, 27 + 5) << 1;
	bits |= dd(d, 27 + 5);
	if (bits < 8)
		return bits + 3;
	if (bits == 13)
		return 23 + rd_mb_type_intra(s, 1);
	if (bits == 14)
		return 11;
	if (bits == 15)
		return 22;
	bits = (bits << 1) | dd(d, 27 + 5);
	return bits - 4;
}

static uint32_t rd_sub_mb_type(struct sl *s)
{
	struct dec *d = &s->d;
	if (s->slice_type == SL_P) {
		if (dd(d, 21))
			return 0;
		if (!dd(d, 22))
			return 1;
		return dd(d, 23) ? 2 : 3;
	}
	if (!dd(d, 36))
		return 0;
	if (!dd(d, 37))
		return 1 + dd(d, 39);
	uint32_t t = 3;
	if (dd(d, 38)) {
		if (dd(d, 39))
			return 11 + dd(d, 39);
		t += 4;
	}
	t += 2 * dd(d, 39);
	t += dd(d, 39);
	return t;
}

/* the 8x8 quadrant / macroblock that holds the 4x4 block left of / above block (x, y) */
static const struct mbinfo *nb_left(const struct sl *s, int x, int y, int *nx, int *ny)
{
	*ny = y;
	if (x > 0) {
		*nx = x - 1;
		return s->c;
	}
	*nx = 3;
	return s->A;
}
static const struct mbinfo *nb_up(const struct sl *s, int x, int y, int *nx, int *ny)
{
	*nx = x;
	if (y > 0) {
		*ny = y - 1;
		return s->c;
	}
	*ny = 3;
	return s->B;
}

/* ref_idx_lX of the partition whose top-left 4x4 block is (x, y): 9.3.3.1.1.6 */
static uint32_t rd_ref_idx(struct sl *s, int list, int x, int y)
{
	int nx, ny;
	uint32_t inc = 0;
	const struct mbinfo *n = nb_left(s, x, y, &nx, &ny);
	if (n && !n->skip && !n->intra && !n->direct8[(ny >> 1) * 2 + (nx >> 1)] && n->ref[list][(ny >> 1) * 2 + (nx >> 1)] > 0)
		inc += 1;
	n = nb_up(s, x, y, &nx, &ny);
	if (n && !n->skip && !n->intra && !n->direct8[(ny >> 1) * 2 + (nx >> 1)] && n->ref[list][(ny >> 1) * 2 + (nx >> 1)] > 0)
		inc += 2;
	uint32_t v = 0;
	if (!dd(&s->d, 54 + inc))
		return 0;
	v = 1;
	if (!dd(&s->d, 54 + 4))
		return 1;
	v = 2;
	while (dd(&s->d, 54 + 5) && v < 64)
		v++;
	return v;
}

/* mvd_lX[..][comp] of the partition whose top-left block is (x, y): 9.3.3.1.1.7, UEG3 uCoff 9 signed */
static int32_t rd_mvd(struct sl *s, int list, int comp, int x, int y)
{
	int nx, ny;
	uint32_t sum = 0;
	const struct mbinfo *n = nb_left(s, x, y, &nx, &ny);
	if (n)
		sum += n->amvd[list][ny * 4 + nx][comp];
	n = nb_up(s, x, y, &nx, &ny);
	if (n)
		sum += n->amvd[list][ny * 4 + nx][comp];
	const uint32_t base = comp ? 47 : 40;
	const uint32_t inc = sum < 3 ? 0 : sum > 32 ? 2 : 1;
	struct dec *d = &s->d;
	if (!dd(d, base + inc))
		return 0;
	uint32_t v = 1;
	uint32_t ctx = 3;
	while (v < 9 && dd(d, base + ctx)) {
		v++;
		if (ctx < 6)
			ctx++;
	}
	if (v >= 9) {
		/* suffix: 3rd order Exp-Golomb, bypass */
		int k = 3;
		while (db(d) && k < 30) {
			v += 1u << k;
			k++;
		}
		while (k--)
			v += db(d) << k;
	}
	return db(d) ? -(int32_t)v : (int32_t)v;
}

/* mb_qp_delta: 9.3.3.1.1.5, unary, mapped as Table 9-3 */
static int32_t rd_qp_delta(struct sl *s)
{
	struct dec *d = &s->d;
	if (!dd(d, 60 + (s->prev_qp_delta_nonzero ? 1 : 0)))
		return 0;
	uint32_t k = 1;
	if (dd(d, 60 + 2)) {
		k = 2;
		while (dd(d, 60 + 3) && k < 256)
			k++;
	}
	return (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)(k >> 1);
}

/* intra_chroma_pred_mode: 9.3.3.1.1.8, TU cMax 3 */
static uint32_t rd_chroma_pred(struct sl *s)
{
	const uint32_t inc = (s->A && s->A->intra && !s->A->pcm && s->A->chroma_pred != 0) +
			     (s->B && s->B->intra && !s->B->pcm && s->B->chroma_pred != 0);
	struct dec *d = &s->d;
	if (!dd(d, 64 + inc))
		return 0;
	if (!dd(d, 64 + 3))
		return 1;
	return dd(d, 64 + 3) ? 3 : 2;
}

/* coded_block_pattern: 9.3.3.1.1.4; prefix = 4 luma bins, suffix = chroma TU cMax 2 */
static uint32_t rd_cbp(struct sl *s)
{
	struct dec *d = &s->d;
	uint32_t luma = 0;
	for (int b8 = 0; b8 < 4; b8++) {
		const int x = b8 & 1, y = b8 >> 1;
		/* condTermFlagN = ((cbp luma bit of the 8x8 block N) == 0), 0 when N is unavailable or I_PCM;
		 * a skipped macroblock has pattern 0 */
		uint32_t ca, cb;
		if (x > 0) {
			ca = !((luma >> (b8 - 1)) & 1);
		} else {
			const struct mbinfo *n = s->A;
			ca = n && !n->pcm && !((n->cbp_luma >> (b8 + 1)) & 1);
		}
		if (y > 0) {
			cb = !((luma >> (b8 - 2)) & 1);
		} else {
			const struct mbinfo *n = s->B;
			cb = n && !n->pcm && !((n->cbp_luma >> (b8 + 2)) & 1);
		}
		luma |= dd(d, 73 + ca + 2 * cb) << b8;
	}
	uint32_t chroma = 0;
	if (s->cat == 1 || s->cat == 2) {
		uint32_t ca = s->A && (s->A->pcm || s->A->cbp_chroma != 0);
		uint32_t cb = s->B && (s->B->pcm || s->B->cbp_chroma != 0);
		if (dd(d, 77 + ca + 2 * cb)) {
			ca = s->A && (s->A->pcm || s->A->cbp_chroma == 2);
			cb = s->B && (s->B->pcm || s->B->cbp_chroma == 2);
			chroma = 1 + dd(d, 77 + 4 + ca + 2 * cb);
		}
	}
	return luma | chroma << 4;
}

/* ---- residual_block_cabac (7.3.5.3.3) ------------------------------------------------------ */
/* Table 9-43: ctxIdxInc of significant_coeff_flag / last_significant_coeff_flag, ctxBlockCat 5, frame */
static const uint8_t sig8x8_frame[63] = {
	0, 1, 2, 3, 4, 5, 5, 4, 4, 3, 3, 4, 4, 4, 5, 5, 4, 4, 4, 4, 3, 3, 6, 7, 7, 7, 8, 9, 10, 9, 8, 7,
	7, 6, 11, 12, 13, 11, 6, 7, 8, 9, 14, 10, 9, 8, 6, 11, 12, 13, 11, 6, 9, 14, 10, 9, 11, 12, 13, 11, 14, 10, 12};
static const uint8_t last8x8[63] = {
	0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2,
	3, 3, 3, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8};
/* Table 9-40: ctxBlockCatOffset */
static const uint16_t cbf_cat_off[5] = {0, 4, 8, 12, 16};
static const uint16_t sig_cat_off[5] = {0, 15, 29, 44, 47};
static const uint16_t abs_cat_off[5] = {0, 10, 20, 30, 39};

/*
 * cat 0 Intra16x16 DC, 1 Intra16x16 AC, 2 luma 4x4, 3 chroma DC, 4 chroma AC, 5 luma 8x8.
 * cbf_inc: ctxIdxInc of coded_block_flag (unused for cat 5).  Returns the number of non-zero
 * coefficients; coef[] in scan order (0 .. maxNumCoeff - 1).
 */
static int rd_residual_block(struct sl *s, int cat, uint32_t cbf_inc, int max_num, int16_t *coef)
{
	struct dec *d = &s->d;
	memset(coef, 0, sizeof(int16_t) * (size_t)max_num);
	if (cat != 5) {
		if (!dd(d, 85 + cbf_cat_off[cat] + cbf_inc))
			return 0;
	}
	const uint32_t sig_base = cat == 5 ? 402 : 105 + sig_cat_off[cat];
	const uint32_t last_base = cat == 5 ? 417 : 166 + sig_cat_off[cat];
	const uint32_t abs_base = cat == 5 ? 426 : 227 + abs_cat_off[cat];
	const int num_c8x8 = s->cat == 2 ? 2 : 1;
	uint8_t sig[64];
	int num_coeff = max_num, i = 0, n_sig = 0;
	memset(sig, 0, sizeof(sig));
	while (i < num_coeff - 1) {
		uint32_t inc_s, inc_l;
		if (cat == 5) {
			inc_s = sig8x8_frame[i];
			inc_l = last8x8[i];
		} else if (cat == 3) {
			inc_s = inc_l = (uint32_t)(i / num_c8x8 < 2 ? i / num_c8x8 : 2);
		} else {
			inc_s = inc_l = (uint32_t)i;
		}
		sig[i] = (uint8_t)dd(d, sig_base + inc_s);
		if (sig[i]) {
			n_sig++;
			if (dd(d, last_base + inc_l))
				num_coeff = i + 1;
		}
		i++;
	}
	if (num_coeff == max_num) { /* the last coefficient of the block is inferred significant */
		sig[max_num - 1] = 1;
		n_sig++;
	}
	/* levels in reverse scan order: 9.3.3.1.3 */
	uint32_t eq1 = 0, gt1 = 0;
	for (i = num_coeff - 1; i >= 0; i--) {
		if (!sig[i])
			continue;
		const uint32_t inc0 = gt1 != 0 ? 0 : (1 + eq1 < 4 ? 1 + eq1 : 4);
		uint32_t v = 0; /* coeff_abs_level_minus1: UEG0, uCoff 14 */
		if (dd(d, abs_base + inc0)) {
			const uint32_t cap = 4u - (cat == 3 ? 1u : 0u);
			const uint32_t inc1 = 5 + (gt1 < cap ? gt1 : cap);
			v = 1;
			while (v < 14 && dd(d, abs_base + inc1))
				v++;
			if (v >= 14) {
				int k = 0;
				while (db(d) && k < 30) {
					v += 1u << k;
					k++;
				}
				while (k--)
					v += db(d) << k;
			}
		}
		const int32_t mag = (int32_t)v + 1;
		coef[i] = (int16_t)(db(d) ? -mag : mag);
		if (v == 0)
			eq1++;
		else
			gt1++;
	}
	return n_sig;
}

/* coded_block_flag ctxIdxInc from the two neighbouring blocks' flags (9.3.3.1.1.9): an
 * unavailable macroblock counts as coded for an intra macroblock, as not coded for an inter one;
 * I_PCM counts as coded; a macroblock without that block (skipped, pattern bit 0) as not coded */
static uint32_t cbf_term(const struct sl *s, const struct mbinfo *n, int has_block, int flag)
{
	if (n == NULL)
		return s->c->intra ? 1 : 0;
	if (n->pcm)
		return 1;
	if (n->skip || !has_block)
		return 0;
	return (uint32_t)flag;
}

/* ---- one macroblock --------------------------------------------------------------------- */
static const uint8_t b_part_pred[9][2] = {{0, 0}, {1, 1}, {0, 1}, {1, 0}, {0, 2}, {1, 2}, {2, 0}, {2, 1}, {2, 2}};

static void set_ref(struct mbinfo *m, int list, int q0, int q1, int v)
{
	for (int q = q0; q <= q1; q++)
		m->ref[list][q] = (int8_t)v;
}

static int decode_mb(struct sl *s, uint32_t *mb_type_out)
{
	struct dec *d = &s->d;
	struct mbinfo *c = s->c;
	const struct h264gpu_slice_params *sp = s->sp;
	const uint32_t raw = rd_mb_type(s);
	uint32_t t = raw; /* intra type number 0..25 when intra */
	int intra = 0, num_part = 0, pm[2] = {PRED_L0, PRED_L0};
	hadd(s, H264GPU_F_RAW_MB_TYPE, 0, raw);
	memset(c->ref, -1, sizeof(c->ref));
	if (s->slice_type == SL_I) {
		intra = 1;
	} else if (s->slice_type == SL_P) {
		if (raw >= 5) {
			intra = 1;
			t = raw - 5;
		} else if (raw == 0) {
			c->mb_type = T_P_16x16;
			num_part = 1;
		} else if (raw <= 2) {
			c->mb_type = raw == 1 ? T_P_16x8 : T_P_8x16;
			num_part = 2;
		} else {
			c->mb_type = T_P_8x8;
			num_part = 4;
		}
	} else {
		if (raw >= 23) {
			intra = 1;
			t = raw - 23;
		} else if (raw == 0) {
			c->mb_type = T_B_Direct_16x16;
			c->direct16 = 1;
			memset(c->direct8, 1, 4);
			pm[0] = PRED_DIRECT;
			num_part = 1;
		} else if (raw <= 3) {
			c->mb_type = T_B_16x16;
			num_part = 1;
			pm[0] = (int)raw - 1;
		} else if (raw <= 21) {
			c->mb_type = ((raw - 4) & 1) ? T_B_8x16 : T_B_16x8;
			num_part = 2;
			pm[0] = b_part_pred[(raw - 4) >> 1][0];
			pm[1] = b_part_pred[(raw - 4) >> 1][1];
		} else {
			c->mb_type = T_B_8x8;
			num_part = 4;
		}
	}
	uint32_t cbp_luma = 0, cbp_chroma = 0;
	int t8 = 0;
	if (intra) {
		c->intra = 1;
		if (t == 0) {
			c->mb_type = T_I_NxN;
			c->inxn = 1;
		} else if (t <= 24) {
			c->mb_type = T_I_16x16;
			c->i16 = 1;
			hadd(s, H264GPU_F_I16_PRED_MODE, 0, (t - 1) % 4);
			cbp_luma = t <= 12 ? 0 : 15;
			cbp_chroma = ((t - 1) / 4) % 3;
		} else if (t == 25) {
			c->mb_type = T_I_PCM;
			c->pcm = 1;
		} else {
			return -EIO;
		}
	}
	*mb_type_out = c->mb_type;

	if (c->pcm) {
		/* pcm_alignment_zero_bit, samples, then the engine starts again (9.3.1.2) */
		while (d->bitpos & 7)
			if (rd_bit(d))
				return -EIO;
		for (uint32_t i = 0; i < 256; i++)
			hadd(s, H264GPU_F_PCM_LUMA, i, rd_bits(d, sp->bit_depth_luma));
		const uint32_t nc = s->cat == 0 ? 0 : s->cat == 1 ? 64 : s->cat == 2 ? 128 : 256;
		for (uint32_t ic = 0; ic < 2; ic++)
			for (uint32_t i = 0; i < nc; i++)
				hadd(s, H264GPU_F_PCM_CHROMA, ic * 256 + i, rd_bits(d, sp->bit_depth_chroma));
		init_engine(d);
		c->cbp_luma = 15;
		c->cbp_chroma = 2;
		memset(c->cbf_luma, 1, sizeof(c->cbf_luma));
		memset(c->cbf_dc, 1, sizeof(c->cbf_dc));
		memset(c->cbf_cac, 1, sizeof(c->cbf_cac));
		s->prev_qp_delta_nonzero = 0;
		return 0;
	}

	int no_sub_lt8 = 1;
	const uint32_t max0 = sp->num_ref_idx_l0_active_minus1, max1 = sp->num_ref_idx_l1_active_minus1;
	if (!intra && num_part == 4) {
		/* sub_mb_pred */
		uint32_t sub[4], nsub[4];
		int spm[4];
		static const uint8_t p_nsub[4] = {1, 2, 2, 4};
		/* Table 7-18: B sub-macroblock types: parts and prediction */
		static const uint8_t b_nsub[13] = {4, 1, 1, 1, 2, 2, 2, 2, 2, 2, 4, 4, 4};
		static const uint8_t b_pred[13] = {PRED_DIRECT, 0, 1, 2, 0, 0, 1, 1, 2, 2, 0, 1, 2};
		/* shape of the sub-partitions: 0 8x8, 1 8x4, 2 4x8, 3 4x4 */
		static const uint8_t b_shape[13] = {3, 0, 0, 0, 1, 2, 1, 2, 1, 2, 3, 3, 3};
		int shape[4];
		for (int i = 0; i < 4; i++) {
			sub[i] = rd_sub_mb_type(s);
			hadd(s, H264GPU_F_RAW_SUB_MB_TYPE, (uint32_t)i, sub[i]);
			if (s->slice_type == SL_P) {
				nsub[i] = p_nsub[sub[i]];
				spm[i] = PRED_L0;
				shape[i] = (int)sub[i];
			} else {
				nsub[i] = b_nsub[sub[i]];
				spm[i] = b_pred[sub[i]];
				shape[i] = b_shape[sub[i]];
			}
			c->direct8[i] = spm[i] == PRED_DIRECT;
			if (spm[i] != PRED_DIRECT) {
				if (nsub[i] > 1)
					no_sub_lt8 = 0;
			} else if (!sp->direct_8x8_inference_flag) {
				no_sub_lt8 = 0;
			}
		}
		/* ref_idx: all of list 0 first, written into the info as they are read (a later
		 * partition's context sees the earlier ones) */
		for (int list = 0; list < 2; list++) {
			const uint32_t mx = list ? max1 : max0;
			for (int i = 0; i < 4; i++) {
				const int used = spm[i] != PRED_DIRECT && spm[i] != (list ? PRED_L0 : PRED_L1);
				if (!used)
					continue;
				uint32_t v = 0;
				if (mx > 0)
					v = rd_ref_idx(s, list, (i & 1) * 2, (i >> 1) * 2);
				c->ref[list][i] = (int8_t)v;
				hadd(s, list ? H264GPU_F_REF_IDX_L1 : H264GPU_F_REF_IDX_L0, (uint32_t)i, (uint8_t)v);
			}
		}
		for (int list = 0; list < 2; list++) {
			for (int i = 0; i < 4; i++) {
				const int used = spm[i] != PRED_DIRECT && spm[i] != (list ? PRED_L0 : PRED_L1);
				if (!used)
					continue;
				for (uint32_t j = 0; j < nsub[i]; j++) {
					/* top-left block of sub-partition j, and its extent in blocks */
					int bx = (i & 1) * 2, by = (i >> 1) * 2, w = 2, h = 2;
					if (shape[i] == 1) {
						h = 1;
						by += (int)j;
					} else if (shape[i] == 2) {
						w = 1;
						bx += (int)j;
					} else if (shape[i] == 3) {
						w = h = 1;
						bx += (int)(j & 1);
						by += (int)(j >> 1);
					}
					for (int comp = 0; comp < 2; comp++) {
						const int32_t v = rd_mvd(s, list, comp, bx, by);
						hadd(s, list ? H264GPU_F_MVD_L1 : H264GPU_F_MVD_L0, (uint32_t)((i * 4 + (int)j) * 2 + comp),
						     (int16_t)v);
						const uint16_t a = (uint16_t)(v < 0 ? -v : v);
						for (int yy = by; yy < by + h; yy++)
							for (int xx = bx; xx < bx + w; xx++)
								c->amvd[list][yy * 4 + xx][comp] = a;
					}
				}
			}
		}
	} else {
		if (sp->transform_8x8_mode_flag && c->inxn) {
			const uint32_t inc = (s->A && s->A->t8) + (s->B && s->B->t8);
			t8 = (int)dd(d, 399 + inc);
		}
		if (intra) {
			if (c->inxn) {
				const int n = t8 ? 4 : 16;
				for (int i = 0; i < n; i++) {
					int64_t v = -1;
					if (!dd(d, 68)) {
						uint32_t r = dd(d, 69);
						r |= dd(d, 69) << 1;
						r |= dd(d, 69) << 2;
						v = r;
					}
					hadd(s, t8 ? H264GPU_F_INTRA8X8_PRED_MODE : H264GPU_F_INTRA4X4_PRED_MODE, (uint32_t)i, v);
				}
			}
			if (s->cat == 1 || s->cat == 2) {
				c->chroma_pred = (uint8_t)rd_chroma_pred(s);
				hadd(s, H264GPU_F_INTRA_CHROMA_PRED_MODE, 0, c->chroma_pred);
			}
		} else if (pm[0] != PRED_DIRECT) {
			/* partition geometry in blocks: 16x16, 16x8 (two rows), 8x16 (two columns) */
			const int is16x8 = c->mb_type == T_P_16x8 || c->mb_type == T_B_16x8;
			for (int list = 0; list < 2; list++) {
				const uint32_t mx = list ? max1 : max0;
				for (int i = 0; i < num_part; i++) {
					if (pm[i] == (list ? PRED_L0 : PRED_L1))
						continue;
					const int bx = (num_part == 2 && !is16x8) ? 2 * i : 0, by = (num_part == 2 && is16x8) ? 2 * i : 0;
					uint32_t v = 0;
					if (mx > 0)
						v = rd_ref_idx(s, list, bx, by);
					if (num_part == 1)
						set_ref(c, list, 0, 3, (int)v);
					else if (is16x8)
						set_ref(c, list, 2 * i, 2 * i + 1, (int)v);
					else {
						c->ref[list][i] = (int8_t)v;
						c->ref[list][i + 2] = (int8_t)v;
					}
					hadd(s, list ? H264GPU_F_REF_IDX_L1 : H264GPU_F_REF_IDX_L0, (uint32_t)i, (uint8_t)v);
				}
			}
			for (int list = 0; list < 2; list++) {
				for (int i = 0; i < num_part; i++) {
					if (pm[i] == (list ? PRED_L0 : PRED_L1))
						continue;
					const int bx = (num_part == 2 && !is16x8) ? 2 * i : 0, by = (num_part == 2 && is16x8) ? 2 * i : 0;
					const int w = (num_part == 2 && !is16x8) ? 2 : 4, h = (num_part == 2 && is16x8) ? 2 : 4;
					for (int comp = 0; comp < 2; comp++) {
						const int32_t v = rd_mvd(s, list, comp, bx, by);
						hadd(s, list ? H264GPU_F_MVD_L1 : H264GPU_F_MVD_L0, (uint32_t)((i * 4) * 2 + comp), (int16_t)v);
						const uint16_t a = (uint16_t)(v < 0 ? -v : v);
						for (int yy = by; yy < by + h; yy++)
							for (int xx = bx; xx < bx + w; xx++)
								c->amvd[list][yy * 4 + xx][comp] = a;
					}
				}
			}
		}
	}

	if (!c->i16) {
		const uint32_t cbp = rd_cbp(s);
		cbp_luma = cbp & 15;
		cbp_chroma = cbp >> 4;
		hadd(s, H264GPU_F_CBP, 0, cbp_luma + 16 * cbp_chroma);
		if (cbp_luma > 0 && sp->transform_8x8_mode_flag && !c->inxn && no_sub_lt8 &&
		    (c->mb_type != T_B_Direct_16x16 || sp->direct_8x8_inference_flag)) {
			const uint32_t inc = (s->A && s->A->t8) + (s->B && s->B->t8);
			t8 = (int)dd(d, 399 + inc);
		}
	}
	c->cbp_luma = (uint8_t)cbp_luma;
	c->cbp_chroma = (uint8_t)cbp_chroma;
	c->t8 = (uint8_t)t8;
	hadd(s, H264GPU_F_TRANSFORM_8X8, 0, t8);
	hadd(s, H264GPU_F_CBP_LUMA, 0, cbp_luma);
	hadd(s, H264GPU_F_CBP_CHROMA, 0, cbp_chroma);

	if (!(cbp_luma > 0 || cbp_chroma > 0 || c->i16)) {
		s->prev_qp_delta_nonzero = 0; /* no mb_qp_delta in this macroblock: inferred 0 */
	} else {
		const int32_t qpd = rd_qp_delta(s); /* context from the macroblock before */
		hadd(s, H264GPU_F_MB_QP_DELTA, 0, qpd);
		s->prev_qp_delta_nonzero = qpd != 0;

		/* residual( 0, 15 ) */
		int16_t coef[64];
		if (c->i16) {
			const uint32_t inc = cbf_term(s, s->A, s->A && s->A->i16, s->A ? s->A->cbf_dc[0] : 0) +
					     2 * cbf_term(s, s->B, s->B && s->B->i16, s->B ? s->B->cbf_dc[0] : 0);
			c->cbf_dc[0] = rd_residual_block(s, 0, inc, 16, coef) > 0;
			for (int i = 0; i < 16; i++)
				hadd(s, H264GPU_F_I16_DC, (uint32_t)i, coef[i]);
		}
		for (int b8 = 0; b8 < 4; b8++) {
			if (!(cbp_luma & (1u << b8)))
				continue;
			if (t8) {
				/* one 8x8 block, coded_block_flag inferred 1; its 64 coefficients interleave
				 * into the four 4x4 blocks the CAVLC syntax would carry (7.3.5.3.2) */
				rd_residual_block(s, 5, 0, 64, coef);
				for (int j = 0; j < 64; j++)
					hadd(s, H264GPU_F_LEVEL4X4, (uint32_t)((b8 * 4 + (j & 3)) * 16 + (j >> 2)), coef[j]);
				for (int k = 0; k < 4; k++)
					c->cbf_luma[b8 * 4 + k] = 1;
				continue;
			}
			for (int k = 0; k < 4; k++) {
				const int blk = b8 * 4 + k, x = blk_x[blk], y = blk_y[blk];
				int nx, ny;
				const struct mbinfo *n = nb_left(s, x, y, &nx, &ny);
				int nb = blk_of_xy[ny][nx];
				const uint32_t ia = cbf_term(s, n, n && ((n->cbp_luma >> (nb >> 2)) & 1), n ? n->cbf_luma[nb] : 0);
				n = nb_up(s, x, y, &nx, &ny);
				nb = blk_of_xy[ny][nx];
				const uint32_t ib = cbf_term(s, n, n && ((n->cbp_luma >> (nb >> 2)) & 1), n ? n->cbf_luma[nb] : 0);
				const int nz = rd_residual_block(s, c->i16 ? 1 : 2, ia + 2 * ib, c->i16 ? 15 : 16, coef);
				c->cbf_luma[blk] = nz > 0;
				for (int i = 0; i < (c->i16 ? 15 : 16); i++)
					hadd(s, c->i16 ? H264GPU_F_I16_AC : H264GPU_F_LEVEL4X4, (uint32_t)(blk * 16 + i), coef[i]);
			}
		}
		if (s->cat == 1 || s->cat == 2) {
			const int num_c8x8 = s->cat == 2 ? 2 : 1;
			if (cbp_chroma & 3) {
				for (int ic = 0; ic < 2; ic++) {
					const uint32_t inc = cbf_term(s, s->A, s->A && s->A->cbp_chroma != 0, s->A ? s->A->cbf_dc[1 + ic] : 0) +
							     2 * cbf_term(s, s->B, s->B && s->B->cbp_chroma != 0, s->B ? s->B->cbf_dc[1 + ic] : 0);
					c->cbf_dc[1 + ic] = rd_residual_block(s, 3, inc, 4 * num_c8x8, coef) > 0;
					for (int i = 0; i < 4 * num_c8x8; i++)
						hadd(s, H264GPU_F_CHROMA_DC, (uint32_t)(ic * 16 + i), coef[i]);
				}
			}
			if (cbp_chroma & 2) {
				const int rows = 2 * num_c8x8; /* chroma block rows (2 columns) */
				for (int ic = 0; ic < 2; ic++) {
					for (int blk = 0; blk < 4 * num_c8x8; blk++) {
						const int x = blk & 1, y = blk >> 1;
						const struct mbinfo *n = x > 0 ? c : s->A;
						const int nbA = x > 0 ? blk - 1 : 2 * y + 1;
						const uint32_t ia = cbf_term(s, n, n && n->cbp_chroma == 2, n ? n->cbf_cac[ic][nbA] : 0);
						n = y > 0 ? c : s->B;
						const int nbB = y > 0 ? blk - 2 : 2 * (rows - 1) + x;
						const uint32_t ib = cbf_term(s, n, n && n->cbp_chroma == 2, n ? n->cbf_cac[ic][nbB] : 0);
						c->cbf_cac[ic][blk] = rd_residual_block(s, 4, ia + 2 * ib, 15, coef) > 0;
						for (int i = 0; i < 15; i++)
							hadd(s, H264GPU_F_CHROMA_AC, (uint32_t)((ic * 16 + blk) * 16 + i), coef[i]);
					}
				}
			}
		}
	}
	return d->overrun ? -EIO : 0;
}

int oracle_cabac_spec_decode(const uint8_t *stream, uint64_t stream_len, const struct h264gpu_slice_params *params,
			     uint32_t n_slices, struct h264gpu_mb_record *records, uint64_t n_records,
			     struct h264gpu_slice_result *results)
{
	for (uint32_t i = 0; i < n_slices; i++) {
		const struct h264gpu_slice_params *sp = &params[i];
		struct h264gpu_slice_result *res = &results[i];
		res->status = 0;
		res->mb_count = 0;
		res->end_bit = 0;
		if (!sp->entropy_coding_mode_flag) {
			res->status = H264GPU_SLICE_SKIPPED;
			continue;
		}
		if (sp->mbaff_frame_flag || sp->field_pic_flag || sp->num_slice_groups_minus1 != 0 ||
		    sp->pic_width_in_mbs == 0 || sp->chroma_array_type == 3 || sp->slice_type > SL_I) {
			res->status = -ENOSYS;
			continue;
		}
		if (sp->nal_off + sp->nal_len > stream_len) {
			res->status = -EINVAL;
			continue;
		}
		struct rbsp r;
		if (unescape(stream + sp->nal_off, sp->nal_len, &r) < 0)
			return -ENOMEM;
		struct sl *s = calloc(1, sizeof(*s));
		const uint32_t pic_size = (uint32_t)sp->pic_width_in_mbs * sp->pic_height_in_mbs;
		s->mb = calloc(pic_size ? pic_size : 1, sizeof(struct mbinfo));
		s->sp = sp;
		s->W = sp->pic_width_in_mbs;
		s->cat = sp->chroma_array_type;
		s->slice_type = sp->slice_type;
		s->d.r = &r;
		/* slice_data() starts at raw bit data_bit_off: in RBSP coordinates, then cabac_alignment_one_bit */
		uint64_t byte = 0;
		while (byte < r.n && r.raw[byte] < (sp->data_bit_off >> 3))
			byte++;
		s->d.bitpos = byte * 8 + (sp->data_bit_off & 7);
		while (s->d.bitpos & 7)
			rd_bit(&s->d);
		init_contexts(&s->d, (int)sp->slice_type, sp->cabac_init_idc, sp->slice_qp);
		init_engine(&s->d);
		uint32_t cur = sp->first_mb_in_slice, count = 0;
		int status = 0;
		for (;;) {
			if (count >= sp->mb_out_cap || cur >= pic_size || sp->mb_out_off + (uint64_t)count >= n_records) {
				status = -ENOBUFS;
				break;
			}
			struct mbinfo *c = &s->mb[cur];
			memset(c, 0, sizeof(*c));
			memset(c->ref, -1, sizeof(c->ref));
			c->avail = 1;
			s->c = c;
			s->cur = cur;
			s->A = (cur % s->W != 0 && cur >= sp->first_mb_in_slice + 1 && s->mb[cur - 1].avail) ? &s->mb[cur - 1] : NULL;
			s->B = (cur >= sp->first_mb_in_slice + s->W && s->mb[cur - s->W].avail) ? &s->mb[cur - s->W] : NULL;
			s->hash = 0;
			uint32_t mb_type = 0;
			int skipped = 0;
			if (s->slice_type != SL_I)
				skipped = (int)rd_mb_skip(s);
			if (skipped) {
				c->skip = 1;
				c->mb_type = s->slice_type == SL_B ? T_B_SKIP : T_P_SKIP;
				c->direct16 = s->slice_type == SL_B;
				if (s->slice_type == SL_B)
					memset(c->direct8, 1, 4);
				else
					set_ref(c, 0, 0, 3, 0);
				mb_type = c->mb_type;
				s->prev_qp_delta_nonzero = 0;
			} else {
				const int e = decode_mb(s, &mb_type);
				if (e < 0) {
					status = e;
					break;
				}
			}
			struct h264gpu_mb_record *rec = &records[sp->mb_out_off + count];
			rec->mb_addr = cur;
			rec->mb_type = mb_type;
			rec->hash = s->hash;
			count++;
			cur++;
			if (s->d.overrun) {
				status = -EIO;
				break;
			}
			if (dt(&s->d)) /* end_of_slice_flag */
				break;
		}
		res->status = status;
		res->mb_count = count;
		/* raw bit position after the last bit read: RBSP bit -> NAL bit */
		{
			const uint64_t b = s->d.bitpos >> 3;
			res->end_bit = (b < r.n ? r.raw[b] : (uint64_t)sp->nal_len) * 8 + (s->d.bitpos & 7);
		}
		free(s->mb);
		free(s);
		free(r.b);
		free(r.raw);
	}
	return 0;
}
