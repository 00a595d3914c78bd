/*
 * cabac_syntax.h — CABAC slice data (Rec. ITU-T H.264 clauses 7.3.4, 7.3.5, 9.3) as ONE
 * bidirectional walk: the same code decodes (GPU kernel, CPU checker) or encodes (the
 * synthetic-stream generator), depending on the coder type it is instantiated with.
 *
 *     c.bin(ctxIdx, v)   decision with context:  encoder codes v and returns it,
 *     c.byp(v)           bypass                  decoder ignores v and returns the bin
 *     c.term(v)          terminate
 *
 * What the reference provides and what it does not: libh264 has NO
 * CABAC decoder — its reader returns before CABAC slice data (src/h264_syntax_slice_data.h:
 * 715-717).  From the reference: context init tables (src/h264_cabac_ctx_tables.c), range /
 * transition tables and the ENCODER engine (src/h264_bac.c:46-358), ctxIdx + binarisation
 * of mb_type, mb_skip_flag, intra_chroma_pred_mode, mb_qp_delta, coded_block_flag,
 * end_of_slice_flag (src/h264_cabac.c:55-975), and the entropy-agnostic macroblock-layer
 * control flow (src/h264_syntax_slice_data.h:422-696).  Everything else here follows the
 * standard's clause 9.3 and is NOT pinned by the reference ("parity unpinned").
 *
 * Scope: frame or field pictures without MBAFF, one slice group, ChromaArrayType 0/1/2.
 */
#ifndef CABAC_SYNTAX_H
#define CABAC_SYNTAX_H

#include <stdint.h>

#include "h264gpu_slice.h"

#ifndef CABAC_HD
#define CABAC_HD
#endif
#ifndef CABAC_OUTLINE
#define CABAC_OUTLINE CABAC_HD
#endif

namespace cabac {

enum { MB_UNKNOWN = 0, MB_I_NxN, MB_I_16x16, MB_I_PCM, MB_SI, MB_P_16x16, MB_P_16x8, MB_P_8x16,
       MB_P_8x8, MB_P_8x8ref0, MB_P_SKIP, MB_B_Direct_16x16, MB_B_16x16, MB_B_16x8, MB_B_8x16,
       MB_B_8x8, MB_B_SKIP };
enum { PM_I4 = 0, PM_I8, PM_I16, PM_L0, PM_L1, PM_BI, PM_DIRECT };
enum { ST_P = 0, ST_B = 1, ST_I = 2, ST_SP = 3, ST_SI = 4 };

/* additional checksum field: coefficients of an 8x8 block (CABAC keeps them together) */
enum { F_LEVEL8X8 = 30 }; /* + 3*comp : [blk8*64 + i] */

/* ---- what later macroblocks need to know about an earlier one (64 bytes) --------------- */
struct Nb {
	uint8_t valid;        /* written in this slice (slot reuse guard) */
	uint8_t skip, intra, pcm;
	uint8_t not_inxn;     /* mb_type ctxIdxInc term for I slices (type is not I_NxN) */
	uint8_t not_si;       /* ... for SI prefix */
	uint8_t not_bdirect;  /* ... for B slices (not B_Skip / B_Direct_16x16) */
	uint8_t t8;
	uint8_t cbp_luma, cbp_chroma;
	uint8_t chroma_mode_nz;
	uint8_t ref_gt0[2];   /* bit b8: partition uses list X, is not direct, ref_idx > 0 */
	uint8_t pad[3];
	uint32_t cbf;         /* bits 0..15 luma 4x4 (blkIdx), 16 luma DC, 17 Cb DC, 18 Cr DC */
	uint32_t cbf_c;       /* bits 0..7 Cb AC blkIdx, 8..15 Cr AC blkIdx */
	uint8_t mvd_right[2][2][4];  /* [list][comp][row]: |mvd| (capped) of the right column */
	uint8_t mvd_bottom[2][2][4]; /* [list][comp][col]: ... of the bottom row */
	uint8_t pad2[8];
};

/* ---- one macroblock's syntax elements ------------------------------------------------------ */
struct Mb {
	uint32_t mb_type;      /* enum above */
	uint32_t raw_type;     /* mb_type syntax element value of the slice type */
	uint32_t num_part, pm[4];
	uint32_t sub_type[4];
	uint8_t t8;
	uint8_t prev_flag[16]; /* intra 4x4 / 8x8: prev_intra_pred_mode_flag, rem mode */
	uint8_t rem_mode[16];
	uint8_t chroma_mode;
	int8_t ref_idx[2][4];
	int16_t mvd[2][16][2]; /* [list][mbPart*4 + subPart][comp] */
	uint32_t cbp_luma, cbp_chroma;
	int32_t qp_delta;
};

/* every coefficient and PCM sample of one macroblock: what the encoder codes; a decoder fills
 * it only when asked to (CPU checker), the GPU kernel only hashes */
struct MbCoef {
	int16_t dc16[16];       /* Intra16x16 DC, scan order */
	int16_t luma[16][16];   /* 4x4 blocks by blkIdx; Intra16x16 AC uses [0..14] */
	int16_t luma8[4][64];   /* 8x8 blocks (transform_size_8x8_flag) */
	int16_t cdc[2][8];      /* chroma DC per iCbCr */
	int16_t cac[2][8][16];  /* chroma AC per iCbCr, blkIdx; [0..14] */
	uint8_t pcm[768];       /* I_PCM samples: 256 luma, 256 slots Cb, 256 slots Cr */
};

/* 16 4x4 luma blocks: blkIdx (z-order) <-> (x, y) in 4-sample units */
CABAC_HD static inline uint32_t blk_x(uint32_t b) { return (b & 1) | ((b >> 1) & 2); }
CABAC_HD static inline uint32_t blk_y(uint32_t b) { return ((b >> 1) & 1) | ((b >> 2) & 2); }
CABAC_HD static inline uint32_t blk_of(uint32_t x, uint32_t y)
{
	return (x & 1) | ((y & 1) << 1) | ((x & 2) << 1) | ((y & 2) << 2);
}

/* Table 9-43: ctxIdxInc of significant_coeff_flag / last_significant_coeff_flag, 8x8 frame blocks */
#ifndef CABAC_CONST
#define CABAC_CONST static const
#endif
CABAC_CONST uint8_t sig8x8_frame[63] = {0, 1, 2, 3, 4, 5, 5, 4, 4, 3, 3, 4, 4, 4, 5, 5, 4, 4, 4, 4, 3,
					3, 6, 7, 7, 7, 8, 9, 10, 9, 8, 7, 7, 6, 11, 12, 13, 11, 6, 7, 8, 9,
					14, 10, 9, 8, 6, 11, 12, 13, 11, 6, 9, 14, 10, 9, 11, 12, 13, 11, 14, 10, 12};
CABAC_CONST uint8_t sig8x8_field[63] = {0, 1, 1, 2, 2, 3, 3, 4, 5, 6, 7, 7, 7, 8, 4, 5, 6, 9, 10, 10, 8,
					11, 12, 11, 9, 9, 10, 10, 8, 11, 12, 11, 9, 9, 10, 10, 8, 11, 12, 11, 9, 9,
					10, 10, 8, 13, 13, 9, 9, 10, 10, 8, 13, 13, 9, 9, 10, 10, 14, 14, 14, 14, 14};
CABAC_CONST uint8_t last8x8[63] = {0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2,
				   2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3, 3, 3, 4, 4,
				   4, 4, 4, 4, 4, 4, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 8, 8, 8};

/* ctxBlockCat: 0 I16 DC, 1 I16 AC, 2 luma 4x4, 3 chroma DC, 4 chroma AC, 5 luma 8x8 */
CABAC_CONST uint16_t cat_cbf_off[5] = {0, 4, 8, 12, 16};          /* + 85 */
CABAC_CONST uint16_t cat_sig_off[6] = {0, 15, 29, 44, 47, 0};     /* + 105 / 166 (402 / 417 for cat 5) */
CABAC_CONST uint16_t cat_abs_off[6] = {0, 10, 20, 30, 39, 0};     /* + 227 (426 for cat 5) */

/* ---- per-slice walker state ------------------------------------------------------------------- */
template <class Coder> struct Walk {
	Coder c;
	const h264gpu_slice_params *sp;
	Nb *ring;        /* (W + 1) entries */
	uint32_t W, cat; /* PicWidthInMbs, ChromaArrayType */
	uint32_t cur, first;
	bool availA, availB;
	const Nb *nbA, *nbB;
	Nb me;
	uint8_t amvd[2][2][16]; /* |mvd| (capped at 64) per list, component, 4x4 block */
	int last_dqp_nz;        /* previous macroblock in decoding order had mb_qp_delta != 0 */
	uint64_t hash;
	bool err;
	int16_t coef[64];
	MbCoef *mc;             /* encoder: levels to code; decoder: optional sink (may be NULL) */

	CABAC_HD Nb *slot(uint32_t mb) { return ring + mb % (W + 1); }

	CABAC_HD void hadd(uint32_t field, uint32_t idx, int64_t v)
	{
		if (v != 0) {
			const uint64_t key = ((uint64_t)field << 16) | idx;
			hash += (uint64_t)v * (((key + 1) * 0x9E3779B97F4A7C15ull) | 1ull);
		}
	}

	/* ---- 9.3.2.5 mb_type ----------------------------------------------------------- */

	/* intra types 1..24 after "not I_NxN, not I_PCM": bins cbp luma, chroma, pred mode */
	CABAC_HD uint32_t intra16_tail(uint32_t base, bool islice, uint32_t v)
	{
		/* v = 1 + pred + 4*chroma + 12*luma15 */
		const uint32_t want = v - 1;
		const uint32_t luma = c.bin(base + (islice ? 3 : 1), want >= 12);
		const uint32_t chroma_nz = c.bin(base + (islice ? 4 : 2), (want % 12) >= 4);
		uint32_t chroma = 0;
		if (chroma_nz)
			chroma = 1 + c.bin(base + (islice ? 5 : 2), (want % 12) >= 8);
		const uint32_t p1 = c.bin(base + (islice ? 6 : 3), (want >> 1) & 1);
		const uint32_t p0 = c.bin(base + (islice ? 7 : 3), want & 1);
		return 1 + (p1 << 1 | p0) + 4 * chroma + 12 * luma;
	}

	/* intra mb_type 0..25 with ctxIdxOffset base: 3 (I slices, ctxIdxInc from A/B for
	 * bin 0), 17 (P suffix), 32 (B suffix) */
	CABAC_HD uint32_t intra_type(uint32_t base, uint32_t v)
	{
		const bool islice = base == 3;
		uint32_t ctx0 = base;
		if (islice)
			ctx0 += (availA && nbA->not_inxn ? 1u : 0u) + (availB && nbB->not_inxn ? 1u : 0u);
		if (!c.bin(ctx0, v != 0))
			return 0; /* I_NxN */
		if (c.term(v == 25))
			return 25; /* I_PCM */
		return intra16_tail(base, islice, v);
	}

	CABAC_HD uint32_t mb_type_p(uint32_t v)
	{
		if (c.bin(14, v >= 5))
			return 5 + intra_type(17, v - 5);
		if (!c.bin(15, v == 1 || v == 2))
			return c.bin(16, v == 3) ? 3 : 0;
		return c.bin(17, v == 1) ? 1 : 2;
	}

	CABAC_HD uint32_t mb_type_b(uint32_t v)
	{
		const uint32_t ctx0 = 27 + (availA && nbA->not_bdirect ? 1u : 0u) + (availB && nbB->not_bdirect ? 1u : 0u);
		if (!c.bin(ctx0, v != 0))
			return 0;
		if (!c.bin(27 + 3, v >= 3))
			return 1 + c.bin(27 + 5, v == 2);
		/* 4 more bins b2..b5; Table 9-37 */
		uint32_t want;
		if (v <= 10)
			want = v - 3; /* 0..7 */
		else if (v == 11)
			want = 14;
		else if (v == 22)
			want = 15;
		else if (v >= 23)
			want = 13;
		else
			want = (v + 4) >> 1; /* 12..21 -> 8..12 */
		uint32_t bits = c.bin(27 + 4, (want >> 3) & 1) << 3;
		bits |= c.bin(27 + 5, (want >> 2) & 1) << 2;
		bits |= c.bin(27 + 5, (want >> 1) & 1) << 1;
		bits |= c.bin(27 + 5, want & 1);
		if (bits < 8)
			return bits + 3;
		if (bits == 13)
			return 23 + intra_type(32, v - 23);
		if (bits == 14)
			return 11;
		if (bits == 15)
			return 22;
		bits = bits << 1 | c.bin(27 + 5, (v + 4) & 1);
		return bits - 4;
	}

	CABAC_HD uint32_t sub_type_p(uint32_t v)
	{
		if (c.bin(21, v == 0))
			return 0;
		if (!c.bin(22, v != 1))
			return 1;
		return c.bin(23, v == 2) ? 2 : 3;
	}

	CABAC_HD uint32_t sub_type_b(uint32_t v)
	{
		if (!c.bin(36, v != 0))
			return 0;
		if (!c.bin(37, v >= 3))
			return 1 + c.bin(39, v == 2);
		uint32_t type = 3;
		if (c.bin(38, v >= 7)) {
			if (c.bin(39, v >= 11))
				return 11 + c.bin(39, v == 12);
			type += 4;
		}
		type += 2 * c.bin(39, ((v - type) >> 1) & 1);
		type += c.bin(39, (v - type) & 1);
		return type;
	}

	/* ---- neighbour-dependent elements --------------------------------------------- */

	CABAC_HD uint32_t skip_flag(uint32_t v)
	{
		const uint32_t base = sp->slice_type == ST_B ? 24 : 11;
		return c.bin(base + (availA && !nbA->skip ? 1u : 0u) + (availB && !nbB->skip ? 1u : 0u), v);
	}

	CABAC_HD uint32_t chroma_pred_mode(uint32_t v)
	{
		const uint32_t inc = (availA && nbA->chroma_mode_nz ? 1u : 0u) + (availB && nbB->chroma_mode_nz ? 1u : 0u);
		if (!c.bin(64 + inc, v != 0))
			return 0;
		if (!c.bin(64 + 3, v != 1))
			return 1;
		return c.bin(64 + 3, v != 2) ? 3 : 2;
	}

	/* 9.3.3.1.1.4: an unavailable or I_PCM neighbour counts as "all coded" for luma */
	CABAC_HD uint32_t nb_cbp_luma(bool avail, const Nb *n) const
	{
		return !avail || n->pcm ? 0xfu : n->cbp_luma;
	}
	CABAC_HD uint32_t nb_cbp_chroma(bool avail, const Nb *n) const
	{
		return !avail ? 0u : n->pcm ? 2u : n->cbp_chroma;
	}

	CABAC_HD void coded_block_pattern(uint32_t &luma, uint32_t &chroma)
	{
		const uint32_t a = nb_cbp_luma(availA, nbA), b = nb_cbp_luma(availB, nbB);
		const uint32_t want = luma;
		uint32_t v = 0;
		v |= c.bin(73 + !(a & 2) + 2 * !(b & 4), want & 1);
		v |= c.bin(73 + !(v & 1) + 2 * !(b & 8), (want >> 1) & 1) << 1;
		v |= c.bin(73 + !(a & 8) + 2 * !(v & 1), (want >> 2) & 1) << 2;
		v |= c.bin(73 + !(v & 4) + 2 * !(v & 2), (want >> 3) & 1) << 3;
		luma = v;
		if (cat == 1 || cat == 2) {
			const uint32_t ca = nb_cbp_chroma(availA, nbA), cb = nb_cbp_chroma(availB, nbB);
			const uint32_t wantc = chroma;
			uint32_t cc = 0;
			if (c.bin(77 + (ca > 0) + 2 * (cb > 0), wantc != 0))
				cc = 1 + c.bin(77 + 4 + (ca == 2) + 2 * (cb == 2), wantc == 2);
			chroma = cc;
		} else {
			chroma = 0;
		}
	}

	CABAC_HD int32_t qp_delta(int32_t v)
	{
		const uint32_t want = v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v);
		if (!c.bin(60 + (last_dqp_nz ? 1u : 0u), want != 0))
			return 0;
		uint32_t val = 1, ctx = 62;
		while (c.bin(ctx, want > val)) {
			ctx = 63;
			if (++val > 104) { /* 2 * (51 + QpBdOffset max) + slack */
				err = true;
				break;
			}
		}
		return (val & 1) ? (int32_t)((val + 1) >> 1) : -(int32_t)(val >> 1);
	}

	/* ref_idx of the partition whose top-left 8x8 block is b8 (0..3) */
	CABAC_OUTLINE uint32_t ref_idx(uint32_t list, uint32_t b8, uint32_t v)
	{
		const uint32_t x = b8 & 1, y = b8 >> 1;
		uint32_t a, b;
		if (x)
			a = (me.ref_gt0[list] >> (b8 - 1)) & 1;
		else
			a = availA ? (nbA->ref_gt0[list] >> (b8 + 1)) & 1 : 0;
		if (y)
			b = (me.ref_gt0[list] >> (b8 - 2)) & 1;
		else
			b = availB ? (nbB->ref_gt0[list] >> (b8 + 2)) & 1 : 0;
		uint32_t ctx = a + 2 * b, ref = 0;
		while (c.bin(54 + ctx, v > ref)) {
			ref++;
			ctx = (ctx >> 2) + 4;
			if (ref >= 32) {
				err = true;
				break;
			}
		}
		return ref;
	}

	/* sum of the |mvd| of the blocks left of and above the 4x4 block (bx, by) */
	CABAC_HD uint32_t mvd_sum(uint32_t list, uint32_t comp, uint32_t bx, uint32_t by) const
	{
		uint32_t a = 0, b = 0;
		if (bx)
			a = amvd[list][comp][blk_of(bx - 1, by)];
		else if (availA)
			a = nbA->mvd_right[list][comp][by];
		if (by)
			b = amvd[list][comp][blk_of(bx, by - 1)];
		else if (availB)
			b = nbB->mvd_bottom[list][comp][bx];
		return a + b;
	}

	/* 9.3.2.3 UEG3, signedValFlag 1, uCoff 9 */
	CABAC_OUTLINE int32_t mvd_comp(uint32_t list, uint32_t comp, uint32_t bx, uint32_t by, int32_t v)
	{
		const uint32_t sum = mvd_sum(list, comp, bx, by);
		const uint32_t base = comp ? 47 : 40;
		const uint32_t want = v < 0 ? (uint32_t)-v : (uint32_t)v;
		if (!c.bin(base + (sum < 3 ? 0 : sum > 32 ? 2 : 1), want != 0))
			return 0;
		uint32_t mag = 1, ctx = base + 3;
		while (mag < 9 && c.bin(ctx, want > mag)) {
			if (mag < 4)
				ctx++;
			mag++;
		}
		if (mag >= 9) {
			uint32_t k = 3, rest = want >= 9 ? want - 9 : 0;
			while (c.byp(rest >= (1u << k))) {
				mag += 1u << k;
				rest -= 1u << k;
				if (++k > 24) {
					err = true;
					return 0;
				}
			}
			while (k--)
				mag += c.byp((rest >> k) & 1) << k;
		}
		return c.byp(v < 0) ? -(int32_t)mag : (int32_t)mag;
	}

	/* both components of one (sub-)partition covering w x h 4x4 blocks at (bx, by) */
	CABAC_HD void mvd_part(Mb &m, uint32_t list, uint32_t idx, uint32_t bx, uint32_t by, uint32_t w, uint32_t h)
	{
		for (uint32_t comp = 0; comp < 2; comp++) {
			const int32_t v = mvd_comp(list, comp, bx, by, m.mvd[list][idx][comp]);
			m.mvd[list][idx][comp] = (int16_t)v;
			const uint32_t mag = v < 0 ? (uint32_t)-v : (uint32_t)v;
			const uint8_t capped = (uint8_t)(mag > 64 ? 64 : mag);
			for (uint32_t yy = 0; yy < h; yy++)
				for (uint32_t xx = 0; xx < w; xx++)
					amvd[list][comp][blk_of(bx + xx, by + yy)] = capped;
		}
	}

	/* ---- 7.3.5.3.3 residual_block_cabac ------------------------------------------------- */

	/* coded_block_flag ctxIdxInc terms of one neighbour (9.3.3.1.1.9) */
	CABAC_HD uint32_t cbf_term(bool avail, const Nb *n, bool chroma_word, uint32_t bit) const
	{
		if (!avail)
			return me.intra ? 1u : 0u;
		if (n->pcm)
			return 1u;
		return ((chroma_word ? n->cbf_c : n->cbf) >> bit) & 1u;
	}

	/*
	 * One block.  blkcat: ctxBlockCat; n: maxNumCoeff; (bit_a, bit_b): where the left /
	 * upper neighbour block's coded_block_flag lives (in_a / in_b: inside this
	 * macroblock); self_bit: where to record this block's flag.  Returns the number of
	 * non-zero coefficients; coef[0..n) holds the levels in scan order.
	 */
	CABAC_OUTLINE uint32_t residual_block(uint32_t blkcat, uint32_t n, bool chroma_word, bool in_a, uint32_t bit_a,
					 bool in_b, uint32_t bit_b, uint32_t self_bit, uint32_t field, uint32_t idx_base)
	{
		/* levels to code (encoder): already in coef[]; count them */
		uint32_t want_nz = 0;
		int32_t want_last = -1;
		for (uint32_t i = 0; i < n; i++)
			if (Coder::kEncode && coef[i] != 0) {
				want_nz++;
				want_last = (int32_t)i;
			}
		uint32_t coded = 1;
		if (blkcat != 5) {
			const uint32_t ta = in_a ? (((chroma_word ? me.cbf_c : me.cbf) >> bit_a) & 1u)
						 : cbf_term(availA, nbA, chroma_word, bit_a);
			const uint32_t tb = in_b ? (((chroma_word ? me.cbf_c : me.cbf) >> bit_b) & 1u)
						 : cbf_term(availB, nbB, chroma_word, bit_b);
			coded = c.bin(85 + cat_cbf_off[blkcat] + ta + 2 * tb, want_nz != 0);
		}
		if (!Coder::kEncode)
			for (uint32_t i = 0; i < n; i++)
				coef[i] = 0;
		if (!coded)
			return 0;
		if (chroma_word)
			me.cbf_c |= 1u << self_bit;
		else
			me.cbf |= 1u << self_bit;

		const bool field_pic = sp->field_pic_flag != 0;
		const uint32_t sig_base = blkcat == 5 ? (field_pic ? 436 : 402) : (field_pic ? 277 : 105) + cat_sig_off[blkcat];
		const uint32_t last_base = blkcat == 5 ? (field_pic ? 451 : 417) : (field_pic ? 338 : 166) + cat_sig_off[blkcat];
		const uint32_t abs_base = blkcat == 5 ? 426 : 227 + cat_abs_off[blkcat];
		const uint32_t numc8 = cat == 1 ? 1 : 2; /* chroma DC: NumC8x8 */

		/* significance map */
		uint64_t sig = 0;
		uint32_t num = n, i = 0;
		for (; i + 1 < num; i++) {
			uint32_t si, li;
			if (blkcat == 5) {
				si = field_pic ? sig8x8_field[i] : sig8x8_frame[i];
				li = last8x8[i];
			} else if (blkcat == 3) {
				si = li = i / numc8 < 2 ? i / numc8 : 2;
			} else {
				si = li = i;
			}
			if (c.bin(sig_base + si, Coder::kEncode && coef[i] != 0)) {
				sig |= 1ull << i;
				if (c.bin(last_base + li, (int32_t)i == want_last)) {
					num = i + 1;
					break;
				}
			}
		}
		if (num == n && !((sig >> (n - 1)) & 1) && i + 1 >= num)
			sig |= 1ull << (n - 1); /* reached the last position: it is significant */

		/* levels, last to first */
		uint32_t eq1 = 0, gt1 = 0, count = 0;
		for (int32_t k = (int32_t)num - 1; k >= 0; k--) {
			if (!((sig >> k) & 1))
				continue;
			const int32_t lv = coef[k];
			const uint32_t want = Coder::kEncode ? (uint32_t)(lv < 0 ? -lv : lv) - 1 : 0;
			const uint32_t ctx0 = abs_base + (gt1 ? 0 : (1 + eq1 < 4 ? 1 + eq1 : 4));
			uint32_t mag = 0;
			if (c.bin(ctx0, want > 0)) {
				const uint32_t lim = 4 - (blkcat == 3 ? 1 : 0);
				const uint32_t ctx1 = abs_base + 5 + (gt1 < lim ? gt1 : lim);
				mag = 1;
				while (mag < 14 && c.bin(ctx1, want > mag))
					mag++;
				if (mag >= 14) { /* UEG0 suffix */
					uint32_t kk = 0, rest = want >= 14 ? want - 14 : 0;
					while (c.byp(rest >= (1u << kk))) {
						mag += 1u << kk;
						rest -= 1u << kk;
						if (++kk > 24) {
							err = true;
							return count;
						}
					}
					while (kk--)
						mag += c.byp((rest >> kk) & 1) << kk;
				}
				gt1++;
			} else {
				eq1++;
			}
			const uint32_t neg = c.byp(lv < 0);
			const int32_t val = neg ? -(int32_t)(mag + 1) : (int32_t)(mag + 1);
			coef[k] = (int16_t)val;
			/* an 8x8 block is hashed where CAVLC (and so the reference's ctx->mb) keeps it:
			 * coefficient k of block b8 is entry k >> 2 of 4x4 block 4 * b8 + (k & 3) */
			hadd(field, blkcat == 5 ? idx_base + ((uint32_t)k & 3) * 16 + ((uint32_t)k >> 2)
						: idx_base + (uint32_t)k, val);
			count++;
		}
		return count;
	}
	/* ---- remaining flags -------------------------------------------------------------- */

	CABAC_HD uint32_t t8_flag(uint32_t v)
	{
		return c.bin(399 + (availA && nbA->t8 ? 1u : 0u) + (availB && nbB->t8 ? 1u : 0u), v);
	}

	/* prev_intra{4x4,8x8}_pred_mode_flag + rem_intra_pred_mode (3 bins, LSB first) */
	CABAC_HD void intra_pred_mode(uint8_t &flag, uint8_t &rem)
	{
		flag = (uint8_t)c.bin(68, flag);
		if (!flag) {
			uint32_t r = c.bin(69, rem & 1);
			r |= c.bin(69, (rem >> 1) & 1) << 1;
			r |= c.bin(69, (rem >> 2) & 1) << 2;
			rem = (uint8_t)r;
		}
	}

	/* ---- mb_type value -> (enum h264_mb_type, partitions, prediction modes); Table 7-11/13/14,
	 * same result as src/h264_slice_data.c:839-969 ------------------------------------------- */
	CABAC_HD bool classify(Mb &m, bool &i16, uint32_t &i16_pred)
	{
		uint32_t type = m.raw_type;
		const uint32_t st = sp->slice_type;
		bool intra = false;
		m.num_part = 0;
		m.pm[0] = m.pm[1] = m.pm[2] = m.pm[3] = 0;
		m.mb_type = MB_UNKNOWN;
		i16 = false;
		if (st == ST_I) {
			intra = true;
		} else if (st == ST_P || st == ST_SP) {
			if (type == 0) {
				m.mb_type = MB_P_16x16;
				m.num_part = 1;
				m.pm[0] = PM_L0;
			} else if (type <= 2) {
				m.mb_type = type == 1 ? MB_P_16x8 : MB_P_8x16;
				m.num_part = 2;
				m.pm[0] = m.pm[1] = PM_L0;
			} else if (type == 3) {
				m.mb_type = MB_P_8x8;
				m.num_part = 4;
			} else if (type == 4) {
				return false; /* P_8x8ref0 cannot be coded with CABAC */
			} else {
				type -= 5;
				intra = true;
			}
		} else if (st == ST_B) {
			if (type == 0) {
				m.mb_type = MB_B_Direct_16x16;
				m.num_part = 1;
				m.pm[0] = PM_DIRECT;
			} else if (type <= 3) {
				m.mb_type = MB_B_16x16;
				m.num_part = 1;
				m.pm[0] = type == 1 ? PM_L0 : type == 2 ? PM_L1 : PM_BI;
			} else if (type <= 21) {
				/* Table 7-14 rows 4..21: (first, second) partition modes */
				const uint32_t p = (type - 4) >> 1;
				const uint8_t first[9] = {0, 1, 0, 1, 0, 1, 2, 2, 2};
				const uint8_t second[9] = {0, 1, 1, 0, 2, 2, 0, 1, 2};
				m.mb_type = ((type - 4) & 1) ? MB_B_8x16 : MB_B_16x8;
				m.num_part = 2;
				m.pm[0] = PM_L0 + first[p];
				m.pm[1] = PM_L0 + second[p];
			} else if (type == 22) {
				m.mb_type = MB_B_8x8;
				m.num_part = 4;
			} else {
				type -= 23;
				intra = true;
			}
		} else {
			return false;
		}
		if (intra) {
			if (type == 0) {
				m.mb_type = MB_I_NxN;
				m.num_part = 1;
				m.pm[0] = PM_I4;
			} else if (type <= 24) {
				m.mb_type = MB_I_16x16;
				m.num_part = 1;
				m.pm[0] = PM_I16;
				i16 = true;
				i16_pred = (type - 1) % 4;
				m.cbp_luma = type <= 12 ? 0 : 15;
				m.cbp_chroma = ((type - 1) / 4) % 3;
			} else if (type == 25) {
				m.mb_type = MB_I_PCM;
			} else {
				return false;
			}
		}
		return true;
	}

	CABAC_HD void set_ref(uint32_t list, uint32_t mask, uint32_t ref)
	{
		if (ref > 0)
			me.ref_gt0[list] |= (uint8_t)mask;
	}

	/* ---- 7.3.5.3 residual (CABAC branch) ----------------------------------------------- */

	CABAC_HD void blk_in(const int16_t *src, uint32_t n)
	{
		if (Coder::kEncode)
			for (uint32_t i = 0; i < n; i++)
				coef[i] = src[i];
	}
	CABAC_HD void blk_out(int16_t *dst, uint32_t n)
	{
		if (!Coder::kEncode && mc != 0)
			for (uint32_t i = 0; i < n; i++)
				dst[i] = coef[i];
	}

	CABAC_HD void residual(bool i16, uint32_t cbp_luma, uint32_t cbp_chroma, bool t8)
	{
		if (i16) {
			blk_in(mc->dc16, 16);
			residual_block(0, 16, false, false, 16, false, 16, 16, H264GPU_F_I16_DC, 0);
			blk_out(mc->dc16, 16);
		}
		for (uint32_t b8 = 0; b8 < 4 && !err; b8++) {
			if (!((cbp_luma >> b8) & 1))
				continue;
			if (t8) {
				blk_in(mc->luma8[b8], 64);
				residual_block(5, 64, false, false, 0, false, 0, 4 * b8, H264GPU_F_LEVEL4X4, b8 * 64);
				blk_out(mc->luma8[b8], 64);
				me.cbf |= 0xfu << (4 * b8); /* no coded_block_flag: inferred 1 for all four */
				continue;
			}
			for (uint32_t b4 = 0; b4 < 4 && !err; b4++) {
				const uint32_t blk = b8 * 4 + b4;
				const uint32_t bx = blk_x(blk), by = blk_y(blk);
				const bool in_a = bx > 0, in_b = by > 0;
				const uint32_t bit_a = in_a ? blk_of(bx - 1, by) : blk_of(3, by);
				const uint32_t bit_b = in_b ? blk_of(bx, by - 1) : blk_of(bx, 3);
				blk_in(mc->luma[blk], i16 ? 15 : 16);
				residual_block(i16 ? 1 : 2, i16 ? 15 : 16, false, in_a, bit_a, in_b, bit_b, blk,
					       i16 ? H264GPU_F_I16_AC : H264GPU_F_LEVEL4X4, blk * 16);
				blk_out(mc->luma[blk], i16 ? 15 : 16);
			}
		}
		if (cat != 1 && cat != 2)
			return;
		const uint32_t nblk = cat == 1 ? 4 : 8;
		if (cbp_chroma & 3) {
			for (uint32_t ic = 0; ic < 2 && !err; ic++) {
				blk_in(mc->cdc[ic], nblk);
				residual_block(3, nblk, false, false, 17 + ic, false, 17 + ic, 17 + ic,
					       H264GPU_F_CHROMA_DC, ic * 16);
				blk_out(mc->cdc[ic], nblk);
			}
		}
		if (cbp_chroma & 2) {
			for (uint32_t ic = 0; ic < 2; ic++) {
				for (uint32_t blk = 0; blk < nblk && !err; blk++) {
					const uint32_t bx = blk & 1, by = blk >> 1;
					const bool in_a = bx > 0, in_b = by > 0;
					const uint32_t bit_a = ic * 8 + (in_a ? blk - 1 : blk + 1);
					const uint32_t bit_b = ic * 8 + (in_b ? blk - 2 : blk + nblk - 2);
					blk_in(mc->cac[ic][blk], 15);
					residual_block(4, 15, true, in_a, bit_a, in_b, bit_b, ic * 8 + blk,
						       H264GPU_F_CHROMA_AC, (ic * 16 + blk) * 16);
					blk_out(mc->cac[ic][blk], 15);
				}
			}
		}
	}

	/* ---- 7.3.5 macroblock_layer; control flow of src/h264_syntax_slice_data.h:604-696 -------- */

	CABAC_HD bool macroblock(Mb &m)
	{
		const uint32_t st = sp->slice_type;
		if (st == ST_I)
			m.raw_type = intra_type(3, m.raw_type);
		else if (st == ST_B)
			m.raw_type = mb_type_b(m.raw_type);
		else
			m.raw_type = mb_type_p(m.raw_type);
		bool i16 = false;
		uint32_t i16_pred = 0;
		if (!Coder::kEncode)
			m.cbp_luma = m.cbp_chroma = 0;
		if (err || !classify(m, i16, i16_pred))
			return false;
		hadd(H264GPU_F_RAW_MB_TYPE, 0, m.raw_type);
		me.intra = m.mb_type <= MB_SI;
		me.not_inxn = m.mb_type != MB_I_NxN;
		me.not_bdirect = m.mb_type != MB_B_Direct_16x16;
		if (i16)
			hadd(H264GPU_F_I16_PRED_MODE, 0, i16_pred);

		if (m.mb_type == MB_I_PCM) {
			const uint32_t mbw_c = cat == 0 ? 0 : (cat == 3 ? 16 : 8), mbh_c = cat == 0 ? 0 : (cat == 1 ? 8 : 16);
			const uint32_t nchroma = mbw_c * mbh_c;
			c.pcm_begin();
			for (uint32_t i = 0; i < 256; i++) {
				const uint32_t v = c.pcm_sample(mc ? mc->pcm[i] : 0, sp->bit_depth_luma);
				hadd(H264GPU_F_PCM_LUMA, i, v);
				if (!Coder::kEncode && mc)
					mc->pcm[i] = (uint8_t)v;
			}
			for (uint32_t ic = 0; ic < 2; ic++)
				for (uint32_t i = 0; i < nchroma; i++) {
					const uint32_t k = 256 + ic * 256 + i;
					const uint32_t v = c.pcm_sample(mc ? mc->pcm[k] : 0, sp->bit_depth_chroma);
					hadd(H264GPU_F_PCM_CHROMA, ic * 256 + i, v);
					if (!Coder::kEncode && mc)
						mc->pcm[k] = (uint8_t)v;
				}
			c.pcm_end();
			me.pcm = 1;
			me.cbp_luma = 0xf;
			me.cbp_chroma = 2;
			me.cbf = 0x7ffffu;
			me.cbf_c = 0xffffu;
			last_dqp_nz = 0;
			return !err && !c.failed();
		}

		bool no_sub_lt_8x8 = true;
		bool t8 = false;
		const uint32_t max0 = sp->num_ref_idx_l0_active_minus1, max1 = sp->num_ref_idx_l1_active_minus1;
		if (m.num_part == 4) {
			/* sub_mb_pred: src/h264_syntax_slice_data.h:422-503 */
			uint32_t nsub[4], spm[4], shape[4];
			bool direct[4];
			for (uint32_t i = 0; i < 4; i++) {
				const uint32_t t = st == ST_B ? sub_type_b(m.sub_type[i]) : sub_type_p(m.sub_type[i]);
				m.sub_type[i] = t;
				hadd(H264GPU_F_RAW_SUB_MB_TYPE, i, t);
				if (st == ST_B) {
					direct[i] = t == 0;
					nsub[i] = (t == 0 || t >= 10) ? 4 : t <= 3 ? 1 : 2;
					spm[i] = t == 0 ? PM_DIRECT
						       : (t == 1 || t == 4 || t == 5 || t == 10) ? PM_L0
						       : (t == 2 || t == 6 || t == 7 || t == 11) ? PM_L1 : PM_BI;
					/* 0: 8x8, 1: 8x4, 2: 4x8, 3: 4x4 */
					shape[i] = t <= 3 ? 0 : t >= 10 ? 3 : ((t - 4) & 1) ? 2 : 1;
				} else {
					direct[i] = false;
					nsub[i] = t == 0 ? 1 : t == 3 ? 4 : 2;
					spm[i] = PM_L0;
					shape[i] = t;
				}
			}
			for (uint32_t list = 0; list < 2; list++) {
				const uint32_t other = list ? PM_L0 : PM_L1;
				if ((list ? max1 : max0) == 0)
					continue;
				for (uint32_t i = 0; i < 4; i++)
					if (!direct[i] && spm[i] != other) {
						const uint32_t r = ref_idx(list, i, (uint32_t)m.ref_idx[list][i]);
						m.ref_idx[list][i] = (int8_t)r;
						hadd(list ? H264GPU_F_REF_IDX_L1 : H264GPU_F_REF_IDX_L0, i, r);
						set_ref(list, 1u << i, r);
					}
			}
			for (uint32_t list = 0; list < 2; list++) {
				const uint32_t other = list ? PM_L0 : PM_L1;
				for (uint32_t i = 0; i < 4; i++) {
					if (direct[i] || spm[i] == other)
						continue;
					const uint32_t ox = 2 * (i & 1), oy = 2 * (i >> 1);
					for (uint32_t j = 0; j < nsub[i]; j++) {
						uint32_t bx = ox, by = oy, w = 2, h = 2;
						if (shape[i] == 1) {
							by += j;
							h = 1;
						} else if (shape[i] == 2) {
							bx += j;
							w = 1;
						} else if (shape[i] == 3) {
							bx += j & 1;
							by += j >> 1;
							w = h = 1;
						}
						mvd_part(m, list, i * 4 + j, bx, by, w, h);
						for (uint32_t comp = 0; comp < 2; comp++)
							hadd(list ? H264GPU_F_MVD_L1 : H264GPU_F_MVD_L0, (i * 4 + j) * 2 + comp,
							     m.mvd[list][i * 4 + j][comp]);
					}
				}
			}
			for (uint32_t i = 0; i < 4; i++) {
				if (!direct[i]) {
					if (nsub[i] > 1)
						no_sub_lt_8x8 = false;
				} else if (!sp->direct_8x8_inference_flag) {
					no_sub_lt_8x8 = false;
				}
			}
		} else {
			if (sp->transform_8x8_mode_flag && m.mb_type == MB_I_NxN) {
				t8 = t8_flag(m.t8) != 0;
				if (t8)
					m.pm[0] = PM_I8;
			}
			/* mb_pred: src/h264_syntax_slice_data.h:506-601 */
			if (m.pm[0] <= PM_I16) {
				if (m.pm[0] == PM_I4 || m.pm[0] == PM_I8) {
					const uint32_t nmode = m.pm[0] == PM_I4 ? 16 : 4;
					for (uint32_t i = 0; i < nmode; i++) {
						intra_pred_mode(m.prev_flag[i], m.rem_mode[i]);
						hadd(m.pm[0] == PM_I4 ? H264GPU_F_INTRA4X4_PRED_MODE : H264GPU_F_INTRA8X8_PRED_MODE,
						     i, m.prev_flag[i] ? -1 : (int64_t)m.rem_mode[i]);
					}
				}
				if (cat == 1 || cat == 2) {
					m.chroma_mode = (uint8_t)chroma_pred_mode(m.chroma_mode);
					hadd(H264GPU_F_INTRA_CHROMA_PRED_MODE, 0, m.chroma_mode);
					me.chroma_mode_nz = m.chroma_mode != 0;
				}
			} else if (m.pm[0] != PM_DIRECT) {
				const bool h2 = m.mb_type == MB_P_16x8 || m.mb_type == MB_B_16x8; /* two 16x8 */
				for (uint32_t list = 0; list < 2; list++) {
					const uint32_t other = list ? PM_L0 : PM_L1;
					if ((list ? max1 : max0) == 0)
						continue;
					for (uint32_t i = 0; i < m.num_part; i++)
						if (m.pm[i] != other) {
							const uint32_t b8 = m.num_part == 1 ? 0 : h2 ? 2 * i : i;
							const uint32_t mask = m.num_part == 1 ? 0xfu : h2 ? (3u << (2 * i)) : (5u << i);
							const uint32_t r = ref_idx(list, b8, (uint32_t)m.ref_idx[list][i]);
							m.ref_idx[list][i] = (int8_t)r;
							hadd(list ? H264GPU_F_REF_IDX_L1 : H264GPU_F_REF_IDX_L0, i, r);
							set_ref(list, mask, r);
						}
				}
				for (uint32_t list = 0; list < 2; list++) {
					const uint32_t other = list ? PM_L0 : PM_L1;
					for (uint32_t i = 0; i < m.num_part; i++)
						if (m.pm[i] != other) {
							if (m.num_part == 1)
								mvd_part(m, list, 0, 0, 0, 4, 4);
							else if (h2)
								mvd_part(m, list, i * 4, 0, 2 * i, 4, 2);
							else
								mvd_part(m, list, i * 4, 2 * i, 0, 2, 4);
							for (uint32_t comp = 0; comp < 2; comp++)
								hadd(list ? H264GPU_F_MVD_L1 : H264GPU_F_MVD_L0, (i * 4) * 2 + comp,
								     m.mvd[list][i * 4][comp]);
						}
				}
			}
		}
		if (err)
			return false;

		if (!i16) {
			coded_block_pattern(m.cbp_luma, m.cbp_chroma);
			hadd(H264GPU_F_CBP, 0, m.cbp_luma + 16 * m.cbp_chroma);
			if (m.cbp_luma > 0 && sp->transform_8x8_mode_flag && m.mb_type != MB_I_NxN && no_sub_lt_8x8 &&
			    (m.mb_type != MB_B_Direct_16x16 || sp->direct_8x8_inference_flag))
				t8 = t8_flag(m.t8) != 0;
		}
		m.t8 = t8;
		me.t8 = t8;
		me.cbp_luma = (uint8_t)m.cbp_luma;
		me.cbp_chroma = (uint8_t)m.cbp_chroma;
		hadd(H264GPU_F_TRANSFORM_8X8, 0, t8 ? 1 : 0);
		hadd(H264GPU_F_CBP_LUMA, 0, m.cbp_luma);
		hadd(H264GPU_F_CBP_CHROMA, 0, m.cbp_chroma);

		if (m.cbp_luma > 0 || m.cbp_chroma > 0 || i16) {
			m.qp_delta = qp_delta(m.qp_delta);
			hadd(H264GPU_F_MB_QP_DELTA, 0, m.qp_delta);
			last_dqp_nz = m.qp_delta != 0;
			residual(i16, m.cbp_luma, m.cbp_chroma, t8);
		} else {
			last_dqp_nz = 0;
		}
		return !err && !c.failed();
	}

	/* ---- slice_data() loop body: 7.3.4 with entropy_coding_mode_flag = 1 ------------------- */

	CABAC_HD void begin_slice(const h264gpu_slice_params *p, Nb *r)
	{
		sp = p;
		ring = r;
		W = p->pic_width_in_mbs;
		cat = p->chroma_array_type;
		first = p->first_mb_in_slice;
		last_dqp_nz = 0;
		err = false;
		mc = 0;
	}

	/* One macroblock at address mb.  skipped / end: encoder inputs, decoder outputs.
	 * Returns false on a syntax error. */
	CABAC_HD bool mb_step(uint32_t mb, bool &skipped, Mb &m, bool &end)
	{
		cur = mb;
		availA = mb >= first + 1 && mb % W != 0;
		availB = mb >= first + W;
		nbA = slot(mb - 1);
		nbB = slot(mb >= W ? mb - W : 0);
		for (uint32_t i = 0; i < sizeof(me); i++)
			((uint8_t *)&me)[i] = 0;
		me.valid = 1;
		for (uint32_t i = 0; i < sizeof(amvd); i++)
			((uint8_t *)amvd)[i] = 0;
		hash = 0;
		const bool inter = sp->slice_type != ST_I;
		bool ok = true;
		if (inter)
			skipped = skip_flag(skipped ? 1u : 0u) != 0;
		else
			skipped = false;
		if (skipped) {
			me.skip = 1;
			m.mb_type = sp->slice_type == ST_B ? MB_B_SKIP : MB_P_SKIP;
			last_dqp_nz = 0;
		} else {
			ok = macroblock(m);
		}
		for (uint32_t list = 0; list < 2; list++)
			for (uint32_t comp = 0; comp < 2; comp++)
				for (uint32_t k = 0; k < 4; k++) {
					me.mvd_right[list][comp][k] = amvd[list][comp][blk_of(3, k)];
					me.mvd_bottom[list][comp][k] = amvd[list][comp][blk_of(k, 3)];
				}
		*slot(mb) = me;
		if (!ok || c.failed())
			return false;
		end = c.term(end ? 1u : 0u) != 0;
		return !c.failed();
	}
};

} /* namespace cabac */

#endif /* CABAC_SYNTAX_H */
