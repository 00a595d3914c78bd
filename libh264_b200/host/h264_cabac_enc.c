/*
 * h264_cabac_enc.c — see h264_cabac_enc.h.  The probability tables are the ones the
 * GPU side uses (csrc/cabac_tables.h: H.264 Tables 9-12..9-33, 9-44, 9-45).
 */
#include "h264_cabac_enc.h"

#define CABAC_TAB static const
#include "cabac_tables.h"

/* 9.3.1.1: preCtxState = Clip3(1, 126, ((m * Clip3(0?1, 51, QP)) >> 4) + n) */
void h264_cabac_enc_init(struct h264_cabac_enc *e, struct h264_bitstream *bs,
			 enum h264_slice_type slice_type, uint32_t cabac_init_idc, int32_t slice_qp)
{
	memset(e, 0, sizeof(*e));
	e->bs = bs;
	e->range = 510;
	e->first_bit = 1;
	const int intra = slice_type == H264_SLICE_TYPE_I || slice_type == H264_SLICE_TYPE_SI;
	if (!intra && cabac_init_idc > 2)
		return; /* the reference leaves the states untouched (all zero) here too */
	const int8_t(*mn)[2] = cabac_init_mn[intra ? 0 : 1 + cabac_init_idc];
	const int32_t qp = slice_qp < 1 ? 1 : slice_qp > 51 ? 51 : slice_qp;
	for (int i = 0; i < 1024; i++) {
		const int32_t pre = ((mn[i][0] * qp) >> 4) + mn[i][1];
		if (pre <= 63)
			e->state[i] = (uint8_t)(63 - (pre < 1 ? 1 : pre));
		else
			e->state[i] = (uint8_t)(((pre > 126 ? 126 : pre) - 64) | 0x40);
	}
}

/* PutBit (Figure 9-9) */
static int put_bit(struct h264_cabac_enc *e, uint32_t bit)
{
	int r = 0;
	if (e->first_bit)
		e->first_bit = 0;
	else
		r = h264_bs_write_bits(e->bs, bit, 1);
	for (; r >= 0 && e->outstanding > 0; e->outstanding--)
		r = h264_bs_write_bits(e->bs, !bit, 1);
	return r < 0 ? r : 0;
}

/* RenormE (Figure 9-8) */
static int renorm(struct h264_cabac_enc *e)
{
	int r = 0;
	while (r >= 0 && e->range < 256) {
		if (e->low < 256) {
			r = put_bit(e, 0);
		} else if (e->low < 512) {
			e->low -= 256;
			e->outstanding++;
		} else {
			e->low -= 512;
			r = put_bit(e, 1);
		}
		e->range <<= 1;
		e->low <<= 1;
	}
	return r;
}

int h264_cabac_enc_decision(struct h264_cabac_enc *e, uint32_t ctx_idx, int bin)
{
	uint8_t *st = &e->state[ctx_idx];
	const uint32_t idx = *st & 0x3f, mps = *st >> 6;
	const uint32_t lps_range = cabac_range_lps[idx][(e->range >> 6) & 3];
	e->range -= lps_range;
	if ((uint32_t)!!bin == mps) {
		*st = (uint8_t)(cabac_trans_mps[idx] | mps << 6);
	} else {
		e->low += e->range;
		e->range = lps_range;
		*st = (uint8_t)(cabac_trans_lps[idx] | (idx == 0 ? !mps : mps) << 6);
	}
	return renorm(e);
}

int h264_cabac_enc_bypass(struct h264_cabac_enc *e, int bin)
{
	e->low <<= 1;
	if (bin)
		e->low += e->range;
	if (e->low >= 1024) {
		e->low -= 1024;
		return put_bit(e, 1);
	}
	if (e->low >= 512) {
		e->low -= 512;
		e->outstanding++;
		return 0;
	}
	return put_bit(e, 0);
}

/* EncodeTerminate + EncodeFlush (Figures 9-11, 9-12); the last bit written is the
 * rbsp_stop_one_bit */
int h264_cabac_enc_terminate(struct h264_cabac_enc *e, int bin)
{
	e->range -= 2;
	if (!bin)
		return renorm(e);
	e->low += e->range;
	e->range = 2;
	int r = renorm(e);
	if (r >= 0)
		r = put_bit(e, (e->low >> 9) & 1);
	if (r >= 0)
		r = h264_bs_write_bits(e->bs, ((e->low >> 7) & 3) | 1, 2);
	return r < 0 ? r : 0;
}

/* neighbours inside the slice, frame macroblocks (src/h264_macroblock.c:306-351) */
static void neighbours(const struct h264_ctx *ctx, uint32_t mb_addr, int *left, int *up)
{
	const uint32_t w = ctx->spsd.PicWidthInMbs, first = ctx->sh.first_mb_in_slice;
	*left = w != 0 && mb_addr % w != 0 && mb_addr >= first + 1;
	*up = mb_addr >= first + w;
}

int h264_cabac_enc_grey_i_mb(struct h264_cabac_enc *e, const struct h264_ctx *ctx, uint32_t mb_addr,
			     uint32_t index_in_slice)
{
	(void)index_in_slice;
	int a, b, r;
	neighbours(ctx, mb_addr, &a, &b);
	/* mb_type 3, Table 9-36 "1 0 0 0 1 0": bin 0 ctx 3 + (A is not I_NxN) + (B ...),
	 * bin 1 terminate, bins 2..5 ctx 3+3, 3+4, 3+6, 3+7 (9.3.3.1.2, b3 = 0) */
	if ((r = h264_cabac_enc_decision(e, 3 + (uint32_t)a + (uint32_t)b, 1)) < 0 ||
	    (r = h264_cabac_enc_terminate(e, 0)) < 0 || (r = h264_cabac_enc_decision(e, 6, 0)) < 0 ||
	    (r = h264_cabac_enc_decision(e, 7, 0)) < 0 || (r = h264_cabac_enc_decision(e, 9, 1)) < 0 ||
	    (r = h264_cabac_enc_decision(e, 10, 0)) < 0)
		return r;
	/* intra_chroma_pred_mode 0: no neighbour has a non-DC mode -> ctx 64 */
	if ((r = h264_cabac_enc_decision(e, 64, 0)) < 0)
		return r;
	/* mb_qp_delta 0 -> ctx 60 */
	if ((r = h264_cabac_enc_decision(e, 60, 0)) < 0)
		return r;
	/* coded_block_flag of Intra16x16DCLevel (ctxBlockCat 0): an unavailable neighbour of
	 * an intra macroblock counts as coded (9.3.3.1.1.9) */
	return h264_cabac_enc_decision(e, 85 + (uint32_t)!a + 2 * (uint32_t)!b, 0);
}

int h264_cabac_enc_skipped_p_mb(struct h264_cabac_enc *e, const struct h264_ctx *ctx, uint32_t mb_addr,
				uint32_t index_in_slice)
{
	(void)ctx;
	(void)mb_addr;
	(void)index_in_slice;
	if (ctx->slice_type != H264_SLICE_TYPE_P && ctx->slice_type != H264_SLICE_TYPE_SP &&
	    ctx->slice_type != H264_SLICE_TYPE_B)
		return -EIO;
	/* every available neighbour is skipped too: ctxIdxInc 0 (9.3.3.1.1.1) */
	return h264_cabac_enc_decision(e, ctx->slice_type == H264_SLICE_TYPE_B ? 24 : 11, 1);
}

int h264_cabac_enc_end_of_slice(struct h264_cabac_enc *e, int last)
{
	return h264_cabac_enc_terminate(e, last);
}
