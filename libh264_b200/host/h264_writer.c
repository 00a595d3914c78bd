/*
 * h264_writer.c — serialise the NAL unit held in a context, the two concealment
 * slices, and the in-place slice header patch (reference: src/h264_writer.c:240-370).
 *
 * One NAL unit at a time is host work (tens of bytes of header syntax through the
 * same bidirectional walk the reader uses); escaping + framing of MANY payloads at
 * once is the GPU stage h264gpu_frame_* (include/h264gpu.h).
 */
#include "h264_priv.h"
#include "h264_cabac_enc.h"

int h264_write_nalu(struct h264_bitstream *bs, struct h264_ctx *ctx)
{
	if (bs == NULL || ctx == NULL)
		return -EINVAL;
	struct h264_io io = {H264_IO_WRITE, bs, ctx, NULL, NULL, NULL};
	return h264_syntax_nalu(&io);
}

/* ---- concealment slices (src/h264_writer.c:49-219) ---------------------------------------- */

enum conceal_kind { GREY_I, SKIPPED_P };

static int align_with(struct h264_bitstream *bs, uint32_t bit)
{
	while (!h264_bs_byte_aligned(bs)) {
		int r = h264_bs_write_bits(bs, bit, 1);
		if (r < 0)
			return r;
	}
	return 0;
}

static int conceal_cavlc(struct h264_bitstream *bs, enum conceal_kind kind, uint32_t mb_count)
{
	int r = 0;
	if (kind == SKIPPED_P) {
		r = h264_bs_write_bits_ue(bs, mb_count); /* one mb_skip_run covers the slice */
	} else {
		/* per MB: mb_type 3 (I_16x16_2_0_0) = 00100, intra_chroma_pred_mode 0 = 1,
		 * mb_qp_delta 0 = 1, coeff_token(nC = 0, no coefficient) = 1 */
		for (uint32_t i = 0; i < mb_count && r >= 0; i++)
			r = h264_bs_write_bits(bs, 0x27, 8);
	}
	return r < 0 ? r : h264_bs_write_rbsp_trailing_bits(bs);
}

static int conceal_cabac(struct h264_bitstream *bs, struct h264_ctx *ctx, enum conceal_kind kind,
			 uint32_t mb_count)
{
	struct h264_cabac_enc enc;
	int r = align_with(bs, 1); /* cabac_alignment_one_bit */
	if (r < 0)
		return r;
	h264_cabac_enc_init(&enc, bs, ctx->slice_type, ctx->sh.cabac_init_idc, ctx->SliceQPLuma);
	for (uint32_t i = 0; i < mb_count && r >= 0; i++) {
		const uint32_t addr = ctx->sh.first_mb_in_slice + i;
		if (kind == SKIPPED_P)
			r = h264_cabac_enc_skipped_p_mb(&enc, ctx, addr, i);
		else
			r = h264_cabac_enc_grey_i_mb(&enc, ctx, addr, i);
		if (r >= 0)
			r = h264_cabac_enc_end_of_slice(&enc, i == mb_count - 1);
	}
	/* the terminate bin already carries rbsp_stop_one_bit */
	return r < 0 ? r : align_with(bs, 0);
}

static int conceal_slice(struct h264_bitstream *bs, struct h264_ctx *ctx, enum conceal_kind kind,
			 uint32_t mb_count)
{
	if (bs == NULL || ctx == NULL || mb_count == 0)
		return -EINVAL;
	int r = h264_ctx_activate_pps(ctx, ctx->sh.pic_parameter_set_id);
	if (r < 0)
		return r;
	if ((r = h264_write_nalu(bs, ctx)) < 0)
		return r;
	return ctx->pps->entropy_coding_mode_flag ? conceal_cabac(bs, ctx, kind, mb_count)
						   : conceal_cavlc(bs, kind, mb_count);
}

int h264_write_grey_i_slice(struct h264_bitstream *bs, struct h264_ctx *ctx, uint32_t mb_count)
{
	return conceal_slice(bs, ctx, GREY_I, mb_count);
}

int h264_write_skipped_p_slice(struct h264_bitstream *bs, struct h264_ctx *ctx, uint32_t mb_count)
{
	return conceal_slice(bs, ctx, SKIPPED_P, mb_count);
}

/* ---- slice header patch, only when the bit length is unchanged ----------------------------- */

/* the new NAL header + slice header in a 64-byte scratch bitstream, as the reference builds it
 * (src/h264_writer.c:327-349); on failure the context's slice header is what it was */
static int rewrite_scratch(struct h264_ctx *ctx, const struct h264_slice_header *sh, struct h264_bitstream *tmp,
			   uint8_t scratch[64])
{
	h264_bs_init(tmp, scratch, 64, 1);
	const struct h264_slice_header saved = ctx->sh;
	const size_t saved_bits = ctx->sh_bits;
	int r = h264_ctx_set_slice_header(ctx, sh);
	if (r < 0)
		return r;
	r = h264_write_nalu(tmp, ctx);
	if (r >= 0 && ctx->sh_bits != saved_bits)
		r = -EPROTO;
	if (r < 0) {
		ctx->sh = saved;
		ctx->sh_bits = saved_bits;
	}
	return r;
}

int h264_rewrite_slice_header(struct h264_bitstream *bs, struct h264_ctx *ctx,
			      const struct h264_slice_header *sh)
{
	if (bs == NULL || ctx == NULL || sh == NULL)
		return -EINVAL;
	uint8_t scratch[64];
	struct h264_bitstream tmp;
	const int r = rewrite_scratch(ctx, sh, &tmp, scratch);
	if (r < 0)
		return r;
	/* whole bytes, then the leading bits of the byte shared with the slice data */
	memcpy(bs->data, tmp.data, tmp.off);
	if (tmp.cachebits != 0) {
		const uint8_t keep = (uint8_t)((1u << (8 - tmp.cachebits)) - 1);
		bs->data[tmp.off] = (uint8_t)((tmp.cache & ~keep) | (bs->data[tmp.off] & keep));
	}
	return 0;
}

/* Extension: the same checks and the same bytes, left in a patch record for
 * h264gpu_patch_slice_headers (a stream resident on the device) instead of being copied over
 * the NAL; the caller sets patch->nal_off */
int h264_rewrite_slice_header_patch(struct h264_ctx *ctx, const struct h264_slice_header *sh,
				    struct h264gpu_hdr_patch *patch)
{
	if (ctx == NULL || sh == NULL || patch == NULL)
		return -EINVAL;
	uint8_t scratch[64];
	struct h264_bitstream tmp;
	memset(scratch, 0, sizeof(scratch));
	const int r = rewrite_scratch(ctx, sh, &tmp, scratch);
	if (r < 0)
		return r;
	if (tmp.off + (tmp.cachebits != 0) > 64)
		return -ENOBUFS;
	memcpy(patch->bytes, scratch, sizeof(patch->bytes));
	patch->nbytes = (uint32_t)tmp.off;
	patch->tail_bits = (uint32_t)tmp.cachebits;
	if (tmp.cachebits != 0 && tmp.off < 64)
		patch->bytes[tmp.off] = (uint8_t)tmp.cache;
	return 0;
}
