/*
 * h264_cabac_enc.h — CABAC arithmetic ENCODER (9.3.4.2) with the few binarisations
 * the writer needs for its concealment slices (reference: src/h264_bac.c:143-358,
 * src/h264_cabac.c:55-232,630-975; used by src/h264_writer.c:79-129,177-219).
 */
#ifndef H264B200_CABAC_ENC_H
#define H264B200_CABAC_ENC_H

#include "h264_priv.h"

struct h264_cabac_enc {
	struct h264_bitstream *bs;
	uint32_t low, range, outstanding;
	int first_bit;
	uint8_t state[1024]; /* pStateIdx | valMPS << 6 */
};

void h264_cabac_enc_init(struct h264_cabac_enc *e, struct h264_bitstream *bs,
			 enum h264_slice_type slice_type, uint32_t cabac_init_idc, int32_t slice_qp);
int h264_cabac_enc_decision(struct h264_cabac_enc *e, uint32_t ctx_idx, int bin);
int h264_cabac_enc_bypass(struct h264_cabac_enc *e, int bin);
int h264_cabac_enc_terminate(struct h264_cabac_enc *e, int bin);

/* I_16x16_2_0_0 with DC chroma prediction, mb_qp_delta 0 and an empty DC block */
int h264_cabac_enc_grey_i_mb(struct h264_cabac_enc *e, const struct h264_ctx *ctx, uint32_t mb_addr,
			     uint32_t index_in_slice);
/* mb_skip_flag = 1 */
int h264_cabac_enc_skipped_p_mb(struct h264_cabac_enc *e, const struct h264_ctx *ctx, uint32_t mb_addr,
				uint32_t index_in_slice);
int h264_cabac_enc_end_of_slice(struct h264_cabac_enc *e, int last);

#endif /* H264B200_CABAC_ENC_H */
