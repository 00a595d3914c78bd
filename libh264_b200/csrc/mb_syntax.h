/*
 * mb_syntax.h — one macroblock's syntax elements, independent of the entropy coder.  Filled
 * by the reference harness from the reference's private ctx->mb (struct h264_macroblock,
 * src/h264_macroblock.h:105-167) and consumed by the CABAC transcoder of the synthetic-stream
 * library, so that a CAVLC slice parsed by the REFERENCE can be re-coded as a CABAC slice
 * carrying exactly the same syntax elements (tests/test_cabac.py, "twin" test).
 * Test / workload infrastructure only.
 */
#ifndef MB_SYNTAX_H
#define MB_SYNTAX_H

#include <stdint.h>

struct h264_mb_syntax {
	uint32_t mb_addr;
	uint32_t mb_type;     /* enum h264_mb_type */
	uint32_t raw_mb_type;
	uint32_t raw_sub_mb_type[4];
	int32_t mb_qp_delta;
	uint8_t transform_size_8x8_flag;
	uint8_t intra_chroma_pred_mode;
	uint8_t cbp_luma, cbp_chroma;
	int8_t intra4x4_pred_mode[16]; /* -1: prev_intra4x4_pred_mode_flag set */
	int8_t intra8x8_pred_mode[4];
	uint8_t ref_idx[2][4];
	int16_t mvd[2][16][2];         /* [list][mbPart * 4 + subMbPart][comp] */
	int16_t dc16[16];
	int16_t ac16[16][16];          /* [blkIdx][0..14] */
	int16_t l4[16][16];
	int16_t cdc[2][16];
	int16_t cac[2][16][16];        /* [iCbCr][blkIdx][0..14] */
	uint8_t pcm[768];              /* 256 luma, 256 slots Cb, 256 slots Cr (8-bit samples) */
};

#endif /* MB_SYNTAX_H */
