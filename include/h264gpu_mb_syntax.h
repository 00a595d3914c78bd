/*
 * h264gpu_mb_syntax.h — one macroblock's syntax elements, independent of the entropy coder: the
 * content of the reference's private ctx->mb (struct h264_macroblock,
 * src/h264_macroblock.h:105-167) that its dump walks (src/h264_dump.c:295-316), as a plain
 * record.
 *   - the FULL-RECORD output of the CAVLC slice kernel (h264gpu_cavlc_parse_full_dev,
 *     include/h264gpu_slice.h): a consumer gets mvd / ref_idx / levels without re-parsing;
 *   - filled by the reference harness from ctx->mb (oracle/ref_harness.c) as the checker of that
 *     output, and consumed by the CABAC transcoder of the synthetic-stream library (a CAVLC slice
 *     parsed by the REFERENCE re-coded as a CABAC slice with the same elements, tests/test_cabac.py).
 * Unset fields are 0; the record is zero-filled first (padding included), so two records of the
 * same macroblock compare equal with memcmp.
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
