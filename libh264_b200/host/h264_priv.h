/*
 * h264_priv.h — private state of the host library (libh264.so of this repo).
 *
 * The layout is this library's own; only the public structs of include/h264/ are
 * ABI.  What the reference keeps per macroblock in ctx->mb (src/h264_priv.h:117-118)
 * lives in the GPU parse records here (include/h264gpu_slice.h).
 */
#ifndef H264B200_PRIV_H
#define H264B200_PRIV_H

#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "h264/h264.h"
#include "h264gpu.h"
#include "h264gpu_slice.h"

#define H264_SPS_MAX 32
#define H264_PPS_MAX 256
#define COUNT_OF(a) (sizeof(a) / sizeof((a)[0]))

struct h264_ctx {
	/* NAL unit being parsed / written */
	enum h264_nalu_type nalu_type;
	struct h264_nalu_header nalu_hdr;
	int nalu_unknown;
	/* access-unit change detection, 7.4.1.2.4 */
	int first_vcl, prev_vcl, prev_filler;
	struct h264_nalu_header prev_slice_nalu_hdr;
	struct h264_slice_header prev_slice_hdr;

	struct h264_aud aud;
	struct h264_sps *sps; /* active (points into sps_tab) */
	struct h264_pps *pps;
	struct h264_sps *sps_tab[H264_SPS_MAX];
	struct h264_pps *pps_tab[H264_PPS_MAX];
	struct h264_sps_derived spsd;
	struct h264_sei *sei_tab;
	uint32_t sei_count;
	size_t filler_len;

	/* current slice */
	enum h264_slice_type slice_type;
	struct h264_slice_header sh;
	size_t sh_bits; /* raw bit length of NAL header + slice header (escapes included) */
	struct {
		uint8_t partial, partialbits; /* unread low bits of the byte already fetched */
		const uint8_t *buf;           /* next raw byte */
		size_t len;
	} rawdata;

	/* variables derived from the active PPS / slice header (7.4.2.2, 7.4.3) */
	uint32_t SliceGroupChangeRate;
	int MbaffFrameFlag;
	uint32_t PicHeightInMbs, PicSizeInMbs, MaxPicNum, CurrPicNum, MapUnitsInSliceGroup0;
	int32_t SliceQPLuma, QSLuma;
};

struct h264_reader {
	struct h264_ctx_cbs cbs;
	void *userdata;
	struct h264_ctx *ctx;
	int stop;
	uint32_t flags;
	h264gpu_ctx *gpu; /* created on first bulk use */
	h264gpu_ctx *gpu_single; /* one-slice launches (h264_reader_parse_nalu, fall-backs) */
	/* bulk parse: macroblock records of every slice of the buffer, filled by the GPU */
	const struct h264gpu_mb_record *records;
	const struct h264gpu_slice_result *results;
	const struct h264gpu_slice_params *params;
	uint32_t n_slices, next_slice;
	const uint8_t *bulk_base; /* buffer the params' nal_off refer to */
};

/* h264_ctx.c */
int h264_ctx_activate_sps(struct h264_ctx *ctx, uint32_t id);
int h264_ctx_activate_pps(struct h264_ctx *ctx, uint32_t id);
int h264_ctx_new_sei(struct h264_ctx *ctx, struct h264_sei **out);
void h264_ctx_drop_sei(struct h264_ctx *ctx);
int h264_sei_fix_pointers(struct h264_sei *sei);

/* h264_syntax.c: one bidirectional walk of the header syntax */
enum h264_io_mode { H264_IO_READ, H264_IO_WRITE };
struct h264_io {
	enum h264_io_mode mode;
	struct h264_bitstream *bs;
	struct h264_ctx *ctx;
	const struct h264_ctx_cbs *cbs; /* READ only, may be NULL */
	void *userdata;
	struct h264_reader *reader; /* READ only: slice-data hand-off, may be NULL */
};
int h264_syntax_nalu(struct h264_io *io);
int h264_syntax_nalu_header(struct h264_io *io, struct h264_nalu_header *nh);
int h264_syntax_sps(struct h264_io *io, struct h264_sps *sps);
int h264_syntax_pps_body(struct h264_io *io, const struct h264_sps *sps, struct h264_pps *pps);
int h264_syntax_slice_header(struct h264_io *io, struct h264_slice_header *sh);
int h264_syntax_sei_payload(struct h264_io *io, struct h264_sei *sei);

/* h264_reader.c: slice data of the NAL being read (GPU) */
int h264_reader_slice_data(struct h264_reader *reader, struct h264_ctx *ctx,
			   const uint8_t *nal, size_t nal_len);
int h264_fill_slice_params(const struct h264_ctx *ctx, struct h264gpu_slice_params *p);

static inline uint32_t h264_ceil_log2(uint32_t v)
{
	/* Ceil(Log2(v)), 0 for v <= 1 */
	uint32_t n = 0;
	while (n < 32 && (1ull << n) < v)
		n++;
	return n;
}

#endif /* H264B200_PRIV_H */
