/*
 * h264/h264_writer.h — serialise path (reference: include/h264/h264_writer.h:31-50).
 * The NAL unit is written WITHOUT start code into a caller-initialised bitstream
 * (emulation prevention on); framing is the caller's job, or the bulk
 * h264gpu_frame_* stage (include/h264gpu.h).
 */
#ifndef H264B200_WRITER_H
#define H264B200_WRITER_H

/* serialise the NAL unit currently described by ctx (header, then SPS / PPS /
 * SEI / AUD / filler / slice header + raw slice data) */
H264_API int h264_write_nalu(struct h264_bitstream *bs, struct h264_ctx *ctx);
/* synthesise a grey I slice / an all-skipped P slice of mb_count macroblocks */
H264_API int h264_write_grey_i_slice(struct h264_bitstream *bs, struct h264_ctx *ctx,
				     uint32_t mb_count);
H264_API int h264_write_skipped_p_slice(struct h264_bitstream *bs, struct h264_ctx *ctx,
					uint32_t mb_count);
/* patch the slice header of the NAL in bs in place (same bit length only) */
H264_API int h264_rewrite_slice_header(struct h264_bitstream *bs, struct h264_ctx *ctx,
				       const struct h264_slice_header *sh);
/* Extension of this library (not in the reference): the same header rewrite (same checks, same
 * return codes, same bytes) left in a patch record for h264gpu_patch_slice_headers
 * (include/h264gpu_slice.h), which applies any number of them to a stream resident on the GPU in
 * one launch; the caller fills patch->nal_off */
struct h264gpu_hdr_patch;
H264_API int h264_rewrite_slice_header_patch(struct h264_ctx *ctx, const struct h264_slice_header *sh,
					     struct h264gpu_hdr_patch *patch);

#endif /* H264B200_WRITER_H */
