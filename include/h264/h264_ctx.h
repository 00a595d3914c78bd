/*
 * h264/h264_ctx.h — parsing context and the callback table of the libh264 API
 * (reference: include/h264/h264_ctx.h:31-227; the order of the 18 function
 * pointers in struct h264_ctx_cbs is ABI).
 *
 * Callback order per NAL unit, identical to the reference
 * (src/h264_syntax.h:1446-1604):
 *   nalu_begin, then for a slice: slice_data_begin, slice_data_mb x N,
 *   slice_data_end, slice; or sps / pps / aud; or per SEI message: sei, sei_<type>;
 *   then au_end if this NAL unit starts a new access unit; then nalu_end.
 * au_end for access unit n therefore fires inside the first NAL unit of access
 * unit n+1, and never for the last one of a stream.
 */
#ifndef H264B200_CTX_H
#define H264B200_CTX_H

struct h264_ctx;

struct h264_ctx_cbs {
	void (*au_end)(struct h264_ctx *ctx, void *userdata);

	void (*nalu_begin)(struct h264_ctx *ctx, enum h264_nalu_type type, const uint8_t *buf,
			   size_t len, const struct h264_nalu_header *nh, void *userdata);
	void (*nalu_end)(struct h264_ctx *ctx, enum h264_nalu_type type, const uint8_t *buf,
			 size_t len, const struct h264_nalu_header *nh, void *userdata);

	void (*slice)(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
		      const struct h264_slice_header *sh, void *userdata);
	void (*slice_data_begin)(struct h264_ctx *ctx, const struct h264_slice_header *sh,
				 void *userdata);
	void (*slice_data_end)(struct h264_ctx *ctx, const struct h264_slice_header *sh,
			       uint32_t mb_count, void *userdata);
	void (*slice_data_mb)(struct h264_ctx *ctx, const struct h264_slice_header *sh,
			      uint32_t mb_addr, enum h264_mb_type mb_type, void *userdata);

	void (*sps)(struct h264_ctx *ctx, const uint8_t *buf, size_t len, const struct h264_sps *sps,
		    void *userdata);
	void (*pps)(struct h264_ctx *ctx, const uint8_t *buf, size_t len, const struct h264_pps *pps,
		    void *userdata);
	void (*aud)(struct h264_ctx *ctx, const uint8_t *buf, size_t len, const struct h264_aud *aud,
		    void *userdata);

	void (*sei)(struct h264_ctx *ctx, enum h264_sei_type type, const uint8_t *buf, size_t len,
		    void *userdata);
	void (*sei_buffering_period)(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
				     const struct h264_sei_buffering_period *sei, void *userdata);
	void (*sei_pic_timing)(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
			       const struct h264_sei_pic_timing *sei, void *userdata);
	void (*sei_pan_scan_rect)(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
				  const struct h264_sei_pan_scan_rect *sei, void *userdata);
	void (*sei_filler_payload)(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
				   const struct h264_sei_filler_payload *sei, void *userdata);
	void (*sei_user_data_registered)(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
					 const struct h264_sei_user_data_registered *sei,
					 void *userdata);
	void (*sei_user_data_unregistered)(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
					   const struct h264_sei_user_data_unregistered *sei,
					   void *userdata);
	void (*sei_recovery_point)(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
				   const struct h264_sei_recovery_point *sei, void *userdata);
};

H264_API int h264_ctx_new(struct h264_ctx **ret_obj);
H264_API int h264_ctx_destroy(struct h264_ctx *ctx);
H264_API int h264_ctx_clear(struct h264_ctx *ctx);
H264_API int h264_ctx_clear_nalu(struct h264_ctx *ctx);
H264_API int h264_ctx_set_nalu_header(struct h264_ctx *ctx, const struct h264_nalu_header *nh);
H264_API int h264_ctx_is_nalu_unknown(struct h264_ctx *ctx);
H264_API int h264_ctx_set_aud(struct h264_ctx *ctx, const struct h264_aud *aud);
H264_API int h264_ctx_set_sps(struct h264_ctx *ctx, const struct h264_sps *sps);
H264_API int h264_ctx_set_pps(struct h264_ctx *ctx, const struct h264_pps *pps);
/* Extension of this library (not in the reference, which keeps the map private in ctx->slice):
 * the macroblock -> slice group map (8.2.2) of the slice whose header is current, one byte per
 * macroblock.  Returns PicSizeInMbs, or a negative errno. */
H264_API int h264_ctx_get_slice_group_map(const struct h264_ctx *ctx, uint8_t *map, size_t cap);
H264_API int h264_ctx_set_filler(struct h264_ctx *ctx, size_t len);
H264_API const struct h264_sps *h264_ctx_get_sps(struct h264_ctx *ctx);
H264_API const struct h264_pps *h264_ctx_get_pps(struct h264_ctx *ctx);
H264_API int h264_ctx_add_sei(struct h264_ctx *ctx, const struct h264_sei *sei);
H264_API int h264_ctx_get_sei_count(struct h264_ctx *ctx);
H264_API uint64_t h264_ctx_sei_pic_timing_to_ts(struct h264_ctx *ctx,
						const struct h264_sei_pic_timing *sei);
H264_API uint64_t h264_ctx_sei_pic_timing_to_us(struct h264_ctx *ctx,
						const struct h264_sei_pic_timing *sei);
H264_API int h264_ctx_set_slice_header(struct h264_ctx *ctx, const struct h264_slice_header *sh);
H264_API int h264_ctx_get_info(struct h264_ctx *ctx, struct h264_info *info);

#endif /* H264B200_CTX_H */
