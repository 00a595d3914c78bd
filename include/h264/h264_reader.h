/*
 * h264/h264_reader.h — the parse entry points (reference: include/h264/h264_reader.h:31-85).
 *
 * In this library h264_reader_parse runs the Annex-B scan and, with
 * H264_READER_FLAGS_SLICE_DATA, the macroblock syntax parse of every CAVLC slice
 * of the buffer on the GPU (libh264gpu.so), then replays the callbacks on the
 * caller's thread in exactly the reference's order.
 */
#ifndef H264B200_READER_H
#define H264B200_READER_H

struct h264_reader;

/* also parse slice data and deliver slice_data_* callbacks (CAVLC slices only,
 * like the reference: CABAC slice data is skipped, src/h264_syntax_slice_data.h:715-717) */
#define H264_READER_FLAGS_SLICE_DATA 0x01
/* Extension of this library (not in the reference): with H264_READER_FLAGS_SLICE_DATA, also parse
 * the slice data of CABAC slices on the GPU and deliver slice_data_begin / slice_data_mb /
 * slice_data_end for them.  Without this bit CABAC slices get no slice-data callbacks, exactly
 * like the reference (src/h264_syntax_slice_data.h:715-717). */
#define H264_READER_FLAGS_SLICE_DATA_CABAC 0x02

H264_API int h264_reader_new(const struct h264_ctx_cbs *cbs, void *userdata,
			     struct h264_reader **ret_obj);
H264_API int h264_reader_destroy(struct h264_reader *reader);
H264_API struct h264_ctx *h264_reader_get_ctx(struct h264_reader *reader);
/* from inside a callback: stop h264_reader_parse after the current NAL unit */
H264_API int h264_reader_stop(struct h264_reader *reader);
/* parse every NAL unit of an Annex-B buffer; *off = bytes consumed; always 0 */
H264_API int h264_reader_parse(struct h264_reader *reader, uint32_t flags, const uint8_t *buf,
			       size_t len, size_t *off);
/* parse one NAL unit given without start code; 0 or the negative errno of the walk */
H264_API int h264_reader_parse_nalu(struct h264_reader *reader, uint32_t flags, const uint8_t *buf,
				    size_t len);
H264_API int h264_parse_nalu_header(const uint8_t *buf, size_t len, struct h264_nalu_header *nh);
H264_API int h264_parse_sps(const uint8_t *buf, size_t len, struct h264_sps *sps);
H264_API int h264_parse_pps(const uint8_t *buf, size_t len, const struct h264_sps *sps,
			    struct h264_pps *pps);

#endif /* H264B200_READER_H */
