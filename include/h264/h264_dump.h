/*
 * h264/h264_dump.h — JSON dump API of the reference (include/h264/h264_dump.h:31-73).
 * The reference implements it with json-c, which this build does not have and
 * which is off the parse/serialise hot path (out of scope): the
 * types and flags are declared for source compatibility, the functions are not
 * provided.
 */
#ifndef H264B200_DUMP_H
#define H264B200_DUMP_H

struct json_object;
struct h264_dump;

#define H264_DUMP_FLAGS_SLICE_DATA 0x01

enum h264_dump_type {
	H264_DUMP_TYPE_JSON,
};

struct h264_dump_cfg {
	enum h264_dump_type type;
};

/* JSON dump of the NAL unit held in a context (reference: include/h264/h264_dump.h:49-73).
 * It needs json-c; this library reports -ENOSYS (out of scope). */
H264_API int h264_dump_new(const struct h264_dump_cfg *cfg, struct h264_dump **ret_obj);
H264_API int h264_dump_destroy(struct h264_dump *dump);
H264_API int h264_dump_clear(struct h264_dump *dump);
H264_API int h264_dump_get_json_object(struct h264_dump *dump, struct json_object **jobj);
H264_API int h264_dump_get_json_str(struct h264_dump *dump, const char **str);
H264_API int h264_dump_nalu(struct h264_dump *dump, struct h264_ctx *ctx, uint32_t flags);

#endif /* H264B200_DUMP_H */
