/*
 * h264_misc.c — the small leaves of the libh264 API: enum names
 * (reference: src/h264_types.c:35-349), Annex-B <-> AVCC in-place conversion
 * (src/h264.c:184-272), and the JSON dump entry points (src/h264_dump.c), which need
 * json-c and are out of scope here: they report -ENOSYS.
 */
#include "h264_priv.h"

#include <arpa/inet.h>

static const char *lookup(const char *const *tab, size_t n, long v, const char *unknown)
{
	return v >= 0 && (size_t)v < n && tab[v] != NULL ? tab[v] : unknown;
}

const char *h264_nalu_type_str(enum h264_nalu_type val)
{
	static const char *const tab[] = {NULL,  "SLICE", "SLICE_DPA", "SLICE_DPB",  "SLICE_DPC",     "SLICE_IDR", "SEI",
					  "SPS", "PPS",   "AUD",       "END_OF_SEQ", "END_OF_STREAM", "FILLER"};
	return lookup(tab, COUNT_OF(tab), (long)val, "UNKNOWN");
}

const char *h264_slice_type_str(enum h264_slice_type val)
{
	static const char *const tab[] = {"P", "B", "I", "SP", "SI"};
	return lookup(tab, COUNT_OF(tab), (long)val, "UNKNOWN");
}

const char *h264_mb_type_str(enum h264_mb_type val)
{
	static const char *const tab[] = {NULL,     "I_NxN",  "I_16x16", "I_PCM",     "SI",     "P_16x16",
					  "P_16x8", "P_8x16", "P_8x8",   "P_8x8ref0", "P_SKIP", "B_Direct_16x16",
					  "B_16x16", "B_16x8", "B_8x16", "B_8x8",     "B_SKIP"};
	return lookup(tab, COUNT_OF(tab), (long)val, "UNKNOWN");
}

int h264_mb_type_is_intra(enum h264_mb_type val)
{
	return val >= H264_MB_TYPE_I_NxN && val <= H264_MB_TYPE_SI;
}

int h264_mb_type_is_inter(enum h264_mb_type val)
{
	return val >= H264_MB_TYPE_P_16x16 && val <= H264_MB_TYPE_B_SKIP;
}

const char *h264_profile_str(enum h264_profile val)
{
	switch (val) {
	case H264_PROFILE_CAVLC_444: return "CAVLC_444";
	case H264_PROFILE_BASELINE: return "BASELINE";
	case H264_PROFILE_MAIN: return "MAIN";
	case H264_PROFILE_EXTENDED: return "EXTENDED";
	case H264_PROFILE_HIGH: return "HIGH";
	case H264_PROFILE_HIGH_10: return "HIGH_10";
	case H264_PROFILE_HIGH_422: return "HIGH_422";
	case H264_PROFILE_HIGH_444: return "HIGH_444";
	default: return "UNKNOWN";
	}
}

const char *h264_color_format_str(enum h264_color_format val)
{
	static const char *const tab[] = {"MONO", "YUV420", "YUV422", "YUV444"};
	return lookup(tab, COUNT_OF(tab), (long)val, "UNKNOWN");
}

char *h264_aspect_ratio_str_alloc(enum h264_aspect_ratio val, uint32_t sar_width, uint32_t sar_height)
{
	static const char *const tab[] = {"UNSPECIFIED", "1:1",   "12:11", "10:11", "16:11", "40:33",
					  "24:11",       "20:11", "32:11", "80:33", "18:11", "15:11",
					  "64:33",       "160:99", "4:3",  "3:2",   "2:1"};
	char *s = NULL;
	if (val == H264_ASPECT_RATIO_EXTENDED_SAR)
		return asprintf(&s, "EXTENDED_SAR_%u:%u", sar_width, sar_height) < 0 ? NULL : s;
	return strdup(lookup(tab, COUNT_OF(tab), (long)val, "UNKNOWN"));
}

const char *h264_sei_type_str(enum h264_sei_type val)
{
	/* payloadType 0..53 in order, then 54 */
	static const char *const tab[] = {
		"BUFFERING_PERIOD", "PIC_TIMING", "PAN_SCAN_RECT", "FILLER_PAYLOAD", "USER_DATA_REGISTERED",
		"USER_DATA_UNREGISTERED", "RECOVERY_POINT", "DEC_REF_PIC_MARKING_REPETITION", "SPARE_PIC",
		"SCENE_INFO", "SUB_SEQ_INFO", "SUB_SEQ_LAYER_CHARACTERISTICS", "SUB_SEQ_CHARACTERISTICS",
		"FULL_FRAME_FREEZE", "FULL_FRAME_FREEZE_RELEASE", "FULL_FRAME_SNAPSHOT",
		"PROGRESSIVE_REFINEMENT_SEGMENT_START", "PROGRESSIVE_REFINEMENT_SEGMENT_END",
		"MOTION_CONSTRAINED_SLICE_GROUP_SET", "FILM_GRAIN_CHARACTERISTICS",
		"DEBLOCKING_FILTER_DISPLAY_PREFERENCE", "STEREO_VIDEO_INFO", "POST_FILTER_HINT",
		"TONE_MAPPING_INFO", "SCALABILITY_INFO", "SUB_PIC_SCALABLE_LAYER", "NON_REQUIRED_LAYER_REP",
		"PRIORITY_LAYER_INFO", "LAYERS_NOT_PRESENT", "LAYER_DEPENDENCY_CHANGE", "SCALABLE_NESTING",
		"BASE_LAYER_TEMPORAL_HRD", "QUALITY_LAYER_INTEGRITY_CHECK", "REDUNDANT_PIC_PROPERTY",
		"TL0_DEP_REP_INDEX", "TL_SWITCHING_POINT", "PARALLEL_DECODING_INFO", "MVC_SCALABLE_NESTING",
		"VIEW_SCALABILITY_INFO", "MULTIVIEW_SCENE_INFO", "MULTIVIEW_ACQUISITION_INFO",
		"NON_REQUIRED_VIEW_COMPONENT", "VIEW_DEPENDENCY_CHANGE", "OPERATION_POINTS_NOT_PRESENT",
		"BASE_VIEW_TEMPORAL_HRD", "FRAME_PACKING_ARRANGEMENT", "MULTIVIEW_VIEW_POSITION",
		"DISPLAY_ORIENTATION", "MVCD_SCALABLE_NESTING", "MVCD_VIEW_SCALABILITY_INFO",
		"DEPTH_REPRESENTATION_INFO", "THREE_DIMENSIONAL_REFERENCE_DISPLAYS_INFO", "DEPTH_TIMING",
		"DEPTH_SAMPLING_INFO",
		"CONSTRAINED_DEPTH_PARAMETER_SET_IDENTIFIER"};
	/* the fallback string is spelt like the reference's (src/h264_types.c:305) */
	return lookup(tab, COUNT_OF(tab), (long)val, "UNKNONW");
}

/* ---- Annex-B (4-byte start codes) <-> AVCC (4-byte big-endian lengths), in place --------- */

static size_t next_code4(const uint8_t *p, size_t len)
{
	for (size_t i = 0; i + 4 <= len; i++)
		if (p[i] == 0 && p[i + 1] == 0 && p[i + 2] == 0 && p[i + 3] == 1)
			return i;
	return SIZE_MAX;
}

int h264_byte_stream_to_avcc(uint8_t *data, size_t len)
{
	if (data == NULL || len == 0)
		return -EINVAL;
	size_t at = next_code4(data, len);
	if (at == SIZE_MAX)
		return 0; /* nothing to convert */
	data += at;
	len -= at;
	while (len > 4) {
		const size_t nxt = next_code4(data + 4, len - 4);
		const size_t nal = nxt == SIZE_MAX ? len - 4 : nxt;
		const uint32_t be = htonl((uint32_t)nal);
		memcpy(data, &be, 4);
		data += 4 + nal;
		len -= 4 + nal;
	}
	return 0;
}

int h264_avcc_to_byte_stream(uint8_t *data, size_t len)
{
	if (data == NULL || len == 0)
		return -EINVAL;
	const uint32_t code = htonl(1);
	for (size_t off = 0; off < len;) {
		uint32_t be;
		memcpy(&be, data + off, 4);
		const uint32_t nal = ntohl(be);
		if (nal == 0)
			return -EPROTO;
		memcpy(data + off, &code, 4);
		off += 4 + (size_t)nal;
	}
	return 0;
}

/* ---- JSON dump: needs json-c, not built (DESIGN.md "out of scope") ------------------------ */

H264_API int h264_dump_new(const struct h264_dump_cfg *cfg, struct h264_dump **ret_obj)
{
	(void)cfg;
	if (ret_obj != NULL)
		*ret_obj = NULL;
	return -ENOSYS;
}
H264_API int h264_dump_destroy(struct h264_dump *dump)
{
	(void)dump;
	return 0;
}
H264_API int h264_dump_clear(struct h264_dump *dump)
{
	(void)dump;
	return -ENOSYS;
}
H264_API int h264_dump_get_json_object(struct h264_dump *dump, struct json_object **jobj)
{
	(void)dump;
	(void)jobj;
	return -ENOSYS;
}
H264_API int h264_dump_get_json_str(struct h264_dump *dump, const char **str)
{
	(void)dump;
	(void)str;
	return -ENOSYS;
}
H264_API int h264_dump_nalu(struct h264_dump *dump, struct h264_ctx *ctx, uint32_t flags)
{
	(void)dump;
	(void)ctx;
	(void)flags;
	return -ENOSYS;
}
