/*
 * h264/h264.h — umbrella header of the libh264-compatible API served by the
 * B200-native library (libh264.so built from libh264_b200/host + libh264gpu.so).
 *
 * Same include order, export macro and top-level entry points as the reference
 * (Parrot-Developers/libh264 include/h264/h264.h:27-89).
 */
#ifndef H264B200_H
#define H264B200_H

#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef H264_API_EXPORTS
#	ifdef _WIN32
#		define H264_API __declspec(dllexport)
#	else
#		define H264_API __attribute__((visibility("default")))
#	endif
#else
#	define H264_API
#endif

#include "h264/h264_types.h"
#include "h264/h264_bitstream.h"
#include "h264/h264_ctx.h"
#include "h264/h264_dump.h"
#include "h264/h264_reader.h"
#include "h264/h264_writer.h"

/* variables derived from an SPS (reference: src/h264.c:120-208) */
H264_API int h264_get_sps_derived(const struct h264_sps *sps, struct h264_sps_derived *sps_derived);

/* parse an SPS and a PPS NAL unit and summarise them (reference: src/h264.c:36-80) */
H264_API int h264_get_info(const uint8_t *sps, size_t sps_len, const uint8_t *pps, size_t pps_len,
			   struct h264_info *info);

/* E.2.1 table lookup; 255 (Extended_SAR) when the ratio is not in the table */
H264_API int h264_sar_to_aspect_ratio_idc(unsigned int sar_width, unsigned int sar_height);

/* In-place Annex-B (4-byte start codes only) <-> AVCC (4-byte big-endian lengths)
 * conversion (reference: src/h264.c:210-272) */
H264_API int h264_byte_stream_to_avcc(uint8_t *data, size_t len);
H264_API int h264_avcc_to_byte_stream(uint8_t *data, size_t len);

#ifdef __cplusplus
}
#endif

#endif /* H264B200_H */
