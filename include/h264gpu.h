/*
 * h264gpu.h — C-ABI of the B200 (sm_100a) data-parallel stages of the H.264
 * bitstream front end.  Plain pointers and sizes only; no CUDA or torch types.
 *
 * Every entry point returns 0 or a negative errno, like libh264 itself
 * (reference convention: include/h264/h264_reader.h, src/h264_reader.c:66-143).
 *
 * What each stage replaces in the reference (Parrot-Developers/libh264):
 *
 *   h264gpu_split_strip_*   Annex-B start-code scan + NAL table + (optionally)
 *                           emulation-prevention-byte removal.
 *                           Replaces the NAL loop of h264_reader_parse
 *                           (src/h264_reader.c:133-140) over h264_find_nalu
 *                           (src/h264_bitstream.c:159-184; start/end code search
 *                           :87-154) and the per-byte EPB test of h264_bs_fetch
 *                           (include/h264/h264_bitstream.h:168-190).
 *   h264gpu_frame_*         Writer side: EPB insertion + start-code framing of
 *                           RBSP payloads.  Replaces h264_bs_flush
 *                           (src/h264_bitstream.c:54-81) driven by
 *                           h264_bs_write_bits (:211-239); the 4-byte start code
 *                           is the one h264_avcc_to_byte_stream writes
 *                           (src/h264.c:251-272).
 *   h264gpu_cavlc_*         Slice-parallel CAVLC macroblock syntax parse.
 *                           Replaces _h264_read_slice_data_internal
 *                           (src/h264_syntax_slice_data.h:701-787) and callees
 *                           (see include/h264gpu_slice.h).
 *
 * Device pointers ("d_*") must come from the same device the context was made
 * on and be 16-byte aligned.  All *_dev calls are asynchronous on `stream`
 * (a cudaStream_t passed as void*, NULL = the legacy default stream).
 *
 * Concurrency: a context owns ONE set of workspaces (tickets, chain words, event buffers,
 * context rings) shared by all its entry points, so at most one operation may be in flight
 * per context; use one context per stream / thread for concurrent work (contexts are cheap:
 * workspaces are allocated on first use and grow only).
 */
#ifndef H264GPU_H
#define H264GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define H264GPU_API __attribute__((visibility("default")))

#define H264GPU_NONE UINT64_MAX

typedef struct h264gpu_ctx h264gpu_ctx;

/* Neighbour bytes of a byte-range shard (multi-GPU byte-range partitioning,
 * SURVEY.md §8e).  A shard that is the whole buffer passes NULL. */
struct h264gpu_shard_edge {
	uint8_t left[2];   /* the 2 bytes before the shard (left[1] is adjacent) */
	uint8_t right[2];  /* the 2 bytes after the shard  (right[0] is adjacent) */
	uint8_t has_left;  /* 0: shard starts the stream */
	uint8_t has_right; /* 0: shard ends the stream (end-of-buffer rule applies) */
	uint8_t assume_in; /* 1: bytes before the shard's first boundary event are
			      assumed to lie inside a NAL (resolved at merge time) */
	uint8_t pad;
};

/* Result of one scan(+strip) launch; lives in device or host memory. */
struct h264gpu_scan_result {
	uint64_t n_nal;           /* NALs whose start code begins in this shard */
	uint64_t rbsp_bytes;      /* bytes written to d_rbsp */
	uint64_t first_event_pos; /* absolute offset of the shard's first boundary
				     event, H264GPU_NONE when it has none */
	uint64_t head_bytes;      /* RBSP bytes emitted before that first event
				     (only non-zero with assume_in) */
	uint32_t first_event_is_sc; /* that event is a start code */
	uint32_t any_event;
	uint32_t end_open;        /* the last boundary event is a start code: the
				     last NAL runs to the shard end / next shard */
	uint32_t reserved;
};

H264GPU_API int h264gpu_device_count(void);
H264GPU_API int h264gpu_create(int device, h264gpu_ctx **out);
H264GPU_API int h264gpu_destroy(h264gpu_ctx *ctx);
H264GPU_API int h264gpu_device(const h264gpu_ctx *ctx);
H264GPU_API const char *h264gpu_version(void);
/* Number of kernels this library launched on ctx since creation. */
H264GPU_API uint64_t h264gpu_launch_count(const h264gpu_ctx *ctx);

/* Plain device/pinned memory helpers so a C host needs no CUDA headers. */
H264GPU_API int h264gpu_malloc(h264gpu_ctx *ctx, size_t bytes, void **d_ptr);
H264GPU_API int h264gpu_free(h264gpu_ctx *ctx, void *d_ptr);
H264GPU_API int h264gpu_host_alloc(h264gpu_ctx *ctx, size_t bytes, void **h_ptr);
H264GPU_API int h264gpu_host_free(h264gpu_ctx *ctx, void *h_ptr);
H264GPU_API int h264gpu_memcpy_h2d(h264gpu_ctx *ctx, void *d_dst, const void *h_src,
				   size_t bytes, void *stream);
H264GPU_API int h264gpu_memcpy_d2h(h264gpu_ctx *ctx, void *h_dst, const void *d_src,
				   size_t bytes, void *stream);
H264GPU_API int h264gpu_sync(h264gpu_ctx *ctx, void *stream);
/* A non-blocking stream of the context's device, as the opaque `stream` argument of the calls above. */
H264GPU_API int h264gpu_stream_create(h264gpu_ctx *ctx, void **stream);
H264GPU_API int h264gpu_stream_destroy(h264gpu_ctx *ctx, void *stream);

/*
 * Annex-B scan (+ EPB strip when d_rbsp != NULL) of d_in[0,len).
 *   base       absolute offset of d_in[0] in the whole stream; every offset
 *              written to the tables is base-relative + base
 *   d_nal_*    tables of nal_cap entries: start (first byte after the start
 *              code), end (exclusive), rbsp (offset of the NAL's RBSP in d_rbsp;
 *              may be NULL when d_rbsp is NULL)
 *   d_result   struct h264gpu_scan_result in device memory
 * NAL k's end is written by whoever sees the boundary event that closes it; the
 * last NAL's end is len+base when the shard ends the stream and is left
 * untouched (end_open=1) when it does not.
 */
H264GPU_API int h264gpu_split_strip_dev(h264gpu_ctx *ctx, const uint8_t *d_in,
					uint64_t len, uint64_t base,
					const struct h264gpu_shard_edge *edge,
					uint8_t *d_rbsp, uint64_t *d_nal_start,
					uint64_t *d_nal_end, uint64_t *d_nal_rbsp,
					uint64_t nal_cap,
					struct h264gpu_scan_result *d_result,
					void *stream);

/*
 * The same scan + strip with the RBSP of every NAL written IN PLACE: NAL k's RBSP is
 * d_rbsp[d_nal_rbsp[k] .. d_nal_rbsp[k] + d_nal_rbsp_len[k]) with d_nal_rbsp[k] =
 * d_nal_start[k] - base, i.e. where the NAL's first byte is in d_in; the bytes of d_rbsp
 * between two NALs' RBSPs are unspecified.  d_rbsp needs len bytes.  Same reference behaviour
 * per NAL (h264_find_nalu bounds, src/h264_bitstream.c:159-184; the byte sequence
 * h264_bs_read_bits(8) yields, include/h264/h264_bitstream.h:168-218), but a byte's output
 * position depends only on its own NAL, which removes the stream-long dependency chain of the
 * packed form.  result->rbsp_bytes = sum of the RBSP lengths; the bytes a shard
 * holds before its first boundary event (a NAL continued from the previous shard) are written
 * at d_rbsp[0 .. head_bytes).  The last NAL's length runs to the shard end when end_open.
 * result->reserved != 0: more boundary events than the workspace holds (4 x nal_cap + 4096),
 * the table is incomplete -> pass a larger nal_cap.
 */
H264GPU_API int h264gpu_split_strip_inplace_dev(h264gpu_ctx *ctx, const uint8_t *d_in,
						uint64_t len, uint64_t base,
						const struct h264gpu_shard_edge *edge,
						uint8_t *d_rbsp, uint64_t *d_nal_start,
						uint64_t *d_nal_end, uint64_t *d_nal_rbsp,
						uint64_t *d_nal_rbsp_len, uint64_t nal_cap,
						struct h264gpu_scan_result *d_result,
						void *stream);

/*
 * Reserve the workspace of h264gpu_split_strip_inplace_dev for streams of up to len bytes and
 * nal_cap NAL units now, instead of at the first call (which otherwise pays a device-wide
 * synchronisation and the allocation).  Until the span tickets of the kernel were spread over
 * several counters, WHEN the buffer was allocated also decided 2.0 against 2.8 - 2.9 ms per 4 GiB:
 * the one ticket word was served at the rate of the L2 slice it happened to land on
 * (profiles/r02_scan_workspace_placement.txt, profiles/r02_scan7_ticket_counters.txt).  With the
 * counters the launch takes 2.05 ms wherever the buffer is; the call is kept for the first reason.
 */
H264GPU_API int h264gpu_scan_reserve(h264gpu_ctx *ctx, uint64_t len, uint64_t nal_cap);

/*
 * Host-side merge of per-shard results (byte-range shards of one stream, in
 * stream order: chunks of the host pipeline, or one shard per GPU).  No device
 * work and no collective: each shard contributes its small result struct and
 * its NAL table.
 *
 *   h264gpu_merge_shard   call once per shard AFTER appending that shard's
 *                         table entries (shard_nals_copied of them) to the
 *                         merged tables at index m->n_nal.  Closes the NAL left
 *                         open by earlier shards, rebases the appended RBSP
 *                         offsets, and tells the caller which slice
 *                         [rbsp_skip, rbsp_skip+rbsp_take) of the shard's RBSP
 *                         output belongs to the merged RBSP (appended at the
 *                         previous m->rbsp_bytes).
 *   h264gpu_merge_finish  closes the last NAL at stream_len and yields the
 *                         *off value of h264_reader_parse (src/h264_reader.c:139).
 */
struct h264gpu_merge {
	uint64_t n_nal;
	uint64_t rbsp_bytes;
	uint32_t open;   /* a NAL is still open at the end of the shards seen so far */
	uint32_t shards;
};

H264GPU_API void h264gpu_merge_init(struct h264gpu_merge *m);
H264GPU_API int h264gpu_merge_shard(struct h264gpu_merge *m,
				    const struct h264gpu_scan_result *r,
				    uint64_t *tab_start, uint64_t *tab_end,
				    uint64_t *tab_rbsp, uint64_t tab_cap,
				    uint64_t shard_nals_copied, uint64_t *rbsp_skip,
				    uint64_t *rbsp_take);
/*
 * Merge step for shards scanned by h264gpu_split_strip_inplace_dev.  RBSP offsets stay
 * shard-relative (NAL k's RBSP is in the d_rbsp of the shard its start lies in); what crosses a
 * seam is the NAL left open by the earlier shards: its length grows by this shard's head
 * bytes (*carry_len, found at the start of this shard's d_rbsp, or at offset nal_start - base
 * when the NAL's start code straddles the seam) and its end becomes this
 * shard's first boundary event.  tab_end / tab_rbsp_len: the merged tables so far.
 */
H264GPU_API int h264gpu_merge_shard_inplace(struct h264gpu_merge *m,
					    const struct h264gpu_scan_result *r,
					    uint64_t *tab_end, uint64_t *tab_rbsp_len,
					    uint64_t tab_cap, uint64_t *carry_len);
H264GPU_API int h264gpu_merge_finish(struct h264gpu_merge *m, uint64_t stream_len,
				     uint64_t *tab_end, uint64_t tab_cap,
				     uint64_t *final_off);

/*
 * Host-buffer form (the call a libh264 user's buffer goes through): copies
 * h_in to the device in chunks overlapped with the kernel, and copies the
 * tables (and RBSP when h_rbsp != NULL) back.  Synchronous.
 *   *n_nal      in: capacity of the h_nal_* arrays; out: NALs found
 *   *final_off  what h264_reader_parse would leave in *off
 * Returns -ENOBUFS when the tables were too small (n_nal still set).
 */
H264GPU_API int h264gpu_split_strip_host(h264gpu_ctx *ctx, const uint8_t *h_in,
					 uint64_t len, uint8_t *h_rbsp,
					 uint64_t *h_nal_start, uint64_t *h_nal_end,
					 uint64_t *h_nal_rbsp, uint64_t *n_nal,
					 uint64_t *rbsp_bytes, uint64_t *final_off);

/*
 * Reader session: the device-resident path h264_reader_parse drives.  The buffer is copied to
 * the GPU ONCE (pooled device memory, the context's own stream; H264GPU_REGISTER_INPUT=1
 * page-locks the caller's buffer for the copy), the scan-only kernel builds the NAL table
 * (replaces the h264_find_nalu loop, src/h264_reader.c:133-140) and the slice kernels then
 * parse out of the same resident copy (include/h264gpu_slice.h: h264gpu_reader_parse_cavlc).
 * *h_start / *h_end point into pooled pinned memory of the context and stay valid until the
 * next h264gpu_reader_* or *_parse_host call.  *final_off: what h264_reader_parse leaves in
 * *off (src/h264_reader.c:139).  One operation per context at a time: the pools and the
 * workspaces are shared by every entry point of a context.
 */
H264GPU_API int h264gpu_reader_scan(h264gpu_ctx *ctx, const uint8_t *h_buf, uint64_t len,
				    const uint64_t **h_start, const uint64_t **h_end,
				    uint64_t *n_nal, uint64_t *final_off);
H264GPU_API int h264gpu_reader_upload(h264gpu_ctx *ctx, const uint8_t *h_buf, uint64_t len);
/* the copy left on the device by h264gpu_reader_scan / _upload */
H264GPU_API int h264gpu_reader_resident(h264gpu_ctx *ctx, const uint8_t **d_stream, uint64_t *len);

/*
 * h264_byte_stream_to_avcc (src/h264.c:210-246) on the GPU: every 4-byte start code
 * 00 00 00 01 of d_data[0,len) is replaced, in place, by the big-endian length of the NAL
 * behind it (up to the next 4-byte code or the end of the buffer; a code within the last 4
 * bytes stays, as in the reference's `while (len > 4)`).  3-byte codes are not start codes here,
 * exactly like the reference's find_start_code (src/h264.c:184-207).
 *   max_nal   capacity of d_pos and of the scan's code table (0 = worst case len / 4)
 *   d_pos     optional: offsets of the codes found, in order;  d_count: how many were found
 *             (if *d_count > max_nal the table overflowed and the conversion is incomplete)
 * The host form uploads the buffer, fetches the code offsets and patches the 4-byte length
 * fields in h_data itself (nothing else changes, so the buffer is not copied back).
 * h264_avcc_to_byte_stream is a pointer chase (each length gives the next position, one
 * dependent load per NAL) and stays on the host (libh264.so).
 */
H264GPU_API int h264gpu_byte_stream_to_avcc_dev(h264gpu_ctx *ctx, uint8_t *d_data, uint64_t len,
						uint64_t max_nal, uint64_t *d_pos, uint64_t *d_count,
						void *stream);
H264GPU_API int h264gpu_byte_stream_to_avcc_host(h264gpu_ctx *ctx, uint8_t *h_data, uint64_t len,
						 uint64_t *n_nal);

/*
 * Writer side: escape n payloads d_rbsp[d_off[k], d_off[k+1]) and put a start
 * code of sc_len (3 or 4; 0 = none) bytes in front of each.
 *   d_out       capacity out_cap bytes (worst case 3/2 * payload + n*sc_len)
 *   d_out_off   n+1 entries: offset of NAL k's start code; entry n = total
 *   d_total     one uint64: total bytes the output needs (also when > out_cap,
 *               in which case nothing beyond out_cap is written)
 */
H264GPU_API int h264gpu_frame_dev(h264gpu_ctx *ctx, const uint8_t *d_rbsp,
				  const uint64_t *d_off, uint64_t n, int sc_len,
				  uint8_t *d_out, uint64_t out_cap,
				  uint64_t *d_out_off, uint64_t *d_total,
				  void *stream);

/*
 * The same from and to host buffers (what a caller of h264_write_nalu over a list of payloads
 * has, src/h264_writer.c:240-243).  Inputs of two chunks (2 x 128 MiB) or more are cut into chunks
 * of whole payloads and go through two device slots on three streams, so that the upload of one
 * chunk, the kernel of the one before and the download of the one before that overlap; shorter
 * inputs take one upload, one launch, one download.  Page-locked h_rbsp / h_out give the full
 * rate (38 GB/s of payload on the measured box), pageable memory works (10 GB/s).
 * Returns -ENOBUFS (with *total = the bytes needed and nothing written past out_cap) when the
 * output does not fit.
 */
H264GPU_API int h264gpu_frame_host(h264gpu_ctx *ctx, const uint8_t *h_rbsp,
				   const uint64_t *h_off, uint64_t n, int sc_len,
				   uint8_t *h_out, uint64_t out_cap,
				   uint64_t *h_out_off, uint64_t *total);

/* Event timing on a stream, for callers without CUDA headers (bench harness). */
H264GPU_API int h264gpu_timer_create(h264gpu_ctx *ctx, void **timer);
H264GPU_API int h264gpu_timer_destroy(h264gpu_ctx *ctx, void *timer);
H264GPU_API int h264gpu_timer_start(h264gpu_ctx *ctx, void *timer, void *stream);
H264GPU_API int h264gpu_timer_stop(h264gpu_ctx *ctx, void *timer, void *stream);
/* Synchronises on the stop event; elapsed milliseconds. */
H264GPU_API int h264gpu_timer_elapsed_ms(h264gpu_ctx *ctx, void *timer, float *ms);

#ifdef __cplusplus
}
#endif

#endif /* H264GPU_H */
