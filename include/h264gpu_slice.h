/*
 * h264gpu_slice.h — C-ABI of the slice-parallel macroblock syntax parse (K4 CAVLC,
 * K5 CABAC).  One independent slice per GPU thread; the host supplies, per slice,
 * the parameter block the reference keeps in struct h264_ctx while it walks
 * slice_data (src/h264_priv.h:67-140), and gets back one record per macroblock —
 * what the reference delivers through h264_ctx_cbs.slice_data_mb
 * (include/h264/h264_ctx.h:78-82) plus a checksum of the macroblock's syntax
 * elements (the reference keeps those in the private ctx->mb,
 * src/h264_macroblock.h:105-167).
 *
 * Replaces (Parrot-Developers/libh264):
 *   _h264_read_slice_data_internal   src/h264_syntax_slice_data.h:701-787
 *   _h264_read_macroblock_layer      src/h264_syntax_slice_data.h:604-696
 *   mb_pred / sub_mb_pred            src/h264_syntax_slice_data.h:422-601
 *   residual / residual_luma / residual_block   src/h264_syntax_slice_data.h:103-419
 *   h264_read_mb_type / _sub_mb_type / _coded_block_pattern / _coeff_token /
 *   _total_zeros / _run_before, h264_new_macroblock   src/h264_slice_data.c:839-1416
 *   neighbour derivation             src/h264_macroblock.c:306-433
 *   bit reader + EPB skip            include/h264/h264_bitstream.h:168-304
 *   h264_bs_more_rbsp_data           src/h264_bitstream.c:325-355
 */
#ifndef H264GPU_SLICE_H
#define H264GPU_SLICE_H

#include "h264gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* status codes of a slice parse (negative errno like the reference, or:) */
#define H264GPU_SLICE_OK 0
#define H264GPU_SLICE_SKIPPED 1 /* CABAC slice without the CABAC flag: the reference
				    parses no slice data either (F2) */

/* Per-slice parameter block (56 bytes). */
struct h264gpu_slice_params {
	uint64_t nal_off;      /* offset of the NAL's header byte in the device stream */
	uint32_t nal_len;      /* NAL length in bytes (escaped, header included) */
	uint32_t data_bit_off; /* raw bit offset of slice_data() inside the NAL
				  (ctx->slice.hdr_len, src/h264_syntax.h:1383) */
	uint32_t first_mb_in_slice;
	uint32_t mb_out_off;   /* first record index of this slice in the record array */
	uint32_t mb_out_cap;   /* records this slice may write */
	uint32_t row_state_off; /* num_slice_groups_minus1 != 0: byte offset of the slice's macroblock ->
				   slice group map in the group map buffer (else unused) */
	uint16_t pic_width_in_mbs;
	uint16_t pic_height_in_mbs; /* PicHeightInMbs of the slice's picture */
	uint8_t slice_type;         /* 0 P, 1 B, 2 I, 3 SP, 4 SI (slice_type % 5) */
	uint8_t chroma_array_type;
	uint8_t bit_depth_luma;
	uint8_t bit_depth_chroma;
	uint8_t transform_8x8_mode_flag;
	uint8_t direct_8x8_inference_flag;
	uint8_t num_ref_idx_l0_active_minus1;
	uint8_t num_ref_idx_l1_active_minus1;
	uint8_t field_pic_flag;
	uint8_t mbaff_frame_flag;
	uint8_t entropy_coding_mode_flag;
	uint8_t num_slice_groups_minus1;
	uint8_t cabac_init_idc;
	int8_t slice_qp;            /* SliceQPY (CABAC context init) */
	uint8_t reserved[2];
};

/* One record per macroblock, in slice_data_mb callback order. */
struct h264gpu_mb_record {
	uint32_t mb_addr;
	uint32_t mb_type; /* enum h264_mb_type value */
	uint64_t hash;    /* h264gpu_mb_hash of the macroblock's syntax elements */
};

struct h264gpu_slice_result {
	int32_t status;    /* 0, H264GPU_SLICE_SKIPPED, or negative errno */
	uint32_t mb_count; /* what slice_data_end reports */
	uint64_t end_bit;  /* raw bit position in the NAL after the last macroblock */
};

/*
 * Syntax-element checksum.  Every element that the reference stores in ctx->mb
 * contributes value * weight(field, index); zero-valued elements contribute
 * nothing, so a parser only touches what it decodes.  Order independent.
 */
enum h264gpu_mb_field {
	H264GPU_F_RAW_MB_TYPE = 1,
	H264GPU_F_TRANSFORM_8X8 = 2,
	H264GPU_F_MB_QP_DELTA = 3,
	H264GPU_F_CBP = 4,
	H264GPU_F_CBP_LUMA = 5,
	H264GPU_F_CBP_CHROMA = 6,
	H264GPU_F_INTRA_CHROMA_PRED_MODE = 7,
	H264GPU_F_INTRA4X4_PRED_MODE = 8,  /* [16] */
	H264GPU_F_INTRA8X8_PRED_MODE = 9,  /* [4] */
	H264GPU_F_REF_IDX_L0 = 10,         /* [4] */
	H264GPU_F_REF_IDX_L1 = 11,
	H264GPU_F_MVD_L0 = 12,             /* [(part*4+sub)*2+comp] */
	H264GPU_F_MVD_L1 = 13,
	H264GPU_F_RAW_SUB_MB_TYPE = 14,    /* [4] */
	H264GPU_F_MB_FIELD_DECODING_FLAG = 15,
	H264GPU_F_I16_DC = 16,             /* + 3*comp : [16]            comp 0 Y,1 Cb,2 Cr */
	H264GPU_F_I16_AC = 17,             /* + 3*comp : [blk*16+i]      */
	H264GPU_F_LEVEL4X4 = 18,           /* + 3*comp : [blk*16+i]      */
	H264GPU_F_CHROMA_DC = 25,          /* [iCbCr*16+i] */
	H264GPU_F_CHROMA_AC = 26,          /* [(iCbCr*16+blk)*16+i] */
	H264GPU_F_PCM_LUMA = 27,           /* [256] */
	H264GPU_F_PCM_CHROMA = 28,         /* [iCbCr*256+i] */
	H264GPU_F_I16_PRED_MODE = 29,
};

static inline uint64_t h264gpu_mb_hash_weight(uint32_t field, uint32_t index)
{
	uint64_t key = ((uint64_t)field << 16) | index;
	return ((key + 1) * 0x9E3779B97F4A7C15ull) | 1ull;
}

static inline uint64_t h264gpu_mb_hash_term(uint32_t field, uint32_t index, int64_t value)
{
	return (uint64_t)value * h264gpu_mb_hash_weight(field, index);
}

/*
 * Parse n_slices CAVLC slices of the device-resident stream d_stream (escaped
 * Annex-B bytes; the bit reader skips emulation prevention bytes itself, exactly
 * like the reference's, keeping the two raw bytes of left context).
 *   d_params   n_slices parameter blocks
 *   d_records  record array (capacity = sum of mb_out_cap)
 *   d_results  n_slices results
 * Frame pictures, field pictures (field_pic_flag) and MBAFF frames (macroblock pairs,
 * mb_field_decoding_flag read / inherited / inferred like the reference's h264_new_macroblock,
 * src/h264_slice_data.c:1144-1205; neighbours of 6.4.12.2, src/h264_macroblock.c:110-231; in
 * MBAFF slices first_mb_in_slice counts pairs and the records carry macroblock addresses).
 * Slices of pictures with several slice groups need their map (h264gpu_cavlc_parse_fmo_dev);
 * without it, or with MBAFF and slice groups together, the slice gets status -ENOSYS.  There is no
 * CPU fallback.
 */
H264GPU_API int h264gpu_cavlc_parse_dev(h264gpu_ctx *ctx, const uint8_t *d_stream,
					uint64_t stream_len,
					const struct h264gpu_slice_params *d_params,
					uint32_t n_slices,
					struct h264gpu_mb_record *d_records,
					struct h264gpu_slice_result *d_results,
					void *stream);

/*
 * The same parse with FULL per-macroblock records: d_syntax[i] (struct h264_mb_syntax,
 * include/h264gpu_mb_syntax.h, 3108 bytes) receives every syntax element of the macroblock of
 * d_records[i] -- what the reference keeps in its private ctx->mb (src/h264_macroblock.h:105-167)
 * and walks in its dump (src/h264_dump.c:295-316): raw types, prediction modes, ref_idx, mvd,
 * cbp, qp delta, DC / AC / 4x4 levels, I_PCM samples.  d_syntax = NULL: records only.
 */
struct h264_mb_syntax;
H264GPU_API int h264gpu_cavlc_parse_full_dev(h264gpu_ctx *ctx, const uint8_t *d_stream,
					     uint64_t stream_len,
					     const struct h264gpu_slice_params *d_params,
					     uint32_t n_slices,
					     struct h264gpu_mb_record *d_records,
					     struct h264gpu_slice_result *d_results,
					     struct h264_mb_syntax *d_syntax, void *stream);

/*
 * Flexible macroblock ordering (several slice groups, 8.2.2).  The reference builds the map
 * unit -> slice group map of the picture (h264_gen_slice_group_map, src/h264_fmo.c:244-291),
 * walks the slice with h264_next_mb_addr (:308-319) and finds neighbours through its per-slice
 * macroblock table; the kernel needs the finished macroblock -> slice group map, one byte per
 * macroblock, of every slice with num_slice_groups_minus1 != 0:
 *   h264gpu_fmo_mb_map        host-side: 8.2.2.1 - 8.2.2.8 from the PPS / slice header fields
 *   d_group_maps              the maps back to back; slice i's map starts at byte
 *                             d_params[i].row_state_off (slices of one picture may share a map)
 */
struct h264gpu_fmo_desc {
	uint32_t num_slice_groups_minus1;   /* 1..7 */
	uint32_t slice_group_map_type;      /* 0..6 */
	const uint32_t *run_length_minus1;  /* type 0: [num_slice_groups_minus1 + 1] */
	const uint32_t *top_left;           /* type 2: [num_slice_groups_minus1] */
	const uint32_t *bottom_right;
	int slice_group_change_direction_flag; /* types 3..5 */
	uint32_t map_units_in_slice_group0; /* types 3..5: Min(slice_group_change_cycle *
					       SliceGroupChangeRate, PicSizeInMapUnits) */
	const uint32_t *slice_group_id;     /* type 6: [n_slice_group_id] */
	uint32_t n_slice_group_id;
	uint32_t pic_width_in_mbs, pic_height_in_map_units;
	int frame_mbs_only_flag, field_pic_flag, mbaff_frame_flag;
	uint32_t pic_size_in_mbs;           /* bytes written to mb_map */
};
H264GPU_API int h264gpu_fmo_mb_map(const struct h264gpu_fmo_desc *desc, uint8_t *mb_map);

H264GPU_API int h264gpu_cavlc_parse_fmo_dev(h264gpu_ctx *ctx, const uint8_t *d_stream,
					    uint64_t stream_len,
					    const struct h264gpu_slice_params *d_params,
					    uint32_t n_slices,
					    struct h264gpu_mb_record *d_records,
					    struct h264gpu_slice_result *d_results,
					    struct h264_mb_syntax *d_syntax, /* or NULL */
					    const uint8_t *d_group_maps, void *stream);

/* Host-buffer forms: the group maps the NEXT h264gpu_cavlc_parse_host / h264gpu_reader_parse_cavlc /
 * h264gpu_reader_parse_slices call on this context uses (copied; forgotten after that call). */
H264GPU_API int h264gpu_reader_set_group_maps(h264gpu_ctx *ctx, const uint8_t *h_maps, uint64_t bytes);

/* Host-buffer form: uploads the stream and parameter blocks, runs the kernel,
 * downloads records and results.  Synchronous. */
H264GPU_API int h264gpu_cavlc_parse_host(h264gpu_ctx *ctx, const uint8_t *h_stream,
					 uint64_t stream_len,
					 const struct h264gpu_slice_params *h_params,
					 uint32_t n_slices,
					 struct h264gpu_mb_record *h_records,
					 uint64_t n_records,
					 struct h264gpu_slice_result *h_results);

/*
 * K5: the same for CABAC slices (entropy_coding_mode_flag = 1).  The reference has no
 * counterpart: its reader returns before CABAC slice data (src/h264_syntax_slice_data.h:
 * 715-717), so slice_data_begin / slice_data_mb / slice_data_end never fire for such slices.
 * Records use the same mb_type values and checksum fields as the CAVLC parse; end_bit is the
 * raw bit position after the last bit the arithmetic decoder consumed (the stop bit).
 * CAVLC slices get H264GPU_SLICE_SKIPPED; MBAFF, several slice groups, 4:4:4 and SI slices
 * -ENOSYS.
 */
H264GPU_API int h264gpu_cabac_parse_dev(h264gpu_ctx *ctx, const uint8_t *d_stream,
					uint64_t stream_len,
					const struct h264gpu_slice_params *d_params,
					uint32_t n_slices,
					struct h264gpu_mb_record *d_records,
					struct h264gpu_slice_result *d_results,
					void *stream);

H264GPU_API int h264gpu_cabac_parse_host(h264gpu_ctx *ctx, const uint8_t *h_stream,
					 uint64_t stream_len,
					 const struct h264gpu_slice_params *h_params,
					 uint32_t n_slices,
					 struct h264gpu_mb_record *h_records,
					 uint64_t n_records,
					 struct h264gpu_slice_result *h_results);

/*
 * The same parses out of the buffer h264gpu_reader_scan / h264gpu_reader_upload left on the
 * device (include/h264gpu.h): no second upload, pooled device and pinned memory, the context's
 * own stream.  *h_records / *h_results point into the context's pinned pool and stay valid until
 * the next reader / parse call on the context.
 */
H264GPU_API int h264gpu_reader_parse_cavlc(h264gpu_ctx *ctx, const struct h264gpu_slice_params *h_params,
					   uint32_t n_slices, uint64_t n_records,
					   const struct h264gpu_mb_record **h_records,
					   const struct h264gpu_slice_result **h_results);
/* a list that mixes CAVLC and CABAC slices: one launch of each kernel over the same records */
H264GPU_API int h264gpu_reader_parse_slices(h264gpu_ctx *ctx, const struct h264gpu_slice_params *h_params,
					    uint32_t n_slices, uint64_t n_records,
					    const struct h264gpu_mb_record **h_records,
					    const struct h264gpu_slice_result **h_results);
H264GPU_API int h264gpu_reader_parse_cabac(h264gpu_ctx *ctx, const struct h264gpu_slice_params *h_params,
					   uint32_t n_slices, uint64_t n_records,
					   const struct h264gpu_mb_record **h_records,
					   const struct h264gpu_slice_result **h_results);


/*
 * N4: bulk synthesis of concealment slices, one slice per GPU lane.  What the reference writes
 * one NAL at a time with h264_write_grey_i_slice / h264_write_skipped_p_slice
 * (src/h264_writer.c:49-219; CABAC: src/h264_cabac.c, src/h264_bac.c:232-345): the NAL + slice
 * header bits come from the host syntax walk (h264_write_nalu without slice data), the slice data
 * of mb_count grey Intra16x16 / skipped macroblocks, CAVLC or CABAC, is generated here.
 *   h_hdr / d_hdr   the slices' unescaped header bytes; slice k: hdr_bits bits from byte hdr_off
 *   _dev            d_payload gets the unescaped NAL payloads back to back, d_off[0..n] their
 *                   offsets (d_off[n] = total; if it exceeds payload_cap nothing is written):
 *                   the input of h264gpu_frame_dev
 *   _host           the finished byte stream: start codes of sc_len bytes, emulation prevention
 *                   bytes inserted (h264gpu_frame_dev); h_out_off[0..n] = NAL offsets
 */
#define H264GPU_CONCEAL_GREY_I 0
#define H264GPU_CONCEAL_SKIPPED_P 1
struct h264gpu_conceal_params {
	uint64_t hdr_off;
	uint32_t hdr_bits;
	uint32_t mb_count;
	uint32_t first_mb_in_slice;
	uint16_t pic_width_in_mbs;
	uint8_t kind;                     /* H264GPU_CONCEAL_* */
	uint8_t entropy_coding_mode_flag;
	uint8_t slice_type;               /* 0 P, 1 B, 2 I, 3 SP, 4 SI */
	uint8_t cabac_init_idc;
	int8_t slice_qp;                  /* SliceQPY */
	uint8_t reserved[5];
};
H264GPU_API int h264gpu_conceal_slices_dev(h264gpu_ctx *ctx, const struct h264gpu_conceal_params *d_params,
					   uint32_t n, const uint8_t *d_hdr, uint8_t *d_payload,
					   uint64_t payload_cap, uint64_t *d_off, void *stream);
H264GPU_API int h264gpu_conceal_slices_host(h264gpu_ctx *ctx, const struct h264gpu_conceal_params *h_params,
					    uint32_t n, const uint8_t *h_hdr, uint64_t hdr_bytes, int sc_len,
					    uint8_t *h_out, uint64_t out_cap, uint64_t *h_out_off,
					    uint64_t *total);

/*
 * N4, second half: h264_rewrite_slice_header (src/h264_writer.c:311-370) in bulk on a stream that
 * is resident on the device.  The reference writes the new NAL header + slice header into a
 * 64-byte scratch bitstream (emulation prevention on), requires the same bit length as the old
 * header, copies the whole bytes over the NAL and blends the leading bits of the byte the header
 * shares with the slice data.  Here the scratch bytes come from the host syntax walk
 * (h264_rewrite_slice_header_patch, include/h264/h264_writer.h, same checks and return codes),
 * one patch per slice, and one launch applies them all:
 *     stream[nal_off + j] = bytes[j]                         for j < nbytes
 *     stream[nal_off + nbytes] = (bytes[nbytes] & ~keep) | (stream[..] & keep),
 *         keep = (1 << (8 - tail_bits)) - 1                  if tail_bits != 0
 * Patches whose bytes would reach past stream_len are skipped (status -EINVAL from the call).
 */
struct h264gpu_hdr_patch {
	uint64_t nal_off;   /* offset of the NAL's header byte in the stream */
	uint32_t nbytes;    /* whole bytes of the new header, NAL header byte included (<= 63) */
	uint32_t tail_bits; /* 0..7 header bits in the byte after them */
	uint8_t bytes[64];  /* the scratch bitstream: nbytes whole bytes, then the byte holding tail_bits */
};
H264GPU_API int h264gpu_patch_slice_headers_dev(h264gpu_ctx *ctx, uint8_t *d_stream, uint64_t stream_len,
						const struct h264gpu_hdr_patch *d_patches, uint32_t n, void *stream);
/* the same with the patches in host memory (pooled device copy, the context's own stream when
 * stream is NULL; returns after the launch has been queued and the patches have been uploaded) */
H264GPU_API int h264gpu_patch_slice_headers(h264gpu_ctx *ctx, uint8_t *d_stream, uint64_t stream_len,
					    const struct h264gpu_hdr_patch *h_patches, uint32_t n, void *stream);

#ifdef __cplusplus
}
#endif

#endif /* H264GPU_SLICE_H */
