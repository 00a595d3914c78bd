/*
 * ref_harness.c — thin driver around the UNMODIFIED reference (libh264) so tests
 * and the CPU-baseline leg of bench.py can call it through ctypes.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled together with the reference sources (read
 * in place from /root/reference, never copied) into oracle/_ref/libh264_ref.so
 * by oracle/Makefile.  All work is done by reference functions; this file only
 * loops over them the way the reference's own callers do:
 *
 *   ref_scan    = the NAL loop of h264_reader_parse (src/h264_reader.c:133-140)
 *                 with h264_reader_parse_nalu replaced by "record (start,end)"
 *   ref_strip   = h264_bs_cinit(..,1) + h264_bs_read_bits(bs,&v,8) until failure
 *   ref_insert  = h264_bs_init(NULL,0,1) + h264_bs_write_bits(bs,byte,8) per byte
 *   ref_mt_*    = the same loops on T threads over byte ranges / NAL ranges
 *                 (the library is re-entrant per object)
 */
#include "h264_priv.h"

#include <pthread.h>
#include <time.h>

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

size_t ref_scan(const uint8_t *buf, size_t len, uint64_t *starts, uint64_t *ends,
		size_t cap, uint64_t *final_off)
{
	size_t off = 0, start = 0, end = 0, n = 0;
	while (off < len) {
		int res = h264_find_nalu(buf + off, len - off, &start, &end);
		if (res < 0 && res != -EAGAIN)
			break;
		if (n < cap) {
			starts[n] = off + start;
			ends[n] = off + end;
		}
		n++;
		off += end;
	}
	if (final_off)
		*final_off = off;
	return n;
}

size_t ref_strip(const uint8_t *nal, size_t len, uint8_t *out, uint64_t *raw_off_at_fail)
{
	struct h264_bitstream bs;
	uint32_t v = 0;
	size_t o = 0;
	h264_bs_cinit(&bs, nal, len, 1);
	while (h264_bs_read_bits(&bs, &v, 8) == 8)
		out[o++] = (uint8_t)v;
	if (raw_off_at_fail)
		*raw_off_at_fail = bs.off;
	return o;
}

/* returns escaped length, or (size_t)-1 if cap is too small */
size_t ref_insert(const uint8_t *rbsp, size_t len, uint8_t *out, size_t cap)
{
	struct h264_bitstream bs;
	size_t n;
	h264_bs_init(&bs, NULL, 0, 1);
	for (size_t i = 0; i < len; i++)
		h264_bs_write_bits(&bs, rbsp[i], 8);
	n = bs.off;
	if (n <= cap && n > 0)
		memcpy(out, bs.data, n);
	h264_bs_clear(&bs);
	return n <= cap ? n : (size_t)-1;
}

/* ------------------------------------------------------------------------- */
/* Multi-threaded CPU baselines (timed inside, seconds returned).             */

struct mt_job {
	const uint8_t *buf;
	size_t len;
	size_t lo, hi; /* this thread owns NALs whose start code begins in [lo,hi) */
	uint8_t *out;  /* scratch, at least hi-lo+8 bytes, or NULL for scan-only */
	uint64_t nals, rbsp_bytes, in_bytes;
};

static void *mt_split_strip_worker(void *arg)
{
	struct mt_job *j = arg;
	size_t off = j->lo, start = 0, end = 0;
	/* back up so a start code straddling lo is seen by exactly one thread:
	 * a code is owned by the range holding its first 00 of "00 00 01" */
	while (off < j->len) {
		int res = h264_find_nalu(j->buf + off, j->len - off, &start, &end);
		if (res < 0 && res != -EAGAIN)
			break;
		/* first byte of the 3-byte code */
		size_t sc = off + start - 3;
		if (sc >= j->hi)
			break;
		j->nals++;
		j->in_bytes += end - start;
		if (j->out != NULL) {
			struct h264_bitstream bs;
			uint32_t v = 0;
			size_t o = 0;
			h264_bs_cinit(&bs, j->buf + off + start, end - start, 1);
			while (h264_bs_read_bits(&bs, &v, 8) == 8)
				j->out[o++] = (uint8_t)v;
			j->rbsp_bytes += o;
		}
		off += end;
	}
	return NULL;
}

/*
 * Scan (+ strip when do_strip) of buf on nthreads threads, byte range i of
 * nthreads each.  scratch must hold len bytes when do_strip.  Returns seconds.
 */
double ref_mt_split_strip(const uint8_t *buf, size_t len, int nthreads, int do_strip,
			  uint8_t *scratch, uint64_t *nals, uint64_t *rbsp_bytes)
{
	pthread_t th[256];
	struct mt_job jobs[256];
	double t0, t1;
	if (nthreads < 1)
		nthreads = 1;
	if (nthreads > 256)
		nthreads = 256;
	for (int i = 0; i < nthreads; i++) {
		memset(&jobs[i], 0, sizeof(jobs[i]));
		jobs[i].buf = buf;
		jobs[i].len = len;
		jobs[i].lo = (size_t)((unsigned __int128)len * i / nthreads);
		jobs[i].hi = (size_t)((unsigned __int128)len * (i + 1) / nthreads);
		jobs[i].out = do_strip ? scratch + jobs[i].lo : NULL;
	}
	t0 = now_s();
	for (int i = 1; i < nthreads; i++)
		pthread_create(&th[i], NULL, mt_split_strip_worker, &jobs[i]);
	mt_split_strip_worker(&jobs[0]);
	for (int i = 1; i < nthreads; i++)
		pthread_join(th[i], NULL);
	t1 = now_s();
	if (nals) {
		*nals = 0;
		for (int i = 0; i < nthreads; i++)
			*nals += jobs[i].nals;
	}
	if (rbsp_bytes) {
		*rbsp_bytes = 0;
		for (int i = 0; i < nthreads; i++)
			*rbsp_bytes += jobs[i].rbsp_bytes;
	}
	return t1 - t0;
}

struct mt_ins_job {
	const uint8_t *rbsp;
	const uint64_t *off;
	size_t k0, k1;
	uint64_t out_bytes;
};

static void *mt_insert_worker(void *arg)
{
	struct mt_ins_job *j = arg;
	for (size_t k = j->k0; k < j->k1; k++) {
		struct h264_bitstream bs;
		h264_bs_init(&bs, NULL, 0, 1);
		for (uint64_t i = j->off[k]; i < j->off[k + 1]; i++)
			h264_bs_write_bits(&bs, j->rbsp[i], 8);
		j->out_bytes += bs.off;
		h264_bs_clear(&bs);
	}
	return NULL;
}

/* EPB insertion of n payloads on nthreads threads (payload ranges). Seconds. */
double ref_mt_insert(const uint8_t *rbsp, const uint64_t *off, size_t n, int nthreads,
		     uint64_t *out_bytes)
{
	pthread_t th[256];
	struct mt_ins_job jobs[256];
	double t0, t1;
	if (nthreads < 1)
		nthreads = 1;
	if (nthreads > 256)
		nthreads = 256;
	for (int i = 0; i < nthreads; i++) {
		jobs[i].rbsp = rbsp;
		jobs[i].off = off;
		jobs[i].k0 = n * i / nthreads;
		jobs[i].k1 = n * (i + 1) / nthreads;
		jobs[i].out_bytes = 0;
	}
	t0 = now_s();
	for (int i = 1; i < nthreads; i++)
		pthread_create(&th[i], NULL, mt_insert_worker, &jobs[i]);
	mt_insert_worker(&jobs[0]);
	for (int i = 1; i < nthreads; i++)
		pthread_join(th[i], NULL);
	t1 = now_s();
	if (out_bytes) {
		*out_bytes = 0;
		for (int i = 0; i < nthreads; i++)
			*out_bytes += jobs[i].out_bytes;
	}
	return t1 - t0;
}

/* ------------------------------------------------------------------------- */
/* Full-reader trace: every callback of h264_reader_parse, in order, plus the
 * slice parameter block and per-macroblock syntax checksum the GPU path must
 * reproduce.  The reference does all the parsing; this only records.        */

#include "h264gpu_slice.h"
#include "../libh264_b200/csrc/mb_syntax.h"

enum {
	TR_NALU_BEGIN = 1, /* u32 type, u32 ref_idc, u64 off, u64 len */
	TR_NALU_END = 2,   /* same */
	TR_AU_END = 3,
	TR_SPS = 4,        /* struct h264_sps bytes */
	TR_PPS = 5,        /* struct h264_pps bytes */
	TR_SLICE = 6,      /* struct h264_slice_header bytes */
	TR_SLICE_DATA_BEGIN = 7,
	TR_SLICE_DATA_END = 8, /* u32 mb_count */
	TR_AUD = 9,        /* u32 primary_pic_type */
	TR_SEI = 10,       /* u32 type, raw payload bytes */
	TR_SLICE_PARAMS = 11, /* struct h264gpu_slice_params (off/out fields filled) */
	TR_GROUP_MAP = 12, /* slices of pictures with several slice groups: the reference's macroblock ->
			      slice group map (ctx->slice.group_map through 8.2.2.8), one byte per macroblock */
};

struct ref_tracer {
	const uint8_t *base;
	uint8_t *log;
	size_t log_cap, log_len;
	struct h264gpu_mb_record *mbs;
	size_t mb_cap, mb_n;
	int overflow;
	uint64_t cur_nal_off;
	uint64_t cur_nal_len;
	size_t nal_mb0; /* records written before the current NAL */
	struct h264_mb_syntax *syn; /* optional: full syntax elements per macroblock */
};

static void tr_put(struct ref_tracer *t, uint32_t tag, const void *p, uint32_t n)
{
	if (t->log_len + 8 + n > t->log_cap) {
		t->overflow = 1;
		return;
	}
	memcpy(t->log + t->log_len, &tag, 4);
	memcpy(t->log + t->log_len + 4, &n, 4);
	if (n)
		memcpy(t->log + t->log_len + 8, p, n);
	t->log_len += 8 + ((n + 7) & ~7u);
}

static uint64_t mb_hash(const struct h264_macroblock *mb)
{
	uint64_t h = 0;
#define T(f, i, v) h += h264gpu_mb_hash_term((f), (uint32_t)(i), (int64_t)(v))
	T(H264GPU_F_RAW_MB_TYPE, 0, mb->raw_mb_type);
	T(H264GPU_F_TRANSFORM_8X8, 0, mb->transform_size_8x8_flag);
	T(H264GPU_F_MB_QP_DELTA, 0, mb->mb_qp_delta);
	T(H264GPU_F_CBP, 0, mb->coded_block_pattern);
	T(H264GPU_F_CBP_LUMA, 0, mb->CodedBlockPatternLuma);
	T(H264GPU_F_CBP_CHROMA, 0, mb->CodedBlockPatternChroma);
	T(H264GPU_F_INTRA_CHROMA_PRED_MODE, 0, mb->intra_chroma_pred_mode);
	T(H264GPU_F_I16_PRED_MODE, 0, mb->Intra16x16PredMode);
	T(H264GPU_F_MB_FIELD_DECODING_FLAG, 0, mb->mb_field_decoding_flag);
	for (int i = 0; i < 16; i++)
		T(H264GPU_F_INTRA4X4_PRED_MODE, i, mb->intra4x4_pred_mode[i]);
	for (int i = 0; i < 4; i++) {
		T(H264GPU_F_INTRA8X8_PRED_MODE, i, mb->intra8x8_pred_mode[i]);
		T(H264GPU_F_REF_IDX_L0, i, mb->ref_idx_l0[i]);
		T(H264GPU_F_REF_IDX_L1, i, mb->ref_idx_l1[i]);
		T(H264GPU_F_RAW_SUB_MB_TYPE, i, mb->raw_sub_mb_type[i]);
	}
	for (int p = 0; p < 4; p++)
		for (int s = 0; s < 4; s++)
			for (int c = 0; c < 2; c++) {
				T(H264GPU_F_MVD_L0, (p * 4 + s) * 2 + c, mb->mvd_l0[p][s][c]);
				T(H264GPU_F_MVD_L1, (p * 4 + s) * 2 + c, mb->mvd_l1[p][s][c]);
			}
	const int16_t(*dc[3])[16] = {&mb->Intra16x16DCLevel, &mb->CbIntra16x16DCLevel,
				     &mb->CrIntra16x16DCLevel};
	const int16_t(*ac[3])[16][15] = {&mb->Intra16x16ACLevel, &mb->CbIntra16x16ACLevel,
					 &mb->CrIntra16x16ACLevel};
	const int16_t(*l4[3])[16][16] = {&mb->LumaLevel4x4, &mb->CbLevel4x4, &mb->CrLevel4x4};
	for (int comp = 0; comp < 3; comp++) {
		for (int i = 0; i < 16; i++)
			T(H264GPU_F_I16_DC + 3 * comp, i, (*dc[comp])[i]);
		for (int b = 0; b < 16; b++) {
			for (int i = 0; i < 15; i++)
				T(H264GPU_F_I16_AC + 3 * comp, b * 16 + i, (*ac[comp])[b][i]);
			for (int i = 0; i < 16; i++)
				T(H264GPU_F_LEVEL4X4 + 3 * comp, b * 16 + i, (*l4[comp])[b][i]);
		}
	}
	for (int c = 0; c < 2; c++) {
		for (int i = 0; i < 16; i++)
			T(H264GPU_F_CHROMA_DC, c * 16 + i, mb->ChromaDCLevel[c][i]);
		for (int b = 0; b < 16; b++)
			for (int i = 0; i < 15; i++)
				T(H264GPU_F_CHROMA_AC, (c * 16 + b) * 16 + i, mb->ChromaACLevel[c][b][i]);
		for (int i = 0; i < 256; i++)
			T(H264GPU_F_PCM_CHROMA, c * 256 + i, mb->pcm_sample_chroma[c][i]);
	}
	for (int i = 0; i < 256; i++)
		T(H264GPU_F_PCM_LUMA, i, mb->pcm_sample_luma[i]);
#undef T
	return h;
}

/* copy the reference's ctx->mb (src/h264_macroblock.h:105-167) into the coder-neutral record */
static void mb_syntax_from_ref(const struct h264_macroblock *mb, uint32_t mb_addr,
			       enum h264_mb_type mb_type, struct h264_mb_syntax *o)
{
	memset(o, 0, sizeof(*o));
	o->mb_addr = mb_addr;
	o->mb_type = (uint32_t)mb_type;
	o->raw_mb_type = mb->raw_mb_type;
	o->mb_qp_delta = mb->mb_qp_delta;
	o->transform_size_8x8_flag = (uint8_t)mb->transform_size_8x8_flag;
	o->intra_chroma_pred_mode = mb->intra_chroma_pred_mode;
	o->cbp_luma = mb->CodedBlockPatternLuma;
	o->cbp_chroma = mb->CodedBlockPatternChroma;
	for (int i = 0; i < 16; i++)
		o->intra4x4_pred_mode[i] = mb->intra4x4_pred_mode[i];
	for (int i = 0; i < 4; i++) {
		o->raw_sub_mb_type[i] = mb->raw_sub_mb_type[i];
		o->intra8x8_pred_mode[i] = mb->intra8x8_pred_mode[i];
		o->ref_idx[0][i] = mb->ref_idx_l0[i];
		o->ref_idx[1][i] = mb->ref_idx_l1[i];
		for (int s = 0; s < 4; s++)
			for (int c = 0; c < 2; c++) {
				o->mvd[0][i * 4 + s][c] = mb->mvd_l0[i][s][c];
				o->mvd[1][i * 4 + s][c] = mb->mvd_l1[i][s][c];
			}
	}
	for (int i = 0; i < 16; i++)
		o->dc16[i] = mb->Intra16x16DCLevel[i];
	for (int b = 0; b < 16; b++) {
		for (int i = 0; i < 15; i++)
			o->ac16[b][i] = mb->Intra16x16ACLevel[b][i];
		for (int i = 0; i < 16; i++)
			o->l4[b][i] = mb->LumaLevel4x4[b][i];
	}
	for (int c = 0; c < 2; c++) {
		for (int i = 0; i < 16; i++)
			o->cdc[c][i] = mb->ChromaDCLevel[c][i];
		for (int b = 0; b < 16; b++)
			for (int i = 0; i < 15; i++)
				o->cac[c][b][i] = mb->ChromaACLevel[c][b][i];
	}
	for (int i = 0; i < 256; i++)
		o->pcm[i] = (uint8_t)mb->pcm_sample_luma[i];
	for (int c = 0; c < 2; c++)
		for (int i = 0; i < 256; i++)
			o->pcm[256 + c * 256 + i] = (uint8_t)mb->pcm_sample_chroma[c][i];
}

static void tr_nalu(struct ref_tracer *t, uint32_t tag, enum h264_nalu_type type,
		    const uint8_t *buf, size_t len, const struct h264_nalu_header *nh)
{
	uint64_t rec[3] = {(uint64_t)type | (uint64_t)nh->nal_ref_idc << 32,
			   (uint64_t)(buf - t->base), (uint64_t)len};
	tr_put(t, tag, rec, sizeof(rec));
}

static void cb_nalu_begin(struct h264_ctx *ctx, enum h264_nalu_type type, const uint8_t *buf,
			  size_t len, const struct h264_nalu_header *nh, void *ud)
{
	struct ref_tracer *t = ud;
	t->cur_nal_off = (uint64_t)(buf - t->base);
	t->cur_nal_len = len;
	t->nal_mb0 = t->mb_n;
	tr_nalu(t, TR_NALU_BEGIN, type, buf, len, nh);
}

static void cb_nalu_end(struct h264_ctx *ctx, enum h264_nalu_type type, const uint8_t *buf,
			size_t len, const struct h264_nalu_header *nh, void *ud)
{
	tr_nalu(ud, TR_NALU_END, type, buf, len, nh);
}

static void cb_au_end(struct h264_ctx *ctx, void *ud)
{
	tr_put(ud, TR_AU_END, NULL, 0);
}

static void cb_sps(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
		   const struct h264_sps *sps, void *ud)
{
	tr_put(ud, TR_SPS, sps, sizeof(*sps));
}

static void cb_pps(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
		   const struct h264_pps *pps, void *ud)
{
	tr_put(ud, TR_PPS, pps, sizeof(*pps));
}

static void cb_aud(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
		   const struct h264_aud *aud, void *ud)
{
	tr_put(ud, TR_AUD, aud, sizeof(*aud));
}

static void cb_sei(struct h264_ctx *ctx, enum h264_sei_type type, const uint8_t *buf,
		   size_t len, void *ud)
{
	struct ref_tracer *t = ud;
	uint32_t ty = type;
	if (t->log_len + 16 + len > t->log_cap) {
		t->overflow = 1;
		return;
	}
	uint8_t *tmp = malloc(len + 4);
	memcpy(tmp, &ty, 4);
	memcpy(tmp + 4, buf, len);
	tr_put(t, TR_SEI, tmp, (uint32_t)len + 4);
	free(tmp);
}

/* the parameter block a slice-parallel parser needs, read off the reference ctx */
static void fill_params(struct ref_tracer *t, struct h264_ctx *ctx,
			struct h264gpu_slice_params *p)
{
	const struct h264_slice_header *sh = &ctx->slice.hdr;
	memset(p, 0, sizeof(*p));
	p->nal_off = t->cur_nal_off;
	p->nal_len = (uint32_t)t->cur_nal_len;
	p->data_bit_off = (uint32_t)ctx->slice.hdr_len;
	p->first_mb_in_slice = sh->first_mb_in_slice;
	p->mb_out_off = (uint32_t)t->nal_mb0;
	p->mb_out_cap = (uint32_t)(t->mb_n - t->nal_mb0);
	p->pic_width_in_mbs = (uint16_t)ctx->sps_derived.PicWidthInMbs;
	p->pic_height_in_mbs = (uint16_t)ctx->derived.PicHeightInMbs;
	p->slice_type = (uint8_t)ctx->slice.type;
	p->chroma_array_type = (uint8_t)ctx->sps_derived.ChromaArrayType;
	p->bit_depth_luma = (uint8_t)ctx->sps_derived.BitDepthLuma;
	p->bit_depth_chroma = (uint8_t)ctx->sps_derived.BitDepthChroma;
	p->transform_8x8_mode_flag = (uint8_t)ctx->pps->transform_8x8_mode_flag;
	p->direct_8x8_inference_flag = (uint8_t)ctx->sps->direct_8x8_inference_flag;
	p->num_ref_idx_l0_active_minus1 = (uint8_t)sh->num_ref_idx_l0_active_minus1;
	p->num_ref_idx_l1_active_minus1 = (uint8_t)sh->num_ref_idx_l1_active_minus1;
	p->field_pic_flag = (uint8_t)sh->field_pic_flag;
	p->mbaff_frame_flag = (uint8_t)ctx->derived.MbaffFrameFlag;
	p->entropy_coding_mode_flag = (uint8_t)ctx->pps->entropy_coding_mode_flag;
	p->num_slice_groups_minus1 = (uint8_t)ctx->pps->num_slice_groups_minus1;
	p->cabac_init_idc = (uint8_t)sh->cabac_init_idc;
	p->slice_qp = (int8_t)ctx->derived.SliceQPLuma;
}

static void cb_slice(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
		     const struct h264_slice_header *sh, void *ud)
{
	struct ref_tracer *t = ud;
	struct h264gpu_slice_params p;
	tr_put(t, TR_SLICE, sh, sizeof(*sh));
	/* emitted after the slice's macroblocks: mb_out_off/mb_out_cap delimit its records */
	fill_params(t, ctx, &p);
	tr_put(t, TR_SLICE_PARAMS, &p, sizeof(p));
	if (ctx->pps->num_slice_groups_minus1 > 0 && ctx->slice.group_map != NULL) {
		/* h264_mb_to_slice_group (static in src/h264_fmo.c:217-235) over the whole picture */
		const uint32_t n = ctx->derived.PicSizeInMbs, W = ctx->sps_derived.PicWidthInMbs;
		uint8_t *m = malloc(n ? n : 1);
		for (uint32_t i = 0; i < n; i++) {
			uint32_t u = i;
			if (!(ctx->sps->frame_mbs_only_flag || sh->field_pic_flag))
				u = ctx->derived.MbaffFrameFlag ? i / 2 : (i / (2 * W)) * W + (i % W);
			m[i] = (uint8_t)ctx->slice.group_map[u];
		}
		tr_put(t, TR_GROUP_MAP, m, n);
		free(m);
	}
}

static void cb_sd_begin(struct h264_ctx *ctx, const struct h264_slice_header *sh, void *ud)
{
	tr_put(ud, TR_SLICE_DATA_BEGIN, NULL, 0);
}

static void cb_sd_end(struct h264_ctx *ctx, const struct h264_slice_header *sh,
		      uint32_t mb_count, void *ud)
{
	tr_put(ud, TR_SLICE_DATA_END, &mb_count, 4);
}

static void cb_sd_mb(struct h264_ctx *ctx, const struct h264_slice_header *sh,
		     uint32_t mb_addr, enum h264_mb_type mb_type, void *ud)
{
	struct ref_tracer *t = ud;
	if (t->mb_n < t->mb_cap) {
		t->mbs[t->mb_n].mb_addr = mb_addr;
		t->mbs[t->mb_n].mb_type = mb_type;
		t->mbs[t->mb_n].hash = t->mbs ? mb_hash(ctx->mb) : 0;
		if (t->syn)
			mb_syntax_from_ref(ctx->mb, mb_addr, mb_type, &t->syn[t->mb_n]);
	}
	t->mb_n++;
}

/*
 * h264_reader_parse(flags) over buf with recording callbacks.
 * Returns 0, or 1 when a buffer was too small (counts are still exact).
 * hash_mbs = 0 skips the (slow) per-MB checksum: used by the CPU baseline timing.
 */
static struct h264_mb_syntax *g_syn_out; /* set by ref_trace_parse_syntax for one call */

int ref_trace_parse(const uint8_t *buf, size_t len, uint32_t flags, uint8_t *log,
		    size_t log_cap, size_t *log_len, struct h264gpu_mb_record *mbs,
		    size_t mb_cap, size_t *mb_n, size_t *final_off);

/* same, and syn[mb_cap] receives every macroblock's syntax elements (single-threaded use) */
int ref_trace_parse_syntax(const uint8_t *buf, size_t len, uint32_t flags, uint8_t *log,
			   size_t log_cap, size_t *log_len, struct h264gpu_mb_record *mbs,
			   size_t mb_cap, size_t *mb_n, size_t *final_off, struct h264_mb_syntax *syn)
{
	g_syn_out = syn;
	int r = ref_trace_parse(buf, len, flags, log, log_cap, log_len, mbs, mb_cap, mb_n, final_off);
	g_syn_out = NULL;
	return r;
}

uint32_t ref_sizeof_mb_syntax(void)
{
	return (uint32_t)sizeof(struct h264_mb_syntax);
}

int ref_trace_parse(const uint8_t *buf, size_t len, uint32_t flags, uint8_t *log,
		    size_t log_cap, size_t *log_len, struct h264gpu_mb_record *mbs,
		    size_t mb_cap, size_t *mb_n, size_t *final_off)
{
	struct h264_ctx_cbs cbs;
	struct h264_reader *reader = NULL;
	struct ref_tracer t;
	size_t off = 0;
	memset(&cbs, 0, sizeof(cbs));
	memset(&t, 0, sizeof(t));
	t.base = buf;
	t.log = log;
	t.log_cap = log_cap;
	t.mbs = mbs;
	t.mb_cap = mb_cap;
	t.syn = g_syn_out;
	cbs.nalu_begin = cb_nalu_begin;
	cbs.nalu_end = cb_nalu_end;
	cbs.au_end = cb_au_end;
	cbs.sps = cb_sps;
	cbs.pps = cb_pps;
	cbs.aud = cb_aud;
	cbs.sei = cb_sei;
	cbs.slice = cb_slice;
	cbs.slice_data_begin = cb_sd_begin;
	cbs.slice_data_end = cb_sd_end;
	cbs.slice_data_mb = cb_sd_mb;
	if (h264_reader_new(&cbs, &t, &reader) < 0)
		return -1;
	h264_reader_parse(reader, flags, buf, len, &off);
	h264_reader_destroy(reader);
	if (log_len)
		*log_len = t.log_len;
	if (mb_n)
		*mb_n = t.mb_n;
	if (final_off)
		*final_off = off;
	return t.overflow || t.mb_n > mb_cap;
}

/* ---- CPU baseline for the macroblock parse: counting callbacks only ---------- */

struct count_ud {
	uint64_t mbs, slices, nals;
};
static void cnt_mb(struct h264_ctx *ctx, const struct h264_slice_header *sh, uint32_t a,
		   enum h264_mb_type ty, void *ud)
{
	((struct count_ud *)ud)->mbs++;
}
static void cnt_slice(struct h264_ctx *ctx, const uint8_t *buf, size_t len,
		      const struct h264_slice_header *sh, void *ud)
{
	((struct count_ud *)ud)->slices++;
}

struct parse_job {
	const uint8_t *buf;
	size_t len;
	uint32_t flags;
	struct count_ud c;
};

static void *parse_worker(void *arg)
{
	struct parse_job *j = arg;
	struct h264_ctx_cbs cbs;
	struct h264_reader *reader = NULL;
	size_t off = 0;
	memset(&cbs, 0, sizeof(cbs));
	cbs.slice_data_mb = cnt_mb;
	cbs.slice = cnt_slice;
	if (h264_reader_new(&cbs, &j->c, &reader) < 0)
		return NULL;
	h264_reader_parse(reader, j->flags, j->buf, j->len, &off);
	h264_reader_destroy(reader);
	return NULL;
}

/*
 * h264_reader_parse(flags) of nstreams independent streams (stream i =
 * bufs[i], lens[i]) on up to nthreads threads, one reader per stream, streams
 * dealt round-robin.  Returns seconds; *mbs = slice_data_mb callbacks seen.
 */
struct parse_thread {
	struct parse_job *jobs;
	size_t n, first, stride;
};
static void *parse_thread_main(void *arg)
{
	struct parse_thread *p = arg;
	for (size_t i = p->first; i < p->n; i += p->stride)
		parse_worker(&p->jobs[i]);
	return NULL;
}

double ref_mt_parse(const uint8_t *const *bufs, const size_t *lens, size_t nstreams,
		    uint32_t flags, int nthreads, uint64_t *mbs, uint64_t *slices)
{
	pthread_t th[256];
	struct parse_thread pt[256];
	struct parse_job *jobs = calloc(nstreams, sizeof(*jobs));
	double t0, t1;
	if (nthreads < 1)
		nthreads = 1;
	if (nthreads > 256)
		nthreads = 256;
	if ((size_t)nthreads > nstreams)
		nthreads = (int)nstreams;
	for (size_t i = 0; i < nstreams; i++) {
		jobs[i].buf = bufs[i];
		jobs[i].len = lens[i];
		jobs[i].flags = flags;
	}
	for (int i = 0; i < nthreads; i++) {
		pt[i].jobs = jobs;
		pt[i].n = nstreams;
		pt[i].first = (size_t)i;
		pt[i].stride = (size_t)nthreads;
	}
	t0 = now_s();
	for (int i = 1; i < nthreads; i++)
		pthread_create(&th[i], NULL, parse_thread_main, &pt[i]);
	parse_thread_main(&pt[0]);
	for (int i = 1; i < nthreads; i++)
		pthread_join(th[i], NULL);
	t1 = now_s();
	if (mbs) {
		*mbs = 0;
		for (size_t i = 0; i < nstreams; i++)
			*mbs += jobs[i].c.mbs;
	}
	if (slices) {
		*slices = 0;
		for (size_t i = 0; i < nstreams; i++)
			*slices += jobs[i].c.slices;
	}
	free(jobs);
	return t1 - t0;
}

/* ---- the reference's CABAC encoder engine driven by a bin script ------------------------
 * ops[i] = kind << 24 | ctxIdx << 8 | bin (kind 0 decision, 1 bypass, 2 terminate); contexts
 * initialised by the reference's h264_bac_state_init from ITS tables for (table, slice_qp):
 * table 0 = I slices, 1..3 = cabac_init_idc 0..2.  Returns bytes written (zero padded to a
 * byte).  Pins the generator's encoder and, through it, the GPU decoder's arithmetic. */
size_t ref_bac_ops(const uint32_t *ops, size_t n, uint32_t table, int32_t slice_qp, uint8_t *out,
		   size_t cap)
{
	struct h264_bitstream bs;
	struct h264_ctx *ctx = NULL;
	static struct h264_cabac cab;
	if (h264_ctx_new(&ctx) < 0)
		return 0;
	ctx->slice.type = table == 0 ? H264_SLICE_TYPE_I : H264_SLICE_TYPE_P;
	ctx->slice.hdr.cabac_init_idc = table == 0 ? 0 : table - 1;
	ctx->derived.SliceQPLuma = slice_qp;
	memset(out, 0, cap);
	h264_bs_init(&bs, out, cap, 0);
	h264_cabac_init_states(&cab, ctx); /* the reference's own tables + h264_bac_state_init */
	h264_bac_encode_init(&cab.enc, &bs, 1);
	for (size_t i = 0; i < n; i++) {
		uint32_t kind = ops[i] >> 24, c = (ops[i] >> 8) & 0x3ff, b = ops[i] & 1;
		if (kind == 0)
			h264_bac_encode_bin(&cab.enc, &cab.states[c], (int)b);
		else if (kind == 1)
			h264_bac_encode_bypass(&cab.enc, (int)b);
		else
			h264_bac_encode_terminate(&cab.enc, (int)b);
	}
	while (bs.cachebits % 8 != 0)
		h264_bs_write_bits(&bs, 0, 1);
	h264_ctx_destroy(ctx);
	return bs.off;
}
