/*
 * h264_reader.c — the libh264 reader API on top of the GPU stages.
 *
 * Same entry points and callback order as the reference (src/h264_reader.c:56-255,
 * src/h264_syntax.h:1446-1604), re-ordered for bulk work:
 *
 *   h264_reader_parse(buf)
 *     1. the buffer goes to the GPU ONCE (pooled device memory) and ONE scan-only launch
 *        builds the NAL table (h264gpu_reader_scan) instead of the h264_find_nalu loop
 *        (src/h264_reader.c:133-140);
 *     2. with H264_READER_FLAGS_SLICE_DATA: a silent header pass over the NAL units (a
 *        shadow context, no callbacks) collects one parameter block per CAVLC slice,
 *        and ONE launch parses the macroblock layer of all of them out of the resident
 *        copy (h264gpu_reader_parse_cavlc) instead of _h264_read_slice_data_internal per
 *        slice (src/h264_syntax_slice_data.h:701-787); the records arrive in pooled
 *        pinned memory.  A slice's record budget ends where the next slice of its picture
 *        begins, so the records of a buffer add up to its macroblocks;
 *     3. the replay pass parses the headers again on the caller's context and fires the
 *        callbacks in the reference's order, slice_data_mb from the GPU records.
 *
 * There is no CPU path for the bulk stages: without a GPU h264_reader_parse and the
 * slice-data part of h264_reader_parse_nalu fail with -ENODEV.
 */
#include <time.h>
#include "h264_priv.h"

int h264_reader_new(const struct h264_ctx_cbs *cbs, void *userdata, struct h264_reader **ret_obj)
{
	if (ret_obj == NULL)
		return -EINVAL;
	*ret_obj = NULL;
	if (cbs == NULL)
		return -EINVAL;
	struct h264_reader *r = calloc(1, sizeof(*r));
	if (r == NULL)
		return -ENOMEM;
	r->cbs = *cbs; /* copied, like the reference (src/h264_reader.c:75) */
	r->userdata = userdata;
	int res = h264_ctx_new(&r->ctx);
	if (res < 0) {
		free(r);
		return res;
	}
	*ret_obj = r;
	return 0;
}

int h264_reader_destroy(struct h264_reader *reader)
{
	if (reader == NULL)
		return 0;
	h264_ctx_destroy(reader->ctx);
	if (reader->gpu != NULL)
		h264gpu_destroy(reader->gpu);
	if (reader->gpu_single != NULL)
		h264gpu_destroy(reader->gpu_single);
	free(reader);
	return 0;
}

struct h264_ctx *h264_reader_get_ctx(struct h264_reader *reader)
{
	return reader == NULL ? NULL : reader->ctx;
}

int h264_reader_stop(struct h264_reader *reader)
{
	if (reader == NULL)
		return -EINVAL;
	reader->stop = 1;
	return 0;
}

static int reader_gpu(struct h264_reader *reader)
{
	if (reader->gpu != NULL)
		return 0;
	int dev = 0;
	const char *e = getenv("H264_GPU_DEVICE");
	if (e != NULL)
		dev = atoi(e);
	return h264gpu_create(dev, &reader->gpu);
}

/* ---- slice data --------------------------------------------------------------------------- */

/* the parameter block the slice kernel needs, from the context
 * right after h264_ctx_set_slice_header */
int h264_fill_slice_params(const struct h264_ctx *ctx, struct h264gpu_slice_params *p)
{
	const struct h264_sps *sps = ctx->sps;
	const struct h264_pps *pps = ctx->pps;
	if (sps == NULL || pps == NULL)
		return -EINVAL;
	memset(p, 0, sizeof(*p));
	p->data_bit_off = (uint32_t)ctx->sh_bits;
	p->first_mb_in_slice = ctx->sh.first_mb_in_slice;
	p->pic_width_in_mbs = (uint16_t)ctx->spsd.PicWidthInMbs;
	p->pic_height_in_mbs = (uint16_t)ctx->PicHeightInMbs;
	p->slice_type = (uint8_t)ctx->slice_type;
	p->chroma_array_type = (uint8_t)ctx->spsd.ChromaArrayType;
	p->bit_depth_luma = (uint8_t)ctx->spsd.BitDepthLuma;
	p->bit_depth_chroma = (uint8_t)ctx->spsd.BitDepthChroma;
	p->transform_8x8_mode_flag = (uint8_t)pps->transform_8x8_mode_flag;
	p->direct_8x8_inference_flag = (uint8_t)sps->direct_8x8_inference_flag;
	p->num_ref_idx_l0_active_minus1 = (uint8_t)ctx->sh.num_ref_idx_l0_active_minus1;
	p->num_ref_idx_l1_active_minus1 = (uint8_t)ctx->sh.num_ref_idx_l1_active_minus1;
	p->field_pic_flag = (uint8_t)ctx->sh.field_pic_flag;
	p->mbaff_frame_flag = (uint8_t)ctx->MbaffFrameFlag;
	p->entropy_coding_mode_flag = (uint8_t)pps->entropy_coding_mode_flag;
	p->num_slice_groups_minus1 = (uint8_t)pps->num_slice_groups_minus1;
	p->cabac_init_idc = (uint8_t)ctx->sh.cabac_init_idc;
	p->slice_qp = (int8_t)ctx->SliceQPLuma;
	const uint32_t first = ctx->sh.first_mb_in_slice * (1 + (ctx->MbaffFrameFlag ? 1 : 0));
	p->mb_out_cap = ctx->PicSizeInMbs > first ? ctx->PicSizeInMbs - first : 0;
	return 0;
}

/* called by the syntax walk right after the slice header of a slice NAL when the
 * caller asked for slice data: deliver slice_data_begin / _mb / _end */
int h264_reader_slice_data(struct h264_reader *reader, struct h264_ctx *ctx, const uint8_t *nal,
			   size_t nal_len)
{
	/* CABAC slice data is not parsed, and nothing is delivered for it: exactly the
	 * reference's behaviour (src/h264_syntax_slice_data.h:715-717), unless the caller opted in
	 * to this library's extension (H264_READER_FLAGS_SLICE_DATA_CABAC) */
	const int cabac = ctx->pps->entropy_coding_mode_flag != 0;
	if (cabac && !(reader->flags & H264_READER_FLAGS_SLICE_DATA_CABAC))
		return 0;
	const struct h264_ctx_cbs *cbs = &reader->cbs;
	const struct h264gpu_mb_record *rec = NULL;
	struct h264gpu_mb_record *own = NULL;
	struct h264gpu_slice_result result;
	memset(&result, 0, sizeof(result));

	if (cbs->slice_data_begin != NULL)
		cbs->slice_data_begin(ctx, &ctx->sh, reader->userdata);

	int bulk = 0;
	if (reader->records != NULL && reader->next_slice < reader->n_slices &&
	    reader->bulk_base + reader->params[reader->next_slice].nal_off == nal) {
		/* bulk mode: this slice was parsed with all the others */
		const struct h264gpu_slice_params *p = &reader->params[reader->next_slice];
		rec = reader->records + p->mb_out_off;
		result = reader->results[reader->next_slice];
		reader->next_slice++;
		/* a slice that runs past the start of the next one (only a broken stream does) ran out
		 * of its record budget: parse it again on its own with the whole picture as budget */
		bulk = result.status != -ENOBUFS;
	}
	if (!bulk) {
		/* single NAL unit: one-slice launch (own context: the bulk records stay untouched) */
		struct h264gpu_slice_params p;
		if (reader->gpu_single == NULL) {
			int dev = 0;
			const char *e = getenv("H264_GPU_DEVICE");
			if (e != NULL)
				dev = atoi(e);
			int res = h264gpu_create(dev, &reader->gpu_single);
			if (res < 0)
				return res;
		}
		int res = h264_fill_slice_params(ctx, &p);
		if (res < 0)
			return res;
		p.nal_off = 0;
		p.nal_len = (uint32_t)nal_len;
		p.mb_out_off = 0;
		own = malloc(((size_t)p.mb_out_cap + 1) * sizeof(*own));
		if (own == NULL)
			return -ENOMEM;
		if (p.num_slice_groups_minus1 != 0 && !cabac) {
			/* several slice groups: the slice's macroblock -> slice group map goes with it */
			uint8_t *map = malloc(ctx->PicSizeInMbs ? ctx->PicSizeInMbs : 1);
			res = map == NULL ? -ENOMEM : h264_ctx_get_slice_group_map(ctx, map, ctx->PicSizeInMbs);
			if (res >= 0)
				res = h264gpu_reader_set_group_maps(reader->gpu_single, map, ctx->PicSizeInMbs);
			free(map);
			if (res < 0) {
				free(own);
				return res;
			}
			p.row_state_off = 0;
		}
		res = cabac ? h264gpu_cabac_parse_host(reader->gpu_single, nal, nal_len, &p, 1, own, p.mb_out_cap,
						       &result)
			    : h264gpu_cavlc_parse_host(reader->gpu_single, nal, nal_len, &p, 1, own, p.mb_out_cap,
						       &result);
		if (res < 0) {
			free(own);
			return res;
		}
		rec = own;
	}
	if (cbs->slice_data_mb != NULL)
		for (uint32_t i = 0; i < result.mb_count; i++)
			cbs->slice_data_mb(ctx, &ctx->sh, rec[i].mb_addr, (enum h264_mb_type)rec[i].mb_type,
					   reader->userdata);
	free(own);
	if (result.status < 0)
		return result.status; /* the walk stops here, like the reference's on a bad MB */
	if (cbs->slice_data_end != NULL)
		cbs->slice_data_end(ctx, &ctx->sh, result.mb_count, reader->userdata);
	return 0;
}

/* ---- one NAL unit ------------------------------------------------------------------------- */

static int parse_one(struct h264_reader *reader, struct h264_ctx *ctx, const struct h264_ctx_cbs *cbs,
		     const uint8_t *buf, size_t len)
{
	struct h264_bitstream bs;
	h264_bs_cinit(&bs, buf, len, 1);
	bs.priv = reader;
	struct h264_io io = {H264_IO_READ, &bs, ctx, cbs, reader->userdata, cbs != NULL ? reader : NULL};
	return h264_syntax_nalu(&io);
}

int h264_reader_parse_nalu(struct h264_reader *reader, uint32_t flags, const uint8_t *buf, size_t len)
{
	if (reader == NULL || buf == NULL)
		return -EINVAL;
	reader->stop = 0;
	reader->flags = flags;
	return parse_one(reader, reader->ctx, &reader->cbs, buf, len);
}

/* ---- whole Annex-B buffer ------------------------------------------------------------------ */

struct slice_list {
	struct h264gpu_slice_params *params;
	uint32_t n, cap;
	uint64_t records;
	int any_cabac;
	/* pictures with several slice groups: macroblock -> slice group maps, back to back */
	uint8_t *maps;
	size_t maps_len, maps_cap;
};

/* the map of the slice whose header the context holds: appended to the list (or the previous
 * slice's map again if nothing changed) */
static int add_group_map(struct slice_list *out, const struct h264_ctx *ctx, struct h264gpu_slice_params *p)
{
	const size_t n = ctx->PicSizeInMbs;
	if (out->maps_len + n > out->maps_cap) {
		const size_t cap = (out->maps_cap ? out->maps_cap * 2 : 65536) + n;
		void *m = realloc(out->maps, cap);
		if (m == NULL)
			return -ENOMEM;
		out->maps = m;
		out->maps_cap = cap;
	}
	uint8_t *map = out->maps + out->maps_len;
	const int r = h264_ctx_get_slice_group_map(ctx, map, n);
	if (r < 0)
		return r;
	if (out->maps_len >= n && memcmp(map - n, map, n) == 0) {
		p->row_state_off = (uint32_t)(out->maps_len - n);
		return 0;
	}
	if (out->maps_len + n > UINT32_MAX)
		return -E2BIG;
	p->row_state_off = (uint32_t)out->maps_len;
	out->maps_len += n;
	return 0;
}

/* header-only pass on a private context: which NAL units are CAVLC slices, and with
 * which parameters */
static int collect_slices(struct h264_reader *reader, const uint8_t *buf, const uint64_t *st,
			  const uint64_t *en, uint64_t n_nal, struct slice_list *out)
{
	struct h264_ctx *shadow = NULL;
	int res = h264_ctx_new(&shadow);
	if (res < 0)
		return res;
	/* the caller's parameter sets are already known to later slices of this buffer */
	for (uint32_t i = 0; i < H264_SPS_MAX; i++)
		if (reader->ctx->sps_tab[i] != NULL)
			h264_ctx_set_sps(shadow, reader->ctx->sps_tab[i]);
	for (uint32_t i = 0; i < H264_PPS_MAX; i++)
		if (reader->ctx->pps_tab[i] != NULL)
			h264_ctx_set_pps(shadow, reader->ctx->pps_tab[i]);
	for (uint64_t k = 0; k < n_nal && res >= 0; k++) {
		const uint8_t *nal = buf + st[k];
		const size_t len = (size_t)(en[k] - st[k]);
		if (len == 0)
			continue;
		const unsigned type = nal[0] & 0x1f;
		if (type != H264_NALU_TYPE_SLICE && type != H264_NALU_TYPE_SLICE_IDR &&
		    type != H264_NALU_TYPE_SPS && type != H264_NALU_TYPE_PPS)
			continue; /* nothing else changes what a slice header means */
		if (parse_one(reader, shadow, NULL, nal, len) < 0)
			continue; /* the replay pass reports it */
		if (type == H264_NALU_TYPE_SPS || type == H264_NALU_TYPE_PPS ||
		    (shadow->pps->entropy_coding_mode_flag && !(reader->flags & H264_READER_FLAGS_SLICE_DATA_CABAC)))
			continue;
		if (shadow->pps->entropy_coding_mode_flag)
			out->any_cabac = 1;
		if (out->n == out->cap) {
			const uint32_t cap = out->cap ? out->cap * 2 : 256;
			void *p = realloc(out->params, (size_t)cap * sizeof(*out->params));
			if (p == NULL) {
				res = -ENOMEM;
				break;
			}
			out->params = p;
			out->cap = cap;
		}
		struct h264gpu_slice_params *p = &out->params[out->n];
		if (h264_fill_slice_params(shadow, p) < 0)
			continue;
		p->nal_off = st[k];
		p->nal_len = (uint32_t)len;
		if (p->num_slice_groups_minus1 != 0 && add_group_map(out, shadow, p) < 0)
			continue; /* parsed on its own by the replay pass */
		/* the slice before this one, if it belongs to the same picture and starts earlier,
		 * ends where this one begins: give its unused record budget back (one slice group only:
		 * slices of different groups interleave) */
		if (out->n > 0 && !shadow->first_vcl && !shadow->MbaffFrameFlag && p->num_slice_groups_minus1 == 0 &&
		    out->params[out->n - 1].num_slice_groups_minus1 == 0) {
			struct h264gpu_slice_params *q = &out->params[out->n - 1];
			if (q->first_mb_in_slice < p->first_mb_in_slice && q->mb_out_off + (uint64_t)q->mb_out_cap == out->records) {
				const uint32_t cap = p->first_mb_in_slice - q->first_mb_in_slice;
				if (cap < q->mb_out_cap) {
					out->records -= q->mb_out_cap - cap;
					q->mb_out_cap = cap;
				}
			}
		}
		if (out->records + p->mb_out_cap > UINT32_MAX)
			break; /* record index space exhausted: the rest parses one by one */
		p->mb_out_off = (uint32_t)out->records;
		out->records += p->mb_out_cap;
		out->n++;
	}
	h264_ctx_destroy(shadow);
	return res;
}

/* H264_READER_TIMING=1: phase times of h264_reader_parse on stderr (diagnostics) */
static double now_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

int h264_reader_parse(struct h264_reader *reader, uint32_t flags, const uint8_t *buf, size_t len,
		      size_t *off)
{
	const int timing = getenv("H264_READER_TIMING") != NULL;
	double t0 = timing ? now_ms() : 0, t1 = 0, t2 = 0, t3 = 0;
	if (reader == NULL || buf == NULL || off == NULL)
		return -EINVAL;
	reader->stop = 0;
	*off = 0;
	if (len == 0)
		return 0;
	int res = reader_gpu(reader);
	if (res < 0)
		return res;

	/* 1. one upload, NAL table of the whole buffer (pooled pinned memory of the GPU context) */
	uint64_t n_nal = 0, final_off = 0;
	const uint64_t *st = NULL, *en = NULL;
	res = h264gpu_reader_scan(reader->gpu, buf, len, &st, &en, &n_nal, &final_off);
	if (res < 0)
		return res;
	if (timing)
		t1 = now_ms();

	/* 2. macroblock layer of every CAVLC slice, one launch on the resident copy */
	struct slice_list sl;
	memset(&sl, 0, sizeof(sl));
	reader->flags = flags;
	if (flags & H264_READER_FLAGS_SLICE_DATA) {
		const struct h264gpu_mb_record *records = NULL;
		const struct h264gpu_slice_result *results = NULL;
		res = collect_slices(reader, buf, st, en, n_nal, &sl);
		if (timing)
			t2 = now_ms();
		if (res >= 0 && sl.n > 0 && sl.maps_len > 0)
			res = h264gpu_reader_set_group_maps(reader->gpu, sl.maps, sl.maps_len);
		if (res >= 0 && sl.n > 0)
			res = sl.any_cabac
				? h264gpu_reader_parse_slices(reader->gpu, sl.params, sl.n, sl.records, &records, &results)
				: h264gpu_reader_parse_cavlc(reader->gpu, sl.params, sl.n, sl.records, &records, &results);
		if (res < 0)
			goto out_slices;
		if (sl.n > 0) {
			reader->records = records;
			reader->results = results;
			reader->params = sl.params;
			reader->n_slices = sl.n;
			reader->next_slice = 0;
			reader->bulk_base = buf;
		}
	}

	if (timing)
		t3 = now_ms();
	/* 3. callbacks, in stream order (per-NAL errors are ignored like the reference
	 * ignores them, src/h264_reader.c:137) */
	reader->flags = flags;
	for (uint64_t k = 0; k < n_nal && !reader->stop; k++) {
		/* skip the records of slices whose NAL the replay does not reach as a slice */
		while (reader->records != NULL && reader->next_slice < reader->n_slices &&
		       reader->params[reader->next_slice].nal_off < st[k])
			reader->next_slice++;
		reader->stop = 0;
		parse_one(reader, reader->ctx, &reader->cbs, buf + st[k], (size_t)(en[k] - st[k]));
		*off = (size_t)en[k];
	}
	if (!reader->stop)
		*off = (size_t)final_off;
	if (timing)
		fprintf(stderr, "h264_reader_parse: %zu bytes, %llu NALs, %u slices: upload + scan %.2f ms, header pass %.2f, "
			"slice kernels + records back %.2f, callbacks %.2f\n", len, (unsigned long long)n_nal, sl.n, t1 - t0,
			t2 > 0 ? t2 - t1 : 0.0, t2 > 0 ? t3 - t2 : 0.0, now_ms() - t3);
	res = 0;
	reader->records = NULL;
	reader->results = NULL;
	reader->params = NULL;
	reader->n_slices = 0;
out_slices:
	free(sl.params);
	free(sl.maps);
	return res;
}

/* ---- standalone header parsers (src/h264_reader.c:165-255) -------------------------------- */

int h264_parse_nalu_header(const uint8_t *buf, size_t len, struct h264_nalu_header *nh)
{
	if (buf == NULL || nh == NULL)
		return -EINVAL;
	struct h264_bitstream bs;
	struct h264_io io = {H264_IO_READ, &bs, NULL, NULL, NULL, NULL};
	memset(nh, 0, sizeof(*nh));
	h264_bs_cinit(&bs, buf, len, 1);
	return h264_syntax_nalu_header(&io, nh);
}

int h264_parse_sps(const uint8_t *buf, size_t len, struct h264_sps *sps)
{
	if (buf == NULL || sps == NULL)
		return -EINVAL;
	struct h264_bitstream bs;
	struct h264_nalu_header nh;
	struct h264_io io = {H264_IO_READ, &bs, NULL, NULL, NULL, NULL};
	memset(sps, 0, sizeof(*sps));
	memset(&nh, 0, sizeof(nh));
	sps->chroma_format_idc = 1;
	h264_bs_cinit(&bs, buf, len, 1);
	int res = h264_syntax_nalu_header(&io, &nh);
	if (res < 0)
		return res;
	if (nh.nal_unit_type != H264_NALU_TYPE_SPS)
		return -EIO;
	return h264_syntax_sps(&io, sps);
}

int h264_parse_pps(const uint8_t *buf, size_t len, const struct h264_sps *sps, struct h264_pps *pps)
{
	if (buf == NULL || sps == NULL || pps == NULL)
		return -EINVAL;
	struct h264_bitstream bs;
	struct h264_nalu_header nh;
	struct h264_io io = {H264_IO_READ, &bs, NULL, NULL, NULL, NULL};
	memset(pps, 0, sizeof(*pps));
	memset(&nh, 0, sizeof(nh));
	h264_bs_cinit(&bs, buf, len, 1);
	int res = h264_syntax_nalu_header(&io, &nh);
	if (res < 0)
		return res;
	if (nh.nal_unit_type != H264_NALU_TYPE_PPS)
		return -EIO;
	uint32_t id = 0;
	if ((res = h264_bs_read_bits_ue(&bs, &id)) < 0)
		return res;
	pps->pic_parameter_set_id = id;
	if ((res = h264_bs_read_bits_ue(&bs, &id)) < 0)
		return res;
	pps->seq_parameter_set_id = id;
	if (sps->seq_parameter_set_id != pps->seq_parameter_set_id)
		return -EIO;
	return h264_syntax_pps_body(&io, sps, pps);
}
