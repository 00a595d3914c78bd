/*
 * h264gpu_slice_api.cu — C-ABI entry points of the slice-parallel macroblock
 * parse (include/h264gpu_slice.h).  Launch only; no CPU implementation exists.
 */
#include "h264gpu_internal.h"

#include "cavlc_steps.cuh"
#include "cabac_parse.cuh"
#include "conceal.cuh"


/*
 * K4, second generation (cavlc_steps.cuh): lanes pull slices from a counter; one ring of
 * bottom-row counts per LANE ((PicWidthInMbs + 1) x 16 bytes, pictures up to 8192 luma samples
 * wide; wider slices get -E2BIG).
 *   H264GPU_CAVLC_LANES_LOG2  log2 of the working lanes per warp (default: by slice count)
 *   H264GPU_CAVLC_PER_LANE    slices per lane the grid is sized for (default 1)
 */
static int cavlc_steps_launch(h264gpu_ctx *ctx, const uint8_t *d_stream, uint64_t stream_len,
			      const struct h264gpu_slice_params *d_params, uint32_t n_slices,
			      struct h264gpu_mb_record *d_records, struct h264gpu_slice_result *d_results,
			      struct h264_mb_syntax *d_syntax, const uint8_t *d_group_maps, cudaStream_t st, int sms)
{
	uint32_t lanes_log2 = 0;
	const char *env = getenv("H264GPU_CAVLC_LANES_LOG2");
	if (env != NULL && atoi(env) >= 0 && atoi(env) <= 5) {
		lanes_log2 = (uint32_t)atoi(env);
	} else {
		/* one slice per warp while the warps fit the machine (~32 per SM), then more lanes per
		 * warp (measured at 16000 slices on 148 SMs: 4 lanes 95 M, 8 lanes 93 M, 16 lanes 71 M,
		 * 32 lanes 57 M macroblocks/s: the warps to switch between are worth more than the shared
		 * instruction issue) */
		while (lanes_log2 < 5 && ((uint64_t)n_slices >> lanes_log2) > (uint64_t)sms * 32)
			lanes_log2++;
	}
	uint32_t per_lane = 1;
	env = getenv("H264GPU_CAVLC_PER_LANE");
	if (env != NULL && atoi(env) >= 1 && atoi(env) <= 64)
		per_lane = (uint32_t)atoi(env);
	const uint64_t lanes = ((uint64_t)n_slices + per_lane - 1) / per_lane;
	const uint32_t threads = CAVLC2_STRIDE;
	const uint64_t warps = (lanes + (1u << lanes_log2) - 1) >> lanes_log2;
	const uint32_t blocks = (uint32_t)((warps * 32 + threads - 1) / threads);
	const uint64_t grid_lanes = ((uint64_t)blocks * threads / 32) << lanes_log2;
	/* rings sized for pictures up to 8192 luma samples wide unless the caller's parameter blocks
	 * were seen on the host (the host forms): then for the widest picture among them */
	const uint32_t ring_w = ctx->ring_w_hint ? ctx->ring_w_hint : 512;
	const uint64_t ring_stride = (uint64_t)(ring_w + 1) * CAVLC2_RING_SLOT;
	const size_t counter_off = (size_t)(grid_lanes * ring_stride);
	const size_t order_off = counter_off + 16;
	int r = h264gpu_ws_reserve(ctx, order_off + (size_t)n_slices * 4);
	if (r < 0)
		return r;
	cavlc2::CavlcArgs a;
	a.stream = d_stream;
	a.stream_len = stream_len;
	a.params = d_params;
	a.n_slices = n_slices;
	a.records = d_records;
	a.results = d_results;
	a.ring = (uint8_t *)ctx->ws;
	a.ring_stride = ring_stride;
	a.ring_w = ring_w;
	a.lanes_log2 = lanes_log2;
	a.syntax = d_syntax;
	a.group_maps = d_group_maps;
	a.repeat = CAVLC2_RES_REPEAT;
	env = getenv("H264GPU_CAVLC_REPEAT");
	if (env != NULL && atoi(env) >= 1 && atoi(env) <= 64)
		a.repeat = (uint32_t)atoi(env);
	a.next_slice = (uint32_t *)((uint8_t *)ctx->ws + counter_off);
	a.order = (uint32_t *)((uint8_t *)ctx->ws + order_off);
	/* longest slices first; the first grid_lanes tickets are the lanes' own numbers */
	env = getenv("H264GPU_CAVLC_SORT");
	cavlc2::order_kernel<<<1, CAVLC2_ORDER_T, 0, st>>>(d_params, n_slices, a.order, a.next_slice, (uint32_t)grid_lanes,
						    env != NULL && atoi(env) == 0 ? 0u : 1u);
	ctx->launches++;
	const size_t smem = (size_t)CAVLC2_SM_WORDS * threads * 4;
	if (d_syntax == NULL) {
		cavlc2::cavlc_steps_kernel<false, false><<<blocks, threads, smem, st>>>(a);
		cavlc2::cavlc_steps_kernel<false, true><<<blocks, threads, smem, st>>>(a);
	} else {
		cavlc2::cavlc_steps_kernel<true, false><<<blocks, threads, smem, st>>>(a);
		cavlc2::cavlc_steps_kernel<true, true><<<blocks, threads, smem, st>>>(a);
	}
	CU_TRY(cudaGetLastError());
	ctx->launches += 2;
	return 0;
}

extern "C" int h264gpu_cavlc_parse_dev(h264gpu_ctx *ctx, const uint8_t *d_stream,
				       uint64_t stream_len,
				       const struct h264gpu_slice_params *d_params,
				       uint32_t n_slices, struct h264gpu_mb_record *d_records,
				       struct h264gpu_slice_result *d_results, void *stream)
{
	return h264gpu_cavlc_parse_fmo_dev(ctx, d_stream, stream_len, d_params, n_slices, d_records, d_results, NULL,
					   NULL, stream);
}

extern "C" int h264gpu_cavlc_parse_full_dev(h264gpu_ctx *ctx, const uint8_t *d_stream, uint64_t stream_len,
					    const struct h264gpu_slice_params *d_params, uint32_t n_slices,
					    struct h264gpu_mb_record *d_records, struct h264gpu_slice_result *d_results,
					    struct h264_mb_syntax *d_syntax, void *stream)
{
	return h264gpu_cavlc_parse_fmo_dev(ctx, d_stream, stream_len, d_params, n_slices, d_records, d_results,
					   d_syntax, NULL, stream);
}

/* A12: 8.2.2 on the host (fmo_map.h), for callers that build the parameter blocks themselves */
#include "fmo_map.h"
extern "C" int h264gpu_fmo_mb_map(const struct h264gpu_fmo_desc *desc, uint8_t *mb_map)
{
	if (desc == NULL || mb_map == NULL || desc->pic_width_in_mbs == 0 || desc->pic_height_in_map_units == 0)
		return -EINVAL;
	struct fmo_desc d;
	memset(&d, 0, sizeof(d));
	d.num_slice_groups_minus1 = desc->num_slice_groups_minus1;
	d.map_type = desc->slice_group_map_type;
	d.run_length_minus1 = desc->run_length_minus1;
	d.top_left = desc->top_left;
	d.bottom_right = desc->bottom_right;
	d.change_direction_flag = desc->slice_group_change_direction_flag;
	d.map_units_in_slice_group0 = desc->map_units_in_slice_group0;
	d.slice_group_id = desc->slice_group_id;
	d.n_slice_group_id = desc->n_slice_group_id;
	d.pic_width_in_mbs = desc->pic_width_in_mbs;
	d.pic_height_in_map_units = desc->pic_height_in_map_units;
	d.frame_mbs_only_flag = desc->frame_mbs_only_flag;
	d.field_pic_flag = desc->field_pic_flag;
	d.mbaff_frame_flag = desc->mbaff_frame_flag;
	d.pic_size_in_mbs = desc->pic_size_in_mbs;
	if ((d.map_type == 0 && d.run_length_minus1 == NULL) ||
	    (d.map_type == 2 && (d.top_left == NULL || d.bottom_right == NULL)) ||
	    (d.map_type == 6 && d.slice_group_id == NULL))
		return -EINVAL;
	const size_t units = (size_t)d.pic_width_in_mbs * d.pic_height_in_map_units;
	uint8_t *u = (uint8_t *)malloc(units);
	if (u == NULL)
		return -ENOMEM;
	const int r = fmo_map_units(&d, u);
	if (r == 0)
		fmo_mb_map(&d, u, mb_map);
	free(u);
	return r < 0 ? -EINVAL : 0;
}

extern "C" int h264gpu_reader_set_group_maps(h264gpu_ctx *ctx, const uint8_t *h_maps, uint64_t bytes)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	ctx->rd_maps_len = 0;
	if (bytes == 0)
		return 0;
	if (h_maps == NULL)
		return -EINVAL;
	cudaStream_t st;
	if ((r = h264gpu_reader_stream(ctx, &st)) < 0 || (r = h264gpu_pool_dev(ctx, &ctx->rd_maps, bytes)) < 0)
		return r;
	CU_TRY(cudaMemcpyAsync(ctx->rd_maps.p, h_maps, bytes, cudaMemcpyHostToDevice, st));
	CU_TRY(cudaStreamSynchronize(st)); /* h_maps is the caller's again */
	ctx->rd_maps_len = bytes;
	return 0;
}

extern "C" int h264gpu_cavlc_parse_fmo_dev(h264gpu_ctx *ctx, const uint8_t *d_stream, uint64_t stream_len,
					   const struct h264gpu_slice_params *d_params, uint32_t n_slices,
					   struct h264gpu_mb_record *d_records, struct h264gpu_slice_result *d_results,
					   struct h264_mb_syntax *d_syntax, const uint8_t *d_group_maps, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (n_slices == 0)
		return 0;
	if (d_stream == NULL || d_params == NULL || d_records == NULL || d_results == NULL)
		return -EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
	return cavlc_steps_launch(ctx, d_stream, stream_len, d_params, n_slices, d_records, d_results, d_syntax,
				  d_group_maps, st, sms);
}

typedef int (*parse_dev_fn)(h264gpu_ctx *, const uint8_t *, uint64_t, const struct h264gpu_slice_params *,
			    uint32_t, struct h264gpu_mb_record *, struct h264gpu_slice_result *, void *);

/* upload (unless the stream is the one resident from h264gpu_reader_scan), launch, download into
 * pooled pinned memory; *out_records / *out_results stay valid until the next reader / parse call */
static int parse_pooled(parse_dev_fn dev, h264gpu_ctx *ctx, const uint8_t *h_stream, uint64_t stream_len,
			const struct h264gpu_slice_params *h_params, uint32_t n_slices, uint64_t n_records,
			const struct h264gpu_mb_record **out_records,
			const struct h264gpu_slice_result **out_results, parse_dev_fn dev2 = NULL)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	cudaStream_t st;
	if ((r = h264gpu_reader_stream(ctx, &st)) < 0)
		return r;
	if (h_stream != NULL) {
		if ((r = h264gpu_reader_upload(ctx, h_stream, stream_len)) < 0)
			return r;
	} else if (ctx->rd_stream_len == 0) {
		return -EINVAL;
	}
	if ((r = h264gpu_pool_dev(ctx, &ctx->rd_params, (size_t)n_slices * sizeof(*h_params))) < 0 ||
	    (r = h264gpu_pool_dev(ctx, &ctx->rd_records, (n_records + 1) * sizeof(struct h264gpu_mb_record))) < 0 ||
	    (r = h264gpu_pool_dev(ctx, &ctx->rd_results, 2 * (size_t)n_slices * sizeof(struct h264gpu_slice_result))) < 0 ||
	    (r = h264gpu_pool_host(ctx, &ctx->rh_records, (n_records + 1) * sizeof(struct h264gpu_mb_record))) < 0 ||
	    (r = h264gpu_pool_host(ctx, &ctx->rh_results, 2 * (size_t)n_slices * sizeof(struct h264gpu_slice_result))) < 0)
		return r;
	CU_TRY(cudaMemcpyAsync(ctx->rd_params.p, h_params, (size_t)n_slices * sizeof(*h_params),
			       cudaMemcpyHostToDevice, st));
	uint32_t wmax = 1;
	for (uint32_t i = 0; i < n_slices; i++)
		if (h_params[i].pic_width_in_mbs > wmax)
			wmax = h_params[i].pic_width_in_mbs;
	ctx->ring_w_hint = wmax;
	r = dev(ctx, (const uint8_t *)ctx->rd_stream.p, ctx->rd_stream_len,
		(const struct h264gpu_slice_params *)ctx->rd_params.p, n_slices,
		(struct h264gpu_mb_record *)ctx->rd_records.p, (struct h264gpu_slice_result *)ctx->rd_results.p, st);
	if (r < 0)
		return r;
	if (dev2 != NULL) {
		/* the other entropy coder's slices: same records array (every kernel writes only the
		 * records of the slices it parses), second half of the results */
		r = dev2(ctx, (const uint8_t *)ctx->rd_stream.p, ctx->rd_stream_len,
			 (const struct h264gpu_slice_params *)ctx->rd_params.p, n_slices,
			 (struct h264gpu_mb_record *)ctx->rd_records.p,
			 (struct h264gpu_slice_result *)ctx->rd_results.p + n_slices, st);
		if (r < 0)
			return r;
	}
	ctx->ring_w_hint = 0;
	if (n_records)
		CU_TRY(cudaMemcpyAsync(ctx->rh_records.p, ctx->rd_records.p, n_records * sizeof(struct h264gpu_mb_record),
				       cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaMemcpyAsync(ctx->rh_results.p, ctx->rd_results.p,
			       (dev2 ? 2 : 1) * (size_t)n_slices * sizeof(struct h264gpu_slice_result),
			       cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaStreamSynchronize(st));
	struct h264gpu_slice_result *res = (struct h264gpu_slice_result *)ctx->rh_results.p;
	if (dev2 != NULL)
		for (uint32_t i = 0; i < n_slices; i++)
			if (res[i].status == H264GPU_SLICE_SKIPPED)
				res[i] = res[n_slices + i];
	*out_records = (const struct h264gpu_mb_record *)ctx->rh_records.p;
	*out_results = res;
	return 0;
}

static int parse_host(parse_dev_fn dev, h264gpu_ctx *ctx, const uint8_t *h_stream, uint64_t stream_len,
		      const struct h264gpu_slice_params *h_params, uint32_t n_slices,
		      struct h264gpu_mb_record *h_records, uint64_t n_records,
		      struct h264gpu_slice_result *h_results)
{
	if (n_slices == 0)
		return 0;
	if (h_stream == NULL || h_params == NULL || h_records == NULL || h_results == NULL)
		return -EINVAL;
	const struct h264gpu_mb_record *rec = NULL;
	const struct h264gpu_slice_result *res = NULL;
	const int r = parse_pooled(dev, ctx, h_stream, stream_len, h_params, n_slices, n_records, &rec, &res);
	if (r < 0)
		return r;
	memcpy(h_records, rec, n_records * sizeof(*rec));
	memcpy(h_results, res, (size_t)n_slices * sizeof(*res));
	return 0;
}

/* the CAVLC parse of the pooled paths: with the group maps h264gpu_reader_set_group_maps left */
static int cavlc_parse_pooled_dev(h264gpu_ctx *ctx, const uint8_t *d_stream, uint64_t stream_len,
				  const struct h264gpu_slice_params *d_params, uint32_t n_slices,
				  struct h264gpu_mb_record *d_records, struct h264gpu_slice_result *d_results, void *stream)
{
	const uint8_t *maps = ctx->rd_maps_len ? (const uint8_t *)ctx->rd_maps.p : NULL;
	ctx->rd_maps_len = 0;
	return h264gpu_cavlc_parse_fmo_dev(ctx, d_stream, stream_len, d_params, n_slices, d_records, d_results, NULL,
					   maps, stream);
}

/* the slices of the buffer h264gpu_reader_scan left on the device (no second upload) */
extern "C" int h264gpu_reader_parse_cavlc(h264gpu_ctx *ctx, const struct h264gpu_slice_params *h_params,
					  uint32_t n_slices, uint64_t n_records,
					  const struct h264gpu_mb_record **h_records,
					  const struct h264gpu_slice_result **h_results)
{
	if (h_params == NULL || h_records == NULL || h_results == NULL || n_slices == 0)
		return -EINVAL;
	return parse_pooled(cavlc_parse_pooled_dev, ctx, NULL, 0, h_params, n_slices, n_records, h_records, h_results);
}

/* CAVLC and CABAC slices of one parameter list (two launches on the same records array) */
extern "C" int h264gpu_reader_parse_slices(h264gpu_ctx *ctx, const struct h264gpu_slice_params *h_params,
					   uint32_t n_slices, uint64_t n_records,
					   const struct h264gpu_mb_record **h_records,
					   const struct h264gpu_slice_result **h_results)
{
	if (h_params == NULL || h_records == NULL || h_results == NULL || n_slices == 0)
		return -EINVAL;
	return parse_pooled(cavlc_parse_pooled_dev, ctx, NULL, 0, h_params, n_slices, n_records, h_records, h_results,
			    h264gpu_cabac_parse_dev);
}

extern "C" int h264gpu_reader_parse_cabac(h264gpu_ctx *ctx, const struct h264gpu_slice_params *h_params,
					  uint32_t n_slices, uint64_t n_records,
					  const struct h264gpu_mb_record **h_records,
					  const struct h264gpu_slice_result **h_results)
{
	if (h_params == NULL || h_records == NULL || h_results == NULL || n_slices == 0)
		return -EINVAL;
	return parse_pooled(h264gpu_cabac_parse_dev, ctx, NULL, 0, h_params, n_slices, n_records, h_records, h_results);
}

extern "C" int h264gpu_cavlc_parse_host(h264gpu_ctx *ctx, const uint8_t *h_stream,
					uint64_t stream_len,
					const struct h264gpu_slice_params *h_params,
					uint32_t n_slices, struct h264gpu_mb_record *h_records,
					uint64_t n_records, struct h264gpu_slice_result *h_results)
{
	return parse_host(cavlc_parse_pooled_dev, ctx, h_stream, stream_len, h_params, n_slices, h_records,
			  n_records, h_results);
}

extern "C" int h264gpu_cabac_parse_host(h264gpu_ctx *ctx, const uint8_t *h_stream,
					uint64_t stream_len,
					const struct h264gpu_slice_params *h_params,
					uint32_t n_slices, struct h264gpu_mb_record *h_records,
					uint64_t n_records, struct h264gpu_slice_result *h_results)
{
	return parse_host(h264gpu_cabac_parse_dev, ctx, h_stream, stream_len, h_params, n_slices, h_records,
			  n_records, h_results);
}

/* ---- K5: CABAC ------------------------------------------------------------------------ */

extern "C" int h264gpu_cabac_parse_dev(h264gpu_ctx *ctx, const uint8_t *d_stream,
				       uint64_t stream_len,
				       const struct h264gpu_slice_params *d_params,
				       uint32_t n_slices, struct h264gpu_mb_record *d_records,
				       struct h264gpu_slice_result *d_results, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (n_slices == 0)
		return 0;
	if (d_stream == NULL || d_params == NULL || d_records == NULL || d_results == NULL)
		return -EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	/* neighbour ring per WORKING LANE: (PicWidthInMbs + 1) records of 64 B, sized for pictures up
	 * to 8192 luma samples wide (512 MBs) unless the host forms saw the parameter blocks; wider
	 * slices get -E2BIG. */
	const uint32_t ring_w = ctx->ring_w_hint ? ctx->ring_w_hint : 512;
	const uint64_t ring_stride = (uint64_t)(ring_w + 1);
	/* One warp per block (a block's shared memory stays at tables + 1.1 KB per slice), a persistent
	 * grid of at most 32 blocks per SM (what the register file holds), slices handed out longest
	 * first through a ticket counter; one slice per warp while the slices fit the grid.
	 *   H264GPU_CABAC_LANES_LOG2  log2 of the working lanes per warp (default: by slice count)
	 *   H264GPU_CABAC_WARPS       warps per block (1, 2, 4; default 1)
	 *   H264GPU_CABAC_PER_SM      resident warps per SM the grid is sized for (default 32)
	 *   H264GPU_CABAC_SORT=0      tickets in list order (A/B) */
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
	uint32_t lanes_log2 = 0;
	const char *env = getenv("H264GPU_CABAC_LANES_LOG2");
	if (env != NULL && atoi(env) >= 0 && atoi(env) <= 5) {
		lanes_log2 = (uint32_t)atoi(env);
	} else {
		/* more slices than resident warps: pack them into lanes (measured at 16000 slices with
		 * the persistent grid: 1 lane 258 ms, 2 lanes 235 ms, 4 lanes 206 ms; at 4000 slices 1 lane
		 * 74 ms, 2 lanes 93 ms: a warp's slices take turns, so packing only pays once every
		 * warp slot is taken) */
		while (lanes_log2 < 5 && ((uint64_t)n_slices >> lanes_log2) > (uint64_t)sms * 32)
			lanes_log2++;
	}
	uint32_t threads = 32;
	env = getenv("H264GPU_CABAC_WARPS");
	if (env != NULL && (atoi(env) == 1 || atoi(env) == 2 || atoi(env) == 4))
		threads = 32u * (uint32_t)atoi(env);
	uint32_t per_sm = 32;
	env = getenv("H264GPU_CABAC_PER_SM");
	if (env != NULL && atoi(env) >= 1 && atoi(env) <= 64)
		per_sm = (uint32_t)atoi(env);
	const size_t smem = cabac::smem_bytes((threads / 32) << lanes_log2);
	if (smem > ctx->cabac_smem_set) { /* function attributes are per device: cached per context */
		CU_TRY(cudaFuncSetAttribute(cabac::cabac_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
					    (int)smem));
		ctx->cabac_smem_set = smem;
	}
	/* the grid must be resident as a whole: a block that starts late starts with one of the longest
	 * slices (the first tickets are the working lanes' own numbers) */
	int resident = 0;
	CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, cabac::cabac_parse_kernel, (int)threads, smem));
	if (resident < 1)
		return -ENOMEM;
	if (per_sm > (uint32_t)resident * (threads / 32))
		per_sm = (uint32_t)resident * (threads / 32);
	uint64_t warps = ((uint64_t)n_slices + (1u << lanes_log2) - 1) >> lanes_log2;
	if (warps > (uint64_t)sms * per_sm)
		warps = (uint64_t)sms * per_sm;
	const uint32_t blocks = (uint32_t)((warps * 32 + threads - 1) / threads);
	const uint64_t grid_lanes = ((uint64_t)blocks * threads / 32) << lanes_log2;
	const size_t counter_off = (size_t)(grid_lanes * ring_stride * sizeof(cabac::Nb));
	const size_t order_off = counter_off + 16;
	r = h264gpu_ws_reserve(ctx, order_off + (size_t)n_slices * 4);
	if (r < 0)
		return r;
	cabac::CabacArgs a;
	a.stream = d_stream;
	a.stream_len = stream_len;
	a.params = d_params;
	a.n_slices = n_slices;
	a.records = d_records;
	a.results = d_results;
	a.ring = (cabac::Nb *)ctx->ws;
	a.ring_stride = ring_stride;
	a.ring_w = ring_w;
	a.lanes_log2 = lanes_log2;
	a.next_slice = (uint32_t *)((uint8_t *)ctx->ws + counter_off);
	a.order = (uint32_t *)((uint8_t *)ctx->ws + order_off);
	env = getenv("H264GPU_CABAC_SORT");
	cavlc2::order_kernel<<<1, CAVLC2_ORDER_T, 0, st>>>(d_params, n_slices, (uint32_t *)((uint8_t *)ctx->ws + order_off),
							    a.next_slice, (uint32_t)grid_lanes,
							    env != NULL && atoi(env) == 0 ? 0u : 1u);
	ctx->launches++;
	cabac::cabac_parse_kernel<<<blocks, threads, smem, st>>>(a);
	CU_TRY(cudaGetLastError());
	ctx->launches++;
	return 0;
}

/* ---- N4: concealment slices ------------------------------------------------------------- */

/* ---- N4: slice headers rewritten in bulk ------------------------------------------------------ */

extern "C" int h264gpu_patch_slice_headers_dev(h264gpu_ctx *ctx, uint8_t *d_stream, uint64_t stream_len,
					       const struct h264gpu_hdr_patch *d_patches, uint32_t n, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (n == 0)
		return 0;
	if (d_stream == NULL || d_patches == NULL)
		return -EINVAL;
	const uint64_t threads = (uint64_t)n * 64;
	conceal::patch_headers_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_stream, stream_len,
													     d_patches, n);
	CU_TRY(cudaGetLastError());
	ctx->launches++;
	return 0;
}

extern "C" int h264gpu_patch_slice_headers(h264gpu_ctx *ctx, uint8_t *d_stream, uint64_t stream_len,
					   const struct h264gpu_hdr_patch *h_patches, uint32_t n, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (n == 0)
		return 0;
	if (d_stream == NULL || h_patches == NULL)
		return -EINVAL;
	for (uint32_t i = 0; i < n; i++) {
		const struct h264gpu_hdr_patch *q = &h_patches[i];
		const uint64_t end = q->nal_off + q->nbytes + (q->tail_bits ? 1u : 0u);
		if (q->nbytes > 63u || q->tail_bits > 7u || end > stream_len || end < q->nal_off)
			return -EINVAL;
	}
	cudaStream_t st = (cudaStream_t)stream;
	if (st == NULL && (r = h264gpu_reader_stream(ctx, &st)) < 0)
		return r;
	if ((r = h264gpu_pool_dev(ctx, &ctx->rd_params, (size_t)n * sizeof(*h_patches))) < 0)
		return r;
	CU_TRY(cudaMemcpyAsync(ctx->rd_params.p, h_patches, (size_t)n * sizeof(*h_patches), cudaMemcpyHostToDevice, st));
	r = h264gpu_patch_slice_headers_dev(ctx, d_stream, stream_len, (const struct h264gpu_hdr_patch *)ctx->rd_params.p, n, st);
	if (r < 0)
		return r;
	CU_TRY(cudaStreamSynchronize(st)); /* h_patches is the caller's again, the stream is patched */
	return 0;
}

extern "C" int h264gpu_conceal_slices_dev(h264gpu_ctx *ctx, const struct h264gpu_conceal_params *d_params,
					  uint32_t n, const uint8_t *d_hdr, uint8_t *d_payload,
					  uint64_t payload_cap, uint64_t *d_off, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (d_off == NULL)
		return -EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	if (n == 0) {
		CU_TRY(cudaMemsetAsync(d_off, 0, 8, st));
		return 0;
	}
	if (d_params == NULL || d_hdr == NULL || d_payload == NULL)
		return -EINVAL;
	conceal::ConcealArgs a;
	a.params = d_params;
	a.n = n;
	a.hdr = d_hdr;
	a.off = d_off;
	a.out = d_payload;
	a.cap = payload_cap;
	const uint32_t blocks = (n + 127) / 128;
	conceal::conceal_kernel<false><<<blocks, 128, 0, st>>>(a); /* lengths */
	conceal::conceal_scan<<<1, 1024, 0, st>>>(d_off, n);
	conceal::conceal_kernel<true><<<blocks, 128, 0, st>>>(a);  /* bytes */
	CU_TRY(cudaGetLastError());
	ctx->launches += 3;
	return 0;
}

extern "C" int h264gpu_conceal_slices_host(h264gpu_ctx *ctx, const struct h264gpu_conceal_params *h_params,
					   uint32_t n, const uint8_t *h_hdr, uint64_t hdr_bytes, int sc_len,
					   uint8_t *h_out, uint64_t out_cap, uint64_t *h_out_off, uint64_t *total)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (total == NULL || (n && (h_params == NULL || h_hdr == NULL || h_out == NULL)))
		return -EINVAL;
	*total = 0;
	if (n == 0)
		return 0;
	/* an upper bound of the payload bytes: header + 2 bytes per macroblock + slack (a CABAC
	 * macroblock of these slices is at most 10 bins; a CAVLC one exactly one byte) */
	uint64_t bound = 0;
	for (uint32_t k = 0; k < n; k++)
		bound += (h_params[k].hdr_bits >> 3) + 2ull * h_params[k].mb_count + 16;
	cudaStream_t st;
	if ((r = h264gpu_reader_stream(ctx, &st)) < 0)
		return r;
	/* pooled buffers of the reader session, reused: params | header bytes | offsets (in) | payload |
	 * offsets (out) + total | stream */
	const size_t p_bytes = (size_t)n * sizeof(*h_params), o_bytes = ((size_t)n + 2) * 8;
	if ((r = h264gpu_pool_dev(ctx, &ctx->rd_params, p_bytes + hdr_bytes + 16)) < 0 ||
	    (r = h264gpu_pool_dev(ctx, &ctx->rd_tab, 2 * o_bytes)) < 0 ||
	    (r = h264gpu_pool_dev(ctx, &ctx->rd_records, bound + 16)) < 0 ||
	    (r = h264gpu_pool_dev(ctx, &ctx->rd_stream, out_cap + 64)) < 0)
		return r;
	ctx->rd_stream_len = 0; /* the resident reader buffer is gone */
	uint8_t *d_params = (uint8_t *)ctx->rd_params.p, *d_hdr = d_params + p_bytes;
	uint64_t *d_off = (uint64_t *)ctx->rd_tab.p, *d_out_off = d_off + n + 2, *d_total = d_out_off + n + 1;
	uint8_t *d_payload = (uint8_t *)ctx->rd_records.p, *d_out = (uint8_t *)ctx->rd_stream.p;
	CU_TRY(cudaMemcpyAsync(d_params, h_params, p_bytes, cudaMemcpyHostToDevice, st));
	CU_TRY(cudaMemcpyAsync(d_hdr, h_hdr, hdr_bytes, cudaMemcpyHostToDevice, st));
	r = h264gpu_conceal_slices_dev(ctx, (const struct h264gpu_conceal_params *)d_params, n, d_hdr, d_payload, bound,
				       d_off, st);
	if (r < 0)
		return r;
	r = h264gpu_frame_dev(ctx, d_payload, d_off, n, sc_len, d_out, out_cap, d_out_off, d_total, st);
	if (r < 0)
		return r;
	CU_TRY(cudaMemcpyAsync(total, d_total, 8, cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaStreamSynchronize(st));
	if (h_out_off)
		CU_TRY(cudaMemcpyAsync(h_out_off, d_out_off, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, st));
	const uint64_t ncopy = *total < out_cap ? *total : out_cap;
	if (ncopy)
		CU_TRY(cudaMemcpyAsync(h_out, d_out, ncopy, cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaStreamSynchronize(st));
	return *total > out_cap ? -ENOBUFS : 0;
}
