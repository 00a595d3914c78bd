/*
 * h264gpu_slice_api.cu — C-ABI entry points of the slice-parallel macroblock
 * parse (include/h264gpu_slice.h).  Launch only; no CPU implementation exists.
 */
#include "h264gpu_internal.h"

#include "cavlc_parse.cuh"
#define CAVLC_NS cavlc_full
#define CAVLC_FULL 1
#include "cavlc_parse.cuh" /* the same parse with struct h264_mb_syntax records (opt-in entry point) */
#include "cabac_parse.cuh"

extern "C" int h264gpu_cavlc_parse_dev(h264gpu_ctx *ctx, const uint8_t *d_stream,
				       uint64_t stream_len,
				       const struct h264gpu_slice_params *d_params,
				       uint32_t n_slices, struct h264gpu_mb_record *d_records,
				       struct h264gpu_slice_result *d_results, void *stream)
{
	return h264gpu_cavlc_parse_full_dev(ctx, d_stream, stream_len, d_params, n_slices, d_records, d_results, NULL,
					    stream);
}

extern "C" int h264gpu_cavlc_parse_full_dev(h264gpu_ctx *ctx, const uint8_t *d_stream, uint64_t stream_len,
					    const struct h264gpu_slice_params *d_params, uint32_t n_slices,
					    struct h264gpu_mb_record *d_records, struct h264gpu_slice_result *d_results,
					    struct h264_mb_syntax *d_syntax, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (n_slices == 0)
		return 0;
	if (d_stream == NULL || d_params == NULL || d_records == NULL || d_results == NULL)
		return -EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	/* nC context ring per slice: (PicWidthInMbs + 1) macroblocks x 48 counts, sized
	 * for pictures up to 8192 luma samples wide (512 MBs); wider slices get -E2BIG. */
	const uint32_t ring_w = 512;
	const uint64_t ring_stride = (uint64_t)(ring_w + 1) * 48;
	const size_t need = (size_t)n_slices * ring_stride;
	r = h264gpu_ws_reserve(ctx, need);
	if (r < 0)
		return r;
	cavlc::CavlcArgs a;
	a.stream = d_stream;
	a.stream_len = stream_len;
	a.params = d_params;
	a.n_slices = n_slices;
	a.records = d_records;
	a.results = d_results;
	a.ring = (uint8_t *)ctx->ws;
	a.ring_stride = ring_stride;
	a.ring_w = ring_w;
	a.syntax = d_syntax;
	/* Slices per warp.  Slices diverge completely, so a warp runs its lanes one after the
	 * other: one slice per warp is best while the warps fit the machine, and lanes are packed
	 * only beyond ~32 warps per SM (measured, profiles/r01_slice_lane_packing.txt: 4000 slices
	 * -> 1 per warp, 16000 slices -> 4 per warp). */
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
	uint32_t lanes_log2 = 0;
	const char *env = getenv("H264GPU_CAVLC_LANES_LOG2");
	if (env != NULL && atoi(env) >= 0 && atoi(env) <= 5) {
		lanes_log2 = (uint32_t)atoi(env);
	} else {
		while (lanes_log2 < 5 && ((uint64_t)n_slices >> lanes_log2) > (uint64_t)sms * 32)
			lanes_log2++;
	}
	a.lanes_log2 = lanes_log2;
	const uint32_t threads = 128; /* 4 warps per block */
	const uint64_t warps = ((uint64_t)n_slices + (1u << lanes_log2) - 1) >> lanes_log2;
	const uint32_t blocks = (uint32_t)((warps * 32 + threads - 1) / threads);
	if (d_syntax == NULL) {
		cavlc::cavlc_parse_kernel<<<blocks, threads, 0, st>>>(a);
	} else {
		cavlc_full::CavlcArgs f;
		static_assert(sizeof(f) == sizeof(a), "same argument block in both instantiations");
		memcpy(&f, &a, sizeof(f));
		cavlc_full::cavlc_parse_kernel<<<blocks, threads, 0, st>>>(f);
	}
	CU_TRY(cudaGetLastError());
	ctx->launches++;
	return 0;
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

/* the slices of the buffer h264gpu_reader_scan left on the device (no second upload) */
extern "C" int h264gpu_reader_parse_cavlc(h264gpu_ctx *ctx, const struct h264gpu_slice_params *h_params,
					  uint32_t n_slices, uint64_t n_records,
					  const struct h264gpu_mb_record **h_records,
					  const struct h264gpu_slice_result **h_results)
{
	if (h_params == NULL || h_records == NULL || h_results == NULL || n_slices == 0)
		return -EINVAL;
	return parse_pooled(h264gpu_cavlc_parse_dev, ctx, NULL, 0, h_params, n_slices, n_records, h_records, h_results);
}

/* CAVLC and CABAC slices of one parameter list (two launches on the same records array) */
extern "C" int h264gpu_reader_parse_slices(h264gpu_ctx *ctx, const struct h264gpu_slice_params *h_params,
					   uint32_t n_slices, uint64_t n_records,
					   const struct h264gpu_mb_record **h_records,
					   const struct h264gpu_slice_result **h_results)
{
	if (h_params == NULL || h_records == NULL || h_results == NULL || n_slices == 0)
		return -EINVAL;
	return parse_pooled(h264gpu_cavlc_parse_dev, ctx, NULL, 0, h_params, n_slices, n_records, h_records, h_results,
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
	return parse_host(h264gpu_cavlc_parse_dev, ctx, h_stream, stream_len, h_params, n_slices, h_records,
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
	(void)stream_len;
	cudaStream_t st = (cudaStream_t)stream;
	/* neighbour ring per slice: (PicWidthInMbs + 1) records of 64 B, sized for pictures up
	 * to 8192 luma samples wide (512 MBs); wider slices get -E2BIG. */
	const uint32_t ring_w = 512;
	const uint64_t ring_stride = (uint64_t)(ring_w + 1);
	r = h264gpu_ws_reserve(ctx, (size_t)n_slices * ring_stride * sizeof(cabac::Nb));
	if (r < 0)
		return r;
	cabac::CabacArgs a;
	a.stream = d_stream;
	a.params = d_params;
	a.n_slices = n_slices;
	a.records = d_records;
	a.results = d_results;
	a.ring = (cabac::Nb *)ctx->ws;
	a.ring_stride = ring_stride;
	a.ring_w = ring_w;
	/* slices per warp: same rule as the CAVLC parse; one warp per block so that a block's
	 * shared memory (tables + 1.1 KB per slice) stays small and many blocks fit an SM */
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
	uint32_t lanes_log2 = 0;
	const char *env = getenv("H264GPU_CABAC_LANES_LOG2");
	if (env != NULL && atoi(env) >= 0 && atoi(env) <= 5) {
		lanes_log2 = (uint32_t)atoi(env);
	} else {
		while (lanes_log2 < 5 && ((uint64_t)n_slices >> lanes_log2) > (uint64_t)sms * 32)
			lanes_log2++;
	}
	a.lanes_log2 = lanes_log2;
	const uint32_t threads = 32; /* one warp per block */
	const uint64_t warps = ((uint64_t)n_slices + (1u << lanes_log2) - 1) >> lanes_log2;
	const uint32_t blocks = (uint32_t)((warps * 32 + threads - 1) / threads);
	const size_t smem = cabac::smem_bytes((threads / 32) << lanes_log2);
	if (smem > ctx->cabac_smem_set) { /* function attributes are per device: cached per context */
		CU_TRY(cudaFuncSetAttribute(cabac::cabac_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
					    (int)smem));
		ctx->cabac_smem_set = smem;
	}
	cabac::cabac_parse_kernel<<<blocks, threads, smem, st>>>(a);
	CU_TRY(cudaGetLastError());
	ctx->launches++;
	return 0;
}
