/*
 * h264gpu_slice_api.cu — C-ABI entry points of the slice-parallel macroblock
 * parse (include/h264gpu_slice.h).  Launch only; no CPU implementation exists.
 */
#include "h264gpu_internal.h"

#include "cavlc_parse.cuh"
#include "cabac_parse.cuh"

extern "C" int h264gpu_cavlc_parse_dev(h264gpu_ctx *ctx, const uint8_t *d_stream,
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
	cavlc::cavlc_parse_kernel<<<blocks, threads, 0, st>>>(a);
	CU_TRY(cudaGetLastError());
	ctx->launches++;
	return 0;
}

typedef int (*parse_dev_fn)(h264gpu_ctx *, const uint8_t *, uint64_t, const struct h264gpu_slice_params *,
			    uint32_t, struct h264gpu_mb_record *, struct h264gpu_slice_result *, void *);

static int parse_host(parse_dev_fn dev, h264gpu_ctx *ctx, const uint8_t *h_stream, uint64_t stream_len,
		      const struct h264gpu_slice_params *h_params, uint32_t n_slices,
		      struct h264gpu_mb_record *h_records, uint64_t n_records,
		      struct h264gpu_slice_result *h_results)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (n_slices == 0)
		return 0;
	if (h_stream == NULL || h_params == NULL || h_records == NULL || h_results == NULL)
		return -EINVAL;
	uint8_t *d_stream = NULL;
	struct h264gpu_slice_params *d_params = NULL;
	struct h264gpu_mb_record *d_records = NULL;
	struct h264gpu_slice_result *d_results = NULL;
	int rc = 0;
	cudaStream_t st = 0;
#define SL_TRY(expr)                                                                       \
	do {                                                                               \
		if ((expr) != cudaSuccess) {                                               \
			rc = -EIO;                                                         \
			goto out;                                                          \
		}                                                                          \
	} while (0)
	SL_TRY(cudaMalloc(&d_stream, stream_len + 16));
	SL_TRY(cudaMalloc(&d_params, (size_t)n_slices * sizeof(*d_params)));
	SL_TRY(cudaMalloc(&d_records, (n_records + 1) * sizeof(*d_records)));
	SL_TRY(cudaMalloc(&d_results, (size_t)n_slices * sizeof(*d_results)));
	SL_TRY(cudaMemcpyAsync(d_stream, h_stream, stream_len, cudaMemcpyHostToDevice, st));
	SL_TRY(cudaMemcpyAsync(d_params, h_params, (size_t)n_slices * sizeof(*d_params),
			       cudaMemcpyHostToDevice, st));
	rc = dev(ctx, d_stream, stream_len, d_params, n_slices, d_records, d_results, st);
	if (rc < 0)
		goto out;
	SL_TRY(cudaMemcpyAsync(h_records, d_records, n_records * sizeof(*d_records),
			       cudaMemcpyDeviceToHost, st));
	SL_TRY(cudaMemcpyAsync(h_results, d_results, (size_t)n_slices * sizeof(*d_results),
			       cudaMemcpyDeviceToHost, st));
	SL_TRY(cudaStreamSynchronize(st));
out:
#undef SL_TRY
	cudaFree(d_stream);
	cudaFree(d_params);
	cudaFree(d_records);
	cudaFree(d_results);
	return rc;
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
	static size_t smem_set = 0;
	if (smem > smem_set) {
		CU_TRY(cudaFuncSetAttribute(cabac::cabac_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
					    (int)smem));
		smem_set = smem;
	}
	cabac::cabac_parse_kernel<<<blocks, threads, smem, st>>>(a);
	CU_TRY(cudaGetLastError());
	ctx->launches++;
	return 0;
}
