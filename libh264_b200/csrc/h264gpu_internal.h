/*
 * h264gpu_internal.h — context object shared by the C-ABI translation units.
 */
#ifndef H264GPU_INTERNAL_H
#define H264GPU_INTERNAL_H

#include <cuda_runtime.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "h264gpu.h"

struct h264gpu_ctx {
	int device;
	uint64_t launches;
	/* scan/frame workspace: [256 B control][tile descriptors / tile tables] */
	void *ws;
	size_t ws_bytes;
	/* gen-7 scan workspace (own buffer: zeroed once, re-armed by the finalize kernels) */
	void *ws7;
	size_t ws7_bytes;
	uint32_t epoch7;
	int ws7_vmm;          /* 1: ws7 is a mapped range of its own (ws7_va_bytes reserved, ws7_handle mapped) */
	size_t ws7_va_bytes;
	unsigned long long ws7_handle;
	uint64_t ws7_spans; /* spans of the last launch: another length moves the arrays behind the chain words */
	uint32_t attr_set; /* per-context (= per-device) cudaFuncSetAttribute done: bit 0 scan7 */
	int sms;
	/* host-buffer pipeline */
	cudaStream_t s_in, s_out, s_tab, s_up; /* kernels, RBSP downloads, table downloads, uploads */
	uint8_t *d_chunk_in[2];
	uint8_t *d_chunk_out[2];
	uint64_t *d_tab[2]; /* start | end | rbsp, tab_cap entries each */
	struct h264gpu_scan_result *d_res[2];
	struct h264gpu_scan_result *h_res; /* pinned, 2 entries */
	size_t chunk_bytes;
	size_t tab_cap;
	cudaEvent_t ev_in[2], ev_k[2], ev_out[2];
	int scan_items; /* 16-byte vectors per thread (1, 2 or 4) */
	/* reader session (h264gpu_reader_*, h264gpu_*_parse_host): grow-only device and pinned
	 * buffers on a private stream, so that a call costs no cudaMalloc */
	cudaStream_t s_rd;
	struct h264gpu_pool {
		void *p;
		size_t cap;
	} rd_stream, rd_tab, rd_res, rd_params, rd_records, rd_results, rd_maps; /* device */
	struct h264gpu_pool rh_tab, rh_res, rh_records, rh_results;         /* pinned host */
	uint64_t rd_stream_len; /* bytes of the buffer resident in rd_stream */
	uint32_t ring_w_hint;   /* widest picture (in macroblocks) of the parameter blocks the host forms have seen
				   for the launch they are about to make (0: unknown, rings sized for 512) */
	uint64_t rd_maps_len;   /* bytes of slice group maps in rd_maps for the next pooled parse (0: none) */
	size_t cabac_smem_set;  /* dynamic shared memory the CABAC kernel is configured for on this device */
	/* h264gpu_frame_host pipeline: upload / kernel / download streams over two chunk slots */
	struct {
		cudaStream_t s_up, s_k, s_dn;
		cudaEvent_t ev_up[2], ev_k[2], ev_dn[2];
		struct h264gpu_pool d_in[2], d_out[2], d_tab[2]; /* device */
		struct h264gpu_pool h_tab[2];                   /* pinned: offsets in, [out_off | total] back */
	} fp;
};

int h264gpu_pool_dev(h264gpu_ctx *ctx, struct h264gpu_ctx::h264gpu_pool *pl, size_t bytes);
int h264gpu_pool_host(h264gpu_ctx *ctx, struct h264gpu_ctx::h264gpu_pool *pl, size_t bytes);
int h264gpu_reader_stream(h264gpu_ctx *ctx, cudaStream_t *st);

#define CU_TRY(expr)                                                                       \
	do {                                                                               \
		cudaError_t _e = (expr);                                                   \
		if (_e != cudaSuccess) {                                                   \
			fprintf(stderr, "h264gpu: %s failed: %s (%s:%d)\n", #expr,         \
				cudaGetErrorString(_e), __FILE__, __LINE__);               \
			return _e == cudaErrorMemoryAllocation ? -ENOMEM : -EIO;           \
		}                                                                          \
	} while (0)

static inline int h264gpu_use(const h264gpu_ctx *ctx)
{
	if (ctx == NULL)
		return -EINVAL;
	CU_TRY(cudaSetDevice(ctx->device));
	return 0;
}

int h264gpu_ws_reserve(h264gpu_ctx *ctx, size_t bytes);

#endif /* H264GPU_INTERNAL_H */
