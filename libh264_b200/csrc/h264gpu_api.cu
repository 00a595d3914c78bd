/*
 * h264gpu_api.cu — C-ABI entry points (include/h264gpu.h): context, memory and
 * timing helpers, the scan(+strip) launcher, the shard merge and the host-buffer
 * pipeline.  Host code here only moves bytes and launches kernels; there is no
 * CPU implementation of any stage (a missing GPU is an error, never a fallback).
 */
#include "h264gpu_internal.h"

#include "annexb_scan.cuh"
#include "annexb_scan2.cuh"
#include "annexb_scan7.cuh"

extern "C" {

const char *h264gpu_version(void)
{
	return "libh264gpu 0.1 (sm_100a)";
}

int h264gpu_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess)
		return -ENODEV;
	return n;
}

static int ws7_reserve(h264gpu_ctx *ctx, size_t bytes);
static void ws7_free(h264gpu_ctx *ctx);
static void frame_pipe_free(h264gpu_ctx *ctx);

int h264gpu_create(int device, h264gpu_ctx **out)
{
	if (out == NULL)
		return -EINVAL;
	*out = NULL;
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0)
		return -ENODEV;
	if (device < 0 || device >= n)
		return -EINVAL;
	CU_TRY(cudaSetDevice(device));
	h264gpu_ctx *ctx = (h264gpu_ctx *)calloc(1, sizeof(*ctx));
	if (ctx == NULL)
		return -ENOMEM;
	ctx->device = device;
	/* tile shape of the packed-RBSP kernel and of the writer kernel: 16-byte chunks per thread
	 * (8 = 32 KiB tiles; 1, 2, 4 = small tiles, used by the tests to get many seams on small inputs) */
	ctx->scan_items = 8;
	const char *e = getenv("H264GPU_SCAN_ITEMS");
	if (e != NULL) {
		int v = atoi(e);
		if (v == 1 || v == 2 || v == 4 || v == 8)
			ctx->scan_items = v;
	}
	size_t chunk_mb = 128;
	e = getenv("H264GPU_CHUNK_MB");
	if (e != NULL && atoi(e) > 0)
		chunk_mb = (size_t)atoi(e);
	ctx->chunk_bytes = chunk_mb << 20;
	*out = ctx;
	return 0;
}

static void free_pipeline(h264gpu_ctx *ctx)
{
	for (int b = 0; b < 2; b++) {
		cudaFree(ctx->d_chunk_in[b]);
		cudaFree(ctx->d_chunk_out[b]);
		cudaFree(ctx->d_tab[b]);
		cudaFree(ctx->d_res[b]);
		ctx->d_chunk_in[b] = ctx->d_chunk_out[b] = NULL;
		ctx->d_tab[b] = NULL;
		ctx->d_res[b] = NULL;
		if (ctx->ev_k[b]) {
			cudaEventDestroy(ctx->ev_in[b]);
			cudaEventDestroy(ctx->ev_k[b]);
			cudaEventDestroy(ctx->ev_out[b]);
			ctx->ev_k[b] = NULL;
		}
	}
	if (ctx->h_res)
		cudaFreeHost(ctx->h_res);
	ctx->h_res = NULL;
	if (ctx->s_in) {
		cudaStreamDestroy(ctx->s_in);
		cudaStreamDestroy(ctx->s_out);
		cudaStreamDestroy(ctx->s_tab);
		if (ctx->s_up)
			cudaStreamDestroy(ctx->s_up);
		ctx->s_in = ctx->s_out = ctx->s_tab = ctx->s_up = NULL;
	}
}

int h264gpu_destroy(h264gpu_ctx *ctx)
{
	if (ctx == NULL)
		return 0;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	free_pipeline(ctx);
	frame_pipe_free(ctx);
	cudaFree(ctx->ws);
	ws7_free(ctx);
	cudaFree(ctx->rd_stream.p);
	cudaFree(ctx->rd_tab.p);
	cudaFree(ctx->rd_res.p);
	cudaFree(ctx->rd_params.p);
	cudaFree(ctx->rd_records.p);
	cudaFree(ctx->rd_results.p);
	cudaFree(ctx->rd_maps.p);
	cudaFreeHost(ctx->rh_tab.p);
	cudaFreeHost(ctx->rh_res.p);
	cudaFreeHost(ctx->rh_records.p);
	cudaFreeHost(ctx->rh_results.p);
	if (ctx->s_rd)
		cudaStreamDestroy(ctx->s_rd);
	free(ctx);
	return 0;
}

int h264gpu_device(const h264gpu_ctx *ctx)
{
	return ctx ? ctx->device : -EINVAL;
}

uint64_t h264gpu_launch_count(const h264gpu_ctx *ctx)
{
	return ctx ? ctx->launches : 0;
}

int h264gpu_malloc(h264gpu_ctx *ctx, size_t bytes, void **d_ptr)
{
	int r = h264gpu_use(ctx);
	if (r < 0 || d_ptr == NULL)
		return r < 0 ? r : -EINVAL;
	CU_TRY(cudaMalloc(d_ptr, bytes ? bytes : 16));
	return 0;
}

int h264gpu_free(h264gpu_ctx *ctx, void *d_ptr)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	CU_TRY(cudaFree(d_ptr));
	return 0;
}

int h264gpu_host_alloc(h264gpu_ctx *ctx, size_t bytes, void **h_ptr)
{
	int r = h264gpu_use(ctx);
	if (r < 0 || h_ptr == NULL)
		return r < 0 ? r : -EINVAL;
	CU_TRY(cudaMallocHost(h_ptr, bytes ? bytes : 16));
	return 0;
}

int h264gpu_host_free(h264gpu_ctx *ctx, void *h_ptr)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	CU_TRY(cudaFreeHost(h_ptr));
	return 0;
}

int h264gpu_memcpy_h2d(h264gpu_ctx *ctx, void *d_dst, const void *h_src, size_t bytes,
		       void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	CU_TRY(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
	return 0;
}

int h264gpu_memcpy_d2h(h264gpu_ctx *ctx, void *h_dst, const void *d_src, size_t bytes,
		       void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	CU_TRY(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
	return 0;
}

int h264gpu_sync(h264gpu_ctx *ctx, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
	return 0;
}

int h264gpu_stream_create(h264gpu_ctx *ctx, void **stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0 || stream == NULL)
		return r < 0 ? r : -EINVAL;
	cudaStream_t st;
	CU_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
	*stream = (void *)st;
	return 0;
}

int h264gpu_stream_destroy(h264gpu_ctx *ctx, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	CU_TRY(cudaStreamDestroy((cudaStream_t)stream));
	return 0;
}

/* ---- timers ------------------------------------------------------------ */

struct gpu_timer {
	cudaEvent_t a, b;
};

int h264gpu_timer_create(h264gpu_ctx *ctx, void **timer)
{
	int r = h264gpu_use(ctx);
	if (r < 0 || timer == NULL)
		return r < 0 ? r : -EINVAL;
	gpu_timer *t = (gpu_timer *)calloc(1, sizeof(*t));
	if (t == NULL)
		return -ENOMEM;
	CU_TRY(cudaEventCreate(&t->a));
	CU_TRY(cudaEventCreate(&t->b));
	*timer = t;
	return 0;
}

int h264gpu_timer_destroy(h264gpu_ctx *ctx, void *timer)
{
	gpu_timer *t = (gpu_timer *)timer;
	if (t == NULL || h264gpu_use(ctx) < 0)
		return -EINVAL;
	cudaEventDestroy(t->a);
	cudaEventDestroy(t->b);
	free(t);
	return 0;
}

int h264gpu_timer_start(h264gpu_ctx *ctx, void *timer, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0 || timer == NULL)
		return r < 0 ? r : -EINVAL;
	CU_TRY(cudaEventRecord(((gpu_timer *)timer)->a, (cudaStream_t)stream));
	return 0;
}

int h264gpu_timer_stop(h264gpu_ctx *ctx, void *timer, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0 || timer == NULL)
		return r < 0 ? r : -EINVAL;
	CU_TRY(cudaEventRecord(((gpu_timer *)timer)->b, (cudaStream_t)stream));
	return 0;
}

int h264gpu_timer_elapsed_ms(h264gpu_ctx *ctx, void *timer, float *ms)
{
	int r = h264gpu_use(ctx);
	if (r < 0 || timer == NULL || ms == NULL)
		return r < 0 ? r : -EINVAL;
	gpu_timer *t = (gpu_timer *)timer;
	CU_TRY(cudaEventSynchronize(t->b));
	CU_TRY(cudaEventElapsedTime(ms, t->a, t->b));
	return 0;
}

} /* extern "C" */

/* ---- workspace ---------------------------------------------------------- */

int h264gpu_ws_reserve(h264gpu_ctx *ctx, size_t bytes)
{
	if (bytes <= ctx->ws_bytes)
		return 0;
	/* growing is rare: drain the device, then replace the buffer */
	CU_TRY(cudaDeviceSynchronize());
	if (ctx->ws)
		CU_TRY(cudaFree(ctx->ws));
	ctx->ws = NULL;
	ctx->ws_bytes = 0;
	size_t want = (bytes + (1u << 20)) & ~(size_t)((1u << 20) - 1);
	CU_TRY(cudaMalloc(&ctx->ws, want));
	ctx->ws_bytes = want;
	return 0;
}

/* ---- scan (+strip) ------------------------------------------------------- */

template <int CPT, int MINB>
static cudaError_t launch_scan2(const annexb::ScanArgs &a, bool strip, cudaStream_t st)
{
	/* function attributes are per device: set on every launch path (a cheap driver call) */
	cudaFuncSetAttribute(annexb2::scan2_kernel<CPT, true, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
	cudaFuncSetAttribute(annexb2::scan2_kernel<CPT, false, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
	if (strip)
		annexb2::scan2_kernel<CPT, true, MINB><<<a.num_tiles, annexb2::kT, 0, st>>>(a);
	else
		annexb2::scan2_kernel<CPT, false, MINB><<<a.num_tiles, annexb2::kT, 0, st>>>(a);
	return cudaGetLastError();
}

static uint64_t scan_tile_bytes(int items)
{
	return (uint64_t)annexb2::kT * items * 16;
}

extern "C" int h264gpu_split_strip_dev(h264gpu_ctx *ctx, const uint8_t *d_in, uint64_t len,
				       uint64_t base, const struct h264gpu_shard_edge *edge,
				       uint8_t *d_rbsp, uint64_t *d_nal_start, uint64_t *d_nal_end,
				       uint64_t *d_nal_rbsp, uint64_t nal_cap,
				       struct h264gpu_scan_result *d_result, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (d_in == NULL || d_result == NULL || (nal_cap && (!d_nal_start || !d_nal_end)))
		return -EINVAL;
	if (((uintptr_t)d_in & 15) || ((uintptr_t)d_rbsp & 15))
		return -EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	if (len == 0) {
		CU_TRY(cudaMemsetAsync(d_result, 0, sizeof(*d_result), st));
		return 0;
	}
	const int items = ctx->scan_items;
	const uint64_t tile = scan_tile_bytes(items);
	const uint64_t ntiles = (len + tile - 1) / tile;
	if (ntiles > 0x3fffffffull)
		return -E2BIG;
	/* workspace: [256 control][32 B descriptor per tile] */
	const size_t need = 256 + (size_t)ntiles * 32;
	r = h264gpu_ws_reserve(ctx, need);
	if (r < 0)
		return r;
	/* one memset arms the ticket and invalidates every tile descriptor */
	CU_TRY(cudaMemsetAsync(ctx->ws, 0xff, need, st));
	CU_TRY(cudaMemsetAsync(d_result, 0xff, sizeof(*d_result), st));

	annexb::ScanArgs a;
	memset(&a, 0, sizeof(a));
	a.in = d_in;
	a.len = len;
	a.base = base;
	a.rbsp = d_rbsp;
	a.nal_start = d_nal_start;
	a.nal_end = d_nal_end;
	a.nal_rbsp = d_nal_rbsp;
	a.nal_cap = nal_cap;
	a.ticket = (uint32_t *)ctx->ws;
	a.desc = (uint64_t *)((uint8_t *)ctx->ws + 256);
	a.result = d_result;
	a.num_tiles = (uint32_t)ntiles;
	a.halo_left = 0xffffffffu;
	a.right[0] = a.right[1] = 0xff;
	if (edge != NULL) {
		if (edge->has_left)
			a.halo_left = 0xffffu | (uint32_t)edge->left[0] << 16 |
				      (uint32_t)edge->left[1] << 24;
		a.has_right = edge->has_right ? 1 : 0;
		a.right[0] = edge->right[0];
		a.right[1] = edge->right[1];
		a.init_in = edge->assume_in ? 1 : 0;
	}
	cudaError_t ce;
	const bool strip = d_rbsp != NULL;
	/* diagnostics: per-tile phase timestamps dumped to the file named by H264GPU_SCAN_TRACE */
	const char *trace_path = getenv("H264GPU_SCAN_TRACE");
	uint64_t *d_trace = NULL;
	if (trace_path != NULL) {
		CU_TRY(cudaMalloc(&d_trace, ntiles * 64));
		CU_TRY(cudaMemsetAsync(d_trace, 0, ntiles * 64, st));
		a.trace = d_trace;
	}
	if (items == 8)
		ce = launch_scan2<8, 4>(a, strip, st);
	else if (items == 4)
		ce = launch_scan2<4, 6>(a, strip, st);
	else if (items == 2)
		ce = launch_scan2<2, 6>(a, strip, st);
	else
		ce = launch_scan2<1, 6>(a, strip, st);
	CU_TRY(ce);
	ctx->launches++;
	if (d_trace != NULL) {
		uint64_t *h = (uint64_t *)malloc(ntiles * 64);
		CU_TRY(cudaStreamSynchronize(st));
		CU_TRY(cudaMemcpy(h, d_trace, ntiles * 64, cudaMemcpyDeviceToHost));
		FILE *f = fopen(trace_path, "wb");
		if (f != NULL) {
			fwrite(h, 64, ntiles, f);
			fclose(f);
		}
		free(h);
		cudaFree(d_trace);
	}
	return 0;
}

/* ---- scan + strip, RBSP in place: gen 7 (annexb_scan7.cuh) ------------------------ */

/* the scan's own workspace: control words are zero at allocation and re-armed by the last
 * finalize block of every launch; chain words carry the launch epoch, so nothing is cleared
 * between launches */
/*
 * Placement of the workspace.  The in-place kernel polls and publishes chain words here, and the
 * same launch runs 2.00 or 2.85 ms per 4 GiB depending on where cudaMalloc happened to put the
 * buffer (profiles/r02_scan_workspace_placement.txt: fast only when it lands directly below the
 * stream buffers, slow next to a page-locked host mapping or among the context's first small
 * allocations).  H264GPU_WS7_VMM=1 takes cudaMalloc out of the picture: the workspace gets a
 * virtual range of its own (1 GiB aligned, 1 GiB reserved, so no other mapping shares its page
 * directory entries) backed by one physical allocation of the recommended (2 MiB) granularity,
 * through the driver's virtual memory management entry points (fetched from the runtime, no
 * link-time dependency on libcuda).
 */
#include <cuda.h>

struct vmm_api {
	CUresult (*granularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags);
	CUresult (*reserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
	CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
	CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
	CUresult (*set_access)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
	CUresult (*unmap)(CUdeviceptr, size_t);
	CUresult (*release)(CUmemGenericAllocationHandle);
	CUresult (*addr_free)(CUdeviceptr, size_t);
	int ok;
};

static const vmm_api *vmm(void)
{
	static vmm_api api;
	static int tried = 0;
	if (!tried) {
		tried = 1;
		cudaDriverEntryPointQueryResult q;
		int ok = 1;
#define H264_VMM_SYM(field, name)                                                                   \
	ok = ok && cudaGetDriverEntryPoint(name, (void **)&api.field, cudaEnableDefault, &q) == cudaSuccess && \
	     q == cudaDriverEntryPointSuccess && api.field != NULL
		H264_VMM_SYM(granularity, "cuMemGetAllocationGranularity");
		H264_VMM_SYM(reserve, "cuMemAddressReserve");
		H264_VMM_SYM(create, "cuMemCreate");
		H264_VMM_SYM(map, "cuMemMap");
		H264_VMM_SYM(set_access, "cuMemSetAccess");
		H264_VMM_SYM(unmap, "cuMemUnmap");
		H264_VMM_SYM(release, "cuMemRelease");
		H264_VMM_SYM(addr_free, "cuMemAddressFree");
#undef H264_VMM_SYM
		api.ok = ok;
	}
	return api.ok ? &api : NULL;
}

static void ws7_free(h264gpu_ctx *ctx)
{
	if (ctx->ws7 == NULL)
		return;
	if (ctx->ws7_vmm) {
		const vmm_api *v = vmm();
		v->unmap((CUdeviceptr)ctx->ws7, ctx->ws7_bytes);
		v->release((CUmemGenericAllocationHandle)ctx->ws7_handle);
		v->addr_free((CUdeviceptr)ctx->ws7, ctx->ws7_va_bytes);
	} else {
		cudaFree(ctx->ws7);
	}
	ctx->ws7 = NULL;
	ctx->ws7_bytes = 0;
	ctx->ws7_vmm = 0;
}

/* 0: mapped, < 0: not available (the caller falls back to cudaMalloc) */
static int ws7_map_own_range(h264gpu_ctx *ctx, size_t want)
{
	const vmm_api *v = vmm();
	if (v == NULL)
		return -ENOSYS;
	CUmemAllocationProp prop;
	memset(&prop, 0, sizeof(prop));
	prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
	prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
	prop.location.id = ctx->device;
	size_t gran = 0;
	if (v->granularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0)
		return -ENOSYS;
	const size_t size = (want + gran - 1) / gran * gran;
	const size_t gib = (size_t)1 << 30;
	const size_t va_bytes = (size + gib - 1) / gib * gib;
	CUdeviceptr va = 0;
	if (v->reserve(&va, va_bytes, gib, 0, 0) != CUDA_SUCCESS)
		return -ENOMEM;
	CUmemGenericAllocationHandle h;
	if (v->create(&h, size, &prop, 0) != CUDA_SUCCESS) {
		v->addr_free(va, va_bytes);
		return -ENOMEM;
	}
	CUmemAccessDesc acc;
	memset(&acc, 0, sizeof(acc));
	acc.location = prop.location;
	acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
	if (v->map(va, size, 0, h, 0) != CUDA_SUCCESS || v->set_access(va, size, &acc, 1) != CUDA_SUCCESS) {
		v->unmap(va, size);
		v->release(h);
		v->addr_free(va, va_bytes);
		return -ENOMEM;
	}
	ctx->ws7 = (void *)va;
	ctx->ws7_bytes = size;
	ctx->ws7_va_bytes = va_bytes;
	ctx->ws7_handle = (unsigned long long)h;
	ctx->ws7_vmm = 1;
	return 0;
}

static int ws7_reserve(h264gpu_ctx *ctx, size_t bytes)
{
	if (bytes <= ctx->ws7_bytes)
		return 0;
	CU_TRY(cudaDeviceSynchronize());
	ws7_free(ctx);
	const size_t want = (bytes + (bytes >> 3) + (1u << 20)) & ~(size_t)((1u << 20) - 1);
	const char *e = getenv("H264GPU_WS7_PAD_MB");
	if (e != NULL && atoi(e) > 0) {
		/* diagnostics: a spacer allocation (kept until the process ends) moves the workspace */
		void *pad = NULL;
		cudaMalloc(&pad, (size_t)atoi(e) << 20);
	}
	e = getenv("H264GPU_WS7_VMM");
	if (e == NULL || atoi(e) == 0 || ws7_map_own_range(ctx, want) < 0) {
		CU_TRY(cudaMalloc(&ctx->ws7, want));
		ctx->ws7_bytes = want;
	}
	if (getenv("H264GPU_DEBUG_WS") != NULL)
		fprintf(stderr, "h264gpu: ws7 %p + %zu%s\n", ctx->ws7, ctx->ws7_bytes, ctx->ws7_vmm ? " (own range)" : "");
	CU_TRY(cudaMemset(ctx->ws7, 0, ctx->ws7_bytes));
	return 0;
}

static int device_sms(h264gpu_ctx *ctx)
{
	if (ctx->sms == 0) {
		int sms = 0;
		cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
		ctx->sms = sms > 0 ? sms : 148;
	}
	return ctx->sms;
}

template <int ROWS, int CTAS>
static int scan7_launch_t(h264gpu_ctx *ctx, const uint8_t *d_in, uint64_t len, uint64_t base,
			  const struct h264gpu_shard_edge *edge, uint8_t *d_rbsp, uint64_t *d_nal_start,
			  uint64_t *d_nal_end, uint64_t *d_nal_rbsp, uint64_t *d_nal_rbsp_len, uint64_t nal_cap,
			  struct h264gpu_scan_result *d_result, cudaStream_t st)
{
	const uint64_t span = (uint64_t)annexb7::Cfg<ROWS>::SPAN;
	/* a boundary event is owned by its third byte: the launch covers the two edge bytes too */
	const uint64_t nspans = (len + 2 + span - 1) / span;
	if (nspans > 0x7fffffffull)
		return -E2BIG;
	/* events: a start code and at most a few terminators per NAL in real streams */
	uint64_t ev_cap = 4 * nal_cap + 4096;
	if (ev_cap > len / 3 + 2)
		ev_cap = len / 3 + 2;
	if (ev_cap >= 0xffffffffull)
		ev_cap = 0xfffffffeull;
	const uint64_t nblk = (nspans + annexb7::kFinT - 1) / annexb7::kFinT;
	/* [control][chain][fin][span_pre][blk: 2 per fin block][totals][events][ordered events] */
	const size_t chain_off = annexb7::kCtrlBytes;
	const size_t fin_off = chain_off + (size_t)nspans * 8;
	const size_t pre_off = fin_off + (size_t)nspans * 8;
	const size_t blk_off = pre_off + (size_t)nspans * 8;
	const size_t tot_off = blk_off + (size_t)nblk * 16;
	const size_t ev_off = tot_off + 64;
	const size_t ord_off = ev_off + (size_t)ev_cap * 16;
	const size_t def_off = ord_off + (size_t)ev_cap * 24;
	const size_t need = def_off + (size_t)nspans * 4;
	const size_t had = ctx->ws7_bytes;
	int r = ws7_reserve(ctx, need);
	if (r < 0)
		return r;
	if (ctx->ws7_bytes != had) {
		ctx->epoch7 = 0; /* fresh zeroed buffer */
	} else if (ctx->ws7_spans != nspans) {
		/* another stream length: the arrays behind the chain words move, and what a shorter
		 * launch left where chain words of this one will be (span words whose top 16 bits are a
		 * start-code count) could pass for a word of this launch's epoch */
		CU_TRY(cudaMemsetAsync((uint8_t *)ctx->ws7 + chain_off, 0, (size_t)nspans * 8, st));
	}
	ctx->ws7_spans = nspans;
	if (++ctx->epoch7 > 0xffffu) {
		/* epoch wrap: stale chain words of 65535 launches ago could look valid */
		CU_TRY(cudaMemsetAsync((uint8_t *)ctx->ws7 + chain_off, 0, (size_t)nspans * 8, st));
		ctx->epoch7 = 1;
	}
	uint8_t *ws = (uint8_t *)ctx->ws7;

	annexb7::Scan7Args a;
	memset(&a, 0, sizeof(a));
	a.in = d_in;
	a.len = len;
	a.rbsp = d_rbsp;
	a.chain = (uint64_t *)(ws + chain_off);
	a.fin = (uint64_t *)(ws + fin_off);
	a.ctrl = (uint32_t *)ws;
	a.evbuf = (uint64_t *)(ws + ev_off);
	a.ev_cap = ev_cap;
	a.deferred = (uint32_t *)(ws + def_off);
	a.num_spans = (uint32_t)nspans;
	a.halo_left = 0xffffffffu;
	a.epoch = ctx->epoch7;
	a.right[0] = a.right[1] = 0xff;
	int assume_in = 0;
	if (edge != NULL) {
		if (edge->has_left)
			a.halo_left = 0xffffu | (uint32_t)edge->left[0] << 16 | (uint32_t)edge->left[1] << 24;
		a.has_right = edge->has_right ? 1 : 0;
		a.right[0] = edge->right[0];
		a.right[1] = edge->right[1];
		assume_in = edge->assume_in ? 1 : 0;
	}
	const int sms = device_sms(ctx);
	const uint64_t ctas_needed = (nspans + annexb7::kW - 1) / annexb7::kW;
	if (d_rbsp != NULL) {
		const uint32_t grid = (uint32_t)(ctas_needed < (uint64_t)sms * CTAS ? ctas_needed : (uint64_t)sms * CTAS);
		/* tickets go round robin over K regions: a span's predecessor was taken K tickets
		 * (K x ~2 ns) before it.  K = a quarter of the resident warps gives ~3 us of lead, the
		 * time a span needs from its ticket to its chain word; every region start defers ~17
		 * spans (up to the first start code) to the second pass, so K is also kept below 1/512
		 * of the spans (at most ~3 % of them deferred).  Measured at 4 GiB: K = 1 1719, 370 2068, 740 2108, 1480 2136, 5920 2116 GB/s */
		uint32_t K = grid * annexb7::kW / 4;
		{
			const char *e = getenv("H264GPU_SCAN7_REGIONS");
			if (e != NULL && atoi(e) > 0)
				K = (uint32_t)atoi(e);
		}
		if (K > nspans / 512)
			K = (uint32_t)(nspans / 512);
		if (K < 1)
			K = 1;
		a.regions = K;
		a.region_len = (uint32_t)((nspans + K - 1) / K);
		a.pf_dist = 1; /* the next span of the region: taken K tickets from now */
		{
			const char *e = getenv("H264GPU_SCAN7_PF");
			if (e != NULL)
				a.pf_dist = (uint32_t)atoi(e);
		}
		/* ticket counters: one word cannot hand out more than a ticket per ~1.9 ns (annexb_scan7.cuh) */
		a.tick_n = 8;
		{
			const char *e = getenv("H264GPU_SCAN7_TICKS");
			if (e != NULL && atoi(e) >= 1 && atoi(e) <= (int)annexb7::kTickMax)
				a.tick_n = (uint32_t)atoi(e);
		}
		{
			const char *e = getenv("H264GPU_SCAN7_NAP");
			a.nap_max = e != NULL && atoi(e) > 0 ? (uint32_t)atoi(e) : 512u;
		}
		if (!(ctx->attr_set & (1u << ROWS))) {
			cudaFuncSetAttribute(annexb7::scan7_kernel<ROWS, CTAS>, cudaFuncAttributePreferredSharedMemoryCarveout,
					     cudaSharedmemCarveoutMaxShared);
			ctx->attr_set |= 1u << ROWS;
		}
		const char *trace_path = getenv("H264GPU_SCAN_TRACE");
		if (trace_path != NULL) {
			/* diagnostics build of the same kernel: per-span phase timestamps to a file */
			uint64_t *d_trace = NULL;
			CU_TRY(cudaMalloc(&d_trace, nspans * 64));
			CU_TRY(cudaMemsetAsync(d_trace, 0, nspans * 64, st));
			a.trace = d_trace;
			cudaFuncSetAttribute(annexb7::scan7_kernel<ROWS, CTAS, true>,
					     cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
			annexb7::scan7_kernel<ROWS, CTAS, true><<<grid, annexb7::kT, 0, st>>>(a);
			uint64_t *h = (uint64_t *)malloc(nspans * 64);
			CU_TRY(cudaStreamSynchronize(st));
			CU_TRY(cudaMemcpy(h, d_trace, nspans * 64, cudaMemcpyDeviceToHost));
			FILE *fp = fopen(trace_path, "wb");
			if (fp != NULL) {
				fwrite(h, 64, nspans, fp);
				fclose(fp);
			}
			free(h);
			cudaFree(d_trace);
		} else {
			annexb7::scan7_kernel<ROWS, CTAS><<<grid, annexb7::kT, 0, st>>>(a);
		}
		if (K > 1) {
			/* second pass: the spans deferred at region starts (every chain word is there now) */
			annexb7::Scan7Args b = a;
			b.pass2 = 1;
			b.regions = 1;
			b.trace = NULL;
			const uint64_t est = (uint64_t)K * 4; /* CTAs worth launching: ~17 spans per region */
			const uint32_t g2 = (uint32_t)(est < grid ? est : grid);
			annexb7::scan7_kernel<ROWS, CTAS><<<g2 ? g2 : 1, annexb7::kT, 0, st>>>(b);
			ctx->launches++;
		}
	} else {
		const uint32_t grid = (uint32_t)(ctas_needed < (uint64_t)sms * 4 ? ctas_needed : (uint64_t)sms * 4);
		annexb7::scan7_only_kernel<ROWS, 4><<<grid, annexb7::kT, 0, st>>>(a);
	}
	CU_TRY(cudaGetLastError());

	annexb7::Fin7Args f;
	memset(&f, 0, sizeof(f));
	f.fin = a.fin;
	f.chain = a.chain;
	f.num_spans = (uint32_t)nspans;
	f.nblk = (uint32_t)nblk;
	f.evbuf = a.evbuf;
	f.ev_cap = ev_cap;
	f.ordered = (uint64_t *)(ws + ord_off);
	f.span_pre = (uint64_t *)(ws + pre_off);
	f.blk = (uint64_t *)(ws + blk_off);
	f.totals = (uint64_t *)(ws + tot_off);
	f.ctrl = a.ctrl;
	f.len = len;
	f.base = base;
	f.nal_start = d_nal_start;
	f.nal_end = d_nal_end;
	f.nal_rbsp = d_nal_rbsp;
	f.nal_rbsp_len = d_nal_rbsp_len;
	f.nal_cap = nal_cap;
	f.result = d_result;
	f.has_right = a.has_right;
	f.strip = d_rbsp != NULL ? 1 : 0;
	f.assume_in = (uint32_t)assume_in;
	annexb7::fin7_spans<<<(uint32_t)nblk, annexb7::kFinT, 0, st>>>(f);
	annexb7::fin7_order<<<(uint32_t)((nspans + 255) / 256), 256, 0, st>>>(f);
	/* grid-stride over the ordered events: any grid covers them all */
	uint64_t tb = (ev_cap + 255) / 256;
	if (tb > (uint64_t)sms * 8)
		tb = (uint64_t)sms * 8;
	annexb7::fin7_table<<<(uint32_t)(tb ? tb : 1), 256, 0, st>>>(f);
	CU_TRY(cudaGetLastError());
	ctx->launches += 4;
	return 0;
}

/* bytes of workspace a launch over len bytes with room for nal_cap NAL units needs (the layout
 * of scan7_launch_t) */
static size_t scan7_workspace_bytes(uint64_t len, uint64_t nal_cap)
{
	const uint64_t span = annexb7::Cfg<8>::SPAN;
	const uint64_t nspans = (len + 2 + span - 1) / span;
	uint64_t ev_cap = 4 * nal_cap + 4096;
	if (ev_cap > len / 3 + 2)
		ev_cap = len / 3 + 2;
	const uint64_t nblk = (nspans + annexb7::kFinT - 1) / annexb7::kFinT;
	return annexb7::kCtrlBytes + (size_t)nspans * 28 + (size_t)nblk * 16 + 64 + (size_t)ev_cap * 40;
}

extern "C" int h264gpu_scan_reserve(h264gpu_ctx *ctx, uint64_t len, uint64_t nal_cap)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	const size_t had = ctx->ws7_bytes;
	r = ws7_reserve(ctx, scan7_workspace_bytes(len, nal_cap));
	if (r >= 0 && ctx->ws7_bytes != had)
		ctx->epoch7 = 0;
	return r;
}

static int scan7_launch(h264gpu_ctx *ctx, const uint8_t *d_in, uint64_t len, uint64_t base,
			const struct h264gpu_shard_edge *edge, uint8_t *d_rbsp, uint64_t *d_nal_start,
			uint64_t *d_nal_end, uint64_t *d_nal_rbsp, uint64_t *d_nal_rbsp_len, uint64_t nal_cap,
			struct h264gpu_scan_result *d_result, cudaStream_t st)
{
	/* 8 rows = 4 KiB spans, 5 CTAs per SM.  3 KiB spans at 6 CTAs per SM (H264GPU_SCAN7_ROWS=6) were
	 * measured 11 % slower when all tickets came from one word (4/3 of the tickets at a fixed ticket
	 * rate); kept selectable for a measurement with the ticket counters */
	const char *e = getenv("H264GPU_SCAN7_ROWS");
	if (e != NULL && atoi(e) == 6)
		return scan7_launch_t<6, 6>(ctx, d_in, len, base, edge, d_rbsp, d_nal_start, d_nal_end, d_nal_rbsp,
					    d_nal_rbsp_len, nal_cap, d_result, st);
	return scan7_launch_t<8, 5>(ctx, d_in, len, base, edge, d_rbsp, d_nal_start, d_nal_end, d_nal_rbsp,
				    d_nal_rbsp_len, nal_cap, d_result, st);
}

/* ---- Annex-B (4-byte start codes) -> AVCC lengths, in place on the device ---------------- */

extern "C" int h264gpu_byte_stream_to_avcc_dev(h264gpu_ctx *ctx, uint8_t *d_data, uint64_t len, uint64_t max_nal,
						uint64_t *d_pos, uint64_t *d_count, void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (d_data == NULL || len == 0 || ((uintptr_t)d_data & 15))
		return -EINVAL;
	if (len >= (1ull << 40))
		return -E2BIG;
	constexpr int ROWS = 8;
	cudaStream_t st = (cudaStream_t)stream;
	const uint64_t span = (uint64_t)annexb7::Cfg<ROWS>::SPAN;
	const uint64_t nspans = (len + span - 1) / span;
	uint64_t ev_cap = max_nal ? max_nal : len / 4 + 2;
	if (ev_cap > len / 4 + 2)
		ev_cap = len / 4 + 2;
	if (ev_cap >= 0xffffffffull)
		ev_cap = 0xfffffffeull;
	const uint64_t nblk = (nspans + annexb7::kFinT - 1) / annexb7::kFinT;
	const size_t fin_off = (size_t)annexb7::kCtrlBytes;
	const size_t pre_off = fin_off + (size_t)nspans * 8;
	const size_t blk_off = pre_off + (size_t)nspans * 8;
	const size_t tot_off = blk_off + (size_t)nblk * 16;
	const size_t ev_off = tot_off + 64;
	const size_t ord_off = ev_off + (size_t)ev_cap * 16;
	const size_t need = ord_off + (size_t)ev_cap * 24;
	const size_t had = ctx->ws7_bytes;
	if ((r = ws7_reserve(ctx, need)) < 0)
		return r;
	if (ctx->ws7_bytes != had)
		ctx->epoch7 = 0;
	uint8_t *ws = (uint8_t *)ctx->ws7;
	annexb7::Scan7Args a;
	memset(&a, 0, sizeof(a));
	a.in = d_data;
	a.len = len;
	a.fin = (uint64_t *)(ws + fin_off);
	a.ctrl = (uint32_t *)ws;
	a.evbuf = (uint64_t *)(ws + ev_off);
	a.ev_cap = ev_cap;
	a.num_spans = (uint32_t)nspans;
	a.halo_left = 0xffffffffu;
	a.right[0] = a.right[1] = 0xff;
	const int sms = device_sms(ctx);
	const uint64_t ctas = (nspans + annexb7::kW - 1) / annexb7::kW;
	const uint32_t grid = (uint32_t)(ctas < (uint64_t)sms * 4 ? ctas : (uint64_t)sms * 4);
	annexb7::scan7_only_kernel<ROWS, 4, true><<<grid, annexb7::kT, 0, st>>>(a);
	annexb7::Fin7Args f;
	memset(&f, 0, sizeof(f));
	f.fin = a.fin;
	f.num_spans = (uint32_t)nspans;
	f.nblk = (uint32_t)nblk;
	f.evbuf = a.evbuf;
	f.ev_cap = ev_cap;
	f.ordered = (uint64_t *)(ws + ord_off);
	f.span_pre = (uint64_t *)(ws + pre_off);
	f.blk = (uint64_t *)(ws + blk_off);
	f.totals = (uint64_t *)(ws + tot_off);
	f.ctrl = a.ctrl;
	annexb7::fin7_spans<<<(uint32_t)nblk, annexb7::kFinT, 0, st>>>(f);
	annexb7::fin7_order<<<(uint32_t)((nspans + 255) / 256), 256, 0, st>>>(f);
	uint64_t tb = (ev_cap + 255) / 256;
	if (tb > (uint64_t)sms * 8)
		tb = (uint64_t)sms * 8;
	annexb7::avcc_write_kernel<<<(uint32_t)(tb ? tb : 1), 256, 0, st>>>(f.ordered, f.totals, ev_cap, d_data, len, d_pos,
									   d_count);
	annexb7::avcc_rearm_kernel<<<1, 1, 0, st>>>(a.ctrl);
	CU_TRY(cudaGetLastError());
	ctx->launches += 5;
	return 0;
}

extern "C" int h264gpu_byte_stream_to_avcc_host(h264gpu_ctx *ctx, uint8_t *h_data, uint64_t len, uint64_t *n_nal)
{
	if (n_nal)
		*n_nal = 0;
	if (h_data == NULL || len == 0)
		return -EINVAL;
	int r = h264gpu_reader_upload(ctx, h_data, len);
	if (r < 0)
		return r;
	cudaStream_t st = ctx->s_rd;
	/* only the 4-byte length fields change: the positions come back and the host patches them */
	uint64_t cap = len / 64 + 4096;
	for (int pass = 0; pass < 2; pass++) {
		if ((r = h264gpu_pool_dev(ctx, &ctx->rd_tab, (cap + 1) * 8)) < 0 ||
		    (r = h264gpu_pool_host(ctx, &ctx->rh_tab, (cap + 1) * 8)) < 0)
			return r;
		uint64_t *d_pos = (uint64_t *)ctx->rd_tab.p, *h_pos = (uint64_t *)ctx->rh_tab.p;
		r = h264gpu_byte_stream_to_avcc_dev(ctx, (uint8_t *)ctx->rd_stream.p, len, cap, d_pos + 1, d_pos, st);
		if (r < 0)
			return r;
		CU_TRY(cudaMemcpyAsync(h_pos, d_pos, 8, cudaMemcpyDeviceToHost, st));
		CU_TRY(cudaStreamSynchronize(st));
		const uint64_t n = h_pos[0];
		if (n > cap) {
			cap = len / 4 + 2;
			/* the device copy was patched with an incomplete table: start over */
			if ((r = h264gpu_reader_upload(ctx, h_data, len)) < 0)
				return r;
			continue;
		}
		if (n) {
			CU_TRY(cudaMemcpyAsync(h_pos + 1, d_pos + 1, n * 8, cudaMemcpyDeviceToHost, st));
			CU_TRY(cudaStreamSynchronize(st));
		}
		for (uint64_t e = 0; e < n; e++) {
			const uint64_t p = h_pos[1 + e];
			const uint64_t nxt = e + 1 < n ? h_pos[2 + e] : len;
			if (len - p <= 4)
				continue;
			const uint32_t nal = (uint32_t)(nxt - p - 4);
			h_data[p] = (uint8_t)(nal >> 24);
			h_data[p + 1] = (uint8_t)(nal >> 16);
			h_data[p + 2] = (uint8_t)(nal >> 8);
			h_data[p + 3] = (uint8_t)nal;
		}
		if (n_nal)
			*n_nal = n;
		return 0;
	}
	return -E2BIG;
}

/* ---- scan + strip, RBSP in place: C-ABI entry -------------------------------------- */

extern "C" int h264gpu_split_strip_inplace_dev(h264gpu_ctx *ctx, const uint8_t *d_in, uint64_t len,
					       uint64_t base, const struct h264gpu_shard_edge *edge,
					       uint8_t *d_rbsp, uint64_t *d_nal_start, uint64_t *d_nal_end,
					       uint64_t *d_nal_rbsp, uint64_t *d_nal_rbsp_len,
					       uint64_t nal_cap, struct h264gpu_scan_result *d_result,
					       void *stream)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (d_in == NULL || d_result == NULL || (nal_cap && (!d_nal_start || !d_nal_end)))
		return -EINVAL;
	if (((uintptr_t)d_in & 15) || ((uintptr_t)d_rbsp & 15))
		return -EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	if (len == 0) {
		CU_TRY(cudaMemsetAsync(d_result, 0, sizeof(*d_result), st));
		return 0;
	}
	if (len >= (1ull << 40))
		return -E2BIG;
	return scan7_launch(ctx, d_in, len, base, edge, d_rbsp, d_nal_start, d_nal_end, d_nal_rbsp, d_nal_rbsp_len,
			    nal_cap, d_result, st);
}

/* ---- reader session: what h264_reader_parse drives -------------------------------- */

int h264gpu_pool_dev(h264gpu_ctx *ctx, struct h264gpu_ctx::h264gpu_pool *pl, size_t bytes)
{
	if (bytes <= pl->cap)
		return 0;
	if (ctx->s_rd)
		CU_TRY(cudaStreamSynchronize(ctx->s_rd));
	if (pl->p)
		CU_TRY(cudaFree(pl->p));
	pl->p = NULL;
	pl->cap = 0;
	const size_t want = ((bytes + (bytes >> 2)) + 4095) & ~(size_t)4095;
	CU_TRY(cudaMalloc(&pl->p, want));
	pl->cap = want;
	return 0;
}

int h264gpu_pool_host(h264gpu_ctx *ctx, struct h264gpu_ctx::h264gpu_pool *pl, size_t bytes)
{
	if (bytes <= pl->cap)
		return 0;
	if (ctx->s_rd)
		CU_TRY(cudaStreamSynchronize(ctx->s_rd));
	if (pl->p)
		CU_TRY(cudaFreeHost(pl->p));
	pl->p = NULL;
	pl->cap = 0;
	const size_t want = ((bytes + (bytes >> 2)) + 4095) & ~(size_t)4095;
	CU_TRY(cudaMallocHost(&pl->p, want));
	pl->cap = want;
	return 0;
}

int h264gpu_reader_stream(h264gpu_ctx *ctx, cudaStream_t *st)
{
	if (ctx->s_rd == NULL)
		CU_TRY(cudaStreamCreateWithFlags(&ctx->s_rd, cudaStreamNonBlocking));
	*st = ctx->s_rd;
	return 0;
}

extern "C" int h264gpu_reader_upload(h264gpu_ctx *ctx, const uint8_t *h_buf, uint64_t len)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (h_buf == NULL && len)
		return -EINVAL;
	cudaStream_t st;
	if ((r = h264gpu_reader_stream(ctx, &st)) < 0)
		return r;
	if ((r = h264gpu_pool_dev(ctx, &ctx->rd_stream, len + 64)) < 0)
		return r;
	ctx->rd_stream_len = 0;
	/* H264GPU_REGISTER_INPUT=1: page-lock the caller's buffer for the copy (pays off for large
	 * buffers that are parsed more than once; a pageable copy is staged by the driver) */
	static int reg = -1;
	if (reg < 0) {
		const char *e = getenv("H264GPU_REGISTER_INPUT");
		reg = e != NULL && atoi(e) > 0;
	}
	bool registered = false;
	if (reg && len >= (1u << 20))
		registered = cudaHostRegister((void *)h_buf, len, cudaHostRegisterReadOnly) == cudaSuccess ||
			     cudaHostRegister((void *)h_buf, len, cudaHostRegisterDefault) == cudaSuccess;
	if (!registered)
		(void)cudaGetLastError();
	cudaError_t ce = len ? cudaMemcpyAsync(ctx->rd_stream.p, h_buf, len, cudaMemcpyHostToDevice, st) : cudaSuccess;
	/* the kernels may read a few bytes past the end of a NAL: keep them defined */
	if (ce == cudaSuccess)
		ce = cudaMemsetAsync((uint8_t *)ctx->rd_stream.p + len, 0xff, 64, st);
	if (registered) {
		cudaStreamSynchronize(st);
		cudaHostUnregister((void *)h_buf);
	}
	CU_TRY(ce);
	ctx->rd_stream_len = len;
	return 0;
}

extern "C" int h264gpu_reader_scan(h264gpu_ctx *ctx, const uint8_t *h_buf, uint64_t len,
				   const uint64_t **h_start, const uint64_t **h_end, uint64_t *n_nal,
				   uint64_t *final_off)
{
	if (h_start == NULL || h_end == NULL || n_nal == NULL)
		return -EINVAL;
	*h_start = *h_end = NULL;
	*n_nal = 0;
	if (final_off)
		*final_off = 0;
	int r = h264gpu_reader_upload(ctx, h_buf, len);
	if (r < 0 || len == 0)
		return r;
	cudaStream_t st = ctx->s_rd;
	/* table capacity: a NAL per 256 bytes is generous for video; the scan reports how many it
	 * found, and a second launch with the exact size follows if that was not enough */
	uint64_t cap = len / 256 + 1024;
	for (int pass = 0; pass < 2; pass++) {
		if ((r = h264gpu_pool_dev(ctx, &ctx->rd_tab, cap * 16)) < 0 ||
		    (r = h264gpu_pool_dev(ctx, &ctx->rd_res, sizeof(struct h264gpu_scan_result))) < 0 ||
		    (r = h264gpu_pool_host(ctx, &ctx->rh_res, sizeof(struct h264gpu_scan_result))) < 0)
			return r;
		uint64_t *d_st = (uint64_t *)ctx->rd_tab.p, *d_en = d_st + cap;
		struct h264gpu_scan_result *d_res = (struct h264gpu_scan_result *)ctx->rd_res.p;
		struct h264gpu_scan_result *h_res = (struct h264gpu_scan_result *)ctx->rh_res.p;
		r = h264gpu_split_strip_inplace_dev(ctx, (const uint8_t *)ctx->rd_stream.p, len, 0, NULL, NULL, d_st,
						    d_en, NULL, NULL, cap, d_res, st);
		if (r < 0)
			return r;
		CU_TRY(cudaMemcpyAsync(h_res, d_res, sizeof(*h_res), cudaMemcpyDeviceToHost, st));
		CU_TRY(cudaStreamSynchronize(st));
		if (h_res->n_nal <= cap && !h_res->reserved) {
			const uint64_t n = h_res->n_nal;
			if ((r = h264gpu_pool_host(ctx, &ctx->rh_tab, (n + 1) * 16)) < 0)
				return r;
			uint64_t *hs = (uint64_t *)ctx->rh_tab.p, *he = hs + n + 1;
			if (n) {
				CU_TRY(cudaMemcpyAsync(hs, d_st, n * 8, cudaMemcpyDeviceToHost, st));
				CU_TRY(cudaMemcpyAsync(he, d_en, n * 8, cudaMemcpyDeviceToHost, st));
				CU_TRY(cudaStreamSynchronize(st));
			}
			*h_start = hs;
			*h_end = he;
			*n_nal = n;
			/* what h264_reader_parse leaves in *off (src/h264_reader.c:139): the end of the last
			 * NAL, which is `len` when it ran to the end of the buffer */
			if (final_off)
				*final_off = n ? he[n - 1] : 0;
			return 0;
		}
		cap = h_res->reserved ? len / 3 + 2 : h_res->n_nal + 16;
	}
	return -E2BIG;
}

extern "C" int h264gpu_reader_resident(h264gpu_ctx *ctx, const uint8_t **d_stream, uint64_t *len)
{
	if (ctx == NULL || d_stream == NULL || len == NULL)
		return -EINVAL;
	*d_stream = (const uint8_t *)ctx->rd_stream.p;
	*len = ctx->rd_stream_len;
	return 0;
}

/* ---- shard merge (host) --------------------------------------------------- */

extern "C" void h264gpu_merge_init(struct h264gpu_merge *m)
{
	memset(m, 0, sizeof(*m));
}

extern "C" int h264gpu_merge_shard(struct h264gpu_merge *m, const struct h264gpu_scan_result *r,
				   uint64_t *tab_start, uint64_t *tab_end, uint64_t *tab_rbsp,
				   uint64_t tab_cap, uint64_t shard_nals_copied,
				   uint64_t *rbsp_skip, uint64_t *rbsp_take)
{
	if (m == NULL || r == NULL)
		return -EINVAL;
	/* Was the byte before this shard inside a NAL?  The shard assumed "yes"
	 * for every shard but the first; undo that if it was wrong. */
	const int assumed_in = m->shards > 0;
	const int actual_in = m->shards > 0 && m->open;
	uint64_t skip = (assumed_in && !actual_in) ? r->head_bytes : 0;
	uint64_t take = r->rbsp_bytes - skip;
	/* the open NAL of earlier shards ends at this shard's first event */
	if (m->open && r->any_event && m->n_nal >= 1 && m->n_nal - 1 < tab_cap)
		tab_end[m->n_nal - 1] = r->first_event_pos;
	/* rebase the RBSP offsets of the entries just appended at [n_nal, n_nal+copied) */
	if (tab_rbsp != NULL) {
		for (uint64_t k = 0; k < shard_nals_copied; k++)
			tab_rbsp[m->n_nal + k] = tab_rbsp[m->n_nal + k] - skip + m->rbsp_bytes;
	}
	(void)tab_start;
	m->n_nal += r->n_nal;
	m->rbsp_bytes += take;
	if (r->any_event)
		m->open = r->end_open ? 1 : 0;
	m->shards++;
	if (rbsp_skip)
		*rbsp_skip = skip;
	if (rbsp_take)
		*rbsp_take = take;
	return 0;
}

extern "C" int h264gpu_merge_shard_inplace(struct h264gpu_merge *m, const struct h264gpu_scan_result *r,
					   uint64_t *tab_end, uint64_t *tab_rbsp_len, uint64_t tab_cap,
					   uint64_t *carry_len)
{
	if (m == NULL || r == NULL)
		return -EINVAL;
	/* every shard but the first reported the bytes before its first event as the head of an
	 * open NAL; they count only if a NAL really is open */
	const int assumed_in = m->shards > 0;
	const int actual_in = m->shards > 0 && m->open;
	const uint64_t head = assumed_in ? r->head_bytes : 0;
	const uint64_t carry = actual_in ? head : 0;
	if (m->open && m->n_nal >= 1 && m->n_nal - 1 < tab_cap) {
		if (tab_rbsp_len != NULL)
			tab_rbsp_len[m->n_nal - 1] += carry;
		if (r->any_event)
			tab_end[m->n_nal - 1] = r->first_event_pos;
	}
	m->n_nal += r->n_nal;
	m->rbsp_bytes += r->rbsp_bytes - (head - carry);
	if (r->any_event)
		m->open = r->end_open ? 1 : 0;
	m->shards++;
	if (carry_len)
		*carry_len = carry;
	return 0;
}

extern "C" int h264gpu_merge_finish(struct h264gpu_merge *m, uint64_t stream_len,
				    uint64_t *tab_end, uint64_t tab_cap, uint64_t *final_off)
{
	if (m == NULL)
		return -EINVAL;
	uint64_t off = 0;
	if (m->n_nal >= 1) {
		if (m->open) {
			if (m->n_nal - 1 < tab_cap)
				tab_end[m->n_nal - 1] = stream_len;
			off = stream_len;
		} else if (m->n_nal - 1 < tab_cap) {
			off = tab_end[m->n_nal - 1];
		}
	}
	if (final_off)
		*final_off = off;
	return 0;
}

/* ---- host-buffer pipeline -------------------------------------------------- */

static int pipeline_init(h264gpu_ctx *ctx)
{
	if (ctx->s_in != NULL)
		return 0;
	CU_TRY(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
	CU_TRY(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
	CU_TRY(cudaStreamCreateWithFlags(&ctx->s_tab, cudaStreamNonBlocking));
	CU_TRY(cudaStreamCreateWithFlags(&ctx->s_up, cudaStreamNonBlocking));
	if (ctx->tab_cap == 0)
		ctx->tab_cap = ctx->chunk_bytes / 64 > 4096 ? ctx->chunk_bytes / 64 : 4096;
	for (int b = 0; b < 2; b++) {
		CU_TRY(cudaMalloc(&ctx->d_chunk_in[b], ctx->chunk_bytes + 16));
		CU_TRY(cudaMalloc(&ctx->d_chunk_out[b], ctx->chunk_bytes + 16));
		CU_TRY(cudaMalloc(&ctx->d_tab[b], ctx->tab_cap * 3 * sizeof(uint64_t)));
		CU_TRY(cudaMalloc(&ctx->d_res[b], sizeof(struct h264gpu_scan_result)));
		CU_TRY(cudaEventCreateWithFlags(&ctx->ev_in[b], cudaEventDisableTiming));
		CU_TRY(cudaEventCreateWithFlags(&ctx->ev_k[b], cudaEventDisableTiming));
		CU_TRY(cudaEventCreateWithFlags(&ctx->ev_out[b], cudaEventDisableTiming));
	}
	CU_TRY(cudaMallocHost(&ctx->h_res, 2 * sizeof(struct h264gpu_scan_result)));
	return 0;
}

static int split_strip_host_impl(h264gpu_ctx *ctx, const uint8_t *h_in, uint64_t len, uint8_t *h_rbsp,
				 uint64_t *h_nal_start, uint64_t *h_nal_end, uint64_t *h_nal_rbsp, uint64_t *n_nal,
				 uint64_t *rbsp_bytes, uint64_t *final_off);

extern "C" int h264gpu_split_strip_host(h264gpu_ctx *ctx, const uint8_t *h_in, uint64_t len,
					uint8_t *h_rbsp, uint64_t *h_nal_start,
					uint64_t *h_nal_end, uint64_t *h_nal_rbsp, uint64_t *n_nal,
					uint64_t *rbsp_bytes, uint64_t *final_off)
{
	const int r = split_strip_host_impl(ctx, h_in, len, h_rbsp, h_nal_start, h_nal_end, h_nal_rbsp, n_nal, rbsp_bytes,
					    final_off);
	if (r < 0 && r != -ENOBUFS && ctx != NULL && ctx->s_in != NULL) {
		/* an early return: copies from / into the caller's buffers may still be in flight */
		cudaStreamSynchronize(ctx->s_up);
		cudaStreamSynchronize(ctx->s_in);
		cudaStreamSynchronize(ctx->s_tab);
		cudaStreamSynchronize(ctx->s_out);
	}
	return r;
}

static int split_strip_host_impl(h264gpu_ctx *ctx, const uint8_t *h_in, uint64_t len, uint8_t *h_rbsp,
				 uint64_t *h_nal_start, uint64_t *h_nal_end, uint64_t *h_nal_rbsp, uint64_t *n_nal,
				 uint64_t *rbsp_bytes, uint64_t *final_off)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (h_in == NULL || n_nal == NULL || (*n_nal && (!h_nal_start || !h_nal_end)))
		return -EINVAL;
	const uint64_t cap = *n_nal;
	*n_nal = 0;
	if (rbsp_bytes)
		*rbsp_bytes = 0;
	if (final_off)
		*final_off = 0;
	if (len == 0)
		return 0;
	r = pipeline_init(ctx);
	if (r < 0)
		return r;

	const uint64_t chunk = ctx->chunk_bytes;
	const uint64_t nchunks = (len + chunk - 1) / chunk;
	struct h264gpu_merge mg;
	h264gpu_merge_init(&mg);
	int overflow = 0;
	bool used[2] = {false, false};

	/* software pipeline: upload+kernel of chunk c overlaps download of chunk c-1 */
	for (uint64_t c = 0; c <= nchunks; c++) {
		if (c < nchunks) {
			const int b = (int)(c & 1);
			const uint64_t lo = c * chunk;
			const uint64_t n = len - lo < chunk ? len - lo : chunk;
			struct h264gpu_shard_edge e;
			memset(&e, 0, sizeof(e));
			if (lo >= 2) {
				e.has_left = 1;
				e.left[0] = h_in[lo - 2];
				e.left[1] = h_in[lo - 1];
			}
			if (lo + n < len) {
				e.has_right = 1;
				e.right[0] = h_in[lo + n];
				e.right[1] = lo + n + 1 < len ? h_in[lo + n + 1] : 0xff;
			}
			e.assume_in = c > 0;
			/* the upload has a stream of its own and waits only for the kernel that read this slot's
			 * input two chunks ago; the kernel waits for the upload and for the download that empties
			 * the slot's output.  (Round 1 had upload and kernel on one stream behind the download:
			 * every upload started a kernel + table copy + merge late, 85 % of the copy ceiling.) */
			if (used[b])
				CU_TRY(cudaStreamWaitEvent(ctx->s_up, ctx->ev_k[b], 0));
			CU_TRY(cudaMemcpyAsync(ctx->d_chunk_in[b], h_in + lo, n, cudaMemcpyHostToDevice,
					       ctx->s_up));
			CU_TRY(cudaEventRecord(ctx->ev_in[b], ctx->s_up));
			CU_TRY(cudaStreamWaitEvent(ctx->s_in, ctx->ev_in[b], 0));
			if (used[b]) /* chunk c-2's download must have left the slot's output */
				CU_TRY(cudaStreamWaitEvent(ctx->s_in, ctx->ev_out[b], 0));
			r = h264gpu_split_strip_dev(ctx, ctx->d_chunk_in[b], n, lo, &e,
						    h_rbsp ? ctx->d_chunk_out[b] : NULL, ctx->d_tab[b],
						    ctx->d_tab[b] + ctx->tab_cap,
						    ctx->d_tab[b] + 2 * ctx->tab_cap, ctx->tab_cap,
						    ctx->d_res[b], ctx->s_in);
			if (r < 0)
				return r;
			CU_TRY(cudaMemcpyAsync(&ctx->h_res[b], ctx->d_res[b], sizeof(ctx->h_res[b]),
					       cudaMemcpyDeviceToHost, ctx->s_in));
			CU_TRY(cudaEventRecord(ctx->ev_k[b], ctx->s_in));
			used[b] = true;
		}
		if (c >= 1) {
			const int b = (int)((c - 1) & 1);
			CU_TRY(cudaEventSynchronize(ctx->ev_k[b]));
			const struct h264gpu_scan_result res = ctx->h_res[b];
			if (res.n_nal > ctx->tab_cap) {
				/* per-chunk table too small: extremely NAL-dense input */
				fprintf(stderr, "h264gpu: chunk holds %llu NALs > table %zu; set "
						"H264GPU_CHUNK_MB lower\n",
					(unsigned long long)res.n_nal, ctx->tab_cap);
				return -E2BIG;
			}
			uint64_t room = mg.n_nal < cap ? cap - mg.n_nal : 0;
			uint64_t ncopy = res.n_nal < room ? res.n_nal : room;
			if (ncopy < res.n_nal)
				overflow = 1;
			if (ncopy) {
				CU_TRY(cudaMemcpyAsync(h_nal_start + mg.n_nal, ctx->d_tab[b],
						       ncopy * 8, cudaMemcpyDeviceToHost, ctx->s_tab));
				CU_TRY(cudaMemcpyAsync(h_nal_end + mg.n_nal, ctx->d_tab[b] + ctx->tab_cap,
						       ncopy * 8, cudaMemcpyDeviceToHost, ctx->s_tab));
				if (h_nal_rbsp)
					CU_TRY(cudaMemcpyAsync(h_nal_rbsp + mg.n_nal,
							       ctx->d_tab[b] + 2 * ctx->tab_cap, ncopy * 8,
							       cudaMemcpyDeviceToHost, ctx->s_tab));
			}
			CU_TRY(cudaStreamSynchronize(ctx->s_tab)); /* tables of this chunk are home */
			const uint64_t dst = mg.rbsp_bytes;
			uint64_t skip = 0, take = 0;
			h264gpu_merge_shard(&mg, &res, h_nal_start, h_nal_end,
					    h_rbsp ? h_nal_rbsp : NULL, cap, ncopy, &skip, &take);
			if (h_rbsp && take)
				CU_TRY(cudaMemcpyAsync(h_rbsp + dst, ctx->d_chunk_out[b] + skip, take,
						       cudaMemcpyDeviceToHost, ctx->s_out));
			CU_TRY(cudaEventRecord(ctx->ev_out[b], ctx->s_out));
		}
	}
	CU_TRY(cudaStreamSynchronize(ctx->s_out));
	uint64_t off = 0;
	h264gpu_merge_finish(&mg, len, h_nal_end, cap, &off);
	*n_nal = mg.n_nal;
	if (rbsp_bytes)
		*rbsp_bytes = mg.rbsp_bytes;
	if (final_off)
		*final_off = off;
	return overflow ? -ENOBUFS : 0;
}

/* ---- writer side: EPB insert + framing --------------------------------------- */

#include "annexb_frame.cuh"
#include "annexb_frame6.cuh"
#include "annexb_frame7.cuh"
#include "annexb_frame8.cuh"

/* writer kernel generation: 6 (block-wide 32 KiB tiles, annexb_frame6.cuh; the default) or 7
 * (warp-autonomous spans, annexb_frame7.cuh; the same speed on B200, see its header):
 * H264GPU_FRAME_GEN */
static int frame_gen(void)
{
	const char *e = getenv("H264GPU_FRAME_GEN");
	return (e != NULL && (atoi(e) == 7 || atoi(e) == 8)) ? atoi(e) : 6;
}

template <int ROWS, int NW, int NBUF>
static cudaError_t launch_frame7_cfg(const frame::FrameArgs &a, cudaStream_t st, int sms)
{
	/* resident CTAs per SM: what 227 KB of shared memory hold (1 KB per CTA is the system's) */
	constexpr int per_cta = NW * (int)sizeof(frame7::WSmem<ROWS, NBUF>) + 1024;
	constexpr int fit = (227 * 1024) / per_cta;
	constexpr int MINB = fit * NW > 48 ? 48 / NW : (fit > 32 ? 32 : fit);
	cudaFuncSetAttribute(frame7::frame7_kernel<ROWS, NW, MINB, NBUF>, cudaFuncAttributePreferredSharedMemoryCarveout,
			     cudaSharedmemCarveoutMaxShared);
	/* persistent warps: as many CTAs as fit, fewer when there are fewer spans than warps */
	const uint64_t cap = (uint64_t)sms * MINB;
	const uint64_t want = ((uint64_t)a.num_tiles + NW - 1) / NW;
	frame7::frame7_kernel<ROWS, NW, MINB, NBUF><<<(unsigned)(want < cap ? want : cap), 32 * NW, 0, st>>>(a);
	return cudaGetLastError();
}

template <int ROWS>
static cudaError_t launch_frame7(const frame::FrameArgs &a, cudaStream_t st)
{
	const uint32_t pthreads = 128;
	const uint32_t pblocks = (a.num_tiles + 1 + pthreads - 1) / pthreads;
	frame7::frame7_prepass<ROWS><<<pblocks, pthreads, 0, st>>>(a);
	int sms = 0, dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (sms <= 0)
		sms = 148;
	/* H264GPU_FRAME7_NBUF = 2 (default): a span's bytes stay staged from classification to
	 * emission, one warp per CTA, 20 warps per SM; 1: staged twice, four warps per CTA, 32 warps per
	 * SM (measured slower: 1064 vs 1207 GB/s, the instruction cache misses of 32 unrelated warps) */
	const char *e = getenv("H264GPU_FRAME7_NBUF");
	if (e != NULL && atoi(e) == 1)
		return launch_frame7_cfg<ROWS, 4, 1>(a, st, sms);
	return launch_frame7_cfg<ROWS, 1, 2>(a, st, sms);
}

/* gen 8: the tiles of gen 6, counts published an iteration ahead over the chain of gen 7 */
template <int ROWS>
static cudaError_t launch_frame8(const frame::FrameArgs &a, cudaStream_t st)
{
	const uint32_t pthreads = 128;
	const uint32_t pblocks = (a.num_tiles + 1 + pthreads - 1) / pthreads;
	frame8::frame8_prepass<ROWS><<<pblocks, pthreads, 0, st>>>(a);
	int sms = 0, dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (sms <= 0)
		sms = 148;
	constexpr int per_cta = (int)sizeof(frame8::Smem8<ROWS>) + 1024;
	constexpr int fit = (227 * 1024) / per_cta;
	constexpr int MINB = fit > 8 ? 8 : fit;
	cudaFuncSetAttribute(frame8::frame8_kernel<ROWS, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout,
			     cudaSharedmemCarveoutMaxShared);
	const uint32_t cap = (uint32_t)sms * MINB;
	frame8::frame8_kernel<ROWS, MINB><<<a.num_tiles < cap ? a.num_tiles : cap, frame6::kT, 0, st>>>(a);
	return cudaGetLastError();
}

template <int ROWS>
static cudaError_t launch_frame6(const frame::FrameArgs &a, cudaStream_t st)
{
	const uint32_t pthreads = 128;
	const uint32_t pblocks = (a.num_tiles + 1 + pthreads - 1) / pthreads;
	frame::frame_prepass<ROWS><<<pblocks, pthreads, 0, st>>>(a);
	int sms = 0, dev = 0;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	if (sms <= 0)
		sms = 148;
	/* resident CTAs per SM the 32 KiB-tile kernel is compiled for: 5 (48 registers, the default) or
	 * 4 (64 registers); H264GPU_FRAME_CTAS */
	static int minb = 0;
	if (minb == 0) {
		const char *e = getenv("H264GPU_FRAME_CTAS");
		minb = (e != NULL && atoi(e) == 4) ? 4 : 5;
	}
	if (ROWS == 8 && minb == 5) {
		cudaFuncSetAttribute(frame6::frame6_kernel<ROWS, 5>, cudaFuncAttributePreferredSharedMemoryCarveout,
				     cudaSharedmemCarveoutMaxShared);
		const uint32_t cap = (uint32_t)sms * 5;
		frame6::frame6_kernel<ROWS, 5><<<a.num_tiles < cap ? a.num_tiles : cap, frame6::kT, 0, st>>>(a);
	} else {
		cudaFuncSetAttribute(frame6::frame6_kernel<ROWS, 4>, cudaFuncAttributePreferredSharedMemoryCarveout,
				     cudaSharedmemCarveoutMaxShared);
		const uint32_t cap = (uint32_t)sms * 4;
		frame6::frame6_kernel<ROWS, 4><<<a.num_tiles < cap ? a.num_tiles : cap, frame6::kT, 0, st>>>(a);
	}
	return cudaGetLastError();
}

/* len_known: the host already knows off[n] (the host forms); otherwise it is read back */
static int frame_dev_impl(h264gpu_ctx *ctx, const uint8_t *d_rbsp, const uint64_t *d_off, uint64_t n, int sc_len,
			  uint8_t *d_out, uint64_t out_cap, uint64_t *d_out_off, uint64_t *d_total, void *stream,
			  bool len_known, uint64_t known_len)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (d_off == NULL || d_out == NULL || d_out_off == NULL || d_total == NULL)
		return -EINVAL;
	if (sc_len != 0 && sc_len != 3 && sc_len != 4)
		return -EINVAL;
	if (((uintptr_t)d_rbsp & 15) || ((uintptr_t)d_out & 15))
		return -EINVAL;
	cudaStream_t st = (cudaStream_t)stream;
	/* total payload length = off[n]: needed on the host to size the grid */
	uint64_t len = known_len;
	if (!len_known) {
		CU_TRY(cudaMemcpyAsync(&len, d_off + n, sizeof(len), cudaMemcpyDeviceToHost, st));
		CU_TRY(cudaStreamSynchronize(st));
	}
	if (len && d_rbsp == NULL)
		return -EINVAL;
	const int items = ctx->scan_items;
	const uint64_t tile = (uint64_t)annexb::kBlock * items * 16;
	uint64_t ntiles = (len + tile - 1) / tile;
	if (ntiles == 0)
		ntiles = 1;
	if (ntiles > 0x7fffffffull)
		return -E2BIG;
	/* workspace: [256 control][desc: ntiles u64][first: ntiles+1 u64][tail: ntiles u32]
	 * [group_w][super_w][super_p] (gen 7: ntiles = spans of 512 x items bytes) */
	const int gen = frame_gen();
	int rows = items; /* 512-byte rows per span of the gen-7 kernel: H264GPU_FRAME7_ROWS = 1, 2, 4, 6 or 8 */
	{
		const char *e = getenv("H264GPU_FRAME7_ROWS");
		const int v = e != NULL ? atoi(e) : 0;
		if (v == 1 || v == 2 || v == 4 || v == 6 || v == 8)
			rows = v;
	}
	if (gen == 7) {
		const uint64_t span = (uint64_t)512 * rows;
		ntiles = (len + span - 1) / span;
		if (ntiles == 0)
			ntiles = 1;
		if (ntiles > 0x7fffffffull)
			return -E2BIG;
	}
	const size_t desc_bytes = 256 + (size_t)ntiles * 8;
	const size_t first_off = (desc_bytes + 15) & ~(size_t)15;
	const size_t tail_off = first_off + ((size_t)ntiles + 1) * 8;
	const size_t group_off = (tail_off + (size_t)ntiles * 4 + 15) & ~(size_t)15;
	const size_t ngroup = (size_t)(ntiles >> 5) + 1, nsuper = (size_t)(ntiles >> 10) + 1;
	const size_t need = group_off + (ngroup + 2 * nsuper) * 8;
	r = h264gpu_ws_reserve(ctx, need);
	if (r < 0)
		return r;
	if (gen == 6) /* gen 7 / 8: the pre-pass clears the chain words */
		CU_TRY(cudaMemsetAsync(ctx->ws, 0xff, desc_bytes, st));

	frame::FrameArgs a;
	memset(&a, 0, sizeof(a));
	a.rbsp = d_rbsp;
	a.len = len;
	a.off = d_off;
	a.n = n;
	a.sc_len = (uint32_t)sc_len;
	a.out = d_out;
	a.out_cap = out_cap;
	a.out_off = d_out_off;
	a.total = d_total;
	a.ticket = (uint32_t *)ctx->ws;
	a.desc = (uint64_t *)((uint8_t *)ctx->ws + 256);
	a.first = (uint64_t *)((uint8_t *)ctx->ws + first_off);
	a.tail = (uint32_t *)((uint8_t *)ctx->ws + tail_off);
	a.num_tiles = (uint32_t)ntiles;
	a.group_w = (uint64_t *)((uint8_t *)ctx->ws + group_off);
	a.super_w = a.group_w + ngroup;
	a.super_p = a.super_w + nsuper;
	/* diagnostics: phase timestamps of every span of the gen-7 kernel, written to the named file */
	const char *trace_path = gen == 7 ? getenv("H264GPU_FRAME_TRACE") : NULL;
	uint64_t *d_trace = NULL;
	if (trace_path != NULL && trace_path[0] != 0) {
		CU_TRY(cudaMalloc(&d_trace, (size_t)ntiles * 64));
		CU_TRY(cudaMemsetAsync(d_trace, 0, (size_t)ntiles * 64, st));
		a.trace = d_trace;
	}
	cudaError_t ce;
	if (gen == 7)
		ce = rows == 1 ? launch_frame7<1>(a, st)
		   : rows == 2 ? launch_frame7<2>(a, st)
		   : rows == 4 ? launch_frame7<4>(a, st)
		   : rows == 6 ? launch_frame7<6>(a, st)
			       : launch_frame7<8>(a, st);
	else if (gen == 8)
		ce = items == 1 ? launch_frame8<1>(a, st)
		   : items == 2 ? launch_frame8<2>(a, st)
		   : items == 4 ? launch_frame8<4>(a, st)
				: launch_frame8<8>(a, st);
	else
		ce = items == 1 ? launch_frame6<1>(a, st)
		   : items == 2 ? launch_frame6<2>(a, st)
		   : items == 4 ? launch_frame6<4>(a, st)
				: launch_frame6<8>(a, st);
	CU_TRY(ce);
	if (d_trace != NULL) {
		uint64_t *h = (uint64_t *)malloc((size_t)ntiles * 64);
		if (h != NULL && cudaMemcpyAsync(h, d_trace, (size_t)ntiles * 64, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
		    cudaStreamSynchronize(st) == cudaSuccess) {
			FILE *f = fopen(trace_path, "wb");
			if (f != NULL) {
				fwrite(h, 64, ntiles, f);
				fclose(f);
			}
		}
		free(h);
		cudaFree(d_trace);
	}
	ctx->launches += 2;
	return 0;
}

extern "C" int h264gpu_frame_dev(h264gpu_ctx *ctx, const uint8_t *d_rbsp, const uint64_t *d_off,
				 uint64_t n, int sc_len, uint8_t *d_out, uint64_t out_cap,
				 uint64_t *d_out_off, uint64_t *d_total, void *stream)
{
	return frame_dev_impl(ctx, d_rbsp, d_off, n, sc_len, d_out, out_cap, d_out_off, d_total, stream, false, 0);
}

/*
 * Host buffers, long inputs: chunks of whole payloads (>= chunk_bytes of payload each) through two
 * device slots and three streams, so that the upload of chunk c + 1, the kernel of chunk c and
 * the download of chunk c - 1 overlap.  A chunk is framed on its own (offsets rebased to the
 * chunk); its output goes where the chunks before it ended, known once their kernels are done.
 */
static void frame_pipe_free(h264gpu_ctx *ctx)
{
	if (ctx->fp.s_up == NULL)
		return;
	for (int b = 0; b < 2; b++) {
		cudaEventDestroy(ctx->fp.ev_up[b]);
		cudaEventDestroy(ctx->fp.ev_k[b]);
		cudaEventDestroy(ctx->fp.ev_dn[b]);
		cudaFree(ctx->fp.d_in[b].p);
		cudaFree(ctx->fp.d_out[b].p);
		cudaFree(ctx->fp.d_tab[b].p);
		cudaFreeHost(ctx->fp.h_tab[b].p);
	}
	cudaStreamDestroy(ctx->fp.s_up);
	cudaStreamDestroy(ctx->fp.s_k);
	cudaStreamDestroy(ctx->fp.s_dn);
	memset(&ctx->fp, 0, sizeof(ctx->fp));
}

static int frame_pipe_init(h264gpu_ctx *ctx)
{
	if (ctx->fp.s_up != NULL)
		return 0;
	CU_TRY(cudaStreamCreateWithFlags(&ctx->fp.s_up, cudaStreamNonBlocking));
	CU_TRY(cudaStreamCreateWithFlags(&ctx->fp.s_k, cudaStreamNonBlocking));
	CU_TRY(cudaStreamCreateWithFlags(&ctx->fp.s_dn, cudaStreamNonBlocking));
	for (int b = 0; b < 2; b++) {
		CU_TRY(cudaEventCreateWithFlags(&ctx->fp.ev_up[b], cudaEventDisableTiming));
		CU_TRY(cudaEventCreateWithFlags(&ctx->fp.ev_k[b], cudaEventDisableTiming));
		CU_TRY(cudaEventCreateWithFlags(&ctx->fp.ev_dn[b], cudaEventDisableTiming));
	}
	return 0;
}

/* a slot's buffer, grown only when nothing of the pipeline is in flight */
static int frame_pipe_grow(h264gpu_ctx *ctx, struct h264gpu_ctx::h264gpu_pool *pl, size_t bytes, bool pinned)
{
	if (bytes <= pl->cap)
		return 0;
	CU_TRY(cudaStreamSynchronize(ctx->fp.s_up));
	CU_TRY(cudaStreamSynchronize(ctx->fp.s_k));
	CU_TRY(cudaStreamSynchronize(ctx->fp.s_dn));
	return pinned ? h264gpu_pool_host(ctx, pl, bytes) : h264gpu_pool_dev(ctx, pl, bytes);
}

static int frame_host_pipelined(h264gpu_ctx *ctx, const uint8_t *h_rbsp, const uint64_t *h_off, uint64_t n, int sc_len,
				uint8_t *h_out, uint64_t out_cap, uint64_t *h_out_off, uint64_t *total)
{
	int r = frame_pipe_init(ctx);
	if (r < 0)
		return r;
	const uint64_t target = ctx->chunk_bytes;
	uint64_t k0[2] = {0, 0}, k1[2] = {0, 0}; /* payloads of the chunk in each slot */
	bool used[2] = {false, false};
	uint64_t base = 0;  /* output bytes of the chunks finished so far */
	uint64_t next_k = 0; /* first payload of the next chunk to issue */
	uint64_t issued = 0, finished = 0;
	while (finished < issued || next_k < n) {
		/* issue the next chunk (one ahead of the chunk being finished) */
		if (next_k < n && issued - finished < 2) {
			const int b = (int)(issued & 1);
			const uint64_t a0 = next_k;
			uint64_t lo = a0 + 1, hi = n; /* smallest a1 > a0 with off[a1] - off[a0] >= target, else n */
			while (lo < hi) {
				const uint64_t mid = (lo + hi) >> 1;
				if (h_off[mid] - h_off[a0] >= target)
					hi = mid;
				else
					lo = mid + 1;
			}
			const uint64_t a1 = lo, nc = a1 - a0, len_c = h_off[a1] - h_off[a0];
			const uint64_t bound = len_c + len_c / 2 + (uint64_t)sc_len * nc + 64;
			if ((r = frame_pipe_grow(ctx, &ctx->fp.d_in[b], len_c + 64, false)) < 0 ||
			    (r = frame_pipe_grow(ctx, &ctx->fp.d_out[b], bound + 64, false)) < 0 ||
			    (r = frame_pipe_grow(ctx, &ctx->fp.d_tab[b], (2 * nc + 4) * 8, false)) < 0 ||
			    (r = frame_pipe_grow(ctx, &ctx->fp.h_tab[b], (2 * nc + 4) * 8, true)) < 0)
				return r;
			uint64_t *roff = (uint64_t *)ctx->fp.h_tab[b].p, *back = roff + nc + 1;
			for (uint64_t i = 0; i <= nc; i++)
				roff[i] = h_off[a0 + i] - h_off[a0];
			uint64_t *d_off = (uint64_t *)ctx->fp.d_tab[b].p, *d_out_off = d_off + nc + 1, *d_total = d_out_off + nc + 1;
			if (used[b]) /* the kernel that read this slot's input is done */
				CU_TRY(cudaStreamWaitEvent(ctx->fp.s_up, ctx->fp.ev_k[b], 0));
			if (len_c)
				CU_TRY(cudaMemcpyAsync(ctx->fp.d_in[b].p, h_rbsp + h_off[a0], len_c, cudaMemcpyHostToDevice, ctx->fp.s_up));
			CU_TRY(cudaMemcpyAsync(d_off, roff, (nc + 1) * 8, cudaMemcpyHostToDevice, ctx->fp.s_up));
			CU_TRY(cudaEventRecord(ctx->fp.ev_up[b], ctx->fp.s_up));
			CU_TRY(cudaStreamWaitEvent(ctx->fp.s_k, ctx->fp.ev_up[b], 0));
			if (used[b]) /* the slot's previous output has been downloaded */
				CU_TRY(cudaStreamWaitEvent(ctx->fp.s_k, ctx->fp.ev_dn[b], 0));
			r = frame_dev_impl(ctx, (const uint8_t *)ctx->fp.d_in[b].p, d_off, nc, sc_len, (uint8_t *)ctx->fp.d_out[b].p,
					   bound, d_out_off, d_total, ctx->fp.s_k, true, len_c);
			if (r < 0)
				return r;
			/* out_off[0 .. nc) and the chunk's total come back with the kernel */
			CU_TRY(cudaMemcpyAsync(back, d_out_off, (nc + 2) * 8, cudaMemcpyDeviceToHost, ctx->fp.s_k));
			CU_TRY(cudaEventRecord(ctx->fp.ev_k[b], ctx->fp.s_k));
			k0[b] = a0;
			k1[b] = a1;
			used[b] = true;
			next_k = a1;
			issued++;
			if (next_k < n && issued - finished < 2)
				continue; /* fill the pipeline before waiting */
		}
		/* finish the oldest chunk in flight: its size places it (and everything after it) */
		const int b = (int)(finished & 1);
		const uint64_t nc = k1[b] - k0[b];
		CU_TRY(cudaEventSynchronize(ctx->fp.ev_k[b]));
		const uint64_t *back = (const uint64_t *)ctx->fp.h_tab[b].p + nc + 1;
		const uint64_t total_c = back[nc + 1];
		if (h_out_off)
			for (uint64_t i = 0; i < nc; i++)
				h_out_off[k0[b] + i] = back[i] + base;
		const uint64_t room = out_cap > base ? out_cap - base : 0;
		const uint64_t ncopy = total_c < room ? total_c : room;
		if (ncopy)
			CU_TRY(cudaMemcpyAsync(h_out + base, ctx->fp.d_out[b].p, ncopy, cudaMemcpyDeviceToHost, ctx->fp.s_dn));
		CU_TRY(cudaEventRecord(ctx->fp.ev_dn[b], ctx->fp.s_dn));
		base += total_c;
		finished++;
	}
	CU_TRY(cudaStreamSynchronize(ctx->fp.s_dn));
	if (h_out_off)
		h_out_off[n] = base;
	*total = base;
	return base > out_cap ? -ENOBUFS : 0;
}

extern "C" int h264gpu_frame_host(h264gpu_ctx *ctx, const uint8_t *h_rbsp, const uint64_t *h_off,
				  uint64_t n, int sc_len, uint8_t *h_out, uint64_t out_cap,
				  uint64_t *h_out_off, uint64_t *total)
{
	int r = h264gpu_use(ctx);
	if (r < 0)
		return r;
	if (h_off == NULL || h_out == NULL || total == NULL)
		return -EINVAL;
	const uint64_t len = h_off[n];
	if (len && h_rbsp == NULL)
		return -EINVAL;
	if (sc_len != 0 && sc_len != 3 && sc_len != 4)
		return -EINVAL;
	/* long inputs: chunked, copies and kernels overlapped (H264GPU_FRAME_PIPE=0: one shot) */
	{
		const char *e = getenv("H264GPU_FRAME_PIPE");
		if (n >= 2 && len >= 2 * (uint64_t)ctx->chunk_bytes && !(e != NULL && atoi(e) == 0)) {
			r = frame_host_pipelined(ctx, h_rbsp, h_off, n, sc_len, h_out, out_cap, h_out_off, total);
			if (r < 0 && r != -ENOBUFS) {
				/* copies from / to the caller's buffers may still be in flight */
				cudaStreamSynchronize(ctx->fp.s_up);
				cudaStreamSynchronize(ctx->fp.s_k);
				cudaStreamSynchronize(ctx->fp.s_dn);
			}
			return r;
		}
	}
	/* pooled device buffers of the reader session on its private stream: a call costs no
	 * cudaMalloc / cudaFree (round 1 allocated four buffers per call on the default stream) */
	cudaStream_t st;
	if ((r = h264gpu_reader_stream(ctx, &st)) < 0 ||
	    (r = h264gpu_pool_dev(ctx, &ctx->rd_stream, len + 64)) < 0 ||
	    (r = h264gpu_pool_dev(ctx, &ctx->rd_records, out_cap + 64)) < 0 ||
	    (r = h264gpu_pool_dev(ctx, &ctx->rd_tab, (2 * n + 4) * 8)) < 0)
		return r;
	ctx->rd_stream_len = 0; /* the reader's resident buffer is gone */
	uint8_t *d_rbsp = (uint8_t *)ctx->rd_stream.p, *d_out = (uint8_t *)ctx->rd_records.p;
	uint64_t *d_off = (uint64_t *)ctx->rd_tab.p, *d_out_off = d_off + n + 1, *d_total = d_out_off + n + 1;
	if (len)
		CU_TRY(cudaMemcpyAsync(d_rbsp, h_rbsp, len, cudaMemcpyHostToDevice, st));
	CU_TRY(cudaMemcpyAsync(d_off, h_off, (n + 1) * 8, cudaMemcpyHostToDevice, st));
	r = frame_dev_impl(ctx, d_rbsp, d_off, n, sc_len, d_out, out_cap, d_out_off, d_total, st, true, len);
	if (r < 0)
		return r;
	CU_TRY(cudaMemcpyAsync(total, d_total, 8, cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaStreamSynchronize(st));
	if (h_out_off)
		CU_TRY(cudaMemcpyAsync(h_out_off, d_out_off, (n + 1) * 8, cudaMemcpyDeviceToHost, st));
	const uint64_t ncopy = *total < out_cap ? *total : out_cap;
	if (ncopy)
		CU_TRY(cudaMemcpyAsync(h_out, d_out, ncopy, cudaMemcpyDeviceToHost, st));
	CU_TRY(cudaStreamSynchronize(st));
	return *total > out_cap ? -ENOBUFS : 0;
}
