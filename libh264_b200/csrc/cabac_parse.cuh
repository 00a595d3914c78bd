/*
 * cabac_parse.cuh — K5: slice-parallel CABAC macroblock syntax parse.  One independent slice
 * per thread; a slice's 460 context states live in SHARED memory (one byte each, interleaved
 * over the block's slices so lanes hit different banks), the range / transition tables too.
 *
 * The reference cannot decode CABAC slice data (src/h264_syntax_slice_data.h:715-717 returns
 * before slice_data_begin).  What this kernel delivers per macroblock is what
 * the reference delivers for a CAVLC slice (h264_ctx_cbs.slice_data_mb: mb_addr, mb_type,
 * include/h264/h264_ctx.h:78-82) plus the same syntax-element checksum as K4, so a CABAC
 * slice and a CAVLC slice carrying the same syntax elements give identical records — that is
 * how tests pin this path to the reference's own parse (tests/test_cabac.py).
 * Syntax walk: cabac_syntax.h (Rec. ITU-T H.264 7.3.4, 7.3.5, 9.3); engine: cabac_engine.h.
 */
#ifndef CABAC_PARSE_CUH
#define CABAC_PARSE_CUH

#include <stdint.h>

#include "h264gpu_slice.h"

#define CABAC_HD __device__
#define CABAC_OUTLINE __device__ __noinline__
#define CABAC_CONST __device__ const
#include "cabac_engine.h"
#include "cabac_syntax.h"
#define CABAC_TAB __device__ const
#include "cabac_tables.h"

#ifndef EIO
#define EIO 5
#endif
#ifndef ENOSYS
#define ENOSYS 38
#endif
#ifndef ENOBUFS
#define ENOBUFS 105
#endif

namespace cabac {

struct CabacArgs {
	const uint8_t *stream;
	uint64_t stream_len;
	const h264gpu_slice_params *params;
	uint32_t n_slices;
	h264gpu_mb_record *records;
	h264gpu_slice_result *results;
	Nb *ring;             /* n_slices x ring_stride neighbour records */
	uint64_t ring_stride; /* in Nb units */
	uint32_t ring_w;      /* widest picture (in MBs) a ring row can hold */
	uint32_t lanes_log2;  /* log2 of the slices carried by one warp (0..5) */
	uint32_t *next_slice; /* ticket counter (starts at the number of working lanes), or NULL */
	const uint32_t *order; /* slice indices, longest NAL first, or NULL: list order */
};

constexpr uint32_t kTabBytes = 128 * 8; /* Tables::fused */
/* per-slice working set kept in SHARED memory (walker state + the macroblock being decoded):
 * as thread-local data it would live in lane-interleaved local memory, where a warp that runs
 * one or a few slices still occupies 32 lanes' worth of cache lines and thrashes L1 */
constexpr uint32_t kSlotBytes = (uint32_t)((sizeof(Walk<Dec>) + sizeof(Mb) + 15) & ~(size_t)15) + 16;

__host__ __device__ inline uint32_t smem_bytes(uint32_t slices_per_block)
{
	return ((kTabBytes + kNumCtx * slices_per_block + 15) & ~15u) + slices_per_block * kSlotBytes;
}

__device__ __forceinline__ void parse_slice(const uint8_t *stream, uint64_t stream_len, const h264gpu_slice_params &sp, Nb *ring,
					     h264gpu_mb_record *rec, h264gpu_slice_result &res, uint8_t *ctx_states,
					     uint32_t ctx_stride, const uint8_t *tabs, uint8_t *slot)
{
	res.status = 0;
	res.mb_count = 0;
	res.end_bit = 0;
	if (!sp.entropy_coding_mode_flag) {
		res.status = H264GPU_SLICE_SKIPPED;
		return;
	}
	if (sp.mbaff_frame_flag || sp.num_slice_groups_minus1 != 0 || sp.pic_width_in_mbs == 0 ||
	    sp.chroma_array_type == 3 || sp.slice_type == ST_SI || sp.slice_type > ST_SI) {
		res.status = -ENOSYS;
		return;
	}
	/* a parameter block that points outside the stream must not be followed */
	if ((uint64_t)sp.nal_off + sp.nal_len > stream_len || sp.data_bit_off >= 8ull * sp.nal_len) {
		res.status = -22; /* -EINVAL */
		return;
	}
	Walk<Dec> &w = *reinterpret_cast<Walk<Dec> *>(slot);
	Mb &m = *reinterpret_cast<Mb *>(slot + ((sizeof(Walk<Dec>) + 7) & ~(size_t)7));
	w.begin_slice(&sp, ring);
	w.c.st = ctx_states;
	w.c.stride = ctx_stride;
	w.c.t.fused = reinterpret_cast<const uint64_t *>(tabs);
	init_contexts(ctx_states, ctx_stride, cabac_init_mn[sp.slice_type == ST_I ? 0 : 1 + (sp.cabac_init_idc % 3)],
		      sp.slice_qp);
	w.c.start(stream + sp.nal_off, sp.nal_len, sp.data_bit_off);
	const uint32_t pic_size = (uint32_t)sp.pic_width_in_mbs * sp.pic_height_in_mbs;
	uint32_t cur = sp.first_mb_in_slice, count = 0;
	int status = w.c.failed() ? -EIO : 0;
	while (!status) {
		if (count >= sp.mb_out_cap || cur >= pic_size) {
			status = -ENOBUFS;
			break;
		}
		bool skipped = false, end = false;
		if (!w.mb_step(cur, skipped, m, end)) {
			status = -EIO;
			break;
		}
		rec[count].mb_addr = cur;
		rec[count].mb_type = m.mb_type;
		rec[count].hash = w.hash;
		count++;
		cur++;
		if (end)
			break;
	}
	res.status = status;
	res.mb_count = count;
	res.end_bit = w.c.raw_bitpos();
}

/*
 * A slice is one serial chain of bins: a warp carries ONE slice (one working lane) while the
 * slices fit the resident warps, because two slices in one warp diverge at every branch and
 * take turns (measured: 2 lanes per warp = half the speed per warp).  What matters is that every
 * SM sub-partition always has its ~8 chains to switch between and that the longest chains start
 * first: the grid is persistent (at most 32 one-warp blocks per SM, the register file's limit at
 * 64 registers x 32 lanes), working lanes pull slices from a ticket counter, and the tickets run
 * over the slices in descending order of NAL length (cavlc2::order_kernel).  The round-1 launch
 * (slice i -> warp i in list order, 4 slices per warp at 16000 slices) ran ~3.4 waves each as
 * long as its longest I slice: 14.7 % of the issue slots.
 * The neighbour ring belongs to the working lane, not to the slice.
 * Dynamic shared memory: [1 KB fused table][460 x (slices per block) context bytes][slices x working set].
 */
__global__ void __launch_bounds__(128, 8) cabac_parse_kernel(const CabacArgs a)
{
	extern __shared__ uint8_t smem[];
	for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x)
		reinterpret_cast<uint64_t *>(smem)[i] = fused_entry(i, cabac_range_lps, cabac_trans_lps, cabac_trans_mps);
	__syncthreads();
	const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t step = 32u >> a.lanes_log2; /* active lanes are multiples of this */
	if (lane & (step - 1))
		return;
	const uint32_t unit = (gwarp << a.lanes_log2) + lane / step; /* working-lane number = first ticket */
	const uint32_t per_block = (blockDim.x >> 5) << a.lanes_log2;
	const uint32_t slot = ((threadIdx.x >> 5) << a.lanes_log2) + lane / step;
	uint8_t *slots = smem + ((kTabBytes + kNumCtx * per_block + 15) & ~15u);
	for (uint32_t t = unit; t < a.n_slices; t = a.next_slice ? atomicAdd(a.next_slice, 1u) : a.n_slices) {
		const uint32_t i = a.order ? a.order[t] : t;
		const h264gpu_slice_params sp = a.params[i];
		h264gpu_slice_result res;
		if (sp.pic_width_in_mbs > a.ring_w) {
			res.status = -7; /* -E2BIG */
			res.mb_count = 0;
			res.end_bit = 0;
			a.results[i] = res;
			continue;
		}
		parse_slice(a.stream, a.stream_len, sp, a.ring + (uint64_t)unit * a.ring_stride, a.records + sp.mb_out_off,
			    res, smem + kTabBytes + slot, per_block, smem, slots + (size_t)slot * kSlotBytes);
		a.results[i] = res;
	}
}

} /* namespace cabac */

#endif /* CABAC_PARSE_CUH */
