/*
 * conceal.cuh — N4: bulk synthesis of concealment slices on the GPU, one slice per lane.
 *
 * What the reference writes one slice at a time with h264_write_grey_i_slice /
 * h264_write_skipped_p_slice (src/h264_writer.c:49-219): after the NAL + slice header (header
 * syntax stays host work: a few dozen bits per slice, passed in as bytes), the slice data of
 *   grey I      every macroblock I_16x16_2_0_0, DC chroma prediction, mb_qp_delta 0, empty DC block
 *   skipped P   every macroblock skipped
 * in CAVLC (a fixed 8-bit pattern per macroblock / one mb_skip_run) or CABAC (the arithmetic
 * encoder of 9.3.4.2 over the handful of contexts these macroblocks touch: src/h264_cabac.c,
 * src/h264_bac.c:232-345).  Output: the unescaped NAL payloads back to back, ready for the
 * escaping + framing stage (h264gpu_frame_dev).  Two passes of the same code: lengths first
 * (STORE = false), then bytes at the scanned offsets.
 */
#ifndef CONCEAL_CUH
#define CONCEAL_CUH

#include "gpu_compat.h"
#include "h264gpu_slice.h"

#ifndef CABAC_TAB
#ifdef H264_EMU
#define CABAC_TAB static const
#else
#define CABAC_TAB __device__ const
#endif
#endif
#include "cabac_tables.h"

namespace conceal {

/* MSB-first bit writer into a byte buffer (or a bit counter only) */
template <bool STORE>
struct BitOut {
	uint8_t *p;
	uint64_t bits; /* bits written */
	uint32_t acc;  /* pending bits, right aligned */
	uint32_t nacc;

	__device__ __forceinline__ void put(uint32_t v, uint32_t n) /* n <= 16 */
	{
		acc = (acc << n) | (v & ((1u << n) - 1));
		nacc += n;
		bits += n;
		while (nacc >= 8) {
			nacc -= 8;
			if (STORE)
				*p = (uint8_t)(acc >> nacc);
			p++;
		}
	}
	__device__ __forceinline__ bool aligned() const { return nacc == 0; }
};

/* the contexts a concealment slice touches, packed into 16 slots */
__device__ __forceinline__ uint32_t ctx_slot(uint32_t ctx_idx)
{
	/* 3..10 -> 0..7, 11 -> 8, 24 -> 9, 60 -> 10, 64 -> 11, 85..88 -> 12..15 */
	return ctx_idx <= 10 ? ctx_idx - 3 : ctx_idx == 11 ? 8 : ctx_idx == 24 ? 9 : ctx_idx == 60 ? 10
	       : ctx_idx == 64 ? 11 : ctx_idx - 85 + 12;
}

/* 9.3.4.2 arithmetic encoder (Figures 9-7 .. 9-12); states as pStateIdx | valMPS << 6 */
template <bool STORE>
struct Enc {
	BitOut<STORE> *o;
	uint32_t low, range, outstanding;
	bool first_bit;
	uint64_t st_lo, st_hi; /* 16 context states, one byte each */

	__device__ __forceinline__ uint32_t get_state(uint32_t s) const
	{
		return (uint32_t)((s < 8 ? st_lo >> (8 * s) : st_hi >> (8 * (s - 8))) & 0xff);
	}
	__device__ __forceinline__ void set_state(uint32_t s, uint32_t v)
	{
		if (s < 8)
			st_lo = (st_lo & ~(0xffull << (8 * s))) | (uint64_t)v << (8 * s);
		else
			st_hi = (st_hi & ~(0xffull << (8 * (s - 8)))) | (uint64_t)v << (8 * (s - 8));
	}
	/* 9.3.1.1: preCtxState = Clip3(1, 126, ((m * Clip3(1, 51, QP)) >> 4) + n) */
	__device__ __forceinline__ void init(BitOut<STORE> *out, bool intra, uint32_t cabac_init_idc, int32_t slice_qp)
	{
		o = out;
		low = 0;
		range = 510;
		outstanding = 0;
		first_bit = true;
		st_lo = st_hi = 0;
		if (!intra && cabac_init_idc > 2)
			return; /* the reference leaves the states untouched (all zero) here too */
		const int8_t(*mn)[2] = cabac_init_mn[intra ? 0 : 1 + cabac_init_idc];
		const int32_t qp = slice_qp < 1 ? 1 : slice_qp > 51 ? 51 : slice_qp;
		const uint32_t idx[16] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 24, 60, 64, 85, 86, 87, 88};
#pragma unroll
		for (uint32_t s = 0; s < 16; s++) {
			const int32_t pre = ((mn[idx[s]][0] * qp) >> 4) + mn[idx[s]][1];
			const uint32_t v = pre <= 63 ? (uint32_t)(63 - (pre < 1 ? 1 : pre))
						     : (uint32_t)(((pre > 126 ? 126 : pre) - 64) | 0x40);
			set_state(s, v);
		}
	}
	__device__ __forceinline__ void put_bit(uint32_t bit)
	{
		if (first_bit)
			first_bit = false;
		else
			o->put(bit, 1);
		for (; outstanding > 0; outstanding--)
			o->put(!bit, 1);
	}
	__device__ __forceinline__ void renorm()
	{
		while (range < 256) {
			if (low < 256) {
				put_bit(0);
			} else if (low < 512) {
				low -= 256;
				outstanding++;
			} else {
				low -= 512;
				put_bit(1);
			}
			range <<= 1;
			low <<= 1;
		}
	}
	__device__ __forceinline__ void decision(uint32_t ctx_idx, uint32_t bin)
	{
		const uint32_t s = ctx_slot(ctx_idx);
		const uint32_t st = get_state(s);
		const uint32_t idx = st & 0x3f, mps = st >> 6;
		const uint32_t lps_range = cabac_range_lps[idx][(range >> 6) & 3];
		range -= lps_range;
		if (bin == mps) {
			set_state(s, cabac_trans_mps[idx] | mps << 6);
		} else {
			low += range;
			range = lps_range;
			set_state(s, cabac_trans_lps[idx] | (idx == 0 ? !mps : mps) << 6);
		}
		renorm();
	}
	/* EncodeTerminate + EncodeFlush; the last bit written is the rbsp_stop_one_bit */
	__device__ __forceinline__ void terminate(uint32_t bin)
	{
		range -= 2;
		if (!bin) {
			renorm();
			return;
		}
		low += range;
		range = 2;
		renorm();
		put_bit((low >> 9) & 1);
		o->put(((low >> 7) & 3) | 1, 2);
	}
};

struct ConcealArgs {
	const struct h264gpu_conceal_params *params;
	uint32_t n;
	const uint8_t *hdr;  /* the slices' header bytes */
	uint64_t *off;       /* STORE = false: off[k + 1] = payload bytes of slice k (off[0] = 0), scanned
				in place by conceal_scan; STORE = true: payload k starts at off[k] */
	uint8_t *out;
	uint64_t cap;        /* bytes at out: nothing is written if the total exceeds it */
};

template <bool STORE>
__global__ void __launch_bounds__(128) conceal_kernel(const ConcealArgs a)
{
	const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= a.n)
		return;
	const h264gpu_conceal_params c = a.params[k];
	BitOut<STORE> o;
	if (STORE && a.off[a.n] > a.cap)
		return;
	o.p = STORE ? a.out + a.off[k] : nullptr;
	o.bits = 0;
	o.acc = 0;
	o.nacc = 0;
	/* NAL header + slice header, written by the host syntax walk (unescaped) */
	const uint8_t *h = a.hdr + c.hdr_off;
	for (uint32_t i = 0; i < (c.hdr_bits >> 3); i++)
		o.put(h[i], 8);
	if (c.hdr_bits & 7)
		o.put((uint32_t)h[c.hdr_bits >> 3] >> (8 - (c.hdr_bits & 7)), c.hdr_bits & 7);

	if (!c.entropy_coding_mode_flag) {
		if (c.kind == H264GPU_CONCEAL_SKIPPED_P) {
			/* one mb_skip_run covers the slice: ue(mb_count) */
			const uint32_t v = c.mb_count + 1;
			const uint32_t len = 32 - (uint32_t)__clz((int)v); /* significant bits */
			for (uint32_t i = 1; i < len; i++)
				o.put(0, 1);
			for (uint32_t i = len; i > 0;) {
				const uint32_t take = i > 16 ? 16 : i;
				i -= take;
				o.put(v >> i, take);
			}
		} else {
			/* per macroblock: mb_type 3 (I_16x16_2_0_0) = 00100, intra_chroma_pred_mode 0 = 1,
			 * mb_qp_delta 0 = 1, coeff_token(nC = 0, no coefficient) = 1 */
			for (uint32_t i = 0; i < c.mb_count; i++)
				o.put(0x27, 8);
		}
		o.put(1, 1); /* rbsp_trailing_bits */
		while (!o.aligned())
			o.put(0, 1);
	} else {
		while (!o.aligned())
			o.put(1, 1); /* cabac_alignment_one_bit */
		Enc<STORE> e;
		const bool intra = c.slice_type == 2 || c.slice_type == 4;
		e.init(&o, intra, c.cabac_init_idc, c.slice_qp);
		const uint32_t w = c.pic_width_in_mbs, first = c.first_mb_in_slice;
		uint32_t col = w ? first % w : 0;
		for (uint32_t i = 0; i < c.mb_count; i++) {
			const uint32_t addr = first + i;
			if (c.kind == H264GPU_CONCEAL_SKIPPED_P) {
				/* mb_skip_flag = 1; every available neighbour is skipped too: ctxIdxInc 0 */
				e.decision(c.slice_type == 1 ? 24 : 11, 1);
			} else {
				/* neighbours inside the slice, frame macroblocks (src/h264_macroblock.c:306-351) */
				const uint32_t av_a = w != 0 && col != 0 && addr >= first + 1;
				const uint32_t av_b = addr >= first + w;
				/* mb_type 3, Table 9-36 "1 0 0 0 1 0": bin 0 ctx 3 + (A is not I_NxN) + (B ...), bin 1
				 * terminate, bins 2..5 ctx 3+3, 3+4, 3+6, 3+7 */
				e.decision(3 + av_a + av_b, 1);
				e.terminate(0);
				e.decision(6, 0);
				e.decision(7, 0);
				e.decision(9, 1);
				e.decision(10, 0);
				e.decision(64, 0); /* intra_chroma_pred_mode 0 */
				e.decision(60, 0); /* mb_qp_delta 0 */
				/* coded_block_flag of Intra16x16DCLevel: an unavailable neighbour of an intra
				 * macroblock counts as coded (9.3.3.1.1.9) */
				e.decision(85 + (av_a ? 0u : 1u) + (av_b ? 0u : 2u), 0);
			}
			e.terminate(i == c.mb_count - 1); /* end_of_slice_flag */
			if (++col == w)
				col = 0;
		}
		while (!o.aligned())
			o.put(0, 1);
	}
	if (!STORE)
		a.off[k + 1] = o.bits >> 3;
}

/* off[1..n] lengths -> off[0..n] offsets (one block; n is the number of slices of a batch) */
__global__ void __launch_bounds__(1024) conceal_scan(uint64_t *off, uint32_t n)
{
#ifndef H264_EMU
	__shared__ uint64_t part[1024];
	const uint32_t t = threadIdx.x;
	const uint32_t per = (n + 1023) / 1024;
	const uint32_t b = t * per, e = b + per < n ? b + per : n;
	uint64_t s = 0;
	for (uint32_t i = b; i < e; i++)
		s += off[i + 1];
	part[t] = s;
	__syncthreads();
	if (t == 0) {
		uint64_t acc = 0;
		for (uint32_t i = 0; i < 1024; i++) {
			const uint64_t v = part[i];
			part[i] = acc;
			acc += v;
		}
		off[0] = 0;
	}
	__syncthreads();
	uint64_t acc = part[t];
	for (uint32_t i = b; i < e; i++) {
		acc += off[i + 1];
		off[i + 1] = acc;
	}
#endif
}

/*
 * h264_rewrite_slice_header in bulk (src/h264_writer.c:351-361: memcpy of the whole header bytes,
 * blend of the byte shared with the slice data): 64 threads per patch, one byte each.  A patch
 * that does not fit the stream (or is malformed) is left out.
 */
__global__ void patch_headers_kernel(uint8_t *stream, uint64_t stream_len, const struct h264gpu_hdr_patch *patches,
				     uint32_t n)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t p = g >> 6, j = g & 63u;
	if (p >= n)
		return;
	const struct h264gpu_hdr_patch &q = patches[p];
	const uint32_t nb = q.nbytes, tb = q.tail_bits;
	if (nb > 63u || tb > 7u)
		return;
	const uint64_t end = q.nal_off + nb + (tb ? 1u : 0u);
	if (end > stream_len || end < q.nal_off)
		return;
	if (j < nb) {
		stream[q.nal_off + j] = q.bytes[j];
	} else if (j == nb && tb) {
		const uint32_t keep = (1u << (8 - tb)) - 1u;
		stream[q.nal_off + j] = (uint8_t)((q.bytes[j] & ~keep) | (stream[q.nal_off + j] & keep));
	}
}

} /* namespace conceal */

#endif /* CONCEAL_CUH */
