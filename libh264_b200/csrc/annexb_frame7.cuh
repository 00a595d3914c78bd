/*
 * annexb_frame7.cuh — K3 third generation (opt-in: H264GPU_FRAME_GEN=7): writer-side
 * emulation-prevention-byte insertion and start-code framing with WARP-AUTONOMOUS SPANS, the
 * insert count of a span published an iteration AHEAD of its look-back, a three-level chain
 * (span / group of 32 spans / supergroup of 32 groups) and soft seams.
 *
 * Why it was written: frame6_kernel (annexb_frame6.cuh, block-wide 32 KiB tiles) issues 18.7 k
 * warp-instructions per tile with 59 % of its issue slots busy (profiles/r02_frame_kernel_raw.csv):
 * 21 % of warp time sits at the barrier behind the look-back (the tile waits for the AGGREGATES
 * of the ~100 tiles in flight before it, i.e. for the slowest of them) and 15 % at the
 * end-of-tile barrier (the slowest of eight warps).  The writer's output is packed, so its chain
 * is a true stream-long prefix: it cannot be cut at start codes the way the in-place scan kernel
 * (annexb_scan7.cuh) cuts its chain.  Here the wait is removed instead of cut:
 *
 *  * A warp owns TWO span buffers.  In one loop iteration it classifies span B, publishes B's
 *    insert count, and only then looks back for span A (classified an iteration ago) and emits A.
 *    Every predecessor of A took its ticket before A did and published right after its own
 *    classification, an iteration ago: the look-back of A finds all of them there.  No barrier,
 *    no polling in the steady state (measured: 0.00 / 0.01 misses per span on the first two
 *    levels, profiles/r02_frame7_trace_two_buffers.txt).
 *  * A TICKET IS NEVER HELD ACROSS A WAIT.  The next span's ticket is taken after the look-back
 *    (the only place a warp can wait) and after the emit, unless the emit is short (no byte-wise
 *    row in span A; then before it, with an L2 prefetch).  The first build took the ticket at the
 *    top of the iteration: a waiting warp then held an unpublished span, its wait (plus a polling
 *    delay) was handed on to every warp that needed that span, and the waits grew linearly along
 *    the stream (16 us per look-back in the first eighth of 1 GiB, 256 us in the last: 10.3 ms
 *    per GiB, profiles/r02_frame7_trace_ticket_at_top.txt, profiles/r02_frame7_steps.txt).
 *  * Publishing that far ahead means the nearest span with a known PREFIX is thousands of spans
 *    back (every span in flight has an aggregate and no prefix), too deep for a linear look-back.
 *    The chain therefore has three levels, all single 64-bit self-validating words:
 *      span_w[s]   valid bit | inserts of span s
 *      group_w[g]  spans published << 40 | sum of their inserts (one atomicAdd per span); the
 *                  warp that completes a group adds its sum into
 *      super_w[q]  groups completed << 40 | sum (4 MiB of source per supergroup)
 *      super_p[q]  1 + inserts before supergroup q: a decoupled look-back over supergroups,
 *                  resolved by the first spans of the supergroup, read as one word by the others
 *    A look-back is three coalesced probes issued together: the spans of the own group before
 *    the span, the groups of the own supergroup before the group, super_p of the supergroup.
 *  * The 16 bytes before and after a span are loaded with it (one bulk copy of SPAN + 32 bytes).
 *    The zero run entering the span is read off the staged bytes, so the pre-pass only finds the
 *    first payload of every span and clears the chain words (no per-tile tail array, no memset).
 *    SOFT SEAMS: a 16-byte unit that straddles two spans is written whole by the span it starts
 *    in (which works out the insert mask of the chunk after its end itself) unless a payload
 *    starts or the input ends next to the seam; a hard seam costs ~290 warp-instructions of
 *    byte stores per side.
 *
 * Deadlock freedom: a warp publishes span B before it waits for anything but B's own bulk copy;
 * its look-backs wait only for spans with smaller tickets.  The smallest unpublished span is
 * owned by a resident warp (tickets are taken by running warps) that is between its ticket and
 * its publication, where nothing waits for another warp.
 *
 * What it reaches (B200, 1 GiB of RBSP, config 5): 0.885 ms = 1213 GB/s of payload, the same as
 * gen 6 (0.876 ms), which stays the default.  The chain wait is gone, the kernel is not faster:
 * both generations issue 2100-2300 warp-instructions per 4 KiB at 56-60 % of the issue slots.
 * Gen 6 loses its slots to barriers; gen 7, at two 4 KiB buffers per warp, has 5 warps per
 * scheduler and loses them to dependent-issue latency; with ONE buffer per warp (NBUF = 1: the
 * bytes are staged a second time for the emit, 11 % more DRAM reads) an SM holds 32 warps, the
 * issue slots are 61 % busy, and instruction fetch becomes the top stall (32 unrelated warps in
 * ~3500 instructions of code): 1.01 ms.  The writer is bound by its instruction count.
 *
 * Output per chunk, the byte-exact units, byte-wise rows and capacity rules are those of gen 6
 * (see annexb_frame6.cuh).
 *
 * Reference behaviour reproduced bit-exactly (Parrot-Developers/libh264):
 *   h264_bs_flush        src/h264_bitstream.c:54-81
 *   h264_bs_write_bits   src/h264_bitstream.c:211-239
 *   start code           src/h264.c:251-272
 */
#ifndef ANNEXB_FRAME7_CUH
#define ANNEXB_FRAME7_CUH

#include "annexb_frame6.cuh"

namespace frame7 {

using annexb::zmask4;
using frame::FrameArgs;
using frame::count_le;
using frame::lower_bound_off;
using frame6::finish_output;
using frame6::inserts_with_starts;
using frame6::put_byte;
using frame6::put_unit;
using frame6::valid16;

constexpr int kSoff = 16; /* payload starts of a span staged in shared memory (more: searched in off[]) */
constexpr uint64_t kValid = 1ull << 63;
constexpr uint64_t kOne = 1ull << 40; /* count field of group_w / super_w */
constexpr uint64_t kSumMask = kOne - 1;
constexpr uint32_t kNapMin = 128, kNapMax = 2048; /* back-off of a look-back poll, ns */
constexpr uint32_t kSoftBefore = 1u << 30, kSoftAfter = 1u << 31; /* SpanRegs::bw: soft seam at the span start / end */
#ifndef FRAME7_TICKET_EARLY
#define FRAME7_TICKET_EARLY 1
#endif
constexpr int kTicketEarly = FRAME7_TICKET_EARLY;
constexpr int kResolvers = 8;                       /* spans at the head of a supergroup that resolve its prefix */
constexpr int kPollSuper = 6;                       /* polls of super_p before any other span resolves it itself */

template <int ROWS> struct Cfg {
	static constexpr int SPAN_CH = 32 * ROWS; /* 16-byte chunks per span */
	static constexpr int SPAN = SPAN_CH * 16; /* bytes per span */
};

/* the staged bytes of a span: the 16 bytes before it, the span, the 16 bytes after it */
template <int ROWS> struct __align__(128) Data {
	uint8_t raw[16 + Cfg<ROWS>::SPAN + 32]; /* [16 bytes before the span][span][16 bytes after][pad] */
	uint64_t bar;
};

/* what classification leaves for the emission of a span */
template <int ROWS> struct __align__(16) Meta {
	uint16_t M[Cfg<ROWS>::SPAN_CH + 8]; /* insert mask per chunk (bit j = 03 before byte j) */
	uint32_t krow[ROWS + 2];            /* payload starts of the span before the row start */
	uint32_t soff[kSoff];               /* the span's first payload starts, relative to the span */
};

template <int ROWS> struct Buf { /* a span as the routines see it */
	Data<ROWS> &d;
	Meta<ROWS> &m;
};

/*
 * The slice of shared memory a warp owns.  NBUF = 2: span B's bytes stay staged from its
 * classification to its emission an iteration later (20 warps per SM at 4 KiB spans).  NBUF = 1:
 * only the masks stay; the bytes are staged a second time (from L2, the first read was ~10 us
 * ago) for the emission, and an SM holds 32 warps.
 */
template <int ROWS, int NBUF> struct __align__(128) WSmem {
	Data<ROWS> d[NBUF];
	Meta<ROWS> m[2];
	uint16_t E[Cfg<ROWS>::SPAN_CH]; /* inserts of the span before the chunk; candidate list before that */
	uint8_t dl[Cfg<ROWS>::SPAN_CH]; /* chunks that take the byte-exact path */
};

/* what a warp carries from the classification of a span to its emission */
struct SpanRegs {
	uint64_t k_lo, k_hi; /* payloads that start inside the span: off[k_lo .. k_hi) */
	uint32_t t;
	uint32_t bw;    /* rows that go byte by byte */
	uint32_t count; /* inserts of the span */
};

template <int ROWS> __global__ void frame7_prepass(const FrameArgs a)
{
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t > a.num_tiles)
		return;
	const uint64_t s = t * Cfg<ROWS>::SPAN < a.len ? t * Cfg<ROWS>::SPAN : a.len;
	a.first[t] = lower_bound_off(a.off, a.n, t == a.num_tiles ? a.len : s);
	if (t < a.num_tiles)
		a.desc[t] = 0;
	if ((t & 31) == 0)
		a.group_w[t >> 5] = 0;
	if ((t & 1023) == 0) {
		a.super_w[t >> 10] = 0;
		a.super_p[t >> 10] = t ? 0 : 1; /* nothing is inserted before the first supergroup */
	}
	if (t == 0)
		*a.ticket = 0;
}

/* payload starts of the span (off[k_lo .. k_hi)) at or before span position x */
template <int ROWS>
__device__ __forceinline__ uint32_t starts_le(const Buf<ROWS> s, const FrameArgs &a, uint64_t span_off, uint64_t k_lo,
					      uint64_t k_hi, uint32_t x)
{
	if (k_hi - k_lo > (uint64_t)kSoff)
		return count_le(a.off, k_lo, k_hi, span_off + x);
	uint32_t lo = 0, hi = (uint32_t)(k_hi - k_lo);
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if (s.m.soff[mid] <= x)
			lo = mid + 1;
		else
			hi = mid;
	}
	return lo;
}

/* span position of the span's payload start number i */
template <int ROWS>
__device__ __forceinline__ uint32_t start_at(const Buf<ROWS> s, const FrameArgs &a, uint64_t span_off, uint64_t k_lo,
					     uint64_t k_hi, uint32_t i)
{
	return k_hi - k_lo > (uint64_t)kSoff ? (uint32_t)(a.off[k_lo + i] - span_off) : s.m.soff[i];
}

/* bytes [from, to) of chunk c's own output sequence (its 03s included), byte stores; dst = where
 * output offset 0 of the chunk goes */
template <int ROWS>
__device__ __noinline__ void chunk_bytes(const Buf<ROWS> s, uint32_t c, uint32_t from, uint32_t to, uint8_t *dst,
					    const uint8_t *capend)
{
	const uint8_t *rawb = s.d.raw + 16 + c * 16;
	const uint32_t m = s.m.M[c];
	uint32_t o = 0;
	for (uint32_t j = 0; j < 16 && o < to; j++) {
		if ((m >> j) & 1) {
			if (o >= from)
				put_byte(dst + o, 3, capend);
			o++;
			if (o >= to)
				break;
		}
		if (o >= from)
			put_byte(dst + o, rawb[j], capend);
		o++;
	}
}

/* the aligned unit that starts ub bytes into chunk c's output (ub < 16 + inserts of c): 16 output
 * bytes from the source window that starts at the byte (or at the 03 before the byte) found there */
template <int ROWS>
__device__ __forceinline__ void gen_unit(const Buf<ROWS> s, uint32_t c, uint32_t ub, uint8_t *dst, const uint8_t *capend)
{
	const uint32_t m = s.m.M[c];
	/* smallest source byte j of the chunk whose output offset j + inserts(<= j) is >= ub */
	uint32_t j = ub < 15u ? ub : 15u;
	while (j > 0) {
		const uint32_t jj = j - 1;
		if (jj + (uint32_t)__popc(m & ((2u << jj) - 1u)) >= ub)
			j = jj;
		else
			break;
	}
	const uint32_t idx = j + (uint32_t)__popc(m & ((2u << j) - 1u));
	uint32_t dm = (m | (uint32_t)s.m.M[c + 1] << 16) >> j;
	if (idx == ub)
		dm &= ~1u; /* the 03 before byte j, if any, belongs to the unit before */
	dm &= 0xffffu;
	const uint32_t S = c * 16 + j;
	const uint32_t *raw32 = (const uint32_t *)(s.d.raw + 16);
	const uint32_t wi = S >> 2, sh = (S & 3) * 8;
	const uint32_t y0 = raw32[wi], y1 = raw32[wi + 1], y2 = raw32[wi + 2], y3 = raw32[wi + 3], y4 = raw32[wi + 4];
	uint64_t q0 = (uint64_t)__funnelshift_r(y0, y1, sh) | (uint64_t)__funnelshift_r(y1, y2, sh) << 32;
	uint64_t q1 = (uint64_t)__funnelshift_r(y2, y3, sh) | (uint64_t)__funnelshift_r(y3, y4, sh) << 32;
	uint32_t added = 0;
	while (dm) {
		const uint32_t pos = (uint32_t)__ffs((int)dm) - 1 + added;
		if (pos >= 16)
			break;
		dm &= dm - 1;
		added++;
		if (pos < 8) {
			const uint64_t lo = (1ull << (8 * pos)) - 1;
			const uint64_t carry = q0 >> 56;
			q0 = (q0 & lo) | (3ull << (8 * pos)) | ((q0 & ~lo) << 8);
			q1 = (q1 << 8) | carry;
		} else {
			const uint32_t pp = pos - 8;
			const uint64_t lo = (1ull << (8 * pp)) - 1;
			q1 = (q1 & lo) | (3ull << (8 * pp)) | ((q1 & ~lo) << 8);
		}
	}
	put_unit(dst, q0, q1, capend);
}

/* chunk c of a byte-wise row: payload starts (out_off, start code), 03s and bytes one by one;
 * d0 = inserts before the span + sc_len * payloads before the span */
template <int ROWS>
__device__ __noinline__ void bytewise_chunk(const Buf<ROWS> s, const uint16_t *E, const FrameArgs &a, uint32_t c,
					       uint64_t span_off, uint32_t nvalid, uint64_t k_lo, uint64_t k_hi, uint64_t d0)
{
	const uint32_t p0 = c * 16;
	if (p0 >= nvalid)
		return;
	const uint32_t nv = nvalid - p0 >= 16u ? 16u : nvalid - p0;
	const uint32_t nb = (uint32_t)(k_hi - k_lo);
	/* payloads of the span that started before the chunk */
	uint32_t next = (nb && p0) ? starts_le<ROWS>(s, a, span_off, k_lo, k_hi, p0 - 1) : 0u;
	uint64_t pos = span_off + p0 + d0 + (uint64_t)E[c] + a.sc_len * next;
	const uint8_t *capend = a.out + a.out_cap;
	const uint8_t *rawb = s.d.raw + 16 + p0;
	const uint32_t m = s.m.M[c];
	uint32_t noff = next < nb ? start_at<ROWS>(s, a, span_off, k_lo, k_hi, next) : 0xffffffffu;
	for (uint32_t j = 0; j < nv; j++) {
		while (noff == p0 + j) {
			a.out_off[k_lo + next] = pos;
			for (uint32_t b = 0; b < a.sc_len; b++)
				put_byte(a.out + pos + b, b + 1 == a.sc_len ? 1 : 0, capend);
			pos += a.sc_len;
			next++;
			noff = next < nb ? start_at<ROWS>(s, a, span_off, k_lo, k_hi, next) : 0xffffffffu;
		}
		if ((m >> j) & 1)
			put_byte(a.out + pos++, 3, capend);
		put_byte(a.out + pos++, rawb[j], capend);
	}
}

/* the span the input ends in (once per launch): plain loads, zero fill past the end */
template <int ROWS>
__device__ __noinline__ void load_partial_span(const Buf<ROWS> s, const FrameArgs &a, uint64_t span_off, uint32_t lane)
{
	uint32_t *raw32 = (uint32_t *)(s.d.raw + 16);
	for (int c = (int)lane - 1; c < Cfg<ROWS>::SPAN_CH; c += 32) {
		const int64_t o = (int64_t)span_off + (int64_t)c * 16;
		uint32_t w[4] = {0, 0, 0, 0};
		if (o >= 0) {
			for (int b = 0; b < 16; b++)
				if ((uint64_t)(o + b) < a.len)
					w[b >> 2] |= (uint32_t)a.rbsp[o + b] << (8 * (b & 3));
		} else {
			w[3] = 0xffffffffu; /* nothing before the input */
		}
		*(uint4 *)(raw32 + 4 * c) = make_uint4(w[0], w[1], w[2], w[3]);
	}
}

/* phase timestamps of a span (diagnostics: H264GPU_FRAME_TRACE) */
__device__ __forceinline__ void trace_mark(const FrameArgs &a, uint32_t t, uint32_t lane, int k, uint64_t extra = 0)
{
#ifndef H264_EMU
	if (a.trace != NULL && lane == 0 && t < a.num_tiles) {
		uint64_t ns;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
		a.trace[(uint64_t)t * 8 + k] = k == 7 ? extra : ns;
	}
#else
	(void)a; (void)t; (void)lane; (void)k; (void)extra;
#endif
}

/* ticket of the next span, every lane gets it */
__device__ __forceinline__ uint32_t take_ticket(const FrameArgs &a, uint32_t lane)
{
	uint32_t t = 0;
	if (lane == 0)
		t = atomicAdd(a.ticket, 1u);
	return __shfl_sync(FULL_MASK, t, 0);
}

/* P0: the span's bulk copy (with the 16 bytes before and, when the input has them, after it) */
template <int ROWS>
__device__ __forceinline__ void copy_issue(const Buf<ROWS> s, const FrameArgs &a, uint32_t t, uint32_t lane)
{
	using C = Cfg<ROWS>;
	const uint64_t span_off = (uint64_t)t * C::SPAN;
	if (lane == 0 && span_off + (uint64_t)C::SPAN <= a.len) {
		const uint32_t right = span_off + (uint64_t)C::SPAN + 16 <= a.len ? 16u : 0u;
		if (t > 0) {
			bulk_load_issue(s.d.raw, a.rbsp + span_off - 16, C::SPAN + 16 + right, &s.d.bar);
		} else {
			((uint32_t *)(s.d.raw + 16))[-1] = 0xffffffffu; /* nothing before the input */
			bulk_load_issue(s.d.raw + 16, a.rbsp, C::SPAN + right, &s.d.bar);
		}
	}
}

/* ... and its arrival (the span the input ends in: plain loads) */
template <int ROWS>
__device__ __forceinline__ void copy_wait(const Buf<ROWS> s, const FrameArgs &a, uint32_t t, uint32_t lane, uint32_t &par,
					  uint32_t bufi)
{
	using C = Cfg<ROWS>;
	const uint64_t span_off = (uint64_t)t * C::SPAN;
	if (span_off + (uint64_t)C::SPAN <= a.len) {
		bulk_load_wait_parity(&s.d.bar, (par >> bufi) & 1u);
		par ^= 1u << bufi;
	} else {
		load_partial_span<ROWS>(s, a, span_off, lane);
	}
	__syncwarp();
}

template <int ROWS> __device__ __forceinline__ void clear_masks(const Buf<ROWS> s, uint32_t lane)
{
	uint4 *m4 = (uint4 *)s.m.M;
	for (uint32_t i = lane; i < (uint32_t)(Cfg<ROWS>::SPAN_CH + 8) / 8; i += 32)
		m4[i] = make_uint4(0, 0, 0, 0);
}

/*
 * P1: wait for the bytes, stage the payload starts, find the insert masks.  Fills r (t must be
 * set) and returns with the span's insert count in r.count.
 */
template <int ROWS>
__device__ __forceinline__ void classify(const Buf<ROWS> s, uint16_t *cand, const FrameArgs &a, SpanRegs &r, uint32_t lane,
					 uint32_t &par, uint32_t bufi)
{
	using C = Cfg<ROWS>;
	const uint32_t ltmask = (1u << lane) - 1u;
	uint32_t *raw32 = (uint32_t *)(s.d.raw + 16);
	const uint8_t *rawb = s.d.raw + 16;
	const uint32_t t = r.t;
	const uint64_t span_off = (uint64_t)t * C::SPAN;
	const uint32_t nvalid = span_off >= a.len ? 0u
						  : (a.len - span_off >= (uint64_t)C::SPAN ? (uint32_t)C::SPAN
											   : (uint32_t)(a.len - span_off));
	const bool full = nvalid == (uint32_t)C::SPAN;
	const uint64_t k_lo = a.first[t], k_hi = a.first[t + 1];
	const bool has_b = k_hi > k_lo;
	r.k_lo = k_lo;
	r.k_hi = k_hi;
	if (lane < (uint32_t)kSoff && k_lo + lane < k_hi)
		s.m.soff[lane] = (uint32_t)(a.off[k_lo + lane] - span_off);
	copy_wait<ROWS>(s, a, t, lane, par, bufi);
	/* payload starts of the span before row `lane` (entry ROWS: all of them); byte-wise rows: a
	 * payload start inside, or the input ends in (or before) the row */
	uint32_t kr = 0;
	if (has_b && lane >= 1 && lane <= (uint32_t)ROWS)
		kr = starts_le<ROWS>(s, a, span_off, k_lo, k_hi, lane * 512u - 1u);
	if (lane <= (uint32_t)ROWS)
		s.m.krow[lane] = kr;
	const uint32_t krn = __shfl_down_sync(FULL_MASK, kr, 1);
	const uint32_t BW = __ballot_sync(FULL_MASK, lane < (uint32_t)ROWS && (krn > kr || (lane + 1) * 512u > nvalid));
	/*
	 * Soft seams.  A unit that straddles two spans is written whole by the span it starts in when
	 * neither the last row of that span nor the first row of the next one goes byte by byte (no
	 * payload start, no end of input there): the left span has the next 16 source bytes (right
	 * halo) and works out their insert mask itself (chunk SPAN_CH).  Both sides evaluate the same
	 * condition from off[]; otherwise the seam is hard and each side writes its own bytes of the
	 * unit one by one.
	 */
	const bool soft_before = t > 0 && !(BW & 1u) && !(k_lo > 0 && a.off[k_lo - 1] + 512u >= span_off);
	const bool soft_after = full && !((BW >> (ROWS - 1)) & 1u) && span_off + (uint64_t)C::SPAN + 16 <= a.len &&
				a.off[k_hi] >= span_off + (uint64_t)C::SPAN + 512u;
	r.bw = BW | (soft_before ? kSoftBefore : 0u) | (soft_after ? kSoftAfter : 0u);
	/* the zero run that reaches the span from the left, parity-faithful code: 0, 1, 2 = even >= 2,
	 * 3 = odd >= 3; it stops at the last payload start before the span */
	uint32_t zt = 0;
	if (t > 0 && !(has_b && a.off[k_lo] == span_off)) {
		const uint64_t lim = k_lo ? a.off[k_lo - 1] : 0;
		const uint32_t nz = __ballot_sync(FULL_MASK, lane < 16 && rawb[-1 - (int)lane] != 0);
		uint64_t z = nz ? (uint64_t)(__ffs((int)nz) - 1) : 16u;
		if (z > span_off - lim)
			z = span_off - lim;
		if (z == 16 && span_off - lim > 16) { /* (3/16)^16 of the spans of a random stream */
			uint64_t p = span_off - 16;
			while (p > lim && a.rbsp[p - 1] == 0)
				p--;
			z = span_off - p;
		}
		zt = z < 2 ? (uint32_t)z : 2u + (uint32_t)(z & 1);
	}

	/* candidates (some byte <= 3 after two zero bytes), then the exact insert mask for them */
	const uint32_t k1 = 0x01010101u, kfc = 0xfcfcfcfcu;
	uint32_t ntot = 0;
#pragma unroll 2
	for (int i = 0; i < ROWS; i++) {
		const uint32_t c = i * 32 + lane;
		const uint4 v = *(const uint4 *)(raw32 + 4 * c);
		const uint32_t pw = raw32[4 * (int)c - 1];
		const uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
		const uint32_t X0 = __funnelshift_l(pw, w0, 16) | __funnelshift_l(pw, w0, 8) | (w0 & kfc);
		const uint32_t X1 = __funnelshift_l(w0, w1, 16) | __funnelshift_l(w0, w1, 8) | (w1 & kfc);
		const uint32_t X2 = __funnelshift_l(w1, w2, 16) | __funnelshift_l(w1, w2, 8) | (w2 & kfc);
		const uint32_t X3 = __funnelshift_l(w2, w3, 16) | __funnelshift_l(w2, w3, 8) | (w3 & kfc);
		const uint32_t acc = ((X0 - k1) & ~X0) | ((X1 - k1) & ~X1) | ((X2 - k1) & ~X2) | ((X3 - k1) & ~X3);
		const bool hit = (acc & 0x80808080u) != 0;
		const uint32_t bal = __ballot_sync(FULL_MASK, hit);
		if (hit)
			cand[ntot + (uint32_t)__popc(bal & ltmask)] = (uint16_t)c;
		ntot += (uint32_t)__popc(bal);
	}
	if (soft_after) { /* the chunk after the span: its mask decides how the last chunk's unit is built */
		if (lane == 0)
			cand[ntot] = (uint16_t)C::SPAN_CH;
		ntot++;
	}
	__syncwarp();
	uint32_t cnt = 0;
	for (uint32_t q = lane; q < ntot; q += 32) {
		const uint32_t c = cand[q];
		const uint32_t p0 = c * 16;
		const uint32_t R = c >> 5;
		const uint32_t vm = full ? 0xffffu : valid16(p0, nvalid);
		/* payload starts: the last one at or before the chunk, those inside it */
		uint32_t B16 = 0;
		int32_t lim = -1;
		if (has_b) {
			uint32_t cnt0 = s.m.krow[R];
			if ((BW >> R) & 1) {
				cnt0 = starts_le<ROWS>(s, a, span_off, k_lo, k_hi, p0);
				const uint32_t cnt1 = starts_le<ROWS>(s, a, span_off, k_lo, k_hi, p0 + 15);
				for (uint32_t n = cnt0; n < cnt1; n++)
					B16 |= 1u << (start_at<ROWS>(s, a, span_off, k_lo, k_hi, n) - p0);
			}
			if (cnt0)
				lim = (int32_t)start_at<ROWS>(s, a, span_off, k_lo, k_hi, cnt0 - 1);
		}
		/* zero run that reaches the chunk from the left */
		const uint32_t lo = lim >= 0 ? (uint32_t)lim : 0u;
		uint32_t p = p0;
		while (p >= lo + 16) {
			const uint4 u = *(const uint4 *)(raw32 + (p >> 2) - 4);
			if (u.x | u.y | u.z | u.w)
				break;
			p -= 16;
		}
		while (p > lo && rawb[p - 1] == 0)
			p--;
		uint32_t zin = p0 - p;
		if (p == 0 && lim < 0)
			zin += zt;
		const uint4 v = *(const uint4 *)(raw32 + 4 * c);
		const uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
		uint32_t ins16 = 0;
		if (B16 == 0) {
			const uint32_t Z16 = zmask4(w0) | zmask4(w1) << 4 | zmask4(w2) << 8 | zmask4(w3) << 12;
			const uint32_t L16 = zmask4(w0 & kfc) | zmask4(w1 & kfc) << 4 | zmask4(w2 & kfc) << 8 |
					     zmask4(w3 & kfc) << 12;
			const uint32_t Zext = Z16 << 2 | (zin >= 1 ? 2u : 0u) | (zin >= 2 ? 1u : 0u);
			uint32_t pc = L16 & (Zext >> 1) & Zext & 0xffffu;
			while (pc) {
				const uint32_t j = (uint32_t)__ffs((int)pc) - 1;
				pc &= pc - 1;
				const uint32_t n1 = (uint32_t)__clz((int)~(Zext << (30 - j)));
				const uint32_t z = n1 < j + 2 ? n1 : j + zin;
				if (!(z & 1))
					ins16 |= 1u << j;
			}
		} else {
			ins16 = inserts_with_starts(w0, w1, w2, w3, B16, zin);
		}
		ins16 &= vm;
		if (ins16)
			s.m.M[c] = (uint16_t)ins16;
		if (c < (uint32_t)C::SPAN_CH) /* the next span counts its own */
			cnt += (uint32_t)__popc(ins16);
	}
	__syncwarp();
	r.count = warp_add(cnt);
}

/* the span's words of the chain: its own, its group's and, from the warp that completes the
 * group, the supergroup's (lane 0) */
__device__ __forceinline__ void publish(const FrameArgs &a, uint32_t t, uint32_t count)
{
	st_relaxed_u64(a.desc + t, kValid | (uint64_t)count);
	const uint32_t g = t >> 5;
	const uint64_t old = (uint64_t)atomicAdd((unsigned long long *)(a.group_w + g), (unsigned long long)(kOne | count));
	if ((old >> 40) == 31u)
		atomicAdd((unsigned long long *)(a.super_w + (g >> 5)),
			  (unsigned long long)(kOne | ((old & kSumMask) + count)));
}

/* the three probes of a look-back, fetched together (and usually an iteration's classification
 * ahead of their use) */
struct Probe {
	uint64_t w1, w2, p;
};

__device__ __forceinline__ Probe lb_fetch(const FrameArgs &a, uint32_t t, uint32_t lane)
{
	const uint32_t i = t & 31u, g = t >> 5, gi = g & 31u, q = g >> 5;
	Probe r;
	r.w1 = lane < i ? ld_relaxed_u64(a.desc + (t - i) + lane) : kValid;
	r.w2 = lane < gi ? ld_relaxed_u64(a.group_w + (g - gi) + lane) : 32ull << 40;
	r.p = lane == 0 ? ld_relaxed_u64(a.super_p + q) : 0;
	return r;
}

/* inserts before supergroup q (> 0): back over the supergroup sums to the nearest known prefix */
__device__ __noinline__ uint64_t resolve_super(const FrameArgs &a, uint32_t q, uint32_t lane)
{
	uint64_t acc = 0;
	int64_t j0 = (int64_t)q - 1;
	uint32_t nap = kNapMin;
	for (;;) {
		const int64_t j = j0 - (int64_t)lane;
		const uint64_t p = j >= 0 ? ld_relaxed_u64(a.super_p + j) : 0;
		const uint64_t w = j >= 0 ? ld_relaxed_u64(a.super_w + j) : 32ull << 40;
		const uint32_t pm = __ballot_sync(FULL_MASK, p != 0);
		const uint32_t okm = __ballot_sync(FULL_MASK, (w >> 40) == 32u);
		const int fp = pm ? __ffs((int)pm) - 1 : 32;
		const uint32_t upto = fp < 31 ? (2u << fp) - 1u : 0xffffffffu;
		if ((okm & upto) != upto) {
			spin_pause(nap);
			nap = nap < kNapMax ? nap * 2 : nap;
			continue;
		}
		acc += (uint64_t)warp_add((int)lane <= fp ? (uint32_t)(w & 0xffffffffu) : 0u);
		if (fp < 32) {
			const uint64_t pv = __shfl_sync(FULL_MASK, p, fp);
			acc += pv - 1;
			break;
		}
		j0 -= 32;
	}
	if (lane == 0)
		st_relaxed_u64(a.super_p + q, acc + 1);
	return acc;
}

/* inserts before span t */
__device__ __forceinline__ uint64_t lb_eval(const FrameArgs &a, uint32_t t, uint32_t lane, Probe r, uint32_t &spins)
{
	const uint32_t i = t & 31u, g = t >> 5, gi = g & 31u, q = g >> 5;
	/* a poll that misses sleeps: every waiting span of a supergroup reads the same few words, and
	 * the publishers' atomics queue behind those reads in the same L2 slice */
	uint32_t nap = kNapMin;
	while (!__all_sync(FULL_MASK, (r.w1 & kValid) != 0)) {
		spins += 1u;
		spin_pause(nap);
		nap = nap < kNapMax ? nap * 2 : nap;
		if (lane < i)
			r.w1 = ld_relaxed_u64(a.desc + (t - i) + lane);
	}
	while (!__all_sync(FULL_MASK, (r.w2 >> 40) == 32u)) {
		spins += 1u << 10;
		spin_pause(nap);
		nap = nap < kNapMax ? nap * 2 : nap;
		if (lane < gi)
			r.w2 = ld_relaxed_u64(a.group_w + (g - gi) + lane);
	}
	const uint32_t s12 = warp_add(((uint32_t)r.w1 & 0xffffffu) + (uint32_t)(r.w2 & 0xffffffffu));
	/* inserts before the supergroup: resolved by its first spans (they get here first), read as
	 * one word by the others, who resolve on their own only after kPollSuper misses; one lane's
	 * copy decides for the warp (the branch holds collectives) */
	uint64_t p0 = __shfl_sync(FULL_MASK, r.p, 0);
	if (p0 == 0 && (t & 1023u) >= (uint32_t)kResolvers) {
		for (int n = 0; n < kPollSuper && p0 == 0; n++) {
			spins += 1u << 20;
			spin_pause(nap);
			nap = nap < kNapMax ? nap * 2 : nap;
			uint64_t v = 0;
			if (lane == 0)
				v = ld_relaxed_u64(a.super_p + q);
			p0 = __shfl_sync(FULL_MASK, v, 0);
		}
	}
	if (p0 == 0)
		spins |= 1u << 31;
	const uint64_t pq = p0 ? p0 - 1 : resolve_super(a, q, lane);
	return pq + (uint64_t)s12;
}

/*
 * The rows of a span.  bwl: bit 0 = a hard seam at the span start, bits 1..ROWS = own rows that go
 * byte by byte, bit ROWS + 1 = a hard seam at the span end.  LEAN (most spans): no byte-wise row, hence no payload start, in the
 * span.  base = where source position 0 of the span goes at E = 0 without the start codes of
 * the span, shl = its low bits.  Returns the number of chunks listed for the byte-exact pass.
 */
template <int ROWS, bool LEAN>
__device__ __forceinline__ uint32_t emit_span(const Buf<ROWS> s, const uint16_t *E, uint8_t *dl, const FrameArgs &a,
					      uint32_t lane, uint32_t bwl, uint8_t *base, uint32_t shl, bool has_b, bool safe,
					      const uint8_t *capend)
{
	const uint32_t ltmask = (1u << lane) - 1u;
	const uint32_t *raw32 = (const uint32_t *)(s.d.raw + 16);
	uint32_t ndirty = 0;
#pragma unroll(LEAN ? ROWS : 1)
	for (int i = 0; i < ROWS; i++) {
		if (!LEAN && ((bwl >> (i + 1)) & 1))
			continue;
		const uint32_t koff = (!LEAN && has_b) ? a.sc_len * s.m.krow[i] : 0u;
		const uint32_t c = i * 32 + lane;
		const uint32_t p0 = c * 16;
		const uint32_t e = E[c];
		const uint32_t m = s.m.M[c];
		const uint32_t mn = s.m.M[c + 1];
		uint8_t *o = base + (p0 + e + koff); /* the chunk's first output byte */
		const uint32_t b = (0u - (shl + e + koff)) & 15u;
		uint32_t bad = m | (mn & ((1u << b) - 1u));
		if ((i == 0 || !LEAN) && ((bwl >> i) & 1) && lane == 0 && b)
			bad = 1; /* the bytes before the row's first unit: written in the byte-exact pass */
		if ((i == ROWS - 1 || !LEAN) && ((bwl >> (i + 2)) & 1) && lane == 31 && (b | m))
			bad = 1; /* the unit runs over a seam */
		if (!bad) {
			const uint32_t S = p0 + b;
			const uint32_t wi = S >> 2, sh = (S & 3) * 8;
			const uint32_t y0 = raw32[wi], y1 = raw32[wi + 1], y2 = raw32[wi + 2];
			const uint32_t y3 = raw32[wi + 3], y4 = raw32[wi + 4];
			const uint4 val = make_uint4(__funnelshift_r(y0, y1, sh), __funnelshift_r(y1, y2, sh),
						     __funnelshift_r(y2, y3, sh), __funnelshift_r(y3, y4, sh));
			if (safe || o + b + 16 <= capend)
				stg_stream16_free(o + b, val);
			else
				put_unit(o + b, (uint64_t)val.x | (uint64_t)val.y << 32, (uint64_t)val.z | (uint64_t)val.w << 32,
					 capend);
		}
		const uint32_t bal = __ballot_sync(FULL_MASK, bad != 0);
		if (bad)
			dl[ndirty + (uint32_t)__popc(bal & ltmask)] = (uint8_t)(i * 32 + lane);
		ndirty += (uint32_t)__popc(bal);
	}
	return ndirty;
}

/* P2..P4 of a classified span: inserts before every chunk, the rows, the byte-exact pass.
 * pin = inserts before the span.  The ticket of the span after next is asked for on the way
 * (next: the atomic's round trip hides behind the rows). */
template <int ROWS>
__device__ __forceinline__ void emit(uint16_t *E, uint8_t *dl, const Buf<ROWS> s, const FrameArgs &a, const SpanRegs &r,
				     uint64_t pin, uint32_t lane, uint32_t &tnext, bool early)
{
	using C = Cfg<ROWS>;
	const uint32_t t = r.t;
	const uint64_t span_off = (uint64_t)t * C::SPAN;
	const uint32_t nvalid = span_off >= a.len ? 0u
						  : (a.len - span_off >= (uint64_t)C::SPAN ? (uint32_t)C::SPAN
											   : (uint32_t)(a.len - span_off));
	const uint64_t k_lo = r.k_lo, k_hi = r.k_hi;
	const bool has_b = k_hi > k_lo;
	const uint32_t BW = r.bw & 0xffffu;
	const bool hard_before = !(r.bw & kSoftBefore), hard_after = !(r.bw & kSoftAfter);
	const uint8_t *capend = a.out + a.out_cap;

	/* inserts of the span before every chunk (a lane sums ROWS consecutive chunks) */
	{
		uint32_t ex[ROWS];
		uint32_t run = 0;
		const uint16_t *mp = s.m.M + lane * ROWS;
#pragma unroll
		for (int k = 0; k < ROWS; k++) {
			ex[k] = run;
			run += (uint32_t)__popc(mp[k]);
		}
		uint32_t inc = run;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
			if (lane >= (uint32_t)d)
				inc += o;
		}
		const uint32_t base = inc - run;
		uint16_t *ep = E + lane * ROWS;
#pragma unroll
		for (int k = 0; k < ROWS; k++)
			ep[k] = (uint16_t)(base + ex[k]);
		__syncwarp();
	}
	if (lane == 0 && t == a.num_tiles - 1)
		finish_output(a, k_hi, pin + r.count);

	/* the span's whole output, with a unit of slack, inside the capacity: no store needs a check */
	const bool safe = span_off + (uint64_t)(C::SPAN + C::SPAN / 2 + 32) + pin + a.sc_len * k_hi <= a.out_cap;
	const uint64_t d0 = pin + a.sc_len * k_lo;
	const uint32_t bwl = BW << 1 | (hard_before ? 1u : 0u) | (hard_after ? 1u << (ROWS + 1) : 0u);
	uint8_t *base = a.out + (span_off + d0);
	uint32_t ndirty;
	if (BW == 0)
		ndirty = emit_span<ROWS, true>(s, E, dl, a, lane, bwl, base, (uint32_t)d0, has_b, safe, capend);
	else
		ndirty = emit_span<ROWS, false>(s, E, dl, a, lane, bwl, base, (uint32_t)d0, has_b, safe, capend);
	__syncwarp();
	/* the ticket asked for before the rows is back: its span on the way into L2 under the byte-exact pass */
	if (early) {
		tnext = __shfl_sync(FULL_MASK, tnext, 0);
		if (lane == 0 && tnext < a.num_tiles && (uint64_t)(tnext + 1) * C::SPAN <= a.len)
			l2_prefetch(a.rbsp + (uint64_t)tnext * C::SPAN, C::SPAN);
		trace_mark(a, tnext, lane, 0);
	}
	for (uint32_t x = BW; x; x &= x - 1)
		bytewise_chunk<ROWS>(s, E, a, ((uint32_t)__ffs((int)x) - 1) * 32 + lane, span_off, nvalid, k_lo, k_hi, d0);
	const uint8_t *cap2 = safe ? (const uint8_t *)~(uintptr_t)0 : capend;
	for (uint32_t g = lane; g < ndirty; g += 32) {
		const uint32_t c = dl[g];
		const uint32_t R = c >> 5;
		const uint64_t rd = d0 + (has_b ? a.sc_len * s.m.krow[R] : 0u);
		const uint32_t e = E[c];
		const uint32_t kins = (uint32_t)__popc(s.m.M[c]);
		uint8_t *o = a.out + (span_off + rd + (uint64_t)c * 16 + e);
		const uint32_t b = (0u - ((uint32_t)rd + e)) & 15u;
		const bool seam_after = (c & 31u) == 31u && (R == (uint32_t)ROWS - 1 ? hard_after : ((BW >> (R + 1)) & 1) != 0);
		if ((c & 31u) == 0 && b && (R == 0 ? hard_before : ((BW >> (R - 1)) & 1) != 0))
			chunk_bytes<ROWS>(s, c, 0, b, o, capend); /* the unit began on the other side of a seam */
		if (seam_after) {
			chunk_bytes<ROWS>(s, c, b, 16 + kins, o, capend);
		} else {
			/* one unit, or two when the inserts push a second boundary into the chunk */
#pragma unroll 1
			for (uint32_t ub = b; ub < 16 + kins; ub += 16)
				gen_unit<ROWS>(s, c, ub, o + ub, cap2);
		}
	}
	__syncwarp(); /* E, dl and the buffer are reused */
}

/*
 * NW warps per CTA, every warp on its own.  Per loop iteration (A = the span classified an
 * iteration ago, B = the span whose ticket was just taken):
 *   bulk copy of B | look-back words of A fetched | B classified and published |
 *   (NBUF = 1: bulk copy of A, again) | look-back of A evaluated | A emitted | next ticket.
 * The ticket is taken AFTER the look-back, the only place a warp can wait, and after the emit
 * unless that is short: a warp never holds an unpublished span for long.  (A first build took it
 * at the top of the iteration: every wait was then handed on, plus a polling delay, to the warps
 * that needed the held span, and the waits grew linearly along the stream: 16 us per look-back
 * in the first eighth of 1 GiB, 256 us in the last, profiles/r02_frame7_trace_ticket_at_top.txt.)
 */
template <int ROWS, int NW, int MINB, int NBUF>
__global__ void __launch_bounds__(32 * NW, MINB) frame7_kernel(const FrameArgs a)
{
	__shared__ WSmem<ROWS, NBUF> sm[NW];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	WSmem<ROWS, NBUF> &ws = sm[warp];
	if (lane == 0) {
#pragma unroll
		for (int i = 0; i < NBUF; i++)
			bulk_bar_init(&ws.d[i].bar);
	}
	__syncwarp();
	uint32_t par = 0; /* bit i: phase parity of buffer i's barrier */
	SpanRegs A, B;
	A.t = 0xffffffffu;
	uint32_t tB = take_ticket(a, lane);
	trace_mark(a, tB, lane, 0);
	uint32_t cur = 0; /* masks (and, NBUF = 2, bytes) of A; B goes to the other slot */
	for (;;) {
		const bool hasA = A.t != 0xffffffffu;
		const bool hasB = tB < a.num_tiles;
		if (!hasA && !hasB)
			break;
		const uint32_t da = NBUF == 2 ? cur : 0u, db = NBUF == 2 ? cur ^ 1u : 0u;
		const Buf<ROWS> bufA = {ws.d[da], ws.m[cur]}, bufB = {ws.d[db], ws.m[cur ^ 1u]};
		if (hasB) {
			trace_mark(a, tB, lane, 1);
			copy_issue<ROWS>(bufB, a, tB, lane);
			clear_masks<ROWS>(bufB, lane);
		}
		Probe pr;
		if (hasA && NBUF == 2)
			pr = lb_fetch(a, A.t, lane);
		if (hasB) {
			B.t = tB;
			classify<ROWS>(bufB, ws.E, a, B, lane, par, db);
			trace_mark(a, tB, lane, 2);
			if (lane == 0)
				publish(a, tB, B.count);
			trace_mark(a, tB, lane, 3);
		}
		uint32_t tC = 0xffffffffu;
		if (hasA) {
			uint32_t spins = 0;
			if (NBUF == 1) { /* the bytes of A again, under its look-back */
				__syncwarp();
				copy_issue<ROWS>(bufA, a, A.t, lane);
				pr = lb_fetch(a, A.t, lane);
			}
			trace_mark(a, A.t, lane, 4);
			const uint64_t pin = lb_eval(a, A.t, lane, pr, spins);
			trace_mark(a, A.t, lane, 5);
			if (NBUF == 1)
				copy_wait<ROWS>(bufA, a, A.t, lane, par, da);
			/* the next ticket before the emit (its span prefetched into L2 on the way) only when the
			 * emit is short and uniform: a span with byte-wise rows can take 30 us, and a ticket held
			 * unpublished that long makes every later span wait (kTicketEarly: 0 never, 1 when A has
			 * no byte-wise row, 2 always) */
			const bool early = hasB && (kTicketEarly == 2 || (kTicketEarly == 1 && (A.bw & 0xffffu) == 0));
			if (early && lane == 0)
				tC = atomicAdd(a.ticket, 1u);
			emit<ROWS>(ws.E, ws.dl, bufA, a, A, pin, lane, tC, early);
			trace_mark(a, A.t, lane, 6);
			trace_mark(a, A.t, lane, 7, spins);
			if (!early && hasB) {
				tC = take_ticket(a, lane);
				trace_mark(a, tC, lane, 0);
			}
		} else {
			tC = take_ticket(a, lane);
			trace_mark(a, tC, lane, 0);
		}
		if (!hasB)
			break;
		A = B;
		tB = tC;
		cur ^= 1u;
	}
}

} /* namespace frame7 */

#endif /* ANNEXB_FRAME7_CUH */
