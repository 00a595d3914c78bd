/*
 * annexb_scan7.cuh — K1/K2 seventh generation: the in-place layout (NAL k's RBSP starts where
 * the NAL starts; out(p) = p - #EPB in [S(p), p)) with WARP-AUTONOMOUS SPANS.
 *
 * Why: gen 6 (block-wide 32 KiB tiles, warps meeting at four barriers per tile) sat at 58 % of
 * its issue slots with 8.2 warp-cycles of barrier stall per issued instruction
 * (profiles/r01_scan_kernel_raw.csv): one look-back per tile behind a barrier, and every phase
 * ending at the pace of the slowest of eight warps.  Here NOTHING is block-wide.  A warp takes a
 * 4 KiB span by ticket, pulls it into its own slice of shared memory with its own bulk copy and
 * mbarrier, classifies, publishes the span's aggregate, looks back on its own, emits its rows
 * and runs its own byte-exact pass.  A warp that waits (bulk load, look-back) costs one of the
 * forty resident warps of the SM, not a whole CTA, and there is no __syncthreads in the kernel.
 *
 * TICKETS GO ROUND ROBIN OVER K REGIONS of the stream.  With tickets in stream order a span's
 * look-back needs the words of ~16 predecessors that took their tickets within nanoseconds of
 * it, i.e. it waits for the slowest of them: 4.8 us of a 12.8 us span life
 * (profiles/r02_scan_phase_trace.txt).  With ticket -> (region tk % K, index tk / K) the
 * predecessor of a span was taken K tickets earlier (K = 1480: ~3 us, the time from a ticket to
 * its chain word), so the look-back finds it published, usually already as a PFX.  The spans at
 * the head of a region (up to its first start code) would have to wait for the END of the
 * region before them: their look-back does not block but defers them to a second, small launch
 * of the same kernel over the deferred list (~17 spans per region), by which time every chain
 * word exists.  1719 -> 2136 GB/s at 4 GiB.
 *
 * TICKETS COME FROM SEVERAL COUNTERS (take_ticket, kTickBase): with one ticket word the kernel
 * ran at the rate same-address atomics are served by the L2 slice that owns the word, one per
 * 1.9 ns at best = 4 KiB / 1.9 ns = 2140 GB/s, one per 2.6 - 2.9 ns on another slice.
 *
 * Measured and NOT kept -- all of it WITH ONE TICKET WORD, i.e. under that bound (3 KiB spans
 * cost 4/3 of the tickets: the 1687 below is 2136 x 3/4 + a little; worth measuring again):
 * a CTA-level ticket pool (spans
 * reserved early are published late: 1636 GB/s), 3 KiB spans at 48 warps per SM (1687), two
 * 16-byte loads + word select instead of five conflicting 4-byte loads (no change: the kernel is
 * not bound by a pipe), and running the emit phase one loop iteration after the classify phase
 * with the masks parked in L2 (look-back wait 0.95 us, but the second read of the span misses
 * L2 59 % of the time and the larger code misses the instruction cache: 1760 GB/s).
 *
 * Chain descriptors are ONE 64-bit word per span, tagged with a 16-bit epoch of the launch
 * (nothing to clear between launches):  epoch << 48 | pfx << 41 | value.
 *   AGG (pfx = 0)  value = EPBs of the span (a span without a reset point)
 *   PFX (pfx = 1)  value = shift at the span end = EPBs since the last reset point
 * A span with a reset publishes PFX at once; a span without one publishes AGG, looks back to
 * the nearest PFX (32 predecessors per probe) and then publishes its own PFX.
 *
 * Boundary events carry the shift at their position (EPBs since the start of the NAL they
 * close), so the NAL table needs no EPB prefix over the stream:
 *   rbsp_len(k) = (end_k - start_k) - shift(end event of k).
 * Finalize = three small kernels (fin7_spans, fin7_order, fin7_table), no memset: the control
 * words are re-armed by the last block of fin7_table.
 *
 * scan7_only_kernel: the table without the strip (what h264_reader_parse needs): no shared
 * memory at all, 16-byte loads straight into registers, 1.0 B of traffic per input byte.
 *
 * Reference behaviour reproduced bit-exactly (Parrot-Developers/libh264):
 *   NAL loop of h264_reader_parse            src/h264_reader.c:133-140
 *   h264_find_nalu / start / end code search src/h264_bitstream.c:87-184
 *   EPB removal in h264_bs_fetch             include/h264/h264_bitstream.h:168-190
 * Closed form: SURVEY.md Appendix A.1/A.2 (boundary events, last-event-wins, local EPB test).
 */
#ifndef ANNEXB_SCAN7_CUH
#define ANNEXB_SCAN7_CUH

#include "gpu_compat.h"
#include "h264gpu.h"

namespace annexb7 {

constexpr int kT = 256;
constexpr int kW = kT / 32;
constexpr size_t kCtrlBytes = 4096; /* control words at the head of the workspace */
/* span tickets of the first pass come from up to kTickMax counters, 256 bytes apart (words
 * kTickBase + c * kTickStride of the control area): ONE counter hands out a ticket per ~1.9 ns at
 * best (same-address atomics are served one after the other by the L2 slice that owns the word:
 * 4 KiB / 1.9 ns = the 2140 GB/s the kernel ran at) and per 2.6-2.9 ns when the word lands on a
 * less lucky slice (the 'workspace placement' of profiles/r02_scan_workspace_placement.txt) */
constexpr uint32_t kTickBase = 64, kTickStride = 64, kTickMax = 15;
constexpr uint64_t kValMask = (1ull << 40) - 1;
constexpr uint64_t kPfxBit = 1ull << 41;

template <int ROWS> struct Cfg {
	static constexpr int SPAN_CH = 32 * ROWS;  /* 16-byte chunks per span */
	static constexpr int SPAN = SPAN_CH * 16;  /* bytes per span */
};

struct Scan7Args {
	const uint8_t *in;
	uint64_t len;
	uint8_t *rbsp;
	uint64_t *chain; /* one word per span (strip only) */
	uint64_t *fin;   /* one word per span: event slot | events << 32 | start codes << 48 */
	uint32_t *ctrl;  /* [0] span ticket  [1] event cursor  [2] r0  [3],[4] finalize block counters
			  * [5] deferred spans  [6] ticket of the second pass */
	uint64_t *evbuf; /* 2 words per event: position | start code << 62, shift */
	uint64_t ev_cap;
	uint32_t num_spans;
	uint32_t halo_left; /* bytes -4..-1 as a little-endian word, 0xff = none */
	uint32_t epoch;     /* 1..65535 */
	uint32_t pf_dist;   /* L2 prefetch distance in spans, 0 = off */
	uint32_t nap_max;   /* longest sleep of a look-back poll, ns */
	uint32_t regions;   /* K: tickets go round robin over K regions of region_len spans */
	uint32_t region_len;
	uint32_t pass2;     /* 1: the launch works off the deferred list */
	uint32_t tick_n;    /* ticket counters of the first pass (0 or 1: the single word ctrl[0]) */
	uint32_t *deferred; /* spans whose look-back would cross a region start before the region before it is done */
	uint8_t right[2];
	uint8_t has_right;
	uint8_t pad;
	uint64_t *trace; /* diagnostics (H264GPU_SCAN_TRACE): 8 timestamps per span, or NULL */
};

/* The next span ticket of this warp.  Counter c hands out c, c + n, c + 2n, ...: the counters
 * advance at the same pace (each serves 1/n of the resident warps), so tickets are still taken
 * in (nearly) increasing order, which is all the K-ticket lead of a span over its successor in
 * the region needs.  Progress: the smallest unfinished ticket is either being worked on (all its
 * predecessors in the stream have smaller tickets: its look-back ends) or is the next one of a
 * counter whose warps have all finished their spans and are asking for it.  With the default of
 * 8 counters = the warps of a CTA, warp w of EVERY CTA asks counter w: any one resident CTA serves
 * all counters, so this holds even when part of the grid is not resident (another kernel on the
 * GPU); more counters than that (the knob goes to 15) rely on the whole grid being resident. */
__device__ __forceinline__ uint32_t take_ticket(const Scan7Args &a, uint32_t c, uint32_t n)
{
	if (a.pass2)
		return atomicAdd(a.ctrl + 6, 1u);
	if (n <= 1)
		return atomicAdd(a.ctrl, 1u);
	return atomicAdd(a.ctrl + kTickBase + c * kTickStride, 1u) * n + c;
}

/* phase timestamps of a span (diagnostics build of the kernel only) */
template <bool TRACE> __device__ __forceinline__ void trace_mark(const Scan7Args &a, uint32_t t, uint32_t lane, int k)
{
#ifndef H264_EMU
	if (TRACE && lane == 0) {
		uint64_t ns;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
		a.trace[(uint64_t)t * 8 + k] = ns;
	}
#else
	(void)a; (void)t; (void)lane; (void)k;
#endif
}

/* the slice of shared memory a warp owns */
template <int ROWS> struct __align__(128) WSmem {
	uint8_t raw[16 + Cfg<ROWS>::SPAN + 32]; /* [left halo pad][span][pad] */
	uint16_t M[Cfg<ROWS>::SPAN_CH + 8];     /* delete mask per chunk (bit j = byte j is an EPB) */
	uint16_t E[Cfg<ROWS>::SPAN_CH];         /* EPBs of the span before the chunk; candidate list before that */
	uint8_t dl[Cfg<ROWS>::SPAN_CH];         /* chunks that take the byte-exact path */
	int32_t rbrel[ROWS];                    /* shift at the row start minus E, rows after a reset */
	uint64_t bar;
};

/* ---- byte tests on 32-bit words ----------------------------------------------------- */

/* 0x80 in every byte of x that is zero, exact */
__device__ __forceinline__ uint32_t zero_bytes_msb(uint32_t x)
{
	const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
	return ~(t | x | 0x7f7f7f7fu);
}
/* msb-per-byte mask (bits 7,15,23,31) -> 4-bit mask, bit j = byte j */
__device__ __forceinline__ uint32_t msb_to_nib(uint32_t m)
{
	return (m * 0x00204081u) >> 28;
}
__device__ __forceinline__ uint32_t zmask4(uint32_t x)
{
	return msb_to_nib(zero_bytes_msb(x));
}
/* non-zero iff x has a zero byte (cheap boolean form) */
__device__ __forceinline__ uint32_t haszero(uint32_t x)
{
	return (x - 0x01010101u) & ~x & 0x80808080u;
}

struct TMasks {
	uint32_t ev16, sc16;
};

/* exact boundary-event masks of a chunk by THIRD-byte position, from the window [-4, 16):
 * bit j of sc16: bytes j-2, j-1, j are 00 00 01; of ev16: a boundary event has its third byte
 * at j (00 00 01, or 00 00 00 not preceded by another zero) */
__device__ __forceinline__ TMasks t_masks(uint32_t pw, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
	const uint32_t k1 = 0x01010101u;
	const uint32_t Z = zmask4(pw) | zmask4(w0) << 4 | zmask4(w1) << 8 | zmask4(w2) << 12 | zmask4(w3) << 16;
	const uint32_t O = zmask4(pw ^ k1) | zmask4(w0 ^ k1) << 4 | zmask4(w1 ^ k1) << 8 | zmask4(w2 ^ k1) << 12 |
			   zmask4(w3 ^ k1) << 16;
	const uint32_t zz = (Z << 2) & (Z << 1) & (Z | O);
	const uint32_t sc = (Z << 2) & (Z << 1) & O;
	const uint32_t ev = zz & (sc | ~(Z << 3));
	TMasks m;
	m.ev16 = (ev >> 4) & 0xffffu;
	m.sc16 = (sc >> 4) & 0xffffu;
	return m;
}

/* bit j: byte j of the chunk is the 01 of a 4-byte start code 00 00 00 01 (window [-4, 16)) */
__device__ __forceinline__ uint32_t code4_mask(uint32_t pw, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
	const uint32_t k1 = 0x01010101u;
	const uint32_t Z = zmask4(pw) | zmask4(w0) << 4 | zmask4(w1) << 8 | zmask4(w2) << 12 | zmask4(w3) << 16;
	const uint32_t O = zmask4(w0 ^ k1) << 4 | zmask4(w1 ^ k1) << 8 | zmask4(w2 ^ k1) << 12 | zmask4(w3 ^ k1) << 16;
	return ((O & (Z << 1) & (Z << 2) & (Z << 3)) >> 4) & 0xffffu;
}

/* bit j: byte j of the chunk is a reset point (the three bytes before it are 00 00 01) */
__device__ __forceinline__ uint32_t reset_mask(uint32_t pw, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
	const uint32_t k1 = 0x01010101u;
#define H264_R7(lo, hi)                                                                          \
	msb_to_nib(zero_bytes_msb(__funnelshift_l(lo, hi, 24) | __funnelshift_l(lo, hi, 16) |    \
				  (__funnelshift_l(lo, hi, 8) ^ k1)))
	const uint32_t r = H264_R7(pw, w0) | H264_R7(w0, w1) << 4 | H264_R7(w1, w2) << 8 | H264_R7(w2, w3) << 12;
#undef H264_R7
	return r;
}

__device__ __forceinline__ uint32_t valid16(uint32_t p0, uint32_t nvalid)
{
	const uint32_t nv = p0 >= nvalid ? 0u : (nvalid - p0 >= 16u ? 16u : nvalid - p0);
	return (1u << nv) - 1u;
}

/* event positions t a chunk may own: q = t - 2 inside the shard (and not before it) */
__device__ __forceinline__ uint32_t event_valid16(uint32_t p0, uint32_t nvalid, bool first_of_shard)
{
	uint32_t m = valid16(p0, nvalid + 2);
	if (first_of_shard && p0 == 0)
		m &= ~3u;
	return m;
}

/* bytes of the stream at shard offset idx, with the shard-edge rules */
__device__ __forceinline__ uint32_t edge_byte(const Scan7Args &a, uint64_t idx)
{
	if (idx < a.len)
		return a.in[idx];
	const uint64_t d = idx - a.len;
	return (a.has_right && d < 2) ? (d == 0 ? a.right[0] : a.right[1]) : 0xffu;
}
__device__ __forceinline__ uint32_t edge_word(const Scan7Args &a, uint64_t idx)
{
	return edge_byte(a, idx) | edge_byte(a, idx + 1) << 8 | edge_byte(a, idx + 2) << 16 |
	       edge_byte(a, idx + 3) << 24;
}

/* the 4 bytes before shard offset off (off > 0 or the left halo) */
__device__ __forceinline__ uint32_t prev_word(const Scan7Args &a, uint64_t off)
{
	if (off == 0)
		return a.halo_left;
	if (off <= a.len)
		return ldg_u32(a.in + off - 4);
	return edge_word(a, off - 4);
}

__device__ __forceinline__ uint64_t chain_word(uint32_t epoch, bool pfx, uint64_t value)
{
	return (uint64_t)epoch << 48 | (pfx ? kPfxBit : 0ull) | (value & kValMask);
}

/*
 * Shift at the start of span t = EPBs since the last reset point before it: fold the
 * predecessors' words back to the nearest PFX.  Whole warp, 32 predecessors per probe.
 */
constexpr uint64_t kDefer = ~0ull;

/*
 * rstart: first span of the region t lies in.  Words inside the region are waited for (their
 * spans were taken earlier and are running); a word BEFORE the region that is not there yet
 * belongs to a span that is taken late (the end of the previous region): the span is deferred
 * to the second pass instead of waiting (kDefer).
 */
__device__ __forceinline__ uint64_t lookback(const Scan7Args &a, uint32_t t, uint32_t rstart, uint32_t lane)
{
	uint64_t acc = 0;
	int64_t j0 = (int64_t)t - 1;
	const uint32_t ep = a.epoch;
	uint32_t nap = 128; /* a waiting warp sleeps (the warps it waits for need the issue slots) */
	for (;;) {
		const int64_t jl = j0 - (int64_t)lane;
		const uint64_t *p = a.chain + jl;
		uint64_t w = (uint64_t)ep << 48 | kPfxBit; /* before the shard: shift 0 */
		uint32_t stop;
		for (;;) {
			if (jl >= 0)
				w = ld_relaxed_u64(p);
			const uint32_t hi = (uint32_t)(w >> 32);
			const bool valid = (hi >> 16) == ep;
			stop = __ballot_sync(FULL_MASK, valid && (hi & (uint32_t)(kPfxBit >> 32)));
			const uint32_t ok = __ballot_sync(FULL_MASK, valid);
			const uint32_t upto = stop ? ((stop & (0u - stop)) << 1) - 1u : 0xffffffffu;
			const uint32_t missing = ~ok & upto;
			if (!missing)
				break;
			if (missing & __ballot_sync(FULL_MASK, jl < (int64_t)rstart))
				return kDefer;
			spin_pause(nap);
			nap = nap * 2 < a.nap_max ? nap * 2 : a.nap_max;
		}
		const int fs = stop ? __ffs((int)stop) - 1 : 32;
		/* AGG values are small (<= SPAN / 3): one 32-bit warp sum; the PFX word has 40 bits */
		acc += (uint64_t)warp_add((int)lane < fs ? (uint32_t)w & 0xffffffu : 0u);
		if (fs < 32) {
			const uint32_t lo = __shfl_sync(FULL_MASK, (uint32_t)w, fs);
			const uint32_t hi = __shfl_sync(FULL_MASK, (uint32_t)(w >> 32) & 0xffu, fs);
			return acc + ((uint64_t)lo | (uint64_t)hi << 32);
		}
		j0 -= 32;
	}
}

/*
 * Byte-exact output of the unit chunk c is responsible for: it starts b kept bytes into the
 * chunk; 16 kept bytes are gathered from there (at most 24 source bytes: an EPB needs two zero
 * bytes before it), cut at `limit` (source position: the next byte-wise row or the span end).
 * shlo = low bits of the shift at E = 0, basep = where source position 0 of the span goes at E = 0.
 */
template <int ROWS>
__device__ __forceinline__ void dirty_chunk(const WSmem<ROWS> &s, uint32_t c, uint32_t shlo, uint8_t *basep,
					    uint32_t limit)
{
	const uint32_t p0 = c * 16;
	const uint32_t m = s.M[c];
	const uint32_t e = s.E[c];
	const uint32_t b = (shlo + e) & 15u;
	if (b + (uint32_t)__popc(m) >= 16u)
		return; /* fewer than b + 1 kept bytes: no unit starts in this chunk */
	/* position of kept byte number b: the smallest fixed point of j = b + deleted(0..j) */
	uint32_t j = b;
	for (;;) {
		const uint32_t nj = b + (uint32_t)__popc(m & ((2u << j) - 1u));
		if (nj == j)
			break;
		j = nj;
	}
	const uint32_t pos = p0 + j;
	uint8_t *g = basep + (p0 - e + b);
	const uint32_t mi = pos >> 4;
	const uint64_t mb = (uint64_t)s.M[mi] | (uint64_t)s.M[mi + 1] << 16 | (uint64_t)s.M[mi + 2] << 32;
	uint32_t dm = (uint32_t)(mb >> (pos & 15));
	/* kept bytes that are ours: the source bytes before `limit` (the squeeze below runs over
	 * the real deletions only; bytes past the limit are squeezed along but never stored) */
	const uint32_t avail = limit - pos;
	const uint32_t kept32 = avail < 32 ? avail - (uint32_t)__popc(dm & ((1u << avail) - 1u)) : 32u - (uint32_t)__popc(dm);
	const uint32_t nout = kept32 < 16u ? kept32 : 16u;
	const uint32_t *raw32 = (const uint32_t *)(s.raw + 16);
	const uint32_t wi = pos >> 2, sh = (pos & 3) * 8;
	uint32_t w[7];
#pragma unroll
	for (int k = 0; k < 7; k++)
		w[k] = raw32[wi + k];
	uint64_t q0 = (uint64_t)__funnelshift_r(w[0], w[1], sh) | (uint64_t)__funnelshift_r(w[1], w[2], sh) << 32;
	uint64_t q1 = (uint64_t)__funnelshift_r(w[2], w[3], sh) | (uint64_t)__funnelshift_r(w[3], w[4], sh) << 32;
	uint64_t q2 = (uint64_t)__funnelshift_r(w[4], w[5], sh) | (uint64_t)__funnelshift_r(w[5], w[6], sh) << 32;
	if ((dm & (dm - 1) & 0x1ffffu) == 0 && (dm & 0xffffu)) {
		/* the common case, one deleted byte k among the first 17: bytes below k stay, the
		 * rest moves down by one */
		const uint32_t k = (uint32_t)__ffs((int)dm) - 1;
		const uint64_t s0 = (q0 >> 8) | (q1 << 56), s1 = (q1 >> 8) | (q2 << 56);
		const uint64_t m0 = k < 8 ? (1ull << (8 * k)) - 1 : ~0ull;
		const uint64_t m1 = k < 8 ? 0ull : (1ull << (8 * (k - 8))) - 1;
		q0 = (q0 & m0) | (s0 & ~m0);
		q1 = (q1 & m1) | (s1 & ~m1);
	} else {
		uint32_t removed = 0;
		while (dm) {
			const uint32_t k = (uint32_t)__ffs((int)dm) - 1 - removed;
			if (k >= 16)
				break;
			dm &= dm - 1;
			removed++;
			if (k < 8) {
				const uint64_t mm = (1ull << (8 * k)) - 1;
				q0 = (q0 & mm) | ((q0 >> 8) & ~mm) | (q1 << 56);
				q1 = (q1 >> 8) | (q2 << 56);
			} else {
				const uint64_t mm = (1ull << (8 * (k - 8))) - 1;
				q1 = (q1 & mm) | ((q1 >> 8) & ~mm) | (q2 << 56);
			}
			q2 >>= 8;
		}
	}
	if (nout == 16) {
		stg_stream16(g, make_uint4((uint32_t)q0, (uint32_t)(q0 >> 32), (uint32_t)q1, (uint32_t)(q1 >> 32)));
	} else {
		/* a unit cut at a seam: g is 16-byte aligned, so 8 + 4 + 2 + 1 byte stores are aligned */
		uint32_t n = 0;
		if (nout & 8) {
			*(uint64_t *)g = q0;
			q0 = q1;
			n = 8;
		}
		if (nout & 4) {
			*(uint32_t *)(g + n) = (uint32_t)q0;
			q0 >>= 32;
			n += 4;
		}
		if (nout & 2) {
			*(uint16_t *)(g + n) = (uint16_t)q0;
			q0 >>= 16;
			n += 2;
		}
		if (nout & 1)
			g[n] = (uint8_t)q0;
	}
}

/* shift base of row i of the span */
template <int ROWS>
__device__ __forceinline__ void row_base(const WSmem<ROWS> &s, uint32_t i, uint32_t rbhas, uint64_t d0,
					 uint8_t *out_span, uint32_t &shlo, uint8_t *&basep)
{
	if ((rbhas >> i) & 1) {
		const int32_t r = s.rbrel[i];
		shlo = (uint32_t)r;
		basep = out_span - (int64_t)r;
	} else {
		shlo = (uint32_t)d0;
		basep = out_span - d0;
	}
}

/* chunk c of a byte-wise row (a reset point in the row, or the shard ends in it): every byte to
 * its place, the shift restarting at each reset point; E[c] holds the shift at the chunk start
 * as left by emit_span */
template <int ROWS>
__device__ __noinline__ void bytewise_chunk(const WSmem<ROWS> &s, uint32_t c, uint32_t rbhas, uint64_t d0,
					    uint8_t *out_span, uint32_t nvalid)
{
	const uint32_t *raw32 = (const uint32_t *)(s.raw + 16);
	const uint8_t *rawb = s.raw + 16;
	const uint32_t p0 = c * 16;
	uint32_t shlo;
	uint8_t *basep;
	row_base<ROWS>(s, c >> 5, rbhas, d0, out_span, shlo, basep);
	const uint32_t enc = s.E[c];
	const uint32_t m = s.M[c];
	const uint4 v = *(const uint4 *)(raw32 + 4 * c);
	const uint32_t rm = reset_mask(raw32[4 * (int)c - 1], v.x, v.y, v.z, v.w) & valid16(p0, nvalid);
	/* out = basec + p0 + j - cur, cur = EPBs since the last reset */
	uint8_t *basec = (enc & 0x8000u) ? out_span : basep;
	uint32_t cur = enc & 0x7fffu;
	const uint32_t nv = p0 >= nvalid ? 0u : (nvalid - p0 >= 16u ? 16u : nvalid - p0);
	for (uint32_t j = 0; j < nv; j++) {
		if ((rm >> j) & 1) {
			cur = 0;
			basec = out_span;
		}
		if ((m >> j) & 1) {
			cur++;
			continue;
		}
		basec[p0 + j - cur] = rawb[p0 + j];
	}
}

/*
 * The rows of a span.  bwl: bit 0 = the row before the span is a seam (always: the span start),
 * bits 1..ROWS = own rows that go byte by byte, bit ROWS + 1 = the row after the span is a seam
 * (always: the span end).  LEAN (4 spans in 5): no byte-wise row and no reset in the span, one
 * shift base d0 for all rows.  Returns the number of chunks listed for the byte-exact pass.
 */
template <int ROWS, bool LEAN>
__device__ __forceinline__ uint32_t emit_span(WSmem<ROWS> &s, uint32_t lane, uint32_t bwl, uint32_t rbhas,
					      uint64_t d0, uint8_t *out_span, uint32_t nvalid)
{
	const uint32_t ltmask = (1u << lane) - 1u;
	const uint32_t *raw32 = (const uint32_t *)(s.raw + 16);
	const uint8_t *rawb = s.raw + 16;
	uint32_t ndirty = 0;
#pragma unroll(LEAN ? ROWS : 1)
	for (int i = 0; i < ROWS; i++) {
		const uint32_t c = i * 32 + lane;
		const uint32_t p0 = c * 16;
		const uint32_t e = s.E[c];
		const uint32_t m = s.M[c];
		uint32_t shlo = (uint32_t)d0;
		uint8_t *basep = out_span - d0;
		if (!LEAN && ((rbhas >> i) & 1)) {
			const int32_t r = s.rbrel[i];
			shlo = (uint32_t)r;
			basep = out_span - (int64_t)r;
		}
		if (!LEAN && ((bwl >> (i + 1)) & 1)) {
			/* a row with a reset point or the shard end goes byte by byte */
			const uint4 v = *(const uint4 *)(raw32 + 4 * c);
			const uint32_t rm = reset_mask(raw32[4 * (int)c - 1], v.x, v.y, v.z, v.w) & valid16(p0, nvalid);
			const uint32_t bal = __ballot_sync(FULL_MASK, rm != 0);
			const uint32_t top = 31u - (uint32_t)__clz((int)(rm | 1u));
			const uint32_t e_last = e + (uint32_t)__popc(m & ((1u << top) - 1u));
			const uint32_t below = bal & ltmask;
			const uint32_t e_prev = __shfl_sync(FULL_MASK, e_last, below ? 31 - __clz((int)below) : 0);
			/* the bytes themselves are written in the byte-exact pass (one lane per chunk);
			 * leave it the chunk's shift: EPBs since the last reset of the row below the
			 * chunk (bit 15), or the span's count as for any other row */
			s.E[c] = (uint16_t)(below ? ((e - e_prev) | 0x8000u) : e);
			continue;
		}
		const uint32_t mn = s.M[c + 1];
		const uint32_t b = (shlo + e) & 15u;
		const uint32_t bk = b + (uint32_t)__popc(m); /* the unit starts here if nothing is deleted from here on */
		if ((i == 0 || !LEAN) && ((bwl >> i) & 1)) {
			/* the unit this row starts in began on the other side of a seam: our bytes of it,
			 * one per lane (kept byte number `lane` of the row) */
			const uint32_t bf = __shfl_sync(FULL_MASK, b, 0);
			if (lane < bf) {
				const uint32_t cf = (uint32_t)i * 32u;
				const uint32_t mm = (uint32_t)s.M[cf] | (uint32_t)s.M[cf + 1] << 16;
				uint32_t pos = lane;
				for (;;) {
					const uint32_t np = lane + (uint32_t)__popc(mm & ((2u << pos) - 1u));
					if (np == pos)
						break;
					pos = np;
				}
				basep[cf * 16 - (uint32_t)s.E[cf] + lane] = rawb[cf * 16 + pos];
			}
		}
		/* deletions inside the unit's 16 source bytes [p0 + bk, p0 + bk + 16) */
		uint32_t bad = (m >> bk) | (mn & ~(~0u << bk));
		if ((i == ROWS - 1 || !LEAN) && ((bwl >> (i + 2)) & 1) && lane == 31 && bk)
			bad = 1; /* the unit runs over a seam */
		const bool resp = bk < 16u;
		if (resp && !bad) {
			const uint32_t S = p0 + bk;
			const uint32_t wi = S >> 2, sh = (S & 3) * 8;
			const uint32_t y0 = raw32[wi], y1 = raw32[wi + 1], y2 = raw32[wi + 2];
			const uint32_t y3 = raw32[wi + 3], y4 = raw32[wi + 4];
			stg_stream16_free(basep + (p0 - e + b),
					  make_uint4(__funnelshift_r(y0, y1, sh), __funnelshift_r(y1, y2, sh),
						     __funnelshift_r(y2, y3, sh), __funnelshift_r(y3, y4, sh)));
		}
		const bool todo = resp && bad;
		const uint32_t bal = __ballot_sync(FULL_MASK, todo);
		if (todo)
			s.dl[ndirty + (uint32_t)__popc(bal & ltmask)] = (uint8_t)(i * 32 + lane);
		ndirty += (uint32_t)__popc(bal);
	}
	return ndirty;
}

/* one span, all of it, by one warp */
template <int ROWS, bool TRACE>
__device__ __forceinline__ void do_span(WSmem<ROWS> &s, const Scan7Args &a, uint32_t t, uint32_t rstart,
					uint32_t lane, uint32_t &parity, uint32_t &next, uint32_t tc, uint32_t tn)
{
	using C = Cfg<ROWS>;
	const uint32_t ltmask = (1u << lane) - 1u;
	uint32_t *raw32 = (uint32_t *)(s.raw + 16);
	const uint8_t *rawb = s.raw + 16;
	const uint64_t span_off = (uint64_t)t * C::SPAN;
	const uint32_t nvalid = span_off >= a.len ? 0u
						  : (a.len - span_off >= (uint64_t)C::SPAN ? (uint32_t)C::SPAN
											   : (uint32_t)(a.len - span_off));
	const bool full = nvalid == (uint32_t)C::SPAN;

	trace_mark<TRACE>(a, t, lane, 0);
	/* ---- P0: bulk load into the warp's slice, L2 prefetch ahead, cleared delete masks ---- */
	if (lane == 0) {
		if (full)
			bulk_load_issue(s.raw + 16, a.in + span_off, C::SPAN, &s.bar);
		/* the span that follows in the region is taken `regions` tickets from now */
		const uint64_t noff = span_off + (uint64_t)a.pf_dist * C::SPAN;
		if (a.pf_dist && noff + (uint64_t)C::SPAN <= a.len)
			l2_prefetch(a.in + noff, C::SPAN);
		raw32[-1] = prev_word(a, span_off);
	}
	{
		uint4 *m4 = (uint4 *)s.M;
		for (uint32_t i = lane; i < (uint32_t)(C::SPAN_CH + 8) / 8; i += 32)
			m4[i] = make_uint4(0, 0, 0, 0);
	}
	if (full) {
		bulk_load_wait_parity(&s.bar, parity);
		parity ^= 1u;
	} else {
		for (uint32_t c = lane; c < (uint32_t)C::SPAN_CH; c += 32) {
			const uint64_t o = span_off + (uint64_t)c * 16;
			uint4 v;
			if (o + 16 <= a.len)
				v = ldg_stream16(a.in + o);
			else
				v = make_uint4(edge_word(a, o), edge_word(a, o + 4), edge_word(a, o + 8), edge_word(a, o + 12));
			*(uint4 *)(raw32 + 4 * c) = v;
		}
	}
	__syncwarp();
	trace_mark<TRACE>(a, t, lane, 1);

	/* ---- P1: classify.  A chunk matters only if some byte <= 3 follows two zero bytes (an EPB
	 * or the third byte of a boundary event): one SIMD-in-register test per chunk finds the
	 * candidates; they are compacted and only they get the exact masks. ---- */
	uint32_t evrows = 0; /* rows with a candidate for a boundary event */
	uint32_t ecnt = 0;   /* EPBs this lane found */
	{
		const uint32_t k1 = 0x01010101u, kfc = 0xfcfcfcfcu;
		uint16_t *cand = s.E;
		uint32_t ntot = 0;
#pragma unroll 2
		for (int i = 0; i < ROWS; i++) {
			const uint32_t c = i * 32 + lane;
			const uint4 v = *(const uint4 *)(raw32 + 4 * c);
			const uint32_t pw = raw32[4 * (int)c - 1];
			const uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
			/* byte p of Xk is 0 <=> b[p-2] = b[p-1] = 0 and b[p] <= 3 */
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
		__syncwarp();
		const uint32_t k3 = 0x03030303u, kfe = 0xfefefefeu;
		for (uint32_t j = lane; j < ntot; j += 32) {
			const uint32_t c = cand[j];
			const uint4 v = *(const uint4 *)(raw32 + 4 * c);
			const uint32_t pw = raw32[4 * (int)c - 1];
			const uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
			const uint32_t AB0 = __funnelshift_l(pw, w0, 16) | __funnelshift_l(pw, w0, 8);
			const uint32_t AB1 = __funnelshift_l(w0, w1, 16) | __funnelshift_l(w0, w1, 8);
			const uint32_t AB2 = __funnelshift_l(w1, w2, 16) | __funnelshift_l(w1, w2, 8);
			const uint32_t AB3 = __funnelshift_l(w2, w3, 16) | __funnelshift_l(w2, w3, 8);
			/* third byte 0 or 1: a boundary event or a reset (rare) */
			if (haszero(AB0 | (w0 & kfe)) | haszero(AB1 | (w1 & kfe)) | haszero(AB2 | (w2 & kfe)) |
			    haszero(AB3 | (w3 & kfe)))
				evrows |= 1u << (c >> 5);
			uint32_t del = msb_to_nib(zero_bytes_msb(AB0 | (w0 ^ k3))) |
				       msb_to_nib(zero_bytes_msb(AB1 | (w1 ^ k3))) << 4 |
				       msb_to_nib(zero_bytes_msb(AB2 | (w2 ^ k3))) << 8 |
				       msb_to_nib(zero_bytes_msb(AB3 | (w3 ^ k3))) << 12;
			if (!full)
				del &= valid16(c * 16, nvalid);
			if (del)
				s.M[c] = (uint16_t)del;
			ecnt += (uint32_t)__popc(del);
		}
	}
	/* a start code ending just before the span resets at its first byte */
	const bool start_reset = rawb[-3] == 0 && rawb[-2] == 0 && rawb[-1] == 1 && nvalid > 0;
	__syncwarp();
	evrows = warp_or(evrows) | (start_reset ? 1u : 0u);
	const bool any_ev = evrows != 0;
	/* a span without a reset point (4 in 5) is known to be one here: its chain word goes out
	 * now, before the prefix, so that the spans after it wait as little as possible */
	ecnt = warp_add(ecnt);
	if (!any_ev && lane == 0)
		st_relaxed_u64(a.chain + t, chain_word(a.epoch, t == 0, ecnt));
	trace_mark<TRACE>(a, t, lane, 2);

	/* ---- P2: EPBs of the span before every chunk (a lane sums ROWS consecutive chunks) ---- */
	uint32_t etot;
	{
		uint32_t ex[ROWS];
		uint32_t run = 0;
		const uint16_t *mp = s.M + lane * ROWS;
		if (ROWS == 8) {
			const uint4 mv = *(const uint4 *)mp;
			const uint32_t mw[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
			for (int k = 0; k < 4; k++) {
				ex[2 * k % ROWS] = run;
				run += (uint32_t)__popc(mw[k] & 0xffffu);
				ex[(2 * k + 1) % ROWS] = run;
				run += (uint32_t)__popc(mw[k] >> 16);
			}
		} else {
#pragma unroll
			for (int k = 0; k < ROWS; k++) {
				ex[k] = run;
				run += (uint32_t)__popc(mp[k]);
			}
		}
		uint32_t inc = run;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
			if (lane >= (uint32_t)d)
				inc += o;
		}
		etot = __shfl_sync(FULL_MASK, inc, 31);
		const uint32_t base = inc - run;
		uint16_t *ep = s.E + lane * ROWS;
		if (ROWS == 8) {
			*(uint4 *)ep = make_uint4((base + ex[0]) | (base + ex[1 % ROWS]) << 16,
						  (base + ex[2 % ROWS]) | (base + ex[3 % ROWS]) << 16,
						  (base + ex[4 % ROWS]) | (base + ex[5 % ROWS]) << 16,
						  (base + ex[6 % ROWS]) | (base + ex[7 % ROWS]) << 16);
		} else {
#pragma unroll
			for (int k = 0; k < ROWS; k++)
				ep[k] = (uint16_t)(base + ex[k]);
		}
		__syncwarp();
	}

	/* ---- P2b (spans with a boundary event or a reset, ~1 in 5): events counted, reset points
	 * turned into per-row shift bases and byte-wise rows ---- */
	uint32_t nev = 0, nsc = 0, rbhas = 0, bw = 0, first_r = 0xffffffffu;
	int32_t cur_rel = 0;
	bool has = false;
	if (any_ev) {
		uint32_t carry = start_reset ? 1u : 0u;
		const uint32_t walk = evrows | evrows << 1; /* a start code ending a row resets in the next */
		for (int i = 0; i < ROWS; i++) {
			if (lane == 0)
				s.rbrel[i] = cur_rel;
			if (has)
				rbhas |= 1u << i;
			if (!((walk >> i) & 1))
				continue;
			const uint32_t c = i * 32 + lane;
			const uint32_t p0 = c * 16;
			const uint4 v = *(const uint4 *)(raw32 + 4 * c);
			const TMasks m = t_masks(raw32[4 * (int)c - 1], v.x, v.y, v.z, v.w);
			const uint32_t evm = event_valid16(p0, nvalid, t == 0);
			nev += (uint32_t)__popc(m.ev16 & evm);
			nsc += (uint32_t)__popc(m.sc16 & evm);
			uint32_t up = __shfl_up_sync(FULL_MASK, m.sc16 >> 15, 1);
			if (lane == 0)
				up = carry;
			carry = __shfl_sync(FULL_MASK, m.sc16 >> 15, 31);
			const uint32_t rm = ((m.sc16 << 1) | up) & valid16(p0, nvalid);
			const uint32_t bal = __ballot_sync(FULL_MASK, rm != 0);
			if (bal) {
				bw |= 1u << i;
				const uint32_t top = 31u - (uint32_t)__clz((int)(rm | 1u));
				const uint32_t e_last = (uint32_t)s.E[c] + (uint32_t)__popc(s.M[c] & ((1u << top) - 1u));
				const int tl = 31 - __clz((int)bal);
				cur_rel = -(int32_t)__shfl_sync(FULL_MASK, e_last, tl);
				const uint32_t fr = __shfl_sync(FULL_MASK, p0 + (uint32_t)__ffs((int)rm) - 1, __ffs((int)bal) - 1);
				if (!has)
					first_r = fr;
				has = true;
			}
		}
		nev = warp_add(nev);
		nsc = warp_add(nsc);
	}
	if (!full) { /* rows the shard ends in (or that lie past its end) go byte by byte */
		for (int i = 0; i < ROWS; i++)
			if ((uint32_t)(i + 1) * 512u > nvalid)
				bw |= 1u << i;
	}

	/* ---- P3: the span's chain word, published at once; event slot asked for; then the shift
	 * at the span start by the look-back ---- */
	const uint32_t e_tail = has ? (uint32_t)((int32_t)etot + cur_rel) : etot;
	if (lane == 0 && any_ev)
		st_relaxed_u64(a.chain + t, chain_word(a.epoch, has || t == 0, e_tail));
	uint64_t d0 = 0;
	trace_mark<TRACE>(a, t, lane, 3);
	if (t > 0) {
		d0 = lookback(a, t, rstart, lane);
		if (d0 == kDefer) {
			/* the shift depends on the region before, which is not done: second pass */
			if (lane == 0) {
				a.deferred[atomicAdd(a.ctrl + 5, 1u)] = t;
				next = take_ticket(a, tc, tn);
			}
			__syncwarp();
			return;
		}
		if (!has && lane == 0)
			st_relaxed_u64(a.chain + t, chain_word(a.epoch, true, d0 + (uint64_t)etot));
	}
	uint32_t evbase = 0;
	if (lane == 0) {
		if (nev)
			evbase = atomicAdd(a.ctrl + 1, nev);
		if (t == 0)
			a.ctrl[2] = first_r < 3 ? first_r : 0u;
		a.fin[t] = (uint64_t)evbase | (uint64_t)nev << 32 | (uint64_t)nsc << 48;
	}
	__syncwarp(); /* rbrel written by lane 0 */
	trace_mark<TRACE>(a, t, lane, 4);

	/* ---- P5 (spans with events): event records with the shift at the event (before the emit
	 * pass reuses E of byte-wise rows) ---- */
	if (nev) {
		uint32_t idx0 = __shfl_sync(FULL_MASK, evbase, 0);
		for (int i = 0; i < ROWS; i++) {
			if (!((evrows >> i) & 1))
				continue;
			const uint32_t c = i * 32 + lane;
			const uint32_t p0 = c * 16;
			const uint4 v = *(const uint4 *)(raw32 + 4 * c);
			const uint32_t pw = raw32[4 * (int)c - 1];
			const TMasks m = t_masks(pw, v.x, v.y, v.z, v.w);
			const uint32_t ev = m.ev16 & event_valid16(p0, nvalid, t == 0);
			const uint32_t mc = s.M[c], ec = s.E[c];
			/* resets of the row: an event after one takes its shift from there */
			uint32_t rm = 0;
			if ((bw >> i) & 1)
				rm = reset_mask(pw, v.x, v.y, v.z, v.w) & valid16(p0, nvalid);
			const uint32_t rbal = __ballot_sync(FULL_MASK, rm != 0);
			const uint32_t rtop = 31u - (uint32_t)__clz((int)(rm | 1u));
			const uint32_t e_last = ec + (uint32_t)__popc(mc & ((1u << rtop) - 1u));
			const uint32_t below = rbal & ltmask;
			const uint32_t e_prev = __shfl_sync(FULL_MASK, e_last, below ? 31 - __clz((int)below) : 0);
			uint64_t rowsh = d0;
			if ((rbhas >> i) & 1)
				rowsh = (uint64_t)(int64_t)s.rbrel[i];
			const uint32_t cnt = (uint32_t)__popc(ev);
			uint32_t einc = cnt;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t o = __shfl_up_sync(FULL_MASK, einc, d);
				if (lane >= (uint32_t)d)
					einc += o;
			}
			uint32_t idx = idx0 + einc - cnt;
			for (uint32_t x = ev; x; idx++) {
				const uint32_t j = (uint32_t)__ffs((int)x) - 1;
				x &= x - 1;
				const uint32_t e_here = ec + (uint32_t)__popc(mc & ((1u << j) - 1u));
				const uint32_t rmb = rm & ((2u << j) - 1u);
				uint64_t shift;
				if (rmb) {
					const uint32_t tp = 31u - (uint32_t)__clz((int)rmb);
					shift = e_here - (ec + (uint32_t)__popc(mc & ((1u << tp) - 1u)));
				} else if (below) {
					shift = e_here - e_prev;
				} else {
					shift = rowsh + (uint64_t)e_here;
				}
				if ((uint64_t)idx < a.ev_cap) {
					a.evbuf[2 * (uint64_t)idx] =
						(span_off + p0 + j - 2) | (uint64_t)((m.sc16 >> j) & 1) << 62;
					a.evbuf[2 * (uint64_t)idx + 1] = shift;
				}
			}
			idx0 += __shfl_sync(FULL_MASK, einc, 31);
		}
	}

	/* ---- P4: emit the rows, then the byte-exact pass (listed chunks, byte-wise rows) ---- */
	{
		const uint32_t bwl = bw << 1 | 1u | 1u << (ROWS + 1);
		uint8_t *const out_span = a.rbsp + span_off;
		uint32_t ndirty;
		if (rbhas == 0 && bw == 0)
			ndirty = emit_span<ROWS, true>(s, lane, bwl, 0, d0, out_span, nvalid);
		else
			ndirty = emit_span<ROWS, false>(s, lane, bwl, rbhas, d0, out_span, nvalid);
		__syncwarp();
		trace_mark<TRACE>(a, t, lane, 5);
		/* the next span's ticket is asked for here: its round trip (the one contended word of
		 * the kernel, ~1.7 us under load) hides behind the byte-exact pass */
		if (lane == 0)
			next = take_ticket(a, tc, tn);
		for (uint32_t g = lane; g < ndirty; g += 32) {
			const uint32_t c = s.dl[g];
			const uint32_t R = c >> 5;
			uint32_t shlo;
			uint8_t *basep;
			row_base<ROWS>(s, R, rbhas, d0, out_span, shlo, basep);
			const bool nextlim = R == (uint32_t)ROWS - 1 || ((bw >> (R + 1)) & 1);
			dirty_chunk<ROWS>(s, c, shlo, basep, nextlim ? (R + 1) * 512u : (uint32_t)C::SPAN + 64u);
		}
		for (uint32_t x = bw; x; x &= x - 1)
			bytewise_chunk<ROWS>(s, ((uint32_t)__ffs((int)x) - 1) * 32 + lane, rbhas, d0, out_span, nvalid);
	}
	__syncwarp(); /* the slice is reused by the next span */
	trace_mark<TRACE>(a, t, lane, 6);
}

template <int ROWS, int MINB, bool TRACE = false>
__global__ void __launch_bounds__(kT, MINB) scan7_kernel(const Scan7Args a)
{
	__shared__ WSmem<ROWS> sm[kW];
	const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	WSmem<ROWS> &s = sm[warp];
	if (lane == 0)
		bulk_bar_init(&s.bar);
	__syncwarp();
	uint32_t parity = 0;
	uint32_t next = 0;
	/* every counter needs a warp of its own: no more counters than warps in the grid */
	const uint32_t gw = gridDim.x * (uint32_t)kW;
	const uint32_t tn = a.pass2 || a.tick_n <= 1 ? 1u : (a.tick_n < gw ? a.tick_n : gw);
	const uint32_t tc = (blockIdx.x * (uint32_t)kW + warp) % tn;
	if (lane == 0)
		next = take_ticket(a, tc, tn);
	/* first pass: ticket -> (region, index): consecutive spans of the stream are taken `regions`
	 * tickets apart, so a span's predecessors published their chain words microseconds ago;
	 * second pass: the spans the first one deferred */
	const uint32_t limit = a.pass2 ? *(volatile uint32_t *)(a.ctrl + 5) : a.regions * a.region_len;
	for (;;) {
		const uint32_t tk = __shfl_sync(FULL_MASK, next, 0);
		if (tk >= limit)
			break;
		uint32_t t, rstart = 0;
		if (a.pass2) {
			t = a.deferred[tk];
		} else {
			const uint32_t i = tk / a.regions, r = tk - i * a.regions;
			rstart = r * a.region_len;
			t = rstart + i;
			if (t >= a.num_spans) { /* the last regions are ragged */
				if (lane == 0)
					next = take_ticket(a, tc, tn);
				continue;
			}
		}
		do_span<ROWS, TRACE>(s, a, t, rstart, lane, parity, next, tc, tn);
	}
}

/*
 * Scan only (no RBSP): the NAL table of h264_reader_parse's loop.  A warp takes spans round
 * robin (no chain, so no ticket), loads its ROWS x 512 bytes straight into registers and tests
 * them there; 1.0 B of traffic per input byte.  Events are rare: only rows with a candidate
 * (00 00 followed by a byte <= 1) evaluate the exact masks.
 */
template <int ROWS, int MINB, bool AVCC = false>
__global__ void __launch_bounds__(kT, MINB) scan7_only_kernel(const Scan7Args a)
{
	using C = Cfg<ROWS>;
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t gw = (blockIdx.x * kT + threadIdx.x) >> 5, nw = (gridDim.x * kT) >> 5;
	const uint32_t kfe = 0xfefefefeu;
	for (uint32_t t = gw; t < a.num_spans; t += nw) {
		const uint64_t span_off = (uint64_t)t * C::SPAN;
		const uint32_t nvalid = span_off >= a.len ? 0u
							  : (a.len - span_off >= (uint64_t)C::SPAN ? (uint32_t)C::SPAN
												   : (uint32_t)(a.len - span_off));
		uint4 v[ROWS];
		if (nvalid == (uint32_t)C::SPAN) {
#pragma unroll
			for (int i = 0; i < ROWS; i++)
				v[i] = ldg_stream16(a.in + span_off + (uint64_t)(i * 32 + lane) * 16);
		} else {
#pragma unroll
			for (int i = 0; i < ROWS; i++) {
				const uint64_t o = span_off + (uint64_t)(i * 32 + lane) * 16;
				if (o + 16 <= a.len)
					v[i] = ldg_stream16(a.in + o);
				else
					v[i] = make_uint4(edge_word(a, o), edge_word(a, o + 4), edge_word(a, o + 8),
							  edge_word(a, o + 12));
			}
		}
		uint32_t carry = lane == 0 ? prev_word(a, span_off) : 0u;
		uint32_t masks[ROWS]; /* ev16 | sc16 << 16 per row */
		uint32_t nev = 0, nsc = 0;
#pragma unroll
		for (int i = 0; i < ROWS; i++) {
			uint32_t pw = __shfl_up_sync(FULL_MASK, v[i].w, 1);
			const uint32_t last = __shfl_sync(FULL_MASK, v[i].w, 31);
			if (lane == 0)
				pw = carry;
			carry = last;
			const uint32_t w0 = v[i].x, w1 = v[i].y, w2 = v[i].z, w3 = v[i].w;
			const uint32_t AB0 = __funnelshift_l(pw, w0, 16) | __funnelshift_l(pw, w0, 8);
			const uint32_t AB1 = __funnelshift_l(w0, w1, 16) | __funnelshift_l(w0, w1, 8);
			const uint32_t AB2 = __funnelshift_l(w1, w2, 16) | __funnelshift_l(w1, w2, 8);
			const uint32_t AB3 = __funnelshift_l(w2, w3, 16) | __funnelshift_l(w2, w3, 8);
			const uint32_t X0 = AB0 | (w0 & kfe), X1 = AB1 | (w1 & kfe), X2 = AB2 | (w2 & kfe), X3 = AB3 | (w3 & kfe);
			const uint32_t k1 = 0x01010101u;
			const uint32_t acc = ((X0 - k1) & ~X0) | ((X1 - k1) & ~X1) | ((X2 - k1) & ~X2) | ((X3 - k1) & ~X3);
			masks[i] = 0;
			if (acc & 0x80808080u) {
				const uint32_t p0 = (uint32_t)(i * 32 + lane) * 16;
				if (AVCC) {
					/* 4-byte start codes 00 00 00 01 only (src/h264.c:184-207), owned by the
					 * position of their 01 byte; the record carries the code's first byte */
					const uint32_t m4 = code4_mask(pw, w0, w1, w2, w3) & valid16(p0, nvalid);
					masks[i] = m4 | m4 << 16;
					nev += (uint32_t)__popc(m4);
					nsc += (uint32_t)__popc(m4);
				} else {
					const TMasks m = t_masks(pw, w0, w1, w2, w3);
					const uint32_t evm = event_valid16(p0, nvalid, t == 0);
					masks[i] = (m.ev16 & evm) | (m.sc16 & evm) << 16;
					nev += (uint32_t)__popc(m.ev16 & evm);
					nsc += (uint32_t)__popc(m.sc16 & evm);
				}
			}
		}
		nev = warp_add(nev);
		nsc = warp_add(nsc);
		uint32_t evbase = 0;
		if (nev) {
			if (lane == 0)
				evbase = atomicAdd(a.ctrl + 1, nev);
			uint32_t idx0 = __shfl_sync(FULL_MASK, evbase, 0);
#pragma unroll
			for (int i = 0; i < ROWS; i++) {
				const uint32_t any = __ballot_sync(FULL_MASK, masks[i] != 0);
				if (!any)
					continue;
				const uint32_t ev = masks[i] & 0xffffu, sc = masks[i] >> 16;
				const uint32_t cnt = (uint32_t)__popc(ev);
				uint32_t einc = cnt;
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const uint32_t o = __shfl_up_sync(FULL_MASK, einc, d);
					if (lane >= (uint32_t)d)
						einc += o;
				}
				uint32_t idx = idx0 + einc - cnt;
				const uint32_t p0 = (uint32_t)(i * 32 + lane) * 16;
				for (uint32_t x = ev; x; idx++) {
					const uint32_t j = (uint32_t)__ffs((int)x) - 1;
					x &= x - 1;
					if ((uint64_t)idx < a.ev_cap) {
						a.evbuf[2 * (uint64_t)idx] = (span_off + p0 + j - (AVCC ? 3 : 2)) | (uint64_t)((sc >> j) & 1) << 62;
						a.evbuf[2 * (uint64_t)idx + 1] = 0;
					}
				}
				idx0 += __shfl_sync(FULL_MASK, einc, 31);
			}
		}
		if (lane == 0)
			a.fin[t] = (uint64_t)evbase | (uint64_t)nev << 32 | (uint64_t)nsc << 48;
	}
}

/* ---- finalize ------------------------------------------------------------------------- */

struct Fin7Args {
	const uint64_t *fin;   /* per span: event slot | events << 32 | start codes << 48 */
	const uint64_t *chain; /* per span (strip): the last one holds the shift at the shard end */
	uint32_t num_spans;
	uint32_t nblk;         /* fin7_spans blocks */
	const uint64_t *evbuf;
	uint64_t ev_cap;
	uint64_t *ordered;  /* 3 words per event: record, shift, NAL index */
	uint64_t *span_pre; /* per span: events | start codes << 32 before it in its fin7_spans block */
	uint64_t *blk;      /* per fin7_spans block: totals, then (last block) exclusive prefixes */
	uint64_t *totals;   /* [0] events  [1] start codes  [2] RBSP byte sum */
	uint32_t *ctrl;
	uint64_t len, base;
	uint64_t *nal_start, *nal_end, *nal_rbsp, *nal_rbsp_len;
	uint64_t nal_cap;
	struct h264gpu_scan_result *result;
	uint32_t has_right, strip, assume_in;
};

constexpr int kFinT = 1024;

/* a span per thread: block-local exclusive prefix of (events, start codes); the last block to
 * finish turns the block totals into prefixes and the stream totals */
__global__ void __launch_bounds__(kFinT, 1) fin7_spans(const Fin7Args f)
{
	__shared__ uint64_t wtot[kFinT / 32];
	__shared__ uint32_t is_last;
	const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint64_t t = (uint64_t)blockIdx.x * kFinT + tid;
	uint64_t mine = 0;
	if (t < f.num_spans) {
		const uint64_t w = f.fin[t];
		mine = ((w >> 32) & 0xffffu) | ((w >> 48) & 0xffffu) << 32;
	}
	uint64_t inc = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint64_t o = __shfl_up_sync(FULL_MASK, inc, d);
		if (lane >= (uint32_t)d)
			inc += o;
	}
	if (lane == 31)
		wtot[warp] = inc;
	__syncthreads();
	if (warp == 0) {
		uint64_t x = wtot[lane];
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint64_t o = __shfl_up_sync(FULL_MASK, x, d);
			if (lane >= (uint32_t)d)
				x += o;
		}
		wtot[lane] = x; /* inclusive over warps */
	}
	__syncthreads();
	const uint64_t wbase = warp ? wtot[warp - 1] : 0;
	if (t < f.num_spans)
		f.span_pre[t] = wbase + inc - mine;
	if (tid == 0) {
		f.blk[blockIdx.x] = wtot[kFinT / 32 - 1];
		__threadfence();
		is_last = atomicAdd(f.ctrl + 3, 1u) == gridDim.x - 1 ? 1u : 0u;
	}
	__syncthreads();
	if (!is_last)
		return;
	__threadfence();
	/* exclusive prefix over the block totals, kFinT at a time */
	uint64_t carry = 0;
	for (uint32_t b0 = 0; b0 < f.nblk; b0 += kFinT) {
		const uint32_t b = b0 + tid;
		const uint64_t v = b < f.nblk ? *(volatile uint64_t *)(f.blk + b) : 0;
		uint64_t x = v;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint64_t o = __shfl_up_sync(FULL_MASK, x, d);
			if (lane >= (uint32_t)d)
				x += o;
		}
		__syncthreads();
		if (lane == 31)
			wtot[warp] = x;
		__syncthreads();
		if (warp == 0) {
			uint64_t y = wtot[lane];
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint64_t o = __shfl_up_sync(FULL_MASK, y, d);
				if (lane >= (uint32_t)d)
					y += o;
			}
			wtot[lane] = y;
		}
		__syncthreads();
		const uint64_t wb = warp ? wtot[warp - 1] : 0;
		if (b < f.nblk)
			f.blk[f.nblk + b] = carry + wb + x - v;
		carry += wtot[kFinT / 32 - 1];
	}
	if (tid == 0) {
		f.totals[0] = carry & 0xffffffffull;
		f.totals[1] = carry >> 32;
		f.totals[2] = 0;
		f.ctrl[3] = 0;
	}
}

/* a span per thread: the span's event records to their ordered places, with their NAL index */
__global__ void __launch_bounds__(256) fin7_order(const Fin7Args f)
{
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= f.num_spans)
		return;
	const uint64_t w = f.fin[t];
	const uint32_t n = (uint32_t)((w >> 32) & 0xffffu);
	if (n == 0)
		return;
	const uint64_t basei = w & 0xffffffffull;
	const uint64_t pre = f.span_pre[t] + f.blk[f.nblk + t / kFinT];
	const uint64_t e = pre & 0xffffffffull;
	uint64_t k = pre >> 32;
	for (uint32_t i = 0; i < n; i++) {
		if (e + i >= f.ev_cap || basei + i >= f.ev_cap)
			break;
		const uint64_t rec = f.evbuf[2 * (basei + i)];
		f.ordered[3 * (e + i)] = rec;
		f.ordered[3 * (e + i) + 1] = f.evbuf[2 * (basei + i) + 1];
		f.ordered[3 * (e + i) + 2] = k;
		k += (rec >> 62) & 1;
	}
}

/* an ordered event per thread (grid-stride): table entries and the RBSP byte sum; the last
 * block writes the result struct and re-arms the control words for the next launch */
__global__ void __launch_bounds__(256) fin7_table(const Fin7Args f)
{
	__shared__ uint32_t is_last;
	const uint64_t total_ev = f.totals[0], total_sc = f.totals[1];
	const uint64_t nordered = total_ev < f.ev_cap ? total_ev : f.ev_cap;
	const uint64_t posmask = (1ull << 40) - 1;
	/* shift at the shard end (strip): the last span's chain word */
	const uint64_t end_shift = f.strip ? (f.chain[f.num_spans - 1] & kValMask) : 0;
	uint64_t rsum = 0;
	for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nordered;
	     e += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t rec = f.ordered[3 * e];
		if (!((rec >> 62) & 1))
			continue;
		const uint64_t q = rec & posmask, k = f.ordered[3 * e + 2];
		const uint64_t start = q + 3;
		uint64_t end, sh;
		bool open = false;
		if (e + 1 < nordered) {
			end = f.ordered[3 * (e + 1)] & posmask;
			sh = f.ordered[3 * (e + 1) + 1];
		} else {
			end = f.len;
			sh = end_shift;
			open = true;
		}
		uint64_t rlen = 0;
		if (start >= end)
			end = start; /* start code in the last bytes of the shard: empty so far */
		else
			rlen = (end - start) - sh;
		if (!f.strip)
			rlen = 0;
		rsum += rlen;
		if (k < f.nal_cap) {
			f.nal_start[k] = f.base + start;
			if (!open || !f.has_right)
				f.nal_end[k] = f.base + end;
			if (f.nal_rbsp)
				f.nal_rbsp[k] = start;
			if (f.nal_rbsp_len)
				f.nal_rbsp_len[k] = rlen;
		}
	}
#pragma unroll
	for (int d = 16; d >= 1; d >>= 1)
		rsum += __shfl_xor_sync(FULL_MASK, rsum, d);
	if ((threadIdx.x & 31) == 0 && rsum)
		atomicAdd((unsigned long long *)&f.totals[2], (unsigned long long)rsum);
	__syncthreads();
	if (threadIdx.x == 0) {
		__threadfence();
		is_last = atomicAdd(f.ctrl + 4, 1u) == gridDim.x - 1 ? 1u : 0u;
	}
	__syncthreads();
	if (!is_last || threadIdx.x != 0)
		return;
	__threadfence();
	struct h264gpu_scan_result r;
	r.n_nal = total_sc;
	r.any_event = total_ev ? 1u : 0u;
	r.first_event_pos = H264GPU_NONE;
	r.first_event_is_sc = 0;
	r.head_bytes = 0;
	r.end_open = 0;
	r.reserved = total_ev > f.ev_cap ? 1u : 0u; /* event buffer overflow: table incomplete */
	const uint64_t r0 = f.ctrl[2];
	if (nordered) {
		const uint64_t first = f.ordered[0], last = f.ordered[3 * (nordered - 1)];
		const uint64_t q = first & posmask;
		r.first_event_pos = f.base + q;
		r.first_event_is_sc = (uint32_t)((first >> 62) & 1);
		r.head_bytes = f.strip ? q - f.ordered[1] : 0;
		r.end_open = (uint32_t)((last >> 62) & 1);
	} else {
		r.head_bytes = f.strip ? f.len - end_shift : 0;
	}
	/* a start code that began in the previous shard: its last bytes are not NAL data */
	r.head_bytes = r.head_bytes > r0 ? r.head_bytes - r0 : 0;
	/* the bytes before the first event belong to a NAL only if an earlier shard left one
	 * open; a shard that assumes so reports them (the merge drops them otherwise) */
	if (!f.assume_in)
		r.head_bytes = 0;
	r.rbsp_bytes = f.strip ? r.head_bytes + *(volatile uint64_t *)(f.totals + 2) : 0;
	*f.result = r;
	f.ctrl[0] = 0;
	f.ctrl[1] = 0;
	f.ctrl[2] = 0;
	f.ctrl[4] = 0;
	f.ctrl[5] = 0;
	f.ctrl[6] = 0;
	for (uint32_t c = 0; c < kTickMax; c++)
		f.ctrl[kTickBase + c * kTickStride] = 0;
}

/*
 * h264_byte_stream_to_avcc (src/h264.c:210-246), the part after the scan: every 4-byte start
 * code becomes the big-endian length of the NAL behind it (up to the next code or the end).
 * ordered: fin7_order's output for a scan7_only_kernel<.., AVCC = true> launch.  The reference
 * stops when 4 bytes or fewer are left (`while (len > 4)`): a code in the last 4 bytes stays.
 */
__global__ void __launch_bounds__(256) avcc_write_kernel(const uint64_t *ordered, const uint64_t *totals,
							  uint64_t ev_cap, uint8_t *data, uint64_t len,
							  uint64_t *pos_out, uint64_t *count_out)
{
	const uint64_t n = totals[0] < ev_cap ? totals[0] : ev_cap;
	const uint64_t posmask = (1ull << 40) - 1;
	for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t p = ordered[3 * e] & posmask;
		const uint64_t nxt = e + 1 < n ? (ordered[3 * (e + 1)] & posmask) : len;
		if (pos_out)
			pos_out[e] = p;
		if (len - p <= 4)
			continue;
		const uint32_t nal = (uint32_t)(nxt - p - 4);
		if (data) {
			data[p] = (uint8_t)(nal >> 24);
			data[p + 1] = (uint8_t)(nal >> 16);
			data[p + 2] = (uint8_t)(nal >> 8);
			data[p + 3] = (uint8_t)nal;
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0 && count_out)
		*count_out = totals[0];
}

/* re-arm the control words after an AVCC scan (fin7_table, which does it for the NAL scan, is not run) */
__global__ void avcc_rearm_kernel(uint32_t *ctrl)
{
	ctrl[0] = ctrl[1] = ctrl[2] = ctrl[3] = ctrl[4] = ctrl[5] = ctrl[6] = 0;
	for (uint32_t c = 0; c < kTickMax; c++)
		ctrl[kTickBase + c * kTickStride] = 0;
}

} /* namespace annexb7 */

#endif /* ANNEXB_SCAN7_CUH */
