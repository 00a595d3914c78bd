/*
 * annexb_scan5.cuh — K1/K2 fifth generation: Annex-B scan + NAL table + EPB strip with the
 * RBSP of every NAL written IN PLACE: NAL k's RBSP starts at the offset of its first byte
 * (d_rbsp + nal_start[k] - base) and is nal_rbsp_len[k] bytes long; bytes between two NALs'
 * RBSPs are unspecified.
 *
 * Why: with a tightly packed RBSP buffer the output position of a byte depends on EVERY
 * earlier tile (one global prefix sum), and a chained scan then runs at the pace of the
 * slowest of the ~500 tiles in flight: measured 43 % of a tile's life in gen 2 was the
 * look-back wait (profiles/r01_scan2_phase_trace.txt).  Here a byte moves left only by the
 * number of emulation prevention bytes since the start of ITS NAL:
 *
 *     out(p) = p - #{ EPB in [S(p), p) },   S(p) = first byte of the NAL p lies in
 *
 * so the chain is cut at every start code.  A tile needs at most the EPB count since the last
 * start code before it (usually the previous tile's aggregate), a tile
 * that contains a start code publishes its inclusive prefix without looking back at all, and
 * bytes outside NALs are simply copied along (they land in the gaps).  The NAL table no longer rides a
 * global chain either: every tile appends its boundary events to a buffer and
 * scan5_finalize (one small kernel) orders them, pairs each start code with the next event
 * and derives the RBSP lengths from a prefix sum of per-tile EPB counts.
 *
 * Reference behaviour reproduced bit-exactly (Parrot-Developers/libh264):
 *   NAL loop of h264_reader_parse            src/h264_reader.c:133-140
 *   h264_find_nalu / start / end code search src/h264_bitstream.c:87-184
 *   EPB removal in h264_bs_fetch             include/h264/h264_bitstream.h:168-190
 * Closed form: see annexb_scan.cuh.
 *
 * Per tile (one CTA, 256 threads x CPT 16-byte chunks, staged once by a bulk async copy):
 *   P1  classify: EPB delete mask M[chunk]; chunks owning a boundary event flagged in EVB
 *   P1b warp 0 (only with events): count events, collect the reset points r = q + 3 of the
 *       start codes (where the shift goes back to 0)
 *   P2  EPB counts: block prefix, tile aggregate (EPBs after the last reset, has-reset flag),
 *       published at once; a tile with a reset publishes its inclusive prefix immediately
 *   P3  look-back by warp 0 (short: ends at the first predecessor with a reset or a prefix)
 *       while the other warps clear the unit tables
 *   P4  emit_tile: all segments in one pass over unit space: bin deletions per 16-byte output
 *       unit, prefix, copy clean units through a funnel shift, squeeze dirty units
 *   P5  last warp: event records (position, kind, EPBs before it in the tile) -> event buffer
 */
#ifndef ANNEXB_SCAN5_CUH
#define ANNEXB_SCAN5_CUH

#include "annexb_scan.cuh"
#include "annexb_scan2.cuh"

namespace annexb5 {

using annexb::ScanArgs;
using annexb::SlowMasks;
using annexb::kInvalid;
using annexb2::msb_to_nib;
using annexb2::trace_mark;

constexpr int kT = 256;
constexpr int kW = kT / 32;
constexpr int kRcap = 64; /* reset points listed per tile; denser tiles take the byte-wise path */

template <int CPT> struct Cfg {
	static constexpr int NCH = kT * CPT;
	static constexpr int TILE = NCH * 16;
	static constexpr int UPT = CPT + 1;
	static constexpr int NUCAP = kT * UPT;
	/* deleting chunks / dirty units listed per tile (~9 % of the chunks at P(00) = 3/16);
	 * denser tiles take the unlisted paths.  Sized so that five CTAs fit an SM. */
	static constexpr int LIST_CAP = NCH / 8 < 64 ? 64 : NCH / 8;
};

template <int CPT> struct __align__(128) Smem {
	uint8_t raw[16 + Cfg<CPT>::TILE + 48]; /* [left halo pad][tile][right halo pad] */
	uint16_t M[Cfg<CPT>::NCH + 16];        /* delete mask per chunk (bit j = byte j is an EPB) */
	uint16_t T[Cfg<CPT>::NUCAP];           /* deletions binned per output unit -> inclusive prefix */
	uint32_t D[(Cfg<CPT>::NUCAP + 31) / 32];
	uint32_t EVB[Cfg<CPT>::NCH / 32];
	uint16_t cl[Cfg<CPT>::LIST_CAP];
	uint16_t dl[Cfg<CPT>::LIST_CAP];
	uint32_t epre[kT]; /* EPBs in the chunks before thread group g (CPT chunks each) */
	uint32_t wsum[kW];
	uint32_t wsum2[kW];
	uint16_t rl[kRcap]; /* reset points, ascending */
	/* segments of the tile in unit space (see emit_tile): start / end x, alignment padding */
	uint16_t segx[kRcap + 1], sege[kRcap + 1], segj[kRcap + 1];
	int64_t segg[kRcap + 1]; /* global byte offset of x = 0 of the segment */
	uint64_t bar;
	uint64_t b0; /* shift at the tile start */
	uint32_t tile, anyev, ndirty, ndel, nreset, big, rlast, nev, nsc, evbase, eall;
};

/* descriptor words of a tile (4 x u64): AGG, PFX (self-validating, bit 63 = not yet), EVT, EPB */
__device__ __forceinline__ uint64_t pack_agg(uint32_t e_tail, bool has)
{
	return (uint64_t)e_tail | (uint64_t)(has ? 1 : 0) << 32;
}

/* event record: shard-relative position | EPBs before it in its tile << 40 | start code << 62 */
__device__ __forceinline__ uint64_t pack_event(uint64_t q, uint32_t epb_before, bool is_sc)
{
	return q | (uint64_t)epb_before << 40 | (uint64_t)(is_sc ? 1 : 0) << 62;
}

template <int CPT>
__device__ __forceinline__ SlowMasks chunk_masks(const Smem<CPT> &s, uint32_t c, uint32_t nvalid)
{
	const uint32_t *raw32 = (const uint32_t *)(s.raw + 16);
	const uint4 v = *(const uint4 *)(raw32 + 4 * c);
	SlowMasks m = annexb::slow_masks(raw32[4 * (int)c - 1], v.x, v.y, v.z, v.w, raw32[4 * c + 4]);
	const uint32_t p0 = c * 16;
	const uint32_t nv = p0 >= nvalid ? 0u : (nvalid - p0 >= 16u ? 16u : nvalid - p0);
	const uint32_t vm = (1u << nv) - 1u;
	m.ev16 &= vm;
	m.sc16 &= vm;
	return m;
}

/* EPBs (deleted bytes) of the tile before source position p */
template <int CPT> __device__ __forceinline__ uint32_t epb_before(const Smem<CPT> &s, uint32_t p)
{
	const uint32_t cl = p >> 4;
	const uint32_t g = cl / CPT;
	uint32_t e = g < (uint32_t)kT ? s.epre[g] : s.eall;
	if (g >= (uint32_t)kT)
		return e;
	for (uint32_t c = g * CPT; c < cl; c++)
		e += (uint32_t)__popc(s.M[c]);
	e += (uint32_t)__popc(s.M[cl] & ((1u << (p & 15)) - 1u));
	return e;
}

/*
 * Shift at the start of tile t = EPBs since the last reset before it: fold the predecessors'
 * aggregates back to the nearest tile that has a reset (its tail count ends the walk) or an
 * inclusive prefix.  Whole warp 0.
 */
__device__ __forceinline__ uint64_t lookback(const ScanArgs &a, uint32_t t, uint32_t lane)
{
	uint64_t acc = 0;
	int64_t j0 = (int64_t)t - 1;
	for (;;) {
		const int64_t jl = j0 - (int64_t)lane;
		uint64_t agg, pfx;
		uint32_t stop, upto;
		for (;;) {
			agg = kInvalid;
			pfx = kInvalid;
			if (jl >= 0) {
				pfx = ld_relaxed_u64(a.desc + jl * 4 + 1);
				agg = ld_relaxed_u64(a.desc + jl * 4);
			} else {
				pfx = 0; /* before the shard: shift 0 */
			}
			const bool ends = !(pfx & kInvalid) || (!(agg & kInvalid) && ((agg >> 32) & 1));
			stop = __ballot_sync(FULL_MASK, ends);
			const uint32_t ok = __ballot_sync(FULL_MASK, ends || !(agg & kInvalid));
			upto = stop ? ((stop & (0u - stop)) << 1) - 1u : 0xffffffffu;
			if ((ok & upto) == upto)
				break;
			spin_pause(32);
		}
		const int fs = stop ? __ffs((int)stop) - 1 : 32;
		uint64_t mine = 0;
		if ((int)lane < fs)
			mine = agg & 0xffffffffull;
		else if ((int)lane == fs)
			mine = !(pfx & kInvalid) ? pfx : (agg & 0xffffffffull);
#pragma unroll
		for (int d = 16; d >= 1; d >>= 1)
			mine += __shfl_xor_sync(FULL_MASK, mine, d);
		acc += mine;
		if (fs < 32)
			return acc;
		j0 -= 32;
	}
}

template <int CPT> __device__ __forceinline__ void list_chunk(Smem<CPT> &s, uint32_t c)
{
	const uint32_t idx = atomicAdd(&s.ndel, 1u);
	if (idx < (uint32_t)Cfg<CPT>::LIST_CAP)
		s.cl[idx] = (uint16_t)c;
}

/* segment of source position p / of the unit starting at x: the last one that begins at or
 * before it (few resets per tile, so a short linear walk) */
template <int CPT> __device__ __forceinline__ uint32_t seg_of_src(const Smem<CPT> &s, uint32_t nreset, uint32_t p)
{
	uint32_t i = 0;
	while (i < nreset && s.rl[i] <= p)
		i++;
	return i;
}
template <int CPT> __device__ __forceinline__ uint32_t seg_of_x(const Smem<CPT> &s, uint32_t nreset, uint32_t x)
{
	uint32_t i = 0;
	while (i < nreset && s.segx[i + 1] <= x)
		i++;
	return i;
}

/*
 * Byte-exact output of the part of unit u that belongs to segment i (a unit with a deletion
 * inside, or the partial first / last unit of a segment).  Fast form: take the 32 source bytes
 * from the unit's first source byte, squeeze out the deleted ones, store 16.  Slow form (unit
 * starts before the segment, or > 16 deletions in the window): walk.
 */
template <int CPT>
__device__ __forceinline__ void dirty_unit(const Smem<CPT> &s, uint32_t u, uint32_t i, uint32_t a0, uint8_t *rbsp)
{
	const uint32_t xs = s.segx[i], xe = s.sege[i];
	const uint32_t sbias = a0 + s.segj[i];
	const uint32_t lo_src = i ? s.rl[i - 1] : 0u;
	const uint32_t x0 = 16 * u;
	const uint32_t lo = x0 > xs ? x0 : xs;
	const uint32_t hi = x0 + 16 < xe ? x0 + 16 : xe;
	if (lo >= hi)
		return;
	uint8_t *g = rbsp + (s.segg[i] + (int64_t)x0);
	uint64_t vlo = 0, vhi = 0;
	bool done = false;
	if (x0 >= xs) {
		const uint32_t pos = x0 - sbias + s.T[u];
		const uint32_t mi = pos >> 4;
		const uint64_t mb = (uint64_t)s.M[mi] | (uint64_t)s.M[mi + 1] << 16 | (uint64_t)s.M[mi + 2] << 32;
		uint32_t dm = (uint32_t)(mb >> (pos & 15));
		if (__popc(dm) <= 16) {
			const uint32_t *raw32 = (const uint32_t *)(s.raw + 16);
			const uint32_t wi = pos >> 2, sh = (pos & 3) * 8;
			uint32_t w[9];
#pragma unroll
			for (int k = 0; k < 9; k++)
				w[k] = raw32[wi + k];
			uint64_t q0 = (uint64_t)__funnelshift_r(w[0], w[1], sh) | (uint64_t)__funnelshift_r(w[1], w[2], sh) << 32;
			uint64_t q1 = (uint64_t)__funnelshift_r(w[2], w[3], sh) | (uint64_t)__funnelshift_r(w[3], w[4], sh) << 32;
			uint64_t q2 = (uint64_t)__funnelshift_r(w[4], w[5], sh) | (uint64_t)__funnelshift_r(w[5], w[6], sh) << 32;
			uint64_t q3 = (uint64_t)__funnelshift_r(w[6], w[7], sh) | (uint64_t)__funnelshift_r(w[7], w[8], sh) << 32;
			if ((dm & (dm - 1) & 0x1ffffu) == 0 && (dm & 0xffffu)) {
				/* the common case, one deleted byte k among the first 17: bytes below k stay,
				 * the rest moves down by one */
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
						const uint64_t m = (1ull << (8 * k)) - 1;
						q0 = (q0 & m) | ((q0 >> 8) & ~m) | (q1 << 56);
						q1 = (q1 >> 8) | (q2 << 56);
					} else {
						const uint64_t m = (1ull << (8 * (k - 8))) - 1;
						q1 = (q1 & m) | ((q1 >> 8) & ~m) | (q2 << 56);
					}
					q2 = (q2 >> 8) | (q3 << 56);
					q3 >>= 8;
				}
			}
			vlo = q0;
			vhi = q1;
			done = true;
		}
	}
	if (!done) {
		uint32_t pos = x0 >= xs ? x0 - sbias + s.T[u] : lo_src;
		const uint8_t *raw = s.raw + 16;
		for (uint32_t x = lo; x < hi; x++) {
			while ((s.M[pos >> 4] >> (pos & 15)) & 1)
				pos++;
			const uint64_t bt = raw[pos++];
			const uint32_t n = x - x0;
			if (n < 8)
				vlo |= bt << (8 * n);
			else
				vhi |= bt << (8 * (n - 8));
		}
	}
	if (lo == x0 && hi == x0 + 16) {
		stg_stream16(g, make_uint4((uint32_t)vlo, (uint32_t)(vlo >> 32), (uint32_t)vhi, (uint32_t)(vhi >> 32)));
	} else {
		for (uint32_t x = lo; x < hi; x++) {
			const uint32_t n = x - x0;
			g[n] = (uint8_t)((n < 8 ? vlo >> (8 * n) : vhi >> (8 * (n - 8))) & 0xff);
		}
	}
}

/* bin the deletions of chunk c by the output unit they land in */
template <int CPT>
__device__ __forceinline__ void bin_chunk(Smem<CPT> &s, uint32_t c, uint32_t nreset, uint32_t a0)
{
	const uint32_t p0 = c * 16;
	const uint32_t mk = s.M[c];
	uint32_t dm = mk;
	if (!dm)
		return;
	const uint32_t keptm = ~mk & 0xffffu;
	const uint32_t o = p0 - epb_before<CPT>(s, p0); /* kept bytes of the tile before the chunk */
	uint32_t *t32 = (uint32_t *)s.T;
	while (dm) {
		const uint32_t j = (uint32_t)__ffs((int)dm) - 1;
		const uint32_t r = (uint32_t)__ffs((int)~(dm >> j)) - 1; /* run of deleted bytes */
		const uint32_t seg = seg_of_src<CPT>(s, nreset, p0 + j);
		const uint32_t x = a0 + s.segj[seg] + o + (uint32_t)__popc(keptm & ((1u << j) - 1u));
		const uint32_t bin = (x + 15) >> 4;
		atomicAdd(&t32[bin >> 1], r << (16 * (bin & 1)));
		if (x & 15) {
			const uint32_t bit = 1u << ((x >> 4) & 31);
			const uint32_t old = atomicOr(&s.D[x >> 9], bit);
			if (!(old & bit)) {
				const uint32_t idx = atomicAdd(&s.ndirty, 1u);
				if (idx < (uint32_t)Cfg<CPT>::LIST_CAP)
					s.dl[idx] = (uint16_t)(x >> 4);
			}
		}
		dm &= ~(((1u << r) - 1u) << j);
	}
}

/*
 * Emit the whole tile in ONE pass over "unit space".  The tile's segments (head, then one per
 * reset) are laid out one after the other on an x axis, x = a0 + kept bytes before p + J(p),
 * where a0 makes the head segment's x congruent (mod 16) to its global output position and the
 * padding J grows by < 16 at every reset so that each later segment is congruent to ITS
 * output position (which restarts at the reset).  A 16-byte unit of x then maps to one aligned
 * 16-byte unit of d_rbsp (segment base + 16u); units inside a segment are copied from source
 * offset 16u - a0 - J + delpre[u] through a funnel shift, units with a deletion inside and the
 * partial first / last units of segments take the byte-exact path.  T and D are already clear.
 */
template <int CPT>
__device__ __forceinline__ void emit_tile(Smem<CPT> &s, const ScanArgs &a, uint64_t tile_off, uint32_t nreset,
					   uint32_t a0)
{
	using C = Cfg<CPT>;
	const uint32_t tid = threadIdx.x;
	const uint32_t lane = tid & 31, warp = tid >> 5;
	const uint32_t *raw32 = (const uint32_t *)(s.raw + 16);
	const uint32_t nseg = nreset + 1;
	const uint32_t NU = ((uint32_t)s.sege[nreset] + 15) >> 4;

	{
		const uint32_t ndel = s.ndel;
		if (ndel <= (uint32_t)C::LIST_CAP) {
			for (uint32_t i = tid; i < ndel; i += kT)
				bin_chunk<CPT>(s, s.cl[i], nreset, a0);
		} else {
			for (uint32_t c = tid; c < (uint32_t)C::NCH; c += kT)
				bin_chunk<CPT>(s, c, nreset, a0);
		}
	}
	__syncthreads();
	if (tid == 0)
		trace_mark(a, s.tile, 5);
	/* inclusive prefix of T: delpre[u] = deletions landing at or before 16u */
	{
		const uint32_t u0 = tid * C::UPT;
		uint32_t loc[C::UPT];
		uint32_t run = 0;
#pragma unroll
		for (int k = 0; k < C::UPT; k++) {
			run += s.T[u0 + k];
			loc[k] = run;
		}
		uint32_t rinc = run;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t o = __shfl_up_sync(FULL_MASK, rinc, d);
			if (lane >= (uint32_t)d)
				rinc += o;
		}
		if (lane == 31)
			s.wsum2[warp] = rinc;
		__syncthreads();
		uint32_t base = rinc - run;
#pragma unroll
		for (int j = 0; j < kW; j++)
			if ((uint32_t)j < warp)
				base += s.wsum2[j];
#pragma unroll
		for (int k = 0; k < C::UPT; k++)
			s.T[u0 + k] = (uint16_t)(base + loc[k]);
	}
	__syncthreads();
	if (tid == 0)
		trace_mark(a, s.tile, 6);
	/* clean units: one aligned 16-byte store each.  The segment of a unit is decided by one
	 * compare in the common cases (no or one reset in the tile). */
	{
		const uint32_t xb1 = nreset >= 1 ? (uint32_t)s.segx[1] : 0xffffffffu;
		const uint32_t e0 = s.sege[0], e1 = nreset >= 1 ? (uint32_t)s.sege[1] : 0u;
		const uint32_t bias0 = a0, bias1 = a0 + (nreset >= 1 ? (uint32_t)s.segj[1] : 0u);
		uint8_t *const out0 = a.rbsp + s.segg[0];
		uint8_t *const out1 = a.rbsp + (nreset >= 1 ? s.segg[1] : 0);
#pragma unroll 3
		for (uint32_t u = tid; u < NU; u += kT) {
			if ((s.D[u >> 5] >> (u & 31)) & 1)
				continue;
			const uint32_t x = 16 * u;
			uint32_t bias, end;
			uint8_t *outp;
			if (nreset <= 1) {
				const bool second = x >= xb1;
				bias = second ? bias1 : bias0;
				end = second ? e1 : e0;
				outp = second ? out1 : out0;
			} else {
				const uint32_t i = seg_of_x<CPT>(s, nreset, x);
				bias = a0 + s.segj[i];
				end = s.sege[i];
				outp = a.rbsp + s.segg[i];
			}
			if (x + 16 > end)
				continue; /* inside the padding of an empty segment */
			const uint32_t S = x - bias + s.T[u];
			const uint32_t wi = S >> 2, sh = (S & 3) * 8;
			const uint32_t y0 = raw32[wi], y1 = raw32[wi + 1], y2 = raw32[wi + 2];
			const uint32_t y3 = raw32[wi + 3], y4 = raw32[wi + 4];
			stg_stream16_free(outp + x, make_uint4(__funnelshift_r(y0, y1, sh), __funnelshift_r(y1, y2, sh),
							       __funnelshift_r(y2, y3, sh), __funnelshift_r(y3, y4, sh)));
		}
	}
	if (tid == 0)
		trace_mark(a, s.tile, 7);
	/* units with a deletion inside */
	const uint32_t nd = s.ndirty;
	if (nd <= (uint32_t)C::LIST_CAP) {
		for (uint32_t k = tid; k < nd; k += kT) {
			const uint32_t u = s.dl[k];
			dirty_unit<CPT>(s, u, seg_of_x<CPT>(s, nreset, 16 * u), a0, a.rbsp);
		}
	} else {
		for (uint32_t u = tid; u < NU; u += kT)
			if ((s.D[u >> 5] >> (u & 31)) & 1) {
				const uint32_t i = seg_of_x<CPT>(s, nreset, 16 * u);
				const bool edge = (16 * u < (uint32_t)s.segx[i]) || (16 * u + 16 > (uint32_t)s.sege[i]);
				if (!edge)
					dirty_unit<CPT>(s, u, i, a0, a.rbsp);
			}
	}
	/* partial first / last unit of every segment (threads from the top, the list ran bottom up) */
	for (uint32_t k = tid; k < nseg; k += kT) {
		const uint32_t i = k;
		const uint32_t xs = s.segx[i], xe = s.sege[i];
		if (xs >= xe)
			continue;
		const uint32_t uf = xs >> 4, ul = (xe - 1) >> 4;
		if (xs & 15)
			dirty_unit<CPT>(s, uf, i, a0, a.rbsp);
		if ((xe & 15) && !(ul == uf && (xs & 15)))
			dirty_unit<CPT>(s, ul, i, a0, a.rbsp);
	}
}

/* a reset at r (first byte of a NAL): the three bytes before it are a start code */
template <int CPT> __device__ __forceinline__ bool is_reset(const Smem<CPT> &s, uint32_t r)
{
	const uint8_t *raw = s.raw + 16;
	return raw[(int)r - 3] == 0 && raw[(int)r - 2] == 0 && raw[(int)r - 1] == 1;
}

/* tiles with more resets than the list holds: every thread walks its own 16*CPT bytes and
 * copies, byte by byte, the segments that START there (slow, only for start-code-dense input) */
template <int CPT>
__device__ __forceinline__ void emit_bytewise(Smem<CPT> &s, const ScanArgs &a, uint64_t tile_off, uint32_t nvalid,
					      uint64_t b0)
{
	using C = Cfg<CPT>;
	const uint32_t tid = threadIdx.x;
	const uint8_t *raw = s.raw + 16;
	const uint32_t p0 = tid * CPT * 16, p1 = p0 + CPT * 16;
	for (uint32_t p = p0; p < p1 && p < nvalid; p++) {
		const bool head = p == 0;
		if (!head && !is_reset<CPT>(s, p))
			continue;
		if (head && is_reset<CPT>(s, 0))
			b0 = 0;
		uint64_t out = tile_off + p - (head ? b0 : 0);
		for (uint32_t x = p; x < nvalid && x < (uint32_t)C::TILE; x++) {
			if (x > p && is_reset<CPT>(s, x))
				break;
			if ((s.M[x >> 4] >> (x & 15)) & 1)
				continue;
			a.rbsp[out++] = raw[x];
		}
	}
}

template <int CPT, bool STRIP, int MINB = 4>
__global__ void __launch_bounds__(kT, MINB) scan5_kernel(const ScanArgs a)
{
	using C = Cfg<CPT>;
	__shared__ Smem<CPT> s;

	const uint32_t tid = threadIdx.x;
	const uint32_t lane = tid & 31, warp = tid >> 5;
	uint32_t *raw32 = (uint32_t *)(s.raw + 16);

	/* ---- P0: ticket, bulk load, clear ---- */
	if (tid == 0) {
		const uint32_t t0 = atomicAdd(a.ticket, 1u) + 1u;
		s.tile = t0;
		const uint64_t off = (uint64_t)t0 * C::TILE;
		if (a.len - off >= (uint64_t)C::TILE)
			bulk_load_start(s.raw + 16, a.in + off, C::TILE, &s.bar);
		s.anyev = 0;
		s.ndirty = 0;
		s.ndel = 0;
		s.nreset = 0;
		s.big = 0;
		s.rlast = 0xffffffffu;
		s.nev = 0;
		s.nsc = 0;
		s.evbase = 0;
		s.b0 = 0;
	}
	if (STRIP) { /* M = 0 (16-byte stores); only candidate chunks are written again */
		uint4 *m4 = (uint4 *)s.M;
		for (uint32_t i = tid; i < (uint32_t)(C::NCH + 16) / 8; i += kT)
			m4[i] = make_uint4(0, 0, 0, 0);
	}
	for (uint32_t i = tid; i < (uint32_t)C::NCH / 32; i += kT)
		s.EVB[i] = 0;
	__syncthreads();

	const uint32_t t = s.tile;
	if (tid == 0)
		trace_mark(a, t, 0);
	const uint64_t tile_off = (uint64_t)t * C::TILE;
	const uint64_t rem = a.len - tile_off;
	const bool full = rem >= (uint64_t)C::TILE;
	const uint32_t nvalid = full ? (uint32_t)C::TILE : (uint32_t)rem;

	if (tid == 32)
		raw32[-1] = tile_off ? ldg_u32(a.in + tile_off - 4) : a.halo_left;
	if (tid == 64)
		raw32[C::TILE / 4] = tile_off + C::TILE + 4 <= a.len ? ldg_u32(a.in + tile_off + C::TILE)
								      : annexb::edge_word(a, tile_off + C::TILE);
	if (full) {
		bulk_load_wait(&s.bar);
	} else {
		for (uint32_t c = tid; c < (uint32_t)C::NCH; c += kT) {
			const uint64_t o = tile_off + (uint64_t)c * 16;
			uint4 v;
			if (o + 16 <= a.len)
				v = ldg_stream16(a.in + o);
			else
				v = make_uint4(annexb::edge_word(a, o), annexb::edge_word(a, o + 4),
					       annexb::edge_word(a, o + 8), annexb::edge_word(a, o + 12));
			*(uint4 *)(raw32 + 4 * c) = v;
		}
	}
	__syncthreads();
	if (tid == 0)
		trace_mark(a, t, 1);

	/* ---- P1: classify, two levels.  A chunk matters only if some byte <= 3 follows two zero
	 * bytes (an EPB, or the third byte of a start code / terminator): one cheap SIMD-in-register
	 * test per chunk finds those candidates (11 % at P(00) = 3/16); they are compacted per warp
	 * (list aliased on T, which is not in use yet) and only they get the exact masks. ---- */
	{
		const uint32_t k1 = 0x01010101u, kfc = 0xfcfcfcfcu;
		uint16_t *cand = s.T + warp * (32 * CPT);
		uint32_t ntot = 0;
#pragma unroll 2
		for (int i = 0; i < CPT; i++) {
			const uint32_t c = warp * (32 * CPT) + i * 32 + lane;
			const uint4 v = *(const uint4 *)(raw32 + 4 * c);
			const uint32_t pw = raw32[4 * (int)c - 1], nw = raw32[4 * c + 4];
			const uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
			/* byte p of Xk is 0 <=> b[p-2] = b[p-1] = 0 and b[p] <= 3 */
			const uint32_t X0 = __funnelshift_l(pw, w0, 16) | __funnelshift_l(pw, w0, 8) | (w0 & kfc);
			const uint32_t X1 = __funnelshift_l(w0, w1, 16) | __funnelshift_l(w0, w1, 8) | (w1 & kfc);
			const uint32_t X2 = __funnelshift_l(w1, w2, 16) | __funnelshift_l(w1, w2, 8) | (w2 & kfc);
			const uint32_t X3 = __funnelshift_l(w2, w3, 16) | __funnelshift_l(w2, w3, 8) | (w3 & kfc);
			const uint32_t X4 = __funnelshift_l(w3, nw, 16) | __funnelshift_l(w3, nw, 8) | (nw & kfc) | 0xffff0000u;
			const uint32_t acc = ((X0 - k1) & ~X0) | ((X1 - k1) & ~X1) | ((X2 - k1) & ~X2) |
					     ((X3 - k1) & ~X3) | ((X4 - k1) & ~X4);
			const bool hit = (acc & 0x80808080u) != 0;
			if (STRIP && !full) { /* bytes past the end of the shard count as deleted */
				const uint32_t p0 = c * 16;
				const uint32_t nv = p0 >= nvalid ? 0u : (nvalid - p0 >= 16u ? 16u : nvalid - p0);
				const uint32_t pad = ~((1u << nv) - 1u) & 0xffffu;
				if (pad) {
					list_chunk<CPT>(s, c);
					s.M[c] = (uint16_t)pad;
				}
			}
			const uint32_t bal = __ballot_sync(FULL_MASK, hit);
			if (hit)
				cand[ntot + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = (uint16_t)c;
			ntot += (uint32_t)__popc(bal);
		}
		__syncwarp();
		const uint32_t k3 = 0x03030303u;
		for (uint32_t j = lane; j < ntot; j += 32) {
			const uint32_t c = cand[j];
			const uint4 v = *(const uint4 *)(raw32 + 4 * c);
			const uint32_t pw = raw32[4 * (int)c - 1], nw = raw32[4 * c + 4];
			const uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
			const uint32_t AB0 = __funnelshift_l(pw, w0, 16) | __funnelshift_l(pw, w0, 8);
			const uint32_t AB1 = __funnelshift_l(w0, w1, 16) | __funnelshift_l(w0, w1, 8);
			const uint32_t AB2 = __funnelshift_l(w1, w2, 16) | __funnelshift_l(w1, w2, 8);
			const uint32_t AB3 = __funnelshift_l(w2, w3, 16) | __funnelshift_l(w2, w3, 8);
			const uint32_t AB4 = __funnelshift_l(w3, nw, 16) | __funnelshift_l(w3, nw, 8) | 0xffff0000u;
			/* third byte 0 or 1 somewhere: a boundary event may start in this chunk (rare;
			 * most candidates are emulation prevention bytes) */
			const uint32_t kfe = 0xfefefefeu;
			const uint32_t evt = annexb::haszero(AB0 | (w0 & kfe)) | annexb::haszero(AB1 | (w1 & kfe)) |
					     annexb::haszero(AB2 | (w2 & kfe)) | annexb::haszero(AB3 | (w3 & kfe)) |
					     annexb::haszero(AB4 | (nw & kfe));
			if (evt) {
				const SlowMasks m = annexb::slow_masks(pw, w0, w1, w2, w3, nw);
				const uint32_t p0 = c * 16;
				const uint32_t nv = p0 >= nvalid ? 0u : (nvalid - p0 >= 16u ? 16u : nvalid - p0);
				if (m.ev16 & ((1u << nv) - 1u)) {
					atomicOr(&s.EVB[c >> 5], 1u << (c & 31));
					s.anyev = 1;
				}
			}
			if (STRIP) {
				const uint32_t del = msb_to_nib(annexb::zero_bytes_msb(AB0 | (w0 ^ k3))) |
						     msb_to_nib(annexb::zero_bytes_msb(AB1 | (w1 ^ k3))) << 4 |
						     msb_to_nib(annexb::zero_bytes_msb(AB2 | (w2 ^ k3))) << 8 |
						     msb_to_nib(annexb::zero_bytes_msb(AB3 | (w3 ^ k3))) << 12;
				const uint32_t old = s.M[c];
				if (del) {
					s.M[c] = (uint16_t)(old | del);
					if (!old)
						list_chunk<CPT>(s, c);
				}
			}
		}
	}
	__syncthreads();

	/* ---- P1b (warp 0): events of the tile: counts and reset points, in stream order.  A
	 * start code that begins in the last 3 bytes of the previous tile resets inside this
	 * one (its event belongs to the previous tile). ---- */
	if (warp == 0) {
		uint32_t nev = 0, nsc = 0, nres = 0, rlast = 0xffffffffu;
		bool big = false;
		if (tile_off != 0 || a.halo_left != 0xffffffffu) {
			for (uint32_t r = 0; r < 3 && r < nvalid; r++)
				if (is_reset<CPT>(s, r)) {
					if (lane == 0)
						s.rl[0] = (uint16_t)r;
					nres = 1;
					rlast = r;
				}
		}
		if (s.anyev) {
			for (uint32_t wg = 0; wg < (uint32_t)C::NCH / 32; wg += 32) {
				uint32_t nz = __ballot_sync(FULL_MASK, wg + lane < (uint32_t)C::NCH / 32 && s.EVB[wg + lane] != 0);
				while (nz) {
					const uint32_t wd = wg + (uint32_t)__ffs((int)nz) - 1;
					nz &= nz - 1;
					const uint32_t bits = s.EVB[wd];
					const uint32_t c = wd * 32 + lane;
					SlowMasks m;
					m.ev16 = m.sc16 = m.insc16 = 0;
					if ((bits >> lane) & 1)
						m = chunk_masks<CPT>(s, c, nvalid);
					/* resets of this chunk's start codes that still fall inside the tile */
					uint32_t rs = 0;
					for (uint32_t e = m.sc16; e;) {
						const uint32_t j = (uint32_t)__ffs((int)e) - 1;
						e &= e - 1;
						if (c * 16 + j + 3 < nvalid)
							rs |= 1u << j;
					}
					uint32_t cnt = (uint32_t)__popc(rs), inc = cnt;
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
						if (lane >= (uint32_t)d)
							inc += o;
					}
					uint32_t slot = nres + inc - cnt;
					for (uint32_t e = rs; e; slot++) {
						const uint32_t j = (uint32_t)__ffs((int)e) - 1;
						e &= e - 1;
						if (slot < (uint32_t)kRcap)
							s.rl[slot] = (uint16_t)(c * 16 + j + 3);
					}
					const uint32_t tot = __shfl_sync(FULL_MASK, inc, 31);
					const uint32_t bres = __ballot_sync(FULL_MASK, rs != 0);
					if (bres) {
						const int tl = 31 - __clz((int)bres);
						const uint32_t top = (uint32_t)(31 - __clz((int)rs | 1));
						rlast = __shfl_sync(FULL_MASK, (wd * 32 + lane) * 16 + top + 3, tl);
					}
					nres += tot;
					uint32_t ne = (uint32_t)__popc(m.ev16), ns = (uint32_t)__popc(m.sc16);
#pragma unroll
					for (int d = 16; d >= 1; d >>= 1) {
						ne += __shfl_xor_sync(FULL_MASK, ne, d);
						ns += __shfl_xor_sync(FULL_MASK, ns, d);
					}
					nev += ne;
					nsc += ns;
				}
			}
		}
		big = nres > (uint32_t)kRcap;
		if (lane == 0) {
			s.nreset = nres;
			s.big = big ? 1u : 0u;
			s.rlast = rlast;
			s.nev = nev;
			s.nsc = nsc;
			if (nev)
				s.evbase = atomicAdd(a.ev_cursor, nev) + 1u;
		}
	} else if (STRIP) {
		/* meanwhile: EPBs per thread group (CPT chunks), for the block prefix of P2 */
		for (uint32_t g = tid - 32; g < (uint32_t)kT; g += kT - 32) {
			uint32_t e = 0;
#pragma unroll
			for (int k = 0; k < CPT; k++)
				e += (uint32_t)__popc(s.M[g * CPT + k]);
			s.epre[g] = e;
		}
	}
	__syncthreads();

	/* ---- P2: EPB counts (the per-group sums were taken by warps 1..7 during P1b):
	 * exclusive block prefix per thread group, tile aggregate ---- */
	uint32_t mine = STRIP ? s.epre[tid] : 0u;
	uint32_t inc = mine;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
		if (lane >= (uint32_t)d)
			inc += o;
	}
	if (lane == 31)
		s.wsum[warp] = inc;
	__syncthreads();
	uint32_t wbase = 0, total = 0;
#pragma unroll
	for (int j = 0; j < kW; j++) {
		const uint32_t x = s.wsum[j];
		if ((uint32_t)j < warp)
			wbase += x;
		total += x;
	}
	s.epre[tid] = wbase + inc - mine;
	if (tid == 0)
		s.eall = total;
	__syncthreads();

	const uint32_t nreset = s.nreset;
	const bool big = s.big != 0;
	const bool has = nreset != 0;
	uint64_t *dt = a.desc + (uint64_t)t * 4;
	if (tid == 0) {
		/* bytes past the end of the shard count as deleted in M: keep them out of the chain */
		const uint32_t pad = (uint32_t)C::TILE - nvalid;
		const uint32_t epb_all = STRIP ? total - pad : 0u;
		const uint32_t e_tail = has ? epb_all - (STRIP ? epb_before<CPT>(s, s.rlast) : 0u) : epb_all;
		st_relaxed_u64(dt, pack_agg(e_tail, has));
		if (has || t == 0)
			st_relaxed_u64(dt + 1, (uint64_t)e_tail);
		/* tile 0 of a shard: a start code begun in the previous shard ends here (bytes 0..1
		 * are not NAL data): tell finalize where the shard's first NAL byte is */
		const uint32_t r0 = (t == 0 && has && s.rl[0] < 3) ? s.rl[0] : 0u;
		st_relaxed_u64(dt + 3, (uint64_t)epb_all | (uint64_t)r0 << 32);
		trace_mark(a, t, 2);
	}

	if (STRIP) {
		/* ---- P3: warp 0: shift at the tile start (short look-back), then the segment table
		 * (x start / end, padding, global offset of x = 0) and the partial first / last units of
		 * the segments marked in D; the other warps clear T meanwhile ---- */
		const uint32_t head_hi = big ? (uint32_t)C::TILE : (has ? s.rl[0] : (uint32_t)C::TILE);
		if (warp == 0) {
			for (uint32_t i = lane; i < (uint32_t)(C::NUCAP + 31) / 32; i += 32)
				s.D[i] = 0;
			uint64_t b0 = 0;
			if (t > 0 && head_hi > 0) {
				b0 = lookback(a, t, lane);
				if (lane == 0 && !has)
					st_relaxed_u64(dt + 1, b0 + (uint64_t)(total - ((uint32_t)C::TILE - nvalid)));
			}
			__syncwarp();
			if (lane == 0) {
				s.b0 = b0;
				trace_mark(a, t, 3);
				const uint32_t a0 = (uint32_t)((0 - b0) & 15);
				uint32_t J = 0, prev = 0, xs_prev = a0;
				s.segx[0] = (uint16_t)a0;
				s.segj[0] = 0;
				s.segg[0] = (int64_t)tile_off - (int64_t)b0 - (int64_t)a0;
				if (!big) {
					for (uint32_t i = 0; i <= nreset; i++) {
						/* close segment `prev` at the next reset (or the tile end) */
						const uint32_t r = i < nreset ? (uint32_t)s.rl[i] : (uint32_t)C::TILE;
						const uint32_t kb = i < nreset ? r - epb_before<CPT>(s, r) : (uint32_t)C::TILE - s.eall;
						const uint32_t xend = a0 + kb + J;
						s.sege[prev] = (uint16_t)xend;
						if (xs_prev < xend) { /* partial units are byte-exact and stay off the dirty list */
							if (xs_prev & 15)
								s.D[xs_prev >> 9] |= 1u << ((xs_prev >> 4) & 31);
							if (xend & 15)
								s.D[(xend - 1) >> 9] |= 1u << (((xend - 1) >> 4) & 31);
						}
						if (i == nreset)
							break;
						J += (r - xend) & 15;
						xs_prev = a0 + kb + J;
						s.segx[i + 1] = (uint16_t)xs_prev;
						s.segj[i + 1] = (uint16_t)J;
						s.segg[i + 1] = (int64_t)tile_off + (int64_t)r - (int64_t)xs_prev;
						prev = i + 1;
					}
				}
			}
		} else {
			uint32_t *t32 = (uint32_t *)s.T;
			for (uint32_t i = tid - 32; i < (uint32_t)C::NUCAP / 2; i += kT - 32)
				t32[i] = 0;
		}
		__syncthreads();
		const uint32_t a0 = (uint32_t)((0 - s.b0) & 15);
		if (big)
			emit_bytewise<CPT>(s, a, tile_off, nvalid, s.b0);
		else
			emit_tile<CPT>(s, a, tile_off, nreset, a0);
	} else if (!has && t > 0 && tid == 0) {
		/* scan only: no shifts to chain, but keep the descriptors complete */
		st_relaxed_u64(dt + 1, 0);
	}
	if (tid == 0)
		trace_mark(a, t, 4);

	/* ---- P5 (last warp): event records; nothing waits on them ---- */
	if (warp == kW - 1) {
		const uint32_t nev = s.nev;
		if (lane == 0)
			dt[2] = (uint64_t)s.evbase | (uint64_t)nev << 32 | (uint64_t)s.nsc << 48;
		if (nev) {
			uint32_t run = 0;
			const uint32_t base = s.evbase;
			for (uint32_t wg = 0; wg < (uint32_t)C::NCH / 32; wg += 32) {
				uint32_t nz = __ballot_sync(FULL_MASK, wg + lane < (uint32_t)C::NCH / 32 && s.EVB[wg + lane] != 0);
				while (nz) {
					const uint32_t wd = wg + (uint32_t)__ffs((int)nz) - 1;
					nz &= nz - 1;
					const uint32_t bits = s.EVB[wd];
					const uint32_t c = wd * 32 + lane;
					SlowMasks m;
					m.ev16 = m.sc16 = m.insc16 = 0;
					if ((bits >> lane) & 1)
						m = chunk_masks<CPT>(s, c, nvalid);
					const uint32_t cnt = (uint32_t)__popc(m.ev16);
					uint32_t einc = cnt;
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const uint32_t o = __shfl_up_sync(FULL_MASK, einc, d);
						if (lane >= (uint32_t)d)
							einc += o;
					}
					uint32_t idx = run + einc - cnt;
					for (uint32_t e = m.ev16; e; idx++) {
						const uint32_t j = (uint32_t)__ffs((int)e) - 1;
						e &= e - 1;
						const uint32_t q = c * 16 + j;
						const uint64_t slot = (uint64_t)(base + idx);
						if (slot < a.ev_cap)
							a.evbuf[slot] = pack_event(tile_off + q, STRIP ? epb_before<CPT>(s, q) : 0u,
										   (m.sc16 >> j) & 1);
					}
					run += __shfl_sync(FULL_MASK, einc, 31);
				}
			}
		}
	}
}

/*
 * Finalize: order the event records by tile, pair every start code with the next event, derive
 * the RBSP lengths from the prefix of per-tile EPB counts, fill the NAL table and the result.
 * Three small kernels (~1 % of the scan):
 *   fin_tiles   a tile per thread: block-local exclusive prefix of (events, start codes, EPBs)
 *   fin_order   a tile per thread: + earlier blocks' totals; events -> ordered array with NAL
 *               index and EPB prefix
 *   fin_head    totals and the result fields that need no reduction
 *   fin_table   a thread per ordered event: table entries, RBSP byte sum
 */
struct FinArgs {
	const uint64_t *desc; /* 4 words per tile */
	uint32_t num_tiles;
	uint32_t tile_bytes;
	const uint64_t *evbuf;
	uint64_t ev_cap;
	uint64_t *ordered;  /* 3 words per event: record, EPB prefix, NAL index */
	uint64_t *tile_pre; /* 2 words per tile: ev_off | sc_off << 32, EPB prefix */
	uint64_t *totals;   /* 4 words: events, start codes, EPBs, spare */
	uint64_t *blk_tot;  /* 3 words per fin_tiles block */
	uint64_t len, base;
	uint64_t *nal_start, *nal_end, *nal_rbsp, *nal_rbsp_len;
	uint64_t nal_cap;
	struct h264gpu_scan_result *result;
	uint32_t has_right, strip, assume_in;
};

constexpr int kFinT = 1024;

/* a tile per thread, kFinT tiles per block: block-local exclusive prefix + block totals */
__global__ void __launch_bounds__(kFinT, 1) fin_tiles(const FinArgs f)
{
	__shared__ uint64_t sh[kFinT][3];
	const uint32_t tid = threadIdx.x;
	const uint64_t t = (uint64_t)blockIdx.x * kFinT + tid;
	uint64_t nev = 0, nsc = 0, epb = 0;
	if (t < f.num_tiles) {
		const uint64_t ev = f.desc[t * 4 + 2];
		nev = (ev >> 32) & 0xffffu;
		nsc = (ev >> 48) & 0x7fffu;
		epb = f.desc[t * 4 + 3] & 0xffffffffull;
	}
	sh[tid][0] = nev;
	sh[tid][1] = nsc;
	sh[tid][2] = epb;
	__syncthreads();
	for (uint32_t d = 1; d < (uint32_t)kFinT; d <<= 1) {
		uint64_t x0 = 0, x1 = 0, x2 = 0;
		if (tid >= d) {
			x0 = sh[tid - d][0];
			x1 = sh[tid - d][1];
			x2 = sh[tid - d][2];
		}
		__syncthreads();
		sh[tid][0] += x0;
		sh[tid][1] += x1;
		sh[tid][2] += x2;
		__syncthreads();
	}
	if (t < f.num_tiles) {
		f.tile_pre[2 * t] = (sh[tid][0] - nev) | (sh[tid][1] - nsc) << 32;
		f.tile_pre[2 * t + 1] = sh[tid][2] - epb;
	}
	if (tid == kFinT - 1) {
		f.blk_tot[3 * (uint64_t)blockIdx.x] = sh[tid][0];
		f.blk_tot[3 * (uint64_t)blockIdx.x + 1] = sh[tid][1];
		f.blk_tot[3 * (uint64_t)blockIdx.x + 2] = sh[tid][2];
	}
}

/* a tile per thread (256 per block, all inside one fin_tiles block): add the totals of the
 * earlier fin_tiles blocks, then write the tile's events to their ordered places */
__global__ void __launch_bounds__(256) fin_order(const FinArgs f)
{
	__shared__ uint64_t grp[3];
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t g = (blockIdx.x * blockDim.x) / kFinT; /* earlier fin_tiles blocks */
	if (threadIdx.x < 32) {
		uint64_t a0 = 0, a1 = 0, a2 = 0;
		for (uint32_t j = threadIdx.x; j < g; j += 32) {
			a0 += f.blk_tot[3 * (uint64_t)j];
			a1 += f.blk_tot[3 * (uint64_t)j + 1];
			a2 += f.blk_tot[3 * (uint64_t)j + 2];
		}
#pragma unroll
		for (int d = 16; d >= 1; d >>= 1) {
			a0 += __shfl_xor_sync(FULL_MASK, a0, d);
			a1 += __shfl_xor_sync(FULL_MASK, a1, d);
			a2 += __shfl_xor_sync(FULL_MASK, a2, d);
		}
		if (threadIdx.x == 0) {
			grp[0] = a0;
			grp[1] = a1;
			grp[2] = a2;
		}
	}
	__syncthreads();
	if (t >= f.num_tiles)
		return;
	const uint64_t ev = f.desc[(uint64_t)t * 4 + 2];
	const uint32_t n = (uint32_t)((ev >> 32) & 0xffffu);
	if (n == 0)
		return;
	const uint64_t basei = ev & 0xffffffffull;
	const uint64_t pre = f.tile_pre[2 * (uint64_t)t];
	const uint64_t e = (pre & 0xffffffffull) + grp[0], gp = f.tile_pre[2 * (uint64_t)t + 1] + grp[2];
	uint64_t k = (pre >> 32) + grp[1];
	for (uint32_t i = 0; i < n; i++) {
		if (e + i >= f.ev_cap || basei + i >= f.ev_cap)
			break;
		const uint64_t rec = f.evbuf[basei + i];
		f.ordered[3 * (e + i)] = rec;
		f.ordered[3 * (e + i) + 1] = gp + ((rec >> 40) & 0x3fffffu);
		f.ordered[3 * (e + i) + 2] = k;
		k += (rec >> 62) & 1;
	}
}

__global__ void __launch_bounds__(256) fin_table(const FinArgs f)
{
	const uint64_t total_ev = f.totals[0], total_sc = f.totals[1], total_epb = f.totals[2];
	const uint64_t nordered = total_ev < f.ev_cap ? total_ev : f.ev_cap;
	const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t rlen = 0;
	if (e < nordered) {
		const uint64_t rec = f.ordered[3 * e];
		if ((rec >> 62) & 1) {
			const uint64_t q = rec & ((1ull << 40) - 1);
			const uint64_t g = f.ordered[3 * e + 1], k = f.ordered[3 * e + 2];
			const uint64_t start = q + 3;
			uint64_t end, gend;
			bool open = false;
			if (e + 1 < nordered) {
				end = f.ordered[3 * (e + 1)] & ((1ull << 40) - 1);
				gend = f.ordered[3 * (e + 1) + 1];
			} else {
				end = f.len;
				gend = total_epb;
				open = true;
			}
			if (start > end)
				end = start; /* start code in the last bytes of the shard: empty so far */
			rlen = (end - start) - (gend - g);
			if (k < f.nal_cap) {
				f.nal_start[k] = f.base + start;
				if (!open || !f.has_right)
					f.nal_end[k] = f.base + end;
				if (f.nal_rbsp)
					f.nal_rbsp[k] = start;
				if (f.nal_rbsp_len)
					f.nal_rbsp_len[k] = f.strip ? rlen : 0;
			}
		}
	}
	/* RBSP byte sum: warp reduce, one atomic per warp */
#pragma unroll
	for (int d = 16; d >= 1; d >>= 1)
		rlen += __shfl_xor_sync(FULL_MASK, rlen, d);
	if ((threadIdx.x & 31) == 0 && rlen && f.strip)
		atomicAdd((unsigned long long *)&f.result->rbsp_bytes, (unsigned long long)rlen);
	if (e == 0) {
		/* everything but rbsp_bytes (pre-set by fin_head, accumulated above) */
	}
}

/* result fields that need no reduction; runs before fin_table (which adds the RBSP sum) */
__global__ void fin_head(const FinArgs f)
{
	uint64_t total_ev = 0, total_sc = 0, total_epb = 0;
	const uint32_t nblk = (f.num_tiles + kFinT - 1) / kFinT;
	for (uint32_t j = 0; j < nblk; j++) {
		total_ev += f.blk_tot[3 * (uint64_t)j];
		total_sc += f.blk_tot[3 * (uint64_t)j + 1];
		total_epb += f.blk_tot[3 * (uint64_t)j + 2];
	}
	f.totals[0] = total_ev;
	f.totals[1] = total_sc;
	f.totals[2] = total_epb;
	f.totals[3] = 0;
	const uint64_t nordered = total_ev < f.ev_cap ? total_ev : f.ev_cap;
	struct h264gpu_scan_result r;
	r.n_nal = total_sc;
	r.any_event = total_ev ? 1u : 0u;
	r.first_event_pos = H264GPU_NONE;
	r.first_event_is_sc = 0;
	r.head_bytes = 0;
	r.end_open = 0;
	r.reserved = total_ev > f.ev_cap ? 1u : 0u; /* event buffer overflow: table incomplete */
	if (nordered) {
		const uint64_t first = f.ordered[0], last = f.ordered[3 * (nordered - 1)];
		const uint64_t q = first & ((1ull << 40) - 1);
		r.first_event_pos = f.base + q;
		r.first_event_is_sc = (uint32_t)((first >> 62) & 1);
		r.head_bytes = f.strip ? q - f.ordered[1] : 0;
		r.end_open = (uint32_t)((last >> 62) & 1);
	} else {
		r.head_bytes = f.strip ? f.len - total_epb : 0;
	}
	/* a start code that began in the previous shard: its last bytes are not NAL data */
	const uint64_t r0 = f.desc[3] >> 32;
	r.head_bytes = r.head_bytes > r0 ? r.head_bytes - r0 : 0;
	/* the bytes before the first event belong to a NAL only if an earlier shard left one
	 * open; a shard that assumes so reports them (the merge drops them otherwise) */
	if (!f.assume_in)
		r.head_bytes = 0;
	r.rbsp_bytes = f.strip ? r.head_bytes : 0;
	*f.result = r;
}

/* launch helper shared by the library and the emulator harness */
#ifndef H264_EMU
static inline cudaError_t launch_finalize(const FinArgs &f, uint64_t ev_bound, cudaStream_t st)
{
	fin_tiles<<<(f.num_tiles + kFinT - 1) / kFinT, kFinT, 0, st>>>(f);
	fin_order<<<(f.num_tiles + 255) / 256, 256, 0, st>>>(f);
	fin_head<<<1, 1, 0, st>>>(f);
	const uint64_t nb = (ev_bound + 255) / 256;
	fin_table<<<(uint32_t)(nb ? nb : 1), 256, 0, st>>>(f);
	return cudaGetLastError();
}
#endif

} /* namespace annexb5 */

#endif /* ANNEXB_SCAN5_CUH */
