/*
 * annexb_scan2.cuh — K1/K2 second generation: Annex-B start-code scan, NAL table
 * and emulation-prevention-byte strip in ONE pass (1 B read + <=1 B written per
 * input byte), restructured around what the data looks like:
 *
 *   - deletions (EPBs, start-code bytes, inter-NAL bytes) are SPARSE, so the
 *     RBSP is the input copied through with a slowly growing shift.  The tile
 *     is staged once in shared memory by a bulk async copy (TMA, mbarrier
 *     completion); a per-16-byte-chunk delete mask table M[] is the only thing
 *     classification produces.  Output is then written straight from the staged
 *     input as destination-aligned 16-byte vectors: unit u reads its 16 source
 *     bytes at (16u - a0 + delpre[u]) where delpre is the prefix count of
 *     deletions; only the few units that contain a deletion ("dirty" units) take
 *     a byte-exact path, compacted over the whole CTA.
 *   - boundary events (start codes / 00 00 00 terminators) are RARE, so chunks
 *     that own one are only flagged in a bitmap during classification and one
 *     warp resolves the last-event-wins "inside a NAL" state over flagged chunks.
 *   - tiles are chained with the same single-pass decoupled look-back as v1.
 *
 * Reference behaviour reproduced bit-exactly (Parrot-Developers/libh264):
 *   NAL loop of h264_reader_parse            src/h264_reader.c:133-140
 *   h264_find_nalu / start / end code search src/h264_bitstream.c:87-184
 *   EPB removal in h264_bs_fetch             include/h264/h264_bitstream.h:168-190
 * Closed form: see annexb_scan.cuh.
 */
#ifndef ANNEXB_SCAN2_CUH
#define ANNEXB_SCAN2_CUH

#include "annexb_scan.cuh"

namespace annexb2 {

using annexb::Agg;
using annexb::ScanArgs;
using annexb::SlowMasks;
using annexb::kInvalid;

/* phase timestamps of a tile (diagnostics only; a.trace is NULL in normal runs) */
__device__ __forceinline__ void trace_mark(const ScanArgs &a, uint32_t t, int k)
{
#ifndef H264_EMU
	if (a.trace != nullptr) {
		uint64_t ns;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
		a.trace[(uint64_t)t * 8 + k] = ns;
	}
#else
	(void)a; (void)t; (void)k;
#endif
}

constexpr int kT = 256;
constexpr int kW = kT / 32;

template <int CPT> struct Cfg {
	static constexpr int NCH = kT * CPT;  /* 16-byte chunks per tile */
	static constexpr int TILE = NCH * 16; /* bytes per tile */
	static constexpr int UPT = CPT + 1;   /* output units per thread in the T scan */
	static constexpr int NUCAP = kT * UPT; /* >= NCH + 2 */
	static constexpr int LIST_CAP = NCH / 4; /* deleting chunks / dirty units listed per tile;
						    denser tiles take the unlisted (slow) paths */
};

template <int CPT> struct __align__(128) Smem {
	uint8_t raw[16 + Cfg<CPT>::TILE + 48]; /* [left halo pad][tile][right halo pad] */
	uint16_t M[Cfg<CPT>::NCH + 16];        /* delete mask per chunk (bit j = byte j dropped) */
	uint16_t T[Cfg<CPT>::NUCAP];           /* deletions binned per output unit -> inclusive prefix */
	uint32_t D[(Cfg<CPT>::NUCAP + 31) / 32]; /* dirty output units */
	uint32_t EVB[Cfg<CPT>::NCH / 32];      /* chunks that own a boundary event */
	uint16_t cl[Cfg<CPT>::LIST_CAP];       /* chunks that delete something (unordered) */
	uint16_t dl[Cfg<CPT>::LIST_CAP];       /* dirty output units (unordered) */
	uint32_t pre[kT];                      /* exclusive kept prefix per thread group: head | body<<16 */
	uint32_t wsum[kW];
	uint32_t wsum2[kW];
	uint64_t bar; /* mbarrier of the bulk load */
	uint64_t pin[2];
	uint32_t tile, anyev, ndirty, ndel, qfirst, tn, tflags;
};

/* msb-per-byte mask (bits 7,15,23,31) -> 4-bit mask, bit j = byte j */
__device__ __forceinline__ uint32_t msb_to_nib(uint32_t m)
{
	return (m * 0x00204081u) >> 28;
}

/* exact boundary-event masks of chunk c from the staged tile (positions >= nvalid cut) */
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

template <int CPT> __device__ __forceinline__ uint32_t next_event_chunk(const Smem<CPT> &s, uint32_t c)
{
	uint32_t wd = c >> 5;
	uint32_t b = s.EVB[wd] & ~((2u << (c & 31)) - 1u);
	while (!b) {
		if (++wd >= (uint32_t)Cfg<CPT>::NCH / 32)
			return Cfg<CPT>::NCH;
		b = s.EVB[wd];
	}
	return wd * 32 + (uint32_t)__ffs((int)b) - 1;
}

/* kept bytes of chunk (mask mk, first byte p0) among its bits `sel`, split at qf */
__device__ __forceinline__ void count_chunk(uint32_t mk, uint32_t p0, uint32_t qf, uint32_t sel,
					    uint32_t &h, uint32_t &b)
{
	const uint32_t kept = ~mk & sel & 0xffffu;
	if (p0 + 16 <= qf) {
		h += (uint32_t)__popc(kept);
	} else if (p0 >= qf) {
		b += (uint32_t)__popc(kept);
	} else {
		const uint32_t low = (1u << (qf - p0)) - 1u;
		h += (uint32_t)__popc(kept & low);
		b += (uint32_t)__popc(kept & ~low);
	}
}

/* output bytes of this tile that precede source position p */
template <int CPT>
__device__ __forceinline__ uint32_t out_before(const Smem<CPT> &s, uint32_t p, uint32_t qf, bool sc_in)
{
	const uint32_t g = p / (16 * CPT);
	const uint32_t pk = s.pre[g];
	uint32_t h = pk & 0xffffu, b = pk >> 16;
	const uint32_t cl = p >> 4;
	for (uint32_t c = g * CPT; c < cl; c++)
		count_chunk(s.M[c], c * 16, qf, 0xffffu, h, b);
	count_chunk(s.M[cl], cl * 16, qf, (1u << (p & 15)) - 1u, h, b);
	return (sc_in ? h : 0u) + b;
}

/* tile-range aggregate and its (non-commutative) fold: L is the range LEFT of R */
struct Fold {
	uint32_t h, b, n; /* kept before the first event, kept after it, start codes */
	bool ev, st;      /* has an event, last event is a start code */
};
__device__ __forceinline__ Fold fold(const Fold &L, const Fold &R)
{
	Fold o;
	o.b = L.b + R.b + ((L.ev && L.st) ? R.h : 0u);
	o.h = L.ev ? L.h : L.h + R.h;
	o.n = L.n + R.n;
	o.ev = L.ev || R.ev;
	o.st = R.ev ? R.st : L.st;
	return o;
}

constexpr int kLB = 1; /* predecessor tiles probed per lane and round (window = 32 * kLB) */

/*
 * Single-pass chained scan: publish this tile's aggregate, look back over the
 * predecessors (aggregate or inclusive prefix, whichever is there), publish the
 * inclusive prefix.  Whole warp 0.  The sustainable tile rate of such a chain is
 * (window / probe latency), so every round probes 128 predecessors: lane l holds
 * tiles j0-4l .. j0-4l-3 (nearest first) and folds them locally before the warp
 * tree.
 */
__device__ __forceinline__ void lookback(const ScanArgs &a, uint32_t t, uint32_t lane, uint32_t th,
					 uint32_t tb, uint32_t tn, bool tev, bool tst,
					 uint64_t &kept_in, uint64_t &nnal_in, bool &sc_in, bool &any_in)
{
	uint64_t *dt = a.desc + (uint64_t)t * 4;
	kept_in = 0;
	nnal_in = 0;
	sc_in = a.init_in != 0;
	any_in = false;
	if (t > 0) {
		if (lane == 0)
			st_relaxed_u64(dt, annexb::pack_agg(th, tb, tn, tev, tst));
		/* acc = fold of the tiles between the prefix tile found and t (exclusive) */
		uint64_t ah = 0, ab = 0, an = 0;
		bool aev = false, ast = false;
		int64_t j0 = (int64_t)t - 1;
		for (bool done = false; !done;) {
			const int64_t jl = j0 - (int64_t)lane * kLB; /* my nearest tile */
			uint64_t agg[kLB], p1[kLB];
			int kp;          /* my first tile carrying a prefix, kLB if none */
			bool ready;      /* everything I need up to kp is there */
			uint32_t pm;
			for (;;) {
#pragma unroll
				for (int k = 0; k < kLB; k++) {
					p1[k] = kInvalid;
					agg[k] = kInvalid;
					if (jl - k >= 0) {
						p1[k] = ld_relaxed_u64(a.desc + (jl - k) * 4 + 1);
						agg[k] = ld_relaxed_u64(a.desc + (jl - k) * 4);
					}
				}
				kp = kLB;
				ready = true;
#pragma unroll
				for (int k = kLB - 1; k >= 0; k--)
					if (!(p1[k] & kInvalid))
						kp = k;
#pragma unroll
				for (int k = 0; k < kLB; k++)
					if (k < kp && (agg[k] & kInvalid))
						ready = false;
				pm = __ballot_sync(FULL_MASK, kp < kLB);
				const uint32_t ok = __ballot_sync(FULL_MASK, ready);
				const uint32_t upto = pm ? ((pm & (0u - pm)) << 1) - 1u : 0xffffffffu;
				if ((ok & upto) == upto)
					break;
				spin_pause(32);
			}
			const int fp = pm ? __ffs((int)pm) - 1 : 32;
			/* local fold of my tiles nearer than kp; lanes beyond fp contribute nothing */
			Fold e;
			e.h = e.b = e.n = 0;
			e.ev = e.st = false;
			if ((int)lane <= fp) {
#pragma unroll
				for (int k = 0; k < kLB; k++) {
					if (k < kp) {
						const Agg g = annexb::unpack_agg(agg[k]);
						Fold L;
						L.h = g.head;
						L.b = g.body;
						L.n = g.nsc;
						L.ev = g.ev;
						L.st = g.st;
						e = fold(L, e);
					}
				}
			}
			uint32_t eh = e.h, eb = e.b, em = e.n | (e.ev ? 1u << 30 : 0u) | (e.st ? 1u << 31 : 0u);
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				uint32_t oh = __shfl_down_sync(FULL_MASK, eh, d);
				uint32_t ob = __shfl_down_sync(FULL_MASK, eb, d);
				uint32_t om = __shfl_down_sync(FULL_MASK, em, d);
				if (lane + d < 32) {
					const bool oev = (om >> 30) & 1, ost = om >> 31;
					const bool mev = (em >> 30) & 1;
					eb = ob + eb + ((oev && ost) ? eh : 0u);
					eh = oev ? oh : oh + eh;
					em = ((om & 0x3fffffffu) + (em & 0x3fffffffu)) |
					     ((oev || mev) ? 1u << 30 : 0u) |
					     (mev ? (em & 0x80000000u) : (om & 0x80000000u));
				}
			}
			{
				const uint32_t wh = __shfl_sync(FULL_MASK, eh, 0);
				const uint32_t wb = __shfl_sync(FULL_MASK, eb, 0);
				const uint32_t wm = __shfl_sync(FULL_MASK, em, 0);
				const bool wev = (wm >> 30) & 1, wst = wm >> 31;
				ab = wb + ab + ((wev && wst) ? ah : 0);
				ah = wev ? wh : wh + ah;
				an += wm & 0x3fffffffu;
				if (!aev)
					ast = wst;
				aev = aev || wev;
			}
			if (fp < 32) {
				uint64_t p2 = kInvalid, P1 = 0;
				if ((int)lane == fp) {
#pragma unroll
					for (int k = 0; k < kLB; k++)
						if (k == kp)
							P1 = p1[k];
					do {
						p2 = ld_relaxed_u64(a.desc + (jl - kp) * 4 + 2);
					} while (p2 & kInvalid);
				}
				p2 = __shfl_sync(FULL_MASK, p2, fp);
				P1 = __shfl_sync(FULL_MASK, P1, fp);
				bool psc = (P1 >> 62) & 1, pany = (P1 >> 61) & 1;
				uint64_t pk = P1 & ((1ull << 61) - 1);
				kept_in = pk + (psc ? ah : 0) + ab;
				nnal_in = p2 + an;
				sc_in = aev ? ast : psc;
				any_in = pany || aev;
				done = true;
			} else {
				j0 -= 32 * kLB;
			}
		}
	}
	const uint64_t kept_out = kept_in + (sc_in ? th : 0u) + tb;
	const uint64_t nnal_out = nnal_in + tn;
	const bool sc_out = tev ? tst : sc_in;
	const bool any_out = any_in || tev;
	if (lane == 0) {
#ifdef H264_EMU
		if (annexb::emu_prefix_every <= 1 || t % annexb::emu_prefix_every == 0)
#endif
		{
			st_relaxed_u64(dt + 1, kept_out | (uint64_t)(sc_out ? 1 : 0) << 62 |
						       (uint64_t)(any_out ? 1 : 0) << 61);
			st_relaxed_u64(dt + 2, nnal_out);
		}
		if (t == a.num_tiles - 1) {
			a.result->n_nal = nnal_out;
			a.result->rbsp_bytes = kept_out;
			a.result->end_open = sc_out ? 1u : 0u;
			a.result->any_event = any_out ? 1u : 0u;
			a.result->reserved = 0;
			if (!any_out)
				a.result->head_bytes = kept_out;
			if (sc_out && !a.has_right && nnal_out >= 1 && nnal_out - 1 < a.nal_cap)
				a.nal_end[nnal_out - 1] = a.base + a.len;
		}
	}
}

/* remember chunk c as "deletes something" (list is unordered; overflow -> unlisted path) */
template <int CPT> __device__ __forceinline__ void list_chunk(Smem<CPT> &s, uint32_t c)
{
	const uint32_t idx = atomicAdd(&s.ndel, 1u);
	if (idx < (uint32_t)Cfg<CPT>::LIST_CAP)
		s.cl[idx] = (uint16_t)c;
}

/*
 * Byte-exact output of unit u (a unit that contains a deletion, or an edge unit).
 * Fast form: the unit's first source byte is known, so take the 32 source bytes from
 * there, squeeze out the deleted ones (delete bits from M) and store 16.  Slow form
 * (first unit shared with the previous tile, or > 16 deletions in the window): walk.
 */
template <int CPT>
__device__ __forceinline__ void dirty_unit(const Smem<CPT> &s, uint32_t u, uint32_t a0,
					   uint32_t tile_kept, uint32_t drop, uint8_t *gout)
{
	const uint32_t x0 = 16 * u;
	const uint32_t lo = x0 > a0 ? x0 : a0;
	const uint32_t xe = a0 + tile_kept;
	const uint32_t hi = x0 + 16 < xe ? x0 + 16 : xe;
	if (lo >= hi)
		return;
	uint8_t *g = gout + x0;
	uint64_t vlo = 0, vhi = 0;
	bool done = false;
	if (x0 >= a0) {
		const uint32_t pos = x0 - a0 + s.T[u];
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
			vlo = q0;
			vhi = q1;
			done = true;
		}
	}
	if (!done) {
		uint32_t pos = x0 >= a0 ? x0 - a0 + s.T[u] : drop;
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
		stg_stream16(g, make_uint4((uint32_t)vlo, (uint32_t)(vlo >> 32), (uint32_t)vhi,
					   (uint32_t)(vhi >> 32)));
	} else {
		for (uint32_t x = lo; x < hi; x++) {
			const uint32_t n = x - x0;
			g[n] = (uint8_t)((n < 8 ? vlo >> (8 * n) : vhi >> (8 * (n - 8))) & 0xff);
		}
	}
}

/*
 * Bin the deletions of chunk c by the 16-byte output unit they land in; a unit that
 * gets a deletion strictly inside it becomes dirty (listed once).
 */
template <int CPT>
__device__ __forceinline__ void bin_chunk(Smem<CPT> &s, uint32_t c, uint32_t qf, bool sc_in,
					  uint32_t drop, uint32_t a0)
{
	const uint32_t p0 = c * 16;
	if (p0 + 16 <= drop)
		return;
	const uint32_t mk = s.M[c];
	const uint32_t low = p0 < drop ? (1u << (drop - p0)) - 1u : 0u;
	const uint32_t keptm = ~mk & ~low & 0xffffu;
	uint32_t dm = mk & ~low;
	if (!dm)
		return;
	/* output bytes before the chunk */
	const uint32_t g0 = c / CPT * CPT;
	const uint32_t pk = s.pre[c / CPT];
	uint32_t h = pk & 0xffffu, b = pk >> 16;
	for (uint32_t cc = g0; cc < c; cc++)
		count_chunk(s.M[cc], cc * 16, qf, 0xffffu, h, b);
	const uint32_t o = (sc_in ? h : 0u) + b;
	uint32_t *t32 = (uint32_t *)s.T;
	while (dm) {
		const uint32_t j = (uint32_t)__ffs((int)dm) - 1;
		const uint32_t r = (uint32_t)__ffs((int)~(dm >> j)) - 1; /* run of deleted bytes */
		const uint32_t x = a0 + o + (uint32_t)__popc(keptm & ((1u << j) - 1u));
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

template <int CPT, bool STRIP, int MINB = (CPT >= 8 ? 4 : 6)>
__global__ void __launch_bounds__(kT, MINB) scan2_kernel(const ScanArgs a)
{
	using C = Cfg<CPT>;
	__shared__ Smem<CPT> s;

	const uint32_t tid = threadIdx.x;
	const uint32_t lane = tid & 31, warp = tid >> 5;
	uint32_t *raw32 = (uint32_t *)(s.raw + 16);

	/* ---- P0: ticket, start the bulk load of the tile, clear the tables ---- */
	if (tid == 0) {
		const uint32_t t0 = atomicAdd(a.ticket, 1u) + 1u;
		s.tile = t0;
		const uint64_t off = (uint64_t)t0 * C::TILE;
		if (a.len - off >= (uint64_t)C::TILE)
			bulk_load_start(s.raw + 16, a.in + off, C::TILE, &s.bar);
		s.anyev = 0;
		s.ndirty = 0;
		s.ndel = 0;
		s.qfirst = C::TILE;
		s.tn = 0;
		s.tflags = 0;
	}
	if (STRIP) {
		uint32_t *t32 = (uint32_t *)s.T;
		for (uint32_t i = tid; i < (uint32_t)C::NUCAP / 2; i += kT)
			t32[i] = 0;
		for (uint32_t i = tid; i < (uint32_t)(C::NUCAP + 31) / 32; i += kT)
			s.D[i] = 0;
		if (tid < 16)
			s.M[C::NCH + tid] = 0;
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
		raw32[C::TILE / 4] = tile_off + C::TILE + 4 <= a.len
					     ? ldg_u32(a.in + tile_off + C::TILE)
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

	/* ---- P1: classify every chunk: delete mask (EPB, start-code bytes, bytes
	 * past the end) and the "owns a boundary event" flag ---- */
	{
		const uint32_t k3 = 0x03030303u, kfe = 0xfefefefeu;
#pragma unroll 2
		for (int i = 0; i < CPT; i++) {
			const uint32_t c = warp * (32 * CPT) + i * 32 + lane;
			const uint4 v = *(const uint4 *)(raw32 + 4 * c);
			const uint32_t pw = raw32[4 * (int)c - 1], nw = raw32[4 * c + 4];
			const uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
			/* byte p of ABk is 0 <=> b[p-2] = b[p-1] = 0 */
			const uint32_t AB0 = __funnelshift_l(pw, w0, 16) | __funnelshift_l(pw, w0, 8);
			const uint32_t AB1 = __funnelshift_l(w0, w1, 16) | __funnelshift_l(w0, w1, 8);
			const uint32_t AB2 = __funnelshift_l(w1, w2, 16) | __funnelshift_l(w1, w2, 8);
			const uint32_t AB3 = __funnelshift_l(w2, w3, 16) | __funnelshift_l(w2, w3, 8);
			const uint32_t AB4 = __funnelshift_l(w3, nw, 16) | __funnelshift_l(w3, nw, 8) | 0xffff0000u;
			/* third byte 0 or 1: a boundary event starts two bytes earlier (rare) */
			const uint32_t evt = annexb::haszero(AB0 | (w0 & kfe)) | annexb::haszero(AB1 | (w1 & kfe)) |
					     annexb::haszero(AB2 | (w2 & kfe)) | annexb::haszero(AB3 | (w3 & kfe)) |
					     annexb::haszero(AB4 | (nw & kfe));
			uint32_t del = 0;
			if (STRIP) {
				/* third byte 3: emulation prevention byte (branch-free) */
				del = msb_to_nib(annexb::zero_bytes_msb(AB0 | (w0 ^ k3))) |
				      msb_to_nib(annexb::zero_bytes_msb(AB1 | (w1 ^ k3))) << 4 |
				      msb_to_nib(annexb::zero_bytes_msb(AB2 | (w2 ^ k3))) << 8 |
				      msb_to_nib(annexb::zero_bytes_msb(AB3 | (w3 ^ k3))) << 12;
			}
			if (evt) {
				const SlowMasks m = annexb::slow_masks(pw, w0, w1, w2, w3, nw);
				const uint32_t p0 = c * 16;
				const uint32_t nv = p0 >= nvalid ? 0u : (nvalid - p0 >= 16u ? 16u : nvalid - p0);
				del |= m.insc16;
				if (m.ev16 & ((1u << nv) - 1u)) {
					atomicOr(&s.EVB[c >> 5], 1u << (c & 31));
					s.anyev = 1;
				}
			}
			if (STRIP) {
				if (!full) {
					const uint32_t p0 = c * 16;
					const uint32_t nv = p0 >= nvalid ? 0u : (nvalid - p0 >= 16u ? 16u : nvalid - p0);
					del |= ~((1u << nv) - 1u) & 0xffffu;
				}
				s.M[c] = (uint16_t)del;
				if (del)
					list_chunk<CPT>(s, c);
			}
		}
	}
	__syncthreads();

	/* ---- P1b (warp 0, only when the tile has events): resolve inside/outside a
	 * NAL around every event, in stream order ---- */
	if (warp == 0 && s.anyev) {
		bool cur_seen = false, cur_sc = false;
		uint32_t nsc_run = 0;
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
			const bool has_ev = m.ev16 != 0;
			const int top = has_ev ? 31 - __clz((int)m.ev16) : 0;
			const bool last_sc = has_ev && ((m.sc16 >> top) & 1);
			const uint32_t bev = __ballot_sync(FULL_MASK, has_ev);
			const uint32_t bls = __ballot_sync(FULL_MASK, last_sc);
			const uint32_t before = bev & ((1u << lane) - 1);
			bool pre_known, pre_in;
			if (before) {
				const int pl = 31 - __clz((int)before);
				pre_known = true;
				pre_in = (bls >> pl) & 1;
			} else {
				pre_known = cur_seen;
				pre_in = cur_sc;
			}
			if (has_ev) {
				const uint32_t fe = (uint32_t)__ffs((int)m.ev16) - 1;
				if (STRIP) {
					const uint32_t bm = (1u << fe) - 1;
					uint32_t inm = 0;
					for (uint32_t e = m.ev16; e;) {
						const int j = __ffs((int)e) - 1;
						e &= e - 1;
						const int nj = e ? __ffs((int)e) - 1 : 16;
						if ((m.sc16 >> j) & 1)
							inm |= ((1u << nj) - 1) & ~((1u << j) - 1);
					}
					const uint32_t keepc = ((pre_known && !pre_in) ? 0u : bm) | inm;
					const uint32_t oldm = s.M[c], newm = oldm | (~keepc & 0xffffu);
					s.M[c] = (uint16_t)newm;
					if (!oldm && newm)
						list_chunk<CPT>(s, c);
					if (!last_sc) {
						const uint32_t nxt = next_event_chunk<CPT>(s, c);
						for (uint32_t c2 = c + 1; c2 < nxt; c2++) {
							if (!s.M[c2])
								list_chunk<CPT>(s, c2);
							s.M[c2] = 0xffffu;
						}
					}
				}
				if (!pre_known)
					s.qfirst = 16 * c + fe;
			}
			uint32_t ns = (uint32_t)__popc(m.sc16);
#pragma unroll
			for (int d = 16; d >= 1; d >>= 1)
				ns += __shfl_xor_sync(FULL_MASK, ns, d);
			nsc_run += ns;
			if (bev) {
				const int tl = 31 - __clz((int)bev);
				cur_seen = true;
				cur_sc = (bls >> tl) & 1;
			}
		  }
		}
		if (lane == 0) {
			s.tn = nsc_run;
			s.tflags = (cur_seen ? 1u : 0u) | (cur_sc ? 2u : 0u);
		}
	}
	__syncthreads();

	/* ---- P2: kept-byte counts (head = before the tile's first event, body = the
	 * rest), exclusive block prefix per group of CPT chunks ---- */
	const uint32_t qf = s.qfirst;
	uint32_t packed = 0;
	if (STRIP) {
		uint32_t h = 0, b = 0;
		const uint32_t c0 = tid * CPT;
#pragma unroll
		for (int k = 0; k < CPT; k++)
			count_chunk(s.M[c0 + k], (c0 + k) * 16, qf, 0xffffu, h, b);
		packed = h | b << 16;
	}
	uint32_t inc = packed;
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
	const uint32_t excl = wbase + inc - packed;
	s.pre[tid] = excl;
	const uint32_t H = total & 0xffffu, B = total >> 16;

	/* ---- P3 (warp 0): chained look-back ---- */
	if (warp == 0) {
		uint64_t kept_in, nnal_in;
		bool sc_in, any_in;
		const uint32_t fl = s.tflags;
		if (lane == 0)
			trace_mark(a, t, 2);
		lookback(a, t, lane, H, B, s.tn, fl & 1, (fl >> 1) & 1, kept_in, nnal_in, sc_in, any_in);
		if (lane == 0) {
			s.pin[0] = kept_in | (uint64_t)(sc_in ? 1 : 0) << 62 | (uint64_t)(any_in ? 1 : 0) << 61;
			s.pin[1] = nnal_in;
			trace_mark(a, t, 3);
		}
	}
	__syncthreads();

	const uint64_t kept_in = s.pin[0] & ((1ull << 61) - 1);
	const bool sc_in = (s.pin[0] >> 62) & 1, any_in = (s.pin[0] >> 61) & 1;
	const uint64_t nnal_in = s.pin[1];
	const uint32_t tile_kept = (sc_in ? H : 0u) + B;
	const uint32_t a0 = (uint32_t)(kept_in & 15);
	const uint32_t drop = sc_in ? 0u : qf; /* source bytes before `drop` are not output */
	const uint32_t NU = (a0 + tile_kept + 15) >> 4;

	if (STRIP && tile_kept != 0) { /* block-uniform */

	/* ---- P4b: bin every deletion by the output unit it lands in ---- */
	{
		if (tid == 0) {
			if (drop) {
				uint32_t *t32 = (uint32_t *)s.T;
				const uint32_t bin = (a0 + 15) >> 4;
				atomicAdd(&t32[bin >> 1], drop << (16 * (bin & 1)));
			}
			if (a0)
				atomicOr(&s.D[0], 1u);
			if ((a0 + tile_kept) & 15)
				atomicOr(&s.D[(NU - 1) >> 5], 1u << ((NU - 1) & 31));
		}
		const uint32_t ndel = s.ndel;
		if (ndel <= (uint32_t)C::LIST_CAP) {
			for (uint32_t i = tid; i < ndel; i += kT)
				bin_chunk<CPT>(s, s.cl[i], qf, sc_in, drop, a0);
		} else {
			for (uint32_t c = tid; c < (uint32_t)C::NCH; c += kT)
				bin_chunk<CPT>(s, c, qf, sc_in, drop, a0);
		}
	}
	__syncthreads();

	/* ---- P5: inclusive prefix of T: delpre[u] = deletions landing at or before 16u ---- */
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

	/* ---- P6: clean units: one aligned 16-byte store each, read through a byte shift ---- */
	uint8_t *gout = a.rbsp + (kept_in - a0);
	for (uint32_t u = tid; u < NU; u += kT) {
		if ((s.D[u >> 5] >> (u & 31)) & 1)
			continue;
		const uint32_t S = 16 * u - a0 + s.T[u];
		const uint32_t wi = S >> 2, sh = (S & 3) * 8;
		const uint32_t x0 = raw32[wi], x1 = raw32[wi + 1], x2 = raw32[wi + 2];
		const uint32_t x3 = raw32[wi + 3], x4 = raw32[wi + 4];
		stg_stream16(gout + 16 * u,
			     make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh),
					__funnelshift_r(x2, x3, sh), __funnelshift_r(x3, x4, sh)));
	}

	/* ---- P7: dirty units (listed by P4b; the two edge units are always tried) ---- */
	const uint32_t nd = s.ndirty;
	if (nd <= (uint32_t)C::LIST_CAP) {
		for (uint32_t i = tid; i < nd; i += kT)
			dirty_unit<CPT>(s, s.dl[i], a0, tile_kept, drop, gout);
		if (tid == kT - 1 && a0)
			dirty_unit<CPT>(s, 0, a0, tile_kept, drop, gout);
		if (tid == kT - 2 && ((a0 + tile_kept) & 15) && !(NU == 1 && a0))
			dirty_unit<CPT>(s, NU - 1, a0, tile_kept, drop, gout);
	} else {
		for (uint32_t u = tid; u < NU; u += kT)
			if ((s.D[u >> 5] >> (u & 31)) & 1)
				dirty_unit<CPT>(s, u, a0, tile_kept, drop, gout);
	}
	} /* STRIP && tile_kept */

	if (tid == 0)
		trace_mark(a, t, 4);
	/* ---- P8 (last warp, only with events): NAL table entries; nothing waits on it ---- */
	if (warp == kW - 1 && s.anyev) {
		bool cur_seen = false, cur_sc = false;
		uint32_t nsc_run = 0;
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
			const bool has_ev = m.ev16 != 0;
			const int top = has_ev ? 31 - __clz((int)m.ev16) : 0;
			const bool last_sc = has_ev && ((m.sc16 >> top) & 1);
			const uint32_t bev = __ballot_sync(FULL_MASK, has_ev);
			const uint32_t bls = __ballot_sync(FULL_MASK, last_sc);
			const uint32_t before = bev & ((1u << lane) - 1);
			bool pre_known, pre_in;
			if (before) {
				const int pl = 31 - __clz((int)before);
				pre_known = true;
				pre_in = (bls >> pl) & 1;
			} else {
				pre_known = cur_seen;
				pre_in = cur_sc;
			}
			const uint32_t ns = (uint32_t)__popc(m.sc16);
			uint32_t sinc = ns;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t o = __shfl_up_sync(FULL_MASK, sinc, d);
				if (lane >= (uint32_t)d)
					sinc += o;
			}
			const uint32_t ns_round = __shfl_sync(FULL_MASK, sinc, 31);
			if (has_ev) {
				bool prev_sc = pre_known ? pre_in : sc_in;
				uint64_t nscb = nnal_in + nsc_run + (sinc - ns);
				bool first = !any_in && !pre_known;
				for (uint32_t e = m.ev16; e;) {
					const int j = __ffs((int)e) - 1;
					e &= e - 1;
					const bool is_sc = (m.sc16 >> j) & 1;
					const uint32_t p = c * 16 + (uint32_t)j;
					const uint64_t q = a.base + tile_off + p;
					const uint64_t keptb = STRIP ? kept_in + out_before<CPT>(s, p, qf, sc_in) : 0;
					if (prev_sc && nscb >= 1 && nscb - 1 < a.nal_cap)
						a.nal_end[nscb - 1] = q;
					if (first) {
						a.result->first_event_pos = q;
						a.result->first_event_is_sc = is_sc ? 1u : 0u;
						a.result->head_bytes = keptb;
						first = false;
					}
					if (is_sc) {
						if (nscb < a.nal_cap) {
							a.nal_start[nscb] = q + 3;
							if (a.nal_rbsp)
								a.nal_rbsp[nscb] = keptb;
						}
						nscb++;
					}
					prev_sc = is_sc;
				}
			}
			nsc_run += ns_round;
			if (bev) {
				const int tl = 31 - __clz((int)bev);
				cur_seen = true;
				cur_sc = (bls >> tl) & 1;
			}
		  }
		}
	}
}

} /* namespace annexb2 */

#endif /* ANNEXB_SCAN2_CUH */
