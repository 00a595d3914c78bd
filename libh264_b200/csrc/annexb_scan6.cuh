/*
 * annexb_scan6.cuh — K1/K2 sixth generation: the in-place layout of gen 5 (NAL k's RBSP starts
 * where the NAL starts, annexb_scan5.cuh) with WARP-AUTONOMOUS tiles.
 *
 * Why: gen 5 ran at 48 % issue-slot utilisation with ~15 k warp-instructions per 32 KiB tile
 * (profiles/r01_scan_kernel_raw.csv): deletions were binned per output unit with shared
 * atomics, a block-wide prefix turned the bins into source offsets, dirty units and chunks went
 * through block-wide lists, and a single warp walked the events and built a segment table while
 * seven waited.  Here every warp owns a contiguous span of ROWS x 512 bytes of the tile and
 * does everything for it alone; the block meets twice (span aggregates, shift at the tile start).
 *
 * The observation that removes the binning: inside a segment (no start code) the output
 * ranges of consecutive 16-byte source chunks partition the output, and a chunk with k
 * deleted bytes covers 16 - k <= 16 output bytes, so it contains AT MOST ONE 16-byte output
 * unit boundary.  With  shift(c) = EPBs since the NAL start before chunk c,
 *
 *     b(c) = shift(c) & 15       bytes from the chunk's first output byte to the next boundary
 *
 * and chunk c is responsible for the unit that starts b(c) kept bytes into it (if b < 16 - k).
 * A chunk without a deletion whose successor has none in its first b bytes copies its unit from
 * source offset 16c + b through a funnel shift (one aligned 16-byte store); the others
 * (~15 %) are listed per warp and squeezed byte-exactly afterwards, 32 at a time.
 *
 * Start codes are rare (one per NAL): the 512-byte row a reset point falls in, the row the
 * shard ends in, and nothing else, is written byte by byte; the units next to such a row or to
 * the tile seam are cut at the seam (each side writes its own bytes).  The byte-exact work
 * (listed chunks, chunks of byte-wise rows) is shared out over the whole block, one item per
 * thread, so that no warp holds the tile back with a second pass.
 *
 * CTAs are persistent (148 x 5, tiles by ticket, taken only when the CTA is ready: a ticket held
 * back delays every later tile's look-back) and pull the tile one grid ahead into L2.
 *
 * Boundary events are owned by the position t = q + 2 of their THIRD byte here (the chunk
 * needs no look-ahead then); the record still carries q.  The launch therefore covers
 * len + 2 bytes (the two bytes of the right shard edge can be third bytes).
 *
 * Descriptor words, event records and the finalize kernels are those of gen 5.
 *
 * Reference behaviour reproduced bit-exactly (Parrot-Developers/libh264):
 *   NAL loop of h264_reader_parse            src/h264_reader.c:133-140
 *   h264_find_nalu / start / end code search src/h264_bitstream.c:87-184
 *   EPB removal in h264_bs_fetch             include/h264/h264_bitstream.h:168-190
 * Closed form: see annexb_scan.cuh.
 */
#ifndef ANNEXB_SCAN6_CUH
#define ANNEXB_SCAN6_CUH

#include "annexb_scan5.cuh"

namespace annexb6 {

using annexb::ScanArgs;
using annexb::kInvalid;
using annexb2::msb_to_nib;
using annexb5::pack_agg;
using annexb5::pack_event;
using annexb2::trace_mark;

constexpr int kT = 256;
constexpr int kW = kT / 32;

template <int ROWS> struct Cfg {
	static constexpr int SPAN_CH = 32 * ROWS;   /* 16-byte chunks per warp span */
	static constexpr int NCH = kW * SPAN_CH;    /* chunks per tile */
	static constexpr int TILE = NCH * 16;       /* bytes per tile */
	static constexpr int NROW = kW * ROWS;      /* 512-byte rows per tile (<= 64) */
};

template <int ROWS> struct __align__(128) Smem {
	uint8_t raw[16 + Cfg<ROWS>::TILE + 32]; /* [left halo pad][tile][pad] */
	uint16_t M[Cfg<ROWS>::NCH + 8];         /* delete mask per chunk (bit j = byte j is an EPB) */
	uint16_t E[Cfg<ROWS>::NCH];             /* EPBs of the span before the chunk; candidate list before that */
	uint8_t dl[Cfg<ROWS>::NCH];             /* per span: chunks that take the byte-exact path */
	int32_t rbrel[kW][ROWS];                /* shift at the row start minus E, rows after a reset */
	uint64_t sp_d0[kW];                     /* shift at the span start */
	uint32_t sp_etail[kW], sp_etot[kW], sp_nev[kW], sp_nsc[kW], sp_nd[kW];
	uint8_t sp_has[kW], sp_rbhas[kW], sp_bw[kW], sp_evrows[kW];
	uint64_t bar;
	uint64_t b0; /* shift at the tile start */
	uint32_t tile, evbase, first_r;
	/* the tile aggregate, kept from its publication to the look-back (pipelined kernel) */
	uint32_t ag_total, ag_nev, ag_nsc, ag_has, ag_evprev;
};

struct TMasks {
	uint32_t ev16, sc16;
};

/* exact boundary-event masks of a chunk by THIRD-byte position t, from the window [-4, 16):
 * bit j of sc16: bytes j-2, j-1, j are 00 00 01; of ev16: a boundary event has its third byte
 * at j (00 00 01, or 00 00 00 not preceded by another zero) */
__device__ __forceinline__ TMasks t_masks(uint32_t pw, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
	using annexb::zmask4;
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

/* bit j: byte j of the chunk is a reset point (the three bytes before it are 00 00 01) */
__device__ __forceinline__ uint32_t reset_mask(uint32_t pw, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
	const uint32_t k1 = 0x01010101u;
#define H264_R6(lo, hi)                                                                                  \
	msb_to_nib(annexb::zero_bytes_msb(__funnelshift_l(lo, hi, 24) | __funnelshift_l(lo, hi, 16) | \
					  (__funnelshift_l(lo, hi, 8) ^ k1)))
	const uint32_t r = H264_R6(pw, w0) | H264_R6(w0, w1) << 4 | H264_R6(w1, w2) << 8 | H264_R6(w2, w3) << 12;
#undef H264_R6
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

/*
 * Byte-exact output of the unit chunk c is responsible for: it starts b kept bytes into the
 * chunk; 16 kept bytes are gathered from there (at most 24 source bytes: an EPB needs two zero
 * bytes before it), cut at `limit` (source position: the next byte-wise row or the tile end).
 * shlo = low bits of the shift at E = 0, basep = where source position 0 of the tile goes at E = 0.
 */
template <int ROWS>
__device__ __forceinline__ void dirty_chunk(const Smem<ROWS> &s, uint32_t c, uint32_t shlo, uint8_t *basep,
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
	const uint32_t avail = limit - pos;
	if (avail < 32)
		dm |= ~0u << avail; /* bytes past the limit: not ours */
	const uint32_t kept32 = 32u - (uint32_t)__popc(dm);
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
		for (uint32_t n = 0; n < nout; n++)
			g[n] = (uint8_t)((n < 8 ? q0 >> (8 * n) : q1 >> (8 * (n - 8))) & 0xff);
	}
}

/*
 * P4: the rows of a span.  bwl: bit 0 = the row before the span is a seam (byte-wise row or
 * tile start), bits 1..ROWS = own rows that go byte by byte, bit ROWS + 1 = the row after the
 * span is a seam.  LEAN (4 spans in 5): no byte-wise row and no reset in the span, one shift
 * base d0 for all rows.
 */
template <int ROWS, bool LEAN>
__device__ __forceinline__ void emit_span(Smem<ROWS> &s, uint32_t warp, uint32_t lane, uint32_t bwl, uint32_t rbhas,
					  uint64_t d0, uint8_t *out_tile, uint32_t nvalid)
{
	const uint32_t c0 = warp * (uint32_t)Cfg<ROWS>::SPAN_CH;
	const uint32_t ltmask = (1u << lane) - 1u;
	const uint32_t *raw32 = (const uint32_t *)(s.raw + 16);
	const uint8_t *rawb = s.raw + 16;
	uint8_t *dlist = s.dl + c0;
	uint32_t ndirty = 0;
#pragma unroll(LEAN ? ROWS : 1)
	for (int i = 0; i < ROWS; i++) {
		const uint32_t c = c0 + i * 32 + lane;
		const uint32_t p0 = c * 16;
		const uint32_t e = s.E[c];
		const uint32_t m = s.M[c];
		uint32_t shlo = (uint32_t)d0;
		uint8_t *basep = out_tile - d0;
		if (!LEAN && ((rbhas >> i) & 1)) {
			const int32_t r = s.rbrel[warp][i];
			shlo = (uint32_t)r;
			basep = out_tile - (int64_t)r;
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
			/* the bytes themselves are written in the block-wide pass (one thread per chunk);
			 * leave it the chunk's shift: EPBs since the last reset of the row below the
			 * chunk (bit 15), or the span's count as for any other row */
			s.E[c] = (uint16_t)(below ? ((e - e_prev) | 0x8000u) : e);
			continue;
		}
		const uint32_t mn = s.M[c + 1];
		const uint32_t b = (shlo + e) & 15u;
		const uint32_t bk = b + (uint32_t)__popc(m); /* the unit starts here if nothing is deleted from here on */
		if ((i == 0 || !LEAN) && ((bwl >> i) & 1) && lane == 0 && b) {
			/* the unit this row starts in began on the other side of a seam: our bytes of it */
			uint32_t pos = p0, cnt = b;
			uint8_t *o = basep + (p0 - e);
			while (cnt) {
				if (!((s.M[pos >> 4] >> (pos & 15)) & 1)) {
					*o++ = rawb[pos];
					cnt--;
				}
				pos++;
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
			dlist[ndirty + (uint32_t)__popc(bal & ltmask)] = (uint8_t)(i * 32 + lane);
		ndirty += (uint32_t)__popc(bal);
	}
	if (lane == 0)
		s.sp_nd[warp] = ndirty;
}

/* shift base of row R (low bits of the shift at E = 0, where tile position 0 goes at E = 0) */
template <int ROWS>
__device__ __forceinline__ void row_base(const Smem<ROWS> &s, uint32_t R, uint8_t *out_tile, uint32_t &shlo,
					 uint8_t *&basep)
{
	const uint32_t w = R / ROWS, i = R % ROWS;
	if ((s.sp_rbhas[w] >> i) & 1) {
		const int32_t r = s.rbrel[w][i];
		shlo = (uint32_t)r;
		basep = out_tile - (int64_t)r;
	} else {
		const uint64_t d0 = s.sp_d0[w];
		shlo = (uint32_t)d0;
		basep = out_tile - d0;
	}
}

/* the byte-exact path for chunk c of the tile: shift base and seam from the span's tables */
template <int ROWS>
__device__ __forceinline__ void dirty_of_tile(const Smem<ROWS> &s, uint32_t c, uint8_t *out_tile)
{
	const uint32_t R = c >> 5;
	uint32_t shlo;
	uint8_t *basep;
	row_base<ROWS>(s, R, out_tile, shlo, basep);
	const bool nextlim =
		R == (uint32_t)Cfg<ROWS>::NROW - 1 || ((s.sp_bw[(R + 1) / ROWS] >> ((R + 1) % ROWS)) & 1);
	const uint32_t limit = nextlim ? (R + 1) * 512u : (uint32_t)Cfg<ROWS>::TILE + 64u;
	dirty_chunk<ROWS>(s, c, shlo, basep, limit);
}

/* chunk c of a byte-wise row (a reset point in the row, or the shard ends in it): every byte to
 * its place, the shift restarting at each reset point; E[c] holds the shift at the chunk start
 * as left by emit_span */
template <int ROWS>
__device__ __noinline__ void bytewise_chunk(const Smem<ROWS> &s, uint32_t c, uint8_t *out_tile, uint32_t nvalid)
{
	const uint32_t *raw32 = (const uint32_t *)(s.raw + 16);
	const uint8_t *rawb = s.raw + 16;
	const uint32_t p0 = c * 16;
	uint32_t shlo;
	uint8_t *basep;
	row_base<ROWS>(s, c >> 5, out_tile, shlo, basep);
	const uint32_t enc = s.E[c];
	const uint32_t m = s.M[c];
	const uint4 v = *(const uint4 *)(raw32 + 4 * c);
	const uint32_t rm = reset_mask(raw32[4 * (int)c - 1], v.x, v.y, v.z, v.w) & valid16(p0, nvalid);
	/* out = basec + p0 + j - cur, cur = EPBs since the last reset */
	uint8_t *basec = (enc & 0x8000u) ? out_tile : basep;
	uint32_t cur = enc & 0x7fffu;
	const uint32_t nv = p0 >= nvalid ? 0u : (nvalid - p0 >= 16u ? 16u : nvalid - p0);
	for (uint32_t j = 0; j < nv; j++) {
		if ((rm >> j) & 1) {
			cur = 0;
			basec = out_tile;
		}
		if ((m >> j) & 1) {
			cur++;
			continue;
		}
		basec[p0 + j - cur] = rawb[p0 + j];
	}
}

/* P0 of a tile into buffer s: ticket, bulk load, L2 prefetch one grid ahead, cleared tables */
template <int ROWS, bool STRIP>
__device__ __forceinline__ void stage_take(Smem<ROWS> &s, const ScanArgs &a)
{
	using C = Cfg<ROWS>;
	const uint32_t tid = threadIdx.x;
	const uint32_t lane = tid & 31, warp = tid >> 5;
	const uint32_t ltmask = (1u << lane) - 1u;
	uint32_t *raw32 = (uint32_t *)(s.raw + 16);
	const uint32_t c0 = warp * (uint32_t)C::SPAN_CH;
	(void)ltmask;
	(void)raw32;
	(void)c0;
	/* ---- P0: ticket, bulk load, clear the span's delete masks ---- */
	if (tid == 0) {
		const uint32_t t0 = atomicAdd(a.ticket, 1u) + 1u;
		s.tile = t0;
		if (t0 < a.num_tiles) {
			const uint64_t off = (uint64_t)t0 * C::TILE;
			if (off + (uint64_t)C::TILE <= a.len)
				bulk_load_issue(s.raw + 16, a.in + off, C::TILE, &s.bar);
			const uint64_t noff = off + (uint64_t)gridDim.x * C::TILE;
			if (noff + (uint64_t)C::TILE <= a.len && !(a.flags & 2u))
				l2_prefetch(a.in + noff, C::TILE);
		}
		s.evbase = 0;
		s.b0 = 0;
		s.first_r = 0xffffffffu;
	}
	if (STRIP) {
		uint4 *m4 = (uint4 *)(s.M + c0);
		for (uint32_t i = lane; i < (uint32_t)C::SPAN_CH / 8; i += 32)
			m4[i] = make_uint4(0, 0, 0, 0);
		if (warp == kW - 1 && lane == 0)
			*(uint4 *)(s.M + C::NCH) = make_uint4(0, 0, 0, 0);
	}
}

/*
 * Stage 1 of a tile (buffer s, tile id in s.tile, bulk load under way): wait for the bytes,
 * classify, per-span prefix and events, tile aggregate published.  Ends with every warp past
 * the aggregate barrier; the look-back and everything that needs it is stage 2.
 */
template <int ROWS, bool STRIP, bool TRACE>
__device__ __forceinline__ void stage_classify(Smem<ROWS> &s, const ScanArgs &a, uint32_t parity)
{
	using C = Cfg<ROWS>;
	const uint32_t tid = threadIdx.x;
	const uint32_t lane = tid & 31, warp = tid >> 5;
	const uint32_t ltmask = (1u << lane) - 1u;
	uint32_t *raw32 = (uint32_t *)(s.raw + 16);
	const uint32_t c0 = warp * (uint32_t)C::SPAN_CH;
	(void)ltmask;
	(void)raw32;
	(void)c0;
	const uint32_t t = s.tile;
	const uint64_t tile_off = (uint64_t)t * C::TILE;
	const uint32_t nvalid = tile_off >= a.len ? 0u
						  : (a.len - tile_off >= (uint64_t)C::TILE ? (uint32_t)C::TILE
											   : (uint32_t)(a.len - tile_off));
	const bool full = nvalid == (uint32_t)C::TILE;
	(void)full;

	if (TRACE && tid == 0)
		trace_mark(a, t, 0);
	if (tid == 0)
		raw32[-1] = tile_off == 0 ? a.halo_left
					  : (tile_off <= a.len ? ldg_u32(a.in + tile_off - 4)
							       : annexb::edge_word(a, tile_off - 4));
	if (full) {
		bulk_load_wait_parity(&s.bar, parity);
		__syncwarp();
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
		__syncthreads();
	}

	if (TRACE && tid == 0)
		trace_mark(a, t, 1);

	/* ---- P1 (per warp): classify.  A chunk matters only if some byte <= 3 follows two zero
	 * bytes (an EPB or the third byte of a boundary event): one SIMD-in-register test per chunk
	 * finds the candidates; they are compacted per warp and only they get the exact masks. ---- */
	uint32_t evrows = 0; /* rows of the span with a candidate for a boundary event */
	{
		const uint32_t k1 = 0x01010101u, kfc = 0xfcfcfcfcu;
		uint16_t *cand = s.E + c0;
		uint32_t ntot = 0;
#pragma unroll 2
		for (int i = 0; i < ROWS; i++) {
			const uint32_t c = c0 + i * 32 + lane;
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
			if (annexb::haszero(AB0 | (w0 & kfe)) | annexb::haszero(AB1 | (w1 & kfe)) |
			    annexb::haszero(AB2 | (w2 & kfe)) | annexb::haszero(AB3 | (w3 & kfe)))
				evrows |= 1u << ((c - c0) >> 5);
			if (STRIP) {
				uint32_t del = msb_to_nib(annexb::zero_bytes_msb(AB0 | (w0 ^ k3))) |
					       msb_to_nib(annexb::zero_bytes_msb(AB1 | (w1 ^ k3))) << 4 |
					       msb_to_nib(annexb::zero_bytes_msb(AB2 | (w2 ^ k3))) << 8 |
					       msb_to_nib(annexb::zero_bytes_msb(AB3 | (w3 ^ k3))) << 12;
				if (!full)
					del &= valid16(c * 16, nvalid);
				if (del)
					s.M[c] = (uint16_t)del;
			}
		}
	}
	if (TRACE && tid == 0)
		trace_mark(a, t, 5);
	/* a start code ending just before the span resets at its first byte */
	const uint8_t *rawb = s.raw + 16;
	const bool start_reset = rawb[(int)(c0 * 16) - 3] == 0 && rawb[(int)(c0 * 16) - 2] == 0 &&
				 rawb[(int)(c0 * 16) - 1] == 1 && c0 * 16 < nvalid;
	__syncwarp();
	evrows = warp_or(evrows) | (start_reset ? 1u : 0u);
	const bool any_ev = evrows != 0;

	/* ---- P2 (per warp): EPBs of the span before every chunk ---- */
	uint32_t etot = 0;
	if (STRIP) {
		uint32_t ex[ROWS];
		uint32_t run = 0;
		const uint16_t *mp = s.M + c0 + lane * ROWS;
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
		uint16_t *ep = s.E + c0 + lane * ROWS;
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
			if (!((walk >> i) & 1)) {
				if (lane == 0)
					s.rbrel[warp][i] = cur_rel;
				if (has)
					rbhas |= 1u << i;
				continue;
			}
			const uint32_t c = c0 + i * 32 + lane;
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
			if (lane == 0)
				s.rbrel[warp][i] = cur_rel;
			if (has)
				rbhas |= 1u << i;
			if (bal) {
				bw |= 1u << i;
				const uint32_t top = 31u - (uint32_t)__clz((int)(rm | 1u));
				const uint32_t e_last =
					STRIP ? (uint32_t)s.E[c] + (uint32_t)__popc(s.M[c] & ((1u << top) - 1u)) : 0u;
				const int tl = 31 - __clz((int)bal);
				cur_rel = -(int32_t)__shfl_sync(FULL_MASK, e_last, tl);
				const uint32_t fr = __shfl_sync(FULL_MASK, p0 + (uint32_t)__ffs((int)rm) - 1, __ffs((int)bal) - 1);
				if (!has)
					first_r = fr;
				has = true;
			}
		}
#pragma unroll
		for (int d = 16; d >= 1; d >>= 1) {
			nev += __shfl_xor_sync(FULL_MASK, nev, d);
			nsc += __shfl_xor_sync(FULL_MASK, nsc, d);
		}
	}
	if (!full) { /* rows the shard ends in (or that lie past its end) go byte by byte */
		for (int i = 0; i < ROWS; i++)
			if ((c0 + (uint32_t)(i + 1) * 32) * 16 > nvalid)
				bw |= 1u << i;
	}
	if (TRACE && tid == 0)
		trace_mark(a, t, 6);
	if (lane == 0) {
		s.sp_etail[warp] = has ? (uint32_t)((int32_t)etot + cur_rel) : etot;
		s.sp_etot[warp] = etot;
		s.sp_nev[warp] = nev;
		s.sp_nsc[warp] = nsc;
		s.sp_has[warp] = has ? 1 : 0;
		s.sp_rbhas[warp] = (uint8_t)rbhas;
		s.sp_bw[warp] = (uint8_t)bw;
		s.sp_evrows[warp] = (uint8_t)evrows;
		if (warp == 0)
			s.first_r = first_r;
	}
	__syncthreads();

	/* ---- P3 (warp 0): tile aggregate, published at once; shift at the tile start by the
	 * short look-back of gen 5 ---- */
	if (TRACE && tid == 0)
		trace_mark(a, t, 2);
	if (warp == 0) {
		uint32_t total = 0, e_tail = 0, nev_t = 0, nsc_t = 0;
		bool has_t = false;
#pragma unroll
		for (int j = 0; j < kW; j++) {
			total += s.sp_etot[j];
			nev_t += s.sp_nev[j];
			nsc_t += s.sp_nsc[j];
			if (s.sp_has[j]) {
				e_tail = s.sp_etail[j];
				has_t = true;
			} else {
				e_tail += s.sp_etail[j];
			}
		}
		uint64_t *dt = a.desc + (uint64_t)t * 4;
		if (lane == 0) {
			st_relaxed_u64(dt, pack_agg(e_tail, has_t));
			if (has_t || t == 0)
				st_relaxed_u64(dt + 1, (uint64_t)e_tail);
			const uint32_t fr = s.first_r;
			const uint32_t r0 = (t == 0 && fr < 3) ? fr : 0u;
			st_relaxed_u64(dt + 3, (uint64_t)total | (uint64_t)r0 << 32);
		}
		/* the slot of the tile's event records: asked for now, needed after the look-back */
		if (lane == 0) {
			s.ag_total = total;
			s.ag_nev = nev_t;
			s.ag_nsc = nsc_t;
			s.ag_has = has_t ? 1u : 0u;
			s.ag_evprev = nev_t ? atomicAdd(a.ev_cursor, nev_t) : 0xffffffffu;
		}
	}
}

/*
 * Stage 2 of a tile: shift at the tile start (look-back by warp 0), event records, rows, the
 * byte-exact pass.  Leaves with the tile's output written; the caller's next barrier frees s.
 */
template <int ROWS, bool STRIP, bool TRACE>
__device__ __forceinline__ void stage_emit(Smem<ROWS> &s, const ScanArgs &a)
{
	using C = Cfg<ROWS>;
	const uint32_t tid = threadIdx.x;
	const uint32_t lane = tid & 31, warp = tid >> 5;
	const uint32_t ltmask = (1u << lane) - 1u;
	uint32_t *raw32 = (uint32_t *)(s.raw + 16);
	const uint32_t c0 = warp * (uint32_t)C::SPAN_CH;
	(void)ltmask;
	(void)raw32;
	(void)c0;
	const uint32_t t = s.tile;
	const uint64_t tile_off = (uint64_t)t * C::TILE;
	const uint32_t nvalid = tile_off >= a.len ? 0u
						  : (a.len - tile_off >= (uint64_t)C::TILE ? (uint32_t)C::TILE
											   : (uint32_t)(a.len - tile_off));
	const bool full = nvalid == (uint32_t)C::TILE;
	(void)full;
	if (warp == 0) {
		__syncwarp();
		const uint32_t total = s.ag_total, nev_t = s.ag_nev, nsc_t = s.ag_nsc;
		const bool has_t = s.ag_has != 0;
		uint64_t *dt = a.desc + (uint64_t)t * 4;
		if (t > 0) {
			if (STRIP) {
				const uint64_t b0 = annexb5::lookback(a, t, lane);
				if (lane == 0) {
					s.b0 = b0;
					if (!has_t)
						st_relaxed_u64(dt + 1, b0 + (uint64_t)total);
				}
			} else if (!has_t && lane == 0) {
				st_relaxed_u64(dt + 1, 0);
			}
		}
		if (lane == 0) {
			const uint32_t evbase = s.ag_evprev + 1u;
			s.evbase = evbase;
			dt[2] = (uint64_t)evbase | (uint64_t)nev_t << 32 | (uint64_t)nsc_t << 48;
		}
	}
	if (TRACE && tid == 0)
		trace_mark(a, t, 3);
	__syncthreads();

	/* ---- P5 (spans with events): event records (before the emit pass reuses E of byte-wise rows) ---- */
	const uint32_t nev = s.sp_nev[warp], evrows = s.sp_evrows[warp], rbhas = s.sp_rbhas[warp];
	if (nev) {
		uint32_t idx0 = s.evbase, epb0 = 0;
		for (uint32_t j = 0; j < warp; j++) {
			idx0 += s.sp_nev[j];
			epb0 += s.sp_etot[j];
		}
		for (int i = 0; i < ROWS; i++) {
			if (!((evrows >> i) & 1))
				continue;
			const uint32_t c = c0 + i * 32 + lane;
			const uint32_t p0 = c * 16;
			const uint4 v = *(const uint4 *)(raw32 + 4 * c);
			const TMasks m = t_masks(raw32[4 * (int)c - 1], v.x, v.y, v.z, v.w);
			const uint32_t ev = m.ev16 & event_valid16(p0, nvalid, t == 0);
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
				const uint32_t eb =
					STRIP ? epb0 + (uint32_t)s.E[c] + (uint32_t)__popc(s.M[c] & ((1u << j) - 1u)) : 0u;
				if ((uint64_t)idx < a.ev_cap)
					a.evbuf[idx] = pack_event(tile_off + p0 + j - 2, eb, (m.sc16 >> j) & 1);
			}
			idx0 += __shfl_sync(FULL_MASK, einc, 31);
		}
	}
	/* ---- P4: emit.  Every warp its span's rows; the chunks that need the byte-exact path are
	 * listed per span and then shared out over the whole block (a span has 28 +- 6 of them:
	 * alone, one warp in four would need a second pass and the rest of the tile would wait) ---- */
	if (STRIP) {
		uint64_t d0 = 0;
		{
			bool found = false;
			for (int j = (int)warp - 1; j >= 0; j--) {
				d0 += s.sp_etail[j];
				if (s.sp_has[j]) {
					found = true;
					break;
				}
			}
			if (!found)
				d0 += s.b0;
		}
		if (lane == 0)
			s.sp_d0[warp] = d0;
		/* seams around and inside the span */
		uint32_t bwl = (uint32_t)s.sp_bw[warp] << 1;
		bwl |= warp == 0 ? 1u : ((uint32_t)s.sp_bw[warp - 1] >> (ROWS - 1)) & 1u;
		bwl |= (warp == kW - 1 ? 1u : (uint32_t)s.sp_bw[warp + 1] & 1u) << (ROWS + 1);
		uint8_t *const out_tile = a.rbsp + tile_off;
		if (rbhas == 0 && (bwl & (((1u << ROWS) - 1u) << 1)) == 0)
			emit_span<ROWS, true>(s, warp, lane, bwl, 0, d0, out_tile, nvalid);
		else
			emit_span<ROWS, false>(s, warp, lane, bwl, rbhas, d0, out_tile, nvalid);
		if (TRACE && tid == 0)
			trace_mark(a, t, 7);
		__syncthreads();
		uint32_t pre[kW + 1];
		pre[0] = 0;
#pragma unroll
		for (int j = 0; j < kW; j++)
			pre[j + 1] = pre[j] + s.sp_nd[j];
		uint64_t BW = 0;
#pragma unroll
		for (int j = 0; j < kW; j++)
			BW |= (uint64_t)s.sp_bw[j] << (j * ROWS);
		const uint32_t nitems = pre[kW] + 32u * (uint32_t)__popcll(BW);
		for (uint32_t g = tid; g < nitems; g += kT) {
			if (g >= pre[kW]) {
				/* chunk (g % 32) of the byte-wise row number (g / 32) */
				const uint32_t k = (g - pre[kW]) >> 5;
				uint64_t x = BW;
				for (uint32_t n = 0; n < k; n++)
					x &= x - 1;
				const uint32_t R = (uint32_t)__ffsll((long long)x) - 1;
				bytewise_chunk<ROWS>(s, R * 32 + ((g - pre[kW]) & 31u), out_tile, nvalid);
				continue;
			}
			uint32_t w = 0;
#pragma unroll
			for (int j = 1; j < kW; j++)
				w += g >= pre[j] ? 1u : 0u;
			uint32_t pw = 0;
#pragma unroll
			for (int j = 1; j < kW; j++)
				pw = g >= pre[j] ? pre[j] : pw;
			const uint32_t c = w * (uint32_t)C::SPAN_CH + s.dl[w * (uint32_t)C::SPAN_CH + (g - pw)];
			dirty_of_tile<ROWS>(s, c, out_tile);
		}
	}
	if (TRACE && tid == 0)
		trace_mark(a, t, 4);
}

template <int ROWS, bool STRIP, int MINB = 4, bool TRACE = false>
__global__ void __launch_bounds__(kT, MINB) scan6_kernel(const ScanArgs a)
{
	__shared__ Smem<ROWS> s;
	/* persistent CTAs: tiles are taken by ticket until none is left (a ticket is taken only when
	 * the CTA is ready for it: a tile held back delays every later tile's look-back).  The tile
	 * one grid ahead, which some CTA will take about a tile's life from now, is pulled into L2. */
	uint32_t parity = 0;
	if (threadIdx.x == 0)
		bulk_bar_init(&s.bar);
	for (;;) {
		stage_take<ROWS, STRIP>(s, a);
		__syncthreads();
		if (s.tile >= a.num_tiles)
			break;
		stage_classify<ROWS, STRIP, TRACE>(s, a, parity);
		if ((uint64_t)s.tile * Cfg<ROWS>::TILE + Cfg<ROWS>::TILE <= a.len)
			parity ^= 1u;
		stage_emit<ROWS, STRIP, TRACE>(s, a);
		if (a.flags & 1u)
			break;
		__syncthreads(); /* the tile's shared memory is reused */
	}
}

/*
 * The same stages, two tiles per CTA in flight: tile B is taken, loaded and classified (its
 * aggregate published) BEFORE tile A's look-back, so that the look-back finds its predecessors'
 * aggregates already there instead of waiting for their timing jitter (20 % of warp time in the
 * single-buffer kernel).  The ticket of a tile is still taken right before its load and
 * classification (no tile is held back).  Two buffers of half the size: ROWS = 4, 16 KiB tiles.
 */
template <int ROWS, bool STRIP, int MINB = 4, bool TRACE = false>
__global__ void __launch_bounds__(kT, MINB) scan6p_kernel(const ScanArgs a)
{
	__shared__ Smem<ROWS> sb[2];
	uint32_t parity[2] = {0, 0};
	if (threadIdx.x == 0) {
		bulk_bar_init(&sb[0].bar);
		bulk_bar_init(&sb[1].bar);
	}
	uint32_t cur = 0;
	bool pending = false;
	for (;;) {
		Smem<ROWS> &s = sb[cur];
		stage_take<ROWS, STRIP>(s, a);
		__syncthreads();
		const bool valid = s.tile < a.num_tiles;
		if (valid) {
			stage_classify<ROWS, STRIP, TRACE>(s, a, parity[cur]);
			if ((uint64_t)s.tile * Cfg<ROWS>::TILE + Cfg<ROWS>::TILE <= a.len)
				parity[cur] ^= 1u;
		}
		if (pending)
			stage_emit<ROWS, STRIP, TRACE>(sb[cur ^ 1], a);
		if (!valid)
			break;
		pending = true;
		cur ^= 1;
		__syncthreads(); /* the older tile's buffer is free for the next ticket */
	}
}

} /* namespace annexb6 */

#endif /* ANNEXB_SCAN6_CUH */
