/*
 * annexb_frame6.cuh — K3 second generation: writer-side emulation-prevention-byte insertion and
 * start-code framing with the tile organisation of the gen-6 scan kernel (annexb_scan6.cuh).
 *
 * Why: frame_kernel (annexb_frame.cuh) spends ~20 k warp-instructions per 16 KiB tile
 * (profiles/r01_frame_gen1_kernel_raw.csv): exact zero-run bookkeeping for every chunk, a byte shifter
 * per inserted 03, a swizzled shared staging buffer and a copy-out pass; 0.14 of the HBM peak.
 *
 * Here the source tile (32 KiB) is staged once with a bulk copy and the OUTPUT is produced in
 * aligned 16-byte units straight from it.  A source byte at absolute position x goes to
 *
 *     out(x) = x + INS(x) + sc_len * K(x)
 *
 * INS(x) = 03 bytes inserted before x (look-back prefix over tiles + the span prefix E),
 * K(x) = number of payloads starting at or before x (closed form: a search in off[], no chain).
 * Inside a row without a payload start the output ranges of consecutive 16-byte source chunks
 * partition the output; chunk c, with k inserted bytes, covers 16 + k >= 16 output bytes and
 * therefore always holds the unit boundary b = (-out(16c)) & 15 bytes into its output.  A chunk
 * without an insert whose successor has none before its byte b copies source bytes
 * [16c + b, 16c + b + 16) with one funnel-shifted aligned store (~85 % of the chunks); the others
 * build their one or two units with a 128-bit byte shifter (block-shared pass, one item per
 * thread).  Rows that contain a payload start (start code, zero-run reset, out_off[k]) or the
 * end of the input go byte by byte, one thread per chunk; units next to such a row or to the
 * tile seam are cut at the seam (each side writes its own bytes).
 *
 * Classification: an insert needs `00 00 [<= 3]`; one SIMD-in-register test per chunk finds the
 * candidate chunks (~10 %), only they get the exact parity-of-the-zero-run evaluation
 * (insert before byte i  <=>  r[i] <= 3, z(i) >= 2, z(i) even).
 *
 * Reference behaviour reproduced bit-exactly (Parrot-Developers/libh264):
 *   h264_bs_flush        src/h264_bitstream.c:54-81
 *   h264_bs_write_bits   src/h264_bitstream.c:211-239
 *   start code           src/h264.c:251-272
 * The pre-pass (per-tile first payload and trailing zero run) is frame::frame_prepass.
 */
#ifndef ANNEXB_FRAME6_CUH
#define ANNEXB_FRAME6_CUH

#include "annexb_frame.cuh"

namespace frame6 {

using annexb::kInvalid;
using annexb::zmask4;
using frame::FrameArgs;
using frame::count_le;
using frame::kPrefix;
using frame::kValueMask;

constexpr int kT = 256;
constexpr int kW = kT / 32;
constexpr int kLbW = 1;   /* look-back: windows of 32 tile descriptors fetched per round trip (4 measured: no gain, the wait is for aggregates, not depth) */
constexpr int kSoff = 64; /* payload starts of a tile staged in shared memory (more: searched in off[]) */

template <int ROWS> struct Cfg {
	static constexpr int SPAN_CH = 32 * ROWS; /* 16-byte chunks per warp span */
	static constexpr int NCH = kW * SPAN_CH;  /* chunks per tile */
	static constexpr int TILE = NCH * 16;     /* bytes per tile */
	static constexpr int NROW = kW * ROWS;    /* 512-byte rows per tile (<= 64) */
};

template <int ROWS> struct __align__(128) Smem {
	uint8_t raw[16 + Cfg<ROWS>::TILE + 32]; /* [left halo pad][tile][pad] */
	uint16_t M[Cfg<ROWS>::NCH + 8];         /* insert mask per chunk (bit j = 03 before byte j) */
	uint16_t E[Cfg<ROWS>::NCH];             /* inserts of the span before the chunk; candidate list before that */
	uint8_t dl[Cfg<ROWS>::NCH];             /* per span: chunks that take the byte-exact path */
	uint32_t krow[Cfg<ROWS>::NROW + 1];     /* payload starts of the tile before the row start */
	uint32_t soff[kSoff];                   /* the tile's first payload starts, relative to the tile */
	uint32_t sp_etot[kW], sp_nd[kW];
	uint64_t sp_base[kW];                   /* inserts before the span + sc_len * payloads before the tile */
	uint64_t bar;
	uint64_t pin; /* inserts before the tile */
	uint32_t tile, zt;
};

__device__ __forceinline__ uint32_t valid16(uint32_t p0, uint32_t nvalid)
{
	const uint32_t nv = p0 >= nvalid ? 0u : (nvalid - p0 >= 16u ? 16u : nvalid - p0);
	return (1u << nv) - 1u;
}

/* payload starts of the tile (off[k_lo .. k_hi), nb of them) at or before tile position x; from
 * the staged copy when they all fit */
template <int ROWS>
__device__ __forceinline__ uint32_t starts_le(const Smem<ROWS> &s, const FrameArgs &a, uint64_t tile_off, uint64_t k_lo,
					      uint64_t k_hi, uint32_t x)
{
	const uint32_t nb = (uint32_t)(k_hi - k_lo);
	if (k_hi - k_lo > (uint64_t)kSoff)
		return count_le(a.off, k_lo, k_hi, tile_off + x);
	uint32_t lo = 0, hi = nb;
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if (s.soff[mid] <= x)
			lo = mid + 1;
		else
			hi = mid;
	}
	return lo;
}

/* tile position of the tile's payload start number i */
template <int ROWS>
__device__ __forceinline__ uint32_t start_at(const Smem<ROWS> &s, const FrameArgs &a, uint64_t tile_off, uint64_t k_lo,
					     uint64_t k_hi, uint32_t i)
{
	return k_hi - k_lo > (uint64_t)kSoff ? (uint32_t)(a.off[k_lo + i] - tile_off) : s.soff[i];
}

/* one byte / one 16-byte unit to the output, never past its capacity */
__device__ __forceinline__ void put_byte(uint8_t *p, uint8_t v, const uint8_t *capend)
{
	if (p < capend)
		*p = v;
}

/* a unit that does not fit whole: byte by byte (only next to the end of the output buffer) */
__device__ __noinline__ void put_unit_bytes(uint8_t *p, uint64_t q0, uint64_t q1, const uint8_t *capend)
{
#pragma unroll 1
	for (uint32_t n = 0; n < 16; n++)
		put_byte(p + n, (uint8_t)((n < 8 ? q0 >> (8 * n) : q1 >> (8 * (n - 8))) & 0xff), capend);
}

__device__ __forceinline__ void put_unit(uint8_t *p, uint64_t q0, uint64_t q1, const uint8_t *capend)
{
	if (p + 16 <= capend)
		stg_stream16(p, make_uint4((uint32_t)q0, (uint32_t)(q0 >> 32), (uint32_t)q1, (uint32_t)(q1 >> 32)));
	else
		put_unit_bytes(p, q0, q1, capend);
}

/* bytes [from, to) of chunk c's own output sequence (its 03s included), byte stores; dst = where
 * output offset 0 of the chunk goes */
template <int ROWS>
__device__ __noinline__ void chunk_bytes(const Smem<ROWS> &s, uint32_t c, uint32_t from, uint32_t to, uint8_t *dst,
					    const uint8_t *capend)
{
	const uint8_t *rawb = s.raw + 16 + c * 16;
	const uint32_t m = s.M[c];
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
__device__ __forceinline__ void gen_unit(const Smem<ROWS> &s, uint32_t c, uint32_t ub, uint8_t *dst, const uint8_t *capend)
{
	const uint32_t m = s.M[c];
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
	uint32_t dm = (m | (uint32_t)s.M[c + 1] << 16) >> j;
	if (idx == ub)
		dm &= ~1u; /* the 03 before byte j, if any, belongs to the unit before */
	dm &= 0xffffu;
	const uint32_t S = c * 16 + j;
	const uint32_t *raw32 = (const uint32_t *)(s.raw + 16);
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

/* chunk c of a byte-wise row: payload starts (out_off, start code), 03s and bytes one by one */
template <int ROWS>
__device__ __noinline__ void bytewise_chunk(const Smem<ROWS> &s, const FrameArgs &a, uint32_t c, uint64_t tile_off,
					       uint32_t nvalid, uint64_t k_lo, uint64_t k_hi)
{
	const uint32_t p0 = c * 16;
	if (p0 >= nvalid)
		return;
	const uint32_t nv = nvalid - p0 >= 16u ? 16u : nvalid - p0;
	const uint32_t nb = (uint32_t)(k_hi - k_lo);
	/* payloads of the tile that started before the chunk */
	uint32_t next = (nb && p0) ? starts_le<ROWS>(s, a, tile_off, k_lo, k_hi, p0 - 1) : 0u;
	/* sp_base holds sc_len * k_lo already */
	const uint64_t d = s.sp_base[c / (uint32_t)Cfg<ROWS>::SPAN_CH] + (uint64_t)s.E[c];
	uint64_t pos = tile_off + p0 + d + a.sc_len * next;
	const uint8_t *capend = a.out + a.out_cap;
	const uint8_t *rawb = s.raw + 16 + p0;
	const uint32_t m = s.M[c];
	uint32_t noff = next < nb ? start_at<ROWS>(s, a, tile_off, k_lo, k_hi, next) : 0xffffffffu;
	for (uint32_t j = 0; j < nv; j++) {
		while (noff == p0 + j) {
			a.out_off[k_lo + next] = pos;
			for (uint32_t b = 0; b < a.sc_len; b++)
				put_byte(a.out + pos + b, b + 1 == a.sc_len ? 1 : 0, capend);
			pos += a.sc_len;
			next++;
			noff = next < nb ? start_at<ROWS>(s, a, tile_off, k_lo, k_hi, next) : 0xffffffffu;
		}
		if ((m >> j) & 1)
			put_byte(a.out + pos++, 3, capend);
		put_byte(a.out + pos++, rawb[j], capend);
	}
}

/* the tile the input ends in (once per launch): plain loads, zero fill past the end */
template <int ROWS>
__device__ __noinline__ void load_partial_tile(Smem<ROWS> &s, const FrameArgs &a, uint64_t tile_off)
{
	uint32_t *raw32 = (uint32_t *)(s.raw + 16);
	for (uint32_t c = threadIdx.x; c < (uint32_t)Cfg<ROWS>::NCH; c += kT) {
		const uint64_t o = tile_off + (uint64_t)c * 16;
		uint4 v;
		if (o + 16 <= a.len) {
			v = ldg_stream16(a.rbsp + o);
		} else {
			uint32_t w[4] = {0, 0, 0, 0};
			for (int b = 0; b < 16; b++)
				if (o + b < a.len)
					w[b >> 2] |= (uint32_t)a.rbsp[o + b] << (8 * (b & 3));
			v = make_uint4(w[0], w[1], w[2], w[3]);
		}
		*(uint4 *)(raw32 + 4 * c) = v;
	}
}

/* insert mask of a chunk with payload starts inside (bits of B16): byte by byte, the zero run
 * restarting at every start */
__device__ __noinline__ uint32_t inserts_with_starts(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t B16,
						     uint32_t zin)
{
	uint32_t ins16 = 0, z = zin;
	const uint32_t wv[4] = {w0, w1, w2, w3};
	for (uint32_t j = 0; j < 16; j++) {
		if ((B16 >> j) & 1)
			z = 0;
		const uint32_t cb = (wv[j >> 2] >> (8 * (j & 3))) & 0xff;
		if (cb <= 3 && z >= 2 && !(z & 1))
			ins16 |= 1u << j;
		z = cb == 0 ? z + 1 : 0;
	}
	return ins16;
}

/* last tile: payloads that start at len (empty, at the very end) and the total */
__device__ __noinline__ void finish_output(const FrameArgs &a, uint64_t k_hi, uint64_t pout)
{
	const uint8_t *capend = a.out + a.out_cap;
	for (uint64_t k = k_hi; k < a.n; k++) {
		const uint64_t pos = a.len + pout + a.sc_len * k;
		a.out_off[k] = pos;
		for (uint32_t b = 0; b < a.sc_len; b++)
			put_byte(a.out + pos + b, b + 1 == a.sc_len ? 1 : 0, capend);
	}
	const uint64_t end = a.len + pout + a.sc_len * a.n;
	a.out_off[a.n] = end;
	*a.total = end;
}

/*
 * P4: the rows of a span.  bwl: bit 0 = the row before the span is a seam (byte-wise row or tile
 * start), bits 1..ROWS = own rows that go byte by byte, bit ROWS + 1 = the row after the span is a
 * seam.  LEAN (most spans): no byte-wise row, hence no payload start, in the span: one start-code
 * offset for all rows and seams only at the two ends.  base = where source position 0 of the
 * tile goes for this span at E = 0 without the start codes of the tile, shl = its low bits.
 */
template <int ROWS, bool LEAN>
__device__ __forceinline__ void emit_span(Smem<ROWS> &s, const FrameArgs &a, uint32_t warp, uint32_t lane, uint32_t bwl,
					  uint8_t *base, uint32_t shl, bool has_b, bool safe, const uint8_t *capend)
{
	const uint32_t c0 = warp * (uint32_t)Cfg<ROWS>::SPAN_CH;
	const uint32_t ltmask = (1u << lane) - 1u;
	const uint32_t *raw32 = (const uint32_t *)(s.raw + 16);
	uint8_t *dl = s.dl + c0;
	uint32_t ndirty = 0;
	const uint32_t koff0 = has_b ? a.sc_len * s.krow[warp * ROWS] : 0u;
#pragma unroll(LEAN ? ROWS : 1)
	for (int i = 0; i < ROWS; i++) {
		if (!LEAN && ((bwl >> (i + 1)) & 1))
			continue;
		const uint32_t koff = LEAN ? koff0 : (has_b ? a.sc_len * s.krow[warp * ROWS + i] : 0u);
		const uint32_t c = c0 + i * 32 + lane;
		const uint32_t p0 = c * 16;
		const uint32_t e = s.E[c];
		const uint32_t m = s.M[c];
		const uint32_t mn = s.M[c + 1];
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
	if (lane == 0)
		s.sp_nd[warp] = ndirty;
}

template <int ROWS, int MINB = 5>
__global__ void __launch_bounds__(kT, MINB) frame6_kernel(const FrameArgs a)
{
	using C = Cfg<ROWS>;
	__shared__ Smem<ROWS> s;
	const uint32_t tid = threadIdx.x;
	const uint32_t lane = tid & 31, warp = tid >> 5;
	const uint32_t ltmask = (1u << lane) - 1u;
	uint32_t *raw32 = (uint32_t *)(s.raw + 16);
	const uint8_t *rawb = s.raw + 16;
	const uint32_t c0 = warp * (uint32_t)C::SPAN_CH;
	const uint8_t *capend = a.out + a.out_cap;
	uint32_t parity = 0;
	if (tid == 0)
		bulk_bar_init(&s.bar);

	for (;;) {
		/* ---- P0: ticket, bulk load, L2 prefetch one grid ahead, cleared masks ---- */
		if (tid == 0) {
			const uint32_t t0 = atomicAdd(a.ticket, 1u) + 1u;
			s.tile = t0;
			if (t0 < a.num_tiles) {
				const uint64_t off = (uint64_t)t0 * C::TILE;
				if (off + (uint64_t)C::TILE <= a.len)
					bulk_load_issue(s.raw + 16, a.rbsp + off, C::TILE, &s.bar);
				const uint64_t noff = off + (uint64_t)gridDim.x * C::TILE;
				if (noff + (uint64_t)C::TILE <= a.len)
					l2_prefetch(a.rbsp + noff, C::TILE);
			}
		}
		{
			uint4 *m4 = (uint4 *)(s.M + c0);
			for (uint32_t i = lane; i < (uint32_t)C::SPAN_CH / 8; i += 32)
				m4[i] = make_uint4(0, 0, 0, 0);
			if (warp == kW - 1 && lane == 0)
				*(uint4 *)(s.M + C::NCH) = make_uint4(0, 0, 0, 0);
		}
		__syncthreads();
		const uint32_t t = s.tile;
		if (t >= a.num_tiles)
			break;
		const uint64_t tile_off = (uint64_t)t * C::TILE;
		const uint32_t nvalid = tile_off >= a.len ? 0u
							  : (a.len - tile_off >= (uint64_t)C::TILE ? (uint32_t)C::TILE
												   : (uint32_t)(a.len - tile_off));
		const bool full = nvalid == (uint32_t)C::TILE;
		const uint64_t k_lo = a.first[t], k_hi = a.first[t + 1];
		const bool has_b = k_hi > k_lo;

		/* ---- P0b: the zero run before the tile, payload counts and byte-wise rows ---- */
		if (tid == kT - 1) {
			/* parity-faithful code: 0, 1, 2 = even >= 2, 3 = odd >= 3 */
			uint32_t zt = 0;
			if (t > 0 && !(has_b && a.off[k_lo] == tile_off)) {
				uint64_t z = 0;
				for (int64_t tt = (int64_t)t - 1;; tt--) {
					const uint32_t ti = a.tail[tt];
					z += ti & 0x7fffffffu;
					if (!(ti >> 31) || tt == 0)
						break;
				}
				zt = z < 2 ? (uint32_t)z : 2u + (uint32_t)(z & 1);
			}
			s.zt = zt;
		}
		/* the four bytes before the tile, for the candidate test of its first chunk */
		uint32_t halo = 0xffffffffu;
		if (tid == 0 && tile_off)
			halo = ldg_u32(a.rbsp + tile_off - 4);
		if (tid >= kT - 32 - kSoff && tid < kT - 32 && k_lo + (tid - (kT - 32 - kSoff)) < k_hi)
			s.soff[tid - (kT - 32 - kSoff)] = (uint32_t)(a.off[k_lo + (tid - (kT - 32 - kSoff))] - tile_off);
		if (tid <= (uint32_t)C::NROW) {
			/* payload starts of the tile before row `tid` (entry NROW: all of them) */
			const uint64_t rs = tile_off + (uint64_t)tid * 512;
			s.krow[tid] = (has_b && rs) ? count_le(a.off, k_lo, k_hi, rs - 1) : 0u;
		}
		if (!full)
			load_partial_tile<ROWS>(s, a, tile_off);
		if (full) {
			bulk_load_wait_parity(&s.bar, parity);
			parity ^= 1u;
			__syncwarp();
		} else {
			__syncthreads();
		}
		uint64_t BW;

		/* ---- P1 (per warp): candidates (some byte <= 3 after two zero bytes), then the exact
		 * insert mask for them ---- */
		{
			const uint32_t k1 = 0x01010101u, kfc = 0xfcfcfcfcu;
			uint16_t *cand = s.E + c0;
			uint32_t ntot = 0;
#pragma unroll 2
			for (int i = 0; i < ROWS; i++) {
				const uint32_t c = c0 + i * 32 + lane;
				const uint4 v = *(const uint4 *)(raw32 + 4 * c);
				const uint32_t pw = c ? raw32[4 * (int)c - 1] : halo;
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
			/* the searches of P0b (payload counts, staged offsets, zero run before the tile) ran
			 * under the bulk load and the candidate test: only the exact evaluation needs them */
			__syncthreads();
			{
				/* byte-wise rows: a payload start inside, or the input ends in (or before) the row */
				const uint32_t r0 = lane, r1 = lane + 32;
				const bool b0 = r0 < (uint32_t)C::NROW && (s.krow[r0 + 1] > s.krow[r0] || (r0 + 1) * 512u > nvalid);
				const bool b1 = r1 < (uint32_t)C::NROW && (s.krow[r1 + 1] > s.krow[r1] || (r1 + 1) * 512u > nvalid);
				BW = (uint64_t)__ballot_sync(FULL_MASK, b0) | (uint64_t)__ballot_sync(FULL_MASK, b1) << 32;
			}
			const uint32_t zt = s.zt;
			for (uint32_t q = lane; q < ntot; q += 32) {
				const uint32_t c = cand[q];
				const uint32_t p0 = c * 16;
				const uint32_t R = c >> 5;
				const uint32_t vm = full ? 0xffffu : valid16(p0, nvalid);
				/* payload starts: the last one at or before the chunk, those inside it */
				uint32_t B16 = 0;
				int32_t lim = -1;
				if (has_b) {
					uint32_t cnt0 = s.krow[R];
					if ((BW >> R) & 1) {
						cnt0 = starts_le<ROWS>(s, a, tile_off, k_lo, k_hi, p0);
						const uint32_t cnt1 = starts_le<ROWS>(s, a, tile_off, k_lo, k_hi, p0 + 15);
						for (uint32_t n = cnt0; n < cnt1; n++)
							B16 |= 1u << (start_at<ROWS>(s, a, tile_off, k_lo, k_hi, n) - p0);
					}
					if (cnt0)
						lim = (int32_t)start_at<ROWS>(s, a, tile_off, k_lo, k_hi, cnt0 - 1);
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
					const uint32_t kfc2 = 0xfcfcfcfcu;
					const uint32_t Z16 = zmask4(w0) | zmask4(w1) << 4 | zmask4(w2) << 8 | zmask4(w3) << 12;
					const uint32_t L16 = zmask4(w0 & kfc2) | zmask4(w1 & kfc2) << 4 | zmask4(w2 & kfc2) << 8 |
							     zmask4(w3 & kfc2) << 12;
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
					s.M[c] = (uint16_t)ins16;
			}
			__syncwarp();
		}

		/* ---- P2 (per warp): inserts of the span before every chunk ---- */
		{
			uint32_t ex[ROWS];
			uint32_t run = 0;
			const uint16_t *mp = s.M + c0 + lane * ROWS;
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
			const uint32_t etot = __shfl_sync(FULL_MASK, inc, 31);
			const uint32_t base = inc - run;
			uint16_t *ep = s.E + c0 + lane * ROWS;
#pragma unroll
			for (int k = 0; k < ROWS; k++)
				ep[k] = (uint16_t)(base + ex[k]);
			if (lane == 0)
				s.sp_etot[warp] = etot;
		}
		__syncthreads();

		/* ---- P3 (warp 0): tile aggregate published at once, look-back over sums ---- */
		if (warp == 0) {
			uint32_t tile_total = 0;
#pragma unroll
			for (int j = 0; j < kW; j++)
				tile_total += s.sp_etot[j];
			uint64_t pin = 0;
			if (t > 0) {
				if (lane == 0)
					st_relaxed_u64(a.desc + t, (uint64_t)tile_total);
				/* the inserts of every tile before this one: aggregates summed back to the nearest
				 * tile with an inclusive prefix, kLbW windows of 32 descriptors per round trip
				 * (tile -1 reads as a prefix of 0 and ends the walk) */
				int64_t j0 = (int64_t)t - 1;
				for (bool done = false; !done;) {
					uint64_t d[kLbW];
#pragma unroll
					for (int k = 0; k < kLbW; k++) {
						const int64_t j = j0 - 32 * k - (int64_t)lane;
						d[k] = j >= 0 ? ld_relaxed_u64(a.desc + j) : kPrefix;
					}
#pragma unroll
					for (int k = 0; k < kLbW; k++) {
						const bool valid = !(d[k] & kInvalid);
						const uint32_t pm = __ballot_sync(FULL_MASK, valid && (d[k] & kPrefix));
						const uint32_t ok = __ballot_sync(FULL_MASK, valid);
						const uint32_t upto = pm ? ((pm & (0u - pm)) << 1) - 1u : 0xffffffffu;
						if ((ok & upto) != upto) {
							spin_pause(20); /* an aggregate is not there yet: fetch again from this window on */
							break;
						}
						const int fp = pm ? __ffs((int)pm) - 1 : 32;
						uint64_t val = (int)lane <= fp ? (d[k] & kValueMask) : 0;
#pragma unroll
						for (int dd = 16; dd >= 1; dd >>= 1)
							val += __shfl_xor_sync(FULL_MASK, val, dd);
						pin += val;
						if (fp < 32) {
							done = true;
							break;
						}
						j0 -= 32;
					}
				}
			}
			const uint64_t pout = pin + tile_total;
			if (lane == 0) {
#ifdef H264_EMU
				/* test hook: withhold most prefixes so look-back must sum aggregates */
				if (t == 0 || annexb::emu_prefix_every <= 1 || t % annexb::emu_prefix_every == 0)
#endif
					st_relaxed_u64(a.desc + t, pout | kPrefix);
				s.pin = pin;
				if (t == a.num_tiles - 1)
					finish_output(a, k_hi, pout);
			}
		}
		__syncthreads();

		/* ---- P4: emit.  Every warp its span's rows; chunks that need the byte-exact path are
		 * listed per span and shared out over the block afterwards ---- */
		const uint64_t pin = s.pin;
		/* the tile's whole output, with a unit of slack, inside the capacity: no store needs a check */
		const bool safe = tile_off + (uint64_t)(C::TILE + C::TILE / 2 + 32) + pin + a.sc_len * k_hi <= a.out_cap;
		{
			uint64_t d0 = pin + a.sc_len * k_lo;
			for (uint32_t j = 0; j < warp; j++)
				d0 += s.sp_etot[j];
			if (lane == 0)
				s.sp_base[warp] = d0;
			/* seams: bit 0 = the row before the span, bits 1..ROWS = own byte-wise rows, bit ROWS + 1 =
			 * the row after the span */
			uint32_t bwl = ((uint32_t)(BW >> (warp * ROWS)) & ((1u << ROWS) - 1u)) << 1;
			bwl |= warp == 0 ? 1u : (uint32_t)(BW >> (warp * ROWS - 1)) & 1u;
			bwl |= (warp == kW - 1 ? 1u : (uint32_t)(BW >> ((warp + 1) * ROWS)) & 1u) << (ROWS + 1);
			uint8_t *base = a.out + (tile_off + d0);
			if ((bwl & (((1u << ROWS) - 1u) << 1)) == 0)
				emit_span<ROWS, true>(s, a, warp, lane, bwl, base, (uint32_t)d0, has_b, safe, capend);
			else
				emit_span<ROWS, false>(s, a, warp, lane, bwl, base, (uint32_t)d0, has_b, safe, capend);
		}
		__syncthreads();
		{
			/* byte-wise rows first and on the last warps: they are the long items (searches in
			 * off[], byte stores) and must not share a warp's pass with the listed chunks */
			const uint32_t nbwi = 32u * (uint32_t)__popcll(BW);
			for (uint32_t g = kT - 1 - tid; g < nbwi; g += kT) {
				uint64_t x = BW;
				for (uint32_t n = 0; n < (g >> 5); n++)
					x &= x - 1;
				const uint32_t R = (uint32_t)__ffsll((long long)x) - 1;
				bytewise_chunk<ROWS>(s, a, R * 32 + (g & 31u), tile_off, nvalid, k_lo, k_hi);
			}
			uint32_t pre[kW + 1];
			pre[0] = 0;
#pragma unroll
			for (int j = 0; j < kW; j++)
				pre[j + 1] = pre[j] + s.sp_nd[j];
			const uint8_t *cap2 = safe ? (const uint8_t *)~(uintptr_t)0 : capend;
			for (uint32_t g = tid; g < pre[kW]; g += kT) {
				uint32_t w = 0, pw = 0;
#pragma unroll
				for (int j = 1; j < kW; j++) {
					w += g >= pre[j] ? 1u : 0u;
					pw = g >= pre[j] ? pre[j] : pw;
				}
				const uint32_t c = w * (uint32_t)C::SPAN_CH + s.dl[w * (uint32_t)C::SPAN_CH + (g - pw)];
				const uint32_t R = c >> 5;
				const uint64_t rd = s.sp_base[w] + (has_b ? a.sc_len * s.krow[R] : 0u);
				const uint32_t e = s.E[c];
				const uint32_t kins = (uint32_t)__popc(s.M[c]);
				uint8_t *o = a.out + (tile_off + rd + (uint64_t)c * 16 + e);
				const uint32_t b = (0u - ((uint32_t)rd + e)) & 15u;
				const bool seam_after = (c & 31u) == 31u && (R == (uint32_t)C::NROW - 1 || ((BW >> (R + 1)) & 1));
				if ((c & 31u) == 0 && b && (R == 0 || ((BW >> (R - 1)) & 1)))
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
		}
		__syncthreads(); /* the tile's shared memory is reused */
	}
}

} /* namespace frame6 */

#endif /* ANNEXB_FRAME6_CUH */
