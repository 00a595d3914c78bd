/*
 * annexb_frame.cuh — K3: writer-side emulation-prevention-byte insertion and
 * start-code framing of RBSP payloads, as the inverse scan-and-scatter of K1/K2.
 *
 * Reference behaviour reproduced bit-exactly (Parrot-Developers/libh264):
 *   - h264_bs_flush            src/h264_bitstream.c:54-81   (03 before a byte <= 3
 *                              when the two previous OUTPUT bytes are 00 00)
 *   - h264_bs_write_bits       src/h264_bitstream.c:211-239 (one byte per flush)
 *   - start code               00 00 00 01 as h264_avcc_to_byte_stream writes it,
 *                              src/h264.c:251-272 (3-byte form also offered)
 *
 * Closed form, r = payload bytes, z(i) = number of
 * consecutive 00 bytes immediately before byte i inside the same payload:
 *   insert 03 before byte i  <=>  r[i] <= 3  and  z(i) >= 2  and  z(i) even
 * The only cross-tile dependency is z at the tile start; a tiny pre-pass records
 * each tile's trailing zero run, so the main kernel knows z exactly and its tile
 * aggregate is a plain byte count (decoupled look-back over sums).
 *
 * Output position of byte i of payload k:  i + #inserts(<= i) + sc_len * (k+1).
 */
#ifndef ANNEXB_FRAME_CUH
#define ANNEXB_FRAME_CUH

#include "annexb_scan.cuh"

namespace frame {

using annexb::kBlock;
using annexb::kWarps;
using annexb::kInvalid;
using annexb::swz;
using annexb::zmask4;

constexpr uint64_t kPrefix = 1ull << 62;
constexpr uint64_t kValueMask = (1ull << 62) - 1;

struct FrameArgs {
	const uint8_t *rbsp;
	uint64_t len;
	const uint64_t *off; /* n+1 entries */
	uint64_t n;
	uint32_t sc_len; /* 0, 3 or 4 */
	uint8_t *out;
	uint64_t out_cap;
	uint64_t *out_off; /* n+1 entries */
	uint64_t *total;
	uint64_t *desc;   /* one word per tile */
	uint32_t *ticket; /* pre-set to 0xffffffff */
	uint64_t *first;  /* num_tiles+1: first payload starting at/after the tile */
	uint32_t *tail;   /* num_tiles: trailing zero run | all-zero << 31 */
	uint32_t num_tiles;
};

/* first k in [0,n) with off[k] >= x */
__device__ __forceinline__ uint64_t lower_bound_off(const uint64_t *off, uint64_t n, uint64_t x)
{
	uint64_t lo = 0, hi = n;
	while (lo < hi) {
		uint64_t mid = (lo + hi) >> 1;
		if (off[mid] < x)
			lo = mid + 1;
		else
			hi = mid;
	}
	return lo;
}

/* number of k in [klo,khi) with off[k] <= x */
__device__ __forceinline__ uint32_t count_le(const uint64_t *off, uint64_t klo, uint64_t khi,
					      uint64_t x)
{
	uint64_t lo = klo, hi = khi;
	while (lo < hi) {
		uint64_t mid = (lo + hi) >> 1;
		if (off[mid] <= x)
			lo = mid + 1;
		else
			hi = mid;
	}
	return (uint32_t)(lo - klo);
}

template <int ITEMS> __global__ void frame_prepass(const FrameArgs a)
{
	constexpr uint64_t TILE = (uint64_t)kBlock * ITEMS * 16;
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t > a.num_tiles)
		return;
	const uint64_t s = t * TILE < a.len ? t * TILE : a.len;
	const uint64_t k0 = lower_bound_off(a.off, a.n, t == a.num_tiles ? a.len : s);
	a.first[t] = k0;
	if (t == a.num_tiles)
		return;
	const uint64_t e = (t + 1) * TILE < a.len ? (t + 1) * TILE : a.len;
	const uint64_t k1 = lower_bound_off(a.off, a.n, e);
	const bool has_b = k1 > k0;
	const uint64_t stop = has_b ? a.off[k1 - 1] : s;
	uint64_t p = e;
	while (p > stop && a.rbsp[p - 1] == 0)
		p--;
	const uint32_t cnt = (uint32_t)(e - p);
	const bool allz = p == s && !has_b && e > s;
	a.tail[t] = cnt | (allz ? 0x80000000u : 0u);
}

/* insert one 03 byte before byte j of the 20-byte little-endian value d[0..4] */
__device__ __forceinline__ void insert03(uint32_t d[5], uint32_t j)
{
	uint32_t s[5];
	s[0] = d[0] << 8;
#pragma unroll
	for (int m = 1; m < 5; m++)
		s[m] = __funnelshift_l(d[m - 1], d[m], 8);
#pragma unroll
	for (int m = 0; m < 5; m++) {
		const uint32_t lo = 4u * m;
		if (j > lo + 3) {
			/* unchanged */
		} else if (j < lo) {
			d[m] = s[m];
		} else {
			const uint32_t r = j - lo; /* 0..3 */
			const uint32_t lm = (1u << (8 * r)) - 1;
			const uint32_t hm = r == 3 ? 0u : ~((1u << (8 * (r + 1))) - 1);
			d[m] = (d[m] & lm) | (3u << (8 * r)) | (s[m] & hm);
		}
	}
}

template <int ITEMS, int MINB = 4>
__global__ void __launch_bounds__(kBlock, MINB) frame_kernel(const FrameArgs a)
{
	constexpr int WARP_BYTES = ITEMS * 512;
	constexpr int TILE = kBlock * ITEMS * 16;
	constexpr uint32_t SMEM_CAP = TILE + TILE / 4;

	__shared__ uint32_t s_tile;
	__shared__ uint32_t s_wtot[kWarps];
	__shared__ uint64_t s_pin;
	__shared__ __align__(16) uint8_t s_out[SMEM_CAP + 64];

	const uint32_t tid = threadIdx.x;
	const uint32_t lane = tid & 31, warp = tid >> 5;

	if (tid == 0)
		s_tile = atomicAdd(a.ticket, 1u) + 1u;
	__syncthreads();
	const uint32_t t = s_tile;
	const uint64_t tile_off = (uint64_t)t * TILE;
	const uint64_t region = tile_off + (uint64_t)warp * WARP_BYTES;
	const bool edge = tile_off + TILE > a.len;
	const uint64_t k_lo = a.first[t], k_hi = a.first[t + 1];
	const bool tile_has_b = k_hi > k_lo;

	/* ---- zero run before the tile (parity-faithful code 0,1,2=even>=2,3=odd>=3) ---- */
	uint32_t zt = 0;
	if (t > 0 && !(tile_has_b && a.off[k_lo] == tile_off)) {
		uint64_t z = 0;
		for (int64_t tt = (int64_t)t - 1;; tt--) {
			const uint32_t ti = a.tail[tt];
			z += ti & 0x7fffffffu;
			if (!(ti >> 31) || tt == 0)
				break;
		}
		zt = z < 2 ? (uint32_t)z : 2u + (uint32_t)(z & 1);
	}

	/* ---- zero run before this warp's region ---- */
	uint32_t zw = 0;
	if (region < a.len) {
		if (region == tile_off) {
			zw = zt;
		} else {
			const uint32_t nb_before = tile_has_b ? count_le(a.off, k_lo, k_hi, region - 1) : 0;
			const bool b_here = tile_has_b &&
					    count_le(a.off, k_lo, k_hi, region) > nb_before;
			if (!b_here) {
				const uint64_t limit = nb_before ? a.off[k_lo + nb_before - 1] : tile_off;
				uint64_t p = region;
				while (p > limit && a.rbsp[p - 1] == 0)
					p--;
				zw = (uint32_t)(region - p);
				if (p == limit && nb_before == 0)
					zw += zt;
			}
		}
	}

	/* ---- load ---- */
	uint4 v[ITEMS];
#pragma unroll
	for (int i = 0; i < ITEMS; i++) {
		const uint64_t o = region + i * 512 + lane * 16;
		if (!edge || o + 16 <= a.len) {
			v[i] = ldg_stream16(a.rbsp + o);
		} else {
			uint32_t w[4] = {0, 0, 0, 0};
			for (int b = 0; b < 16; b++)
				if (o + b < a.len)
					w[b >> 2] |= (uint32_t)a.rbsp[o + b] << (8 * (b & 3));
			v[i] = make_uint4(w[0], w[1], w[2], w[3]);
		}
	}

	/* ---- phase A: insertion masks and byte counts ---- */
	uint32_t ib[ITEMS];   /* ins16 | boundary16 << 16 */
	uint32_t offp[ITEMS]; /* exclusive output offset inside the warp region */
	uint32_t run = 0, cz = zw;
	uint32_t bitems = 0; /* items that contain a payload start */

#pragma unroll
	for (int i = 0; i < ITEMS; i++) {
		const uint64_t item_abs = region + i * 512;
		const uint64_t x = item_abs + lane * 16;
		uint32_t vm = 0xffffu;
		if (edge) {
			uint64_t nv = x >= a.len ? 0 : (a.len - x >= 16 ? 16 : a.len - x);
			vm = (1u << nv) - 1;
		}
		const uint32_t w0 = v[i].x, w1 = v[i].y, w2 = v[i].z, w3 = v[i].w;
		const uint32_t kfc = 0xfcfcfcfcu;
		const uint32_t Z16 = (zmask4(w0) | zmask4(w1) << 4 | zmask4(w2) << 8 | zmask4(w3) << 12) & vm;
		const uint32_t L16 = (zmask4(w0 & kfc) | zmask4(w1 & kfc) << 4 | zmask4(w2 & kfc) << 8 |
				      zmask4(w3 & kfc) << 12) & vm;

		/* payload starts inside this item (warp-uniform test first) */
		uint32_t B16 = 0, nbc = 0;
		bool item_b = false;
		if (tile_has_b) {
			const uint32_t c_lo = item_abs ? count_le(a.off, k_lo, k_hi, item_abs - 1) : 0;
			const uint32_t c_hi = count_le(a.off, k_lo, k_hi, item_abs + 511);
			item_b = c_hi > c_lo;
			if (item_b) {
				const uint32_t c0 = x ? count_le(a.off, k_lo, k_hi, x - 1) : 0;
				const uint32_t c1 = count_le(a.off, k_lo, k_hi, x + 15);
				nbc = c1 - c0;
				for (uint32_t q = c0; q < c1; q++)
					B16 |= 1u << (uint32_t)(a.off[k_lo + q] - x);
				bitems |= 1u << i;
			}
		}

		/* zero run reaching into this lane from the left */
		const uint32_t lastb = B16 ? 31u - (uint32_t)__clz((int)B16) : 0u;
		uint32_t tzl = (uint32_t)__clz((int)~(Z16 << 16));
		if (tzl > 16)
			tzl = 16;
		if (B16 && tzl > 16 - lastb)
			tzl = 16 - lastb;
		const bool allz = Z16 == 0xffffu && B16 == 0;
		const uint32_t az = __ballot_sync(FULL_MASK, allz);
		uint32_t r = 0;
		if (lane) {
			const uint32_t m = az & ((1u << lane) - 1);
			r = (uint32_t)__clz((int)~(m << (32 - lane)));
		}
		const int src = (int)lane - (int)r - 1;
		const uint32_t tsrc = __shfl_sync(FULL_MASK, tzl, src < 0 ? 0 : src);
		uint32_t zin = 16 * r + (src >= 0 ? tsrc : cz);
		if (B16 & 1)
			zin = 0;
		/* carry to the next item: the run ending at lane 31's last byte */
		const uint32_t incl = allz ? 16 + zin : tzl;
		cz = __shfl_sync(FULL_MASK, incl, 31);

		uint32_t ins16 = 0;
		if (B16 == 0) {
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
			uint32_t z = zin;
			const uint32_t wv[4] = {w0, w1, w2, w3};
			for (uint32_t j = 0; j < 16; j++) {
				if (!(vm >> j & 1))
					break;
				if (B16 >> j & 1)
					z = 0;
				const uint32_t c = (wv[j >> 2] >> (8 * (j & 3))) & 0xff;
				if (c <= 3 && z >= 2 && !(z & 1))
					ins16 |= 1u << j;
				z = c == 0 ? z + 1 : 0;
			}
		}
		ib[i] = ins16 | B16 << 16;
		const uint32_t cnt = (uint32_t)__popc(vm) + (uint32_t)__popc(ins16) + a.sc_len * nbc;
		uint32_t inc = cnt;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
			if (lane >= (uint32_t)d)
				inc += o;
		}
		offp[i] = run + inc - cnt;
		run += __shfl_sync(FULL_MASK, inc, 31);
	}
	if (lane == 0)
		s_wtot[warp] = run;
	__syncthreads();

	/* ---- phase B: tile total, look-back over sums, publish ---- */
	uint32_t wbase = 0, tile_total = 0;
#pragma unroll
	for (int j = 0; j < kWarps; j++) {
		if ((uint32_t)j < warp)
			wbase += s_wtot[j];
		tile_total += s_wtot[j];
	}
	if (warp == 0) {
		uint64_t pin = 0;
		if (t > 0) {
			if (lane == 0)
				st_relaxed_u64(a.desc + t, (uint64_t)tile_total);
			int64_t j0 = (int64_t)t - 1;
			for (bool done = false; !done;) {
				const int64_t j = j0 - lane;
				const bool need = j >= 0;
				uint64_t d = kInvalid;
				uint32_t pm;
				for (;;) {
					if (need)
						d = ld_relaxed_u64(a.desc + j);
					pm = __ballot_sync(FULL_MASK, need && !(d & kInvalid) && (d & kPrefix));
					uint32_t ok = __ballot_sync(FULL_MASK, !need || !(d & kInvalid));
					uint32_t upto = pm ? ((pm & (0u - pm)) << 1) - 1u : 0xffffffffu;
					if ((ok & upto) == upto)
						break;
				}
				const int fp = pm ? __ffs((int)pm) - 1 : 32;
				uint64_t val = (need && (int)lane <= fp) ? (d & kValueMask) : 0;
#pragma unroll
				for (int dd = 16; dd >= 1; dd >>= 1)
					val += __shfl_xor_sync(FULL_MASK, val, dd);
				pin += val;
				if (fp < 32)
					done = true;
				else
					j0 -= 32;
			}
		}
		const uint64_t pout = pin + tile_total;
		if (lane == 0) {
#ifdef H264_EMU
			/* test hook: withhold most prefixes so look-back must sum aggregates */
			if (t == 0 || annexb::emu_prefix_every <= 1 || t % annexb::emu_prefix_every == 0)
#endif
				st_relaxed_u64(a.desc + t, pout | kPrefix);
			s_pin = pin;
			if (t == a.num_tiles - 1) {
				/* payloads that start at len (empty, at the very end) and the total */
				uint64_t pos = pout;
				for (uint64_t k = k_hi; k < a.n; k++) {
					a.out_off[k] = pos;
					for (uint32_t b = 0; b < a.sc_len; b++)
						if (pos + b < a.out_cap)
							a.out[pos + b] = b + 1 == a.sc_len ? 1 : 0;
					pos += a.sc_len;
				}
				a.out_off[a.n] = pos;
				*a.total = pos;
			}
		}
	}
	__syncthreads();

	/* ---- phase C: expand into shared (or straight to global), copy out ---- */
	const uint64_t pin = s_pin;
	const uint32_t shift = (uint32_t)(pin & 15);
	const bool staged = shift + tile_total <= SMEM_CAP;
	uint32_t *so32 = (uint32_t *)s_out;

	uint32_t Gb[ITEMS];
#pragma unroll
	for (int i = 0; i < ITEMS; i++) {
		uint32_t vm = 0xffffu;
		if (edge) {
			const uint64_t x = region + i * 512 + lane * 16;
			uint64_t nv = x >= a.len ? 0 : (a.len - x >= 16 ? 16 : a.len - x);
			vm = (1u << nv) - 1;
		}
		const uint32_t ni = (uint32_t)__popc(ib[i] & 0xffffu);
		const uint32_t c = (uint32_t)__popc(vm) + ni;
		Gb[i] = __ballot_sync(FULL_MASK, staged && (ib[i] >> 16) == 0 && ni <= 4 && c >= 4);
	}

	uint32_t carry_tail = 0;
#pragma unroll
	for (int i = 0; i < ITEMS; i++) {
		const uint64_t x = region + i * 512 + lane * 16;
		uint32_t vm = 0xffffu;
		if (edge) {
			uint64_t nv = x >= a.len ? 0 : (a.len - x >= 16 ? 16 : a.len - x);
			vm = (1u << nv) - 1;
		}
		const uint32_t ins16 = ib[i] & 0xffffu, B16 = ib[i] >> 16;
		const uint32_t nvb = (uint32_t)__popc(vm);
		const uint32_t c = nvb + (uint32_t)__popc(ins16);
		const uint32_t o = wbase + offp[i]; /* tile-local output offset */
		const bool myG = (Gb[i] >> lane) & 1;

		uint32_t d[5] = {v[i].x, v[i].y, v[i].z, v[i].w, 0};
		uint32_t tailw = 0, tw = 0, e = 0, wi = 0, s = 0;
		uint32_t xw[6] = {0, 0, 0, 0, 0, 0};
		if (myG) {
			for (uint32_t m = ins16; m;) {
				const uint32_t j = 31u - (uint32_t)__clz((int)m);
				m &= ~(1u << j);
				insert03(d, j);
			}
			const uint32_t os = shift + o;
			s = os & 3;
			wi = os >> 2;
			const uint32_t sh = 8 * s;
			e = s + c;
			xw[0] = d[0] << sh;
			xw[1] = __funnelshift_rc(d[0], d[1], 32 - sh);
			xw[2] = __funnelshift_rc(d[1], d[2], 32 - sh);
			xw[3] = __funnelshift_rc(d[2], d[3], 32 - sh);
			xw[4] = __funnelshift_rc(d[3], d[4], 32 - sh);
			xw[5] = __funnelshift_rc(d[4], 0u, 32 - sh);
			tw = e >> 2;
			tailw = tw == 0 ? xw[0] : tw == 1 ? xw[1] : tw == 2 ? xw[2] : tw == 3 ? xw[3]
				: tw == 4 ? xw[4] : xw[5];
		}
		uint32_t ptail = __shfl_up_sync(FULL_MASK, tailw, 1);
		if (lane == 0)
			ptail = carry_tail;
		carry_tail = __shfl_sync(FULL_MASK, tailw, 31);
		const bool prevG = lane ? ((Gb[i] >> (lane - 1)) & 1)
					: (i > 0 && (Gb[i > 0 ? i - 1 : 0] >> 31));
		const bool nextG = lane < 31 ? ((Gb[i] >> (lane + 1)) & 1)
					     : (i + 1 < ITEMS && (Gb[i + 1 < ITEMS ? i + 1 : i] & 1));
		if (myG) {
			if (s == 0) {
				so32[swz(wi)] = xw[0];
			} else if (prevG) {
				const uint32_t lom = (1u << (8 * s)) - 1;
				so32[swz(wi)] = (xw[0] & ~lom) | (ptail & lom);
			} else {
				for (uint32_t b = s; b < 4; b++)
					s_out[(swz(wi) << 2) | b] = (uint8_t)(xw[0] >> (8 * b));
			}
#pragma unroll
			for (uint32_t m = 1; m < 6; m++)
				if (4 * m + 4 <= e)
					so32[swz(wi + m)] = xw[m];
			if ((e & 3) && !nextG) {
				for (uint32_t b = 0; b < (e & 3); b++)
					s_out[(swz(wi + tw) << 2) | b] = (uint8_t)(tailw >> (8 * b));
			}
		} else if (nvb) {
			/* generic path: payload starts, many inserts, short tails, unstaged tiles */
			uint32_t op = o;
			uint32_t q = 0; /* index of the next payload start in this lane */
			uint32_t c0 = 0;
			if (B16)
				c0 = x ? count_le(a.off, k_lo, k_hi, x - 1) : 0;
			for (uint32_t j = 0; j < 16 && (vm >> j & 1); j++) {
				if (B16 >> j & 1) {
					/* every payload starting at byte x+j (empty ones included) */
					while (k_lo + c0 + q < k_hi && a.off[k_lo + c0 + q] == x + j) {
						a.out_off[k_lo + c0 + q] = pin + op;
						for (uint32_t b = 0; b < a.sc_len; b++) {
							const uint8_t val = b + 1 == a.sc_len ? 1 : 0;
							if (staged) {
								uint32_t ob = shift + op;
								s_out[(swz(ob >> 2) << 2) | (ob & 3)] = val;
							} else if (pin + op < a.out_cap) {
								a.out[pin + op] = val;
							}
							op++;
						}
						q++;
					}
				}
				for (int rep = (ins16 >> j & 1) ? 0 : 1; rep < 2; rep++) {
					const uint8_t val = rep == 0 ? 3 : (uint8_t)(d[j >> 2] >> (8 * (j & 3)));
					if (staged) {
						uint32_t ob = shift + op;
						s_out[(swz(ob >> 2) << 2) | (ob & 3)] = val;
					} else if (pin + op < a.out_cap) {
						a.out[pin + op] = val;
					}
					op++;
				}
			}
		}
	}

	if (staged) {
		__syncthreads();
		const uint32_t total = shift + tile_total;
		const uint32_t nvec = (total + 15) >> 4;
		const uint64_t gb = pin - shift;
		for (uint32_t xv = tid; xv < nvec; xv += kBlock) {
			uint4 raw = *(const uint4 *)(s_out + 16 * xv);
			const uint32_t r = (xv >> 3) & 3;
			uint4 val;
			val.x = r == 0 ? raw.x : r == 1 ? raw.y : r == 2 ? raw.z : raw.w;
			val.y = r == 0 ? raw.y : r == 1 ? raw.x : r == 2 ? raw.w : raw.z;
			val.z = r == 0 ? raw.z : r == 1 ? raw.w : r == 2 ? raw.x : raw.y;
			val.w = r == 0 ? raw.w : r == 1 ? raw.z : r == 2 ? raw.y : raw.x;
			const uint32_t lo = xv * 16;
			if (lo >= shift && lo + 16 <= total && gb + lo + 16 <= a.out_cap) {
				stg_stream16(a.out + gb + lo, val);
			} else {
				const uint32_t b0 = lo >= shift ? lo : shift;
				const uint32_t b1 = lo + 16 <= total ? lo + 16 : total;
				const uint32_t wv[4] = {val.x, val.y, val.z, val.w};
				for (uint32_t b = b0; b < b1; b++)
					if (gb + b < a.out_cap)
						a.out[gb + b] = (uint8_t)(wv[(b >> 2) & 3] >> (8 * (b & 3)));
			}
		}
	}
}

} /* namespace frame */

#endif /* ANNEXB_FRAME_CUH */
