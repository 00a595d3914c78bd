/*
 * annexb_frame8.cuh — K3 fourth generation (opt-in: H264GPU_FRAME_GEN=8): the block-wide 32 KiB
 * tiles of gen 6 (annexb_frame6.cuh: its phases, its emit, its byte-exact pass, unchanged) with the
 * chain of gen 7 (annexb_frame7.cuh): a tile's insert count is PUBLISHED ONE ITERATION AHEAD of
 * its look-back, over the three-level chain (tile words, packed atomic sums over groups of 32
 * tiles, supergroup sums and prefixes).
 *
 * Why: gen 7 showed that the look-back wait can be removed completely (no span waits for a
 * predecessor) and that warp-autonomous spans then lose what they won to dependent-issue latency
 * (two 4 KiB buffers per warp: 20 warps per SM) or to instruction fetch (32 unrelated warps).
 * Gen 6 has neither problem (40 warps per SM, the 8 warps of a CTA run the same phase) but spends
 * 21 % of its warp time at the barrier behind the look-back.  Here a CTA, per loop iteration,
 *   takes the ticket of tile B | stages B | classifies B | publishes B's count |
 *   stages tile A AGAIN (classified an iteration ago; its bytes come from L2) under A's look-back |
 *   emits A.
 * Only the insert masks and payload-start tables of A are kept across the iteration (4.6 KB,
 * swapped with B's after B's classification); the 32 KiB of bytes are staged twice, which costs a
 * bulk copy, not instructions.  The ticket is taken after the emit: a CTA never holds an
 * unpublished tile across a wait (annexb_frame7.cuh).
 *
 * Reference behaviour reproduced bit-exactly (Parrot-Developers/libh264):
 *   h264_bs_flush        src/h264_bitstream.c:54-81
 *   h264_bs_write_bits   src/h264_bitstream.c:211-239
 *   start code           src/h264.c:251-272
 */
#ifndef ANNEXB_FRAME8_CUH
#define ANNEXB_FRAME8_CUH

#include "annexb_frame7.cuh"

namespace frame8 {

using namespace frame6; /* Cfg, Smem, kT, kW, kSoff and every per-chunk routine of gen 6 */

/* what classification leaves for the emission of a tile, parked for an iteration */
template <int ROWS> struct __align__(16) Park {
	uint16_t M[Cfg<ROWS>::NCH + 8];
	uint32_t krow[Cfg<ROWS>::NROW + 1];
	uint32_t soff[kSoff];
};

template <int ROWS> struct __align__(128) Smem8 {
	Smem<ROWS> s;
	Park<ROWS> park;
};

/* first payload and trailing zero run of every tile (as frame::frame_prepass), cleared chain */
template <int ROWS> __global__ void frame8_prepass(const FrameArgs a)
{
	constexpr uint64_t TILE = (uint64_t)Cfg<ROWS>::TILE;
	const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t > a.num_tiles)
		return;
	if ((t & 31) == 0)
		a.group_w[t >> 5] = 0;
	if ((t & 1023) == 0) {
		a.super_w[t >> 10] = 0;
		a.super_p[t >> 10] = t ? 0 : 1;
	}
	if (t == 0)
		*a.ticket = 0;
	const uint64_t s = t * TILE < a.len ? t * TILE : a.len;
	const uint64_t k0 = frame::lower_bound_off(a.off, a.n, t == a.num_tiles ? a.len : s);
	a.first[t] = k0;
	if (t == a.num_tiles)
		return;
	a.desc[t] = 0;
	const uint64_t e = (t + 1) * TILE < a.len ? (t + 1) * TILE : a.len;
	const uint64_t k1 = frame::lower_bound_off(a.off, a.n, e);
	const bool has_b = k1 > k0;
	const uint64_t stop = has_b ? a.off[k1 - 1] : s;
	uint64_t p = e;
	while (p > stop && a.rbsp[p - 1] == 0)
		p--;
	const uint32_t cnt = (uint32_t)(e - p);
	const bool allz = p == s && !has_b && e > s;
	a.tail[t] = cnt | (allz ? 0x80000000u : 0u);
}

template <int ROWS, int MINB>
__global__ void __launch_bounds__(kT, MINB) frame8_kernel(const FrameArgs a)
{
	using C = Cfg<ROWS>;
	__shared__ Smem8<ROWS> sm;
	Smem<ROWS> &s = sm.s;
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
	/* tile A: classified and published an iteration ago, its masks in the park */
	uint32_t tA = 0xffffffffu;
	uint64_t klo_A = 0, khi_A = 0, BW_A = 0;

	for (;;) {
		/* ---- ticket of tile B (after the emit of the iteration before: never held across a wait),
		 * bulk load, L2 prefetch one grid ahead, cleared masks ---- */
		if (tid == 0) {
			const uint32_t t0 = atomicAdd(a.ticket, 1u);
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
		const uint32_t tB = s.tile;
		const bool hasA = tA != 0xffffffffu, hasB = tB < a.num_tiles;
		if (!hasA && !hasB)
			break;
		uint64_t klo_B = 0, khi_B = 0, BW_B = 0;

		/* ==== tile B: classification (the phases of gen 6) and its insert count ==== */
		if (hasB) {
			const uint32_t t = tB;
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

			/* inserts of the warp's span (the tile's count is all the chain needs now) */
			{
				uint32_t run = 0;
				const uint16_t *mp = s.M + c0 + lane * ROWS;
#pragma unroll
				for (int k = 0; k < ROWS; k++)
					run += (uint32_t)__popc(mp[k]);
				run = warp_add(run);
				if (lane == 0)
					s.sp_etot[warp] = run;
			}
			klo_B = k_lo;
			khi_B = k_hi;
			BW_B = BW;
		}
		__syncthreads();

		/* ==== B's count published; A's bytes staged again under A's look-back; masks swapped ==== */
		if (warp == 0) {
			if (hasB && lane == 0) {
				uint32_t tile_total = 0;
#pragma unroll
				for (int j = 0; j < kW; j++)
					tile_total += s.sp_etot[j];
				frame7::publish(a, tB, tile_total);
			}
			if (hasA) {
				const uint64_t off = (uint64_t)tA * C::TILE;
				if (lane == 0 && off + (uint64_t)C::TILE <= a.len)
					bulk_load_issue(s.raw + 16, a.rbsp + off, C::TILE, &s.bar);
				uint32_t spins = 0;
				const frame7::Probe pr = frame7::lb_fetch(a, tA, lane);
				const uint64_t pin = frame7::lb_eval(a, tA, lane, pr, spins);
				if (lane == 0)
					s.pin = pin;
			}
		}
		{
			/* s <-> park: A's masks and payload-start tables come back, B's are parked */
			uint4 *x = (uint4 *)s.M, *y = (uint4 *)sm.park.M;
			for (uint32_t i = tid; i < (uint32_t)(C::NCH + 8) / 8; i += kT) {
				const uint4 u = x[i], v = y[i];
				x[i] = v;
				y[i] = u;
			}
			for (uint32_t i = tid; i < (uint32_t)C::NROW + 1; i += kT) {
				const uint32_t u = s.krow[i], v = sm.park.krow[i];
				s.krow[i] = v;
				sm.park.krow[i] = u;
			}
			for (uint32_t i = tid; i < (uint32_t)kSoff; i += kT) {
				const uint32_t u = s.soff[i], v = sm.park.soff[i];
				s.soff[i] = v;
				sm.park.soff[i] = u;
			}
		}
		__syncthreads();

		/* ==== tile A: the emit of gen 6 ==== */
		if (hasA) {
			const uint32_t t = tA;
			const uint64_t tile_off = (uint64_t)t * C::TILE;
			const uint32_t nvalid = tile_off >= a.len ? 0u
								  : (a.len - tile_off >= (uint64_t)C::TILE ? (uint32_t)C::TILE
													   : (uint32_t)(a.len - tile_off));
			const bool full = nvalid == (uint32_t)C::TILE;
			const uint64_t k_lo = klo_A, k_hi = khi_A;
			const bool has_b = k_hi > k_lo;
			const uint64_t BW = BW_A;
			if (full) {
				bulk_load_wait_parity(&s.bar, parity);
				parity ^= 1u;
				__syncwarp();
			} else {
				load_partial_tile<ROWS>(s, a, tile_off);
				__syncthreads();
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
			if (tid == 0 && t == a.num_tiles - 1) {
				uint32_t tile_total = 0;
#pragma unroll
				for (int j = 0; j < kW; j++)
					tile_total += s.sp_etot[j];
				finish_output(a, k_hi, s.pin + tile_total);
			}
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
		tA = hasB ? tB : 0xffffffffu;
		klo_A = klo_B;
		khi_A = khi_B;
		BW_A = BW_B;
	}
}

} /* namespace frame8 */

#endif /* ANNEXB_FRAME8_CUH */
