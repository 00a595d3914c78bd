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
	/* gen 7 (annexb_frame7.cuh): desc = one word per span, plus the upper levels of its chain */
	uint64_t *group_w; /* num_tiles / 32 + 1 */
	uint64_t *super_w; /* num_tiles / 1024 + 1 */
	uint64_t *super_p; /* num_tiles / 1024 + 1 */
	uint64_t *trace;   /* diagnostics (H264GPU_FRAME_TRACE): 8 words per span, or NULL */
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

} /* namespace frame */

#endif /* ANNEXB_FRAME_CUH */
