/*
 * annexb_scan.cuh — K1/K2: Annex-B start-code scan, NAL table compaction and
 * emulation-prevention-byte strip, fused in ONE pass over the byte stream
 * (single-pass chained scan with decoupled look-back; 1 B read + <=1 B written
 * per input byte).
 *
 * Reference behaviour reproduced bit-exactly (Parrot-Developers/libh264):
 *   - NAL loop of h264_reader_parse            src/h264_reader.c:133-140
 *   - h264_find_nalu / start / end code search src/h264_bitstream.c:87-184
 *   - EPB removal in h264_bs_fetch             include/h264/h264_bitstream.h:168-190
 *
 * Closed form used, b = input bytes:
 *   zz(q) : q+2 < len, b[q]=0, b[q+1]=0, b[q+2]<=1        sc(q) : zz(q), b[q+2]=1
 *   boundary events = { q : zz(q) and (sc(q) or not zz(q-1)) }, in order
 *   every sc(q) starts a NAL at q+3; that NAL ends at the NEXT event (or len)
 *   byte p is inside a NAL  <=>  the last event at or before p is a start code
 *                                and p is not one of its 3 bytes
 *   byte p is an EPB        <=>  b[p]=3, b[p-1]=0, b[p-2]=0  (raw bytes only)
 * so "inside a NAL" is a last-event-wins carry and the RBSP offset of a byte is
 * a prefix sum over (inside and not EPB) — both folded into one tile aggregate.
 *
 * This header holds what the packed-RBSP kernel (annexb_scan2.cuh) and the writer kernels share:
 * the argument block, the byte tests on 32-bit words and the shard-edge rules.  (The
 * first-generation register-resident kernel that used to live here was removed in round 2.)
 */
#ifndef ANNEXB_SCAN_CUH
#define ANNEXB_SCAN_CUH

#include "gpu_compat.h"
#include "h264gpu.h"

namespace annexb {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;
/* descriptor words are pre-set to all-ones; a word is valid once bit 63 is 0 */
constexpr uint64_t kInvalid = 1ull << 63;

struct ScanArgs {
	const uint8_t *in;
	uint64_t len;
	uint64_t base;
	uint8_t *rbsp;
	uint64_t *nal_start;
	uint64_t *nal_end;
	uint64_t *nal_rbsp;
	uint64_t nal_cap;
	uint64_t *desc;   /* 4 words per tile: AGG, P1, P2, pad */
	uint32_t *ticket; /* pre-set to 0xffffffff */
	struct h264gpu_scan_result *result; /* pre-set to all-ones */
	uint32_t num_tiles;
	uint32_t halo_left; /* bytes -4..-1 as a little-endian word, 0xff = none */
	uint8_t right[2];
	uint8_t has_right;
	uint8_t init_in;
	uint64_t *trace; /* diagnostics (H264GPU_SCAN_TRACE): 8 timestamps per tile, or NULL */
	/* gen 5 (in-place RBSP, annexb_scan5.cuh): boundary-event append buffer */
	uint32_t *ev_cursor; /* pre-set to 0xffffffff */
	uint64_t *evbuf;
	uint64_t ev_cap;
	uint32_t flags; /* gen 6 tuning: 1 = one tile per CTA (grid = tiles), 2 = no L2 prefetch */
};

/* 0x80 in every byte of x that is zero, exact */
__device__ __forceinline__ uint32_t zero_bytes_msb(uint32_t x)
{
	uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
	return ~(t | x | 0x7f7f7f7fu);
}

/* bit k set iff byte k of x is zero */
__device__ __forceinline__ uint32_t zmask4(uint32_t x)
{
	return (((zero_bytes_msb(x) >> 7) * 0x00204081u) >> 21) & 0xfu;
}

/* non-zero iff x has a zero byte (cheap boolean form) */
__device__ __forceinline__ uint32_t haszero(uint32_t x)
{
	return (x - 0x01010101u) & ~x & 0x80808080u;
}

__device__ __forceinline__ uint32_t swz(uint32_t w)
{
	/* shared-memory word swizzle: lanes 8 apart write words 32 apart */
	return w ^ ((w >> 5) & 3u);
}

#ifdef H264_EMU
static uint32_t emu_prefix_every = 1;
#endif

struct SlowMasks {
	uint32_t ev16, sc16, insc16;
};

/* exact event masks of a lane's 16 bytes from the 24-byte window [-4, 20) */
__device__ __forceinline__ SlowMasks slow_masks(uint32_t pw, uint32_t w0, uint32_t w1,
						 uint32_t w2, uint32_t w3, uint32_t nw)
{
	const uint32_t k1 = 0x01010101u;
	uint32_t Z = zmask4(pw) | zmask4(w0) << 4 | zmask4(w1) << 8 | zmask4(w2) << 12 |
		     zmask4(w3) << 16 | zmask4(nw) << 20;
	uint32_t O = zmask4(pw ^ k1) | zmask4(w0 ^ k1) << 4 | zmask4(w1 ^ k1) << 8 |
		     zmask4(w2 ^ k1) << 12 | zmask4(w3 ^ k1) << 16 | zmask4(nw ^ k1) << 20;
	uint32_t zz = Z & (Z >> 1) & ((Z | O) >> 2);
	uint32_t sc = Z & (Z >> 1) & (O >> 2);
	uint32_t ev = zz & (sc | ~(zz << 1));
	SlowMasks m;
	m.ev16 = (ev >> 4) & 0xffffu;
	m.sc16 = (sc >> 4) & 0xffffu;
	m.insc16 = ((sc | sc << 1 | sc << 2) >> 4) & 0xffffu;
	return m;
}

/* bytes of the stream at absolute shard offset idx, with the shard-edge rules */
__device__ __forceinline__ uint32_t edge_byte(const ScanArgs &a, uint64_t idx)
{
	if (idx < a.len)
		return a.in[idx];
	uint64_t d = idx - a.len;
	return (a.has_right && d < 2) ? a.right[d] : 0xffu;
}

__device__ __forceinline__ uint32_t edge_word(const ScanArgs &a, uint64_t idx)
{
	return edge_byte(a, idx) | edge_byte(a, idx + 1) << 8 | edge_byte(a, idx + 2) << 16 |
	       edge_byte(a, idx + 3) << 24;
}

/* remove the bytes whose keep bit is 0, packing the kept ones downwards */
__device__ __forceinline__ void compact16(uint32_t &d0, uint32_t &d1, uint32_t &d2,
					   uint32_t &d3, uint32_t keep16)
{
	uint64_t lo = (uint64_t)d0 | ((uint64_t)d1 << 32);
	uint64_t hi = (uint64_t)d2 | ((uint64_t)d3 << 32);
	uint32_t drop = ~keep16 & 0xffffu;
	while (drop) {
		int j = 31 - __clz((int)drop);
		drop &= ~(1u << j);
		if (j >= 8) {
			uint64_t m = (1ull << (8 * (j - 8))) - 1;
			hi = (hi & m) | ((hi >> 8) & ~m);
		} else {
			uint64_t m = (1ull << (8 * j)) - 1;
			lo = (lo & m) | ((lo >> 8) & ~m) | (hi << 56);
			hi >>= 8;
		}
	}
	d0 = (uint32_t)lo;
	d1 = (uint32_t)(lo >> 32);
	d2 = (uint32_t)hi;
	d3 = (uint32_t)(hi >> 32);
}

/* unpack a warp/tile aggregate word */
struct Agg {
	uint32_t head, body, nsc;
	bool ev, st;
};
__device__ __forceinline__ Agg unpack_agg(uint64_t g)
{
	Agg a;
	a.head = (uint32_t)(g & 0xffffu);
	a.body = (uint32_t)((g >> 16) & 0xffffu);
	a.nsc = (uint32_t)((g >> 32) & 0xffffu);
	a.ev = (g >> 48) & 1;
	a.st = (g >> 49) & 1;
	return a;
}
__device__ __forceinline__ uint64_t pack_agg(uint32_t head, uint32_t body, uint32_t nsc,
					      bool ev, bool st)
{
	return (uint64_t)head | (uint64_t)body << 16 | (uint64_t)nsc << 32 |
	       (uint64_t)(ev ? 1 : 0) << 48 | (uint64_t)(st ? 1 : 0) << 49;
}

} /* namespace annexb */

#endif /* ANNEXB_SCAN_CUH */
