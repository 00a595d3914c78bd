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
 * Work decomposition: one CTA = one tile of kBlock*ITEMS*16 bytes; a warp owns a
 * contiguous ITEMS*512-byte region; a lane owns one 16-byte vector per item
 * (coalesced 512-byte warp loads).  Tiles take tickets so that look-back only
 * ever waits on tiles that are already running.
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

template <int ITEMS, bool STRIP>
__global__ void __launch_bounds__(kBlock, 4) scan_kernel(const ScanArgs a)
{
	constexpr int WARP_BYTES = ITEMS * 512;
	constexpr int TILE = kBlock * ITEMS * 16;

	__shared__ uint32_t s_tile;
	__shared__ uint64_t s_wagg[kWarps];
	__shared__ uint64_t s_pin[3]; /* kept_in|flags, nnal_in, tile kept */
	__shared__ __align__(16) uint8_t s_out[STRIP ? TILE + 64 : 16];

	const uint32_t tid = threadIdx.x;
	const uint32_t lane = tid & 31, warp = tid >> 5;

	if (tid == 0)
		s_tile = atomicAdd(a.ticket, 1u) + 1u;
	__syncthreads();
	const uint32_t t = s_tile;
	const uint64_t tile_off = (uint64_t)t * TILE;
	const uint64_t region = tile_off + (uint64_t)warp * WARP_BYTES;
	const bool edge = tile_off + TILE + 4 > a.len;

	/* ---- load: ITEMS coalesced 16-byte vectors per lane + 4-byte halos ---- */
	uint4 v[ITEMS];
	uint32_t hl, hr;
	if (!edge) {
#pragma unroll
		for (int i = 0; i < ITEMS; i++)
			v[i] = ldg_stream16(a.in + region + i * 512 + lane * 16);
		hr = ldg_u32(a.in + region + WARP_BYTES);
		hl = region ? ldg_u32(a.in + region - 4) : a.halo_left;
	} else {
#pragma unroll
		for (int i = 0; i < ITEMS; i++) {
			uint64_t o = region + i * 512 + lane * 16;
			if (o + 16 <= a.len)
				v[i] = ldg_stream16(a.in + o);
			else
				v[i] = make_uint4(edge_word(a, o), edge_word(a, o + 4),
						  edge_word(a, o + 8), edge_word(a, o + 12));
		}
		hr = edge_word(a, region + WARP_BYTES);
		hl = region ? edge_word(a, region - 4) : a.halo_left;
	}

	/* ---- phase A: classify, count, warp-level prefix ---- */
	uint32_t kh[ITEMS];   /* keep-candidate mask | head mask << 16 */
	uint32_t offp[ITEMS]; /* exclusive prefix in the warp region: head | body << 16 */
	uint32_t slowmask = 0;
	bool cur_seen = false, cur_sc = false;
	uint32_t run = 0, nsc_run = 0;

#pragma unroll
	for (int i = 0; i < ITEMS; i++) {
		const uint32_t w0 = v[i].x, w1 = v[i].y, w2 = v[i].z, w3 = v[i].w;
		uint32_t up = __shfl_up_sync(FULL_MASK, w3, 1);
		uint32_t p31 = __shfl_sync(FULL_MASK, i > 0 ? v[i > 0 ? i - 1 : 0].w : 0u, 31);
		uint32_t pw = lane ? up : (i > 0 ? p31 : hl);
		uint32_t dn = __shfl_down_sync(FULL_MASK, w0, 1);
		uint32_t n0 = __shfl_sync(FULL_MASK, i + 1 < ITEMS ? v[i + 1 < ITEMS ? i + 1 : i].x : 0u, 0);
		uint32_t nw = lane < 31 ? dn : (i + 1 < ITEMS ? n0 : hr);

		/* per word: A = bytes two back, B = bytes one back */
		uint32_t ab0 = __funnelshift_l(pw, w0, 16) | __funnelshift_l(pw, w0, 8);
		uint32_t ab1 = __funnelshift_l(w0, w1, 16) | __funnelshift_l(w0, w1, 8);
		uint32_t ab2 = __funnelshift_l(w1, w2, 16) | __funnelshift_l(w1, w2, 8);
		uint32_t ab3 = __funnelshift_l(w2, w3, 16) | __funnelshift_l(w2, w3, 8);
		uint32_t abn = __funnelshift_l(w3, nw, 16) | __funnelshift_l(w3, nw, 8);
		const uint32_t kfe = 0xfefefefeu, k3 = 0x03030303u;
		uint32_t trig = haszero(ab0 | (w0 & kfe)) | haszero(ab1 | (w1 & kfe)) |
				haszero(ab2 | (w2 & kfe)) | haszero(ab3 | (w3 & kfe)) |
				haszero(abn | (nw & kfe) | 0xffff0000u);
		uint32_t epb16 = 0;
		if (STRIP)
			epb16 = zmask4(ab0 | (w0 ^ k3)) | zmask4(ab1 | (w1 ^ k3)) << 4 |
				zmask4(ab2 | (w2 ^ k3)) << 8 | zmask4(ab3 | (w3 ^ k3)) << 12;

		uint32_t vm = 0xffffu;
		if (edge) {
			uint64_t o = region + i * 512 + lane * 16;
			uint64_t nv = o >= a.len ? 0 : (a.len - o >= 16 ? 16 : a.len - o);
			vm = (1u << nv) - 1;
		}

		uint32_t km, hdm;
		const bool slow = __any_sync(FULL_MASK, trig != 0);
		if (!slow) {
			km = ~epb16 & 0xffffu;
			if (cur_seen) {
				hdm = 0;
				if (!cur_sc)
					km = 0;
			} else {
				hdm = 0xffffu;
			}
		} else {
			slowmask |= 1u << i;
			SlowMasks m = slow_masks(pw, w0, w1, w2, w3, nw);
			m.ev16 &= vm;
			m.sc16 &= vm;
			const bool has_ev = m.ev16 != 0;
			const int top = has_ev ? 31 - __clz((int)m.ev16) : 0;
			const bool last_sc = has_ev && ((m.sc16 >> top) & 1);
			const uint32_t bev = __ballot_sync(FULL_MASK, has_ev);
			const uint32_t bls = __ballot_sync(FULL_MASK, last_sc);
			const uint32_t before = bev & ((1u << lane) - 1);
			bool pre_known, pre_in;
			if (before) {
				int pl = 31 - __clz((int)before);
				pre_known = true;
				pre_in = (bls >> pl) & 1;
			} else {
				pre_known = cur_seen;
				pre_in = cur_sc;
			}
			const uint32_t fe = has_ev ? (uint32_t)__ffs((int)m.ev16) - 1 : 16;
			const uint32_t bm = (1u << fe) - 1;
			uint32_t inm = 0;
			for (uint32_t e = m.ev16; e;) {
				int j = __ffs((int)e) - 1;
				e &= e - 1;
				int nj = e ? __ffs((int)e) - 1 : 16;
				if ((m.sc16 >> j) & 1)
					inm |= ((1u << nj) - 1) & ~((1u << j) - 1);
			}
			const uint32_t cand = ~epb16 & ~m.insc16 & 0xffffu;
			km = cand & ((pre_known ? (pre_in ? bm : 0u) : bm) | inm);
			hdm = pre_known ? 0u : bm;
			uint32_t ns = (uint32_t)__popc(m.sc16);
#pragma unroll
			for (int d = 16; d >= 1; d >>= 1)
				ns += __shfl_xor_sync(FULL_MASK, ns, d);
			nsc_run += ns;
			if (bev) {
				int tl = 31 - __clz((int)bev);
				cur_seen = true;
				cur_sc = (bls >> tl) & 1;
			}
		}
		km &= vm;
		kh[i] = km | hdm << 16;
		if (STRIP) {
			uint32_t cnt = (uint32_t)__popc(km & hdm) | (uint32_t)__popc(km & ~hdm) << 16;
			uint32_t inc = cnt;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
				if (lane >= (uint32_t)d)
					inc += o;
			}
			offp[i] = run + inc - cnt;
			run += __shfl_sync(FULL_MASK, inc, 31);
		} else {
			offp[i] = 0;
		}
	}
	if (lane == 0)
		s_wagg[warp] = pack_agg(run & 0xffffu, run >> 16, nsc_run, cur_seen, cur_sc);
	__syncthreads();

	/* ---- phase B (warp 0): tile aggregate, decoupled look-back, publish ---- */
	if (warp == 0) {
		uint32_t th = 0, tb = 0, tn = 0;
		bool tev = false, tst = false;
#pragma unroll
		for (int j = 0; j < kWarps; j++) {
			Agg g = unpack_agg(s_wagg[j]);
			if (!tev) {
				th += g.head;
				tb += g.body;
			} else {
				tb += (tst ? g.head : 0u) + g.body;
			}
			tn += g.nsc;
			if (g.ev) {
				tev = true;
				tst = g.st;
			}
		}
		uint64_t *dt = a.desc + (uint64_t)t * 4;
		uint64_t kept_in = 0, nnal_in = 0;
		bool sc_in = a.init_in != 0, any_in = false;
		if (t > 0) {
			if (lane == 0)
				st_relaxed_u64(dt, pack_agg(th, tb, tn, tev, tst));
			/* running aggregate of the tiles between the prefix found and t */
			uint64_t ah = 0, ab = 0, an = 0;
			bool aev = false, ast = false;
			int64_t j0 = (int64_t)t - 1;
			for (bool done = false; !done;) {
				const int64_t j = j0 - lane;
				const bool need = j >= 0;
				uint64_t agg = kInvalid, p1 = kInvalid;
				uint32_t pm;
				for (;;) {
					if (need) {
						p1 = ld_relaxed_u64(a.desc + j * 4 + 1);
						if (p1 & kInvalid)
							agg = ld_relaxed_u64(a.desc + j * 4);
					}
					pm = __ballot_sync(FULL_MASK, need && !(p1 & kInvalid));
					uint32_t ok = __ballot_sync(
						FULL_MASK, !need || !(p1 & kInvalid) || !(agg & kInvalid));
					uint32_t upto = pm ? ((pm & (0u - pm)) << 1) - 1u : 0xffffffffu;
					if ((ok & upto) == upto)
						break;
				}
				const int fp = pm ? __ffs((int)pm) - 1 : 32;
				/* Ordered tree reduction of the window's aggregates: lane k
				 * holds tile j0-k, so a higher lane is the LEFT operand. */
				uint32_t eh = 0, eb = 0, em = 0; /* em: nsc | ev<<30 | st<<31 */
				if ((int)lane < fp && need) {
					Agg g = unpack_agg(agg);
					eh = g.head;
					eb = g.body;
					em = g.nsc | (g.ev ? 1u << 30 : 0u) | (g.st ? 1u << 31 : 0u);
				}
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
					/* acc = combine(window, acc): the window is LEFT of acc */
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
					uint64_t p2 = kInvalid;
					if ((int)lane == fp) {
						do {
							p2 = ld_relaxed_u64(a.desc + j * 4 + 2);
						} while (p2 & kInvalid);
					}
					p2 = __shfl_sync(FULL_MASK, p2, fp);
					uint64_t P1 = __shfl_sync(FULL_MASK, p1, fp);
					bool psc = (P1 >> 62) & 1, pany = (P1 >> 61) & 1;
					uint64_t pk = P1 & ((1ull << 61) - 1);
					kept_in = pk + (psc ? ah : 0) + ab;
					nnal_in = p2 + an;
					sc_in = aev ? ast : psc;
					any_in = pany || aev;
					done = true;
				} else {
					j0 -= 32;
				}
			}
		}
		const uint64_t kept_out = kept_in + (sc_in ? th : 0u) + tb;
		const uint64_t nnal_out = nnal_in + tn;
		const bool sc_out = tev ? tst : sc_in;
		const bool any_out = any_in || tev;
		if (lane == 0) {
#ifdef H264_EMU
			/* test hook: withhold most prefixes so look-back must fold aggregates */
			if (emu_prefix_every <= 1 || t % emu_prefix_every == 0)
#endif
			{
				st_relaxed_u64(dt + 1, kept_out | (uint64_t)(sc_out ? 1 : 0) << 62 |
							       (uint64_t)(any_out ? 1 : 0) << 61);
				st_relaxed_u64(dt + 2, nnal_out);
			}
			s_pin[0] = kept_in | (uint64_t)(sc_in ? 1 : 0) << 62 |
				   (uint64_t)(any_in ? 1 : 0) << 61;
			s_pin[1] = nnal_in;
			s_pin[2] = kept_out - kept_in;
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
	__syncthreads();

	/* ---- phase C: NAL table entries, scatter to shared, vector copy-out ---- */
	const uint64_t kept_in = s_pin[0] & ((1ull << 61) - 1);
	const bool sc_in = (s_pin[0] >> 62) & 1, any_in = (s_pin[0] >> 61) & 1;
	const uint64_t nnal_in = s_pin[1];
	const uint32_t tile_kept = (uint32_t)s_pin[2];
	if (!STRIP && slowmask == 0)
		return; /* warp-uniform; no barriers below for !STRIP */

	bool seenB = false, stB = false;
	uint32_t baseOff = 0, baseNsc = 0;
	for (uint32_t j = 0; j < warp; j++) {
		Agg g = unpack_agg(s_wagg[j]);
		bool hk = seenB ? stB : sc_in;
		baseOff += (hk ? g.head : 0u) + g.body;
		baseNsc += g.nsc;
		if (g.ev) {
			seenB = true;
			stB = g.st;
		}
	}
	const bool headKeep = seenB ? stB : sc_in;
	const uint32_t shift = (uint32_t)(kept_in & 15);

	uint32_t fk[ITEMS], Gb[ITEMS];
#pragma unroll
	for (int i = 0; i < ITEMS; i++) {
		uint32_t km = kh[i] & 0xffffu, hdm = kh[i] >> 16;
		fk[i] = headKeep ? km : (km & ~hdm);
		Gb[i] = STRIP ? __ballot_sync(FULL_MASK, __popc(fk[i]) >= 4) : 0u;
	}

	cur_seen = false;
	cur_sc = false;
	nsc_run = 0;
	uint32_t carry_tail = 0;
	uint32_t *so32 = (uint32_t *)s_out;

#pragma unroll
	for (int i = 0; i < ITEMS; i++) {
		uint32_t d0 = v[i].x, d1 = v[i].y, d2 = v[i].z, d3 = v[i].w;
		const uint32_t laneoff = (headKeep ? (offp[i] & 0xffffu) : 0u) + (offp[i] >> 16);

		if (slowmask >> i & 1) {
			uint32_t up = __shfl_up_sync(FULL_MASK, d3, 1);
			uint32_t p31 = __shfl_sync(FULL_MASK, i > 0 ? v[i > 0 ? i - 1 : 0].w : 0u, 31);
			uint32_t pw = lane ? up : (i > 0 ? p31 : hl);
			uint32_t dn = __shfl_down_sync(FULL_MASK, d0, 1);
			uint32_t n0 = __shfl_sync(FULL_MASK, i + 1 < ITEMS ? v[i + 1 < ITEMS ? i + 1 : i].x : 0u, 0);
			uint32_t nw = lane < 31 ? dn : (i + 1 < ITEMS ? n0 : hr);
			uint32_t vm = 0xffffu;
			const uint64_t lane_abs = region + i * 512 + lane * 16;
			if (edge) {
				uint64_t nv = lane_abs >= a.len ? 0 : (a.len - lane_abs >= 16 ? 16 : a.len - lane_abs);
				vm = (1u << nv) - 1;
			}
			SlowMasks m = slow_masks(pw, d0, d1, d2, d3, nw);
			m.ev16 &= vm;
			m.sc16 &= vm;
			const bool has_ev = m.ev16 != 0;
			const int top = has_ev ? 31 - __clz((int)m.ev16) : 0;
			const bool last_sc = has_ev && ((m.sc16 >> top) & 1);
			const uint32_t bev = __ballot_sync(FULL_MASK, has_ev);
			const uint32_t bls = __ballot_sync(FULL_MASK, last_sc);
			const uint32_t before = bev & ((1u << lane) - 1);
			bool pre_known, pre_in;
			if (before) {
				int pl = 31 - __clz((int)before);
				pre_known = true;
				pre_in = (bls >> pl) & 1;
			} else {
				pre_known = cur_seen;
				pre_in = cur_sc;
			}
			uint32_t ns = (uint32_t)__popc(m.sc16), inc = ns;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				uint32_t o = __shfl_up_sync(FULL_MASK, inc, d);
				if (lane >= (uint32_t)d)
					inc += o;
			}
			const uint32_t ns_item = __shfl_sync(FULL_MASK, inc, 31);
			if (has_ev) {
				bool prev_sc = pre_known ? pre_in : headKeep;
				uint64_t nscb = nnal_in + baseNsc + nsc_run + (inc - ns);
				bool first = !any_in && !seenB && !cur_seen && before == 0;
				for (uint32_t e = m.ev16; e;) {
					int j = __ffs((int)e) - 1;
					e &= e - 1;
					const bool is_sc = (m.sc16 >> j) & 1;
					const uint64_t q = a.base + lane_abs + j;
					const uint64_t keptb = kept_in + baseOff + laneoff +
							       (uint32_t)__popc(fk[i] & ((1u << j) - 1));
					if (prev_sc && nscb >= 1 && nscb - 1 < a.nal_cap)
						a.nal_end[nscb - 1] = q;
					if (first) {
						a.result->first_event_pos = q;
						a.result->first_event_is_sc = is_sc ? 1u : 0u;
						a.result->head_bytes = STRIP ? keptb : 0;
						first = false;
					}
					if (is_sc) {
						if (nscb < a.nal_cap) {
							a.nal_start[nscb] = q + 3;
							if (a.nal_rbsp)
								a.nal_rbsp[nscb] = STRIP ? keptb : 0;
						}
						nscb++;
					}
					prev_sc = is_sc;
				}
			}
			nsc_run += ns_item;
			if (bev) {
				int tl = 31 - __clz((int)bev);
				cur_seen = true;
				cur_sc = (bls >> tl) & 1;
			}
		}

		if (STRIP) {
			const uint32_t keep = fk[i];
			const uint32_t c = (uint32_t)__popc(keep);
			if (keep != 0xffffu && c != 0)
				compact16(d0, d1, d2, d3, keep);
			const uint32_t o = shift + baseOff + laneoff;
			const uint32_t s = o & 3, wi = o >> 2, sh = 8 * s;
			const uint32_t e = s + c;
			/* the lane's bytes as 5 words aligned to shared-memory words */
			const uint32_t x0 = d0 << sh;
			const uint32_t x1 = __funnelshift_rc(d0, d1, 32 - sh);
			const uint32_t x2 = __funnelshift_rc(d1, d2, 32 - sh);
			const uint32_t x3 = __funnelshift_rc(d2, d3, 32 - sh);
			const uint32_t x4 = __funnelshift_rc(d3, 0u, 32 - sh);
			const uint32_t tw = e >> 2; /* word holding the partial tail */
			uint32_t tailw = tw == 0 ? x0 : tw == 1 ? x1 : tw == 2 ? x2 : tw == 3 ? x3 : x4;
			uint32_t ptail = __shfl_up_sync(FULL_MASK, tailw, 1);
			if (lane == 0)
				ptail = carry_tail;
			carry_tail = __shfl_sync(FULL_MASK, tailw, 31);
			const bool myG = c >= 4;
			const bool prevG = lane ? ((Gb[i] >> (lane - 1)) & 1)
						: (i > 0 && (Gb[i > 0 ? i - 1 : 0] >> 31));
			const bool nextG = lane < 31 ? ((Gb[i] >> (lane + 1)) & 1)
						     : (i + 1 < ITEMS && (Gb[i + 1 < ITEMS ? i + 1 : i] & 1));
			if (myG) {
				if (s == 0) {
					so32[swz(wi)] = x0;
				} else if (prevG) {
					const uint32_t lom = (1u << sh) - 1;
					so32[swz(wi)] = (x0 & ~lom) | (ptail & lom);
				} else {
					for (uint32_t b = s; b < 4; b++)
						s_out[(swz(wi) << 2) | b] = (uint8_t)(x0 >> (8 * b));
				}
				if (8 <= e)
					so32[swz(wi + 1)] = x1;
				if (12 <= e)
					so32[swz(wi + 2)] = x2;
				if (16 <= e)
					so32[swz(wi + 3)] = x3;
				if (20 <= e)
					so32[swz(wi + 4)] = x4;
				if ((e & 3) && !nextG) {
					for (uint32_t b = 0; b < (e & 3); b++)
						s_out[(swz(wi + tw) << 2) | b] = (uint8_t)(tailw >> (8 * b));
				}
			} else {
				for (uint32_t b = 0; b < c; b++) {
					uint32_t ob = o + b;
					s_out[(swz(ob >> 2) << 2) | (ob & 3)] = (uint8_t)(d0 >> (8 * b));
				}
			}
		}
	}

	if (STRIP) {
		__syncthreads();
		const uint32_t total = shift + tile_kept;
		const uint32_t nvec = (total + 15) >> 4;
		uint8_t *gbase = a.rbsp + (kept_in - shift);
		for (uint32_t x = tid; x < nvec; x += kBlock) {
			/* word m of vector x lives at 4x + (m ^ r); r is constant per thread */
			const uint32_t r = (x >> 3) & 3;
			uint4 val;
			val.x = so32[4 * x + r];
			val.y = so32[4 * x + (1 ^ r)];
			val.z = so32[4 * x + (2 ^ r)];
			val.w = so32[4 * x + (3 ^ r)];
			const uint32_t lo = x * 16;
			if (lo >= shift && lo + 16 <= total) {
				stg_stream16(gbase + lo, val);
			} else {
				const uint32_t b0 = lo >= shift ? lo : shift;
				const uint32_t b1 = lo + 16 <= total ? lo + 16 : total;
				const uint32_t wv[4] = {val.x, val.y, val.z, val.w};
				for (uint32_t b = b0; b < b1; b++)
					gbase[b] = (uint8_t)(wv[(b >> 2) & 3] >> (8 * (b & 3)));
			}
		}
	}
}

} /* namespace annexb */

#endif /* ANNEXB_SCAN_CUH */
