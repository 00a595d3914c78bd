/*
 * cavlc_parse.cuh — K4: slice-parallel CAVLC macroblock syntax parse, one
 * independent slice per GPU thread.
 *
 * Reference behaviour reproduced (Parrot-Developers/libh264), frame/field pictures
 * without MBAFF, one slice group:
 *   slice_data loop, skip runs, end test   src/h264_syntax_slice_data.h:701-787
 *   macroblock_layer                       src/h264_syntax_slice_data.h:604-696
 *   mb_pred / sub_mb_pred                  src/h264_syntax_slice_data.h:422-601
 *   residual, residual_luma, residual_block src/h264_syntax_slice_data.h:103-419
 *   mb_type / sub_mb_type / cbp mapping    src/h264_slice_data.c:839-1080
 *   coeff_token (nC from the slice-local neighbours), total_zeros, run_before
 *                                          src/h264_slice_data.c:1239-1416
 *   neighbouring 4x4 blocks, non-MBAFF     src/h264_macroblock.c:84-108,306-433
 *   bit reader with EPB skip, ue/se/te     include/h264/h264_bitstream.h:168-304,
 *                                          src/h264_bitstream.c:190-208
 *   h264_bs_more_rbsp_data (incl. its one-trailing-zero-byte tolerance)
 *                                          src/h264_bitstream.c:325-355
 *
 * Design: the parse of one slice is inherently serial (every code's position
 * depends on all earlier ones), so parallelism is across slices.  Per thread:
 * a 64-bit bit cache refilled bytewise from the ESCAPED stream (EPB skipped on the
 * fly, raw offsets kept exact for the end-of-slice test), VLCs decoded by
 * count-leading-zeros + one table probe (the reference probes once per bit),
 * nC context in a ring of the last PicWidthInMbs+1 macroblocks' counts.
 * Output per macroblock: (mb_addr, mb_type) — what slice_data_mb delivers — plus
 * the order-independent checksum of every decoded syntax element
 * (include/h264gpu_slice.h).
 */
/*
 * The body below is compiled once per namespace: CAVLC_NS = cavlc with CAVLC_FULL 0 (the
 * default kernel: no trace of the full-record path in its code) and, when the includer asks
 * for it by including this file again with CAVLC_NS = cavlc_full and CAVLC_FULL 1, the
 * variant that also fills struct h264_mb_syntax records.
 */
#ifndef CAVLC_PARSE_PRELUDE
#define CAVLC_PARSE_PRELUDE

#include "gpu_compat.h"
#include "h264gpu_slice.h"
#include "h264gpu_mb_syntax.h"

#ifdef H264_EMU
#define CAVLC_TAB static const
#else
#define CAVLC_TAB __device__ const
#endif
#include "cavlc_luts.h"

#ifndef EIO
#define EIO 5
#endif
#ifndef ENOSYS
#define ENOSYS 38
#endif
#ifndef ENOBUFS
#define ENOBUFS 105
#endif

#endif /* CAVLC_PARSE_PRELUDE */

#ifndef CAVLC_NS
#define CAVLC_NS cavlc
#define CAVLC_FULL 0
#endif

namespace CAVLC_NS {

/* enum h264_mb_type values (include/h264/h264_types.h:70-90) */
enum {
	MB_UNKNOWN = 0, MB_I_NxN, MB_I_16x16, MB_I_PCM, MB_SI, MB_P_16x16, MB_P_16x8, MB_P_8x16,
	MB_P_8x8, MB_P_8x8ref0, MB_P_SKIP, MB_B_Direct_16x16, MB_B_16x16, MB_B_16x8, MB_B_8x16,
	MB_B_8x8, MB_B_SKIP,
};
enum { PM_I4 = 0, PM_I8, PM_I16, PM_L0, PM_L1, PM_BI, PM_DIRECT };
enum { ST_P = 0, ST_B = 1, ST_I = 2, ST_SP = 3, ST_SI = 4 };

/* ---- bit reader over the escaped NAL ------------------------------------------ */
struct BitReader {
	const uint8_t *p; /* NAL start (header byte) */
	uint32_t len;
	uint32_t pos;     /* raw offset of the next byte to load */
	uint64_t cache;   /* unread bits, MSB first */
	int nbits;
	uint32_t zeros;   /* run of raw zero bytes just before pos */
	uint32_t epbq;    /* bit k: an EPB was skipped right before the k-th most recent byte */
	bool err;
};

/* 4 raw bytes at p (any alignment) as a big-endian word; reads the two aligned words around them */
__device__ __forceinline__ uint32_t load_be32(const uint8_t *p)
{
#ifdef H264_EMU
	return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
#else
	const uintptr_t a = (uintptr_t)p;
	const uint32_t *w = (const uint32_t *)(a & ~(uintptr_t)3);
	const uint32_t v = __funnelshift_r(w[0], w[1], (uint32_t)(a & 3) * 8);
	return __byte_perm(v, 0, 0x0123);
#endif
}

__device__ __noinline__ void br_refill(BitReader &b)
{
	/* word at a time: a word without a 03 byte holds no emulation prevention byte, whatever came
	 * before it (8 bytes of the NAL must be left so that the aligned loads stay inside it) */
	while (b.nbits <= 32 && b.pos + 8 <= b.len) {
		const uint32_t w = load_be32(b.p + b.pos);
		const uint32_t x = w ^ 0x03030303u;
		if ((x - 0x01010101u) & ~x & 0x80808080u)
			break; /* a 03 byte: the byte path decides */
		b.cache |= (uint64_t)w << (32 - b.nbits);
		b.nbits += 32;
		b.pos += 4;
		b.epbq <<= 4;
		b.zeros = (w & 0xffffu) == 0 ? 2u : ((w & 0xffu) == 0 ? 1u : 0u);
	}
	if (b.nbits > 32)
		return; /* no reader asks for more than 32 bits at once */
	while (b.nbits <= 56 && b.pos < b.len) {
		/* up to 5 raw bytes in flight at once (independent loads), then the EPB
		 * state machine runs on registers */
		const uint32_t avail = b.len - b.pos;
		uint64_t pk = 0;
#pragma unroll
		for (int j = 0; j < 5; j++)
			pk |= (uint64_t)((uint32_t)j < avail ? b.p[b.pos + j] : 0u) << (8 * j);
		uint32_t k = 0;
		while (b.nbits <= 56 && k < 4 && k < avail) {
			uint32_t c = (uint32_t)(pk >> (8 * k)) & 0xffu;
			uint32_t skipped = 0;
			if (b.zeros >= 2 && c == 3) {
				if (k + 1 >= avail) {
					b.pos += k;
					return; /* the reference fails the fetch here */
				}
				k++;
				c = (uint32_t)(pk >> (8 * k)) & 0xffu;
				skipped = 1;
				b.zeros = 0;
			}
			b.zeros = c == 0 ? b.zeros + 1 : 0;
			k++;
			b.cache |= (uint64_t)c << (56 - b.nbits);
			b.nbits += 8;
			b.epbq = (b.epbq << 1) | skipped;
		}
		b.pos += k;
	}
}

/* position the reader at raw bit offset bit_off of the NAL */
__device__ __forceinline__ void br_init(BitReader &b, const uint8_t *nal, uint32_t len,
					 uint32_t bit_off)
{
	b.p = nal;
	b.len = len;
	b.cache = 0;
	b.nbits = 0;
	b.epbq = 0;
	b.err = false;
	/* raw byte holding the first slice_data bit; if bit_off is byte aligned the
	 * reference has not fetched that byte yet */
	uint32_t byte = bit_off >> 3;
	b.pos = byte;
	b.zeros = 0;
	if (byte >= 1 && nal[byte - 1] == 0)
		b.zeros = (byte >= 2 && nal[byte - 2] == 0) ? 2 : 1;
	/* hdr_len counts raw bits, EPBs included: if `byte` itself is an EPB position the
	 * header parse already skipped it, which only happens when bit_off is past it */
	br_refill(b);
	uint32_t drop = bit_off & 7;
	if (drop) {
		b.cache <<= drop;
		b.nbits -= (int)drop;
	}
}

__device__ __forceinline__ uint32_t br_peek(BitReader &b, int n)
{
	if (b.nbits < n)
		br_refill(b);
	return n ? (uint32_t)(b.cache >> (64 - n)) : 0;
}

__device__ __forceinline__ void br_skip(BitReader &b, int n)
{
	if (b.nbits < n) {
		br_refill(b);
		if (b.nbits < n) {
			b.err = true;
			b.cache = 0;
			b.nbits = 0;
			return;
		}
	}
	b.cache <<= n;
	b.nbits -= n;
}

__device__ __noinline__ uint32_t br_get(BitReader &b, int n)
{
	uint32_t v = br_peek(b, n);
	br_skip(b, n);
	return v;
}

/* 9.1 Exp-Golomb; same result as h264_bs_read_bits_ue for codes the reference
 * handles without undefined behaviour (fewer than 32 leading zeros) */
__device__ __noinline__ uint32_t br_ue(BitReader &b)
{
	uint32_t top = br_peek(b, 32);
	if (top == 0) {
		b.err = true;
		return 0;
	}
	int lz = __clz((int)top);
	if (lz <= 15)
		return br_get(b, 2 * lz + 1) - 1;
	br_skip(b, lz);
	return br_get(b, lz + 1) - 1;
}

__device__ __noinline__ int32_t br_se(BitReader &b)
{
	uint32_t u = br_ue(b);
	return (u & 1) ? (int32_t)((u + 1) >> 1) : -(int32_t)(u >> 1);
}

__device__ __noinline__ uint32_t br_te(BitReader &b, uint32_t max)
{
	if (max == 1)
		return br_get(b, 1) ^ 1;
	return br_ue(b);
}

__device__ __forceinline__ bool br_aligned(const BitReader &b)
{
	return (b.nbits & 7) == 0;
}

/* raw bit position of the next unread bit (for the result record) */
__device__ __forceinline__ uint64_t br_raw_bitpos(const BitReader &b)
{
	int whole = b.nbits >> 3;
	uint32_t epbs = (uint32_t)__popc(b.epbq & ((1u << whole) - 1));
	uint32_t next_byte = b.pos - (uint32_t)whole - epbs; /* first raw byte not begun */
	return (uint64_t)next_byte * 8 - (uint32_t)(b.nbits & 7);
}

/* h264_bs_more_rbsp_data, src/h264_bitstream.c:325-355 */
__device__ __noinline__ bool br_more_rbsp_data(BitReader &b)
{
	if (b.nbits < 8)
		br_refill(b);
	if (b.nbits == 0)
		return false;
	int k = (b.nbits & 7) ? (b.nbits & 7) : 8;
	uint32_t v = (uint32_t)(b.cache >> (64 - k));
	if (v != (1u << (k - 1)))
		return true;
	/* stop bit pattern: what lies after that byte decides */
	int whole = (b.nbits - k) >> 3;
	uint32_t epbs = (uint32_t)__popc(b.epbq & ((1u << whole) - 1));
	uint32_t off = b.pos - (uint32_t)whole - epbs;
	if (off >= b.len)
		return false;
	return off + 1 < b.len || b.p[off] != 0;
}

/* ---- VLC probes ----------------------------------------------------------------- */
template <int K>
__device__ __forceinline__ uint32_t vlc_probe(BitReader &b, const uint16_t *lut)
{
	uint32_t top = br_peek(b, 32);
	int lz = top ? __clz((int)top) : 32;
	if (lz > 15)
		lz = 15;
	uint32_t rest = K ? ((top << (lz + 1)) >> (32 - K)) : 0;
	uint32_t e = lut[(lz << K) | rest];
	if (e == 0) {
		b.err = true;
		return 0;
	}
	br_skip(b, (int)(e & 0xff));
	return e >> 8;
}

/* ---- per-slice state -------------------------------------------------------------- */
struct SliceCtx {
	BitReader br;
	const h264gpu_slice_params *sp;
	uint32_t W;
	uint32_t cat;      /* ChromaArrayType */
	uint32_t mbw_c4, mbh_c4; /* chroma MB size in 4x4 blocks */
	uint32_t cur;      /* current mbAddr */
	uint8_t *ring;     /* (W+1) x 48 bytes: total_coeff of the last W+1 macroblocks */
	uint8_t nz[48];    /* current macroblock */
	bool availA, availB;
	uint64_t hash;
	h264_mb_syntax *syn; /* full record of the current macroblock, or NULL */
};

#if CAVLC_FULL
/* the element's place in the full record (same indexing as the checksum fields) */
__device__ __noinline__ void syn_store(h264_mb_syntax *o, uint32_t field, uint32_t idx, int64_t v)
{
	switch (field) {
	case H264GPU_F_RAW_MB_TYPE: o->raw_mb_type = (uint32_t)v; break;
	case H264GPU_F_TRANSFORM_8X8: o->transform_size_8x8_flag = (uint8_t)v; break;
	case H264GPU_F_MB_QP_DELTA: o->mb_qp_delta = (int32_t)v; break;
	case H264GPU_F_CBP_LUMA: o->cbp_luma = (uint8_t)v; break;
	case H264GPU_F_CBP_CHROMA: o->cbp_chroma = (uint8_t)v; break;
	case H264GPU_F_INTRA_CHROMA_PRED_MODE: o->intra_chroma_pred_mode = (uint8_t)v; break;
	case H264GPU_F_INTRA4X4_PRED_MODE: o->intra4x4_pred_mode[idx & 15] = (int8_t)v; break;
	case H264GPU_F_INTRA8X8_PRED_MODE: o->intra8x8_pred_mode[idx & 3] = (int8_t)v; break;
	case H264GPU_F_REF_IDX_L0: o->ref_idx[0][idx & 3] = (uint8_t)v; break;
	case H264GPU_F_REF_IDX_L1: o->ref_idx[1][idx & 3] = (uint8_t)v; break;
	case H264GPU_F_MVD_L0: o->mvd[0][(idx >> 1) & 15][idx & 1] = (int16_t)v; break;
	case H264GPU_F_MVD_L1: o->mvd[1][(idx >> 1) & 15][idx & 1] = (int16_t)v; break;
	case H264GPU_F_RAW_SUB_MB_TYPE: o->raw_sub_mb_type[idx & 3] = (uint32_t)v; break;
	case H264GPU_F_I16_DC: o->dc16[idx & 15] = (int16_t)v; break;
	case H264GPU_F_I16_AC: o->ac16[(idx >> 4) & 15][idx & 15] = (int16_t)v; break;
	case H264GPU_F_LEVEL4X4: o->l4[(idx >> 4) & 15][idx & 15] = (int16_t)v; break;
	case H264GPU_F_CHROMA_DC: o->cdc[(idx >> 4) & 1][idx & 15] = (int16_t)v; break;
	case H264GPU_F_CHROMA_AC: o->cac[(idx >> 8) & 1][(idx >> 4) & 15][idx & 15] = (int16_t)v; break;
	case H264GPU_F_PCM_LUMA: o->pcm[idx & 255] = (uint8_t)v; break;
	case H264GPU_F_PCM_CHROMA: o->pcm[256 + (idx & 511)] = (uint8_t)v; break;
	default: break; /* CBP (derived), I16 pred mode (in raw_mb_type), Cb / Cr of 4:4:4: not in the record */
	}
}
#endif

__device__ __forceinline__ void hash_add(SliceCtx &s, uint32_t field, uint32_t idx, int64_t v)
{
#if CAVLC_FULL
	if (s.syn != nullptr && v != 0)
		syn_store(s.syn, field, idx, v);
#endif
	/* device-side twin of h264gpu_mb_hash_term (include/h264gpu_slice.h) */
	if (v != 0) {
		const uint64_t key = ((uint64_t)field << 16) | idx;
		s.hash += (uint64_t)v * (((key + 1) * 0x9E3779B97F4A7C15ull) | 1ull);
	}
}

__device__ __forceinline__ uint8_t *ring_slot(SliceCtx &s, uint32_t mb)
{
	return s.ring + (size_t)(mb % (s.W + 1)) * 48;
}

/* 6.4.11.4 / 6.4.11.5: total_coeff of the left and upper 4x4 neighbours -> nC */
__device__ __noinline__ uint32_t calc_nc(SliceCtx &s, uint32_t comp, uint32_t blk, bool chroma_ac)
{
	uint32_t nA = 0, nB = 0;
	bool aA, aB;
	if (!chroma_ac) {
		/* x,y of the block in 4-sample units from the z-order index */
		const uint32_t x = (blk & 1) | ((blk >> 1) & 2);
		const uint32_t y = ((blk >> 1) & 1) | ((blk >> 2) & 2);
		/* inverse: idx(x,y) */
#define LIDX(xx, yy) ((((xx) & 1) | (((yy) & 1) << 1) | (((xx) & 2) << 1) | (((yy) & 2) << 2)))
		if (x > 0) {
			aA = true;
			nA = s.nz[comp * 16 + LIDX(x - 1, y)];
		} else if ((aA = s.availA)) {
			nA = ring_slot(s, s.cur - 1)[comp * 16 + LIDX(3, y)];
		}
		if (y > 0) {
			aB = true;
			nB = s.nz[comp * 16 + LIDX(x, y - 1)];
		} else if ((aB = s.availB)) {
			nB = ring_slot(s, s.cur - s.W)[comp * 16 + LIDX(x, 3)];
		}
#undef LIDX
	} else {
		const uint32_t x = blk & 1, y = blk >> 1;
		if (x > 0) {
			aA = true;
			nA = s.nz[comp * 16 + blk - 1];
		} else if ((aA = s.availA)) {
			nA = ring_slot(s, s.cur - 1)[comp * 16 + 2 * y + (s.mbw_c4 - 1)];
		}
		if (y > 0) {
			aB = true;
			nB = s.nz[comp * 16 + blk - 2];
		} else if ((aB = s.availB)) {
			nB = ring_slot(s, s.cur - s.W)[comp * 16 + 2 * (s.mbh_c4 - 1) + x];
		}
	}
	if (aA && aB)
		return (nA + nB + 1) >> 1;
	return aA ? nA : aB ? nB : 0;
}

/*
 * residual_block (CAVLC).  tab: 0..2 VLC by nC, 3 FLC, 4 chroma DC 4:2:0, 5 chroma DC
 * 4:2:2.  Coefficient i of the block contributes to the checksum as
 * (field, idx_base + startIdx + position).  Returns total_coeff.
 */
__device__ __noinline__ uint32_t residual_block(SliceCtx &s, uint32_t tab, uint32_t max_num_coeff,
						    uint32_t field, uint32_t idx_base)
{
	BitReader &b = s.br;
	uint32_t tok;
	if (tab == 3) {
		uint32_t e = cavlc_coeff_token_flc[br_get(b, 6)];
		if (e == 0) {
			b.err = true;
			return 0;
		}
		tok = e & 0x7f;
	} else {
		tok = vlc_probe<CAVLC_COEFF_TOKEN_K>(b, cavlc_coeff_token[tab <= 2 ? tab : tab - 1]);
	}
	const uint32_t t1 = (tok >> 5) & 3, tc = tok & 0x1f;
	if (tc == 0 || b.err)
		return 0;
	if (tc > max_num_coeff) {
		b.err = true;
		return 0;
	}
	int16_t lev[16];
	uint32_t suffix_len = (tc > 10 && t1 < 3) ? 1 : 0;
	for (uint32_t i = 0; i < tc; i++) {
		if (i < t1) {
			lev[i] = (int16_t)(1 - 2 * (int)br_get(b, 1));
			continue;
		}
		uint32_t top = br_peek(b, 32);
		uint32_t prefix = top ? (uint32_t)__clz((int)top) : 32;
		if (prefix > 25) {
			b.err = true; /* the reference caps level_prefix at 25 */
			return 0;
		}
		br_skip(b, (int)prefix + 1);
		uint32_t code = (prefix < 15 ? prefix : 15) << suffix_len;
		if (suffix_len > 0 || prefix >= 14) {
			uint32_t sz = (prefix == 14 && suffix_len == 0) ? 4 : prefix >= 15 ? prefix - 3 : suffix_len;
			if (sz)
				code += br_get(b, (int)sz);
		}
		if (prefix >= 15 && suffix_len == 0)
			code += 15;
		if (prefix >= 16)
			code += (1u << (prefix - 3)) - 4096;
		if (i == t1 && t1 < 3)
			code += 2;
		int32_t v = (code & 1) ? (-(int32_t)code - 1) >> 1 : (int32_t)((code + 2) >> 1);
		lev[i] = (int16_t)v;
		if (suffix_len == 0)
			suffix_len = 1;
		int32_t a = lev[i] < 0 ? -lev[i] : lev[i];
		if (a > (3 << (suffix_len - 1)) && suffix_len < 6)
			suffix_len++;
	}
	uint32_t zeros_left = 0;
	if (tc < max_num_coeff) {
		if (max_num_coeff == 4)
			zeros_left = vlc_probe<CAVLC_TOTAL_ZEROS_2X2_K>(b, cavlc_total_zeros_2x2[tc]);
		else if (max_num_coeff == 8)
			zeros_left = vlc_probe<CAVLC_TOTAL_ZEROS_2X4_K>(b, cavlc_total_zeros_2x4[tc]);
		else
			zeros_left = vlc_probe<CAVLC_TOTAL_ZEROS_4X4_K>(b, cavlc_total_zeros_4x4[tc]);
	}
	if (b.err)
		return 0;
	/* positions: coefficient i sits run_before(i) zeros above coefficient i+1 */
	int32_t pos = (int32_t)(tc + zeros_left) - 1; /* index of the highest coefficient */
	for (uint32_t i = 0; i < tc; i++) {
		if (pos < 0 || pos >= (int32_t)max_num_coeff) {
			b.err = true;
			return 0;
		}
		hash_add(s, field, idx_base + (uint32_t)pos, lev[i]);
		uint32_t run = 0;
		if (i + 1 < tc && zeros_left > 0) {
			run = vlc_probe<CAVLC_RUN_BEFORE_K>(b, cavlc_run_before[zeros_left > 6 ? 7 : zeros_left]);
			if (run > zeros_left) {
				b.err = true;
				return 0;
			}
			zeros_left -= run;
		}
		pos -= (int32_t)run + 1;
	}
	return tc;
}

__device__ __forceinline__ uint32_t tab_for_nc(uint32_t nc)
{
	return nc < 2 ? 0 : nc < 4 ? 1 : nc < 8 ? 2 : 3;
}

/*
 * The residual blocks of a macroblock as a bit mask of SLOTS, in the order residual() reads them
 * (src/h264_syntax_slice_data.h:334-419, residual_luma :247-331):
 *   17 * c + 0        Intra16x16 DC of colour component c (c > 0 only with ChromaArrayType 3)
 *   17 * c + 1 + blk  4x4 block blk of component c (AC only for Intra16x16)
 *   17, 18            chroma DC Cb, Cr            (ChromaArrayType 1, 2)
 *   19 + 8 * ic + blk chroma AC block             (ChromaArrayType 1, 2)
 * A lane walks its mask lowest bit first; the k-th coded block of every lane of a warp goes
 * through the same residual_block code together.
 */
__device__ __forceinline__ uint64_t residual_slots(const SliceCtx &s, bool i16, uint32_t cbp_luma,
						    uint32_t cbp_chroma)
{
	uint64_t luma = i16 ? 1u : 0u;
#pragma unroll
	for (uint32_t q = 0; q < 4; q++)
		if (cbp_luma & (1u << q))
			luma |= 0xfull << (1 + 4 * q);
	uint64_t m = luma;
	if (s.cat == 3) {
		m |= luma << 17 | luma << 34;
	} else if (s.cat == 1 || s.cat == 2) {
		const uint64_t blks = s.cat == 1 ? 0xfull : 0xffull;
		if (cbp_chroma & 3)
			m |= 3ull << 17;
		if (cbp_chroma & 2)
			m |= blks << 19 | blks << 27;
	}
	return m;
}

__device__ __forceinline__ void residual_slot(SliceCtx &s, bool i16, uint32_t slot)
{
	uint32_t tab, maxc, field, idx, out;
	if (s.cat == 3 || slot < 17) {
		const uint32_t comp = slot / 17, k = slot - 17 * comp;
		if (k == 0) {
			tab = tab_for_nc(calc_nc(s, comp, 0, false));
			maxc = 16;
			field = H264GPU_F_I16_DC + 3 * comp;
			idx = 0;
			out = comp * 16;
		} else {
			const uint32_t blk = k - 1;
			tab = tab_for_nc(calc_nc(s, comp, blk, false));
			maxc = i16 ? 15 : 16;
			field = (i16 ? H264GPU_F_I16_AC : H264GPU_F_LEVEL4X4) + 3 * comp;
			idx = blk * 16;
			out = comp * 16 + blk;
		}
	} else if (slot < 19) {
		const uint32_t ic = slot - 17;
		tab = s.cat == 1 ? 4 : 5;
		maxc = s.cat == 1 ? 4 : 8;
		field = H264GPU_F_CHROMA_DC;
		idx = ic * 16;
		out = (1 + ic) * 16;
	} else {
		const uint32_t jj = slot - 19, ic = jj >> 3, blk = jj & 7;
		tab = tab_for_nc(calc_nc(s, 1 + ic, blk, true));
		maxc = 15;
		field = H264GPU_F_CHROMA_AC;
		idx = (ic * 16 + blk) * 16;
		out = (1 + ic) * 16 + blk;
	}
	const uint32_t tc = residual_block(s, tab, maxc, field, idx);
	s.nz[out] = (uint8_t)tc;
}

/* B mb_type 4..21: prediction modes of the two partitions (src/h264_slice_data.c:847-866) */
__device__ __forceinline__ void b_part_modes(uint32_t t, uint32_t &m0, uint32_t &m1)
{
	/* pairs in raw mb_type order, two types (16x8, 8x16) per pair */
	const uint32_t p = (t - 4) >> 1;
	/* table: (L0,L0)(L1,L1)(L0,L1)(L1,L0)(L0,Bi)(L1,Bi)(Bi,L0)(Bi,L1)(Bi,Bi) */
	const uint32_t first[9] = {0, 1, 0, 1, 0, 1, 2, 2, 2};
	const uint32_t second[9] = {0, 1, 1, 0, 2, 2, 0, 1, 2};
	m0 = PM_L0 + first[p];
	m1 = PM_L0 + second[p];
}

struct MbInfo {
	uint32_t mb_type, raw_type, num_part;
	uint32_t pm[4];
};

/* h264_read_mb_type: src/h264_slice_data.c:839-969 */
__device__ __forceinline__ bool decode_mb_type(SliceCtx &s, MbInfo &m, uint32_t &cbp_luma,
						uint32_t &cbp_chroma)
{
	uint32_t type = br_ue(s.br);
	m.raw_type = type;
	m.num_part = 0;
	m.pm[0] = m.pm[1] = m.pm[2] = m.pm[3] = 0;
	m.mb_type = MB_UNKNOWN;
	const uint32_t st = s.sp->slice_type;
	bool intra = false;
	if (st == ST_I) {
		intra = true;
	} else if (st == ST_SI) {
		if (type == 0) {
			m.mb_type = MB_SI;
			m.num_part = 1;
			m.pm[0] = PM_I4;
			return true;
		}
		type -= 1;
		intra = true;
	} else if (st == ST_P || st == ST_SP) {
		if (type == 0) {
			m.mb_type = MB_P_16x16;
			m.num_part = 1;
			m.pm[0] = PM_L0;
		} else if (type <= 2) {
			m.mb_type = type == 1 ? MB_P_16x8 : MB_P_8x16;
			m.num_part = 2;
			m.pm[0] = m.pm[1] = PM_L0;
		} else if (type == 3) {
			m.mb_type = MB_P_8x8;
			m.num_part = 4;
		} else if (type == 4) {
			m.mb_type = MB_P_8x8ref0;
			m.num_part = 4;
		} else {
			type -= 5;
			intra = true;
		}
	} else { /* B */
		if (type == 0) {
			m.mb_type = MB_B_Direct_16x16;
			m.num_part = 1;
			m.pm[0] = PM_DIRECT;
		} else if (type <= 3) {
			m.mb_type = MB_B_16x16;
			m.num_part = 1;
			m.pm[0] = type == 1 ? PM_L0 : type == 2 ? PM_L1 : PM_BI;
		} else if (type <= 21) {
			m.mb_type = ((type - 4) & 1) ? MB_B_8x16 : MB_B_16x8;
			m.num_part = 2;
			b_part_modes(type, m.pm[0], m.pm[1]);
		} else if (type == 22) {
			m.mb_type = MB_B_8x8;
			m.num_part = 4;
		} else {
			type -= 23;
			intra = true;
		}
	}
	if (intra) {
		if (type == 0) {
			m.mb_type = MB_I_NxN;
			m.num_part = 1;
			m.pm[0] = PM_I4;
		} else if (type <= 24) {
			m.mb_type = MB_I_16x16;
			m.num_part = 1;
			m.pm[0] = PM_I16;
			hash_add(s, H264GPU_F_I16_PRED_MODE, 0, (type - 1) % 4);
			cbp_luma = type <= 12 ? 0 : 15;
			cbp_chroma = ((type - 1) / 4) % 3;
		} else if (type == 25) {
			m.mb_type = MB_I_PCM;
			m.num_part = 0;
		} else {
			return false;
		}
	}
	return true;
}

/* macroblock_layer: src/h264_syntax_slice_data.h:604-696 */
__device__ __forceinline__ bool macroblock_layer(SliceCtx &s, uint32_t &mb_type_out, uint64_t &slots, bool &i16_out)
{
	BitReader &b = s.br;
	const h264gpu_slice_params &sp = *s.sp;
	MbInfo m;
	uint32_t cbp_luma = 0, cbp_chroma = 0;
	slots = 0;
	i16_out = false;
	if (!decode_mb_type(s, m, cbp_luma, cbp_chroma) || b.err)
		return false;
	mb_type_out = m.mb_type;
	hash_add(s, H264GPU_F_RAW_MB_TYPE, 0, m.raw_type);
	const uint32_t st = sp.slice_type;

	if (m.mb_type == MB_I_PCM) {
		while (!br_aligned(b))
			if (br_get(b, 1) != 0 || b.err)
				return false;
		for (uint32_t i = 0; i < 256; i++)
			hash_add(s, H264GPU_F_PCM_LUMA, i, br_get(b, sp.bit_depth_luma));
		const uint32_t nc = 16 * s.mbw_c4 * s.mbh_c4;
		for (uint32_t ic = 0; ic < 2; ic++)
			for (uint32_t i = 0; i < nc; i++)
				hash_add(s, H264GPU_F_PCM_CHROMA, ic * 256 + i, br_get(b, sp.bit_depth_chroma));
		for (int i = 0; i < 48; i++)
			s.nz[i] = 16;
		return !b.err;
	}

	bool no_sub_lt_8x8 = true;
	bool t8 = false;
	if (m.mb_type != MB_I_NxN && m.pm[0] != PM_I16 && m.num_part == 4) {
		/* sub_mb_pred: src/h264_syntax_slice_data.h:422-503 */
		uint32_t nsub[4], spm[4];
		bool direct[4];
		for (int i = 0; i < 4; i++) {
			uint32_t t = br_ue(b);
			hash_add(s, H264GPU_F_RAW_SUB_MB_TYPE, (uint32_t)i, t);
			direct[i] = false;
			nsub[i] = 0;
			spm[i] = 0;
			if (st == ST_P || st == ST_SP) {
				if (t >= 4)
					return false;
				nsub[i] = t == 0 ? 1 : t == 3 ? 4 : 2;
				spm[i] = PM_L0;
			} else if (st == ST_B) {
				if (t >= 13)
					return false;
				direct[i] = t == 0;
				nsub[i] = (t == 0 || t >= 10) ? 4 : t <= 3 ? 1 : 2;
				spm[i] = t == 0 ? PM_DIRECT
					       : (t == 1 || t == 4 || t == 5 || t == 10) ? PM_L0
					       : (t == 2 || t == 6 || t == 7 || t == 11) ? PM_L1 : PM_BI;
			}
		}
		const uint32_t max0 = sp.num_ref_idx_l0_active_minus1, max1 = sp.num_ref_idx_l1_active_minus1;
		if (max0 > 0 && m.mb_type != MB_P_8x8ref0)
			for (uint32_t i = 0; i < 4; i++)
				if (!direct[i] && spm[i] != PM_L1)
					hash_add(s, H264GPU_F_REF_IDX_L0, i, (uint8_t)br_te(b, max0));
		if (max1 > 0)
			for (uint32_t i = 0; i < 4; i++)
				if (!direct[i] && spm[i] != PM_L0)
					hash_add(s, H264GPU_F_REF_IDX_L1, i, (uint8_t)br_te(b, max1));
		for (uint32_t i = 0; i < 4; i++)
			if (!direct[i] && spm[i] != PM_L1)
				for (uint32_t j = 0; j < nsub[i]; j++)
					for (uint32_t c = 0; c < 2; c++)
						hash_add(s, H264GPU_F_MVD_L0, (i * 4 + j) * 2 + c, (int16_t)br_se(b));
		for (uint32_t i = 0; i < 4; i++)
			if (!direct[i] && spm[i] != PM_L0)
				for (uint32_t j = 0; j < nsub[i]; j++)
					for (uint32_t c = 0; c < 2; c++)
						hash_add(s, H264GPU_F_MVD_L1, (i * 4 + j) * 2 + c, (int16_t)br_se(b));
		for (int i = 0; i < 4; i++) {
			if (!direct[i]) {
				if (nsub[i] > 1)
					no_sub_lt_8x8 = false;
			} else if (!sp.direct_8x8_inference_flag) {
				no_sub_lt_8x8 = false;
			}
		}
	} else {
		if (sp.transform_8x8_mode_flag && m.mb_type == MB_I_NxN) {
			t8 = br_get(b, 1);
			if (t8)
				m.pm[0] = PM_I8;
		}
		/* mb_pred: src/h264_syntax_slice_data.h:506-601 */
		if (m.pm[0] <= PM_I16) {
			if (m.pm[0] == PM_I4) {
				for (uint32_t i = 0; i < 16; i++) {
					if (br_get(b, 1))
						hash_add(s, H264GPU_F_INTRA4X4_PRED_MODE, i, -1);
					else
						hash_add(s, H264GPU_F_INTRA4X4_PRED_MODE, i, br_get(b, 3));
				}
			} else if (m.pm[0] == PM_I8) {
				for (uint32_t i = 0; i < 4; i++) {
					if (br_get(b, 1))
						hash_add(s, H264GPU_F_INTRA8X8_PRED_MODE, i, -1);
					else
						hash_add(s, H264GPU_F_INTRA8X8_PRED_MODE, i, br_get(b, 3));
				}
			}
			if (s.cat == 1 || s.cat == 2)
				hash_add(s, H264GPU_F_INTRA_CHROMA_PRED_MODE, 0, (uint8_t)br_ue(b));
		} else if (m.pm[0] != PM_DIRECT) {
			const uint32_t max0 = sp.num_ref_idx_l0_active_minus1, max1 = sp.num_ref_idx_l1_active_minus1;
			if (max0 > 0)
				for (uint32_t i = 0; i < m.num_part; i++)
					if (m.pm[i] != PM_L1)
						hash_add(s, H264GPU_F_REF_IDX_L0, i, (uint8_t)br_te(b, max0));
			if (max1 > 0)
				for (uint32_t i = 0; i < m.num_part; i++)
					if (m.pm[i] != PM_L0)
						hash_add(s, H264GPU_F_REF_IDX_L1, i, (uint8_t)br_te(b, max1));
			for (uint32_t i = 0; i < m.num_part; i++)
				if (m.pm[i] != PM_L1)
					for (uint32_t c = 0; c < 2; c++)
						hash_add(s, H264GPU_F_MVD_L0, (i * 4) * 2 + c, (int16_t)br_se(b));
			for (uint32_t i = 0; i < m.num_part; i++)
				if (m.pm[i] != PM_L0)
					for (uint32_t c = 0; c < 2; c++)
						hash_add(s, H264GPU_F_MVD_L1, (i * 4) * 2 + c, (int16_t)br_se(b));
		}
	}
	if (b.err)
		return false;

	if (m.pm[0] != PM_I16) {
		/* h264_read_coded_block_pattern: src/h264_slice_data.c:1041-1080 */
		uint32_t code = br_ue(b);
		const bool intra_map = m.mb_type == MB_I_NxN || m.mb_type == MB_I_16x16 || m.mb_type == MB_SI;
		uint32_t cbp;
		if (s.cat == 1 || s.cat == 2) {
			if (code >= 48)
				return false;
			cbp = cavlc_cbp_chroma[code][intra_map ? 0 : 1];
		} else {
			if (code >= 16)
				return false;
			cbp = cavlc_cbp_nochroma[code][intra_map ? 0 : 1];
		}
		hash_add(s, H264GPU_F_CBP, 0, cbp);
		cbp_luma = cbp % 16;
		cbp_chroma = cbp / 16;
		if (cbp_luma > 0 && sp.transform_8x8_mode_flag && m.mb_type != MB_I_NxN && no_sub_lt_8x8 &&
		    (m.mb_type != MB_B_Direct_16x16 || sp.direct_8x8_inference_flag))
			t8 = br_get(b, 1);
	}
	hash_add(s, H264GPU_F_TRANSFORM_8X8, 0, t8 ? 1 : 0);
	hash_add(s, H264GPU_F_CBP_LUMA, 0, cbp_luma);
	hash_add(s, H264GPU_F_CBP_CHROMA, 0, cbp_chroma);

	if (cbp_luma > 0 || cbp_chroma > 0 || m.pm[0] == PM_I16) {
		hash_add(s, H264GPU_F_MB_QP_DELTA, 0, br_se(b));
		i16_out = m.pm[0] == PM_I16;
		slots = residual_slots(s, i16_out, cbp_luma, cbp_chroma);
	}
	return !b.err;
}

/*
 * One slice (src/h264_syntax_slice_data.h:701-787) as three steps per macroblock, so that the
 * lanes of a warp (one slice each) can take every step together:
 *   mb_begin   skip run + macroblock header up to the residual -> the slots to decode
 *   residual_slot x popcount(slots)
 *   mb_end     context ring, record, end-of-slice test
 */
struct SliceRun {
	SliceCtx s;
	h264gpu_mb_record *rec;
	h264_mb_syntax *syn_base; /* full records, same indexing as rec, or NULL */
	uint32_t pic_size, first, cur, count, mb_type;
	int status;
	bool inter, done, i16;
};

__device__ __forceinline__ void syn_open(h264_mb_syntax *o, uint32_t mb_addr, uint32_t mb_type)
{
	uint32_t *w = (uint32_t *)o;
	for (uint32_t i = 0; i < (uint32_t)sizeof(*o) / 4; i++)
		w[i] = 0;
	o->mb_addr = mb_addr;
	o->mb_type = mb_type;
}

__device__ __forceinline__ void slice_begin(SliceRun &r, const uint8_t *stream, uint64_t stream_len,
					     const h264gpu_slice_params &sp, uint8_t *ring, h264gpu_mb_record *rec,
					     h264_mb_syntax *syn = nullptr)
{
	r.status = 0;
	r.count = 0;
	r.done = true;
	r.rec = rec;
	r.syn_base = syn;
	r.s.syn = nullptr;
	r.s.sp = &sp;
	r.s.br.p = stream;
	r.s.br.len = r.s.br.pos = 0;
	r.s.br.nbits = 0;
	r.s.br.cache = 0;
	r.s.br.epbq = 0;
	if (sp.entropy_coding_mode_flag) {
		r.status = H264GPU_SLICE_SKIPPED;
		return;
	}
	if (sp.mbaff_frame_flag || sp.num_slice_groups_minus1 != 0 || sp.pic_width_in_mbs == 0) {
		r.status = -ENOSYS;
		return;
	}
	/* a parameter block that points outside the stream must not be followed */
	if ((uint64_t)sp.nal_off + sp.nal_len > stream_len || sp.data_bit_off >= 8ull * sp.nal_len) {
		r.status = -22; /* -EINVAL */
		return;
	}
	SliceCtx &s = r.s;
	s.W = sp.pic_width_in_mbs;
	s.cat = sp.chroma_array_type;
	s.mbw_c4 = s.cat == 0 ? 0 : (s.cat == 3 ? 4 : 2);
	s.mbh_c4 = s.cat == 0 ? 0 : (s.cat == 1 ? 2 : 4);
	s.ring = ring;
	br_init(s.br, stream + sp.nal_off, sp.nal_len, sp.data_bit_off);
	r.pic_size = (uint32_t)sp.pic_width_in_mbs * sp.pic_height_in_mbs;
	r.first = sp.first_mb_in_slice;
	r.inter = sp.slice_type != ST_I && sp.slice_type != ST_SI;
	r.cur = r.first;
	r.done = false;
}

/* returns the residual slots of the macroblock; r.done set when the slice ends here */
__device__ __forceinline__ uint64_t mb_begin(SliceRun &r)
{
	SliceCtx &s = r.s;
	const h264gpu_slice_params &sp = *s.sp;
	if (r.inter) {
		const uint32_t run = br_ue(s.br);
		if (s.br.err) {
			r.status = -EIO;
			r.done = true;
			return 0;
		}
		for (uint32_t i = 0; i < run; i++) {
			if (r.count >= sp.mb_out_cap || r.cur >= r.pic_size) {
				r.status = -ENOBUFS;
				r.done = true;
				return 0;
			}
			uint32_t *slot = (uint32_t *)ring_slot(s, r.cur);
			for (int k = 0; k < 12; k++)
				slot[k] = 0;
			r.rec[r.count].mb_addr = r.cur;
			r.rec[r.count].mb_type = sp.slice_type == ST_B ? MB_B_SKIP : MB_P_SKIP;
			r.rec[r.count].hash = 0;
			if (CAVLC_FULL && r.syn_base)
				syn_open(r.syn_base + r.count, r.cur, sp.slice_type == ST_B ? MB_B_SKIP : MB_P_SKIP);
			r.count++;
			r.cur++;
		}
		if (run > 0 && !br_more_rbsp_data(s.br)) {
			r.done = true;
			return 0;
		}
	}
	if (r.count >= sp.mb_out_cap || r.cur >= r.pic_size) {
		r.status = -ENOBUFS;
		r.done = true;
		return 0;
	}
	s.cur = r.cur;
	s.availA = r.cur >= r.first + 1 && r.cur % s.W != 0;
	s.availB = r.cur >= r.first + s.W;
	for (int k = 0; k < 48; k++)
		s.nz[k] = 0;
	s.hash = 0;
	s.syn = CAVLC_FULL && r.syn_base ? r.syn_base + r.count : nullptr;
	if (CAVLC_FULL && s.syn)
		syn_open(s.syn, r.cur, 0);
	r.mb_type = MB_UNKNOWN;
	uint64_t slots = 0;
	if (!macroblock_layer(s, r.mb_type, slots, r.i16)) {
		r.status = -EIO;
		r.done = true;
		return 0;
	}
	return slots;
}

__device__ __forceinline__ void mb_end(SliceRun &r)
{
	SliceCtx &s = r.s;
	if (s.br.err) {
		r.status = -EIO;
		r.done = true;
		return;
	}
	uint32_t *slot = (uint32_t *)ring_slot(s, r.cur);
	for (int k = 0; k < 12; k++)
		slot[k] = (uint32_t)s.nz[4 * k] | (uint32_t)s.nz[4 * k + 1] << 8 |
			  (uint32_t)s.nz[4 * k + 2] << 16 | (uint32_t)s.nz[4 * k + 3] << 24;
	r.rec[r.count].mb_addr = r.cur;
	r.rec[r.count].mb_type = r.mb_type;
	r.rec[r.count].hash = s.hash;
	if (CAVLC_FULL && s.syn)
		s.syn->mb_type = r.mb_type;
	r.count++;
	r.cur++;
	if (!br_more_rbsp_data(s.br))
		r.done = true;
}

__device__ __forceinline__ void slice_end(const SliceRun &r, h264gpu_slice_result &res)
{
	res.status = r.status;
	res.mb_count = r.count;
	res.end_bit = (r.status == 0 || r.status == -EIO || r.status == -ENOBUFS) && r.s.br.len ? br_raw_bitpos(r.s.br) : 0;
}

/* the serial form (emulator harness, one slice per call) */
__device__ __forceinline__ void parse_slice(const uint8_t *stream, uint64_t stream_len,
					     const h264gpu_slice_params &sp, uint8_t *ring, h264gpu_mb_record *rec,
					     h264gpu_slice_result &res, h264_mb_syntax *syn = nullptr)
{
	SliceRun r;
	slice_begin(r, stream, stream_len, sp, ring, rec, syn);
	while (!r.done) {
		uint64_t slots = mb_begin(r);
		if (r.done)
			break;
		while (slots && !r.s.br.err) {
			const uint32_t slot = (uint32_t)__ffsll((long long)slots) - 1;
			slots &= slots - 1;
			residual_slot(r.s, r.i16, slot);
		}
		mb_end(r);
	}
	slice_end(r, res);
}

struct CavlcArgs {
	const uint8_t *stream;
	uint64_t stream_len;
	const h264gpu_slice_params *params;
	uint32_t n_slices;
	h264gpu_mb_record *records;
	h264gpu_slice_result *results;
	uint8_t *ring; /* n_slices x ring_stride bytes */
	uint64_t ring_stride;
	uint32_t ring_w; /* widest picture (in MBs) a ring slot row can hold */
	uint32_t lanes_log2; /* log2 of the slices carried by one warp (0..5) */
	h264_mb_syntax *syntax; /* full records (index = record index), or NULL */
};

/*
 * Warp-synchronous: every lane of a warp carries one slice and the warp takes the macroblock
 * steps together.  Left alone, the lanes of a warp never meet again after the first divergent
 * branch of a slice (round 1: 1.00 active threads per instruction, packing slices into lanes
 * made it slower).  Here the warp re-converges at every macroblock (the loop condition is a warp
 * vote), the header paths diverge only by macroblock type, and the residual blocks -- where the
 * bits are -- run in lock-step: iteration k of the slot loop decodes the k-th coded block of
 * every lane through the same code.  Few slices are still spread over more warps
 * (lanes_log2 < 5) so that the SMs have warps to switch between.
 */
__global__ void __launch_bounds__(128) cavlc_parse_kernel(const CavlcArgs a)
{
	const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t step = 32u >> a.lanes_log2; /* active lanes are multiples of this */
	const uint32_t i = (gwarp << a.lanes_log2) + lane / step;
	const bool mine = !(lane & (step - 1)) && i < a.n_slices;
	SliceRun r;
	r.done = true;
	r.status = 0;
	r.count = 0;
	r.s.br.len = 0;
	if (mine) {
		const h264gpu_slice_params &sp = a.params[i];
		if (sp.pic_width_in_mbs > a.ring_w)
			r.status = -7; /* -E2BIG */
		else
			slice_begin(r, a.stream, a.stream_len, sp, a.ring + (uint64_t)i * a.ring_stride,
				    a.records + sp.mb_out_off, a.syntax ? a.syntax + sp.mb_out_off : nullptr);
	}
	while (__any_sync(FULL_MASK, !r.done)) {
		uint64_t slots = 0;
		if (!r.done)
			slots = mb_begin(r);
		while (slots && !r.s.br.err) {
			const uint32_t slot = (uint32_t)__ffsll((long long)slots) - 1;
			slots &= slots - 1;
			residual_slot(r.s, r.i16, slot);
		}
		if (!r.done)
			mb_end(r);
	}
	if (mine) {
		h264gpu_slice_result res;
		slice_end(r, res);
		a.results[i] = res;
	}
}

} /* namespace CAVLC_NS */

#undef CAVLC_NS
#undef CAVLC_FULL
