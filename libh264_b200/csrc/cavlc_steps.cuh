/*
 * cavlc_steps.cuh — K4, second generation: slice-parallel CAVLC macroblock syntax parse as a
 * per-lane state machine.  One independent slice per GPU lane; ONE syntax element per lane per
 * step; all lanes of a warp take the step together.
 *
 * Why: a slice parse is serial, so the parallelism is across slices, and the first generation
 * (cavlc_parse.cuh) ran every lane through its own copy of the call tree: lanes diverged at the
 * first branch and a warp issued for ~3 of its 32 lanes (ncu: 3.06 active threads per
 * instruction with 32 slices per warp; 1.0 in round 1).  Here the control flow of a slice is DATA
 * (lane.state), the code is one loop body, and every iteration has a long common part -- refill
 * of the bit cache, Exp-Golomb decode, bit consumption, checksum update -- around a short
 * per-state switch.  The whole reader state sits in registers (the first generation kept it in
 * local memory behind __noinline__ calls); the macroblock's coefficient counts and the levels of
 * the block being decoded sit in shared memory (dynamic index), the row of counts above in a
 * 16-byte-per-macroblock ring in global memory.
 *
 * Reference behaviour reproduced (Parrot-Developers/libh264): frame and field pictures, MBAFF
 * frames (macroblock pairs, 6.4.12.2 neighbours: mbaff_nb.h), any number of slice groups:
 *   slice_data loop, skip runs, end test   src/h264_syntax_slice_data.h:701-787
 *   macroblock_layer                       src/h264_syntax_slice_data.h:604-696
 *   mb_pred / sub_mb_pred                  src/h264_syntax_slice_data.h:422-601
 *   residual, residual_luma, residual_block src/h264_syntax_slice_data.h:103-419
 *   mb_type / sub_mb_type / cbp mapping    src/h264_slice_data.c:839-1080
 *   coeff_token (nC from the slice-local neighbours), total_zeros, run_before
 *                                          src/h264_slice_data.c:1239-1416
 *   neighbouring macroblocks / 4x4 blocks  src/h264_macroblock.c:84-108,306-433
 *   next macroblock of the slice group     src/h264_fmo.c:217-235,308-319
 *   bit reader with EPB skip, ue/se/te     include/h264/h264_bitstream.h:168-304,
 *                                          src/h264_bitstream.c:190-208
 *   h264_bs_more_rbsp_data (incl. its one-trailing-zero-byte tolerance)
 *                                          src/h264_bitstream.c:325-355
 */
#ifndef CAVLC_STEPS_CUH
#define CAVLC_STEPS_CUH

#include "gpu_compat.h"
#include "h264gpu_slice.h"
#include "h264gpu_mb_syntax.h"
#include "mbaff_nb.h"

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

namespace cavlc2 {

/* enum h264_mb_type values (include/h264/h264_types.h:70-90) */
enum {
	MB_UNKNOWN = 0, MB_I_NxN, MB_I_16x16, MB_I_PCM, MB_SI, MB_P_16x16, MB_P_16x8, MB_P_8x16,
	MB_P_8x8, MB_P_8x8ref0, MB_P_SKIP, MB_B_Direct_16x16, MB_B_16x16, MB_B_16x8, MB_B_8x16,
	MB_B_8x8, MB_B_SKIP,
};
enum { ST_P = 0, ST_B = 1, ST_I = 2, ST_SP = 3, ST_SI = 4 };

/* what a lane reads next */
enum {
	S_SKIP_RUN = 0, /* mb_skip_run                                   ue */
	S_SKIP_EMIT,    /* one skipped macroblock                        -  */
	S_MB_TYPE,      /* mb_type (+ transform_size_8x8_flag of I_NxN)  ue */
	S_SUB_TYPE,     /* sub_mb_type[k]                                ue */
	S_REF,          /* ref_idx_l0 / l1 of the lowest bit of ref_mask te */
	S_MVD,          /* mvd_l0 / l1 component, lowest bit of mvd_mask se */
	S_PRED,         /* prev_intra pred flag + rem mode, block k      1 or 4 bits */
	S_CHROMA_PRED,  /* intra_chroma_pred_mode                        ue */
	S_CBP,          /* coded_block_pattern (+ transform_size_8x8_flag) ue */
	S_QPD,          /* mb_qp_delta                                   se */
	S_TOKEN,        /* coeff_token + trailing ones signs of the lowest slot */
	S_LEVEL,        /* one level (an escape takes two steps)         */
	S_ZEROS,        /* total_zeros                                   */
	S_RUN,          /* coefficients into place + one run_before      */
	S_MB_END,       /* counts row, record, end-of-slice test         -  */
	S_PCM_ALIGN,    /* pcm_alignment_zero_bits                       */
	S_PCM,          /* one pcm sample                                */
	S_MB_FIELD,     /* MBAFF: mb_field_decoding_flag (top macroblock of a pair, or the bottom one
			   after a skipped top)                          1 bit */
	S_DONE,
};
#define CAVLC2_UE_STATES                                                                               \
	((1u << S_SKIP_RUN) | (1u << S_MB_TYPE) | (1u << S_SUB_TYPE) | (1u << S_MVD) | (1u << S_CHROMA_PRED) | \
	 (1u << S_CBP) | (1u << S_QPD))

/* lane flags */
enum {
	F_INTER = 1u << 0,   /* P / SP / B slice: mb_skip_run is read */
	F_B = 1u << 1,
	F_T8MODE = 1u << 2,  /* transform_8x8_mode_flag */
	F_D8INF = 1u << 3,   /* direct_8x8_inference_flag */
	F_AVAIL_A = 1u << 4,
	F_AVAIL_B = 1u << 5,
	F_I16 = 1u << 6,
	F_T8 = 1u << 7,
	F_NO_SUB_LT8 = 1u << 8,
	F_PREV_ZERO = 1u << 9, /* the macroblock before (prev_addr) was skipped: all counts 0 */
	F_NXN = 1u << 10,      /* I_NxN (the intra cbp table, no second transform flag) */
	F_INTRA_CBP = 1u << 11,
	F_DIRECT16 = 1u << 12,
	F_WHICH = 1u << 13,    /* which shared count buffer is the current macroblock's */
	F_PRED8 = 1u << 14,    /* S_PRED reads intra 8x8 modes */
	F_MBAFF = 1u << 15,    /* MbaffFrameFlag: macroblock pairs */
	F_CUR_FIELD = 1u << 16,   /* mb_field_decoding_flag of the current pair */
	F_TOP_SKIPPED = 1u << 17, /* the top macroblock of the current pair was skipped */
	F_A_FIELD = 1u << 18,     /* field flag of the pair before (the pair to the left when F_AVAIL_A) */
	F_B_FIELD = 1u << 19,     /* field flag of the pair above */
	F_PREV_SKIPPED = 1u << 20, /* the mb_skip_run read last was not 0 (prevMbSkipped, 7.3.4) */
	F_FIELD_PIC = 1u << 21,    /* field_pic_flag: every macroblock is a field macroblock */
};

/* MBAFF slice in an instantiation that carries the macroblock-pair code */
#define CAVLC2_MBAFF(l) ((l).can_mbaff && ((l).flags & F_MBAFF) != 0)

#ifdef H264_EMU
#define CAVLC2_STRIDE 1u
#else
#define CAVLC2_STRIDE 128u /* threads per block: shared words of a lane are interleaved by thread */
#endif
#define CAVLC2_SM_WORDS 56u /* per lane: 4 x 12 words of counts (this and the previous macroblock or pair), 8 words of levels */
#define CAVLC2_RING_SLOT 32u /* bytes per macroblock (or pair) in the ring of bottom-row counts */

struct Lane {
	/* bit reader over the escaped NAL */
	const uint8_t *p; /* NAL start (header byte) */
	uint32_t len;
	uint32_t pos;     /* raw offset of the next byte to load */
	uint64_t cache;   /* unread bits, MSB first */
	int nbits;
	uint32_t zeros;   /* run of raw zero bytes just before pos */
	uint32_t epbq;    /* bit k: an EPB was skipped right before the k-th most recent byte */
	uint32_t pend;    /* first half of a long code already consumed: leading zeros (ue) / prefix (level) */
	/* slice */
	const h264gpu_slice_params *sp;
	const uint8_t *gmap; /* macroblock -> slice group, or NULL (one group) */
	uint8_t *ring;       /* (W + 1) x 16 bytes: bottom-row counts of the last W + 1 macroblocks */
	h264gpu_mb_record *rec;
	h264_mb_syntax *syn; /* full records (same indexing as rec) or NULL */
	uint32_t *sm;        /* this lane's shared words (stride CAVLC2_STRIDE) */
	uint32_t W, pic_size, first, cap, cat, max0, max1;
	uint32_t flags, state, cur, count, prev_addr;
	bool can_mbaff;      /* constant of the kernel instantiation: false removes the macroblock-pair code */
	int status;
	/* macroblock */
	uint64_t hash, slots, mvd_mask;
	uint32_t ref_mask, k, kend, mb_type, cbp;
	uint32_t top0, top1, top2; /* bottom-row counts of the macroblock above (MBAFF: top macroblock of the pair above) */
	uint32_t tb0, tb1, tb2;    /* MBAFF: bottom-row counts of the bottom macroblock of the pair above */
	/* residual block */
	uint32_t tc, ci, suffix_len, zeros_left, field, idx_base, out, maxc;
	int32_t cpos;
};

/* ---- shared memory of a lane ------------------------------------------------------------ */
__device__ __forceinline__ uint8_t *nz_ptr(const Lane &l, uint32_t buf, uint32_t i)
{
	return (uint8_t *)(l.sm + (buf * 12u + (i >> 2)) * CAVLC2_STRIDE) + (i & 3u);
}
__device__ __forceinline__ int16_t *lev_ptr(const Lane &l, uint32_t i)
{
	return (int16_t *)(l.sm + (48u + (i >> 1)) * CAVLC2_STRIDE) + (i & 1u);
}
/* count buffers: 2 * which + (bottom macroblock of an MBAFF pair); the other two hold the
 * macroblock (pair) before */
__device__ __forceinline__ uint32_t cur_buf(const Lane &l)
{
	return ((l.flags & F_WHICH) ? 2u : 0u) | (CAVLC2_MBAFF(l) ? (l.cur & 1u) : 0u);
}
__device__ __forceinline__ uint32_t prev_buf(const Lane &l) { return (l.flags & F_WHICH) ? 0u : 2u; }

/* ---- bit reader ---------------------------------------------------------------------------- */
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

/* at least 32 unread bits afterwards, or every byte of the NAL loaded */
__device__ __forceinline__ void refill(Lane &b)
{
	/* word at a time: a word without a 03 byte holds no emulation prevention byte, whatever came
	 * before it (8 bytes of the NAL must be left so that the aligned loads stay inside it) */
	if (b.pos + 8 <= b.len) {
		const uint32_t w = load_be32(b.p + b.pos);
		const uint32_t x = w ^ 0x03030303u;
		if (!((x - 0x01010101u) & ~x & 0x80808080u)) {
			b.cache |= (uint64_t)w << (32 - b.nbits);
			b.nbits += 32;
			b.pos += 4;
			b.epbq <<= 4;
			b.zeros = (w & 0xffffu) == 0 ? 2u : ((w & 0xffu) == 0 ? 1u : 0u);
			return;
		}
	}
	/* byte path: a 03 byte in the word, or the tail of the NAL */
	while (b.nbits <= 56 && b.pos < b.len) {
		uint32_t c = b.p[b.pos];
		uint32_t skipped = 0;
		if (b.zeros >= 2 && c == 3) {
			if (b.pos + 1 >= b.len)
				return; /* the reference fails the fetch here */
			b.pos++;
			c = b.p[b.pos];
			skipped = 1;
			b.zeros = 0;
		}
		b.zeros = c == 0 ? b.zeros + 1 : 0;
		b.pos++;
		b.cache |= (uint64_t)c << (56 - b.nbits);
		b.nbits += 8;
		b.epbq = (b.epbq << 1) | skipped;
	}
}

/* raw bit position of the next unread bit (for the result record) */
__device__ __forceinline__ uint64_t raw_bitpos(const Lane &b)
{
	const int whole = b.nbits >> 3;
	const uint32_t epbs = (uint32_t)__popc(b.epbq & ((1u << whole) - 1));
	const uint32_t next_byte = b.pos - (uint32_t)whole - epbs; /* first raw byte not begun */
	return (uint64_t)next_byte * 8 - (uint32_t)(b.nbits & 7);
}

/* h264_bs_more_rbsp_data, src/h264_bitstream.c:325-355 (the cache holds 32 bits or the whole rest) */
__device__ __forceinline__ bool more_rbsp_data(const Lane &b)
{
	if (b.nbits == 0)
		return false;
	const int k = (b.nbits & 7) ? (b.nbits & 7) : 8;
	const uint32_t v = (uint32_t)(b.cache >> (64 - k));
	if (v != (1u << (k - 1)))
		return true;
	/* stop bit pattern: what lies after that byte decides */
	const int whole = (b.nbits - k) >> 3;
	const uint32_t epbs = (uint32_t)__popc(b.epbq & ((1u << whole) - 1));
	const uint32_t off = b.pos - (uint32_t)whole - epbs;
	if (off >= b.len)
		return false;
	return off + 1 < b.len || b.p[off] != 0;
}

/* one probe of a (leading zeros, K bits after the first one) table */
template <int K>
__device__ __forceinline__ uint32_t vlc_entry(uint32_t top, const uint16_t *lut)
{
	int lz = __clz((int)top);
	if (lz > 15)
		lz = 15;
	const uint32_t rest = K ? ((top << (lz + 1)) >> (32 - K)) : 0;
	return lut[(lz << K) | rest];
}

/* ---- checksum / full record ------------------------------------------------------------- */
__device__ __forceinline__ uint64_t hash_term(uint32_t field, uint32_t idx, int64_t v)
{
	/* device-side twin of h264gpu_mb_hash_term (include/h264gpu_slice.h) */
	const uint64_t key = ((uint64_t)field << 16) | idx;
	const uint64_t w = ((key + 1) * 0x9E3779B97F4A7C15ull) | 1ull;
	/* every value is a 32-bit quantity: (int64)v * w from two 32-bit multiplies */
	const uint32_t vl = (uint32_t)v, wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
	const uint32_t hi = vl * wh - (v < 0 ? wl : 0u);
	return (uint64_t)vl * wl + ((uint64_t)hi << 32);
}

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

__device__ __forceinline__ void syn_open(h264_mb_syntax *o, uint32_t mb_addr, uint32_t mb_type)
{
	uint32_t *w = (uint32_t *)o;
	for (uint32_t i = 0; i < (uint32_t)sizeof(*o) / 4; i++)
		w[i] = 0;
	o->mb_addr = mb_addr;
	o->mb_type = mb_type;
}

template <bool FULL>
__device__ __forceinline__ void hash_add(Lane &l, uint32_t field, uint32_t idx, int64_t v)
{
	if (v != 0) {
		if (FULL && l.syn != nullptr)
			syn_store(l.syn + l.count, field, idx, v);
		l.hash += hash_term(field, idx, v);
	}
}

/* ---- slice group map (8.2.2): next macroblock of the slice, neighbour availability --------- */
__device__ __forceinline__ uint32_t next_mb_addr(const Lane &l, uint32_t mb)
{
	if (l.gmap == nullptr)
		return mb + 1;
	const uint8_t g = l.gmap[mb];
	uint32_t i = mb + 1;
	while (i < l.pic_size && l.gmap[i] != g)
		i++;
	return i;
}

/* x, y of a luma-like 4x4 block in block units from its z-order index, and back */
__device__ __forceinline__ uint32_t blk_x(uint32_t blk) { return (blk & 1) | ((blk >> 1) & 2); }
__device__ __forceinline__ uint32_t blk_y(uint32_t blk) { return ((blk >> 1) & 1) | ((blk >> 2) & 2); }
__device__ __forceinline__ uint32_t blk_idx(uint32_t x, uint32_t y)
{
	return (x & 1) | ((y & 1) << 1) | ((x & 2) << 1) | ((y & 2) << 2);
}

/* 6.4.11.4 / 6.4.11.5 + 9.2.1: nC of a block from the counts left of and above it */
__device__ __forceinline__ uint32_t calc_nc(const Lane &l, uint32_t comp, uint32_t blk, bool chroma_ac)
{
	const uint32_t cb = cur_buf(l), pb = prev_buf(l);
	uint32_t nA = 0, nB = 0;
	bool aA, aB;
	uint32_t x, y, left_in, up_in;
	if (!chroma_ac) {
		x = blk_x(blk);
		y = blk_y(blk);
		left_in = blk_idx(x - 1, y);
		up_in = blk_idx(x, y - 1);
	} else {
		x = blk & 1;
		y = blk >> 1;
		left_in = blk - 1;
		up_in = blk - 2;
	}
	const bool mbaff = CAVLC2_MBAFF(l);
	if (x > 0) {
		aA = true;
		nA = *nz_ptr(l, cb, comp * 16 + left_in);
	} else if (!mbaff) {
		if ((aA = (l.flags & F_AVAIL_A) != 0))
			nA = (l.flags & F_PREV_ZERO) ? 0u : *nz_ptr(l, pb, comp * 16 + (chroma_ac ? 2 * y + 1 : blk_idx(3, y)));
	} else {
		/* 6.4.12.2: which macroblock of the pair to the left, and which of its rows */
		const int maxH = chroma_ac ? (l.cat == 1 ? 8 : 16) : 16;
		int yM = 0;
		const int w = mbaff_left((l.flags & F_CUR_FIELD) != 0, (int)(l.cur & 1), (l.flags & F_AVAIL_A) != 0,
					 (l.flags & F_A_FIELD) != 0, (int)(4 * y), maxH, &yM);
		if ((aA = w != MBAFF_NB_NONE)) {
			const uint32_t row = (uint32_t)yM >> 2;
			nA = *nz_ptr(l, pb | (w == MBAFF_NB_A_BOT ? 1u : 0u), comp * 16 + (chroma_ac ? 2 * row + 1 : blk_idx(3, row)));
		}
	}
	if (y > 0) {
		aB = true;
		nB = *nz_ptr(l, cb, comp * 16 + up_in);
	} else if (!mbaff) {
		if ((aB = (l.flags & F_AVAIL_B) != 0))
			nB = ((comp == 0 ? l.top0 : comp == 1 ? l.top1 : l.top2) >> (8 * x)) & 0xffu;
	} else {
		const int w = mbaff_up((l.flags & F_CUR_FIELD) != 0, (int)(l.cur & 1), (l.flags & F_AVAIL_B) != 0,
				       (l.flags & F_B_FIELD) != 0);
		aB = w != MBAFF_NB_NONE;
		if (w == MBAFF_NB_CUR_TOP) {
			const uint32_t bot = chroma_ac ? (l.cat == 1 ? 2u : 6u) + x : blk_idx(x, 3);
			nB = *nz_ptr(l, cb & 2u, comp * 16 + bot);
		} else if (w == MBAFF_NB_B_TOP) {
			nB = ((comp == 0 ? l.top0 : comp == 1 ? l.top1 : l.top2) >> (8 * x)) & 0xffu;
		} else if (w == MBAFF_NB_B_BOT) {
			nB = ((comp == 0 ? l.tb0 : comp == 1 ? l.tb1 : l.tb2) >> (8 * x)) & 0xffu;
		}
	}
	if (aA && aB)
		return (nA + nB + 1) >> 1;
	return aA ? nA : aB ? nB : 0;
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
 */
__device__ __forceinline__ uint64_t residual_slots(uint32_t cat, bool i16, uint32_t cbp_luma, uint32_t cbp_chroma)
{
	uint64_t luma = i16 ? 1u : 0u;
#pragma unroll
	for (uint32_t q = 0; q < 4; q++)
		if (cbp_luma & (1u << q))
			luma |= 0xfull << (1 + 4 * q);
	uint64_t m = luma;
	if (cat == 3) {
		m |= luma << 17 | luma << 34;
	} else if (cat == 1 || cat == 2) {
		const uint64_t blks = cat == 1 ? 0xfull : 0xffull;
		if (cbp_chroma & 3)
			m |= 3ull << 17;
		if (cbp_chroma & 2)
			m |= blks << 19 | blks << 27;
	}
	return m;
}

/* the bottom-row counts of the current macroblock, one word per colour component */
__device__ __forceinline__ void bottom_rows(const Lane &l, uint32_t w[3])
{
	const uint32_t cb = cur_buf(l);
#define NZ(i) ((uint32_t)*nz_ptr(l, cb, (i)))
	w[0] = NZ(10) | NZ(11) << 8 | NZ(14) << 16 | NZ(15) << 24;
	if (l.cat == 3) {
		w[1] = NZ(26) | NZ(27) << 8 | NZ(30) << 16 | NZ(31) << 24;
		w[2] = NZ(42) | NZ(43) << 8 | NZ(46) << 16 | NZ(47) << 24;
	} else if (l.cat == 1) {
		w[1] = NZ(18) | NZ(19) << 8;
		w[2] = NZ(34) | NZ(35) << 8;
	} else if (l.cat == 2) {
		w[1] = NZ(22) | NZ(23) << 8;
		w[2] = NZ(38) | NZ(39) << 8;
	} else {
		w[1] = w[2] = 0;
	}
#undef NZ
}

__device__ __forceinline__ uint32_t *ring_slot(const Lane &l, uint32_t mb)
{
	return (uint32_t *)(l.ring + (size_t)(mb % (l.W + 1)) * CAVLC2_RING_SLOT);
}

/* ---- MBAFF: macroblock pairs ----------------------------------------------------------------- */
/* the top macroblock of a pair begins (coded or skipped): neighbouring pairs inside the slice
 * (6.4.10; the reference: h264_compute_neighbouring_macroblocks, src/h264_macroblock.c:327-336) */
__device__ __forceinline__ void pair_begin(Lane &l)
{
	const uint32_t half = l.cur >> 1;
	const bool aA = half >= l.first + 1 && half % l.W != 0;
	const bool aB = half >= l.first + l.W;
	l.flags &= ~(F_AVAIL_A | F_AVAIL_B | F_B_FIELD | F_CUR_FIELD | F_TOP_SKIPPED);
	l.flags |= (aA ? F_AVAIL_A : 0u) | (aB ? F_AVAIL_B : 0u);
	if (aB) {
		const uint32_t *up = ring_slot(l, half - l.W);
		l.top0 = up[0];
		l.top1 = up[1];
		l.top2 = up[2];
		if (up[3])
			l.flags |= F_B_FIELD;
		l.tb0 = up[4];
		l.tb1 = up[5];
		l.tb2 = up[6];
	}
}

/* the bottom macroblock of a pair is done: the pair becomes the pair before */
__device__ __forceinline__ void pair_end(Lane &l)
{
	const bool field = (l.flags & F_CUR_FIELD) != 0;
	ring_slot(l, l.cur >> 1)[3] = field ? 1u : 0u;
	l.flags = ((l.flags & ~F_A_FIELD) | (field ? F_A_FIELD : 0u)) ^ F_WHICH;
}

/* where a lane goes for its next coded macroblock */
__device__ __forceinline__ uint32_t mb_start_state(const Lane &l)
{
	if (CAVLC2_MBAFF(l) && (!(l.cur & 1u) || (l.flags & F_PREV_SKIPPED)))
		return S_MB_FIELD;
	return S_MB_TYPE;
}

/* ---- slice begin / end ------------------------------------------------------------------- */
__device__ __forceinline__ void slice_begin(Lane &l, const uint8_t *stream, uint64_t stream_len,
					     const h264gpu_slice_params &sp, uint8_t *ring, h264gpu_mb_record *rec,
					     h264_mb_syntax *syn, const uint8_t *group_maps, uint32_t ring_w)
{
	l.status = 0;
	l.count = 0;
	l.state = S_DONE;
	l.sp = &sp;
	l.rec = rec;
	l.syn = syn;
	l.p = stream;
	l.len = l.pos = 0;
	l.nbits = 0;
	l.cache = 0;
	l.epbq = 0;
	l.pend = 0;
	l.zeros = 0;
	if (sp.entropy_coding_mode_flag) {
		l.status = H264GPU_SLICE_SKIPPED;
		return;
	}
	if (sp.pic_width_in_mbs == 0 || (sp.num_slice_groups_minus1 != 0 && (group_maps == nullptr || sp.mbaff_frame_flag))) {
		l.status = -ENOSYS;
		return;
	}
	if (sp.pic_width_in_mbs > ring_w) {
		l.status = -7; /* -E2BIG */
		return;
	}
	/* a parameter block that points outside the stream must not be followed */
	if ((uint64_t)sp.nal_off + sp.nal_len > stream_len || sp.data_bit_off >= 8ull * sp.nal_len) {
		l.status = -22; /* -EINVAL */
		return;
	}
	l.W = sp.pic_width_in_mbs;
	l.cat = sp.chroma_array_type;
	l.ring = ring;
	l.gmap = sp.num_slice_groups_minus1 != 0 ? group_maps + sp.row_state_off : nullptr;
	l.pic_size = (uint32_t)sp.pic_width_in_mbs * sp.pic_height_in_mbs;
	l.first = sp.first_mb_in_slice;
	l.cap = sp.mb_out_cap;
	l.max0 = sp.num_ref_idx_l0_active_minus1;
	l.max1 = sp.num_ref_idx_l1_active_minus1;
	const bool inter = sp.slice_type != ST_I && sp.slice_type != ST_SI;
	l.flags = (inter ? F_INTER : 0u) | (sp.slice_type == ST_B ? F_B : 0u) |
		  (sp.transform_8x8_mode_flag ? F_T8MODE : 0u) | (sp.direct_8x8_inference_flag ? F_D8INF : 0u) |
		  (sp.mbaff_frame_flag ? F_MBAFF : 0u) | (sp.field_pic_flag ? F_FIELD_PIC : 0u);
	l.cur = sp.mbaff_frame_flag ? 2 * l.first : l.first;
	l.prev_addr = 0xffffffffu;
	/* position the reader at raw bit offset data_bit_off of the NAL: if the offset is byte
	 * aligned the reference has not fetched that byte yet */
	const uint8_t *nal = stream + sp.nal_off;
	l.p = nal;
	l.len = sp.nal_len;
	const uint32_t byte = sp.data_bit_off >> 3;
	l.pos = byte;
	if (byte >= 1 && nal[byte - 1] == 0)
		l.zeros = (byte >= 2 && nal[byte - 2] == 0) ? 2 : 1;
	refill(l);
	const uint32_t drop = sp.data_bit_off & 7;
	if (drop) {
		l.cache <<= drop;
		l.nbits -= (int)drop;
	}
	l.state = inter ? (uint32_t)S_SKIP_RUN : mb_start_state(l);
}

__device__ __forceinline__ void slice_end(const Lane &l, h264gpu_slice_result &res)
{
	res.status = l.status;
	res.mb_count = l.count;
	res.end_bit = (l.status == 0 || l.status == -EIO || l.status == -ENOBUFS) && l.len ? raw_bitpos(l) : 0;
}

__device__ __forceinline__ void fail(Lane &l, int status)
{
	l.status = status;
	l.state = S_DONE;
}

/* where a slot's block lives: table, size, checksum field / index, count index */
__device__ __forceinline__ uint32_t slot_setup(Lane &l, uint32_t slot)
{
	const bool i16 = (l.flags & F_I16) != 0;
	uint32_t tab;
	if (l.cat == 3 || slot < 17) {
		const uint32_t comp = slot / 17, k = slot - 17 * comp;
		if (k == 0) {
			tab = tab_for_nc(calc_nc(l, comp, 0, false));
			l.maxc = 16;
			l.field = H264GPU_F_I16_DC + 3 * comp;
			l.idx_base = 0;
			l.out = comp * 16;
		} else {
			const uint32_t blk = k - 1;
			tab = tab_for_nc(calc_nc(l, comp, blk, false));
			l.maxc = i16 ? 15 : 16;
			l.field = (i16 ? H264GPU_F_I16_AC : H264GPU_F_LEVEL4X4) + 3 * comp;
			l.idx_base = blk * 16;
			l.out = comp * 16 + blk;
		}
	} else if (slot < 19) {
		const uint32_t ic = slot - 17;
		tab = l.cat == 1 ? 4 : 5;
		l.maxc = l.cat == 1 ? 4 : 8;
		l.field = H264GPU_F_CHROMA_DC;
		l.idx_base = ic * 16;
		l.out = (1 + ic) * 16;
	} else {
		const uint32_t jj = slot - 19, ic = jj >> 3, blk = jj & 7;
		tab = tab_for_nc(calc_nc(l, 1 + ic, blk, true));
		l.maxc = 15;
		l.field = H264GPU_F_CHROMA_AC;
		l.idx_base = (ic * 16 + blk) * 16;
		l.out = (1 + ic) * 16 + blk;
	}
	return tab;
}

/* the block of the lowest slot is complete */
__device__ __forceinline__ void block_done(Lane &l, uint32_t tc)
{
	*nz_ptr(l, cur_buf(l), l.out) = (uint8_t)tc;
	l.slots &= l.slots - 1;
	l.state = l.slots ? S_TOKEN : S_MB_END;
}

#define CAVLC2_IS_RES(s) ((s) >= S_TOKEN && (s) <= S_RUN)

/*
 * One step of a lane inside a residual block (states S_TOKEN .. S_RUN): where the bits are.  The
 * three table-coded elements (coeff_token, total_zeros, run_before) share ONE probe -- each
 * state only picks its table -- so lanes in different residual states diverge for a few
 * instructions around a common load; levels are arithmetic (9.2.2.1).
 */
template <bool FULL>
__device__ __forceinline__ void res_step1(Lane &l)
{
	if (l.nbits < 32)
		refill(l);
	const uint32_t top = (uint32_t)(l.cache >> 32);
	const uint32_t state = l.state;
	uint32_t n = 0;          /* bits this step consumes */
	const uint16_t *lut = nullptr;
	uint32_t K = 0, e = 0;
	bool probe = false;

	/* which table */
	if (state == S_TOKEN) {
		const uint32_t slot = (uint32_t)__ffsll((long long)l.slots) - 1;
		const uint32_t tab = slot_setup(l, slot);
		probe = true;
		if (tab == 3) { /* nC >= 8: 6-bit fixed length code */
			const uint32_t f = cavlc_coeff_token_flc[top >> 26];
			e = f ? ((f & 0x7fu) << 8 | 6u) : 0u;
		} else {
			lut = cavlc_coeff_token[tab <= 2 ? tab : tab - 1];
			K = CAVLC_COEFF_TOKEN_K;
		}
	} else if (state == S_ZEROS) {
		probe = true;
		if (l.maxc == 4) {
			lut = cavlc_total_zeros_2x2[l.tc];
			K = CAVLC_TOTAL_ZEROS_2X2_K;
		} else if (l.maxc == 8) {
			lut = cavlc_total_zeros_2x4[l.tc];
			K = CAVLC_TOTAL_ZEROS_2X4_K;
		} else {
			lut = cavlc_total_zeros_4x4[l.tc];
			K = CAVLC_TOTAL_ZEROS_4X4_K;
		}
	} else if (state == S_RUN) {
		/* coefficient ci sits run_before(ci) zeros above coefficient ci + 1: put coefficients in
		 * place until a run_before has to be read (one per step) */
		for (;;) {
			if (l.cpos < 0 || l.cpos >= (int32_t)l.maxc) {
				fail(l, -EIO);
				return;
			}
			hash_add<FULL>(l, l.field, l.idx_base + (uint32_t)l.cpos, *lev_ptr(l, l.ci));
			l.ci++;
			if (l.ci == l.tc) {
				block_done(l, l.tc);
				break;
			}
			if (l.zeros_left > 0) {
				probe = true;
				lut = cavlc_run_before[l.zeros_left > 6 ? 7 : l.zeros_left];
				K = CAVLC_RUN_BEFORE_K;
				break;
			}
			l.cpos -= 1;
		}
	}

	/* the probe: leading zeros (at most 15) and the K bits after the first one */
	if (lut != nullptr) {
		int lz = __clz((int)top);
		if (lz > 15)
			lz = 15;
		const uint32_t rest = ((top << (lz + 1)) >> 1) >> (31 - K);
		e = lut[((uint32_t)lz << K) | rest];
	}
	if (probe) {
		if (e == 0) {
			fail(l, -EIO);
			return;
		}
		n = e & 0xff;
	}
	const uint32_t v = e >> 8;

	if (state == S_TOKEN) {
		const uint32_t t1 = (v >> 5) & 3, tc = v & 0x1f;
		if (tc == 0) {
			block_done(l, 0);
		} else if (tc > l.maxc) {
			l.cache <<= n;
			l.nbits -= (int)n;
			fail(l, -EIO);
			return;
		} else {
			/* trailing ones: their sign bits follow the token (at most 16 + 3 bits together) */
			for (uint32_t i = 0; i < t1; i++)
				*lev_ptr(l, i) = (int16_t)(1 - 2 * (int)((top >> (31 - n - i)) & 1u));
			n += t1;
			l.tc = tc;
			l.ci = t1;
			l.suffix_len = (tc > 10 && t1 < 3) ? 1 : 0;
			l.k = t1; /* the level at this index gets the +2 when fewer than 3 trailing ones */
			l.zeros_left = 0;
			l.cpos = (int32_t)tc - 1;
			if (t1 == tc) {
				l.ci = 0;
				l.state = tc < l.maxc ? S_ZEROS : S_RUN;
			} else {
				l.state = S_LEVEL;
			}
		}
	} else if (state == S_ZEROS) {
		l.zeros_left = v;
		l.cpos = (int32_t)(l.tc + v) - 1; /* index of the highest coefficient */
		l.state = S_RUN;
	} else if (state == S_RUN) {
		if (probe) {
			if (v > l.zeros_left) {
				l.cache <<= n;
				l.nbits -= (int)n;
				fail(l, -EIO);
				return;
			}
			l.zeros_left -= v;
			l.cpos -= (int32_t)v + 1;
		}
	} else {
		/* S_LEVEL: 9.2.2.1, src/h264_syntax_slice_data.h:147-199 */
		uint32_t prefix, suffix_off;
		if (l.pend) {
			prefix = l.pend - 1;
			l.pend = 0;
			suffix_off = 0;
		} else {
			prefix = (uint32_t)__clz((int)top);
			if (prefix > 25) {
				fail(l, -EIO); /* the reference caps level_prefix at 25 */
				return;
			}
			if (prefix >= 15) { /* escape: prefix now, up to 22 suffix bits in the next step */
				l.cache <<= prefix + 1;
				l.nbits -= (int)prefix + 1;
				if (l.nbits < 0) {
					l.cache = 0;
					l.nbits = 0;
					fail(l, -EIO);
					return;
				}
				l.pend = prefix + 1;
				return;
			}
			suffix_off = prefix + 1;
		}
		const uint32_t sl = l.suffix_len;
		uint32_t code = (prefix < 15 ? prefix : 15) << sl;
		uint32_t sz = 0;
		if (sl > 0 || prefix >= 14)
			sz = (prefix == 14 && sl == 0) ? 4 : prefix >= 15 ? prefix - 3 : sl;
		if (sz)
			code += (top << suffix_off) >> (32 - sz);
		n = suffix_off + sz;
		if (prefix >= 15 && sl == 0)
			code += 15;
		if (prefix >= 16)
			code += (1u << (prefix - 3)) - 4096;
		if (l.ci == l.k && l.k < 3)
			code += 2;
		const int32_t lv = (code & 1) ? (-(int32_t)code - 1) >> 1 : (int32_t)((code + 2) >> 1);
		*lev_ptr(l, l.ci) = (int16_t)lv;
		uint32_t nsl = sl == 0 ? 1 : sl;
		const int32_t a16 = (int16_t)lv;
		const int32_t a = a16 < 0 ? -a16 : a16;
		if (a > (3 << (nsl - 1)) && nsl < 6)
			nsl++;
		l.suffix_len = nsl;
		if (++l.ci == l.tc) {
			l.ci = 0;
			l.state = l.tc < l.maxc ? S_ZEROS : S_RUN;
		}
	}

	/* consume */
	if ((int)n > l.nbits) {
		l.cache = 0;
		l.nbits = 0;
		fail(l, -EIO);
		return;
	}
	l.cache <<= n;
	l.nbits -= (int)n;
}

/*
 * One step of a lane outside the residual blocks (macroblock header, skip runs, macroblock end,
 * I_PCM samples): at most 32 bits read.  Called with a state that is neither S_DONE nor a
 * residual state.
 */
template <bool FULL>
__device__ __forceinline__ void hdr_step(Lane &l)
{
	if (l.nbits < 32)
		refill(l);
	const uint32_t top = (uint32_t)(l.cache >> 32);
	uint32_t n = 0;          /* bits this step consumes */
	uint32_t hf = 0, hi = 0; /* the step's checksum term (field, index, value) */
	int64_t hv = 0;
	uint32_t val = 0;
	uint32_t state = l.state;

	/* Exp-Golomb, 9.1: same result as h264_bs_read_bits_ue for codes the reference handles without
	 * undefined behaviour (fewer than 32 leading zeros).  A code longer than 32 bits takes two
	 * steps: its leading zeros first (pend), then the rest. */
	bool want_ue = (CAVLC2_UE_STATES >> state) & 1u;
	uint32_t rbit = 0, rmax = 0;
	if (state == S_REF) {
		rbit = (uint32_t)__ffs((int)l.ref_mask) - 1;
		rmax = (rbit & 4) ? l.max1 : l.max0;
		if (CAVLC2_MBAFF(l) && (l.flags & F_CUR_FIELD))
			rmax = 2 * rmax + 1;
		want_ue = rmax > 1;
	}
	if (want_ue) {
		if (l.pend) {
			n = l.pend + 1;
			val = (top >> (32 - n)) - 1;
			l.pend = 0;
		} else {
			if (top == 0) {
				fail(l, -EIO);
				return;
			}
			const uint32_t lz = (uint32_t)__clz((int)top);
			if (lz > 15) {
				l.cache <<= lz;
				l.nbits -= (int)lz;
				l.pend = lz;
				return;
			}
			n = 2 * lz + 1;
			val = (top >> (32 - n)) - 1;
		}
	}

	switch (state) {
	case S_SKIP_RUN:
		if (val > 0) {
			l.k = val;
			l.flags |= F_PREV_SKIPPED;
			l.state = S_SKIP_EMIT;
		} else {
			l.flags &= ~F_PREV_SKIPPED;
			l.state = mb_start_state(l);
		}
		break;

	case S_SKIP_EMIT: {
		if (l.count >= l.cap || l.cur >= l.pic_size) {
			fail(l, -ENOBUFS);
			return;
		}
		const uint32_t t = (l.flags & F_B) ? MB_B_SKIP : MB_P_SKIP;
		/* what the reference holds in mb_field_decoding_flag when it delivers the macroblock
		 * (h264_new_macroblock, src/h264_slice_data.c:1144-1175) */
		uint32_t field = (l.flags & F_FIELD_PIC) ? 1u : 0u;
		if (CAVLC2_MBAFF(l)) {
			const uint32_t bottom = l.cur & 1u;
			if (!bottom) {
				pair_begin(l);
				l.flags |= F_TOP_SKIPPED; /* the flag waits for the bottom macroblock: 0 for now */
			} else if (l.flags & F_TOP_SKIPPED) {
				/* both skipped: inferred from the pair to the left, else the pair above, else frame */
				const bool f = (l.flags & F_AVAIL_A) ? (l.flags & F_A_FIELD) != 0
							       : (l.flags & F_AVAIL_B) ? (l.flags & F_B_FIELD) != 0 : false;
				l.flags = (l.flags & ~F_CUR_FIELD) | (f ? F_CUR_FIELD : 0u);
				field = f;
			} else {
				field = (l.flags & F_CUR_FIELD) ? 1u : 0u; /* the top macroblock's */
			}
			uint32_t *slot = ring_slot(l, l.cur >> 1) + (bottom ? 4 : 0);
			slot[0] = slot[1] = slot[2] = 0;
			const uint32_t cb = cur_buf(l);
#pragma unroll
			for (uint32_t w = 0; w < 12; w++)
				l.sm[(cb * 12u + w) * CAVLC2_STRIDE] = 0;
		} else {
			uint32_t *slot = ring_slot(l, l.cur);
			slot[0] = slot[1] = slot[2] = 0;
		}
		l.rec[l.count].mb_addr = l.cur;
		l.rec[l.count].mb_type = t;
		l.rec[l.count].hash = field ? hash_term(H264GPU_F_MB_FIELD_DECODING_FLAG, 0, 1) : 0;
		if (FULL && l.syn)
			syn_open(l.syn + l.count, l.cur, t);
		l.count++;
		if (CAVLC2_MBAFF(l)) {
			if (l.cur & 1u)
				pair_end(l);
			l.cur++;
		} else {
			l.prev_addr = l.cur;
			l.flags |= F_PREV_ZERO;
			l.cur = next_mb_addr(l, l.cur);
		}
		if (--l.k == 0) {
			if (!more_rbsp_data(l)) {
				l.state = S_DONE;
				return;
			}
			l.state = mb_start_state(l);
		}
		break;
	}

	case S_MB_TYPE: {
		if (l.count >= l.cap || l.cur >= l.pic_size) {
			fail(l, -ENOBUFS);
			return;
		}
		/* new macroblock: neighbours (6.4.9; with several slice groups a neighbour also has to
		 * be of this slice's group), counts cleared */
		const uint32_t cur = l.cur;
		l.flags &= ~(F_I16 | F_T8 | F_NO_SUB_LT8 | F_NXN | F_INTRA_CBP | F_DIRECT16 | F_PRED8);
		l.flags |= F_NO_SUB_LT8;
		if (!CAVLC2_MBAFF(l)) { /* MBAFF: pair_begin did this for the pair */
			bool aA = cur >= l.first + 1 && cur % l.W != 0;
			bool aB = cur >= l.first + l.W;
			if (l.gmap != nullptr) {
				aA = aA && l.gmap[cur - 1] == l.gmap[cur];
				aB = aB && l.gmap[cur - l.W] == l.gmap[cur];
			}
			if (aA && l.prev_addr != cur - 1)
				aA = false; /* cannot happen: the macroblock before cur in the slice is cur - 1 then */
			l.flags &= ~(F_AVAIL_A | F_AVAIL_B);
			l.flags |= (aA ? F_AVAIL_A : 0u) | (aB ? F_AVAIL_B : 0u);
			if (aB) {
				const uint32_t *up = ring_slot(l, cur - l.W);
				l.top0 = up[0];
				l.top1 = up[1];
				l.top2 = up[2];
			}
		}
		{
			const uint32_t cb = cur_buf(l);
#pragma unroll
			for (uint32_t w = 0; w < 12; w++)
				l.sm[(cb * 12u + w) * CAVLC2_STRIDE] = 0;
		}
		l.hash = 0;
		l.cbp = 0;
		l.slots = 0;
		l.ref_mask = 0;
		l.mvd_mask = 0;
		l.mb_type = MB_UNKNOWN;
		if (FULL && l.syn)
			syn_open(l.syn + l.count, cur, 0);
		/* field macroblock: field picture, or the pair's flag in an MBAFF frame; ref_idx then
		 * ranges over fields (src/h264_slice_data.c:1199-1205) and is sent even for one frame */
		const bool mb_field = CAVLC2_MBAFF(l) ? (l.flags & F_CUR_FIELD) != 0 : (l.flags & F_FIELD_PIC) != 0;
		const bool mf = CAVLC2_MBAFF(l) && mb_field;
		const bool r0 = l.max0 > 0 || mf, r1 = l.max1 > 0 || mf;
		hash_add<FULL>(l, H264GPU_F_MB_FIELD_DECODING_FLAG, 0, mb_field ? 1 : 0);

		/* h264_read_mb_type: src/h264_slice_data.c:839-969 */
		uint32_t type = val;
		hash_add<FULL>(l, H264GPU_F_RAW_MB_TYPE, 0, type);
		const uint32_t st = l.sp->slice_type;
		bool intra = false;
		uint32_t num_part = 0, m0 = 0, m1 = 0; /* partition prediction: 0 L0, 1 L1, 2 Bi */
		if (st == ST_I) {
			intra = true;
		} else if (st == ST_SI) {
			if (type == 0) {
				l.mb_type = MB_SI;
			} else {
				type -= 1;
				intra = true;
			}
		} else if (st == ST_P || st == ST_SP) {
			if (type == 0) {
				l.mb_type = MB_P_16x16;
				num_part = 1;
			} else if (type <= 2) {
				l.mb_type = type == 1 ? MB_P_16x8 : MB_P_8x16;
				num_part = 2;
			} else if (type == 3) {
				l.mb_type = MB_P_8x8;
			} else if (type == 4) {
				l.mb_type = MB_P_8x8ref0;
			} else {
				type -= 5;
				intra = true;
			}
		} else { /* B */
			if (type == 0) {
				l.mb_type = MB_B_Direct_16x16;
				l.flags |= F_DIRECT16;
			} else if (type <= 3) {
				l.mb_type = MB_B_16x16;
				num_part = 1;
				m0 = type - 1;
			} else if (type <= 21) {
				l.mb_type = ((type - 4) & 1) ? MB_B_8x16 : MB_B_16x8;
				num_part = 2;
				/* pairs in raw mb_type order (src/h264_slice_data.c:847-866):
				 * (L0,L0)(L1,L1)(L0,L1)(L1,L0)(L0,Bi)(L1,Bi)(Bi,L0)(Bi,L1)(Bi,Bi), 2 bits each */
				const uint32_t pr = (type - 4) >> 1;
				m0 = (0x2a444u >> (2 * pr)) & 3u;
				m1 = (0x24a14u >> (2 * pr)) & 3u;
			} else if (type == 22) {
				l.mb_type = MB_B_8x8;
			} else {
				type -= 23;
				intra = true;
			}
		}
		if (intra) {
			if (type == 0) {
				l.mb_type = MB_I_NxN;
				l.flags |= F_NXN;
			} else if (type <= 24) {
				l.mb_type = MB_I_16x16;
				l.flags |= F_I16;
				hash_add<FULL>(l, H264GPU_F_I16_PRED_MODE, 0, (type - 1) % 4);
				const uint32_t cl = type <= 12 ? 0 : 15, cc = ((type - 1) / 4) % 3;
				l.cbp = cl | cc << 4;
				hash_add<FULL>(l, H264GPU_F_CBP_LUMA, 0, cl);
				hash_add<FULL>(l, H264GPU_F_CBP_CHROMA, 0, cc);
			} else if (type == 25) {
				l.mb_type = MB_I_PCM;
			} else {
				fail(l, -EIO);
				return;
			}
		}
		const uint32_t t = l.mb_type;
		if (t == MB_I_NxN || t == MB_I_16x16 || t == MB_SI)
			l.flags |= F_INTRA_CBP;
		if (t == MB_I_PCM) {
			l.state = S_PCM_ALIGN;
		} else if (t == MB_P_8x8 || t == MB_P_8x8ref0 || t == MB_B_8x8) {
			l.k = 0;
			l.state = S_SUB_TYPE;
		} else if (t == MB_I_NxN || t == MB_SI) {
			l.k = 0;
			l.kend = 16;
			if (t == MB_I_NxN && (l.flags & F_T8MODE)) {
				/* transform_size_8x8_flag right after mb_type (at most 9 + 1 bits) */
				if ((top >> (31 - n)) & 1u) {
					l.flags |= F_T8 | F_PRED8;
					l.kend = 4;
				}
				n += 1;
			}
			l.state = S_PRED;
		} else if (t == MB_I_16x16) {
			l.state = (l.cat == 1 || l.cat == 2) ? S_CHROMA_PRED : S_QPD;
		} else if (t == MB_B_Direct_16x16) {
			l.state = S_CBP;
		} else {
			/* mb_pred of an inter macroblock: src/h264_syntax_slice_data.h:506-601 */
			for (uint32_t i = 0; i < num_part; i++) {
				const uint32_t m = i ? m1 : m0;
				if (m != 1) { /* uses list 0 */
					if (r0)
						l.ref_mask |= 1u << i;
					l.mvd_mask |= 3ull << (8 * i);
				}
				if (m != 0) { /* uses list 1 */
					if (r1)
						l.ref_mask |= 16u << i;
					l.mvd_mask |= 3ull << (32 + 8 * i);
				}
			}
			l.state = l.ref_mask ? S_REF : S_MVD;
		}
		break;
	}

	case S_SUB_TYPE: {
		/* sub_mb_pred: src/h264_syntax_slice_data.h:422-503 */
		const uint32_t t = val, i = l.k;
		hf = H264GPU_F_RAW_SUB_MB_TYPE;
		hi = i;
		hv = t;
		bool direct = false;
		uint32_t nsub, m; /* m: 0 L0, 1 L1, 2 Bi */
		if (!(l.flags & F_B)) {
			if (t >= 4) {
				fail(l, -EIO);
				return;
			}
			nsub = t == 0 ? 1 : t == 3 ? 4 : 2;
			m = 0;
		} else {
			if (t >= 13) {
				fail(l, -EIO);
				return;
			}
			direct = t == 0;
			nsub = (t == 0 || t >= 10) ? 4 : t <= 3 ? 1 : 2;
			m = (t == 1 || t == 4 || t == 5 || t == 10) ? 0 : (t == 2 || t == 6 || t == 7 || t == 11) ? 1 : 2;
		}
		if (!direct) {
			const uint64_t bits = (1ull << (2 * nsub)) - 1;
			const bool mf = CAVLC2_MBAFF(l) && (l.flags & F_CUR_FIELD);
			if (m != 1) {
				if ((l.max0 > 0 || mf) && l.mb_type != MB_P_8x8ref0)
					l.ref_mask |= 1u << i;
				l.mvd_mask |= bits << (8 * i);
			}
			if (m != 0) {
				if (l.max1 > 0 || mf)
					l.ref_mask |= 16u << i;
				l.mvd_mask |= bits << (32 + 8 * i);
			}
			if (nsub > 1)
				l.flags &= ~F_NO_SUB_LT8;
		} else if (!(l.flags & F_D8INF)) {
			l.flags &= ~F_NO_SUB_LT8;
		}
		if (++l.k == 4)
			l.state = l.ref_mask ? S_REF : l.mvd_mask ? S_MVD : S_CBP;
		break;
	}

	case S_REF: {
		if (rmax == 1) { /* te() with range 1: one inverted bit */
			n = 1;
			val = (top >> 31) ^ 1u;
		}
		hf = (rbit & 4) ? H264GPU_F_REF_IDX_L1 : H264GPU_F_REF_IDX_L0;
		hi = rbit & 3;
		hv = (uint8_t)val;
		l.ref_mask &= l.ref_mask - 1;
		if (l.ref_mask == 0)
			l.state = l.mvd_mask ? S_MVD : S_CBP;
		break;
	}

	case S_MVD: {
		const uint32_t bit = (uint32_t)__ffsll((long long)l.mvd_mask) - 1;
		hf = (bit & 32) ? H264GPU_F_MVD_L1 : H264GPU_F_MVD_L0;
		hi = bit & 31;
		hv = (int16_t)((val & 1) ? (int32_t)((val + 1) >> 1) : -(int32_t)(val >> 1));
		l.mvd_mask &= l.mvd_mask - 1;
		if (l.mvd_mask == 0)
			l.state = S_CBP;
		break;
	}

	case S_PRED: {
		if (top >> 31) {
			n = 1;
			hv = -1;
		} else {
			n = 4;
			hv = (top >> 28) & 7u;
		}
		hf = (l.flags & F_PRED8) ? H264GPU_F_INTRA8X8_PRED_MODE : H264GPU_F_INTRA4X4_PRED_MODE;
		hi = l.k;
		if (++l.k == l.kend)
			l.state = (l.cat == 1 || l.cat == 2) ? S_CHROMA_PRED : S_CBP;
		break;
	}

	case S_CHROMA_PRED:
		hf = H264GPU_F_INTRA_CHROMA_PRED_MODE;
		hv = (uint8_t)val;
		l.state = (l.flags & F_I16) ? S_QPD : S_CBP;
		break;

	case S_CBP: {
		/* h264_read_coded_block_pattern: src/h264_slice_data.c:1041-1080 */
		uint32_t cbp;
		const uint32_t col = (l.flags & F_INTRA_CBP) ? 0 : 1;
		if (l.cat == 1 || l.cat == 2) {
			if (val >= 48) {
				fail(l, -EIO);
				return;
			}
			cbp = cavlc_cbp_chroma[val][col];
		} else {
			if (val >= 16) {
				fail(l, -EIO);
				return;
			}
			cbp = cavlc_cbp_nochroma[val][col];
		}
		hash_add<FULL>(l, H264GPU_F_CBP, 0, cbp);
		const uint32_t cl = cbp % 16, cc = cbp / 16;
		l.cbp = cl | cc << 4;
		if (cl > 0 && (l.flags & F_T8MODE) && !(l.flags & F_NXN) && (l.flags & F_NO_SUB_LT8) &&
		    (!(l.flags & F_DIRECT16) || (l.flags & F_D8INF))) {
			/* transform_size_8x8_flag right after the pattern (at most 11 + 1 bits) */
			if ((top >> (31 - n)) & 1u)
				l.flags |= F_T8;
			n += 1;
		}
		hash_add<FULL>(l, H264GPU_F_TRANSFORM_8X8, 0, (l.flags & F_T8) ? 1 : 0);
		hash_add<FULL>(l, H264GPU_F_CBP_LUMA, 0, cl);
		hash_add<FULL>(l, H264GPU_F_CBP_CHROMA, 0, cc);
		l.state = cbp ? S_QPD : S_MB_END;
		break;
	}

	case S_QPD:
		hf = H264GPU_F_MB_QP_DELTA;
		hv = (val & 1) ? (int32_t)((val + 1) >> 1) : -(int32_t)(val >> 1);
		l.slots = residual_slots(l.cat, (l.flags & F_I16) != 0, l.cbp & 15u, l.cbp >> 4);
		l.state = l.slots ? S_TOKEN : S_MB_END;
		break;

	case S_MB_END: {
		uint32_t w[3];
		bottom_rows(l, w);
		const bool mbaff = CAVLC2_MBAFF(l);
		uint32_t *slot = mbaff ? ring_slot(l, l.cur >> 1) + ((l.cur & 1u) ? 4 : 0) : ring_slot(l, l.cur);
		slot[0] = w[0];
		slot[1] = w[1];
		slot[2] = w[2];
		l.rec[l.count].mb_addr = l.cur;
		l.rec[l.count].mb_type = l.mb_type;
		l.rec[l.count].hash = l.hash;
		if (FULL && l.syn)
			l.syn[l.count].mb_type = l.mb_type;
		l.count++;
		if (mbaff) {
			if (l.cur & 1u)
				pair_end(l);
			l.cur++;
		} else {
			l.prev_addr = l.cur;
			l.flags = (l.flags & ~F_PREV_ZERO) ^ F_WHICH;
			l.cur = next_mb_addr(l, l.cur);
		}
		if (!more_rbsp_data(l)) {
			l.state = S_DONE;
			return;
		}
		l.state = (l.flags & F_INTER) ? (uint32_t)S_SKIP_RUN : mb_start_state(l);
		break;
	}

	case S_PCM_ALIGN: {
		n = (uint32_t)l.nbits & 7u;
		if (n && (top >> (32 - n)) != 0) {
			fail(l, -EIO);
			return;
		}
		l.k = 0;
		l.kend = 256 + 32 * (l.cat == 0 ? 0u : l.cat == 1 ? 4u : l.cat == 2 ? 8u : 16u);
		l.state = S_PCM;
		break;
	}

	case S_PCM: {
		const uint32_t k = l.k;
		if (k < 256) {
			n = l.sp->bit_depth_luma;
			hf = H264GPU_F_PCM_LUMA;
			hi = k;
		} else {
			const uint32_t nc = (l.kend - 256) >> 1, j = k - 256;
			n = l.sp->bit_depth_chroma;
			hf = H264GPU_F_PCM_CHROMA;
			hi = j >= nc ? 256 + (j - nc) : j;
		}
		hv = n ? (top >> (32 - n)) : 0;
		if (++l.k >= l.kend) {
			const uint32_t cb = cur_buf(l);
#pragma unroll
			for (uint32_t w = 0; w < 12; w++)
				l.sm[(cb * 12u + w) * CAVLC2_STRIDE] = 0x10101010u;
			l.state = S_MB_END;
		}
		break;
	}

	case S_MB_FIELD:
		/* 7.3.4: read with the top macroblock of a pair, or with the bottom one when the top was
		 * skipped (then it also becomes the skipped top's flag) */
		if (!(l.cur & 1u))
			pair_begin(l);
		n = 1;
		l.flags = (l.flags & ~F_CUR_FIELD) | ((top >> 31) ? F_CUR_FIELD : 0u);
		l.state = S_MB_TYPE;
		break;

	default:
		break;
	}

	/* consume */
	if ((int)n > l.nbits) {
		l.cache = 0;
		l.nbits = 0;
		fail(l, -EIO);
		return;
	}
	l.cache <<= n;
	l.nbits -= (int)n;
	if (hv != 0) {
		if (FULL && l.syn != nullptr)
			syn_store(l.syn + l.count, hf, hi, hv);
		l.hash += hash_term(hf, hi, hv);
	}
}

/* up to CAVLC2_RES_REPEAT elements while the lane stays inside residual blocks: the warp loop's
 * vote is paid once for them (the levels and runs of a block follow each other) */
#ifndef CAVLC2_RES_REPEAT
#define CAVLC2_RES_REPEAT 4
#endif
template <bool FULL>
__device__ __forceinline__ void res_step(Lane &l, uint32_t repeat = CAVLC2_RES_REPEAT)
{
#pragma unroll 1
	for (uint32_t k = 0; k < repeat; k++) {
		res_step1<FULL>(l);
		if (!CAVLC2_IS_RES(l.state))
			break;
	}
}

template <bool FULL>
__device__ __forceinline__ void step(Lane &l)
{
	if (CAVLC2_IS_RES(l.state))
		res_step<FULL>(l);
	else
		hdr_step<FULL>(l);
}

/* the serial form (emulator harness, one slice per call) */
template <bool FULL>
__device__ __forceinline__ void parse_slice(const uint8_t *stream, uint64_t stream_len, const h264gpu_slice_params &sp,
					     uint8_t *ring, h264gpu_mb_record *rec, h264gpu_slice_result &res,
					     h264_mb_syntax *syn, const uint8_t *group_maps, uint32_t *sm)
{
	Lane l;
	l.sm = sm;
	l.can_mbaff = true;
	slice_begin(l, stream, stream_len, sp, ring, rec, syn, group_maps, 0xffffu);
	while (l.state != S_DONE)
		step<FULL>(l);
	slice_end(l, res);
}

struct CavlcArgs {
	const uint8_t *stream;
	uint64_t stream_len;
	const h264gpu_slice_params *params;
	uint32_t n_slices;
	h264gpu_mb_record *records;
	h264gpu_slice_result *results;
	uint8_t *ring; /* one ring of ring_stride bytes per LANE of the grid */
	uint64_t ring_stride;
	uint32_t ring_w; /* widest picture (in MBs) a ring can hold */
	uint32_t lanes_log2; /* log2 of the lanes of a warp that carry slices (0..5) */
	h264_mb_syntax *syntax; /* full records (index = record index), or NULL */
	const uint8_t *group_maps; /* macroblock -> slice group maps (slice i: + params[i].row_state_off), or NULL */
	uint32_t *next_slice; /* [0] ticket counter of the frame / field kernel, [1] of the MBAFF kernel,
				 [2] number of MBAFF slices; set by order_kernel */
	uint32_t *order;      /* slice indices, longest NAL first */
	uint32_t repeat;      /* residual elements a lane takes per vote of the warp loop */
};

/*
 * Slices in descending order of NAL length (counting sort on the top 8 bits, one block).  The
 * lanes of a warp take consecutive tickets: slices of similar length share a warp, so its lanes
 * finish together (with the list order, a warp waited for its one I slice with 3 of 16 lanes
 * busy), and the longest slices start first.
 */
#define CAVLC2_ORDER_T 1024
__global__ void __launch_bounds__(CAVLC2_ORDER_T) order_kernel(const h264gpu_slice_params *params, uint32_t n,
								 uint32_t *order, uint32_t *counter, uint32_t counter_init, uint32_t sort)
{
#ifndef H264_EMU
	if (!sort) { /* A/B: list order */
		for (uint32_t i = threadIdx.x; i < n; i += CAVLC2_ORDER_T)
			order[i] = i;
		if (threadIdx.x == 0) {
			counter[0] = counter[1] = counter_init;
			counter[2] = 1; /* unknown: let the MBAFF kernel look */
		}
		return;
	}
	__shared__ uint32_t hist[256];
	__shared__ uint32_t smax;
	const uint32_t t = threadIdx.x;
	if (t < 256)
		hist[t] = 0;
	if (t == 0)
		smax = 0;
	__syncthreads();
	uint32_t m = 0, n_mbaff = 0;
	for (uint32_t i = t; i < n; i += CAVLC2_ORDER_T) {
		m = max(m, params[i].nal_len);
		n_mbaff += params[i].mbaff_frame_flag && !params[i].entropy_coding_mode_flag;
	}
	if (t == 0)
		counter[2] = 0;
	__syncthreads();
	if (n_mbaff)
		atomicAdd(&counter[2], n_mbaff);
	m = __reduce_max_sync(FULL_MASK, m);
	if ((t & 31) == 0)
		atomicMax(&smax, m);
	__syncthreads();
	const uint32_t bits = 32 - (uint32_t)__clz((int)(smax | 1u));
	const uint32_t shift = bits > 8 ? bits - 8 : 0;
	for (uint32_t i = t; i < n; i += CAVLC2_ORDER_T)
		atomicAdd(&hist[255 - (params[i].nal_len >> shift)], 1u);
	__syncthreads();
	if (t == 0) {
		uint32_t acc = 0;
		for (uint32_t b = 0; b < 256; b++) {
			const uint32_t c = hist[b];
			hist[b] = acc;
			acc += c;
		}
		counter[0] = counter[1] = counter_init;
	}
	__syncthreads();
	for (uint32_t i = t; i < n; i += CAVLC2_ORDER_T)
		order[atomicAdd(&hist[255 - (params[i].nal_len >> shift)], 1u)] = i;
#endif
}

/*
 * Every working lane pulls slices from a counter until none is left (slices differ a lot in
 * length: I / P / B, skip runs) and the warp steps its lanes together.  Few slices are spread
 * over more warps (lanes_log2 < 5) so that the SMs have warps to switch between.
 */
template <bool FULL, bool MBAFF>
__global__ void __launch_bounds__(CAVLC2_STRIDE) cavlc_steps_kernel(const CavlcArgs a)
{
#ifndef H264_EMU
	extern __shared__ uint32_t smem[];
	/* two instantiations share a launch's slices: frame / field slices here, MBAFF slices in the
	 * one that carries the macroblock-pair code (launched behind it, gone at once when the list
	 * has no MBAFF slice) */
	if (MBAFF && a.next_slice[2] == 0)
		return;
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t step_l = 32u >> a.lanes_log2; /* working lanes are multiples of this */
	const bool worker = !(lane & (step_l - 1));
	const uint32_t gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const uint32_t glane = (gwarp << a.lanes_log2) + lane / step_l;
	Lane l;
	l.sm = smem + threadIdx.x;
	l.can_mbaff = MBAFF;
	l.state = S_DONE;
	uint32_t slice = 0xffffffffu;
	bool out = !worker, first_done = false;
	for (;;) {
		if (!out && l.state == S_DONE) {
			if (slice != 0xffffffffu) {
				h264gpu_slice_result res;
				slice_end(l, res);
				a.results[slice] = res;
			}
			/* first ticket = the lane's own number (consecutive within the warp), then the counter */
			const uint32_t ticket = slice == 0xffffffffu && !first_done ? glane : atomicAdd(a.next_slice + (MBAFF ? 1 : 0), 1u);
			first_done = true;
			if (ticket >= a.n_slices) {
				out = true;
			} else {
				slice = a.order[ticket];
				const h264gpu_slice_params &sp = a.params[slice];
				if ((sp.mbaff_frame_flag != 0 && !sp.entropy_coding_mode_flag) != MBAFF) {
					slice = 0xffffffffu; /* the other instantiation's */
					continue;
				}
				slice_begin(l, a.stream, a.stream_len, sp, a.ring + (uint64_t)glane * a.ring_stride,
					    a.records + sp.mb_out_off, FULL && a.syntax ? a.syntax + sp.mb_out_off : nullptr,
					    a.group_maps, a.ring_w);
			}
		}
		if (__all_sync(FULL_MASK, out))
			break;
		/* the residual blocks are where the steps are: while most of the warp's running lanes are
		 * inside one, only those lanes step (a short loop body: one table probe or one level); then
		 * every lane outside takes one header step.  (Measured against a fixed burst of 2 / 4 / 8
		 * residual steps per header step, with and without a vote: this rule is the fastest at 300
		 * and at 16000 slices, profiles/r02_k4_loop_variants.txt) */
		for (;;) {
			const bool run = !out; /* a lane waiting for its next slice votes for leaving too */
			const bool res = run && CAVLC2_IS_RES(l.state);
			const uint32_t n_res = (uint32_t)__popc(__ballot_sync(FULL_MASK, res));
			const uint32_t n_run = (uint32_t)__popc(__ballot_sync(FULL_MASK, run));
			if (n_res == 0 || 2 * n_res < n_run)
				break;
			if (res)
				res_step<FULL>(l, a.repeat);
		}
		if (!out && l.state != S_DONE && !CAVLC2_IS_RES(l.state))
			hdr_step<FULL>(l);
	}
#endif
}

} /* namespace cavlc2 */

#endif /* CAVLC_STEPS_CUH */
