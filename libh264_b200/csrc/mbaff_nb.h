/*
 * mbaff_nb.h — H.264 6.4.12.2: neighbouring locations in MBAFF frames, restricted to what the
 * CAVLC parse needs (nC of a 4x4 block: the block to the left, xN < 0, and the block above,
 * yN < 0).  Header only; used by the slice kernel (cavlc_steps.cuh) and by the synthetic stream
 * generator (synth_video.c).  The reference's form: h264_get_neighbouring_locations_mbaff,
 * src/h264_macroblock.c:110-231.
 *
 * A macroblock pair is "field" when its macroblocks have mb_field_decoding_flag = 1.  For the
 * current macroblock (cur_field, cur_bottom) the neighbour sits in one of five macroblocks.
 */
#ifndef MBAFF_NB_H
#define MBAFF_NB_H

#ifndef H264_HD
#ifdef __CUDACC__
#define H264_HD __host__ __device__
#else
#define H264_HD
#endif
#endif

enum {
	MBAFF_NB_NONE = -1,
	MBAFF_NB_A_TOP = 0, /* top macroblock of the pair to the left (mbAddrA) */
	MBAFF_NB_A_BOT,     /* bottom macroblock of that pair (mbAddrA + 1) */
	MBAFF_NB_B_TOP,     /* top macroblock of the pair above (mbAddrB) */
	MBAFF_NB_B_BOT,     /* bottom macroblock of that pair (mbAddrB + 1) */
	MBAFF_NB_CUR_TOP,   /* top macroblock of the current pair (CurrMbAddr - 1) */
};

/* xN < 0: the row yN (0 .. maxH - 1, in samples of the component) of the current macroblock
 * continues in row *yM of the returned macroblock */
H264_HD static inline int mbaff_left(int cur_field, int cur_bottom, int a_avail, int a_field, int yN, int maxH,
				     int *yM)
{
	if (!a_avail)
		return MBAFF_NB_NONE;
	if (!cur_field) {
		if (!a_field) {
			*yM = yN;
			return cur_bottom ? MBAFF_NB_A_BOT : MBAFF_NB_A_TOP;
		}
		*yM = cur_bottom ? (yN + maxH) >> 1 : yN >> 1;
		return (yN & 1) ? MBAFF_NB_A_BOT : MBAFF_NB_A_TOP;
	}
	if (a_field) {
		*yM = yN;
		return cur_bottom ? MBAFF_NB_A_BOT : MBAFF_NB_A_TOP;
	}
	if (yN < maxH / 2) {
		*yM = (yN << 1) + (cur_bottom ? 1 : 0);
		return MBAFF_NB_A_TOP;
	}
	*yM = (yN << 1) + (cur_bottom ? 1 : 0) - maxH;
	return MBAFF_NB_A_BOT;
}

/* yN = -1: always the bottom row of samples (or the one above it) of the returned macroblock,
 * i.e. its bottom row of 4x4 blocks */
H264_HD static inline int mbaff_up(int cur_field, int cur_bottom, int b_avail, int b_field)
{
	if (!cur_field)
		return cur_bottom ? MBAFF_NB_CUR_TOP : (b_avail ? MBAFF_NB_B_BOT : MBAFF_NB_NONE);
	if (!b_avail)
		return MBAFF_NB_NONE;
	if (cur_bottom)
		return MBAFF_NB_B_BOT;
	return b_field ? MBAFF_NB_B_TOP : MBAFF_NB_B_BOT;
}

#endif /* MBAFF_NB_H */
