/*
 * cabac_engine.h — the two coder types cabac_syntax.h is instantiated with.
 *
 *   cabac::Dec  binary arithmetic DECODER (Rec. ITU-T H.264 9.3.1.2, 9.3.3.2), host + device.
 *               libh264 has no CABAC decoder (h264_bac_decode_init,
 *               src/h264_bac.c:236-247, is dead code); this follows the standard's flowcharts
 *               and is the exact inverse of the reference's encoder engine
 *               (src/h264_bac.c:150-358), which the tests drive bin by bin against it.
 *               Reads the ESCAPED NAL and drops emulation prevention bytes like
 *               h264_bs_fetch (include/h264/h264_bitstream.h:168-190).
 *   cabac::Enc  binary arithmetic ENCODER (9.3.4), host only, used by the synthetic-stream
 *               generator; same arithmetic as src/h264_bac.c:150-358 (tests compare the bits).
 *
 * Context state byte: pStateIdx << 1 | valMPS.  Initialisation 9.3.1.1 as
 * h264_bac_state_init (src/h264_bac.c:216-230); (m, n) from cabac_tables.h, which is
 * generated from the compiled reference's tables.
 */
#ifndef CABAC_ENGINE_H
#define CABAC_ENGINE_H

#include <stdint.h>

#ifndef CABAC_HD
#define CABAC_HD
#endif
/* hot, many-call-site routines stay real functions on the device (code size: the fully
 * inlined kernel does not fit the instruction cache) */
#ifndef CABAC_OUTLINE
#define CABAC_OUTLINE CABAC_HD
#endif

namespace cabac {

enum { kNumCtx = 460 }; /* ctxIdx 0..459: everything below the 4:4:4-only range */

struct Tables {
	const uint8_t *range_lps; /* [64][4] */
	const uint8_t *trans_lps; /* [64] */
	const uint8_t *trans_mps; /* [64] */
	/* device: one 8-byte entry per context state byte s = pStateIdx << 1 | valMPS:
	 * bytes 0..3 rangeTabLPS[pStateIdx][0..3], byte 4 the state byte after an MPS, byte 5 after
	 * an LPS (valMPS flipped at pStateIdx 0).  One shared-memory load per bin instead of a
	 * chain of three (fused_entry() builds it from the three tables above). */
	const uint64_t *fused; /* [128] */
};

CABAC_HD static inline uint64_t fused_entry(uint32_t s, const uint8_t (*range_lps)[4], const uint8_t *trans_lps,
					     const uint8_t *trans_mps)
{
	const uint32_t ps = s >> 1, mps = s & 1u;
	const uint32_t after_mps = (uint32_t)trans_mps[ps] << 1 | mps;
	const uint32_t after_lps = (uint32_t)trans_lps[ps] << 1 | (ps == 0 ? mps ^ 1u : mps);
	return (uint64_t)range_lps[ps][0] | (uint64_t)range_lps[ps][1] << 8 | (uint64_t)range_lps[ps][2] << 16 |
	       (uint64_t)range_lps[ps][3] << 24 | (uint64_t)after_mps << 32 | (uint64_t)after_lps << 40;
}

CABAC_HD static inline uint8_t init_state(int m, int n, int slice_qp)
{
	const int qp = slice_qp < 1 ? 1 : slice_qp > 51 ? 51 : slice_qp;
	const int idx = ((m * qp) >> 4) + n;
	if (idx <= 63) {
		const int i = idx < 1 ? 1 : idx;
		return (uint8_t)((63 - i) << 1);
	}
	return (uint8_t)((((idx > 126 ? 126 : idx) - 64) << 1) | 1);
}

/* mn: cabac_init_mn[table] with table = 0 for I slices, 1 + cabac_init_idc otherwise */
CABAC_HD static inline void init_contexts(uint8_t *st, uint32_t stride, const int8_t (*mn)[2], int slice_qp)
{
	for (uint32_t i = 0; i < kNumCtx; i++)
		st[i * stride] = init_state(mn[i][0], mn[i][1], slice_qp);
}

CABAC_HD static inline int clz32(uint32_t x)
{
#ifdef __CUDA_ARCH__
	return __clz((int)x);
#else
	return x ? __builtin_clz(x) : 32;
#endif
}

struct Dec {
	static const bool kEncode = false;
	const uint8_t *p; /* NAL, escaped */
	uint32_t len, pos, zeros;
	uint32_t cache;
	int nbits;
	uint32_t range, offset;
	uint8_t *st;
	uint32_t stride;
	Tables t;
	bool bad;

	CABAC_HD bool failed() const { return bad; }

	/* next RBSP byte (emulation prevention bytes dropped) */
	CABAC_OUTLINE uint32_t fetch()
	{
		if (pos >= len) {
			bad = true;
			return 0;
		}
		uint32_t c = p[pos++];
		if (zeros >= 2 && c == 3) {
			zeros = 0;
			if (pos >= len) {
				bad = true;
				return 0;
			}
			c = p[pos++];
		}
		zeros = c == 0 ? zeros + 1 : 0;
		return c;
	}

	CABAC_HD uint32_t get(int n)
	{
		while (nbits < n) {
			cache = (cache << 8) | fetch();
			nbits += 8;
		}
		nbits -= n;
		return (cache >> nbits) & ((1u << n) - 1u);
	}

	/* position at raw bit offset bit_off of the NAL, skip cabac_alignment_one_bits,
	 * initialise the engine (9.3.1.2) */
	CABAC_HD void start(const uint8_t *nal, uint32_t nal_len, uint32_t bit_off)
	{
		p = nal;
		len = nal_len;
		bad = false;
		cache = 0;
		nbits = 0;
		pos = (bit_off + 7) >> 3; /* slice data is byte aligned */
		zeros = 0;
		if (pos >= 1 && pos <= len && nal[pos - 1] == 0)
			zeros = (pos >= 2 && nal[pos - 2] == 0) ? 2 : 1;
		restart();
	}

	CABAC_HD void restart()
	{
		range = 510;
		offset = get(9);
		if (offset >= 510)
			bad = true;
	}

	/* raw bit position in the NAL of the next unread bit */
	CABAC_HD uint64_t raw_bitpos() const { return (uint64_t)pos * 8 - (uint64_t)nbits; }

	CABAC_OUTLINE uint32_t bin(uint32_t ctx, uint32_t)
	{
#ifdef __CUDA_ARCH__
		/* the serial chain of a slice: state byte -> ONE table entry -> compare -> state byte;
		 * context bytes and table addressed as shared memory (through the generic pointers the
		 * loads are LD.E and the store makes the compiler reload range from memory) */
		const uint32_t sa = (uint32_t)__cvta_generic_to_shared(st) + ctx * stride;
		uint32_t s, lo, next;
		asm volatile("ld.shared.u8 %0, [%1];" : "=r"(s) : "r"(sa) : "memory"); /* after init_contexts' plain stores */
		asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];"
			     : "=r"(lo), "=r"(next)
			     : "r"((uint32_t)__cvta_generic_to_shared(t.fused) + s * 8));
		uint32_t r = range, o = offset;
		const uint32_t rlps = (lo >> (((r >> 6) & 3) * 8)) & 0xffu;
		r -= rlps;
		const bool lps = o >= r;
		if (lps) {
			o -= r;
			r = rlps;
		}
		asm volatile("st.shared.u8 [%0], %1;" : : "r"(sa), "r"(lps ? next >> 8 : next) : "memory");
		if (r < 256) {
			const int sh = clz32(r) - 23;
			r <<= sh;
			o = (o << sh) | get(sh);
		}
		range = r;
		offset = o;
		return (s & 1u) ^ (lps ? 1u : 0u);
#else
		uint8_t *sp = st + ctx * stride;
		const uint32_t s = *sp;
		uint32_t ps = s >> 1, mps = s & 1u;
		const uint32_t rlps = t.range_lps[ps * 4 + ((range >> 6) & 3)];
		range -= rlps;
		uint32_t b;
		if (offset >= range) {
			b = mps ^ 1u;
			offset -= range;
			range = rlps;
			if (ps == 0)
				mps ^= 1u;
			ps = t.trans_lps[ps];
		} else {
			b = mps;
			ps = t.trans_mps[ps];
		}
		*sp = (uint8_t)(ps << 1 | mps);
		if (range < 256) {
			const int sh = clz32(range) - 23;
			range <<= sh;
			offset = (offset << sh) | get(sh);
		}
		return b;
#endif
	}

	CABAC_OUTLINE uint32_t byp(uint32_t)
	{
		offset = (offset << 1) | get(1);
		if (offset >= range) {
			offset -= range;
			return 1;
		}
		return 0;
	}

	CABAC_HD uint32_t term(uint32_t)
	{
		range -= 2;
		if (offset >= range)
			return 1; /* no renormalisation: every bit up to the stop bit has been read */
		if (range < 256) {
			range <<= 1;
			offset = (offset << 1) | get(1);
		}
		return 0;
	}

	/* I_PCM: samples start at the next byte boundary, the engine restarts after them */
	CABAC_HD void pcm_begin() { nbits = 0; }
	CABAC_HD uint32_t pcm_sample(uint32_t, uint32_t bits) { return get((int)bits); }
	CABAC_HD void pcm_end() { restart(); }
};

#ifndef __CUDACC__
/* ---- encoder (host, generator only) ---------------------------------------------------------- */
struct Enc {
	static const bool kEncode = true;
	uint8_t *out; /* RBSP bytes (not escaped) */
	uint64_t cap, n;
	uint32_t acc;
	int nacc;
	uint32_t low, range, outstanding;
	bool first;
	uint8_t st[1024];
	Tables t;
	uint64_t bins;

	bool failed() const { return false; }

	void wbit(uint32_t b)
	{
		acc = (acc << 1) | (b & 1u);
		if (++nacc == 8) {
			if (n < cap)
				out[n] = (uint8_t)acc;
			n++;
			acc = 0;
			nacc = 0;
		}
	}
	void put(uint32_t b)
	{
		if (first)
			first = false;
		else
			wbit(b);
		for (; outstanding > 0; outstanding--)
			wbit(!b);
	}
	void renorm()
	{
		while (range < 256) {
			if (low < 256) {
				put(0);
			} else if (low < 512) {
				low -= 256;
				outstanding++;
			} else {
				low -= 512;
				put(1);
			}
			range <<= 1;
			low <<= 1;
		}
	}
	void start(uint8_t *buf, uint64_t capacity)
	{
		out = buf;
		cap = capacity;
		n = 0;
		acc = 0;
		nacc = 0;
		bins = 0;
		restart();
	}
	void restart()
	{
		low = 0;
		range = 510;
		first = true;
		outstanding = 0;
	}
	uint32_t bin(uint32_t ctx, uint32_t v)
	{
		v = v ? 1u : 0u;
		uint32_t ps = st[ctx] >> 1, mps = st[ctx] & 1u;
		const uint32_t rlps = t.range_lps[ps * 4 + ((range >> 6) & 3)];
		range -= rlps;
		if (v == mps) {
			ps = t.trans_mps[ps];
		} else {
			low += range;
			range = rlps;
			if (ps == 0)
				mps ^= 1u;
			ps = t.trans_lps[ps];
		}
		st[ctx] = (uint8_t)(ps << 1 | mps);
		renorm();
		bins++;
		return v;
	}
	uint32_t byp(uint32_t v)
	{
		v = v ? 1u : 0u;
		low <<= 1;
		if (v)
			low += range;
		if (low >= 1024) {
			put(1);
			low -= 1024;
		} else if (low >= 512) {
			low -= 512;
			outstanding++;
		} else {
			put(0);
		}
		bins++;
		return v;
	}
	uint32_t term(uint32_t v)
	{
		v = v ? 1u : 0u;
		range -= 2;
		if (v) {
			low += range;
			range = 2;
			renorm();
			put((low >> 9) & 1);
			wbit((low >> 8) & 1);
			wbit(1); /* rbsp_stop_one_bit when this ends the slice */
		} else {
			renorm();
		}
		bins++;
		return v;
	}
	void pcm_begin()
	{
		while (nacc)
			wbit(0); /* pcm_alignment_zero_bit */
	}
	uint32_t pcm_sample(uint32_t v, uint32_t bits)
	{
		for (int k = (int)bits - 1; k >= 0; k--)
			wbit((v >> k) & 1);
		return v;
	}
	void pcm_end() { restart(); }
	/* zero bits up to the byte boundary (rbsp_alignment_zero_bit after the stop bit) */
	void align_zero()
	{
		while (nacc)
			wbit(0);
	}
};
#endif

} /* namespace cabac */

#endif /* CABAC_ENGINE_H */
