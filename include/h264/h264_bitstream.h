/*
 * h264/h264_bitstream.h — bit-level reader/writer object of the libh264 API.
 *
 * struct h264_bitstream is part of the ABI: the reference's public header reads
 * it from inline functions (include/h264/h264_bitstream.h:31-60,117-317), so
 * callers compiled against either header must see the same layout and the same
 * inline semantics (EPB skip on fetch, one cached byte, MSB first).
 */
#ifndef H264B200_BITSTREAM_H
#define H264B200_BITSTREAM_H

struct h264_bitstream {
	union {
		const uint8_t *cdata; /* read side */
		uint8_t *data;        /* write side */
	};
	size_t len;  /* buffer length (capacity when writing) */
	size_t off;  /* next byte to fetch / to flush */
	uint8_t cache;     /* byte being consumed / assembled */
	uint8_t cachebits; /* unread bits left in cache / bits already assembled */
	int emulation_prevention; /* skip / insert 0x03 after 00 00 */
	int dynamic; /* write side: buffer is grown with realloc */
	void *priv;
};

/* next NAL unit of an Annex-B buffer: *start = first byte after the first start
 * code, *end = first 00 00 00 / 00 00 01 at or after it; -EAGAIN (with *end = len)
 * when the buffer ends first, -ENOENT without a start code */
H264_API int h264_find_nalu(const uint8_t *buf, size_t len, size_t *start, size_t *end);

H264_API int h264_bs_write_bits(struct h264_bitstream *bs, uint64_t v, uint32_t n);
H264_API int h264_bs_read_bits_ue(struct h264_bitstream *bs, uint32_t *v);
H264_API int h264_bs_write_bits_ue(struct h264_bitstream *bs, uint32_t v);
H264_API int h264_bs_read_bits_ff_coded(struct h264_bitstream *bs, uint32_t *v);
H264_API int h264_bs_write_bits_ff_coded(struct h264_bitstream *bs, uint32_t v);
H264_API int h264_bs_more_rbsp_data(const struct h264_bitstream *bs);
H264_API int h264_bs_next_bits(const struct h264_bitstream *bs, uint32_t *v, uint32_t n);
H264_API int h264_bs_read_rbsp_trailing_bits(struct h264_bitstream *bs);
H264_API int h264_bs_write_rbsp_trailing_bits(struct h264_bitstream *bs);
H264_API int h264_bs_read_raw_bytes(struct h264_bitstream *bs, uint8_t *buf, size_t len);
H264_API int h264_bs_write_raw_bytes(struct h264_bitstream *bs, const uint8_t *buf, size_t len);
H264_API int h264_bs_acquire_buf(struct h264_bitstream *bs, uint8_t **buf, size_t *len);

static inline void h264_bs_cinit(struct h264_bitstream *bs, const uint8_t *buf, size_t len,
				 int emulation_prevention)
{
	memset(bs, 0, sizeof(*bs));
	bs->cdata = buf;
	bs->len = len;
	bs->emulation_prevention = emulation_prevention;
}

static inline void h264_bs_init(struct h264_bitstream *bs, uint8_t *buf, size_t len,
				int emulation_prevention)
{
	memset(bs, 0, sizeof(*bs));
	bs->data = buf;
	bs->len = len;
	bs->dynamic = buf == NULL && len == 0;
	bs->emulation_prevention = emulation_prevention;
}

static inline void h264_bs_clear(struct h264_bitstream *bs)
{
	if (bs->dynamic)
		free(bs->data);
	memset(bs, 0, sizeof(*bs));
}

static inline int h264_bs_byte_aligned(const struct h264_bitstream *bs)
{
	return (bs->cachebits & 7) == 0;
}

static inline int h264_bs_eos(const struct h264_bitstream *bs)
{
	return bs->cachebits == 0 && bs->off >= bs->len;
}

static inline size_t h264_bs_rem_raw_bits(const struct h264_bitstream *bs)
{
	return 8 * (bs->len - bs->off) + bs->cachebits;
}

/* load the next payload byte into the cache, stepping over an emulation
 * prevention byte (a 03 that follows two raw zero bytes) */
static inline int h264_bs_fetch(struct h264_bitstream *bs)
{
	size_t o = bs->off;
	if (o >= bs->len)
		return -EIO;
	if (bs->emulation_prevention && o >= 2 && bs->cdata[o] == 0x03 && bs->cdata[o - 1] == 0x00 &&
	    bs->cdata[o - 2] == 0x00) {
		if (o + 1 >= bs->len)
			return -EIO;
		o++;
	}
	bs->cache = bs->cdata[o];
	bs->cachebits = 8;
	bs->off = o + 1;
	return 0;
}

static inline int h264_bs_read_bits(struct h264_bitstream *bs, uint32_t *v, uint32_t n)
{
	uint32_t acc = 0;
	int got = 0;
	*v = 0;
	while (n > 0) {
		if (bs->cachebits == 0 && h264_bs_fetch(bs) < 0)
			return -EIO;
		uint32_t take = n < bs->cachebits ? n : bs->cachebits;
		bs->cachebits -= take;
		acc = (acc << take) | ((bs->cache >> bs->cachebits) & ((1u << take) - 1));
		*v = acc;
		n -= take;
		got += (int)take;
	}
	return got;
}

static inline int h264_bs_read_bits_u(struct h264_bitstream *bs, uint32_t *v, uint32_t n)
{
	return h264_bs_read_bits(bs, v, n);
}

static inline int h264_bs_write_bits_u(struct h264_bitstream *bs, uint32_t v, uint32_t n)
{
	return h264_bs_write_bits(bs, v, n);
}

static inline int h264_bs_read_bits_i(struct h264_bitstream *bs, int32_t *v, uint32_t n)
{
	uint32_t u = 0;
	int res = h264_bs_read_bits(bs, &u, n);
	if (res >= 0) {
		if (u & (1u << (n - 1)))
			u |= ((uint32_t)-1) << n; /* sign extend */
		*v = (int32_t)u;
	}
	return res;
}

static inline int h264_bs_write_bits_i(struct h264_bitstream *bs, int32_t v, uint32_t n)
{
	return h264_bs_write_bits_u(bs, ((uint32_t)v) & ((1 << n) - 1), n);
}

/* 9.1 se(v): codeNum k -> (-1)^(k+1) * ceil(k / 2) */
static inline int h264_bs_read_bits_se(struct h264_bitstream *bs, int32_t *v)
{
	uint32_t k = 0;
	int res = h264_bs_read_bits_ue(bs, &k);
	if (res >= 0)
		*v = (k & 1) ? (((int32_t)k + 1) / 2) : (-((int32_t)k + 1) / 2);
	return res;
}

static inline int h264_bs_write_bits_se(struct h264_bitstream *bs, int32_t v)
{
	return h264_bs_write_bits_ue(bs, v <= 0 ? (uint32_t)(-2 * v) : (uint32_t)(2 * v - 1));
}

/* 9.1 te(v): a single inverted bit when the range is 0..1, ue(v) otherwise */
static inline int h264_bs_read_bits_te(struct h264_bitstream *bs, uint32_t *v, uint32_t m)
{
	if (m != 1)
		return h264_bs_read_bits_ue(bs, v);
	int res = h264_bs_read_bits(bs, v, 1);
	*v = !*v;
	return res;
}

static inline int h264_bs_write_bits_te(struct h264_bitstream *bs, uint32_t v, uint32_t m)
{
	return m == 1 ? h264_bs_write_bits(bs, !v, 1) : h264_bs_write_bits_ue(bs, v);
}

#endif /* H264B200_BITSTREAM_H */
