/*
 * h264_bits.c — the non-inline half of the libh264 bitstream API
 * (reference behaviour: src/h264_bitstream.c:30-451; declarations
 * include/h264/h264_bitstream.h:63-114).
 *
 * These are the single-NAL, bit-granular utilities the public header exposes to
 * callers (its read side is inline in the header).  They are host code in the
 * reference's API by construction; the BULK Annex-B paths of this library
 * (h264_reader_parse, the framing of many payloads) run on the GPU through
 * libh264gpu.so and never through these loops.
 */
#include "h264_priv.h"

/* ---- Annex-B, one NAL unit at a time (B.1; src/h264_bitstream.c:87-184) ---------- */

/* closed form of the reference's two byte loops:
 * start = 3 + first i with b[i..i+2] = 00 00 01; end = first q >= start with
 * b[q..q+2] = 00 00 0x, x <= 1 */
int h264_find_nalu(const uint8_t *buf, size_t len, size_t *start, size_t *end)
{
	if (buf == NULL || start == NULL || end == NULL)
		return -EINVAL;
	size_t i = 0;
	int found = 0;
	while (i + 2 < len) {
		/* jump over bytes that cannot be the third byte of a start code */
		if (buf[i + 2] > 1) {
			i += 3;
		} else if (buf[i] == 0 && buf[i + 1] == 0 && buf[i + 2] == 1) {
			found = 1;
			break;
		} else {
			i++;
		}
	}
	if (!found)
		return -ENOENT;
	*start = i + 3;
	size_t q = *start;
	while (q + 2 < len) {
		if (buf[q + 2] > 1) {
			q += 3;
		} else if (buf[q] == 0 && buf[q + 1] == 0) {
			*end = q;
			return 0;
		} else {
			q++;
		}
	}
	*end = len;
	return -EAGAIN;
}

/* ---- write side (src/h264_bitstream.c:30-81,211-239) ------------------------------ */

static int bs_reserve(struct h264_bitstream *bs, size_t want)
{
	if (want <= bs->len)
		return 0;
	if (!bs->dynamic)
		return -EIO;
	want = (want + 255) & ~(size_t)255; /* the reference grows in 256-byte steps */
	uint8_t *p = realloc(bs->data, want);
	if (p == NULL)
		return -ENOMEM;
	bs->data = p;
	bs->len = want;
	return 0;
}

/* a completed byte leaves the cache; 03 goes in front of it when the two previous
 * OUTPUT bytes are 00 00 and it is <= 3 */
static int bs_emit_byte(struct h264_bitstream *bs)
{
	const int escape = bs->emulation_prevention && bs->off >= 2 && bs->data[bs->off - 1] == 0 &&
			   bs->data[bs->off - 2] == 0 && bs->cache <= 3;
	int res = bs_reserve(bs, bs->off + 1 + escape);
	if (res < 0)
		return res;
	if (escape)
		bs->data[bs->off++] = 3;
	bs->data[bs->off++] = bs->cache;
	bs->cache = 0;
	bs->cachebits = 0;
	return 0;
}

int h264_bs_write_bits(struct h264_bitstream *bs, uint64_t v, uint32_t n)
{
	if (n > 64)
		return -EINVAL;
	int done = 0;
	while (n > 0) {
		uint32_t room = 8u - bs->cachebits;
		uint32_t take = n < room ? n : room;
		uint32_t part = (uint32_t)(v >> (n - take)) & ((1u << take) - 1);
		bs->cache |= (uint8_t)(part << (room - take));
		bs->cachebits += take;
		n -= take;
		done += (int)take;
		if (bs->cachebits == 8 && bs_emit_byte(bs) < 0)
			return -EIO;
	}
	return done;
}

/* ---- Exp-Golomb and friends (9.1; src/h264_bitstream.c:190-319) ------------------- */

int h264_bs_read_bits_ue(struct h264_bitstream *bs, uint32_t *v)
{
	int zeros = 0;
	uint32_t bit = 0, rest = 0;
	for (;;) {
		if (h264_bs_read_bits(bs, &bit, 1) < 0)
			return -EIO;
		if (bit)
			break;
		zeros++;
	}
	if (zeros && h264_bs_read_bits(bs, &rest, (uint32_t)zeros) < 0)
		return -EIO;
	/* validity domain: zeros < 32 (the reference shifts by it unchecked) */
	*v = (uint32_t)((1ull << zeros) - 1) + rest;
	return 2 * zeros + 1;
}

int h264_bs_write_bits_ue(struct h264_bitstream *bs, uint32_t v)
{
	if (v == 0)
		return h264_bs_write_bits(bs, 1, 1);
	const uint32_t code = v + 1;
	const uint32_t width = 32u - (uint32_t)__builtin_clz(code);
	return h264_bs_write_bits(bs, code, 2 * width - 1);
}

int h264_bs_read_bits_ff_coded(struct h264_bitstream *bs, uint32_t *v)
{
	int n = 0;
	uint32_t byte = 0;
	*v = 0;
	do {
		int res = h264_bs_read_bits(bs, &byte, 8);
		if (res < 0)
			return res;
		n += res;
		*v += byte;
	} while (byte == 0xff);
	return n;
}

int h264_bs_write_bits_ff_coded(struct h264_bitstream *bs, uint32_t v)
{
	int n = 0;
	uint32_t byte;
	do {
		byte = v > 0xff ? 0xff : v;
		int res = h264_bs_write_bits(bs, byte, 8);
		if (res < 0)
			return res;
		n += res;
		v -= byte;
	} while (byte == 0xff);
	return n;
}

/* ---- end-of-RBSP tests (7.2, 7.3.2.11; src/h264_bitstream.c:325-415) -------------- */

/* on a copy, stop bit then zeros to alignment; exactly one
 * trailing 00 byte after that still counts as "no more data" */
int h264_bs_more_rbsp_data(const struct h264_bitstream *bs)
{
	struct h264_bitstream t = *bs;
	uint32_t bit = 0;
	if (h264_bs_read_bits(&t, &bit, 1) < 0)
		return 0;
	if (bit != 1)
		return 1;
	while (!h264_bs_byte_aligned(&t)) {
		if (h264_bs_read_bits(&t, &bit, 1) < 0)
			return 0;
		if (bit)
			return 1;
	}
	if (h264_bs_eos(&t))
		return 0;
	return t.off + 1 < t.len || t.cdata[t.off] != 0;
}

int h264_bs_next_bits(const struct h264_bitstream *bs, uint32_t *v, uint32_t n)
{
	struct h264_bitstream t = *bs;
	return h264_bs_read_bits(&t, v, n);
}

int h264_bs_read_rbsp_trailing_bits(struct h264_bitstream *bs)
{
	uint32_t bit = 0;
	int res = h264_bs_read_bits(bs, &bit, 1);
	if (res < 0)
		return res;
	if (bit != 1)
		return -EIO;
	while (!h264_bs_byte_aligned(bs)) {
		res = h264_bs_read_bits(bs, &bit, 1);
		if (res < 0)
			return res;
		if (bit)
			return -EIO;
	}
	return 0;
}

int h264_bs_write_rbsp_trailing_bits(struct h264_bitstream *bs)
{
	int res = h264_bs_write_bits(bs, 1, 1);
	while (res >= 0 && !h264_bs_byte_aligned(bs))
		res = h264_bs_write_bits(bs, 0, 1);
	return res < 0 ? res : 0;
}

/* ---- raw bytes and buffer hand-over (src/h264_bitstream.c:418-451) ---------------- */

int h264_bs_read_raw_bytes(struct h264_bitstream *bs, uint8_t *buf, size_t len)
{
	if (!h264_bs_byte_aligned(bs) || bs->len - bs->off != len)
		return -EIO;
	memcpy(buf, bs->cdata + bs->off, len);
	bs->off += len;
	return 0;
}

int h264_bs_write_raw_bytes(struct h264_bitstream *bs, const uint8_t *buf, size_t len)
{
	if (!h264_bs_byte_aligned(bs))
		return -EIO;
	int res = bs_reserve(bs, bs->off + len);
	if (res < 0)
		return res;
	memcpy(bs->data + bs->off, buf, len);
	bs->off += len;
	return 0;
}

int h264_bs_acquire_buf(struct h264_bitstream *bs, uint8_t **buf, size_t *len)
{
	if (!h264_bs_byte_aligned(bs) || !bs->dynamic)
		return -EIO;
	*buf = bs->data;
	*len = bs->off;
	bs->dynamic = 0;
	return 0;
}
