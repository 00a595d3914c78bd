/*
 * oracle_annexb.c — CPU restatement of libh264's Annex-B framing path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked, imported or
 * executed by the product path (libh264_b200/, include/); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may use it, and only as the checker or the reported CPU baseline.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle.py)
 * against (a) the known answers of SURVEY.md Appendix A.4, captured from the
 * compiled reference, and (b) the compiled reference itself
 * (oracle/_ref/libh264_ref.so, built by oracle/Makefile from /root/reference)
 * on seeded random inputs.
 *
 * Each function is a closed-form / state-machine restatement, written from the
 * behaviour of the reference, not a transcription of its loops:
 *
 *   oracle_scan         <- h264_reader_parse's NAL loop, src/h264_reader.c:133-140,
 *                          over h264_find_nalu, src/h264_bitstream.c:159-184
 *                          (h264_find_start_code :88-120, h264_find_end_code :125-154)
 *   oracle_strip        <- h264_bs_fetch, include/h264/h264_bitstream.h:168-190,
 *                          driven 8 bits at a time by h264_bs_read_bits :194-218
 *   oracle_insert       <- h264_bs_flush, src/h264_bitstream.c:54-81, driven by
 *                          h264_bs_write_bits(...,8) :211-239
 *   oracle_split_strip  <- the two above fused: what the CUDA scan+strip kernel emits
 *   oracle_frame        <- oracle_insert per payload + a start code in front of each
 *                          (the reference leaves framing to the caller,
 *                          src/h264_writer.c:240-243; 4-byte codes are what
 *                          h264_avcc_to_byte_stream writes, src/h264.c:251-272)
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* zz(i): 00 00 0x with x<=1 fully inside the buffer (i+2 < len). */
static int zz_at(const uint8_t *b, size_t len, size_t i)
{
	return i + 2 < len && b[i] == 0 && b[i + 1] == 0 && b[i + 2] <= 1;
}

/*
 * NAL table of an Annex-B buffer, exactly as the reference's reader loop visits it.
 * starts[k]/ends[k] are absolute byte offsets (NAL k = buf[starts[k], ends[k])).
 * Returns the number of NALs found (entries beyond cap are counted, not stored).
 * *final_off gets the value h264_reader_parse leaves in *off.
 * *last_open is 1 when the last NAL ran to the end of the buffer (-EAGAIN case).
 */
size_t oracle_scan(const uint8_t *b, size_t len, uint64_t *starts, uint64_t *ends,
		   size_t cap, uint64_t *final_off, int *last_open)
{
	size_t n = 0, i = 0;
	uint64_t off = 0;
	int open = 0;

	while (i + 2 < len) {
		/* next start code at or after i */
		if (!(b[i] == 0 && b[i + 1] == 0 && b[i + 2] == 1)) {
			i++;
			continue;
		}
		size_t s = i + 3, e = s;
		while (e < len && !zz_at(b, len, e))
			e++;
		open = (e >= len);
		if (open)
			e = len;
		if (n < cap) {
			starts[n] = s;
			ends[n] = e;
		}
		n++;
		off = e;
		i = e;
	}
	if (final_off)
		*final_off = n ? off : 0;
	if (last_open)
		*last_open = n ? open : 0;
	return n;
}

/*
 * RBSP bytes of one NAL (escaped bytes without start code): the sequence that
 * h264_bs_read_bits(bs,&v,8) yields on an emulation_prevention=1 stream until it
 * fails.  Returns the number of bytes written to out (out may alias nothing;
 * capacity len is always enough).
 */
size_t oracle_strip(const uint8_t *nal, size_t len, uint8_t *out)
{
	size_t o = 0;
	for (size_t i = 0; i < len; i++) {
		if (i >= 2 && nal[i] == 3 && nal[i - 1] == 0 && nal[i - 2] == 0)
			continue; /* emulation prevention byte: dropped */
		out[o++] = nal[i];
	}
	return o;
}

/*
 * Escaped bytes of one payload: what a dynamic emulation_prevention=1 bitstream
 * holds after h264_bs_write_bits(bs, byte, 8) for every input byte.
 * Worst case output is len + len/2 (one 03 per two zeros) -> capacity 3*len/2+1.
 */
size_t oracle_insert(const uint8_t *rbsp, size_t len, uint8_t *out)
{
	size_t o = 0;
	unsigned zeros = 0; /* trailing 00 bytes already in the output (0..2) */
	for (size_t i = 0; i < len; i++) {
		uint8_t c = rbsp[i];
		if (zeros == 2 && c <= 3) {
			out[o++] = 3;
			zeros = 0;
		}
		out[o++] = c;
		zeros = (c == 0) ? zeros + 1 : 0;
	}
	return o;
}

/* Number of bytes oracle_insert would add (no output). */
size_t oracle_insert_count(const uint8_t *rbsp, size_t len)
{
	size_t add = 0;
	unsigned zeros = 0;
	for (size_t i = 0; i < len; i++) {
		uint8_t c = rbsp[i];
		if (zeros == 2 && c <= 3) {
			add++;
			zeros = 0;
		}
		zeros = (c == 0) ? zeros + 1 : 0;
	}
	return add;
}

/*
 * Fused split + strip over a whole Annex-B buffer: NAL table plus the
 * concatenated RBSP of every NAL.  rbsp_off[k] is the offset of NAL k's RBSP in
 * rbsp (rbsp_off has cap+1 usable entries; entry n is the total).  Returns n.
 */
size_t oracle_split_strip(const uint8_t *b, size_t len, uint64_t *starts,
			  uint64_t *ends, uint64_t *rbsp_off, size_t cap,
			  uint8_t *rbsp, uint64_t *rbsp_total, uint64_t *final_off)
{
	size_t n = oracle_scan(b, len, starts, ends, cap, final_off, NULL);
	size_t m = n < cap ? n : cap;
	uint64_t o = 0;
	for (size_t k = 0; k < m; k++) {
		rbsp_off[k] = o;
		o += oracle_strip(b + starts[k], (size_t)(ends[k] - starts[k]), rbsp + o);
	}
	if (m <= cap)
		rbsp_off[m] = o;
	if (rbsp_total)
		*rbsp_total = o;
	return n;
}

/*
 * Writer-side framing: for payload k = rbsp[off[k], off[k+1]) emit a start code
 * (sc_len = 3 or 4 bytes) followed by the escaped payload.  out_off[k] = offset
 * of NAL k's start code in out; out_off[n] = total.  Returns total bytes.
 */
size_t oracle_frame(const uint8_t *rbsp, const uint64_t *off, size_t n, int sc_len,
		    uint8_t *out, uint64_t *out_off)
{
	size_t o = 0;
	for (size_t k = 0; k < n; k++) {
		if (out_off)
			out_off[k] = o;
		if (sc_len == 4)
			out[o++] = 0;
		out[o++] = 0;
		out[o++] = 0;
		out[o++] = 1;
		o += oracle_insert(rbsp + off[k], (size_t)(off[k + 1] - off[k]), out + o);
	}
	if (out_off)
		out_off[n] = o;
	return o;
}
