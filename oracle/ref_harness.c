/*
 * ref_harness.c — thin driver around the UNMODIFIED reference (libh264) so tests
 * and the CPU-baseline leg of bench.py can call it through ctypes.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled together with the reference sources (read
 * in place from /root/reference, never copied) into oracle/_ref/libh264_ref.so
 * by oracle/Makefile.  All work is done by reference functions; this file only
 * loops over them the way the reference's own callers do:
 *
 *   ref_scan    = the NAL loop of h264_reader_parse (src/h264_reader.c:133-140)
 *                 with h264_reader_parse_nalu replaced by "record (start,end)"
 *   ref_strip   = h264_bs_cinit(..,1) + h264_bs_read_bits(bs,&v,8) until failure
 *   ref_insert  = h264_bs_init(NULL,0,1) + h264_bs_write_bits(bs,byte,8) per byte
 *   ref_mt_*    = the same loops on T threads over byte ranges / NAL ranges
 *                 (the library is re-entrant per object)
 */
#include "h264_priv.h"

#include <pthread.h>
#include <time.h>

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + ts.tv_nsec * 1e-9;
}

size_t ref_scan(const uint8_t *buf, size_t len, uint64_t *starts, uint64_t *ends,
		size_t cap, uint64_t *final_off)
{
	size_t off = 0, start = 0, end = 0, n = 0;
	while (off < len) {
		int res = h264_find_nalu(buf + off, len - off, &start, &end);
		if (res < 0 && res != -EAGAIN)
			break;
		if (n < cap) {
			starts[n] = off + start;
			ends[n] = off + end;
		}
		n++;
		off += end;
	}
	if (final_off)
		*final_off = off;
	return n;
}

size_t ref_strip(const uint8_t *nal, size_t len, uint8_t *out, uint64_t *raw_off_at_fail)
{
	struct h264_bitstream bs;
	uint32_t v = 0;
	size_t o = 0;
	h264_bs_cinit(&bs, nal, len, 1);
	while (h264_bs_read_bits(&bs, &v, 8) == 8)
		out[o++] = (uint8_t)v;
	if (raw_off_at_fail)
		*raw_off_at_fail = bs.off;
	return o;
}

/* returns escaped length, or (size_t)-1 if cap is too small */
size_t ref_insert(const uint8_t *rbsp, size_t len, uint8_t *out, size_t cap)
{
	struct h264_bitstream bs;
	size_t n;
	h264_bs_init(&bs, NULL, 0, 1);
	for (size_t i = 0; i < len; i++)
		h264_bs_write_bits(&bs, rbsp[i], 8);
	n = bs.off;
	if (n <= cap && n > 0)
		memcpy(out, bs.data, n);
	h264_bs_clear(&bs);
	return n <= cap ? n : (size_t)-1;
}

/* ------------------------------------------------------------------------- */
/* Multi-threaded CPU baselines (timed inside, seconds returned).             */

struct mt_job {
	const uint8_t *buf;
	size_t len;
	size_t lo, hi; /* this thread owns NALs whose start code begins in [lo,hi) */
	uint8_t *out;  /* scratch, at least hi-lo+8 bytes, or NULL for scan-only */
	uint64_t nals, rbsp_bytes, in_bytes;
};

static void *mt_split_strip_worker(void *arg)
{
	struct mt_job *j = arg;
	size_t off = j->lo, start = 0, end = 0;
	/* back up so a start code straddling lo is seen by exactly one thread:
	 * a code is owned by the range holding its first 00 of "00 00 01" */
	while (off < j->len) {
		int res = h264_find_nalu(j->buf + off, j->len - off, &start, &end);
		if (res < 0 && res != -EAGAIN)
			break;
		/* first byte of the 3-byte code */
		size_t sc = off + start - 3;
		if (sc >= j->hi)
			break;
		j->nals++;
		j->in_bytes += end - start;
		if (j->out != NULL) {
			struct h264_bitstream bs;
			uint32_t v = 0;
			size_t o = 0;
			h264_bs_cinit(&bs, j->buf + off + start, end - start, 1);
			while (h264_bs_read_bits(&bs, &v, 8) == 8)
				j->out[o++] = (uint8_t)v;
			j->rbsp_bytes += o;
		}
		off += end;
	}
	return NULL;
}

/*
 * Scan (+ strip when do_strip) of buf on nthreads threads, byte range i of
 * nthreads each.  scratch must hold len bytes when do_strip.  Returns seconds.
 */
double ref_mt_split_strip(const uint8_t *buf, size_t len, int nthreads, int do_strip,
			  uint8_t *scratch, uint64_t *nals, uint64_t *rbsp_bytes)
{
	pthread_t th[256];
	struct mt_job jobs[256];
	double t0, t1;
	if (nthreads < 1)
		nthreads = 1;
	if (nthreads > 256)
		nthreads = 256;
	for (int i = 0; i < nthreads; i++) {
		memset(&jobs[i], 0, sizeof(jobs[i]));
		jobs[i].buf = buf;
		jobs[i].len = len;
		jobs[i].lo = (size_t)((unsigned __int128)len * i / nthreads);
		jobs[i].hi = (size_t)((unsigned __int128)len * (i + 1) / nthreads);
		jobs[i].out = do_strip ? scratch + jobs[i].lo : NULL;
	}
	t0 = now_s();
	for (int i = 1; i < nthreads; i++)
		pthread_create(&th[i], NULL, mt_split_strip_worker, &jobs[i]);
	mt_split_strip_worker(&jobs[0]);
	for (int i = 1; i < nthreads; i++)
		pthread_join(th[i], NULL);
	t1 = now_s();
	if (nals) {
		*nals = 0;
		for (int i = 0; i < nthreads; i++)
			*nals += jobs[i].nals;
	}
	if (rbsp_bytes) {
		*rbsp_bytes = 0;
		for (int i = 0; i < nthreads; i++)
			*rbsp_bytes += jobs[i].rbsp_bytes;
	}
	return t1 - t0;
}

struct mt_ins_job {
	const uint8_t *rbsp;
	const uint64_t *off;
	size_t k0, k1;
	uint64_t out_bytes;
};

static void *mt_insert_worker(void *arg)
{
	struct mt_ins_job *j = arg;
	for (size_t k = j->k0; k < j->k1; k++) {
		struct h264_bitstream bs;
		h264_bs_init(&bs, NULL, 0, 1);
		for (uint64_t i = j->off[k]; i < j->off[k + 1]; i++)
			h264_bs_write_bits(&bs, j->rbsp[i], 8);
		j->out_bytes += bs.off;
		h264_bs_clear(&bs);
	}
	return NULL;
}

/* EPB insertion of n payloads on nthreads threads (payload ranges). Seconds. */
double ref_mt_insert(const uint8_t *rbsp, const uint64_t *off, size_t n, int nthreads,
		     uint64_t *out_bytes)
{
	pthread_t th[256];
	struct mt_ins_job jobs[256];
	double t0, t1;
	if (nthreads < 1)
		nthreads = 1;
	if (nthreads > 256)
		nthreads = 256;
	for (int i = 0; i < nthreads; i++) {
		jobs[i].rbsp = rbsp;
		jobs[i].off = off;
		jobs[i].k0 = n * i / nthreads;
		jobs[i].k1 = n * (i + 1) / nthreads;
		jobs[i].out_bytes = 0;
	}
	t0 = now_s();
	for (int i = 1; i < nthreads; i++)
		pthread_create(&th[i], NULL, mt_insert_worker, &jobs[i]);
	mt_insert_worker(&jobs[0]);
	for (int i = 1; i < nthreads; i++)
		pthread_join(th[i], NULL);
	t1 = now_s();
	if (out_bytes) {
		*out_bytes = 0;
		for (int i = 0; i < nthreads; i++)
			*out_bytes += jobs[i].out_bytes;
	}
	return t1 - t0;
}
