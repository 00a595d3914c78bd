/*
 * synth.c — deterministic synthetic workload generator (host, multi-threaded).
 *
 * Produces the inputs BASELINE.json's configs name: RBSP payloads
 * with NAL sizes log-uniform in [lo,hi], bytes i.i.d. with P(00)=3/16, and the
 * Annex-B stream made of them (start codes 3 or 4 bytes, optional trailing zero
 * bytes, payloads escaped so the stream is valid).  This is workload generation
 * for bench.py and the tests, not part of the parse/serialise path; its output
 * is checked against the oracle writer in tests/test_synth.py.
 *
 * Content of NAL k depends only on (seed, k), so the result does not depend on
 * the number of threads.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t xs64(uint64_t *s)
{
	uint64_t x = *s;
	x ^= x << 13;
	x ^= x >> 7;
	x ^= x << 17;
	return *s = x;
}

static inline uint64_t mix(uint64_t a, uint64_t b)
{
	uint64_t x = a * 0x9E3779B97F4A7C15ull + b + 0x632BE59BD9B4E019ull;
	x ^= x >> 30;
	x *= 0xBF58476D1CE4E5B9ull;
	x ^= x >> 27;
	x *= 0x94D049BB133111EBull;
	x ^= x >> 31;
	return x ? x : 1;
}

/*
 * Payload sizes, log-uniform in [lo,hi], until their sum reaches target_bytes
 * (the last one is trimmed to land exactly).  offs gets n+1 entries.  Returns n
 * (or the number needed when cap is too small).
 */
uint64_t synth_sizes(uint64_t seed, uint64_t target_bytes, uint32_t lo, uint32_t hi,
		     uint64_t *offs, uint64_t cap)
{
	uint64_t s = mix(seed, 0x5151), n = 0, tot = 0;
	const double llo = log((double)lo), lhi = log((double)hi + 1.0);
	if (cap)
		offs[0] = 0;
	while (tot < target_bytes) {
		double u = (double)(xs64(&s) >> 11) * (1.0 / 9007199254740992.0);
		uint64_t sz = (uint64_t)exp(llo + u * (lhi - llo));
		if (sz < lo)
			sz = lo;
		if (sz > hi)
			sz = hi;
		if (tot + sz > target_bytes)
			sz = target_bytes - tot;
		tot += sz;
		n++;
		if (n < cap)
			offs[n] = tot;
	}
	return n;
}

struct job {
	uint64_t seed;
	const uint64_t *offs;
	uint64_t k0, k1;
	uint8_t *rbsp;
	/* framing */
	const uint64_t *out_off;
	uint8_t *out;
	uint64_t *esc_len;
	int mixed_sc, trailing;
};

/* byte distribution: P(00) = 3/16, the other 13/16 uniform over 1..255 */
static void fill_payload(uint64_t seed, uint64_t k, uint8_t *p, uint64_t n)
{
	uint64_t s = mix(seed, k);
	for (uint64_t i = 0; i < n; i++) {
		if ((i & 3) == 0)
			xs64(&s);
		uint32_t r = (uint32_t)(s >> (16 * (i & 3))) & 0xffff;
		p[i] = (r & 15) < 3 ? 0 : (uint8_t)(1 + ((r >> 4) * 255 >> 12));
	}
	if (n) /* legal NAL header: forbidden_zero_bit 0, nal_unit_type 1..23 */
		p[0] = (uint8_t)(((s >> 5) & 3) << 5 | (1 + (s >> 9) % 23));
	if (n >= 2) /* like rbsp_trailing_bits: a real RBSP never ends in 00 */
		p[n - 1] = 0x80;
}

static void *fill_worker(void *arg)
{
	struct job *j = arg;
	for (uint64_t k = j->k0; k < j->k1; k++)
		fill_payload(j->seed, k, j->rbsp + j->offs[k], j->offs[k + 1] - j->offs[k]);
	return NULL;
}

static int sc_len_of(uint64_t seed, uint64_t k, int mixed)
{
	return mixed ? (int)(3 + (mix(seed ^ 0xabcdef, k) & 1)) : 4;
}

static int trail_of(uint64_t seed, uint64_t k, int trailing)
{
	if (!trailing)
		return 0;
	uint64_t r = mix(seed ^ 0x7777, k);
	return (r & 3) == 0 ? (int)((r >> 2) % 3) : 0; /* 0..2 zero bytes on 1/4 of NALs */
}

/* escape one payload; out == NULL only counts */
static uint64_t escape(const uint8_t *p, uint64_t n, uint8_t *out)
{
	uint64_t o = 0;
	unsigned z = 0;
	for (uint64_t i = 0; i < n; i++) {
		uint8_t c = p[i];
		if (z == 2 && c <= 3) {
			if (out)
				out[o] = 3;
			o++;
			z = 0;
		}
		if (out)
			out[o] = c;
		o++;
		z = c == 0 ? z + 1 : 0;
	}
	return o;
}

static void *count_worker(void *arg)
{
	struct job *j = arg;
	for (uint64_t k = j->k0; k < j->k1; k++)
		j->esc_len[k] = escape(j->rbsp + j->offs[k], j->offs[k + 1] - j->offs[k], NULL);
	return NULL;
}

static void *frame_worker(void *arg)
{
	struct job *j = arg;
	for (uint64_t k = j->k0; k < j->k1; k++) {
		uint8_t *o = j->out + j->out_off[k];
		int sc = sc_len_of(j->seed, k, j->mixed_sc);
		if (sc == 4)
			*o++ = 0;
		*o++ = 0;
		*o++ = 0;
		*o++ = 1;
		o += escape(j->rbsp + j->offs[k], j->offs[k + 1] - j->offs[k], o);
		int tz = trail_of(j->seed, k, j->trailing);
		for (int i = 0; i < tz; i++)
			*o++ = 0;
	}
	return NULL;
}

static void run_jobs(void *(*fn)(void *), struct job *proto, uint64_t n, int nthreads)
{
	pthread_t th[256];
	struct job jobs[256];
	if (nthreads < 1)
		nthreads = 1;
	if (nthreads > 256)
		nthreads = 256;
	for (int i = 0; i < nthreads; i++) {
		jobs[i] = *proto;
		jobs[i].k0 = n * (uint64_t)i / nthreads;
		jobs[i].k1 = n * (uint64_t)(i + 1) / nthreads;
	}
	for (int i = 1; i < nthreads; i++)
		pthread_create(&th[i], NULL, fn, &jobs[i]);
	fn(&jobs[0]);
	for (int i = 1; i < nthreads; i++)
		pthread_join(th[i], NULL);
}

void synth_fill(uint64_t seed, const uint64_t *offs, uint64_t n, uint8_t *rbsp, int nthreads)
{
	struct job j;
	memset(&j, 0, sizeof(j));
	j.seed = seed;
	j.offs = offs;
	j.rbsp = rbsp;
	run_jobs(fill_worker, &j, n, nthreads);
}

/*
 * Frame + escape n payloads into out (capacity out_cap).  out_off gets n+1
 * entries (offset of each NAL's start code; [n] = total).  Returns the total, or
 * the size needed if out_cap is too small (nothing written then).
 */
uint64_t synth_frame(uint64_t seed, const uint8_t *rbsp, const uint64_t *offs, uint64_t n,
		     int mixed_sc, int trailing, uint8_t *out, uint64_t out_cap,
		     uint64_t *out_off, int nthreads)
{
	struct job j;
	memset(&j, 0, sizeof(j));
	j.seed = seed;
	j.offs = offs;
	j.rbsp = (uint8_t *)rbsp;
	j.mixed_sc = mixed_sc;
	j.trailing = trailing;
	uint64_t *esc = malloc((n + 1) * sizeof(uint64_t));
	j.esc_len = esc;
	run_jobs(count_worker, &j, n, nthreads);
	uint64_t tot = 0;
	for (uint64_t k = 0; k < n; k++) {
		out_off[k] = tot;
		tot += (uint64_t)sc_len_of(seed, k, mixed_sc) + esc[k] +
		       (uint64_t)trail_of(seed, k, trailing);
	}
	out_off[n] = tot;
	free(esc);
	if (tot > out_cap)
		return tot;
	j.out = out;
	j.out_off = out_off;
	run_jobs(frame_worker, &j, n, nthreads);
	return tot;
}
