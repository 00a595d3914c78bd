/*
 * host_harness.c — drives ANY libh264-compatible shared library (this repo's
 * libh264_b200/libh264.so, or the compiled reference oracle/_ref/libh264_ref.so)
 * through the PUBLIC API only, so the two can be compared call for call.
 *
 * TEST INFRASTRUCTURE ONLY.
 *   hh_trace   h264_reader_parse / per-NAL h264_reader_parse_nalu with recording
 *              callbacks: every callback appends (tag, payload bytes) to a log
 *   hh_gen     builds a stream of seeded-random but valid SPS / PPS / AUD / SEI / slice
 *              header NAL units with the library's WRITER (h264_ctx_set_*, h264_write_nalu)
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <errno.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "h264/h264.h"

struct api {
	void *h;
	int (*reader_new)(const struct h264_ctx_cbs *, void *, struct h264_reader **);
	int (*reader_destroy)(struct h264_reader *);
	int (*reader_parse)(struct h264_reader *, uint32_t, const uint8_t *, size_t, size_t *);
	int (*reader_parse_nalu)(struct h264_reader *, uint32_t, const uint8_t *, size_t);
	int (*find_nalu)(const uint8_t *, size_t, size_t *, size_t *);
	int (*ctx_new)(struct h264_ctx **);
	int (*ctx_destroy)(struct h264_ctx *);
	int (*ctx_clear_nalu)(struct h264_ctx *);
	int (*ctx_set_nalu_header)(struct h264_ctx *, const struct h264_nalu_header *);
	int (*ctx_set_sps)(struct h264_ctx *, const struct h264_sps *);
	int (*ctx_set_pps)(struct h264_ctx *, const struct h264_pps *);
	int (*ctx_set_aud)(struct h264_ctx *, const struct h264_aud *);
	int (*ctx_set_filler)(struct h264_ctx *, size_t);
	int (*ctx_add_sei)(struct h264_ctx *, const struct h264_sei *);
	int (*ctx_set_slice_header)(struct h264_ctx *, const struct h264_slice_header *);
	int (*ctx_get_info)(struct h264_ctx *, struct h264_info *);
	int (*write_nalu)(struct h264_bitstream *, struct h264_ctx *);
	int (*write_grey_i_slice)(struct h264_bitstream *, struct h264_ctx *, uint32_t);
	int (*write_skipped_p_slice)(struct h264_bitstream *, struct h264_ctx *, uint32_t);
	int (*bs_write_bits)(struct h264_bitstream *, uint64_t, uint32_t);
	int (*bs_acquire_buf)(struct h264_bitstream *, uint8_t **, size_t *);
	int (*get_info)(const uint8_t *, size_t, const uint8_t *, size_t, struct h264_info *);
	int (*rewrite_slice_header)(struct h264_bitstream *, struct h264_ctx *, const struct h264_slice_header *);
};

#define SYM(field, name)                                                                       \
	do {                                                                                   \
		*(void **)&a->field = dlsym(a->h, name);                                       \
		if (a->field == NULL) {                                                        \
			fprintf(stderr, "host_harness: %s lacks %s\n", path, name);            \
			return -ENOENT;                                                        \
		}                                                                              \
	} while (0)

static int api_open(struct api *a, const char *path)
{
	memset(a, 0, sizeof(*a));
	a->h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
	if (a->h == NULL) {
		fprintf(stderr, "host_harness: dlopen %s: %s\n", path, dlerror());
		return -ENOENT;
	}
	SYM(reader_new, "h264_reader_new");
	SYM(reader_destroy, "h264_reader_destroy");
	SYM(reader_parse, "h264_reader_parse");
	SYM(reader_parse_nalu, "h264_reader_parse_nalu");
	SYM(find_nalu, "h264_find_nalu");
	SYM(ctx_new, "h264_ctx_new");
	SYM(ctx_destroy, "h264_ctx_destroy");
	SYM(ctx_clear_nalu, "h264_ctx_clear_nalu");
	SYM(ctx_set_nalu_header, "h264_ctx_set_nalu_header");
	SYM(ctx_set_sps, "h264_ctx_set_sps");
	SYM(ctx_set_pps, "h264_ctx_set_pps");
	SYM(ctx_set_aud, "h264_ctx_set_aud");
	SYM(ctx_set_filler, "h264_ctx_set_filler");
	SYM(ctx_add_sei, "h264_ctx_add_sei");
	SYM(ctx_set_slice_header, "h264_ctx_set_slice_header");
	SYM(ctx_get_info, "h264_ctx_get_info");
	SYM(write_nalu, "h264_write_nalu");
	SYM(write_grey_i_slice, "h264_write_grey_i_slice");
	SYM(write_skipped_p_slice, "h264_write_skipped_p_slice");
	SYM(bs_write_bits, "h264_bs_write_bits");
	SYM(bs_acquire_buf, "h264_bs_acquire_buf");
	SYM(get_info, "h264_get_info");
	SYM(rewrite_slice_header, "h264_rewrite_slice_header");
	return 0;
}

/* ---- trace ------------------------------------------------------------------------------- */

enum { T_NALU_BEGIN = 1, T_NALU_END, T_AU_END, T_SPS, T_PPS, T_SLICE, T_SD_BEGIN, T_SD_END, T_SD_MB,
       T_AUD, T_SEI, T_SEI_BP, T_SEI_PT, T_SEI_PAN, T_SEI_FILLER, T_SEI_UDR, T_SEI_UDU, T_SEI_RP,
       T_RESULT, T_INFO };

struct trace {
	struct api *a;
	const uint8_t *base;
	uint8_t *log;
	size_t cap, used;
	int overflow;
};

static void put(struct trace *t, uint32_t tag, const void *p, size_t n)
{
	const uint32_t hdr[2] = {tag, (uint32_t)n};
	if (t->used + sizeof(hdr) + n > t->cap) {
		t->overflow = 1;
		return;
	}
	memcpy(t->log + t->used, hdr, sizeof(hdr));
	if (n)
		memcpy(t->log + t->used + sizeof(hdr), p, n);
	t->used += sizeof(hdr) + n;
}

static void cb_nalu(struct trace *t, uint32_t tag, enum h264_nalu_type type, const uint8_t *buf, size_t len,
		    const struct h264_nalu_header *nh)
{
	uint64_t rec[6] = {(uint64_t)type, (uint64_t)(buf - t->base), len, nh->forbidden_zero_bit,
			   nh->nal_ref_idc, nh->nal_unit_type};
	put(t, tag, rec, sizeof(rec));
}
static void cb_nalu_begin(struct h264_ctx *c, enum h264_nalu_type type, const uint8_t *buf, size_t len,
			  const struct h264_nalu_header *nh, void *u)
{
	(void)c;
	cb_nalu(u, T_NALU_BEGIN, type, buf, len, nh);
}
static void cb_nalu_end(struct h264_ctx *c, enum h264_nalu_type type, const uint8_t *buf, size_t len,
			const struct h264_nalu_header *nh, void *u)
{
	struct trace *t = u;
	cb_nalu(t, T_NALU_END, type, buf, len, nh);
	struct h264_info info;
	memset(&info, 0, sizeof(info));
	int r = t->a->ctx_get_info(c, &info);
	if (r == 0)
		put(t, T_INFO, &info, sizeof(info));
}
static void cb_au_end(struct h264_ctx *c, void *u)
{
	(void)c;
	put(u, T_AU_END, NULL, 0);
}
static void cb_sps(struct h264_ctx *c, const uint8_t *buf, size_t len, const struct h264_sps *s, void *u)
{
	(void)c;
	(void)buf;
	(void)len;
	put(u, T_SPS, s, sizeof(*s));
}
static void cb_pps(struct h264_ctx *c, const uint8_t *buf, size_t len, const struct h264_pps *p, void *u)
{
	(void)c;
	(void)buf;
	(void)len;
	put(u, T_PPS, p, sizeof(*p));
}
static void cb_aud(struct h264_ctx *c, const uint8_t *buf, size_t len, const struct h264_aud *a, void *u)
{
	(void)c;
	(void)buf;
	(void)len;
	put(u, T_AUD, a, sizeof(*a));
}
static void cb_slice(struct h264_ctx *c, const uint8_t *buf, size_t len, const struct h264_slice_header *sh,
		     void *u)
{
	(void)c;
	(void)buf;
	(void)len;
	put(u, T_SLICE, sh, sizeof(*sh));
}
static void cb_sd_begin(struct h264_ctx *c, const struct h264_slice_header *sh, void *u)
{
	(void)c;
	put(u, T_SD_BEGIN, &sh->first_mb_in_slice, 4);
}
static void cb_sd_end(struct h264_ctx *c, const struct h264_slice_header *sh, uint32_t n, void *u)
{
	(void)c;
	(void)sh;
	put(u, T_SD_END, &n, 4);
}
static void cb_sd_mb(struct h264_ctx *c, const struct h264_slice_header *sh, uint32_t addr, enum h264_mb_type ty,
		     void *u)
{
	(void)c;
	(void)sh;
	uint32_t rec[2] = {addr, (uint32_t)ty};
	put(u, T_SD_MB, rec, sizeof(rec));
}
static void cb_sei(struct h264_ctx *c, enum h264_sei_type type, const uint8_t *buf, size_t len, void *u)
{
	(void)c;
	uint32_t ty = (uint32_t)type;
	put(u, T_SEI, &ty, 4);
	put(u, T_SEI, buf, len);
}
#define SEI_CB(name, tag, type)                                                                \
	static void cb_##name(struct h264_ctx *c, const uint8_t *buf, size_t len, const type *s, void *u) \
	{                                                                                      \
		(void)c;                                                                       \
		(void)buf;                                                                     \
		(void)len;                                                                     \
		put(u, tag, s, sizeof(*s));                                                    \
	}
SEI_CB(sei_bp, T_SEI_BP, struct h264_sei_buffering_period)
SEI_CB(sei_pt, T_SEI_PT, struct h264_sei_pic_timing)
SEI_CB(sei_pan, T_SEI_PAN, struct h264_sei_pan_scan_rect)
SEI_CB(sei_rp, T_SEI_RP, struct h264_sei_recovery_point)
static void cb_sei_filler(struct h264_ctx *c, const uint8_t *buf, size_t len,
			  const struct h264_sei_filler_payload *s, void *u)
{
	(void)c;
	uint64_t rec[2] = {(uint64_t)(s->buf - buf), s->len};
	(void)len;
	put(u, T_SEI_FILLER, rec, sizeof(rec));
}
static void cb_sei_udr(struct h264_ctx *c, const uint8_t *buf, size_t len,
		       const struct h264_sei_user_data_registered *s, void *u)
{
	(void)c;
	(void)len;
	uint64_t rec[4] = {s->country_code, s->country_code_extension_byte, (uint64_t)(s->buf - buf), s->len};
	put(u, T_SEI_UDR, rec, sizeof(rec));
}
static void cb_sei_udu(struct h264_ctx *c, const uint8_t *buf, size_t len,
		       const struct h264_sei_user_data_unregistered *s, void *u)
{
	(void)c;
	(void)len;
	uint64_t rec[4];
	memcpy(rec, s->uuid, 16);
	rec[2] = (uint64_t)(s->buf - buf);
	rec[3] = s->len;
	put(u, T_SEI_UDU, rec, sizeof(rec));
}

/* mode 0: h264_reader_parse on the whole buffer; mode 1: split with the library's own
 * h264_find_nalu and h264_reader_parse_nalu per NAL unit (return values logged) */
int hh_trace(const char *libpath, const uint8_t *buf, size_t len, uint32_t flags, int mode, uint8_t *log,
	     size_t cap, size_t *used)
{
	struct api a;
	int r = api_open(&a, libpath);
	if (r < 0)
		return r;
	struct trace t = {&a, buf, log, cap, 0, 0};
	struct h264_ctx_cbs cbs;
	memset(&cbs, 0, sizeof(cbs));
	cbs.au_end = cb_au_end;
	cbs.nalu_begin = cb_nalu_begin;
	cbs.nalu_end = cb_nalu_end;
	cbs.slice = cb_slice;
	cbs.slice_data_begin = cb_sd_begin;
	cbs.slice_data_end = cb_sd_end;
	cbs.slice_data_mb = cb_sd_mb;
	cbs.sps = cb_sps;
	cbs.pps = cb_pps;
	cbs.aud = cb_aud;
	cbs.sei = cb_sei;
	cbs.sei_buffering_period = cb_sei_bp;
	cbs.sei_pic_timing = cb_sei_pt;
	cbs.sei_pan_scan_rect = cb_sei_pan;
	cbs.sei_filler_payload = cb_sei_filler;
	cbs.sei_user_data_registered = cb_sei_udr;
	cbs.sei_user_data_unregistered = cb_sei_udu;
	cbs.sei_recovery_point = cb_sei_rp;
	struct h264_reader *rd = NULL;
	r = a.reader_new(&cbs, &t, &rd);
	if (r < 0)
		return r;
	if (mode == 0) {
		size_t off = 0;
		int64_t res[2];
		res[0] = a.reader_parse(rd, flags, buf, len, &off);
		res[1] = (int64_t)off;
		put(&t, T_RESULT, res, sizeof(res));
	} else {
		size_t off = 0, start = 0, end = 0;
		while (off < len) {
			int fr = a.find_nalu(buf + off, len - off, &start, &end);
			if (fr < 0 && fr != -EAGAIN)
				break;
			int64_t res[2];
			res[0] = a.reader_parse_nalu(rd, flags, buf + off + start, end - start);
			res[1] = (int64_t)(off + start);
			put(&t, T_RESULT, res, sizeof(res));
			off += end;
		}
	}
	a.reader_destroy(rd);
	*used = t.used;
	dlclose(a.h);
	return t.overflow ? -ENOBUFS : 0;
}

/* ---- A12: the macroblock -> slice group map of every slice (this library's extension
 * h264_ctx_get_slice_group_map), NAL by NAL without slice data: no GPU involved ------------- */
struct gm_out {
	int (*get)(const struct h264_ctx *, uint8_t *, size_t);
	uint8_t *out;
	size_t cap, used;
	int err;
};
static void gm_slice(struct h264_ctx *c, const uint8_t *buf, size_t len, const struct h264_slice_header *sh,
		     void *u)
{
	struct gm_out *g = u;
	(void)buf;
	(void)len;
	(void)sh;
	int n = g->get(c, g->out + g->used, g->cap - g->used);
	if (n < 0)
		g->err = n;
	else
		g->used += (size_t)n;
}
int hh_group_maps(const char *libpath, const uint8_t *buf, size_t len, uint8_t *out, size_t cap, size_t *used)
{
	struct api a;
	int r = api_open(&a, libpath);
	if (r < 0)
		return r;
	struct gm_out g = {NULL, out, cap, 0, 0};
	*(void **)&g.get = dlsym(a.h, "h264_ctx_get_slice_group_map");
	if (g.get == NULL)
		return -ENOSYS;
	struct h264_ctx_cbs cbs;
	memset(&cbs, 0, sizeof(cbs));
	cbs.slice = gm_slice;
	struct h264_reader *rd = NULL;
	r = a.reader_new(&cbs, &g, &rd);
	if (r < 0)
		return r;
	size_t off = 0, start = 0, end = 0;
	while (off < len) {
		int fr = a.find_nalu(buf + off, len - off, &start, &end);
		if (fr < 0 && fr != -EAGAIN)
			break;
		a.reader_parse_nalu(rd, 0, buf + off + start, end - start);
		off += end;
	}
	a.reader_destroy(rd);
	*used = g.used;
	dlclose(a.h);
	return g.err;
}

/* ---- timing: h264_reader_parse with counting callbacks ------------------------------------
 * counts: [0] nalu_begin  [1] slices  [2] slice_data_mb  [3] sps + pps  [4] sum of mb_addr ^ mb_type
 * (so that the macroblock callbacks cannot be optimised into nothing and both libraries can be
 * checked to have delivered the same thing).  Returns the best wall time of `reps` calls after
 * one warm-up call, or a negative errno. */
#include <time.h>
struct counts {
	uint64_t c[5];
};
static void ct_nalu_begin(struct h264_ctx *c, enum h264_nalu_type type, const uint8_t *buf, size_t len,
			  const struct h264_nalu_header *nh, void *u)
{
	((struct counts *)u)->c[0]++;
}
static void ct_slice(struct h264_ctx *c, const uint8_t *buf, size_t len, const struct h264_slice_header *sh,
		     void *u)
{
	((struct counts *)u)->c[1]++;
}
static void ct_sd_mb(struct h264_ctx *c, const struct h264_slice_header *sh, uint32_t addr, enum h264_mb_type ty,
		     void *u)
{
	struct counts *k = u;
	k->c[2]++;
	k->c[4] += addr ^ ((uint32_t)ty << 20);
}
static void ct_sps(struct h264_ctx *c, const uint8_t *buf, size_t len, const struct h264_sps *s, void *u)
{
	((struct counts *)u)->c[3]++;
}
static void ct_pps(struct h264_ctx *c, const uint8_t *buf, size_t len, const struct h264_pps *p, void *u)
{
	((struct counts *)u)->c[3]++;
}

double hh_time_parse(const char *libpath, const uint8_t *buf, size_t len, uint32_t flags, int reps,
		     uint64_t *counts_out)
{
	struct api a;
	int r = api_open(&a, libpath);
	if (r < 0)
		return (double)r;
	struct h264_ctx_cbs cbs;
	memset(&cbs, 0, sizeof(cbs));
	cbs.nalu_begin = ct_nalu_begin;
	cbs.slice = ct_slice;
	cbs.slice_data_mb = ct_sd_mb;
	cbs.sps = ct_sps;
	cbs.pps = ct_pps;
	struct counts k;
	struct h264_reader *rd = NULL;
	r = a.reader_new(&cbs, &k, &rd);
	if (r < 0)
		return (double)r;
	double best = -1;
	for (int i = 0; i <= reps; i++) {
		struct timespec t0, t1;
		size_t off = 0;
		memset(&k, 0, sizeof(k));
		clock_gettime(CLOCK_MONOTONIC, &t0);
		r = a.reader_parse(rd, flags, buf, len, &off);
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (r < 0) {
			a.reader_destroy(rd);
			return (double)r;
		}
		const double dt = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
		if (i > 0 && (best < 0 || dt < best))
			best = dt;
	}
	if (counts_out)
		memcpy(counts_out, k.c, sizeof(k.c));
	a.reader_destroy(rd);
	dlclose(a.h);
	return best;
}

/* ---- h264_rewrite_slice_header: patch every slice of a stream in place ---------------------
 * mode 0: frame_num ^ 1 (a fixed-width field: same bit length, must succeed and change bytes)
 * mode 1: slice_qp_delta + 17 (an se(v) that grows: must fail with -EPROTO and leave the NAL and
 *         the context's slice header untouched)
 * `out` receives a copy of the stream with the patched NALs; rcs[k] = return code for slice k. */
struct rw_state {
	struct api *a;
	const uint8_t *base;
	uint8_t *out;
	int mode;
	int32_t *rcs;
	uint32_t n, cap;
};

static void rw_slice(struct h264_ctx *c, const uint8_t *buf, size_t len, const struct h264_slice_header *sh,
		     void *u)
{
	struct rw_state *st = u;
	struct h264_slice_header nsh = *sh;
	struct h264_bitstream bs;
	if (st->mode == 0)
		nsh.frame_num ^= 1;
	else
		nsh.slice_qp_delta += 17;
	h264_bs_init(&bs, st->out + (buf - st->base), len, 1);
	const int r = st->a->rewrite_slice_header(&bs, c, &nsh);
	if (st->n < st->cap)
		st->rcs[st->n] = r;
	st->n++;
}

long hh_rewrite(const char *libpath, const uint8_t *buf, size_t len, int mode, uint8_t *out, int32_t *rcs,
		uint32_t cap)
{
	struct api a;
	int r = api_open(&a, libpath);
	if (r < 0)
		return r;
	memcpy(out, buf, len);
	struct rw_state st = {&a, buf, out, mode, rcs, 0, cap};
	struct h264_ctx_cbs cbs;
	memset(&cbs, 0, sizeof(cbs));
	cbs.slice = rw_slice;
	struct h264_reader *rd = NULL;
	r = a.reader_new(&cbs, &st, &rd);
	if (r < 0)
		return r;
	/* NAL by NAL (headers only: no GPU stage is involved) */
	size_t off = 0, start = 0, end = 0;
	while (off < len) {
		int fr = a.find_nalu(buf + off, len - off, &start, &end);
		if (fr < 0 && fr != -EAGAIN)
			break;
		a.reader_parse_nalu(rd, 0, buf + off + start, end - start);
		off += end;
	}
	a.reader_destroy(rd);
	dlclose(a.h);
	return (long)st.n;
}

/* the same edits through this library's h264_rewrite_slice_header_patch: one patch record per
 * slice whose rewrite succeeds (nal_off relative to buf), rcs[k] = return code for slice k */
#include "h264gpu_slice.h"
struct rwp_state {
	int (*patch_fn)(struct h264_ctx *, const struct h264_slice_header *, struct h264gpu_hdr_patch *);
	const uint8_t *base;
	struct h264gpu_hdr_patch *patches;
	int mode;
	int32_t *rcs;
	uint32_t n, np, cap;
};

static void rwp_slice(struct h264_ctx *c, const uint8_t *buf, size_t len, const struct h264_slice_header *sh,
		      void *u)
{
	struct rwp_state *st = u;
	struct h264_slice_header nsh = *sh;
	(void)len;
	if (st->mode == 0)
		nsh.frame_num ^= 1;
	else
		nsh.slice_qp_delta += 17;
	struct h264gpu_hdr_patch pt;
	memset(&pt, 0, sizeof(pt));
	const int r = st->patch_fn(c, &nsh, &pt);
	if (st->n < st->cap) {
		st->rcs[st->n] = r;
		if (r == 0) {
			pt.nal_off = (uint64_t)(buf - st->base);
			st->patches[st->np++] = pt;
		}
	}
	st->n++;
}

long hh_rewrite_patches(const char *libpath, const uint8_t *buf, size_t len, int mode,
			struct h264gpu_hdr_patch *patches, int32_t *rcs, uint32_t cap, uint32_t *n_patches)
{
	struct api a;
	int r = api_open(&a, libpath);
	if (r < 0)
		return r;
	struct rwp_state st = {NULL, buf, patches, mode, rcs, 0, 0, cap};
	*(void **)&st.patch_fn = dlsym(a.h, "h264_rewrite_slice_header_patch");
	if (st.patch_fn == NULL)
		return -ENOSYS;
	struct h264_ctx_cbs cbs;
	memset(&cbs, 0, sizeof(cbs));
	cbs.slice = rwp_slice;
	struct h264_reader *rd = NULL;
	r = a.reader_new(&cbs, &st, &rd);
	if (r < 0)
		return r;
	size_t off = 0, start = 0, end = 0;
	while (off < len) {
		int fr = a.find_nalu(buf + off, len - off, &start, &end);
		if (fr < 0 && fr != -EAGAIN)
			break;
		a.reader_parse_nalu(rd, 0, buf + off + start, end - start);
		off += end;
	}
	a.reader_destroy(rd);
	dlclose(a.h);
	*n_patches = st.np;
	return (long)st.n;
}

/* ---- generator ----------------------------------------------------------------------------- */

static uint64_t rng_state;
static uint32_t rnd(void)
{
	rng_state ^= rng_state << 13;
	rng_state ^= rng_state >> 7;
	rng_state ^= rng_state << 17;
	return (uint32_t)(rng_state >> 16);
}
static uint32_t rn(uint32_t n) { return n ? rnd() % n : 0; } /* 0..n-1 */
static int32_t rs(int32_t lim) { return (int32_t)rn(2 * lim + 1) - lim; }
static int coin(void) { return rnd() & 1; }

struct out {
	uint8_t *buf;
	size_t cap, len;
};

static int emit_nalu(struct api *a, struct h264_ctx *ctx, struct out *o, const uint8_t *tail, size_t ntail)
{
	struct h264_bitstream bs;
	h264_bs_init(&bs, NULL, 0, 1);
	int r = a->write_nalu(&bs, ctx);
	if (r < 0) {
		h264_bs_clear(&bs);
		return r;
	}
	for (size_t i = 0; i < ntail && r >= 0; i++)
		r = a->bs_write_bits(&bs, tail[i], 8);
	if (r >= 0 && o->len + 4 + bs.off <= o->cap && bs.cachebits == 0) {
		static const uint8_t sc[4] = {0, 0, 0, 1};
		memcpy(o->buf + o->len, sc, 4);
		memcpy(o->buf + o->len + 4, bs.data, bs.off);
		o->len += 4 + bs.off;
	} else if (r >= 0) {
		r = -ENOBUFS;
	}
	h264_bs_clear(&bs);
	return r < 0 ? r : 0;
}

static void gen_hrd(struct h264_hrd *h)
{
	h->cpb_cnt_minus1 = rn(3);
	h->bit_rate_scale = rn(16);
	h->cpb_size_scale = rn(16);
	for (uint32_t i = 0; i <= h->cpb_cnt_minus1; i++) {
		h->cpb[i].bit_rate_value_minus1 = rn(100000);
		h->cpb[i].cpb_size_value_minus1 = rn(100000);
		h->cpb[i].cbr_flag = coin();
	}
	h->initial_cpb_removal_delay_length_minus1 = rn(32);
	h->cpb_removal_delay_length_minus1 = rn(32);
	h->dpb_output_delay_length_minus1 = rn(32);
	h->time_offset_length = rn(32);
}

static void gen_matrix(struct h264_scaling_matrix *m, uint32_t count)
{
	for (uint32_t i = 0; i < count; i++) {
		m->scaling_list_present_flag[i] = coin();
		if (!m->scaling_list_present_flag[i])
			continue;
		int32_t *l = i < 6 ? m->scaling_list_4x4[i] : m->scaling_list_8x8[i - 6];
		const uint32_t n = i < 6 ? 16 : 64;
		const uint32_t keep = 1 + rn(n); /* constant tail after `keep` values */
		for (uint32_t k = 0; k < n; k++)
			l[k] = k < keep ? (int32_t)(1 + rn(255)) : l[keep - 1];
		*(i < 6 ? &m->_optimized_4x4[i] : &m->_optimized_8x8[i - 6]) = coin();
	}
}

static void gen_sps(struct h264_sps *s, uint32_t id)
{
	static const uint32_t profiles[] = {66, 77, 88, 100, 110, 122, 244, 44};
	memset(s, 0, sizeof(*s));
	s->profile_idc = profiles[rn(8)];
	s->constraint_set0_flag = coin();
	s->constraint_set1_flag = coin();
	s->constraint_set3_flag = coin();
	s->level_idc = 10 + rn(42);
	s->seq_parameter_set_id = id;
	s->chroma_format_idc = 1;
	if (s->profile_idc >= 100 || s->profile_idc == 44) {
		s->chroma_format_idc = rn(4);
		if (s->chroma_format_idc == 3)
			s->separate_colour_plane_flag = coin();
		s->bit_depth_luma_minus8 = rn(7);
		s->bit_depth_chroma_minus8 = rn(7);
		s->qpprime_y_zero_transform_bypass_flag = coin();
		s->seq_scaling_matrix_present_flag = coin();
		if (s->seq_scaling_matrix_present_flag)
			gen_matrix(&s->seq_scaling_matrix, s->chroma_format_idc != 3 ? 8 : 12);
	}
	s->log2_max_frame_num_minus4 = rn(13);
	s->pic_order_cnt_type = rn(3);
	if (s->pic_order_cnt_type == 0) {
		s->log2_max_pic_order_cnt_lsb_minus4 = rn(13);
	} else if (s->pic_order_cnt_type == 1) {
		s->delta_pic_order_always_zero_flag = coin();
		s->offset_for_non_ref_pic = rs(100000);
		s->offset_for_top_to_bottom_field = rs(100000);
		s->num_ref_frames_in_pic_order_cnt_cycle = rn(6);
		for (uint32_t i = 0; i < s->num_ref_frames_in_pic_order_cnt_cycle; i++)
			s->offset_for_ref_frame[i] = rs(50000);
	}
	s->max_num_ref_frames = rn(17);
	s->gaps_in_frame_num_value_allowed_flag = coin();
	s->pic_width_in_mbs_minus1 = rn(120);
	s->pic_height_in_map_units_minus1 = rn(68);
	s->frame_mbs_only_flag = rn(4) != 0;
	if (!s->frame_mbs_only_flag)
		s->mb_adaptive_frame_field_flag = coin();
	s->direct_8x8_inference_flag = coin();
	s->frame_cropping_flag = coin();
	if (s->frame_cropping_flag) {
		s->frame_crop_left_offset = rn(4);
		s->frame_crop_right_offset = rn(4);
		s->frame_crop_top_offset = rn(4);
		s->frame_crop_bottom_offset = rn(4);
	}
	s->vui_parameters_present_flag = coin();
	if (s->vui_parameters_present_flag) {
		struct h264_vui *v = &s->vui;
		v->aspect_ratio_info_present_flag = coin();
		if (v->aspect_ratio_info_present_flag) {
			v->aspect_ratio_idc = coin() ? 255 : rn(17);
			if (v->aspect_ratio_idc == 255) {
				v->sar_width = 1 + rn(65535);
				v->sar_height = 1 + rn(65535);
			}
		}
		v->overscan_info_present_flag = coin();
		v->overscan_appropriate_flag = v->overscan_info_present_flag && coin();
		v->video_signal_type_present_flag = coin();
		if (v->video_signal_type_present_flag) {
			v->video_format = rn(8);
			v->video_full_range_flag = coin();
			v->colour_description_present_flag = coin();
			if (v->colour_description_present_flag) {
				v->colour_primaries = rn(256);
				v->transfer_characteristics = rn(256);
				v->matrix_coefficients = rn(256);
			}
		}
		v->chroma_loc_info_present_flag = coin();
		if (v->chroma_loc_info_present_flag) {
			v->chroma_sample_loc_type_top_field = rn(6);
			v->chroma_sample_loc_type_bottom_field = rn(6);
		}
		v->timing_info_present_flag = coin();
		if (v->timing_info_present_flag) {
			v->num_units_in_tick = 1 + rnd();
			v->time_scale = 1 + rnd();
			v->fixed_frame_rate_flag = coin();
		}
		v->nal_hrd_parameters_present_flag = coin();
		if (v->nal_hrd_parameters_present_flag)
			gen_hrd(&v->nal_hrd);
		v->vcl_hrd_parameters_present_flag = coin();
		if (v->vcl_hrd_parameters_present_flag)
			gen_hrd(&v->vcl_hrd);
		if (v->nal_hrd_parameters_present_flag || v->vcl_hrd_parameters_present_flag)
			v->low_delay_hrd_flag = coin();
		v->pic_struct_present_flag = coin();
		v->bitstream_restriction_flag = coin();
		if (v->bitstream_restriction_flag) {
			v->motion_vectors_over_pic_boundaries_flag = coin();
			v->max_bytes_per_pic_denom = rn(17);
			v->max_bits_per_mb_denom = rn(17);
			v->log2_max_mv_length_horizontal = rn(17);
			v->log2_max_mv_length_vertical = rn(17);
			v->max_num_reorder_frames = rn(17);
			v->max_dec_frame_buffering = rn(17);
		}
	}
}

static void gen_pps(struct h264_pps *p, const struct h264_sps *s, uint32_t id, int allow_cabac)
{
	memset(p, 0, sizeof(*p));
	p->pic_parameter_set_id = id;
	p->seq_parameter_set_id = s->seq_parameter_set_id;
	p->entropy_coding_mode_flag = allow_cabac && coin();
	p->bottom_field_pic_order_in_frame_present_flag = coin();
	if (rn(3) == 0) {
		p->num_slice_groups_minus1 = 1 + rn(3);
		p->slice_group_map_type = rn(7);
		const uint32_t units = (s->pic_width_in_mbs_minus1 + 1) * (s->pic_height_in_map_units_minus1 + 1);
		switch (p->slice_group_map_type) {
		case 0:
			for (uint32_t i = 0; i <= p->num_slice_groups_minus1; i++)
				p->run_length_minus1[i] = rn(50);
			break;
		case 2:
			for (uint32_t i = 0; i < p->num_slice_groups_minus1; i++) {
				p->top_left[i] = rn(units);
				p->bottom_right[i] = rn(units);
			}
			break;
		case 3:
		case 4:
		case 5:
			p->slice_group_change_direction_flag = coin();
			p->slice_group_change_rate_minus1 = rn(units);
			break;
		case 6:
			p->pic_size_in_map_units_minus1 = rn(units < 256 ? units : 256);
			for (uint32_t i = 0; i <= p->pic_size_in_map_units_minus1 && i < 256; i++)
				p->slice_group_id[i] = rn(p->num_slice_groups_minus1 + 1);
			break;
		default:
			break;
		}
	}
	p->num_ref_idx_l0_default_active_minus1 = rn(4);
	p->num_ref_idx_l1_default_active_minus1 = rn(4);
	p->weighted_pred_flag = coin();
	p->weighted_bipred_idc = rn(3);
	p->pic_init_qp_minus26 = rs(20);
	p->pic_init_qs_minus26 = rs(20);
	p->chroma_qp_index_offset = rs(12);
	p->deblocking_filter_control_present_flag = coin();
	p->constrained_intra_pred_flag = coin();
	p->redundant_pic_cnt_present_flag = coin();
	if (s->profile_idc >= 100 && coin()) {
		p->_more_rbsp_data_present = 1;
		p->transform_8x8_mode_flag = coin();
		p->pic_scaling_matrix_present_flag = coin();
		if (p->pic_scaling_matrix_present_flag)
			gen_matrix(&p->pic_scaling_matrix,
				   6 + (p->transform_8x8_mode_flag ? (s->chroma_format_idc != 3 ? 2 : 6) : 0));
		p->second_chroma_qp_index_offset = rs(12);
	}
}

static void gen_sei(struct h264_sei *sei, const struct h264_sps *s, uint8_t *blob, size_t blobcap)
{
	static const uint32_t types[] = {0, 1, 2, 3, 4, 5, 6};
	memset(sei, 0, sizeof(*sei));
	sei->type = (enum h264_sei_type)types[rn(7)];
	/* a picture timing SEI without HRD and pic_struct would be empty, which neither
	 * library accepts (raw.len == 0) */
	if (sei->type == H264_SEI_TYPE_PIC_TIMING && !s->vui.nal_hrd_parameters_present_flag &&
	    !s->vui.vcl_hrd_parameters_present_flag && !s->vui.pic_struct_present_flag)
		sei->type = H264_SEI_TYPE_RECOVERY_POINT;
	const size_t n = 1 + rn((uint32_t)blobcap - 1);
	for (size_t i = 0; i < n; i++)
		blob[i] = (uint8_t)(rn(5) == 0 ? 0 : rnd());
	switch (sei->type) {
	case H264_SEI_TYPE_BUFFERING_PERIOD:
		sei->buffering_period.seq_parameter_set_id = s->seq_parameter_set_id;
		for (int i = 0; i < 4; i++) {
			const uint32_t nb = s->vui.nal_hrd.initial_cpb_removal_delay_length_minus1 + 1;
			const uint32_t vb = s->vui.vcl_hrd.initial_cpb_removal_delay_length_minus1 + 1;
			sei->buffering_period.nal_hrd_cpb[i].initial_cpb_removal_delay = rnd() & (uint32_t)((1ull << nb) - 1);
			sei->buffering_period.nal_hrd_cpb[i].initial_cpb_removal_delay_offset = rnd() & (uint32_t)((1ull << nb) - 1);
			sei->buffering_period.vcl_hrd_cpb[i].initial_cpb_removal_delay = rnd() & (uint32_t)((1ull << vb) - 1);
			sei->buffering_period.vcl_hrd_cpb[i].initial_cpb_removal_delay_offset = rnd() & (uint32_t)((1ull << vb) - 1);
		}
		break;
	case H264_SEI_TYPE_PIC_TIMING: {
		const struct h264_hrd *h = s->vui.nal_hrd_parameters_present_flag   ? &s->vui.nal_hrd
					   : s->vui.vcl_hrd_parameters_present_flag ? &s->vui.vcl_hrd
										    : NULL;
		if (h != NULL) {
			sei->pic_timing.cpb_removal_delay = rnd() & (uint32_t)((1ull << (h->cpb_removal_delay_length_minus1 + 1)) - 1);
			sei->pic_timing.dpb_output_delay = rnd() & (uint32_t)((1ull << (h->dpb_output_delay_length_minus1 + 1)) - 1);
		}
		sei->pic_timing.pic_struct = rn(9);
		const uint32_t tol = h != NULL ? h->time_offset_length : 24;
		for (int i = 0; i < 3; i++) {
			sei->pic_timing.clk_ts[i].clock_timestamp_flag = coin();
			sei->pic_timing.clk_ts[i].ct_type = rn(3);
			sei->pic_timing.clk_ts[i].nuit_field_based_flag = coin();
			sei->pic_timing.clk_ts[i].counting_type = rn(7);
			sei->pic_timing.clk_ts[i].full_timestamp_flag = coin();
			sei->pic_timing.clk_ts[i].discontinuity_flag = coin();
			sei->pic_timing.clk_ts[i].cnt_dropped_flag = coin();
			sei->pic_timing.clk_ts[i].n_frames = rn(256);
			sei->pic_timing.clk_ts[i].seconds_flag = coin();
			sei->pic_timing.clk_ts[i].minutes_flag = coin();
			sei->pic_timing.clk_ts[i].hours_flag = coin();
			sei->pic_timing.clk_ts[i].seconds_value = rn(60);
			sei->pic_timing.clk_ts[i].minutes_value = rn(60);
			sei->pic_timing.clk_ts[i].hours_value = rn(24);
			if (tol > 0)
				sei->pic_timing.clk_ts[i].time_offset = rs((int32_t)((1u << (tol - 1)) - 1));
		}
		break;
	}
	case H264_SEI_TYPE_PAN_SCAN_RECT:
		sei->pan_scan_rect.pan_scan_rect_id = rn(1000);
		sei->pan_scan_rect.pan_scan_rect_cancel_flag = coin();
		sei->pan_scan_rect.pan_scan_cnt_minus1 = rn(3);
		for (int i = 0; i < 3; i++) {
			sei->pan_scan_rect.pan_scan_rect[i].left_offset = rs(5000);
			sei->pan_scan_rect.pan_scan_rect[i].right_offset = rs(5000);
			sei->pan_scan_rect.pan_scan_rect[i].top_offset = rs(5000);
			sei->pan_scan_rect.pan_scan_rect[i].bottom_offset = rs(5000);
		}
		sei->pan_scan_rect.pan_scan_rect_repetition_period = rn(100);
		break;
	case H264_SEI_TYPE_FILLER_PAYLOAD:
		memset(blob, 0xff, n);
		sei->filler_payload.buf = blob;
		sei->filler_payload.len = n;
		break;
	case H264_SEI_TYPE_USER_DATA_REGISTERED:
		sei->user_data_registered.country_code = coin() ? 0xff : rn(255);
		sei->user_data_registered.country_code_extension_byte = rn(256);
		sei->user_data_registered.buf = blob;
		sei->user_data_registered.len = n;
		break;
	case H264_SEI_TYPE_USER_DATA_UNREGISTERED:
		for (int i = 0; i < 16; i++)
			sei->user_data_unregistered.uuid[i] = (uint8_t)rnd();
		sei->user_data_unregistered.buf = blob;
		sei->user_data_unregistered.len = n;
		break;
	default:
		sei->recovery_point.recovery_frame_cnt = rn(1000);
		sei->recovery_point.exact_match_flag = coin();
		sei->recovery_point.broken_link_flag = coin();
		sei->recovery_point.changing_slice_group_idc = rn(4);
		break;
	}
}

static void gen_slice_header(struct h264_slice_header *sh, const struct h264_sps *s, const struct h264_pps *p,
			     int idr, uint32_t nal_ref_idc)
{
	memset(sh, 0, sizeof(*sh));
	const uint32_t pic_mbs = (s->pic_width_in_mbs_minus1 + 1) * (s->pic_height_in_map_units_minus1 + 1);
	sh->first_mb_in_slice = rn(pic_mbs);
	const uint32_t type = idr ? 2 : rn(5);
	sh->slice_type = type + (coin() ? 5 : 0);
	sh->pic_parameter_set_id = p->pic_parameter_set_id;
	if (s->separate_colour_plane_flag)
		sh->colour_plane_id = rn(3);
	sh->frame_num = rnd() & ((1u << (s->log2_max_frame_num_minus4 + 4)) - 1);
	if (!s->frame_mbs_only_flag) {
		sh->field_pic_flag = coin();
		if (sh->field_pic_flag)
			sh->bottom_field_flag = coin();
	}
	if (idr)
		sh->idr_pic_id = rn(65536);
	if (s->pic_order_cnt_type == 0) {
		sh->pic_order_cnt_lsb = rnd() & ((1u << (s->log2_max_pic_order_cnt_lsb_minus4 + 4)) - 1);
		if (p->bottom_field_pic_order_in_frame_present_flag && !sh->field_pic_flag)
			sh->delta_pic_order_cnt_bottom = rs(1000);
	}
	if (s->pic_order_cnt_type == 1 && !s->delta_pic_order_always_zero_flag) {
		sh->delta_pic_order_cnt[0] = rs(1000);
		if (p->bottom_field_pic_order_in_frame_present_flag && !sh->field_pic_flag)
			sh->delta_pic_order_cnt[1] = rs(1000);
	}
	if (p->redundant_pic_cnt_present_flag)
		sh->redundant_pic_cnt = rn(128);
	sh->num_ref_idx_l0_active_minus1 = p->num_ref_idx_l0_default_active_minus1;
	sh->num_ref_idx_l1_active_minus1 = p->num_ref_idx_l1_default_active_minus1;
	if (type == 1)
		sh->direct_spatial_mv_pred_flag = coin();
	if (type == 0 || type == 3 || type == 1) {
		sh->num_ref_idx_active_override_flag = coin();
		if (sh->num_ref_idx_active_override_flag) {
			sh->num_ref_idx_l0_active_minus1 = rn(8);
			if (type == 1)
				sh->num_ref_idx_l1_active_minus1 = rn(8);
		}
	}
	for (int l = 0; l < 2; l++) {
		if ((l == 0 && (type == 2 || type == 4)) || (l == 1 && type != 1))
			continue;
		int *flag = l == 0 ? &sh->rplm.ref_pic_list_modification_flag_l0
				   : &sh->rplm.ref_pic_list_modification_flag_l1;
		struct h264_rplm_item *it = l == 0 ? sh->rplm.pic_num_l0 : sh->rplm.pic_num_l1;
		*flag = coin();
		if (*flag) {
			const uint32_t n = rn(5);
			for (uint32_t i = 0; i < n; i++) {
				it[i].modification_of_pic_nums_idc = rn(3);
				it[i].abs_diff_pic_num_minus1 = rn(100);
			}
			it[n].modification_of_pic_nums_idc = 3;
		}
	}
	if ((p->weighted_pred_flag && (type == 0 || type == 3)) || (p->weighted_bipred_idc == 1 && type == 1)) {
		const uint32_t cat = s->separate_colour_plane_flag ? 0 : s->chroma_format_idc;
		sh->pwt.luma_log2_weight_denom = rn(8);
		if (cat)
			sh->pwt.chroma_log2_weight_denom = rn(8);
		for (int l = 0; l < 2; l++) {
			struct h264_pwt_item *w = l == 0 ? sh->pwt.l0 : sh->pwt.l1;
			const uint32_t n = l == 0 ? sh->num_ref_idx_l0_active_minus1 : sh->num_ref_idx_l1_active_minus1;
			if (l == 1 && type != 1)
				break;
			for (uint32_t i = 0; i <= n; i++) {
				w[i].luma_weight_flag = coin();
				if (w[i].luma_weight_flag) {
					w[i].luma_weight = rs(127);
					w[i].luma_offset = rs(127);
				}
				if (cat) {
					w[i].chroma_weight_flag = coin();
					if (w[i].chroma_weight_flag)
						for (int c = 0; c < 2; c++) {
							w[i].chroma_weight[c] = rs(127);
							w[i].chroma_offset[c] = rs(127);
						}
				}
			}
		}
	}
	if (nal_ref_idc != 0) {
		if (idr) {
			sh->drpm.no_output_of_prior_pics_flag = coin();
			sh->drpm.long_term_reference_flag = coin();
		} else {
			sh->drpm.adaptive_ref_pic_marking_mode_flag = coin();
			if (sh->drpm.adaptive_ref_pic_marking_mode_flag) {
				const uint32_t n = rn(5);
				for (uint32_t i = 0; i < n; i++) {
					struct h264_drpm_item *m = &sh->drpm.mm[i];
					m->memory_management_control_operation = 1 + rn(6);
					const uint32_t op = m->memory_management_control_operation;
					if (op == 1 || op == 3)
						m->difference_of_pic_nums_minus1 = rn(100);
					if (op == 2)
						m->long_term_pic_num = rn(100);
					if (op == 3 || op == 6)
						m->long_term_frame_idx = rn(16);
					if (op == 4)
						m->max_long_term_frame_idx_plus1 = rn(17);
				}
				sh->drpm.mm[n].memory_management_control_operation = 0;
			}
		}
	}
	if (p->entropy_coding_mode_flag && type != 2 && type != 4)
		sh->cabac_init_idc = rn(3);
	sh->slice_qp_delta = rs(10);
	if (type == 3 || type == 4) {
		if (type == 3)
			sh->sp_for_switch_flag = coin();
		sh->slice_qs_delta = rs(10);
	}
	if (p->deblocking_filter_control_present_flag) {
		sh->disable_deblocking_filter_idc = rn(3);
		if (sh->disable_deblocking_filter_idc != 1) {
			sh->slice_alpha_c0_offset_div2 = rs(6);
			sh->slice_beta_offset_div2 = rs(6);
		}
	}
	if (p->num_slice_groups_minus1 > 0 && p->slice_group_map_type >= 3 && p->slice_group_map_type <= 5) {
		uint32_t units = pic_mbs / (p->slice_group_change_rate_minus1 + 1) + 1, bits = 0;
		while ((1u << bits) < units)
			bits++;
		sh->slice_group_change_cycle = bits ? rnd() & ((1u << bits) - 1) : 0;
	}
}

/*
 * A seeded stream of `rounds` groups: SPS, PPS, [AUD], [SEI x k], [filler], slice NAL units
 * (random header followed by opaque bytes, or a concealment slice), all written by the
 * library under test.  Returns the stream length or a negative errno.
 */
/* N4 test input: for every concealment slice hh_gen writes, what a bulk synthesiser needs (the
 * unescaped NAL + slice header bits from h264_write_nalu, the slice's parameters) and where the
 * library's own h264_write_grey_i_slice / h264_write_skipped_p_slice output sits in the stream */
struct conceal_rec {
	uint64_t hdr_off;
	uint32_t hdr_bits, mb_count, first_mb_in_slice;
	uint16_t pic_width_in_mbs;
	uint8_t kind, entropy_coding_mode_flag, slice_type, cabac_init_idc;
	int8_t slice_qp;
	uint8_t reserved[5];
	uint64_t ref_off, ref_len; /* start code included */
};
static struct {
	struct conceal_rec *recs;
	size_t cap, n;
	uint8_t *hdr;
	size_t hdr_cap, hdr_used;
} clog;

long hh_gen(const char *libpath, uint64_t seed, int rounds, int conceal, uint8_t *outbuf, size_t cap);
long hh_gen_conceal(const char *libpath, uint64_t seed, int rounds, uint8_t *outbuf, size_t cap,
		    struct conceal_rec *recs, size_t recs_cap, size_t *nrecs, uint8_t *hdr, size_t hdr_cap)
{
	clog.recs = recs;
	clog.cap = recs_cap;
	clog.n = 0;
	clog.hdr = hdr;
	clog.hdr_cap = hdr_cap;
	clog.hdr_used = 0;
	long r = hh_gen(libpath, seed, rounds, 1, outbuf, cap);
	*nrecs = clog.n;
	clog.recs = NULL;
	return r;
}

long hh_gen(const char *libpath, uint64_t seed, int rounds, int conceal, uint8_t *outbuf, size_t cap)
{
	struct api a;
	int r = api_open(&a, libpath);
	if (r < 0)
		return r;
	rng_state = seed * 0x9E3779B97F4A7C15ull + 0x1234567;
	for (int i = 0; i < 8; i++)
		rnd();
	struct out o = {outbuf, cap, 0};
	struct h264_ctx *ctx = NULL;
	if ((r = a.ctx_new(&ctx)) < 0)
		return r;
	struct h264_sps *sps = calloc(1, sizeof(*sps));
	struct h264_pps *pps = calloc(1, sizeof(*pps));
	struct h264_slice_header *sh = calloc(1, sizeof(*sh));
	uint8_t blob[64], tail[96];
	for (int g = 0; g < rounds && r >= 0; g++) {
		struct h264_nalu_header nh = {0, 1 + rn(3), H264_NALU_TYPE_SPS};
		gen_sps(sps, rn(4));
		if (conceal) { /* concealment slices need a plain frame picture */
			sps->frame_mbs_only_flag = 1;
			sps->mb_adaptive_frame_field_flag = 0;
			sps->separate_colour_plane_flag = 0;
		}
		a.ctx_clear_nalu(ctx);
		a.ctx_set_nalu_header(ctx, &nh);
		if ((r = a.ctx_set_sps(ctx, sps)) < 0 || (r = emit_nalu(&a, ctx, &o, NULL, 0)) < 0)
			break;
		gen_pps(pps, sps, rn(8), 1);
		if (conceal)
			pps->num_slice_groups_minus1 = 0;
		nh.nal_unit_type = H264_NALU_TYPE_PPS;
		a.ctx_clear_nalu(ctx);
		a.ctx_set_nalu_header(ctx, &nh);
		if ((r = a.ctx_set_pps(ctx, pps)) < 0 || (r = emit_nalu(&a, ctx, &o, NULL, 0)) < 0)
			break;
		const int nslices = 1 + (int)rn(4);
		for (int k = 0; k < nslices && r >= 0; k++) {
			if (coin()) {
				struct h264_aud aud = {rn(8)};
				nh.nal_ref_idc = 0;
				nh.nal_unit_type = H264_NALU_TYPE_AUD;
				a.ctx_clear_nalu(ctx);
				a.ctx_set_nalu_header(ctx, &nh);
				a.ctx_set_aud(ctx, &aud);
				if ((r = emit_nalu(&a, ctx, &o, NULL, 0)) < 0)
					break;
			}
			if (coin()) {
				nh.nal_ref_idc = 0;
				nh.nal_unit_type = H264_NALU_TYPE_SEI;
				a.ctx_clear_nalu(ctx);
				a.ctx_set_nalu_header(ctx, &nh);
				const int nsei = 1 + (int)rn(3);
				for (int j = 0; j < nsei && r >= 0; j++) {
					struct h264_sei sei;
					gen_sei(&sei, sps, blob, sizeof(blob));
					r = a.ctx_add_sei(ctx, &sei);
				}
				if (r < 0 || (r = emit_nalu(&a, ctx, &o, NULL, 0)) < 0)
					break;
			}
			if (rn(4) == 0) {
				nh.nal_ref_idc = 0;
				nh.nal_unit_type = H264_NALU_TYPE_FILLER;
				a.ctx_clear_nalu(ctx);
				a.ctx_set_nalu_header(ctx, &nh);
				a.ctx_set_filler(ctx, rn(40));
				if ((r = emit_nalu(&a, ctx, &o, NULL, 0)) < 0)
					break;
			}
			const int idr = rn(4) == 0;
			nh.nal_ref_idc = idr ? 1 + rn(3) : rn(4);
			nh.nal_unit_type = idr ? H264_NALU_TYPE_SLICE_IDR : H264_NALU_TYPE_SLICE;
			a.ctx_clear_nalu(ctx);
			a.ctx_set_nalu_header(ctx, &nh);
			gen_slice_header(sh, sps, pps, idr, nh.nal_ref_idc);
			if (conceal) {
				const uint32_t w = sps->pic_width_in_mbs_minus1 + 1;
				const uint32_t pic = w * (sps->pic_height_in_map_units_minus1 + 1);
				const int grey = idr || coin();
				sh->slice_type = grey ? 7 : 5;
				if (grey) {
					memset(&sh->rplm, 0, sizeof(sh->rplm));
					memset(&sh->pwt, 0, sizeof(sh->pwt));
					sh->num_ref_idx_active_override_flag = 0;
					sh->direct_spatial_mv_pred_flag = 0;
					sh->cabac_init_idc = 0;
				}
				sh->sp_for_switch_flag = 0;
				sh->slice_qs_delta = 0;
				if ((r = a.ctx_set_slice_header(ctx, sh)) < 0)
					break;
				struct h264_bitstream bs;
				h264_bs_init(&bs, NULL, 0, 1);
				const uint32_t n = 1 + rn(pic - sh->first_mb_in_slice);
				if (clog.recs != NULL && clog.n < clog.cap) {
					/* header only, no emulation prevention: the synthesiser's input */
					struct h264_bitstream hb;
					h264_bs_init(&hb, NULL, 0, 0);
					if (a.write_nalu(&hb, ctx) >= 0 && clog.hdr_used + hb.off + 1 <= clog.hdr_cap) {
						struct conceal_rec *c = &clog.recs[clog.n];
						memset(c, 0, sizeof(*c));
						c->hdr_off = clog.hdr_used;
						c->hdr_bits = (uint32_t)(hb.off * 8 + hb.cachebits);
						memcpy(clog.hdr + clog.hdr_used, hb.data, hb.off);
						clog.hdr[clog.hdr_used + hb.off] = (uint8_t)hb.cache;
						clog.hdr_used += hb.off + 1;
						c->mb_count = n;
						c->first_mb_in_slice = sh->first_mb_in_slice;
						c->pic_width_in_mbs = (uint16_t)w;
						c->kind = grey ? 0 : 1;
						c->entropy_coding_mode_flag = (uint8_t)pps->entropy_coding_mode_flag;
						c->slice_type = (uint8_t)(sh->slice_type % 5);
						c->cabac_init_idc = (uint8_t)sh->cabac_init_idc;
						c->slice_qp = (int8_t)(26 + pps->pic_init_qp_minus26 + sh->slice_qp_delta);
						c->ref_off = o.len;
					}
					h264_bs_clear(&hb);
				}
				r = grey ? a.write_grey_i_slice(&bs, ctx, n) : a.write_skipped_p_slice(&bs, ctx, n);
				if (r >= 0 && clog.recs != NULL && clog.n < clog.cap && clog.recs[clog.n].hdr_bits)
					clog.recs[clog.n++].ref_len = 4 + bs.off;
				if (r >= 0 && o.len + 4 + bs.off <= o.cap) {
					static const uint8_t sc[4] = {0, 0, 0, 1};
					memcpy(o.buf + o.len, sc, 4);
					memcpy(o.buf + o.len + 4, bs.data, bs.off);
					o.len += 4 + bs.off;
				} else if (r >= 0) {
					r = -ENOBUFS;
				}
				h264_bs_clear(&bs);
			} else {
				if ((r = a.ctx_set_slice_header(ctx, sh)) < 0)
					break;
				/* opaque "slice data": stays byte-exact as long as the header ends the
				 * NAL on a byte boundary, so pad the header with ones then add bytes */
				struct h264_bitstream bs;
				h264_bs_init(&bs, NULL, 0, 1);
				r = a.write_nalu(&bs, ctx);
				while (r >= 0 && bs.cachebits != 0)
					r = a.bs_write_bits(&bs, 1, 1);
				const size_t nt = 1 + rn(sizeof(tail) - 1);
				for (size_t i = 0; i < nt && r >= 0; i++)
					r = a.bs_write_bits(&bs, i + 1 == nt ? 0x80 : (rn(4) == 0 ? 0 : rnd() & 0xff), 8);
				if (r >= 0 && o.len + 4 + bs.off <= o.cap) {
					static const uint8_t sc[4] = {0, 0, 0, 1};
					memcpy(o.buf + o.len, sc, 4);
					memcpy(o.buf + o.len + 4, bs.data, bs.off);
					o.len += 4 + bs.off;
				} else if (r >= 0) {
					r = -ENOBUFS;
				}
				h264_bs_clear(&bs);
				(void)tail;
			}
		}
	}
	free(sps);
	free(pps);
	free(sh);
	a.ctx_destroy(ctx);
	dlclose(a.h);
	return r < 0 ? r : (long)o.len;
}

/* ---- ABI: struct sizes as this repo's headers lay them out ---------- */
int hh_sizes(uint32_t *out, int cap)
{
	const uint32_t v[] = {sizeof(struct h264_bitstream), sizeof(struct h264_nalu_header), sizeof(struct h264_sps),
			      sizeof(struct h264_pps),       sizeof(struct h264_vui),         sizeof(struct h264_hrd),
			      sizeof(struct h264_scaling_matrix), sizeof(struct h264_slice_header), sizeof(struct h264_rplm),
			      sizeof(struct h264_pwt),       sizeof(struct h264_drpm),        sizeof(struct h264_aud),
			      sizeof(struct h264_sei),       sizeof(struct h264_sps_derived), sizeof(struct h264_info),
			      sizeof(struct h264_ctx_cbs)};
	int n = (int)(sizeof(v) / sizeof(v[0]));
	for (int i = 0; i < n && i < cap; i++)
		out[i] = v[i];
	return n;
}
