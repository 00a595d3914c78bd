/*
 * synth_cabac.cpp — CABAC side of the synthetic-stream library (host C++; WORKLOAD AND TEST
 * INFRASTRUCTURE, never on the parse path):
 *
 *   synth_cabac_slice      seeded-random but valid CABAC slice_data() for the generator in
 *                          synth_video.c (BASELINE.json configs 3 and 4)
 *   synth_cabac_transcode  re-code macroblock syntax elements (struct h264_mb_syntax, dumped
 *                          by the REFERENCE's CAVLC parse) as a CABAC slice: the twin stream
 *                          that pins the CABAC parse to the reference's own parse results
 *   synth_cabac_decode     the walker of cabac_syntax.h instantiated with the decoder on the
 *                          CPU: the checker the GPU kernel is compared against
 *   synth_cabac_ops        raw engine access (encode a bin script) for the bit-exact
 *                          comparison with the reference's encoder (src/h264_bac.c:150-358)
 *
 * The walker is the same source the GPU kernel compiles (cabac_syntax.h); binarisations and
 * ctxIdx follow Rec. ITU-T H.264 9.3; the pieces libh264 implements are compared with it in
 * tests/test_cabac.py.
 */
#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "h264gpu_slice.h"
#include "mb_syntax.h"

#define CABAC_HD
#define CABAC_CONST static const
#include "cabac_engine.h"
#include "cabac_syntax.h"
#define CABAC_TAB static const
#include "cabac_tables.h"

using namespace cabac;

namespace {

Tables host_tables()
{
	Tables t;
	t.range_lps = &cabac_range_lps[0][0];
	t.trans_lps = cabac_trans_lps;
	t.trans_mps = cabac_trans_mps;
	return t;
}

struct Rng {
	uint64_t s;
	uint32_t next()
	{
		uint64_t x = s;
		x ^= x << 13;
		x ^= x >> 7;
		x ^= x << 17;
		s = x;
		return (uint32_t)(x >> 24);
	}
	uint32_t n(uint32_t k) { return k ? next() % k : 0; }
	bool pct(uint32_t p) { return n(100) < p; }
};

struct SliceCfg {
	uint32_t width_mbs, height_mbs, first_mb, count;
	uint32_t slice_type, chroma_array_type, transform_8x8, direct_8x8_inference;
	uint32_t num_ref_l0, num_ref_l1, cabac_init_idc;
	int32_t slice_qp;
	uint32_t pct_skip, pct_intra_in_inter, pct_pcm, coef_density;
};

void params_from_cfg(const SliceCfg &c, h264gpu_slice_params &sp)
{
	memset(&sp, 0, sizeof(sp));
	sp.first_mb_in_slice = c.first_mb;
	sp.pic_width_in_mbs = (uint16_t)c.width_mbs;
	sp.pic_height_in_mbs = (uint16_t)c.height_mbs;
	sp.slice_type = (uint8_t)c.slice_type;
	sp.chroma_array_type = (uint8_t)c.chroma_array_type;
	sp.bit_depth_luma = sp.bit_depth_chroma = 8;
	sp.transform_8x8_mode_flag = (uint8_t)c.transform_8x8;
	sp.direct_8x8_inference_flag = (uint8_t)c.direct_8x8_inference;
	sp.num_ref_idx_l0_active_minus1 = (uint8_t)(c.num_ref_l0 ? c.num_ref_l0 - 1 : 0);
	sp.num_ref_idx_l1_active_minus1 = (uint8_t)(c.num_ref_l1 ? c.num_ref_l1 - 1 : 0);
	sp.entropy_coding_mode_flag = 1;
	sp.cabac_init_idc = (uint8_t)c.cabac_init_idc;
	sp.slice_qp = (int8_t)c.slice_qp;
}

void start_encoder(Walk<Enc> &w, const h264gpu_slice_params &sp, uint8_t *out, uint64_t cap)
{
	w.c.t = host_tables();
	init_contexts(w.c.st, 1, cabac_init_mn[sp.slice_type == ST_I ? 0 : 1 + (sp.cabac_init_idc % 3)], sp.slice_qp);
	w.c.start(out, cap);
}

/* ---- random syntax ---------------------------------------------------------------------- */

int16_t rand_level(Rng &r)
{
	const uint32_t m = r.n(1000);
	int mag;
	if (m < 600)
		mag = 1;
	else if (m < 850)
		mag = 2 + (int)r.n(3);
	else if (m < 980)
		mag = 5 + (int)r.n(30); /* crosses the 14-bin prefix: UEG0 suffix */
	else
		mag = 35 + (int)r.n(900);
	return (int16_t)((r.next() & 1) ? -mag : mag);
}

/* n slots; returns the count of non-zeros (at least 1 when must) */
int rand_block(Rng &r, const SliceCfg &c, int16_t *d, int n, bool must)
{
	for (int i = 0; i < n; i++)
		d[i] = 0;
	if (!must && !r.pct(c.coef_density))
		return 0;
	int tc;
	const uint32_t sel = r.n(100);
	if (sel < 45)
		tc = 1 + (int)r.n(3);
	else if (sel < 80)
		tc = 1 + (int)r.n((uint32_t)(n < 8 ? n : 8));
	else
		tc = 1 + (int)r.n((uint32_t)n);
	if (tc > n)
		tc = n;
	int placed = 0;
	while (placed < tc) {
		const int span = r.n(4) == 0 ? n : (tc + 3 < n ? tc + 3 : n);
		const int p = (int)r.n((uint32_t)span);
		if (d[p])
			continue;
		d[p] = rand_level(r);
		placed++;
	}
	return tc;
}

int32_t rand_mvd(Rng &r)
{
	const uint32_t s = r.n(100);
	const int32_t m = s < 50 ? 0 : s < 85 ? (int32_t)r.n(8) : (int32_t)r.n(300);
	return (r.next() & 1) ? -m : m;
}

void rand_cbp(Rng &r, const SliceCfg &c, Mb &m)
{
	m.cbp_luma = r.n(100) < 30 ? 0 : (r.n(100) < 40 ? 15 : r.n(16));
	m.cbp_chroma = (c.chroma_array_type == 1 || c.chroma_array_type == 2) ? r.n(3) : 0;
}

void rand_residual(Rng &r, const SliceCfg &c, const Mb &m, bool i16, MbCoef &mc)
{
	if (i16)
		rand_block(r, c, mc.dc16, 16, false);
	for (uint32_t b8 = 0; b8 < 4; b8++) {
		if (!((m.cbp_luma >> b8) & 1))
			continue;
		if (m.t8) {
			rand_block(r, c, mc.luma8[b8], 64, true);
		} else {
			for (uint32_t b4 = 0; b4 < 4; b4++)
				rand_block(r, c, mc.luma[b8 * 4 + b4], i16 ? 15 : 16, false);
		}
	}
	const uint32_t nblk = c.chroma_array_type == 1 ? 4 : 8;
	if (c.chroma_array_type == 1 || c.chroma_array_type == 2) {
		if (m.cbp_chroma & 3)
			for (int ic = 0; ic < 2; ic++)
				rand_block(r, c, mc.cdc[ic], (int)nblk, false);
		if (m.cbp_chroma & 2)
			for (int ic = 0; ic < 2; ic++)
				for (uint32_t b = 0; b < nblk; b++)
					rand_block(r, c, mc.cac[ic][b], 15, false);
	}
}

/* fills m (+ mc) with a random intra macroblock; type_base: 0 (I), 5 (P), 23 (B) */
void rand_intra(Rng &r, const SliceCfg &c, uint32_t type_base, Mb &m, MbCoef &mc)
{
	const uint32_t sel = r.n(1000);
	const bool chroma = c.chroma_array_type == 1 || c.chroma_array_type == 2;
	if (sel < c.pct_pcm) {
		m.raw_type = type_base + 25;
		for (int i = 0; i < 768; i++)
			mc.pcm[i] = (uint8_t)r.n(256);
		return;
	}
	if (sel < 500) {
		m.raw_type = type_base;
		m.t8 = c.transform_8x8 ? (uint8_t)(r.next() & 1) : 0;
		for (int i = 0; i < 16; i++) {
			m.prev_flag[i] = (uint8_t)(r.next() & 1);
			m.rem_mode[i] = (uint8_t)r.n(8);
		}
		m.chroma_mode = chroma ? (uint8_t)r.n(4) : 0;
		rand_cbp(r, c, m);
		m.qp_delta = (int32_t)r.n(7) - 3;
		rand_residual(r, c, m, false, mc);
		return;
	}
	const uint32_t pm = r.n(4), cc = chroma ? r.n(3) : 0, cl = (r.next() & 1) ? 15 : 0;
	m.raw_type = type_base + 1 + pm + 4 * cc + (cl ? 12 : 0);
	m.cbp_luma = cl;
	m.cbp_chroma = cc;
	m.chroma_mode = chroma ? (uint8_t)r.n(4) : 0;
	m.qp_delta = (int32_t)r.n(7) - 3;
	rand_residual(r, c, m, true, mc);
}

void rand_inter_tail(Rng &r, const SliceCfg &c, Mb &m, MbCoef &mc, bool allow_t8)
{
	rand_cbp(r, c, m);
	m.t8 = (m.cbp_luma && c.transform_8x8 && allow_t8) ? (uint8_t)(r.next() & 1) : 0;
	m.qp_delta = (int32_t)r.n(7) - 3;
	rand_residual(r, c, m, false, mc);
}

void rand_p(Rng &r, const SliceCfg &c, Mb &m, MbCoef &mc)
{
	const uint32_t sel = r.n(100);
	for (int i = 0; i < 4; i++)
		m.ref_idx[0][i] = (int8_t)r.n(c.num_ref_l0);
	for (int i = 0; i < 16; i++)
		for (int k = 0; k < 2; k++)
			m.mvd[0][i][k] = (int16_t)rand_mvd(r);
	if (sel < 45) {
		m.raw_type = 0;
		rand_inter_tail(r, c, m, mc, true);
	} else if (sel < 70) {
		m.raw_type = 1 + (r.next() & 1);
		rand_inter_tail(r, c, m, mc, true);
	} else {
		m.raw_type = 3;
		bool no_small = true;
		for (int i = 0; i < 4; i++) {
			m.sub_type[i] = r.n(4);
			if (m.sub_type[i])
				no_small = false;
		}
		rand_inter_tail(r, c, m, mc, no_small);
	}
}

void rand_b(Rng &r, const SliceCfg &c, Mb &m, MbCoef &mc)
{
	for (int l = 0; l < 2; l++) {
		for (int i = 0; i < 4; i++)
			m.ref_idx[l][i] = (int8_t)r.n(l ? c.num_ref_l1 : c.num_ref_l0);
		for (int i = 0; i < 16; i++)
			for (int k = 0; k < 2; k++)
				m.mvd[l][i][k] = (int16_t)rand_mvd(r);
	}
	const uint32_t sel = r.n(100);
	if (sel < 15) {
		m.raw_type = 0; /* B_Direct_16x16 */
		rand_inter_tail(r, c, m, mc, c.direct_8x8_inference != 0);
	} else if (sel < 45) {
		m.raw_type = 1 + r.n(3);
		rand_inter_tail(r, c, m, mc, true);
	} else if (sel < 75) {
		m.raw_type = 4 + r.n(18);
		rand_inter_tail(r, c, m, mc, true);
	} else {
		m.raw_type = 22;
		bool no_small = true;
		for (int i = 0; i < 4; i++) {
			const uint32_t t = r.n(13);
			m.sub_type[i] = t;
			if (t == 0 ? !c.direct_8x8_inference : t >= 4)
				no_small = false;
		}
		rand_inter_tail(r, c, m, mc, no_small);
	}
}

uint64_t finish_slice(Walk<Enc> &w)
{
	w.c.align_zero(); /* rbsp_alignment_zero_bit after the stop bit the flush wrote */
	return w.c.n;
}

} /* namespace */

extern "C" {

/* optional sink: the records (mb_addr, mb_type, checksum of the INTENDED syntax elements) of
 * every macroblock the generator codes, in stream order; what a correct parse must return */
static thread_local h264gpu_mb_record *g_sink;
static thread_local uint64_t g_sink_cap, g_sink_n;

void synth_cabac_record_sink(struct h264gpu_mb_record *recs, uint64_t cap)
{
	g_sink = recs;
	g_sink_cap = cap;
	g_sink_n = 0;
}

uint64_t synth_cabac_records_written(void) { return g_sink_n; }

uint64_t synth_cabac_slice(const SliceCfg *cfg, uint64_t *rng_state, uint8_t *out, uint64_t cap)
{
	Rng r;
	r.s = *rng_state;
	h264gpu_slice_params sp;
	params_from_cfg(*cfg, sp);
	std::vector<Nb> ring(cfg->width_mbs + 1);
	Walk<Enc> w;
	w.begin_slice(&sp, ring.data());
	start_encoder(w, sp, out, cap);
	static thread_local MbCoef mc;
	std::vector<h264gpu_mb_record> local;
	for (uint32_t a = cfg->first_mb; a < cfg->first_mb + cfg->count; a++) {
		Mb m;
		memset(&m, 0, sizeof(m));
		memset(&mc, 0, sizeof(mc));
		bool skipped = cfg->slice_type != ST_I && r.pct(cfg->pct_skip);
		if (!skipped) {
			if (cfg->slice_type == ST_I)
				rand_intra(r, *cfg, 0, m, mc);
			else if (r.pct(cfg->pct_intra_in_inter))
				rand_intra(r, *cfg, cfg->slice_type == ST_P ? 5 : 23, m, mc);
			else if (cfg->slice_type == ST_P)
				rand_p(r, *cfg, m, mc);
			else
				rand_b(r, *cfg, m, mc);
		}
		bool end = a + 1 == cfg->first_mb + cfg->count;
		w.mc = &mc;
		w.mb_step(a, skipped, m, end);
		if (g_sink) {
			h264gpu_mb_record rec;
			rec.mb_addr = a;
			rec.mb_type = m.mb_type;
			rec.hash = w.hash;
			local.push_back(rec);
		}
	}
	*rng_state = r.s;
	if (g_sink && w.c.n + (w.c.nacc ? 1 : 0) <= cap) { /* not the undersized first try */
		for (size_t i = 0; i < local.size(); i++, g_sink_n++)
			if (g_sink_n < g_sink_cap)
				g_sink[g_sink_n] = local[i];
	}
	return finish_slice(w);
}

/*
 * Re-code n macroblocks (in slice order, skipped ones included with mb_type P_Skip/B_Skip) as
 * CABAC slice data.  sp: the parameter block of the CAVLC slice; cabac_init_idc / slice_qp are
 * taken from it.  Returns RBSP bytes needed, or 0 if a macroblock cannot be carried by CABAC
 * (P_8x8ref0, empty 8x8-transform block with its cbp bit set).
 */
uint64_t synth_cabac_transcode(const h264gpu_slice_params *sp_in, const struct h264_mb_syntax *mbs, uint32_t n,
			       uint8_t *out, uint64_t cap)
{
	h264gpu_slice_params sp = *sp_in;
	sp.entropy_coding_mode_flag = 1;
	std::vector<Nb> ring((size_t)sp.pic_width_in_mbs + 1);
	Walk<Enc> w;
	w.begin_slice(&sp, ring.data());
	start_encoder(w, sp, out, cap);
	static thread_local MbCoef mc;
	for (uint32_t k = 0; k < n; k++) {
		const h264_mb_syntax &s = mbs[k];
		Mb m;
		memset(&m, 0, sizeof(m));
		memset(&mc, 0, sizeof(mc));
		bool skipped = s.mb_type == MB_P_SKIP || s.mb_type == MB_B_SKIP;
		if (s.mb_type == MB_P_8x8ref0)
			return 0;
		m.raw_type = s.raw_mb_type;
		m.t8 = s.transform_size_8x8_flag;
		m.chroma_mode = s.intra_chroma_pred_mode;
		m.cbp_luma = s.cbp_luma;
		m.cbp_chroma = s.cbp_chroma;
		m.qp_delta = s.mb_qp_delta;
		const bool i8 = s.mb_type == MB_I_NxN && s.transform_size_8x8_flag;
		for (int i = 0; i < 16; i++) {
			const int8_t v = i8 ? (i < 4 ? s.intra8x8_pred_mode[i] : 0) : s.intra4x4_pred_mode[i];
			m.prev_flag[i] = v < 0;
			m.rem_mode[i] = v < 0 ? 0 : (uint8_t)v;
		}
		for (int l = 0; l < 2; l++) {
			for (int i = 0; i < 4; i++) {
				m.ref_idx[l][i] = (int8_t)s.ref_idx[l][i];
				m.sub_type[i] = s.raw_sub_mb_type[i];
			}
			memcpy(m.mvd[l], s.mvd[l], sizeof(m.mvd[l]));
		}
		const bool i16 = s.mb_type == MB_I_16x16;
		memcpy(mc.dc16, s.dc16, sizeof(mc.dc16));
		for (int b = 0; b < 16; b++)
			memcpy(mc.luma[b], i16 ? s.ac16[b] : s.l4[b], sizeof(mc.luma[b]));
		if (s.transform_size_8x8_flag) {
			for (int b8 = 0; b8 < 4; b8++) {
				bool any = false;
				for (int j = 0; j < 64; j++) {
					mc.luma8[b8][j] = s.l4[b8 * 4 + (j & 3)][j >> 2];
					any = any || mc.luma8[b8][j] != 0;
				}
				if (!any && ((s.cbp_luma >> b8) & 1))
					return 0;
			}
		}
		for (int ic = 0; ic < 2; ic++) {
			memcpy(mc.cdc[ic], s.cdc[ic], sizeof(mc.cdc[ic]));
			for (int b = 0; b < 8; b++)
				memcpy(mc.cac[ic][b], s.cac[ic][b], sizeof(mc.cac[ic][b]));
		}
		memcpy(mc.pcm, s.pcm, sizeof(mc.pcm));
		bool end = k + 1 == n;
		w.mc = &mc;
		if (!w.mb_step(s.mb_addr, skipped, m, end))
			return 0;
	}
	return finish_slice(w);
}

/*
 * CPU run of the walker with the decoder (same semantics as h264gpu_cabac_parse_host).
 * coefs (optional): n_records MbCoef-sized blobs, filled per macroblock.
 */
int synth_cabac_decode(const uint8_t *stream, uint64_t stream_len, const h264gpu_slice_params *params,
		       uint32_t n_slices, struct h264gpu_mb_record *records, uint64_t n_records,
		       struct h264gpu_slice_result *results)
{
	(void)stream_len;
	static uint8_t states[kNumCtx];
	for (uint32_t i = 0; i < n_slices; i++) {
		const h264gpu_slice_params &sp = params[i];
		h264gpu_slice_result &res = results[i];
		res.status = 0;
		res.mb_count = 0;
		res.end_bit = 0;
		if (!sp.entropy_coding_mode_flag) {
			res.status = H264GPU_SLICE_SKIPPED;
			continue;
		}
		if (sp.mbaff_frame_flag || sp.num_slice_groups_minus1 != 0 || sp.pic_width_in_mbs == 0 ||
		    sp.chroma_array_type == 3 || sp.slice_type >= ST_SP + 1) {
			res.status = -ENOSYS;
			continue;
		}
		std::vector<Nb> ring((size_t)sp.pic_width_in_mbs + 1);
		Walk<Dec> w;
		w.begin_slice(&sp, ring.data());
		w.c.st = states;
		w.c.stride = 1;
		w.c.t = host_tables();
		init_contexts(states, 1, cabac_init_mn[sp.slice_type == ST_I ? 0 : 1 + (sp.cabac_init_idc % 3)], sp.slice_qp);
		w.c.start(stream + sp.nal_off, sp.nal_len, sp.data_bit_off);
		const uint32_t pic_size = (uint32_t)sp.pic_width_in_mbs * sp.pic_height_in_mbs;
		uint32_t cur = sp.first_mb_in_slice, count = 0;
		int status = w.c.failed() ? -EIO : 0;
		while (!status) {
			if (count >= sp.mb_out_cap || cur >= pic_size || sp.mb_out_off + (uint64_t)count >= n_records) {
				status = -ENOBUFS;
				break;
			}
			Mb m;
			bool skipped = false, end = false;
			if (!w.mb_step(cur, skipped, m, end)) {
				status = -EIO;
				break;
			}
			h264gpu_mb_record &rec = records[sp.mb_out_off + count];
			rec.mb_addr = cur;
			rec.mb_type = m.mb_type;
			rec.hash = w.hash;
			count++;
			cur++;
			if (end)
				break;
		}
		res.status = status;
		res.mb_count = count;
		res.end_bit = w.c.raw_bitpos();
	}
	return 0;
}

/*
 * Encode a bin script with the engine: ops[i] = kind << 24 | ctx << 8 | bin, kind 0 decision,
 * 1 bypass, 2 terminate.  Contexts initialised for (table, slice_qp).  Returns bytes written
 * (partial last byte zero padded); used to compare with the reference's h264_bac_encode_*.
 */
uint64_t synth_cabac_ops(const uint32_t *ops, uint64_t n, uint32_t table, int32_t slice_qp, uint8_t *out,
			 uint64_t cap)
{
	Enc e;
	e.t = host_tables();
	init_contexts(e.st, 1, cabac_init_mn[table & 3], slice_qp);
	e.start(out, cap);
	for (uint64_t i = 0; i < n; i++) {
		const uint32_t kind = ops[i] >> 24, ctx = (ops[i] >> 8) & 0x3ff, b = ops[i] & 1;
		if (kind == 0)
			e.bin(ctx, b);
		else if (kind == 1)
			e.byp(b);
		else
			e.term(b);
	}
	e.align_zero();
	return e.n;
}

/* The matching decode of a bin script: fills bins[i] with the decoded values. */
int synth_cabac_ops_decode(const uint32_t *ops, uint64_t n, uint32_t table, int32_t slice_qp,
			   const uint8_t *buf, uint32_t len, uint8_t *bins)
{
	static uint8_t states[kNumCtx];
	Dec d;
	d.st = states;
	d.stride = 1;
	d.t = host_tables();
	init_contexts(states, 1, cabac_init_mn[table & 3], slice_qp);
	d.start(buf, len, 0);
	for (uint64_t i = 0; i < n; i++) {
		const uint32_t kind = ops[i] >> 24, ctx = (ops[i] >> 8) & 0x3ff;
		bins[i] = (uint8_t)(kind == 0 ? d.bin(ctx, 0) : kind == 1 ? d.byp(0) : d.term(0));
	}
	return d.failed() ? -EIO : 0;
}

uint32_t synth_cabac_sizeof_mb_syntax(void) { return (uint32_t)sizeof(struct h264_mb_syntax); }

} /* extern "C" */
