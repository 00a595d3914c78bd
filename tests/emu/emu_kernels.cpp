/*
 * emu_kernels.cpp — runs the CUDA kernels' source on the CPU SIMT emulator
 * (tests/emu/cuda_emu.h) so their logic is checked against the oracle without a
 * GPU.  TEST INFRASTRUCTURE ONLY; built by tests/emu/build.py with g++ -DH264_EMU.
 */
#include "annexb_scan.cuh"
#include "annexb_scan2.cuh"
#include "annexb_scan7.cuh"
#include "annexb_frame.cuh"
#include "annexb_frame6.cuh"
#include "annexb_frame7.cuh"
#include "annexb_frame8.cuh"

#include <vector>

extern "C" void emu_set_prefix_every(uint32_t n) { annexb::emu_prefix_every = n; }

extern "C" int emu_split_strip(const uint8_t *in, uint64_t len, uint64_t base,
			       const struct h264gpu_shard_edge *edge, uint8_t *rbsp,
			       uint64_t *nal_start, uint64_t *nal_end, uint64_t *nal_rbsp,
			       uint64_t nal_cap, struct h264gpu_scan_result *result, int items)
{
	using namespace annexb;
	if (len == 0)
		return -1;
	/* items = 16-byte chunks per thread of the packed-RBSP kernel (1, 2 or 8; 1xx accepted too) */
	const int cpt = items % 100;
	const uint64_t tile = (uint64_t)annexb2::kT * (cpt == 1 || cpt == 2 ? cpt : 8) * 16;
	const uint32_t ntiles = (uint32_t)((len + tile - 1) / tile);
	std::vector<uint64_t> desc((size_t)ntiles * 4, ~0ull);
	uint32_t ticket = 0xffffffffu;
	memset(result, 0xff, sizeof(*result));
	/* 16-byte aligned private copy, exactly len bytes (catches over-reads under ASan) */
	uint8_t *buf = (uint8_t *)aligned_alloc(16, (len + 15) & ~15ull);
	memcpy(buf, in, len);
	ScanArgs a;
	memset(&a, 0, sizeof(a));
	a.in = buf;
	a.len = len;
	a.base = base;
	a.rbsp = rbsp;
	a.nal_start = nal_start;
	a.nal_end = nal_end;
	a.nal_rbsp = nal_rbsp;
	a.nal_cap = nal_cap;
	a.desc = desc.data();
	a.ticket = &ticket;
	a.result = result;
	a.num_tiles = ntiles;
	a.halo_left = 0xffffffffu;
	a.right[0] = a.right[1] = 0xff;
	if (edge) {
		if (edge->has_left)
			a.halo_left = 0xffffu | (uint32_t)edge->left[0] << 16 | (uint32_t)edge->left[1] << 24;
		a.has_right = edge->has_right;
		a.right[0] = edge->right[0];
		a.right[1] = edge->right[1];
		a.init_in = edge->assume_in;
	}
	dim3 grid(ntiles), b2(annexb2::kT);
	if (rbsp) {
		if (cpt == 1) EMU_LAUNCH((annexb2::scan2_kernel<1, true>), grid, b2, a);
		else if (cpt == 2) EMU_LAUNCH((annexb2::scan2_kernel<2, true>), grid, b2, a);
		else EMU_LAUNCH((annexb2::scan2_kernel<8, true>), grid, b2, a);
	} else {
		if (cpt == 1) EMU_LAUNCH((annexb2::scan2_kernel<1, false>), grid, b2, a);
		else if (cpt == 2) EMU_LAUNCH((annexb2::scan2_kernel<2, false>), grid, b2, a);
		else EMU_LAUNCH((annexb2::scan2_kernel<8, false>), grid, b2, a);
	}
	free(buf);
	return 0;
}

/* gen 7: warp-autonomous spans (annexb_scan7.cuh); rows = 512-byte rows per span (1, 2 or 8).
 * `epoch` and a dirty (non-zero, stale-epoch) chain buffer check that nothing needs clearing. */
template <int ROWS>
static void emu_scan7_run(const annexb7::Scan7Args &a, const annexb7::Fin7Args &f, bool strip, uint64_t ev_cap)
{
	const uint32_t nctas = (a.num_spans + annexb7::kW - 1) / annexb7::kW;
	dim3 grid(nctas < 3 ? nctas : 3), block(annexb7::kT);
	if (strip) {
		EMU_LAUNCH((annexb7::scan7_kernel<ROWS, 1>), grid, block, a);
		if (a.regions > 1) {
			annexb7::Scan7Args b = a;
			b.pass2 = 1;
			b.regions = 1;
			EMU_LAUNCH((annexb7::scan7_kernel<ROWS, 1>), grid, block, b);
		}
	} else
		EMU_LAUNCH((annexb7::scan7_only_kernel<ROWS, 1>), grid, block, a);
	dim3 gs(f.nblk), bs(annexb7::kFinT), b256(256);
	EMU_LAUNCH((annexb7::fin7_spans), gs, bs, f);
	dim3 go((a.num_spans + 255) / 256);
	EMU_LAUNCH((annexb7::fin7_order), go, b256, f);
	uint64_t tb = (ev_cap + 255) / 256;
	if (tb > 5)
		tb = 5; /* fewer blocks than events: the grid-stride loop is exercised */
	dim3 gt((uint32_t)(tb ? tb : 1));
	EMU_LAUNCH((annexb7::fin7_table), gt, b256, f);
}

static int emu_scan7(const uint8_t *buf, uint64_t len, uint64_t base, const struct h264gpu_shard_edge *edge,
		     uint8_t *rbsp, uint64_t *nal_start, uint64_t *nal_end, uint64_t *nal_rbsp,
		     uint64_t *nal_rbsp_len, uint64_t nal_cap, struct h264gpu_scan_result *result, int rows,
		     uint64_t ev_cap)
{
	using namespace annexb7;
	const uint64_t span = (uint64_t)rows * 512;
	const uint32_t nspans = (uint32_t)((len + 2 + span - 1) / span);
	const uint32_t nblk = (nspans + kFinT - 1) / kFinT;
	static uint32_t epoch = 0;
	static std::vector<uint64_t> chain; /* kept across launches: they see each other's stale words */
	if (++epoch > 0xffffu) {
		std::fill(chain.begin(), chain.end(), 0);
		epoch = 1;
	}
	if (chain.size() < nspans)
		chain.resize(nspans, 0);
	std::vector<uint64_t> fin(nspans, ~0ull), pre(nspans, ~0ull), blk(2 * (size_t)nblk, ~0ull);
	std::vector<uint64_t> evbuf(2 * (ev_cap ? ev_cap : 1), 0), ordered(3 * (ev_cap ? ev_cap : 1), 0);
	uint64_t totals[8] = {0};
	static uint32_t ctrl[annexb7::kCtrlBytes / 4] = {0}; /* zero once; the finalize kernels re-arm it */
	Scan7Args a;
	memset(&a, 0, sizeof(a));
	a.in = buf;
	a.len = len;
	a.rbsp = rbsp;
	a.chain = chain.data();
	a.fin = fin.data();
	a.ctrl = ctrl;
	a.evbuf = evbuf.data();
	a.ev_cap = ev_cap;
	a.num_spans = nspans;
	a.halo_left = 0xffffffffu;
	a.epoch = epoch;
	a.pf_dist = 1;
	a.nap_max = 256;
	/* regions: 1, 3, 5, 7 in turn (small streams get regions of a few spans: lots of deferrals) */
	static uint32_t kreg = 0;
	kreg = (kreg + 1) % 4;
	a.regions = 2 * kreg + 1;
	if (a.regions > nspans)
		a.regions = nspans;
	a.region_len = (nspans + a.regions - 1) / a.regions;
	/* ticket counters: 0 (the single word), 2, 3, 8, 5 in turn.  Never more than the warps of a
	 * CTA: the emulator runs one CTA at a time, so the first one must serve every counter. */
	static uint32_t ktick = 0;
	static const uint32_t tick_ns[5] = {0, 2, 3, 8, 5};
	a.tick_n = tick_ns[ktick++ % 5];
	std::vector<uint32_t> deferred(nspans + 1, 0);
	a.deferred = deferred.data();
	a.right[0] = a.right[1] = 0xff;
	int assume_in = 0;
	if (edge) {
		if (edge->has_left)
			a.halo_left = 0xffffu | (uint32_t)edge->left[0] << 16 | (uint32_t)edge->left[1] << 24;
		a.has_right = edge->has_right;
		a.right[0] = edge->right[0];
		a.right[1] = edge->right[1];
		assume_in = edge->assume_in;
	}
	Fin7Args f;
	memset(&f, 0, sizeof(f));
	f.fin = fin.data();
	f.chain = chain.data();
	f.num_spans = nspans;
	f.nblk = nblk;
	f.evbuf = evbuf.data();
	f.ev_cap = ev_cap;
	f.ordered = ordered.data();
	f.span_pre = pre.data();
	f.blk = blk.data();
	f.totals = totals;
	f.ctrl = ctrl;
	f.len = len;
	f.base = base;
	f.nal_start = nal_start;
	f.nal_end = nal_end;
	f.nal_rbsp = nal_rbsp;
	f.nal_rbsp_len = nal_rbsp_len;
	f.nal_cap = nal_cap;
	f.result = result;
	f.has_right = a.has_right;
	f.strip = rbsp ? 1 : 0;
	f.assume_in = (uint32_t)assume_in;
	if (rows == 1) emu_scan7_run<1>(a, f, rbsp != NULL, ev_cap);
	else if (rows == 2) emu_scan7_run<2>(a, f, rbsp != NULL, ev_cap);
	else if (rows == 6) emu_scan7_run<6>(a, f, rbsp != NULL, ev_cap);
	else emu_scan7_run<8>(a, f, rbsp != NULL, ev_cap);
	for (size_t i = 0; i < sizeof(ctrl) / 4; i++)
		if (ctrl[i] != 0)
			return -2; /* control words (ticket counters included) not re-armed */
	return 0;
}

/* h264_byte_stream_to_avcc on the emulator: AVCC-mode scan + ordering + length write, in place */
extern "C" int emu_avcc(uint8_t *data, uint64_t len, uint64_t ev_cap, uint64_t *pos, uint64_t *count, int rows)
{
	using namespace annexb7;
	if (len == 0)
		return -1;
	const uint64_t span = (uint64_t)rows * 512;
	const uint32_t nspans = (uint32_t)((len + span - 1) / span);
	const uint32_t nblk = (nspans + kFinT - 1) / kFinT;
	std::vector<uint64_t> fin(nspans, ~0ull), pre(nspans, ~0ull), blk(2 * (size_t)nblk, ~0ull);
	std::vector<uint64_t> evbuf(2 * (ev_cap ? ev_cap : 1), 0), ordered(3 * (ev_cap ? ev_cap : 1), 0);
	uint64_t totals[8] = {0};
	static uint32_t ctrl[annexb7::kCtrlBytes / 4] = {0};
	uint8_t *buf = (uint8_t *)aligned_alloc(16, (len + 15) & ~15ull);
	memcpy(buf, data, len);
	Scan7Args a;
	memset(&a, 0, sizeof(a));
	a.in = buf;
	a.len = len;
	a.fin = fin.data();
	a.ctrl = ctrl;
	a.evbuf = evbuf.data();
	a.ev_cap = ev_cap;
	a.num_spans = nspans;
	a.halo_left = 0xffffffffu;
	a.right[0] = a.right[1] = 0xff;
	Fin7Args f;
	memset(&f, 0, sizeof(f));
	f.fin = fin.data();
	f.num_spans = nspans;
	f.nblk = nblk;
	f.evbuf = evbuf.data();
	f.ev_cap = ev_cap;
	f.ordered = ordered.data();
	f.span_pre = pre.data();
	f.blk = blk.data();
	f.totals = totals;
	f.ctrl = ctrl;
	const uint32_t nctas = (nspans + kW - 1) / kW;
	dim3 grid(nctas < 3 ? nctas : 3), block(kT), gs(nblk), bs(kFinT), b256(256), go((nspans + 255) / 256), g2(2), one(1);
	if (rows == 1) EMU_LAUNCH((scan7_only_kernel<1, 1, true>), grid, block, a);
	else if (rows == 2) EMU_LAUNCH((scan7_only_kernel<2, 1, true>), grid, block, a);
	else EMU_LAUNCH((scan7_only_kernel<8, 1, true>), grid, block, a);
	EMU_LAUNCH((fin7_spans), gs, bs, f);
	EMU_LAUNCH((fin7_order), go, b256, f);
	EMU_LAUNCH((avcc_write_kernel), g2, b256, f.ordered, f.totals, ev_cap, buf, len, pos, count);
	EMU_LAUNCH((avcc_rearm_kernel), one, one, ctrl);
	memcpy(data, buf, len);
	free(buf);
	for (int i = 0; i < 7; i++)
		if (ctrl[i] != 0)
			return -2;
	return 0;
}

/* in-place RBSP (annexb_scan7.cuh): cpt % 10 = 512-byte rows per span (1, 2, 6 or 8) */
extern "C" int emu_split_strip_inplace(const uint8_t *in, uint64_t len, uint64_t base,
				       const struct h264gpu_shard_edge *edge, uint8_t *rbsp,
				       uint64_t *nal_start, uint64_t *nal_end, uint64_t *nal_rbsp,
				       uint64_t *nal_rbsp_len, uint64_t nal_cap,
				       struct h264gpu_scan_result *result, int cpt, uint64_t ev_cap)
{
	if (len == 0)
		return -1;
	memset(result, 0xff, sizeof(*result));
	uint8_t *buf = (uint8_t *)aligned_alloc(16, (len + 15) & ~15ull);
	memcpy(buf, in, len);
	const int rc = emu_scan7(buf, len, base, edge, rbsp, nal_start, nal_end, nal_rbsp, nal_rbsp_len, nal_cap,
				 result, cpt % 10, ev_cap);
	free(buf);
	return rc;
}

extern "C" int emu_frame(const uint8_t *rbsp, uint64_t len, const uint64_t *off, uint64_t n,
			 int sc_len, uint8_t *out, uint64_t out_cap, uint64_t *out_off,
			 uint64_t *total, int items)
{
	using namespace frame;
	/* 90 + rows: frame8_kernel (tiles of gen 6, chain of gen 7) with 1/2/4/8 rows per warp */
	if (items > 90) {
		const int rows = items - 90;
		const uint64_t tile8 = (uint64_t)kBlock * rows * 16;
		uint32_t nt = (uint32_t)((len + tile8 - 1) / tile8);
		if (nt == 0)
			nt = 1;
		std::vector<uint64_t> desc(nt, ~0ull), first(nt + 1, ~0ull), gw((nt >> 5) + 1, ~0ull), sw((nt >> 10) + 1, ~0ull),
			sp((nt >> 10) + 1, ~0ull);
		std::vector<uint32_t> tail(nt, ~0u);
		uint32_t ticket = 0xffffffffu;
		uint8_t *buf = (uint8_t *)aligned_alloc(16, ((len + 15) & ~15ull) + 16);
		memcpy(buf, rbsp, len);
		FrameArgs a;
		memset(&a, 0, sizeof(a));
		a.rbsp = buf;
		a.len = len;
		a.off = off;
		a.n = n;
		a.sc_len = (uint32_t)sc_len;
		a.out = out;
		a.out_cap = out_cap;
		a.out_off = out_off;
		a.total = total;
		a.desc = desc.data();
		a.ticket = &ticket;
		a.first = first.data();
		a.tail = tail.data();
		a.num_tiles = nt;
		a.group_w = gw.data();
		a.super_w = sw.data();
		a.super_p = sp.data();
		dim3 pgrid((nt + 1 + 127) / 128), pblock(128), block(kBlock), grid(nt < 2 ? nt : 2);
		switch (rows) {
		case 1:
			EMU_LAUNCH((frame8::frame8_prepass<1>), pgrid, pblock, a);
			EMU_LAUNCH((frame8::frame8_kernel<1, 1>), grid, block, a);
			break;
		case 2:
			EMU_LAUNCH((frame8::frame8_prepass<2>), pgrid, pblock, a);
			EMU_LAUNCH((frame8::frame8_kernel<2, 1>), grid, block, a);
			break;
		case 4:
			EMU_LAUNCH((frame8::frame8_prepass<4>), pgrid, pblock, a);
			EMU_LAUNCH((frame8::frame8_kernel<4, 1>), grid, block, a);
			break;
		default:
			EMU_LAUNCH((frame8::frame8_prepass<8>), pgrid, pblock, a);
			EMU_LAUNCH((frame8::frame8_kernel<8, 1>), grid, block, a);
			break;
		}
		free(buf);
		return 0;
	}
	/* 70 + rows: frame7_kernel with 1/2/4/6/8 rows per span, three warps per CTA, the bytes of a
	 * span staged twice; 80 + rows: staged once (two buffers per warp) */
	if (items > 70) {
		const int nbuf = items > 80 ? 2 : 1;
		const int rows = items > 80 ? items - 80 : items - 70;
		const uint64_t span = 512ull * rows;
		uint32_t ns = (uint32_t)((len + span - 1) / span);
		if (ns == 0)
			ns = 1;
		std::vector<uint64_t> desc(ns, ~0ull), first(ns + 1, ~0ull), gw((ns >> 5) + 1, ~0ull),
			sw((ns >> 10) + 1, ~0ull), sp((ns >> 10) + 1, ~0ull);
		uint32_t ticket = 0xffffffffu;
		uint8_t *buf = (uint8_t *)aligned_alloc(16, ((len + 15) & ~15ull) + 16);
		memcpy(buf, rbsp, len);
		FrameArgs a;
		memset(&a, 0, sizeof(a));
		a.rbsp = buf;
		a.len = len;
		a.off = off;
		a.n = n;
		a.sc_len = (uint32_t)sc_len;
		a.out = out;
		a.out_cap = out_cap;
		a.out_off = out_off;
		a.total = total;
		a.desc = desc.data();
		a.ticket = &ticket;
		a.first = first.data();
		a.num_tiles = ns;
		a.group_w = gw.data();
		a.super_w = sw.data();
		a.super_p = sp.data();
		dim3 pgrid((ns + 1 + 127) / 128), pblock(128), block(96), grid(2);
		switch (rows) {
		case 1:
			EMU_LAUNCH((frame7::frame7_prepass<1>), pgrid, pblock, a);
			if (nbuf == 2)
				EMU_LAUNCH((frame7::frame7_kernel<1, 3, 1, 2>), grid, block, a);
			else
				EMU_LAUNCH((frame7::frame7_kernel<1, 3, 1, 1>), grid, block, a);
			break;
		case 2:
			EMU_LAUNCH((frame7::frame7_prepass<2>), pgrid, pblock, a);
			if (nbuf == 2)
				EMU_LAUNCH((frame7::frame7_kernel<2, 3, 1, 2>), grid, block, a);
			else
				EMU_LAUNCH((frame7::frame7_kernel<2, 3, 1, 1>), grid, block, a);
			break;
		case 4:
			EMU_LAUNCH((frame7::frame7_prepass<4>), pgrid, pblock, a);
			if (nbuf == 2)
				EMU_LAUNCH((frame7::frame7_kernel<4, 3, 1, 2>), grid, block, a);
			else
				EMU_LAUNCH((frame7::frame7_kernel<4, 3, 1, 1>), grid, block, a);
			break;
		case 6:
			EMU_LAUNCH((frame7::frame7_prepass<6>), pgrid, pblock, a);
			if (nbuf == 2)
				EMU_LAUNCH((frame7::frame7_kernel<6, 3, 1, 2>), grid, block, a);
			else
				EMU_LAUNCH((frame7::frame7_kernel<6, 3, 1, 1>), grid, block, a);
			break;
		default:
			EMU_LAUNCH((frame7::frame7_prepass<8>), pgrid, pblock, a);
			if (nbuf == 2)
				EMU_LAUNCH((frame7::frame7_kernel<8, 3, 1, 2>), grid, block, a);
			else
				EMU_LAUNCH((frame7::frame7_kernel<8, 3, 1, 1>), grid, block, a);
			break;
		}
		free(buf);
		return 0;
	}
	/* items (or 60 + items): frame6_kernel with 1/2/4/8 rows per warp */
	if (items > 60)
		items -= 60;
	const uint64_t tile = (uint64_t)kBlock * items * 16;
	uint32_t ntiles = (uint32_t)((len + tile - 1) / tile);
	if (ntiles == 0)
		ntiles = 1;
	std::vector<uint64_t> desc(ntiles, ~0ull), first(ntiles + 1, 0);
	std::vector<uint32_t> tail(ntiles, 0);
	uint32_t ticket = 0xffffffffu;
	uint8_t *buf = (uint8_t *)aligned_alloc(16, ((len + 15) & ~15ull) + 16);
	memcpy(buf, rbsp, len);
	FrameArgs a;
	memset(&a, 0, sizeof(a));
	a.rbsp = buf;
	a.len = len;
	a.off = off;
	a.n = n;
	a.sc_len = (uint32_t)sc_len;
	a.out = out;
	a.out_cap = out_cap;
	a.out_off = out_off;
	a.total = total;
	a.desc = desc.data();
	a.ticket = &ticket;
	a.first = first.data();
	a.tail = tail.data();
	a.num_tiles = ntiles;
	dim3 pgrid((ntiles + 1 + 127) / 128), pblock(128), block(kBlock);
	/* persistent CTAs: a few of them share the tiles */
	dim3 g6(ntiles < 3 ? ntiles : 3);
	if (items == 1) {
		EMU_LAUNCH((frame_prepass<1>), pgrid, pblock, a);
		EMU_LAUNCH((frame6::frame6_kernel<1, 1>), g6, block, a);
	} else if (items == 2) {
		EMU_LAUNCH((frame_prepass<2>), pgrid, pblock, a);
		EMU_LAUNCH((frame6::frame6_kernel<2, 1>), g6, block, a);
	} else if (items == 4) {
		EMU_LAUNCH((frame_prepass<4>), pgrid, pblock, a);
		EMU_LAUNCH((frame6::frame6_kernel<4, 1>), g6, block, a);
	} else {
		EMU_LAUNCH((frame_prepass<8>), pgrid, pblock, a);
		EMU_LAUNCH((frame6::frame6_kernel<8, 1>), g6, block, a);
	}
	free(buf);
	return 0;
}

/* K4: the same step function the kernel's lanes run, one slice at a time */
#include "cavlc_steps.cuh"

extern "C" int emu_cavlc_steps(const uint8_t *stream, uint64_t stream_len,
			       const struct h264gpu_slice_params *params, uint32_t n_slices,
			       struct h264gpu_mb_record *records, struct h264gpu_slice_result *results,
			       struct h264_mb_syntax *syn, const uint8_t *group_maps)
{
	for (uint32_t i = 0; i < n_slices; i++) {
		const h264gpu_slice_params &sp = params[i];
		std::vector<uint8_t> ring(((size_t)sp.pic_width_in_mbs + 1) * CAVLC2_RING_SLOT + 64, 0xEE);
		uint32_t sm[CAVLC2_SM_WORDS];
		memset(sm, 0xEE, sizeof(sm));
		if (syn)
			cavlc2::parse_slice<true>(stream, stream_len, sp, ring.data(), records + sp.mb_out_off, results[i],
						  syn + sp.mb_out_off, group_maps, sm);
		else
			cavlc2::parse_slice<false>(stream, stream_len, sp, ring.data(), records + sp.mb_out_off, results[i],
						   nullptr, group_maps, sm);
	}
	return 0;
}

/* N4: concealment slice synthesis, both passes of the kernel; the scan in between on the host */
#include "conceal.cuh"

/* N4: bulk slice header patches */
extern "C" int emu_patch_headers(uint8_t *stream, uint64_t stream_len, const struct h264gpu_hdr_patch *patches, uint32_t n)
{
	if (n == 0)
		return 0;
	dim3 grid((unsigned)(((uint64_t)n * 64 + 255) / 256)), block(256);
	EMU_LAUNCH((conceal::patch_headers_kernel), grid, block, stream, stream_len, patches, n);
	return 0;
}


extern "C" int emu_conceal(const struct h264gpu_conceal_params *params, uint32_t n, const uint8_t *hdr,
			   uint8_t *out, uint64_t cap, uint64_t *off)
{
	conceal::ConcealArgs a;
	a.params = params;
	a.n = n;
	a.hdr = hdr;
	a.off = off;
	a.out = out;
	a.cap = cap;
	dim3 grid((n + 127) / 128), block(128);
	off[0] = 0;
	EMU_LAUNCH((conceal::conceal_kernel<false>), grid, block, a);
	for (uint32_t k = 0; k < n; k++)
		off[k + 1] += off[k];
	EMU_LAUNCH((conceal::conceal_kernel<true>), grid, block, a);
	return 0;
}
