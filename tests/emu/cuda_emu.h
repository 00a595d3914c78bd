/*
 * cuda_emu.h — a tiny SIMT emulator so the CUDA kernels' LOGIC can be unit-tested
 * on a machine without a GPU (the dev container and the driver's CPU test run).
 *
 * TEST INFRASTRUCTURE ONLY.  The product never compiles with H264_EMU.
 *
 * Model: one CTA at a time, CTAs in launch order; every CUDA thread is a ucontext
 * fiber; warp collectives (__shfl*_sync, __ballot_sync, __any_sync, ...) and
 * __syncthreads are rendezvous points between fibers.  Because CTAs run to
 * completion in ticket order, decoupled look-back never has to spin.
 * It checks indexing / prefix / masking logic, not memory-model races.
 */
#ifndef CUDA_EMU_H
#define CUDA_EMU_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
	unsigned x, y, z;
	dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
	uint4 r = {x, y, z, w};
	return r;
}
static inline uint2 make_uint2(uint32_t x, uint32_t y)
{
	uint2 r = {x, y};
	return r;
}

namespace emu {

struct Fiber {
	ucontext_t ctx;
	char *stack;
	uint3 tid;
	bool done;
};

struct WarpState {
	unsigned arrived;
	unsigned gen;
	uint64_t vals[2][32];
};

struct State {
	ucontext_t sched;
	std::vector<Fiber> fibers;
	std::vector<WarpState> warps;
	int cur;
	unsigned alive;
	unsigned bar_arrived, bar_gen;
	std::function<void()> body;
	unsigned long switches_without_progress;
};

inline State &st()
{
	static State s;
	return s;
}

inline uint3 &tidx() { static uint3 v; return v; }
inline uint3 &bidx() { static uint3 v; return v; }
inline dim3 &bdim() { static dim3 v; return v; }
inline dim3 &gdim() { static dim3 v; return v; }

inline void yield()
{
	State &s = st();
	if (++s.switches_without_progress > 50000000ul) {
		fprintf(stderr, "cuda_emu: deadlock (all fibers waiting)\n");
		abort();
	}
	swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

inline void progress() { st().switches_without_progress = 0; }

inline void trampoline()
{
	State &s = st();
	s.body();
	s.fibers[s.cur].done = true;
	s.alive--;
	progress();
	swapcontext(&s.fibers[s.cur].ctx, &s.sched);
}

inline unsigned linear_tid()
{
	return tidx().x + bdim().x * (tidx().y + bdim().y * tidx().z);
}

/* all lanes named in mask deposit v, everyone gets the 32 values back */
inline void exchange(unsigned mask, uint64_t v, uint64_t out[32])
{
	State &s = st();
	unsigned lt = linear_tid();
	WarpState &w = s.warps[lt / 32];
	unsigned lane = lt % 32;
	unsigned g = w.gen, b = g & 1;
	w.vals[b][lane] = v;
	w.arrived++;
	progress();
	if (w.arrived == (unsigned)__builtin_popcount(mask)) {
		w.arrived = 0;
		w.gen++;
	} else {
		while (w.gen == g)
			yield();
	}
	memcpy(out, w.vals[b], sizeof(uint64_t) * 32);
}

inline void syncthreads()
{
	State &s = st();
	unsigned g = s.bar_gen;
	s.bar_arrived++;
	progress();
	if (s.bar_arrived == s.alive) {
		s.bar_arrived = 0;
		s.bar_gen++;
	} else {
		while (s.bar_gen == g)
			yield();
	}
}

template <class F> inline void launch(dim3 grid, dim3 block, F f)
{
	State &s = st();
	const size_t kStack = 256 * 1024;
	unsigned nthreads = block.x * block.y * block.z;
	gdim() = grid;
	bdim() = block;
	s.body = f;
	s.fibers.resize(nthreads);
	for (unsigned i = 0; i < nthreads; i++)
		s.fibers[i].stack = (char *)malloc(kStack);
	for (unsigned bz = 0; bz < grid.z; bz++)
	for (unsigned by = 0; by < grid.y; by++)
	for (unsigned bx = 0; bx < grid.x; bx++) {
		uint3 b = {bx, by, bz};
		bidx() = b;
		s.warps.assign((nthreads + 31) / 32, WarpState());
		s.alive = nthreads;
		s.bar_arrived = 0;
		s.bar_gen = 0;
		s.switches_without_progress = 0;
		for (unsigned i = 0; i < nthreads; i++) {
			Fiber &fb = s.fibers[i];
			getcontext(&fb.ctx);
			fb.ctx.uc_stack.ss_sp = fb.stack;
			fb.ctx.uc_stack.ss_size = kStack;
			fb.ctx.uc_link = &s.sched;
			fb.done = false;
			fb.tid.x = i % block.x;
			fb.tid.y = (i / block.x) % block.y;
			fb.tid.z = i / (block.x * block.y);
			makecontext(&fb.ctx, (void (*)())trampoline, 0);
		}
		while (s.alive > 0) {
			for (unsigned i = 0; i < nthreads; i++) {
				if (s.fibers[i].done)
					continue;
				s.cur = (int)i;
				tidx() = s.fibers[i].tid;
				swapcontext(&s.sched, &s.fibers[i].ctx);
			}
		}
	}
	for (unsigned i = 0; i < nthreads; i++)
		free(s.fibers[i].stack);
}

} /* namespace emu */

#define threadIdx (emu::tidx())
#define blockIdx (emu::bidx())
#define blockDim (emu::bdim())
#define gridDim (emu::gdim())

static inline void __syncthreads() { emu::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu)
{
	uint64_t o[32];
	emu::exchange(mask, 0, o);
}
static inline void __threadfence() {}

template <class T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
	uint64_t o[32], in = 0;
	static_assert(sizeof(T) <= 8, "shfl type");
	memcpy(&in, &v, sizeof(T));
	emu::exchange(mask, in, o);
	unsigned lane = emu::linear_tid() % 32;
	unsigned base = lane & ~(unsigned)(width - 1);
	T r;
	memcpy(&r, &o[base + ((unsigned)src & (unsigned)(width - 1))], sizeof(T));
	return r;
}
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32)
{
	uint64_t o[32], in = 0;
	memcpy(&in, &v, sizeof(T));
	emu::exchange(mask, in, o);
	unsigned lane = emu::linear_tid() % 32;
	unsigned base = lane & ~(unsigned)(width - 1);
	T r = v;
	if (lane - base >= d)
		memcpy(&r, &o[lane - d], sizeof(T));
	return r;
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32)
{
	uint64_t o[32], in = 0;
	memcpy(&in, &v, sizeof(T));
	emu::exchange(mask, in, o);
	unsigned lane = emu::linear_tid() % 32;
	unsigned base = lane & ~(unsigned)(width - 1);
	T r = v;
	if (lane - base + d < (unsigned)width)
		memcpy(&r, &o[lane + d], sizeof(T));
	return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int x, int width = 32)
{
	(void)width;
	uint64_t o[32], in = 0;
	memcpy(&in, &v, sizeof(T));
	emu::exchange(mask, in, o);
	unsigned lane = emu::linear_tid() % 32;
	T r;
	memcpy(&r, &o[lane ^ (unsigned)x], sizeof(T));
	return r;
}
static inline unsigned __ballot_sync(unsigned mask, int pred)
{
	uint64_t o[32];
	emu::exchange(mask, pred ? 1 : 0, o);
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if ((mask >> i & 1) && o[i])
			r |= 1u << i;
	return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v)
{
	unsigned r = 0;
	for (int i = 0; i < 32; i++)
		if (v >> i & 1)
			r |= 1u << (31 - i);
	return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned s)
{
	uint64_t v = ((uint64_t)hi << 32) | lo;
	return (unsigned)((v << (s & 31)) >> 32);
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{
	uint64_t v = ((uint64_t)hi << 32) | lo;
	return (unsigned)(v >> (s & 31));
}
static inline unsigned __funnelshift_lc(unsigned lo, unsigned hi, unsigned s)
{
	uint64_t v = ((uint64_t)hi << 32) | lo;
	s = s > 32 ? 32 : s;
	return s == 32 ? lo : (unsigned)((v << s) >> 32);
}
static inline unsigned __funnelshift_rc(unsigned lo, unsigned hi, unsigned s)
{
	uint64_t v = ((uint64_t)hi << 32) | lo;
	s = s > 32 ? 32 : s;
	return s == 32 ? hi : (unsigned)(v >> s);
}
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel)
{
	uint64_t v = ((uint64_t)b << 32) | a;
	unsigned r = 0;
	for (int i = 0; i < 4; i++) {
		unsigned n = (sel >> (4 * i)) & 0xf;
		unsigned byte = (unsigned)(v >> (8 * (n & 7))) & 0xff;
		if (n & 8)
			byte = (byte & 0x80) ? 0xff : 0x00;
		r |= byte << (8 * i);
	}
	return r;
}
template <class T> static inline T atomicAdd(T *p, T v)
{
	T o = *p;
	*p = o + v;
	return o;
}
template <class T> static inline T atomicMin(T *p, T v)
{
	T o = *p;
	if (v < o)
		*p = v;
	return o;
}
template <class T> static inline T atomicMax(T *p, T v)
{
	T o = *p;
	if (v > o)
		*p = v;
	return o;
}
template <class T> static inline T atomicExch(T *p, T v)
{
	T o = *p;
	*p = v;
	return o;
}
template <class T> static inline T atomicCAS(T *p, T cmp, T v)
{
	T o = *p;
	if (o == cmp)
		*p = v;
	return o;
}
static inline void __threadfence_block() {}
template <class T> static inline T atomicOr(T *p, T v)
{
	T o = *p;
	*p = o | v;
	return o;
}
using std::max;
using std::min;

#define EMU_LAUNCH(kernel, grid, block, ...) emu::launch((grid), (block), [&]() { kernel(__VA_ARGS__); })

#endif /* CUDA_EMU_H */
