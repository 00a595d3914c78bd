/*
 * gpu_compat.h — the few primitives whose spelling differs between the real
 * sm_100a build (inline PTX) and the CPU logic-emulation build used by the
 * non-GPU tests (H264_EMU, tests/emu/cuda_emu.h).  The product is always built
 * by nvcc without H264_EMU.
 */
#ifndef H264GPU_COMPAT_H
#define H264GPU_COMPAT_H

#include <stdint.h>

#ifdef H264_EMU
#include "cuda_emu.h"
#define H264_HD
static inline uint4 ldg_stream16(const void *p) { return *(const uint4 *)p; }
static inline uint32_t ldg_u32(const void *p) { return *(const uint32_t *)p; }
static inline uint64_t ld_relaxed_u64(const uint64_t *p) { return *(const volatile uint64_t *)p; }
static inline void st_relaxed_u64(uint64_t *p, uint64_t v) { *(volatile uint64_t *)p = v; }
static inline void stg_stream16(void *p, uint4 v) { *(uint4 *)p = v; }
static inline void stg_stream16_free(void *p, uint4 v) { *(uint4 *)p = v; }
static inline void bulk_load_start(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
	(void)bar;
	memcpy(sdst, gsrc, bytes);
}
static inline void bulk_load_wait(uint64_t *bar) { (void)bar; }
static inline void bulk_bar_init(uint64_t *bar) { (void)bar; }
static inline void bulk_load_issue(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
	(void)bar;
	memcpy(sdst, gsrc, bytes);
}
static inline void bulk_load_wait_parity(uint64_t *bar, uint32_t parity) { (void)bar; (void)parity; }
static inline void l2_prefetch(const void *gsrc, uint32_t bytes) { (void)gsrc; (void)bytes; }
/* a polling loop gives the other fibers (warps of the CTA) a turn */
static inline void spin_pause(unsigned ns) { (void)ns; emu::yield(); }
static inline uint4 ldcg16(const void *p) { return *(const uint4 *)p; }
static inline uint32_t ldcg_u32(const void *p) { return *(const volatile uint32_t *)p; }
static inline uint32_t ld_relaxed_u32(const uint32_t *p) { return *(const volatile uint32_t *)p; }
static inline void st_relaxed_u32(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }
#else
#include <cuda_runtime.h>
#define H264_HD __host__ __device__

/* streaming 16-byte load: read-only path, do not keep the line in L1 */
__device__ __forceinline__ uint4 ldg_stream16(const void *p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
		     : "l"(p));
	return r;
}
__device__ __forceinline__ uint32_t ldg_u32(const void *p)
{
	uint32_t r;
	asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(r) : "l"(p));
	return r;
}
/* single-word, self-validating tile descriptors: relaxed gpu-scope accesses */
__device__ __forceinline__ uint64_t ld_relaxed_u64(const uint64_t *p)
{
	uint64_t v;
	asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed_u64(uint64_t *p, uint64_t v)
{
	asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
/* streaming 16-byte store */
__device__ __forceinline__ void stg_stream16(void *p, uint4 v)
{
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p),
		     "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
		     : "memory");
}
/* the same without the compiler-level memory barrier: output bytes nobody in this kernel reads
 * back, so loads of the next iterations may be hoisted above the store */
__device__ __forceinline__ void stg_stream16_free(void *p, uint4 v)
{
	asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p),
		     "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
/* loads that bypass the (non-coherent) L1: data another CTA wrote */
__device__ __forceinline__ uint4 ldcg16(const void *p)
{
	uint4 r;
	asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
		     : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
		     : "l"(p)
		     : "memory");
	return r;
}
__device__ __forceinline__ uint32_t ldcg_u32(const void *p)
{
	uint32_t r;
	asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
	return r;
}
__device__ __forceinline__ uint32_t ld_relaxed_u32(const uint32_t *p)
{
	uint32_t v;
	asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_relaxed_u32(uint32_t *p, uint32_t v)
{
	asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
/* back off inside a polling loop so the spinning warp leaves the issue slots alone */
__device__ __forceinline__ void spin_pause(unsigned ns)
{
	asm volatile("nanosleep.u32 %0;" ::"r"(ns));
}
/*
 * Bulk asynchronous copy global -> shared (TMA, 1-D form) completing on an
 * mbarrier: one thread arms the barrier with the byte count and issues the copy;
 * every consumer waits on phase 0.  dst/src 16-byte aligned, bytes % 16 == 0.
 */
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
	return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void bulk_load_start(void *sdst, const void *gsrc, uint32_t bytes,
						uint64_t *bar)
{
	const uint32_t b = smem_u32(bar), d = smem_u32(sdst);
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes)
		     : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
		     "[%0], [%1], %2, [%3];" ::"r"(d),
		     "l"(gsrc), "r"(bytes), "r"(b)
		     : "memory");
}
__device__ __forceinline__ void bulk_load_wait(uint64_t *bar)
{
	const uint32_t b = smem_u32(bar);
	asm volatile("{\n"
		     ".reg .pred p;\n"
		     "BULK_WAIT:\n"
		     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
		     "@p bra BULK_DONE;\n"
		     "bra BULK_WAIT;\n"
		     "BULK_DONE:\n"
		     "}" ::"r"(b)
		     : "memory");
}
/* the same split up for a barrier that is reused tile after tile (persistent CTAs): init once,
 * then arm + copy per tile and wait on the phase parity (0, 1, 0, ...) */
__device__ __forceinline__ void bulk_bar_init(uint64_t *bar)
{
	const uint32_t b = smem_u32(bar);
	asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_issue(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
	const uint32_t b = smem_u32(bar), d = smem_u32(sdst);
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes)
		     : "memory");
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
		     "[%0], [%1], %2, [%3];" ::"r"(d),
		     "l"(gsrc), "r"(bytes), "r"(b)
		     : "memory");
}
__device__ __forceinline__ void bulk_load_wait_parity(uint64_t *bar, uint32_t parity)
{
	const uint32_t b = smem_u32(bar);
	asm volatile("{\n"
		     ".reg .pred p;\n"
		     "BULK_WAITP:\n"
		     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		     "@p bra BULK_DONEP;\n"
		     "bra BULK_WAITP;\n"
		     "BULK_DONEP:\n"
		     "}" ::"r"(b),
		     "r"(parity)
		     : "memory");
}
/* pull a range of global memory into L2 ahead of its bulk load */
__device__ __forceinline__ void l2_prefetch(const void *gsrc, uint32_t bytes)
{
	asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
#endif

#define FULL_MASK 0xffffffffu

/* atomic add on a shared-memory word (the generic-address form costs a space check and a loop) */
#ifdef H264_EMU
static inline uint32_t smem_atomic_add(uint32_t *p, uint32_t v)
{
	const uint32_t o = *p;
	*p = o + v;
	return o;
}
#else
__device__ __forceinline__ uint32_t smem_atomic_add(uint32_t *p, uint32_t v)
{
	uint32_t o;
	asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(smem_u32(p)), "r"(v) : "memory");
	return o;
}
#endif

/* OR of a value over the warp (all lanes take part) */
#ifdef H264_EMU
static inline uint32_t warp_or(uint32_t v)
{
	for (int d = 16; d >= 1; d >>= 1)
		v |= __shfl_xor_sync(FULL_MASK, v, d);
	return v;
}
#else
__device__ __forceinline__ uint32_t warp_or(uint32_t v)
{
	return __reduce_or_sync(FULL_MASK, v);
}
#endif

/* sum of a 32-bit value over the warp (all lanes take part) */
#ifdef H264_EMU
static inline uint32_t warp_add(uint32_t v)
{
	for (int d = 16; d >= 1; d >>= 1)
		v += __shfl_xor_sync(FULL_MASK, v, d);
	return v;
}
#else
__device__ __forceinline__ uint32_t warp_add(uint32_t v)
{
	return __reduce_add_sync(FULL_MASK, v);
}
#endif

#endif /* H264GPU_COMPAT_H */
