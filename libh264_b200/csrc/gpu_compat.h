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
#endif

#define FULL_MASK 0xffffffffu

#endif /* H264GPU_COMPAT_H */
