// Build switch: real CUDA (nvcc, sm_100a) is the product; -DB200_USE_EMU (tests/cudaemu, g++) is a CPU debugging
// harness for the same sources and is never shipped nor loaded by the package.
#pragma once
#ifdef B200_USE_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define B200_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define B200_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char b200_dyn_smem_[]; \
    type* name = reinterpret_cast<type*>(b200_dyn_smem_)
#endif

// asynchronous global->shared copies (LDGSTS on sm_100a); the emulator copies synchronously.
#ifdef B200_USE_EMU
__device__ __forceinline__ void cp_async16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
__device__ __forceinline__ void cp_async8(void* dst, const void* src) { std::memcpy(dst, src, 8); }
__device__ __forceinline__ void cp_async_wait_all() {}
__device__ __forceinline__ void cp_async_commit() {}
__device__ __forceinline__ void cp_async_wait_but_last() {}
__device__ __forceinline__ void prefetch_l2(const void*) {}
__device__ __forceinline__ void named_bar_sync(int id, int count) { emu::bar_sync(id, count); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { emu::bar_arrive(id, count); }
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) { emu::dmma884(c0, c1, a, b); }
#else
// fp64 tensor-core MMA (DMMA on sm_100a): C(8 x 8) += A(8 x 4) B(4 x 8).  Lane (g = lane >> 2, t = lane & 3) supplies A[g][t] and
// B[t][g] and owns C[g][2t], C[g][2t+1].  IEEE fp64 fused multiply-adds: same rounding class as the scalar DFMA path.
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
// software pipelines: close the current group of copies / wait for every group but the most recent one
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_but_last() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// named barriers (ids 1..15; 0 is __syncthreads): producer/consumer hand-over between warp-specialised roles of one CTA.
// count = number of THREADS (multiple of 32) that arrive in total; memory ordering as for bar.sync.
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
#endif
