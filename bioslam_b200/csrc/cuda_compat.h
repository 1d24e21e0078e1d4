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
// Tensor Memory (see the sm_100a versions below): the emulator keeps one 128-lane x 512-column array of 32-bit cells per block
__device__ __forceinline__ void tmem_alloc512(unsigned* smem_slot) { *smem_slot = 0u; }
__device__ __forceinline__ void tmem_dealloc512(unsigned) {}
__device__ __forceinline__ void tmem_fence_before_sync() {}
__device__ __forceinline__ void tmem_fence_after_sync() {}
__device__ __forceinline__ void tmem_wait_st() {}
__device__ __forceinline__ void tmem_st4(unsigned taddr, const double (&v)[4]) { std::memcpy(emu::tmem_cell((int)(taddr >> 16) + emu::lane_id(), (int)(taddr & 0xffffu)), v, 32); }
__device__ __forceinline__ void tmem_ld4(unsigned taddr, double (&v)[4]) { std::memcpy(v, emu::tmem_cell((int)(taddr >> 16) + emu::lane_id(), (int)(taddr & 0xffffu)), 32); }
struct TmemRaw { int r[8]; };
__device__ __forceinline__ void tmem_ld4_issue(unsigned taddr, TmemRaw& raw) { std::memcpy(raw.r, emu::tmem_cell((int)(taddr >> 16) + emu::lane_id(), (int)(taddr & 0xffffu)), 32); }
__device__ __forceinline__ void tmem_ld4_wait(TmemRaw& raw, double (&v)[4]) { std::memcpy(v, raw.r, 32); }
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
// Tensor Memory (TMEM: 128 lanes x 512 columns x 32 bit per SM, reached through tcgen05.ld / tcgen05.st only) used as a
// WARP-PRIVATE scratch for fp64 row tiles: warp w owns lanes [32 (w & 3), +32); with the 32x32b shape thread l of the warp reads /
// writes N consecutive columns of lane 32 (w & 3) + l, i.e. every thread gets its own data back.  The sweep keeps its panel row
// tiles there as accumulator-layout fragments (4 doubles = 8 columns per 8-column block) and feeds them back to the fp64 MMA as A
// operands (tools/ubench/tmem_probe.cu: 29 TFLOP/s from 8 independent warps, against 33 with the same tile in shared memory).
// taddr = (lane << 16) | column.  alloc / dealloc are executed by ONE warp; the base address arrives through shared memory.
__device__ __forceinline__ void tmem_alloc512(unsigned* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((unsigned)__cvta_generic_to_shared(smem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(unsigned base) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory"); }
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st4(unsigned taddr, const double (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__double2loint(v[0])), "r"(__double2hiint(v[0])),
                 "r"(__double2loint(v[1])), "r"(__double2hiint(v[1])), "r"(__double2loint(v[2])), "r"(__double2hiint(v[2])), "r"(__double2loint(v[3])),
                 "r"(__double2hiint(v[3])) : "memory");
}
__device__ __forceinline__ void tmem_ld4(unsigned taddr, double (&v)[4]) {
    int r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5),
                 "=r"(r6), "=r"(r7) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v[0] = __hiloint2double(r1, r0); v[1] = __hiloint2double(r3, r2); v[2] = __hiloint2double(r5, r4); v[3] = __hiloint2double(r7, r6);
}
// the same in two halves, for software pipelining: the load of the NEXT fragment is in flight while the MMAs of the current one
// issue.  The destination registers are outputs of the issue statement and in/out operands of the wait statement, so the compiler
// can neither read nor move them in between (tcgen05.ld writes them asynchronously until tcgen05.wait::ld).
struct TmemRaw { int r[8]; };
__device__ __forceinline__ void tmem_ld4_issue(unsigned taddr, TmemRaw& raw) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(raw.r[0]), "=r"(raw.r[1]), "=r"(raw.r[2]), "=r"(raw.r[3]),
                 "=r"(raw.r[4]), "=r"(raw.r[5]), "=r"(raw.r[6]), "=r"(raw.r[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4_wait(TmemRaw& raw, double (&v)[4]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(raw.r[0]), "+r"(raw.r[1]), "+r"(raw.r[2]), "+r"(raw.r[3]), "+r"(raw.r[4]), "+r"(raw.r[5]),
                 "+r"(raw.r[6]), "+r"(raw.r[7]) :: "memory");
    v[0] = __hiloint2double(raw.r[1], raw.r[0]); v[1] = __hiloint2double(raw.r[3], raw.r[2]); v[2] = __hiloint2double(raw.r[5], raw.r[4]);
    v[3] = __hiloint2double(raw.r[7], raw.r[6]);
}
#endif
