// TEST TOOL ONLY -- a minimal CUDA-thread emulator so the .cu sources under bioslam_b200/csrc can be compiled with
// plain g++ and executed on a CPU for debugging (indexing, barrier placement, host-side control flow) in the
// build container, which has no GPU.  One fiber per CUDA thread, blocks run one after another, __syncthreads() and
// warp shuffles are real rendezvous points (a barrier that not every thread reaches is reported as a deadlock).
//
// This is NOT a product path and NOT a fallback: nothing in bioslam_b200/, bench.py or __graft_entry__.py loads a
// library built with this header.  Only tests/test_emu_*.py (marked "not gpu") build and load it, as
// tests/cudaemu/libbioslam_b200_emu.so.
#pragma once
#ifndef __CUDACC__
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <functional>
#include <vector>

#define B200_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct emu_uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
extern emu_uint3 threadIdx, blockIdx, blockDim, gridDim;

typedef int cudaError_t;
typedef struct emu_stream_s* cudaStream_t;
struct emu_event_s { std::chrono::steady_clock::time_point t; };
typedef emu_event_s* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1 };

inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emu error" : "no error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::calloc(n ? n : 1, 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event_s(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
struct cudaFuncAttributes { size_t sharedSizeBytes = 0; int numRegs = 0; };
template <class F> inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { *a = cudaFuncAttributes(); return cudaSuccess; }

namespace emu {
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);
void sync_block();
void sync_warp();
void bar_sync(int id, int count);     // named barrier (bar.sync id, count)
void bar_arrive(int id, int count);   // named barrier, non-blocking arrival (bar.arrive id, count)
double shfl_exchange(double v, int src_lane);   // generic lane exchange within the caller's warp
unsigned ballot(int pred);
void dmma884(double& c0, double& c1, double a, double b);   // warp-collective fp64 MMA m8n8k4 (row.col)
unsigned* tmem_cell(int lane, int col);   // Tensor Memory of the running block: 128 lanes x 512 columns of 32-bit cells
extern char* dyn_smem;
extern long long launch_counter;
int lane_id();
}  // namespace emu

#define B200_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define B200_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem)

inline void __syncthreads() { emu::sync_block(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::sync_warp(); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline double __shfl_sync(unsigned, double v, int src, int width = 32) {
    int lane = emu::lane_id();
    int base = lane - (lane % width);
    return emu::shfl_exchange(v, base + (src % width));
}
inline int __shfl_sync(unsigned m, int v, int src, int width = 32) { return (int)__shfl_sync(m, (double)v, src, width); }
inline double __shfl_xor_sync(unsigned, double v, int mask, int = 32) { return emu::shfl_exchange(v, emu::lane_id() ^ mask); }
inline int __shfl_xor_sync(unsigned m, int v, int mask, int w = 32) { return (int)__shfl_xor_sync(m, (double)v, mask, w); }
inline double __shfl_down_sync(unsigned, double v, unsigned d, int = 32) {
    int lane = emu::lane_id();
    return emu::shfl_exchange(v, lane + (int)d < 32 ? lane + (int)d : lane);
}
inline unsigned __ballot_sync(unsigned, int pred) { return emu::ballot(pred); }
inline double atomicAdd(double* a, double v) { double o = *a; *a = o + v; return o; }
inline int atomicAdd(int* a, int v) { int o = *a; *a = o + v; return o; }
inline int atomicMax(int* a, int v) { int o = *a; if (v > o) *a = v; return o; }
inline int atomicMin(int* a, int v) { int o = *a; if (v < o) *a = v; return o; }
inline int atomicCAS(int* a, int cmp, int v) { int o = *a; if (o == cmp) *a = v; return o; }
inline int atomicExch(int* a, int v) { int o = *a; *a = v; return o; }
inline void sincos(double x, double* s, double* c) { *s = std::sin(x); *c = std::cos(x); }
inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double __dsqrt_rn(double x) { return std::sqrt(x); }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
using std::fabs; using std::sqrt; using std::sin; using std::cos; using std::acos; using std::atan2; using std::tanh; using std::tan;
using std::fmax; using std::fmin;
#endif
