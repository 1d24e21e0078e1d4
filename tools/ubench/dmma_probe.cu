// Development micro-benchmark (not part of the product): fp64 GEMM micro-kernel variants fed from shared memory on one CTA per SM.
//   v0: 4x4 register tile per thread, LDS.128 operand pairs (the r01 micro-kernel)      v1: mma.sync.m8n8k4.f64, warp tile 16x32
//   v2: mma.sync.m8n8k4.f64, warp tile 32x32       r0 / r1: register-only DFMA / DMMA issue peaks
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
constexpr int KD = 128;
template <int V> __global__ void __launch_bounds__(1024) k(double* out, int reps, int ld) {
    extern __shared__ double sm[];
    double* A = sm;                 // 64 x KD
    double* B = sm + 64 * ld;       // 128 x KD
    for (int i = threadIdx.x; i < (64 + 128) * ld; i += blockDim.x) sm[i] = 1e-3 * ((i * 7 + blockIdx.x) % 13);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double s = 0;
    if (V == 0) {
        const int ty = lane >> 3, tx = lane & 7, m0 = (warp >> 2) << 4, r0 = (warp & 3) << 5;
        double acc[4][4] = {};
        for (int r = 0; r < reps; r++) {
            for (int kk = 0; kk < KD; kk += 2) {
                double2 a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) { a[i] = ld2(A + (m0 + ty + 4 * i) * ld + kk); b[i] = ld2(B + (r0 + tx + 8 * i) * ld + kk); }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i].y, b[j].y, fma(a[i].x, b[j].x, acc[i][j]));
            }
        }
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += acc[i][j];
    } else if (V == 1) {
        const int g = lane >> 2, t = lane & 3, m0 = (warp >> 2) << 4, r0 = (warp & 3) << 5;
        double acc[2][4][2] = {};
        for (int r = 0; r < reps; r++) {
#pragma unroll 4
            for (int kk = 0; kk < KD; kk += 4) {
                double a[2], b[4];
#pragma unroll
                for (int i = 0; i < 2; i++) a[i] = A[(m0 + 8 * i + g) * ld + kk + t];
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = B[(r0 + 8 * j + g) * ld + kk + t];
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        for (int i = 0; i < 2; i++) for (int j = 0; j < 4; j++) s += acc[i][j][0] + acc[i][j][1];
    } else if (V == 2) {
        // 16 warps: 2 x 4 warp tiles of 32 x 32 over a 64 x 128 output, each computed twice (same flops per warp as v1 x 2)
        const int g = lane >> 2, t = lane & 3, m0 = ((warp >> 2) & 1) << 5, r0 = (warp & 3) << 5;
        double acc[4][4][2] = {};
        for (int r = 0; r < reps; r++) {
#pragma unroll 2
            for (int kk = 0; kk < KD; kk += 4) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) a[i] = A[(m0 + 8 * i + g) * ld + kk + t];
#pragma unroll
                for (int j = 0; j < 4; j++) b[j] = B[(r0 + 8 * j + g) * ld + kk + t];
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) s += acc[i][j][0] + acc[i][j][1];
    } else if (V == 3) {   // register-only DFMA
        double acc[16]; for (int i = 0; i < 16; i++) acc[i] = i;
        double a = sm[lane], b = sm[lane + 32];
        for (int r = 0; r < reps; r++)
            for (int kk = 0; kk < KD; kk++)
#pragma unroll
                for (int i = 0; i < 16; i++) acc[i] = fma(a, b, acc[i]);
        for (int i = 0; i < 16; i++) s += acc[i];
    } else {               // register-only DMMA: same FMA count per thread per kk-quad as v1
        double acc[8][2] = {};
        double a = sm[lane], b = sm[lane + 32];
        for (int r = 0; r < reps; r++)
            for (int kk = 0; kk < KD; kk += 4)
#pragma unroll
                for (int i = 0; i < 8; i++) dmma(acc[i][0], acc[i][1], a, b);
        for (int i = 0; i < 8; i++) s += acc[i][0] + acc[i][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int V> void run(const char* name, double flop_per_cta_rep, int ld, int threads = 512) {
    double* out; cudaMalloc(&out, 148 * 512 * 8);
    size_t smem = (size_t)(64 + 128) * ld * 8;
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int reps = 400;
    k<V><<<148, threads, smem>>>(out, 10, ld);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<V><<<148, threads, smem>>>(out, reps, ld);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    printf("%-28s ld %3d  %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, ld, ms, flop_per_cta_rep * reps * 148 / (ms * 1e-3) / 1e12, cudaGetErrorString(err));
    cudaFree(out);
}
int main() {
    const double tile = 2.0 * 64 * 128 * KD;   // flops of one 64x128xKD product
    for (int ld : {130, 132, 136}) {
        run<0>("v0 fma 4x4 LDS.128", tile, ld);
        run<1>("v1 dmma warp 16x32", tile, ld);
        run<2>("v2 dmma warp 32x32 (x2 work)", 2 * tile, ld);
    }
    run<3>("r0 dfma registers", 2.0 * 512 * 16 * KD, 130);
    run<4>("r1 dmma registers", 2.0 * 16 * 8 * 256 * (KD / 4), 130);
    // DMMA issue interval per warp: the same register-only loop with 8 / 4 / 32 warps per SM (flops scale with the warp count)
    run<4>("r1 dmma registers,  8 warps", 2.0 * 8 * 8 * 256 * (KD / 4), 130, 256);
    run<4>("r1 dmma registers,  4 warps", 2.0 * 4 * 8 * 256 * (KD / 4), 130, 128);
    run<4>("r1 dmma registers, 32 warps", 2.0 * 32 * 8 * 256 * (KD / 4), 130, 1024);
    run<3>("r0 dfma registers,  8 warps", 2.0 * 256 * 16 * KD, 130, 256);
    return 0;
}
