// Micro-benchmark: Tensor Memory (TMEM) as a warp-private scratch for fp64 row tiles that feed mma.sync m8n8k4.f64 (DMMA).
// A warp keeps a 16 x 112 fp64 row tile as accumulator-layout fragments in TMEM (tcgen05.st 32x32b), reads it back with
// tcgen05.ld as the A operand (contraction index permuted so that a C fragment IS an A fragment) and multiplies it with a
// 112 x 112 matrix held in shared memory (16 x 16 tiles, row stride 20).  Variants: A from TMEM / from shared memory / registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const double* v) {   // 4 doubles = 8 columns
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, double* v) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5),
                 "=r"(r6), "=r"(r7) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    v[0] = __hiloint2double(r1, r0); v[1] = __hiloint2double(r3, r2); v[2] = __hiloint2double(r5, r4); v[3] = __hiloint2double(r7, r6);
}

#define TLD 20
#define TSZ (16 * TLD)
__device__ __forceinline__ int tile_off(int kb, int cb) { return (cb * (cb + 1) / 2 + kb) * TSZ; }

// mode 0: A from TMEM, 1: A from shared memory (row-major, ld 116), 2: A constant registers
template <int MODE>
__global__ void __launch_bounds__(512) probe(double* out, int reps, int nwarps_active) {
    extern __shared__ __align__(16) double sm[];
    double* Wt = sm;                       // 28 tiles
    double* As = sm + 28 * TSZ;            // 16 warps x 16 rows x 116
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int i = threadIdx.x; i < 28 * TSZ; i += blockDim.x) Wt[i] = 1e-3 * ((i * 7) % 13);
    for (int i = threadIdx.x; i < 16 * 116; i += blockDim.x) As[i] = 1e-3 * ((i * 5) % 11);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tb = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(warp >> 2) * 112;   // lane quarter, 112 columns per warp
    // row tile -> TMEM: fragment (i, j): rows 8i + g, columns 8j + 2t + e ; stored as [j][i][e] -> 4 doubles (8 columns) per 8-column block j
    if (MODE == 0) {
        for (int j = 0; j < 14; j++) {
            double v[4];
            for (int i = 0; i < 2; i++) for (int e = 0; e < 2; e++) v[2 * i + e] = As[(8 * i + g) * 116 + 8 * j + 2 * t + e];
            tmem_st8(tb + 8 * j, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    double sum = 0.0;
    if (warp < nwarps_active) {
        const double* arow[2] = {As + g * 116, As + (8 + g) * 116};
        for (int rep = 0; rep < reps; rep++) {
            // y = p W^T for the 16-column block pairs (bA, bB) = (s, 6 - s): 4 "slots" per row tile, all done by this warp
            for (int s = 0; s < 4; s++) {
                const int bA = s, bB = 6 - s;
                double acc[2][4][2];
                for (int i = 0; i < 2; i++) for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
                for (int kb = 0; kb <= bB; kb++) {
                    const bool doA = bA != bB && kb <= bA;
                    const double* tB = Wt + tile_off(kb, bB);
                    const double* tA = Wt + tile_off(kb, doA ? bA : bB);
#pragma unroll
                    for (int h = 0; h < 2; h++) {   // 8-column block 2 kb + h of the row tile: contraction slots k = 8 (2kb + h) + 2t + e
                        double a[4];
                        if (MODE == 0) tmem_ld8(tb + 8 * (2 * kb + h), a);
                        else if (MODE == 1) { for (int i = 0; i < 2; i++) for (int e = 0; e < 2; e++) a[2 * i + e] = arow[i][16 * kb + 8 * h + 2 * t + e]; }
                        else { a[0] = 1.0 + rep; a[1] = 2.0; a[2] = 3.0; a[3] = 4.0 + s; }
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int kr = 8 * h + 2 * t + e;   // row inside the 16 x 16 tile
                            const double b2 = tB[kr * TLD + g], b3 = tB[kr * TLD + 8 + g];
                            if (doA) {
                                const double b0 = tA[kr * TLD + g], b1 = tA[kr * TLD + 8 + g];
                                for (int i = 0; i < 2; i++) { dmma884(acc[i][0][0], acc[i][0][1], a[2 * i + e], b0); dmma884(acc[i][1][0], acc[i][1][1], a[2 * i + e], b1); }
                            }
                            for (int i = 0; i < 2; i++) { dmma884(acc[i][2][0], acc[i][2][1], a[2 * i + e], b2); dmma884(acc[i][3][0], acc[i][3][1], a[2 * i + e], b3); }
                        }
                    }
                }
                for (int i = 0; i < 2; i++) for (int j = 0; j < 4; j++) sum += acc[i][j][0] + acc[i][j][1];
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

template <int MODE>
static void run(const char* name, int nw, double* out) {
    const int reps = 200;
    const size_t smem = sizeof(double) * (28 * TSZ + 16 * 116);
    cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<MODE><<<148, 512, smem>>>(out, 2, nw);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<148, 512, smem>>>(out, reps, nw);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // DMMAs per row tile: B blocks 16 (bB + 1) + A blocks 16 (bA + 1) summed over the slots = 448
    const double flops = 148.0 * nw * reps * 448.0 * 512.0;
    double h[4]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s warps %2d  %8.3f ms  %6.2f TFLOP/s  (check %.6e)  %s\n", name, nw, ms, flops / ms * 1e-9, h[0], cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 512);
    for (int nw : {4, 8, 12, 16}) {
        run<0>("A from TMEM (tcgen05.ld)", nw, out);
        run<1>("A from shared memory", nw, out);
        run<2>("A in registers", nw, out);
    }
    return 0;
}
