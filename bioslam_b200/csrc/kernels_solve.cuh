// Sparse solve (north_star subsystem 3): block-tridiagonal time chain + arrowhead (static variables) Cholesky.
//
//   (H + lambda I) delta = -g,   H = [ D_0  O_0^T            A_0^T ]
//                                    [ O_0  D_1   O_1^T      A_1^T ]
//                                    [        ...            ...   ]
//                                    [ A_0  A_1   ...        S     ]
//
// Nested partitioned (SPIKE-style) factorisation so one solve exposes P-way parallelism:
//   1. chain_sweep_kernel, one CTA per chunk: eliminate the chunk's interior keyframes left to right.  The coupling to
//      the chunk's LEFT separator keyframe is carried as extra "arrow" rows, exactly like the static variables, and the
//      right-hand side is carried as one more arrow row.
//   2. chain_schur_kernel, one CTA per chunk: the arrow Schur complement  S - sum_k Y_k Y_k^T  of the chunk, computed
//      from the stored arrow panels in one pass (off the sequential path, accumulators in registers, no atomics).
//   3. chain_reduce_kernel: gather the chunks' Schur complements onto the separator keyframes (= nodes of the next level).
//   4. the same three steps on the separators (level 1), then on the group separators by a single CTA (top level),
//      top_finish_kernel: static system + back-substitution of the top-level nodes.
//   5. chain_backsub_prepass_kernel (all nodes in parallel: y_k = Y_k^T coef) and chain_backsub_kernel (one CTA per
//      chunk, sequential over its nodes, matrix-vector products only).
//
// One elimination step (node k, pivot block M = D_k + lambda I - fill, n x n):
//   a. right-looking blocked Cholesky of M fused with the inversion of its factor: the identity is carried along as
//      appended rows, which turns them into W^T = L^-T.  L lives in the lower triangle of the shared-memory block while it
//      is needed, W^T (upper triangle, diagonal included) is what remains; 16 x 16 diagonal blocks are factorised and
//      inverted by one warp in registers (shuffles), panels and trailing updates run on all warps.
//   b. every row p of the panel [O_k ; arrow rows] (processed in chunks of <= 64 rows that fit in shared memory; two
//      independent groups of warps on alternate 32-row chunks where O_k is block-sparse):
//        y = p W^T            (forward substitution, one triangular GEMM against the resident W^T)
//        z = y W  = p M^-1    (second triangular GEMM against the same block)
//        fill row = z O_k^T   (O_k is block-sparse: packed rows, <= 16 terms per entry for the IMU chain)
//      y is stored for the Schur complement / back-substitution; the fill rows are the corrections of the next node.
// All GEMMs run on the fp64 tensor pipe (DMMA, mma.sync m8n8k4; warp tile 16 x 32, operands read from shared memory as
// scalars, see "micro-kernels" below and DESIGN.md section 3 for why this departs from the north_star's "no tensor cores").
#pragma once
#include "kernels_assemble.cuh"

namespace b200d {

#define B200_NB 16        // Cholesky block width
#define B200_SOLVE_THREADS 512
#define B200_SCHUR_THREADS 384
#define B200_PREPASS_THREADS 256

// optional per-phase cycle counters of the sweep (debug builds: -DB200_PHASE_CLOCKS), accumulated by thread 0 of the last block
#ifdef B200_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[24];
__device__ int g_phase_on;   // set by the level-0 (block-sparse) sweep, cleared by the dense ones: only level 0 is profiled
#define PH_T0() unsigned long long ph_t_ = clock64()
#define PH_ADD(i) do { if (threadIdx.x == 0 && blockIdx.x == 1 && g_phase_on) { unsigned long long n_ = clock64(); g_phase_clk[i] += n_ - ph_t_; ph_t_ = n_; } } while (0)
// the same from an arbitrary thread t of the last block (worker-warp timelines of the warp-specialised Cholesky)
#define PHW_T0() unsigned long long phw_t_ = clock64()
#define PHW_ADD(i, t) do { if (threadIdx.x == (t) && blockIdx.x == 1 && g_phase_on) { unsigned long long n_ = clock64(); g_phase_clk[i] += n_ - phw_t_; phw_t_ = n_; } } while (0)
#else
#define PH_T0()
#define PH_ADD(i)
#define PHW_T0()
#define PHW_ADD(i, t)
#endif

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void st2(double* p, double x, double y) { *reinterpret_cast<double2*>(p) = make_double2(x, y); }

// ------------------------------------------------------------------------------------------------- micro-kernels
// Warp tile 16 x 32 on the fp64 tensor cores (DMMA, mma.sync m8n8k4 -- see dmma884 in cuda_compat.h): per contraction step
// of 4 a warp issues 6 scalar operand loads and 8 MMAs (2048 FMAs); the r01 scalar micro-kernel needed 16 LDS.128 for the same
// work and was shared-memory bound at ~55 % of the fp64 rate (tools/ubench/dmma_probe.cu: 22 vs 37 TFLOP/s).
// lane -> (g = lane >> 2, t = lane & 3).  Accumulator element acc[i][j][e]: row m0 + 8 i + g, column (base of 8-column
// block j) + 2 t + e.  Contraction ranges are arbitrary: operands at or beyond the end of a range are read as zero.
typedef double Acc[2][4][2];
__device__ __forceinline__ void acc_zero(Acc& acc) {
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
}

// NN:  C += A(16 x k) B(k x 32), B row-major indexed [k][c]; 8-column blocks at c0 + 8 j.
__device__ __forceinline__ void warp_mma_nn(const double* __restrict__ A, int lda, int m0, int M, const double* __restrict__ B, int ldb,
                                            int c0, int N, int ka, int kb, Acc& acc) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double* ap[2];
    const double* bp[4];
#pragma unroll
    for (int i = 0; i < 2; i++) { int m = m0 + 8 * i + g; if (m > M - 1) m = M - 1; ap[i] = A + (long long)m * lda + t; }
#pragma unroll
    for (int j = 0; j < 4; j++) { int c = c0 + 8 * j + g; if (c > N - 1) c = N - 1; bp[j] = B + (long long)t * ldb + c; }
    for (int k = ka; k < kb; k += 4) {
        const bool ok = k + t < kb;
        double a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; i++) a[i] = ok ? ap[i][k] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = ok ? bp[j][(long long)k * ldb] : 0.0;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// NT:  C += A(16 x k) B(32 x k)^T, both row-major along k; 8-column blocks = rows r0 + 8 j of B.
__device__ __forceinline__ void warp_mma_nt(const double* __restrict__ A, int lda, int m0, int M, const double* __restrict__ B, int ldb,
                                            int r0, int R, int ka, int kb, Acc& acc) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double* ap[2];
    const double* bp[4];
#pragma unroll
    for (int i = 0; i < 2; i++) { int m = m0 + 8 * i + g; if (m > M - 1) m = M - 1; ap[i] = A + (long long)m * lda + t; }
#pragma unroll
    for (int j = 0; j < 4; j++) { int r = r0 + 8 * j + g; if (r > R - 1) r = R - 1; bp[j] = B + (long long)r * ldb + t; }
    for (int k = ka; k < kb; k += 4) {
        const bool ok = k + t < kb;
        double a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; i++) a[i] = ok ? ap[i][k] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = ok ? bp[j][k] : 0.0;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// Balanced variants for the TRIANGULAR products: a warp owns two 16-column blocks A (acc[.][0..1]) and B (acc[.][2..3]) whose
// contraction ranges are complementary (short + long), so that every warp of a row group carries the same work.
// NN: block A = columns [cA, cA+16) over k in [ka, kendA), block B = columns [cB, cB+16) over k in [ka, kendB); kendA <= kendB.
__device__ __forceinline__ void warp_mma_nn2(const double* __restrict__ A, int lda, int m0, int M, const double* __restrict__ B, int ldb,
                                             int cA, int cB, int N, int ka, int kendA, int kendB, Acc& acc) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double* ap[2];
    const double* bp[4];
#pragma unroll
    for (int i = 0; i < 2; i++) { int m = m0 + 8 * i + g; if (m > M - 1) m = M - 1; ap[i] = A + (long long)m * lda + t; }
#pragma unroll
    for (int j = 0; j < 4; j++) { int c = ((j < 2) ? cA : cB) + 8 * (j & 1) + g; if (c > N - 1) c = N - 1; bp[j] = B + (long long)t * ldb + c; }
    int k = ka;
    for (; k < kendA; k += 4) {
        const bool ok = k + t < kendA;
        double a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; i++) a[i] = ok ? ap[i][k] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = ok ? bp[j][(long long)k * ldb] : 0.0;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    for (k = (kendA > ka ? kendA : ka); k < kendB; k += 4) {
        const bool ok = k + t < kendB;
        double a[2], b[2];
#pragma unroll
        for (int i = 0; i < 2; i++) a[i] = ok ? ap[i][k] : 0.0;
#pragma unroll
        for (int j = 0; j < 2; j++) b[j] = ok ? bp[2 + j][(long long)k * ldb] : 0.0;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 2; j++) dmma884(acc[i][2 + j][0], acc[i][2 + j][1], a[i], b[j]);
    }
}
// NT: block A = rows [kA, kA+16) of B over c in [sA, kend), block B = rows [kB, kB+16) of B over c in [sB, kend).
__device__ __forceinline__ void warp_mma_nt2(const double* __restrict__ A, int lda, int m0, int M, const double* __restrict__ B, int ldb,
                                             int kA, int kB, int R, int sA, int sB, int kend, Acc& acc) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double* ap[2];
    const double* bp[4];
#pragma unroll
    for (int i = 0; i < 2; i++) { int m = m0 + 8 * i + g; if (m > M - 1) m = M - 1; ap[i] = A + (long long)m * lda + t; }
#pragma unroll
    for (int j = 0; j < 4; j++) { int r = ((j < 2) ? kA : kB) + 8 * (j & 1) + g; if (r > R - 1) r = R - 1; bp[j] = B + (long long)r * ldb + t; }
    // the block that starts earlier runs alone over [lo, hi)  (static register indices: two copies of the loop)
    const int lo = sA < sB ? sA : sB, hi = sA < sB ? sB : sA;
    const int hic = hi < kend ? hi : kend;
    if (sA < sB) {
        for (int k = lo; k < hic; k += 4) {
            const bool ok = k + t < hic;
            double a[2], b[2];
#pragma unroll
            for (int i = 0; i < 2; i++) a[i] = ok ? ap[i][k] : 0.0;
#pragma unroll
            for (int j = 0; j < 2; j++) b[j] = ok ? bp[j][k] : 0.0;
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    } else {
        for (int k = lo; k < hic; k += 4) {
            const bool ok = k + t < hic;
            double a[2], b[2];
#pragma unroll
            for (int i = 0; i < 2; i++) a[i] = ok ? ap[i][k] : 0.0;
#pragma unroll
            for (int j = 0; j < 2; j++) b[j] = ok ? bp[2 + j][k] : 0.0;
#pragma unroll
            for (int i = 0; i < 2; i++)
#pragma unroll
                for (int j = 0; j < 2; j++) dmma884(acc[i][2 + j][0], acc[i][2 + j][1], a[i], b[j]);
        }
    }
    for (int k = hi; k < kend; k += 4) {
        const bool ok = k + t < kend;
        double a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; i++) a[i] = ok ? ap[i][k] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = ok ? bp[j][k] : 0.0;
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// NT, warp tile 32 x 32 (Schur complement kernel): C += A[m0.., :] A[r0.., :]^T; acc[i][j][e]: row m0 + 8 i + g, column r0 + 8 j + 2 t + e.
typedef double Acc32[4][4][2];
__device__ __forceinline__ void warp_mma_nt32(const double* __restrict__ A, int lda, int m0, int M, int r0, int ka, int kb, Acc32& acc) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double* ap[4];
    const double* bp[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { int m = m0 + 8 * i + g; if (m > M - 1) m = M - 1; ap[i] = A + (long long)m * lda + t; }
#pragma unroll
    for (int j = 0; j < 4; j++) { int r = r0 + 8 * j + g; if (r > M - 1) r = M - 1; bp[j] = A + (long long)r * lda + t; }
    for (int k = ka; k < kb; k += 4) {
        const bool ok = k + t < kb;
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; i++) a[i] = ok ? ap[i][k] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = ok ? bp[j][k] : 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// legacy 32 x 16 NT tile loop over the whole CTA (used by the small static system only)
template <class Out>
__device__ __forceinline__ void cta_gemm_nt(const double* __restrict__ A, int lda, int M, const double* __restrict__ B, int ldb, int N,
                                            int k0, int k1, Out out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int tx = lane & 3, ty = lane >> 2;
    const int mt = (M + 31) >> 5, ntl = (N + 15) >> 4;
    for (int tile = warp; tile < mt * ntl; tile += nwarps) {
        const int m0 = (tile / ntl) << 5, n0 = (tile % ntl) << 4;
        const double* ap[4];
        const double* bp[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            int m = m0 + ty + 8 * i; if (m > M - 1) m = M - 1;
            int n = n0 + tx + 4 * i; if (n > N - 1) n = N - 1;
            ap[i] = A + (long long)m * lda;
            bp[i] = B + (long long)n * ldb;
        }
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = 0.0;
        for (int k = k0; k < k1; k += 2) {
            double2 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; i++) { a[i] = ld2(ap[i] + k); b[i] = ld2(bp[i] + k); }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fma(a[i].y, b[j].y, fma(a[i].x, b[j].x, acc[i][j]));
        }
        __syncwarp();  // in-place use (C aliases A within the warp's own rows): all lanes have finished reading
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int m = m0 + ty + 8 * i;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int n = n0 + tx + 4 * j;
                if (m < M && n < N) out(m, n, acc[i][j]);
            }
        }
    }
}

// The same with the structurally dead 8 x 8 blocks of a Schur tile left out: only the first NI row blocks exist (last tile
// row of an h that is not a multiple of 32) and, on a diagonal tile, only the blocks on or below the block diagonal.
template <int NI, bool DIAG>
__device__ __forceinline__ void warp_mma_nt32_t(const double* __restrict__ A, int lda, int m0, int M, int r0, int ka, int kb, Acc32& acc) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    constexpr int NJ = DIAG ? NI : 4;
    const double* ap[NI];
    const double* bp[NJ];
#pragma unroll
    for (int i = 0; i < NI; i++) { int m = m0 + 8 * i + g; if (m > M - 1) m = M - 1; ap[i] = A + (long long)m * lda + t; }
#pragma unroll
    for (int j = 0; j < NJ; j++) { int r = r0 + 8 * j + g; if (r > M - 1) r = M - 1; bp[j] = A + (long long)r * lda + t; }
    for (int k = ka; k < kb; k += 4) {
        const bool ok = k + t < kb;
        double a[NI], b[NJ];
#pragma unroll
        for (int i = 0; i < NI; i++) a[i] = ok ? ap[i][k] : 0.0;
#pragma unroll
        for (int j = 0; j < NJ; j++) b[j] = ok ? bp[j][k] : 0.0;
#pragma unroll
        for (int i = 0; i < NI; i++)
#pragma unroll
            for (int j = 0; j < (DIAG ? i + 1 : 4); j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}
__device__ __forceinline__ void warp_mma_nt32_kind(int ni, bool diag, const double* __restrict__ A, int lda, int m0, int M, int r0, int ka, int kb,
                                                   Acc32& acc) {
    switch (ni * 2 + (diag ? 1 : 0)) {
        case 2: warp_mma_nt32_t<1, false>(A, lda, m0, M, r0, ka, kb, acc); break;
        case 3: warp_mma_nt32_t<1, true>(A, lda, m0, M, r0, ka, kb, acc); break;
        case 4: warp_mma_nt32_t<2, false>(A, lda, m0, M, r0, ka, kb, acc); break;
        case 5: warp_mma_nt32_t<2, true>(A, lda, m0, M, r0, ka, kb, acc); break;
        case 6: warp_mma_nt32_t<3, false>(A, lda, m0, M, r0, ka, kb, acc); break;
        case 7: warp_mma_nt32_t<3, true>(A, lda, m0, M, r0, ka, kb, acc); break;
        case 8: warp_mma_nt32_t<4, false>(A, lda, m0, M, r0, ka, kb, acc); break;
        default: warp_mma_nt32_t<4, true>(A, lda, m0, M, r0, ka, kb, acc); break;
    }
}

// ------------------------------------------------------------------------------------------------- 16 x 16 building blocks
#define B200_WLD 20  // leading dimension of the 16 x 16 inverse diagonal block in shared memory (== 4 mod 8: conflict-free DMMA operand loads)

// Diagonal block by ONE warp, entirely in registers: lane l < 16 owns row l of the jb x jb block (jb <= 16, padded with
// identity); lane 16 + r carries row r of an appended identity, which the same column operations turn into row r of
// W^T = L^-T (the inverse of the factor comes for free, no separate substitution).  Pivots / column entries travel by warp
// shuffle.  Wd[i * WLD + l] = W[i][l] (row-major, lower).  If inv_in_place: the block itself becomes W^T (upper triangle
// incl. diagonal, strictly lower part zeroed); else L is written back in place (lower) with the strictly lower part of W
// stashed above the diagonal, and Tdiag / invd receive L^T / the reciprocal diagonal (legacy static solver).
// Returns 0 or 1 + index of a non-positive pivot (uniform over the warp).
__device__ __forceinline__ int warp_potrf16(double* Db, int ld, int jb, double* Tdiag, double* invd, double* Wd, bool inv_in_place) {
    const int l = threadIdx.x & 31;
    const int rr = l & 15;
    double r[B200_NB];
    if (l < B200_NB) {
#pragma unroll
        for (int c = 0; c < B200_NB; c++) r[c] = (l < jb && c < jb && c <= l) ? Db[(long long)l * ld + c] : ((l == c) ? 1.0 : 0.0);
    } else {
#pragma unroll
        for (int c = 0; c < B200_NB; c++) r[c] = (c == rr) ? 1.0 : 0.0;
    }
    int fail = 0;
#pragma unroll
    for (int c = 0; c < B200_NB; c++) {
        double d = __shfl_sync(0xffffffffu, r[c], c);
        if (!(d > 0.0) || !(d < 1.0e300)) { if (!fail) fail = 1 + c; d = 1.0; }
        const double inv = rsqrt(d);
        const double x = r[c] * inv;   // lane c: sqrt(d) ; lanes c < l < 16: L[l][c] ; lane 16 + r: W[c][r]
        r[c] = x;
        if (!inv_in_place && l == c) invd[c] = inv;
#pragma unroll
        for (int c2 = c + 1; c2 < B200_NB; c2++) {
            const double lc2 = __shfl_sync(0xffffffffu, x, c2);
            r[c2] = fma(-x, lc2, r[c2]);
        }
    }
    if (l < B200_NB) {
        if (!inv_in_place) {
#pragma unroll
            for (int c = 0; c < B200_NB; c++) {
                if (c <= l) {
                    if (l < jb && c < jb) Db[(long long)l * ld + c] = r[c];
                    Tdiag[c * B200_NB + l] = r[c];
                }
            }
        }
    } else {
        // lane 16 + rr holds row rr of W^T: r[c] = W[c][rr] (zero for c < rr)
#pragma unroll
        for (int c = 0; c < B200_NB; c++) {
            Wd[c * B200_WLD + rr] = r[c];
            if (inv_in_place) {
                if (rr < jb && c < jb) Db[(long long)rr * ld + c] = r[c];
            } else {
                if (c > rr && c < jb) Db[(long long)rr * ld + c] = r[c];
            }
        }
    }
    return fail;
}

// ---- the same diagonal block WITHOUT shuffles on the pivot chain: the block is split 8 + 8 and each 8 x 8 half is factorised and
// inverted REDUNDANTLY IN EVERY LANE, entirely in registers (36 + 36 doubles, statically indexed): per pivot the chain is
// rsqrt -> scale -> update, no warp shuffle and no shared-memory round trip (the shuffle version measured ~590 cycles per pivot
// next to 15 warps that saturate the MIO pipe with operand loads; r2 profile).  The three 8 x 8 products that join the halves
//   L21 = A21 W11^T,  A22 -= L21 L21^T,  W21 = -W22 L21 W11
// run on the tensor pipe.  An accumulator fragment (row g, columns 2t, 2t+1) is fed back as an A or B operand by permuting the
// contraction index (slot t of step e <-> k = 2t + e), so L21 and T never leave the registers.
#define B200_TRI(i, j) ((i) * ((i) + 1) / 2 + (j))
// a: packed lower triangle of an SPD 8 x 8 block, identical in all lanes.  Returns W = L^-1 (packed lower) in w; a holds L.
__device__ __forceinline__ int potrf8_regs(double (&a)[36], double (&w)[36]) {
    int fail = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        double d = a[B200_TRI(c, c)];
        if (!(d > 0.0) || !(d < 1.0e300)) { if (!fail) fail = 1 + c; d = 1.0; }
        const double inv = rsqrt(d);
        // row c of W: W[c][j] = -inv * sum_{k = j}^{c-1} L[c][k] W[k][j]   (the sums do not wait for the reciprocal root)
        w[B200_TRI(c, c)] = inv;
#pragma unroll
        for (int j = 0; j < c; j++) {
            double sacc = 0.0;
#pragma unroll
            for (int k = j; k < c; k++) sacc = fma(a[B200_TRI(c, k)], w[B200_TRI(k, j)], sacc);
            w[B200_TRI(c, j)] = -inv * sacc;
        }
#pragma unroll
        for (int i = c + 1; i < 8; i++) a[B200_TRI(i, c)] *= inv;
#pragma unroll
        for (int j = c + 1; j < 8; j++)
#pragma unroll
            for (int i = j; i < 8; i++) a[B200_TRI(i, j)] = fma(-a[B200_TRI(i, c)], a[B200_TRI(j, c)], a[B200_TRI(i, j)]);
    }
    return fail;
}
// row (lane & 7) of the packed lower-triangular w as 8 doubles (zeros right of the diagonal): a select tree instead of
// dynamically indexed registers
__device__ __forceinline__ void tri_row_select(const double (&w)[36], int row, double (&out)[8]) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        double v = 0.0;
#pragma unroll
        for (int i = j; i < 8; i++) v = (row == i) ? w[B200_TRI(i, j)] : v;
        out[j] = v;
    }
}
// In: lower triangle of the jb x jb diagonal block at Db (row-major, ld).  Out: the block holds W^T = L^-T (upper triangle incl.
// the diagonal, strictly lower part zero), Wd[i * WLD + l] = W[i][l] (16 x 16, identity padding beyond jb).  One warp.
__device__ __forceinline__ int warp_potrf16_r8(double* Db, int ld, int jb, double* Wd) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double a[36], w[36];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) a[B200_TRI(i, j)] = (i < jb) ? Db[(long long)i * ld + j] : ((i == j) ? 1.0 : 0.0);
    // A21 as an A operand (row 8 + g, columns 4 s + t) and A22 as an accumulator fragment (row 8 + g, columns 8 + 2t + e)
    double a21[2], a22[2];
    {
        const int r = 8 + g;
#pragma unroll
        for (int s = 0; s < 2; s++) a21[s] = (r < jb) ? Db[(long long)r * ld + 4 * s + t] : 0.0;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int c = 8 + 2 * t + e;
            a22[e] = (r < jb && c <= r) ? Db[(long long)r * ld + c] : ((r == c) ? 1.0 : 0.0);
        }
    }
    int fail = potrf8_regs(a, w);
    {   // rows 0..7 of W (columns 8..15 are zero)
        double row[8];
        tri_row_select(w, lane & 7, row);
        if (lane < 8) {
            double* wr = Wd + lane * B200_WLD;
#pragma unroll
            for (int j = 0; j < 8; j += 2) { st2(wr + j, row[j], row[j + 1]); st2(wr + 8 + j, 0.0, 0.0); }
        }
    }
    __syncwarp();
    // L21 = A21 W11^T (natural contraction order) -> fragment l21[e] = L21[g][2t + e]
    double l21[2] = {0.0, 0.0};
#pragma unroll
    for (int s = 0; s < 2; s++) dmma884(l21[0], l21[1], a21[s], Wd[g * B200_WLD + 4 * s + t]);
    // A22 - L21 L21^T: the fragment is both operands (slot t of step e <-> k = 2t + e)
    double u[2] = {0.0, 0.0};
#pragma unroll
    for (int e = 0; e < 2; e++) dmma884(u[0], u[1], l21[e], l21[e]);
    st2(Wd + (8 + g) * B200_WLD + 8 + 2 * t, a22[0] - u[0], a22[1] - u[1]);   // scratch: the W22 area of Wd
    // T^T = W11^T L21^T (needed after the second half; issued here, off the chain): tT[e] = T^T[g][2t + e]
    double tT[2] = {0.0, 0.0};
#pragma unroll
    for (int e = 0; e < 2; e++) dmma884(tT[0], tT[1], Wd[(2 * t + e) * B200_WLD + g], l21[e]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) a[B200_TRI(i, j)] = Wd[(8 + i) * B200_WLD + 8 + j];
    __syncwarp();   // every lane has read the scratch before rows 8..15 of W overwrite it
    {
        const int f2 = potrf8_regs(a, w);
        if (!fail && f2) fail = 8 + f2;
    }
    {   // W22 into rows 8..15, columns 8..15
        double row[8];
        tri_row_select(w, lane & 7, row);
        if (lane < 8) {
            double* wr = Wd + (8 + lane) * B200_WLD + 8;
#pragma unroll
            for (int j = 0; j < 8; j += 2) st2(wr + j, row[j], row[j + 1]);
        }
    }
    __syncwarp();
    // W21^T = -T^T W22^T: x[e] = (T^T W22^T)[g][2t + e] = -W[8 + 2t + e][g]
    double x[2] = {0.0, 0.0};
#pragma unroll
    for (int e = 0; e < 2; e++) dmma884(x[0], x[1], tT[e], Wd[(8 + g) * B200_WLD + 8 + 2 * t + e]);
    Wd[(8 + 2 * t) * B200_WLD + g] = -x[0];
    Wd[(8 + 2 * t + 1) * B200_WLD + g] = -x[1];
    __syncwarp();
    // the block itself <- W^T (upper triangle incl. the diagonal; zero below), rows / columns < jb only
    {
        const int r = lane >> 1, c0 = (lane & 1) << 3;
        if (r < jb) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
                const int c = c0 + j;
                if (c < jb) {   // (jb is even or the last column is written alone)
                    const double v0 = (c >= r) ? Wd[c * B200_WLD + r] : 0.0;
                    if (c + 1 < jb) { const double v1 = (c + 1 >= r) ? Wd[(c + 1) * B200_WLD + r] : 0.0; st2(Db + (long long)r * ld + c, v0, v1); }
                    else Db[(long long)r * ld + c] = v0;
                }
            }
        }
    }
    return fail;
}

// small static shared scratch of the diagonal-block routines (one instance per kernel)
__device__ __forceinline__ void solve_scratch(double*& Tdiag, double*& Wd, double*& invd) {
    __shared__ __align__(16) double s_Tdiag[B200_NB * B200_NB];
    __shared__ __align__(16) double s_Wd[B200_NB * B200_WLD];
    __shared__ double s_invd[B200_NB];
    Tdiag = s_Tdiag; Wd = s_Wd; invd = s_invd;
}

// ------------------------------------------------------------------------------------------------- fused Cholesky + inverse
// In: the lower triangle (diagonal included) of the SPD matrix M in Ms (n x n, row-major, ld), strictly upper part ZERO.
// Out: the upper triangle (diagonal included) holds W^T = L^-T (Ms[k][c] = W[c][k], c >= k); below the diagonal blocks what is
// left of L remains (the caller clears it).
// Right-looking, block width 16.  Step j: (a) diagonal block -> W_jj^T (one warp); (b) column block j of every other
// row times W_jj^T (rows below: the L panel; rows above: the appended identity rows); (c) trailing update with the L panel
// of the rows that are alive: r >= c (Cholesky) or r < j0 + 16 (appended rows already started).
// Returns (to all threads) 0 or 1 + index of the first non-positive pivot.  Must be called by all threads of the CTA.
// Warp-specialised: warp 0 (the "pivot" warp) runs the sequential chain of diagonal blocks -- for block s+1 it scales the
// 16 x 16 panel block (s+1, s), applies it to the diagonal tile and factorises/inverts it -- while the other warps scale the
// rest of column block s and apply the rank-16 trailing update.  The two roles meet at named barriers (ids 1..6, two
// alternating instances each): W_READY (W_ss in Wd[s&1]), ROW_READY (block row s+1 scaled), T_DONE (trailing update s done).
// Runs on a GROUP of gnw consecutive warps (warp index in the group gw; default: the whole CTA); gsync = named barrier id the group
// may use for its own full-group synchronisation (0: __syncthreads, i.e. the group is the CTA).
// lower_only: plain blocked Cholesky -- the appended identity rows (the upper triangle) are left alone: on return the strictly lower
// blocks hold L, the diagonal 16 x 16 tiles hold W_ss^T = L_ss^-T (upper incl. the diagonal, zero below) and nothing else is touched.
// Half the trailing-update work of the fused version; the caller builds W^T from L by forward substitution (kernels_sweep_pipe.cuh).
__device__ int cta_chol_inv(double* Ms, int ld, int n, int gw = -1, int gnw = 0, int gsync = 0, bool lower_only = false) {
    __shared__ int fail;
    __shared__ __align__(16) double s_Wd2[2][B200_NB * B200_WLD];
    double *Tdiag, *Wd0, *invd;
    solve_scratch(Tdiag, Wd0, invd);
    const int lane = threadIdx.x & 31;
    const int warp = gw >= 0 ? gw : (int)(threadIdx.x >> 5), nwarps = gw >= 0 ? gnw : (int)(blockDim.x >> 5);
    const int mg = lane >> 2, mt = lane & 3;   // MMA fragment coordinates
    const int nb = (n + B200_NB - 1) / B200_NB;
#define CHOL_GSYNC() do { if (gsync == 0) __syncthreads(); else named_bar_sync(gsync, nwarps << 5); } while (0)
    // the pivot warp's chain is latency-bound: the warps that share its scheduler (warp & 3 == 0) stay out of the way
    const bool quiet = false;   // (measured: idling the 3 warps on the pivot scheduler does not pay -- the chain waits on the shared MIO pipe)
    const int nwk = quiet ? nwarps - (nwarps + 3) / 4 : nwarps - 1;        // worker warps
    const int nthr = 32 * (nwk + 1);                                       // threads taking part in the named barriers
    if (warp == 0 && lane == 0) fail = 0;
    CHOL_GSYNC();
    PH_T0();
    if (warp == 0) {
        // ------------------------------------------------------------------ pivot warp
        {
            const int f = warp_potrf16_r8(Ms, ld, n < B200_NB ? n : B200_NB, s_Wd2[0]);
            if (f && lane == 0) fail = f;
        }
        __syncwarp();   // W_00 (written by lanes 16..31) is read by the whole warp below
        __threadfence_block();
        named_bar_arrive(1, nthr);
        PHW_T0();
        PHW_ADD(8, 0);
        for (int s = 0; s + 1 < nb; s++) {
            const int j0 = s * B200_NB, cb0 = j0 + B200_NB;
            const int jb1 = (n - cb0 < B200_NB) ? (n - cb0) : B200_NB;
            const double* Wd = s_Wd2[s & 1];
            if (s >= 1) named_bar_sync(5 + ((s - 1) & 1), nthr);     // trailing update s-1 has reached block row s+1
            PHW_ADD(11, 0);
            // block (s+1, s) <- block * W_ss^T  (16 x 16 x 16 on the tensor cores: 2 x 2 MMA blocks, 4 contraction steps)
            const double* xa[2];
            bool rok[2];
#pragma unroll
            for (int i = 0; i < 2; i++) {
                int r = cb0 + 8 * i + mg; rok[i] = r < n; if (r > n - 1) r = n - 1;
                xa[i] = Ms + (long long)r * ld + j0;
            }
            {
                double sacc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
                for (int k = 0; k < B200_NB; k += 4) {
                    double av[2], bv[2];
#pragma unroll
                    for (int i = 0; i < 2; i++) { av[i] = xa[i][k + mt]; bv[i] = Wd[(8 * i + mg) * B200_WLD + k + mt]; }
#pragma unroll
                    for (int i = 0; i < 2; i++)
#pragma unroll
                        for (int j = 0; j < 2; j++) dmma884(sacc[i][j][0], sacc[i][j][1], av[i], bv[j]);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; i++)
#pragma unroll
                    for (int j = 0; j < 2; j++)
                        if (rok[i]) st2(const_cast<double*>(xa[i]) + 8 * j + 2 * mt, sacc[i][j][0], sacc[i][j][1]);
            }
            __syncwarp();
            // diagonal tile (s+1, s+1), lower part: -= L L^T with the block just scaled
            {
                const double* xb[2];
#pragma unroll
                for (int j = 0; j < 2; j++) { int c = cb0 + 8 * j + mg; if (c > n - 1) c = n - 1; xb[j] = Ms + (long long)c * ld + j0; }
                double dacc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
                for (int k = 0; k < B200_NB; k += 4) {
                    double av[2], bv[2];
#pragma unroll
                    for (int i = 0; i < 2; i++) { av[i] = xa[i][k + mt]; bv[i] = xb[i][k + mt]; }
                    dmma884(dacc[0][0][0], dacc[0][0][1], av[0], bv[0]);
                    dmma884(dacc[1][0][0], dacc[1][0][1], av[1], bv[0]);
                    dmma884(dacc[1][1][0], dacc[1][1][1], av[1], bv[1]);
                }
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int r = cb0 + 8 * i + mg;
#pragma unroll
                    for (int j = 0; j <= i; j++) {
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int c = cb0 + 8 * j + 2 * mt + e;
                            if (r < n && c <= r) Ms[(long long)r * ld + c] -= dacc[i][j][e];
                        }
                    }
                }
            }
#ifdef B200_EMU_DBG
            if (lane == 0) printf("pivot s=%d: L[%d][%d]=%g  diag[%d]=%g\n", s, cb0, j0, Ms[(long long)cb0 * ld + j0], cb0, Ms[(long long)cb0 * ld + cb0]);
#endif
            __threadfence_block();
            named_bar_arrive(3 + (s & 1), nthr);                      // block row s+1 of column block s is scaled
            __syncwarp();
            PHW_ADD(9, 0);
            {
                const int f = warp_potrf16_r8(Ms + (long long)cb0 * ld + cb0, ld, jb1, s_Wd2[(s + 1) & 1]);
                if (f && lane == 0 && fail == 0) fail = cb0 + f;
            }
            __syncwarp();
            __threadfence_block();
            named_bar_arrive(1 + ((s + 1) & 1), nthr);                // W_{s+1} ready
            PHW_ADD(8, 0);
            if (fail) break;                                          // (uniform: set by lane 0 before the fence)
        }
        // consume the last trailing-update barrier so that every named barrier is back to zero arrivals for the next call
        if (nb >= 2 && !fail) named_bar_sync(5 + ((nb - 2) & 1), nthr);
    } else if (!quiet || (warp & 3) != 0) {
        // ------------------------------------------------------------------ worker warps
        const int wk = quiet ? warp - 1 - (warp >> 2) : warp - 1;
        PHW_T0();
        for (int s = 0; s < nb; s++) {
            const int j0 = s * B200_NB, cb0 = j0 + B200_NB;
            const int jb = (n - j0 < B200_NB) ? (n - j0) : B200_NB;
            const double* Wd = s_Wd2[s & 1];
            named_bar_sync(1 + (s & 1), nthr);                        // W_ss ready (and every worker has finished update s-1)
            PHW_ADD(12, 32);
            if (fail) break;
            // rows r in [0, j0) U [cb0 + 16, n): X[r][j0 .. j0+jb) <- X * W_ss^T, 16 (virtual) rows per warp, in place
            const int skip = (cb0 < n) ? ((n - cb0 < B200_NB) ? (n - cb0) : B200_NB) : 0;   // block row s+1 belongs to the pivot warp
            const int vabove = lower_only ? 0 : j0;   // (lower_only: no appended rows above the block)
            const int nvr = n - jb - skip - (j0 - vabove);
            for (int t = wk; t * 16 < nvr; t += nwk) {
                const double* ap[2];
                bool ok[2];
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    int vr = t * 16 + 8 * i + mg;
                    ok[i] = vr < nvr;
                    if (!ok[i]) vr = nvr - 1;
                    const int r = (vr < vabove) ? vr : vr - vabove + j0 + jb + skip;
                    ap[i] = Ms + (long long)r * ld + j0;
                }
                double sacc[2][2][2] = {{{0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}}};
#pragma unroll
                for (int k = 0; k < B200_NB; k += 4) {
                    const bool kin = j0 + k + mt < n;
                    double av[2], bv[2];
#pragma unroll
                    for (int i = 0; i < 2; i++) { av[i] = kin ? ap[i][k + mt] : 0.0; bv[i] = kin ? Wd[(8 * i + mg) * B200_WLD + k + mt] : 0.0; }
#pragma unroll
                    for (int i = 0; i < 2; i++)
#pragma unroll
                        for (int j = 0; j < 2; j++) dmma884(sacc[i][j][0], sacc[i][j][1], av[i], bv[j]);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    if (!ok[i]) continue;
                    double* xw = const_cast<double*>(ap[i]);
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        const int c = 8 * j + 2 * mt;
                        if (c < jb) st2(xw + c, sacc[i][j][0], sacc[i][j][1]);   // jb is even: a pair is inside or outside
                    }
                }
            }
            if (cb0 >= n) break;
            __threadfence_block();
            PHW_ADD(13, 32);
            named_bar_sync(3 + (s & 1), nthr);                        // whole column block s scaled (pivot: block row s+1)
            PHW_ADD(14, 32);
            // trailing update: Ms[r][c] -= sum_t Ms[r][j0+t] Ms[c][j0+t], c >= cb0, alive rows (r >= c or r < cb0),
            // except the diagonal tile (s+1, s+1), which the pivot warp has already updated
            if (lower_only) {
                // live tiles only: 16-row tiles below block row s+1, 32-column tiles up to the diagonal
                int idx = 0;
                for (int m0 = cb0 + B200_NB; m0 < n; m0 += 16) {
                    for (int c0 = cb0; c0 <= m0; c0 += 32) {
                        if ((idx++) % nwk != wk) continue;
                        Acc acc;
                        acc_zero(acc);
                        warp_mma_nt(Ms, ld, m0, n, Ms, ld, c0, n, j0, cb0, acc);
#pragma unroll
                        for (int i = 0; i < 2; i++) {
                            const int r = m0 + 8 * i + mg;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int c = c0 + 8 * j + 2 * mt;
                                if (r < n && c <= r) {   // (c even, n even: the pair is inside the matrix; c == r rewrites (r, r+1) unchanged)
                                    double* q = Ms + (long long)r * ld + c;
                                    double2 v = ld2(q);
                                    v.x -= acc[i][j][0];
                                    if (c + 1 <= r) v.y -= acc[i][j][1];
                                    st2(q, v.x, v.y);
                                }
                            }
                        }
                    }
                }
            } else {
                const int ntr = (n + 15) >> 4, ntc = (n - cb0 + 31) >> 5;
                for (int tile = wk; tile < ntr * ntc; tile += nwk) {
                    const int m0 = (tile / ntc) << 4, c0 = cb0 + ((tile % ntc) << 5);
                    if (m0 >= cb0 && m0 + 15 < c0) continue;   // dead zone: not-yet-started appended rows
                    Acc acc;
                    acc_zero(acc);
                    warp_mma_nt(Ms, ld, m0, n, Ms, ld, c0, n, j0, cb0, acc);
#pragma unroll
                    for (int i = 0; i < 2; i++) {
                        const int r = m0 + 8 * i + mg;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
#pragma unroll
                            for (int e = 0; e < 2; e++) {
                                const int c = c0 + 8 * j + 2 * mt + e;
                                const bool pivots = (r >= cb0 && r < cb0 + B200_NB && c < cb0 + B200_NB);
                                if (r < n && c < n && (r >= c || r < cb0) && !pivots) Ms[(long long)r * ld + c] -= acc[i][j][e];
                            }
                        }
                    }
                }
            }
            __threadfence_block();
            named_bar_arrive(5 + (s & 1), nthr);                      // trailing update s done
            PHW_ADD(15, 32);
        }
    }
    CHOL_GSYNC();
    PH_ADD(10);
    return fail;
#undef CHOL_GSYNC
}

// ------------------------------------------------------------------------------------------------- legacy in-smem Cholesky
// (only for the small static system of the top level) Lower Cholesky of the leading n x n block of the (nrows x n) panel
// in shared memory, blocked left-looking; rows [n, nrows) are carried along (they end up multiplied by L^-T).
__device__ int cta_chol_smem(double* Ms, int ld, int n, int nrows) {
    __shared__ int fail;
    double *Tdiag, *Wd, *invd;
    solve_scratch(Tdiag, Wd, invd);
    if (threadIdx.x == 0) fail = 0;
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += B200_NB) {
        const int jb = (n - j0 < B200_NB) ? (n - j0) : B200_NB;
        if (j0 > 0) {
            double* Cb = Ms + (long long)j0 * ld + j0;
            cta_gemm_nt(Ms + (long long)j0 * ld, ld, nrows - j0, Ms + (long long)j0 * ld, ld, jb, 0, j0,
                        [&](int m, int c, double v) { Cb[(long long)m * ld + c] -= v; });
            __syncthreads();
        }
        if (threadIdx.x < 32) {
            const int f = warp_potrf16(Ms + (long long)j0 * ld + j0, ld, jb, Tdiag, invd, Wd, false);
            if (f && threadIdx.x == 0 && fail == 0) fail = j0 + f;
        }
        __syncthreads();
        // rows below the diagonal block: X <- X * L_jj^-T, right-looking in registers, one row per thread
        for (int row = j0 + jb + threadIdx.x; row < nrows; row += blockDim.x) {
            double* xr = Ms + (long long)row * ld + j0;
            double x[B200_NB];
#pragma unroll
            for (int c = 0; c < B200_NB; c++) x[c] = (c < jb) ? xr[c] : 0.0;
#pragma unroll
            for (int c = 0; c < B200_NB; c++) {
                if (c < jb) {
                    x[c] *= invd[c];
#pragma unroll
                    for (int c2 = c + 1; c2 < B200_NB; c2++)
                        if (c2 < jb) x[c2] = fma(-x[c], Tdiag[c * B200_NB + c2], x[c2]);
                }
            }
#pragma unroll
            for (int c = 0; c < B200_NB; c++) if (c < jb) xr[c] = x[c];
        }
        __syncthreads();
    }
    return fail;
}

// The same factorisation for a SMALL block (the static system: 58 x 58 plus the rhs row), right-looking with one barrier per pivot:
// after pivot c every entry (r, c2), c < c2 <= r, takes  M[r][c2] -= M[r][c] M[c2][c] / d_c  from the still unscaled column c, all
// threads of the CTA sharing the entries; the columns are scaled by 1 / sqrt(d_c) in one pass at the end.  ~2 x n barriers instead of
// the blocked version's serial 16 x 16 steps (one warp factorises while 15 wait): 58 x 58 in ~7 us instead of ~40.
// rs: n doubles of scratch.  Rows [n, nrows) ride along (forward substitution of the rhs row).  Returns 0 or 1 + failing pivot.
__device__ int cta_chol_rl_smem(double* Ms, int ld, int n, int nrows, double* rs) {
    __shared__ int fail;
    if (threadIdx.x == 0) fail = 0;
    __syncthreads();
    for (int c = 0; c < n; c++) {
        double d = Ms[(long long)c * ld + c];
        if (!(d > 0.0) || !(d < 1.0e300)) { if (threadIdx.x == 0 && fail == 0) fail = c + 1; d = 1.0; }
        const double inv = 1.0 / d;
        const int m = nrows - 1 - c;   // rows below the pivot
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int i = e / m, r = c + 1 + i, c2 = c + 1 + (e - i * m);
            if (c2 <= r && c2 < n) Ms[(long long)r * ld + c2] = fma(-(Ms[(long long)r * ld + c] * inv), Ms[(long long)c2 * ld + c], Ms[(long long)r * ld + c2]);
        }
        __syncthreads();
    }
    for (int c = threadIdx.x; c < n; c += blockDim.x) { const double d = Ms[(long long)c * ld + c]; rs[c] = (d > 0.0 && d < 1.0e300) ? rsqrt(d) : 1.0; }
    __syncthreads();
    for (int e = threadIdx.x; e < nrows * n; e += blockDim.x) {
        const int r = e / n, c = e - r * n;
        if (c <= r) Ms[(long long)r * ld + c] *= rs[c];   // (diagonal: d / sqrt(d))
    }
    __syncthreads();
    return fail;
}

// x <- L^-T x for the n x n lower factor in shared memory (ld); x in shared memory (tv).  Blocked: the 16 x 16 diagonal
// solves run in one warp (registers + shuffles), the updates of the remaining entries on all threads.
__device__ void cta_solve_LT(const double* Ms, int ld, int n, double* tv) {
    const int nblk = (n + B200_NB - 1) / B200_NB;
    for (int b = nblk - 1; b >= 0; b--) {
        const int j0 = b * B200_NB;
        const int jb = (n - j0 < B200_NB) ? (n - j0) : B200_NB;
        if (threadIdx.x < 32) {
            const int l = threadIdx.x;
            double t = (l < jb) ? tv[j0 + l] : 0.0;
            const double dinv = (l < jb) ? 1.0 / Ms[(long long)(j0 + l) * ld + j0 + l] : 0.0;
#pragma unroll
            for (int m = B200_NB - 1; m >= 0; m--) {
                if (m < jb) {
                    const double xm = __shfl_sync(0xffffffffu, t * dinv, m);
                    if (l == m) t = xm;
                    else if (l < m) t = fma(-Ms[(long long)(j0 + m) * ld + j0 + l], xm, t);
                }
            }
            if (l < jb) tv[j0 + l] = t;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < j0; c += blockDim.x) {
            double t = tv[c];
            for (int r = 0; r < jb; r++) t = fma(-Ms[(long long)(j0 + r) * ld + c], tv[j0 + r], t);
            tv[c] = t;
        }
        __syncthreads();
    }
}

// copy the lower triangle (rows [0, nr), columns [0, r] rounded up to a pair) of a global row-major matrix into shared rows
__device__ __forceinline__ void cp_lower_async(double* dst, int ld, const double* __restrict__ src, int ldg, int nr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < nr; r += nwarps)
        for (int c2 = lane; 2 * c2 <= r; c2 += 32) cp_async16(dst + (long long)r * ld + 2 * c2, src + (long long)r * ldg + 2 * c2);
}
// copy the upper triangle (columns [r & ~1, w)) of a global row-major matrix into shared rows
__device__ __forceinline__ void cp_upper_async(double* dst, int ld, const double* __restrict__ src, int ldg, int nr, int w) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < nr; r += nwarps)
        for (int c2 = (r >> 1) + lane; 2 * c2 < w; c2 += 32) cp_async16(dst + (long long)r * ld + 2 * c2, src + (long long)r * ldg + 2 * c2);
}

// ------------------------------------------------------------------------------------------------- the sweep
struct SweepArgs {
    // sources, indexed by src index (src0 + i)
    const double* Dsrc;  // [ntp*ntp]
    const double* Osrc;  // [nc*nc]   coupling (next node chain) x (this node chain)
    const double* Asrc;  // [ns1*ntp] static rows + rhs row
    const double* Ssrc;  // [ns1*ns1]
    int src0, n;
    const int* st_idx;   // storage index per node (nullptr: storage index == src index)
    int has_right;       // the last node is coupled to a right neighbour through Osrc
    const double* Oleft; // O block of the left separator (nullptr: none) -> arrow rows [ns1, ns1+nc)
    int h;               // arrow height of this sweep: ns1 (+ nc with a left separator)
    double lambda;
    double* Wst;         // [storage idx][ntp*ntp]   W^T = L^-T of the pivot block (upper triangle, lower part zero)
    double* Yst;         // [storage idx][hmax*ntp]  forward-substituted arrow rows y = p W^T
    double* F;           // 2 x [(nc+hmax)*nc] ping-pong fill buffers of this sweep (rows [0,nc): lower triangle only)
    double* Sacc;        // [hmax*hmax] (lower triangle) Schur complement of this sweep (chain_schur_kernel)
    int* status;
    int dense_o;         // the coupling blocks are dense (reduced systems of the upper levels), else block-sparse (dm.orow)
    int ar_lo, ar_hi;    // arrow rows handled by this CTA (row split of a chunk over several CTAs on the upper levels)
    int lead;            // this CTA stores the shared results (W) of the chunk
};

// shared-memory carve-up of the sweep (doubles): Ms [ntp*ld] | Yb [cr*ld] | Op [max(nc*owp, 2*nc*OSLD)]
#define B200_OSLD 20   // row stride of a 16-column slab of a dense coupling block (== 4 mod 8)
__host__ __device__ inline long long sweep_op_doubles(int nc, int owp) {
    const long long a = ((long long)nc * owp + 1) & ~1LL, b = 2LL * nc * B200_OSLD;
    return a > b ? a : b;
}
__host__ __device__ inline long long sweep_smem_doubles(int ntp, int ld, int cr, int nc, int owp) {
    return (long long)ntp * ld + (long long)cr * ld + sweep_op_doubles(nc, owp);
}
__host__ __device__ inline int static_lds(int ns) { int lds = ns + (ns & 1); while ((lds & 3) != 2) lds += 2; return lds; }

// [ka, kb) (even bounds, time-block coordinates) of the columns in which rows [r0, r0+16) of O can be non-zero
__device__ __forceinline__ void xrow_range(const Dims& dm, const int* orow, bool dense, int r0, int& ka, int& kb, int nrows = 16) {
    if (dense) { ka = dm.nl & ~1; kb = dm.ntp; return; }
    const int lane = threadIdx.x & 31;
    int lo = dm.ntp, hi = 0;
    const int r = r0 + (lane & 15);
    if (r < dm.nc && (lane & 15) < nrows) {
        const int c0 = orow[2 * r], c1 = orow[2 * r + 1];
        if (c1 > c0) { lo = (dm.nl + c0) & ~1; hi = (dm.nl + c1 + 1) & ~1; }
    }
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) {
        const int lo2 = __shfl_xor_sync(0xffffffffu, lo, o), hi2 = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = lo2 < lo ? lo2 : lo; hi = hi2 > hi ? hi2 : hi;
    }
    ka = lo; kb = hi > dm.ntp ? dm.ntp : hi;
}

// DENSE: the coupling blocks are dense (upper levels) -- a compile-time switch, so that the level-0 kernel carries none of the
// dense-level code (slab streaming) and the dense kernel none of the block-sparse paths
template <bool DENSE>
__device__ void cta_sweep(const Dims& dm, const SweepArgs& a, double* sm) {
    const int n = dm.ntp, nl = dm.nl, nc = dm.nc, ns1 = dm.ns1, ld = dm.ld, hmax = dm.hmax, CR = dm.cr, ow = dm.ow, owp = dm.owp;
    const long long Fsz = (long long)(nc + hmax) * nc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int mg = lane >> 2, mt = lane & 3;   // MMA fragment coordinates (micro-kernels above)
    double* Ms = sm;
    double* Yb = sm + (long long)n * ld;
    double* Op = Yb + (long long)CR * ld;
    const int kz0 = nl & ~1;                        // first column of z that is needed (chain columns, pair aligned)
#ifdef B200_PHASE_CLOCKS
    if (threadIdx.x == 0) g_phase_on = DENSE ? 0 : 1;
#endif
    // structure tables of O in shared memory (they are consulted at the start of every GEMM phase: keep them off the L2 path)
    __shared__ int s_orow[2 * 256], s_ozs[256], s_fblk[64], s_nfblk;
    for (int e = threadIdx.x; e < 2 * nc; e += blockDim.x) s_orow[e] = dm.orow[e];
    for (int e = threadIdx.x; e < nc; e += blockDim.x) s_ozs[e] = dm.ozs[e];
    __syncthreads();
    if (threadIdx.x == 0) {
        // 8-column blocks of the block-sparse fill: every run of rows of O that share a packed window is covered by blocks that
        // stay inside the run (the last one is shifted back and overlaps its neighbour), so a block contracts over ONE window
        int nbk = 0, i0 = 0;
        while (i0 < nc && nbk < 60) {
            int e = i0;
            while (e < nc && s_ozs[e] == s_ozs[i0]) e++;
            for (int j = i0; j < e && nbk < 60; j += 8) s_fblk[nbk++] = (j + 8 <= e) ? j : ((e - 8 > i0) ? e - 8 : i0);
            i0 = e;
        }
        s_nfblk = nbk;
    }
    __syncthreads();
    for (int i = 0; i < a.n; i++) {
        const long long src = a.src0 + i;
        const long long st = a.st_idx ? a.st_idx[i] : src;
        const bool has_x = nc > 0 && ((i + 1 < a.n) || a.has_right);
        const double* Fin = a.F + (long long)(i & 1) * Fsz;
        double* Fout = a.F + (long long)((i + 1) & 1) * Fsz;
        double* Wg = a.Wst + st * n * n;
        double* Yg = a.Yst + st * (long long)hmax * n;
        const double* Og = a.Osrc + src * nc * nc;
        const double* Ag = a.Asrc + src * ns1 * n;
        // ---------------- phase A: pivot block  M = D + lambda I - fill ; W^T = chol(M)^-T
        PH_T0();
        __syncthreads();   // previous node done with Ms / Yb / Op
        cp_lower_async(Ms, ld, a.Dsrc + src * n * n, n, n);
        if (has_x && !DENSE) {   // packed rows of O: Op[r][t] = O[r][ozs[r] - nl + t]
            for (int e = threadIdx.x; e < nc * ow; e += blockDim.x) {
                const int r = e / ow, t = e - r * ow;
                const int cc = s_ozs[r] - nl + t;
                if (cc >= 0 && cc < nc) cp_async8(Op + r * owp + t, Og + (long long)r * nc + cc);
                else Op[r * owp + t] = 0.0;
            }
        }
        if (i + 1 < a.n) {  // pull the next node's pivot block towards L2 while this one is factorised
            const double* Dn = a.Dsrc + (src + 1) * n * n;
            for (int e = threadIdx.x * 16; e < n * n; e += blockDim.x * 16) prefetch_l2(Dn + e);
        }
        // ... and this node's panel sources (read in phase B, after the factorisation)
        for (int e = threadIdx.x * 16; e < ns1 * n; e += blockDim.x * 16) prefetch_l2(Ag + e);
        if (has_x) for (int e = threadIdx.x * 16; e < nc * nc; e += blockDim.x * 16) prefetch_l2(Og + e);
        // fill of the previous node (chain x chain, lower triangle): loads issued before the wait, 8 rows x 4 columns per lane
        for (int rb = 0; rb < n; rb += 8 * nwarps) {
            for (int cb = 0; cb < n; cb += 128) {
                double f[8][4];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int r = rb + warp + q * nwarps;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int c = cb + lane + 32 * u;
                        f[q][u] = (i > 0 && r < n && r >= nl && c >= nl && c <= r) ? Fin[(long long)(r - nl) * nc + (c - nl)] : 0.0;
                    }
                }
                if (rb == 0 && cb == 0) { cp_async_wait_all(); __syncthreads(); }
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int r = rb + warp + q * nwarps;
                    if (r >= n) continue;
                    double* row = Ms + (long long)r * ld;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int c = cb + lane + 32 * u;
                        if (c >= n) continue;
                        if (c > r) row[c] = 0.0;
                        else row[c] = row[c] + ((c == r) ? a.lambda : 0.0) - f[q][u];
                    }
                }
            }
        }
        __syncthreads();
        PH_ADD(0);
        const int fail = cta_chol_inv(Ms, ld, n);
        PH_ADD(1);
        if (fail) {
#ifdef B200_USE_EMU
            if (threadIdx.x == 0 && std::getenv("B200_EMU_DEBUG")) printf("pivot failure: node %d (storage %lld) index %d\n", i, st, fail - 1);
#endif
            if (threadIdx.x == 0) atomicCAS(a.status, 0, (int)(st + 1));
            return;
        }
        // clear what is left of L below the diagonal (the GEMMs below rely on W^T being upper triangular) and store W^T
        for (int r = warp; r < n; r += nwarps)
            for (int c = 2 * lane; c < n; c += 64) {
                double2 v = ld2(Ms + (long long)r * ld + c);
                if (c < r) { v.x = 0.0; if (c + 1 < r) v.y = 0.0; st2(Ms + (long long)r * ld + c, v.x, v.y); }
                if (a.lead) st2(Wg + (long long)r * n + c, v.x, v.y);
            }
        PH_ADD(2);
        // ---------------- phase B: panel rows [X (nc, if has_x) ; arrow (h)] in chunks of <= CRg rows.
        // Block-sparse levels: the 16 warps form TWO groups of 8 that work through alternate chunks independently, each in its
        // own half of the panel buffer and with its own named barrier (ids 7, 8): one group's load / barrier / epilogue phases
        // overlap the other group's MMA phases.  Dense levels: one group of all warps (the O slabs stream through a shared buffer).
        const int xo = has_x ? nc : 0;
        const int ngroups = (!DENSE && (nwarps & 1) == 0 && ((CR >> 1) & ~15) >= 16) ? 2 : 1;
        const int gnw = nwarps / ngroups, gid = warp / gnw, gw = warp - gid * gnw;
        const int gthreads = gnw << 5, gtid = threadIdx.x - gid * gthreads;
        const int CRg = ngroups == 2 ? ((CR >> 1) & ~15) : CR;
        double* Ybg = Yb + (long long)gid * CRg * ld;
#define BSYNC() do { if (ngroups == 1) __syncthreads(); else named_bar_sync(7 + gid, gthreads); } while (0)
        const int cnt0 = xo, cnt1 = a.ar_hi - a.ar_lo;
        const int nch0 = (cnt0 + CRg - 1) / CRg, nch1 = (cnt1 + CRg - 1) / CRg;
        for (int qch = gid; qch < nch0 + nch1; qch += ngroups) {
            const int seg = qch < nch0 ? 0 : 1, ch = seg ? qch - nch0 : qch;
            const int base = seg == 0 ? 0 : xo + a.ar_lo, cnt = seg == 0 ? cnt0 : cnt1, nch = seg == 0 ? nch0 : nch1;
            int csz = (cnt + nch - 1) / nch;
            csz = (csz + 15) & ~15;
            const int r0 = base + ch * csz;
            const int cr = (base + cnt - r0 < csz) ? (base + cnt - r0) : csz;
            BSYNC();   // Ybg free
            // ---- load rows (one gw per row): sources through cp.async; the fill rows of the previous node are fetched
            // into registers before the wait (their L2 latency overlaps the copies) and applied afterwards
            for (int lr = gw; lr < cr; lr += gnw) {
                const int pr = r0 + lr;
                double* dst = Ybg + (long long)lr * ld;
                if (seg == 0) {
                    const double* og = Og + (long long)pr * nc - nl;
                    for (int c = lane; c < n; c += 32) {
                        if (c >= nl) cp_async8(dst + c, og + c);
                        else dst[c] = 0.0;
                    }
                } else {
                    const int ar = pr - xo;
                    if (ar < ns1) {
                        const double* ag = Ag + (long long)ar * n;
                        for (int c = 2 * lane; c < n; c += 64) cp_async16(dst + c, ag + c);
                    } else if (i == 0 && a.Oleft) {
                        const double* ol = a.Oleft + (ar - ns1);
                        for (int c = lane; c < n; c += 32) dst[c] = (c >= nl) ? ol[(long long)(c - nl) * nc] : 0.0;
                    } else {
                        for (int c = lane; c < nl; c += 32) dst[c] = 0.0;
                        if (i == 0) for (int c = nl + lane; c < n; c += 32) dst[c] = 0.0;
                    }
                }
            }
            if (seg == 1 && i > 0) {
                const int ar0 = r0 - xo;
                for (int rb = 0; rb < cr; rb += 4 * gnw) {
                    for (int cb = 0; cb < nc; cb += 128) {
                        double f[4][4];
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int lr = rb + gw + q * gnw;
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int c = cb + lane + 32 * u;
                                f[q][u] = (lr < cr && c < nc) ? Fin[(long long)(nc + ar0 + lr) * nc + c] : 0.0;
                            }
                        }
                        if (rb == 0 && cb == 0) { cp_async_wait_all(); BSYNC(); }
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const int lr = rb + gw + q * gnw;
                            if (lr >= cr) continue;
                            double* dr = Ybg + (long long)lr * ld + nl;
                            const bool src_row = (ar0 + lr) < ns1;   // static / rhs rows start from A, left rows from zero
#pragma unroll
                            for (int u = 0; u < 4; u++) {
                                const int c = cb + lane + 32 * u;
                                if (c < nc) dr[c] = (src_row ? dr[c] : 0.0) - f[q][u];
                            }
                        }
                    }
                }
            } else {
                cp_async_wait_all();
            }
            BSYNC();
            PH_ADD(3);
            const int rtiles = (cr + 15) >> 4;
            const int tpw = gnw / rtiles;              // column tiles per wave (rtiles <= CRg/16 <= gnw)
            // ---- y = p W^T, results written back in place after a barrier.  Balanced scheme (one wave): every gw of a row
            // group owns the 16-column blocks s and nb-1-s (short + long contraction); else column-tile waves from the right.
            const int nb16 = (n + 15) >> 4, npair = (nb16 + 1) >> 1;
            if (npair <= tpw) {
                const int slot = gw % tpw, rt = gw / tpw;
                const bool act = rt < rtiles && slot < npair;
                const int bA = slot, bB = nb16 - 1 - slot;   // bA <= bB ; equal: the middle block of an odd count
                Acc acc;
                acc_zero(acc);
                const int m0 = rt << 4;
                if (act) {
                    int ka = 0, kb = n;
                    if (seg == 0) xrow_range(dm, s_orow, DENSE, r0 + m0, ka, kb);
                    int kendA = (bA << 4) + 16; if (kendA > kb) kendA = kb;
                    int kendB = (bB << 4) + 16; if (kendB > kb) kendB = kb;
                    if (bA == bB) kendA = ka;   // single block: only the B half is used
                    warp_mma_nn2(Ybg, ld, m0, cr, Ms, ld, bA << 4, bB << 4, n, ka, kendA, kendB, acc);
                }
                BSYNC();
                if (act) {
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int lr = m0 + 8 * q + mg;
                        if (lr >= cr) continue;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (j < 2 && bA == bB) continue;
                            const int c = ((j < 2 ? bA : bB) << 4) + 8 * (j & 1) + 2 * mt;
                            if (c >= n) continue;
                            st2(Ybg + (long long)lr * ld + c, acc[q][j][0], acc[q][j][1]);
                            if (seg == 1) st2(Yg + (long long)(r0 - xo + lr) * n + c, acc[q][j][0], acc[q][j][1]);
                        }
                    }
                }
                BSYNC();
            } else {
                const int nct = (n + 31) >> 5;
                for (int hi = nct; hi > 0; hi -= tpw) {
                    const int lo = hi - tpw > 0 ? hi - tpw : 0;
                    const int slot = gw % tpw, rt = gw / tpw;
                    // spread long and short column tiles over the four schedulers (gw & 3)
                    const int ct = lo + (slot + rt) % tpw;
                    const bool act = rt < rtiles && ct < hi;
                    Acc acc;
                    acc_zero(acc);
                    const int m0 = rt << 4, c0 = ct << 5;
                    if (act) {
                        int ka = 0, kb = n;
                        if (seg == 0) xrow_range(dm, s_orow, DENSE, r0 + m0, ka, kb);
                        const int kend = (c0 + 32 < kb) ? c0 + 32 : kb;
                        warp_mma_nn(Ybg, ld, m0, cr, Ms, ld, c0, n, ka, kend, acc);
                    }
                    BSYNC();
                    if (act) {
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const int lr = m0 + 8 * q + mg;
                            if (lr >= cr) continue;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int c = c0 + 8 * j + 2 * mt;
                                if (c >= n) continue;
                                st2(Ybg + (long long)lr * ld + c, acc[q][j][0], acc[q][j][1]);
                                if (seg == 1) st2(Yg + (long long)(r0 - xo + lr) * n + c, acc[q][j][0], acc[q][j][1]);
                            }
                        }
                    }
                    BSYNC();
                }
            }
            PH_ADD(4);
            if (!has_x) continue;
            // ---- z = y W (chain columns only), in place after a barrier; same balanced pairing of 16-column blocks
            const int nbz = (n - kz0 + 15) >> 4, npz = (nbz + 1) >> 1;
            if (npz <= tpw) {
                const int slot = gw % tpw, rt = gw / tpw;
                const bool act = rt < rtiles && slot < npz;
                const int bA = slot, bB = nbz - 1 - slot;
                const int kA = kz0 + (bA << 4), kB = kz0 + (bB << 4);
                Acc acc;
                acc_zero(acc);
                const int m0 = rt << 4;
                if (act) {
                    int sA = kA, sB = kB;
                    if (seg == 0) {
                        // rows of O: y is zero left of ka; z is only needed up to the last column that feeds the LOWER triangle
                        int ka, kb; xrow_range(dm, s_orow, DENSE, r0 + m0, ka, kb);
                        if (ka > sA) sA = ka;
                        if (ka > sB) sB = ka;
                        if (!DENSE) { if (kA >= kb) sA = n; if (kB >= kb) sB = n; }
                    }
                    if (bA == bB) sA = n;
                    warp_mma_nt2(Ybg, ld, m0, cr, Ms, ld, kA, kB, n, sA, sB, n, acc);
                }
                BSYNC();
                if (act) {
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int lr = m0 + 8 * q + mg;
                        if (lr >= cr) continue;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            if (j < 2 && bA == bB) continue;
                            const int c = ((j < 2) ? kA : kB) + 8 * (j & 1) + 2 * mt;
                            if (c < n) st2(Ybg + (long long)lr * ld + c, acc[q][j][0], acc[q][j][1]);
                        }
                    }
                }
                BSYNC();
            } else {
                const int nct = (n - kz0 + 31) >> 5;
                for (int lo = 0; lo < nct; lo += tpw) {
                    const int hi = lo + tpw < nct ? lo + tpw : nct;
                    const int slot = gw % tpw, rt = gw / tpw;
                    const int ct = lo + (slot + rt) % tpw;
                    const bool act = rt < rtiles && ct < hi;
                    Acc acc;
                    acc_zero(acc);
                    const int m0 = rt << 4, k0 = kz0 + (ct << 5);
                    if (act) {
                        int ca = k0;
                        bool need = true;
                        if (seg == 0) {
                            int ka, kb; xrow_range(dm, s_orow, DENSE, r0 + m0, ka, kb);
                            if (ka > ca) ca = ka;
                            if (!DENSE && k0 >= kb) need = false;
                        }
                        if (need) warp_mma_nt(Ybg, ld, m0, cr, Ms, ld, k0, n, ca, n, acc);
                    }
                    BSYNC();
                    if (act) {
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const int lr = m0 + 8 * q + mg;
                            if (lr >= cr) continue;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
                                const int c = k0 + 8 * j + 2 * mt;
                                if (c < n) st2(Ybg + (long long)lr * ld + c, acc[q][j][0], acc[q][j][1]);
                            }
                        }
                    }
                    BSYNC();
                }
            }
            PH_ADD(5);
            // ---- fill rows  F = z[:, chain] O^T
            if (DENSE) {
                // dense coupling block: O streams through two 16-column slabs (cp.async), accumulators stay in registers
                const int nsl = (n - kz0 + 15) >> 4;
                const int nctf = (nc + 31) >> 5;
                auto load_slab = [&](int sidx, double* Os) {
                    const int cbase = kz0 - nl + 16 * sidx;
                    for (int e = gtid; e < nc * 16; e += gthreads) {
                        const int j = e >> 4, t = e & 15, cc = cbase + t;
                        if (cc >= 0 && cc < nc) cp_async8(Os + j * B200_OSLD + t, Og + (long long)j * nc + cc);
                        else Os[j * B200_OSLD + t] = 0.0;
                    }
                };
                for (int lo = 0; lo < nctf; lo += tpw) {
                    const int slot = gw % tpw, rt = gw / tpw;
                    const int ct = lo + slot;
                    const bool act = rt < rtiles && ct < nctf;
                    Acc acc;
                    acc_zero(acc);
                    const int m0 = rt << 4;
                    load_slab(0, Op);
                    for (int sidx = 0; sidx < nsl; sidx++) {
                        cp_async_wait_all();
                        BSYNC();
                        if (sidx + 1 < nsl) load_slab(sidx + 1, Op + ((sidx + 1) & 1) * nc * B200_OSLD);
                        int kw = n - kz0 - 16 * sidx; if (kw > 16) kw = 16;
                        if (act) warp_mma_nt(Ybg + kz0 + 16 * sidx, ld, m0, cr, Op + (sidx & 1) * nc * B200_OSLD, B200_OSLD, ct << 5, nc, 0, kw, acc);
                    }
                    if (act) {
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const int lr = m0 + 8 * q + mg;
                            if (lr >= cr) continue;
#pragma unroll
                            for (int j = 0; j < 4; j++) {
#pragma unroll
                                for (int e = 0; e < 2; e++) {
                                    const int ci = (ct << 5) + 8 * j + 2 * mt + e;
                                    if (ci < nc) Fout[(long long)(r0 + lr) * nc + ci] = acc[q][j][e];
                                }
                            }
                        }
                    }
                    BSYNC();
                }
            } else {
                // block-sparse coupling on the tensor cores: a warp task = 16 panel rows x four 8-column blocks of the table
                // s_fblk (8 rows of O sharing one packed window: the contraction runs over that window only, <= ow columns); the
                // four blocks advance in lockstep so that 8 independent accumulators are in flight.
                const int nfb = s_nfblk, ncq = (nfb + 3) >> 2, nst = (ow + 3) >> 2;
                for (int task = gw; task < rtiles * ncq; task += gnw) {
                    const int rt = task % rtiles, cq = task / rtiles;
                    const int m0 = rt << 4;
                    int zs[4], cs[4];
                    bool cval[4];
                    const double* bp[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int fb = (cq << 2) + j;
                        cs[j] = fb < nfb ? s_fblk[fb] : nc;
                        zs[j] = fb < nfb ? s_ozs[cs[j]] : n;
                        const int ci = cs[j] + mg;                        // this lane's B column of block j = row of O
                        cval[j] = ci < nc && s_ozs[ci < nc ? ci : 0] == zs[j];
                        bp[j] = Op + (cval[j] ? ci : 0) * owp;
                    }
                    const double* ap[2];
#pragma unroll
                    for (int q = 0; q < 2; q++) { int lr = m0 + 8 * q + mg; if (lr > cr - 1) lr = cr - 1; ap[q] = Ybg + (long long)lr * ld; }
                    Acc acc;
                    acc_zero(acc);
                    for (int sidx = 0; sidx < nst; sidx++) {
                        const int idx = 4 * sidx + mt;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int kk = zs[j] + idx;
                            const bool ok = kk < n && idx < ow;
                            const double bv = (cval[j] && idx < ow) ? bp[j][idx] : 0.0;
#pragma unroll
                            for (int q = 0; q < 2; q++) dmma884(acc[q][j][0], acc[q][j][1], ok ? ap[q][kk] : 0.0, bv);
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const int lr = m0 + 8 * q + mg;
                        if (lr >= cr) continue;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
#pragma unroll
                            for (int e = 0; e < 2; e++) {
                                const int co = cs[j] + 2 * mt + e;
                                if (co < nc && s_ozs[co] == zs[j]) Fout[(long long)(r0 + lr) * nc + co] = acc[q][j][e];
                            }
                        }
                    }
                }
            }
            PH_ADD(6);
        }
#undef BSYNC
        __threadfence_block();
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------- levels
// One level of the nested partition.  Level 0: nodes = keyframes (sources = assembled blocks, storage index = node index).
// Level l+1: nodes = the separators of level l (sources = the reduced blocks written by chain_reduce_kernel, storage index =
// the separator's keyframe).  The nodes of a level are cut into P chunks; the last node of every chunk but the last is that
// chunk's separator and becomes a node of the next level.  The top level is a Level with P == 1.
struct Level {
    const double *Dsrc, *Osrc, *Asrc, *Ssrc;  // per node: [ntp*ntp], [nc*nc], [ns1*ntp], [ns1*ns1]
    const int* st;                            // storage (keyframe) index per node, nullptr = identity
    const int* begin;                         // [P+1] chunk c covers nodes [begin[c], begin[c+1])
    int P;
    double *F;                                // [P][2][(nc+hmax)*nc] fill buffers (ping-pong) of each chunk sweep
    double *S;                                // [P][hmax*hmax] arrow Schur complement of each chunk sweep
    double *Dred, *Ored, *Ared, *Sred;        // [P-1] reduced sources == next level's node sources
    int* sep_st;                              // [P-1] storage index of separator c == next level's st
    int dense_o;                              // coupling blocks of this level are dense (reduced systems)
    int nsplit;                               // CTAs per chunk sweep: each one factorises the pivot blocks and processes the rows
                                              // of O (redundantly, bit-identical) plus its share of the arrow rows; F is [P][nsplit][2][..]
};

struct SolveBuffers {
    double *Wst, *Yst;                        // factors, per storage index (keyframe)
    double *ybuf;                             // [storage idx][ntp]  y = Y^T coef of the back-substitution pre-pass
    double *delta, *dstat;                    // solution (internal order), delta per storage index
    double* scal;                             // scalar outputs
    int* status;
    const double* lam;                        // lambda of the trial being solved (device scalar: written by the LM judge kernel / the host)
    int lam_in_sweep;                         // the sweeps add lambda to the pivot blocks (0: already in the sources, see kernels_local.cuh)
    const int* skip;                          // non-null and *skip != 0: this lambda trial is not needed any more, every kernel returns at once
};
__device__ __forceinline__ bool solve_skipped(const SolveBuffers& sb) { return sb.skip != nullptr && *sb.skip != 0; }

// arrow rows [lo, hi) of split y out of ns (h rows in total)
// (boundaries on multiples of 16 rows = whole row tiles of the sweep; a split may own no rows at all)
__device__ __forceinline__ int split_lo(int h, int ns, int y) { const int lo = 16 * (int)((long long)y * ((h + 15) >> 4) / ns); return lo < h ? lo : h; }
__device__ __forceinline__ int split_owner(int h, int ns, int row) { int y = 0; while (y + 1 < ns && split_lo(h, ns, y + 1) <= row) y++; return y; }

__device__ __forceinline__ void level_chunk_args(const Dims& dm, const Level& lv, const SolveBuffers& sb, int c, SweepArgs& a, int y = 0) {
    const int kb = lv.begin[c], ke = lv.begin[c + 1];
    a.Dsrc = lv.Dsrc; a.Osrc = lv.Osrc; a.Asrc = lv.Asrc; a.Ssrc = lv.Ssrc;
    a.src0 = kb;
    a.n = (c < lv.P - 1) ? (ke - 1 - kb) : (ke - kb);   // the last node of every chunk but the last is a separator
    a.st_idx = lv.st ? lv.st + kb : nullptr;
    a.has_right = (c < lv.P - 1);
    a.Oleft = (c > 0) ? lv.Osrc + (long long)(kb - 1) * dm.nc * dm.nc : nullptr;
    a.h = dm.ns1 + ((c > 0) ? dm.nc : 0);
    a.Wst = sb.Wst; a.Yst = sb.Yst;
    const int ns = lv.nsplit > 0 ? lv.nsplit : 1;
    a.F = lv.F + ((long long)c * ns + y) * 2 * (dm.nc + dm.hmax) * dm.nc;
    a.ar_lo = split_lo(a.h, ns, y); a.ar_hi = (y + 1 < ns) ? split_lo(a.h, ns, y + 1) : a.h;
    a.lead = (y == 0);
    a.Sacc = lv.S + (long long)c * dm.hmax * dm.hmax;
    a.status = sb.status;
    a.lambda = 0.0;
    a.dense_o = lv.dense_o;
}

__device__ void cta_sweep_pipe(const Dims& dm, const SweepArgs& a, double* sm);   // kernels_sweep_pipe.cuh
__device__ void cta_sweep_dense_stream(const Dims& dm, const SweepArgs& a, double* sm);

// grid = number of chunks of this level owned by this rank (chunk = c_first + blockIdx.x)
template <bool DENSE>
__global__ void __launch_bounds__(B200_SOLVE_THREADS) chain_sweep_kernel(Dims dm, Level lv, SolveBuffers sb, int c_first) {
    B200_DYN_SMEM(double, sm);
    if (solve_skipped(sb)) return;
    SweepArgs a;
    level_chunk_args(dm, lv, sb, c_first + blockIdx.x, a, blockIdx.y);
    a.lambda = sb.lam_in_sweep ? *sb.lam : 0.0;
    if (!DENSE && dm.pipe && gridDim.y == 1 && a.n > 0) cta_sweep_pipe(dm, a, sm);   // keyframe-pipelined variant (kernels_sweep_pipe.cuh)
    else if (DENSE && dm.pipe == 1 && a.n > 0) cta_sweep_dense_stream(dm, a, sm);      // dense levels, row tiles streamed through Tensor Memory
    else cta_sweep<DENSE>(dm, a, sm);
}

// Arrow Schur complement of one chunk sweep:  Sacc = sum_nodes ( S_node - Y_node Y_node^T ), lower triangle of h x h.
// One CTA per chunk, 12 warps; every warp owns up to two 32 x 32 output tiles whose accumulators stay in registers over
// all nodes of the chunk (fixed summation order, no atomics).  The stored panel of a node is staged through shared memory.
__global__ void __launch_bounds__(B200_SCHUR_THREADS, 1) chain_schur_kernel(Dims dm, Level lv, SolveBuffers sb, int c_first, int ldy, int kslab) {
    B200_DYN_SMEM(double, Ys);
    if (solve_skipped(sb)) return;
    SweepArgs a;
    level_chunk_args(dm, lv, sb, c_first + blockIdx.x, a);
    const int n = dm.ntp, ns1 = dm.ns1, hmax = dm.hmax, h = a.h;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int mg = lane >> 2, mt = lane & 3;   // MMA fragment coordinates
    if (*sb.status != 0) return;
    const int nrt = (h + 31) >> 5;
    const int ntiles = nrt * (nrt + 1) / 2;
    const int nsplit = gridDim.y, part = blockIdx.y;          // output tiles dealt round-robin to the CTAs of a chunk
    const int ntl = (ntiles - part + nsplit - 1) / nsplit;    // tiles of this CTA: part, part + nsplit, ...
    // Tiles of this CTA are dealt to (warp, slot) pairs so that the MMA work -- live 8 x 8 blocks: 16 for a full tile, 10 on the
    // diagonal, fewer in the last tile row -- is balanced over the four schedulers (warp & 3): heaviest tile first, to the
    // scheduler with the least work so far, within it to the least loaded warp that still has a free slot.
    __shared__ int s_asg[32][2], s_wgt[64];
    auto tile_rc = [](int t, int& r, int& c) { r = 0; while ((r + 1) * (r + 2) / 2 <= t) r++; c = t - r * (r + 1) / 2; };
    auto tile_ni = [&](int r) { const int rows = h - (r << 5); return rows >= 32 ? 4 : (rows + 7) >> 3; };
    auto tile_weight = [&](int t) { int r, c; tile_rc(t, r, c); const int ni = tile_ni(r); return r == c ? ni * (ni + 1) / 2 : ni * 4; };
    for (int base = 0; base < ntl; base += 2 * nwarps) {
        const int nhere = (ntl - base < 2 * nwarps) ? (ntl - base) : 2 * nwarps;
        __syncthreads();
        const bool balance = a.n >= 8;   // the serial assignment below costs ~10 us: only worth it for a long chunk (level 0)
        if (!balance) {
            if ((int)threadIdx.x < 2 * nwarps) {
                const int w = threadIdx.x % nwarps, sl = threadIdx.x / nwarps;
                s_asg[w][sl] = ((int)threadIdx.x < nhere) ? (base + (int)threadIdx.x) * nsplit + part : -1;
            }
        } else if ((int)threadIdx.x < nhere) s_wgt[threadIdx.x] = tile_weight((base + (int)threadIdx.x) * nsplit + part);
        __syncthreads();
        if (balance && threadIdx.x == 0) {
            int wload[32], sload[4] = {0, 0, 0, 0}, nslot[32];
            for (int w = 0; w < nwarps; w++) { wload[w] = 0; nslot[w] = 0; s_asg[w][0] = -1; s_asg[w][1] = -1; }
            for (int it = 0; it < nhere; it++) {
                int best = -1, bw = -1;
                for (int i = 0; i < nhere; i++) {
                    const int w = s_wgt[i];
                    if (w > bw) { bw = w; best = i; }
                }
                s_wgt[best] = -1;
                int tw = -1;
                for (int w = 0; w < nwarps; w++) {
                    if (nslot[w] >= 2) continue;
                    if (tw < 0 || sload[w & 3] < sload[tw & 3] || (sload[w & 3] == sload[tw & 3] && wload[w] < wload[tw])) tw = w;
                }
                s_asg[tw][nslot[tw]++] = (base + best) * nsplit + part;
                wload[tw] += bw; sload[tw & 3] += bw;
            }
        }
        __syncthreads();
        int tr[2], tc[2], ni[2];
        bool act[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int t = s_asg[warp][s];
            act[s] = t >= 0;
            tile_rc(act[s] ? t : 0, tr[s], tc[s]);
            ni[s] = tile_ni(tr[s]);
        }
        Acc32 acc0, acc1;
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) { acc0[i][j][0] = 0.0; acc0[i][j][1] = 0.0; acc1[i][j][0] = 0.0; acc1[i][j][1] = 0.0; }
        for (int i = 0; i < a.n; i++) {
            const long long st = a.st_idx ? a.st_idx[i] : (long long)a.src0 + i;
            const double* Yg = a.Yst + st * (long long)hmax * n;
            for (int k0 = 0; k0 < n; k0 += kslab) {
                const int kw = (n - k0 < kslab) ? (n - k0) : kslab;
                __syncthreads();
                for (int r = warp; r < h; r += nwarps)
                    for (int c = 2 * lane; c < kw; c += 64) cp_async16(Ys + (long long)r * ldy + c, Yg + (long long)r * n + k0 + c);
                cp_async_wait_all();
                __syncthreads();
                if (act[0]) warp_mma_nt32_kind(ni[0], tr[0] == tc[0], Ys, ldy, tr[0] << 5, h, tc[0] << 5, 0, kw, acc0);
                if (act[1]) warp_mma_nt32_kind(ni[1], tr[1] == tc[1], Ys, ldy, tr[1] << 5, h, tc[1] << 5, 0, kw, acc1);
            }
        }
        auto store_tile = [&](int trs, int tcs, const Acc32& acc) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int a1 = (trs << 5) + 8 * i + mg;
#pragma unroll
                for (int j = 0; j < 4; j++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int a2 = (tcs << 5) + 8 * j + 2 * mt + e;
                        if (a1 < h && a2 <= a1) a.Sacc[(long long)a1 * hmax + a2] = -acc[i][j][e];
                    }
                }
            }
        };
        if (act[0]) store_tile(tr[0], tc[0], acc0);
        if (act[1]) store_tile(tr[1], tc[1], acc1);
    }
    __syncthreads();
    // the nodes' own static blocks
    for (int e = threadIdx.x; e < ns1 * ns1; e += blockDim.x) {
        const int a1 = e / ns1, a2 = e - a1 * ns1;
        if (a2 > a1) continue;
        const int trr = a1 >> 5, tcc = a2 >> 5;
        if ((trr * (trr + 1) / 2 + tcc) % nsplit != part) continue;   // the CTA that wrote the entry adds to it
        double s = 0.0;
        for (int i = 0; i < a.n; i++) s += a.Ssrc[((long long)a.src0 + i) * ns1 * ns1 + e];
        a.Sacc[(long long)a1 * hmax + a2] += s;
    }
}

// The same Schur complement for the upper levels (few chunks of 2..4 nodes, most SMs idle): ONE 32 x 32 output tile per CTA
// (grid = chunks x tiles of the lower triangle), only the two 32-row slices of the panel that the tile needs are staged, and the four
// warps split the contraction range (partial tiles added in warp order: fixed summation order).  The nodes' own static blocks are
// added in the same pass.  27-36 us per launch against 45-52 us of chain_schur_kernel with its tiles dealt to 21 CTAs.
#define B200_SCHUR_TILE_THREADS 128
__host__ __device__ inline long long schur_tile_smem_doubles(int ldy) { return 64LL * ldy; }   // (>= 4 x 32 x 33 for the partial tiles)
__global__ void __launch_bounds__(B200_SCHUR_TILE_THREADS) chain_schur_tile_kernel(Dims dm, Level lv, SolveBuffers sb, int c_first, int ldy) {
    B200_DYN_SMEM(double, Ys);
    if (solve_skipped(sb)) return;
    SweepArgs a;
    level_chunk_args(dm, lv, sb, c_first + blockIdx.x, a);
    const int n = dm.ntp, ns1 = dm.ns1, hmax = dm.hmax, h = a.h;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int mg = lane >> 2, mt = lane & 3;
    if (*sb.status != 0) return;
    const int nrt = (h + 31) >> 5;
    const int t = blockIdx.y;
    if (t >= nrt * (nrt + 1) / 2) return;
    int tr = 0;
    while ((tr + 1) * (tr + 2) / 2 <= t) tr++;
    const int tc = t - tr * (tr + 1) / 2;
    const bool diag = tr == tc;
    const int rowsA = (h - (tr << 5) < 32) ? (h - (tr << 5)) : 32;     // valid rows of the tile
    const int ni = (rowsA + 7) >> 3;
    const int M = diag ? rowsA : 64, r0 = diag ? 0 : 32;                // staged rows: [0, 32) = output rows, [32, 64) = output columns
    const int kq = ((n + 4 * nwarps - 1) / (4 * nwarps)) * 4;           // contraction range of a warp (multiple of 4)
    const int ka = warp * kq, kb = (ka + kq < n) ? (ka + kq) : n;
    Acc32 acc;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    for (int r = rowsA + warp; r < 32; r += nwarps)   // output rows beyond h (last tile row): defined zeros, results discarded
        for (int c = lane; c < n; c += 32) Ys[(long long)r * ldy + c] = 0.0;
    for (int i = 0; i < a.n; i++) {
        const long long st = a.st_idx ? a.st_idx[i] : (long long)a.src0 + i;
        const double* Yg = a.Yst + st * (long long)hmax * n;
        __syncthreads();
        for (int r = warp; r < (diag ? rowsA : 64); r += nwarps) {
            const int gr = (r < 32) ? (tr << 5) + r : (tc << 5) + r - 32;
            if (r < 32 && r >= rowsA) continue;
            for (int c = 2 * lane; c < n; c += 64) cp_async16(Ys + (long long)r * ldy + c, Yg + (long long)gr * n + c);
        }
        cp_async_wait_all();
        __syncthreads();
        if (ka < kb) warp_mma_nt32_kind(ni, diag, Ys, ldy, 0, M, r0, ka, kb, acc);
    }
    __syncthreads();
    // partial tiles -> shared memory (stride 33), then entry (r, c) = sum over the warps in order, minus sign, plus the static blocks
    double* red = Ys;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int e = 0; e < 2; e++) red[(warp * 32 + 8 * i + mg) * 33 + 8 * j + 2 * mt + e] = acc[i][j][e];
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
        const int r = e >> 5, c = e & 31;
        const int a1 = (tr << 5) + r, a2 = (tc << 5) + c;
        if (a1 >= h || a2 > a1) continue;
        double s = 0.0;
        for (int w = 0; w < nwarps; w++) s += red[(w * 32 + r) * 33 + c];
        double v = -s;
        if (a1 < ns1) {   // (a2 <= a1 < ns1) the nodes' own static blocks
            double q = 0.0;
            for (int i = 0; i < a.n; i++) q += a.Ssrc[((long long)a.src0 + i) * ns1 * ns1 + a1 * ns1 + a2];
            v += q;
        }
        a.Sacc[(long long)a1 * hmax + a2] = v;
    }
}

// Schur complements onto separator c of this level: from chunk c (its right side) and chunk c+1 (its left side).
// grid = separators handled (separator = c_first + blockIdx.x)
__global__ void __launch_bounds__(256) chain_reduce_kernel(Dims dm, Level lv, int c_first, const int* __restrict__ skip) {
    if (skip != nullptr && *skip != 0) return;
    const int c = c_first + blockIdx.x;
    const int P = lv.P;
    const int ntp = dm.ntp, nl = dm.nl, nc = dm.nc, ns1 = dm.ns1, hmax = dm.hmax;
    const long long q = lv.begin[c + 1] - 1;                 // node index of the separator
    const long long Fsz = (long long)(nc + hmax) * nc;
    const int n_c = (int)(q - lv.begin[c]);                  // interior nodes of chunk c
    const int nsp = lv.nsplit > 0 ? lv.nsplit : 1;
    // fill buffers: rows of O from split 0, arrow row a from the split that owns it
    const double* Fc = lv.F + ((long long)c * nsp * 2 + (n_c & 1)) * Fsz;   // written by its last node (split 0)
    const int h_c = ns1 + (c > 0 ? nc : 0), h_n = ns1 + nc;
    const double* Sn = lv.S + (long long)(c + 1) * hmax * hmax;        // chunk c+1 (has left rows)
    if (threadIdx.x == 0 && blockIdx.y == 0) lv.sep_st[c] = lv.st ? lv.st[q] : (int)q;
    double* Dr = lv.Dred + (long long)c * ntp * ntp;
    const double* Dg = lv.Dsrc + q * ntp * ntp;
    double* Ar = lv.Ared + (long long)c * ns1 * ntp;
    const double* Ag = lv.Asrc + q * ns1 * ntp;
    double* Sr = lv.Sred + (long long)c * ns1 * ns1;
    const double* Sg = lv.Ssrc + q * ns1 * ns1;
    const double* S0 = lv.S;
    // coupling separator c -> separator c+1 through chunk c+1: -(fill of the left rows) of that chunk's last node
    const bool has_o = c < P - 2;
    const int n_n = has_o ? (int)(lv.begin[c + 2] - 1 - lv.begin[c + 1]) : 0;
    const double* Fn = lv.F + ((long long)(c + 1) * nsp * 2 + (n_n & 1)) * Fsz;
    double* Or = lv.Ored + (long long)c * nc * nc;
    const int e1 = ntp * ntp, e2 = e1 + ns1 * ntp, e3 = e2 + ns1 * ns1, e4 = e3 + (has_o ? nc * nc : 0);
    // value and destination of entry e.  Four entries per pass: their loads are all issued before the first store (the source and
    // destination pointers may alias as far as the compiler knows, so a load written after a store would wait for it)
    auto entry = [&](int e, double*& dst) -> double {
        if (e < e1) {
            const int r = e / ntp, col = e % ntp;
            double v = Dg[e];
            if (r >= nl && col >= nl) {
                const int i = r - nl, j = col - nl;
                const int hi = i > j ? i : j, lo = i > j ? j : i;
                if (n_c > 0) v -= Fc[(long long)hi * nc + lo];
                v += Sn[(long long)(ns1 + hi) * hmax + (ns1 + lo)];
            }
            dst = Dr + e;
            return v;
        } else if (e < e2) {
            const int f = e - e1, ar = f / ntp, col = f % ntp;
            double v = Ag[f];
            if (col >= nl) {
                if (n_c > 0) v -= Fc[(long long)split_owner(h_c, nsp, ar) * 2 * Fsz + (long long)(nc + ar) * nc + (col - nl)];
                v += Sn[(long long)(ns1 + (col - nl)) * hmax + ar];
            }
            dst = Ar + f;
            return v;
        } else if (e < e3) {
            const int f = e - e2, r = f / ns1, col = f % ns1;
            const int hi = r > col ? r : col, lo = r > col ? col : r;
            double v = Sg[f] + Sn[(long long)hi * hmax + lo];
            if (c == 0) v += S0[(long long)hi * hmax + lo];
            dst = Sr + f;
            return v;
        } else {
            const int f = e - e3, i = f / nc, j = f % nc;   // Or[j][i]: row = chain j of separator c+1, col = chain i of separator c
            dst = Or + (long long)j * nc + i;
            return -Fn[(long long)split_owner(h_n, nsp, ns1 + i) * 2 * Fsz + (long long)(nc + ns1 + i) * nc + j];
        }
    };
    const int stride = blockDim.x * gridDim.y;
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < e4; e += 4 * stride) {
        double v[4];
        double* d[4];
#pragma unroll
        for (int q = 0; q < 4; q++) { d[q] = nullptr; v[q] = 0.0; if (e + q * stride < e4) v[q] = entry(e + q * stride, d[q]); }
#pragma unroll
        for (int q = 0; q < 4; q++) if (d[q]) *d[q] = v[q];
    }
}

// ------------------------------------------------------------------------------------------------- back-substitution
// delta_k = -W^T ( W_c O_k^T delta_{k+1}(chain) + Y_k^T coef ),  coef = [delta_static (ns), -1, delta_left_chain (nc)].
// The second term does not depend on the neighbour: chain_backsub_prepass_kernel computes it for all nodes in parallel.
__device__ __forceinline__ void load_coef(const Dims& dm, const Level& lv, const SolveBuffers& sb, int c, double* coef) {
    const int ns = dm.ns, ns1 = dm.ns1, nc = dm.nc, nl = dm.nl, n = dm.ntp;
    for (int i = threadIdx.x; i < ns; i += blockDim.x) coef[i] = sb.dstat[i];
    if (threadIdx.x == 0) coef[ns] = -1.0;
    if (c > 0) {
        const int kb = lv.begin[c];
        const long long lst = lv.st ? lv.st[kb - 1] : kb - 1;
        for (int j = threadIdx.x; j < nc; j += blockDim.x) coef[ns1 + j] = sb.delta[lst * n + nl + j];
    }
}

// y[c] = sum_r Y[r][c] coef[r]   (h x n panel in global memory, n even).  Thread = (column pair, row group): 16-byte loads, up to four
// row groups (the first 256 threads), two independent partial sums per column; part: 512 doubles of scratch.
__device__ __forceinline__ void cta_yt_coef(const double* __restrict__ Yg, int n, int h, const double* coef, double* part, double* y) {
    const int cp = threadIdx.x & 63, rg = threadIdx.x >> 6;
    const int nrg = ((int)(blockDim.x >> 6) < 4) ? (int)(blockDim.x >> 6) : 4;
    for (int cb = 0; cb < n; cb += 128) {
        const int c = cb + 2 * cp;
        if (rg < nrg) {
            double2 t0 = make_double2(0.0, 0.0), t1 = make_double2(0.0, 0.0);
            if (c < n) {
                int r = rg;
                for (; r + nrg < h; r += 2 * nrg) {
                    const double2 a0 = ld2(Yg + (long long)r * n + c), a1 = ld2(Yg + (long long)(r + nrg) * n + c);
                    const double c0 = coef[r], c1 = coef[r + nrg];
                    t0.x = fma(a0.x, c0, t0.x); t0.y = fma(a0.y, c0, t0.y);
                    t1.x = fma(a1.x, c1, t1.x); t1.y = fma(a1.y, c1, t1.y);
                }
                if (r < h) {
                    const double2 a0 = ld2(Yg + (long long)r * n + c);
                    t0.x = fma(a0.x, coef[r], t0.x); t0.y = fma(a0.y, coef[r], t0.y);
                }
            }
            part[2 * (rg * 64 + cp)] = t0.x + t1.x;
            part[2 * (rg * 64 + cp) + 1] = t0.y + t1.y;
        }
        __syncthreads();
        if (rg == 0 && c < n) {
            double s0 = 0.0, s1 = 0.0;
            for (int q = 0; q < nrg; q++) { s0 += part[2 * (q * 64 + cp)]; s1 += part[2 * (q * 64 + cp) + 1]; }
            y[c] = s0; y[c + 1] = s1;
        }
        __syncthreads();
    }
}

// grid = (max nodes per chunk, chunks owned)
__global__ void __launch_bounds__(B200_PREPASS_THREADS) chain_backsub_prepass_kernel(Dims dm, Level lv, SolveBuffers sb, int c_first) {
    __shared__ double coef[512];
    __shared__ double part[512];
    if (solve_skipped(sb)) return;
    const int c = c_first + blockIdx.y;
    SweepArgs a;
    level_chunk_args(dm, lv, sb, c, a);
    if ((int)blockIdx.x >= a.n) return;
    const long long st = a.st_idx ? a.st_idx[blockIdx.x] : (long long)a.src0 + blockIdx.x;
    load_coef(dm, lv, sb, c, coef);
    __syncthreads();
    cta_yt_coef(a.Yst + st * (long long)dm.hmax * dm.ntp, dm.ntp, a.h, coef, part, sb.ybuf + st * dm.ntp);
}

// shared-memory carve-up of the back-substitution (doubles): Ms [ntp*ld] | u [256] | v [256] | cx [256] | yv [256] | coef [512] | part [512]
// | Os [nc*nc] (the coupling block of the node in flight, only when dm.bs_os: it fits next to the rest)
__host__ __device__ inline long long backsub_smem_doubles(int ntp, int ld, int nc_os = 0) {
    return (long long)ntp * ld + 4 * 256 + 512 + 512 + (long long)nc_os * nc_os;
}

// back-substitution over the nodes of a sweep (reverse order).  right_st: storage index of the right neighbour (or -1).
// y_ready: y = Y^T coef was computed by the pre-pass (sb.ybuf), else it is computed here from coef (shared memory).
// Nothing on the chain waits for global memory after the first node: the factor W^T and y of the NEXT node travel through
// registers while this node is processed, its coupling block O is copied (cp.async) into shared memory as soon as this node's
// O^T delta is done, and the step of this node is handed to the next one in shared memory.
__device__ void cta_backsub(const Dims& dm, const SweepArgs& a, const double* coef, long long right_st, const SolveBuffers& sb, bool y_ready,
                            double* sm) {
    const int n = dm.ntp, nl = dm.nl, nc = dm.nc, ld = dm.ld, hmax = dm.hmax;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    double* Ms = sm;
    double* u = sm + (long long)n * ld;
    double* v = u + 256;
    double* cx = v + 256;
    double* yv = cx + 256;
    double* part = yv + 256 + 512;
    double* Os = part + 512;
    double* delta = sb.delta;
    const bool os_ok = dm.bs_os != 0 && nc > 0;
    // structural part of the coupling block of node `src` -> Os (same row-major layout as in global memory).  Sparse coupling
    // (level 0): every thread owns up to NOS structural entries, their offsets found once (registers), not per node.
    constexpr int NOS = 6;
    int ooff[NOS];
    bool o_list = false;
    if (os_ok && !a.dense_o) {
        int total = 0;
        for (int r = 0; r < nc; r++) total += dm.orow[2 * r + 1] - dm.orow[2 * r];
        o_list = total <= NOS * (int)blockDim.x;
        int r = 0, row_begin = 0, len = nc > 0 ? dm.orow[1] - dm.orow[0] : 0;
#pragma unroll
        for (int j = 0; j < NOS; j++) {
            const int e = j * (int)blockDim.x + (int)threadIdx.x;
            ooff[j] = -1;
            if (o_list && e < total) {
                while (e >= row_begin + len) { row_begin += len; r++; len = dm.orow[2 * r + 1] - dm.orow[2 * r]; }
                ooff[j] = r * nc + dm.orow[2 * r] + (e - row_begin);
            }
        }
    }
    auto stage_o = [&](long long src) {
        const double* Og = a.Osrc + src * nc * nc;
        if (a.dense_o) {
            if ((nc & 1) == 0) for (int e = 2 * threadIdx.x; e < nc * nc; e += 2 * blockDim.x) cp_async16(Os + e, Og + e);
            else for (int e = threadIdx.x; e < nc * nc; e += blockDim.x) cp_async8(Os + e, Og + e);
        } else if (o_list) {
#pragma unroll
            for (int j = 0; j < NOS; j++) if (ooff[j] >= 0) cp_async8(Os + ooff[j], Og + ooff[j]);
        } else {
            for (int r = warp; r < nc; r += nwarps) {
                const int c0 = dm.orow[2 * r], c1 = dm.orow[2 * r + 1];
                for (int c = c0 + lane; c < c1; c += 32) cp_async8(Os + (long long)r * nc + c, Og + (long long)r * nc + c);
            }
        }
    };
    // structural row range of this thread's column of O (u = O^T delta_next), found once
    int ur0 = 0, ur1 = nc;
    if (!a.dense_o && (int)threadIdx.x < nc) { ur0 = dm.ocol[2 * threadIdx.x]; ur1 = dm.ocol[2 * threadIdx.x + 1]; }
    // every thread owns up to NPRE 16-byte pairs of the upper triangle of W^T (rows r, columns from r & ~1): loaded right after the
    // barrier that opens a node, written to shared memory after the node's last barrier
    constexpr int NPRE = 8;
    int goff[NPRE], soff[NPRE];
    {
        const int half = (n + 1) >> 1;
        long long total = 0;
        for (int r = 0; r < n; r++) total += half - (r >> 1);
        const bool fits = total <= (long long)NPRE * blockDim.x;
        int r = 0;
        long long row_begin = 0;   // flat index of the first pair of row r
#pragma unroll
        for (int j = 0; j < NPRE; j++) {
            const long long e = (long long)j * blockDim.x + threadIdx.x;
            goff[j] = -1; soff[j] = -1;
            if (fits && e < total) {
                while (e >= row_begin + (half - (r >> 1))) { row_begin += half - (r >> 1); r++; }
                const int c = 2 * ((r >> 1) + (int)(e - row_begin));
                goff[j] = r * n + c; soff[j] = r * ld + c;
            }
        }
        if (!fits) goff[0] = -2;   // block too large for the register path: every node is copied synchronously
    }
    const bool reg_path = goff[0] != -2;
    const bool y_pre = y_ready && reg_path && n <= (int)blockDim.x;   // y of the next node through a register as well
    bool staged = false;           // Ms / yv already hold this node's factor and y (written from registers at the end of the previous node)
    bool first = true;
    for (int i = a.n - 1; i >= 0; i--) {
        const long long src = a.src0 + i;
        const long long st = a.st_idx ? a.st_idx[i] : src;
        const bool has_x = nc > 0 && ((i + 1 < a.n) || a.has_right);
        long long nxt = -1;
        if (i + 1 < a.n) nxt = a.st_idx ? a.st_idx[i + 1] : src + 1;
        else if (a.has_right) nxt = right_st;
        __syncthreads();
        if (!staged) cp_upper_async(Ms, ld, a.Wst + st * n * n, n, n, n);
        if (first && has_x && os_ok) stage_o(src);
        double2 pre[NPRE];
        double ypre = 0.0;
        const bool prefetch = reg_path && i > 0;
        if (prefetch) {
            const long long stn = a.st_idx ? a.st_idx[i - 1] : src - 1;
            const double* Wn = a.Wst + stn * n * n;
#pragma unroll
            for (int j = 0; j < NPRE; j++) pre[j] = goff[j] >= 0 ? ld2(Wn + goff[j]) : make_double2(0.0, 0.0);
            if (y_pre && (int)threadIdx.x < n) ypre = sb.ybuf[stn * n + threadIdx.x];
        }
        if (first) {
            if (has_x) for (int j = threadIdx.x; j < nc; j += blockDim.x) cx[j] = delta[nxt * n + nl + j];
            if (os_ok) cp_async_wait_all();
        }
        if (y_ready && !(staged && y_pre)) for (int c = threadIdx.x; c < n; c += blockDim.x) yv[c] = sb.ybuf[st * n + c];
        if (first || !(staged && y_pre)) __syncthreads();
        if (!y_ready) cta_yt_coef(a.Yst + st * (long long)hmax * n, n, a.h, coef, part, yv);
        if (has_x) {
            // u = O^T delta_next (structural row range of every column)
            const double* Ob = os_ok ? Os : a.Osrc + src * nc * nc;
            for (int k = threadIdx.x; k < nc; k += blockDim.x) {
                const int r0 = (k == (int)threadIdx.x) ? ur0 : (a.dense_o ? 0 : dm.ocol[2 * k]), r1 = (k == (int)threadIdx.x) ? ur1 : (a.dense_o ? nc : dm.ocol[2 * k + 1]);
                double s0 = 0.0, s1 = 0.0;
                int r = r0;
                for (; r + 1 < r1; r += 2) {
                    s0 = fma(Ob[(long long)r * nc + k], cx[r], s0);
                    s1 = fma(Ob[(long long)(r + 1) * nc + k], cx[r + 1], s1);
                }
                if (r < r1) s0 = fma(Ob[(long long)r * nc + k], cx[r], s0);
                u[k] = s0 + s1;
            }
        }
        cp_async_wait_all();
        __syncthreads();
        if (os_ok && i > 0) stage_o(src - 1);   // (lands during the two products below; waited for before the next node opens)
        // v = y + W[:, chain] u :  v[c] = y[c] + sum_{k: nl+k <= c} Wt[nl+k][c] u[k].  Four adjacent lanes share a column
        // (k = q, q + 4, ...), partial sums meet through two shuffles: the dependent FMA chain is a quarter as long
        for (int base = 0; base < 4 * n; base += blockDim.x) {
            const int idx = base + threadIdx.x;
            const int c = idx >> 2, q = idx & 3;
            double s = 0.0;
            if (has_x && c < n) {
                const int kmax = c - nl;   // inclusive
                for (int k = q; k <= kmax; k += 4) s = fma(Ms[(long long)(nl + k) * ld + c], u[k], s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (q == 0 && c < n) v[c] = yv[c] + s;
        }
        __syncthreads();
        // delta[k] = - sum_{c >= k} Wt[k][c] v[c]; the chain part is handed to the next node in shared memory (cx)
        for (int base = 0; base < 4 * n; base += blockDim.x) {   // four adjacent lanes per row (c = k + q, k + q + 4, ...)
            const int idx = base + threadIdx.x;
            const int k = idx >> 2, q = idx & 3;
            double s = 0.0;
            if (k < n) {
                const double* row = Ms + (long long)k * ld;
                for (int c = k + q; c < n; c += 4) s = fma(row[c], v[c], s);
            }
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (q == 0 && k < n) {
                delta[st * n + k] = -s;
                if (k >= nl) cx[k - nl] = -s;
            }
        }
        staged = false;
        if (prefetch) {
            __syncthreads();   // every warp is done with this node's factor
#pragma unroll
            for (int j = 0; j < NPRE; j++) if (soff[j] >= 0) st2(Ms + soff[j], pre[j].x, pre[j].y);
            if (y_pre && (int)threadIdx.x < n) yv[threadIdx.x] = ypre;
            staged = true;
        }
        cp_async_wait_all();
        first = false;
    }
    __syncthreads();
}

// grid = chunks of this level owned by this rank: back-substitution of the chunk's interior nodes
__global__ void __launch_bounds__(B200_SOLVE_THREADS) chain_backsub_kernel(Dims dm, Level lv, SolveBuffers sb, int c_first) {
    B200_DYN_SMEM(double, sm);
    if (solve_skipped(sb)) return;
    const int c = c_first + blockIdx.x;
    const int ke = lv.begin[c + 1];
    SweepArgs a;
    level_chunk_args(dm, lv, sb, c, a);
    double* coef = sm + (long long)dm.ntp * dm.ld + 4 * 256;
    load_coef(dm, lv, sb, c, coef);
    __syncthreads();
    const long long rst = (c < lv.P - 1) ? (lv.st ? lv.st[ke - 1] : ke - 1) : -1;
    cta_backsub(dm, a, coef, rst, sb, true, sm);
}

// single CTA, after the top-level sweep + Schur kernels (Level with P == 1): static system, back-substitution of the top nodes
__global__ void __launch_bounds__(B200_SOLVE_THREADS) top_finish_kernel(Dims dm, Level lv, SolveBuffers sb) {
    B200_DYN_SMEM(double, sm);
    const int ns = dm.ns, ns1 = dm.ns1, hmax = dm.hmax;
    if (solve_skipped(sb)) return;
    if (*sb.status != 0) return;
    const double lambda = *sb.lam;
    SweepArgs a;
    level_chunk_args(dm, lv, sb, 0, a);
    // static system: (Sacc[0:ns,0:ns] + lambda I) ds = Sacc[ns][0:ns]; the rhs row rides along as a panel row
    double* coef = sm + (long long)dm.ntp * dm.ld + 4 * 256;
    const int lds = static_lds(ns);
    double* Ss = sm;
    for (int e = threadIdx.x; e < ns1 * ns1; e += blockDim.x) {
        const int r = e / ns1, c = e % ns1;
        if (c <= r && c < ns) Ss[r * lds + c] = a.Sacc[(long long)r * hmax + c] + ((r == c) ? lambda : 0.0);
    }
    __syncthreads();
    if (ns > 0) {
        const int fail = (ns <= 256) ? cta_chol_rl_smem(Ss, lds, ns, ns1, coef) : cta_chol_smem(Ss, lds, ns, ns1);
        if (fail) { if (threadIdx.x == 0) atomicCAS(sb.status, 0, 1 << 30); return; }
        for (int i = threadIdx.x; i < ns; i += blockDim.x) coef[i] = Ss[ns * lds + i];
        __syncthreads();
        cta_solve_LT(Ss, lds, ns, coef);
    }
    if (threadIdx.x == 0) coef[ns] = -1.0;
    __syncthreads();
    for (int i = threadIdx.x; i < ns; i += blockDim.x) sb.dstat[i] = coef[i];
    cta_backsub(dm, a, coef, -1, sb, false, sm);
}

}  // namespace b200d
