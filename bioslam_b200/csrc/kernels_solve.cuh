// Sparse solve (north_star subsystem 3): block-tridiagonal time chain + arrowhead (static variables) Cholesky.
//
//   (H + lambda I) delta = -g,   H = [ D_0  O_0^T            A_0^T ]
//                                    [ O_0  D_1   O_1^T      A_1^T ]
//                                    [        ...            ...   ]
//                                    [ A_0  A_1   ...        S     ]
//
// Two-level partitioned (SPIKE-style) factorisation so one solve exposes P-way parallelism:
//   1. chain_sweep_kernel, one CTA per chunk: eliminate the chunk's interior keyframes left to right.  The coupling to
//      the chunk's LEFT separator keyframe is carried as extra "arrow" rows, exactly like the static variables, and the
//      right-hand side is carried as one more arrow row (so forward substitution, the Schur complement onto the
//      statics/separators and the reduced right-hand side all come out of the same rank-k update).
//   2. reduced_assemble_kernel: gather the chunks' Schur complements onto the separator keyframes.
//   3. top_solve_kernel, one CTA: the same sweep over the P-1 separators, the static system, back-substitution of the
//      separators.
//   4. chain_backsub_kernel, one CTA per chunk: back-substitution of the interior keyframes.
// Per keyframe the work is: in-shared-memory blocked Cholesky of the (ntp x ntp) pivot, a blocked triangular solve of
// the panel [O_k ; arrow rows] against it, and a symmetric rank-ntp update (panel * panel^T) that yields the fill for
// the next keyframe and the arrow Schur complement.  All three run on a register-tiled fp64 micro-kernel (4x4 per
// thread, operands in shared memory read as 16-byte pairs).
#pragma once
#include "kernels_assemble.cuh"

namespace b200d {

#define B200_NB 16        // Cholesky / TRSM block width
#define B200_SOLVE_THREADS 512

// optional per-phase cycle counters of the sweep (debug builds: -DB200_PHASE_CLOCKS), accumulated by thread 0 of block 0
#ifdef B200_PHASE_CLOCKS
__device__ unsigned long long g_phase_clk[16];
#define PH_T0() unsigned long long ph_t_ = clock64()
#define PH_ADD(i) do { if (threadIdx.x == 0 && blockIdx.x == gridDim.x - 1) { unsigned long long n_ = clock64(); g_phase_clk[i] += n_ - ph_t_; ph_t_ = n_; } } while (0)
#else
#define PH_T0()
#define PH_ADD(i)
#endif

// ------------------------------------------------------------------------------------------------- micro-kernel
// out(m, n, sum_{k in [k0,k1)} A[m][k]*B[n][k]) for m<M, n<N.  A, B in shared memory, lda/ldb even, k0 even, (k1-k0) even.
// Warp tile 32 x 16: lane -> (ty = lane>>2, tx = lane&3), thread rows m0+ty+8i, cols n0+tx+4j.
template <class Out>
__device__ __forceinline__ void cta_gemm_nt(const double* __restrict__ A, int lda, int M, const double* __restrict__ B, int ldb, int N,
                                            int k0, int k1, bool lower_only, Out out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int tx = lane & 3, ty = lane >> 2;
    const int mt = (M + 31) >> 5, ntl = (N + 15) >> 4;
    for (int tile = warp; tile < mt * ntl; tile += nwarps) {
        const int m0 = (tile / ntl) << 5, n0 = (tile % ntl) << 4;
        if (lower_only && n0 > m0 + 31) continue;
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
            for (int i = 0; i < 4; i++) {
                a[i] = *reinterpret_cast<const double2*>(ap[i] + k);
                b[i] = *reinterpret_cast<const double2*>(bp[i] + k);
            }
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

// ------------------------------------------------------------------------------------------------- 16 x 16 building blocks
// Diagonal block factorisation by ONE warp, entirely in registers: lane l owns row l of the jb x jb block (jb <= 16, padded
// with identity), pivots / column entries travel by warp shuffle.  Writes L back in place, its transpose into Tdiag
// (Tdiag[c*16+c2] = L[c2][c]) and the reciprocal diagonal into invd.  Returns 0 or 1 + index of a non-positive pivot.
#define B200_WLD 18  // leading dimension of the 16 x 16 inverse diagonal block in shared memory (== 2 mod 4)
__device__ __forceinline__ int warp_potrf16(double* Db, int ld, int jb, double* Tdiag, double* invd, double* Wd) {
    const int l = threadIdx.x & 31;
    double r[B200_NB];
#pragma unroll
    for (int c = 0; c < B200_NB; c++) r[c] = (l < jb && c < jb && c <= l) ? Db[(long long)l * ld + c] : ((l == c) ? 1.0 : 0.0);
    int fail = 0;
#pragma unroll
    for (int c = 0; c < B200_NB; c++) {
        double d = __shfl_sync(0xffffffffu, r[c], c);
        if (!(d > 0.0) || !(d < 1.0e300)) { if (!fail) fail = 1 + c; d = 1.0; }
        const double inv = rsqrt(d);
        const double x = r[c] * inv;   // lane c: sqrt(d) ; lanes > c: L[l][c]
        r[c] = x;
        if (l == c) invd[c] = inv;
#pragma unroll
        for (int c2 = c + 1; c2 < B200_NB; c2++) {
            const double lc2 = __shfl_sync(0xffffffffu, x, c2);
            r[c2] = fma(-x, lc2, r[c2]);
        }
    }
    if (l < B200_NB) {
#pragma unroll
        for (int c = 0; c < B200_NB; c++) {
            if (c <= l) {
                if (l < jb && c < jb) Db[(long long)l * ld + c] = r[c];
                Tdiag[c * B200_NB + l] = r[c];
            }
        }
    }
    __syncwarp();
    // W = L^-1: lane c solves L w = e_c by right-looking forward substitution (L entries broadcast from Tdiag)
    if (l < B200_NB) {
        double w[B200_NB];
#pragma unroll
        for (int i = 0; i < B200_NB; i++) w[i] = (i == l) ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < B200_NB; i++) {
            w[i] *= invd[i];
#pragma unroll
            for (int i2 = i + 1; i2 < B200_NB; i2++) w[i2] = fma(-Tdiag[i * B200_NB + i2], w[i], w[i2]);
        }
#pragma unroll
        for (int i = 0; i < B200_NB; i++) {
            Wd[i * B200_WLD + l] = (i >= l) ? w[i] : 0.0;
            // stash the strictly lower part of W in the (otherwise unused) upper triangle of the diagonal block of L
            if (i > l && i < jb) Db[(long long)l * ld + i] = w[i];
        }
    }
    return fail;
}

// rows x 16 block X (shared, ld) <- X * W^T  (== X * Ljj^-T) on the tensor-less fp64 micro-kernel; jb even.
__device__ __forceinline__ void cta_rows_times_WT(double* X, int ld, int nrows, int jb, const double* Wd) {
    cta_gemm_nt(X, ld, nrows, Wd, B200_WLD, jb, 0, jb, false, [&](int m, int c, double v) { X[(long long)m * ld + c] = v; });
}

// x (one row, 16 entries at xr) <- x * Ljj^-T, right-looking in registers.  Tdiag / invd as written by warp_potrf16
// (or rebuilt from a slab of L).  jb < 16: only the first jb entries exist.
__device__ __forceinline__ void row_trisolve16(double* xr, int jb, const double* Tdiag, const double* invd) {
    double x[B200_NB];
    if (jb == B200_NB) {
#pragma unroll
        for (int c = 0; c < B200_NB; c += 2) {
            const double2 v = *reinterpret_cast<const double2*>(xr + c);
            x[c] = v.x; x[c + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int c = 0; c < B200_NB; c++) x[c] = (c < jb) ? xr[c] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < B200_NB; c++) {
        if (c < jb) {
            x[c] *= invd[c];
#pragma unroll
            for (int c2 = c + 1; c2 < B200_NB; c2++)
                if (c2 < jb) x[c2] = fma(-x[c], Tdiag[c * B200_NB + c2], x[c2]);
        }
    }
    if (jb == B200_NB) {
#pragma unroll
        for (int c = 0; c < B200_NB; c += 2) *reinterpret_cast<double2*>(xr + c) = make_double2(x[c], x[c + 1]);
    } else {
#pragma unroll
        for (int c = 0; c < B200_NB; c++) if (c < jb) xr[c] = x[c];
    }
}

// small static shared scratch shared by the Cholesky and the triangular solve (one instance per kernel)
__device__ __forceinline__ void solve_scratch(double*& Tdiag, double*& Wd, double*& invd) {
    __shared__ __align__(16) double s_Tdiag[B200_NB * B200_NB];
    __shared__ __align__(16) double s_Wd[B200_NB * B200_WLD];
    __shared__ double s_invd[B200_NB];
    Tdiag = s_Tdiag; Wd = s_Wd; invd = s_invd;
}

// ------------------------------------------------------------------------------------------------- in-smem Cholesky
// Lower Cholesky of the leading n x n block of the (nrows x n) panel in shared memory (row-major, ld), blocked
// left-looking; rows [n, nrows) are carried along (they end up multiplied by L^-T, i.e. forward-substituted).
// Returns (to all threads) 0 or 1 + index of the first non-positive pivot.  Must be called by all threads of the CTA.
__device__ int cta_chol_smem(double* Ms, int ld, int n, int nrows) {
    __shared__ int fail;
    double *Tdiag, *Wd, *invd;
    solve_scratch(Tdiag, Wd, invd);
    if (threadIdx.x == 0) fail = 0;
    __syncthreads();
    PH_T0();
    for (int j0 = 0; j0 < n; j0 += B200_NB) {
        const int jb = (n - j0 < B200_NB) ? (n - j0) : B200_NB;
        if (j0 > 0) {
            double* Cb = Ms + (long long)j0 * ld + j0;
            cta_gemm_nt(Ms + (long long)j0 * ld, ld, nrows - j0, Ms + (long long)j0 * ld, ld, jb, 0, j0, false,
                        [&](int m, int c, double v) { Cb[(long long)m * ld + c] -= v; });
            __syncthreads();
        }
        PH_ADD(8);
        if (threadIdx.x < 32) {
            const int f = warp_potrf16(Ms + (long long)j0 * ld + j0, ld, jb, Tdiag, invd, Wd);
            if (f && threadIdx.x == 0 && fail == 0) fail = j0 + f;
        }
        PH_ADD(9);
        __syncthreads();
        if ((jb & 1) == 0) {
            if (nrows > j0 + jb) cta_rows_times_WT(Ms + (long long)(j0 + jb) * ld + j0, ld, nrows - j0 - jb, jb, Wd);
        } else {
            for (int row = j0 + jb + threadIdx.x; row < nrows; row += blockDim.x)
                row_trisolve16(Ms + (long long)row * ld + j0, jb, Tdiag, invd);
        }
        __syncthreads();
        PH_ADD(10);
    }
    return fail;
}

// copy rows [r0, r0+nr) x cols [0, w) (w even) of a global row-major matrix (ldg even) into shared rows (ld even), async
__device__ __forceinline__ void cp_rows_async(double* dst, int ld, const double* __restrict__ src, int ldg, int nr, int w) {
    const int w2 = (w + 1) >> 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < nr; r += nwarps)
        for (int c2 = lane; c2 < w2; c2 += 32) cp_async16(dst + (long long)r * ld + 2 * c2, src + (long long)r * ldg + 2 * c2);
}

// rows (cr x n, shared, ld) <- rows * L^-T where L (n x n lower, upper part zero) is streamed block-row by block-row from
// global memory (ldg) through two shared slabs (NB x ld each), the next slab being fetched (cp.async) while the current
// block column is updated and solved.
__device__ void cta_trsm_rows(double* Ps, int ld, int cr, int n, const double* __restrict__ Lg, int ldg, double* slab0, double* slab1) {
    double *Tdiag, *Wd, *invd;
    solve_scratch(Tdiag, Wd, invd);
    {
        const int jb0 = (n < B200_NB) ? n : B200_NB;
        cp_rows_async(slab0, ld, Lg, ldg, jb0, jb0 + (jb0 & 1));
    }
    int blk = 0;
    PH_T0();
    for (int j0 = 0; j0 < n; j0 += B200_NB, blk++) {
        const int jb = (n - j0 < B200_NB) ? (n - j0) : B200_NB;
        double* Ls = (blk & 1) ? slab1 : slab0;
        double* Ln = (blk & 1) ? slab0 : slab1;
        cp_async_wait_all();
        __syncthreads();
        PH_ADD(11);
        if (j0 + jb < n) {
            const int j1 = j0 + jb;
            const int jb1 = (n - j1 < B200_NB) ? (n - j1) : B200_NB;
            const int w1 = j1 + jb1;
            cp_rows_async(Ln, ld, Lg + (long long)j1 * ldg, ldg, jb1, w1 + (w1 & 1));
        }
        if (threadIdx.x < B200_NB * B200_NB) {
            const int c = threadIdx.x / B200_NB, c2 = threadIdx.x % B200_NB;
            if (c2 >= c && c2 < jb) Tdiag[c * B200_NB + c2] = Ls[(long long)c2 * ld + j0 + c];
            if (c2 == c && c < jb) invd[c] = 1.0 / Ls[(long long)c * ld + j0 + c];
            // W (inverse of the diagonal block): strictly lower part stashed above the diagonal of L, diagonal = 1/L_cc
            double wv = 0.0;
            if (c < jb && c2 < jb) {
                if (c2 == c) wv = 1.0 / Ls[(long long)c * ld + j0 + c];
                else if (c2 < c) wv = Ls[(long long)c2 * ld + j0 + c];
            } else if (c == c2) wv = 1.0;
            Wd[c * B200_WLD + c2] = wv;
        }
        if (j0 > 0) {
            double* Cb = Ps + j0;
            cta_gemm_nt(Ps, ld, cr, Ls, ld, jb, 0, j0, false, [&](int m, int c, double v) { Cb[(long long)m * ld + c] -= v; });
        }
        __syncthreads();
        PH_ADD(12);
        if ((jb & 1) == 0) cta_rows_times_WT(Ps + j0, ld, cr, jb, Wd);
        else for (int row = threadIdx.x; row < cr; row += blockDim.x) row_trisolve16(Ps + (long long)row * ld + j0, jb, Tdiag, invd);
        PH_ADD(13);
    }
    __syncthreads();
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
    double* Lst;         // [storage idx][ntp*ntp]
    double* Pst;         // [storage idx][(nc+hmax)*ntp]  rows [0,nc): X, rows [nc, nc+h): arrow
    double* F;           // 2 x [(nc+hmax)*nc] ping-pong fill buffers of this sweep
    double* Sacc;        // [hmax*hmax] (lower triangle) Schur complement accumulated by this sweep
    int* status;
    int rmax;            // rows of the panel resident in shared memory at a time
};

// shared-memory carve-up (doubles): phase A uses [0, ntp*ld) ; phase B: Ps [rmax*ld] | Ls [NB*ld] | Bs [32*ld]
__host__ __device__ inline long long sweep_smem_doubles(int ntp, int ld, int rmax) {
    long long a = (long long)ntp * ld + 256 + 256 + 1024 + 512, b = (long long)(rmax + B200_NB + 32) * ld;
    return a > b ? a : b;
}
__host__ __device__ inline int static_lds(int ns) { int lds = ns + (ns & 1); while ((lds & 3) != 2) lds += 2; return lds; }

__device__ void cta_sweep(const Dims& dm, const SweepArgs& a, double* sm) {
    const int ntp = dm.ntp, nl = dm.nl, nc = dm.nc, ns1 = dm.ns1, ld = dm.ld, hmax = dm.hmax;
    const int h = a.h;
    const long long Fsz = (long long)(nc + hmax) * nc;
    for (int i = threadIdx.x; i < hmax * hmax; i += blockDim.x) a.Sacc[i] = 0.0;
    __syncthreads();
    for (int i = 0; i < a.n; i++) {
        const long long src = a.src0 + i;
        const long long st = a.st_idx ? a.st_idx[i] : src;
        const bool has_x = (i + 1 < a.n) || a.has_right;
        const double* Fin = a.F + (long long)(i & 1) * Fsz;
        double* Fout = a.F + (long long)((i + 1) & 1) * Fsz;
        double* Lg = a.Lst + st * ntp * ntp;
        double* Pg = a.Pst + st * (long long)(nc + hmax) * ntp;
        const double* Og = a.Osrc + src * nc * nc;
        const double* Ag = a.Asrc + src * ns1 * ntp;
        // ---------------- phase A: pivot block  Dacc = D + lambda I - fill
        PH_T0();
        {
            const double* Dg = a.Dsrc + src * ntp * ntp;
            cp_rows_async(sm, ld, Dg, ntp, ntp, ntp);
            if (i + 1 < a.n) {  // pull the next keyframe's blocks towards L2 while this one is factorised
                const double* Dn = a.Dsrc + (src + 1) * ntp * ntp;
                for (int e = threadIdx.x * 16; e < ntp * ntp; e += blockDim.x * 16) prefetch_l2(Dn + e);
            }
            cp_async_wait_all();
            __syncthreads();
            for (int r = threadIdx.x; r < ntp; r += blockDim.x) sm[r * ld + r] += a.lambda;
            if (i > 0) {
                const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
                for (int r = warp; r < nc; r += nwarps) {
                    const double* fr = Fin + (long long)r * nc;
                    double* dr = sm + (nl + r) * ld + nl;
#pragma unroll 4
                    for (int c = lane; c < nc; c += 32) dr[c] -= fr[c];
                }
            }
            __syncthreads();
            PH_ADD(0);
            const int fail = cta_chol_smem(sm, ld, ntp, ntp);
            PH_ADD(1);
            if (fail) {
                if (threadIdx.x == 0) atomicCAS(a.status, 0, (int)(st + 1));
                return;
            }
            {
                const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
                for (int r = warp; r < ntp; r += nwarps)
                    for (int c = 2 * lane; c < ntp; c += 64) {
                        const double2 v = *reinterpret_cast<const double2*>(sm + r * ld + c);
                        const bool diag_blk = (c / B200_NB) == (r / B200_NB);   // (c even: c and c+1 share the block)
                        *reinterpret_cast<double2*>(Lg + (long long)r * ntp + c) =
                            make_double2((c <= r || diag_blk) ? v.x : 0.0, (c + 1 <= r || diag_blk) ? v.y : 0.0);
                    }
            }
            __threadfence_block();
            __syncthreads();
            PH_ADD(2);
        }
        // ---------------- phase B: panel rows [X (nc, if has_x) ; arrow (h)]
        double* Ps = sm;
        double* Ls = sm + (long long)a.rmax * ld;
        double* Bs = Ls + (long long)B200_NB * ld;
        const int xo = has_x ? nc : 0;
        const int Rn = xo + h;
        const double* Sg = a.Ssrc + src * ns1 * ns1;   // this node's own static block, folded into the Schur update below
        const int nchunks = (Rn + a.rmax - 1) / a.rmax;
        const int csz = (Rn + nchunks - 1) / nchunks;
        for (int ch = 0; ch < nchunks; ch++) {
            const int r0 = ch * csz;
            const int cr = (Rn - r0 < csz) ? (Rn - r0) : csz;
            __syncthreads();
            // load rows (one warp per row): sources through cp.async, then the fill correction in shared memory
            {
                const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
                for (int lr = warp; lr < cr; lr += nwarps) {
                    const int pr = r0 + lr;
                    double* dst = Ps + lr * ld;
                    if (pr < xo) {
                        const double* og = Og + (long long)pr * nc - nl;
                        for (int c = lane; c < ntp; c += 32) {
                            if (c >= nl) cp_async8(dst + c, og + c);
                            else dst[c] = 0.0;
                        }
                    } else {
                        const int ar = pr - xo;
                        if (ar < ns1) {
                            const double* ag = Ag + (long long)ar * ntp;
                            for (int c = 2 * lane; c < ntp; c += 64) cp_async16(dst + c, ag + c);
                        } else if (i == 0 && a.Oleft) {
                            const double* ol = a.Oleft + (ar - ns1);
                            for (int c = lane; c < ntp; c += 32) dst[c] = (c >= nl) ? ol[(long long)(c - nl) * nc] : 0.0;
                        } else {
                            for (int c = lane; c < ntp; c += 32) dst[c] = 0.0;
                        }
                    }
                }
                cp_async_wait_all();
                __syncthreads();
                if (i > 0) {
                    // arrow rows of this chunk: subtract Y_prev X_prev^T (chain columns)
                    for (int lr = warp; lr < cr; lr += nwarps) {
                        const int pr = r0 + lr;
                        if (pr < xo) continue;
                        const double* fr = Fin + (long long)(nc + pr - xo) * nc;
                        double* dr = Ps + lr * ld + nl;
#pragma unroll 4
                        for (int c = lane; c < nc; c += 32) dr[c] -= fr[c];
                    }
                }
            }
            // (cta_trsm_rows starts with a barrier)
            __syncthreads();
            PH_ADD(3);
            cta_trsm_rows(Ps, ld, cr, ntp, Lg, ntp, Ls, Bs);
            PH_ADD(4);
            // store
            {
                const int sro_ = has_x ? 0 : nc;
                const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
                for (int lr = warp; lr < cr; lr += nwarps)
                    for (int c = 2 * lane; c < ntp; c += 64)
                        *reinterpret_cast<double2*>(Pg + (long long)(sro_ + r0 + lr) * ntp + c) = *reinterpret_cast<const double2*>(Ps + lr * ld + c);
            }
            __syncthreads();
            PH_ADD(5);
            // rank update, diagonal chunk pair
            auto emit = [&](int sr1, int sr2, double v) {
                // sr1 >= sr2 (storage rows)
                if (sr1 < nc) {
                    Fout[(long long)sr1 * nc + sr2] = v;
                    if (sr1 != sr2) Fout[(long long)sr2 * nc + sr1] = v;
                } else if (sr2 < nc) {
                    Fout[(long long)sr1 * nc + sr2] = v;
                } else {
                    // exactly one reduction per entry and step (own static block - Y Y^T): deterministic, never mixed with
                    // plain loads of Sacc inside the sweep (atomics resolve in L2)
                    const int a1 = sr1 - nc, a2 = sr2 - nc;
                    double add = -v;
                    if (a1 < ns1) add += Sg[a1 * ns1 + a2];
                    atomicAdd(a.Sacc + (long long)a1 * hmax + a2, add);
                }
            };
            const int sro = has_x ? 0 : nc;  // storage row of present row 0
            cta_gemm_nt(Ps, ld, cr, Ps, ld, cr, 0, ntp, true, [&](int m, int n2, double v) {
                if (n2 <= m) emit(sro + r0 + m, sro + r0 + n2, v);
            });
            // cross terms with earlier chunks, streamed back from global memory 32 rows at a time
            for (int p0 = 0; p0 < r0; p0 += 32) {
                const int pb = (r0 - p0 < 32) ? (r0 - p0) : 32;
                __syncthreads();
                cp_rows_async(Bs, ld, Pg + (long long)(sro + p0) * ntp, ntp, pb, ntp);
                cp_async_wait_all();
                __syncthreads();
                cta_gemm_nt(Ps, ld, cr, Bs, ld, pb, 0, ntp, false,
                            [&](int m, int n2, double v) { emit(sro + r0 + m, sro + p0 + n2, v); });
            }
        }
        __threadfence_block();
        __syncthreads();
        PH_ADD(6);
    }
}

// back-substitution over the nodes of a sweep (reverse order).  coefA = [delta_static (ns), -1, delta_left_chain (nc)].
// delta (global) [storage idx][ntp].  right_st: storage index of the right neighbour (or -1).
__device__ void cta_backsub(const Dims& dm, const SweepArgs& a, const double* coefA, long long right_st, double* delta, double* sm) {
    const int ntp = dm.ntp, nl = dm.nl, nc = dm.nc, ld = dm.ld, hmax = dm.hmax;
    const int h = a.h;
    double* Ms = sm;                 // L in shared memory, then scratch vectors (ntp, nc <= 256)
    double* tvec = sm + (long long)ntp * ld;
    double* cx = tvec + 256;
    double* tpart = cx + 256;        // blockDim.x doubles
    for (int i = a.n - 1; i >= 0; i--) {
        const long long src = a.src0 + i;
        const long long st = a.st_idx ? a.st_idx[i] : src;
        const bool has_x = (i + 1 < a.n) || a.has_right;
        const double* Lg = a.Lst + st * ntp * ntp;
        const double* Pg = a.Pst + st * (long long)(nc + hmax) * ntp;
        long long nxt = -1;
        if (i + 1 < a.n) nxt = a.st_idx ? a.st_idx[i + 1] : src + 1;
        else if (a.has_right) nxt = right_st;
        __syncthreads();
        cp_rows_async(Ms, ld, Lg, ntp, ntp, ntp);
        if (has_x) for (int j = threadIdx.x; j < nc; j += blockDim.x) cx[j] = delta[nxt * ntp + nl + j];
        __syncthreads();
        // t[c] = - sum_r P[r][c] * coef[r]; thread groups of 128 split the rows, column passes of 128
        {
            const int grp = threadIdx.x >> 7, ngrp = blockDim.x >> 7;
            for (int cb = 0; cb < ntp; cb += 128) {
                const int c = cb + (threadIdx.x & 127);
                double t = 0.0;
                if (c < ntp) {
                    if (has_x) {
#pragma unroll 4
                        for (int r = grp; r < nc; r += ngrp) t = fma(-Pg[(long long)r * ntp + c], cx[r], t);
                    }
#pragma unroll 4
                    for (int r = grp; r < h; r += ngrp) t = fma(-Pg[(long long)(nc + r) * ntp + c], coefA[r], t);
                }
                tpart[threadIdx.x] = t;
                __syncthreads();
                if (grp == 0 && c < ntp) {
                    double acc = 0.0;
                    for (int q = 0; q < ngrp; q++) acc += tpart[q * 128 + (threadIdx.x & 127)];
                    tvec[c] = acc;
                }
                __syncthreads();
            }
        }
        cp_async_wait_all();
        __syncthreads();
        cta_solve_LT(Ms, ld, ntp, tvec);
        for (int c = threadIdx.x; c < ntp; c += blockDim.x) delta[st * ntp + c] = tvec[c];
        __threadfence_block();
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------- kernels
// One level of the nested partition.  Level 0: nodes = keyframes (sources = assembled blocks, storage index = node index).
// Level l+1: nodes = the separators of level l (sources = the reduced blocks written by level_reduce_kernel, storage index =
// the separator's keyframe).  The nodes of a level are cut into P chunks; the last node of every chunk but the last is that
// chunk's separator and becomes a node of the next level.
struct Level {
    const double *Dsrc, *Osrc, *Asrc, *Ssrc;  // per node: [ntp*ntp], [nc*nc], [ns1*ntp], [ns1*ns1]
    const int* st;                            // storage (keyframe) index per node, nullptr = identity
    const int* begin;                         // [P+1] chunk c covers nodes [begin[c], begin[c+1])
    int P;
    double *F;                                // [P][2][(nc+hmax)*nc] fill buffers (ping-pong) of each chunk sweep
    double *S;                                // [P][hmax*hmax] arrow Schur complement of each chunk sweep
    double *Dred, *Ored, *Ared, *Sred;        // [P-1] reduced sources == next level's node sources
    int* sep_st;                              // [P-1] storage index of separator c == next level's st
};

struct SolveBuffers {
    double *Lst, *Pst;                        // factors, per storage index (keyframe)
    double *Ftop, *Stop;                      // scratch of the top-level sweep
    double *delta, *dstat;                    // solution (internal order), delta per storage index
    double* scal;                             // scalar outputs
    int* status;
};

__device__ __forceinline__ void level_chunk_args(const Dims& dm, const Level& lv, const SolveBuffers& sb, int c, SweepArgs& a) {
    const int kb = lv.begin[c], ke = lv.begin[c + 1];
    a.Dsrc = lv.Dsrc; a.Osrc = lv.Osrc; a.Asrc = lv.Asrc; a.Ssrc = lv.Ssrc;
    a.src0 = kb;
    a.n = (c < lv.P - 1) ? (ke - 1 - kb) : (ke - kb);   // the last node of every chunk but the last is a separator
    a.st_idx = lv.st ? lv.st + kb : nullptr;
    a.has_right = (c < lv.P - 1);
    a.Oleft = (c > 0) ? lv.Osrc + (long long)(kb - 1) * dm.nc * dm.nc : nullptr;
    a.h = dm.ns1 + ((c > 0) ? dm.nc : 0);
    a.Lst = sb.Lst; a.Pst = sb.Pst;
    a.F = lv.F + (long long)c * 2 * (dm.nc + dm.hmax) * dm.nc;
    a.Sacc = lv.S + (long long)c * dm.hmax * dm.hmax;
    a.status = sb.status;
}

// grid = number of chunks of this level owned by this rank (chunk = c_first + blockIdx.x)
__global__ void __launch_bounds__(B200_SOLVE_THREADS) level_sweep_kernel(Dims dm, Level lv, SolveBuffers sb, double lambda, int rmax, int c_first) {
    B200_DYN_SMEM(double, sm);
    SweepArgs a;
    level_chunk_args(dm, lv, sb, c_first + blockIdx.x, a);
    a.lambda = lambda;
    a.rmax = rmax;
    cta_sweep(dm, a, sm);
}

// Schur complements onto separator c of this level: from chunk c (its right side) and chunk c+1 (its left side).
// grid = separators handled (separator = c_first + blockIdx.x)
__global__ void __launch_bounds__(256) level_reduce_kernel(Dims dm, Level lv, int c_first) {
    const int c = c_first + blockIdx.x;
    const int P = lv.P;
    const int ntp = dm.ntp, nl = dm.nl, nc = dm.nc, ns1 = dm.ns1, hmax = dm.hmax;
    const long long q = lv.begin[c + 1] - 1;                 // node index of the separator
    const long long Fsz = (long long)(nc + hmax) * nc;
    const int n_c = (int)(q - lv.begin[c]);                  // interior nodes of chunk c
    const double* Fc = lv.F + ((long long)c * 2 + (n_c & 1)) * Fsz;   // written by its last node
    const double* Sn = lv.S + (long long)(c + 1) * hmax * hmax;        // chunk c+1 (has left rows)
    if (threadIdx.x == 0) lv.sep_st[c] = lv.st ? lv.st[q] : (int)q;
    double* Dr = lv.Dred + (long long)c * ntp * ntp;
    const double* Dg = lv.Dsrc + q * ntp * ntp;
    for (int e = threadIdx.x; e < ntp * ntp; e += blockDim.x) {
        const int r = e / ntp, col = e % ntp;
        double v = Dg[e];
        if (r >= nl && col >= nl) {
            const int i = r - nl, j = col - nl;
            if (n_c > 0) v -= Fc[(long long)i * nc + j];
            const int hi = i > j ? i : j, lo = i > j ? j : i;
            v += Sn[(long long)(ns1 + hi) * hmax + (ns1 + lo)];
        }
        Dr[e] = v;
    }
    double* Ar = lv.Ared + (long long)c * ns1 * ntp;
    const double* Ag = lv.Asrc + q * ns1 * ntp;
    for (int e = threadIdx.x; e < ns1 * ntp; e += blockDim.x) {
        const int ar = e / ntp, col = e % ntp;
        double v = Ag[e];
        if (col >= nl) {
            if (n_c > 0) v -= Fc[(long long)(nc + ar) * nc + (col - nl)];
            v += Sn[(long long)(ns1 + (col - nl)) * hmax + ar];
        }
        Ar[e] = v;
    }
    double* Sr = lv.Sred + (long long)c * ns1 * ns1;
    const double* Sg = lv.Ssrc + q * ns1 * ns1;
    const double* S0 = lv.S;
    for (int e = threadIdx.x; e < ns1 * ns1; e += blockDim.x) {
        const int r = e / ns1, col = e % ns1;
        const int hi = r > col ? r : col, lo = r > col ? col : r;
        double v = Sg[e] + Sn[(long long)hi * hmax + lo];
        if (c == 0) v += S0[(long long)hi * hmax + lo];
        Sr[e] = v;
    }
    if (c < P - 2) {
        // coupling separator c -> separator c+1 through chunk c+1: -(Y_left X^T) of that chunk's last node
        const int n_n = (int)(lv.begin[c + 2] - 1 - lv.begin[c + 1]);
        const double* Fn = lv.F + ((long long)(c + 1) * 2 + (n_n & 1)) * Fsz;
        double* Or = lv.Ored + (long long)c * nc * nc;
        for (int e = threadIdx.x; e < nc * nc; e += blockDim.x) {
            const int j = e / nc, i = e % nc;   // Or[j][i]: row = chain j of separator c+1, col = chain i of separator c
            Or[e] = -Fn[(long long)(nc + ns1 + i) * nc + j];
        }
    }
}

// single CTA: sweep over the n nodes of the top level, static solve, back-substitution of the top-level nodes
__global__ void __launch_bounds__(B200_SOLVE_THREADS) top_solve_kernel(Dims dm, const double* Dsrc, const double* Osrc, const double* Asrc,
                                                                       const double* Ssrc, const int* st, int n, SolveBuffers sb, double lambda,
                                                                       int rmax) {
    B200_DYN_SMEM(double, sm);
    const int ns = dm.ns, ns1 = dm.ns1, hmax = dm.hmax;
    SweepArgs a;
    a.Dsrc = Dsrc; a.Osrc = Osrc; a.Asrc = Asrc; a.Ssrc = Ssrc;
    a.src0 = 0; a.n = n; a.st_idx = st;
    a.has_right = 0; a.Oleft = nullptr; a.h = ns1; a.lambda = lambda;
    a.Lst = sb.Lst; a.Pst = sb.Pst; a.F = sb.Ftop; a.Sacc = sb.Stop; a.status = sb.status; a.rmax = rmax;
    cta_sweep(dm, a, sm);
    __syncthreads();
    if (*sb.status != 0) return;
    // static system: (Sacc[0:ns,0:ns] + lambda I) ds = Sacc[ns][0:ns]; the rhs row rides along as a panel row
    double* coefA = sm + dm.sm_doubles - 512;   // hmax <= 512, lives at the end of the dynamic region
    const int lds = static_lds(ns);
    double* Ss = sm;
    for (int e = threadIdx.x; e < ns1 * ns1; e += blockDim.x) {
        const int r = e / ns1, c = e % ns1;
        if (c <= r && c < ns) Ss[r * lds + c] = __ldcg(a.Sacc + (long long)r * hmax + c) + ((r == c) ? lambda : 0.0);
    }
    __syncthreads();
    if (ns > 0) {
        const int fail = cta_chol_smem(Ss, lds, ns, ns1);
        if (fail) { if (threadIdx.x == 0) atomicCAS(sb.status, 0, 1 << 30); return; }
        for (int i = threadIdx.x; i < ns; i += blockDim.x) coefA[i] = Ss[ns * lds + i];
        __syncthreads();
        cta_solve_LT(Ss, lds, ns, coefA);
    }
    if (threadIdx.x == 0) coefA[ns] = -1.0;
    __syncthreads();
    for (int i = threadIdx.x; i < ns; i += blockDim.x) sb.dstat[i] = coefA[i];
    cta_backsub(dm, a, coefA, -1, sb.delta, sm);
}

// grid = chunks of this level owned by this rank: back-substitution of the chunk's interior nodes
__global__ void __launch_bounds__(B200_SOLVE_THREADS) level_backsub_kernel(Dims dm, Level lv, SolveBuffers sb, int c_first) {
    B200_DYN_SMEM(double, sm);
    const int c = c_first + blockIdx.x;
    const int kb = lv.begin[c], ke = lv.begin[c + 1];
    const int ns = dm.ns, ns1 = dm.ns1, nc = dm.nc, nl = dm.nl, ntp = dm.ntp;
    SweepArgs a;
    level_chunk_args(dm, lv, sb, c, a);
    double* coefA = sm + dm.sm_doubles - 512;
    for (int i = threadIdx.x; i < ns; i += blockDim.x) coefA[i] = sb.dstat[i];
    if (threadIdx.x == 0) coefA[ns] = -1.0;
    if (c > 0) {
        const long long lst = lv.st ? lv.st[kb - 1] : kb - 1;
        for (int j = threadIdx.x; j < nc; j += blockDim.x) coefA[ns1 + j] = sb.delta[lst * ntp + nl + j];
    }
    __syncthreads();
    const long long rst = (c < lv.P - 1) ? (lv.st ? lv.st[ke - 1] : ke - 1) : -1;
    cta_backsub(dm, a, coefA, rst, sb.delta, sm);
}

}  // namespace b200d
