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
#define B200_SOLVE_THREADS 256

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

// ------------------------------------------------------------------------------------------------- in-smem Cholesky
// Lower Cholesky of the n x n matrix in shared memory (row-major, ld), blocked left-looking.  Returns (to all threads)
// 0 on success or 1 + index of the first non-positive pivot.  Must be called by all threads of the CTA.
__device__ int cta_chol_smem(double* Ms, int ld, int n) {
    __shared__ int fail;
    if (threadIdx.x == 0) fail = 0;
    __syncthreads();
    for (int j0 = 0; j0 < n; j0 += B200_NB) {
        const int jb = (n - j0 < B200_NB) ? (n - j0) : B200_NB;
        if (j0 > 0) {
            double* Cb = Ms + (long long)j0 * ld + j0;
            cta_gemm_nt(Ms + (long long)j0 * ld, ld, n - j0, Ms + (long long)j0 * ld, ld, jb, 0, j0, false,
                        [&](int m, int c, double v) { Cb[(long long)m * ld + c] -= v; });
        }
        __syncthreads();
        if (threadIdx.x < 32) {  // diagonal block: right-looking, lane = row
            const int l = threadIdx.x;
            double* Db = Ms + (long long)j0 * ld + j0;
            for (int c = 0; c < jb; c++) {
                if (l == c) {
                    const double d = Db[c * ld + c];
                    if (!(d > 0.0) || !(d < 1.0e300)) { if (fail == 0) fail = 1 + j0 + c; Db[c * ld + c] = 1.0; }
                    else Db[c * ld + c] = sqrt(d);
                }
                __syncwarp();
                if (l > c && l < jb) Db[l * ld + c] /= Db[c * ld + c];
                __syncwarp();
                if (l > c && l < jb) {
                    const double lc = Db[l * ld + c];
                    for (int c2 = c + 1; c2 <= l; c2++) Db[l * ld + c2] -= lc * Db[c2 * ld + c];
                }
                __syncwarp();
            }
        }
        __syncthreads();
        // rows below the diagonal block: x <- x * Ljj^-T (thread per row)
        for (int row = j0 + jb + threadIdx.x; row < n; row += blockDim.x) {
            double* xr = Ms + (long long)row * ld + j0;
            const double* Db = Ms + (long long)j0 * ld + j0;
            double x[B200_NB];
            for (int c = 0; c < jb; c++) {
                double v = xr[c];
                for (int l2 = 0; l2 < c; l2++) v -= x[l2] * Db[c * ld + l2];
                x[c] = v / Db[c * ld + c];
            }
            for (int c = 0; c < jb; c++) xr[c] = x[c];
        }
        __syncthreads();
    }
    return fail;
}

// rows (cr x n, shared, ld) <- rows * L^-T where L (n x n lower) is read block-row by block-row from global memory
// (ldg) through the shared slab Ls (NB x ld).
__device__ void cta_trsm_rows(double* Ps, int ld, int cr, int n, const double* __restrict__ Lg, int ldg, double* Ls) {
    for (int j0 = 0; j0 < n; j0 += B200_NB) {
        const int jb = (n - j0 < B200_NB) ? (n - j0) : B200_NB;
        const int w = j0 + jb;
        __syncthreads();
        for (int i = threadIdx.x; i < jb * w; i += blockDim.x) {
            const int r = i / w, c = i % w;
            Ls[r * ld + c] = Lg[(long long)(j0 + r) * ldg + c];
        }
        __syncthreads();
        if (j0 > 0) {
            double* Cb = Ps + j0;
            cta_gemm_nt(Ps, ld, cr, Ls, ld, jb, 0, j0, false, [&](int m, int c, double v) { Cb[(long long)m * ld + c] -= v; });
            __syncthreads();
        }
        for (int row = threadIdx.x; row < cr; row += blockDim.x) {
            double* xr = Ps + (long long)row * ld + j0;
            const double* Db = Ls + j0;
            double x[B200_NB];
            for (int c = 0; c < jb; c++) {
                double v = xr[c];
                for (int l2 = 0; l2 < c; l2++) v -= x[l2] * Db[c * ld + l2];
                x[c] = v / Db[c * ld + c];
            }
            for (int c = 0; c < jb; c++) xr[c] = x[c];
        }
    }
    __syncthreads();
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
    long long a = (long long)ntp * ld, b = (long long)(rmax + B200_NB + 32) * ld;
    return a > b ? a : b;
}

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
        // ---------------- phase A: pivot block
        {
            const double* Dg = a.Dsrc + src * ntp * ntp;
            for (int e = threadIdx.x; e < ntp * ntp; e += blockDim.x) {
                const int r = e / ntp, c = e % ntp;
                double v = Dg[e];
                if (r == c) v += a.lambda;
                if (i > 0 && r >= nl && c >= nl) v -= Fin[(long long)(r - nl) * nc + (c - nl)];
                sm[r * ld + c] = v;
            }
            __syncthreads();
            const int fail = cta_chol_smem(sm, ld, ntp);
            if (fail) {
                if (threadIdx.x == 0) atomicCAS(a.status, 0, (int)(st + 1));
                return;
            }
            for (int e = threadIdx.x; e < ntp * ntp; e += blockDim.x) {
                const int r = e / ntp, c = e % ntp;
                Lg[e] = (c <= r) ? sm[r * ld + c] : 0.0;
            }
            __threadfence_block();
            __syncthreads();
        }
        // ---------------- phase B: panel rows [X (nc, if has_x) ; arrow (h)]
        double* Ps = sm;
        double* Ls = sm + (long long)a.rmax * ld;
        double* Bs = Ls + (long long)B200_NB * ld;
        const int xo = has_x ? nc : 0;
        const int Rn = xo + h;
        // add this node's own static block to the arrow Schur complement (lower triangle)
        {
            const double* Sg = a.Ssrc + src * ns1 * ns1;
            for (int e = threadIdx.x; e < ns1 * ns1; e += blockDim.x) {
                const int r = e / ns1, c = e % ns1;
                if (c <= r) a.Sacc[(long long)r * hmax + c] += Sg[e];
            }
        }
        const int nchunks = (Rn + a.rmax - 1) / a.rmax;
        const int csz = (Rn + nchunks - 1) / nchunks;
        for (int ch = 0; ch < nchunks; ch++) {
            const int r0 = ch * csz;
            const int cr = (Rn - r0 < csz) ? (Rn - r0) : csz;
            __syncthreads();
            // load rows
            const double* Og = a.Osrc + src * nc * nc;
            const double* Ag = a.Asrc + src * ns1 * ntp;
            for (int e = threadIdx.x; e < cr * ntp; e += blockDim.x) {
                const int lr = e / ntp, c = e % ntp;
                const int pr = r0 + lr;
                double v = 0.0;
                if (pr < xo) {
                    if (c >= nl) v = Og[(long long)pr * nc + (c - nl)];
                } else {
                    const int ar = pr - xo;
                    if (ar < ns1) v = Ag[(long long)ar * ntp + c];
                    else if (i == 0 && a.Oleft && c >= nl) v = a.Oleft[(long long)(c - nl) * nc + (ar - ns1)];
                    if (i > 0 && c >= nl) v -= Fin[(long long)(nc + ar) * nc + (c - nl)];
                }
                Ps[lr * ld + c] = v;
            }
            // (cta_trsm_rows starts with a barrier)
            cta_trsm_rows(Ps, ld, cr, ntp, Lg, ntp, Ls);
            // store
            for (int e = threadIdx.x; e < cr * ntp; e += blockDim.x) {
                const int lr = e / ntp, c = e % ntp;
                const int pr = r0 + lr;
                const int sr = has_x ? pr : nc + pr;
                Pg[(long long)sr * ntp + c] = Ps[lr * ld + c];
            }
            // rank update, diagonal chunk pair
            auto emit = [&](int sr1, int sr2, double v) {
                // sr1 >= sr2 (storage rows)
                if (sr1 < nc) {
                    Fout[(long long)sr1 * nc + sr2] = v;
                    if (sr1 != sr2) Fout[(long long)sr2 * nc + sr1] = v;
                } else if (sr2 < nc) {
                    Fout[(long long)sr1 * nc + sr2] = v;
                } else {
                    a.Sacc[(long long)(sr1 - nc) * hmax + (sr2 - nc)] -= v;
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
                for (int e = threadIdx.x; e < pb * ntp; e += blockDim.x) {
                    const int lr = e / ntp, c = e % ntp;
                    Bs[lr * ld + c] = Pg[(long long)(sro + p0 + lr) * ntp + c];
                }
                __syncthreads();
                cta_gemm_nt(Ps, ld, cr, Bs, ld, pb, 0, ntp, false,
                            [&](int m, int n2, double v) { emit(sro + r0 + m, sro + p0 + n2, v); });
            }
        }
        __threadfence_block();
        __syncthreads();
    }
}

// back-substitution over the nodes of a sweep (reverse order).  coefA = [delta_static (ns), -1, delta_left_chain (nc)].
// delta (global) [storage idx][ntp].  right_st: storage index of the right neighbour (or -1).
__device__ void cta_backsub(const Dims& dm, const SweepArgs& a, const double* coefA, long long right_st, double* delta, double* sm) {
    const int ntp = dm.ntp, nl = dm.nl, nc = dm.nc, ld = dm.ld, hmax = dm.hmax;
    const int h = a.h;
    double* Ms = sm;                 // L in shared memory
    __shared__ double tvec[256];     // ntp <= 256
    __shared__ double cx[256];
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
        if (has_x) for (int j = threadIdx.x; j < nc; j += blockDim.x) cx[j] = delta[nxt * ntp + nl + j];
        for (int e = threadIdx.x; e < ntp * ntp; e += blockDim.x) {
            const int r = e / ntp, c = e % ntp;
            if (c <= r) Ms[r * ld + c] = Lg[e];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < ntp; c += blockDim.x) {
            double t = 0.0;
            if (has_x) for (int r = 0; r < nc; r++) t -= Pg[(long long)r * ntp + c] * cx[r];
            for (int r = 0; r < h; r++) t -= Pg[(long long)(nc + r) * ntp + c] * coefA[r];
            tvec[c] = t;
        }
        __syncthreads();
        if (threadIdx.x < 32) {  // L^T x = t, right-looking over rows of L (contiguous in shared memory)
            const int lane = threadIdx.x;
            for (int r = ntp - 1; r >= 0; r--) {
                const double xr = tvec[r] / Ms[r * ld + r];
                __syncwarp();
                if (lane == 0) tvec[r] = xr;
                for (int c = lane; c < r; c += 32) tvec[c] -= Ms[r * ld + c] * xr;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < ntp; c += blockDim.x) delta[st * ntp + c] = tvec[c];
        __threadfence_block();
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------- kernels
struct ChunkPlan {            // device arrays describing the partition
    const int* begin;         // [P+1] chunk c covers keyframes [begin[c], begin[c+1])
    int P;
};

struct SolveBuffers {
    const double *Dsrc, *Osrc, *Asrc, *Ssrc;  // assembled per-keyframe blocks
    double *Lst, *Pst;                        // factors, per keyframe
    double *Fchunk;                           // [P][2][(nc+hmax)*nc]
    double *Schunk;                           // [P][hmax*hmax]
    double *Dred, *Ored, *Ared, *Sred;        // reduced (separator) system sources, [P-1] each
    double *Ftop, *Stop;                      // scratch of the top-level sweep
    int* sep_idx;                             // [P-1] separator keyframe of chunk c
    double *delta, *dstat;                    // solution (internal order)
    double* scal;                             // scalar outputs
    int* status;
};

// grid = number of chunks owned by this rank (chunk = c_first + blockIdx.x)
__global__ void __launch_bounds__(B200_SOLVE_THREADS) chain_sweep_kernel(Dims dm, ChunkPlan pl, SolveBuffers sb, double lambda, int rmax,
                                                                         int c_first) {
    B200_DYN_SMEM(double, sm);
    const int c = c_first + blockIdx.x;
    const int P = pl.P;
    const int kb = pl.begin[c], ke = pl.begin[c + 1];
    SweepArgs a;
    a.Dsrc = sb.Dsrc; a.Osrc = sb.Osrc; a.Asrc = sb.Asrc; a.Ssrc = sb.Ssrc;
    a.src0 = kb;
    a.n = (c < P - 1) ? (ke - 1 - kb) : (ke - kb);   // the last keyframe of every chunk but the last is a separator
    a.st_idx = nullptr;
    a.has_right = (c < P - 1);
    a.Oleft = (c > 0) ? sb.Osrc + (long long)(kb - 1) * dm.nc * dm.nc : nullptr;
    a.h = dm.ns1 + ((c > 0) ? dm.nc : 0);
    a.lambda = lambda;
    a.Lst = sb.Lst; a.Pst = sb.Pst;
    a.F = sb.Fchunk + (long long)c * 2 * (dm.nc + dm.hmax) * dm.nc;
    a.Sacc = sb.Schunk + (long long)c * dm.hmax * dm.hmax;
    a.status = sb.status;
    a.rmax = rmax;
    cta_sweep(dm, a, sm);
}

// grid = P-1: Schur complements of chunk c (right side = separator c) and chunk c+1 (left side = separator c)
__global__ void __launch_bounds__(256) reduced_assemble_kernel(Dims dm, ChunkPlan pl, SolveBuffers sb) {
    const int c = blockIdx.x;
    const int P = pl.P;
    const int ntp = dm.ntp, nl = dm.nl, nc = dm.nc, ns1 = dm.ns1, hmax = dm.hmax;
    const long long q = pl.begin[c + 1] - 1;
    const long long Fsz = (long long)(nc + hmax) * nc;
    const int n_c = (int)(q - pl.begin[c]);                 // interior nodes of chunk c
    const double* Fc = sb.Fchunk + ((long long)c * 2 + (n_c & 1)) * Fsz;   // written by its last node
    const double* Sn = sb.Schunk + (long long)(c + 1) * hmax * hmax;        // chunk c+1 (has left rows)
    if (threadIdx.x == 0) sb.sep_idx[c] = (int)q;
    double* Dr = sb.Dred + (long long)c * ntp * ntp;
    const double* Dg = sb.Dsrc + q * ntp * ntp;
    for (int e = threadIdx.x; e < ntp * ntp; e += blockDim.x) {
        const int r = e / ntp, col = e % ntp;
        double v = Dg[e];
        if (r >= nl && col >= nl) {
            const int i = r - nl, j = col - nl;
            v -= Fc[(long long)i * nc + j];
            const int hi = i > j ? i : j, lo = i > j ? j : i;
            v += Sn[(long long)(ns1 + hi) * hmax + (ns1 + lo)];
        }
        Dr[e] = v;
    }
    double* Ar = sb.Ared + (long long)c * ns1 * ntp;
    const double* Ag = sb.Asrc + q * ns1 * ntp;
    for (int e = threadIdx.x; e < ns1 * ntp; e += blockDim.x) {
        const int ar = e / ntp, col = e % ntp;
        double v = Ag[e];
        if (col >= nl) {
            v -= Fc[(long long)(nc + ar) * nc + (col - nl)];
            v += Sn[(long long)(ns1 + (col - nl)) * hmax + ar];
        }
        Ar[e] = v;
    }
    double* Sr = sb.Sred + (long long)c * ns1 * ns1;
    const double* Sg = sb.Ssrc + q * ns1 * ns1;
    const double* S0 = sb.Schunk;
    for (int e = threadIdx.x; e < ns1 * ns1; e += blockDim.x) {
        const int r = e / ns1, col = e % ns1;
        const int hi = r > col ? r : col, lo = r > col ? col : r;
        double v = Sg[e] + Sn[(long long)hi * hmax + lo];
        if (c == 0) v += S0[(long long)hi * hmax + lo];
        Sr[e] = v;
    }
    if (c < P - 2) {
        // coupling separator c -> separator c+1 through chunk c+1: -(Y_left X^T) of that chunk's last node
        const int n_n = (int)(pl.begin[c + 2] - 1 - pl.begin[c + 1]);
        const double* Fn = sb.Fchunk + ((long long)(c + 1) * 2 + (n_n & 1)) * Fsz;
        double* Or = sb.Ored + (long long)c * nc * nc;
        for (int e = threadIdx.x; e < nc * nc; e += blockDim.x) {
            const int j = e / nc, i = e % nc;   // Or[j][i]: row = chain j of separator c+1, col = chain i of separator c
            Or[e] = -Fn[(long long)(nc + ns1 + i) * nc + j];
        }
    }
}

// single CTA: top-level sweep (over separators, or over all keyframes when P == 1), static solve, top-level back-substitution
__global__ void __launch_bounds__(B200_SOLVE_THREADS) top_solve_kernel(Dims dm, ChunkPlan pl, SolveBuffers sb, double lambda, int rmax) {
    B200_DYN_SMEM(double, sm);
    const int P = pl.P;
    const int ns = dm.ns, ns1 = dm.ns1, hmax = dm.hmax, nc = dm.nc;
    SweepArgs a;
    if (P == 1) {
        a.Dsrc = sb.Dsrc; a.Osrc = sb.Osrc; a.Asrc = sb.Asrc; a.Ssrc = sb.Ssrc;
        a.src0 = 0; a.n = dm.K; a.st_idx = nullptr;
    } else {
        a.Dsrc = sb.Dred; a.Osrc = sb.Ored; a.Asrc = sb.Ared; a.Ssrc = sb.Sred;
        a.src0 = 0; a.n = P - 1; a.st_idx = sb.sep_idx;
    }
    a.has_right = 0; a.Oleft = nullptr; a.h = ns1; a.lambda = lambda;
    a.Lst = sb.Lst; a.Pst = sb.Pst; a.F = sb.Ftop; a.Sacc = sb.Stop; a.status = sb.status; a.rmax = rmax;
    cta_sweep(dm, a, sm);
    __syncthreads();
    if (*sb.status != 0) return;
    // static system: (Sacc[0:ns,0:ns] + lambda I) ds = Sacc[ns][0:ns]
    __shared__ double coefA[512];   // hmax <= 512
    const int lds = ns + 2;
    double* Ss = sm;
    for (int e = threadIdx.x; e < ns1 * ns1; e += blockDim.x) {
        const int r = e / ns1, c = e % ns1;
        if (c <= r) Ss[r * lds + c] = a.Sacc[(long long)r * hmax + c] + ((r == c && r < ns) ? lambda : 0.0);
    }
    __syncthreads();
    if (ns > 0) {
        if (threadIdx.x == 0) {
            int fail = 0;
            for (int j = 0; j < ns && !fail; j++) {
                double d = Ss[j * lds + j];
                for (int l = 0; l < j; l++) d -= Ss[j * lds + l] * Ss[j * lds + l];
                if (!(d > 0.0) || !(d < 1.0e300)) { fail = 1; break; }
                const double ljj = sqrt(d);
                Ss[j * lds + j] = ljj;
                for (int i = j + 1; i < ns1; i++) {   // includes the rhs row (i == ns): forward substitution
                    double v = Ss[i * lds + j];
                    for (int l = 0; l < j; l++) v -= Ss[i * lds + l] * Ss[j * lds + l];
                    Ss[i * lds + j] = v / ljj;
                }
            }
            if (fail) atomicCAS(sb.status, 0, 1 << 30);
            else {
                for (int i = ns - 1; i >= 0; i--) {
                    double v = Ss[ns * lds + i];
                    for (int l = i + 1; l < ns; l++) v -= Ss[l * lds + i] * coefA[l];
                    coefA[i] = v / Ss[i * lds + i];
                }
            }
        }
        __syncthreads();
        if (*sb.status != 0) return;
    }
    if (threadIdx.x == 0) coefA[ns] = -1.0;
    __syncthreads();
    for (int i = threadIdx.x; i < ns; i += blockDim.x) sb.dstat[i] = coefA[i];
    (void)nc;
    cta_backsub(dm, a, coefA, -1, sb.delta, sm);
}

// grid = chunks owned by this rank: back-substitution of interior keyframes
__global__ void __launch_bounds__(B200_SOLVE_THREADS) chain_backsub_kernel(Dims dm, ChunkPlan pl, SolveBuffers sb, int c_first) {
    B200_DYN_SMEM(double, sm);
    const int c = c_first + blockIdx.x;
    const int P = pl.P;
    const int kb = pl.begin[c], ke = pl.begin[c + 1];
    const int ns = dm.ns, ns1 = dm.ns1, nc = dm.nc, nl = dm.nl, ntp = dm.ntp;
    SweepArgs a;
    a.src0 = kb;
    a.n = (c < P - 1) ? (ke - 1 - kb) : (ke - kb);
    a.st_idx = nullptr;
    a.has_right = (c < P - 1);
    a.h = ns1 + ((c > 0) ? nc : 0);
    a.Lst = sb.Lst; a.Pst = sb.Pst;
    __shared__ double coefA[512];
    for (int i = threadIdx.x; i < ns; i += blockDim.x) coefA[i] = sb.dstat[i];
    if (threadIdx.x == 0) coefA[ns] = -1.0;
    if (c > 0) for (int j = threadIdx.x; j < nc; j += blockDim.x) coefA[ns1 + j] = sb.delta[(long long)(kb - 1) * ntp + nl + j];
    __syncthreads();
    cta_backsub(dm, a, coefA, (long long)ke - 1, sb.delta, sm);
}

}  // namespace b200d
