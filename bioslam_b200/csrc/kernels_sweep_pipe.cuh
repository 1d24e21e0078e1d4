// Level-0 chain sweep, software-pipelined ACROSS keyframes (the kernel VERDICT r1 names: chain_sweep_kernel<0>).
//
// The plain sweep (cta_sweep in kernels_solve.cuh) runs, per keyframe, the latency-bound fused Cholesky + inverse of the pivot block
// (one warp walks 106 sequential pivots) and THEN the panel GEMMs; the fp64 tensor pipe idles during the first and the pivot warp
// during the second.  Only the rows of the coupling block O_k feed the NEXT pivot block; the arrow rows (statics, right-hand side,
// left-separator rows: 165 of the 271 panel rows) are not needed before the next keyframe's own arrow rows.  So, per keyframe k:
//
//   S2  all 16 warps   W_k^T (upper triangle of the pivot buffer) -> 16 x 16 tile layout `Wt` (+ global store for the
//                      back-substitution); packed rows of O_k -> Op
//   S1  all 16 warps   rows of O_k:  y = p W^T, z = y W, fill = z O_k^T        (two groups of 8 warps on alternate 32-row chunks)
//   S3  warps 8..15    arrow rows of keyframe k against Wt / Op                 |  concurrently
//       warps 0..7     pivot block of keyframe k+1: load D - fill, Cholesky + inverse in the pivot buffer
//
// Shared memory: pivot buffer n x ld | Wt: n_tiles x (16 x 20) | panel buffer 32 x ld | Op   (~208 KB for n = 106).
// All panel GEMMs read W^T from the tile layout: tile (kb, cb), kb <= cb, row stride 20 doubles (== 4 mod 8: the DMMA operand
// fragments of both the NN use (y) and the NT use (z) fall into 16 distinct bank pairs per half warp).
// Same arithmetic as cta_sweep up to the summation order inside a dot product (contraction in whole 16-blocks).
#pragma once
#include "kernels_solve.cuh"

namespace b200d {

#define B200_TLD 20
#define B200_TSZ (16 * B200_TLD)
__host__ __device__ inline int pipe_ntiles(int n) { const int nb = (n + 15) >> 4; return nb * (nb + 1) / 2; }
__host__ __device__ inline long long sweep_pipe_smem_doubles(int n, int ld, int nc, int owp) {
    return (long long)n * ld + (long long)pipe_ntiles(n) * B200_TSZ + 32LL * ld + (((long long)nc * owp + 1) & ~1LL);
}
__device__ __forceinline__ int tile_off(int kb, int cb) { return (cb * (cb + 1) / 2 + kb) * B200_TSZ; }

// y tile: acc[.][0..1] += A[16 x k] W^T[k][block bA], acc[.][2..3] += ... [block bB]; k-blocks kb in [kb_lo, min(bA, .)] for A and
// [kb_lo, min(bB, kb_hi)] for B (W^T is upper triangular: only tiles kb <= cb exist).  A is row-major (ld), zero beyond column n.
__device__ __forceinline__ void tile_mma_y(const double* __restrict__ A, int lda, int m0, int M, int n, const double* __restrict__ Wt, int bA, int bB,
                                           int kb_lo, int kb_hi, bool useA, Acc& acc) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double* ap[2];
#pragma unroll
    for (int i = 0; i < 2; i++) { int m = m0 + 8 * i + g; if (m > M - 1) m = M - 1; ap[i] = A + (long long)m * lda + t; }
    const int kend = bB < kb_hi ? bB : kb_hi;   // (the rows of A are zero beyond k-block kb_hi)
    for (int kb = kb_lo; kb <= kend; kb++) {
        const bool doA = useA && kb <= bA;
        const double* tB = Wt + tile_off(kb, bB) + t * B200_TLD + g;
        const double* tA = Wt + tile_off(kb, doA ? bA : bB) + t * B200_TLD + g;
        const int k0 = kb << 4;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            const int k = k0 + 4 * ks;
            const bool ok = k + t < n;
            double a[2];
#pragma unroll
            for (int i = 0; i < 2; i++) a[i] = ok ? ap[i][k] : 0.0;
            const double b2 = tB[4 * ks * B200_TLD], b3 = tB[4 * ks * B200_TLD + 8];
            if (doA) {
                const double b0 = tA[4 * ks * B200_TLD], b1 = tA[4 * ks * B200_TLD + 8];
#pragma unroll
                for (int i = 0; i < 2; i++) { dmma884(acc[i][0][0], acc[i][0][1], a[i], b0); dmma884(acc[i][1][0], acc[i][1][1], a[i], b1); }
            }
#pragma unroll
            for (int i = 0; i < 2; i++) { dmma884(acc[i][2][0], acc[i][2][1], a[i], b2); dmma884(acc[i][3][0], acc[i][3][1], a[i], b3); }
        }
    }
}
// z tile: acc[.][0..1] += sum_{cb >= loA} Y[:, block cb] tile(bA, cb)^T, acc[.][2..3] likewise for bB (loX >= bX; loX >= nb: nothing)
__device__ __forceinline__ void tile_mma_z(const double* __restrict__ A, int lda, int m0, int M, int n, const double* __restrict__ Wt, int bA, int bB,
                                           int loA, int loB, int nb, Acc& acc) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const double* ap[2];
#pragma unroll
    for (int i = 0; i < 2; i++) { int m = m0 + 8 * i + g; if (m > M - 1) m = M - 1; ap[i] = A + (long long)m * lda + t; }
    const int lo = loA < loB ? loA : loB;
    for (int cb = lo; cb < nb; cb++) {
        const bool doA = cb >= loA, doB = cb >= loB;
        const double* tA = Wt + tile_off(doA ? bA : 0, cb) + g * B200_TLD + t;
        const double* tB = Wt + tile_off(doB ? bB : 0, cb) + g * B200_TLD + t;
        const int k0 = cb << 4;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
            const int k = k0 + 4 * ks;
            const bool ok = k + t < n;
            double a[2];
#pragma unroll
            for (int i = 0; i < 2; i++) a[i] = ok ? ap[i][k] : 0.0;
            if (doA) {
                const double b0 = tA[4 * ks], b1 = tA[8 * B200_TLD + 4 * ks];
#pragma unroll
                for (int i = 0; i < 2; i++) { dmma884(acc[i][0][0], acc[i][0][1], a[i], b0); dmma884(acc[i][1][0], acc[i][1][1], a[i], b1); }
            }
            if (doB) {
                const double b2 = tB[4 * ks], b3 = tB[8 * B200_TLD + 4 * ks];
#pragma unroll
                for (int i = 0; i < 2; i++) { dmma884(acc[i][2][0], acc[i][2][1], a[i], b2); dmma884(acc[i][3][0], acc[i][3][1], a[i], b3); }
            }
        }
    }
}

// requires: dm.nl == 0 (solver layout), block-sparse O, blockDim.x == 512, a.n >= 1
__device__ void cta_sweep_pipe(const Dims& dm, const SweepArgs& a, double* sm) {
    const int n = dm.ntp, nc = dm.nc, ns1 = dm.ns1, ld = dm.ld, hmax = dm.hmax, ow = dm.ow, owp = dm.owp;
    const long long Fsz = (long long)(nc + hmax) * nc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mg = lane >> 2, mt = lane & 3;
    const int nb16 = (n + 15) >> 4, npair = (nb16 + 1) >> 1;
    double* Ms = sm;
    double* Wt = Ms + (long long)n * ld;
    double* Pb = Wt + (long long)pipe_ntiles(n) * B200_TSZ;
    double* Op = Pb + 32LL * ld;
    __shared__ int s_orow[2 * 256], s_ozs[256], s_fblk[64], s_nfblk, s_abort;
    for (int e = threadIdx.x; e < 2 * nc; e += blockDim.x) s_orow[e] = dm.orow[e];
    for (int e = threadIdx.x; e < nc; e += blockDim.x) s_ozs[e] = dm.ozs[e];
    if (threadIdx.x == 0) s_abort = 0;
    __syncthreads();
    if (threadIdx.x == 0) {   // 8-column blocks of the block-sparse fill (see cta_sweep)
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
    // two groups of 8 warps; group barrier ids 7 / 8
    const int gnw = 8, gid = warp >> 3, gw = warp & 7, gthreads = gnw << 5, gtid = threadIdx.x & (gthreads - 1);
#define GSYNC() named_bar_sync(7 + gid, gthreads)

#ifdef B200_PHASE_CLOCKS
    unsigned long long pp_t = 0;
#endif
    // ---- pivot block of node j: M = D - fill (lower triangle), fused Cholesky + inverse.  Executed by group 0 only.
    auto factor_node = [&](int j) {
        const long long src = a.src0 + j;
        const double* Fin = a.F + (long long)(j & 1) * Fsz;
        const double* Dg = a.Dsrc + src * n * n;
        for (int r = gw; r < n; r += gnw)
            for (int c2 = lane; 2 * c2 <= r; c2 += 32) cp_async16(Ms + (long long)r * ld + 2 * c2, Dg + (long long)r * n + 2 * c2);
        if (j + 1 < a.n) {
            const double* Dn = a.Dsrc + (src + 1) * n * n;
            for (int e = gtid * 16; e < n * n; e += gthreads * 16) prefetch_l2(Dn + e);
        }
        for (int rb = 0; rb < n; rb += 8 * gnw) {
            for (int cb = 0; cb < n; cb += 128) {
                double f[8][4];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int r = rb + gw + q * gnw;
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int c = cb + lane + 32 * u;
                        f[q][u] = (j > 0 && r < n && c <= r) ? Fin[(long long)r * nc + c] : 0.0;
                    }
                }
                if (rb == 0 && cb == 0) { cp_async_wait_all(); GSYNC(); }
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int r = rb + gw + q * gnw;
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
        GSYNC();
#ifdef B200_PHASE_CLOCKS
        if (threadIdx.x == 0 && blockIdx.x == 1 && j > 0) { unsigned long long n_ = clock64(); g_phase_clk[3] += n_ - pp_t; pp_t = n_; }
#endif
        const int fail = cta_chol_inv(Ms, ld, n, gw, gnw, 7);
        if (fail && gtid == 0) {
            const long long st = a.st_idx ? a.st_idx[j] : src;
            atomicCAS(a.status, 0, (int)(st + 1));
            s_abort = 1;
        }
    };

    // ---- one chunk of <= 32 panel rows [r0, r0 + cr) of node i by this thread's group, in buffer Yb.  seg 0: rows of O, seg 1: arrow rows.
    auto panel_chunk = [&](int i, int seg, int r0, int cr, double* Yb, bool has_x) {
        const long long src = a.src0 + i;
        const long long st = a.st_idx ? a.st_idx[i] : src;
        const int xo = has_x ? nc : 0;
        const double* Fin = a.F + (long long)(i & 1) * Fsz;
        double* Fout = a.F + (long long)((i + 1) & 1) * Fsz;
        double* Yg = a.Yst + st * (long long)hmax * n;
        const double* Og = a.Osrc + src * nc * nc;
        const double* Ag = a.Asrc + src * ns1 * n;
        GSYNC();   // Yb free
        for (int lr = gw; lr < cr; lr += gnw) {
            const int pr = r0 + lr;
            double* dst = Yb + (long long)lr * ld;
            if (seg == 0) {
                const double* og = Og + (long long)pr * nc;
                for (int c = lane; c < n; c += 32) cp_async8(dst + c, og + c);
            } else {
                const int ar = pr - xo;
                if (ar < ns1) {
                    const double* ag = Ag + (long long)ar * n;
                    for (int c = 2 * lane; c < n; c += 64) cp_async16(dst + c, ag + c);
                } else if (i == 0 && a.Oleft) {
                    const double* ol = a.Oleft + (ar - ns1);
                    for (int c = lane; c < n; c += 32) dst[c] = ol[(long long)c * nc];
                } else if (i == 0) {
                    for (int c = lane; c < n; c += 32) dst[c] = 0.0;
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
                    if (rb == 0 && cb == 0) { cp_async_wait_all(); GSYNC(); }
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int lr = rb + gw + q * gnw;
                        if (lr >= cr) continue;
                        double* dr = Yb + (long long)lr * ld;
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
        GSYNC();
        // ---- y = p W^T : every warp owns one 16-row tile and the 16-column blocks (s, nb16 - 1 - s)
        const int rtiles = (cr + 15) >> 4;
        const int slot = gw & 3, rt = gw >> 2;
        const bool act = rt < rtiles && slot < npair;
        const int bA = slot, bB = nb16 - 1 - slot;
        const int m0 = rt << 4;
        int ka = 0, kb = n;
        if (seg == 0 && act) xrow_range(dm, s_orow, false, r0 + m0, ka, kb);
        {
            Acc acc;
            acc_zero(acc);
            if (act && kb > ka) tile_mma_y(Yb, ld, m0, cr, n, Wt, bA, bB, ka >> 4, (kb - 1) >> 4, bA != bB, acc);
            GSYNC();
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
                        st2(Yb + (long long)lr * ld + c, acc[q][j][0], acc[q][j][1]);
                        if (seg == 1) st2(Yg + (long long)(r0 - xo + lr) * n + c, acc[q][j][0], acc[q][j][1]);
                    }
                }
            }
            GSYNC();
        }
        if (!has_x) return;
        // ---- z = y W, in place after a barrier
        {
            Acc acc;
            acc_zero(acc);
            if (act && kb > ka) {
                int loA = bA, loB = bB;
                if (seg == 0) {
                    // rows of O: y is zero left of ka; z is only needed left of kb (the columns that feed the lower triangle of the fill)
                    const int klo = ka >> 4;
                    if (klo > loA) loA = klo;
                    if (klo > loB) loB = klo;
                    if ((bA << 4) >= kb) loA = nb16;
                    if ((bB << 4) >= kb) loB = nb16;
                }
                if (bA == bB) loA = nb16;
                tile_mma_z(Yb, ld, m0, cr, n, Wt, bA, bB, loA, loB, nb16, acc);
            }
            GSYNC();
            if (act) {
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int lr = m0 + 8 * q + mg;
                    if (lr >= cr) continue;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (j < 2 && bA == bB) continue;
                        const int c = ((j < 2 ? bA : bB) << 4) + 8 * (j & 1) + 2 * mt;
                        if (c < n) st2(Yb + (long long)lr * ld + c, acc[q][j][0], acc[q][j][1]);
                    }
                }
            }
            GSYNC();
        }
        // ---- fill rows F = z O^T (block-sparse coupling: 8-column blocks of rows of O that share one packed window)
        {
            const int nfb = s_nfblk, ncq = (nfb + 3) >> 2, nst = (ow + 3) >> 2;
            for (int task = gw; task < rtiles * ncq; task += gnw) {
                const int trt = task % rtiles, cq = task / rtiles;
                const int tm0 = trt << 4;
                int zs[4], cs[4];
                bool cval[4];
                const double* bp[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int fb = (cq << 2) + j;
                    cs[j] = fb < nfb ? s_fblk[fb] : nc;
                    zs[j] = fb < nfb ? s_ozs[cs[j]] : n;
                    const int ci = cs[j] + mg;
                    cval[j] = ci < nc && s_ozs[ci < nc ? ci : 0] == zs[j];
                    bp[j] = Op + (cval[j] ? ci : 0) * owp;
                }
                const double* ap[2];
#pragma unroll
                for (int q = 0; q < 2; q++) { int lr = tm0 + 8 * q + mg; if (lr > cr - 1) lr = cr - 1; ap[q] = Yb + (long long)lr * ld; }
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
                    const int lr = tm0 + 8 * q + mg;
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
    };
    // chunks of a segment of cnt rows starting at panel row `base`: balanced sizes, multiples of 16, <= 32
    auto seg_chunks = [&](int cnt, int& nch, int& csz) {
        nch = (cnt + 31) / 32;
        csz = nch ? (((cnt + nch - 1) / nch + 15) & ~15) : 0;
    };

    // ---- prologue: pivot block of the first node
#ifdef B200_PHASE_CLOCKS
    if (threadIdx.x == 0) g_phase_on = 1;
    pp_t = clock64();
#define PP_ADD(i, t) do { if (threadIdx.x == (t) && blockIdx.x == 1) { unsigned long long n_ = clock64(); g_phase_clk[i] += n_ - pp_t; pp_t = n_; } } while (0)
#else
#define PP_ADD(i, t)
#endif
    if (gid == 0) factor_node(0);
    PP_ADD(7, 0); PP_ADD(7, 256);
    for (int i = 0; i < a.n; i++) {
        const long long src = a.src0 + i;
        const long long st = a.st_idx ? a.st_idx[i] : src;
        const bool has_x = nc > 0 && ((i + 1 < a.n) || a.has_right);
        __syncthreads();   // [A] W_i^T complete in Ms; the arrow rows of node i-1 are done (Pb, Wt, Op free)
        PP_ADD(0, 0); PP_ADD(5, 256);
        if (s_abort) return;
        // ---- S2: W^T -> tile layout (zero padded to whole tiles) + global store; packed rows of O
        {
            double* Wg = a.Wst + st * n * n;
            const int nt = pipe_ntiles(n);
            for (int e = threadIdx.x; e < nt * 128; e += blockDim.x) {   // 128 pairs per tile
                const int tile = e >> 7, w = e & 127, tr = w >> 3, tc2 = (w & 7) << 1;
                int cb = 0;
                while ((cb + 1) * (cb + 2) / 2 <= tile) cb++;
                const int kb = tile - cb * (cb + 1) / 2;
                const int r = (kb << 4) + tr, c = (cb << 4) + tc2;
                double2 v = make_double2(0.0, 0.0);
                if (r < n && c < n) {   // (n is even: the pair is inside or outside)
                    v = ld2(Ms + (long long)r * ld + c);
                    if (c < r) v.x = 0.0;          // what is left of L below the diagonal is not part of W^T
                    if (c + 1 < r) v.y = 0.0;
                }
                st2(Wt + tile * B200_TSZ + tr * B200_TLD + tc2, v.x, v.y);
            }
            if (a.lead) {
                for (int r = warp; r < n; r += 16)
                    for (int c = 2 * lane; c < n; c += 64) {
                        double2 v = ld2(Ms + (long long)r * ld + c);
                        if (c < r) { v.x = 0.0; if (c + 1 < r) v.y = 0.0; }
                        st2(Wg + (long long)r * n + c, v.x, v.y);
                    }
            }
            if (has_x) {
                const double* Og = a.Osrc + src * nc * nc;
                for (int e = threadIdx.x; e < nc * ow; e += blockDim.x) {
                    const int r = e / ow, t = e - r * ow;
                    const int cc = s_ozs[r] + t;
                    if (cc >= 0 && cc < nc) cp_async8(Op + r * owp + t, Og + (long long)r * nc + cc);
                    else Op[r * owp + t] = 0.0;
                }
                for (int e = threadIdx.x * 16; e < nc * nc; e += blockDim.x * 16) prefetch_l2(Og + e);
            }
            {
                const double* Ag = a.Asrc + src * ns1 * n;
                for (int e = threadIdx.x * 16; e < ns1 * n; e += blockDim.x * 16) prefetch_l2(Ag + e);
            }
            cp_async_wait_all();
        }
        __syncthreads();   // [B] Wt / Op complete, Ms free
        PP_ADD(1, 0); PP_ADD(16, 256);
        // ---- S1: rows of O (they feed the next pivot block): both groups, alternate chunks; group 0 in Pb, group 1 in the pivot buffer
        if (has_x) {
            int nch, csz;
            seg_chunks(nc, nch, csz);
            double* Yb = gid == 0 ? Pb : Ms;
            for (int ch = gid; ch < nch; ch += 2) {
                const int r0 = ch * csz;
                const int cr = (nc - r0 < csz) ? (nc - r0) : csz;
                panel_chunk(i, 0, r0, cr, Yb, has_x);
            }
        }
        __threadfence_block();
        __syncthreads();   // [C] the fill of the next pivot block is complete
        PP_ADD(2, 0); PP_ADD(17, 256);
        // ---- S3: group 1: arrow rows of node i   ||   group 0: pivot block of node i + 1
        if (gid == 1) {
            const int xo = has_x ? nc : 0;
            const int cnt = a.ar_hi - a.ar_lo;
            int nch, csz;
            seg_chunks(cnt, nch, csz);
            for (int ch = 0; ch < nch; ch++) {
                const int r0 = xo + a.ar_lo + ch * csz;
                const int cr = (xo + a.ar_lo + cnt - r0 < csz) ? (xo + a.ar_lo + cnt - r0) : csz;
                panel_chunk(i, 1, r0, cr, Pb, has_x);
            }
            __threadfence_block();
            PP_ADD(6, 256);
        } else if (i + 1 < a.n) {
            factor_node(i + 1);
            PP_ADD(4, 0);
        }
    }
#undef GSYNC
    __syncthreads();
}

}  // namespace b200d
