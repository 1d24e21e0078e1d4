// Level-0 chain sweep, software-pipelined ACROSS keyframes, panel rows streamed through Tensor Memory (the kernel VERDICT r1
// names: chain_sweep_kernel<0>).
//
// Per keyframe k the sweep has a latency-bound part -- the fused Cholesky + inverse of the pivot block, a chain of 106 sequential
// pivots -- and a throughput-bound part: the panel rows [O_k ; statics ; rhs ; left separator] (271 rows) that go through
//     y = p W^T  ->  z = y W  ->  fill = z O_k^T          (two triangular GEMMs + one block-sparse GEMM per row)
// Only the rows of O_k feed the NEXT pivot block; the arrow rows are not needed before the next keyframe's own arrow rows.
//
//   S2  all 16 warps    W_k^T (upper triangle of the pivot buffer) -> 16 x 16 tile layout `Wt` (+ global store for the
//                       back-substitution); packed rows of O_k -> Op
//   S1  all 16 warps    rows of O_k (7 row tiles): their fill goes straight into the pivot buffer, M_k+1 = D_k+1 + lambda I - fill
//   S3  warps 5..15     arrow rows of keyframe k (11 row tiles, one per warp)        |  concurrently
//       warps 0..4      pivot block of keyframe k+1: Cholesky + inverse in the pivot buffer
//
// ROW TILES LIVE IN TENSOR MEMORY.  A row tile (16 rows x n columns) is owned by ONE warp from its load to its last store: the
// warp keeps it in its private slice of TMEM as accumulator-layout fragments (lane (g, t): rows g and 8 + g, columns 8 j + 2 t + e
// of every 8-column block j), reads it back with tcgen05.ld as the A operand of the fp64 MMA -- the contraction index is permuted,
// slot t of step e <-> k = 8 j + 2 t + e, so that an accumulator fragment IS an A fragment, no shuffle, no shared-memory round
// trip -- and overwrites it in place (y over p from the last column block down, z over y from the first one up).  No panel
// buffer in shared memory, no barrier between the three GEMMs, no barrier between row tiles: every warp streams MMAs
// independently against the read-only W^T tiles and the packed O rows (r2 profile of the previous version: 32-row chunks shared
// by 8 warps, six group barriers per chunk, 33 % of the tensor pipe in the arrow phase; tools/ubench/tmem_probe.cu: 79 % with 8
// independent warps and the A operand in TMEM).
//
// Shared memory: pivot buffer n x ld | Wt: n_tiles x (16 x 20) | Op   (~181 KB for n = 106).  TMEM: 512 columns, warp w owns
// lanes 32 (w & 3).. and columns 128 (w >> 2)..
// W^T tiles: tile (kb, cb), kb <= cb, row stride 20 doubles (== 4 mod 8): the B fragments of the NN use (y: rows 2t + e, column
// g, two LDS.64) and of the NT use (z: row g, columns 2t, 2t + 1, one LDS.128) are both bank-conflict free.
#pragma once
#include "kernels_solve.cuh"

namespace b200d {

#define B200_TLD 20
#define B200_TSZ (16 * B200_TLD)
#define B200_PIVOT_WARPS 5
__host__ __device__ inline int pipe_ntiles(int n) { const int nb = (n + 15) >> 4; return nb * (nb + 1) / 2; }
__host__ __device__ inline long long sweep_pipe_smem_doubles(int n, int ld, int nc, int owp) {
    return (long long)n * ld + (long long)pipe_ntiles(n) * B200_TSZ + (((long long)nc * owp + 1) & ~1LL);
}
__device__ __forceinline__ int tile_off(int kb, int cb) { return (cb * (cb + 1) / 2 + kb) * B200_TSZ; }

struct RowTile {
    int seg;        // 0: rows of O (fill of the next pivot block), 1: arrow rows
    int pr0;        // first panel row of the tile (row of the fill buffer); arrow row index = pr0 - xo
    int cr;         // rows in the tile (<= 16)
    int xo;         // nc when the node has rows of O, else 0
    bool has_x;
};

// One row tile, start to finish, by the calling warp (all 32 lanes, convergent).  tb: TMEM address of the warp's slice.
// DENSE: the coupling block is dense (reduced systems of the upper levels): Op points at the whole block O (row-major, stride dm.ld)
// and the fill is a dense GEMM z O^T; else O is block-sparse and Op holds its packed rows.
template <int NH, bool DENSE = false>
__device__ __forceinline__ void stream_row_tile(const Dims& dm, const SweepArgs& a, int i, const RowTile rt, const double* __restrict__ Wt,
                                                const double* __restrict__ Op, const int* s_orow, const int* s_ozs, const int* s_fk, unsigned tb,
                                                double* Mnext = nullptr) {
    const int n = dm.ntp, nc = dm.nc, ns1 = dm.ns1, hmax = dm.hmax, ow = dm.ow, owp = dm.owp;
    const long long Fsz = (long long)(nc + hmax) * nc;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nb16 = (n + 15) >> 4;
    const long long src = a.src0 + i;
    const long long st = a.st_idx ? a.st_idx[i] : src;
    const double* Fin = a.F + (long long)(i & 1) * Fsz;
    double* Fout = a.F + (long long)((i + 1) & 1) * Fsz;
    const bool rv[2] = {g < rt.cr, NH > 1 && 8 + g < rt.cr};   // the row halves of this lane exist (NH = 1: 8-row tile)
    // ---- k range of the rows of O of this tile (block-sparse coupling): p is zero outside [ka, kb)
    int ka = 0, kb = n;
    if (rt.seg == 0) xrow_range(dm, s_orow, DENSE, rt.pr0, ka, kb, rt.cr);
    if (kb <= ka) {   // rows without structure: their fill is zero
        for (int c = 2 * t; c < nc; c += 8)
#pragma unroll
            for (int ih = 0; ih < NH; ih++) {
                if (!rv[ih]) continue;
                const int r = rt.pr0 + 8 * ih + g;
                if (Mnext) {
                    const double2 d = ld2(a.Dsrc + (src + 1) * n * n + (long long)r * n + c);
                    st2(Mnext + (long long)r * dm.ld + c, (c <= r) ? d.x + ((c == r) ? a.lambda : 0.0) : 0.0, (c + 1 <= r) ? d.y + ((c + 1 == r) ? a.lambda : 0.0) : 0.0);
                } else st2(Fout + (long long)r * nc + c, 0.0, 0.0);
            }
        return;
    }
    const int kb_lo = ka >> 4, kb_hi = (kb - 1) >> 4;
    // ---- load p into TMEM (fragment layout)
    if (rt.seg == 0) {
        const double* Og = a.Osrc + src * nc * nc;
        for (int j = 2 * kb_lo; j <= 2 * kb_hi + 1; j++) {
            const int c = 8 * j + 2 * t;
            double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int ih = 0; ih < NH; ih++) {
                double2 x = make_double2(0.0, 0.0);
                if (rv[ih] && c < nc) x = ld2(Og + (long long)(rt.pr0 + 8 * ih + g) * nc + c);   // (nc even: the pair is inside or outside)
                v[2 * ih] = x.x; v[2 * ih + 1] = x.y;
            }
            tmem_st4(tb + 8 * j, v);
        }
    } else {
        const double* Ag = a.Asrc + src * ns1 * n;
        const int ar0 = rt.pr0 - rt.xo;
        for (int j0 = 0; j0 < 2 * nb16; j0 += 2) {
            double2 x[2][2], f[2][2];
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                const int c = 8 * (j0 + jj) + 2 * t;
#pragma unroll
                for (int ih = 0; ih < NH; ih++) {
                    const int ar = ar0 + 8 * ih + g;
                    x[jj][ih] = make_double2(0.0, 0.0); f[jj][ih] = make_double2(0.0, 0.0);
                    if (!rv[ih] || c >= n) continue;
                    if (ar < ns1) x[jj][ih] = ld2(Ag + (long long)ar * n + c);
                    else if (i == 0 && a.Oleft) { const double* ol = a.Oleft + (ar - ns1); x[jj][ih].x = ol[(long long)c * nc]; x[jj][ih].y = ol[(long long)(c + 1) * nc]; }
                    if (i > 0 && c < nc) f[jj][ih] = ld2(Fin + (long long)(nc + ar) * nc + c);
                }
            }
#pragma unroll
            for (int jj = 0; jj < 2; jj++) {
                double v[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int ih = 0; ih < NH; ih++) { v[2 * ih] = x[jj][ih].x - f[jj][ih].x; v[2 * ih + 1] = x[jj][ih].y - f[jj][ih].y; }
                tmem_st4(tb + 8 * (j0 + jj), v);
            }
        }
    }
    tmem_wait_st();
    // ---- y = p W^T, in place from the last 16-column block down: pass over the blocks (c1, c0 = c1 - 1); y[:, c] needs p[:, k <= c]
    double* Yg = a.Yst + st * (long long)hmax * n;
    for (int c1 = nb16 - 1; c1 >= kb_lo; c1 -= 2) {
        const int c0 = c1 - 1;
        const bool use0 = c0 >= kb_lo;
        Acc acc;
        acc_zero(acc);
        const int kend = c1 < kb_hi ? c1 : kb_hi;
        for (int j8 = 2 * kb_lo; j8 <= 2 * kend + 1; j8++) {
            const int kbk = j8 >> 1, h = j8 & 1;
            const bool do0 = use0 && kbk <= c0;
            const double* t1 = Wt + tile_off(kbk, c1) + g;
            const double* t0 = Wt + tile_off(kbk, do0 ? c0 : c1) + g;
            {
                double av[4];
                tmem_ld4(tb + 8 * j8, av);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int kr = (8 * h + 2 * t + e) * B200_TLD;
                    const double b2 = t1[kr], b3 = t1[kr + 8];
                    if (do0) {
                        const double b0 = t0[kr], b1 = t0[kr + 8];
#pragma unroll
                        for (int ih = 0; ih < NH; ih++) { dmma884(acc[ih][0][0], acc[ih][0][1], av[2 * ih + e], b0); dmma884(acc[ih][1][0], acc[ih][1][1], av[2 * ih + e], b1); }
                    }
#pragma unroll
                    for (int ih = 0; ih < NH; ih++) { dmma884(acc[ih][2][0], acc[ih][2][1], av[2 * ih + e], b2); dmma884(acc[ih][3][0], acc[ih][3][1], av[2 * ih + e], b3); }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (q < 2 && !use0) continue;
            const int j = 2 * (q < 2 ? c0 : c1) + (q & 1);
            double v[4] = {acc[0][q][0], acc[0][q][1], NH > 1 ? acc[NH - 1][q][0] : 0.0, NH > 1 ? acc[NH - 1][q][1] : 0.0};
            tmem_st4(tb + 8 * j, v);
            if (DENSE && rt.seg == 0 && Mnext) {   // dense levels: the rows of X = O W^T, B operand of every fill of this node
                const int c = 8 * j + 2 * t;
                if (c < n) {
#pragma unroll
                    for (int ih = 0; ih < NH; ih++)
                        if (rv[ih]) st2(Mnext + (long long)(rt.pr0 + 8 * ih + g) * dm.ld + c, acc[ih][q][0], acc[ih][q][1]);
                }
            }
            if (rt.seg == 1) {
                const int c = 8 * j + 2 * t;
                if (c < n) {
#pragma unroll
                    for (int ih = 0; ih < NH; ih++)
                        if (rv[ih]) st2(Yg + (long long)(rt.pr0 - rt.xo + 8 * ih + g) * n + c, acc[ih][q][0], acc[ih][q][1]);
                }
            }
        }
    }
    if (DENSE) { tmem_wait_st(); return; }   // dense levels: the fill is y X^T (dense_fill_tile), after every tile of the node has its y
    if (!rt.has_x) return;
    tmem_wait_st();
    // ---- z = y W, in place from the first 16-column block up: z[:, c] = sum_{k >= max(c, kb_lo)} y[:, k] tile(c, k)^T.
    // Rows of O: z is only needed left of kb (the columns that feed the lower triangle of the fill).
    const int zc_hi = rt.seg == 0 ? kb_hi : nb16 - 1;
    for (int c0 = 0; c0 <= zc_hi; c0 += 2) {
        const int c1 = c0 + 1;
        const bool use1 = c1 <= zc_hi;
        Acc acc;
        acc_zero(acc);
        const int k0 = c0 > kb_lo ? c0 : kb_lo;
        for (int j8 = 2 * k0; j8 < 2 * nb16; j8++) {
            const int kbk = j8 >> 1, h = j8 & 1;
            const bool do1 = use1 && kbk >= c1;
            const double* t0 = Wt + tile_off(c0, kbk) + g * B200_TLD + 2 * t;
            const double* t1 = Wt + tile_off(do1 ? c1 : c0, kbk) + g * B200_TLD + 2 * t;
            {
                double av[4];
                tmem_ld4(tb + 8 * j8, av);
                const double2 b0 = ld2(t0 + 8 * h), b1 = ld2(t0 + 8 * B200_TLD + 8 * h);
#pragma unroll
                for (int ih = 0; ih < NH; ih++) {
                    dmma884(acc[ih][0][0], acc[ih][0][1], av[2 * ih], b0.x); dmma884(acc[ih][0][0], acc[ih][0][1], av[2 * ih + 1], b0.y);
                    dmma884(acc[ih][1][0], acc[ih][1][1], av[2 * ih], b1.x); dmma884(acc[ih][1][0], acc[ih][1][1], av[2 * ih + 1], b1.y);
                }
                if (do1) {
                    const double2 b2 = ld2(t1 + 8 * h), b3 = ld2(t1 + 8 * B200_TLD + 8 * h);
#pragma unroll
                    for (int ih = 0; ih < NH; ih++) {
                        dmma884(acc[ih][2][0], acc[ih][2][1], av[2 * ih], b2.x); dmma884(acc[ih][2][0], acc[ih][2][1], av[2 * ih + 1], b2.y);
                        dmma884(acc[ih][3][0], acc[ih][3][1], av[2 * ih], b3.x); dmma884(acc[ih][3][0], acc[ih][3][1], av[2 * ih + 1], b3.y);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (q >= 2 && !use1) continue;
            const int j = 2 * (q < 2 ? c0 : c1) + (q & 1);
            double v[4] = {acc[0][q][0], acc[0][q][1], NH > 1 ? acc[NH - 1][q][0] : 0.0, NH > 1 ? acc[NH - 1][q][1] : 0.0};
            tmem_st4(tb + 8 * j, v);
        }
    }
    tmem_wait_st();
    // ---- fill rows F = z O^T (block-sparse coupling): aligned 8-column output blocks; row co of O is packed in Op as the window
    // [ozs[co], ozs[co] + ow) and the block contracts over the union of its rows' windows (s_fk: first / last 8-column block of k)
    {
        const int jz_hi = 2 * zc_hi + 1;   // last 8-column block of z that exists
        const int jo_hi = rt.seg == 0 ? ((rt.pr0 + rt.cr - 1) >> 3) : ((nc - 1) >> 3);   // rows of O: lower triangle of the fill only
        // rows of O with a next node in this chunk: the fill goes STRAIGHT into the pivot buffer, M_next = D_next + lambda I - fill
        // (lower triangle; zeros above the diagonal), no round trip through global memory and no separate load phase
        const double* Dn = Mnext ? a.Dsrc + (src + 1) * n * n : nullptr;
        const int ld = dm.ld;
        // D_next of output block jo + 1 is fetched while block jo is computed (the L2 latency stays off the chain of fill blocks)
        double2 dnx[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
        if (Mnext && 2 * t < n) {
#pragma unroll
            for (int ih = 0; ih < NH; ih++)
                if (rv[ih]) dnx[ih] = ld2(Dn + (long long)(rt.pr0 + 8 * ih + g) * n + 2 * t);
        }
        for (int jo = 0; jo <= jo_hi && 8 * jo < nc; jo++) {
            const int co = 8 * jo + g;
            const bool cok = co < nc;
            const int zs = cok ? s_ozs[co] : 0;
            const double* orow_p = Op + (long long)(cok ? co : 0) * owp;
            double f[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
            const int jk0 = s_fk[2 * jo];
            const int jk1 = s_fk[2 * jo + 1] < jz_hi ? s_fk[2 * jo + 1] : jz_hi;
            const double2 dn[2] = {dnx[0], dnx[1]};
            if (Mnext && jo + 1 <= jo_hi && 8 * (jo + 1) + 2 * t < n) {
#pragma unroll
                for (int ih = 0; ih < NH; ih++)
                    if (rv[ih]) dnx[ih] = ld2(Dn + (long long)(rt.pr0 + 8 * ih + g) * n + 8 * (jo + 1) + 2 * t);
            }
            for (int jk = jk0; jk <= jk1; jk++) {
                double av[4];
                tmem_ld4(tb + 8 * jk, av);
                const int off = 8 * jk + 2 * t - zs;
                double2 b = make_double2(0.0, 0.0);
                if (cok && off >= 0 && off < ow) b = ld2(orow_p + off);
#pragma unroll
                for (int ih = 0; ih < NH; ih++) { dmma884(f[ih][0], f[ih][1], av[2 * ih], b.x); dmma884(f[ih][0], f[ih][1], av[2 * ih + 1], b.y); }
            }
            const int cw = 8 * jo + 2 * t;
            if (cw < nc) {
#pragma unroll
                for (int ih = 0; ih < NH; ih++) {
                    if (!rv[ih]) continue;
                    const int r = rt.pr0 + 8 * ih + g;
                    if (Mnext) {
                        const double m0 = (cw <= r) ? dn[ih].x + ((cw == r) ? a.lambda : 0.0) - f[ih][0] : 0.0;
                        const double m1 = (cw + 1 <= r) ? dn[ih].y + ((cw + 1 == r) ? a.lambda : 0.0) - f[ih][1] : 0.0;
                        st2(Mnext + (long long)r * ld + cw, m0, m1);
                    } else st2(Fout + (long long)r * nc + cw, f[ih][0], f[ih][1]);
                }
            }
        }
        if (Mnext) {   // the rest of the strictly upper part of these rows
            for (int cw = 8 * (jo_hi + 1) + 2 * t; cw < n; cw += 8)
#pragma unroll
                for (int ih = 0; ih < NH; ih++)
                    if (rv[ih]) st2(Mnext + (long long)(rt.pr0 + 8 * ih + g) * ld + cw, 0.0, 0.0);
        }
    }
}

// Dense levels: fill rows of one tile, F = y X^T with X = O W^T (= z O^T, one triangular GEMM less): y from the warp's TMEM slice,
// X rows in shared memory (row-major, stride ld; written by the rows-of-O tiles of the same node).  Four 8-column output blocks
// per pass; B[slot][nn] = X[co + nn][8 jk + 2 t + e] as one LDS.128 per block (row g, columns 2t, 2t + 1: conflict free for ld == 4 mod 8).
template <int NH>
__device__ __forceinline__ void dense_fill_tile(const Dims& dm, const SweepArgs& a, int i, const RowTile rt, const double* __restrict__ Xs, unsigned tb) {
    const int n = dm.ntp, nc = dm.nc, hmax = dm.hmax, ldo = dm.ld;
    const long long Fsz = (long long)(nc + hmax) * nc;
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nb16 = (n + 15) >> 4;
    double* Fout = a.F + (long long)((i + 1) & 1) * Fsz;
    const bool rv[2] = {g < rt.cr, NH > 1 && 8 + g < rt.cr};
    const int jz_hi = 2 * nb16 - 1;
    const int jo_hi = rt.seg == 0 ? ((rt.pr0 + rt.cr - 1) >> 3) : ((nc - 1) >> 3);   // rows of O: lower triangle of the fill only
    for (int jo = 0; jo <= jo_hi && 8 * jo < nc; jo += 4) {
        Acc acc;
        acc_zero(acc);
        const double* op[4];
        bool ok[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int co = 8 * (jo + q) + g;
            ok[q] = (jo + q <= jo_hi) && co < nc;
            op[q] = Xs + (long long)(ok[q] ? co : 0) * ldo + 2 * t;
        }
        for (int jk = 0; jk <= jz_hi; jk++) {
            double av[4];
            tmem_ld4(tb + 8 * jk, av);
            const bool kin = 8 * jk + 2 * t < n;
            double2 b[4];
#pragma unroll
            for (int q = 0; q < 4; q++) b[q] = (ok[q] && kin) ? ld2(op[q] + 8 * jk) : make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int ih = 0; ih < NH; ih++) dmma884(acc[ih][q][0], acc[ih][q][1], av[2 * ih], b[q].x);
#pragma unroll
            for (int q = 0; q < 4; q++)
#pragma unroll
                for (int ih = 0; ih < NH; ih++) dmma884(acc[ih][q][0], acc[ih][q][1], av[2 * ih + 1], b[q].y);
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int cw = 8 * (jo + q) + 2 * t;
            if (jo + q > jo_hi || cw >= nc) continue;
#pragma unroll
            for (int ih = 0; ih < NH; ih++)
                if (rv[ih]) st2(Fout + (long long)(rt.pr0 + 8 * ih + g) * nc + cw, acc[ih][q][0], acc[ih][q][1]);
        }
    }
}

// Rows [16 j + 8 ih0, +8) of W^T = L^-T from the plain Cholesky factor left in the pivot buffer (strictly lower blocks = L, diagonal
// tiles = W_cc^T): block forward substitution of an identity row tile, right-looking and entirely in registers,
//     U[:, j] = W_jj^T;   for k = j, j+1, ...:  T_c += U[:, k] L[c][k]^T for every c > k (independent accumulators),
//                                               U[:, k+1] = -T_{k+1} W_{k+1,k+1}^T
// An accumulator fragment is fed back as the A operand through the permuted contraction index (slot t of step (q, e) <->
// k = 8 q + 2 t + e), so nothing leaves the registers between the steps; the dependent chain is 8 MMAs per column block.
// Writes the tile layout Wt (all tiles (j, c >= j), zero padded) and the global copy for the back-substitution.
__device__ __forceinline__ void build_wt_rows(const double* __restrict__ Ms, int ld, int n, int j, int ih0, double* __restrict__ Wt,
                                              double* __restrict__ Wg) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nb16 = (n + 15) >> 4;   // <= 8
    const int lr = 8 * ih0 + g, r = 16 * j + lr;   // this lane's row (local / global)
    const bool rok = r < n;
    // zeros left of the diagonal tile (the global copy holds whole rows)
    if (Wg && rok) for (int c = 2 * t; c < 16 * j; c += 8) st2(Wg + (long long)r * n + c, 0.0, 0.0);
    double u[2][2];   // current column block: fragment (q = 8-column half, e)
#pragma unroll
    for (int q = 0; q < 2; q++) {   // U[:, j] = W_jj^T: upper triangle incl. the diagonal of the diagonal tile
        const int c = 16 * j + 8 * q + 2 * t;
        double2 v = make_double2(0.0, 0.0);
        if (rok && c < n) { v = ld2(Ms + (long long)r * ld + c); if (c < r) v.x = 0.0; if (c + 1 < r) v.y = 0.0; }
        u[q][0] = v.x; u[q][1] = v.y;
    }
    double T[8][2][2];
#pragma unroll
    for (int c = 0; c < 8; c++) { T[c][0][0] = 0.0; T[c][0][1] = 0.0; T[c][1][0] = 0.0; T[c][1][1] = 0.0; }
    for (int k = j; k < nb16; k++) {
        // store block k: tile layout, global copy
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int cc = 16 * k + 8 * q + 2 * t;
            const bool in = rok && cc < n;
            st2(Wt + tile_off(j, k) + lr * B200_TLD + 8 * q + 2 * t, in ? u[q][0] : 0.0, in ? u[q][1] : 0.0);
            if (Wg && in) st2(Wg + (long long)r * n + cc, u[q][0], u[q][1]);
        }
        if (k + 1 >= nb16) break;
        // T_c += U[:, k] L[c][k]^T, c > k   (B[slot][nn] = L[16 c + nn][16 k + 8 q + 2 t + e])
#pragma unroll
        for (int c = 1; c < 8; c++) {
            if (c > k && c < nb16) {
                const int rb0 = 16 * c + g, rb1 = 16 * c + 8 + g;
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int kc = 16 * k + 8 * q + 2 * t;
                    const double2 b0 = rb0 < n ? ld2(Ms + (long long)rb0 * ld + kc) : make_double2(0.0, 0.0);
                    const double2 b1 = rb1 < n ? ld2(Ms + (long long)rb1 * ld + kc) : make_double2(0.0, 0.0);
                    dmma884(T[c][0][0], T[c][0][1], u[q][0], b0.x); dmma884(T[c][0][0], T[c][0][1], u[q][1], b0.y);
                    dmma884(T[c][1][0], T[c][1][1], u[q][0], b1.x); dmma884(T[c][1][0], T[c][1][1], u[q][1], b1.y);
                }
            }
        }
        // U[:, k+1] = -T_{k+1} W_cc^T, cc = k + 1   (A: T fragments; B[slot][nn] = W_cc^T[8 q + 2 t + e][nn])
        double tk[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int c = 1; c < 8; c++)
            if (c == k + 1) { tk[0][0] = T[c][0][0]; tk[0][1] = T[c][0][1]; tk[1][0] = T[c][1][0]; tk[1][1] = T[c][1][1]; }
        double o[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        const int cb = 16 * (k + 1);
#pragma unroll
        for (int q = 0; q < 2; q++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int kr = cb + 8 * q + 2 * t + e;
                const int cc0 = cb + g, cc1 = cb + 8 + g;
                const double b0 = (kr < n && cc0 < n && cc0 >= kr) ? Ms[(long long)kr * ld + cc0] : 0.0;
                const double b1 = (kr < n && cc1 < n && cc1 >= kr) ? Ms[(long long)kr * ld + cc1] : 0.0;
                dmma884(o[0][0], o[0][1], tk[q][e], b0);
                dmma884(o[1][0], o[1][1], tk[q][e], b1);
            }
#pragma unroll
        for (int q = 0; q < 2; q++) { u[q][0] = -o[q][0]; u[q][1] = -o[q][1]; }
    }
}

// requires: dm.nl == 0 (solver layout), block-sparse O, blockDim.x == 512, a.n >= 1, ntp <= 128, ow even
__device__ void cta_sweep_pipe(const Dims& dm, const SweepArgs& a, double* sm) {
    const int n = dm.ntp, nc = dm.nc, ns1 = dm.ns1, ld = dm.ld, hmax = dm.hmax, ow = dm.ow, owp = dm.owp;
    const long long Fsz = (long long)(nc + hmax) * nc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* Ms = sm;
    double* Wt = Ms + (long long)n * ld;
    double* Op = Wt + (long long)pipe_ntiles(n) * B200_TSZ;
    __shared__ int s_orow[2 * 128], s_ozs[128], s_fk[2 * 16], s_abort;
    __shared__ unsigned s_tmem;
    for (int e = threadIdx.x; e < 2 * nc; e += blockDim.x) s_orow[e] = dm.orow[e];
    for (int e = threadIdx.x; e < nc; e += blockDim.x) s_ozs[e] = dm.ozs[e];
    if (threadIdx.x == 0) s_abort = 0;
    if (warp == 0) tmem_alloc512(&s_tmem);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    if (threadIdx.x < 16) {   // contraction range (8-column blocks of z) of every aligned 8-column output block of the fill
        const int jo = threadIdx.x;
        int lo = n, hi = 0;
        for (int co = 8 * jo; co < 8 * jo + 8 && co < nc; co++) {
            if (s_ozs[co] < lo) lo = s_ozs[co];
            if (s_ozs[co] + ow > hi) hi = s_ozs[co] + ow;
        }
        if (hi > n) hi = n;
        s_fk[2 * jo] = lo >> 3; s_fk[2 * jo + 1] = hi > lo ? (hi - 1) >> 3 : (lo >> 3) - 1;
    }
    __syncthreads();
    const unsigned tb = s_tmem + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(128 * (warp >> 2));
    // pivot group: warps [0, B200_PIVOT_WARPS), named barrier 7
    const int gnw = B200_PIVOT_WARPS, gthreads = gnw << 5;
    const bool pivot_grp = warp < gnw;
    const int gw = warp, gtid = threadIdx.x;
#define GSYNC() named_bar_sync(7, gthreads)
#ifdef B200_PHASE_CLOCKS
    unsigned long long pp_t = 0;
#endif

    // ---- pivot block of node j: M = D - fill (lower triangle), fused Cholesky + inverse.  Executed by the pivot group only.
    auto factor_node = [&](int j) {
        const long long src = a.src0 + j;
        if (j == 0) {   // first node of the chunk: M = D + lambda I from global memory (later nodes: written in place by the rows of O)
            const double* Dg = a.Dsrc + src * n * n;
            for (int r = gw; r < n; r += gnw)
                for (int c = 2 * lane; c < n; c += 64) {
                    double2 v = make_double2(0.0, 0.0);
                    if (c <= r) {
                        v = ld2(Dg + (long long)r * n + c);
                        if (c == r) v.x += a.lambda;
                        if (c + 1 == r) v.y += a.lambda;
                        if (c + 1 > r) v.y = 0.0;
                    }
                    st2(Ms + (long long)r * ld + c, v.x, v.y);
                }
            GSYNC();
        }   // (later nodes: the pivot buffer was completed before barrier [C], which every warp of the group has just left)
#ifdef B200_PHASE_CLOCKS
        if (threadIdx.x == 0 && blockIdx.x == 1 && j > 0) { unsigned long long n_ = clock64(); g_phase_clk[3] += n_ - pp_t; pp_t = n_; }
#endif
        const int fail = cta_chol_inv(Ms, ld, n, gw, gnw, 7, true);
        if (fail && gtid == 0) {
            const long long st = a.st_idx ? a.st_idx[j] : src;
            atomicCAS(a.status, 0, (int)(st + 1));
            s_abort = 1;
        }
    };

    // ---- prologue: pivot block of the first node
#ifdef B200_PHASE_CLOCKS
    if (threadIdx.x == 0) g_phase_on = 1;
    pp_t = clock64();
#define PP_ADD(i, t) do { if (threadIdx.x == (t) && blockIdx.x == 1) { unsigned long long n_ = clock64(); g_phase_clk[i] += n_ - pp_t; pp_t = n_; } } while (0)
#else
#define PP_ADD(i, t)
#endif
    if (pivot_grp) factor_node(0);
    PP_ADD(7, 0); PP_ADD(7, 256);
    for (int i = 0; i < a.n; i++) {
        const long long src = a.src0 + i;
        const long long st = a.st_idx ? a.st_idx[i] : src;
        const bool has_x = nc > 0 && ((i + 1 < a.n) || a.has_right);
        const int xo = has_x ? nc : 0;
        __syncthreads();   // [A] W_i^T complete in Ms; the arrow rows of node i-1 are done (Wt, Op free)
        PP_ADD(0, 0); PP_ADD(5, 256);
        if (s_abort) break;
        // ---- S2: W^T = L^-T row tiles by forward substitution (two warps per 16-row tile) -> tile layout + global store; packed rows of O
        {
            double* Wg = a.Wst + st * n * n;
            const int nbt = (n + 15) >> 4;
            if (has_x) {
                const double* Og = a.Osrc + src * nc * nc;
                for (int e = threadIdx.x; e < nc * ow; e += blockDim.x) {
                    const int r = e / ow, t = e - r * ow;
                    const int cc = s_ozs[r] + t;
                    if (cc >= 0 && cc < nc) cp_async8(Op + r * owp + t, Og + (long long)r * nc + cc);
                    else Op[r * owp + t] = 0.0;
                }
            }
            for (int task = warp; task < 2 * nbt; task += 16) build_wt_rows(Ms, ld, n, task >> 1, task & 1, Wt, a.lead ? Wg : nullptr);
            cp_async_wait_all();
        }
        __syncthreads();   // [B] Wt / Op complete, Ms free
        PP_ADD(1, 0); PP_ADD(16, 256);
        // ---- S1: rows of O (they feed the next pivot block): one 16-row tile per warp
        if (has_x) {
            const int ntile = (nc + 7) >> 3;   // 8-row tiles: 14 tasks for the 16 warps
            for (int tl = warp; tl < ntile; tl += 16) {
                RowTile rt;
                rt.seg = 0; rt.pr0 = tl << 3; rt.cr = (nc - rt.pr0 < 8) ? (nc - rt.pr0) : 8; rt.xo = xo; rt.has_x = true;
                stream_row_tile<1>(dm, a, i, rt, Wt, Op, s_orow, s_ozs, s_fk, tb, (i + 1 < a.n) ? Ms : nullptr);
            }
        }
        __threadfence_block();
        __syncthreads();   // [C] the fill of the next pivot block is complete
        PP_ADD(2, 0); PP_ADD(17, 256);
        // ---- S3: panel warps: arrow rows of node i   ||   pivot group: pivot block of node i + 1
        if (!pivot_grp) {
            const int cnt = a.ar_hi - a.ar_lo;
            const int ntile = (cnt + 15) >> 4;
            for (int tl = warp - gnw; tl < ntile; tl += 16 - gnw) {
                RowTile rt;
                rt.seg = 1; rt.pr0 = xo + a.ar_lo + (tl << 4); rt.cr = (cnt - (tl << 4) < 16) ? (cnt - (tl << 4)) : 16; rt.xo = xo; rt.has_x = has_x;
                stream_row_tile<2>(dm, a, i, rt, Wt, Op, s_orow, s_ozs, s_fk, tb);
            }
            // idle until the pivot group is done: pull the sources of the next nodes towards L2 (pivot block of node i + 2 for the
            // rows of O of node i + 1, arrow rows and coupling block of node i + 1)
            {
                const int pt = threadIdx.x - gthreads, pn = blockDim.x - gthreads;
                if (i + 2 < a.n) { const double* Dn = a.Dsrc + (src + 2) * n * n; for (int e = pt * 16; e < n * n; e += pn * 16) prefetch_l2(Dn + e); }
                if (i + 1 < a.n) {
                    const double* An = a.Asrc + (src + 1) * ns1 * n; for (int e = pt * 16; e < ns1 * n; e += pn * 16) prefetch_l2(An + e);
                    const double* On = a.Osrc + (src + 1) * nc * nc; for (int e = pt * 16; e < nc * nc; e += pn * 16) prefetch_l2(On + e);
                }
            }
            __threadfence_block();
            PP_ADD(6, 256);
        } else if (i + 1 < a.n) {
            factor_node(i + 1);
            PP_ADD(4, 0);
        }
    }
#undef GSYNC
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc512(s_tmem);
}

// Dense levels (the separators of the level below: dense coupling blocks, 1..3 nodes per chunk, the arrow rows of a chunk split over
// gridDim.y CTAs): the same building blocks without the cross-keyframe pipeline -- plain Cholesky by all 16 warps, W^T by block
// substitution into the tile layout, then the dense coupling block takes the pivot buffer's place in shared memory and every panel
// row tile (rows of O and this CTA's arrow rows) is streamed through Tensor Memory by one warp.  Replaces cta_sweep<true> (fused
// Cholesky + inverse, 64-row panel chunks through shared memory, O streamed in slabs: ~135 us per dense node).
__device__ void cta_sweep_dense_stream(const Dims& dm, const SweepArgs& a, double* sm) {
    const int n = dm.ntp, nc = dm.nc, ns1 = dm.ns1, ld = dm.ld, hmax = dm.hmax;
    const long long Fsz = (long long)(nc + hmax) * nc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    double* Ms = sm;
    double* Wt = Ms + (long long)n * ld;
    __shared__ unsigned s_tmem;
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_abort = 0;
    if (warp == 0) tmem_alloc512(&s_tmem);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const unsigned tb = s_tmem + ((unsigned)(32 * (warp & 3)) << 16) + (unsigned)(128 * (warp >> 2));
#ifdef B200_PHASE_CLOCKS
    unsigned long long dd_t = clock64();
#define DD_ADD(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) { unsigned long long n_ = clock64(); g_phase_clk[i] += n_ - dd_t; dd_t = n_; } } while (0)
#else
#define DD_ADD(i)
#endif
    for (int i = 0; i < a.n; i++) {
        const long long src = a.src0 + i;
        const long long st = a.st_idx ? a.st_idx[i] : src;
        const bool has_x = nc > 0 && ((i + 1 < a.n) || a.has_right);
        const int xo = has_x ? nc : 0;
        const double* Fin = a.F + (long long)(i & 1) * Fsz;
        // ---- pivot block M = D + lambda I - fill (lower triangle; zeros above), fill of the previous node from global memory (L2)
        {
            const double* Dg = a.Dsrc + src * n * n;
            for (int r0 = warp; r0 < n; r0 += 4 * nwarps) {   // 4 rows x 2 column pairs per lane requested before the first use
                double2 d[4][2], f[4][2];
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int r = r0 + q * nwarps, c = 2 * lane + 64 * u;
                        d[q][u] = make_double2(0.0, 0.0); f[q][u] = make_double2(0.0, 0.0);
                        if (r < n && c < n && c <= r) {
                            d[q][u] = ld2(Dg + (long long)r * n + c);
                            if (i > 0) f[q][u] = ld2(Fin + (long long)r * nc + c);
                        }
                    }
#pragma unroll
                for (int q = 0; q < 4; q++)
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int r = r0 + q * nwarps, c = 2 * lane + 64 * u;
                        if (r >= n || c >= n) continue;
                        double2 v = make_double2(0.0, 0.0);
                        if (c <= r) {
                            v.x = d[q][u].x - f[q][u].x; v.y = d[q][u].y - f[q][u].y;
                            if (c == r) v.x += a.lambda;
                            if (c + 1 == r) v.y += a.lambda;
                            if (c + 1 > r) v.y = 0.0;
                        }
                        st2(Ms + (long long)r * ld + c, v.x, v.y);
                    }
            }
        }
        __syncthreads();
        DD_ADD(18);
        const int fail = cta_chol_inv(Ms, ld, n, -1, 0, 0, true);
        DD_ADD(19);
        if (fail) {
            if (threadIdx.x == 0) atomicCAS(a.status, 0, (int)(st + 1));
            break;
        }
        // ---- W^T = L^-T: tile layout + global copy
        {
            double* Wg = a.Wst + st * n * n;
            const int nbt = (n + 15) >> 4;
            for (int task = warp; task < 2 * nbt; task += nwarps) build_wt_rows(Ms, ld, n, task >> 1, task & 1, Wt, a.lead ? Wg : nullptr);
        }
        __syncthreads();
        DD_ADD(20);
        DD_ADD(21);
        // ---- panel row tiles, two phases.  A: y = p W^T of every tile (rows of O first: their y is X = O W^T, written to the pivot
        // buffer, which is free now).  B: fill = y X^T.  A warp keeps the y of its tile in Tensor Memory between the phases.
        {
            const int nt0 = has_x ? (nc + 15) >> 4 : 0;
            const int cnt = a.ar_hi - a.ar_lo;
            const int nt1 = (cnt + 15) >> 4;
            auto tile_of = [&](int tl, RowTile& rt) {
                rt.xo = xo; rt.has_x = has_x;
                if (tl < nt0) { rt.seg = 0; rt.pr0 = tl << 4; rt.cr = (nc - rt.pr0 < 16) ? (nc - rt.pr0) : 16; }
                else { const int u = tl - nt0; rt.seg = 1; rt.pr0 = xo + a.ar_lo + (u << 4); rt.cr = (cnt - (u << 4) < 16) ? (cnt - (u << 4)) : 16; }
            };
            for (int base = 0; base < nt0 + nt1; base += nwarps) {   // (rounds of one tile per warp; the rows of O are all in the first one)
                const int tl = base + warp;
                RowTile rt;
                if (tl < nt0 + nt1) {
                    tile_of(tl, rt);
                    stream_row_tile<2, true>(dm, a, i, rt, Wt, nullptr, nullptr, nullptr, nullptr, tb, has_x ? Ms : nullptr);
                }
                __syncthreads();   // X complete (first round)
                if (has_x && tl < nt0 + nt1) dense_fill_tile<2>(dm, a, i, rt, Ms, tb);
                if (base + nwarps < nt0 + nt1) __syncthreads();
            }
        }
        __threadfence_block();
        __syncthreads();
        DD_ADD(22);
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc512(s_tmem);
}

}  // namespace b200d
