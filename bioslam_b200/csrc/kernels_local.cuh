// Keyframe-local dofs (the angular velocities of the 7-IMU graph: SURVEY.md section 7 step 6a, section 8(e) "ang-vel is not
// time-coupled") are eliminated for ALL keyframes in one parallel kernel before the sequential chain sweep, so the chain
// pivot shrinks from ntp (126) to the time-chained dofs only (105 -> 106 with padding):
//
//   [ D_ll + lambda I   D_lc   A_l^T ] [d_l]   [-g_l]        L L^T = D_ll + lambda I,   V = [D_cl ; A_l ; -g_l^T] L^-T
//   [ D_cl              D_cc   A_c^T ] [d_c] = [-g_c]   ->   Dc = D_cc + lambda I - V_c V_c^T,  Ac = [A_c ; -g_c^T] - V_a V_c^T,
//   [ A_l               A_c    S     ] [d_s]   [-g_s]        Sc = S_k - V_a V_a^T            (V_c: first nc rows, V_a: last ns1)
//
// and recovered after the chain back-substitution:  d_l = -L^-T ( V_c^T d_c + V_a^T [d_s, -1] )   (L stored, back substitution).
// L and V come out of ONE right-looking elimination of the stacked block [D_ll + lambda I ; D_cl ; A_l ; -g_l^T] (a thread per row);
// the rank-nl updates run on the fp64 tensor pipe (same 16 x 32 DMMA warp tile as the sweep).  The kernel is latency-bound (0.65 ms
// per trial against ~0.2 ms of HBM traffic): what is exposed is one DRAM round trip per half tile of the epilogue and the nl pivots.
#pragma once
#include "kernels_solve.cuh"

namespace b200d {

#define B200_LE_THREADS 256
#define B200_LE_MAXNL 24

struct LocalElim {
    int nl, nc, ncp, ns1, ntp;      // local dofs, chain dofs (incl. the padding dof of an odd nt), chain dim of the solver (even), statics + rhs
    int ldl;                        // shared-memory stride of the nl-wide panels (== 4 mod 8: conflict-free DMMA operand loads)
    const double *D, *A, *S;        // assembled blocks: [K][ntp*ntp] (lower), [K][ns1*ntp], [K][ns1*ns1]
    double *Dc, *Ac, *Sc;           // reduced blocks: [K][ncp*ncp] (lower), [K][ns1*ncp], [K][ns1*ns1]
    double *V, *Wl;                 // [K][(nc+ns1)*nl] forward-substituted local panel, [K][nl*nl] Cholesky factor L of the local block (lower)
};
__host__ __device__ inline int le_ldl(int nl) { int l = nl; while ((l & 7) != 4) l++; return l; }
__host__ __device__ inline long long le_smem_doubles(int nl, int nc, int ns1) { return (long long)(2 * nl + nc + ns1) * le_ldl(nl); }

// grid = keyframes [k_first, k_first + gridDim.x)
// fresh (device flag, may be null = always): non-zero on the first lambda trial after a linearization -> every reduced entry is written.
// On the later trials of the same linearization only lambda has changed: V V^T touches the entries whose row AND column couple to a
// local dof (6 of every 15 chain dofs of the 7-IMU graph: 16 % of Dc) plus the diagonals -- everything else still holds the right value
// from the first trial and is neither read nor written (the kernel is bound by exactly that traffic).
__global__ void __launch_bounds__(B200_LE_THREADS, 3) local_elim_kernel(LocalElim le, int k_first, const double* __restrict__ lambda_p, int* status,
                                                                        const int* __restrict__ skip, const int* __restrict__ fresh) {
    B200_DYN_SMEM(double, sm);
    if (skip && *skip) return;
    const bool full = (fresh == nullptr) || (*fresh != 0);
    __shared__ unsigned char s_row[512];   // panel row r (chain rows, then statics + rhs) has a non-zero row in V
    const int nl = le.nl, nc = le.nc, ncp = le.ncp, ns1 = le.ns1, ntp = le.ntp, ldl = le.ldl, hp = nc + ns1;
    const long long k = k_first + (long long)blockIdx.x;
    const double lambda = *lambda_p;
    double* Ls = sm;                                   // nl x ldl: L (lower)
    double* Vs = Ls + 2LL * nl * ldl;                  // (nc + ns1) x ldl: V
    const double* Dk = le.D + k * ntp * ntp;
    const double* Ak = le.A + k * ns1 * ntp;
    const double* Sk = le.S + k * ns1 * ns1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    // pull everything this keyframe reads towards L2 now (lower triangle of D_k, A_k, S_k: ~150 KB): the epilogue below reads
    // D_cc / A_c / S_k entry by entry and would otherwise wait for DRAM tile after tile
    if (full) {
        for (int r = threadIdx.x; r < ntp; r += blockDim.x)
            for (int c = 0; c <= r; c += 16) prefetch_l2(Dk + (long long)r * ntp + c);
        for (int e = threadIdx.x * 16; e < ns1 * ntp; e += blockDim.x * 16) prefetch_l2(Ak + e);
        for (int e = threadIdx.x * 16; e < ns1 * ns1; e += blockDim.x * 16) prefetch_l2(Sk + e);
    }
    // Stacked elimination of [D_ll + lambda I ; D_cl ; A_l ; -g_l^T] ((nl + hp) x nl): ONE thread per row, the row in registers,
    // right-looking.  Pivot c: the rows of the local block publish their column-c entries (shared memory), one barrier, then EVERY
    // row scales its entry (x = r[c] / sqrt(d)) and updates its trailing entries r[c2] -= x * L[c2][c] -- the Cholesky factor L of the
    // local block and V = P L^-T (the forward substitution of the panel rows) come out of the same nl steps.
    // Thread map: warp 0, lane l < nl = row l of the local block; thread 32 + r = panel row r (hp <= blockDim.x - 32).
    __shared__ double s_col[2][32];
    const bool top = warp == 0 && lane < nl;
    const int prow = (int)threadIdx.x - 32;                      // panel row of this thread (warps >= 1)
    const bool pan = warp > 0 && prow < hp;
    double r[B200_LE_MAXNL];
    if (top) {
#pragma unroll
        for (int c = 0; c < B200_LE_MAXNL; c++) r[c] = (c <= lane) ? Dk[(long long)lane * ntp + c] + (c == lane ? lambda : 0.0) : 0.0;
        s_col[0][lane] = r[0];
    } else {
        const double* src = (prow < nc) ? Dk + (long long)(nl + (pan ? prow : 0)) * ntp : Ak + (long long)((pan ? prow : nc) - nc) * ntp;
#pragma unroll
        for (int c = 0; c < B200_LE_MAXNL; c++) r[c] = (pan && c < nl) ? __ldg(src + c) : 0.0;
    }
    __syncthreads();
    int fail = 0;
#pragma unroll
    for (int c = 0; c < B200_LE_MAXNL; c++) {
        if (c < nl) {
            const double* col = s_col[c & 1];
            double d = col[c];
            if (!(d > 0.0) || !(d < 1.0e300)) { fail = 1; d = 1.0; }
            const double rs = rsqrt(d);
            const double x = r[c] * rs;
            r[c] = x;
#pragma unroll
            for (int c2 = c + 1; c2 < B200_LE_MAXNL; c2++)
                if (c2 < nl) r[c2] = fma(-x, col[c2] * rs, r[c2]);
            if (c + 1 < nl) {
                if (top) s_col[(c + 1) & 1][lane] = r[c + 1];
                __syncthreads();
            }
        }
    }
    if (fail) { if (threadIdx.x == 0) atomicCAS(status, 0, (int)(k + 1)); return; }   // (the same d on every thread: uniform)
    if (top) {
#pragma unroll
        for (int c = 0; c < B200_LE_MAXNL; c++) if (c <= lane) Ls[lane * ldl + c] = r[c];
    } else if (pan) {
        bool nz = false;
#pragma unroll
        for (int c = 0; c < B200_LE_MAXNL; c++) { if (c < nl) { Vs[prow * ldl + c] = r[c]; nz |= (r[c] != 0.0); } }
        s_row[prow] = nz ? 1 : 0;
        for (int c = nl; c < ldl; c++) Vs[prow * ldl + c] = 0.0;
    }
    __syncthreads();
    {   // factors of the local block for the back-substitution: V and L (lower, row-major nl x nl)
        double* Vg = le.V + k * (long long)hp * nl;
        for (int e = threadIdx.x; e < hp * nl; e += blockDim.x) { const int r = e / nl, c = e - r * nl; Vg[e] = Vs[r * ldl + c]; }
        double* Wg = le.Wl + k * (long long)nl * nl;
        for (int e = threadIdx.x; e < nl * nl; e += blockDim.x) { const int r = e / nl, c = e - r * nl; Wg[e] = (c <= r) ? Ls[r * ldl + c] : 0.0; }
    }
    // rank-nl updates, 16 x 32 warp tiles: [0, nD) Dc (lower), [nD, nD + nA) Ac, then Sc (lower)
    const int rtc = (nc + 15) >> 4, ctc = (nc + 31) >> 5, rts = (ns1 + 15) >> 4;
    int nD = 0;
    for (int rt = 0; rt < rtc; rt++) nD += ((16 * rt + 15) >> 5) + 1;
    const int nA = rts * ctc;
    int nS = 0;
    for (int rt = 0; rt < rts; rt++) nS += ((16 * rt + 15) >> 5) + 1;
    double* Dck = le.Dc + k * (long long)ncp * ncp;
    double* Ack = le.Ac + k * (long long)ns1 * ncp;
    double* Sck = le.Sc + k * (long long)ns1 * ns1;
    const int mg = lane >> 2, mt = lane & 3;
    const double* Va = Vs + (long long)nc * ldl;
    for (int tile = warp; tile < nD + nA + nS; tile += nwarps) {
        int kind, rt = 0, ct = 0;
        if (tile < nD) { kind = 0; int t = tile; while (t >= ((16 * rt + 15) >> 5) + 1) { t -= ((16 * rt + 15) >> 5) + 1; rt++; } ct = t; }
        else if (tile < nD + nA) { kind = 1; rt = (tile - nD) / ctc; ct = (tile - nD) % ctc; }
        else { kind = 2; int t = tile - nD - nA; while (t >= ((16 * rt + 15) >> 5) + 1) { t -= ((16 * rt + 15) >> 5) + 1; rt++; } ct = t; }
        Acc acc;
        acc_zero(acc);
        const int m0 = rt << 4, c0 = ct << 5;
        if (kind == 0) warp_mma_nt(Vs, ldl, m0, nc, Vs, ldl, c0, nc, 0, nl, acc);
        else if (kind == 1) warp_mma_nt(Va, ldl, m0, ns1, Vs, ldl, c0, nc, 0, nl, acc);
        else warp_mma_nt(Va, ldl, m0, ns1, Va, ldl, c0, ns1, 0, nl, acc);
        // epilogue: reduced = source - acc.  The source entries of half a tile (8 per lane) are loaded in ONE batch before the first
        // store: source and destination pointers may alias as far as the compiler knows, so a load written after a store waits for it
        // -- entry by entry that is 16 exposed L2 / DRAM round trips per tile instead of 2.
        const int R = (kind == 0) ? nc : ns1;                       // rows of this tile's matrix
        const int rb = (kind == 0) ? 0 : nc, cb = (kind == 2) ? nc : 0;   // s_row offsets of its rows / columns
        const bool tri = kind != 1;                                 // lower triangle only
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int r = m0 + 8 * i + mg;
            const bool rok = r < R;
            const int rr = rok ? r : 0;
            const bool rnz = s_row[rb + rr] != 0;
            const double* srcp = (kind == 0) ? Dk + (long long)(nl + rr) * ntp + nl : (kind == 1) ? Ak + (long long)rr * ntp + nl : Sk + (long long)rr * ns1;
            double* dstp = (kind == 0) ? Dck + (long long)rr * ncp : (kind == 1) ? Ack + (long long)rr * ncp : Sck + (long long)rr * ns1;
            const int cend = tri ? rr + 1 : nc;
            double src[4][2];
            bool w[4][2];
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int c = c0 + 8 * j + 2 * mt + e;
                    bool ok = rok && c < cend;
                    if (ok && !full) ok = (kind == 0 && r == c) || (rnz && s_row[cb + c] != 0);
                    w[j][e] = ok;
                    src[j][e] = ok ? __ldg(srcp + c) : 0.0;
                }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int c = c0 + 8 * j + 2 * mt + e;
                    if (w[j][e]) dstp[c] = src[j][e] + ((kind == 0 && r == c) ? lambda : 0.0) - acc[i][j][e];
                }
            }
        }
    }
    // padding dof of the solver's chain block (ncp > nc): identity
    if (threadIdx.x == 0 && ncp > nc) Dck[(long long)nc * ncp + nc] = 1.0;
}

// d_l = -L^-T ( V_c^T d_c + V_a^T [d_s, -1] ) for every keyframe of [k_first, ...), and the full step in the internal order
// [local | chain] that retract_kernel / model_error_rows_kernel consume.  delta_c: [K][ncp] (solver), delta: [K][ntp].
__global__ void __launch_bounds__(128) local_backsub_kernel(LocalElim le, int k_first, int k_end, const double* __restrict__ delta_c,
                                                            const double* __restrict__ dstat, double* __restrict__ delta, const int* __restrict__ status,
                                                            const int* __restrict__ skip) {
    if (skip && *skip) return;
    if (*status != 0) return;
    const int nl = le.nl, nc = le.nc, ncp = le.ncp, ns1 = le.ns1, ntp = le.ntp, hp = nc + ns1;
    const long long k = k_first + (long long)blockIdx.x;
    if (k >= k_end) return;
    __shared__ double x[512 + 256], part[4][32], t[32];
    for (int j = threadIdx.x; j < nc; j += blockDim.x) x[j] = delta_c[k * ncp + j];
    for (int a = threadIdx.x; a < ns1; a += blockDim.x) x[nc + a] = (a < ns1 - 1) ? dstat[a] : -1.0;
    __syncthreads();
    const double* Vg = le.V + k * (long long)hp * nl;
    const int c = threadIdx.x & 31, q = threadIdx.x >> 5;
    double s = 0.0;
    if (c < nl) for (int r = q; r < hp; r += 4) s = fma(Vg[(long long)r * nl + c], x[r], s);
    part[q][c] = s;
    __syncthreads();
    if (threadIdx.x < (unsigned)nl) t[threadIdx.x] = (part[0][threadIdx.x] + part[1][threadIdx.x]) + (part[2][threadIdx.x] + part[3][threadIdx.x]);
    __syncthreads();
    double* dk = delta + k * ntp;
    if (threadIdx.x < 32) {
        // d_l = -L^-T t: backward substitution with the transpose, lane l holds entry l and COLUMN l of L (loaded before the chain
        // starts: rows of L are read coalesced, no memory access sits between two dependent steps)
        const double* Lg = le.Wl + k * (long long)nl * nl;
        const int l = threadIdx.x;
        double Lc[B200_LE_MAXNL];
#pragma unroll
        for (int i = 0; i < B200_LE_MAXNL; i++) Lc[i] = (i < nl && l <= i) ? Lg[(long long)i * nl + l] : 0.0;
        double v = (l < nl) ? t[l] : 0.0;
#pragma unroll
        for (int i = B200_LE_MAXNL - 1; i >= 0; i--) {
            if (i < nl) {
                const double yi = __shfl_sync(0xffffffffu, v, i) / __shfl_sync(0xffffffffu, Lc[i], i);
                if (l == i) v = yi;
                else if (l < i) v = fma(-Lc[i], yi, v);
            }
        }
        if (l < nl) dk[l] = -v;
    }
    for (int j = threadIdx.x; j < nc; j += blockDim.x) dk[nl + j] = x[j];
}

}  // namespace b200d
