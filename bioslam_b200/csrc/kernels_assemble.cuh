// Hessian / gradient accumulation (north_star subsystem 2): per-keyframe dense blocks of J^T J built in shared memory from the
// whitened per-factor Jacobians, one CTA per (keyframe, half):
//   half 0:  D_k  (ntp x ntp, time block of keyframe k)            + Sx_k ((ns+1) x (ns+1): static block and -g_static)
//   half 1:  O_k  (nc x nc, chain part of keyframe k+1 vs k)        + Ax_k ((ns+1) x ntp: static rows, last row = -g_k)
// Every block is owned by exactly one CTA and slots are visited in table order, so the sums are deterministic
// (no atomics).  Internal tangent order inside a keyframe: [keyframe-local variables (nl) | time-chained variables (nc)].
#pragma once
#include "kernels_factor.cuh"

namespace b200d {

struct Dims {
    int K;
    long long Kld;
    int nt, ntp, nl, nc, ns, ns1;  // ntp = nt padded to even, nc = ntp - nl, ns1 = ns + 1
    int ld;                        // shared-memory leading dimension (>= ntp, == 2 mod 4)
    int hmax;                      // ns1 + nc
    int sm_doubles;                // dynamic shared memory of the solver kernels, in doubles
};

// which (slot, instance) pairs touch keyframe k: instance k (variables AT_K, AT_STATIC, and the K1 x K cross block) and
// instance k-1 (its AT_K1 variables live at keyframe k).
__global__ void __launch_bounds__(256) assemble_kernel(Dims dm, const DevSlot* __restrict__ slots, int n_slots,
                                                       const double* __restrict__ Jbuf, const double* __restrict__ rbuf,
                                                       const int* __restrict__ klist, int k_first, int win_lo, int win_hi,
                                                       double* __restrict__ Dsrc, double* __restrict__ Osrc, double* __restrict__ Asrc,
                                                       double* __restrict__ Ssrc) {
    B200_DYN_SMEM(double, sm);
    const int k = klist ? klist[blockIdx.x] : k_first + (int)blockIdx.x;
    const int half = blockIdx.y;
    const int ntp = dm.ntp, nc = dm.nc, nl = dm.nl, ns = dm.ns, ns1 = dm.ns1;
    const int n1 = half == 0 ? ntp * ntp : nc * nc;
    const int n2 = half == 0 ? ns1 * ns1 : ns1 * ntp;
    double* B1 = sm;            // D or O
    double* B2 = sm + n1;       // Sx or Ax
    double* Js = B2 + n2;       // staged Jacobian (<= 15*30) followed by residual (<= 15)
    __shared__ unsigned char colvar[B200_MAX_FACTOR_COLS + 1];
    for (int i = threadIdx.x; i < n1 + n2; i += blockDim.x) sm[i] = 0.0;
    if (half == 0 && ntp > dm.nt && threadIdx.x == 0) B1[(ntp - 1) * ntp + ntp - 1] = 1.0;  // padding variable: identity
    __syncthreads();
    if (half == 0 && ntp > dm.nt && threadIdx.x == 0) B1[(ntp - 1) * ntp + ntp - 1] = 1.0;
    for (int si = 0; si < n_slots; si++) {
        const DevSlot& s = slots[si];
        bool has_k1 = false;
        for (int v = 0; v < s.nvars; v++) has_k1 |= (s.var_kind[v] == B200_AT_K1);
        for (int sel = 0; sel < 2; sel++) {
            const int ki = k - sel;
            if (sel == 1 && !has_k1) continue;
            if (ki < s.k_begin || ki >= s.k_end) continue;
            if (ki < win_lo || ki >= win_hi) continue;  // factor instance owned by another rank's window
            const int rows = s.rows, cols = s.cols;
            const long long n = s.n_inst, inst = ki - s.k_begin;
            __syncthreads();
            for (int i = threadIdx.x; i < rows * cols; i += blockDim.x) Js[i] = Jbuf[s.J_off + (long long)i * n + inst];
            for (int i = threadIdx.x; i < rows; i += blockDim.x) Js[rows * cols + i] = rbuf[s.r_off + (long long)i * n + inst];
            if (threadIdx.x < (unsigned)cols) {
                int v = 0;
                while (v + 1 < s.nvars && (int)threadIdx.x >= s.col_off[v + 1]) v++;
                colvar[threadIdx.x] = (unsigned char)v;
            }
            __syncthreads();
            const double* rr = Js + rows * cols;
            const int ne = cols * (cols + 1);
            for (int e = threadIdx.x; e < ne; e += blockDim.x) {
                const int ca = e / (cols + 1), cb = e % (cols + 1);
                const int va = colvar[ca];
                const int kda = s.var_kind[va];
                const bool a_static = kda == B200_AT_STATIC;
                const bool a_atk = (kda == B200_AT_K && sel == 0) || (kda == B200_AT_K1 && sel == 1);
                const bool a_atk1 = (kda == B200_AT_K1 && sel == 0);
                const int ta = s.tan_off[va] + (ca - s.col_off[va]);
                if (cb == cols) {  // gradient column: -J^T r
                    if (half == 1 && a_atk) {
                        double acc = 0;
                        for (int i = 0; i < rows; i++) acc += Js[i * cols + ca] * rr[i];
                        B2[ns * ntp + ta] -= acc;
                    } else if (half == 0 && a_static && sel == 0) {
                        double acc = 0;
                        for (int i = 0; i < rows; i++) acc += Js[i * cols + ca] * rr[i];
                        B2[ta * ns1 + ns] -= acc;
                        B2[ns * ns1 + ta] -= acc;
                    }
                    continue;
                }
                const int vb = colvar[cb];
                const int kdb = s.var_kind[vb];
                const bool b_static = kdb == B200_AT_STATIC;
                const bool b_atk = (kdb == B200_AT_K && sel == 0) || (kdb == B200_AT_K1 && sel == 1);
                const int tb = s.tan_off[vb] + (cb - s.col_off[vb]);
                double* dst = nullptr;
                if (half == 0) {
                    if (a_atk && b_atk) dst = B1 + ta * ntp + tb;
                    else if (a_static && b_static && sel == 0) dst = B2 + ta * ns1 + tb;
                } else {
                    if (a_atk1 && b_atk) dst = B1 + (ta - nl) * nc + (tb - nl);
                    else if (a_static && b_atk) dst = B2 + ta * ntp + tb;
                }
                if (!dst) continue;
                double acc = 0;
                for (int i = 0; i < rows; i++) acc += Js[i * cols + ca] * Js[i * cols + cb];
                *dst += acc;
            }
        }
    }
    __syncthreads();
    double* g1 = half == 0 ? Dsrc + (long long)k * ntp * ntp : Osrc + (long long)k * nc * nc;
    double* g2 = half == 0 ? Ssrc + (long long)k * ns1 * ns1 : Asrc + (long long)k * ns1 * ntp;
    for (int i = threadIdx.x; i < n1; i += blockDim.x) g1[i] = B1[i];
    for (int i = threadIdx.x; i < n2; i += blockDim.x) g2[i] = B2[i];
}

}  // namespace b200d
