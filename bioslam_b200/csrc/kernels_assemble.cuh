// Hessian / gradient accumulation (north_star subsystem 2): per-keyframe dense blocks of J^T J built from the whitened
// per-factor Jacobians (staged in shared memory), one CTA per keyframe, both halves from the same staged Jacobians:
//   half 0:  D_k  (ntp x ntp, time block of keyframe k; only the sub-blocks on or below the block diagonal: the solver reads
//            the lower triangle)                                   + Sx_k ((ns+1) x (ns+1): static block and -g_static)
//   half 1:  O_k  (nc x nc, stride nco, chain part of keyframe k+1 vs k)        + Ax_k ((ns+1) x ntp: static rows, last row = -g_k)
// Every block is owned by exactly one CTA and slots are visited in table order, so the sums are deterministic
// (no atomics).  Internal tangent order inside a keyframe: [keyframe-local variables (nl) | time-chained variables (nc)].
#pragma once
#include "kernels_factor.cuh"

namespace b200d {

struct Dims {
    int K;
    long long Kld;
    int nt, ntp, nl, nc, ns, ns1;  // ntp = nt padded to even, nc = ntp - nl, ns1 = ns + 1
    int nco;                       // row stride / dimension of the stored coupling blocks O_k (the SOLVER's chain dimension, >= nc)
    int ld;                        // shared-memory leading dimension (>= ntp, == 2 mod 4)
    int hmax;                      // ns1 + nc
    int sm_doubles;                // dynamic shared memory of the solver kernels, in doubles
    // solver: structure of the time-coupling block O (rows = chain dofs of keyframe k+1, columns = chain dofs of k)
    int cr;                        // panel rows resident in shared memory at a time (multiple of 16)
    int pipe;                      // level-0 sweep: use the keyframe-pipelined variant (kernels_sweep_pipe.cuh)
    int bs_os;                     // back-substitution: the coupling block of the node in flight is staged in shared memory
    int ow, owp;                   // packed width of a row of O (even) and its shared-memory stride (ow + 2)
    const int* orow;               // [nc][2] structural non-zero column range [c0, c1) of every row of O
    const int* ocol;               // [nc][2] structural non-zero row range [r0, r1) of every column of O
    const int* ozs;                // [nc] even start column (time-block coordinates) of the packed row
    const int* ogrp;               // [n_ogrp][3] groups (first row, stride, count <= 4) of rows of O that share the packed window ozs
    int n_ogrp;
};

// Assembly plan, built once per problem on the host (bioslam_b200.cu: build_asm_plan):
//   items  = the (slot, sel) factor instances that touch keyframe k: sel 0 = instance k (its AT_K / static variables and the
//            K1 x K cross block), sel 1 = instance k-1 (its AT_K1 variables live at keyframe k);
//   blocks = destination sub-blocks (one per pair of variables) of D / Sx (half 0) or O / Ax (half 1), each with the ordered
//            list of terms J[:, a]^T J[:, b] (or -J[:, a]^T r for the gradient rows) that sum into it.
// One warp owns one destination block at a time and accumulates all of its terms in registers, in table order: no
// atomics, no barriers between factors, deterministic.
struct AsmItem { int slot, sel, stage_off; };
// mode 0: J_a^T J_b ; 1: -J_a^T r (column block) ; 2: -(J_a^T r)^T (row block).  stage_off / rows / cols repeat the item's staging
// offset and its slot's Jacobian shape so that the term loop needs ONE metadata load per term (no dependent item -> slot chain)
struct __align__(16) AsmTerm { int item, ca, cb, mode, stage_off, rows, cols, pad; };
struct AsmBlock { int which, row, col, da, db, term_begin, term_end; };   // which: 0 = B1 (D or O), 1 = B2 (Sx or Ax)
// one destination ENTRY: block index, position inside the block (the kernel walks entries, one per thread: full lanes whatever the
// block shapes are; neighbouring lanes mostly belong to the same block and share its term list)
struct AsmElem { int blk; short ia, ib; };
struct AsmPlan {
    const AsmItem* items; int n_items;
    const AsmTerm* terms;
    const AsmBlock* blocks[2]; int n_blocks[2];
    const AsmElem* elems[2]; int n_elems[2];
    int stage_doubles;
};

// The dense blocks live in global memory and are ZEROED ONCE at problem creation: the set of destination sub-blocks is the
// same at every keyframe and every linearization, each one is rewritten completely (a block whose factors are all outside
// their keyframe range writes zeros), everything else stays zero.  So a CTA only needs shared memory for the staged
// Jacobians (several CTAs per SM) and writes ~4 k doubles per keyframe instead of the 38 k of the dense blocks.
__global__ void __launch_bounds__(512) assemble_kernel(Dims dm, const DevSlot* __restrict__ slots, AsmPlan pl,
                                                       const double* __restrict__ Jbuf, const double* __restrict__ rbuf,
                                                       int k_first, int win_lo, int win_hi,
                                                       double* __restrict__ Dsrc, double* __restrict__ Osrc, double* __restrict__ Asrc,
                                                       double* __restrict__ Ssrc) {
    B200_DYN_SMEM(double, St);  // staged Jacobians (rows*cols) + residuals (rows) of every item
    const int k = k_first + (int)blockIdx.x;
    const int ntp = dm.ntp, nc = dm.nco, ns1 = dm.ns1;   // nc: stride of O_k
    __shared__ unsigned char valid[256];
    // ---- which items exist at this keyframe; stage their Jacobians (contiguous per-instance records, issued in bulk)
    for (int it = threadIdx.x; it < pl.n_items; it += blockDim.x) {
        const AsmItem im = pl.items[it];
        const DevSlot& s = slots[im.slot];
        const int ki = k - im.sel;
        valid[it] = (ki >= s.k_begin && ki < s.k_end && ki >= win_lo && ki < win_hi) ? 1 : 0;
    }
    __syncthreads();
    for (int it = 0; it < pl.n_items; it++) {
        if (!valid[it]) continue;
        const AsmItem im = pl.items[it];
        const DevSlot& s = slots[im.slot];
        const long long inst = (k - im.sel) - s.k_begin;
        const int nj = s.rows * s.cols;
        double* dst = St + im.stage_off;
        const double* jg = Jbuf + s.J_off + inst * nj;
        const double* rg = rbuf + s.r_off + inst * s.rows;
        for (int i = threadIdx.x; i < nj; i += blockDim.x) cp_async8(dst + i, jg + i);
        if (threadIdx.x < (unsigned)s.rows) cp_async8(dst + nj + threadIdx.x, rg + threadIdx.x);
    }
    cp_async_wait_all();
    __syncthreads();
    if (ntp > dm.nt && threadIdx.x == 0) Dsrc[(long long)k * ntp * ntp + (ntp - 1) * ntp + ntp - 1] = 1.0;  // padding variable: identity
    // ---- one thread per destination entry, terms in table order (fixed summation order), result written straight to global
    // memory; both halves (D + Sx, then O + Ax) from the same staged Jacobians
    for (int half = 0; half < 2; half++) {
    const int ld1 = half == 0 ? ntp : nc, ld2 = half == 0 ? ns1 : ntp;
    double* g1 = half == 0 ? Dsrc + (long long)k * ntp * ntp : Osrc + (long long)k * nc * nc;
    double* g2 = half == 0 ? Ssrc + (long long)k * ns1 * ns1 : Asrc + (long long)k * ns1 * ntp;
    const AsmBlock* blocks = pl.blocks[half];
    const AsmElem* elems = pl.elems[half];
    for (int e = threadIdx.x; e < pl.n_elems[half]; e += blockDim.x) {
        const AsmElem el = elems[e];
        const AsmBlock bk = blocks[el.blk];
        double acc = 0.0;
        for (int t = bk.term_begin; t < bk.term_end; t++) {
            const AsmTerm tm = pl.terms[t];
            if (!valid[tm.item]) continue;
            const int rows = tm.rows, cols = tm.cols;
            const double* Js = St + tm.stage_off;
            if (tm.mode == 0) {
                const double* pa = Js + tm.ca + el.ia;
                const double* pb = Js + tm.cb + el.ib;
#pragma unroll 3
                for (int r = 0; r < rows; r++, pa += cols, pb += cols) acc = fma(*pa, *pb, acc);
            } else {
                const double* pj = Js + tm.ca + (tm.mode == 1 ? el.ia : el.ib);
                const double* rr = Js + rows * cols;
#pragma unroll 3
                for (int r = 0; r < rows; r++, pj += cols) acc = fma(-*pj, rr[r], acc);
            }
        }
        double* B = bk.which == 0 ? g1 : g2;
        const int ld = bk.which == 0 ? ld1 : ld2;
        B[(bk.row + el.ia) * ld + bk.col + el.ib] = acc;
    }
    }
}

}  // namespace b200d
