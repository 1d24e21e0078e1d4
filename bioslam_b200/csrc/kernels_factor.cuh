// Factor-level kernels: linearization (residual + whitened Jacobian per factor instance), error-only evaluation,
// deterministic cost reduction, linear-model error (GaussianFactorGraph::error(delta)) and retraction.
//
// Layout in HBM (DESIGN.md "data layout"): keyframe-major SoA.  Time-variable values tv[component][Kld], per-keyframe
// constants cst[component][Kld]: consecutive lanes (consecutive keyframes of one factor slot) touch consecutive addresses.
// Whitened Jacobians / residuals are per-instance records Jbuf[slot][inst][rows*cols], rbuf[slot][inst][rows].
#pragma once
#include "factors.cuh"

namespace b200d {

struct DevSlot {
    int type, k_begin, k_end, n_inst;
    int nvars, rows, cols, const_off;   // const_off: first component row in cst (or -1)
    int var_kind[B200_MAX_FACTOR_VARS];
    int var_type[B200_MAX_FACTOR_VARS];
    int val_off[B200_MAX_FACTOR_VARS];  // time var: first component row in tv ; static var: offset in sv
    int tan_off[B200_MAX_FACTOR_VARS];  // internal tangent offset inside the time block / static block
    int col_off[B200_MAX_FACTOR_VARS];  // first Jacobian column of the variable
    int tan_dim[B200_MAX_FACTOR_VARS];
    long long J_off, r_off, e_off;      // offsets (doubles) into Jbuf / rbuf / ebuf
    double params[B200_MAX_FACTOR_PARAMS];
};

__device__ __forceinline__ int val_dim_of(int vt) { return vt == B200_VAR_POSE3 ? 12 : (vt == B200_VAR_BIAS6 ? 6 : 3); }

// gather the variable values of one factor instance into registers/local memory
__device__ __forceinline__ void load_vars(const DevSlot& s, int k, const double* __restrict__ tv, long long Kld,
                                          const double* __restrict__ sv, double (*xv)[12], const double** x) {
    for (int v = 0; v < s.nvars; v++) {
        const int n = val_dim_of(s.var_type[v]);
        if (s.var_kind[v] == B200_AT_STATIC) {
            for (int c = 0; c < n; c++) xv[v][c] = sv[s.val_off[v] + c];
        } else {
            const long long kk = k + (s.var_kind[v] == B200_AT_K1 ? 1 : 0);
            for (int c = 0; c < n; c++) xv[v][c] = tv[(long long)(s.val_off[v] + c) * Kld + kk];
        }
        x[v] = xv[v];
    }
}

// grid: (ceil(max_inst/128), n_slots), block 128.  One thread = one factor instance.
template <bool WJ>
__global__ void __launch_bounds__(128, 3) linearize_kernel(const DevSlot* __restrict__ slots, const double* __restrict__ tv,
                                                        const double* __restrict__ sv, const double* __restrict__ cst, long long Kld,
                                                        double* __restrict__ Jbuf, double* __restrict__ rbuf, double* __restrict__ ebuf,
                                                        int k_lo, int k_hi, int e_lo, const int* __restrict__ skip) {
    if (skip != nullptr && *skip != 0) return;   // (lambda trial no longer needed: see lm_judge_kernel)
    // keyframes [k_lo, k_hi) are evaluated; only [e_lo, k_hi) count towards this rank's cost (a sharded rank also linearizes
    // the keyframe left of its window, whose factors belong to the neighbour's cost)
    const DevSlot& s = slots[blockIdx.y];
    const int inst = blockIdx.x * blockDim.x + threadIdx.x;
    if (inst >= s.n_inst) return;
    const int k = s.k_begin + inst;
    if (k < k_lo || k >= k_hi) {  // outside this rank's window: contributes nothing to the local cost
        ebuf[s.e_off + inst] = 0.0;
        return;
    }
    double xv[B200_MAX_FACTOR_VARS][12];
    const double* x[B200_MAX_FACTOR_VARS];
    load_vars(s, k, tv, Kld, sv, xv, x);
    CstRef c;
    c.p = cst + (long long)(s.const_off < 0 ? 0 : s.const_off) * Kld + k;
    c.stride = Kld;
    double r[B200_MAX_FACTOR_ROWS];
    double J[WJ ? B200_MAX_FACTOR_ROWS * B200_MAX_FACTOR_COLS : 1];
    factor_eval<WJ>(s.type, s.params, c, x, r, J, s.cols);
    double* jout = WJ ? Jbuf + s.J_off + (long long)inst * s.rows * s.cols : nullptr;
    const bool j_written = factor_whiten<WJ>(s.type, s.params, c, r, J, s.cols, jout);
    double e = 0.0;
    for (int i = 0; i < s.rows; i++) e += r[i] * r[i];
    ebuf[s.e_off + inst] = (k >= e_lo) ? 0.5 * e : 0.0;
    if (WJ) {
        // per-instance records (rows*cols then rows): the consumer (assemble_kernel) copies one keyframe's records as
        // contiguous chunks; each thread streams its own record, full 32-byte sectors
        double* rp = rbuf + s.r_off + (long long)inst * s.rows;
        for (int i = 0; i < s.rows; i++) rp[i] = r[i];
        if (!j_written) {
            const int ne = s.rows * s.cols;
            for (int i = 0; i < ne; i++) jout[i] = J[i];
        }
    }
}

// One factor instance from explicit inputs (parity tests over every factor type): thread 0 evaluates residual + Jacobian of
// `type` exactly as linearize_kernel does (same factor_eval / factor_whiten device functions).  x: the variables' values
// concatenated in factor order; consts: the per-keyframe constant record (stride 1).
__global__ void factor_eval_kernel(int type, int nvars, const int* __restrict__ var_type, int cols, const double* __restrict__ params,
                                   const double* __restrict__ consts, const double* __restrict__ xcat, int whiten, double* __restrict__ r_out,
                                   double* __restrict__ J_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double xv[B200_MAX_FACTOR_VARS][12];
    const double* x[B200_MAX_FACTOR_VARS];
    int o = 0;
    for (int v = 0; v < nvars; v++) {
        const int n = val_dim_of(var_type[v]);
        for (int c = 0; c < n; c++) xv[v][c] = xcat[o + c];
        x[v] = xv[v];
        o += n;
    }
    double prm[B200_MAX_FACTOR_PARAMS];
    for (int i = 0; i < B200_MAX_FACTOR_PARAMS; i++) prm[i] = params[i];
    CstRef c; c.p = consts; c.stride = 1;
    double r[B200_MAX_FACTOR_ROWS];
    double J[B200_MAX_FACTOR_ROWS * B200_MAX_FACTOR_COLS];
    const int rows = factor_rows(type);
    for (int i = 0; i < rows * cols; i++) J[i] = 0.0;
    factor_eval<true>(type, prm, c, x, r, J, cols);
    bool written = false;
    if (whiten) written = factor_whiten<true>(type, prm, c, r, J, cols, J_out);
    for (int i = 0; i < rows; i++) r_out[i] = r[i];
    if (!written) for (int i = 0; i < rows * cols; i++) J_out[i] = J[i];
}

// deterministic sum of n doubles over B200_REDUCE_BLOCKS CTAs: fixed strided partials per thread, fixed shared-memory tree per CTA,
// and the CTA that finishes last (ticket counter) adds the per-CTA partials in index order -- the result does not depend on which CTA
// that is.  partial: B200_REDUCE_BLOCKS doubles, ticket: one int (zero before the first launch; reset by the last CTA).
#define B200_REDUCE_BLOCKS 64
#define B200_REDUCE_THREADS 256
__global__ void __launch_bounds__(B200_REDUCE_THREADS) reduce_sum_kernel(const double* __restrict__ v, long long n, double* __restrict__ out,
                                                                         const int* __restrict__ skip, double* partial, int* ticket) {
    __shared__ double sh[B200_REDUCE_THREADS];
    __shared__ int s_last;
    if (skip != nullptr && *skip != 0) return;
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * B200_REDUCE_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * B200_REDUCE_THREADS) acc += v[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = B200_REDUCE_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = sh[0];
        __threadfence();
        s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        double s = 0.0;
        for (int b = 0; b < (int)gridDim.x; b++) s += ((volatile double*)partial)[b];
        *out = s;
        *ticket = 0;
    }
}

// linear.error(delta) = sum over all factor rows of 0.5 (J[row] . delta + r[row])^2 on the stored linearization; delta in internal order:
// dtime[K][ntp], dstat[ns].  One thread per factor ROW (15 per IMU factor instance; a thread per instance, with the step gathered into
// a local array, took 0.12 ms at K = 6000): erows[row], rows numbered like rbuf; the caller sums them (reduce_sum_kernel).
__global__ void __launch_bounds__(128) model_error_rows_kernel(const DevSlot* __restrict__ slots, const double* __restrict__ Jbuf,
                                                               const double* __restrict__ rbuf, const double* __restrict__ dtime, int ntp,
                                                               const double* __restrict__ dstat, double* __restrict__ erows, int k_lo, int k_hi,
                                                               const int* __restrict__ skip) {
    if (skip != nullptr && *skip != 0) return;
    const DevSlot& s = slots[blockIdx.y];
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)s.n_inst * s.rows) return;
    const int inst = (int)(t / s.rows);
    const int k = s.k_begin + inst;
    const long long ro = s.r_off + t;
    if (k < k_lo || k >= k_hi) { erows[ro] = 0.0; return; }
    const double* jp = Jbuf + s.J_off + t * s.cols;
    double a = rbuf[ro];
    for (int v = 0; v < s.nvars; v++) {
        const double* src = (s.var_kind[v] == B200_AT_STATIC)
                                ? dstat + s.tan_off[v]
                                : dtime + (long long)(k + (s.var_kind[v] == B200_AT_K1 ? 1 : 0)) * ntp + s.tan_off[v];
        const double* jc = jp + s.col_off[v];
        for (int j = 0; j < s.tan_dim[v]; j++) a += jc[j] * src[j];
    }
    erows[ro] = 0.5 * a * a;
}

struct DevVar {
    int type, val_off, tan_off;  // val_off: component row (time) or offset (static); tan_off internal
};

// retraction of every variable: thread = (keyframe, time variable) ; static variables by block y==1.
__global__ void __launch_bounds__(128) retract_kernel(const DevVar* __restrict__ tvars, int ntv, const DevVar* __restrict__ svars, int nsv,
                                                      int K, long long Kld, const double* __restrict__ tv, const double* __restrict__ sv,
                                                      const double* __restrict__ dtime, int ntp, const double* __restrict__ dstat,
                                                      double* __restrict__ tv_out, double* __restrict__ sv_out, const int* __restrict__ skip) {
    if (skip != nullptr && *skip != 0) return;
    if (blockIdx.y == 1) {
        const int v = blockIdx.x * blockDim.x + threadIdx.x;
        if (v >= nsv) return;
        const DevVar var = svars[v];
        const int n = val_dim_of(var.type);
        double xin[12], xout[12], d[6];
        for (int c = 0; c < n; c++) xin[c] = sv[var.val_off + c];
        const int td = (var.type == B200_VAR_UNIT3) ? 2 : (var.type == B200_VAR_VEC3 ? 3 : 6);
        for (int c = 0; c < td; c++) d[c] = dstat[var.tan_off + c];
        switch (var.type) {
        case B200_VAR_POSE3: pose3_retract(xout, xin, d); break;
        case B200_VAR_UNIT3: unit3_retract(xout, xin, d); break;
        default: for (int c = 0; c < n; c++) xout[c] = xin[c] + d[c]; break;
        }
        for (int c = 0; c < n; c++) sv_out[var.val_off + c] = xout[c];
        return;
    }
    // time variables: blockIdx.z = variable, x-dimension = keyframes (coalesced)
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.z;
    if (k >= K || v >= ntv) return;
    const DevVar var = tvars[v];
    const int n = val_dim_of(var.type);
    double xin[12], xout[12], d[6];
    for (int c = 0; c < n; c++) xin[c] = tv[(long long)(var.val_off + c) * Kld + k];
    const int td = (var.type == B200_VAR_UNIT3) ? 2 : (var.type == B200_VAR_VEC3 ? 3 : 6);
    for (int c = 0; c < td; c++) d[c] = dtime[(long long)k * ntp + var.tan_off + c];
    switch (var.type) {
    case B200_VAR_POSE3: pose3_retract(xout, xin, d); break;
    case B200_VAR_UNIT3: unit3_retract(xout, xin, d); break;
    default: for (int c = 0; c < n; c++) xout[c] = xin[c] + d[c]; break;
    }
    for (int c = 0; c < n; c++) tv_out[(long long)(var.val_off + c) * Kld + k] = xout[c];
}

// [K][dim] (host/ABI layout) <-> [dim][Kld] (device SoA)
__global__ void transpose_in_kernel(const double* __restrict__ src, int K, int dim, long long Kld, double* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)K * dim) return;
    const int c = (int)(i / K), k = (int)(i % K);
    dst[(long long)c * Kld + k] = src[(long long)k * dim + c];
}
__global__ void transpose_out_kernel(const double* __restrict__ src, int K, int dim, long long Kld, double* __restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)K * dim) return;
    const int k = (int)(i / dim), c = (int)(i % dim);
    dst[(long long)k * dim + c] = src[(long long)c * Kld + k];
}

}  // namespace b200d
