// ORACLE -- TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library; the product (bioslam_b200/) never does.
//
// CPU restatement of the hot path the CUDA library replaces:
//   graph.linearize / graph.error        -> linearize(), error()
//   LevenbergMarquardtOptimizer::iterate -> orc_lm_iterate()  (gtsam/nonlinear/LevenbergMarquardtOptimizer.cpp @4.2:
//                                           iterate(), tryLambda(), State::increaseLambda/decreaseLambda)
//   GaussianFactorGraph::optimize        -> solve(): dense block-tridiagonal + arrowhead Cholesky (ordering-independent
//                                           result, same normal equations J^T J + lambda I as buildDampedSystem)
//   Values::retract                      -> retract_all()
//   bioslam outer loops                  -> orc_optimize() (reference src/lowerBodyPoseEstimator.cpp:651-667,
//                                           src/imuPoseEstimator.cpp:330-359)
// PARITY UNPINNED for the GTSAM-owned pieces (GTSAM 4.2 is an un-vendored dependency that cannot be built here); the
// bioslam-owned factor arithmetic is pinned by the reference's own test method (analytic == numerical Jacobian,
// optimise-to-zero), see tests/test_oracle_*.py.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "factors.hpp"
#include "preint.hpp"
#include "jointangles.hpp"
#include "fastsolve.hpp"

using namespace orc;

struct orc_problem {
    int K = 0, ntv = 0, nsv = 0, cstride = 0;
    std::vector<int> tvtype, svtype;
    std::vector<b200_factor_slot> slots;
    std::vector<double> consts;
    std::vector<int> tvoff, tvtoff, svoff, svtoff;
    int tvd = 0, nt = 0, svd = 0, ns = 0;
    std::vector<double> tv, sv;      // current values
    std::vector<double> tv_try, sv_try;
    // linearization (external tangent order)
    std::vector<double> D, O, A, S, g, gs;   // D[K][nt*nt], O[K][nt*nt] (H[(k+1),(k)]), A[K][ns*nt], S[ns*ns]
    std::vector<std::vector<double>> Jw, rw; // per slot: [inst][rows*cols], [inst][rows]
    // factorization workspace
    std::vector<double> L, X, Y, Ls, z, zs;
    std::vector<double> dt_, ds_;
    // LM state
    b200_lm_params lm{};
    double err = 0, lambda = 0;
    int iterations = 0, total_trials = 0;
    int threads = 1;
    // solver selection: 0 = plain dense block sweep (solve(), the pinned restatement), 1 = orcfast::Solver (fastsolve.hpp)
    int fast = 0, fast_chunks = 16;
    bool fast_ready = false;
    orcfast::Solver fs;
};

static void var_ptrs(const orc_problem& p, const b200_factor_slot& s, int k, const std::vector<double>& tv, const std::vector<double>& sv,
                     const double** x) {
    for (int v = 0; v < s.nvars; v++) {
        if (s.var_kind[v] == B200_AT_STATIC) x[v] = sv.data() + p.svoff[s.var_index[v]];
        else {
            int kk = k + (s.var_kind[v] == B200_AT_K1 ? 1 : 0);
            x[v] = tv.data() + (size_t)kk * p.tvd + p.tvoff[s.var_index[v]];
        }
    }
}

static double graph_error(const orc_problem& p, const std::vector<double>& tv, const std::vector<double>& sv) {
    double total = 0;
    for (const auto& s : p.slots) {
        const FactorInfo& fi = factor_info(s.type);
        const int n = s.k_end - s.k_begin;
        // fixed blocks of 64 instances summed in block order: the result does not depend on the number of threads
        const int nblk = (n + 63) / 64;
        std::vector<double> part(nblk, 0.0);
#pragma omp parallel for num_threads(p.threads) schedule(static)
        for (int b = 0; b < nblk; b++) {
            double sub = 0;
            for (int i = b * 64; i < std::min(n, (b + 1) * 64); i++) {
                int k = s.k_begin + i;
                const double* x[B200_MAX_FACTOR_VARS];
                var_ptrs(p, s, k, tv, sv, x);
                const double* c = s.const_offset >= 0 ? p.consts.data() + (size_t)k * p.cstride + s.const_offset : nullptr;
                double r[B200_MAX_FACTOR_ROWS];
                factor_eval(s.type, s.params, c, x, r, nullptr);
                factor_whiten(s.type, s.params, c, r, nullptr);
                double e = 0;
                for (int j = 0; j < fi.rows; j++) e += r[j] * r[j];
                sub += 0.5 * e;
            }
            part[b] = sub;
        }
        for (int b = 0; b < nblk; b++) total += part[b];
    }
    return total;
}

static void linearize(orc_problem& p) {
    const int nt = p.nt, ns = p.ns, K = p.K;
    std::fill(p.D.begin(), p.D.end(), 0.0);
    std::fill(p.O.begin(), p.O.end(), 0.0);
    std::fill(p.A.begin(), p.A.end(), 0.0);
    std::fill(p.S.begin(), p.S.end(), 0.0);
    std::fill(p.g.begin(), p.g.end(), 0.0);
    std::fill(p.gs.begin(), p.gs.end(), 0.0);
    p.Jw.resize(p.slots.size());
    p.rw.resize(p.slots.size());
    // ---- whitened residuals and Jacobians of every factor instance
    for (size_t si = 0; si < p.slots.size(); si++) {
        const auto& s = p.slots[si];
        const FactorInfo& fi = factor_info(s.type);
        const int cols = factor_ncols(s.type), rows = fi.rows, n = s.k_end - s.k_begin;
        p.Jw[si].assign((size_t)n * rows * cols, 0.0);
        p.rw[si].assign((size_t)n * rows, 0.0);
#pragma omp parallel for num_threads(p.threads) schedule(static)
        for (int i = 0; i < n; i++) {
            int k = s.k_begin + i;
            const double* x[B200_MAX_FACTOR_VARS];
            var_ptrs(p, s, k, p.tv, p.sv, x);
            const double* c = s.const_offset >= 0 ? p.consts.data() + (size_t)k * p.cstride + s.const_offset : nullptr;
            double* r = p.rw[si].data() + (size_t)i * rows;
            double* J = p.Jw[si].data() + (size_t)i * rows * cols;
            factor_eval(s.type, s.params, c, x, r, J);
            factor_whiten(s.type, s.params, c, r, J);
        }
    }
    // ---- J^T J and J^T r gathered per DESTINATION keyframe kk (parallel over kk, fixed order inside: slots in order, for every
    // slot first the instance kk-1 (its variables at k+1), then the instance kk (its variables at k)).  Static x static terms
    // are summed per block of 64 keyframes and the blocks are added in order: thread-count independent.
    const int nblk = (K + 63) / 64;
    std::vector<double> Spart((size_t)nblk * ns * ns, 0.0), gspart((size_t)nblk * std::max(ns, 1), 0.0);
#pragma omp parallel for num_threads(p.threads) schedule(static)
    for (int blk = 0; blk < nblk; blk++) {
        double* Sb = Spart.data() + (size_t)blk * ns * ns;
        double* gsb = gspart.data() + (size_t)blk * std::max(ns, 1);
        for (int kk = blk * 64; kk < std::min(K, (blk + 1) * 64); kk++) {
            for (size_t si = 0; si < p.slots.size(); si++) {
                const auto& s = p.slots[si];
                const FactorInfo& fi = factor_info(s.type);
                const int cols = factor_ncols(s.type), rows = fi.rows;
                int coff[B200_MAX_FACTOR_VARS], cdim[B200_MAX_FACTOR_VARS];
                {
                    int o = 0;
                    for (int v = 0; v < fi.nvars; v++) { coff[v] = o; cdim[v] = kTanDim[fi.vartype[v]]; o += cdim[v]; }
                }
                for (int sel = 1; sel >= 0; sel--) {   // sel 1: instance kk-1 seen from its k+1 side ; sel 0: instance kk
                    const int k = kk - sel;
                    if (k < s.k_begin || k >= s.k_end) continue;
                    const int i = k - s.k_begin;
                    const double* r = p.rw[si].data() + (size_t)i * rows;
                    const double* J = p.Jw[si].data() + (size_t)i * rows * cols;
                    for (int a = 0; a < fi.nvars; a++) {
                        const int ka = s.var_kind[a];
                        const bool a_static = ka == B200_AT_STATIC;
                        const bool a_here = (ka == B200_AT_K && sel == 0) || (ka == B200_AT_K1 && sel == 1);
                        const bool a_next = (ka == B200_AT_K1 && sel == 0);
                        const int ta = a_static ? p.svtoff[s.var_index[a]] : p.tvtoff[s.var_index[a]];
                        if (a_here || (a_static && sel == 0)) {
                            for (int ia = 0; ia < cdim[a]; ia++) {
                                double acc = 0;
                                for (int rr = 0; rr < rows; rr++) acc += J[rr * cols + coff[a] + ia] * r[rr];
                                if (a_static) gsb[ta + ia] += acc;
                                else p.g[(size_t)kk * nt + ta + ia] += acc;
                            }
                        }
                        for (int b = 0; b < fi.nvars; b++) {
                            const int kb = s.var_kind[b];
                            const bool b_static = kb == B200_AT_STATIC;
                            const bool b_here = (kb == B200_AT_K && sel == 0) || (kb == B200_AT_K1 && sel == 1);
                            const int tb = b_static ? p.svtoff[s.var_index[b]] : p.tvtoff[s.var_index[b]];
                            double* dst; int ldd;
                            if (a_here && b_here) { dst = p.D.data() + ((size_t)kk * nt + ta) * nt + tb; ldd = nt; }
                            else if (a_static && b_static && sel == 0) { dst = Sb + (size_t)ta * ns + tb; ldd = ns; }
                            else if (a_next && b_here) { dst = p.O.data() + ((size_t)kk * nt + ta) * nt + tb; ldd = nt; }   // O[k]: rows (k+1), cols (k)
                            else if (a_static && b_here) { dst = p.A.data() + ((size_t)kk * ns + ta) * nt + tb; ldd = nt; }
                            else continue;
                            for (int ia = 0; ia < cdim[a]; ia++)
                                for (int ib = 0; ib < cdim[b]; ib++) {
                                    double acc = 0;
                                    for (int rr = 0; rr < rows; rr++) acc += J[rr * cols + coff[a] + ia] * J[rr * cols + coff[b] + ib];
                                    dst[(size_t)ia * ldd + ib] += acc;
                                }
                        }
                    }
                }
            }
        }
    }
    for (int blk = 0; blk < nblk; blk++) {
        for (int i = 0; i < ns * ns; i++) p.S[i] += Spart[(size_t)blk * ns * ns + i];
        for (int i = 0; i < ns; i++) p.gs[i] += gspart[(size_t)blk * std::max(ns, 1) + i];
    }
}

// in-place lower Cholesky of n x n (ld), returns false on non-positive pivot
static bool chol_lower(double* M, int n, int ld) {
    for (int j = 0; j < n; j++) {
        double d = M[j * ld + j];
        for (int l = 0; l < j; l++) d -= M[j * ld + l] * M[j * ld + l];
        if (!(d > 0.0) || !std::isfinite(d)) return false;
        double ljj = std::sqrt(d);
        M[j * ld + j] = ljj;
        for (int i = j + 1; i < n; i++) {
            double v = M[i * ld + j];
            const double* mi = M + i * ld; const double* mj = M + j * ld;
            for (int l = 0; l < j; l++) v -= mi[l] * mj[l];
            M[i * ld + j] = v / ljj;
        }
    }
    return true;
}
// rows: R (m x n, ld) <- R L^-T  (forward substitution per row)
static void trsm_rows(double* R, int m, int n, int ldr, const double* L, int ldl, int threads) {
#pragma omp parallel for num_threads(threads) schedule(static) if (m >= 16)
    for (int i = 0; i < m; i++) {
        double* x = R + (size_t)i * ldr;
        for (int j = 0; j < n; j++) {
            double v = x[j];
            const double* lj = L + (size_t)j * ldl;
            for (int l = 0; l < j; l++) v -= x[l] * lj[l];
            x[j] = v / lj[j];
        }
    }
}
// C (m x n, ldc) -= A (m x kd) B^T (n x kd)
static void gemm_nt_sub(double* C, int m, int n, int ldc, const double* A, int lda, const double* B, int ldb, int kd, int threads) {
#pragma omp parallel for num_threads(threads) schedule(static) if (m >= 16)
    for (int i = 0; i < m; i++) {
        const double* a = A + (size_t)i * lda;
        for (int j = 0; j < n; j++) {
            const double* b = B + (size_t)j * ldb;
            double acc = 0;
#pragma omp simd reduction(+ : acc)
            for (int l = 0; l < kd; l++) acc += a[l] * b[l];
            C[(size_t)i * ldc + j] -= acc;
        }
    }
}

// solve (H + lambda I) delta = -g with the block-tridiagonal + arrow structure. returns false if indeterminate.
static bool solve(orc_problem& p, double lambda, std::vector<double>& dt, std::vector<double>& ds) {
    const int nt = p.nt, ns = p.ns, K = p.K, th = p.threads;
    p.L.resize((size_t)K * nt * nt);
    p.X.resize((size_t)K * nt * nt);
    p.Y.resize((size_t)K * std::max(ns, 1) * nt);
    p.z.resize((size_t)K * nt);
    p.Ls.assign((size_t)ns * ns, 0.0);
    p.zs.assign(ns, 0.0);
    std::vector<double> Dacc((size_t)nt * nt), Aacc((size_t)std::max(ns, 1) * nt), b(nt);
    std::vector<double> Sacc(p.S);
    for (int i = 0; i < ns; i++) { Sacc[(size_t)i * ns + i] += lambda; p.zs[i] = -p.gs[i]; }
    std::vector<double> Dnext_fill((size_t)nt * nt, 0.0), Anext_fill((size_t)std::max(ns, 1) * nt, 0.0), bnext_fill(nt, 0.0);
    for (int k = 0; k < K; k++) {
        double* Lk = p.L.data() + (size_t)k * nt * nt;
        for (int i = 0; i < nt * nt; i++) Lk[i] = p.D[(size_t)k * nt * nt + i] - Dnext_fill[i];
        for (int i = 0; i < nt; i++) Lk[i * nt + i] += lambda;
        if (!chol_lower(Lk, nt, nt)) return false;
        double* Yk = p.Y.data() + (size_t)k * std::max(ns, 1) * nt;
        for (int i = 0; i < ns * nt; i++) Yk[i] = p.A[(size_t)k * ns * nt + i] - Anext_fill[i];
        trsm_rows(Yk, ns, nt, nt, Lk, nt, th);
        // rhs
        double* zk = p.z.data() + (size_t)k * nt;
        for (int i = 0; i < nt; i++) zk[i] = -p.g[(size_t)k * nt + i] - bnext_fill[i];
        trsm_rows(zk, 1, nt, nt, Lk, nt, 1);
        // static accumulations
        gemm_nt_sub(Sacc.data(), ns, ns, ns, Yk, nt, Yk, nt, nt, th);
        for (int i = 0; i < ns; i++) { double acc = 0; for (int l = 0; l < nt; l++) acc += Yk[(size_t)i * nt + l] * zk[l]; p.zs[i] -= acc; }
        if (k + 1 < K) {
            double* Xk = p.X.data() + (size_t)k * nt * nt;
            std::copy(p.O.begin() + (size_t)k * nt * nt, p.O.begin() + (size_t)(k + 1) * nt * nt, Xk);
            trsm_rows(Xk, nt, nt, nt, Lk, nt, th);
            std::fill(Dnext_fill.begin(), Dnext_fill.end(), 0.0);
            // fill = +X X^T etc. (stored positive, subtracted at the next step)
            std::vector<double>& F = Dnext_fill;
#pragma omp parallel for num_threads(th) schedule(static)
            for (int i = 0; i < nt; i++)
                for (int j = 0; j < nt; j++) {
                    double acc = 0;
                    const double* a = Xk + (size_t)i * nt; const double* bb = Xk + (size_t)j * nt;
#pragma omp simd reduction(+ : acc)
                    for (int l = 0; l < nt; l++) acc += a[l] * bb[l];
                    F[(size_t)i * nt + j] = acc;
                }
            for (int i = 0; i < ns; i++)
                for (int j = 0; j < nt; j++) {
                    double acc = 0;
                    for (int l = 0; l < nt; l++) acc += Yk[(size_t)i * nt + l] * Xk[(size_t)j * nt + l];
                    Anext_fill[(size_t)i * nt + j] = acc;
                }
            for (int j = 0; j < nt; j++) { double acc = 0; for (int l = 0; l < nt; l++) acc += Xk[(size_t)j * nt + l] * zk[l]; bnext_fill[j] = acc; }
        }
    }
    ds.assign(ns, 0.0);
    if (ns > 0) {
        p.Ls = Sacc;
        if (!chol_lower(p.Ls.data(), ns, ns)) return false;
        trsm_rows(p.zs.data(), 1, ns, ns, p.Ls.data(), ns, 1);
        // back: L^T ds = zs
        for (int i = ns - 1; i >= 0; i--) {
            double v = p.zs[i];
            for (int l = i + 1; l < ns; l++) v -= p.Ls[(size_t)l * ns + i] * ds[l];
            ds[i] = v / p.Ls[(size_t)i * ns + i];
        }
    }
    dt.assign((size_t)K * nt, 0.0);
    std::vector<double> t(nt);
    for (int k = K - 1; k >= 0; k--) {
        const double* Lk = p.L.data() + (size_t)k * nt * nt;
        const double* Yk = p.Y.data() + (size_t)k * std::max(ns, 1) * nt;
        const double* zk = p.z.data() + (size_t)k * nt;
        for (int i = 0; i < nt; i++) t[i] = zk[i];
        for (int s = 0; s < ns; s++) { double d = ds[s]; const double* y = Yk + (size_t)s * nt; for (int i = 0; i < nt; i++) t[i] -= y[i] * d; }
        if (k + 1 < K) {
            const double* Xk = p.X.data() + (size_t)k * nt * nt;
            const double* dn = dt.data() + (size_t)(k + 1) * nt;
            for (int j = 0; j < nt; j++) { double d = dn[j]; const double* x = Xk + (size_t)j * nt; for (int i = 0; i < nt; i++) t[i] -= x[i] * d; }
        }
        double* dk = dt.data() + (size_t)k * nt;
        for (int i = nt - 1; i >= 0; i--) {
            double v = t[i];
            for (int l = i + 1; l < nt; l++) v -= Lk[(size_t)l * nt + i] * dk[l];
            dk[i] = v / Lk[(size_t)i * nt + i];
        }
    }
    return true;
}

static void retract_var(int type, const double* x, const double* d, double* out) {
    switch (type) {
    case B200_VAR_POSE3: pose3_retract(out, x, d); break;
    case B200_VAR_VEC3: for (int i = 0; i < 3; i++) out[i] = x[i] + d[i]; break;
    case B200_VAR_BIAS6: for (int i = 0; i < 6; i++) out[i] = x[i] + d[i]; break;
    case B200_VAR_UNIT3: unit3_retract(out, x, d); break;
    }
}
static void retract_all(const orc_problem& p, const std::vector<double>& dt, const std::vector<double>& ds, std::vector<double>& tvn,
                        std::vector<double>& svn) {
    tvn.resize(p.tv.size()); svn.resize(p.sv.size());
#pragma omp parallel for num_threads(p.threads) schedule(static)
    for (int k = 0; k < p.K; k++)
        for (int v = 0; v < p.ntv; v++)
            retract_var(p.tvtype[v], p.tv.data() + (size_t)k * p.tvd + p.tvoff[v], dt.data() + (size_t)k * p.nt + p.tvtoff[v],
                        tvn.data() + (size_t)k * p.tvd + p.tvoff[v]);
    for (int v = 0; v < p.nsv; v++) retract_var(p.svtype[v], p.sv.data() + p.svoff[v], ds.data() + p.svtoff[v], svn.data() + p.svoff[v]);
}

// linear.error(delta) = 0.5 * sum ||J delta + r||^2 over the stored linearization (GaussianFactorGraph::error)
static double linear_error(const orc_problem& p, const std::vector<double>& dt, const std::vector<double>& ds, bool zero) {
    double total = 0;
    for (size_t si = 0; si < p.slots.size(); si++) {
        const auto& s = p.slots[si];
        const FactorInfo& fi = factor_info(s.type);
        const int cols = factor_ncols(s.type), rows = fi.rows, n = s.k_end - s.k_begin;
        const int nblk = (n + 63) / 64;   // fixed blocks, summed in order (thread-count independent)
        std::vector<double> part(nblk, 0.0);
#pragma omp parallel for num_threads(p.threads) schedule(static)
        for (int b = 0; b < nblk; b++) {
            double sub = 0;
            for (int i = b * 64; i < std::min(n, (b + 1) * 64); i++) {
                int k = s.k_begin + i;
                const double* r = p.rw[si].data() + (size_t)i * rows;
                const double* J = p.Jw[si].data() + (size_t)i * rows * cols;
                double dl[B200_MAX_FACTOR_COLS];
                int o = 0;
                for (int v = 0; v < fi.nvars; v++) {
                    int dim = kTanDim[fi.vartype[v]];
                    const double* src;
                    if (s.var_kind[v] == B200_AT_STATIC) src = ds.data() + p.svtoff[s.var_index[v]];
                    else src = dt.data() + (size_t)(k + (s.var_kind[v] == B200_AT_K1)) * p.nt + p.tvtoff[s.var_index[v]];
                    for (int j = 0; j < dim; j++) dl[o++] = zero ? 0.0 : src[j];
                }
                double e = 0;
                for (int rr = 0; rr < rows; rr++) {
                    double v = r[rr];
                    for (int c = 0; c < cols; c++) v += J[rr * cols + c] * dl[c];
                    e += v * v;
                }
                sub += 0.5 * e;
            }
            part[b] = sub;
        }
        for (int b = 0; b < nblk; b++) total += part[b];
    }
    return total;
}

// solver dispatch: the plain restatement or the structure-aware fast solver (same system, pinned against each other)
static bool solve_any(orc_problem& p, double lambda, std::vector<double>& dt, std::vector<double>& ds) {
    if (!p.fast) return solve(p, lambda, dt, ds);
    if (!p.fast_ready) {
        // time-chained dofs: variables of a factor that spans k and k+1; per chain row of O the first column that can be non-zero
        std::vector<char> chained(p.nt, 0);
        std::vector<int> ostart(p.nt, p.nt);
        for (const auto& s : p.slots) {
            bool spans = false;
            for (int v = 0; v < s.nvars; v++) spans |= (s.var_kind[v] == B200_AT_K1);
            if (!spans) continue;
            const FactorInfo& fi = factor_info(s.type);
            int lo = p.nt;
            for (int v = 0; v < s.nvars; v++) if (s.var_kind[v] == B200_AT_K) lo = std::min(lo, p.tvtoff[s.var_index[v]]);
            for (int v = 0; v < s.nvars; v++) {
                if (s.var_kind[v] == B200_AT_STATIC) continue;
                const int t0 = p.tvtoff[s.var_index[v]], d = kTanDim[fi.vartype[v]];
                for (int j = 0; j < d; j++) { chained[t0 + j] = 1; if (s.var_kind[v] == B200_AT_K1) ostart[t0 + j] = std::min(ostart[t0 + j], lo); }
            }
        }
        for (int i = 0; i < p.nt; i++) if (ostart[i] >= p.nt) ostart[i] = 0;
        // a row's start column must be a chain dof: move it down to the first chain dof at or after it (conservative: lower)
        for (int i = 0; i < p.nt; i++) { int e = ostart[i]; while (e > 0 && !chained[e]) e--; if (!chained[e]) e = 0; ostart[i] = e; }
        p.fs.setup(p.K, p.nt, p.ns, chained, ostart, p.fast_chunks);
        p.fast_ready = true;
    }
    dt.assign((size_t)p.K * p.nt, 0.0);
    ds.assign(p.ns, 0.0);
    return p.fs.solve(p.D.data(), p.O.data(), p.A.data(), p.g.data(), p.S.data(), p.gs.data(), lambda, p.threads, dt.data(), ds.data());
}

extern "C" {

orc_problem* orc_create(const b200_problem_desc* d) {
    orc_problem* p = new orc_problem();
    p->K = d->K; p->ntv = d->n_time_vars; p->nsv = d->n_static_vars; p->cstride = d->const_stride;
    p->tvtype.assign(d->time_var_type, d->time_var_type + d->n_time_vars);
    if (d->n_static_vars) p->svtype.assign(d->static_var_type, d->static_var_type + d->n_static_vars);
    p->slots.assign(d->slots, d->slots + d->n_slots);
    if (d->const_stride > 0) p->consts.assign(d->consts, d->consts + (size_t)d->K * d->const_stride);
    for (int t : p->tvtype) { p->tvoff.push_back(p->tvd); p->tvtoff.push_back(p->nt); p->tvd += kValDim[t]; p->nt += kTanDim[t]; }
    for (int t : p->svtype) { p->svoff.push_back(p->svd); p->svtoff.push_back(p->ns); p->svd += kValDim[t]; p->ns += kTanDim[t]; }
    p->tv.assign((size_t)p->K * p->tvd, 0.0); p->sv.assign(p->svd, 0.0);
    p->D.assign((size_t)p->K * p->nt * p->nt, 0.0);
    p->O.assign((size_t)p->K * p->nt * p->nt, 0.0);
    p->A.assign((size_t)p->K * p->ns * p->nt, 0.0);
    p->S.assign((size_t)p->ns * p->ns, 0.0);
    p->g.assign((size_t)p->K * p->nt, 0.0); p->gs.assign(p->ns, 0.0);
    return p;
}
void orc_destroy(orc_problem* p) { delete p; }
void orc_set_threads(orc_problem* p, int n) { p->threads = n < 1 ? 1 : n; }
// fast != 0: structure-aware solver (fastsolve.hpp) with `chunks` time chunks (results depend on chunks, not on threads)
void orc_set_solver(orc_problem* p, int fast, int chunks) { p->fast = fast; p->fast_chunks = chunks < 1 ? 1 : chunks; p->fast_ready = false; }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
int orc_time_value_dim(const orc_problem* p) { return p->tvd; }
int orc_static_value_dim(const orc_problem* p) { return p->svd; }
int orc_time_tangent_dim(const orc_problem* p) { return p->nt; }
int orc_static_tangent_dim(const orc_problem* p) { return p->ns; }
void orc_set_values(orc_problem* p, const double* tv, const double* sv) {
    std::copy(tv, tv + p->tv.size(), p->tv.begin());
    if (p->svd) std::copy(sv, sv + p->sv.size(), p->sv.begin());
}
void orc_get_values(const orc_problem* p, double* tv, double* sv) {
    std::copy(p->tv.begin(), p->tv.end(), tv);
    if (p->svd) std::copy(p->sv.begin(), p->sv.end(), sv);
}
double orc_error(orc_problem* p) { return graph_error(*p, p->tv, p->sv); }
void orc_linearize(orc_problem* p) { linearize(*p); }
void orc_get_gradient(const orc_problem* p, double* g, double* gs) {
    std::copy(p->g.begin(), p->g.end(), g);
    std::copy(p->gs.begin(), p->gs.end(), gs);
}
void orc_get_hessian_blocks(const orc_problem* p, int k, double* D, double* O, double* A) {
    size_t n2 = (size_t)p->nt * p->nt;
    std::copy(p->D.begin() + k * n2, p->D.begin() + (k + 1) * n2, D);
    // O returned as H[(k),(k+1)] ... keep the ABI convention: O = block (k, k+1) i.e. rows k, cols k+1
    for (int i = 0; i < p->nt; i++)
        for (int j = 0; j < p->nt; j++) O[(size_t)i * p->nt + j] = (k + 1 < p->K) ? p->O[k * n2 + (size_t)j * p->nt + i] : 0.0;
    // A returned as block (k, static): nt x ns
    for (int i = 0; i < p->nt; i++)
        for (int j = 0; j < p->ns; j++) A[(size_t)i * p->ns + j] = p->A[((size_t)k * p->ns + j) * p->nt + i];
}
void orc_get_static_hessian(const orc_problem* p, double* S) { std::copy(p->S.begin(), p->S.end(), S); }
int orc_solve(orc_problem* p, double lambda, double* dt, double* ds) {
    bool ok = solve_any(*p, lambda, p->dt_, p->ds_);
    if (!ok) return B200_INDETERMINATE_SYSTEM;
    std::copy(p->dt_.begin(), p->dt_.end(), dt);
    std::copy(p->ds_.begin(), p->ds_.end(), ds);
    return B200_OK;
}
void orc_retract(orc_problem* p, const double* dt, const double* ds) {
    std::vector<double> a(dt, dt + (size_t)p->K * p->nt), b(ds, ds + p->ns), tvn, svn;
    retract_all(*p, a, b, tvn, svn);
    p->tv.swap(tvn); p->sv.swap(svn);
}

void orc_lm_begin(orc_problem* p, const b200_lm_params* prm, double* initial_error) {
    p->lm = *prm;
    p->lambda = prm->gauss_newton ? 0.0 : prm->lambda_initial;
    p->iterations = 0; p->total_trials = 0;
    p->err = graph_error(*p, p->tv, p->sv);
    if (initial_error) *initial_error = p->err;
}

// LevenbergMarquardtOptimizer::iterate() @4.2
int orc_lm_iterate(orc_problem* p, b200_lm_state* st) {
    linearize(*p);
    int status = B200_OK, inner = 0;
    const b200_lm_params& P = p->lm;
    while (true) {  // tryLambda
        inner++;
        bool solved = solve_any(*p, p->lambda, p->dt_, p->ds_);
        double modelFidelity = 0.0, newError = 0.0;
        bool step_ok = false, stop = false;
        if (solved) {
            double oldLin = linear_error(*p, p->dt_, p->ds_, true);
            double newLin = linear_error(*p, p->dt_, p->ds_, false);
            double linChange = oldLin - newLin;
            if (P.gauss_newton) {
                retract_all(*p, p->dt_, p->ds_, p->tv_try, p->sv_try);
                newError = graph_error(*p, p->tv_try, p->sv_try);
                step_ok = true;
            } else if (linChange >= 0) {
                retract_all(*p, p->dt_, p->ds_, p->tv_try, p->sv_try);
                newError = graph_error(*p, p->tv_try, p->sv_try);
                double costChange = p->err - newError;
                if (linChange > std::numeric_limits<double>::epsilon() * oldLin) {
                    modelFidelity = costChange / linChange;
                    step_ok = modelFidelity > P.min_model_fidelity;
                }
                if (std::fabs(costChange) < P.relative_error_tol * p->err) stop = true;
            }
        } else status = B200_INDETERMINATE_SYSTEM;
        if (step_ok) {
            p->tv.swap(p->tv_try); p->sv.swap(p->sv_try);
            p->err = newError;
            if (!P.gauss_newton) p->lambda = std::max(P.lambda_lower_bound, p->lambda / P.lambda_factor);
            p->iterations++; p->total_trials++;
            status = B200_OK;
            break;
        } else if (!stop) {
            p->lambda *= P.lambda_factor; p->total_trials++;
            if (p->lambda >= P.lambda_upper_bound) { status = B200_LAMBDA_MAXED; break; }
            if (P.gauss_newton) { status = B200_INDETERMINATE_SYSTEM; break; }
        } else break;
    }
    if (st) {
        st->error = p->err; st->lambda = p->lambda; st->iterations = p->iterations; st->inner_trials = inner;
        st->total_trials = p->total_trials; st->status = status;
    }
    return status;
}

// bioslam fastOptimize loops
int orc_optimize(orc_problem* p, const b200_lm_params* lm, const b200_outer_params* op, double* hist, b200_optimize_result* res) {
    auto t0 = std::chrono::steady_clock::now();
    double e0;
    orc_lm_begin(p, lm, &e0);
    double current = e0, previous = e0;
    int nh = 0, consec = 0, calls = 0, status = B200_OK;
    if (hist && nh < op->max_history) hist[nh++] = e0;
    b200_lm_state st{};
    if (op->single_imu_rule) {
        // reference src/imuPoseEstimator.cpp:330-359
        double absDec = 9.0e9, relDec = 9.0e9;
        while (relDec > op->rel_err_decrease_limit && absDec > op->abs_err_decrease_limit && current > op->abs_err_limit &&
               p->iterations < op->max_iterations) {
            int before = p->iterations;
            status = orc_lm_iterate(p, &st); calls++;
            previous = current; current = st.error;
            absDec = previous - current; relDec = absDec / previous;
            if (hist && nh < op->max_history) hist[nh++] = current;
            if (p->iterations == before) break;  // "iteration number did not increase": :344-348
        }
    } else {
        // reference src/lowerBodyPoseEstimator.cpp:656-667
        while (consec < op->converge_at_consec_small_iter && p->iterations < op->max_iterations) {
            int before = p->iterations;
            previous = current;
            status = orc_lm_iterate(p, &st); calls++;
            current = st.error;
            double absDec = previous - current, relDec = absDec / previous;
            if (absDec < op->abs_err_decrease_limit || relDec < op->rel_err_decrease_limit || current < op->abs_err_limit) consec++;
            else consec = 0;
            if (hist && nh < op->max_history) hist[nh++] = current;
            if (p->iterations == before && status != B200_OK) break;  // LM gave up: nothing can change any more
        }
    }
    if (res) {
        res->initial_error = e0; res->final_error = p->err; res->final_lambda = p->lambda; res->iterations = p->iterations;
        res->outer_calls = calls; res->total_trials = p->total_trials; res->n_history = nh; res->status = status;
        res->device_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    return status;
}

// ---- single-factor entry points for the unit tests ------------------------------------------------------
int orc_factor_rows(int type) { return factor_info(type).rows; }
int orc_factor_cols(int type) { return factor_ncols(type); }
int orc_factor_nvars(int type) { return factor_info(type).nvars; }
int orc_factor_vartype(int type, int v) { return factor_info(type).vartype[v]; }
int orc_factor_nconst(int type) { return factor_info(type).nconst; }
// x: concatenated variable values in factor order
void orc_factor_eval(int type, const double* params, const double* consts, const double* xcat, double* r, double* J, int whiten) {
    const FactorInfo& fi = factor_info(type);
    const double* x[B200_MAX_FACTOR_VARS] = {nullptr};
    int o = 0;
    for (int v = 0; v < fi.nvars; v++) { x[v] = xcat + o; o += kValDim[fi.vartype[v]]; }
    factor_eval(type, params, consts, x, r, J);
    if (whiten) factor_whiten(type, params, consts, r, J);
}
void orc_retract_var(int vartype, const double* x, const double* d, double* out) { retract_var(vartype, x, d, out); }
void orc_unit3_basis(const double* p, double* B) { unit3_basis(B, p); }
void orc_so3_expmap(const double* w, double* R) { so3_expmap(R, w); }
void orc_so3_logmap(const double* R, double* w) { so3_logmap(w, R); }

// derived joint angles (jointangles.hpp)
void orc_R_imu_to_segment(const double* right_axis, const double* to_prox, const double* to_dist, int prox_is_master, double* R) {
    R_imu_to_segment(right_axis, to_prox, to_dist, prox_is_master != 0, R);
}
void orc_R_sacrum_to_pelvis(const double* up, const double* to_rhip, const double* to_lhip, int hip_is_master, double* R) {
    R_sacrum_to_pelvis(up, to_rhip, to_lhip, hip_is_master != 0, R);
}
void orc_R_segment_to_N(const double* R_imu_to_N, const double* R_imu_to_seg, double* R) { R_segment_to_N(R_imu_to_N, R_imu_to_seg, R); }
void orc_jcs_angles(const double* Rp, const double* Rd, int clinical, int right_leg, double* ang) {
    if (clinical) clinical_jcs_angles(Rp, Rd, right_leg != 0, ang); else consistent_jcs_angles(Rp, Rd, ang);
}
void orc_lower_body_joint_angles(int K, const double* R_imu, const double* statics, double* out) { lower_body_joint_angles(K, R_imu, statics, out); }

int orc_preintegrate(int n_imu, int n_samples, int n_per, int n_intervals, const double* acc, const double* gyro, double dt,
                     const double* bias_hat, const b200_imu_noise* noise, int combined, double* out) {
    const int rec = combined ? 190 : 115;
    int bad = 0;
    for (int m = 0; m < n_imu; m++)
        for (int i = 0; i < n_intervals; i++) {
            size_t s0 = (size_t)m * n_samples + (size_t)i * n_per;
            if ((size_t)(i + 1) * n_per > (size_t)n_samples) return B200_BAD_ARG;
            if (!preintegrate_record(acc + 3 * s0, gyro + 3 * s0, n_per, dt, bias_hat + 6 * m, *noise, combined != 0,
                                     out + ((size_t)m * n_intervals + i) * rec)) bad++;
        }
    return bad ? B200_INDETERMINATE_SYSTEM : B200_OK;
}

}  // extern "C"
