// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.cpp header).
//
// Structure-aware CPU solver of the damped normal equations (GaussianFactorGraph::optimize of the reference path, SURVEY.md
// section 8(a) row A14), the solver the CPU baseline / `bench.py --impl reference` arm runs.  Same linear system as solve() in
// oracle.cpp (which stays as the plain, obviously-correct restatement and pins this one in tests/test_oracle_solver.py):
//
//   (H + lambda I) delta = -g,  H block-tridiagonal in time (only the IMU factors couple k and k+1, through the time-chained
//   dofs) plus the arrow of the static variables.
//
// What makes it a fair CPU baseline rather than a strawman:
//   * the keyframe-local dofs (the angular velocities of the 7-IMU graph) are eliminated per keyframe first, in parallel;
//   * the chain is cut into chunks, one per thread; every chunk eliminates its interior keyframes with a blocked upper
//     Cholesky / blocked triangular solves / rank-k updates built on one register-tiled AVX2 micro-kernel (4 x 12 doubles),
//     carrying the coupling to its left separator; the few separators and the static block are solved sequentially;
//   * the block-sparse structure of the time coupling (one 15 x 15 block per IMU) shortens the triangular solves of its rows;
//   * one OpenMP parallel region per phase (local elimination, chunk sweeps, chunk back-substitution, local back-substitution).
// All sums have a fixed order for a given chunk count: results do not depend on the number of threads.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace orcfast {

typedef double v4d __attribute__((vector_size(32)));
typedef double v4du __attribute__((vector_size(32), aligned(8)));
static inline v4d ldv(const double* p) { return *(const v4du*)p; }
static inline void stv(double* p, v4d x) { *(v4du*)p = x; }
static inline int r4(int n) { return (n + 3) & ~3; }

// C[i][j] += sgn * sum_{k<kd} A(i,k) B[k][j],  A(i,k) = A[i*a_rs + k*a_ks] (read as scalars: any strides), B and C row-major.
// n is rounded up to a multiple of 4: the callers keep leading dimensions that are multiples of 4 and zero padding columns.
template <int MR, int NV>
static inline void mk(double* C, int ldc, const double* A, int a_rs, int a_ks, const double* B, int ldb, int kd, double sgn) {
    v4d acc[MR][NV];
    for (int i = 0; i < MR; i++)
        for (int v = 0; v < NV; v++) acc[i][v] = v4d{0, 0, 0, 0};
    for (int k = 0; k < kd; k++) {
        v4d b[NV];
        const double* bk = B + (size_t)k * ldb;
        for (int v = 0; v < NV; v++) b[v] = ldv(bk + 4 * v);
        for (int i = 0; i < MR; i++) {
            const double a = A[(size_t)i * a_rs + (size_t)k * a_ks];
            const v4d av = v4d{a, a, a, a};
            for (int v = 0; v < NV; v++) acc[i][v] += av * b[v];
        }
    }
    const v4d sv = v4d{sgn, sgn, sgn, sgn};
    for (int i = 0; i < MR; i++)
        for (int v = 0; v < NV; v++) stv(C + (size_t)i * ldc + 4 * v, ldv(C + (size_t)i * ldc + 4 * v) + sv * acc[i][v]);
}
template <int MR>
static inline void mk_row(double* C, int ldc, int n4, const double* A, int a_rs, int a_ks, const double* B, int ldb, int kd, double sgn) {
    int j = 0;
    for (; j + 12 <= n4; j += 12) mk<MR, 3>(C + j, ldc, A, a_rs, a_ks, B + j, ldb, kd, sgn);
    if (j + 8 <= n4) { mk<MR, 2>(C + j, ldc, A, a_rs, a_ks, B + j, ldb, kd, sgn); j += 8; }
    if (j + 4 <= n4) { mk<MR, 1>(C + j, ldc, A, a_rs, a_ks, B + j, ldb, kd, sgn); j += 4; }
}
static inline void gemm_acc(double* C, int ldc, int m, int n, int kd, const double* A, int a_rs, int a_ks, const double* B, int ldb, double sgn) {
    if (m <= 0 || n <= 0 || kd <= 0) return;
    const int n4 = r4(n);
    int i = 0;
    for (; i + 4 <= m; i += 4) mk_row<4>(C + (size_t)i * ldc, ldc, n4, A + (size_t)i * a_rs, a_rs, a_ks, B, ldb, kd, sgn);
    for (; i < m; i++) mk_row<1>(C + (size_t)i * ldc, ldc, n4, A + (size_t)i * a_rs, a_rs, a_ks, B, ldb, kd, sgn);
}

// upper Cholesky in place: M = U^T U, only the upper triangle (diagonal included) of the n x n block is read and written.
// Blocked left-looking by rows (block NB), returns false on a non-positive pivot.
static inline bool chol_upper(double* U, int ld, int n) {
    const int NB = 12;
    for (int j0 = 0; j0 < n; j0 += NB) {
        const int jb = std::min(NB, n - j0);
        // rows [j0, j0+jb) -= (rows above)^T-products : C[i][j] -= sum_k U[k][j0+i] U[k][j0+j]
        if (j0 > 0) gemm_acc(U + (size_t)j0 * ld + j0, ld, jb, n - j0, j0, U + j0, 1, ld, U + j0, ld, -1.0);
        for (int r = j0; r < j0 + jb; r++) {
            double* ur = U + (size_t)r * ld;
            for (int q = j0; q < r; q++) {
                const double* uq = U + (size_t)q * ld;
                const double f = uq[r];
                for (int c = r; c < n; c++) ur[c] -= f * uq[c];
            }
            const double d = ur[r];
            if (!(d > 0.0) || !std::isfinite(d)) return false;
            const double inv = 1.0 / std::sqrt(d);
            for (int c = r; c < n; c++) ur[c] *= inv;
        }
    }
    return true;
}

// rows: R (m x n) <- R U^-1 for upper-triangular U (n x n); columns < k0 of R are known to be zero (block-sparse rows).
// Blocked: the 12 x 12 diagonal blocks of U are inverted (back substitution on the identity) and applied as small GEMMs.
static inline void trsm_upper(double* R, int ldr, int m, int n, const double* U, int ld, int k0) {
    const int NB = 12;
    k0 = (k0 / NB) * NB;
    double Ui[NB * NB];
    std::vector<double> tmp((size_t)m * NB);
    for (int j0 = k0; j0 < n; j0 += NB) {
        const int jb = std::min(NB, n - j0);
        if (j0 > k0) gemm_acc(R + j0, ldr, m, jb, j0 - k0, R + k0, ldr, 1, U + (size_t)k0 * ld + j0, ld, -1.0);
        // Ui = inverse of the diagonal block (upper triangular), rows padded with zeros up to NB columns
        for (int r = 0; r < NB * NB; r++) Ui[r] = 0.0;
        for (int c = 0; c < jb; c++) {
            // column c of the inverse: solve Ujj x = e_c
            for (int r = c; r >= 0; r--) {
                double v = (r == c) ? 1.0 : 0.0;
                for (int q = r + 1; q <= c; q++) v -= U[(size_t)(j0 + r) * ld + j0 + q] * Ui[q * NB + c];
                Ui[r * NB + c] = v / U[(size_t)(j0 + r) * ld + j0 + r];
            }
        }
        for (int i = 0; i < m; i++) {
            double* x = R + (size_t)i * ldr + j0;
            for (int c = 0; c < jb; c++) { tmp[(size_t)i * NB + c] = x[c]; x[c] = 0.0; }
        }
        gemm_acc(R + j0, ldr, m, jb, jb, tmp.data(), NB, 1, Ui, NB, 1.0);
    }
}

// x <- U^-1 x (back substitution), U upper n x n
static inline void solve_upper(const double* U, int ld, int n, double* x) {
    for (int r = n - 1; r >= 0; r--) {
        const double* ur = U + (size_t)r * ld;
        double v = x[r];
        for (int c = r + 1; c < n; c++) v -= ur[c] * x[c];
        x[r] = v / ur[r];
    }
}
// x <- U^-T x (forward substitution with the transpose)
static inline void solve_upper_t(const double* U, int ld, int n, double* x) {
    for (int r = 0; r < n; r++) {
        x[r] /= U[(size_t)r * ld + r];
        const double f = x[r];
        const double* ur = U + (size_t)r * ld;
        for (int c = r + 1; c < n; c++) x[c] -= ur[c] * f;
    }
}
static inline void transpose(const double* A, int lda, int m, int n, double* T, int ldt) {   // T (n x m) = A^T
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) T[(size_t)j * ldt + i] = A[(size_t)i * lda + j];
}

struct Solver {
    // dimensions
    int K = 0, nt = 0, nl = 0, nc = 0, ns = 0, ns1 = 1;
    int ldl = 0, ldc = 0, lds = 0, ldh = 0, h = 0;   // padded leading dimensions: local, chain, static(+rhs), arrow (ns1 + nc)
    std::vector<int> perm_l, perm_c;                   // external tangent index of local / chain dof
    std::vector<int> ostart;                           // per chain row of O: first chain column that can be non-zero
    int P = 1;                                         // chunks
    std::vector<int> cb;                               // chunk begin [P+1]
    // per keyframe storage
    std::vector<double> Ul, V, Dc, Pc, Oc, Uc, Y, dl;  // local factor, local panel, reduced blocks, chain factor, panel, delta
    // per chunk
    std::vector<double> fD, fP, fL, Sacc;
    // separators
    std::vector<double> Dr, Pr, Or, Ur, Yr, dr;
    std::vector<double> S1;
    size_t szUl() const { return (size_t)nl * ldl; }
    size_t szV() const { return (size_t)(nc + ns1) * ldl; }
    size_t szC() const { return (size_t)nc * ldc; }
    size_t szP() const { return (size_t)ns1 * ldc; }
    size_t szY() const { return (size_t)(nc + ns1 + nc) * ldc; }
    size_t szS() const { return (size_t)h * ldh; }

    void setup(int K_, int nt_, int ns_, const std::vector<char>& chained_dof, const std::vector<int>& ostart_ext, int chunks) {
        K = K_; nt = nt_; ns = ns_; ns1 = ns + 1;
        perm_l.clear(); perm_c.clear();
        for (int i = 0; i < nt; i++) (chained_dof[i] ? perm_c : perm_l).push_back(i);
        nl = (int)perm_l.size(); nc = (int)perm_c.size();
        ldl = std::max(4, r4(nl)); ldc = std::max(4, r4(nc)); lds = r4(ns1); h = ns1 + nc; ldh = r4(h);
        ostart.assign(std::max(nc, 1), 0);
        std::vector<int> ext2c(nt, -1);
        for (int i = 0; i < nc; i++) ext2c[perm_c[i]] = i;
        for (int i = 0; i < nc; i++) {
            const int e = ostart_ext[perm_c[i]];
            ostart[i] = (e >= 0 && e < nt && ext2c[e] >= 0) ? ext2c[e] : 0;
        }
        P = std::max(1, std::min(chunks, K / 4));
        if (nc == 0) P = 1;
        cb.resize(P + 1);
        for (int c = 0; c <= P; c++) cb[c] = (int)((long long)c * K / P);
        Ul.assign(K * szUl(), 0.0); V.assign(K * szV(), 0.0); Dc.assign(K * szC(), 0.0); Pc.assign(K * szP(), 0.0);
        Oc.assign(K * szC(), 0.0); Uc.assign(K * szC(), 0.0); Y.assign(K * szY(), 0.0); dl.assign((size_t)K * nt, 0.0);
        fD.assign(P * szC(), 0.0); fP.assign(P * szP(), 0.0); fL.assign(P * szC(), 0.0); Sacc.assign(P * szS(), 0.0);
        const int nsep = std::max(P - 1, 1);
        Dr.assign(nsep * szC(), 0.0); Pr.assign(nsep * szP(), 0.0); Or.assign(nsep * szC(), 0.0); Ur.assign(nsep * szC(), 0.0);
        Yr.assign(nsep * szY(), 0.0); dr.assign((size_t)nsep * std::max(nc, 1), 0.0);
        S1.assign((size_t)ns1 * lds, 0.0);
    }

    // Eliminate nodes [0, n) of a chain left to right.  node(i) gives (D, O, Pn) sources and (U, Yp) storage.  O of node i couples it to
    // node i+1 (rows: next, cols: this); sparse_o: O rows start at ostart[].  OleftT != nullptr: coupling of node 0 to a left
    // separator (rows: this node ... given as the O block (rows node 0, cols separator)), carried as nc extra arrow rows.
    // Outputs: fill blocks of the node after the last one (if has_right) and the arrow Schur sum Sa (h x ldh, upper part valid
    // where row <= col; only the first ns1 + (left ? nc : 0) rows/cols are used).
    struct Node { const double *D, *O, *Pn; double *U, *Yp; };
    template <class GetNode>
    bool sweep(int n, GetNode node, bool has_right, bool sparse_o, const double* Oleft, double* fDo, double* fPo, double* fLo, double* Sa) {
        const bool left = Oleft != nullptr;
        const int ha = ns1 + (left ? nc : 0);
        std::vector<double> XT(szC()), YaT((size_t)nc * ldh), cD(szC(), 0.0), cP(szP(), 0.0), cL(szC(), 0.0);
        std::fill(Sa, Sa + szS(), 0.0);
        for (int i = 0; i < n; i++) {
            Node nd = node(i);
            const bool has_x = nc > 0 && (i + 1 < n || has_right);
            double* U = nd.U;
            // pivot block (upper triangle) = D - fill
            for (int r = 0; r < nc; r++)
                for (int c = r; c < nc; c++) U[(size_t)r * ldc + c] = nd.D[(size_t)r * ldc + c] - (i > 0 ? cD[(size_t)r * ldc + c] : 0.0);
            if (!chol_upper(U, ldc, nc)) return false;
            double* Yx = nd.Yp;                       // rows of O (coupling to the next node)
            double* Ya = nd.Yp + (size_t)nc * ldc;    // arrow rows: statics + rhs, then the left-separator rows
            if (has_x) {
                std::memcpy(Yx, nd.O, sizeof(double) * szC());
                if (sparse_o) {
                    // rows with the same start column are solved together
                    for (int r0 = 0; r0 < nc;) {
                        int r1 = r0 + 1;
                        while (r1 < nc && ostart[r1] == ostart[r0]) r1++;
                        trsm_upper(Yx + (size_t)r0 * ldc, ldc, r1 - r0, nc, U, ldc, ostart[r0]);
                        r0 = r1;
                    }
                } else trsm_upper(Yx, ldc, nc, nc, U, ldc, 0);
            }
            for (int r = 0; r < ns1; r++)
                for (int c = 0; c < ldc; c++) Ya[(size_t)r * ldc + c] = nd.Pn[(size_t)r * ldc + c] - (i > 0 ? cP[(size_t)r * ldc + c] : 0.0);
            if (left) {
                double* Yl = Ya + (size_t)ns1 * ldc;
                if (i == 0) transpose(Oleft, ldc, nc, nc, Yl, ldc);   // H[sep, node0] = O^T
                else for (size_t e = 0; e < szC(); e++) Yl[e] = -cL[e];
            }
            trsm_upper(Ya, ldc, ha, nc, U, ldc, 0);
            // arrow Schur sum (upper triangle by row blocks): Sa += Ya Ya^T
            transpose(Ya, ldc, ha, nc, YaT.data(), ldh);
            for (int r0 = 0; r0 < ha; r0 += 12) {
                const int rb = std::min(12, ha - r0), c0 = r0 & ~3;
                gemm_acc(Sa + (size_t)r0 * ldh + c0, ldh, rb, ha - c0, nc, Ya + (size_t)r0 * ldc, ldc, 1, YaT.data() + c0, ldh, 1.0);
            }
            if (has_x) {
                transpose(Yx, ldc, nc, nc, XT.data(), ldc);
                std::fill(cD.begin(), cD.end(), 0.0); std::fill(cP.begin(), cP.end(), 0.0);
                for (int r0 = 0; r0 < nc; r0 += 12) {
                    const int rb = std::min(12, nc - r0), c0 = r0 & ~3;
                    gemm_acc(cD.data() + (size_t)r0 * ldc + c0, ldc, rb, nc - c0, nc, Yx + (size_t)r0 * ldc, ldc, 1, XT.data() + c0, ldc, 1.0);
                }
                gemm_acc(cP.data(), ldc, ns1, nc, nc, Ya, ldc, 1, XT.data(), ldc, 1.0);
                if (left) {
                    std::fill(cL.begin(), cL.end(), 0.0);
                    gemm_acc(cL.data(), ldc, nc, nc, nc, Ya + (size_t)ns1 * ldc, ldc, 1, XT.data(), ldc, 1.0);
                }
            }
        }
        if (has_right) {
            if (n > 0) { std::memcpy(fDo, cD.data(), sizeof(double) * szC()); std::memcpy(fPo, cP.data(), sizeof(double) * szP()); }
            else { std::fill(fDo, fDo + szC(), 0.0); std::fill(fPo, fPo + szP(), 0.0); }
            if (left) std::memcpy(fLo, cL.data(), sizeof(double) * szC());
        }
        return true;
    }

    // delta_i = -U^-1 ( X^T delta_next + Ya^T coef ), nodes in reverse order; coef = [delta_static, -1, delta_left]
    template <class GetNode, class GetDelta>
    void backsub(int n, GetNode node, GetDelta delta_of, bool has_right, const double* d_right, bool left, const double* coef) {
        const int ha = ns1 + (left ? nc : 0);
        std::vector<double> t(std::max(ldc, 4));
        for (int i = n - 1; i >= 0; i--) {
            Node nd = node(i);
            const bool has_x = nc > 0 && (i + 1 < n || has_right);
            std::fill(t.begin(), t.end(), 0.0);
            if (has_x) {
                const double* dn = (i + 1 < n) ? delta_of(i + 1) : d_right;
                for (int j = 0; j < nc; j++) {
                    const double f = dn[j];
                    const double* x = nd.Yp + (size_t)j * ldc;
                    for (int c = 0; c < nc; c++) t[c] += x[c] * f;
                }
            }
            const double* Ya = nd.Yp + (size_t)nc * ldc;
            for (int r = 0; r < ha; r++) {
                const double f = coef[r];
                const double* y = Ya + (size_t)r * ldc;
                for (int c = 0; c < nc; c++) t[c] += y[c] * f;
            }
            solve_upper(nd.U, ldc, nc, t.data());
            double* d = delta_of(i);
            for (int c = 0; c < nc; c++) d[c] = -t[c];
        }
    }

    // D, O, A, g, S, gs in the layout of orc_problem (external tangent order).  Returns false if a pivot is not positive.
    bool solve(const double* D, const double* O, const double* A, const double* g, const double* S, const double* gs, double lambda,
               int threads, double* dt, double* ds) {
        const int ntl = nt;
        int bad = 0;
        std::vector<double> Tloc((size_t)K * ns1 * lds, 0.0);
        // ---- phase 1: per keyframe, eliminate the local dofs; build the reduced chain blocks in internal order
#pragma omp parallel for num_threads(threads) schedule(static) reduction(| : bad)
        for (int k = 0; k < K; k++) {
            const double* Dk = D + (size_t)k * ntl * ntl;
            const double* Ak = A + (size_t)k * ns * ntl;
            const double* gk = g + (size_t)k * ntl;
            double* ul = Ul.data() + k * szUl();
            double* v = V.data() + k * szV();
            double* dc = Dc.data() + k * szC();
            double* pc = Pc.data() + k * szP();
            for (int r = 0; r < nc; r++)
                for (int c = r; c < nc; c++) dc[(size_t)r * ldc + c] = Dk[(size_t)perm_c[r] * ntl + perm_c[c]] + (r == c ? lambda : 0.0);
            for (int a = 0; a < ns; a++)
                for (int c = 0; c < nc; c++) pc[(size_t)a * ldc + c] = Ak[(size_t)a * ntl + perm_c[c]];
            for (int c = 0; c < nc; c++) pc[(size_t)ns * ldc + c] = -gk[perm_c[c]];
            if (k + 1 < K && nc > 0) {
                const double* Ok = O + (size_t)k * ntl * ntl;
                double* oc = Oc.data() + k * szC();
                for (int r = 0; r < nc; r++)
                    for (int c = 0; c < nc; c++) oc[(size_t)r * ldc + c] = Ok[(size_t)perm_c[r] * ntl + perm_c[c]];
            }
            if (nl == 0) continue;
            for (int r = 0; r < nl; r++)
                for (int c = r; c < nl; c++) ul[(size_t)r * ldl + c] = Dk[(size_t)perm_l[r] * ntl + perm_l[c]] + (r == c ? lambda : 0.0);
            if (!chol_upper(ul, ldl, nl)) { bad |= 1; continue; }
            for (int r = 0; r < nc; r++)
                for (int c = 0; c < nl; c++) v[(size_t)r * ldl + c] = Dk[(size_t)perm_c[r] * ntl + perm_l[c]];
            for (int a = 0; a < ns; a++)
                for (int c = 0; c < nl; c++) v[(size_t)(nc + a) * ldl + c] = Ak[(size_t)a * ntl + perm_l[c]];
            for (int c = 0; c < nl; c++) v[(size_t)(nc + ns) * ldl + c] = -gk[perm_l[c]];
            trsm_upper(v, ldl, nc + ns1, nl, ul, ldl, 0);
            std::vector<double> vT((size_t)nl * (r4(nc + ns1) + 4));
            const int ldv = r4(nc + ns1) + 4;
            transpose(v, ldl, nc + ns1, nl, vT.data(), ldv);
            // D' -= Vc Vc^T (upper), P' -= Vp Vc^T, T = Vp Vp^T
            for (int r0 = 0; r0 < nc; r0 += 12) {
                const int rb = std::min(12, nc - r0), c0 = r0 & ~3;
                gemm_acc(dc + (size_t)r0 * ldc + c0, ldc, rb, nc - c0, nl, v + (size_t)r0 * ldl, ldl, 1, vT.data() + c0, ldv, -1.0);
            }
            // (the rounding of the column count up to a multiple of 4 may touch columns >= nc of vT == rows of Vp: keep padding clean)
            gemm_acc(pc, ldc, ns1, nc, nl, v + (size_t)nc * ldl, ldl, 1, vT.data(), ldv, -1.0);
            for (int r = 0; r < nc; r++) for (int c = nc; c < ldc; c++) dc[(size_t)r * ldc + c] = 0.0;
            for (int r = 0; r < ns1; r++) for (int c = nc; c < ldc; c++) pc[(size_t)r * ldc + c] = 0.0;
            double* T = Tloc.data() + (size_t)k * ns1 * lds;
            gemm_acc(T, lds, ns1, ns1, nl, v + (size_t)nc * ldl, ldl, 1, vT.data() + nc, ldv, 1.0);
        }
        if (bad) return false;
        // ---- phase 2: chunk sweeps (interior keyframes), one chunk per task
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1) reduction(| : bad)
        for (int c = 0; c < P; c++) {
            const int k0 = cb[c], k1 = cb[c + 1];
            const int n = (c < P - 1) ? (k1 - 1 - k0) : (k1 - k0);
            auto node = [&](int i) {
                const int k = k0 + i;
                return Node{Dc.data() + k * szC(), Oc.data() + k * szC(), Pc.data() + k * szP(), Uc.data() + k * szC(), Y.data() + k * szY()};
            };
            const double* Oleft = (c > 0) ? Oc.data() + (size_t)(k0 - 1) * szC() : nullptr;
            if (!sweep(n, node, c < P - 1, true, Oleft, fD.data() + c * szC(), fP.data() + c * szP(), fL.data() + c * szC(), Sacc.data() + c * szS())) bad |= 1;
        }
        if (bad) return false;
        // ---- phase 3: separators (sequential): reduced blocks, sweep, static system
        for (int r = 0; r < ns1; r++)
            for (int c = 0; c < ns1; c++) {
                double v = 0.0;
                if (r < ns && c < ns) v = S[(size_t)r * ns + c] + (r == c ? lambda : 0.0);
                else if (r < ns) v = -gs[r];
                else if (c < ns) v = -gs[c];
                S1[(size_t)r * lds + c] = v;
            }
        for (int k = 0; k < K && nl > 0; k++) {
            const double* T = Tloc.data() + (size_t)k * ns1 * lds;
            for (int r = 0; r < ns1; r++) for (int c = 0; c < ns1; c++) S1[(size_t)r * lds + c] -= T[(size_t)r * lds + c];
        }
        auto sym = [&](const double* Sa, int r, int c) { return r <= c ? Sa[(size_t)r * ldh + c] : Sa[(size_t)c * ldh + r]; };
        for (int c = 0; c < P; c++) {
            const double* Sa = Sacc.data() + c * szS();
            for (int r = 0; r < ns1; r++) for (int q = 0; q < ns1; q++) S1[(size_t)r * lds + q] -= sym(Sa, r, q);
        }
        const int nsep = P - 1;
        for (int s = 0; s < nsep; s++) {
            const int k = cb[s + 1] - 1;
            const double* Sn = Sacc.data() + (s + 1) * szS();   // chunk s+1: its left separator is s
            double* dr_ = Dr.data() + s * szC();
            double* pr = Pr.data() + s * szP();
            const double* dc = Dc.data() + k * szC();
            const double* pc = Pc.data() + k * szP();
            const double* fd = fD.data() + s * szC();
            const double* fp = fP.data() + s * szP();
            for (int r = 0; r < nc; r++)
                for (int c = r; c < nc; c++) dr_[(size_t)r * ldc + c] = dc[(size_t)r * ldc + c] - fd[(size_t)r * ldc + c] - sym(Sn, ns1 + r, ns1 + c);
            for (int a = 0; a < ns1; a++)
                for (int c = 0; c < nc; c++) pr[(size_t)a * ldc + c] = pc[(size_t)a * ldc + c] - fp[(size_t)a * ldc + c] - sym(Sn, a, ns1 + c);
            if (s + 1 < nsep) {   // coupling separator s -> s+1 through chunk s+1: -(fill of its left rows)^T
                const double* fl = fL.data() + (s + 1) * szC();
                double* orr = Or.data() + s * szC();
                for (int i = 0; i < nc; i++) for (int j = 0; j < nc; j++) orr[(size_t)i * ldc + j] = -fl[(size_t)j * ldc + i];
            }
        }
        std::vector<double> Stop(szS(), 0.0);
        if (nsep > 0) {
            auto node = [&](int i) { return Node{Dr.data() + i * szC(), Or.data() + i * szC(), Pr.data() + i * szP(), Ur.data() + i * szC(), Yr.data() + i * szY()}; };
            if (!sweep(nsep, node, false, false, nullptr, nullptr, nullptr, nullptr, Stop.data())) return false;
            for (int r = 0; r < ns1; r++) for (int q = 0; q < ns1; q++) S1[(size_t)r * lds + q] -= sym(Stop.data(), r, q);
        }
        std::vector<double> coef(h + 4, 0.0);
        if (ns > 0) {
            std::vector<double> Us((size_t)ns * lds);
            for (int r = 0; r < ns; r++) for (int c = 0; c < lds; c++) Us[(size_t)r * lds + c] = (c < ns) ? S1[(size_t)r * lds + c] : 0.0;
            if (!chol_upper(Us.data(), lds, ns)) return false;
            for (int i = 0; i < ns; i++) coef[i] = S1[(size_t)i * lds + ns];
            solve_upper_t(Us.data(), lds, ns, coef.data());
            solve_upper(Us.data(), lds, ns, coef.data());
            for (int i = 0; i < ns; i++) ds[i] = coef[i];
        }
        coef[ns] = -1.0;
        if (nsep > 0) {
            auto node = [&](int i) { return Node{nullptr, nullptr, nullptr, Ur.data() + i * szC(), Yr.data() + i * szY()}; };
            auto dof = [&](int i) { return dr.data() + (size_t)i * nc; };
            backsub(nsep, node, dof, false, nullptr, false, coef.data());
        }
        // ---- phase 4: chunk back-substitution, then the local dofs
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
        for (int c = 0; c < P; c++) {
            const int k0 = cb[c], k1 = cb[c + 1];
            const int n = (c < P - 1) ? (k1 - 1 - k0) : (k1 - k0);
            std::vector<double> cf(coef.begin(), coef.end());
            if (c > 0) for (int j = 0; j < nc; j++) cf[ns1 + j] = dr[(size_t)(c - 1) * nc + j];
            auto node = [&](int i) { const int k = k0 + i; return Node{nullptr, nullptr, nullptr, Uc.data() + k * szC(), Y.data() + k * szY()}; };
            auto dof = [&](int i) { return dl.data() + (size_t)(k0 + i) * nt + nl; };
            if (c < P - 1) std::memcpy(dl.data() + (size_t)(k1 - 1) * nt + nl, dr.data() + (size_t)c * nc, sizeof(double) * nc);
            backsub(n, node, dof, c < P - 1, (c < P - 1) ? dr.data() + (size_t)c * nc : nullptr, c > 0, cf.data());
        }
#pragma omp parallel for num_threads(threads) schedule(static)
        for (int k = 0; k < K; k++) {
            double* d = dl.data() + (size_t)k * nt;   // internal order: [local | chain]
            if (nl > 0) {
                const double* v = V.data() + k * szV();
                double t[64];
                std::vector<double> tb;
                double* tp = t;
                if (nl > 64) { tb.resize(nl); tp = tb.data(); }
                for (int c = 0; c < nl; c++) tp[c] = 0.0;
                for (int j = 0; j < nc; j++) { const double f = d[nl + j]; const double* x = v + (size_t)j * ldl; for (int c = 0; c < nl; c++) tp[c] += x[c] * f; }
                for (int a = 0; a < ns1; a++) { const double f = coef[a]; const double* x = v + (size_t)(nc + a) * ldl; for (int c = 0; c < nl; c++) tp[c] += x[c] * f; }
                solve_upper(Ul.data() + k * szUl(), ldl, nl, tp);
                for (int c = 0; c < nl; c++) d[c] = -tp[c];
            }
            double* out = dt + (size_t)k * nt;
            for (int c = 0; c < nl; c++) out[perm_l[c]] = d[c];
            for (int c = 0; c < nc; c++) out[perm_c[c]] = d[nl + c];
        }
        return true;
    }
};

}  // namespace orcfast
