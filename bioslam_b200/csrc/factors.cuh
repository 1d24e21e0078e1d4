// Device-side residuals + analytic Jacobians of every factor on the hot path, fp64.
// Closed forms of SURVEY.md Appendix B, i.e. the collapsed chain-rule products of reference src/factors/*.cpp and
// src/mathutils.cpp (cited per case) and of gtsam::CombinedImuFactor / ImuFactor / PriorFactor<T> @4.2.
// One CUDA thread evaluates one factor instance; J is row-major rows x cols (cols = sum of the variables' tangent dims,
// variable order = reference constructor key order).  Constants are read through a strided accessor so that the
// keyframe-major SoA arrays in HBM are read coalesced across the warp (consecutive lanes = consecutive keyframes).
#pragma once
#include "../../include/bioslam_b200.h"
#include "geometry.cuh"

namespace b200d {

struct CstRef {  // per-keyframe constant record: element i lives at p[i*stride]
    const double* p;
    long long stride;
    __device__ __forceinline__ double operator[](int i) const { return p[(long long)i * stride]; }
};

__device__ __forceinline__ void put33(double* J, int ld, int r0, int c0, const double* M, double s) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) J[(r0 + i) * ld + c0 + j] = s * M[3 * i + j];
}

// gradient of the Euclidean norm as gtsam::norm3 defines it (gtsam/geometry/Point3.cpp @4.2): v / |v|, and (1, 1, 1) when
// |v| <= 1e-10 -- every norm on the path goes through it (reference src/mathutils.cpp:75,754, src/factors/*.cpp), so a
// degenerate geometry (parallel / anti-parallel vectors, coincident points) keeps a finite Jacobian exactly as upstream
__device__ __forceinline__ void norm3_grad(const double* v, double n, double* u) {
    if (fabs(n) > 1e-10) { u[0] = v[0] / n; u[1] = v[1] / n; u[2] = v[2] / n; }
    else { u[0] = 1.0; u[1] = 1.0; u[2] = 1.0; }
}

// atan2(|a x b|, a.b) and its gradients (reference src/mathutils.cpp:63-91)
template <bool WJ>
__device__ __forceinline__ double unsigned_angle(const double* a, const double* b, double* Ha, double* Hb) {
    double c[3]; cross3(c, a, b);
    const double y = sqrt(dot3(c, c)), x = dot3(a, b);
    if (WJ) {
        const double D = x * x + y * y;
        const double fy = x / D, fx = -y / D;
        double u[3]; norm3_grad(c, y, u);
        // d|c|/da = u^T(-[b]x) = (b x u)... as row vectors: u^T(-[b]x) = (u x b)^T ... computed explicitly below
        double ub[3], au[3];
        cross3(ub, b, u);   // u^T(-[b]x) = (b x u)^T
        cross3(au, u, a);   // u^T([a]x)  = (u x a)^T
        for (int j = 0; j < 3; j++) { Ha[j] = fy * ub[j] + fx * b[j]; Hb[j] = fy * au[j] + fx * a[j]; }
    }
    return atan2(y, x);
}

// per-type dimensions
__host__ __device__ __forceinline__ int factor_rows(int t) {
    switch (t) {
    case B200_F_IMU_COMBINED: return 15;
    case B200_F_IMU_STATICBIAS: return 9;
    case B200_F_PRIOR_BIAS6: case B200_F_PRIOR_POSE3: return 6;
    case B200_F_PRIOR_UNIT3: return 2;
    case B200_F_HINGE_NORM: case B200_F_JOINTPOS_NORM: case B200_F_JOINTVEL_NORM: case B200_F_SEGLEN_MAG: case B200_F_SEGLEN_MAX: case B200_F_SEGLEN_MIN:
    case B200_F_SEGLEN_DISCREPANCY: case B200_F_ANGLE_AXIS_SEG: case B200_F_POSE3_COMPASS_PRIOR: return 1;
    default: return 3;
    }
}

// hinge: m = R_A^T R_B w_B - w_A ; e = (I - k k^T) m      (reference src/factors/HingeJointFactors.cpp:51-109)
template <bool WJ>
__device__ __forceinline__ void hinge_vec(const double* xA, const double* wA, const double* xB, const double* wB, const double* k,
                                          double* e, double* J3 /*3x20*/) {
    double RAB[9], wBA[3], m[3];
    mTm(RAB, xA, xB);
    mat_vec(wBA, RAB, wB);
    for (int i = 0; i < 3; i++) m[i] = wBA[i] - wA[i];
    const double mk = dot3(m, k);
    for (int i = 0; i < 3; i++) e[i] = m[i] - mk * k[i];
    if (WJ) {
        const int ld = 20;
        double P[9], T[9], PR[9], B[6];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) P[3 * i + j] = (i == j ? 1.0 : 0.0) - k[i] * k[j];
        for (int i = 0; i < 60; i++) J3[i] = 0.0;
        m_skew(T, P, wBA); put33(J3, ld, 0, 0, T, 1.0);       // P [R_A^T R_B w_B]x
        put33(J3, ld, 0, 6, P, -1.0);                         // -P
        mm(PR, P, RAB);
        m_skew(T, PR, wB); put33(J3, ld, 0, 9, T, -1.0);      // -P R_A^T R_B [w_B]x
        put33(J3, ld, 0, 15, PR, 1.0);                        // P R_A^T R_B
        unit3_basis(B, k);
        for (int i = 0; i < 3; i++)
            for (int c = 0; c < 2; c++) {
                double s = 0;
                for (int j = 0; j < 3; j++) s -= (k[i] * m[j] + (i == j ? mk : 0.0)) * B[2 * j + c];
                J3[i * ld + 18 + c] = s;                      // -(k m^T + (m.k) I) B
            }
    }
}

// Unwhitened residual r and Jacobian J (rows x cols row-major, zero-filled here).
template <bool WJ>
__device__ void factor_eval(int type, const double* params, CstRef cst, const double* const* x, double* r, double* J, int cols) {
    const int rows = factor_rows(type);
    if (WJ) for (int i = 0; i < rows * cols; i++) J[i] = 0.0;
    const int ld = cols;
    switch (type) {
    case B200_F_IMU_COMBINED:
    case B200_F_IMU_STATICBIAS: {
        const bool comb = (type == B200_F_IMU_COMBINED);
        const double *Ti = x[0], *vi = x[1], *Tj = x[2], *vj = x[3], *bi = x[4];
        const double dt = cst[69];
        double dba[3], dbg[3], zc[9];
        for (int i = 0; i < 3; i++) { dba[i] = bi[i] - cst[63 + i]; dbg[i] = bi[3 + i] - cst[66 + i]; }
        for (int i = 0; i < 9; i++)
            zc[i] = cst[i] + cst[9 + 3 * i] * dba[0] + cst[10 + 3 * i] * dba[1] + cst[11 + 3 * i] * dba[2] +
                    cst[36 + 3 * i] * dbg[0] + cst[37 + 3 * i] * dbg[1] + cst[38 + 3 * i] * dbg[2];
        const double *th = zc, *p = zc + 3, *v = zc + 6, *g = params;
        const double *Ri = Ti, *pi = Ti + 9, *Rj = Tj, *pj = Tj + 9;
        double E[9], A[9], M[9], rR[3], rp[3], rv[3], t1[3], t2[3];
        so3_exp(E, th);
        mTm(A, Rj, Ri);
        mm(M, A, E);
        so3_log(rR, M);
        mat_vec(t1, Ri, p);
        for (int i = 0; i < 3; i++) t1[i] += pi[i] + dt * vi[i] + 0.5 * dt * dt * g[i] - pj[i];
        matT_vec(rp, Rj, t1);
        mat_vec(t2, Ri, v);
        for (int i = 0; i < 3; i++) t2[i] += vi[i] + dt * g[i] - vj[i];
        matT_vec(rv, Rj, t2);
        for (int i = 0; i < 3; i++) { r[i] = rR[i]; r[3 + i] = rp[i]; r[6 + i] = rv[i]; }
        if (comb) for (int i = 0; i < 6; i++) r[9 + i] = bi[i] - x[5][i];
        if (WJ) {
            double Jri[9], Jrth[9], T[9], RjT[9], S[9], H[9], L[9];
            so3_dlog(Jri, rR);
            so3_dexp(Jrth, th);
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) RjT[3 * i + j] = Rj[3 * j + i];
            mmT(T, Jri, E); put33(J, ld, 0, 0, T, 1.0);            // d rR / d theta_i
            m_skew(T, A, p); put33(J, ld, 3, 0, T, -1.0);          // d rp / d theta_i
            put33(J, ld, 3, 3, A, 1.0);                            // d rp / d p_i
            m_skew(T, A, v); put33(J, ld, 6, 0, T, -1.0);          // d rv / d theta_i
            put33(J, ld, 3, 6, RjT, dt);                           // d rp / d v_i
            put33(J, ld, 6, 6, RjT, 1.0);                          // d rv / d v_i
            mmT(T, Jri, M); put33(J, ld, 0, 9, T, -1.0);           // d rR / d theta_j
            skew3(S, rp); put33(J, ld, 3, 9, S, 1.0);
            for (int i = 0; i < 3; i++) J[(3 + i) * ld + 12 + i] = -1.0;
            skew3(S, rv); put33(J, ld, 6, 9, S, 1.0);
            put33(J, ld, 6, 15, RjT, -1.0);                        // d rv / d v_j
            mm(L, Jri, Jrth);
            for (int blk = 0; blk < 3; blk++)
                for (int half = 0; half < 2; half++) {
                    const int base = (half == 0 ? 9 : 36) + 9 * blk;
                    for (int i = 0; i < 9; i++) H[i] = cst[base + i];
                    mm(T, blk == 0 ? L : A, H);
                    put33(J, ld, 3 * blk, 18 + 3 * half, T, 1.0);
                }
            if (comb) for (int i = 0; i < 6; i++) { J[(9 + i) * ld + 18 + i] = 1.0; J[(9 + i) * ld + 24 + i] = -1.0; }
        }
    } break;
    case B200_F_ANGVEL: {  // reference src/factors/AngularVelocityFactor.cpp:18-31
        for (int i = 0; i < 3; i++) r[i] = x[0][i] + x[1][3 + i] - cst[i];
        if (WJ) for (int i = 0; i < 3; i++) { J[i * ld + i] = 1.0; J[i * ld + 6 + i] = 1.0; }
    } break;
    case B200_F_HINGE_VEC: hinge_vec<WJ>(x[0], x[1], x[2], x[3], x[4], r, J); break;
    case B200_F_HINGE_NORM: {  // reference src/factors/HingeJointFactors.cpp:118-173
        double e[3], J3[60];
        hinge_vec<WJ>(x[0], x[1], x[2], x[3], x[4], e, J3);
        const double n = sqrt(dot3(e, e));
        r[0] = n;
        if (WJ) { double u[3]; norm3_grad(e, n, u); for (int c = 0; c < 20; c++) J[c] = u[0] * J3[c] + u[1] * J3[20 + c] + u[2] * J3[40 + c]; }
    } break;
    case B200_F_JOINTPOS_VEC:
    case B200_F_JOINTPOS_NORM: {  // reference src/mathutils.cpp:735-747, src/factors/ConstrainedJointCenterPositionFactor.cpp:12-123
        const double *xA = x[0], *xB = x[1], *sA = x[2], *sB = x[3];
        double a[3], b[3], e[3], J3[54];
        mat_vec(a, xA, sA); mat_vec(b, xB, sB);
        for (int i = 0; i < 3; i++) e[i] = a[i] + xA[9 + i] - b[i] - xB[9 + i];
        double* Jt = (type == B200_F_JOINTPOS_VEC) ? J : J3;
        if (WJ) {
            double T[9];
            if (type != B200_F_JOINTPOS_VEC) for (int i = 0; i < 54; i++) J3[i] = 0.0;
            m_skew(T, xA, sA); put33(Jt, 18, 0, 0, T, -1.0); put33(Jt, 18, 0, 3, xA, 1.0);
            m_skew(T, xB, sB); put33(Jt, 18, 0, 6, T, 1.0); put33(Jt, 18, 0, 9, xB, -1.0);
            put33(Jt, 18, 0, 12, xA, 1.0); put33(Jt, 18, 0, 15, xB, -1.0);
        }
        if (type == B200_F_JOINTPOS_VEC) { r[0] = e[0]; r[1] = e[1]; r[2] = e[2]; }
        else {
            const double n = sqrt(dot3(e, e));
            r[0] = n - params[1];
            if (WJ) { double u[3]; norm3_grad(e, n, u); for (int c = 0; c < 18; c++) J[c] = u[0] * J3[c] + u[1] * J3[18 + c] + u[2] * J3[36 + c]; }
        }
    } break;
    case B200_F_JOINTVEL_VEC:
    case B200_F_JOINTVEL_NORM: {  // reference src/factors/ConstrainedJointCenterVelocityFactor.cpp:68-131 (vector), :137-214 (norm)
        const bool vec = type == B200_F_JOINTVEL_VEC;
        double e[3], J3[90];
        double* Jt = vec ? J : J3;
        const int ldt = vec ? ld : 30;
        if (WJ && !vec) for (int i = 0; i < 90; i++) J3[i] = 0.0;
        for (int side = 0; side < 2; side++) {
            const double *X = x[4 * side], *V = x[4 * side + 1], *W = x[4 * side + 2], *S_ = x[4 * side + 3];
            const double sg = side == 0 ? 1.0 : -1.0;
            double ws[3], Rws[3];
            cross3(ws, W, S_); mat_vec(Rws, X, ws);
            for (int i = 0; i < 3; i++) { const double t = V[i] + Rws[i]; if (side == 0) e[i] = t; else e[i] -= t; }
            if (WJ) {
                const int c0 = 15 * side;
                double T[9];
                m_skew(T, X, ws); put33(Jt, ldt, 0, c0, T, -sg);
                for (int i = 0; i < 3; i++) Jt[i * ldt + c0 + 6 + i] = sg;
                m_skew(T, X, S_); put33(Jt, ldt, 0, c0 + 9, T, -sg);
                m_skew(T, X, W); put33(Jt, ldt, 0, c0 + 12, T, sg);
            }
        }
        if (vec) { r[0] = e[0]; r[1] = e[1]; r[2] = e[2]; }
        else {
            const double n = sqrt(dot3(e, e));
            r[0] = n;
            if (WJ) { double u[3]; norm3_grad(e, n, u); for (int c = 0; c < 30; c++) J[c] = u[0] * J3[c] + u[1] * J3[30 + c] + u[2] * J3[60 + c]; }
        }
    } break;
    case B200_F_SEGLEN_MAG:
    case B200_F_SEGLEN_MAX:
    case B200_F_SEGLEN_MIN: {  // reference src/factors/SegmentLengthMagnitudeFactor.cpp:31-144, src/mathutils.cpp:148-180
        double d[3] = {x[0][0] - x[1][0], x[0][1] - x[1][1], x[0][2] - x[1][2]};
        const double n = sqrt(dot3(d, d));
        double e = 0.0, dedn = 0.0;
        if (type == B200_F_SEGLEN_MAG) { e = n - params[1]; dedn = 1.0; }
        else if (type == B200_F_SEGLEN_MAX) {
            const double xm = params[1], a = params[2];
            if (n > xm) { const double t = tanh(a * (n - xm)); e = (n - xm) * t * t; dedn = t * t - 2.0 * a * t * (t * t - 1.0) * (n - xm); }
        } else {
            const double xm = params[1], a = params[2];
            if (n < xm) {
                const double t0 = tanh(a * (xm - n)); e = (xm - n) * t0 * t0;
                const double t = tanh(a * (n - xm)); dedn = 2.0 * a * t * (t * t - 1.0) * (n - xm) - t * t;
            }
        }
        r[0] = e;
        if (WJ) { double u[3]; norm3_grad(d, n, u); for (int i = 0; i < 3; i++) { J[i] = dedn * u[i]; J[3 + i] = -dedn * u[i]; } }
    } break;
    case B200_F_SEGLEN_DISCREPANCY: {  // reference src/factors/SegmentLengthDiscrepancyFactor.cpp:43-72
        double a[3], b[3];
        for (int i = 0; i < 3; i++) { a[i] = x[0][i] - x[1][i]; b[i] = x[2][i] - x[3][i]; }
        const double na = sqrt(dot3(a, a)), nb = sqrt(dot3(b, b));
        r[0] = na - nb;
        if (WJ) { double ua[3], ub[3]; norm3_grad(a, na, ua); norm3_grad(b, nb, ub); for (int i = 0; i < 3; i++) { J[i] = ua[i]; J[3 + i] = -ua[i]; J[6 + i] = -ub[i]; J[9 + i] = ub[i]; } }
    } break;
    case B200_F_ANGLE_AXIS_SEG: {  // reference src/factors/AngleBetweenAxisAndSegmentFactor.cpp:38-92
        const double* k = x[0];
        double v[3] = {x[1][0] - x[2][0], x[1][1] - x[2][1], x[1][2] - x[2][2]}, Hv[3], Hk[3];
        const double ang = unsigned_angle<WJ>(v, k, Hv, Hk);
        r[0] = params[1] - ang;
        if (WJ) {
            double B[6]; unit3_basis(B, k);
            for (int c = 0; c < 2; c++) J[c] = -(Hk[0] * B[c] + Hk[1] * B[2 + c] + Hk[2] * B[4 + c]);
            for (int i = 0; i < 3; i++) { J[2 + i] = -Hv[i]; J[5 + i] = Hv[i]; }
        }
    } break;
    case B200_F_POSE3_TRANSLATION_PRIOR: {  // reference src/factors/Pose3Priors.cpp:22-32
        for (int i = 0; i < 3; i++) r[i] = x[0][9 + i] - params[3 + i];
        if (WJ) put33(J, ld, 0, 3, x[0], 1.0);
    } break;
    case B200_F_POSE3_COMPASS_PRIOR: {  // reference src/factors/Pose3Priors.cpp:35-74, src/mathutils.cpp:771-781,93-119
        const double* R = x[0];
        const double *lB = params + 2, *ref = params + 5, *up = params + 8;
        double lN[3], t[3], h[3], Hh[3], Hd[3], axb[3], dmy[3];
        mat_vec(lN, R, lB);
        cross3(t, lN, up);
        cross3(h, t, up);   // (lN x up) x up, the reference's literal (negated) in-plane projection
        double ang = unsigned_angle<WJ>(h, ref, Hh, Hd);
        cross3(axb, h, ref);
        const bool neg = unsigned_angle<false>(up, axb, dmy, dmy) > B200_PI / 2;
        if (neg) ang = -ang;
        r[0] = ang - params[1];
        if (WJ) {
            double Su[9], SS[9], RS[9], T[9];
            skew3(Su, up); mm(SS, Su, Su);
            m_skew(RS, R, lB);       // dlN/dtheta = -R [lB]x
            mm(T, SS, RS);
            const double sg = neg ? -1.0 : 1.0;
            for (int j = 0; j < 3; j++) J[j] = -sg * (Hh[0] * T[j] + Hh[1] * T[3 + j] + Hh[2] * T[6 + j]);
        }
    } break;
    case B200_F_PRIOR_VEC3: {
        for (int i = 0; i < 3; i++) r[i] = x[0][i] - params[3 + i];
        if (WJ) for (int i = 0; i < 3; i++) J[i * ld + i] = 1.0;
    } break;
    case B200_F_PRIOR_BIAS6: {
        for (int i = 0; i < 6; i++) r[i] = x[0][i] - params[6 + i];
        if (WJ) for (int i = 0; i < 6; i++) J[i * ld + i] = 1.0;
    } break;
    case B200_F_PRIOR_UNIT3: {  // -x.localCoordinates(prior), Jacobian := I (GTSAM PriorFactor approximation)
        double l[2]; unit3_local(l, x[0], params + 2);
        r[0] = -l[0]; r[1] = -l[1];
        if (WJ) { J[0] = 1.0; J[ld + 1] = 1.0; }
    } break;
    case B200_F_PRIOR_POSE3: {
        const double* R = x[0]; const double* t = x[0] + 9;
        double dR[9], d[3], dt3[3], xi[6];
        mTm(dR, R, params + 6);
        for (int i = 0; i < 3; i++) d[i] = params[15 + i] - t[i];
        matT_vec(dt3, R, d);
        pose3_log(xi, dR, dt3);
        for (int i = 0; i < 6; i++) r[i] = -xi[i];
        if (WJ) for (int i = 0; i < 6; i++) J[i * ld + i] = 1.0;
    } break;
    case B200_F_MAG_POSE3: {  // reference src/factors/MagPose3Factor.cpp:25-38
        const double m[3] = {cst[0], cst[1], cst[2]};
        double mN[3]; mat_vec(mN, x[0], m);
        for (int i = 0; i < 3; i++) r[i] = mN[i] - params[3 + i];
        if (WJ) { double T[9]; m_skew(T, x[0], m); put33(J, ld, 0, 0, T, -1.0); }
    } break;
    default: break;
    }
}

// whitening: diagonal sigmas in params[0..rows) ; Gaussian factors: packed upper sqrt-information at consts[70..]
// Gaussian whitening of the IMU factors with compile-time shapes: row i of R J is accumulated in registers (COLS static
// accumulators) while the rows j >= i of J stream by, every entry of the packed upper sqrt-information is read once
// (Jout != nullptr: the whitened rows go straight to the caller's record in global memory instead of back into J)
template <bool WJ, int ROWS, int COLS>
__device__ __forceinline__ void whiten_gaussian(CstRef cst, double* r, double* J, double* Jout) {
    int off = 70;
    for (int i = 0; i < ROWS; i++) {
        double acc[WJ ? COLS : 1];
        double accr = 0.0;
        if (WJ) {
#pragma unroll
            for (int c = 0; c < COLS; c++) acc[c] = 0.0;
        }
        for (int j = i; j < ROWS; j++) {
            const double rj = cst[off + j - i];
            accr = fma(rj, r[j], accr);
            if (WJ) {
                const double* Jr = J + j * COLS;
#pragma unroll
                for (int c = 0; c < COLS; c++) acc[c] = fma(rj, Jr[c], acc[c]);
            }
        }
        r[i] = accr;
        if (WJ) {
            double* Ji = (Jout ? Jout : J) + i * COLS;
#pragma unroll
            for (int c = 0; c < COLS; c++) Ji[c] = acc[c];
        }
        off += ROWS - i;
    }
}

// returns true if the whitened Jacobian was written to Jout (the caller then skips its own copy)
template <bool WJ>
__device__ bool factor_whiten(int type, const double* params, CstRef cst, double* r, double* J, int cols, double* Jout = nullptr) {
    const int rows = factor_rows(type);
    if (type == B200_F_IMU_COMBINED && cols == 30) { whiten_gaussian<WJ, 15, 30>(cst, r, J, Jout); return WJ && Jout != nullptr; }
    if (type == B200_F_IMU_STATICBIAS && cols == 24) { whiten_gaussian<WJ, 9, 24>(cst, r, J, Jout); return WJ && Jout != nullptr; }
    if (type == B200_F_IMU_COMBINED || type == B200_F_IMU_STATICBIAS) {
        int off = 70;
        for (int i = 0; i < rows; i++) {
            double Rrow[15];
            for (int j = i; j < rows; j++) Rrow[j] = cst[off + j - i];
            double acc = 0;
            for (int j = i; j < rows; j++) acc += Rrow[j] * r[j];
            r[i] = acc;
            if (WJ)
                for (int c = 0; c < cols; c++) {
                    double a = 0;
                    for (int j = i; j < rows; j++) a += Rrow[j] * J[j * cols + c];
                    J[i * cols + c] = a;
                }
            off += rows - i;
        }
    } else {
        for (int i = 0; i < rows; i++) {
            const double s = 1.0 / params[i];
            r[i] *= s;
            if (WJ) for (int c = 0; c < cols; c++) J[i * cols + c] *= s;
        }
    }
    return false;
}

}  // namespace b200d
