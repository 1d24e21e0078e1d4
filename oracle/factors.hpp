// ORACLE -- TEST INFRASTRUCTURE ONLY (see geometry.hpp header).
//
// CPU restatement of every factor on the hot path: unwhitened residual and
// analytic Jacobian, then whitening.  bioslam-owned factors follow reference src/factors/*.cpp and
// src/mathutils.cpp (cited per function); GTSAM-owned factors (CombinedImuFactor, ImuFactor, PriorFactor<T>) follow
// gtsam @4.2 as recalled in SURVEY.md Appendix A/B -- PARITY UNPINNED for those (GTSAM cannot run here).
// The reference's own oracle for these factors is "analytic Jacobian == numerical derivative" + "optimises to zero"
// (SURVEY.md section 4); tests/test_oracle_factors.py applies exactly that to this file.
#pragma once
#include "../include/bioslam_b200.h"
#include "geometry.hpp"

namespace orc {

struct FactorInfo {
    int rows;
    int nvars;
    int vartype[B200_MAX_FACTOR_VARS];
    int nconst;  // per-keyframe constants
    int whiten;  // 0: diagonal sigmas in params[0..rows), 1: packed upper sqrt-information at consts[whiten_off]
    int whiten_off;
};

static const int kValDim[4] = {12, 3, 6, 3};
static const int kTanDim[4] = {6, 3, 6, 2};

inline const FactorInfo& factor_info(int type) {
    static const FactorInfo T[B200_F_NUM_TYPES] = {
        /* IMU_COMBINED   */ {15, 6, {B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_BIAS6, B200_VAR_BIAS6}, 190, 1, 70},
        /* IMU_STATICBIAS */ {9, 5, {B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_BIAS6}, 115, 1, 70},
        /* ANGVEL         */ {3, 2, {B200_VAR_VEC3, B200_VAR_BIAS6}, 3, 0, 0},
        /* HINGE_VEC      */ {3, 5, {B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_UNIT3}, 0, 0, 0},
        /* HINGE_NORM     */ {1, 5, {B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_UNIT3}, 0, 0, 0},
        /* JOINTPOS_VEC   */ {3, 4, {B200_VAR_POSE3, B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_VEC3}, 0, 0, 0},
        /* JOINTPOS_NORM  */ {1, 4, {B200_VAR_POSE3, B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_VEC3}, 0, 0, 0},
        /* JOINTVEL_VEC   */ {3, 8, {B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_VEC3, B200_VAR_VEC3, B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_VEC3, B200_VAR_VEC3}, 0, 0, 0},
        /* SEGLEN_MAG     */ {1, 2, {B200_VAR_VEC3, B200_VAR_VEC3}, 0, 0, 0},
        /* SEGLEN_MAX     */ {1, 2, {B200_VAR_VEC3, B200_VAR_VEC3}, 0, 0, 0},
        /* SEGLEN_MIN     */ {1, 2, {B200_VAR_VEC3, B200_VAR_VEC3}, 0, 0, 0},
        /* SEGLEN_DISCREP */ {1, 4, {B200_VAR_VEC3, B200_VAR_VEC3, B200_VAR_VEC3, B200_VAR_VEC3}, 0, 0, 0},
        /* ANGLE_AXIS_SEG */ {1, 3, {B200_VAR_UNIT3, B200_VAR_VEC3, B200_VAR_VEC3}, 0, 0, 0},
        /* POSE3_TRANS_PR */ {3, 1, {B200_VAR_POSE3}, 0, 0, 0},
        /* POSE3_COMPASS  */ {1, 1, {B200_VAR_POSE3}, 0, 0, 0},
        /* PRIOR_VEC3     */ {3, 1, {B200_VAR_VEC3}, 0, 0, 0},
        /* PRIOR_BIAS6    */ {6, 1, {B200_VAR_BIAS6}, 0, 0, 0},
        /* PRIOR_UNIT3    */ {2, 1, {B200_VAR_UNIT3}, 0, 0, 0},
        /* PRIOR_POSE3    */ {6, 1, {B200_VAR_POSE3}, 0, 0, 0},
        /* MAG_POSE3      */ {3, 1, {B200_VAR_POSE3}, 3, 0, 0},
        /* JOINTVEL_NORM  */ {1, 8, {B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_VEC3, B200_VAR_VEC3, B200_VAR_POSE3, B200_VAR_VEC3, B200_VAR_VEC3, B200_VAR_VEC3}, 0, 0, 0},
    };
    return T[type];
}

inline int factor_ncols(int type) {
    const FactorInfo& fi = factor_info(type);
    int n = 0;
    for (int i = 0; i < fi.nvars; i++) n += kTanDim[fi.vartype[i]];
    return n;
}

// write a 3x3 block (row-major src) into J at (r0,c0)
inline void put33(double* J, int ld, int r0, int c0, const double* M, double s = 1.0) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) J[(r0 + i) * ld + c0 + j] = s * M[3 * i + j];
}
inline void putI3(double* J, int ld, int r0, int c0, double s) {
    for (int i = 0; i < 3; i++) J[(r0 + i) * ld + c0 + i] = s;
}

// gradient of gtsam::norm3 (gtsam/geometry/Point3.cpp @4.2): v / |v|, and (1, 1, 1) when |v| <= 1e-10.  Every norm of the
// reference's factors goes through it (src/mathutils.cpp:75,754; src/factors/*.cpp), so degenerate geometry keeps a finite Jacobian.
inline void norm3_grad(const double* v, double n, double* u) {
    if (std::fabs(n) > 1e-10) { u[0] = v[0] / n; u[1] = v[1] / n; u[2] = v[2] / n; }
    else { u[0] = 1.0; u[1] = 1.0; u[2] = 1.0; }
}

// mathutils::unsignedAngle (reference src/mathutils.cpp:63-91): atan2(||a x b||, a.b) with Jacobians
inline double unsigned_angle(const double* a, const double* b, double* Ha, double* Hb) {
    double c[3]; v3cross(c, a, b);
    double y = v3norm(c), x = v3dot(a, b);
    if (Ha || Hb) {
        double D = x * x + y * y;
        double dfdy = x / D, dfdx = -y / D;
        double dydc[3]; norm3_grad(c, y, dydc);
        // dc/da = -[b]x, dc/db = [a]x ; dx/da = b^T, dx/db = a^T
        if (Ha) {
            double Sb[9]; skew(Sb, b);
            for (int j = 0; j < 3; j++) {
                double t = 0; for (int i = 0; i < 3; i++) t += dydc[i] * (-Sb[3 * i + j]);
                Ha[j] = dfdy * t + dfdx * b[j];
            }
        }
        if (Hb) {
            double Sa[9]; skew(Sa, a);
            for (int j = 0; j < 3; j++) {
                double t = 0; for (int i = 0; i < 3; i++) t += dydc[i] * Sa[3 * i + j];
                Hb[j] = dfdy * t + dfdx * a[j];
            }
        }
    }
    return std::atan2(y, x);
}
// mathutils::signedAngle (reference src/mathutils.cpp:93-119)
inline double signed_angle(const double* a, const double* b, const double* ref, double* Ha, double* Hb) {
    double ang = unsigned_angle(a, b, Ha, Hb);
    double axb[3]; v3cross(axb, a, b);
    if (unsigned_angle(ref, axb, nullptr, nullptr) > M_PI / 2) {
        ang = -ang;
        if (Ha) for (int i = 0; i < 3; i++) Ha[i] = -Ha[i];
        if (Hb) for (int i = 0; i < 3; i++) Hb[i] = -Hb[i];
    }
    return ang;
}

// ---- hinge residual m, e and the pieces the Jacobians need (reference src/factors/HingeJointFactors.cpp:51-109)
struct HingeTmp { double m[3], e[3], RAtRB[9], wBA[3], P[9], B[6], k[3]; };
inline void hinge_core(HingeTmp& h, const double* xA, const double* wA, const double* xB, const double* wB, const double* k) {
    m3Tmul(h.RAtRB, xA, xB);          // R_A^T R_B
    m3v(h.wBA, h.RAtRB, wB);          // R_A^T R_B w_B
    v3sub(h.m, h.wBA, wA);
    double mk = v3dot(h.m, k);
    for (int i = 0; i < 3; i++) h.e[i] = h.m[i] - mk * k[i];  // m - projmk(m,k)   (src/mathutils.cpp:717-733)
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) h.P[3 * i + j] = (i == j ? 1.0 : 0.0) - k[i] * k[j];
    unit3_basis(h.B, k);
    v3copy(h.k, k);
}
// 3x20 Jacobian of the vector hinge error: cols [xA(6) wA(3) xB(6) wB(3) k(2)]
inline void hinge_jac(const HingeTmp& h, const double* wB, double* J /*3x20*/) {
    const int ld = 20;
    std::memset(J, 0, sizeof(double) * 3 * ld);
    double S[9], T[9];
    skew(S, h.wBA); m3mul(T, h.P, S); put33(J, ld, 0, 0, T);                 // P [R_A^T R_B w_B]x
    put33(J, ld, 0, 6, h.P, -1.0);                                           // -P
    double PR[9]; m3mul(PR, h.P, h.RAtRB);
    skew(S, wB); m3mul(T, PR, S); put33(J, ld, 0, 9, T, -1.0);               // -P R_A^T R_B [w_B]x
    put33(J, ld, 0, 15, PR);                                                 // P R_A^T R_B
    // de/dk(R3) = -(k m^T + (m.k) I); times basis B (3x2)
    double mk = v3dot(h.m, h.k);
    for (int i = 0; i < 3; i++)
        for (int c = 0; c < 2; c++) {
            double s = 0;
            for (int j = 0; j < 3; j++) s += -(h.k[i] * h.m[j] + (i == j ? mk : 0.0)) * h.B[2 * j + c];
            J[i * ld + 18 + c] = s;
        }
}

// Unwhitened residual r (rows) and Jacobian J (rows x ncols row-major, may be null).
inline void factor_eval(int type, const double* params, const double* consts, const double* const* x, double* r, double* J) {
    const FactorInfo& fi = factor_info(type);
    const int ld = factor_ncols(type);
    if (J) std::memset(J, 0, sizeof(double) * fi.rows * ld);
    switch (type) {
    case B200_F_IMU_COMBINED:
    case B200_F_IMU_STATICBIAS: {
        // gtsam::CombinedImuFactor::evaluateError / PreintegrationBase::computeErrorAndJacobians @4.2
        // (tangent preintegration), closed forms of SURVEY.md Appendix B.8.
        const bool comb = (type == B200_F_IMU_COMBINED);
        const double *Ti = x[0], *vi = x[1], *Tj = x[2], *vj = x[3], *bi = x[4];
        const double* bj = comb ? x[5] : nullptr;
        const double *zeta = consts, *Hba = consts + 9, *Hbg = consts + 36, *bhat = consts + 63;
        const double dt = consts[69];
        const double* g = params;
        double dba[3], dbg[3];
        v3sub(dba, bi, bhat); v3sub(dbg, bi + 3, bhat + 3);
        double zc[9];
        for (int i = 0; i < 9; i++)
            zc[i] = zeta[i] + Hba[3 * i] * dba[0] + Hba[3 * i + 1] * dba[1] + Hba[3 * i + 2] * dba[2] +
                    Hbg[3 * i] * dbg[0] + Hbg[3 * i + 1] * dbg[1] + Hbg[3 * i + 2] * dbg[2];
        const double *th = zc, *p = zc + 3, *v = zc + 6;
        const double *Ri = Ti, *pi = Ti + 9, *Rj = Tj, *pj = Tj + 9;
        double E[9], A[9], M[9];
        so3_expmap(E, th);
        m3Tmul(A, Rj, Ri);
        m3mul(M, A, E);
        double rR[3]; so3_logmap(rR, M);
        double tmp[3], rp[3], rv[3], Rip[3], Riv[3];
        m3v(Rip, Ri, p); m3v(Riv, Ri, v);
        for (int i = 0; i < 3; i++) tmp[i] = pi[i] + Rip[i] + dt * vi[i] + 0.5 * dt * dt * g[i] - pj[i];
        m3Tv(rp, Rj, tmp);
        for (int i = 0; i < 3; i++) tmp[i] = vi[i] + Riv[i] + dt * g[i] - vj[i];
        m3Tv(rv, Rj, tmp);
        v3copy(r, rR); v3copy(r + 3, rp); v3copy(r + 6, rv);
        if (comb) for (int i = 0; i < 6; i++) r[9 + i] = bi[i] - bj[i];
        if (J) {
            double Jri[9], Jrth[9], T[9], T2[9], S[9], RjT[9];
            so3_logmap_derivative(Jri, rR);
            so3_dexp(Jrth, th);
            m3T(RjT, Rj);
            // pose_i
            m3mulT(T, Jri, E); put33(J, ld, 0, 0, T);
            skew(S, p); m3mul(T, A, S); put33(J, ld, 3, 0, T, -1.0);
            put33(J, ld, 3, 3, A);
            skew(S, v); m3mul(T, A, S); put33(J, ld, 6, 0, T, -1.0);
            // vel_i
            put33(J, ld, 3, 6, RjT, dt);
            put33(J, ld, 6, 6, RjT);
            // pose_j
            m3mulT(T, Jri, M); put33(J, ld, 0, 9, T, -1.0);
            skew(S, rp); put33(J, ld, 3, 9, S);
            putI3(J, ld, 3, 12, -1.0);
            skew(S, rv); put33(J, ld, 6, 9, S);
            // vel_j
            put33(J, ld, 6, 15, RjT, -1.0);
            // bias_i (cols 18..23): [H_ba | H_bg] blocks
            m3mul(T, Jri, Jrth);
            for (int blk = 0; blk < 3; blk++) {          // row block: theta, p, v
                const double* L = (blk == 0) ? T : A;
                for (int half = 0; half < 2; half++) {   // acc / gyro
                    const double* H = (half == 0 ? Hba : Hbg) + 9 * blk;  // 3x3 row-major slice rows 3*blk..
                    m3mul(T2, L, H);
                    put33(J, ld, 3 * blk, 18 + 3 * half, T2);
                }
            }
            if (comb) {
                for (int i = 0; i < 6; i++) { J[(9 + i) * ld + 18 + i] = 1.0; J[(9 + i) * ld + 24 + i] = -1.0; }
            }
        }
    } break;
    case B200_F_ANGVEL: {  // src/factors/AngularVelocityFactor.cpp:18-31
        const double *w = x[0], *b = x[1];
        for (int i = 0; i < 3; i++) r[i] = w[i] + b[3 + i] - consts[i];
        if (J) { putI3(J, ld, 0, 0, 1.0); putI3(J, ld, 0, 3 + 3, 1.0); }
    } break;
    case B200_F_HINGE_VEC:
    case B200_F_HINGE_NORM: {
        HingeTmp h;
        hinge_core(h, x[0], x[1], x[2], x[3], x[4]);
        if (type == B200_F_HINGE_VEC) {
            v3copy(r, h.e);
            if (J) hinge_jac(h, x[3], J);
        } else {  // src/factors/HingeJointFactors.cpp:155-173: norm3 chain
            double n = v3norm(h.e);
            r[0] = n;
            if (J) {
                double J3[60]; hinge_jac(h, x[3], J3);
                double u[3]; norm3_grad(h.e, n, u);
                for (int c = 0; c < 20; c++) J[c] = u[0] * J3[c] + u[1] * J3[20 + c] + u[2] * J3[40 + c];
            }
        }
    } break;
    case B200_F_JOINTPOS_VEC:
    case B200_F_JOINTPOS_NORM: {  // mathutils::ptSeparation src/mathutils.cpp:735-747 ; cols [xA(6) xB(6) sA(3) sB(3)]
        const double *xA = x[0], *xB = x[1], *sA = x[2], *sB = x[3];
        double a[3], b[3], e[3];
        m3v(a, xA, sA); m3v(b, xB, sB);
        for (int i = 0; i < 3; i++) e[i] = a[i] + xA[9 + i] - b[i] - xB[9 + i];
        double J3[54];
        if (J) {
            std::memset(J3, 0, sizeof(J3));
            double S[9], T[9];
            skew(S, sA); m3mul(T, xA, S); put33(J3, 18, 0, 0, T, -1.0); put33(J3, 18, 0, 3, xA);
            skew(S, sB); m3mul(T, xB, S); put33(J3, 18, 0, 6, T); put33(J3, 18, 0, 9, xB, -1.0);
            put33(J3, 18, 0, 12, xA); put33(J3, 18, 0, 15, xB, -1.0);
        }
        if (type == B200_F_JOINTPOS_VEC) {
            v3copy(r, e);
            if (J) std::memcpy(J, J3, sizeof(J3));
        } else {  // ConstrainedJointCenterNormPositionFactor: ||e|| - expectedDist
            double n = v3norm(e);
            r[0] = n - params[1];
            if (J) { double u[3]; norm3_grad(e, n, u); for (int c = 0; c < 18; c++) J[c] = u[0] * J3[c] + u[1] * J3[18 + c] + u[2] * J3[36 + c]; }
        }
    } break;
    case B200_F_JOINTVEL_VEC:
    case B200_F_JOINTVEL_NORM: {  // src/factors/ConstrainedJointCenterVelocityFactor.cpp:68-131 (vector error), :137-214 (its norm)
        // cols: xA(6) vA(3) wA(3) sA(3) xB(6) vB(3) wB(3) sB(3) = 30
        const bool vec = type == B200_F_JOINTVEL_VEC;
        double e[3], J3[90] = {0};
        double* Jt = vec ? J : J3;
        const int ldt = vec ? ld : 30;
        for (int side = 0; side < 2; side++) {
            const double *X = x[4 * side], *V = x[4 * side + 1], *W = x[4 * side + 2], *S_ = x[4 * side + 3];
            double sg = side == 0 ? 1.0 : -1.0;
            double ws[3], Rws[3];
            v3cross(ws, W, S_); m3v(Rws, X, ws);
            for (int i = 0; i < 3; i++) { double t = V[i] + Rws[i]; if (side == 0) e[i] = t; else e[i] -= t; }
            if (J) {
                int c0 = 15 * side;
                double Sk[9], T[9];
                skew(Sk, ws); m3mul(T, X, Sk); put33(Jt, ldt, 0, c0, T, -sg);        // -R [w x s]x
                putI3(Jt, ldt, 0, c0 + 6, sg);
                skew(Sk, S_); m3mul(T, X, Sk); put33(Jt, ldt, 0, c0 + 9, T, -sg);    // -R [s]x
                skew(Sk, W); m3mul(T, X, Sk); put33(Jt, ldt, 0, c0 + 12, T, sg);     //  R [w]x
            }
        }
        if (vec) v3copy(r, e);
        else {  // ConstrainedJointCenterNormVelocityFactor: ||e|| through gtsam::norm3
            double n = v3norm(e);
            r[0] = n;
            if (J) { double u[3]; norm3_grad(e, n, u); for (int c = 0; c < 30; c++) J[c] = u[0] * J3[c] + u[1] * J3[30 + c] + u[2] * J3[60 + c]; }
        }
    } break;
    case B200_F_SEGLEN_MAG:
    case B200_F_SEGLEN_MAX:
    case B200_F_SEGLEN_MIN: {  // src/factors/SegmentLengthMagnitudeFactor.cpp:31-65,90-144 ; src/mathutils.cpp:148-180
        double d[3]; v3sub(d, x[0], x[1]);
        double n = v3norm(d);
        double e, dedn;
        if (type == B200_F_SEGLEN_MAG) { e = n - params[1]; dedn = 1.0; }
        else if (type == B200_F_SEGLEN_MAX) {
            double xm = params[1], a = params[2];
            e = 0.0; dedn = 0.0;
            if (n > xm) {
                double t = std::tanh(a * (n - xm));
                e = (n - xm) * t * t;
                dedn = t * t - 2.0 * a * t * (t * t - 1.0) * (n - xm);
            }
        } else {
            double xm = params[1], a = params[2];
            e = 0.0; dedn = 0.0;
            if (n < xm) {
                double t0 = std::tanh(a * (xm - n));
                e = (xm - n) * t0 * t0;
                double t = std::tanh(a * (n - xm));
                dedn = 2.0 * a * t * (t * t - 1.0) * (n - xm) - t * t;
            }
        }
        r[0] = e;
        if (J) { double u[3]; norm3_grad(d, n, u); for (int i = 0; i < 3; i++) { J[i] = dedn * u[i]; J[3 + i] = -dedn * u[i]; } }
    } break;
    case B200_F_SEGLEN_DISCREPANCY: {  // src/factors/SegmentLengthDiscrepancyFactor.cpp:43-72
        double a[3], b[3]; v3sub(a, x[0], x[1]); v3sub(b, x[2], x[3]);
        double na = v3norm(a), nb = v3norm(b);
        r[0] = na - nb;
        if (J) { double ua[3], ub[3]; norm3_grad(a, na, ua); norm3_grad(b, nb, ub); for (int i = 0; i < 3; i++) { J[i] = ua[i]; J[3 + i] = -ua[i]; J[6 + i] = -ub[i]; J[9 + i] = ub[i]; } }
    } break;
    case B200_F_ANGLE_AXIS_SEG: {  // src/factors/AngleBetweenAxisAndSegmentFactor.cpp:38-92 ; cols [k(2) v1(3) v2(3)]
        const double *k = x[0];
        double v[3]; v3sub(v, x[1], x[2]);
        double Hv[3], Hk[3];
        double m = unsigned_angle(v, k, J ? Hv : nullptr, J ? Hk : nullptr);
        r[0] = params[1] - m;
        if (J) {
            double B[6]; unit3_basis(B, k);
            for (int c = 0; c < 2; c++) J[c] = -(Hk[0] * B[c] + Hk[1] * B[2 + c] + Hk[2] * B[4 + c]);
            for (int i = 0; i < 3; i++) { J[2 + i] = -Hv[i]; J[5 + i] = Hv[i]; }
        }
    } break;
    case B200_F_POSE3_TRANSLATION_PRIOR: {  // src/factors/Pose3Priors.cpp:22-32
        for (int i = 0; i < 3; i++) r[i] = x[0][9 + i] - params[3 + i];
        if (J) put33(J, ld, 0, 3, x[0]);
    } break;
    case B200_F_POSE3_COMPASS_PRIOR: {  // src/factors/Pose3Priors.cpp:35-74 ; src/mathutils.cpp:771-781,93-119
        const double* R = x[0];
        const double *lB = params + 2, *ref = params + 5, *up = params + 8;
        double lN[3]; m3v(lN, R, lB);
        double t[3], h[3];
        v3cross(t, lN, up);   // NB: cross(v,n) (the reference's comment says cross(n,v); the code is cross(v,n))
        v3cross(h, t, up);
        double Hh[3];
        double ang = signed_angle(h, ref, up, J ? Hh : nullptr, nullptr);
        r[0] = ang - params[1];
        if (J) {
            double Su[9], SS[9], Sl[9], RS[9], T[9];
            skew(Su, up); m3mul(SS, Su, Su);        // dh/dlN = [up]x[up]x
            skew(Sl, lB); m3mul(RS, R, Sl);         // dlN/dR = -R [lB]x
            m3mul(T, SS, RS);
            for (int j = 0; j < 3; j++) J[j] = -(Hh[0] * T[j] + Hh[1] * T[3 + j] + Hh[2] * T[6 + j]);
        }
    } break;
    case B200_F_PRIOR_VEC3: {  // gtsam::PriorFactor<Vector3>: -Local(x, prior) = x - prior
        for (int i = 0; i < 3; i++) r[i] = x[0][i] - params[3 + i];
        if (J) putI3(J, ld, 0, 0, 1.0);
    } break;
    case B200_F_PRIOR_BIAS6: {
        for (int i = 0; i < 6; i++) r[i] = x[0][i] - params[6 + i];
        if (J) for (int i = 0; i < 6; i++) J[i * ld + i] = 1.0;
    } break;
    case B200_F_PRIOR_UNIT3: {  // -x.localCoordinates(prior), H := I (GTSAM approximation)
        double l[2]; unit3_local(l, x[0], params + 2);
        r[0] = -l[0]; r[1] = -l[1];
        if (J) { J[0] = 1.0; J[ld + 1] = 1.0; }
    } break;
    case B200_F_PRIOR_POSE3: {  // -Logmap(x^-1 * prior), H := I
        const double* R = x[0]; const double* t = x[0] + 9;
        const double* Rp = params + 6; const double* tp = params + 15;
        double dR[9], dt3[3], d[3], xi[6];
        m3Tmul(dR, R, Rp);
        v3sub(d, tp, t); m3Tv(dt3, R, d);
        pose3_logmap(xi, dR, dt3);
        for (int i = 0; i < 6; i++) r[i] = -xi[i];
        if (J) for (int i = 0; i < 6; i++) J[i * ld + i] = 1.0;
    } break;
    case B200_F_MAG_POSE3: {  // src/factors/MagPose3Factor.cpp:25-38
        const double* R = x[0];
        double mN[3]; m3v(mN, R, consts);
        for (int i = 0; i < 3; i++) r[i] = mN[i] - params[3 + i];
        if (J) { double S[9], T[9]; skew(S, consts); m3mul(T, R, S); put33(J, ld, 0, 0, T, -1.0); }
    } break;
    default: break;
    }
}

// whiten in place: diagonal -> divide rows by sigma ; Gaussian -> multiply by packed upper R
inline void factor_whiten(int type, const double* params, const double* consts, double* r, double* J) {
    const FactorInfo& fi = factor_info(type);
    const int rows = fi.rows, ld = factor_ncols(type);
    if (!fi.whiten) {
        for (int i = 0; i < rows; i++) {
            double s = 1.0 / params[i];
            r[i] *= s;
            if (J) for (int c = 0; c < ld; c++) J[i * ld + c] *= s;
        }
    } else {
        const double* Rp = consts + fi.whiten_off;
        int off = 0;
        for (int i = 0; i < rows; i++) {  // row i of upper-triangular R touches rows i..rows-1: safe in place, ascending
            double acc = 0;
            for (int j = i; j < rows; j++) acc += Rp[off + j - i] * r[j];
            r[i] = acc;
            if (J)
                for (int c = 0; c < ld; c++) {
                    double a = 0;
                    for (int j = i; j < rows; j++) a += Rp[off + j - i] * J[j * ld + c];
                    J[i * ld + c] = a;
                }
            off += rows - i;
        }
    }
}

}  // namespace orc
