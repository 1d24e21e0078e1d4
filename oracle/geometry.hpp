// ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
//
// CPU restatement of the GTSAM 4.2 geometry conventions bioslam's factors rely on (GTSAM is an un-vendored
// third-party dependency of the reference, pinned at tag 4.2 -- reference README.md:72-76,
// docker/bioslam1.2-gtsam4.2-ubu23.Dockerfile:27,51).  PARITY UNPINNED against real GTSAM: it cannot be built or
// imported in this environment (no Eigen/Boost/GTSAM, no network); the published algorithms are restated from
// gtsam/geometry/{SO3,Rot3M,Pose3,Unit3}.cpp @4.2 as recalled in SURVEY.md Appendix A.
#pragma once
#include <cmath>
#include <cstring>
#include <limits>

namespace orc {

constexpr double kEps = std::numeric_limits<double>::epsilon();

// ---- 3-vectors / 3x3 row-major matrices on raw arrays -------------------------------------------------
inline void v3set(double* o, double x, double y, double z) { o[0] = x; o[1] = y; o[2] = z; }
inline void v3copy(double* o, const double* a) { o[0] = a[0]; o[1] = a[1]; o[2] = a[2]; }
inline void v3add(double* o, const double* a, const double* b) { for (int i = 0; i < 3; i++) o[i] = a[i] + b[i]; }
inline void v3sub(double* o, const double* a, const double* b) { for (int i = 0; i < 3; i++) o[i] = a[i] - b[i]; }
inline void v3scale(double* o, const double* a, double s) { for (int i = 0; i < 3; i++) o[i] = a[i] * s; }
inline double v3dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double v3norm(const double* a) { return std::sqrt(v3dot(a, a)); }
inline void v3cross(double* o, const double* a, const double* b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
// skew-symmetric [v]x  (gtsam skewSymmetric; reference src/mathutils.cpp:136-146)
inline void skew(double* W, const double* v) {
    W[0] = 0; W[1] = -v[2]; W[2] = v[1];
    W[3] = v[2]; W[4] = 0; W[5] = -v[0];
    W[6] = -v[1]; W[7] = v[0]; W[8] = 0;
}
inline void m3ident(double* M) { std::memset(M, 0, 9 * sizeof(double)); M[0] = M[4] = M[8] = 1.0; }
inline void m3copy(double* o, const double* a) { std::memcpy(o, a, 9 * sizeof(double)); }
inline void m3mul(double* o, const double* a, const double* b) {  // o = a*b (o may not alias)
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
inline void m3mulT(double* o, const double* a, const double* b) {  // o = a * b^T
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
inline void m3Tmul(double* o, const double* a, const double* b) {  // o = a^T * b
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
inline void m3T(double* o, const double* a) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * j + i];
}
inline void m3v(double* o, const double* M, const double* v) {  // o = M v
    double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    double y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
inline void m3Tv(double* o, const double* M, const double* v) {  // o = M^T v
    double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    double y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    double z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
inline void m3scale(double* o, const double* a, double s) { for (int i = 0; i < 9; i++) o[i] = a[i] * s; }
inline void m3add(double* o, const double* a, const double* b) { for (int i = 0; i < 9; i++) o[i] = a[i] + b[i]; }
// closed-form inverse (cofactors), what Eigen's fixed-size 3x3 inverse() computes
inline void m3inv(double* o, const double* a) {
    double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    double id = 1.0 / det;
    o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// ---- SO(3): gtsam/geometry/SO3.cpp @4.2 -----------------------------------------------------------------
// so3::ExpmapFunctor::expmap(): I + sin(t) K + (1-cos t) K^2, K = W/t ; near zero: I + W.
inline void so3_expmap(double* R, const double* w) {
    double theta2 = v3dot(w, w);
    double W[9]; skew(W, w);
    if (theta2 <= kEps) { m3ident(R); for (int i = 0; i < 9; i++) R[i] += W[i]; return; }
    double theta = std::sqrt(theta2);
    double st = std::sin(theta), s2 = std::sin(theta / 2.0), omc = 2.0 * s2 * s2;
    double K[9], KK[9];
    m3scale(K, W, 1.0 / theta); m3mul(KK, K, K);
    m3ident(R);
    for (int i = 0; i < 9; i++) R[i] += st * K[i] + omc * KK[i];
}
// so3::DexpFunctor: right Jacobian dexp = I - a K + b K^2, a=(1-cos t)/t, b=1-sin t/t; near zero I - W/2
inline void so3_dexp(double* D, const double* w) {
    double theta2 = v3dot(w, w);
    double W[9]; skew(W, w);
    m3ident(D);
    if (theta2 <= kEps) { for (int i = 0; i < 9; i++) D[i] -= 0.5 * W[i]; return; }
    double theta = std::sqrt(theta2);
    double st = std::sin(theta), s2 = std::sin(theta / 2.0), omc = 2.0 * s2 * s2;
    double a = omc / theta, b = 1.0 - st / theta;
    double K[9], KK[9];
    m3scale(K, W, 1.0 / theta); m3mul(KK, K, K);
    for (int i = 0; i < 9; i++) D[i] += -a * K[i] + b * KK[i];
}
// SO3::LogmapDerivative: inverse right Jacobian
inline void so3_logmap_derivative(double* D, const double* w) {
    double theta2 = v3dot(w, w);
    m3ident(D);
    if (theta2 <= kEps) return;
    double theta = std::sqrt(theta2);
    double W[9], WW[9]; skew(W, w); m3mul(WW, W, W);
    double c = 1.0 / theta2 - (1.0 + std::cos(theta)) / (2.0 * theta * std::sin(theta));
    for (int i = 0; i < 9; i++) D[i] += 0.5 * W[i] + c * WW[i];
}
// SO3::Logmap (matrix form)
inline void so3_logmap(double* w, const double* R) {
    const double R11 = R[0], R12 = R[1], R13 = R[2], R21 = R[3], R22 = R[4], R23 = R[5], R31 = R[6], R32 = R[7], R33 = R[8];
    double tr = R11 + R22 + R33;
    if (tr + 1.0 < 1e-10) {
        if (std::fabs(R33 + 1.0) > 1e-5) { double s = M_PI / std::sqrt(2.0 + 2.0 * R33); v3set(w, s * R13, s * R23, s * (1.0 + R33)); }
        else if (std::fabs(R22 + 1.0) > 1e-5) { double s = M_PI / std::sqrt(2.0 + 2.0 * R22); v3set(w, s * R12, s * (1.0 + R22), s * R32); }
        else { double s = M_PI / std::sqrt(2.0 + 2.0 * R11); v3set(w, s * (1.0 + R11), s * R21, s * R31); }
    } else {
        double magnitude;
        double tr_3 = tr - 3.0;
        if (tr_3 < -1e-7) { double theta = std::acos((tr - 1.0) / 2.0); magnitude = theta / (2.0 * std::sin(theta)); }
        else { magnitude = 0.5 - tr_3 / 12.0; }
        v3set(w, magnitude * (R32 - R23), magnitude * (R13 - R31), magnitude * (R21 - R12));
    }
}

// ---- Unit3: gtsam/geometry/Unit3.cpp @4.2 -----------------------------------------------------------------
// basis B (3x2 row-major): axis with minimal |component|; b1 = normalize(p x axis), b2 = p x b1
inline void unit3_basis(double* B, const double* p) {
    double axis[3] = {0, 0, 1};
    double mx = std::fabs(p[0]), my = std::fabs(p[1]), mz = std::fabs(p[2]);
    if (mx <= my && mx <= mz) v3set(axis, 1, 0, 0);
    else if (my <= mx && my <= mz) v3set(axis, 0, 1, 0);
    double B1[3], b1[3], b2[3];
    v3cross(B1, p, axis);
    double n = v3norm(B1);
    v3scale(b1, B1, 1.0 / n);
    v3cross(b2, p, b1);
    B[0] = b1[0]; B[1] = b2[0]; B[2] = b1[1]; B[3] = b2[1]; B[4] = b1[2]; B[5] = b2[2];
}
inline void unit3_retract(double* q, const double* p, const double* v2) {
    double B[6]; unit3_basis(B, p);
    double xi[3] = {B[0] * v2[0] + B[1] * v2[1], B[2] * v2[0] + B[3] * v2[1], B[4] * v2[0] + B[5] * v2[1]};
    double theta = v3norm(xi);
    double c = std::cos(theta);
    double y[3];
    if (theta < kEps) { for (int i = 0; i < 3; i++) y[i] = c * p[i] + xi[i]; }
    else { double st = std::sin(theta) / theta; for (int i = 0; i < 3; i++) y[i] = c * p[i] + xi[i] * st; }
    double n = v3norm(y);
    v3scale(q, y, 1.0 / n);
}
// p.localCoordinates(q)
inline void unit3_local(double* v2, const double* p, const double* q) {
    double x = v3dot(p, q);
    double z = 1.0 - x * x;
    double y;
    if (z < kEps) {
        if (x > 0) y = 1.0 - (x - 1.0) / 3.0;
        else { v2[0] = M_PI; v2[1] = 0.0; return; }
    } else {
        y = std::acos(x) / std::sqrt(z);
    }
    double B[6]; unit3_basis(B, p);
    double d[3] = {y * (q[0] - x * p[0]), y * (q[1] - x * p[1]), y * (q[2] - x * p[2])};
    v2[0] = B[0] * d[0] + B[2] * d[1] + B[4] * d[2];
    v2[1] = B[1] * d[0] + B[3] * d[1] + B[5] * d[2];
}

// ---- Pose3: value = R(9 row-major), t(3).  retract = T * Expmap(xi), xi=[w,v] (GTSAM_POSE3_EXPMAP) ----------
inline void pose3_expmap(double* R, double* t, const double* xi) {
    const double* w = xi; const double* v = xi + 3;
    so3_expmap(R, w);
    double theta2 = v3dot(w, w);
    if (theta2 > kEps) {
        double wv = v3dot(w, v);
        double tpar[3] = {w[0] * wv, w[1] * wv, w[2] * wv};
        double wxv[3]; v3cross(wxv, w, v);
        double Rwxv[3]; m3v(Rwxv, R, wxv);
        for (int i = 0; i < 3; i++) t[i] = (wxv[i] - Rwxv[i] + tpar[i]) / theta2;
    } else v3copy(t, v);
}
inline void pose3_retract(double* out, const double* in, const double* xi) {
    double Re[9], te[3];
    pose3_expmap(Re, te, xi);
    double Rn[9], tn[3];
    m3mul(Rn, in, Re);
    m3v(tn, in, te);
    for (int i = 0; i < 3; i++) tn[i] += in[9 + i];
    std::memcpy(out, Rn, 9 * sizeof(double));
    std::memcpy(out + 9, tn, 3 * sizeof(double));
}
// Pose3::Logmap for PriorFactor<Pose3>
inline void pose3_logmap(double* xi, const double* R, const double* t) {
    double w[3]; so3_logmap(w, R);
    double th = v3norm(w);
    if (th < 1e-10) { v3copy(xi, w); v3copy(xi + 3, t); return; }
    double W[9]; skew(W, w); for (int i = 0; i < 9; i++) W[i] /= th;
    double Tan = std::tan(0.5 * th);
    double Wt[3]; m3v(Wt, W, t);
    double WWt[3]; m3v(WWt, W, Wt);
    v3copy(xi, w);
    for (int i = 0; i < 3; i++) xi[3 + i] = t[i] - (0.5 * th) * Wt[i] + (1.0 - th / (2.0 * Tan)) * WWt[i];
}

}  // namespace orc
