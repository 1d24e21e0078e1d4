// Device-side fp64 geometry with GTSAM 4.2 conventions: SO(3) exp/log and right Jacobians,
// Unit3 basis / retract / local coordinates, SE(3) exponential.  3x3 matrices are row-major double[9].
#pragma once
#include "cuda_compat.h"

namespace b200d {

#define B200_EPS 2.220446049250313e-16
#define B200_PI 3.14159265358979323846

__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3(double* o, const double* a, const double* b) {
    double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void mat_vec(double* o, const double* M, const double* v) {
    double x = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    double y = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    double z = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void matT_vec(double* o, const double* M, const double* v) {
    double x = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    double y = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    double z = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
    o[0] = x; o[1] = y; o[2] = z;
}
// o = a*b, o = a^T*b, o = a*b^T (o must not alias)
__device__ __forceinline__ void mm(double* o, const double* a, const double* b) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
}
__device__ __forceinline__ void mTm(double* o, const double* a, const double* b) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
__device__ __forceinline__ void mmT(double* o, const double* a, const double* b) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) o[3 * i + j] = a[3 * i] * b[3 * j] + a[3 * i + 1] * b[3 * j + 1] + a[3 * i + 2] * b[3 * j + 2];
}
// o = a * [v]x   (columns: a * skew(v));  a*[v]x = rows_i(a) x v -> row i = cross(a_i, v) with sign: (a [v]x)_{i,:} = -(v x a_i)... computed directly
__device__ __forceinline__ void m_skew(double* o, const double* a, const double* v) {
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double a0 = a[3 * i], a1 = a[3 * i + 1], a2 = a[3 * i + 2];
        // [v]x = [[0,-v2,v1],[v2,0,-v0],[-v1,v0,0]]
        o[3 * i] = a1 * v[2] - a2 * v[1];
        o[3 * i + 1] = a2 * v[0] - a0 * v[2];
        o[3 * i + 2] = a0 * v[1] - a1 * v[0];
    }
}
__device__ __forceinline__ void skew3(double* W, const double* v) {
    W[0] = 0; W[1] = -v[2]; W[2] = v[1];
    W[3] = v[2]; W[4] = 0; W[5] = -v[0];
    W[6] = -v[1]; W[7] = v[0]; W[8] = 0;
}

// I + a*W + b*W^2 for W=[w]x
__device__ __forceinline__ void i_aW_bW2(double* R, const double* w, double a, double b) {
    const double x = w[0], y = w[1], z = w[2];
    const double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = 1.0 - b * (yy + zz); R[1] = -a * z + b * xy;       R[2] = a * y + b * xz;
    R[3] = a * z + b * xy;      R[4] = 1.0 - b * (xx + zz);   R[5] = -a * x + b * yz;
    R[6] = -a * y + b * xz;     R[7] = a * x + b * yz;        R[8] = 1.0 - b * (xx + yy);
}
// so3::ExpmapFunctor::expmap
__device__ __forceinline__ void so3_exp(double* R, const double* w) {
    const double t2 = dot3(w, w);
    if (t2 <= B200_EPS) { i_aW_bW2(R, w, 1.0, 0.0); return; }
    const double t = sqrt(t2);
    const double s2 = sin(0.5 * t);
    i_aW_bW2(R, w, sin(t) / t, 2.0 * s2 * s2 / t2);
}
// so3::DexpFunctor::dexp (right Jacobian): I - a K + b K^2, a=(1-cos)/t, b=1-sin/t, K=W/t
__device__ __forceinline__ void so3_dexp(double* D, const double* w) {
    const double t2 = dot3(w, w);
    if (t2 <= B200_EPS) { i_aW_bW2(D, w, -0.5, 0.0); return; }
    const double t = sqrt(t2);
    const double s2 = sin(0.5 * t);
    i_aW_bW2(D, w, -(2.0 * s2 * s2) / t2, (1.0 - sin(t) / t) / t2);
}
// SO3::LogmapDerivative (inverse right Jacobian)
__device__ __forceinline__ void so3_dlog(double* D, const double* w) {
    const double t2 = dot3(w, w);
    if (t2 <= B200_EPS) { i_aW_bW2(D, w, 0.0, 0.0); return; }
    const double t = sqrt(t2);
    i_aW_bW2(D, w, 0.5, 1.0 / t2 - (1.0 + cos(t)) / (2.0 * t * sin(t)));
}
// SO3::Logmap
__device__ __forceinline__ void so3_log(double* w, const double* R) {
    const double R11 = R[0], R12 = R[1], R13 = R[2], R21 = R[3], R22 = R[4], R23 = R[5], R31 = R[6], R32 = R[7], R33 = R[8];
    const double tr = R11 + R22 + R33;
    if (tr + 1.0 < 1e-10) {
        if (fabs(R33 + 1.0) > 1e-5) { double s = B200_PI / sqrt(2.0 + 2.0 * R33); w[0] = s * R13; w[1] = s * R23; w[2] = s * (1.0 + R33); }
        else if (fabs(R22 + 1.0) > 1e-5) { double s = B200_PI / sqrt(2.0 + 2.0 * R22); w[0] = s * R12; w[1] = s * (1.0 + R22); w[2] = s * R32; }
        else { double s = B200_PI / sqrt(2.0 + 2.0 * R11); w[0] = s * (1.0 + R11); w[1] = s * R21; w[2] = s * R31; }
    } else {
        double mag;
        const double tr_3 = tr - 3.0;
        if (tr_3 < -1e-7) { double th = acos((tr - 1.0) / 2.0); mag = th / (2.0 * sin(th)); }
        else mag = 0.5 - tr_3 / 12.0;
        w[0] = mag * (R32 - R23); w[1] = mag * (R13 - R31); w[2] = mag * (R21 - R12);
    }
}
__device__ __forceinline__ void inv3(double* o, const double* a) {
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double id = 1.0 / (a[0] * c00 + a[1] * c01 + a[2] * c02);
    o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}

// Unit3::basis: B (3x2 row-major)
__device__ __forceinline__ void unit3_basis(double* B, const double* p) {
    double ax[3] = {0.0, 0.0, 1.0};
    const double mx = fabs(p[0]), my = fabs(p[1]), mz = fabs(p[2]);
    if (mx <= my && mx <= mz) { ax[0] = 1.0; ax[2] = 0.0; }
    else if (my <= mx && my <= mz) { ax[1] = 1.0; ax[2] = 0.0; }
    double b1[3], b2[3];
    cross3(b1, p, ax);
    const double inv = 1.0 / sqrt(dot3(b1, b1));
    b1[0] *= inv; b1[1] *= inv; b1[2] *= inv;
    cross3(b2, p, b1);
    B[0] = b1[0]; B[1] = b2[0]; B[2] = b1[1]; B[3] = b2[1]; B[4] = b1[2]; B[5] = b2[2];
}
__device__ __forceinline__ void unit3_retract(double* q, const double* p, const double* v2) {
    double B[6]; unit3_basis(B, p);
    const double xi[3] = {B[0] * v2[0] + B[1] * v2[1], B[2] * v2[0] + B[3] * v2[1], B[4] * v2[0] + B[5] * v2[1]};
    const double th = sqrt(dot3(xi, xi));
    const double c = cos(th);
    const double st = (th < B200_EPS) ? 1.0 : sin(th) / th;
    double y[3] = {c * p[0] + xi[0] * st, c * p[1] + xi[1] * st, c * p[2] + xi[2] * st};
    const double inv = 1.0 / sqrt(dot3(y, y));
    q[0] = y[0] * inv; q[1] = y[1] * inv; q[2] = y[2] * inv;
}
__device__ __forceinline__ void unit3_local(double* v2, const double* p, const double* q) {
    const double x = dot3(p, q);
    const double z = 1.0 - x * x;
    double y;
    if (z < B200_EPS) {
        if (x > 0) y = 1.0 - (x - 1.0) / 3.0;
        else { v2[0] = B200_PI; v2[1] = 0.0; return; }
    } else y = acos(x) / sqrt(z);
    double B[6]; unit3_basis(B, p);
    const double d[3] = {y * (q[0] - x * p[0]), y * (q[1] - x * p[1]), y * (q[2] - x * p[2])};
    v2[0] = B[0] * d[0] + B[2] * d[1] + B[4] * d[2];
    v2[1] = B[1] * d[0] + B[3] * d[1] + B[5] * d[2];
}
// Pose3 retract: T * Expmap([w,v])  (value = R(9) t(3))
__device__ __forceinline__ void pose3_retract(double* out, const double* in, const double* xi) {
    const double* w = xi; const double* v = xi + 3;
    double E[9], te[3];
    so3_exp(E, w);
    const double t2 = dot3(w, w);
    if (t2 > B200_EPS) {
        const double wv = dot3(w, v);
        double wxv[3], Rwxv[3];
        cross3(wxv, w, v);
        mat_vec(Rwxv, E, wxv);
        for (int i = 0; i < 3; i++) te[i] = (wxv[i] - Rwxv[i] + w[i] * wv) / t2;
    } else { te[0] = v[0]; te[1] = v[1]; te[2] = v[2]; }
    double Rn[9], tn[3];
    mm(Rn, in, E);
    mat_vec(tn, in, te);
    for (int i = 0; i < 9; i++) out[i] = Rn[i];
    for (int i = 0; i < 3; i++) out[9 + i] = tn[i] + in[9 + i];
}
// Pose3::Logmap
__device__ __forceinline__ void pose3_log(double* xi, const double* R, const double* t) {
    double w[3]; so3_log(w, R);
    const double th = sqrt(dot3(w, w));
    xi[0] = w[0]; xi[1] = w[1]; xi[2] = w[2];
    if (th < 1e-10) { xi[3] = t[0]; xi[4] = t[1]; xi[5] = t[2]; return; }
    const double k[3] = {w[0] / th, w[1] / th, w[2] / th};
    double Wt[3], WWt[3];
    cross3(Wt, k, t);
    cross3(WWt, k, Wt);
    const double Tan = tan(0.5 * th);
    for (int i = 0; i < 3; i++) xi[3 + i] = t[i] - (0.5 * th) * Wt[i] + (1.0 - th / (2.0 * Tan)) * WWt[i];
}

}  // namespace b200d
