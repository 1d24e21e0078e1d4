// IMU preintegration on device: GTSAM 4.2 tangent preintegration as driven by the reference
// (src/imuPoseEstimator.cpp:188-213 CombinedImuFactor, :271-296 ImuFactor).  One thread per (IMU, keyframe interval) folds
// n_per raw samples into zeta=[theta,p,v], d zeta/d bias (9x3, 9x3), deltaTij and the 15x15 (9x9) covariance, then the
// packed upper square-root information.  Output record == consts layout of B200_F_IMU_COMBINED / _STATICBIAS, written
// straight into the keyframe-major SoA constant array (out[(off + i) * stride + interval]).
#pragma once
#include "geometry.cuh"
#include "../../include/bioslam_b200.h"

namespace b200d {

// d(dexp(omega) * v)/d omega  (so3::DexpFunctor::applyDexp, H1)
__device__ inline void dexp_apply_H(double* H, const double* om, const double* v) {
    const double t2 = dot3(om, om);
    if (t2 <= B200_EPS) { skew3(H, v); for (int i = 0; i < 9; i++) H[i] *= 0.5; return; }
    const double t = sqrt(t2);
    const double st = sin(t), s2 = sin(0.5 * t), omc = 2.0 * s2 * s2;
    const double a = omc / t, b = 1.0 - st / t;
    double K[9]; skew3(K, om); for (int i = 0; i < 9; i++) K[i] /= t;
    double Kv[3]; mat_vec(Kv, K, v);
    const double Da = (st - 2.0 * a) / t2, Db = (omc - 3.0 * b) / t2;
    double M1[9]; for (int i = 0; i < 9; i++) M1[i] = Db * K[i];
    M1[0] -= Da; M1[4] -= Da; M1[8] -= Da;
    double M1Kv[3]; mat_vec(M1Kv, M1, Kv);
    const double tt[3] = {Kv[0] * b / t, Kv[1] * b / t, Kv[2] * b / t};
    double S1[9]; skew3(S1, tt);
    double M2[9]; for (int i = 0; i < 9; i++) M2[i] = -b * K[i];
    M2[0] += a; M2[4] += a; M2[8] += a;
    const double vt[3] = {v[0] / t, v[1] / t, v[2] / t};
    double M2S2[9]; m_skew(M2S2, M2, vt);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) H[3 * i + j] = M1Kv[i] * om[j] - S1[3 * i + j] + M2S2[3 * i + j];
}

// acc/gyro: [n_imu][n_samples][3]; bias_hat [n_imu][6]; out element i of (imu m, interval q) at out[(m*rec + i)*stride + q]
__global__ void __launch_bounds__(64) preintegrate_kernel(int n_imu, int n_samples, int n_per, int n_intervals,
                                                          const double* __restrict__ acc, const double* __restrict__ gyro, double dt,
                                                          const double* __restrict__ bias_hat, b200_imu_noise nz, int combined,
                                                          double* __restrict__ out, long long stride, int* __restrict__ bad) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y;
    if (q >= n_intervals || m >= n_imu) return;
    const int n = combined ? 15 : 9;
    const int rec = combined ? 190 : 115;
    const double* bh = bias_hat + 6 * m;
    double z[9], Hba[27], Hbg[27], P[225];
    for (int i = 0; i < 9; i++) z[i] = 0.0;
    for (int i = 0; i < 27; i++) { Hba[i] = 0.0; Hbg[i] = 0.0; }
    for (int i = 0; i < 225; i++) P[i] = 0.0;
    double dtij = 0.0;
    const double aCov = nz.accel_noise_sigma * nz.accel_noise_sigma, wCov = nz.gyro_noise_sigma * nz.gyro_noise_sigma;
    const double iCov = nz.integration_err_sigma * nz.integration_err_sigma;
    const double baCov = nz.accel_bias_rw_sigma * nz.accel_bias_rw_sigma, bgCov = nz.gyro_bias_rw_sigma * nz.gyro_bias_rw_sigma;
    const double bInt = nz.preint_bias_sigma * nz.preint_bias_sigma;
    const double dt22 = 0.5 * dt * dt;
    for (int j = 0; j < n_per; j++) {
        const long long sidx = ((long long)m * n_samples + (long long)q * n_per + j) * 3;
        double a[3], w[3];
        for (int i = 0; i < 3; i++) { a[i] = acc[sidx + i] - bh[i]; w[i] = gyro[sidx + i] - bh[3 + i]; }
        double dexp[9], inv[9], R[9], wt[3], Dd[9], wtH[9], an[3], aH[9], T[9];
        so3_dexp(dexp, z);
        inv3(inv, dexp);
        so3_exp(R, z);
        mat_vec(wt, inv, w);
        dexp_apply_H(Dd, z, wt);
        mm(wtH, inv, Dd);
        for (int i = 0; i < 9; i++) wtH[i] = -wtH[i];
        mat_vec(an, R, a);
        const double na[3] = {-a[0], -a[1], -a[2]};
        m_skew(T, R, na);
        mm(aH, T, dexp);
        // F = [A, cross ; 0, I6] (15x15, or A 9x9)
        double F[225];
        for (int i = 0; i < n * n; i++) F[i] = 0.0;
        for (int i = 0; i < n; i++) F[i * n + i] = 1.0;
        for (int i = 0; i < 3; i++) {
            for (int c = 0; c < 3; c++) {
                F[i * n + c] += wtH[3 * i + c] * dt;
                F[(3 + i) * n + c] = aH[3 * i + c] * dt22;
                F[(6 + i) * n + c] = aH[3 * i + c] * dt;
            }
            F[(3 + i) * n + 6 + i] = dt;
        }
        // B = [0; R dt22; R dt], C = [inv dt; 0; 0]
        // mean
        double zn[9];
        for (int i = 0; i < 3; i++) {
            zn[i] = z[i] + wt[i] * dt;
            zn[3 + i] = z[3 + i] + z[6 + i] * dt + an[i] * dt22;
            zn[6 + i] = z[6 + i] + an[i] * dt;
        }
        for (int i = 0; i < 9; i++) z[i] = zn[i];
        dtij += dt;
        // H <- A H - {B, C}
        double Hn[27];
        for (int pass = 0; pass < 2; pass++) {
            double* H = pass == 0 ? Hba : Hbg;
            for (int i = 0; i < 9; i++)
                for (int c = 0; c < 3; c++) {
                    double s = 0;
                    for (int l = 0; l < 9; l++) s += F[i * n + l] * H[3 * l + c];
                    double sub = 0.0;
                    if (pass == 0) { if (i >= 6) sub = R[3 * (i - 6) + c] * dt; else if (i >= 3) sub = R[3 * (i - 3) + c] * dt22; }
                    else if (i < 3) sub = inv[3 * i + c] * dt;
                    Hn[3 * i + c] = s - sub;
                }
            for (int i = 0; i < 27; i++) H[i] = Hn[i];
        }
        // covariance
        double G[225];
        for (int i = 0; i < n * n; i++) G[i] = 0.0;
        if (combined) {
            for (int i = 0; i < 3; i++)
                for (int c = 0; c < 3; c++) {
                    F[i * n + 12 + c] = -inv[3 * i + c] * dt;          // theta_H_biasOmega = -C
                    F[(3 + i) * n + 9 + c] = -R[3 * i + c] * dt22;      // pos_H_biasAcc = -B[3:6]
                    F[(6 + i) * n + 9 + c] = -R[3 * i + c] * dt;        // vel_H_biasAcc = -B[6:9]
                }
            for (int i = 0; i < 3; i++) {
                G[(3 + i) * n + 3 + i] = dt * iCov;
                G[(9 + i) * n + 9 + i] = dt * baCov;
                G[(12 + i) * n + 12 + i] = dt * bgCov;
                for (int c = 0; c < 3; c++) {
                    double vv = 0, rr = 0;
                    for (int l = 0; l < 3; l++) { vv += R[3 * i + l] * R[3 * c + l]; rr += inv[3 * i + l] * inv[3 * c + l]; }
                    G[(6 + i) * n + 6 + c] = (1.0 / dt) * (vv * dt * dt) * (aCov + bInt);
                    G[i * n + c] = (1.0 / dt) * (rr * dt * dt) * (wCov + bInt);
                }
            }
        } else {
            // B (aCov/dt) B^T + C (wCov/dt) C^T + iCov*dt on the position block
            for (int i = 0; i < 9; i++)
                for (int c = 0; c < 9; c++) {
                    double bb = 0, cc = 0;
                    for (int l = 0; l < 3; l++) {
                        const double bi = i >= 6 ? R[3 * (i - 6) + l] * dt : (i >= 3 ? R[3 * (i - 3) + l] * dt22 : 0.0);
                        const double bc = c >= 6 ? R[3 * (c - 6) + l] * dt : (c >= 3 ? R[3 * (c - 3) + l] * dt22 : 0.0);
                        const double ci = i < 3 ? inv[3 * i + l] * dt : 0.0, cc2 = c < 3 ? inv[3 * c + l] * dt : 0.0;
                        bb += bi * bc; cc += ci * cc2;
                    }
                    G[i * n + c] = bb * (aCov / dt) + cc * (wCov / dt);
                }
            for (int i = 0; i < 3; i++) G[(3 + i) * n + 3 + i] += iCov * dt;
        }
        // P <- F P F^T + G   (through a row buffer: FP row by row would need a full temporary; use two passes)
        double FP[225];
        for (int i = 0; i < n; i++)
            for (int c = 0; c < n; c++) { double s = 0; for (int l = 0; l < n; l++) s += F[i * n + l] * P[l * n + c]; FP[i * n + c] = s; }
        for (int i = 0; i < n; i++)
            for (int c = 0; c < n; c++) { double s = 0; for (int l = 0; l < n; l++) s += FP[i * n + l] * F[c * n + l]; P[i * n + c] = s + G[i * n + c]; }
    }
    // sqrt information: Sigma = U U^T (U upper), R = U^-1
    double U[225], Rm[225];
    for (int i = 0; i < n * n; i++) { U[i] = 0.0; Rm[i] = 0.0; }
    bool ok = true;
    for (int j = n - 1; j >= 0; j--) {
        double d = P[j * n + j];
        for (int l = j + 1; l < n; l++) d -= U[j * n + l] * U[j * n + l];
        if (!(d > 0.0)) { ok = false; d = 1.0; }
        const double ujj = sqrt(d);
        U[j * n + j] = ujj;
        for (int i = 0; i < j; i++) {
            double v = P[i * n + j];
            for (int l = j + 1; l < n; l++) v -= U[i * n + l] * U[j * n + l];
            U[i * n + j] = v / ujj;
        }
    }
    for (int c = 0; c < n; c++)
        for (int i = c; i >= 0; i--) {
            double v = (i == c) ? 1.0 : 0.0;
            for (int l = i + 1; l <= c; l++) v -= U[i * n + l] * Rm[l * n + c];
            Rm[i * n + c] = v / U[i * n + i];
        }
    if (!ok) atomicAdd(bad, 1);
    double* o = out + (long long)m * rec * stride + q;
    for (int i = 0; i < 9; i++) o[(long long)i * stride] = z[i];
    for (int i = 0; i < 27; i++) { o[(long long)(9 + i) * stride] = Hba[i]; o[(long long)(36 + i) * stride] = Hbg[i]; }
    for (int i = 0; i < 6; i++) o[(long long)(63 + i) * stride] = bh[i];
    o[(long long)69 * stride] = dtij;
    int off = 70;
    for (int r = 0; r < n; r++)
        for (int c = r; c < n; c++) o[(long long)(off++) * stride] = Rm[r * n + c];
}

}  // namespace b200d
