// ORACLE -- TEST INFRASTRUCTURE ONLY (see geometry.hpp header).
//
// CPU restatement of GTSAM 4.2 IMU preintegration as bioslam drives it (reference src/imuPoseEstimator.cpp:188-213 for
// CombinedImuFactor, :271-296 for ImuFactor; noise from src/imu/imuNoiseModelHandler.cpp:11-29).  Algorithm restated
// from gtsam/navigation/{TangentPreintegration,CombinedImuFactor,ImuFactor}.cpp @4.2 (GTSAM_TANGENT_PREINTEGRATION=ON,
// the 4.2 default): PARITY UNPINNED against real GTSAM (not installable here).
//
// Output record layout == B200_F_IMU_COMBINED / B200_F_IMU_STATICBIAS consts (include/bioslam_b200.h).
#pragma once
#include "geometry.hpp"

namespace orc {

// so3::DexpFunctor::applyDexp derivative wrt omega (gtsam/geometry/SO3.cpp @4.2), v applied vector
inline void dexp_apply_H_omega(double* H, const double* omega, const double* v) {
    double theta2 = v3dot(omega, omega);
    if (theta2 <= kEps) { double S[9]; skew(S, v); m3scale(H, S, 0.5); return; }
    double theta = std::sqrt(theta2);
    double st = std::sin(theta), s2 = std::sin(theta / 2.0), omc = 2.0 * s2 * s2;
    double a = omc / theta, b = 1.0 - st / theta;
    double W[9], K[9]; skew(W, omega); m3scale(K, W, 1.0 / theta);
    double Kv[3]; m3v(Kv, K, v);
    double Da = (st - 2.0 * a) / theta2;
    double Db = (omc - 3.0 * b) / theta2;
    // (Db*K - Da*I) * Kv * omega^T - skew(Kv*b/theta) + (a*I - b*K) * skew(v/theta)
    double M1[9];
    for (int i = 0; i < 9; i++) M1[i] = Db * K[i];
    M1[0] -= Da; M1[4] -= Da; M1[8] -= Da;
    double M1Kv[3]; m3v(M1Kv, M1, Kv);
    double t[3] = {Kv[0] * b / theta, Kv[1] * b / theta, Kv[2] * b / theta};
    double S1[9]; skew(S1, t);
    double M2[9];
    for (int i = 0; i < 9; i++) M2[i] = -b * K[i];
    M2[0] += a; M2[4] += a; M2[8] += a;
    double vt[3] = {v[0] / theta, v[1] / theta, v[2] / theta};
    double S2[9], M2S2[9]; skew(S2, vt); m3mul(M2S2, M2, S2);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) H[3 * i + j] = M1Kv[i] * omega[j] - S1[3 * i + j] + M2S2[3 * i + j];
}

struct PreintState {
    double zeta[9];       // [theta, p, v]
    double Hba[27];       // 9x3 d zeta / d biasAcc
    double Hbg[27];       // 9x3 d zeta / d biasOmega
    double dtij;
    double cov[225];      // 15x15 (combined) or top-left 9x9 used (ImuFactor), row-major ld 15
};

inline void preint_reset(PreintState& s) { std::memset(&s, 0, sizeof(s)); }

// TangentPreintegration::UpdatePreintegrated + ::update, then the covariance propagation of
// PreintegratedCombinedMeasurements::integrateMeasurement (combined=true) or
// PreintegratedImuMeasurements::integrateMeasurement (combined=false).
inline void preint_integrate(PreintState& s, const double* acc_meas, const double* gyro_meas, double dt, const double* bias_hat,
                             const b200_imu_noise& nz, bool combined) {
    double a[3], w[3];
    v3sub(a, acc_meas, bias_hat);        // correctAccelerometer
    v3sub(w, gyro_meas, bias_hat + 3);   // correctGyroscope
    const double* theta = s.zeta;
    double dexp[9], invDexp[9], R[9];
    so3_dexp(dexp, theta);
    m3inv(invDexp, dexp);
    so3_expmap(R, theta);
    double w_tan[3]; m3v(w_tan, invDexp, w);
    // w_tangent_H_theta = -invDexp * D(applyDexp(c))/D(omega) at c = w_tan
    double Dd[9], wtH[9];
    dexp_apply_H_omega(Dd, theta, w_tan);
    m3mul(wtH, invDexp, Dd);
    for (int i = 0; i < 9; i++) wtH[i] = -wtH[i];
    double a_nav[3]; m3v(a_nav, R, a);
    const double dt22 = 0.5 * dt * dt;
    // a_nav_H_theta = R * skew(-a) * dexp
    double na[3] = {-a[0], -a[1], -a[2]}, Sa[9], RS[9], aH[9];
    skew(Sa, na); m3mul(RS, R, Sa); m3mul(aH, RS, dexp);
    // A (9x9), B (9x3), C (9x3)
    double A[81], B[27], C[27];
    std::memset(A, 0, sizeof(A)); std::memset(B, 0, sizeof(B)); std::memset(C, 0, sizeof(C));
    for (int i = 0; i < 9; i++) A[10 * i] = 1.0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            A[9 * i + j] += wtH[3 * i + j] * dt;
            A[9 * (3 + i) + j] = aH[3 * i + j] * dt22;
            A[9 * (6 + i) + j] = aH[3 * i + j] * dt;
            B[3 * (3 + i) + j] = R[3 * i + j] * dt22;
            B[3 * (6 + i) + j] = R[3 * i + j] * dt;
            C[3 * i + j] = invDexp[3 * i + j] * dt;
        }
    for (int i = 0; i < 3; i++) A[9 * (3 + i) + 6 + i] = dt;
    // mean propagation
    double zn[9];
    for (int i = 0; i < 3; i++) {
        zn[i] = s.zeta[i] + w_tan[i] * dt;
        zn[3 + i] = s.zeta[3 + i] + s.zeta[6 + i] * dt + a_nav[i] * dt22;
        zn[6 + i] = s.zeta[6 + i] + a_nav[i] * dt;
    }
    std::memcpy(s.zeta, zn, sizeof(zn));
    s.dtij += dt;
    // bias Jacobians: H <- A*H - {B,C}
    double Hn[27];
    for (int pass = 0; pass < 2; pass++) {
        double* H = pass == 0 ? s.Hba : s.Hbg;
        const double* Sub = pass == 0 ? B : C;
        for (int i = 0; i < 9; i++)
            for (int j = 0; j < 3; j++) {
                double acc = 0;
                for (int l = 0; l < 9; l++) acc += A[9 * i + l] * H[3 * l + j];
                Hn[3 * i + j] = acc - Sub[3 * i + j];
            }
        std::memcpy(H, Hn, sizeof(Hn));
    }
    // covariance
    const double aCov = nz.accel_noise_sigma * nz.accel_noise_sigma;
    const double wCov = nz.gyro_noise_sigma * nz.gyro_noise_sigma;
    const double iCov = nz.integration_err_sigma * nz.integration_err_sigma;
    if (!combined) {
        // P <- A P A^T + B (aCov/dt) B^T + C (wCov/dt) C^T ; P[3:6,3:6] += iCov*dt
        double AP[81], P[81];
        for (int i = 0; i < 9; i++)
            for (int j = 0; j < 9; j++) { double acc = 0; for (int l = 0; l < 9; l++) acc += A[9 * i + l] * s.cov[15 * l + j]; AP[9 * i + j] = acc; }
        for (int i = 0; i < 9; i++)
            for (int j = 0; j < 9; j++) {
                double acc = 0;
                for (int l = 0; l < 9; l++) acc += AP[9 * i + l] * A[9 * j + l];
                double bb = 0, cc = 0;
                for (int l = 0; l < 3; l++) { bb += B[3 * i + l] * B[3 * j + l]; cc += C[3 * i + l] * C[3 * j + l]; }
                P[9 * i + j] = acc + bb * (aCov / dt) + cc * (wCov / dt);
            }
        for (int i = 0; i < 3; i++) P[9 * (3 + i) + 3 + i] += iCov * dt;
        for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) s.cov[15 * i + j] = P[9 * i + j];
        return;
    }
    const double baCov = nz.accel_bias_rw_sigma * nz.accel_bias_rw_sigma;
    const double bgCov = nz.gyro_bias_rw_sigma * nz.gyro_bias_rw_sigma;
    const double bInt = nz.preint_bias_sigma * nz.preint_bias_sigma;  // biasAccOmegaInt = I6 * bInt
    // F (15x15): [A, [0 theta_H_bg; pos_H_ba 0; vel_H_ba 0]; 0 I6], theta_H_biasOmega=-C[0:3], pos_H_biasAcc=-B[3:6], vel_H_biasAcc=-B[6:9]
    double F[225]; std::memset(F, 0, sizeof(F));
    for (int i = 0; i < 9; i++) for (int j = 0; j < 9; j++) F[15 * i + j] = A[9 * i + j];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            F[15 * i + 12 + j] = -C[3 * i + j];
            F[15 * (3 + i) + 9 + j] = -B[3 * (3 + i) + j];
            F[15 * (6 + i) + 9 + j] = -B[3 * (6 + i) + j];
        }
    for (int i = 9; i < 15; i++) F[16 * i] = 1.0;
    double G[225]; std::memset(G, 0, sizeof(G));
    // D_t_t = dt*iCov ; D_v_v = (1/dt) vHa (aCov + bInt) vHa^T ; D_R_R = (1/dt) tHg (wCov + bInt) tHg^T ; D_a_a = dt*baCov ; D_g_g = dt*bgCov
    // off-diagonal D_v_R = vHa * biasAccOmegaInt.block(3,0) * tHg^T = 0 for the diagonal biasAccOmegaInt bioslam sets
    for (int i = 0; i < 3; i++) {
        G[15 * (3 + i) + 3 + i] = dt * iCov;
        G[15 * (9 + i) + 9 + i] = dt * baCov;
        G[15 * (12 + i) + 12 + i] = dt * bgCov;
        for (int j = 0; j < 3; j++) {
            double vv = 0, rr = 0;
            for (int l = 0; l < 3; l++) { vv += B[3 * (6 + i) + l] * B[3 * (6 + j) + l]; rr += C[3 * i + l] * C[3 * j + l]; }
            G[15 * (6 + i) + 6 + j] = (1.0 / dt) * vv * (aCov + bInt);
            G[15 * i + j] = (1.0 / dt) * rr * (wCov + bInt);
        }
    }
    double FP[225], P[225];
    for (int i = 0; i < 15; i++)
        for (int j = 0; j < 15; j++) { double acc = 0; for (int l = 0; l < 15; l++) acc += F[15 * i + l] * s.cov[15 * l + j]; FP[15 * i + j] = acc; }
    for (int i = 0; i < 15; i++)
        for (int j = 0; j < 15; j++) { double acc = 0; for (int l = 0; l < 15; l++) acc += FP[15 * i + l] * F[15 * j + l]; P[15 * i + j] = acc + G[15 * i + j]; }
    std::memcpy(s.cov, P, sizeof(P));
}

// sqrt information R (upper, R^T R = Sigma^-1) of an n x n covariance (ld 15), packed row-major upper (row r: cols r..n-1).
// noiseModel::Gaussian::Covariance(Sigma) == Information(Sigma^-1) -> R = chol(Sigma^-1).upper.  Computed as the inverse of
// the upper factor U of Sigma = U U^T (identical in exact arithmetic: Sigma^-1 = U^-T U^-1, U^-1 upper with positive diagonal).
inline bool sqrt_info_packed(const double* cov, int n, double* Rp) {
    double U[225];
    std::memset(U, 0, sizeof(U));
    // backward Cholesky: for j = n-1..0: U_jj = sqrt(S_jj - sum_{l>j} U_jl^2) ; U_ij = (S_ij - sum_{l>j} U_il U_jl)/U_jj, i<j
    for (int j = n - 1; j >= 0; j--) {
        double d = cov[15 * j + j];
        for (int l = j + 1; l < n; l++) d -= U[15 * j + l] * U[15 * j + l];
        if (!(d > 0.0)) return false;
        double ujj = std::sqrt(d);
        U[15 * j + j] = ujj;
        for (int i = 0; i < j; i++) {
            double v = cov[15 * i + j];
            for (int l = j + 1; l < n; l++) v -= U[15 * i + l] * U[15 * j + l];
            U[15 * i + j] = v / ujj;
        }
    }
    // R = U^-1 (upper): column by column back substitution
    double Rm[225];
    std::memset(Rm, 0, sizeof(Rm));
    for (int c = 0; c < n; c++) {
        // solve U x = e_c ; x_i = 0 for i > c
        for (int i = c; i >= 0; i--) {
            double v = (i == c) ? 1.0 : 0.0;
            for (int l = i + 1; l <= c; l++) v -= U[15 * i + l] * Rm[15 * l + c];
            Rm[15 * i + c] = v / U[15 * i + i];
        }
    }
    int off = 0;
    for (int r = 0; r < n; r++)
        for (int c = r; c < n; c++) Rp[off++] = Rm[15 * r + c];
    return true;
}

// one record (190 doubles combined / 115 doubles ImuFactor) from n_per consecutive samples
inline bool preintegrate_record(const double* acc /*[n_per][3]*/, const double* gyro, int n_per, double dt, const double* bias_hat,
                                const b200_imu_noise& nz, bool combined, double* out) {
    PreintState s; preint_reset(s);
    for (int j = 0; j < n_per; j++) preint_integrate(s, acc + 3 * j, gyro + 3 * j, dt, bias_hat, nz, combined);
    std::memcpy(out, s.zeta, 9 * sizeof(double));
    std::memcpy(out + 9, s.Hba, 27 * sizeof(double));
    std::memcpy(out + 36, s.Hbg, 27 * sizeof(double));
    std::memcpy(out + 63, bias_hat, 6 * sizeof(double));
    out[69] = s.dtij;
    return sqrt_info_packed(s.cov, combined ? 15 : 9, out + 70);
}

}  // namespace orc
