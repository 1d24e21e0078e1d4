#include "ekf.h"
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace bioslam_b200 {
namespace ekf {
namespace {

inline void mat4mul(const double* A, const double* B, double* C, bool transposeB) {   // 4x4 row-major
    double T[16];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += A[4 * i + k] * (transposeB ? B[4 * j + k] : B[4 * k + j]);
            T[4 * i + j] = s;
        }
    std::memcpy(C, T, sizeof(T));
}
inline void normalize4(double* x) { const double n = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]); for (int i = 0; i < 4; i++) x[i] /= n; }
inline bool inv3(const double* S, double* Si) {
    const double a = S[0], b = S[1], c = S[2], d = S[3], e = S[4], f = S[5], g = S[6], h = S[7], i = S[8];
    const double det = a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
    if (det == 0.0) return false;
    const double id = 1.0 / det;
    Si[0] = (e * i - f * h) * id; Si[1] = (c * h - b * i) * id; Si[2] = (b * f - c * e) * id;
    Si[3] = (f * g - d * i) * id; Si[4] = (a * i - c * g) * id; Si[5] = (c * d - a * f) * id;
    Si[6] = (d * h - e * g) * id; Si[7] = (b * g - a * h) * id; Si[8] = (a * e - b * d) * id;
    return true;
}

}  // namespace

// reference src/mathutils.cpp:489-568
std::vector<double> forwardEkf(const std::vector<double>& gyros, const std::vector<double>& accels, double dt, const double initX[4], double P[16]) {
    const double noise_gyro = 2.4e-3, noise_accel = 2.83e-2, gravity = 9.81007;
    const size_t N = gyros.size() / 3;
    if (accels.size() != gyros.size() || N == 0) throw std::runtime_error("ekf: gyro / accel size mismatch");
    std::vector<double> allX(N * 4, 0.0);
    double x[4] = {initX[0], initX[1], initX[2], initX[3]};
    std::memcpy(allX.data(), x, sizeof(x));
    for (size_t k = 1; k < N; k++) {
        // 1. propagation
        const double* w = &gyros[3 * k];
        const double Om[16] = {0, -w[0], -w[1], -w[2], w[0], 0, w[2], -w[1], w[1], -w[2], 0, w[0], w[2], w[1], -w[0], 0};
        double F[16];
        for (int i = 0; i < 16; i++) F[i] = ((i % 5 == 0) ? 1.0 : 0.0) + Om[i] * dt * 0.5;
        const double G[12] = {-0.5 * x[1], -0.5 * x[2], -0.5 * x[3], 0.5 * x[0], -0.5 * x[3], 0.5 * x[2],
                              0.5 * x[3], 0.5 * x[0], -0.5 * x[1], -0.5 * x[2], 0.5 * x[1], 0.5 * x[0]};
        double Q[16];
        const double qs = std::pow(noise_gyro * dt, 2.0);
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) Q[4 * i + j] = (G[3 * i] * G[3 * j] + G[3 * i + 1] * G[3 * j + 1] + G[3 * i + 2] * G[3 * j + 2]) * qs;
        double xn[4];
        for (int i = 0; i < 4; i++) xn[i] = F[4 * i] * x[0] + F[4 * i + 1] * x[1] + F[4 * i + 2] * x[2] + F[4 * i + 3] * x[3];
        std::memcpy(x, xn, sizeof(x));
        normalize4(x);
        double FP[16];
        mat4mul(F, P, FP, false);
        mat4mul(FP, F, P, true);
        for (int i = 0; i < 16; i++) P[i] += Q[i];
        // 2. update with the direction of the measured specific force
        const double* a = &accels[3 * k];
        const double an = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        const double ea[3] = {a[0] / an, a[1] / an, a[2] / an};
        const double pred[3] = {2 * (x[1] * x[3] - x[0] * x[2]), 2 * (x[2] * x[3] + x[0] * x[1]), x[0] * x[0] - x[1] * x[1] - x[2] * x[2] + x[3] * x[3]};
        const double y[3] = {ea[0] - pred[0], ea[1] - pred[1], ea[2] - pred[2]};
        const double H[12] = {-2 * x[2], 2 * x[3], -2 * x[0], 2 * x[1], 2 * x[1], 2 * x[0], 2 * x[3], 2 * x[2], 2 * x[0], -2 * x[1], -2 * x[2], 2 * x[3]};
        const double r = std::pow(noise_accel / an, 2.0) + std::pow(1.0 - gravity / an, 2.0);
        double PHt[12];   // 4x3
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 3; j++) {
                double s = 0.0;
                for (int c = 0; c < 4; c++) s += P[4 * i + c] * H[4 * j + c];
                PHt[3 * i + j] = s;
            }
        double S[9], Si[9];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) {
                double s = 0.0;
                for (int c = 0; c < 4; c++) s += H[4 * i + c] * PHt[3 * c + j];
                S[3 * i + j] = s + (i == j ? r : 0.0);
            }
        if (!inv3(S, Si)) throw std::runtime_error("ekf: singular innovation covariance");
        double Kg[12];    // 4x3
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 3; j++) Kg[3 * i + j] = PHt[3 * i] * Si[j] + PHt[3 * i + 1] * Si[3 + j] + PHt[3 * i + 2] * Si[6 + j];
        for (int i = 0; i < 4; i++) x[i] += Kg[3 * i] * y[0] + Kg[3 * i + 1] * y[1] + Kg[3 * i + 2] * y[2];
        double IKH[16];
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++)
                IKH[4 * i + j] = (i == j ? 1.0 : 0.0) - (Kg[3 * i] * H[j] + Kg[3 * i + 1] * H[4 + j] + Kg[3 * i + 2] * H[8 + j]);
        mat4mul(IKH, P, P, false);
        normalize4(x);
        for (int i = 0; i < 4; i++)
            for (int j = i + 1; j < 4; j++) { const double m = 0.5 * (P[4 * i + j] + P[4 * j + i]); P[4 * i + j] = m; P[4 * j + i] = m; }
        std::memcpy(&allX[4 * k], x, sizeof(x));
    }
    return allX;
}

// reference src/mathutils.cpp:570-610: every pass = a forward run from (1,0,0,0) with the covariance the previous pass ended
// with, then a run over the time-reversed data (gyro negated) started from the forward run's last state
std::vector<double> forwardBackwardEkf(const std::vector<double>& gyros, const std::vector<double>& accels, double dt, unsigned numFBPasses) {
    const size_t N = gyros.size() / 3;
    const double initX[4] = {1.0, 0.0, 0.0, 0.0};
    double initP[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::vector<double> rg(gyros.size()), ra(accels.size()), X;
    for (size_t k = 0; k < N; k++)
        for (int c = 0; c < 3; c++) { rg[3 * k + c] = -gyros[3 * (N - 1 - k) + c]; ra[3 * k + c] = accels[3 * (N - 1 - k) + c]; }
    for (unsigned pass = 0; pass < numFBPasses; pass++) {
        double P[16];
        std::memcpy(P, initP, sizeof(P));
        X = forwardEkf(gyros, accels, dt, initX, P);
        const double lastX[4] = {X[4 * (N - 1)], X[4 * (N - 1) + 1], X[4 * (N - 1) + 2], X[4 * (N - 1) + 3]};
        std::vector<double> Xb = forwardEkf(rg, ra, dt, lastX, P);
        for (size_t k = 0; k < N; k++) std::memcpy(&X[4 * k], &Xb[4 * (N - 1 - k)], 4 * sizeof(double));
        std::memcpy(initP, P, sizeof(P));
    }
    return X;
}

void quatToRotBodyToNav(const double q[4], double R[9]) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    // R[N->B] (Eigen / gtsam quaternion -> matrix), returned transposed
    const double M[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                         2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                         2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) R[3 * i + j] = M[3 * j + i];
}

}  // namespace ekf
}  // namespace bioslam_b200
