// oracle_gtsam: GTSAM 4.2 (the real reference) and libbioslam_b200 on the SAME graph and initial values, iteration by iteration.
// NOT compiled in this repository's image (GTSAM absent).  Usage: oracle_gtsam <APDM-v5 trial.h5>
//   builds the 7-IMU lowerBodyPoseEstimator exactly like examples/exampleLowerBodyEstimation.cpp (reference :14-66), then
//   (A) runs gtsam::LevenbergMarquardtOptimizer with the estimator's own parameters (reference src/lowerBodyPoseEstimator.cpp:644-667),
//   (B) flattens m_graph / m_initialValues (b200Adapter) and runs b200_optimize,
//   and prints per iteration: error (A), error (B), relative difference; at the end: iterations, final error, max |value difference|.
#include <gtsam/nonlinear/LevenbergMarquardtOptimizer.h>
#include <cstdio>
#include "b200Adapter.h"
#include "imu/imu.h"
#include "imuPoseEstimator.h"
#include "lowerBodyPoseEstimator.h"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s trial.h5\n", argv[0]); return 2; }
    const std::string file = argv[1];
    std::map<std::string, imu> ImuMap = imu::getImuMapFromDataFile(file);
    const char* labels[7] = {"Sacrum", "Right Thigh", "Right Tibia", "Right Foot", "Left Thigh", "Left Tibia", "Left Foot"};
    std::vector<imuPoseEstimator> est;
    for (int i = 0; i < 7; i++) {
        est.emplace_back(ImuMap[labels[i]], std::string("imu") + std::to_string(i));
        est.back().setImuBiasModelMode(1);     // dynamic bias (examples/exampleLowerBodyEstimation.cpp:24,40)
        est.back().setup();
    }
    lowerBodyPoseEstimator lb(est[0], est[1], est[2], est[3], est[4], est[5], est[6], "lowerbody");
    lb.setup();
    // ---- (A) GTSAM, the reference's own loop
    gtsam::LevenbergMarquardtParams params;
    params.setVerbosityLM("SUMMARY"); params.setlambdaInitial(lb.m_lambdaInitial); params.setlambdaFactor(lb.m_lambdaFactor);
    params.setlambdaLowerBound(lb.m_lambdaLowerBound); params.setlambdaUpperBound(lb.m_lambdaUpperBound);
    params.setRelativeErrorTol(1.0e-20); params.setAbsoluteErrorTol(1.0e-20); params.setLinearSolverType("MULTIFRONTAL_CHOLESKY");
    gtsam::LevenbergMarquardtOptimizer opt(lb.m_graph, lb.m_initialValues, params);
    std::vector<double> errA{opt.error()};
    int consec = 0;
    while (consec < (int)lb.m_convergeAtConsecSmallIter && (int)opt.iterations() < (int)lb.m_maxIterations) {
        const double prev = opt.error();
        opt.iterate();
        const double cur = opt.error(), absDec = prev - cur, relDec = absDec / prev;
        consec = (absDec < lb.m_absErrDecreaseLimit || relDec < lb.m_relErrDecreaseLimit || cur < lb.m_absErrLimit) ? consec + 1 : 0;
        errA.push_back(cur);
    }
    // ---- (B) libbioslam_b200 behind the same graph
    b200::FlatGraph fg = b200::flatten(lb.m_graph, lb.m_initialValues);
    b200::RunResult rb = b200::optimizeLowerBody(fg, lb.m_lambdaInitial, lb.m_lambdaFactor, lb.m_lambdaLowerBound, lb.m_lambdaUpperBound, lb.m_absErrDecreaseLimit,
                                                 lb.m_relErrDecreaseLimit, lb.m_absErrLimit, (int)lb.m_maxIterations, (int)lb.m_convergeAtConsecSmallIter);
    const size_t n = std::min(errA.size(), rb.errorHistory.size());
    for (size_t i = 0; i < n; i++) std::printf("it %4zu  gtsam %.12e  b200 %.12e  rel %.3e\n", i, errA[i], rb.errorHistory[i], std::abs(errA[i] - rb.errorHistory[i]) / errA[i]);
    const gtsam::Values vb = b200::unflatten(fg, lb.m_initialValues), va = opt.values();
    double dmax = 0;
    for (const gtsam::Key k : va.keys()) dmax = std::max(dmax, va.at(k).localCoordinates_(vb.at(k)).cwiseAbs().maxCoeff());
    std::printf("gtsam: %zu iterations, final %.12e | b200: %d iterations (%d lambda trials, %.1f ms on the device), final %.12e | max |local(a, b)| %.3e\n",
                errA.size() - 1, errA.back(), rb.iterations, rb.trials, rb.deviceMs, rb.finalError, dmax);
    return 0;
}
