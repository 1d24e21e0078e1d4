// see imuPoseEstimator.h.  Graph topology follows reference src/imuPoseEstimator.cpp:47-60,117-312 line by line.
#include "imuPoseEstimator.h"
#include "ekf.h"
#include <chrono>
#include <iostream>

namespace bioslam_b200 {

static int g_device = 0;
void setDevice(int device) { g_device = device; }
int getDevice() { return g_device; }

static b200_imu_noise opalV2() {  // reference include/imu/imuNoiseModelHandler.h:28-35
    return b200_imu_noise{0.04903325, 0.005905, 0.00015707963267949, 1.93925e-5, 3.162e-3, 1.0e-4};
}

void imuPoseEstimator::setImuBiasModelMode(unsigned modelMode) {
    if (modelMode > 1) throw std::runtime_error("imuBiasModelMode not recognized. choose 0 (static) or 1 (dynamic).");
    imuBiasModelMode = modelMode;
}

void imuPoseEstimator::setup() {  // reference src/imuPoseEstimator.cpp:47-60
    switch (imuBiasModelMode) {
    case 0: setupImuFactorStaticBias(); break;
    case 1: setupCombinedImuFactor(); break;
    default: throw std::runtime_error("imuBiasModelMode not recognized.");
    }
}

std::vector<unsigned> imuPoseEstimator::getImuIndecesCorrespondingToKeyframes() const {  // :62-70
    std::vector<unsigned> idx(m_nKeyframes ? m_nKeyframes : getExpectedNumberOfKeyframesFromImuData());
    if (idx.empty()) return idx;
    idx[0] = 0;
    for (unsigned k = 1; k < idx.size(); k++) idx[k] = k * m_numMeasToPreint - 1;
    return idx;
}

unsigned imuPoseEstimator::getExpectedNumberOfKeyframesFromImuData() const {  // :117-125
    const unsigned long n = m_imu.length();
    return (unsigned)(n / m_numMeasToPreint + (n % m_numMeasToPreint == 0 ? 0 : 1));
}

void imuPoseEstimator::setupCombinedImuFactor() { setupCommon(true); }
void imuPoseEstimator::setupImuFactorStaticBias() { setupCommon(false); }

void imuPoseEstimator::setupCommon(bool combined) {
    if (m_initializationScheme > 1) throw std::runtime_error("unknown initilization scheme. choose 0 or 1.");   // :111-113
    if (m_imu.length() < 2 * (unsigned long)m_numMeasToPreint) throw std::runtime_error("imu data too short for the requested preintegration length");
    const double dt = m_imu.getDeltaT();
    m_nKeyframes = getExpectedNumberOfKeyframesFromImuData();
    const unsigned K = m_nKeyframes;
    // initial values: zeros / identity (getInitializationData(), scheme 0, :85-91)
    m_initPose.assign(K, Pose3());
    if (m_initializationScheme == 1) {
        // forward/backward orientation EKF, 3 passes; it estimates R[N->B], the graph wants R[B->N] (getInitializationData(), :92-110)
        const size_t n = m_imu.length();
        std::vector<double> gy(n * 3), ac(n * 3);
        for (size_t i = 0; i < n; i++) {
            gy[3 * i] = m_imu.gx[i]; gy[3 * i + 1] = m_imu.gy[i]; gy[3 * i + 2] = m_imu.gz[i];
            ac[3 * i] = m_imu.ax[i]; ac[3 * i + 1] = m_imu.ay[i]; ac[3 * i + 2] = m_imu.az[i];
        }
        const std::vector<double> q = ekf::forwardBackwardEkf(gy, ac, dt, 3);
        const std::vector<unsigned> idx = getImuIndecesCorrespondingToKeyframes();
        // upstream's setup loop inserts initPose[k] at keyframe k + 1 (src/imuPoseEstimator.cpp:221-223; keyframe 0 gets initPose[0],
        // :170): keyframes 0 and 1 start from the same EKF sample and keyframe k >= 1 from the sample of keyframe k - 1.  Reproduced
        // literally so that the LM trajectory starts where upstream's does.
        for (unsigned k = 0; k < K; k++) ekf::quatToRotBodyToNav(&q[4 * (size_t)idx[k > 0 ? k - 1 : 0]], m_initPose[k].R.m);
    }
    m_initVel.assign(K, Vector3(0, 0, 0));
    m_initBias.assign(combined ? K : 1, ConstantBias());
    // estimated time domain (:174, :225): t_0 = relTimeSec[0], t_{k+1} = relTimeSec[j] with j = (k+1)*numMeasToPreint
    m_time.assign(K, 0.0);
    m_time[0] = m_imu.relTimeSec[0];
    for (unsigned k = 0; k + 1 < K; k++) {
        const size_t j = (size_t)(k + 1) * m_numMeasToPreint;
        m_time[k + 1] = m_imu.relTimeSec[j < m_imu.relTimeSec.size() ? j : m_imu.relTimeSec.size() - 1];
    }
    // preintegration of every interval on the device (== the integrateMeasurement loop, :200-213 / :283-296), biasHat = initial bias
    const int n = (int)m_imu.length(), nInt = (int)K - 1, rec = combined ? 190 : 115;
    std::vector<double> acc((size_t)n * 3), gyro((size_t)n * 3);
    for (int i = 0; i < n; i++) {
        acc[3 * i] = m_imu.ax[i]; acc[3 * i + 1] = m_imu.ay[i]; acc[3 * i + 2] = m_imu.az[i];
        gyro[3 * i] = m_imu.gx[i]; gyro[3 * i + 1] = m_imu.gy[i]; gyro[3 * i + 2] = m_imu.gz[i];
    }
    const double biasHat[6] = {0, 0, 0, 0, 0, 0};
    const b200_imu_noise nz = opalV2();
    m_preintRecords.assign((size_t)nInt * rec, 0.0);
    const b200_status st = b200_preintegrate(getDevice(), 1, n, (int)m_numMeasToPreint, nInt, acc.data(), gyro.data(), dt, biasHat, &nz,
                                             combined ? 1 : 0, m_preintRecords.data());
    if (st != B200_OK) throw std::runtime_error(std::string("b200_preintegrate failed: ") + b200_last_error());
    m_isSetup = true;
}

void imuPoseEstimator::setOptimizedValuesToInitialValues() {  // :42-45: m_estimate = m_initialValues; setMemberStatesFromValues
    if (!m_isSetup) throw std::runtime_error("setup() must be called first");
    m_pose = m_initPose;
    m_velocity = m_initVel;
    m_orientation.clear(); m_position.clear(); m_accelBias.clear(); m_gyroBias.clear();
    for (const auto& p : m_pose) { m_orientation.push_back(p.R); m_position.push_back(p.t); }
    for (const auto& b : m_initBias) { m_accelBias.push_back(b.acc); m_gyroBias.push_back(b.gyro); }
}

void imuPoseEstimator::clear() { m_pose.clear(); m_orientation.clear(); m_position.clear(); m_velocity.clear(); m_accelBias.clear(); m_gyroBias.clear(); }

imuPoseEstimator::GraphVars imuPoseEstimator::appendToGraph(GraphBuilder& gb) const {
    if (!m_isSetup) throw std::runtime_error("setup() must be called first");
    const int K = (int)m_nKeyframes;
    if (gb.K() != K) throw std::runtime_error("imuPoseEstimator: keyframe count does not match the graph");
    const bool combined = imuBiasModelMode == 1;
    // values: the estimator's current states when present (== m_estimate), else the initial values
    const bool have = (int)m_pose.size() == K;
    std::vector<double> pv((size_t)K * 12), vv((size_t)K * 3);
    for (int k = 0; k < K; k++) {
        const Pose3& P = have ? m_pose[k] : m_initPose[k];
        const Vector3& V = have ? m_velocity[k] : m_initVel[k];
        for (int c = 0; c < 9; c++) pv[(size_t)k * 12 + c] = P.R.m[c];
        for (int c = 0; c < 3; c++) { pv[(size_t)k * 12 + 9 + c] = P.t[c]; vv[(size_t)k * 3 + c] = V[c]; }
    }
    GraphVars gv;
    gv.staticBias = !combined;
    gv.pose = gb.addTimeVar(B200_VAR_POSE3, pv);
    gv.vel = gb.addTimeVar(B200_VAR_VEC3, vv);
    auto biasAt = [&](int k) { return have ? ConstantBias(m_accelBias[k], m_gyroBias[k]) : m_initBias[k]; };
    if (combined) {
        std::vector<double> bv((size_t)K * 6);
        for (int k = 0; k < K; k++) { ConstantBias b = biasAt(k); for (int c = 0; c < 3; c++) { bv[(size_t)k * 6 + c] = b.acc[c]; bv[(size_t)k * 6 + 3 + c] = b.gyro[c]; } }
        gv.bias = gb.addTimeVar(B200_VAR_BIAS6, bv);
    } else {
        ConstantBias b = biasAt(0);
        const double v[6] = {b.acc[0], b.acc[1], b.acc[2], b.gyro[0], b.gyro[1], b.gyro[2]};
        gv.bias = gb.addStaticVar(B200_VAR_BIAS6, v);
    }
    // priors on the first keyframe (setAnyRequestedPriors, :127-150), in the reference's order
    if (m_usePositionPrior)
        gb.addSlot(B200_F_POSE3_TRANSLATION_PRIOR, {{B200_AT_K, gv.pose}},
                   {m_priorPositionNoise[0], m_priorPositionNoise[1], m_priorPositionNoise[2], m_priorPosition[0], m_priorPosition[1], m_priorPosition[2]}, 0, 1);
    if (m_useVelocityPrior)
        gb.addSlot(B200_F_PRIOR_VEC3, {{B200_AT_K, gv.vel}},
                   {m_priorVelocityNoise[0], m_priorVelocityNoise[1], m_priorVelocityNoise[2], m_priorVelocity[0], m_priorVelocity[1], m_priorVelocity[2]}, 0, 1);
    if (m_useImuBiasPrior)
        gb.addSlot(B200_F_PRIOR_BIAS6, {{combined ? B200_AT_K : B200_AT_STATIC, gv.bias}},
                   {m_priorAccelBiasConstantNoise[0], m_priorAccelBiasConstantNoise[1], m_priorAccelBiasConstantNoise[2], m_priorGyroBiasConstantNoise[0],
                    m_priorGyroBiasConstantNoise[1], m_priorGyroBiasConstantNoise[2], m_priorAccelBias[0], m_priorAccelBias[1], m_priorAccelBias[2],
                    m_priorGyroBias[0], m_priorGyroBias[1], m_priorGyroBias[2]}, 0, 1);
    if (m_useCompassPrior) {
        const Vector3 up = m_globalAcc.normalized() * -1.0;
        gb.addSlot(B200_F_POSE3_COMPASS_PRIOR, {{B200_AT_K, gv.pose}},
                   {m_priorCompassAngleNoise, m_priorCompassAngle, m_refVecLocal[0], m_refVecLocal[1], m_refVecLocal[2], m_refVecNav[0], m_refVecNav[1],
                    m_refVecNav[2], up[0], up[1], up[2]}, 0, 1);
    }
    // IMU factors k -> k+1 (:215 / :298) with the preintegrated records as per-keyframe constants
    const int rec = combined ? 190 : 115;
    std::vector<double> cb((size_t)K * rec, 0.0);
    std::copy(m_preintRecords.begin(), m_preintRecords.end(), cb.begin());
    const int coff = gb.addConstBlock(rec, std::move(cb));
    const std::vector<double> g = {m_globalAcc[0], m_globalAcc[1], m_globalAcc[2]};
    if (combined)
        gb.addSlot(B200_F_IMU_COMBINED, {{B200_AT_K, gv.pose}, {B200_AT_K, gv.vel}, {B200_AT_K1, gv.pose}, {B200_AT_K1, gv.vel}, {B200_AT_K, gv.bias}, {B200_AT_K1, gv.bias}}, g, 0, K - 1, coff);
    else
        gb.addSlot(B200_F_IMU_STATICBIAS, {{B200_AT_K, gv.pose}, {B200_AT_K, gv.vel}, {B200_AT_K1, gv.pose}, {B200_AT_K1, gv.vel}, {B200_AT_STATIC, gv.bias}}, g, 0, K - 1, coff);
    if (m_useMagnetometer) {  // MagPose3Factor on pose_{k+1} (:216-220): measured field at sample j=(k+1)*n, B_global = |B| * dir(B) + 0
        std::vector<double> mb((size_t)K * 3, 0.0);
        for (int k = 0; k + 1 < K; k++) {
            const size_t j = std::min((size_t)(k + 1) * m_numMeasToPreint, (size_t)m_imu.length() - 1);
            if (j < m_imu.mx.size()) { mb[(size_t)(k + 1) * 3] = m_imu.mx[j]; mb[(size_t)(k + 1) * 3 + 1] = m_imu.my[j]; mb[(size_t)(k + 1) * 3 + 2] = m_imu.mz[j]; }
        }
        const int moff = gb.addConstBlock(3, std::move(mb));
        gb.addSlot(B200_F_MAG_POSE3, {{B200_AT_K, gv.pose}},
                   {m_magnetometerNoise[0], m_magnetometerNoise[1], m_magnetometerNoise[2], m_globalB[0], m_globalB[1], m_globalB[2]}, 1, K, moff);
    }
    return gv;
}

void imuPoseEstimator::setMemberStatesFromGraph(const GraphBuilder& gb, const GraphVars& gv) {  // == setMemberStatesFromValues, :392-407
    const int K = gb.K();
    m_pose.assign(K, Pose3()); m_orientation.assign(K, Rot3()); m_position.assign(K, Point3()); m_velocity.assign(K, Vector3());
    const int nb = gv.staticBias ? 1 : K;
    m_accelBias.assign(nb, Vector3()); m_gyroBias.assign(nb, Vector3());
    for (int k = 0; k < K; k++) {
        const double* p = gb.timeValue(k, gv.pose);
        for (int c = 0; c < 9; c++) m_pose[k].R.m[c] = p[c];
        for (int c = 0; c < 3; c++) m_pose[k].t[c] = p[9 + c];
        m_orientation[k] = m_pose[k].R; m_position[k] = m_pose[k].t;
        const double* v = gb.timeValue(k, gv.vel);
        m_velocity[k] = Vector3(v[0], v[1], v[2]);
        if (!gv.staticBias) { const double* b = gb.timeValue(k, gv.bias); m_accelBias[k] = Vector3(b[0], b[1], b[2]); m_gyroBias[k] = Vector3(b[3], b[4], b[5]); }
    }
    if (gv.staticBias) { const double* b = gb.staticValue(gv.bias); m_accelBias[0] = Vector3(b[0], b[1], b[2]); m_gyroBias[0] = Vector3(b[3], b[4], b[5]); }
}

// use the iterate() loop of the reference (src/imuPoseEstimator.cpp:317-390) on the device
void imuPoseEstimator::fastOptimize() {
    if (!m_isSetup) throw std::runtime_error("setup() must be called first");
    GraphBuilder gb((int)m_nKeyframes);
    const GraphVars gv = appendToGraph(gb);
    gb.finalize();
    b200_problem* prob = nullptr;
    if (b200_problem_create(&gb.desc(), getDevice(), &prob) != B200_OK) throw std::runtime_error(std::string("b200_problem_create: ") + b200_last_error());
    b200_status st = b200_set_values(prob, gb.timeValues().data(), gb.staticValues().data());
    // LM parameters: :324-325 (lambda in [1e-8, 1e10]); tolerances from the members
    b200_lm_params lm{1.0e-5, 10.0, 1.0e-8, 1.0e10, m_relErrDecreaseLimit, m_absErrDecreaseLimit, 1.0e-3, 0, 0};
    // (gauss_newton = 0 whatever m_optType says: upstream's fastOptimize() always constructs a LevenbergMarquardtOptimizer, the
    // Gauss-Newton line is commented out, src/imuPoseEstimator.cpp:327-329; m_optType is only a stored setting)
    std::vector<double> hist(m_maxIterations + 2, 0.0);
    b200_outer_params outer{m_absErrDecreaseLimit, m_relErrDecreaseLimit, m_absErrLimit, (int)m_maxIterations, 1, 1, (int)hist.size()};
    b200_optimize_result res{};
    const auto t0 = std::chrono::steady_clock::now();
    if (st == B200_OK) st = b200_optimize(prob, &lm, &outer, hist.data(), &res);
    if (st == B200_OK || st == B200_LAMBDA_MAXED || st == B200_INDETERMINATE_SYSTEM)
        b200_get_values(prob, gb.timeValues().data(), gb.staticValues().data());
    b200_problem_destroy(prob);
    if (st == B200_INDETERMINATE_SYSTEM) std::cerr << "IndeterminantLinearSystem in " << m_strId << " (reference src/imuPoseEstimator.cpp:360)" << std::endl;
    else if (st != B200_OK && st != B200_LAMBDA_MAXED) throw std::runtime_error(std::string("b200_optimize: ") + b200_last_error());
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    m_optimizationTotalError.assign(hist.begin(), hist.begin() + res.n_history);
    m_optimizationTotalTime.assign(res.n_history, 0.0);
    for (int i = 0; i < res.n_history; i++) m_optimizationTotalTime[i] = secs * i / std::max(1, res.n_history - 1);
    setMemberStatesFromGraph(gb, gv);
}

}  // namespace bioslam_b200
