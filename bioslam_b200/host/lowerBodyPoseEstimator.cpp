// see lowerBodyPoseEstimator.h.  IMU order everywhere: sacrum, rthigh, rshank, rfoot, lthigh, lshank, lfoot.
#include <cstdio>
#include "lowerBodyPoseEstimator.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <iostream>

namespace bioslam_b200 {

enum { SAC = 0, RT, RS, RF, LT, LS, LF };
// static variable order (12 IMU -> joint-centre vectors, then 11 hinge axes)
enum { sacrumImuToRHipCtr = 0, sacrumImuToLHipCtr, rthighImuToKneeCtr, rthighImuToHipCtr, rshankImuToKneeCtr, rshankImuToAnkleCtr, rfootImuToAnkleCtr,
       lthighImuToKneeCtr, lthighImuToHipCtr, lshankImuToKneeCtr, lshankImuToAnkleCtr, lfootImuToAnkleCtr,
       rkneeAxisThigh, rkneeAxisShank, lkneeAxisThigh, lkneeAxisShank, hipAxisSacrum, rhipAxisThigh, lhipAxisThigh,
       rankleAxisShank, lankleAxisShank, rankleAxisFoot, lankleAxisFoot, N_STATIC };

lowerBodyPoseEstimator::lowerBodyPoseEstimator(const imuPoseEstimator& sacrum, const imuPoseEstimator& rthigh, const imuPoseEstimator& rshank,
                                               const imuPoseEstimator& rfoot, const imuPoseEstimator& lthigh, const imuPoseEstimator& lshank,
                                               const imuPoseEstimator& lfoot, const std::string& id)
    : m_strId(id), m_rthighImuPoseProblem(rthigh), m_rshankImuPoseProblem(rshank), m_rfootImuPoseProblem(rfoot), m_lthighImuPoseProblem(lthigh),
      m_lshankImuPoseProblem(lshank), m_lfootImuPoseProblem(lfoot), m_sacrumImuPoseProblem(sacrum) {}  // copies, like the reference (:579)

imuPoseEstimator* lowerBodyPoseEstimator::imuProblem(int i) {
    imuPoseEstimator* a[7] = {&m_sacrumImuPoseProblem, &m_rthighImuPoseProblem, &m_rshankImuPoseProblem, &m_rfootImuPoseProblem,
                              &m_lthighImuPoseProblem, &m_lshankImuPoseProblem, &m_lfootImuPoseProblem};
    return a[i];
}

static std::vector<double> v3(const Vector3& s) { return {s[0], s[1], s[2]}; }

void lowerBodyPoseEstimator::setup() {
    // all seven IMU problems must share the keyframe grid
    const unsigned K = m_sacrumImuPoseProblem.m_nKeyframes;
    for (int i = 0; i < 7; i++) {
        if (!imuProblem(i)->isSetup()) throw std::runtime_error("lowerBodyPoseEstimator: every imuPoseEstimator must be setup() first");
        if (imuProblem(i)->m_nKeyframes != K) throw std::runtime_error("lowerBodyPoseEstimator: IMU problems have different numbers of keyframes");
        if (imuProblem(i)->m_pose.size() != K) throw std::runtime_error("lowerBodyPoseEstimator: IMU problem has no estimate (call fastOptimize() or setOptimizedValuesToInitialValues())");
    }
    if (m_initializationScheme != 0) throw std::runtime_error("lowerBodyPoseEstimator: only initialization scheme 0 is available in this build");
    m_time = m_rthighImuPoseProblem.m_time;
    m_gb = std::make_shared<GraphBuilder>((int)K);
    GraphBuilder& gb = *m_gb;
    // () the seven IMU sub-graphs with their estimates as initial values (addImuProblemSolutionsToInitialValues / addImuProblemGraphsToGraph, :137-156)
    for (int i = 0; i < 7; i++) m_imuVars[i] = imuProblem(i)->appendToGraph(gb);
    // () instantaneous angular velocity variables + AngularVelocityFactor (:29-135), init = gyro - bias (initMode 2)
    const std::vector<unsigned> idx = m_sacrumImuPoseProblem.getImuIndecesCorrespondingToKeyframes();
    for (int i = 0; i < 7; i++) {
        const imuPoseEstimator& ipe = *imuProblem(i);
        const size_t nb = ipe.m_gyroBias.size();
        if (nb != 1 && nb != K) throw std::runtime_error("wrong size imu bias keys");
        std::vector<double> w((size_t)K * 3), gy((size_t)K * 3);
        for (unsigned k = 0; k < K; k++) {
            const Vector3 g = ipe.m_imu.gyroVec(idx[k]);
            const Vector3& b = ipe.m_gyroBias[nb == 1 ? 0 : k];
            for (int c = 0; c < 3; c++) { gy[(size_t)k * 3 + c] = g[c]; w[(size_t)k * 3 + c] = g[c] - b[c]; }
        }
        m_angVelVar[i] = gb.addTimeVar(B200_VAR_VEC3, w);
        const int coff = gb.addConstBlock(3, gy);
        const int bkind = m_imuVars[i].staticBias ? B200_AT_STATIC : B200_AT_K;
        gb.addSlot(B200_F_ANGVEL, {{B200_AT_K, m_angVelVar[i]}, {bkind, m_imuVars[i].bias}}, v3(m_angVelNoise), 0, (int)K, coff);
    }
    // () static variables at their prior means (:372-399, :228-231, :280-284)
    const Point3 pts[12] = {priorSacrumImuToRHipCtr, priorSacrumImuToLHipCtr, priorRThighImuToKneeCtr, priorRThighImuToHipCtr, priorRShankImuToKneeCtr,
                            priorRShankImuToAnkleCtr, priorRFootImuToAnkleCtr, priorLThighImuToKneeCtr, priorLThighImuToHipCtr, priorLShankImuToKneeCtr,
                            priorLShankImuToAnkleCtr, priorLFootImuToAnkleCtr};
    const Unit3 axes[11] = {priorKneeAxisRThigh, priorKneeAxisRShank, priorKneeAxisLThigh, priorKneeAxisLShank, priorHipAxisSacrum, priorRHipAxisThigh,
                            priorLHipAxisThigh, priorAnkleAxisRShank, priorAnkleAxisLShank, priorAnkleAxisRFoot, priorAnkleAxisLFoot};
    const Vector2 axisStd[11] = {rkneeHingeAxisThighPriorStd, rkneeHingeAxisShankPriorStd, lkneeHingeAxisThighPriorStd, lkneeHingeAxisShankPriorStd,
                                 hipHingeAxisSacrumPriorStd, rhipHingeAxisThighPriorStd, lhipHingeAxisThighPriorStd, rankleHingeAxisShankPriorStd,
                                 lankleHingeAxisShankPriorStd, rankleHingeAxisFootPriorStd, lankleHingeAxisFootPriorStd};
    for (int i = 0; i < 12; i++) m_staticVar[i] = gb.addStaticVar(B200_VAR_VEC3, pts[i].v);
    for (int i = 0; i < 11; i++) m_staticVar[12 + i] = gb.addStaticVar(B200_VAR_UNIT3, axes[i].p);
    auto S = [&](int s) { return VarRef{B200_AT_STATIC, m_staticVar[s]}; };
    auto pose = [&](int i) { return VarRef{B200_AT_K, m_imuVars[i].pose}; };
    auto vel = [&](int i) { return VarRef{B200_AT_K, m_imuVars[i].vel}; };
    auto angv = [&](int i) { return VarRef{B200_AT_K, m_angVelVar[i]}; };
    // () hinge constraints: knee (:238-256), hip (:206-236), ankle (:258-290)
    const int hingeType = m_hingeAxisFactorDim == 3 ? B200_F_HINGE_VEC : (m_hingeAxisFactorDim == 1 ? B200_F_HINGE_NORM : -1);
    if (hingeType < 0) throw std::runtime_error("unknown factor dimension");
    auto hingeSig = [&](const Vector3& n) { return m_hingeAxisFactorDim == 3 ? v3(n) : std::vector<double>{n.norm()}; };
    auto hinge = [&](int a, int b, int axis, const Vector3& noise) { gb.addSlot(hingeType, {pose(a), angv(a), pose(b), angv(b), S(axis)}, hingeSig(noise), 0, (int)K); };
    hinge(RT, RS, rkneeAxisThigh, m_kneeHingeAxisNoise); hinge(RS, RT, rkneeAxisShank, m_kneeHingeAxisNoise);
    hinge(LT, LS, lkneeAxisThigh, m_kneeHingeAxisNoise); hinge(LS, LT, lkneeAxisShank, m_kneeHingeAxisNoise);
    if (m_assumeHipHinge) {
        hinge(SAC, RT, hipAxisSacrum, m_hipHingeAxisNoise); hinge(SAC, LT, hipAxisSacrum, m_hipHingeAxisNoise);
        hinge(RT, SAC, rhipAxisThigh, m_hipHingeAxisNoise); hinge(LT, SAC, lhipAxisThigh, m_hipHingeAxisNoise);
    }
    if (m_assumeAnkleHinge) {
        hinge(RF, RS, rankleAxisFoot, m_ankleHingeAxisNoise); hinge(LF, LS, lankleAxisFoot, m_ankleHingeAxisNoise);
        hinge(RS, RF, rankleAxisShank, m_ankleHingeAxisNoise); hinge(LS, LF, lankleAxisShank, m_ankleHingeAxisNoise);
    }
    // () shared joint centres (:158-181)
    const int jpType = m_jointPosFactorDim == 3 ? B200_F_JOINTPOS_VEC : (m_jointPosFactorDim == 1 ? B200_F_JOINTPOS_NORM : -1);
    if (jpType < 0) throw std::runtime_error("unknown factor dimension");
    auto jpSig = [&](const Vector3& n) { return m_jointPosFactorDim == 3 ? v3(n) : std::vector<double>{n.norm(), 0.0}; };
    auto joint = [&](int a, int b, int sa, int sb, const Vector3& noise) { gb.addSlot(jpType, {pose(a), pose(b), S(sa), S(sb)}, jpSig(noise), 0, (int)K); };
    joint(SAC, RT, sacrumImuToRHipCtr, rthighImuToHipCtr, m_hipJointCtrConnectionNoise);
    joint(RT, RS, rthighImuToKneeCtr, rshankImuToKneeCtr, m_kneeJointCtrConnectionNoise);
    joint(RS, RF, rshankImuToAnkleCtr, rfootImuToAnkleCtr, m_ankleJointCtrConnectionNoise);
    joint(SAC, LT, sacrumImuToLHipCtr, lthighImuToHipCtr, m_hipJointCtrConnectionNoise);
    joint(LT, LS, lthighImuToKneeCtr, lshankImuToKneeCtr, m_kneeJointCtrConnectionNoise);
    joint(LS, LF, lshankImuToAnkleCtr, lfootImuToAnkleCtr, m_ankleJointCtrConnectionNoise);
    if (m_useJointVelFactor) {  // :183-204
        const int jvType = m_jointVelFactorDim == 3 ? B200_F_JOINTVEL_VEC : (m_jointVelFactorDim == 1 ? B200_F_JOINTVEL_NORM : -1);   // :186-204
        if (jvType < 0) throw std::runtime_error("unknown factor dimension");
        const std::vector<double> jvSig = m_jointVelFactorDim == 3 ? v3(m_jointCtrVelConnectionNoise) : std::vector<double>{m_jointCtrVelConnectionNoise.norm()};
        auto jvel = [&](int a, int b, int sa, int sb) { gb.addSlot(jvType, {pose(a), vel(a), angv(a), S(sa), pose(b), vel(b), angv(b), S(sb)}, jvSig, 0, (int)K); };
        jvel(SAC, RT, sacrumImuToRHipCtr, rthighImuToHipCtr); jvel(RT, RS, rthighImuToKneeCtr, rshankImuToKneeCtr); jvel(RS, RF, rshankImuToAnkleCtr, rfootImuToAnkleCtr);
        jvel(SAC, LT, sacrumImuToLHipCtr, lthighImuToHipCtr); jvel(LT, LS, lthighImuToKneeCtr, lshankImuToKneeCtr); jvel(LS, LF, lshankImuToAnkleCtr, lfootImuToAnkleCtr);
    }
    // () priors on the statics (:381-414, :232-236, :285-290)
    auto axisPrior = [&](int a) { const Unit3& u = axes[a - 12]; gb.addSlot(B200_F_PRIOR_UNIT3, {S(a)}, {axisStd[a - 12][0], axisStd[a - 12][1], u.p[0], u.p[1], u.p[2]}, 0, 1); };
    if (m_usePriorsOnKneeHingeAxes) for (int a : {rkneeAxisThigh, rkneeAxisShank, lkneeAxisThigh, lkneeAxisShank}) axisPrior(a);
    if (m_assumeHipHinge && m_usePriorsOnHipHingeAxes) for (int a : {hipAxisSacrum, rhipAxisThigh, lhipAxisThigh}) axisPrior(a);
    if (m_assumeAnkleHinge && m_usePriorsOnAnkleHingeAxes) for (int a : {rankleAxisShank, lankleAxisShank, rankleAxisFoot, lankleAxisFoot}) axisPrior(a);
    if (m_usePriorsOnImuToJointCtrVectors)
        for (int s = 0; s < 12; s++) gb.addSlot(B200_F_PRIOR_VEC3, {S(s)}, {imuToJointCtrPriorStd[0], imuToJointCtrPriorStd[1], imuToJointCtrPriorStd[2], pts[s][0], pts[s][1], pts[s][2]}, 0, 1);
    // () anthropometry (:430-504)
    const double A4 = 4.000000000000001;  // reference include/factors/SegmentLengthMagnitudeFactor.h:60,97
    auto seg = [&](bool use, int a, int b, double L, double sig, double mx, double mn, bool swap) {
        if (!use) return;
        gb.addSlot(B200_F_SEGLEN_MAG, {S(a), S(b)}, {sig, L}, 0, 1);
        const int a2 = swap ? b : a, b2 = swap ? a : b;
        if (m_useMaxAnthroConstraint) gb.addSlot(B200_F_SEGLEN_MAX, {S(a2), S(b2)}, {m_segmentLengthMaxNoise, mx, A4}, 0, 1);
        if (m_useMinAnthroConstraint) gb.addSlot(B200_F_SEGLEN_MIN, {S(a2), S(b2)}, {m_segmentLengthMinNoise, mn, A4}, 0, 1);
    };
    seg(m_usePelvicWidthFactor, sacrumImuToRHipCtr, sacrumImuToLHipCtr, m_pelvicWidth, m_pelvicWidthSigma, 0.409, 0.10, true);
    seg(m_useRFemurLengthFactor, rthighImuToHipCtr, rthighImuToKneeCtr, m_rfemurLength, m_rfemurLengthSigma, 0.48, 0.326, false);
    seg(m_useRTibiaLengthFactor, rshankImuToKneeCtr, rshankImuToAnkleCtr, m_rtibiaLength, m_rtibiaLengthSigma, 0.479, 0.344, false);
    seg(m_useLFemurLengthFactor, lthighImuToHipCtr, lthighImuToKneeCtr, m_lfemurLength, m_lfemurLengthSigma, 0.48, 0.326, false);
    seg(m_useLTibiaLengthFactor, lshankImuToKneeCtr, lshankImuToAnkleCtr, m_ltibiaLength, m_ltibiaLengthSigma, 0.479, 0.344, false);
    if (m_useFemurLengthDiscrepancyFactor) gb.addSlot(B200_F_SEGLEN_DISCREPANCY, {S(rthighImuToHipCtr), S(rthighImuToKneeCtr), S(lthighImuToHipCtr), S(lthighImuToKneeCtr)}, {m_femurLengthDiscrepancyNoise}, 0, 1);
    if (m_useTibiaLengthDiscrepancyFactor) gb.addSlot(B200_F_SEGLEN_DISCREPANCY, {S(rshankImuToKneeCtr), S(rshankImuToAnkleCtr), S(lshankImuToKneeCtr), S(lshankImuToAnkleCtr)}, {m_tibiaLengthDiscrepancyNoise}, 0, 1);
    if (m_useKneeAxisTibiaOrthogonalityFactor) {
        gb.addSlot(B200_F_ANGLE_AXIS_SEG, {S(rkneeAxisShank), S(rshankImuToKneeCtr), S(rshankImuToAnkleCtr)}, {m_angBwKneeAxisAndSegmentStdTibia, m_angBwKneeAxisAndSegmentMeanRTibia}, 0, 1);
        gb.addSlot(B200_F_ANGLE_AXIS_SEG, {S(lkneeAxisShank), S(lshankImuToKneeCtr), S(lshankImuToAnkleCtr)}, {m_angBwKneeAxisAndSegmentStdTibia, m_angBwKneeAxisAndSegmentMeanLTibia}, 0, 1);
    }
    if (m_useKneeAxisFemurOrthogonalityFactor) {
        gb.addSlot(B200_F_ANGLE_AXIS_SEG, {S(rkneeAxisThigh), S(rthighImuToHipCtr), S(rthighImuToKneeCtr)}, {m_angBwKneeAxisAndSegmentStdFemur, m_angBwKneeAxisAndSegmentMeanRFemur}, 0, 1);
        gb.addSlot(B200_F_ANGLE_AXIS_SEG, {S(lkneeAxisThigh), S(lthighImuToHipCtr), S(lthighImuToKneeCtr)}, {m_angBwKneeAxisAndSegmentStdFemur, m_angBwKneeAxisAndSegmentMeanLFemur}, 0, 1);
    }
    gb.finalize();
}

// pose of IMU i at keyframe k inside the GraphBuilder's value array (hipDrift.h accessor)
static double* lbpePoseAt(void* ctx, int imu, int k) {
    auto* self = static_cast<lowerBodyPoseEstimator*>(ctx);
    return const_cast<double*>(self->graph().timeValue(k, self->poseVarOfImu(imu)));
}
hipdrift::Statics lowerBodyPoseEstimator::hipDriftStatics() const {
    const GraphBuilder& gb = *m_gb;
    auto S = [&](int i) { return gb.staticValue(m_staticVar[i]); };
    // m_staticVar: 12 IMU-to-joint-centre vectors then 11 axes, in the order of setMemberEstimatedStatesFromValues()
    return hipdrift::Statics{S(0), S(1), S(3), S(2), S(4), S(5), S(6), S(8), S(7), S(9), S(10), S(11), S(12 + 0), S(12 + 2)};
}
// reference src/lowerBodyPoseEstimator.cpp:1507-1558
void lowerBodyPoseEstimator::computeHipIeRotAngleRegressionSlopesFromValues(double& rhipIntExtRotSlope, double& lhipIntExtRotSlope) {
    if (!m_gb) throw std::runtime_error("setup() must be called first");
    hipdrift::PoseAccess pa{&lbpePoseAt, this, m_gb->K()};
    hipdrift::hipIeSlopes(pa, hipDriftStatics(), m_time, rhipIntExtRotSlope, lhipIntExtRotSlope);
}
// reference src/lowerBodyPoseEstimator.cpp:1560-1652 (edits the poses of the six leg IMUs in the current values)
void lowerBodyPoseEstimator::adjustLegsForCorrectedHipIeAngles(bool verbose) {
    if (!m_gb) throw std::runtime_error("setup() must be called first");
    if (verbose) std::printf("\tcorrecting hip angles for zero hip i/e angle slope...\n");
    hipdrift::PoseAccess pa{&lbpePoseAt, this, m_gb->K()};
    hipdrift::adjustLegsForCorrectedHipIeAngles(pa, hipDriftStatics(), m_time);
}

void lowerBodyPoseEstimator::fastOptimize(double prevGlobalErr) {
    if (!m_gb) throw std::runtime_error("setup() must be called first");
    GraphBuilder& gb = *m_gb;
    b200_problem* prob = nullptr;
    if (b200_problem_create(&gb.desc(), getDevice(), &prob) != B200_OK) throw std::runtime_error(std::string("b200_problem_create: ") + b200_last_error());
    b200_status st = b200_set_values(prob, gb.timeValues().data(), gb.staticValues().data());
    b200_lm_params lm{m_lambdaInitial, m_lambdaFactor, m_lambdaLowerBound, m_lambdaUpperBound, 1.0e-20, 1.0e-20, 1.0e-3, 0, 0};  // :644-647
    // (always Levenberg-Marquardt, as upstream: m_optimizerType is only written to the results file, src/lowerBodyPoseEstimator.cpp:650, 1064)
    std::vector<double> hist(m_maxIterations + 2, 0.0);
    b200_outer_params outer{m_absErrDecreaseLimit, m_relErrDecreaseLimit, m_absErrLimit, (int)m_maxIterations, (int)m_convergeAtConsecSmallIter, 0, (int)hist.size()};
    b200_optimize_result res{};
    const auto t0 = std::chrono::steady_clock::now();
    if (st == B200_OK) st = b200_optimize(prob, &lm, &outer, hist.data(), &res);
    if (st == B200_OK || st == B200_LAMBDA_MAXED || st == B200_INDETERMINATE_SYSTEM) b200_get_values(prob, gb.timeValues().data(), gb.staticValues().data());
    b200_problem_destroy(prob);
    if (st != B200_OK && st != B200_LAMBDA_MAXED && st != B200_INDETERMINATE_SYSTEM) throw std::runtime_error(std::string("b200_optimize: ") + b200_last_error());
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    m_optimizationTotalError.insert(m_optimizationTotalError.end(), hist.begin(), hist.begin() + res.n_history);   // :655,666 (push_back semantics)
    for (int i = 0; i < res.n_history; i++) m_optimizationTotalTime.push_back(secs * i / std::max(1, res.n_history - 1));
    // restart condition on hip I/E drift (reference :668-688): a drifting hip int/ext rotation angle is straightened by
    // re-seeding the leg IMU poses, and the optimisation runs again from there
    double rhipIeSlope = 0.0, lhipIeSlope = 0.0;
    computeHipIeRotAngleRegressionSlopesFromValues(rhipIeSlope, lhipIeSlope);
    const bool doRestart = (std::fabs(rhipIeSlope) > m_hipIeAngDriftRestartThreshold || std::fabs(lhipIeSlope) > m_hipIeAngDriftRestartThreshold) &&
                           res.final_error < prevGlobalErr * 0.99 && m_nRestarts < m_maxNumRestarts;
    if (doRestart) {
        m_nRestarts++;
        adjustLegsForCorrectedHipIeAngles(false);
        fastOptimize(res.final_error);
        return;
    }
    setMemberEstimatedStatesFromValues();               // reference :700-702
    postHocAxesCheck(false);
    setDerivedJointAnglesFromEstimatedStates();
}

// reference src/lowerBodyPoseEstimator.cpp:711-797: (1) the two knee axes of a leg (thigh frame / shank frame) must point the
// same way in the navigation frame -- mean angle between them over the trial <= 90 deg, else the SHANK axis member is flipped;
// (2) the knee flexion angle should stay inside (-179, 20) deg, which the reference only reports.  Like the reference this
// changes the members (m_[rl]kneeAxisShank), not the optimised values; the derived joint angles are computed from the members.
bool lowerBodyPoseEstimator::postHocAxesCheck(bool verbose) {
    if (!m_gb) throw std::runtime_error("setup() must be called first");
    const double maxFlexExLimRad = 20.0 * M_PI / 180.0, minFlexExLimRad = -179.0 * M_PI / 180.0;
    auto meanAngleDeg = [](const std::vector<Rot3>& Ra, const Unit3& a, const std::vector<Rot3>& Rb, const Unit3& b) {
        double sum = 0.0;
        for (size_t k = 0; k < Ra.size(); k++) {
            const Vector3 u = Ra[k].rotate(a.unitVector()), w = Rb[k].rotate(b.unitVector());
            const Vector3 c(u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]);
            sum += std::atan2(c.norm(), u[0] * w[0] + u[1] * w[1] + u[2] * w[2]) * 180.0 / M_PI;   // mathutils::unsignedAngle (src/mathutils.cpp:63-91)
        }
        return Ra.empty() ? 0.0 : sum / (double)Ra.size();
    };
    bool anyChangesMade = false;
    struct Leg { const char* name; const std::vector<Rot3>* Rthigh; const std::vector<Rot3>* Rshank; const Unit3* thigh; Unit3* shank; };
    Leg legs[2] = {{"right", &m_rthighImuOrientation, &m_rshankImuOrientation, &m_rkneeAxisThigh, &m_rkneeAxisShank},
                   {"left", &m_lthighImuOrientation, &m_lshankImuOrientation, &m_lkneeAxisThigh, &m_lkneeAxisShank}};
    for (Leg& leg : legs) {
        const double mean = meanAngleDeg(*leg.Rthigh, *leg.thigh, *leg.Rshank, *leg.shank);
        if (mean > 90.0) {
            *leg.shank = Unit3(-leg.shank->p[0], -leg.shank->p[1], -leg.shank->p[2]);
            anyChangesMade = true;
        }
        if (verbose) std::printf("postHocAxesCheck: %s knee axes, mean angle between %.3f deg%s\n", leg.name, mean, mean > 90.0 ? " -> shank axis flipped" : "");
    }
    // knee flexion range (the reference tests the RIGHT knee's range for both legs, :771 -- kept)
    setDerivedJointAnglesFromEstimatedStates();
    if (!m_rkneeFlexExAng.empty()) {
        const double mx = *std::max_element(m_rkneeFlexExAng.begin(), m_rkneeFlexExAng.end()), mn = *std::min_element(m_rkneeFlexExAng.begin(), m_rkneeFlexExAng.end());
        if (!(mx < maxFlexExLimRad && mn > minFlexExLimRad)) std::printf("knee angle out of bounds!\n");
    }
    if (anyChangesMade && postHocAxesCheck(false)) std::fprintf(stderr, "Why did something changes on the second pass?\n");
    return anyChangesMade;
}

// reference src/lowerBodyPoseEstimator.cpp:1147-1228
void lowerBodyPoseEstimator::setMemberEstimatedStatesFromValues() {
    const GraphBuilder& gb = *m_gb;
    const int K = gb.K();
    std::vector<Pose3>* poses[7] = {&m_sacrumImuPose, &m_rthighImuPose, &m_rshankImuPose, &m_rfootImuPose, &m_lthighImuPose, &m_lshankImuPose, &m_lfootImuPose};
    std::vector<Rot3>* oris[7] = {&m_sacrumImuOrientation, &m_rthighImuOrientation, &m_rshankImuOrientation, &m_rfootImuOrientation, &m_lthighImuOrientation, &m_lshankImuOrientation, &m_lfootImuOrientation};
    std::vector<Point3>* poss[7] = {&m_sacrumImuPosition, &m_rthighImuPosition, &m_rshankImuPosition, &m_rfootImuPosition, &m_lthighImuPosition, &m_lshankImuPosition, &m_lfootImuPosition};
    std::vector<Vector3>* vels[7] = {&m_sacrumImuVelocity, &m_rthighImuVelocity, &m_rshankImuVelocity, &m_rfootImuVelocity, &m_lthighImuVelocity, &m_lshankImuVelocity, &m_lfootImuVelocity};
    std::vector<Vector3>* gbs[7] = {&m_sacrumImuGyroBias, &m_rthighImuGyroBias, &m_rshankImuGyroBias, &m_rfootImuGyroBias, &m_lthighImuGyroBias, &m_lshankImuGyroBias, &m_lfootImuGyroBias};
    std::vector<Vector3>* abs_[7] = {&m_sacrumImuAccelBias, &m_rthighImuAccelBias, &m_rshankImuAccelBias, &m_rfootImuAccelBias, &m_lthighImuAccelBias, &m_lshankImuAccelBias, &m_lfootImuAccelBias};
    std::vector<Vector3>* avs[7] = {&m_sacrumImuAngVel, &m_rthighImuAngVel, &m_rshankImuAngVel, &m_rfootImuAngVel, &m_lthighImuAngVel, &m_lshankImuAngVel, &m_lfootImuAngVel};
    for (int i = 0; i < 7; i++) {
        imuPoseEstimator& ipe = *imuProblem(i);
        ipe.setMemberStatesFromGraph(gb, m_imuVars[i]);
        *poses[i] = ipe.m_pose; *oris[i] = ipe.m_orientation; *poss[i] = ipe.m_position; *vels[i] = ipe.m_velocity;
        *gbs[i] = ipe.m_gyroBias; *abs_[i] = ipe.m_accelBias;
        avs[i]->assign(K, Vector3());
        for (int k = 0; k < K; k++) { const double* w = gb.timeValue(k, m_angVelVar[i]); (*avs[i])[k] = Vector3(w[0], w[1], w[2]); }
    }
    Point3* pts[12] = {&m_sacrumImuToRHipCtr, &m_sacrumImuToLHipCtr, &m_rthighImuToKneeCtr, &m_rthighImuToHipCtr, &m_rshankImuToKneeCtr, &m_rshankImuToAnkleCtr,
                       &m_rfootImuToAnkleCtr, &m_lthighImuToKneeCtr, &m_lthighImuToHipCtr, &m_lshankImuToKneeCtr, &m_lshankImuToAnkleCtr, &m_lfootImuToAnkleCtr};
    Unit3* axes[11] = {&m_rkneeAxisThigh, &m_rkneeAxisShank, &m_lkneeAxisThigh, &m_lkneeAxisShank, &m_hipAxisSacrum, &m_rhipAxisThigh, &m_lhipAxisThigh,
                       &m_rankleAxisShank, &m_lankleAxisShank, &m_rankleAxisFoot, &m_lankleAxisFoot};
    for (int s = 0; s < 12; s++) { const double* v = gb.staticValue(m_staticVar[s]); *pts[s] = Point3(v[0], v[1], v[2]); }
    for (int a = 0; a < 11; a++) { const double* v = gb.staticValue(m_staticVar[12 + a]); axes[a]->p[0] = v[0]; axes[a]->p[1] = v[1]; axes[a]->p[2] = v[2]; }
}

// reference src/lowerBodyPoseEstimator.cpp:813-835 (segment frames + clinical Grood-Suntay angles: b200_lower_body_joint_angles)
void lowerBodyPoseEstimator::setDerivedJointAnglesFromEstimatedStates(bool verbose) {
    if (!m_gb) throw std::runtime_error("setup() must be called first");
    GraphBuilder& gb = *m_gb;
    b200_joint_angle_desc d{};
    const int imus[5] = {0, 1, 2, 4, 5};                                  // sacrum, rthigh, rshank, lthigh, lshank
    for (int i = 0; i < 5; i++) d.pose_var[i] = m_imuVars[imus[i]].pose;
    // m_staticVar: 12 IMU-to-joint-centre vectors then 11 axes, in the order of setMemberEstimatedStatesFromValues()
    const int st[14] = {0, 1, 12 + 0, 3, 2, 12 + 1, 4, 5, 12 + 2, 8, 7, 12 + 3, 9, 10};
    for (int i = 0; i < 14; i++) d.static_var[i] = m_staticVar[st[i]];
    b200_problem* prob = nullptr;
    if (b200_problem_create(&gb.desc(), getDevice(), &prob) != B200_OK) throw std::runtime_error(std::string("b200_problem_create: ") + b200_last_error());
    std::vector<double> ang((size_t)gb.K() * 12);
    // the knee axes come from the MEMBERS (postHocAxesCheck may have flipped a shank axis), everything else from the values
    std::vector<double> svals = gb.staticValues();
    const Unit3* knee[4] = {&m_rkneeAxisThigh, &m_rkneeAxisShank, &m_lkneeAxisThigh, &m_lkneeAxisShank};
    if (!m_rthighImuOrientation.empty())
        for (int a = 0; a < 4; a++) {
            double* v = svals.data() + (gb.staticValue(m_staticVar[12 + a]) - gb.staticValues().data());
            v[0] = knee[a]->p[0]; v[1] = knee[a]->p[1]; v[2] = knee[a]->p[2];
        }
    b200_status stt = b200_set_values(prob, gb.timeValues().data(), svals.data());
    if (stt == B200_OK) stt = b200_lower_body_joint_angles(prob, &d, ang.data());
    b200_problem_destroy(prob);
    if (stt != B200_OK) throw std::runtime_error(std::string("b200_lower_body_joint_angles: ") + b200_last_error());
    std::vector<double>* out[12] = {&m_rkneeFlexExAng, &m_rkneeAbAdAng, &m_rkneeIntExtRotAng, &m_lkneeFlexExAng, &m_lkneeAbAdAng, &m_lkneeIntExtRotAng,
                                    &m_rhipFlexExAng, &m_rhipAbAdAng, &m_rhipIntExtRotAng, &m_lhipFlexExAng, &m_lhipAbAdAng, &m_lhipIntExtRotAng};
    for (int a = 0; a < 12; a++) {
        out[a]->resize(gb.K());
        for (int k = 0; k < gb.K(); k++) (*out[a])[k] = ang[(size_t)k * 12 + a];
    }
    if (verbose) std::printf("joint angles set for %d keyframes\n", gb.K());
}

}  // namespace bioslam_b200
