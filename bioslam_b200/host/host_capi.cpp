// extern "C" shim over the host classes so that Python (ctypes) tests and tools can drive them; exceptions become status codes.
#include <cstring>
#include <map>
#include "lowerBodyPoseEstimator.h"
#include "ekf.h"

using namespace bioslam_b200;
static thread_local std::string g_err;
#define GUARD(stmt)                                      \
    try { stmt; return 0; }                              \
    catch (const std::exception& e) { g_err = e.what(); return 1; }

static bool setVec3(Vector3& dst, const double* v, int n) { if (n != 3) return false; dst = Vector3(v[0], v[1], v[2]); return true; }

extern "C" {

const char* b200h_last_error(void) { return g_err.c_str(); }
void b200h_set_device(int d) { setDevice(d); }

// acc, gyro: [n][3]; mag may be null; t: [n]
void* b200h_imu_create(int n, const double* acc, const double* gyro, const double* mag, const double* t, const char* label) {
    imu* m = new imu();
    m->label = label ? label : "";
    for (int i = 0; i < n; i++) {
        m->ax.push_back(acc[3 * i]); m->ay.push_back(acc[3 * i + 1]); m->az.push_back(acc[3 * i + 2]);
        m->gx.push_back(gyro[3 * i]); m->gy.push_back(gyro[3 * i + 1]); m->gz.push_back(gyro[3 * i + 2]);
        if (mag) { m->mx.push_back(mag[3 * i]); m->my.push_back(mag[3 * i + 1]); m->mz.push_back(mag[3 * i + 2]); }
        m->relTimeSec.push_back(t[i]);
    }
    return m;
}
void b200h_imu_destroy(void* m) { delete (imu*)m; }

void* b200h_ipe_create(void* m, const char* id) { return new imuPoseEstimator(*(imu*)m, id ? id : ""); }
void b200h_ipe_destroy(void* p) { delete (imuPoseEstimator*)p; }

// set a public setting by its member name
int b200h_ipe_set(void* pv, const char* name, const double* v, int n) {
    imuPoseEstimator& p = *(imuPoseEstimator*)pv;
    const std::string s(name);
    bool ok = true;
    if (s == "m_numMeasToPreint") p.m_numMeasToPreint = (unsigned)v[0];
    else if (s == "imuBiasModelMode") { GUARD(p.setImuBiasModelMode((unsigned)v[0])); }
    else if (s == "m_useMagnetometer") p.m_useMagnetometer = v[0] != 0;
    else if (s == "m_usePositionPrior") p.m_usePositionPrior = v[0] != 0;
    else if (s == "m_useVelocityPrior") p.m_useVelocityPrior = v[0] != 0;
    else if (s == "m_useImuBiasPrior") p.m_useImuBiasPrior = v[0] != 0;
    else if (s == "m_useCompassPrior") p.m_useCompassPrior = v[0] != 0;
    else if (s == "m_globalAcc") ok = setVec3(p.m_globalAcc, v, n);
    else if (s == "m_globalB") ok = setVec3(p.m_globalB, v, n);
    else if (s == "m_refVecLocal") ok = setVec3(p.m_refVecLocal, v, n);
    else if (s == "m_refVecNav") ok = setVec3(p.m_refVecNav, v, n);
    else if (s == "m_priorCompassAngle") p.m_priorCompassAngle = v[0];
    else if (s == "m_magnetometerNoise") ok = setVec3(p.m_magnetometerNoise, v, n);
    else if (s == "m_maxIterations") p.m_maxIterations = (unsigned)v[0];
    else if (s == "m_initializationScheme") p.m_initializationScheme = (unsigned)v[0];
    else if (s == "m_optType") p.m_optType = (unsigned)v[0];
    else if (s == "m_absErrLimit") p.m_absErrLimit = v[0];
    else if (s == "m_relErrDecreaseLimit") p.m_relErrDecreaseLimit = v[0];
    else if (s == "m_absErrDecreaseLimit") p.m_absErrDecreaseLimit = v[0];
    else if (s == "initialOrientations") {  // [K][9]: overrides the zero initialisation (stand-in for the EKF initialiser)
        if (!p.isSetup() || n != (int)p.m_nKeyframes * 9) ok = false;
        else for (unsigned k = 0; k < p.m_nKeyframes; k++) for (int c = 0; c < 9; c++) p.m_initPose[k].R.m[c] = v[(size_t)k * 9 + c];
    } else ok = false;
    if (!ok) { g_err = "unknown setting or wrong size: " + s; return 1; }
    return 0;
}
int b200h_ipe_setup(void* p) { GUARD(((imuPoseEstimator*)p)->setup()); }
int b200h_ipe_set_optimized_to_initial(void* p) { GUARD(((imuPoseEstimator*)p)->setOptimizedValuesToInitialValues()); }
int b200h_ipe_fast_optimize(void* p) { GUARD(((imuPoseEstimator*)p)->fastOptimize()); }
int b200h_ipe_n_keyframes(void* p) { return (int)((imuPoseEstimator*)p)->m_nKeyframes; }

static int copyOut(const std::vector<double>& src, double* out, int cap) {
    if ((int)src.size() > cap) { g_err = "output buffer too small"; return -1; }
    std::memcpy(out, src.data(), sizeof(double) * src.size());
    return (int)src.size();
}
static std::vector<double> flatPoses(const std::vector<Pose3>& v) {
    std::vector<double> o;
    for (const auto& p : v) { o.insert(o.end(), p.R.m, p.R.m + 9); o.insert(o.end(), p.t.v, p.t.v + 3); }
    return o;
}
static std::vector<double> flatVecs(const std::vector<Vector3>& v) { std::vector<double> o; for (const auto& x : v) o.insert(o.end(), x.v, x.v + 3); return o; }

// results by member name; returns the number of doubles written or -1
int b200h_ipe_get(void* pv, const char* name, double* out, int cap) {
    imuPoseEstimator& p = *(imuPoseEstimator*)pv;
    const std::string s(name);
    if (s == "m_pose") return copyOut(flatPoses(p.m_pose), out, cap);
    if (s == "m_velocity") return copyOut(flatVecs(p.m_velocity), out, cap);
    if (s == "m_gyroBias") return copyOut(flatVecs(p.m_gyroBias), out, cap);
    if (s == "m_accelBias") return copyOut(flatVecs(p.m_accelBias), out, cap);
    if (s == "m_time") return copyOut(p.m_time, out, cap);
    if (s == "m_optimizationTotalError") return copyOut(p.m_optimizationTotalError, out, cap);
    g_err = "unknown result: " + s;
    return -1;
}

void* b200h_lbpe_create(void* const* ipe, const char* id) {
    auto g = [&](int i) -> const imuPoseEstimator& { return *(const imuPoseEstimator*)ipe[i]; };
    return new lowerBodyPoseEstimator(g(0), g(1), g(2), g(3), g(4), g(5), g(6), id ? id : "");
}
void b200h_lbpe_destroy(void* p) { delete (lowerBodyPoseEstimator*)p; }
int b200h_lbpe_set(void* pv, const char* name, const double* v, int n) {
    lowerBodyPoseEstimator& p = *(lowerBodyPoseEstimator*)pv;
    const std::string s(name);
    (void)n;
    if (s == "m_useJointVelFactor") p.m_useJointVelFactor = v[0] != 0;
    else if (s == "m_jointVelFactorDim") p.m_jointVelFactorDim = (unsigned)v[0];
    else if (s == "m_assumeHipHinge") p.m_assumeHipHinge = v[0] != 0;
    else if (s == "m_assumeAnkleHinge") p.m_assumeAnkleHinge = v[0] != 0;
    else if (s == "m_hingeAxisFactorDim") p.m_hingeAxisFactorDim = (unsigned)v[0];
    else if (s == "m_jointPosFactorDim") p.m_jointPosFactorDim = (unsigned)v[0];
    else if (s == "m_maxIterations") p.m_maxIterations = (unsigned)v[0];
    else if (s == "m_convergeAtConsecSmallIter") p.m_convergeAtConsecSmallIter = (unsigned)v[0];
    else if (s == "m_optimizerType") p.m_optimizerType = (unsigned)v[0];
    else if (s == "m_lambdaInitial") p.m_lambdaInitial = v[0];
    else if (s == "m_useMaxAnthroConstraint") p.m_useMaxAnthroConstraint = v[0] != 0;
    else if (s == "m_useMinAnthroConstraint") p.m_useMinAnthroConstraint = v[0] != 0;
    else if (s == "m_usePriorsOnImuToJointCtrVectors") p.m_usePriorsOnImuToJointCtrVectors = v[0] != 0;
    else if (s == "m_maxNumRestarts") p.m_maxNumRestarts = (unsigned)v[0];
    else if (s == "m_hipIeAngDriftRestartThreshold") p.m_hipIeAngDriftRestartThreshold = v[0];
    else { g_err = "unknown setting: " + s; return 1; }
    return 0;
}
int b200h_lbpe_setup(void* p) { GUARD(((lowerBodyPoseEstimator*)p)->setup()); }
int b200h_lbpe_fast_optimize(void* p) { GUARD(((lowerBodyPoseEstimator*)p)->fastOptimize()); }
int b200h_lbpe_set_joint_angles(void* p) { GUARD(((lowerBodyPoseEstimator*)p)->setDerivedJointAnglesFromEstimatedStates()); }
// forward/backward orientation EKF on raw samples: q_out [n][4] = (w, x, y, z) of R[N->B]
int b200h_ekf_forward_backward(int n, const double* gyro, const double* acc, double dt, int passes, double* q_out) {
    try {
        std::vector<double> g(gyro, gyro + (size_t)n * 3), a(acc, acc + (size_t)n * 3);
        const std::vector<double> q = bioslam_b200::ekf::forwardBackwardEkf(g, a, dt, (unsigned)passes);
        std::memcpy(q_out, q.data(), sizeof(double) * q.size());
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return 1; }
}
// returns 1 if an axis was flipped, 0 if not, -1 on error
int b200h_lbpe_post_hoc_axes_check(void* p) {
    try { return ((lowerBodyPoseEstimator*)p)->postHocAxesCheck(false) ? 1 : 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
}
// out[0], out[1] = regression slopes (rad/s) of the right / left hip int/ext rotation angle of the current values
int b200h_lbpe_hip_ie_slopes(void* p, double* out) { GUARD(((lowerBodyPoseEstimator*)p)->computeHipIeRotAngleRegressionSlopesFromValues(out[0], out[1])); }
int b200h_lbpe_adjust_legs(void* p) { GUARD(((lowerBodyPoseEstimator*)p)->adjustLegsForCorrectedHipIeAngles(false)); }
int b200h_lbpe_n_restarts(void* p) { return (int)((lowerBodyPoseEstimator*)p)->m_nRestarts; }
int b200h_lbpe_set_members_from_values(void* p) { GUARD(((lowerBodyPoseEstimator*)p)->setMemberEstimatedStatesFromValues()); }
// the flat description setup() produced (valid until the estimator is destroyed) and its current values
const b200_problem_desc* b200h_lbpe_desc(void* p) { auto& e = *(lowerBodyPoseEstimator*)p; return e.isSetup() ? &e.graph().desc() : nullptr; }
int b200h_lbpe_values(void* pv, double* tv, double* sv) {
    auto& e = *(lowerBodyPoseEstimator*)pv;
    if (!e.isSetup()) { g_err = "setup() must be called first"; return 1; }
    std::memcpy(tv, e.graph().timeValues().data(), sizeof(double) * e.graph().timeValues().size());
    std::memcpy(sv, e.graph().staticValues().data(), sizeof(double) * e.graph().staticValues().size());
    return 0;
}
int b200h_lbpe_set_values(void* pv, const double* tv, const double* sv) {
    auto& e = *(lowerBodyPoseEstimator*)pv;
    if (!e.isSetup()) { g_err = "setup() must be called first"; return 1; }
    std::memcpy(e.graph().timeValues().data(), tv, sizeof(double) * e.graph().timeValues().size());
    std::memcpy(e.graph().staticValues().data(), sv, sizeof(double) * e.graph().staticValues().size());
    return 0;
}
int b200h_lbpe_get(void* pv, const char* name, double* out, int cap) {
    lowerBodyPoseEstimator& p = *(lowerBodyPoseEstimator*)pv;
    const std::string s(name);
    if (s == "m_optimizationTotalError") return copyOut(p.m_optimizationTotalError, out, cap);
    if (s == "m_time") return copyOut(p.m_time, out, cap);
    const std::map<std::string, const std::vector<double>*> angs = {{"m_rkneeFlexExAng", &p.m_rkneeFlexExAng}, {"m_rkneeAbAdAng", &p.m_rkneeAbAdAng},
        {"m_rkneeIntExtRotAng", &p.m_rkneeIntExtRotAng}, {"m_lkneeFlexExAng", &p.m_lkneeFlexExAng}, {"m_lkneeAbAdAng", &p.m_lkneeAbAdAng},
        {"m_lkneeIntExtRotAng", &p.m_lkneeIntExtRotAng}, {"m_rhipFlexExAng", &p.m_rhipFlexExAng}, {"m_rhipAbAdAng", &p.m_rhipAbAdAng},
        {"m_rhipIntExtRotAng", &p.m_rhipIntExtRotAng}, {"m_lhipFlexExAng", &p.m_lhipFlexExAng}, {"m_lhipAbAdAng", &p.m_lhipAbAdAng},
        {"m_lhipIntExtRotAng", &p.m_lhipIntExtRotAng}};
    auto ig = angs.find(s);
    if (ig != angs.end()) return copyOut(*ig->second, out, cap);
    const std::map<std::string, const std::vector<Pose3>*> poses = {{"m_sacrumImuPose", &p.m_sacrumImuPose}, {"m_rthighImuPose", &p.m_rthighImuPose},
        {"m_rshankImuPose", &p.m_rshankImuPose}, {"m_rfootImuPose", &p.m_rfootImuPose}, {"m_lthighImuPose", &p.m_lthighImuPose},
        {"m_lshankImuPose", &p.m_lshankImuPose}, {"m_lfootImuPose", &p.m_lfootImuPose}};
    auto it = poses.find(s);
    if (it != poses.end()) return copyOut(flatPoses(*it->second), out, cap);
    // every per-keyframe Vector3 member of the reference (include/lowerBodyPoseEstimator.h:118-131): <imu>Imu{AngVel, Velocity, GyroBias, AccelBias}
    const char* imuName[7] = {"sacrum", "rthigh", "rshank", "rfoot", "lthigh", "lshank", "lfoot"};
    const std::vector<Vector3>* angv[7] = {&p.m_sacrumImuAngVel, &p.m_rthighImuAngVel, &p.m_rshankImuAngVel, &p.m_rfootImuAngVel, &p.m_lthighImuAngVel, &p.m_lshankImuAngVel, &p.m_lfootImuAngVel};
    const std::vector<Vector3>* vel[7] = {&p.m_sacrumImuVelocity, &p.m_rthighImuVelocity, &p.m_rshankImuVelocity, &p.m_rfootImuVelocity, &p.m_lthighImuVelocity, &p.m_lshankImuVelocity, &p.m_lfootImuVelocity};
    const std::vector<Vector3>* gb[7] = {&p.m_sacrumImuGyroBias, &p.m_rthighImuGyroBias, &p.m_rshankImuGyroBias, &p.m_rfootImuGyroBias, &p.m_lthighImuGyroBias, &p.m_lshankImuGyroBias, &p.m_lfootImuGyroBias};
    const std::vector<Vector3>* ab[7] = {&p.m_sacrumImuAccelBias, &p.m_rthighImuAccelBias, &p.m_rshankImuAccelBias, &p.m_rfootImuAccelBias, &p.m_lthighImuAccelBias, &p.m_lshankImuAccelBias, &p.m_lfootImuAccelBias};
    for (int i = 0; i < 7; i++) {
        const std::string base = std::string("m_") + imuName[i] + "Imu";
        if (s == base + "AngVel") return copyOut(flatVecs(*angv[i]), out, cap);
        if (s == base + "Velocity") return copyOut(flatVecs(*vel[i]), out, cap);
        if (s == base + "GyroBias") return copyOut(flatVecs(*gb[i]), out, cap);
        if (s == base + "AccelBias") return copyOut(flatVecs(*ab[i]), out, cap);
    }
    const std::map<std::string, const Unit3*> ax = {{"m_rkneeAxisThigh", &p.m_rkneeAxisThigh}, {"m_rkneeAxisShank", &p.m_rkneeAxisShank},
        {"m_lkneeAxisThigh", &p.m_lkneeAxisThigh}, {"m_lkneeAxisShank", &p.m_lkneeAxisShank}, {"m_hipAxisSacrum", &p.m_hipAxisSacrum},
        {"m_rhipAxisThigh", &p.m_rhipAxisThigh}, {"m_lhipAxisThigh", &p.m_lhipAxisThigh}, {"m_rankleAxisShank", &p.m_rankleAxisShank},
        {"m_rankleAxisFoot", &p.m_rankleAxisFoot}, {"m_lankleAxisShank", &p.m_lankleAxisShank}, {"m_lankleAxisFoot", &p.m_lankleAxisFoot}};
    auto ia = ax.find(s);
    if (ia != ax.end()) return copyOut({ia->second->p[0], ia->second->p[1], ia->second->p[2]}, out, cap);
    const std::map<std::string, const Point3*> pt = {{"m_sacrumImuToRHipCtr", &p.m_sacrumImuToRHipCtr}, {"m_sacrumImuToLHipCtr", &p.m_sacrumImuToLHipCtr},
        {"m_rthighImuToHipCtr", &p.m_rthighImuToHipCtr}, {"m_rthighImuToKneeCtr", &p.m_rthighImuToKneeCtr}, {"m_rshankImuToKneeCtr", &p.m_rshankImuToKneeCtr},
        {"m_rshankImuToAnkleCtr", &p.m_rshankImuToAnkleCtr}, {"m_rfootImuToAnkleCtr", &p.m_rfootImuToAnkleCtr}, {"m_lthighImuToHipCtr", &p.m_lthighImuToHipCtr},
        {"m_lthighImuToKneeCtr", &p.m_lthighImuToKneeCtr}, {"m_lshankImuToKneeCtr", &p.m_lshankImuToKneeCtr}, {"m_lshankImuToAnkleCtr", &p.m_lshankImuToAnkleCtr},
        {"m_lfootImuToAnkleCtr", &p.m_lfootImuToAnkleCtr}};
    auto ip = pt.find(s);
    if (ip != pt.end()) return copyOut({ip->second->v[0], ip->second->v[1], ip->second->v[2]}, out, cap);
    g_err = "unknown result: " + s;
    return -1;
}

}  // extern "C"
