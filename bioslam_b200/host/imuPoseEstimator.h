// Host-side mirror of bioslam's imuPoseEstimator (reference include/imuPoseEstimator.h): same public settings, setup(),
// fastOptimize(), state vectors -- the graph is described as flat factor slots and solved by the CUDA library through the
// C ABI (include/bioslam_b200.h) instead of gtsam::LevenbergMarquardtOptimizer.
#pragma once
#include "GraphBuilder.h"
#include "imu.h"

namespace bioslam_b200 {

class imuPoseEstimator {
  public:
    // ++++++++++++++++ SETTINGS (names and defaults: reference include/imuPoseEstimator.h:23-53) ++++++++++++++++ //
    unsigned m_numMeasToPreint = 10;
    bool m_useMagnetometer = false;
    bool m_usePositionPrior = true, m_useVelocityPrior = true, m_useImuBiasPrior = true, m_useCompassPrior = false;
    Vector3 m_globalB = Vector3(30.5, 14.91, -56.3);
    Vector3 m_globalAcc = Vector3(0.0, 0.0, -9.81);
    Vector3 m_priorPosition = Vector3(1.0e-8, 1.0e-8, 1.0e-8), m_priorVelocity = Vector3(1.0e-8, 1.0e-8, 1.0e-8);
    Vector3 m_priorAccelBias = Vector3(1.0e-8, 1.0e-8, 1.0e-8), m_priorGyroBias = Vector3(1.0e-8, 1.0e-8, 1.0e-8);
    double m_priorCompassAngle = 0.0;
    Vector3 m_priorAccelBiasConstantNoise = Vector3(1.0e-3, 1.0e-3, 1.0e-3), m_priorGyroBiasConstantNoise = Vector3(1.0e-3, 1.0e-3, 1.0e-3);
    Vector3 m_priorPositionNoise = Vector3(1.0e-3, 1.0e-3, 1.0e-3), m_priorVelocityNoise = Vector3(1.0e-3, 1.0e-3, 1.0e-3);
    double m_priorCompassAngleNoise = 1.0e-3;
    Vector3 m_magnetometerNoise = Vector3(3.0, 3.0, 3.0);
    unsigned m_initializationScheme = 0;  // 0: zeros; 1: forward/backward orientation EKF (host/ekf.h, reference src/mathutils.cpp:489-610)
    Vector3 m_refVecLocal = Vector3(1.0, 0.0, 0.0), m_refVecNav = Vector3(1.0, 0.0, 0.0);
    double m_relErrDecreaseLimit = 1.0e-8, m_absErrDecreaseLimit = 1.0e-10, m_absErrLimit = 1.0e-12;
    unsigned m_maxIterations = 5000;
    unsigned m_optType = 1;  // 0: Gauss-Newton, 1: Levenberg-Marquardt
    // ++++++++++++++++ DATA / RESULTS ++++++++++++++++ //
    imu m_imu;
    std::string m_strId;
    std::vector<Pose3> m_pose;
    std::vector<Rot3> m_orientation;
    std::vector<Point3> m_position;
    std::vector<Vector3> m_velocity, m_accelBias, m_gyroBias;
    std::vector<double> m_time;
    unsigned m_nKeyframes = 0;
    std::vector<double> m_optimizationTotalError, m_optimizationTotalTime;
    // initial values as set up by setup() (== m_initialValues of the reference)
    std::vector<Pose3> m_initPose;
    std::vector<Vector3> m_initVel;
    std::vector<ConstantBias> m_initBias;   // size K (dynamic bias) or 1 (static bias)
    std::vector<double> m_preintRecords;    // (K-1) x 190 (CombinedImuFactor) or x 115 (ImuFactor)

    explicit imuPoseEstimator(const imu& imuIn, const std::string& id) : m_imu(imuIn), m_strId(id) {}
    imuPoseEstimator() = default;
    void setup();
    void setupImuFactorStaticBias();
    void setupCombinedImuFactor();
    void setOptimizedValuesToInitialValues();
    unsigned getImuBiasModelMode(bool printVerbose = false) const { (void)printVerbose; return imuBiasModelMode; }
    void setImuBiasModelMode(unsigned modelMode);
    void fastOptimize();
    std::vector<unsigned> getImuIndecesCorrespondingToKeyframes() const;
    unsigned getExpectedNumberOfKeyframesFromImuData() const;
    void clear();
    // variables this estimator contributed to a GraphBuilder
    struct GraphVars { int pose = -1, vel = -1, bias = -1; bool staticBias = false; };
    // append variables (initial values = current member states), IMU factors and priors to a graph (used by fastOptimize()
    // and by lowerBodyPoseEstimator::setup(), which weaves seven of these together)
    GraphVars appendToGraph(GraphBuilder& gb) const;
    void setMemberStatesFromGraph(const GraphBuilder& gb, const GraphVars& gv);
    bool isSetup() const { return m_isSetup; }

  protected:
    unsigned imuBiasModelMode = 0;  // 0: ImuFactor + static bias, 1: CombinedImuFactor + dynamic bias
    bool m_isSetup = false;
    void setupCommon(bool combined);
};

}  // namespace bioslam_b200
