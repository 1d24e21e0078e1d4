// Host-side mirror of bioslam's lowerBodyPoseEstimator (reference include/lowerBodyPoseEstimator.h): same public model /
// solver settings, setup(), fastOptimize(), estimated-state members.  setup() weaves the seven IMU sub-graphs together
// exactly as reference src/lowerBodyPoseEstimator.cpp:361-505 does, as flat factor slots; fastOptimize() is the loop of
// :639-667 run by the CUDA library.
#pragma once
#include <cmath>
#include "hipDrift.h"
#include <memory>
#include "imuPoseEstimator.h"

namespace bioslam_b200 {

class lowerBodyPoseEstimator {
  public:
    // ----------------- biomechanical model options (reference include/lowerBodyPoseEstimator.h:24-37) ----------------- //
    bool m_useRFemurLengthFactor = true, m_useRTibiaLengthFactor = true, m_useLFemurLengthFactor = true, m_useLTibiaLengthFactor = true, m_usePelvicWidthFactor = true;
    bool m_useMaxAnthroConstraint = true, m_useMinAnthroConstraint = true;
    bool m_useJointVelFactor = false;
    bool m_useFemurLengthDiscrepancyFactor = true, m_useTibiaLengthDiscrepancyFactor = true;
    bool m_useKneeAxisTibiaOrthogonalityFactor = true, m_useKneeAxisFemurOrthogonalityFactor = true;
    bool m_assumeHipHinge = true, m_assumeAnkleHinge = true;
    unsigned m_initializationScheme = 0;
    double m_rfemurLength = 0.3943, m_rtibiaLength = 0.4110, m_lfemurLength = 0.3943, m_ltibiaLength = 0.4110, m_pelvicWidth = 0.1866;
    double m_rfemurLengthSigma = 0.0302, m_rtibiaLengthSigma = 0.0264, m_lfemurLengthSigma = 0.0302, m_ltibiaLengthSigma = 0.0264, m_pelvicWidthSigma = 0.0098;
    double m_angBwKneeAxisAndSegmentMeanRFemur = 84.0 * M_PI / 180.0, m_angBwKneeAxisAndSegmentStdFemur = 2.4 * M_PI / 180.0;
    double m_angBwKneeAxisAndSegmentMeanLFemur = 96.0 * M_PI / 180.0;
    double m_angBwKneeAxisAndSegmentMeanRTibia = 92.0 * M_PI / 180.0, m_angBwKneeAxisAndSegmentStdTibia = 1.2 * M_PI / 180.0;
    double m_angBwKneeAxisAndSegmentMeanLTibia = 88.0 * M_PI / 180.0;
    // ------------------ other model noise options (:45-56) ------------------- //
    unsigned m_jointPosFactorDim = 3, m_hingeAxisFactorDim = 3, m_jointVelFactorDim = 3;
    Vector3 m_hipJointCtrConnectionNoise = Vector3(9.135, 16.517, 29.378) * 1.0e-2 * 1.2;
    Vector3 m_kneeJointCtrConnectionNoise = Vector3(6.7348, 12.3006, 10.1090) * 1.0e-2 * 1.2;
    Vector3 m_ankleJointCtrConnectionNoise = Vector3(7.3741, 6.2212, 4.6993) * 1.0e-2 * 1.2;
    Vector3 m_hipHingeAxisNoise = Vector3(1.0, 1.0, 1.0) * 20.0, m_kneeHingeAxisNoise = Vector3(1.0, 1.0, 1.0) * 10.0, m_ankleHingeAxisNoise = Vector3(1.0, 1.0, 1.0) * 20.0;
    double m_segmentLengthMaxNoise = 1.0e-3, m_segmentLengthMinNoise = 1.0e-3;
    Vector3 m_angVelNoise = Vector3(0.000157079633, 0.000157079633, 0.000157079633) * 0.5;
    Vector3 m_jointCtrVelConnectionNoise = Vector3(1.0, 1.0, 1.0) * 10.0;
    double m_femurLengthDiscrepancyNoise = 0.002, m_tibiaLengthDiscrepancyNoise = 0.002;
    // -------------------- model prior settings (:58-79) --------------------- //
    bool m_usePriorsOnKneeHingeAxes = true, m_usePriorsOnHipHingeAxes = true, m_usePriorsOnAnkleHingeAxes = true, m_usePriorsOnImuToJointCtrVectors = true;
    Unit3 priorKneeAxisRThigh = Unit3(0.05, 0.05, 0.95), priorKneeAxisRShank = Unit3(-0.05, 0.05, 0.95), priorKneeAxisLThigh = Unit3(0.05, 0.05, -0.95), priorKneeAxisLShank = Unit3(-0.05, 0.05, -0.95);
    Vector2 rkneeHingeAxisThighPriorStd = Vector2(0.15, 0.15), rkneeHingeAxisShankPriorStd = Vector2(0.15, 0.15), lkneeHingeAxisThighPriorStd = Vector2(0.15, 0.15), lkneeHingeAxisShankPriorStd = Vector2(0.15, 0.15);
    Unit3 priorHipAxisSacrum = Unit3(0.05, 0.95, 0.05), priorRHipAxisThigh = Unit3(-0.05, 0.05, 0.95), priorLHipAxisThigh = Unit3(-0.05, 0.05, -0.95);
    Vector2 hipHingeAxisSacrumPriorStd = Vector2(0.5, 0.5), rhipHingeAxisThighPriorStd = Vector2(0.5, 0.5), lhipHingeAxisThighPriorStd = Vector2(0.5, 0.5);
    Unit3 priorAnkleAxisRFoot = Unit3(0.05, -0.95, -0.05), priorAnkleAxisLFoot = Unit3(-0.05, -0.95, 0.05), priorAnkleAxisRShank = Unit3(0.05, 0.05, 0.95), priorAnkleAxisLShank = Unit3(0.05, 0.05, -0.95);
    Vector2 rankleHingeAxisFootPriorStd = Vector2(0.15, 0.15), lankleHingeAxisFootPriorStd = Vector2(0.15, 0.15), rankleHingeAxisShankPriorStd = Vector2(0.15, 0.15), lankleHingeAxisShankPriorStd = Vector2(0.15, 0.15);
    Point3 priorSacrumImuToRHipCtr = Point3(.026, .135, -.215), priorSacrumImuToLHipCtr = Point3(.024, -.132, -.213);
    Point3 priorRThighImuToKneeCtr = Point3(.26, -1.0e-5, -1.0e-5), priorRThighImuToHipCtr = Point3(-.24, 1.0e-5, 1.0e-5);
    Point3 priorRShankImuToKneeCtr = Point3(-.11, 1.0e-5, 1.0e-5), priorRShankImuToAnkleCtr = Point3(.31, .11, 1.0e-5);
    Point3 priorRFootImuToAnkleCtr = Point3(-.051, -1.0e-5, -.05);
    Point3 priorLThighImuToKneeCtr = Point3(.26, -1.0e-5, -1.0e-5), priorLThighImuToHipCtr = Point3(-.24, 1.0e-5, 1.0e-5);
    Point3 priorLShankImuToKneeCtr = Point3(-.1, 1.0e-5, 1.0e-5), priorLShankImuToAnkleCtr = Point3(.29, .09, -1.0e-5);
    Point3 priorLFootImuToAnkleCtr = Point3(-.049, 1.0e-5, -.055);
    Vector3 imuToJointCtrPriorStd = Vector3(0.5, 0.5, 0.5);
    // ---------------------- solver settings (:81-89) ------------------------ //
    unsigned m_optimizerType = 1;  // 0: Gauss-Newton, 1: Levenberg-Marquardt
    double m_absErrDecreaseLimit = 1.0e-6, m_relErrDecreaseLimit = 1.0e-5, m_absErrLimit = 1.0e-5;
    unsigned m_maxIterations = 10000;
    double m_lambdaUpperBound = 1.0e25, m_lambdaLowerBound = 1.0e-10, m_lambdaInitial = 1.0e-5, m_lambdaFactor = 10.0;
    unsigned m_convergeAtConsecSmallIter = 6;
    std::string m_linearSolverType = "B200_NESTED_BLOCK_CHOLESKY";  // (the reference's GTSAM solver names do not apply)
    // --------------------------------------------------------------- //
    std::string m_strId;
    imuPoseEstimator m_rthighImuPoseProblem, m_rshankImuPoseProblem, m_rfootImuPoseProblem, m_lthighImuPoseProblem, m_lshankImuPoseProblem, m_lfootImuPoseProblem, m_sacrumImuPoseProblem;
    // output (estimated) states (:118-131)
    std::vector<Pose3> m_sacrumImuPose, m_rthighImuPose, m_rshankImuPose, m_rfootImuPose, m_lthighImuPose, m_lshankImuPose, m_lfootImuPose;
    std::vector<Rot3> m_sacrumImuOrientation, m_rthighImuOrientation, m_rshankImuOrientation, m_rfootImuOrientation, m_lthighImuOrientation, m_lshankImuOrientation, m_lfootImuOrientation;
    std::vector<Point3> m_sacrumImuPosition, m_rthighImuPosition, m_rshankImuPosition, m_rfootImuPosition, m_lthighImuPosition, m_lshankImuPosition, m_lfootImuPosition;
    std::vector<Vector3> m_sacrumImuVelocity, m_rthighImuVelocity, m_rshankImuVelocity, m_rfootImuVelocity, m_lthighImuVelocity, m_lshankImuVelocity, m_lfootImuVelocity;
    std::vector<Vector3> m_sacrumImuGyroBias, m_rthighImuGyroBias, m_rshankImuGyroBias, m_rfootImuGyroBias, m_lthighImuGyroBias, m_lshankImuGyroBias, m_lfootImuGyroBias;
    std::vector<Vector3> m_sacrumImuAccelBias, m_rthighImuAccelBias, m_rshankImuAccelBias, m_rfootImuAccelBias, m_lthighImuAccelBias, m_lshankImuAccelBias, m_lfootImuAccelBias;
    Unit3 m_rkneeAxisThigh, m_rkneeAxisShank, m_lkneeAxisThigh, m_lkneeAxisShank;
    Unit3 m_hipAxisSacrum, m_rhipAxisThigh, m_lhipAxisThigh;
    Unit3 m_rankleAxisShank, m_rankleAxisFoot, m_lankleAxisShank, m_lankleAxisFoot;
    Point3 m_sacrumImuToRHipCtr, m_rthighImuToHipCtr, m_rthighImuToKneeCtr, m_rshankImuToKneeCtr, m_rshankImuToAnkleCtr, m_rfootImuToAnkleCtr;
    Point3 m_sacrumImuToLHipCtr, m_lthighImuToHipCtr, m_lthighImuToKneeCtr, m_lshankImuToKneeCtr, m_lshankImuToAnkleCtr, m_lfootImuToAnkleCtr;
    std::vector<double> m_time;
    std::vector<Vector3> m_sacrumImuAngVel, m_rthighImuAngVel, m_rshankImuAngVel, m_rfootImuAngVel, m_lthighImuAngVel, m_lshankImuAngVel, m_lfootImuAngVel;
    std::vector<double> m_optimizationTotalError, m_optimizationTotalTime;
    // derived joint angles, radians (reference include/lowerBodyPoseEstimator.h:133-137)
    std::vector<double> m_rkneeFlexExAng, m_rkneeIntExtRotAng, m_rkneeAbAdAng, m_lkneeFlexExAng, m_lkneeIntExtRotAng, m_lkneeAbAdAng;
    std::vector<double> m_rhipFlexExAng, m_rhipIntExtRotAng, m_rhipAbAdAng, m_lhipFlexExAng, m_lhipIntExtRotAng, m_lhipAbAdAng;

    explicit lowerBodyPoseEstimator(const imuPoseEstimator& sacrumImuPoseProblem, const imuPoseEstimator& rthighImuPoseProblem,
                                    const imuPoseEstimator& rshankImuPoseProblem, const imuPoseEstimator& rfootImuPoseProblem,
                                    const imuPoseEstimator& lthighImuPoseProblem, const imuPoseEstimator& lshankImuPoseProblem,
                                    const imuPoseEstimator& lfootImuPoseProblem, const std::string& id);
    lowerBodyPoseEstimator() = default;
    void setup();
    void fastOptimize(double prevGlobalErr = 9e9);
    void setMemberEstimatedStatesFromValues();
    // == reference src/lowerBodyPoseEstimator.cpp:813-835; runs b200_lower_body_joint_angles on the device
    void setDerivedJointAnglesFromEstimatedStates(bool verbose = false);
    bool postHocAxesCheck(bool verbose = false);   // reference :711-797: flips a knee shank axis that points against its thigh axis
    // hip int/ext rotation drift restart (reference :668-688, :1507-1661; hipDrift.h)
    double m_hipIeAngDriftRestartThreshold = 100.0 * M_PI / 180.0 / 30.0;   // [rad/s] reference include/lowerBodyPoseEstimator.h:87
    unsigned m_maxNumRestarts = 100, m_nRestarts = 0;                        // :88
    void computeHipIeRotAngleRegressionSlopesFromValues(double& rhipIntExtRotSlope, double& lhipIntExtRotSlope);
    void adjustLegsForCorrectedHipIeAngles(bool verbose = false);
    int poseVarOfImu(int imu) const { return m_imuVars[imu].pose; }
    hipdrift::Statics hipDriftStatics() const;
    // the flat description setup() produced and its current values (== m_graph / m_initialValues of the reference)
    const GraphBuilder& graph() const { return *m_gb; }
    GraphBuilder& graph() { return *m_gb; }
    bool isSetup() const { return (bool)m_gb; }

  private:
    std::shared_ptr<GraphBuilder> m_gb;
    imuPoseEstimator* imuProblem(int i);
    imuPoseEstimator::GraphVars m_imuVars[7];
    int m_angVelVar[7] = {-1, -1, -1, -1, -1, -1, -1};
    int m_staticVar[23];
};

}  // namespace bioslam_b200
