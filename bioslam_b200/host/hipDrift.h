// Hip internal/external-rotation drift heuristic of lowerBodyPoseEstimator::fastOptimize, host side.
// Restates, on the flat value arrays of the GraphBuilder:
//   lowerBodyPoseEstimator::computeHipIeRotAngleRegressionSlopesFromValues   reference src/lowerBodyPoseEstimator.cpp:1507-1558
//   lowerBodyPoseEstimator::adjustLegsForCorrectedHipIeAngles                reference src/lowerBodyPoseEstimator.cpp:1560-1652
//   lowerBodyPoseEstimator::correctImuPosesForConnectedJointCenters          reference src/lowerBodyPoseEstimator.cpp:1654-1661
//   mathutils::angleUnwrapOnKnownDomain / linearRegression / adjustRegressionSlopeToTargetSlope   src/mathutils.cpp:403-457
//   mathutils::vectorFromPointToAxis / rotateImuPoseAboutPointAxisAngle                          src/mathutils.cpp:815-858
//   bioutils segment frames and Grood-Suntay hip angles (femur frame with the PROXIMAL vector as master, as the reference
//   uses here -- the derived joint angles of setDerivedJointAnglesFromEstimatedStates use the knee axis)   src/bioutils.cpp:12-237
// O(K) scalar arithmetic on 6 poses per keyframe, run once per LM run: it stays on the host (D2H of the values is part of
// fastOptimize anyway).  Poses are [R row-major 9 | t 3], R = R[B->N].
#pragma once
#include <vector>

namespace bioslam_b200 {
namespace hipdrift {

struct Statics {   // 3 doubles each
    const double *sacrumImuToRHipCtr, *sacrumImuToLHipCtr, *rthighImuToHipCtr, *rthighImuToKneeCtr, *rshankImuToKneeCtr, *rshankImuToAnkleCtr,
        *rfootImuToAnkleCtr, *lthighImuToHipCtr, *lthighImuToKneeCtr, *lshankImuToKneeCtr, *lshankImuToAnkleCtr, *lfootImuToAnkleCtr,
        *rkneeAxisThigh, *lkneeAxisThigh;
};
// poses[i][k] -> 12 doubles of IMU i (0 sacrum, 1 rthigh, 2 rshank, 3 rfoot, 4 lthigh, 5 lshank, 6 lfoot) at keyframe k
struct PoseAccess {
    double* (*at)(void* ctx, int imu, int k);
    void* ctx;
    int K;
};

// unwrapped clinical hip int/ext rotation angles (rad) of both legs at every keyframe
void hipIeAngles(const PoseAccess& pa, const Statics& st, std::vector<double>& rhipIe, std::vector<double>& lhipIe);
// slope of the least-squares line through (x, y)
void linearRegression(const std::vector<double>& x, const std::vector<double>& y, double& m, double& b);
void hipIeSlopes(const PoseAccess& pa, const Statics& st, const std::vector<double>& time, double& rslope, double& lslope);
// straightens both hip I/E angle series (regression slope 0, first sample 0) by rotating every leg IMU about its femur's long
// axis through the hip centre, then reconnects the joint centres from the sacrum down; poses are edited in place
void adjustLegsForCorrectedHipIeAngles(const PoseAccess& pa, const Statics& st, const std::vector<double>& time);

}  // namespace hipdrift
}  // namespace bioslam_b200
