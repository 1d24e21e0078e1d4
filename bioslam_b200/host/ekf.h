// Forward/backward orientation EKF that seeds imuPoseEstimator (m_initializationScheme == 1).
// Restates mathutils::simpleImuOrientationForwardEkf / simpleImuOrientationForwardBackwardEkfWrapper
// (reference src/mathutils.cpp:489-610; call site src/imuPoseEstimator.cpp:92-110).  Serial in time (O(N), 4-state filter),
// runs once in setup(): host code.
#pragma once
#include <vector>

namespace bioslam_b200 {
namespace ekf {

// gyros, accels: [N][3] row-major.  Returns the quaternions (w, x, y, z) of R[N->B] at every sample, [N][4].
// final_P (4x4 row-major, in/out): covariance at the last sample.
std::vector<double> forwardEkf(const std::vector<double>& gyros, const std::vector<double>& accels, double dt, const double initX[4], double P[16]);
std::vector<double> forwardBackwardEkf(const std::vector<double>& gyros, const std::vector<double>& accels, double dt, unsigned numFBPasses);
// gtsam::Rot3(w, x, y, z).inverse() as a row-major matrix: what upstream stores as the keyframe's R[B->N] (it labels the
// filter's quaternion R[N->B]; see tests/test_host_layer.py::test_host_forward_backward_ekf for the convention actually measured)
void quatToRotBodyToNav(const double q[4], double R[9]);

}  // namespace ekf
}  // namespace bioslam_b200
