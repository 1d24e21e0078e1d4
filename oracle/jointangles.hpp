// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle.cpp).
// Derived joint angles: CPU restatement of
//   bioutils::get_R_Imu_to_Segment        reference src/bioutils.cpp:12-44
//   bioutils::get_R_SacrumImu_to_Pelvis   reference src/bioutils.cpp:64-97
//   bioutils::consistent3DofJcsAngles     reference src/bioutils.cpp:117-196  (Grood & Suntay joint coordinate system)
//   bioutils::clinical3DofJcsAngles       reference src/bioutils.cpp:216-237
//   mathutils::unsignedAngle/signedAngle  reference src/mathutils.cpp:63-119
//   lowerBodyPoseEstimator::setDerivedJointAnglesFromEstimatedStates   reference src/lowerBodyPoseEstimator.cpp:813-835
// Rotations are row-major 3x3 R[B->N].  Pinned by the reference's own test method (test/unit/testJointAngleCalculation.cpp):
// zero angles for identical segments, a spin of the distal IMU about the flexion / rotation axis adds to exactly one angle.
#pragma once
#include <cmath>

namespace orc {

inline void ja_cross(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
inline double ja_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline double ja_norm(const double* a) { return std::sqrt(ja_dot(a, a)); }
inline void ja_normalize(double* a) { const double n = ja_norm(a); a[0] /= n; a[1] /= n; a[2] /= n; }
inline double ja_unsigned_angle(const double* a, const double* b) {
    double c[3]; ja_cross(a, b, c);
    return std::atan2(ja_norm(c), ja_dot(a, b));
}
inline double ja_signed_angle(const double* a, const double* b, const double* ref) {
    double ang = ja_unsigned_angle(a, b);
    double c[3]; ja_cross(a, b, c);
    if (ja_unsigned_angle(ref, c) > M_PI / 2) ang = -ang;
    return ang;
}

// rows of the result are the segment axes (x right, y anterior, z proximal) expressed in the IMU frame: R[IMU->Segment]
inline void R_imu_to_segment(const double* right_axis, const double* to_prox, const double* to_dist, bool prox_is_master, double* R) {
    double Fx[3] = {right_axis[0], right_axis[1], right_axis[2]};
    double Fz[3] = {to_prox[0] - to_dist[0], to_prox[1] - to_dist[1], to_prox[2] - to_dist[2]};
    ja_normalize(Fz);
    double Fy[3]; ja_cross(Fz, Fx, Fy); ja_normalize(Fy);
    if (prox_is_master) ja_cross(Fy, Fz, Fx); else ja_cross(Fx, Fy, Fz);
    for (int c = 0; c < 3; c++) { R[c] = Fx[c]; R[3 + c] = Fy[c]; R[6 + c] = Fz[c]; }
}
inline void R_sacrum_to_pelvis(const double* up, const double* to_rhip, const double* to_lhip, bool hip_is_master, double* R) {
    double Fx[3] = {to_rhip[0] - to_lhip[0], to_rhip[1] - to_lhip[1], to_rhip[2] - to_lhip[2]};
    ja_normalize(Fx);
    double Fz[3] = {up[0], up[1], up[2]}; ja_normalize(Fz);
    double Fy[3]; ja_cross(Fz, Fx, Fy); ja_normalize(Fy);
    if (hip_is_master) ja_cross(Fx, Fy, Fz); else ja_cross(Fy, Fz, Fx);
    for (int c = 0; c < 3; c++) { R[c] = Fx[c]; R[3 + c] = Fy[c]; R[6 + c] = Fz[c]; }
}
// R[Segment->N] = R[IMU->N] * R[IMU->Segment]^T
inline void R_segment_to_N(const double* R_imu_to_N, const double* R_imu_to_seg, double* R) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0;
            for (int k = 0; k < 3; k++) s += R_imu_to_N[3 * i + k] * R_imu_to_seg[3 * j + k];
            R[3 * i + j] = s;
        }
}
// [flexion/extension, ab/adduction, internal/external rotation], "consistent" sign convention
inline void consistent_jcs_angles(const double* Rp, const double* Rd, double* ang) {
    const double i_[3] = {Rd[0], Rd[3], Rd[6]}; (void)i_;
    const double j[3] = {Rd[1], Rd[4], Rd[7]}, k[3] = {Rd[2], Rd[5], Rd[8]};
    const double I[3] = {Rp[0], Rp[3], Rp[6]}, J[3] = {Rp[1], Rp[4], Rp[7]};
    double e2[3]; ja_cross(k, I, e2); ja_normalize(e2);
    ang[2] = ja_signed_angle(e2, j, k);
    ang[0] = ja_signed_angle(J, e2, I);
    ang[1] = ja_unsigned_angle(I, k) - M_PI / 2.0;
}
inline void clinical_jcs_angles(const double* Rp, const double* Rd, bool right_leg, double* ang) {
    consistent_jcs_angles(Rp, Rd, ang);
    if (!right_leg) { ang[1] = -ang[1]; ang[2] = -ang[2]; }
}

// out[k][12] = [rknee(3), lknee(3), rhip(3), lhip(3)], each (flexEx, abAd, intExtRot), radians.
// R_imu [5][K][9]: sacrum, rthigh, rshank, lthigh, lshank.  statics (3 doubles each, in this order): sacrumImuToRHipCtr,
// sacrumImuToLHipCtr, rkneeAxisThigh, rthighImuToHipCtr, rthighImuToKneeCtr, rkneeAxisShank, rshankImuToKneeCtr,
// rshankImuToAnkleCtr, lkneeAxisThigh, lthighImuToHipCtr, lthighImuToKneeCtr, lkneeAxisShank, lshankImuToKneeCtr, lshankImuToAnkleCtr
inline void lower_body_joint_angles(int K, const double* R_imu, const double* st, double* out) {
    const double up[3] = {-1.0, 0.0, 0.0};
    double Rs[5][9];
    R_sacrum_to_pelvis(up, st + 0, st + 3, true, Rs[0]);
    R_imu_to_segment(st + 6, st + 9, st + 12, false, Rs[1]);    // right femur: the knee axis is the master vector
    R_imu_to_segment(st + 15, st + 18, st + 21, true, Rs[2]);   // right tibia
    R_imu_to_segment(st + 24, st + 27, st + 30, false, Rs[3]);  // left femur
    R_imu_to_segment(st + 33, st + 36, st + 39, true, Rs[4]);   // left tibia
    for (int k = 0; k < K; k++) {
        double Rn[5][9];
        for (int s = 0; s < 5; s++) R_segment_to_N(R_imu + ((size_t)s * K + k) * 9, Rs[s], Rn[s]);
        clinical_jcs_angles(Rn[1], Rn[2], true, out + (size_t)k * 12 + 0);
        clinical_jcs_angles(Rn[3], Rn[4], false, out + (size_t)k * 12 + 3);
        clinical_jcs_angles(Rn[0], Rn[1], true, out + (size_t)k * 12 + 6);
        clinical_jcs_angles(Rn[0], Rn[3], false, out + (size_t)k * 12 + 9);
    }
}

}  // namespace orc
