// Derived joint angles, one thread per keyframe:
//   segment frames   bioutils::get_R_SacrumImu_to_Pelvis / get_R_Imu_to_Segment   reference src/bioutils.cpp:12-97
//   R[Seg->N]        bioutils::get_R_Segment_to_N / get_R_Pelvis_to_N             reference src/bioutils.cpp:46-115
//   angles           bioutils::clinical3DofJcsAngles (Grood & Suntay)             reference src/bioutils.cpp:117-237
//   caller           lowerBodyPoseEstimator::setDerivedJointAnglesFromEstimatedStates  reference src/lowerBodyPoseEstimator.cpp:813-835
// Values are read in the solver's SoA layout ([component][Kld]): consecutive threads read consecutive addresses.
#pragma once
#include "kernels_factor.cuh"

namespace b200d {

struct JointAngleArgs {
    int pose_off[5];     // first value component of the five Pose3 variables (rotation = 9 row-major entries)
    int static_off[14];  // offsets of the 14 static 3-vectors in the static value array
};

__device__ __forceinline__ void ja_cross(const double* a, const double* b, double* c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double ja_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void ja_normalize(double* a) { const double n = sqrt(ja_dot(a, a)); a[0] /= n; a[1] /= n; a[2] /= n; }
__device__ __forceinline__ double ja_unsigned_angle(const double* a, const double* b) {
    double c[3]; ja_cross(a, b, c);
    return atan2(sqrt(ja_dot(c, c)), ja_dot(a, b));
}
__device__ __forceinline__ double ja_signed_angle(const double* a, const double* b, const double* ref) {
    double ang = ja_unsigned_angle(a, b);
    double c[3]; ja_cross(a, b, c);
    if (ja_unsigned_angle(ref, c) > 1.5707963267948966) ang = -ang;
    return ang;
}
// rows = segment axes (x right, y anterior, z proximal) in the IMU frame
__device__ __forceinline__ void ja_frame(double* Fx, double* Fz, bool z_is_master, double* R) {
    double Fy[3]; ja_cross(Fz, Fx, Fy); ja_normalize(Fy);
    if (z_is_master) ja_cross(Fy, Fz, Fx); else ja_cross(Fx, Fy, Fz);
#pragma unroll
    for (int c = 0; c < 3; c++) { R[c] = Fx[c]; R[3 + c] = Fy[c]; R[6 + c] = Fz[c]; }
}
__device__ __forceinline__ void ja_clinical(const double* Rp, const double* Rd, bool right_leg, double* ang) {
    const double j[3] = {Rd[1], Rd[4], Rd[7]}, k[3] = {Rd[2], Rd[5], Rd[8]};
    const double I[3] = {Rp[0], Rp[3], Rp[6]}, J[3] = {Rp[1], Rp[4], Rp[7]};
    double e2[3]; ja_cross(k, I, e2); ja_normalize(e2);
    const double rot = ja_signed_angle(e2, j, k), flex = ja_signed_angle(J, e2, I), abd = ja_unsigned_angle(I, k) - 1.5707963267948966;
    ang[0] = flex; ang[1] = right_leg ? abd : -abd; ang[2] = right_leg ? rot : -rot;
}

__global__ void __launch_bounds__(128) joint_angles_kernel(JointAngleArgs ja, int K, long long Kld, const double* __restrict__ tv,
                                                           const double* __restrict__ sv, double* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double st[14][3];
#pragma unroll
    for (int v = 0; v < 14; v++)
#pragma unroll
        for (int c = 0; c < 3; c++) st[v][c] = sv[ja.static_off[v] + c];
    // constant IMU -> segment rotations (the statics are shared by all keyframes: recomputed per thread, a few hundred flops)
    double Rs[5][9];
    {
        double Fx[3] = {st[0][0] - st[1][0], st[0][1] - st[1][1], st[0][2] - st[1][2]};
        ja_normalize(Fx);
        double Fz[3] = {-1.0, 0.0, 0.0};
        ja_frame(Fx, Fz, false, Rs[0]);                       // pelvis: the hip-centre line is the master vector
    }
    const int seg[4][3] = {{2, 3, 4}, {5, 6, 7}, {8, 9, 10}, {11, 12, 13}};   // (axis, to proximal, to distal) of femur R, tibia R, femur L, tibia L
#pragma unroll
    for (int s = 0; s < 4; s++) {
        double Fx[3] = {st[seg[s][0]][0], st[seg[s][0]][1], st[seg[s][0]][2]};
        double Fz[3] = {st[seg[s][1]][0] - st[seg[s][2]][0], st[seg[s][1]][1] - st[seg[s][2]][1], st[seg[s][1]][2] - st[seg[s][2]][2]};
        ja_normalize(Fz);
        ja_frame(Fx, Fz, (s & 1) != 0, Rs[1 + s]);            // femur: knee axis is master; tibia: proximal vector is master
    }
    double Rn[5][9];
#pragma unroll
    for (int s = 0; s < 5; s++) {
        double Ri[9];
#pragma unroll
        for (int c = 0; c < 9; c++) Ri[c] = tv[(long long)(ja.pose_off[s] + c) * Kld + k];
#pragma unroll
        for (int i = 0; i < 3; i++)
#pragma unroll
            for (int j = 0; j < 3; j++) Rn[s][3 * i + j] = Ri[3 * i] * Rs[s][3 * j] + Ri[3 * i + 1] * Rs[s][3 * j + 1] + Ri[3 * i + 2] * Rs[s][3 * j + 2];
    }
    double* o = out + (long long)k * 12;
    ja_clinical(Rn[1], Rn[2], true, o + 0);
    ja_clinical(Rn[3], Rn[4], false, o + 3);
    ja_clinical(Rn[0], Rn[1], true, o + 6);
    ja_clinical(Rn[0], Rn[3], false, o + 9);
}

}  // namespace b200d
