#include "hipDrift.h"
#include <cmath>
#include <stdexcept>

namespace bioslam_b200 {
namespace hipdrift {
namespace {

struct V3 { double x, y, z; };
inline V3 v3(const double* p) { return {p[0], p[1], p[2]}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 unit(V3 a) { const double n = norm(a); return {a.x / n, a.y / n, a.z / n}; }
struct M3 { double m[9]; };   // row-major
inline V3 mul(const M3& R, V3 p) { return {R.m[0] * p.x + R.m[1] * p.y + R.m[2] * p.z, R.m[3] * p.x + R.m[4] * p.y + R.m[5] * p.z, R.m[6] * p.x + R.m[7] * p.y + R.m[8] * p.z}; }
inline V3 mulT(const M3& R, V3 p) { return {R.m[0] * p.x + R.m[3] * p.y + R.m[6] * p.z, R.m[1] * p.x + R.m[4] * p.y + R.m[7] * p.z, R.m[2] * p.x + R.m[5] * p.y + R.m[8] * p.z}; }
inline M3 matmul(const M3& A, const M3& B, bool transposeB) {
    M3 C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            double s = 0.0;
            for (int k = 0; k < 3; k++) s += A.m[3 * i + k] * (transposeB ? B.m[3 * j + k] : B.m[3 * k + j]);
            C.m[3 * i + j] = s;
        }
    return C;
}
inline M3 rows(V3 a, V3 b, V3 c) { return {{a.x, a.y, a.z, b.x, b.y, b.z, c.x, c.y, c.z}}; }
// gtsam::Rot3::AxisAngle(k, theta): Rodrigues
inline M3 axisAngle(V3 k, double th) {
    const double c = std::cos(th), s = std::sin(th), v = 1.0 - c;
    return {{c + k.x * k.x * v, k.x * k.y * v - k.z * s, k.x * k.z * v + k.y * s,
             k.y * k.x * v + k.z * s, c + k.y * k.y * v, k.y * k.z * v - k.x * s,
             k.z * k.x * v - k.y * s, k.z * k.y * v + k.x * s, c + k.z * k.z * v}};
}
inline double unsignedAngle(V3 a, V3 b) { return std::atan2(norm(cross(a, b)), dot(a, b)); }            // src/mathutils.cpp:63-91
inline double signedAngle(V3 a, V3 b, V3 ref) {                                                         // src/mathutils.cpp:93-119
    double ang = unsignedAngle(a, b);
    if (unsignedAngle(ref, cross(a, b)) > M_PI / 2) ang = -ang;
    return ang;
}
// R[IMU->Segment] rows = segment axes (x right, y anterior, z proximal) in the IMU frame   src/bioutils.cpp:12-40
inline M3 R_imu_to_segment(V3 rightAxis, V3 toProx, V3 toDist, bool proxIsMaster) {
    V3 Fx = rightAxis, Fz = unit(toProx - toDist);
    V3 Fy = unit(cross(Fz, Fx));
    if (proxIsMaster) Fx = cross(Fy, Fz); else Fz = cross(Fx, Fy);
    return rows(Fx, Fy, Fz);
}
inline M3 R_sacrum_to_pelvis(V3 up, V3 toRHip, V3 toLHip) {                                            // src/bioutils.cpp:61-89, hip line is master
    V3 Fx = unit(toRHip - toLHip), Fz = unit(up);
    V3 Fy = unit(cross(Fz, Fx));
    Fz = cross(Fx, Fy);
    return rows(Fx, Fy, Fz);
}
// third Grood-Suntay angle (int/ext rotation), clinical sign convention                                src/bioutils.cpp:117-237
inline double clinicalIntExtRot(const M3& Rp, const M3& Rd, bool rightLeg) {
    const V3 j{Rd.m[1], Rd.m[4], Rd.m[7]}, k{Rd.m[2], Rd.m[5], Rd.m[8]}, I{Rp.m[0], Rp.m[3], Rp.m[6]};
    const V3 e2 = unit(cross(k, I));
    const double a = signedAngle(e2, j, k);
    return rightLeg ? a : -a;
}
inline void angleUnwrapOnKnownDomain(std::vector<double>& out, double domainMin, double domainMax, double tol) {   // src/mathutils.cpp:403-424
    const double width = domainMax - domainMin;
    for (size_t i = 1; i < out.size(); i++) {
        const double d = out[i] - out[i - 1];
        if (std::fabs(d) > tol) {
            const double m = std::ceil(std::fabs(d) / width), s = d > 0 ? -1.0 : 1.0;
            for (size_t j = i; j < out.size(); j++) out[j] += m * s * width;
        }
    }
}
struct Pose { M3 R; V3 t; };
inline Pose load(const double* p) { Pose x; for (int i = 0; i < 9; i++) x.R.m[i] = p[i]; x.t = v3(p + 9); return x; }
inline void store(double* p, const Pose& x) { for (int i = 0; i < 9; i++) p[i] = x.R.m[i]; p[9] = x.t.x; p[10] = x.t.y; p[11] = x.t.z; }
inline V3 transformFrom(const Pose& x, V3 p) { return x.t + mul(x.R, p); }
// src/mathutils.cpp:832-858 (the orientation is composed with R[B0->B1]^T exactly as the reference does)
inline void rotateImuPoseAboutPointAxisAngle(Pose& x0, V3 rotPointNav, V3 rotVecNav, double rotAngle) {
    const V3 kN = unit(rotVecNav);
    const V3 kB = unit(mulT(x0.R, rotVecNav));
    const M3 R_B0_to_B1 = axisAngle(kB, rotAngle);
    const M3 R_B1_to_N = matmul(x0.R, R_B0_to_B1, true);
    const V3 d = x0.t - rotPointNav;
    const V3 a0 = d - dot(d, kN) * kN;        // == -vectorFromPointToAxis(p0, a, k): from the rotation line to p0
    const V3 c0 = x0.t - a0;
    const V3 a1 = mul(axisAngle(kN, rotAngle), a0);
    x0.R = R_B1_to_N; x0.t = c0 + a1;
}
inline void connect(const Pose& sacrum, Pose& thigh, Pose& shank, Pose& foot, V3 sacrumToHip, V3 thighToHip, V3 thighToKnee, V3 shankToKnee,
                    V3 shankToAnkle, V3 footToAnkle) {   // src/lowerBodyPoseEstimator.cpp:1654-1661
    thigh.t = thigh.t + (transformFrom(sacrum, sacrumToHip) - transformFrom(thigh, thighToHip));
    shank.t = shank.t + (transformFrom(thigh, thighToKnee) - transformFrom(shank, shankToKnee));
    foot.t = foot.t + (transformFrom(shank, shankToAnkle) - transformFrom(foot, footToAnkle));
}

}  // namespace

void linearRegression(const std::vector<double>& x, const std::vector<double>& y, double& m, double& b) {   // src/mathutils.cpp:433-440
    if (x.size() != y.size() || x.empty()) throw std::runtime_error("linearRegression: size mismatch");
    double mx = 0.0, my = 0.0;
    for (size_t i = 0; i < x.size(); i++) { mx += x[i]; my += y[i]; }
    mx /= (double)x.size(); my /= (double)y.size();
    double sxy = 0.0, sxx = 0.0;
    for (size_t i = 0; i < x.size(); i++) { sxy += (x[i] - mx) * (y[i] - my); sxx += (x[i] - mx) * (x[i] - mx); }
    m = sxy / sxx; b = my - m * mx;
}

void hipIeAngles(const PoseAccess& pa, const Statics& st, std::vector<double>& rhipIe, std::vector<double>& lhipIe) {
    const M3 Rsp = R_sacrum_to_pelvis({-1.0, 0.0, 0.0}, v3(st.sacrumImuToRHipCtr), v3(st.sacrumImuToLHipCtr));
    const M3 Rrf = R_imu_to_segment(v3(st.rkneeAxisThigh), v3(st.rthighImuToHipCtr), v3(st.rthighImuToKneeCtr), true);
    const M3 Rlf = R_imu_to_segment(v3(st.lkneeAxisThigh), v3(st.lthighImuToHipCtr), v3(st.lthighImuToKneeCtr), true);
    rhipIe.assign(pa.K, 0.0); lhipIe.assign(pa.K, 0.0);
    for (int k = 0; k < pa.K; k++) {
        const M3 Rp = matmul(load(pa.at(pa.ctx, 0, k)).R, Rsp, true);      // R[Segment->N] = R[IMU->N] R[IMU->Segment]^T
        const M3 Rr = matmul(load(pa.at(pa.ctx, 1, k)).R, Rrf, true);
        const M3 Rl = matmul(load(pa.at(pa.ctx, 4, k)).R, Rlf, true);
        rhipIe[k] = clinicalIntExtRot(Rp, Rr, true);
        lhipIe[k] = clinicalIntExtRot(Rp, Rl, false);
    }
    angleUnwrapOnKnownDomain(rhipIe, -M_PI, M_PI, 0.95 * 2.0 * M_PI);
    angleUnwrapOnKnownDomain(lhipIe, -M_PI, M_PI, 0.95 * 2.0 * M_PI);
}

void hipIeSlopes(const PoseAccess& pa, const Statics& st, const std::vector<double>& time, double& rslope, double& lslope) {
    std::vector<double> r, l;
    hipIeAngles(pa, st, r, l);
    double b;
    linearRegression(time, r, rslope, b);
    linearRegression(time, l, lslope, b);
}

void adjustLegsForCorrectedHipIeAngles(const PoseAccess& pa, const Statics& st, const std::vector<double>& time) {
    std::vector<double> r, l;
    hipIeAngles(pa, st, r, l);
    // adjustRegressionSlopeToTargetSlope(time, angle, 0): keep intercept and residuals, drop the slope; then shift so that the first sample is zero
    auto straighten = [&](const std::vector<double>& y) {
        double m, b;
        linearRegression(time, y, m, b);
        std::vector<double> ynew(y.size());
        for (size_t i = 0; i < y.size(); i++) ynew[i] = b + (y[i] - m * time[i] - b);
        const double y0 = ynew[0];
        for (double& v : ynew) v -= y0;
        return ynew;
    };
    const std::vector<double> rnew = straighten(r), lnew = straighten(l);
    const V3 sRHip = v3(st.sacrumImuToRHipCtr), sLHip = v3(st.sacrumImuToLHipCtr);
    const V3 rtHip = v3(st.rthighImuToHipCtr), rtKnee = v3(st.rthighImuToKneeCtr), rsKnee = v3(st.rshankImuToKneeCtr), rsAnkle = v3(st.rshankImuToAnkleCtr), rfAnkle = v3(st.rfootImuToAnkleCtr);
    const V3 ltHip = v3(st.lthighImuToHipCtr), ltKnee = v3(st.lthighImuToKneeCtr), lsKnee = v3(st.lshankImuToKneeCtr), lsAnkle = v3(st.lshankImuToAnkleCtr), lfAnkle = v3(st.lfootImuToAnkleCtr);
    for (int k = 0; k < pa.K; k++) {
        const Pose sacrum = load(pa.at(pa.ctx, 0, k));
        Pose rt = load(pa.at(pa.ctx, 1, k)), rs = load(pa.at(pa.ctx, 2, k)), rf = load(pa.at(pa.ctx, 3, k));
        Pose lt = load(pa.at(pa.ctx, 4, k)), ls = load(pa.at(pa.ctx, 5, k)), lf = load(pa.at(pa.ctx, 6, k));
        const double rhipAdj = rnew[k] - r[k], lhipAdj = lnew[k] - l[k];
        // rotation axes: the femur's long axis (right: distal direction, left: proximal direction), through the hip centre
        const V3 rVec = -1.0 * mul(rt.R, rtHip - rtKnee), lVec = mul(lt.R, ltHip - ltKnee);
        const V3 rPt = transformFrom(rt, rtHip), lPt = transformFrom(lt, ltHip);
        rotateImuPoseAboutPointAxisAngle(rt, rPt, rVec, rhipAdj);
        rotateImuPoseAboutPointAxisAngle(rs, rPt, rVec, rhipAdj);
        rotateImuPoseAboutPointAxisAngle(rf, rPt, rVec, rhipAdj);
        rotateImuPoseAboutPointAxisAngle(lt, lPt, lVec, lhipAdj);
        rotateImuPoseAboutPointAxisAngle(ls, lPt, lVec, lhipAdj);
        rotateImuPoseAboutPointAxisAngle(lf, lPt, lVec, lhipAdj);
        connect(sacrum, rt, rs, rf, sRHip, rtHip, rtKnee, rsKnee, rsAnkle, rfAnkle);
        connect(sacrum, lt, ls, lf, sLHip, ltHip, ltKnee, lsKnee, lsAnkle, lfAnkle);
        store(pa.at(pa.ctx, 1, k), rt); store(pa.at(pa.ctx, 2, k), rs); store(pa.at(pa.ctx, 3, k), rf);
        store(pa.at(pa.ctx, 4, k), lt); store(pa.at(pa.ctx, 5, k), ls); store(pa.at(pa.ctx, 6, k), lf);
    }
}

}  // namespace hipdrift
}  // namespace bioslam_b200
