// Minimal value types of the host layer (stand-ins for gtsam::Pose3 / Rot3 / Point3 / Unit3 / imuBias::ConstantBias at
// the estimator API; GTSAM is not a dependency of this library).  Storage conventions == include/bioslam_b200.h.
#pragma once
#include <array>
#include <cmath>
#include <string>
#include <vector>

namespace bioslam_b200 {

struct Vector3 {
    double v[3] = {0, 0, 0};
    Vector3() = default;
    Vector3(double x, double y, double z) : v{x, y, z} {}
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    Vector3 normalized() const { double n = norm(); return Vector3(v[0] / n, v[1] / n, v[2] / n); }
    Vector3 operator-(const Vector3& o) const { return Vector3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3 operator+(const Vector3& o) const { return Vector3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vector3 operator*(double s) const { return Vector3(v[0] * s, v[1] * s, v[2] * s); }
};
using Point3 = Vector3;
struct Vector2 { double v[2] = {0, 0}; Vector2() = default; Vector2(double a, double b) : v{a, b} {} double operator[](int i) const { return v[i]; } };

struct Rot3 {  // row-major 3x3, body -> nav
    double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double* data() const { return m; }
    Vector3 rotate(const Vector3& p) const {
        return Vector3(m[0] * p[0] + m[1] * p[1] + m[2] * p[2], m[3] * p[0] + m[4] * p[1] + m[5] * p[2], m[6] * p[0] + m[7] * p[1] + m[8] * p[2]);
    }
};
struct Pose3 {
    Rot3 R;
    Point3 t;
    const Rot3& rotation() const { return R; }
    const Point3& translation() const { return t; }
};
struct Unit3 {  // unit 3-vector (constructor normalises, like gtsam::Unit3(x,y,z))
    double p[3] = {0, 0, 1};
    Unit3() = default;
    Unit3(double x, double y, double z) { double n = std::sqrt(x * x + y * y + z * z); p[0] = x / n; p[1] = y / n; p[2] = z / n; }
    Vector3 unitVector() const { return Vector3(p[0], p[1], p[2]); }
};
struct ConstantBias {
    Vector3 acc, gyro;
    ConstantBias() = default;
    ConstantBias(const Vector3& a, const Vector3& g) : acc(a), gyro(g) {}
    const Vector3& accelerometer() const { return acc; }
    const Vector3& gyroscope() const { return gyro; }
};

// CUDA device ordinal every estimator of this process uses (default 0)
void setDevice(int device);
int getDevice();

}  // namespace bioslam_b200
