// IMU data container mirroring the reference's `imu` class (reference include/imu/imu.h:16-41): same public data members;
// the HDF5 reader is out of scope -- data are supplied as arrays.
#pragma once
#include "b200_types.h"

namespace bioslam_b200 {

class imu {
  public:
    imu() = default;
    std::vector<double> gx, gy, gz, ax, ay, az, mx, my, mz, qs, qx, qy, qz;
    std::vector<double> relTimeSec;
    std::vector<double> unixTimeUtc;
    unsigned id = 0;
    std::string label;
    unsigned long length() const { return (unsigned long)ax.size(); }
    Vector3 gyroVec(unsigned i) const { return Vector3(gx[i], gy[i], gz[i]); }
    Vector3 accelVec(unsigned i) const { return Vector3(ax[i], ay[i], az[i]); }
    Vector3 magVec(unsigned i) const { return Vector3(mx[i], my[i], mz[i]); }
    // mean sample period (reference src/imu/imu.cpp:261-269)
    double getDeltaT() const {
        double s = 0;
        for (size_t i = 1; i < relTimeSec.size(); i++) s += relTimeSec[i] - relTimeSec[i - 1];
        return relTimeSec.size() > 1 ? s / (double)(relTimeSec.size() - 1) : 0.0;
    }
};

}  // namespace bioslam_b200
