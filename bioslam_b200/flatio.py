"""Flat-file stand-in for the reference's HDF5 I/O: the image has no HDF5 library, so the two file formats
of the reference are kept BY NAME in a NumPy `.npz` container (a zip of arrays): every key is the HDF5 path of the dataset or attribute
the reference reads or writes, attributes spelled `<group path>@<attribute name>`.

  * input  == APDM-v5 Opal trial, as read by imu::getImuMapFromDataFile / readSingleImuDataFromApdmOpalH5FileByLabel
             (reference src/imu/imu.cpp:146-259): `/@FileFormatVersion` (5), `/Sensors/<id>/{Accelerometer, Gyroscope, Magnetometer}` [N][3],
             `/Sensors/<id>/Time` (unix microseconds; the reader returns seconds relative to the first sample),
             `/Sensors/<id>/Configuration@Label 0` (the label the estimators look sensors up by), `/Processed/<id>/Orientation` [N][4] (w x y z).
  * output == lowerBodyPoseEstimator::saveResultsToH5File (reference src/lowerBodyPoseEstimator.cpp:848-1074): `/Estimated/*`, `/Derived/*`,
             `/General/*`, `/ModelSettings/*` with the reference's dataset names.

A bioslam maintainer with h5py can convert either way with ten lines (`convert_npz_to_h5` below, skipped when h5py is absent); the MATLAB
post-processing of the reference reads the same names."""
import numpy as np

LOWER_BODY_LABELS = ("Sacrum", "Right Thigh", "Right Tibia", "Right Foot", "Left Thigh", "Left Tibia", "Left Foot")
_IMU_H5 = ("SacrumImu", "RThighImu", "RShankImu", "RFootImu", "LThighImu", "LShankImu", "LFootImu")


def write_apdm_trial(path, trial, labels=LOWER_BODY_LABELS, t0_unix_us=1_491_925_666_000_000, orientation=True):
    """synthetic trial (bioslam_b200.synth.Trial: acc / gyro [n_imu][N][3], dt) -> APDM-v5 names"""
    d = {"/@FileFormatVersion": np.array(5.0)}
    N = trial.acc.shape[1]
    t_us = t0_unix_us + np.round(np.arange(N) * trial.dt * 1e6)
    for i, label in enumerate(labels):
        sid = "XI-%06d" % (1000 + i)
        d[f"/Sensors/{sid}/Accelerometer"] = np.asarray(trial.acc[i], dtype=np.float64)
        d[f"/Sensors/{sid}/Gyroscope"] = np.asarray(trial.gyro[i], dtype=np.float64)
        mag = getattr(trial, "mag", None)
        d[f"/Sensors/{sid}/Magnetometer"] = np.asarray(mag[i], dtype=np.float64) if mag is not None else np.zeros((N, 3))
        d[f"/Sensors/{sid}/Time"] = t_us.astype(np.float64)      # (APDM stores uint64 or double depending on the exporter version: imu.cpp:163-166)
        d[f"/Sensors/{sid}/Configuration@Label 0"] = np.array(label)
        if orientation:
            d[f"/Processed/{sid}/Orientation"] = np.tile(np.array([1.0, 0.0, 0.0, 0.0]), (N, 1))
    np.savez_compressed(path, **d)


def is_apdm_v5(path):
    """reference is_apdm_h5_version5 (src/imu/imu.cpp:216-234)"""
    try:
        with np.load(path, allow_pickle=False) as z:
            return float(z["/@FileFormatVersion"]) == 5.0
    except Exception:
        return False


def read_imu_map(path):
    """== imu::getImuMapFromDataFile: {label: dict(acc, gyro, mag [N][3], time [N] seconds from the first sample, unix_us, quat [N][4] or None, id)}.
    Raises like the reference when the file is not APDM v5 or a sensor has no 'Label 0'."""
    if not is_apdm_v5(path):
        raise RuntimeError("this is not a valid v5 apdm file")
    out = {}
    with np.load(path, allow_pickle=False) as z:
        sids = sorted({k.split("/")[2] for k in z.files if k.startswith("/Sensors/")})
        for sid in sids:
            key = f"/Sensors/{sid}/Configuration@Label 0"
            if key not in z.files:
                raise RuntimeError("ERROR: label 'Label 0' not found")
            label = str(z[key])
            t_us = np.asarray(z[f"/Sensors/{sid}/Time"], dtype=np.float64)
            q = z[f"/Processed/{sid}/Orientation"] if f"/Processed/{sid}/Orientation" in z.files else None
            out[label] = {"acc": z[f"/Sensors/{sid}/Accelerometer"], "gyro": z[f"/Sensors/{sid}/Gyroscope"], "mag": z[f"/Sensors/{sid}/Magnetometer"],
                          "time": (t_us - t_us[0]) / 1.0e6, "unix_us": t_us, "quat": q, "id": int("".join(c for c in sid if c.isdigit()))}
    return out


def lower_body_results(lb, model_id="lowerbody"):
    """the datasets of saveResultsToH5File from a bioslam_b200.host.LowerBodyPoseEstimator after fastOptimize()"""
    g = lambda name: np.asarray(lb.get(name))
    K = g("m_time").size
    d = {"/General/ID": np.array(model_id), "/General/GeneratingModelType": np.array("lowerBodyPoseEstimator"), "/General/nKeyframes": np.array(K),
         "/Time": g("m_time")}
    member = ("sacrum", "rthigh", "rshank", "rfoot", "lthigh", "lshank", "lfoot")
    for m, h5 in zip(member, _IMU_H5):
        pose = g(f"m_{m}ImuPose").reshape(K, 12)
        R = pose[:, :9].reshape(K, 3, 3)
        d[f"/Estimated/R_{h5}_to_N"] = R
        d[f"/Estimated/q_{h5}_to_N"] = _rot_to_quat(R)
        d[f"/Estimated/p_{h5}"] = pose[:, 9:12]
        d[f"/Estimated/v_{h5}"] = g(f"m_{m}ImuVelocity").reshape(-1, 3)
        d[f"/Estimated/gyrobias_{h5}"] = g(f"m_{m}ImuGyroBias").reshape(-1, 3)
        d[f"/Estimated/accelbias_{h5}"] = g(f"m_{m}ImuAccelBias").reshape(-1, 3)
        d[f"/Estimated/{h5}AngVel"] = g(f"m_{m}ImuAngVel").reshape(-1, 3)
    statics = {"SacrumImuToRHipCtr": "m_sacrumImuToRHipCtr", "SacrumImuToLHipCtr": "m_sacrumImuToLHipCtr", "RThighImuToRHipCtr": "m_rthighImuToHipCtr",
               "RThighImuToRKneeCtr": "m_rthighImuToKneeCtr", "RShankImuToRKneeCtr": "m_rshankImuToKneeCtr", "RShankImuToRAnkleCtr": "m_rshankImuToAnkleCtr",
               "RFootImuToRAnkleCtr": "m_rfootImuToAnkleCtr", "LThighImuToLHipCtr": "m_lthighImuToHipCtr", "LThighImuToLKneeCtr": "m_lthighImuToKneeCtr",
               "LShankImuToLKneeCtr": "m_lshankImuToKneeCtr", "LShankImuToLAnkleCtr": "m_lshankImuToAnkleCtr", "LFootImuToLAnkleCtr": "m_lfootImuToAnkleCtr",
               "RKneeAxis_RThigh": "m_rkneeAxisThigh", "RKneeAxis_RShank": "m_rkneeAxisShank", "LKneeAxis_LThigh": "m_lkneeAxisThigh",
               "LKneeAxis_LShank": "m_lkneeAxisShank", "HipAxis_Sacrum": "m_hipAxisSacrum", "RHipAxis_RThigh": "m_rhipAxisThigh", "LHipAxis_LThigh": "m_lhipAxisThigh",
               # (the reference writes the ankle axes under these four names, src/lowerBodyPoseEstimator.cpp:848-1074)
               "RAnkleAxis_RThigh": "m_rankleAxisShank", "RAnkleAxis_Foot": "m_rankleAxisFoot", "LAnkleAxis_LThigh": "m_lankleAxisShank", "LAnkleAxis_Foot": "m_lankleAxisFoot"}
    for h5, mem in statics.items():
        d[f"/Estimated/{h5}"] = g(mem)
    for joint, mem in (("RKnee", "m_rknee"), ("LKnee", "m_lknee"), ("RHip", "m_rhip"), ("LHip", "m_lhip")):
        for ang in ("FlexEx", "AbAd", "IntExtRot"):
            d[f"/Derived/{joint}Angles_{ang}"] = g(f"{mem}{ang}Ang")
    return d


def write_results(path, lb, model_id="lowerbody"):
    np.savez_compressed(path, **lower_body_results(lb, model_id))


def _rot_to_quat(R):
    """[K][3][3] -> [K][4] (w x y z), w >= 0"""
    K = R.shape[0]
    q = np.zeros((K, 4))
    tr = np.trace(R, axis1=1, axis2=2)
    for k in range(K):
        m = R[k]
        if tr[k] > 0:
            s = np.sqrt(tr[k] + 1.0) * 2; q[k] = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
        else:
            i = int(np.argmax(np.diag(m))); j, l = (i + 1) % 3, (i + 2) % 3
            s = np.sqrt(1.0 + m[i, i] - m[j, j] - m[l, l]) * 2
            q[k, 0] = (m[l, j] - m[j, l]) / s; q[k, 1 + i] = 0.25 * s; q[k, 1 + j] = (m[j, i] + m[i, j]) / s; q[k, 1 + l] = (m[l, i] + m[i, l]) / s
        if q[k, 0] < 0: q[k] = -q[k]
    return q


def convert_npz_to_h5(npz_path, h5_path):
    """for a machine that has h5py (not this image): the same names as real HDF5 datasets / attributes"""
    import h5py
    with np.load(npz_path, allow_pickle=False) as z, h5py.File(h5_path, "w") as f:
        for k in z.files:
            if "@" in k:
                grp, attr = k.split("@")
                f.require_group(grp or "/").attrs[attr] = z[k].item() if z[k].ndim == 0 else z[k]
            else:
                f.create_dataset(k, data=z[k])
