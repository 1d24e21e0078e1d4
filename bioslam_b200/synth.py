"""Seeded synthetic 7-IMU treadmill-gait generator.

The reference's test data (test/data/*.h5) is not available, so BASELINE.json's configs run on synthetic trials:
a kinematic chain pelvis -> thigh -> shank -> foot per leg with hinge knees/ankles and a hinge-dominant hip, sampled
at `rate_hz`, turned into body-frame gyro / specific-force streams with Opal-v2 white noise and a constant bias.

IMU order everywhere: sacrum, rthigh, rshank, rfoot, lthigh, lshank, lfoot (the constructor order of
lowerBodyPoseEstimator, reference include/lowerBodyPoseEstimator.h:152).
Frames follow the reference's prior means (include/lowerBodyPoseEstimator.h:61-78): nav z up, g=(0,0,-9.81)
(examples/exampleLowerBodyEstimation.cpp:18); the subject's right is nav +x (the sacrum compass prior maps the sacrum
IMU's +y to nav +x, examples/exampleLowerBodyEstimation.cpp:25).
"""
import numpy as np

IMU_NAMES = ("sacrum", "rthigh", "rshank", "rfoot", "lthigh", "lshank", "lfoot")
G_NAV = np.array([0.0, 0.0, -9.81])

# prior means, reference include/lowerBodyPoseEstimator.h:61-78
PRIOR_AXES = {
    "rkneeAxisThigh": (0.05, 0.05, 0.95), "rkneeAxisShank": (-0.05, 0.05, 0.95),
    "lkneeAxisThigh": (0.05, 0.05, -0.95), "lkneeAxisShank": (-0.05, 0.05, -0.95),
    "hipAxisSacrum": (0.05, 0.95, 0.05), "rhipAxisThigh": (-0.05, 0.05, 0.95), "lhipAxisThigh": (-0.05, 0.05, -0.95),
    "rankleAxisFoot": (0.05, -0.95, -0.05), "lankleAxisFoot": (-0.05, -0.95, 0.05),
    "rankleAxisShank": (0.05, 0.05, 0.95), "lankleAxisShank": (0.05, 0.05, -0.95),
}
PRIOR_OFFSETS = {
    "sacrumImuToRHipCtr": (.026, .135, -.215), "sacrumImuToLHipCtr": (.024, -.132, -.213),
    "rthighImuToKneeCtr": (.26, -1.0e-5, -1.0e-5), "rthighImuToHipCtr": (-.24, 1.0e-5, 1.0e-5),
    "rshankImuToKneeCtr": (-.11, 1.0e-5, 1.0e-5), "rshankImuToAnkleCtr": (.31, .11, 1.0e-5),
    "rfootImuToAnkleCtr": (-.051, -1.0e-5, -.05),
    "lthighImuToKneeCtr": (.26, -1.0e-5, -1.0e-5), "lthighImuToHipCtr": (-.24, 1.0e-5, 1.0e-5),
    "lshankImuToKneeCtr": (-.1, 1.0e-5, 1.0e-5), "lshankImuToAnkleCtr": (.29, .09, -1.0e-5),
    "lfootImuToAnkleCtr": (-.049, 1.0e-5, -.055),
}


def _unit(v):
    v = np.asarray(v, dtype=np.float64)
    return v / np.linalg.norm(v)


def exp_so3(w):
    """batched Rodrigues: w [...,3] -> R [...,3,3]"""
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w, axis=-1)[..., None, None]
    K = np.zeros(w.shape[:-1] + (3, 3))
    K[..., 0, 1], K[..., 0, 2] = -w[..., 2], w[..., 1]
    K[..., 1, 0], K[..., 1, 2] = w[..., 2], -w[..., 0]
    K[..., 2, 0], K[..., 2, 1] = -w[..., 1], w[..., 0]
    small = th < 1e-9
    ths = np.where(small, 1.0, th)
    a = np.where(small, 1.0 - th * th / 6, np.sin(ths) / ths)
    b = np.where(small, 0.5 - th * th / 24, (1 - np.cos(ths)) / (ths * ths))
    return np.eye(3) + a * K + b * (K @ K)


def log_so3(R):
    """batched log for small/moderate rotations: R [...,3,3] -> w [...,3]"""
    tr = np.clip((np.trace(R, axis1=-2, axis2=-1) - 1) / 2, -1, 1)
    th = np.arccos(tr)
    v = np.stack([R[..., 2, 1] - R[..., 1, 2], R[..., 0, 2] - R[..., 2, 0], R[..., 1, 0] - R[..., 0, 1]], axis=-1)
    s = np.where(th < 1e-7, 0.5 + th * th / 12, th / (2 * np.sin(np.where(th < 1e-7, 1.0, th))))
    return v * s[..., None]


def _neutral_frames():
    """IMU axes in nav at neutral standing, columns = IMU x,y,z (right=+x nav, forward=+y nav, up=+z nav)."""
    right, fwd, up = np.eye(3)
    def frame(x, y, z): return np.stack([x, y, z], axis=1)
    return {
        "sacrum": frame(-fwd, right, up),
        "rthigh": frame(-up, fwd, right), "rshank": frame(-up, fwd, right), "rfoot": frame(fwd, -right, up),
        "lthigh": frame(-up, -fwd, -right), "lshank": frame(-up, -fwd, -right), "lfoot": frame(fwd, -right, up),
    }


class GaitTrial:
    pass


def make_trial(seed=0, duration_s=60.0, rate_hz=100.0, noise=True, bias=True, stride_hz=None):
    """Returns a GaitTrial with acc, gyro [7][N][3] (body frame), dt, true poses and true static parameters."""
    rng = np.random.default_rng(seed)
    N = int(round(duration_s * rate_hz))
    dt = 1.0 / rate_hz
    f = stride_hz if stride_hz is not None else rng.uniform(1.0, 1.2)
    # true static parameters = prior means + perturbation
    axes = {}
    for k, v in PRIOR_AXES.items():
        u = _unit(v)
        axes[k] = exp_so3(rng.normal(scale=np.deg2rad(5.0) / np.sqrt(3), size=3)) @ u
    offs = {k: np.asarray(v) + rng.normal(scale=0.02, size=3) for k, v in PRIOR_OFFSETS.items()}
    N0 = _neutral_frames()

    phase_l = np.pi
    amp = dict(hip=np.deg2rad(25.0), knee_lo=np.deg2rad(5.0), knee_hi=np.deg2rad(65.0), ankle=np.deg2rad(15.0))
    hip_nonhinge = np.deg2rad(2.0)
    pel_ang = np.deg2rad(3.0)

    def poses(t):
        """t [M] -> dict name -> (R [M,3,3], p [M,3])"""
        w = 2 * np.pi * f
        out = {}
        # pelvis: sway +-2 cm, small rotations
        p_s = np.stack([0.02 * np.sin(w * t / 2), 0.015 * np.sin(w * t + 0.3), 1.0 + 0.02 * np.cos(w * t)], axis=-1)
        rot_s = np.stack([pel_ang * np.sin(w * t / 2 + 0.4), 0.5 * pel_ang * np.sin(w * t + 1.0), pel_ang * np.sin(w * t / 2)], axis=-1)
        R_s = exp_so3(rot_s) @ N0["sacrum"]
        out["sacrum"] = (R_s, p_s)
        for side, ph in (("r", 0.0), ("l", phase_l)):
            th_h = amp["hip"] * np.sin(w * t + ph)
            th_k = 0.5 * (amp["knee_lo"] + amp["knee_hi"]) - 0.5 * (amp["knee_hi"] - amp["knee_lo"]) * np.cos(w * t + ph - 0.6)
            th_a = amp["ankle"] * np.sin(w * t + ph + 1.1)
            # hip: hinge about hipAxisSacrum (sacrum frame) plus a small non-hinge wobble ("hinge-dominant")
            k_h = axes["hipAxisSacrum"]
            Rab0 = N0["sacrum"].T @ N0[side + "thigh"]
            wob = np.stack([hip_nonhinge * np.sin(w * t + ph + 0.7), 0 * t, hip_nonhinge * np.cos(w * t + ph)], axis=-1)
            R_rel = exp_so3(th_h[:, None] * k_h) @ exp_so3(wob) @ Rab0
            R_t = R_s @ R_rel
            hip = p_s + np.einsum("mij,j->mi", R_s, offs["sacrumImuTo%sHipCtr" % side.upper()])
            p_t = hip - np.einsum("mij,j->mi", R_t, offs[side + "thighImuToHipCtr"])
            out[side + "thigh"] = (R_t, p_t)
            # knee: perfect hinge about kneeAxisThigh (thigh frame)
            k_k = axes[side + "kneeAxisThigh"]
            Rab0 = N0[side + "thigh"].T @ N0[side + "shank"]
            R_k = R_t @ (exp_so3(th_k[:, None] * k_k) @ Rab0)
            knee = p_t + np.einsum("mij,j->mi", R_t, offs[side + "thighImuToKneeCtr"])
            p_k = knee - np.einsum("mij,j->mi", R_k, offs[side + "shankImuToKneeCtr"])
            out[side + "shank"] = (R_k, p_k)
            # ankle: perfect hinge about ankleAxisShank (shank frame)
            k_a = axes[side + "ankleAxisShank"]
            Rab0 = N0[side + "shank"].T @ N0[side + "foot"]
            R_f = R_k @ (exp_so3(th_a[:, None] * k_a) @ Rab0)
            ankle = p_k + np.einsum("mij,j->mi", R_k, offs[side + "shankImuToAnkleCtr"])
            p_f = ankle - np.einsum("mij,j->mi", R_f, offs[side + "footImuToAnkleCtr"])
            out[side + "foot"] = (R_f, p_f)
        return out

    # true axes on the distal side follow from the hinge geometry (k_B = R_AB0^T k_A)
    for side in ("r", "l"):
        axes[side + "kneeAxisShank"] = (N0[side + "thigh"].T @ N0[side + "shank"]).T @ axes[side + "kneeAxisThigh"]
        axes[side + "ankleAxisFoot"] = (N0[side + "shank"].T @ N0[side + "foot"]).T @ axes[side + "ankleAxisShank"]
        axes[side + "hipAxisThigh"] = (N0["sacrum"].T @ N0[side + "thigh"]).T @ axes["hipAxisSacrum"]

    h = 1.0e-4
    t_mid = (np.arange(N) + 0.5) * dt          # measurements: interval midpoints
    P0, Pm, Pp = poses(t_mid), poses(t_mid - h), poses(t_mid + h)
    acc = np.zeros((7, N, 3)); gyro = np.zeros((7, N, 3))
    for i, name in enumerate(IMU_NAMES):
        R0, p0 = P0[name]; Rm, pm = Pm[name]; Rp, pp = Pp[name]
        gyro[i] = log_so3(np.swapaxes(Rm, -1, -2) @ Rp) / (2 * h)
        a_nav = (pp - 2 * p0 + pm) / (h * h)
        acc[i] = np.einsum("mji,mj->mi", R0, a_nav - G_NAV)
    true_bias = np.zeros((7, 6))
    if bias:
        true_bias[:, :3] = rng.normal(scale=0.02, size=(7, 3))
        true_bias[:, 3:] = rng.normal(scale=0.002, size=(7, 3))
        acc += true_bias[:, None, :3]; gyro += true_bias[:, None, 3:]
    if noise:
        # Opal v2 continuous sigmas (reference include/imu/imuNoiseModelHandler.h:29-32); sigma_d = sigma_c / sqrt(dt)
        acc += rng.normal(scale=0.04903325 / np.sqrt(dt), size=acc.shape)
        gyro += rng.normal(scale=0.00015707963267949 / np.sqrt(dt), size=gyro.shape)

    tr = GaitTrial()
    tr.seed, tr.N, tr.dt, tr.rate_hz, tr.stride_hz = seed, N, dt, rate_hz, f
    tr.acc, tr.gyro = acc, gyro
    tr.time = np.arange(N) * dt
    tr.true_axes, tr.true_offsets, tr.true_bias = axes, offs, true_bias
    tr.poses_at = poses
    return tr
