"""Python-side construction of the reference's factor graphs as flat GraphDesc objects (used by bench.py and the tests).

Independent of the C++ host layer (bioslam_b200/host) so that the two can be checked against each other.
`preint` is the preintegration callable (acc, gyro, dt, n_per, bias_hat) -> records; the product passes the device
implementation (bioslam_b200.lib.preintegrate), the tests may pass the CPU checker.
Topology follows reference src/imuPoseEstimator.cpp:152-233 (single IMU chain) and
src/lowerBodyPoseEstimator.cpp:29-135,158-290,361-505 (7-IMU weave) with the defaults of
include/lowerBodyPoseEstimator.h:24-89 and the example's priors (examples/exampleLowerBodyEstimation.cpp:21-55).
"""
import numpy as np
from bioslam_b200 import abi, synth

P, V, B, U = abi.VAR_POSE3, abi.VAR_VEC3, abi.VAR_BIAS6, abi.VAR_UNIT3
K_, K1, ST = abi.AT_K, abi.AT_K1, abi.AT_STATIC
G = (0.0, 0.0, -9.81)

STATIC_POINTS = ["sacrumImuToRHipCtr", "sacrumImuToLHipCtr", "rthighImuToKneeCtr", "rthighImuToHipCtr", "rshankImuToKneeCtr",
                 "rshankImuToAnkleCtr", "rfootImuToAnkleCtr", "lthighImuToKneeCtr", "lthighImuToHipCtr", "lshankImuToKneeCtr",
                 "lshankImuToAnkleCtr", "lfootImuToAnkleCtr"]
STATIC_AXES = ["rkneeAxisThigh", "rkneeAxisShank", "lkneeAxisThigh", "lkneeAxisShank", "hipAxisSacrum", "rhipAxisThigh", "lhipAxisThigh",
               "rankleAxisShank", "lankleAxisShank", "rankleAxisFoot", "lankleAxisFoot"]
AXIS_PRIOR_STD = {"hipAxisSacrum": 0.5, "rhipAxisThigh": 0.5, "lhipAxisThigh": 0.5}  # others 0.15


def n_keyframes(n_samples, n_per):
    return (n_samples + n_per - 1) // n_per


def keyframe_sample_idx(K, n_per):
    """reference src/imuPoseEstimator.cpp:62-70"""
    idx = np.arange(K) * n_per - 1
    idx[0] = 0
    return idx


def lower_body_graph(trial, n_per=10, preint=None, use_joint_vel=False, hinge_dim=3, joint_pos_dim=3, magnetometer=False, global_B=(19.5, -5.1, -47.8),
                     joint_vel_dim=3):
    """Returns (GraphDesc, time_values [K][168], static_values [69]) for the 7-IMU configuration.  Defaults = the reference's
    (include/lowerBodyPoseEstimator.h:24-89); hinge_dim / joint_pos_dim = 1 select the norm-error factors
    (m_hingeAxisFactorDim / m_jointPosFactorDim, sigma = norm of the 3-vector sigma, src/lowerBodyPoseEstimator.cpp:296-312);
    magnetometer adds a MagPose3Factor per IMU on every pose k+1 (m_useMagnetometer, src/imuPoseEstimator.cpp:216-220) with a
    synthetic field measurement R_true^T B_global (sigma 3 uT, include/imuPoseEstimator.h:43)."""
    if preint is None:
        raise ValueError("preint callable required (e.g. bioslam_b200.lib.preintegrate)")
    N = trial.acc.shape[1]
    K = n_keyframes(N, n_per)
    tvt = []
    for _ in range(7):
        tvt += [P, V, B]
    tvt += [V] * 7
    pose = lambda i: 3 * i
    vel = lambda i: 3 * i + 1
    bias = lambda i: 3 * i + 2
    angv = lambda i: 21 + i
    svt = [V] * 12 + [U] * 11
    sidx = {n: i for i, n in enumerate(STATIC_POINTS + STATIC_AXES)}
    g = abi.GraphDesc(K, tvt, svt, const_stride=7 * 190 + 7 * 3 + (21 if magnetometer else 0))
    rec = preint(trial.acc, trial.gyro, trial.dt, n_per, np.zeros((7, 6)))  # [7][K-1][190]
    kidx = keyframe_sample_idx(K, n_per)
    for i in range(7):
        g.consts[:K - 1, 190 * i:190 * (i + 1)] = rec[i]
        g.consts[:, 1330 + 3 * i:1333 + 3 * i] = trial.gyro[i, kidx]
    # --- per-IMU sub-graphs (imuPoseEstimator::setupCombinedImuFactor)
    for i in range(7):
        if i == 0:  # sacrum: position, velocity and compass priors (examples/exampleLowerBodyEstimation.cpp:36-41)
            g.add_slot(abi.F_POSE3_TRANSLATION_PRIOR, [(K_, pose(0))], [1e-3] * 3 + [1e-8] * 3, 0, 1)
            g.add_slot(abi.F_PRIOR_VEC3, [(K_, vel(0))], [1e-3] * 3 + [1e-8] * 3, 0, 1)
            g.add_slot(abi.F_POSE3_COMPASS_PRIOR, [(K_, pose(0))], [1e-3, 0.0, 0, 1, 0, 1, 0, 0, 0, 0, 1], 0, 1)
        g.add_slot(abi.F_IMU_COMBINED, [(K_, pose(i)), (K_, vel(i)), (K1, pose(i)), (K1, vel(i)), (K_, bias(i)), (K1, bias(i))], G,
                   const_offset=190 * i)
    # --- angular velocity factors (lowerBodyPoseEstimator.cpp:97-103); sigma = 0.000157079633*0.5
    for i in range(7):
        g.add_slot(abi.F_ANGVEL, [(K_, angv(i)), (K_, bias(i))], [0.000157079633 * 0.5] * 3, const_offset=1330 + 3 * i)
    SAC, RT, RS, RF, LT, LS, LF = range(7)
    if magnetometer:
        t_k = (kidx + 1) * trial.dt
        t_k[0] = 0.0
        poses_true = trial.poses_at(t_k)
        Bn = np.asarray(global_B, dtype=float)
        for i, name in enumerate(synth.IMU_NAMES):
            Rt = poses_true[name][0]                                  # [K][3][3] body -> nav
            g.consts[:, 1351 + 3 * i:1354 + 3 * i] = np.einsum("kji,j->ki", Rt, Bn)
            g.add_slot(abi.F_MAG_POSE3, [(K_, pose(i))], [3.0, 3.0, 3.0] + list(Bn), 1, K, const_offset=1351 + 3 * i)
    def hinge(a, b, axis, sig):
        if hinge_dim == 1:
            g.add_slot(abi.F_HINGE_NORM, [(K_, pose(a)), (K_, angv(a)), (K_, pose(b)), (K_, angv(b)), (ST, sidx[axis])], [sig * np.sqrt(3.0)])
            return
        g.add_slot(abi.F_HINGE_VEC, [(K_, pose(a)), (K_, angv(a)), (K_, pose(b)), (K_, angv(b)), (ST, sidx[axis])], [sig] * 3)
    # knee (:243-246), hip (:213-216), ankle (:265-268)
    hinge(RT, RS, "rkneeAxisThigh", 10.0); hinge(RS, RT, "rkneeAxisShank", 10.0)
    hinge(LT, LS, "lkneeAxisThigh", 10.0); hinge(LS, LT, "lkneeAxisShank", 10.0)
    hinge(SAC, RT, "hipAxisSacrum", 20.0); hinge(SAC, LT, "hipAxisSacrum", 20.0)
    hinge(RT, SAC, "rhipAxisThigh", 20.0); hinge(LT, SAC, "lhipAxisThigh", 20.0)
    hinge(RF, RS, "rankleAxisFoot", 20.0); hinge(LF, LS, "lankleAxisFoot", 20.0)
    hinge(RS, RF, "rankleAxisShank", 20.0); hinge(LS, LF, "lankleAxisShank", 20.0)
    hipN = np.array([9.135, 16.517, 29.378]) * 1.0e-2 * 1.2
    kneeN = np.array([6.7348, 12.3006, 10.1090]) * 1.0e-2 * 1.2
    ankN = np.array([7.3741, 6.2212, 4.6993]) * 1.0e-2 * 1.2
    def joint(a, b, sa, sb, sig):
        if joint_pos_dim == 1:   # expected distance 0 (include/factors/ConstrainedJointCenterPositionFactor.h default)
            g.add_slot(abi.F_JOINTPOS_NORM, [(K_, pose(a)), (K_, pose(b)), (ST, sidx[sa]), (ST, sidx[sb])], [float(np.linalg.norm(sig)), 0.0])
            return
        g.add_slot(abi.F_JOINTPOS_VEC, [(K_, pose(a)), (K_, pose(b)), (ST, sidx[sa]), (ST, sidx[sb])], sig)
    joint(SAC, RT, "sacrumImuToRHipCtr", "rthighImuToHipCtr", hipN)
    joint(RT, RS, "rthighImuToKneeCtr", "rshankImuToKneeCtr", kneeN)
    joint(RS, RF, "rshankImuToAnkleCtr", "rfootImuToAnkleCtr", ankN)
    joint(SAC, LT, "sacrumImuToLHipCtr", "lthighImuToHipCtr", hipN)
    joint(LT, LS, "lthighImuToKneeCtr", "lshankImuToKneeCtr", kneeN)
    joint(LS, LF, "lshankImuToAnkleCtr", "lfootImuToAnkleCtr", ankN)
    if use_joint_vel:
        def jvel(a, b, sa, sb):
            # m_jointVelFactorDim == 1: ConstrainedJointCenterNormVelocityFactor (reference src/lowerBodyPoseEstimator.cpp:194-204)
            ft, sig = (abi.F_JOINTVEL_VEC, [10.0] * 3) if joint_vel_dim == 3 else (abi.F_JOINTVEL_NORM, [float(np.linalg.norm([10.0] * 3))])
            g.add_slot(ft, [(K_, pose(a)), (K_, vel(a)), (K_, angv(a)), (ST, sidx[sa]),
                            (K_, pose(b)), (K_, vel(b)), (K_, angv(b)), (ST, sidx[sb])], sig)
        jvel(SAC, RT, "sacrumImuToRHipCtr", "rthighImuToHipCtr"); jvel(RT, RS, "rthighImuToKneeCtr", "rshankImuToKneeCtr")
        jvel(RS, RF, "rshankImuToAnkleCtr", "rfootImuToAnkleCtr"); jvel(SAC, LT, "sacrumImuToLHipCtr", "lthighImuToHipCtr")
        jvel(LT, LS, "lthighImuToKneeCtr", "lshankImuToKneeCtr"); jvel(LS, LF, "lshankImuToAnkleCtr", "lfootImuToAnkleCtr")
    # --- static priors
    for n in STATIC_AXES:
        std = AXIS_PRIOR_STD.get(n, 0.15)
        g.add_slot(abi.F_PRIOR_UNIT3, [(ST, sidx[n])], [std, std] + list(synth._unit(synth.PRIOR_AXES[n])))
    for n in STATIC_POINTS:
        g.add_slot(abi.F_PRIOR_VEC3, [(ST, sidx[n])], [0.5] * 3 + list(synth.PRIOR_OFFSETS[n]))
    # --- anthropometry (lowerBodyPoseEstimator.cpp:430-504)
    A4 = 4.000000000000001
    def seg(a, b, L, sig, mx, mn, swap=False):
        g.add_slot(abi.F_SEGLEN_MAG, [(ST, sidx[a]), (ST, sidx[b])], [sig, L])
        a2, b2 = (b, a) if swap else (a, b)
        g.add_slot(abi.F_SEGLEN_MAX, [(ST, sidx[a2]), (ST, sidx[b2])], [1e-3, mx, A4])
        g.add_slot(abi.F_SEGLEN_MIN, [(ST, sidx[a2]), (ST, sidx[b2])], [1e-3, mn, A4])
    seg("sacrumImuToRHipCtr", "sacrumImuToLHipCtr", 0.1866, 0.0098, 0.409, 0.10, swap=True)
    seg("rthighImuToHipCtr", "rthighImuToKneeCtr", 0.3943, 0.0302, 0.48, 0.326)
    seg("rshankImuToKneeCtr", "rshankImuToAnkleCtr", 0.4110, 0.0264, 0.479, 0.344)
    seg("lthighImuToHipCtr", "lthighImuToKneeCtr", 0.3943, 0.0302, 0.48, 0.326)
    seg("lshankImuToKneeCtr", "lshankImuToAnkleCtr", 0.4110, 0.0264, 0.479, 0.344)
    g.add_slot(abi.F_SEGLEN_DISCREPANCY, [(ST, sidx[n]) for n in ("rthighImuToHipCtr", "rthighImuToKneeCtr", "lthighImuToHipCtr", "lthighImuToKneeCtr")], [0.002])
    g.add_slot(abi.F_SEGLEN_DISCREPANCY, [(ST, sidx[n]) for n in ("rshankImuToKneeCtr", "rshankImuToAnkleCtr", "lshankImuToKneeCtr", "lshankImuToAnkleCtr")], [0.002])
    d2r = np.pi / 180
    g.add_slot(abi.F_ANGLE_AXIS_SEG, [(ST, sidx["rkneeAxisShank"]), (ST, sidx["rshankImuToKneeCtr"]), (ST, sidx["rshankImuToAnkleCtr"])], [1.2 * d2r, 92 * d2r])
    g.add_slot(abi.F_ANGLE_AXIS_SEG, [(ST, sidx["lkneeAxisShank"]), (ST, sidx["lshankImuToKneeCtr"]), (ST, sidx["lshankImuToAnkleCtr"])], [1.2 * d2r, 88 * d2r])
    g.add_slot(abi.F_ANGLE_AXIS_SEG, [(ST, sidx["rkneeAxisThigh"]), (ST, sidx["rthighImuToHipCtr"]), (ST, sidx["rthighImuToKneeCtr"])], [2.4 * d2r, 84 * d2r])
    g.add_slot(abi.F_ANGLE_AXIS_SEG, [(ST, sidx["lkneeAxisThigh"]), (ST, sidx["lthighImuToHipCtr"]), (ST, sidx["lthighImuToKneeCtr"])], [2.4 * d2r, 96 * d2r])
    # --- initial values: poses identity/zero, vel 0, bias 0, ang-vel = gyro - bias (lowerBodyPoseEstimator.cpp:116-126), statics at prior means
    tv = np.zeros((K, g.time_value_dim))
    toff = g.time_value_offsets()
    for i in range(7):
        tv[:, toff[pose(i)]:toff[pose(i)] + 9] = np.eye(3).ravel()
        tv[:, toff[angv(i)]:toff[angv(i)] + 3] = trial.gyro[i, kidx]
    sv = np.zeros(g.static_value_dim)
    soff = g.static_value_offsets()
    for n in STATIC_POINTS:
        sv[soff[sidx[n]]:soff[sidx[n]] + 3] = synth.PRIOR_OFFSETS[n]
    for n in STATIC_AXES:
        sv[soff[sidx[n]]:soff[sidx[n]] + 3] = synth._unit(synth.PRIOR_AXES[n])
    return g, tv, sv


JOINT_ANGLE_POSE_IMUS = (0, 1, 2, 4, 5)   # sacrum, right thigh, right shank, left thigh, left shank (IMU order of lower_body_graph)
JOINT_ANGLE_STATICS = ["sacrumImuToRHipCtr", "sacrumImuToLHipCtr", "rkneeAxisThigh", "rthighImuToHipCtr", "rthighImuToKneeCtr",
                       "rkneeAxisShank", "rshankImuToKneeCtr", "rshankImuToAnkleCtr", "lkneeAxisThigh", "lthighImuToHipCtr",
                       "lthighImuToKneeCtr", "lkneeAxisShank", "lshankImuToKneeCtr", "lshankImuToAnkleCtr"]


def lower_body_joint_angle_desc():
    """variable indices of lower_body_graph() for b200_lower_body_joint_angles (reference src/lowerBodyPoseEstimator.cpp:813-835)"""
    sidx = {n: i for i, n in enumerate(STATIC_POINTS + STATIC_AXES)}
    d = abi.JointAngleDesc()
    for j, i in enumerate(JOINT_ANGLE_POSE_IMUS):
        d.pose_var[j] = 3 * i
    for j, n in enumerate(JOINT_ANGLE_STATICS):
        d.static_var[j] = sidx[n]
    return d


def single_imu_graph(acc, gyro, dt, n_per=10, static_bias=False, preint=None, init_R=None, ref_local=(0.0, 1.0, 0.0)):
    """imuPoseEstimator graph for one IMU (acc, gyro [N][3]) with position / velocity / bias / compass priors
    (examples/exampleImuPoseEstimation.cpp:21-39).  Returns (GraphDesc, tv, sv)."""
    if preint is None:
        raise ValueError("preint callable required (e.g. bioslam_b200.lib.preintegrate)")
    N = acc.shape[0]
    K = n_keyframes(N, n_per)
    rec = preint(acc[None], gyro[None], dt, n_per, np.zeros((1, 6)))[0]
    nrec = rec.shape[1]
    if static_bias:
        g = abi.GraphDesc(K, [P, V], [B], const_stride=nrec)
        g.consts[:K - 1, :] = rec
        g.add_slot(abi.F_IMU_STATICBIAS, [(K_, 0), (K_, 1), (K1, 0), (K1, 1), (ST, 0)], G, const_offset=0)
        g.add_slot(abi.F_PRIOR_BIAS6, [(ST, 0)], [1e-3] * 6 + [1e-8] * 6)
    else:
        g = abi.GraphDesc(K, [P, V, B], [], const_stride=nrec)
        g.consts[:K - 1, :] = rec
        g.add_slot(abi.F_IMU_COMBINED, [(K_, 0), (K_, 1), (K1, 0), (K1, 1), (K_, 2), (K1, 2)], G, const_offset=0)
        g.add_slot(abi.F_PRIOR_BIAS6, [(K_, 2)], [1e-3] * 6 + [1e-8] * 6, 0, 1)
    g.add_slot(abi.F_POSE3_TRANSLATION_PRIOR, [(K_, 0)], [1e-3] * 3 + [1e-8] * 3, 0, 1)
    g.add_slot(abi.F_PRIOR_VEC3, [(K_, 1)], [1e-3] * 3 + [1e-8] * 3, 0, 1)
    g.add_slot(abi.F_POSE3_COMPASS_PRIOR, [(K_, 0)], [1e-3, 0.0] + list(ref_local) + [1, 0, 0, 0, 0, 1], 0, 1)
    tv = np.zeros((K, g.time_value_dim))
    tv[:, 0:9] = np.eye(3).ravel() if init_R is None else init_R.reshape(K, 9)
    sv = np.zeros(g.static_value_dim)
    return g, tv, sv


def perturbed_truth_values(trial, g, tv, sv, n_per=10, rot_noise=0.05, pos_noise=0.02, seed=1):
    """Initial values near the generator's truth (keeps LM tests short): true IMU poses with small noise."""
    rng = np.random.default_rng(seed)
    K = g.K
    t_k = (keyframe_sample_idx(K, n_per) + 1) * trial.dt
    t_k[0] = 0.0
    poses = trial.poses_at(t_k)
    toff = g.time_value_offsets()
    tv = tv.copy()
    for i, name in enumerate(synth.IMU_NAMES):
        R, p = poses[name]
        Rn = R @ synth.exp_so3(rng.normal(scale=rot_noise, size=(K, 3)))
        tv[:, toff[3 * i]:toff[3 * i] + 9] = Rn.reshape(K, 9)
        tv[:, toff[3 * i] + 9:toff[3 * i] + 12] = p + rng.normal(scale=pos_noise, size=(K, 3))
        vtrue = (trial.poses_at(t_k + 1e-4)[name][1] - trial.poses_at(t_k - 1e-4)[name][1]) / 2e-4
        tv[:, toff[3 * i + 1]:toff[3 * i + 1] + 3] = vtrue
    return tv, sv
