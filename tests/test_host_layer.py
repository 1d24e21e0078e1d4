"""The C++ host layer (bioslam_b200/host: imuPoseEstimator / lowerBodyPoseEstimator mirrors of the reference classes).
CPU (-m "not gpu"): under the emulator build, setup() must produce exactly the flat graph the independent Python builder
(bioslam_b200/graphs.py, which follows the reference's setup() line by line) produces.  GPU: fastOptimize() == driving the
C ABI directly."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
from bioslam_b200 import abi, host, lib, synth
import graphs

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "cudaemu", "libbioslam_b200_emu.so")


def slots_equal(a, b):
    for f, _ in a._fields_:
        x, y = getattr(a, f), getattr(b, f)
        if hasattr(x, "__len__"):
            if not np.allclose(list(x), list(y), rtol=4e-16, atol=0.0):     # numeric (-0.0 == 0.0; Unit3 normalisation may differ by 1 ulp)
                return False
        elif x != y:
            return False
    return True


def check_same_graph(gh, tvh, svh, gp, tvp, svp):
    assert gh.K == gp.K and gh.time_var_types == gp.time_var_types and gh.static_var_types == gp.static_var_types
    assert gh.const_stride == gp.const_stride and len(gh.slots) == len(gp.slots)
    for i, (a, b) in enumerate(zip(gh.slots, gp.slots)):
        assert slots_equal(a, b), f"slot {i}: type {a.type} vs {b.type}"
    assert np.array_equal(tvh, tvp) and np.allclose(svh, svp, rtol=4e-16, atol=0.0)
    assert np.abs(gh.consts[:, :70] - gp.consts[:, :70]).max() < 1e-13   # device vs oracle preintegration (zeta part)
    assert np.array_equal(gh.consts[:, 1330:], gp.consts[:, 1330:])       # gyro samples


def test_host_setup_builds_the_reference_graph(oracle_lib):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cudaemu")])
    tr = synth.make_trial(seed=8, duration_s=1.5)
    for jv in (0, 1):
        lb = host.lower_body_from_trial(tr, so_path=EMU)
        lb.set(m_useJointVelFactor=jv)
        lb.setup()
        gh, tvh, svh = lb.graph()
        gp, tvp, svp = graphs.lower_body_graph(tr, use_joint_vel=bool(jv))
        check_same_graph(gh, tvh, svh, gp, tvp, svp)


def test_host_error_behaviour(oracle_lib):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cudaemu")])
    tr = synth.make_trial(seed=8, duration_s=1.0)
    m = host.Imu(tr.acc[0], tr.gyro[0], tr.time, so_path=EMU)
    p = host.ImuPoseEstimator(m, "x", so_path=EMU)
    with pytest.raises(host.HostError):
        p.setImuBiasModelMode(2)            # reference: "imuBiasModelMode not recognized"
    with pytest.raises(host.HostError):
        p.setOptimizedValuesToInitialValues()   # before setup()
    p.set(m_initializationScheme=2)
    with pytest.raises(host.HostError):
        p.setup()                            # reference: "unknown initilization scheme. choose 0 or 1."


def _quat_to_R(q):
    w, x, y, z = q.T
    return np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                     np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], -1),
                     np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], -1)], -2)


def test_host_forward_backward_ekf(oracle_lib):
    """SURVEY 8(f) N4, by the reference's own test (test/unit/testEkf.cpp): the F/B EKF orientation of the right-thigh IMU agrees
    with the known orientation to < 5 deg RMSE in roll and pitch (yaw is unobservable without a magnetometer).  Measured fact:
    in the Eigen / gtsam quaternion-to-matrix convention the filter's quaternion is the body->nav rotation (its kinematics are
    q' = q (x) (0, w_body) / 2 and its measurement model predicts M(q)^T e_z for the accelerometer direction).  Upstream labels
    it R[N->B] and seeds the keyframe poses with gtsam::Rot3(q).inverse() (src/imuPoseEstimator.cpp:103-106); scheme 1 of
    setup() reproduces that literally."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cudaemu")])
    tr = synth.make_trial(seed=2, duration_s=20.0)
    q = host.ekf_forward_backward(tr.gyro[1], tr.acc[1], tr.dt, passes=4, so_path=EMU)
    assert np.allclose(np.linalg.norm(q, axis=1), 1.0, atol=1e-12)
    Mq = _quat_to_R(q)
    R_true = tr.poses_at((np.arange(tr.N) + 0.5) * tr.dt)["rthigh"][0]  # R[B->N]
    Rrel = R_true @ np.swapaxes(Mq, -1, -2)
    roll = np.arctan2(Rrel[:, 2, 1], Rrel[:, 2, 2]); pitch = np.arctan2(-Rrel[:, 2, 0], np.hypot(Rrel[:, 0, 0], Rrel[:, 1, 0]))
    rmse = lambda a: np.degrees(np.sqrt(np.mean(a ** 2)))
    assert rmse(roll) < 5.0 and rmse(pitch) < 5.0, (rmse(roll), rmse(pitch))
    # setup() with m_initializationScheme = 1 (3 passes): keyframe rotations = gtsam::Rot3(q).inverse() at the keyframe samples,
    # inserted the way upstream's setup loop does it (src/imuPoseEstimator.cpp:170, 221-223): keyframe 0 <- initPose[0],
    # keyframe k + 1 <- initPose[k]
    m = host.Imu(tr.acc[1], tr.gyro[1], tr.time, so_path=EMU)
    p = host.ImuPoseEstimator(m, "rthigh", so_path=EMU)
    p.set(m_initializationScheme=1)
    p.setup()
    p.setOptimizedValuesToInitialValues()
    pose = p.get("m_pose").reshape(-1, 12)
    q3 = host.ekf_forward_backward(tr.gyro[1], tr.acc[1], tr.dt, passes=3, so_path=EMU)
    idx = graphs.keyframe_sample_idx(p.nKeyframes, 10)
    seed = np.concatenate([idx[:1], idx[:-1]])
    assert np.allclose(pose[:, :9].reshape(-1, 3, 3), np.swapaxes(_quat_to_R(q3[seed]), -1, -2), atol=1e-15)
    assert np.all(pose[:, 9:] == 0.0)


def test_host_post_hoc_axes_check(oracle_lib):
    """reference src/lowerBodyPoseEstimator.cpp:711-797: a knee SHANK axis that points against its thigh axis (mean angle in
    the navigation frame > 90 deg) is flipped in the members; the second pass changes nothing; the derived joint angles
    are those of the consistent axes."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cudaemu")])
    tr = synth.make_trial(seed=8, duration_s=1.5)
    lb = host.lower_body_from_trial(tr, so_path=EMU)
    lb.setup()
    g, tv, sv = lb.graph()
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)           # kinematically consistent values: knee axes agree
    lb.set_values(tv, sv)
    lb.setMemberEstimatedStatesFromValues()
    assert lb.postHocAxesCheck() is False
    lb.setDerivedJointAnglesFromEstimatedStates()
    flex_ok = lb.get("m_rkneeFlexExAng").copy(); lflex_ok = lb.get("m_lkneeFlexExAng").copy()
    shank_ok = lb.get("m_rkneeAxisShank").copy()
    # flip the right knee's shank axis in the values (what an optimisation that converged to the mirrored axis would return)
    so = g.static_value_offsets()
    sv2 = sv.copy()
    i_shank = 12 + 1                                                # static table: 12 joint-centre vectors, then the axes (rkneeAxisThigh, rkneeAxisShank, ...)
    sv2[so[i_shank]:so[i_shank] + 3] *= -1.0
    lb.set_values(tv, sv2)
    lb.setMemberEstimatedStatesFromValues()
    assert np.allclose(lb.get("m_rkneeAxisShank"), -shank_ok)
    assert lb.postHocAxesCheck() is True
    assert np.allclose(lb.get("m_rkneeAxisShank"), shank_ok)        # flipped back in the members
    assert lb.postHocAxesCheck() is False
    lb.setDerivedJointAnglesFromEstimatedStates()                   # uses the members' axes
    assert np.allclose(lb.get("m_rkneeFlexExAng"), flex_ok, atol=1e-12) and np.allclose(lb.get("m_lkneeFlexExAng"), lflex_ok, atol=1e-12)
    _, _, sv_after = lb.graph()
    assert np.array_equal(sv_after, sv2)                            # like the reference, the optimised values are not touched


def test_host_hip_ie_drift_heuristic(oracle_lib):
    """reference src/lowerBodyPoseEstimator.cpp:668-688, 1507-1661: a hip int/ext rotation angle that drifts
    linearly in time is detected by its regression slope; adjustLegsForCorrectedHipIeAngles removes the slope by rotating the
    leg IMUs about the femur's long axis and leaves every joint centre connected."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cudaemu")])
    tr = synth.make_trial(seed=8, duration_s=4.0)
    lb = host.lower_body_from_trial(tr, so_path=EMU)
    lb.setup()
    g, tv, sv = lb.graph()
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv, rot_noise=0.0, pos_noise=0.0)
    lb.set_values(tv, sv)
    r0, l0 = lb.computeHipIeRotAngleRegressionSlopesFromValues()
    thr = 100.0 * np.pi / 180.0 / 30.0
    assert abs(r0) < thr and abs(l0) < thr                       # consistent gait: no drift
    # let the whole right leg drift about the right femur's long axis at 0.2 rad/s, the left leg at -0.15 rad/s
    t = lb.get("m_time")
    assert t.size == g.K
    off, so = g.time_value_offsets(), g.static_value_offsets()
    pose_var = lambda imu: 3 * imu                              # per IMU: pose, velocity, bias (the angular velocities follow)
    stat = lambda i: sv[so[i]:so[i] + 3]
    tv2 = tv.copy()
    for imus, hip, knee, rate in (((1, 2, 3), stat(3), stat(2), 0.2), ((4, 5, 6), stat(8), stat(7), -0.15)):
        for k in range(g.K):
            Rt = tv[k, off[pose_var(imus[0])]:off[pose_var(imus[0])] + 9].reshape(3, 3)
            axis = Rt @ (hip - knee); axis /= np.linalg.norm(axis)
            dR = synth.exp_so3((axis * rate * t[k])[None, :])[0]
            for i in imus:
                o = off[pose_var(i)]
                tv2[k, o:o + 9] = (dR @ tv[k, o:o + 9].reshape(3, 3)).reshape(-1)
    lb.set_values(tv2, sv)
    r1, l1 = lb.computeHipIeRotAngleRegressionSlopesFromValues()
    assert abs(abs(r1) - 0.2) < 0.02 and abs(abs(l1) - 0.15) < 0.02 and abs(r1) > thr and abs(l1) > thr
    lb.adjustLegsForCorrectedHipIeAngles()
    r2, l2 = lb.computeHipIeRotAngleRegressionSlopesFromValues()
    assert abs(r2) < 1e-9 and abs(l2) < 1e-9                     # adjustRegressionSlopeToTargetSlope asserts 1e-10 upstream
    _, tv3, sv3 = lb.graph()
    assert np.array_equal(sv3, sv)
    assert np.array_equal(tv3[:, off[0]:off[0] + 12], tv2[:, off[0]:off[0] + 12])      # the sacrum IMU does not move
    def joint(k, imu, vec):
        o = off[pose_var(imu)]
        return tv3[k, o:o + 9].reshape(3, 3) @ vec + tv3[k, o + 9:o + 12]
    for k in (0, g.K // 2, g.K - 1):
        for (ia, va, ib, vb) in ((0, stat(0), 1, stat(3)), (1, stat(2), 2, stat(4)), (2, stat(5), 3, stat(6)),
                                 (0, stat(1), 4, stat(8)), (4, stat(7), 5, stat(9)), (5, stat(10), 6, stat(11))):
            assert np.abs(joint(k, ia, va) - joint(k, ib, vb)).max() < 1e-12


@pytest.mark.gpu
def test_host_fast_optimize_matches_c_abi():
    tr = synth.make_trial(seed=8, duration_s=3.0)
    lb = host.lower_body_from_trial(tr)
    lb.set(m_maxIterations=12, m_maxNumRestarts=0)           # one LM run: the hip I/E restart hook is tested on its own
    lb.setup()
    g, tv, sv = lb.graph()
    p = lib.Problem(g); p.set_values(tv, sv)
    res, hist = p.optimize(abi.LmParams.lower_body(), abi.OuterParams.lower_body(max_iterations=12))
    lb.fastOptimize()
    h = lb.get("m_optimizationTotalError")
    assert np.array_equal(h, hist)                       # deterministic: bit-identical error history
    tvo, svo = p.get_values()
    _, tvh, svh = lb.graph()
    assert np.array_equal(tvh, tvo) and np.array_equal(svh, svo)
    pose = lb.get("m_rthighImuPose").reshape(-1, 12)
    off = g.time_value_offsets()
    assert np.array_equal(pose, tvo[:, off[3]:off[3] + 12])
    assert np.allclose(np.linalg.norm(lb.get("m_rkneeAxisThigh")), 1.0)
    # derived joint angles through the host class == the C ABI call on the same values
    lb.setDerivedJointAnglesFromEstimatedStates()
    ja = p.joint_angles(graphs.lower_body_joint_angle_desc())
    assert np.array_equal(lb.get("m_rkneeFlexExAng"), ja[:, 0]) and np.array_equal(lb.get("m_lhipIntExtRotAng"), ja[:, 11])
    assert np.array_equal(lb.get("m_rhipAbAdAng"), ja[:, 7])


@pytest.mark.gpu
def test_host_single_imu_fast_optimize():
    tr = synth.make_trial(seed=4, duration_s=10.0)
    m = host.Imu(tr.acc[0], tr.gyro[0], tr.time)
    for mode in (1, 0):
        p = host.ImuPoseEstimator(m, "sacrum")
        p.set(m_useCompassPrior=1, m_refVecLocal=[0, 1, 0], m_refVecNav=[1, 0, 0])
        p.setImuBiasModelMode(mode)
        p.setup()
        K = p.nKeyframes
        R0 = tr.poses_at((graphs.keyframe_sample_idx(K, 10) + 1) * tr.dt)["sacrum"][0] @ synth.exp_so3(np.random.default_rng(0).normal(scale=0.05, size=(K, 3)))
        p.set(initialOrientations=R0.reshape(-1))
        p.fastOptimize()
        err = p.get("m_optimizationTotalError")
        assert err[-1] < 1e-5 * err[0]              # reference solution test: final error <= 1e-5 (testImuEstimationSolution.cpp:20,42)
        assert p.get("m_pose").size == K * 12 and p.get("m_gyroBias").size == (3 * K if mode == 1 else 3)


@pytest.mark.gpu
def test_host_single_imu_reference_solution_test():
    """the reference's own solution test, by its settings and thresholds (test/solution/testImuEstimationSolution.cpp:17-51): one IMU,
    dynamic bias model, zero initialisation, m_absErrLimit = 0.99e-5 -> the final total graph error must be <= 1e-5 ABSOLUTE and the
    estimated velocity must agree with the discrete derivative of the estimated position to < 0.5 m/s RMSE (synthetic right-thigh
    stream here: the reference's HDF5 trial is not in the mount)."""
    tr = synth.make_trial(seed=4, duration_s=10.0)
    m = host.Imu(tr.acc[1], tr.gyro[1], tr.time)
    p = host.ImuPoseEstimator(m, "solutionTest")
    p.setImuBiasModelMode(1)
    p.set(m_useMagnetometer=0, m_initializationScheme=0, m_absErrLimit=0.99e-5)
    p.setup()
    p.fastOptimize()
    err = p.get("m_optimizationTotalError")
    assert err[-1] <= 1.0e-5, err[-5:]
    K = p.nKeyframes
    pos = p.get("m_pose").reshape(K, 12)[:, 9:]
    vel = p.get("m_velocity").reshape(K, 3)
    t = p.get("m_time")
    vd = (pos[1:] - pos[:-1]) / (t[1] - t[0])
    rmse = np.sqrt(np.mean(np.sum((vel[:-1] - vd) ** 2, axis=1)))
    assert rmse < 0.5, rmse


@pytest.mark.gpu
def test_host_single_imu_with_ekf_initialisation():
    """BASELINE configs[0] (exampleImuPoseEstimation.cpp:21-39): single IMU, static-bias ImuFactor, EKF initialisation scheme 1,
    compass + position + velocity + bias priors; the reference's solution test asks for a final error <= 1e-5 x initial."""
    tr = synth.make_trial(seed=4, duration_s=10.0)
    m = host.Imu(tr.acc[0], tr.gyro[0], tr.time)
    p = host.ImuPoseEstimator(m, "sacrum")
    p.set(m_useCompassPrior=1, m_refVecLocal=[0, 1, 0], m_refVecNav=[1, 0, 0], m_initializationScheme=1)
    p.setImuBiasModelMode(0)
    p.setup()
    p.fastOptimize()
    err = p.get("m_optimizationTotalError")
    assert err.size >= 2 and np.all(np.diff(err) <= 1e-9 * err[:-1])      # monotone LM history
    assert err[-1] < 1e-5 * err[0]
    R = p.get("m_pose").reshape(-1, 12)[:, :9].reshape(-1, 3, 3)
    assert np.abs(R @ np.swapaxes(R, 1, 2) - np.eye(3)).max() < 1e-9


@pytest.mark.gpu
def test_host_fast_optimize_restarts_on_hip_ie_drift():
    """the restart recursion of fastOptimize END TO END on the GPU (reference src/lowerBodyPoseEstimator.cpp:668-688): initial values
    whose right leg drifts about the femur's long axis at 0.2 rad/s, LM runs capped at 2 iterations so the drift survives the first
    run -> slope above the threshold -> adjustLegsForCorrectedHipIeAngles -> second LM run from the straightened poses (-> ...);
    the error history is the concatenation of the runs, the final estimate has no drift and a lower cost than the first run."""
    tr = synth.make_trial(seed=8, duration_s=4.0)
    lb = host.lower_body_from_trial(tr)
    lb.set(m_maxIterations=2, m_maxNumRestarts=2)
    lb.setup()
    g, tv, sv = lb.graph()
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv, rot_noise=0.0, pos_noise=0.0)
    t = lb.get("m_time")
    off, so = g.time_value_offsets(), g.static_value_offsets()
    stat = lambda i: sv[so[i]:so[i] + 3]
    tv2 = tv.copy()
    for k in range(g.K):
        Rt = tv[k, off[3]:off[3] + 9].reshape(3, 3)
        axis = Rt @ (stat(3) - stat(2)); axis /= np.linalg.norm(axis)
        dR = synth.exp_so3((axis * 0.2 * t[k])[None, :])[0]
        for i in (1, 2, 3):
            o = off[3 * i]
            tv2[k, o:o + 9] = (dR @ tv[k, o:o + 9].reshape(3, 3)).reshape(-1)
    lb.set_values(tv2, sv)
    thr = 100.0 * np.pi / 180.0 / 30.0
    r_before, _ = lb.computeHipIeRotAngleRegressionSlopesFromValues()
    assert abs(r_before) > thr
    lb.fastOptimize()
    n_restarts = lb.nRestarts
    assert 1 <= n_restarts <= 2
    h = lb.get("m_optimizationTotalError")
    assert h.size == (n_restarts + 1) * 3                      # every run: initial error + 2 iterations
    runs = h.reshape(n_restarts + 1, 3)
    assert np.all(np.diff(runs, axis=1) <= 0)                   # LM never accepts an increase inside a run
    assert runs[-1, -1] < runs[0, -1]                           # the restarted run ends lower than the first one
    r_after, l_after = lb.computeHipIeRotAngleRegressionSlopesFromValues()
    assert abs(r_after) < abs(r_before) and (abs(r_after) < thr or n_restarts == 2)
    assert lb.get("m_rhipIntExtRotAng").size == g.K             # derived angles filled after the last run
