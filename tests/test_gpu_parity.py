"""-m gpu: the parity tests proper.  The CUDA library (through the C ABI) against the oracle on identical seeded inputs,
at sizes the oracle finishes in seconds, plus size-independent properties at the bench size.

Tolerances (north_star): final cost within 1e-6 relative; per-factor sums 1e-11 relative (fp64, different summation
order); positions within 1e-6 m and hip/knee angles within 1e-3 deg at convergence (see test_converged_solution_parity)."""
import numpy as np
import pytest
from bioslam_b200 import abi, lib, synth
import graphs
import parity_checks as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def po(oracle_lib):
    return oracle_lib


def lower_body(seconds, seed=3, near_truth=True, device_preint=False):
    tr = synth.make_trial(seed=seed, duration_s=seconds)
    pre = (lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh)) if device_preint else None
    g, tv, sv = graphs.lower_body_graph(tr, preint=pre)
    if near_truth:
        tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    return tr, g, tv, sv


def test_library_is_the_cuda_build():
    L = lib.load()
    assert not hasattr(L, "emu_marker")
    import torch
    assert torch.cuda.is_available()


@pytest.mark.parametrize("chunks", [1, 2, 5, None])
def test_linearization_and_solve_parity(po, chunks):
    _, g, tv, sv = lower_body(6.0)
    o = po.OracleProblem(g, threads=8); o.set_values(tv, sv)
    p = lib.Problem(g, chunks=chunks); p.set_values(tv, sv)
    tv2, sv2 = p.get_values()
    assert np.array_equal(tv2, tv) and np.array_equal(sv2, sv)
    pc.assert_linearization_parity(p, o, g)
    pc.assert_solve_parity(p, o, g, lams=(1e-5, 1e-3, 1.0, 1e3))
    p.close()


def test_preintegration_parity(po):
    tr = synth.make_trial(seed=9, duration_s=20.0)
    bh = np.random.default_rng(1).normal(scale=0.01, size=(7, 6))
    for combined in (True, False):
        a = po.preintegrate(tr.acc, tr.gyro, tr.dt, 10, bh, combined=combined)
        b = lib.preintegrate(tr.acc, tr.gyro, tr.dt, 10, bh, combined=combined)
        pc.assert_preintegration_parity(a, b, combined)
    # ragged: n_per = 7 does not divide the sample count; n_per = 1 (one keyframe per sample)
    for n_per in (7, 1):
        a = po.preintegrate(tr.acc[:, :300], tr.gyro[:, :300], tr.dt, n_per, bh)
        b = lib.preintegrate(tr.acc[:, :300], tr.gyro[:, :300], tr.dt, n_per, bh)
        pc.assert_preintegration_parity(a, b, True)


@pytest.mark.parametrize("chunks", [1, 4])
def test_lm_iteration_parity_zero_init(po, chunks):
    """reference initial values (identity poses): same accept/reject sequence and errors for the first iterations"""
    _, g, tv, sv = lower_body(4.0, seed=1, near_truth=False)
    o = po.OracleProblem(g, threads=8); o.set_values(tv, sv)
    p = lib.Problem(g, chunks=chunks); p.set_values(tv, sv)
    pc.assert_lm_parity(p, o, abi.LmParams.lower_body(), 8, rel=2e-2)
    p.close()


def test_converged_solution_parity(po):
    """full LM to a fixed iteration budget from the reference's initial values: final cost within 1e-6 relative, IMU
    positions within 1e-6 m, rotations within 1e-3 deg"""
    _, g, tv, sv = lower_body(3.0, seed=2, near_truth=False)
    lm, outer = abi.LmParams.lower_body(), abi.OuterParams.lower_body(max_iterations=400)
    o = po.OracleProblem(g, threads=8); o.set_values(tv, sv)
    ro, ho = o.optimize(lm, outer)
    p = lib.Problem(g, chunks=3); p.set_values(tv, sv)
    rp, hp = p.optimize(lm, outer)
    assert rp.status == ro.status
    assert abs(rp.final_error - ro.final_error) <= 1e-6 * ro.final_error, (rp.final_error, ro.final_error, rp.iterations, ro.iterations)
    tvo, svo = o.get_values(); tvp, svp = p.get_values()
    off = g.time_value_offsets()
    for i in range(7):
        po_, pp_ = tvo[:, off[3 * i] + 9:off[3 * i] + 12], tvp[:, off[3 * i] + 9:off[3 * i] + 12]
        assert np.abs(po_ - pp_).max() < 1e-6
        Ro, Rp = tvo[:, off[3 * i]:off[3 * i] + 9].reshape(-1, 3, 3), tvp[:, off[3 * i]:off[3 * i] + 9].reshape(-1, 3, 3)
        ang = np.degrees(np.linalg.norm(synth.log_so3(np.swapaxes(Ro, 1, 2) @ Rp), axis=-1))
        assert ang.max() < 1e-3
    # north_star: estimated hip / knee angles within 1e-3 deg (derived joint angles, device kernel vs oracle solution),
    # and the kernel against the oracle restatement on identical values (same arithmetic: 1e-12 rad)
    from bioslam_b200 import graphs as bgraphs
    d = bgraphs.lower_body_joint_angle_desc()
    ja_gpu = p.joint_angles(d)
    ja_orc = po.lower_body_joint_angles(g, tvo, svo, d)
    assert ja_gpu.shape == (g.K, 12) and np.all(np.isfinite(ja_gpu))
    assert np.degrees(np.abs(ja_gpu - ja_orc)).max() < 1e-3
    assert np.abs(ja_gpu - po.lower_body_joint_angles(g, tvp, svp, d)).max() < 1e-12
    p.close()


@pytest.mark.parametrize("static_bias", [False, True])
def test_single_imu_parity(po, static_bias):
    tr = synth.make_trial(seed=4, duration_s=10.0)
    K = graphs.n_keyframes(tr.N, 10)
    R0 = tr.poses_at((graphs.keyframe_sample_idx(K, 10) + 1) * tr.dt)["sacrum"][0] @ synth.exp_so3(np.random.default_rng(0).normal(scale=0.05, size=(K, 3)))
    g, tv, sv = graphs.single_imu_graph(tr.acc[0], tr.gyro[0], tr.dt, static_bias=static_bias, init_R=R0, ref_local=(0, 1, 0))
    o = po.OracleProblem(g); o.set_values(tv, sv)
    p = lib.Problem(g, chunks=4); p.set_values(tv, sv)
    pc.assert_linearization_parity(p, o, g)
    pc.assert_solve_parity(p, o, g, lams=(1e-2, 10.0))
    pc.assert_lm_parity(p, o, abi.LmParams.single_imu(), 3, rel=2e-2)
    p.close()


def test_joint_velocity_factor_variant(po):
    tr = synth.make_trial(seed=6, duration_s=3.0)
    g, tv, sv = graphs.lower_body_graph(tr, use_joint_vel=True)
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    o = po.OracleProblem(g, threads=8); o.set_values(tv, sv)
    p = lib.Problem(g, chunks=2); p.set_values(tv, sv)
    pc.assert_linearization_parity(p, o, g)
    p.close()


def test_partition_independence_and_determinism():
    """size-independent properties at a larger size (K = 600): the solve does not depend on the partition beyond
    rounding, and repeated runs are bit-identical (no atomics with more than one contributor)"""
    _, g, tv, sv = lower_body(60.0, seed=5, device_preint=True)
    sols = []
    for chunks in (12, 36, 36):
        p = lib.Problem(g, chunks=chunks); p.set_values(tv, sv)
        p.linearize()
        st, dt, ds = p.solve(1e-2)
        assert st == abi.OK
        sols.append((dt, ds))
        p.close()
    assert np.array_equal(sols[1][0], sols[2][0]) and np.array_equal(sols[1][1], sols[2][1])
    assert np.abs(sols[0][0] - sols[1][0]).max() <= 1e-5 * np.abs(sols[1][0]).max()
    # automatic partition: recursive multi-level tree over the separators, chunk sweeps of the upper levels split over
    # several CTAs -- same solution, bit-reproducible
    auto = []
    for _ in range(2):
        p = lib.Problem(g); p.set_values(tv, sv)
        assert len(p.levels) >= 3, p.levels
        p.linearize()
        st, dt, ds = p.solve(1e-2)
        assert st == abi.OK
        auto.append((dt, ds))
        p.close()
    assert np.array_equal(auto[0][0], auto[1][0]) and np.array_equal(auto[0][1], auto[1][1])
    assert np.abs(auto[0][0] - sols[1][0]).max() <= 1e-5 * np.abs(sols[1][0]).max()


def test_full_size_properties_10min_trial():
    """BASELINE configs[3] at full size (K = 6000, 756 058 unknowns; the oracle would need minutes here): size-independent
    properties only -- the solve does not depend on the partition beyond rounding, repeated solves are bit-identical, the
    reduced cost predicted by the linear model is positive, LM decreases the cost monotonically from the reference's initial
    values and the cost it reports equals the cost recomputed from the final values."""
    _, g, tv, sv = lower_body(600.0, seed=0, near_truth=False, device_preint=True)
    assert g.K == 6000
    sols = []
    for chunks in (None, None, 74):
        p = lib.Problem(g, chunks=chunks); p.set_values(tv, sv)
        if chunks is None:
            assert p.chunks == 148 and len(p.levels) >= 3
        p.linearize()
        st, dt, ds = p.solve(1e-2)
        assert st == abi.OK
        sols.append((dt, ds))
        p.close()
    assert np.array_equal(sols[0][0], sols[1][0]) and np.array_equal(sols[0][1], sols[1][1])
    assert np.abs(sols[0][0] - sols[2][0]).max() <= 1e-5 * np.abs(sols[0][0]).max()
    assert np.abs(sols[0][1] - sols[2][1]).max() <= 1e-5 * max(np.abs(sols[0][1]).max(), 1e-300)
    p = lib.Problem(g); p.set_values(tv, sv)
    e_prev = p.lm_begin(abi.LmParams.lower_body())
    for _ in range(4):
        st = p.lm_iterate()
        lin0, lin1 = p.last_trial_scalars()[:2]
        assert lin1 <= lin0                                   # the accepted step reduces the linearised cost
        assert st.error < e_prev
        e_prev = st.error
    assert abs(p.error() - st.error) <= 1e-12 * st.error
    p.close()


def test_lm_monotone_and_matches_error_recompute():
    _, g, tv, sv = lower_body(60.0, seed=5, near_truth=False, device_preint=True)
    p = lib.Problem(g); p.set_values(tv, sv)
    e_prev = p.lm_begin(abi.LmParams.lower_body())
    for _ in range(6):
        st = p.lm_iterate()
        assert st.error <= e_prev
        e_prev = st.error
    assert abs(p.error() - st.error) <= 1e-12 * st.error
    tv2, sv2 = p.get_values()
    # rotations stay orthonormal, Unit3 statics stay unit
    R = tv2[:, :9].reshape(-1, 3, 3)
    assert np.abs(R @ np.swapaxes(R, 1, 2) - np.eye(3)).max() < 1e-9
    axes = sv2[36:].reshape(-1, 3)
    assert np.abs(np.linalg.norm(axes, axis=1) - 1).max() < 1e-12
    p.close()


def test_indeterminate_and_bad_args():
    g = abi.GraphDesc(3, [abi.VAR_VEC3, abi.VAR_VEC3], [])
    g.add_slot(abi.F_PRIOR_VEC3, [(abi.AT_K, 0)], [1, 1, 1, 0, 0, 0])
    p = lib.Problem(g, chunks=1)
    p.set_values(np.ones((3, 6)), np.zeros(0))
    p.linearize()
    assert p.solve(0.0)[0] == abi.INDETERMINATE_SYSTEM
    st, dt, _ = p.solve(1.0)
    assert st == abi.OK and np.allclose(dt[:, :3], -0.5)
    p.close()
    g.slots[0].k_end = 9
    with pytest.raises(lib.B200Error):
        lib.Problem(g)


def test_every_factor_type_on_device_matches_fixtures(po):
    """all 21 factor types (incl. the non-default HingeJointConstraintNormErrEstAngVel, ConstrainedJointCenterNormPositionFactor,
    MagPose3Factor and PriorFactor<Pose3>) evaluated by the device functions of linearize_kernel against tests/golden/factors.npz
    (residual 1e-13, Jacobian 1e-12) and, whitened, against the oracle on the same inputs"""
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "factors.npz"))
    assert len(abi.FACTOR_INFO) == 21
    for ft in sorted(abi.FACTOR_INFO):
        params, consts, x = d[f"f{ft}_params"], d[f"f{ft}_consts"], d[f"f{ft}_x"]
        r, J = lib.factor_eval(ft, params, consts, x)
        assert np.allclose(r, d[f"f{ft}_r"], rtol=1e-13, atol=1e-13), ft
        assert np.allclose(J, d[f"f{ft}_J"], rtol=1e-12, atol=1e-12), ft
        rw, Jw = lib.factor_eval(ft, params, consts, x, whiten=True)
        ro, Jo = po.factor_eval(ft, params, consts, x, whiten=True)
        assert np.allclose(rw, ro, rtol=1e-11, atol=1e-11 * max(1.0, np.abs(ro).max())), ft
        assert np.allclose(Jw, Jo, rtol=1e-11, atol=1e-11 * max(1.0, np.abs(Jo).max())), ft


@pytest.mark.parametrize("variant", ["hinge_norm", "jointpos_norm", "magnetometer", "jointvel_norm", "all"])
def test_non_default_factor_variants_in_the_graph(po, variant):
    """m_hingeAxisFactorDim = 1, m_jointPosFactorDim = 1, m_useMagnetometer = true (reference src/lowerBodyPoseEstimator.cpp:170,218,
    248,270; src/imuPoseEstimator.cpp:216-219): linearization, solve and LM parity with the oracle on the CUDA path"""
    kw = {"hinge_norm": dict(hinge_dim=1), "jointpos_norm": dict(joint_pos_dim=1), "magnetometer": dict(magnetometer=True),
          "jointvel_norm": dict(use_joint_vel=True, joint_vel_dim=1),
          "all": dict(hinge_dim=1, joint_pos_dim=1, magnetometer=True, use_joint_vel=True, joint_vel_dim=1)}[variant]
    tr = synth.make_trial(seed=8, duration_s=4.0)
    g, tv, sv = graphs.lower_body_graph(tr, **kw)
    types = {s.type for s in g.slots}
    if "hinge_dim" in kw: assert abi.F_HINGE_NORM in types and abi.F_HINGE_VEC not in types
    if "joint_pos_dim" in kw: assert abi.F_JOINTPOS_NORM in types and abi.F_JOINTPOS_VEC not in types
    if "magnetometer" in kw: assert abi.F_MAG_POSE3 in types
    if "joint_vel_dim" in kw: assert abi.F_JOINTVEL_NORM in types and abi.F_JOINTVEL_VEC not in types
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    o = po.OracleProblem(g, threads=8); o.set_values(tv, sv)
    p = lib.Problem(g, chunks=3); p.set_values(tv, sv)
    pc.assert_linearization_parity(p, o, g)
    pc.assert_solve_parity(p, o, g, lams=(1e-3, 1.0))
    o.set_values(tv, sv); p.set_values(tv, sv)
    pc.assert_lm_parity(p, o, abi.LmParams.lower_body(), 4, rel=2e-2)
    p.close()


def test_pose3_prior_factor_in_a_graph(po):
    """gtsam::PriorFactor<Pose3> (A11) on the first pose of a single-IMU chain"""
    tr = synth.make_trial(seed=4, duration_s=3.0)
    K = graphs.n_keyframes(tr.N, 10)
    R0 = tr.poses_at((graphs.keyframe_sample_idx(K, 10) + 1) * tr.dt)["sacrum"][0] @ synth.exp_so3(np.random.default_rng(0).normal(scale=0.05, size=(K, 3)))
    g, tv, sv = graphs.single_imu_graph(tr.acc[0], tr.gyro[0], tr.dt, init_R=R0, ref_local=(0, 1, 0))
    prior = np.concatenate([synth.exp_so3(np.array([[0.1, -0.2, 0.3]]))[0].ravel(), [0.01, 0.02, -0.03]])
    g.add_slot(abi.F_PRIOR_POSE3, [(abi.AT_K, 0)], [0.1, 0.1, 0.1, 0.05, 0.05, 0.05] + list(prior), 0, 1)
    o = po.OracleProblem(g); o.set_values(tv, sv)
    p = lib.Problem(g, chunks=2); p.set_values(tv, sv)
    pc.assert_linearization_parity(p, o, g)
    pc.assert_lm_parity(p, o, abi.LmParams.single_imu(), 3, rel=2e-2)
    p.close()


def test_apdm_rate_128hz_trial(po):
    """BASELINE configs[1] stand-in at the sensor's own rate (APDM Opal: 128 Hz; the reference's test trial is a ~30 s TUG): 128 Hz x
    6.05 s = 774 samples, not a multiple of the 10-sample preintegration -> floor(774 / 10) + 1 = 78 keyframes
    (getExpectedNumberOfKeyframesFromImuData, src/imuPoseEstimator.cpp:117-125) -- preintegration on the device, linearization, damped
    solves and the first LM iterations against the oracle."""
    tr = synth.make_trial(seed=6, duration_s=6.05, rate_hz=128.0)
    assert abs(tr.dt - 1.0 / 128.0) < 1e-15 and tr.acc.shape[1] % 10 != 0
    g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh))
    go, _, _ = graphs.lower_body_graph(tr)                                  # the same graph with the oracle's preintegration
    assert g.K == go.K
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    o = po.OracleProblem(go, threads=8); o.set_values(tv, sv)
    p = lib.Problem(g, chunks=None); p.set_values(tv, sv)
    pc.assert_linearization_parity(p, o, g)
    pc.assert_solve_parity(p, o, g, lams=(1e-3, 1.0))
    pc.assert_lm_parity(p, o, abi.LmParams.lower_body(), 4, rel=2e-2)
    p.close()
