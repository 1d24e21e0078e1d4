"""-m "not gpu": the CUDA sources executed under the CUDA-thread emulator (tests/cudaemu, TEST TOOL ONLY) against the
oracle.  Catches indexing / barrier / host-control-flow bugs in the build container, which has no GPU; numerical parity
on real hardware is tests/test_gpu_parity.py."""
import os
import subprocess
import numpy as np
import pytest
from bioslam_b200 import abi, lib, synth
from oracle import pyoracle as po
import graphs
import parity_checks as pc

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "cudaemu", "libbioslam_b200_emu.so")


@pytest.fixture(scope="module")
def emu(oracle_lib):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cudaemu")])
    return EMU


def lower_body(seconds, seed=3):
    tr = synth.make_trial(seed=seed, duration_s=seconds)
    g, tv, sv = graphs.lower_body_graph(tr)
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    return tr, g, tv, sv


@pytest.mark.parametrize("chunks", [1, 3])
def test_lower_body_linearization_and_solve(emu, chunks):
    _, g, tv, sv = lower_body(1.2)
    o = po.OracleProblem(g); o.set_values(tv, sv)
    p = lib.Problem(g, so_path=emu, chunks=chunks); p.set_values(tv, sv)
    assert p.chunks == chunks
    tv2, sv2 = p.get_values()
    assert np.array_equal(tv2, tv) and np.array_equal(sv2, sv)
    pc.assert_linearization_parity(p, o, g)
    pc.assert_solve_parity(p, o, g)


def test_lower_body_lm_iterations(emu):
    _, g, tv, sv = lower_body(1.0)
    o = po.OracleProblem(g); o.set_values(tv, sv)
    p = lib.Problem(g, so_path=emu, chunks=2); p.set_values(tv, sv)
    # lambda <= 1e-5 in these first iterations: the damped systems have condition numbers ~1e16, two correct solvers agree to ~1e-2 in cost
    pc.assert_lm_parity(p, o, abi.LmParams.lower_body(), 3, rel=2e-2)
    assert p.launch_count() > 10


@pytest.mark.parametrize("static_bias", [False, True])
def test_single_imu(emu, static_bias):
    tr = synth.make_trial(seed=4, duration_s=2.0)
    K = graphs.n_keyframes(tr.N, 10)
    R0 = tr.poses_at((graphs.keyframe_sample_idx(K, 10) + 1) * tr.dt)["sacrum"][0] @ synth.exp_so3(np.random.default_rng(0).normal(scale=0.05, size=(K, 3)))
    g, tv, sv = graphs.single_imu_graph(tr.acc[0], tr.gyro[0], tr.dt, static_bias=static_bias, init_R=R0, ref_local=(0, 1, 0))
    o = po.OracleProblem(g); o.set_values(tv, sv)
    p = lib.Problem(g, so_path=emu, chunks=2); p.set_values(tv, sv)
    pc.assert_linearization_parity(p, o, g)
    pc.assert_solve_parity(p, o, g, lams=(1e-2, 10.0))
    pc.assert_lm_parity(p, o, abi.LmParams.single_imu(), 2, rel=2e-2)


def test_preintegration_kernel(emu):
    tr = synth.make_trial(seed=9, duration_s=1.0)
    bh = np.random.default_rng(1).normal(scale=0.01, size=(7, 6))
    for combined in (True, False):
        a = po.preintegrate(tr.acc, tr.gyro, tr.dt, 10, bh, combined=combined)
        b = lib.preintegrate(tr.acc, tr.gyro, tr.dt, 10, bh, combined=combined, so_path=emu)
        pc.assert_preintegration_parity(a, b, combined)


def test_indeterminate_system_is_reported(emu):
    """a pose with no factor attached makes the undamped system singular: status, not a crash
    (== gtsam::IndeterminantLinearSystemException, reference src/imuPoseEstimator.cpp:360)"""
    g = abi.GraphDesc(3, [abi.VAR_VEC3, abi.VAR_VEC3], [])
    g.add_slot(abi.F_PRIOR_VEC3, [(abi.AT_K, 0)], [1, 1, 1, 0, 0, 0])
    p = lib.Problem(g, so_path=emu, chunks=1)
    p.set_values(np.ones((3, 6)), np.zeros(0))
    p.linearize()
    st, _, _ = p.solve(0.0)
    assert st == abi.INDETERMINATE_SYSTEM
    st, dt, _ = p.solve(1.0)
    assert st == abi.OK and np.allclose(dt[:, :3], -0.5) and np.allclose(dt[:, 3:], 0.0)


def test_bad_arguments_are_rejected(emu):
    g = abi.GraphDesc(2, [abi.VAR_POSE3], [])
    g.add_slot(abi.F_POSE3_TRANSLATION_PRIOR, [(abi.AT_K, 0)], [1, 1, 1, 0, 0, 0])
    g.slots[0].k_end = 5
    with pytest.raises(lib.B200Error) as e:
        lib.Problem(g, so_path=emu)
    assert e.value.status == abi.BAD_ARG


def test_every_factor_type_device_functions(emu):
    """the device residual / Jacobian functions of all 20 factor types (emulated) against tests/golden/factors.npz"""
    d = np.load(os.path.join(HERE, "golden", "factors.npz"))
    for ft in sorted(abi.FACTOR_INFO):
        r, J = lib.factor_eval(ft, d[f"f{ft}_params"], d[f"f{ft}_consts"], d[f"f{ft}_x"], so_path=emu)
        assert np.allclose(r, d[f"f{ft}_r"], rtol=1e-13, atol=1e-13), ft
        assert np.allclose(J, d[f"f{ft}_J"], rtol=1e-12, atol=1e-12), ft
        rw, Jw = lib.factor_eval(ft, d[f"f{ft}_params"], d[f"f{ft}_consts"], d[f"f{ft}_x"], whiten=True, so_path=emu)
        ro, Jo = po.factor_eval(ft, d[f"f{ft}_params"], d[f"f{ft}_consts"], d[f"f{ft}_x"], whiten=True)
        assert np.allclose(rw, ro, rtol=1e-11, atol=1e-11 * max(1.0, np.abs(ro).max())), ft
        assert np.allclose(Jw, Jo, rtol=1e-11, atol=1e-11 * max(1.0, np.abs(Jo).max())), ft


def test_non_default_factor_variants(emu):
    """norm-error hinge / joint-centre position / joint-centre velocity factors and the magnetometer factor inside the lower-body graph"""
    tr = synth.make_trial(seed=8, duration_s=1.0)
    g, tv, sv = graphs.lower_body_graph(tr, hinge_dim=1, joint_pos_dim=1, magnetometer=True, use_joint_vel=True, joint_vel_dim=1)
    assert abi.F_JOINTVEL_NORM in {s.type for s in g.slots}
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    o = po.OracleProblem(g); o.set_values(tv, sv)
    p = lib.Problem(g, so_path=emu, chunks=2); p.set_values(tv, sv)
    pc.assert_linearization_parity(p, o, g)
    pc.assert_solve_parity(p, o, g)


def test_degenerate_geometry_on_device_functions(emu):
    """same guard as gtsam::norm3 on the device: finite Jacobians at parallel / anti-parallel / coincident configurations,
    equal to the oracle's"""
    eye = np.concatenate([np.eye(3).ravel(), np.zeros(3)])
    half_turn = np.concatenate([np.diag([-1.0, -1.0, 1.0]).ravel(), np.zeros(3)])
    cases = [(abi.F_POSE3_COMPASS_PRIOR, [1e-3, 0.0, 1, 0, 0, 1, 0, 0, 0, 0, 1], eye),
             (abi.F_POSE3_COMPASS_PRIOR, [1e-3, 0.0, 1, 0, 0, 1, 0, 0, 0, 0, 1], half_turn),
             (abi.F_SEGLEN_MAG, [0.01, 0.4], np.array([0.1, 0.2, 0.3, 0.1, 0.2, 0.3])),
             (abi.F_SEGLEN_DISCREPANCY, [0.002], np.array([0.1, 0.2, 0.3, 0.1, 0.2, 0.3, 0, 0, 0, 0, 0, 1.0]))]
    for ft, prm, x in cases:
        r, J = lib.factor_eval(ft, prm, None, x, so_path=emu)
        ro, Jo = po.factor_eval(ft, prm, None, x)
        assert np.all(np.isfinite(J)) and np.allclose(J, Jo, atol=1e-14) and np.allclose(r, ro, atol=1e-14)


def test_local_dof_elimination_is_exact_to_rounding(emu):
    """kernels_local.cuh: the angular velocities eliminated per keyframe before the chain sweep (Schur complement against mpmath)"""
    _, g, tv, sv = lower_body(1.0)
    p = lib.Problem(g, so_path=emu, chunks=2); p.set_values(tv, sv); p.linearize()
    assert pc.assert_local_elimination_parity(p, g, 1e-3, [0, 4, g.K - 1])
    # a graph without keyframe-local dofs has nothing to pre-eliminate
    tr = synth.make_trial(seed=4, duration_s=1.0)
    g1, tv1, sv1 = graphs.single_imu_graph(tr.acc[0], tr.gyro[0], tr.dt)
    p1 = lib.Problem(g1, so_path=emu, chunks=1); p1.set_values(tv1, sv1); p1.linearize()
    assert p1.local_elim(0, 1.0) is None


def test_backsub_without_staged_coupling_block(emu, monkeypatch):
    """kernels_solve.cuh cta_backsub: the coupling block of the node in flight is staged in shared memory when it fits next to the
    factor; a block too large for that is read from global memory.  Both paths give the same step, bit for bit."""
    _, g, tv, sv = lower_body(1.2)
    steps = []
    for off in (False, True):
        if off: monkeypatch.setenv("B200_NO_BS_OSTAGE", "1")
        p = lib.Problem(g, so_path=emu, chunks=3); p.set_values(tv, sv); p.linearize()
        steps.append(p.solve(1e-3))
    assert all(np.array_equal(a, b) for a, b in zip(steps[0], steps[1]))
    o = po.OracleProblem(g); o.set_values(tv, sv)
    p = lib.Problem(g, so_path=emu, chunks=3); p.set_values(tv, sv)
    pc.assert_linearization_parity(p, o, g)
    pc.assert_solve_parity(p, o, g)   # (the global-memory path against the oracle)


def test_schur_tile_kernel_against_the_chunk_kernel(emu, monkeypatch):
    """kernels_solve.cuh: on the upper levels the arrow Schur complement is computed one 32 x 32 tile per CTA (contraction split over
    four warps); B200_NO_SCHUR_TILE forces the level-0 kernel (whole panels, one warp per tile) everywhere.  Same complement up to
    the summation order: the steps agree far below the solver's own residual."""
    _, g, tv, sv = lower_body(1.6)
    steps = []
    for off in (False, True):
        if off: monkeypatch.setenv("B200_NO_SCHUR_TILE", "1")
        p = lib.Problem(g, so_path=emu, chunks=4); p.set_values(tv, sv); p.linearize()
        steps.append(p.solve(1.0))
    (st0, dt0, ds0), (st1, dt1, ds1) = steps
    assert st0 == 0 and st1 == 0
    assert np.abs(dt0 - dt1).max() <= 1e-9 * np.abs(dt0).max() and np.abs(ds0 - ds1).max() <= 1e-9 * np.abs(ds0).max()
    assert not (np.array_equal(dt0, dt1) and np.array_equal(ds0, ds1)), "the switch did not change the kernel"


def test_apdm_rate_128hz_trial(emu):
    """the sensor's own rate (APDM Opal, 128 Hz) with a sample count that is not a multiple of the preintegration length:
    154 samples -> floor(154 / 10) + 1 = 16 keyframes (src/imuPoseEstimator.cpp:117-125); device preintegration kernel, linearization
    and a damped solve against the oracle (the GPU twin of this test runs 6 s and LM iterations)."""
    tr = synth.make_trial(seed=6, duration_s=1.2, rate_hz=128.0)
    assert tr.acc.shape[1] == 154
    g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh, so_path=emu))
    go, _, _ = graphs.lower_body_graph(tr)
    assert g.K == go.K == 16
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    o = po.OracleProblem(go); o.set_values(tv, sv)
    p = lib.Problem(g, so_path=emu, chunks=2); p.set_values(tv, sv)
    pc.assert_linearization_parity(p, o, g)
    pc.assert_solve_parity(p, o, g)
