"""-m "not gpu": the oracle's linear algebra, preintegration and LM loop against independent numpy/scipy computations."""
import numpy as np
import pytest
import scipy.linalg as sl
from bioslam_b200 import abi, synth
from oracle import pyoracle as po
import graphs


def small_problem(seconds=1.2, seed=3):
    tr = synth.make_trial(seed=seed, duration_s=seconds)
    g, tv, sv = graphs.lower_body_graph(tr)
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    return tr, g, tv, sv


def dense_system(o, g):
    K, nt, ns = g.K, g.nt, g.ns
    n = K * nt + ns
    H = np.zeros((n, n)); gg = np.zeros(n)
    gt, gs = o.gradient()
    for k in range(K):
        D, O, A = o.hessian_blocks(k)
        H[k * nt:(k + 1) * nt, k * nt:(k + 1) * nt] = D
        if k + 1 < K:
            H[k * nt:(k + 1) * nt, (k + 1) * nt:(k + 2) * nt] = O
            H[(k + 1) * nt:(k + 2) * nt, k * nt:(k + 1) * nt] = O.T
        H[k * nt:(k + 1) * nt, K * nt:] = A
        H[K * nt:, k * nt:(k + 1) * nt] = A.T
    H[K * nt:, K * nt:] = o.static_hessian()
    gg[:K * nt] = gt.ravel(); gg[K * nt:] = gs
    return H, gg


def test_block_arrow_solver_matches_dense_cholesky(oracle_lib):
    _, g, tv, sv = small_problem()
    o = po.OracleProblem(g); o.set_values(tv, sv); o.linearize()
    H, gg = dense_system(o, g)
    assert np.abs(H - H.T).max() < 1e-6 * np.abs(H).max()
    for lam in (1e-3, 1.0, 1e3):
        st, dt, ds = o.solve(lam)
        assert st == 0
        x = np.concatenate([dt.ravel(), ds])
        Hl = H + lam * np.eye(H.shape[0])
        assert np.linalg.norm(Hl @ x + gg) <= 1e-8 * np.linalg.norm(gg)


def test_error_is_half_squared_whitened_residual(oracle_lib):
    _, g, tv, sv = small_problem()
    o = po.OracleProblem(g); o.set_values(tv, sv)
    e = o.error()
    # gradient finite difference along a random direction
    o.linearize(); gt, gs = o.gradient()
    rng = np.random.default_rng(0)
    d_t = rng.normal(size=gt.shape) * 1e-7; d_s = rng.normal(size=gs.shape) * 1e-7
    o.retract(d_t, d_s); e1 = o.error()
    o.set_values(tv, sv); o.retract(-d_t, -d_s); e2 = o.error()
    pred = 2 * (np.sum(gt * d_t) + np.sum(gs * d_s))
    assert abs((e1 - e2) - pred) <= 1e-4 * abs(pred) + 1e-9 * e


def test_preintegration_mean_and_bias_jacobians(oracle_lib):
    tr = synth.make_trial(seed=2, duration_s=1.0)
    acc, gyro = tr.acc[:1], tr.gyro[:1]
    bh = np.zeros((1, 6))
    rec = po.preintegrate(acc, gyro, tr.dt, 10, bh)[0]           # [K-1][190]
    # independent Euler integration on the manifold for comparison of orders of magnitude, and exact tangent recursion
    def tangent_recursion(a, w, dt):
        th = np.zeros(3); p = np.zeros(3); v = np.zeros(3)
        for j in range(10):
            R = synth.exp_so3(th)
            # inverse right Jacobian applied to w
            t = np.linalg.norm(th)
            W = np.array([[0, -th[2], th[1]], [th[2], 0, -th[0]], [-th[1], th[0], 0]])
            Jr = np.eye(3) if t < 1e-8 else np.eye(3) - (1 - np.cos(t)) / t**2 * W + (t - np.sin(t)) / t**3 * W @ W
            wt = np.linalg.solve(Jr, w[j])
            an = R @ a[j]
            th, p, v = th + wt * dt, p + v * dt + 0.5 * an * dt * dt, v + an * dt
        return np.concatenate([th, p, v])
    for q in (0, 3):
        z = tangent_recursion(acc[0, 10 * q:10 * q + 10], gyro[0, 10 * q:10 * q + 10], tr.dt)
        assert np.abs(rec[q, :9] - z).max() < 1e-12
        assert abs(rec[q, 69] - 10 * tr.dt) < 1e-15
    # bias Jacobians by finite differences of the recursion wrt bias_hat
    h = 1e-6
    for j in range(6):
        bp = bh.copy(); bp[0, j] += h; bm = bh.copy(); bm[0, j] -= h
        d = (po.preintegrate(acc, gyro, tr.dt, 10, bp)[0][2, :9] - po.preintegrate(acc, gyro, tr.dt, 10, bm)[0][2, :9]) / (2 * h)
        H = rec[2, 9:36].reshape(9, 3) if j < 3 else rec[2, 36:63].reshape(9, 3)
        assert np.abs(H[:, j % 3] - d).max() < 1e-7


def test_sqrt_information_is_upper_and_whitens_to_unit_covariance(oracle_lib):
    tr = synth.make_trial(seed=2, duration_s=0.5)
    rec = po.preintegrate(tr.acc[:1], tr.gyro[:1], tr.dt, 10, np.zeros((1, 6)))[0][0]
    R = np.zeros((15, 15)); off = 70
    for r in range(15):
        R[r, r:] = rec[off:off + 15 - r]; off += 15 - r
    assert np.all(np.diag(R) > 0)
    info = R.T @ R
    assert np.all(np.linalg.eigvalsh(info) > 0)
    # static-bias twin: 9x9
    rec9 = po.preintegrate(tr.acc[:1], tr.gyro[:1], tr.dt, 10, np.zeros((1, 6)), combined=False)[0][0]
    assert rec9.size == 115 and np.abs(rec9[:9] - rec[:9]).max() == 0.0


def test_lm_converges_and_matches_reference_exit_rule(oracle_lib):
    _, g, tv, sv = small_problem(seconds=1.0)
    o = po.OracleProblem(g); o.set_values(tv, sv)
    res, hist = o.optimize(abi.LmParams.lower_body(), abi.OuterParams.lower_body(max_iterations=40))
    assert res.final_error < 1e-3 * res.initial_error
    assert np.all(np.diff(hist) <= 1e-12 * hist[:-1])  # LM never accepts an increase
    assert res.n_history == res.outer_calls + 1


def test_gauss_newton_mode(oracle_lib):
    _, g, tv, sv = small_problem(seconds=0.8)
    o = po.OracleProblem(g); o.set_values(tv, sv)
    lm = abi.LmParams.lower_body(); lm.gauss_newton = 1
    o.lm_begin(lm)
    st = o.lm_iterate()
    # lambda == 0, exactly one solve, the step is taken unconditionally (matlab/src/lowerBodyPoseEstimator.m:1108-1110)
    assert st.lambda_ == 0.0 and st.iterations == 1 and st.inner_trials == 1 and abs(st.error - o.error()) <= 1e-12 * st.error


def test_fast_solver_matches_plain_restatement(oracle_lib):
    """oracle/fastsolve.hpp (local dofs pre-eliminated, chunked, blocked AVX2 kernels: the CPU baseline solver) against the plain
    dense block sweep on the same damped systems: same normal-equation residual class, for 1..7 chunks, any thread count"""
    import parity_checks as pc
    _, g, tv, sv = small_problem(seconds=3.0)
    o = po.OracleProblem(g, threads=4); o.set_values(tv, sv); o.linearize()
    for lam in (1e-5, 1e-3, 1.0, 1e3):
        o.set_solver(False)
        st, dt0, ds0 = o.solve(lam)
        r0 = pc.normal_eq_residual(o, g, lam, dt0, ds0)
        for chunks in (1, 2, 3, 7):
            o.set_solver(True, chunks)
            st, dt, ds = o.solve(lam)
            assert st == 0
            assert pc.normal_eq_residual(o, g, lam, dt, ds) <= max(10 * r0, 1e-9), (lam, chunks)
            if lam >= 1.0:
                assert np.abs(dt - dt0).max() <= 1e-5 * np.abs(dt0).max()
    # thread count does not change a single bit
    o.set_solver(True, 3); _, dta, dsa = o.solve(1e-3)
    o2 = po.OracleProblem(g, threads=1); o2.set_values(tv, sv); o2.linearize(); o2.set_solver(True, 3)
    _, dtb, dsb = o2.solve(1e-3)
    assert np.array_equal(dta, dtb) and np.array_equal(dsa, dsb)


@pytest.mark.parametrize("static_bias", [False, True])
def test_fast_solver_single_imu(oracle_lib, static_bias):
    """no keyframe-local dofs, no statics (dynamic bias) / a 6-dof static arrow (static bias)"""
    import parity_checks as pc
    tr = synth.make_trial(seed=4, duration_s=4.0)
    K = graphs.n_keyframes(tr.N, 10)
    R0 = tr.poses_at((graphs.keyframe_sample_idx(K, 10) + 1) * tr.dt)["sacrum"][0]
    g, tv, sv = graphs.single_imu_graph(tr.acc[0], tr.gyro[0], tr.dt, static_bias=static_bias, init_R=R0, ref_local=(0, 1, 0))
    o = po.OracleProblem(g); o.set_values(tv, sv); o.linearize()
    for lam in (1e-2, 10.0):
        o.set_solver(False); _, dt0, ds0 = o.solve(lam)
        r0 = pc.normal_eq_residual(o, g, lam, dt0, ds0)
        for chunks in (1, 4):
            o.set_solver(True, chunks); st, dt, ds = o.solve(lam)
            assert st == 0 and pc.normal_eq_residual(o, g, lam, dt, ds) <= max(10 * r0, 1e-9)


def test_fast_solver_reports_indefinite_system(oracle_lib):
    _, g, tv, sv = small_problem(seconds=1.0)
    o = po.OracleProblem(g); o.set_values(tv, sv); o.linearize()
    o.set_solver(True, 2)
    st, _, _ = o.solve(-1e12)
    assert st == abi.INDETERMINATE_SYSTEM
