"""-m "not gpu": closed-form / first-principles pins of the GTSAM-OWNED arithmetic that the oracle restates from SURVEY.md Appendix A
(GTSAM 4.2 is an un-vendored dependency, absent from this image and from /root/reference: the oracle cannot be run against it here).
None of these tests can prove bit-parity with GTSAM; they pin what the published algorithms must satisfy whatever the code path:

  * tangent preintegration (PreintegratedCombinedMeasurements): constant specific force -> dp = a t^2 / 2, dv = a t; constant body rate
    -> dR = Exp(w t); the manifold recursion R_j+1 = R_j Exp(w dt), v += R_j a dt, p += v dt + R_j a dt^2 / 2 (the non-tangent GTSAM
    variant, same mean by construction);
  * preintegrated measurement covariance: the discrete white-noise model propagated through the recursion, against a seeded Monte Carlo;
  * Pose3::Expmap (GTSAM_POSE3_EXPMAP = ON) against the matrix exponential of the twist; Pose3 retract = T * Expmap(xi);
  * Unit3 retract: unit norm, cos|v| along p, tangent basis orthonormal and orthogonal to p, local o retract = id to first order;
  * PriorFactor / tryLambda on a one-variable quadratic: the LM step is x + (H + lambda I)^-1 H (mu - x), lambda <- lambda / 10 on accept,
    lambda <- 10 lambda with the SAME linearization on reject (a cost whose quadratic model overshoots).
"""
import numpy as np
import scipy.linalg as sl
from bioslam_b200 import abi, synth
from oracle import pyoracle as po


def _rec_cov(rec, dim=15):
    """covariance from the packed upper-triangular sqrt-information R of a preintegration record (Sigma^-1 = R^T R)"""
    R = np.zeros((dim, dim)); off = 70
    for r in range(dim):
        R[r, r:] = rec[off:off + dim - r]; off += dim - r
    return np.linalg.inv(R.T @ R)


def test_constant_specific_force_and_constant_rate(oracle_lib):
    dt, n = 0.01, 10
    t = n * dt
    a = np.array([0.3, -1.2, 9.7]); w = np.array([0.4, -0.25, 0.6])
    z = np.zeros((1, 2 * n, 3))
    # pure translation
    rec = po.preintegrate(z + a, z, dt, n, np.zeros((1, 6)))[0][0]
    assert np.abs(rec[0:3]).max() == 0.0
    assert np.abs(rec[3:6] - 0.5 * a * t * t).max() < 1e-15 and np.abs(rec[6:9] - a * t).max() < 1e-15
    assert abs(rec[69] - t) < 1e-16
    # pure rotation about a fixed axis: theta = w t exactly (Log o Exp on one axis), no translation
    rec = po.preintegrate(z, z + w, dt, n, np.zeros((1, 6)))[0][0]
    assert np.abs(rec[0:3] - w * t).max() < 1e-14 and np.abs(rec[3:9]).max() == 0.0
    # both: the manifold recursion (independent numpy code)
    rec = po.preintegrate(z + a, z + w, dt, n, np.zeros((1, 6)))[0][0]
    R = np.eye(3); p = np.zeros(3); v = np.zeros(3)
    for _ in range(n):
        an = R @ a
        p = p + v * dt + 0.5 * an * dt * dt
        v = v + an * dt
        R = R @ sl.expm(np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]) * dt)
    assert np.abs(synth.exp_so3(rec[0:3]) - R).max() < 1e-13
    assert np.abs(rec[3:6] - p).max() < 1e-14 and np.abs(rec[6:9] - v).max() < 1e-14
    # bias_hat is subtracted from the raw samples before integration
    bh = np.array([[0.05, -0.02, 0.01, 0.003, -0.001, 0.002]])
    rec_b = po.preintegrate(z + a + bh[0, :3], z + w + bh[0, 3:], dt, n, bh)[0][0]
    assert np.abs(rec_b[:9] - rec[:9]).max() < 1e-13 and np.abs(rec_b[63:69] - bh[0]).max() == 0.0


def test_preintegrated_covariance_against_monte_carlo(oracle_lib):
    """[theta, p, v] block of preintMeasCov against 40 000 seeded noisy replicas of one interval pushed through the recursion.
    GTSAM 4.2 (CombinedImuFactor.cpp, integrateMeasurement) injects per sample (aCov + biasAccInt) / dt on the velocity and
    (wCov + biasOmegaInt) / dt on the rotation -- the `biasAccOmegaInt` uncertainty of the bias used for preintegration rides as
    white measurement noise -- and NOTHING directly on the position except the integration-error term dt * iCov (block-diagonal
    G Sigma G^T): theta and v must match first principles, p comes out a few per cent BELOW them (the a dt^2 / 2 injection and its
    cross term with v are left out upstream; the oracle follows upstream)."""
    rng = np.random.default_rng(7)
    dt, n, N = 0.01, 10, 40000
    noise = abi.ImuNoise.opal_v2()
    b_int = noise.preint_bias_sigma ** 2
    a = np.array([0.8, -0.5, 9.6]); w = np.array([0.9, -0.4, 0.7])
    acc = a + rng.normal(size=(N, n, 3)) * np.sqrt((noise.accel_noise_sigma ** 2 + b_int) / dt)
    gyr = w + rng.normal(size=(N, n, 3)) * np.sqrt((noise.gyro_noise_sigma ** 2 + b_int) / dt)
    # the replicas ride as N "IMUs" with two intervals each (the second interval is padding)
    pad = np.zeros((N, n, 3))
    recs = po.preintegrate(np.concatenate([acc, pad + a], 1), np.concatenate([gyr, pad + w], 1), dt, n, np.zeros((N, 6)), noise=noise)[:, 0, :9]
    mc = np.cov(recs.T)
    clean = po.preintegrate(np.tile(a, (1, 2 * n, 1)), np.tile(w, (1, 2 * n, 1)), dt, n, np.zeros((1, 6)), noise=noise)[0][0]
    full = _rec_cov(clean)
    cov = full[:9, :9]
    t = n * dt
    sd_mc, sd = np.sqrt(np.diag(mc)), np.sqrt(np.diag(cov))
    # a standard deviation estimated from N replicas scatters by 1 / sqrt(2 N) = 0.35 %
    assert np.abs(sd[0:3] / sd_mc[0:3] - 1.0).max() < 0.02, sd[0:3] / sd_mc[0:3]      # rotation
    assert np.abs(sd[6:9] / sd_mc[6:9] - 1.0).max() < 0.02, sd[6:9] / sd_mc[6:9]      # velocity
    var_p_model = np.diag(cov)[3:6] - noise.integration_err_sigma ** 2 * t             # (the integration error is not in the samples)
    assert np.all(var_p_model < np.diag(mc)[3:6]) and np.all(var_p_model > 0.80 * np.diag(mc)[3:6]), var_p_model / np.diag(mc)[3:6]
    # correlations of the dominant pair (p, v) per axis
    for i in range(3):
        c_mc = mc[3 + i, 6 + i] / (sd_mc[3 + i] * sd_mc[6 + i]); c = cov[3 + i, 6 + i] / (sd[3 + i] * sd[6 + i])
        assert abs(c - c_mc) < 0.08, (c, c_mc)
    # closed forms of the diagonal: var(theta) = (sg^2 + b_int) t for a small rotation, var(v) = (sa^2 + b_int) t
    assert np.abs(np.diag(cov)[0:3] / ((noise.gyro_noise_sigma ** 2 + b_int) * t) - 1.0).max() < 2e-2
    assert np.abs(np.diag(cov)[6:9] / ((noise.accel_noise_sigma ** 2 + b_int) * t) - 1.0).max() < 2e-2
    # bias block of the combined covariance: random walk, sigma_b^2 t per axis, uncorrelated with the rest
    assert np.abs(np.diag(full)[9:12] / (noise.accel_bias_rw_sigma ** 2 * t) - 1.0).max() < 1e-6
    assert np.abs(np.diag(full)[12:15] / (noise.gyro_bias_rw_sigma ** 2 * t) - 1.0).max() < 1e-6


def test_pose3_expmap_is_the_matrix_exponential_of_the_twist(oracle_lib):
    rng = np.random.default_rng(3)
    for scale in (1e-9, 1e-4, 0.3, 2.5):
        w = rng.normal(size=3); w *= scale / np.linalg.norm(w)
        v = rng.normal(size=3)
        X = np.zeros((4, 4)); X[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]; X[:3, 3] = v
        E = sl.expm(X)
        ident = np.concatenate([np.eye(3).ravel(), np.zeros(3)])
        out = po.retract_var(abi.VAR_POSE3, ident, np.concatenate([w, v]))      # tangent order [rot, trans] (GTSAM Pose3)
        # (GTSAM: for |w|^2 <= eps the translation is v itself, i.e. the w x v / 2 term is dropped: error <= |w| |v| / 2 there)
        ttol = 1e-13 if scale * scale > np.finfo(float).eps else 0.5 * scale * np.linalg.norm(v) + 1e-13
        assert np.abs(out[:9].reshape(3, 3) - E[:3, :3]).max() < 1e-13 and np.abs(out[9:] - E[:3, 3]).max() < ttol
        # retract = T * Expmap(xi)
        R0 = synth.exp_so3(rng.normal(size=3)); t0 = rng.normal(size=3)
        out = po.retract_var(abi.VAR_POSE3, np.concatenate([R0.ravel(), t0]), np.concatenate([w, v]))
        assert np.abs(out[:9].reshape(3, 3) - R0 @ E[:3, :3]).max() < 1e-13 and np.abs(out[9:] - (t0 + R0 @ E[:3, 3])).max() < ttol
    assert np.abs(po.so3_logmap(po.so3_expmap(np.array([0.2, -0.7, 0.4]))) - np.array([0.2, -0.7, 0.4])).max() < 1e-14


def test_unit3_retract_properties(oracle_lib):
    rng = np.random.default_rng(5)
    for _ in range(20):
        p = rng.normal(size=3); p /= np.linalg.norm(p)
        assert np.abs(po.retract_var(abi.VAR_UNIT3, p, np.zeros(2)) - p).max() < 1e-15
        # tangent basis from two unit steps: B = d retract / d v at 0
        h = 1e-7
        B = np.stack([(po.retract_var(abi.VAR_UNIT3, p, h * e) - po.retract_var(abi.VAR_UNIT3, p, -h * e)) / (2 * h) for e in np.eye(2)], axis=1)
        assert np.abs(B.T @ B - np.eye(2)).max() < 1e-8 and np.abs(B.T @ p).max() < 1e-8
        v = rng.normal(size=2) * 0.7
        q = po.retract_var(abi.VAR_UNIT3, p, v)
        th = np.linalg.norm(v)
        assert abs(np.linalg.norm(q) - 1.0) < 1e-14 and abs(q @ p - np.cos(th)) < 1e-12
        assert np.abs(q - (np.cos(th) * p + np.sin(th) / th * (B @ v))).max() < 1e-7     # p cos + B v sin / |v| (GTSAM Unit3::retract)


def _one_variable_prior_graph(mu, sigma):
    g = abi.GraphDesc(1, [abi.VAR_VEC3])
    g.add_slot(abi.F_PRIOR_VEC3, [(abi.AT_K, 0)], params=list(sigma) + list(mu))
    return g


def test_try_lambda_on_a_quadratic(oracle_lib):
    """PriorFactor<Vector3> alone: cost = 1/2 |(x - mu) / sigma|^2, H = diag(1 / sigma^2).  iterate(): x <- x + (H + lambda I)^-1 H (mu - x),
    accepted (the model is exact: fidelity 1), lambda <- lambda / 10."""
    mu, sigma = np.array([1.0, -2.0, 0.5]), np.array([0.1, 2.0, 1.0])
    g = _one_variable_prior_graph(mu, sigma)
    o = po.OracleProblem(g)
    x = np.array([[0.3, 0.3, 0.3]])
    o.set_values(x, np.zeros(0))
    lm = abi.LmParams.lower_body(); lm.lambda_initial = 0.5
    e0 = o.lm_begin(lm)
    assert abs(e0 - 0.5 * np.sum(((x[0] - mu) / sigma) ** 2)) < 1e-13
    lam = lm.lambda_initial
    for it in range(1, 4):
        st = o.lm_iterate()
        H = 1.0 / sigma ** 2
        x = x + (H / (H + lam)) * (mu - x)
        assert st.iterations == it and st.inner_trials == 1
        assert abs(st.error - 0.5 * np.sum(((x[0] - mu) / sigma) ** 2)) <= 1e-12 * e0
        lam = max(lam / 10.0, lm.lambda_lower_bound)
        assert abs(st.lambda_ - lam) <= 1e-15 * lam
        xv, _ = o.get_values()
        assert np.abs(xv - x).max() < 1e-13


def test_try_lambda_rejects_and_retries_with_the_same_linearization(oracle_lib):
    """the 7-IMU graph from the reference's zero initialisation rejects its first lambdas: an iterate() that needed m trials must
    (i) leave lambda = lambda_0 10^(m-1) / 10 behind, (ii) have taken exactly the step (H + lambda_0 10^(m-1) I)^-1 (-g) of the
    linearization AT THE VALUES BEFORE the call (rejected rungs do not relinearize), (iii) never increase the cost."""
    import graphs
    tr = synth.make_trial(seed=11, duration_s=1.0)
    g, tv, sv = graphs.lower_body_graph(tr)
    o = po.OracleProblem(g); o.set_values(tv, sv)
    lm = abi.LmParams.lower_body()
    e = o.lm_begin(lm)
    lam = lm.lambda_initial
    seen = 0
    for _ in range(6):
        pre_t, pre_s = o.get_values()
        st = o.lm_iterate()
        assert st.error <= e * (1 + 1e-15) and st.iterations > 0
        m = st.inner_trials
        lam_acc = lam * 10.0 ** (m - 1)
        assert abs(st.lambda_ - max(lam_acc / 10.0, lm.lambda_lower_bound)) <= 1e-12 * st.lambda_
        if m > 1:
            seen += 1
            o2 = po.OracleProblem(g); o2.set_values(pre_t, pre_s); o2.linearize()
            stat, dt_, ds_ = o2.solve(lam_acc)
            assert stat == 0
            o2.retract(dt_, ds_)
            t2, s2 = o2.get_values()
            post_t, post_s = o.get_values()
            assert np.abs(t2 - post_t).max() <= 1e-9 * max(1.0, np.abs(post_t).max()) and np.abs(s2 - post_s).max() <= 1e-9
            assert abs(o2.error() - st.error) <= 1e-9 * st.error
        e, lam = st.error, st.lambda_
    assert seen >= 1
