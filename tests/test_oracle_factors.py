"""-m "not gpu": pins the CPU oracle's factor arithmetic the way the reference pins its own factors:
analytic Jacobian == numerical derivative on seeded random inputs at the reference's tolerances, the reference's
known-answer cases, and "factor + priors optimises to zero"."""
import numpy as np
import pytest
from bioslam_b200 import abi
from oracle import pyoracle as po
import graphs

TOL = {  # reference tolerances: test/unit/test*.cpp (1e-9 hinge/joint-pos, 1e-7 ang-vel/seg-length/joint-vel, 1e-5 norm/angle, 5e-3 compass)
    abi.F_HINGE_VEC: 1e-8, abi.F_JOINTPOS_VEC: 1e-8, abi.F_ANGVEL: 1e-7, abi.F_SEGLEN_MAG: 1e-7, abi.F_SEGLEN_MAX: 1e-7,
    abi.F_SEGLEN_MIN: 1e-7, abi.F_SEGLEN_DISCREPANCY: 1e-7, abi.F_JOINTVEL_VEC: 1e-7, abi.F_HINGE_NORM: 1e-5, abi.F_JOINTPOS_NORM: 1e-5,
    abi.F_ANGLE_AXIS_SEG: 1e-5, abi.F_POSE3_COMPASS_PRIOR: 5e-3, abi.F_POSE3_TRANSLATION_PRIOR: 1e-8, abi.F_PRIOR_VEC3: 1e-9,
    abi.F_PRIOR_BIAS6: 1e-9, abi.F_MAG_POSE3: 1e-8, abi.F_IMU_COMBINED: 1e-6, abi.F_IMU_STATICBIAS: 1e-6, abi.F_JOINTVEL_NORM: 1e-5,
}


def rand_val(rng, t, rot_scale=1.0):
    if t == abi.VAR_POSE3:
        return np.concatenate([po.so3_expmap(rng.normal(size=3) * rot_scale).ravel(), rng.normal(size=3)])
    if t == abi.VAR_VEC3:
        return rng.normal(size=3)
    if t == abi.VAR_BIAS6:
        return rng.normal(size=6) * 0.1
    v = rng.normal(size=3)
    return v / np.linalg.norm(v)


def rand_factor(rng, ft):
    rows, vts, nc = abi.FACTOR_INFO[ft]
    xs = [rand_val(rng, t, 0.5 if ft in (abi.F_IMU_COMBINED, abi.F_IMU_STATICBIAS) else 1.0) for t in vts]
    params = np.abs(rng.normal(size=24)) + 0.5
    if ft in (abi.F_IMU_COMBINED, abi.F_IMU_STATICBIAS):
        params[:3] = [0, 0, -9.81]
    if ft == abi.F_POSE3_COMPASS_PRIOR:
        params[2:5] = [0, 1, 0]; params[5:8] = [1, 0, 0]; params[8:11] = [0, 0, 1]
    if ft == abi.F_SEGLEN_MAX:
        params[1] = 0.3
    if ft == abi.F_SEGLEN_MIN:
        params[1] = 5.0
    consts = rng.normal(size=max(nc, 1)) * 0.3
    if nc >= 115:
        consts[69] = 0.1
    return xs, params, consts


def numerical_jacobian(ft, params, consts, xs, h=1e-6):
    _, vts, _ = abi.FACTOR_INFO[ft]
    x = np.concatenate(xs)
    r0 = po.factor_eval(ft, params, consts, x, jac=False)
    J = np.zeros((r0.size, sum(abi.TAN_DIM[t] for t in vts)))
    col = off = 0
    for vi, t in enumerate(vts):
        for j in range(abi.TAN_DIM[t]):
            d = np.zeros(abi.TAN_DIM[t])
            d[j] = h; xp = x.copy(); xp[off:off + abi.VAL_DIM[t]] = po.retract_var(t, xs[vi], d)
            d[j] = -h; xm = x.copy(); xm[off:off + abi.VAL_DIM[t]] = po.retract_var(t, xs[vi], d)
            J[:, col] = (po.factor_eval(ft, params, consts, xp, jac=False) - po.factor_eval(ft, params, consts, xm, jac=False)) / (2 * h)
            col += 1
        off += abi.VAL_DIM[t]
    return J


@pytest.mark.parametrize("ft", sorted(TOL))
def test_analytic_jacobian_matches_numerical(ft, oracle_lib):
    rng = np.random.default_rng(1000 + ft)
    worst = 0.0
    for _ in range(60):
        xs, params, consts = rand_factor(rng, ft)
        _, J = po.factor_eval(ft, params, consts, np.concatenate(xs))
        Jn = numerical_jacobian(ft, params, consts, xs)
        worst = max(worst, np.abs(J - Jn).max() / max(1.0, np.abs(Jn).max()))
    assert worst < TOL[ft], (ft, worst)


def test_known_answer_linear_velocity(oracle_lib):
    """reference test/unit/testConstrainedJointCenterVelocityFactor.cpp:81-96: v=0, w=(2,0,0), r=(0,4,0), R=I -> (0,0,8)"""
    I = np.concatenate([np.eye(3).ravel(), np.zeros(3)])
    xs = [I, np.zeros(3), np.array([2.0, 0, 0]), np.array([0, 4.0, 0]), I, np.zeros(3), np.zeros(3), np.zeros(3)]
    r = po.factor_eval(abi.F_JOINTVEL_VEC, [1, 1, 1], None, np.concatenate(xs), jac=False)
    assert np.allclose(r, [0, 0, 8.0], atol=1e-15)


def test_known_answer_hinge_zero_when_relative_angvel_parallel_to_axis(oracle_lib):
    rng = np.random.default_rng(5)
    for _ in range(20):
        xA, xB = rand_val(rng, abi.VAR_POSE3), rand_val(rng, abi.VAR_POSE3)
        k = rand_val(rng, abi.VAR_UNIT3)
        wA = rng.normal(size=3)
        RA, RB = xA[:9].reshape(3, 3), xB[:9].reshape(3, 3)
        wB = RB.T @ RA @ (wA + 1.7 * k)  # R_A^T R_B w_B - w_A = 1.7 k
        r = po.factor_eval(abi.F_HINGE_VEC, [1, 1, 1], None, np.concatenate([xA, wA, xB, wB, k]), jac=False)
        assert np.abs(r).max() < 1e-13


def test_known_answer_joint_centre_zero_for_consistent_pair(oracle_lib):
    rng = np.random.default_rng(6)
    for _ in range(20):
        xA, xB = rand_val(rng, abi.VAR_POSE3), rand_val(rng, abi.VAR_POSE3)
        sA = rng.normal(size=3)
        c = xA[:9].reshape(3, 3) @ sA + xA[9:]
        sB = xB[:9].reshape(3, 3).T @ (c - xB[9:])
        r = po.factor_eval(abi.F_JOINTPOS_VEC, [1, 1, 1], None, np.concatenate([xA, xB, sA, sB]), jac=False)
        assert np.abs(r).max() < 1e-13


def test_known_answer_compass_literal_projection(oracle_lib):
    """R=I, l_B=(0,1,0), ref_N=(1,0,0), up=(0,0,1) -> +pi/2 (the reference's negated projection)"""
    I = np.concatenate([np.eye(3).ravel(), np.zeros(3)])
    r = po.factor_eval(abi.F_POSE3_COMPASS_PRIOR, [1, 0.0, 0, 1, 0, 1, 0, 0, 0, 0, 1], None, I, jac=False)
    assert abs(r[0] - np.pi / 2) < 1e-14


def test_seglen_penalties_are_one_sided(oracle_lib):
    a, b = np.array([0.3, 0, 0]), np.zeros(3)
    assert po.factor_eval(abi.F_SEGLEN_MAX, [1e-3, 0.4, 4.000000000000001], None, np.concatenate([a, b]), jac=False)[0] == 0.0
    assert po.factor_eval(abi.F_SEGLEN_MIN, [1e-3, 0.2, 4.000000000000001], None, np.concatenate([a, b]), jac=False)[0] == 0.0
    assert po.factor_eval(abi.F_SEGLEN_MAX, [1e-3, 0.2, 4.000000000000001], None, np.concatenate([a, b]), jac=False)[0] > 0.0


def _tiny_graph(ft, rng, tight):
    """one factor instance + priors on all of its variables (tight on `tight`, loose on the rest): the reference's
    optimise-to-zero test (e.g. testHingeJointConstraintVectorErrorEstAngVel.cpp:41-123)"""
    rows, vts, nc = abi.FACTOR_INFO[ft]
    g = abi.GraphDesc(1, list(vts), [], const_stride=max(nc, 0))
    sig = [1e-2] * rows
    g.add_slot(ft, [(abi.AT_K, i) for i in range(len(vts))], sig + ([0.0] if rows == 1 else []), 0, 1, const_offset=-1)
    xs = [rand_val(rng, t) for t in vts]
    prior_type = {abi.VAR_POSE3: abi.F_PRIOR_POSE3, abi.VAR_VEC3: abi.F_PRIOR_VEC3, abi.VAR_BIAS6: abi.F_PRIOR_BIAS6, abi.VAR_UNIT3: abi.F_PRIOR_UNIT3}
    for i, t in enumerate(vts):
        s = 1e-4 if i in tight else 1e3
        n = abi.TAN_DIM[t]
        g.add_slot(prior_type[t], [(abi.AT_K, i)], [s] * n + list(xs[i]), 0, 1)
    return g, np.concatenate(xs)[None, :]


@pytest.mark.parametrize("ft,tight", [(abi.F_HINGE_VEC, (0, 1)), (abi.F_HINGE_VEC, (2, 3, 4)), (abi.F_JOINTPOS_VEC, (0,)),
                                      (abi.F_JOINTPOS_VEC, (1, 2)), (abi.F_JOINTPOS_VEC, (0, 1, 3))])
def test_factor_with_priors_optimises_to_zero(ft, tight, oracle_lib):
    rng = np.random.default_rng(77)
    for _ in range(5):
        g, tv = _tiny_graph(ft, rng, set(tight))
        o = po.OracleProblem(g)
        o.set_values(tv, np.zeros(0))
        lm = abi.LmParams(1e-5, 10.0, 0.0, 1e5, 1e-5, 1e-5, 1e-3, 0, 0)  # GTSAM defaults
        o.lm_begin(lm)
        for _ in range(60):
            st = o.lm_iterate()
            if st.error < 1e-10:
                break
        tv2, _ = o.get_values()
        x = tv2[0]
        r = po.factor_eval(ft, [1] * 24, None, x, jac=False)
        assert np.abs(r).max() < 1e-3, (ft, tight, r)


def test_degenerate_geometry_keeps_finite_jacobians_like_gtsam_norm3(oracle_lib):
    """gtsam::norm3 returns the gradient (1, 1, 1) for |v| <= 1e-10 (gtsam/geometry/Point3.cpp @4.2); every norm of the reference's
    factors goes through it (src/mathutils.cpp:75,754).  Reachable case: Pose3CompassPrior with the class defaults
    refVecLocal == refVecNav == [1,0,0]: the literal projection (v x n) x n is MINUS the horizontal projection
    (src/mathutils.cpp:771-781), so an identity pose gives the anti-parallel case (angle pi) and a half-turn about z the parallel one."""
    from bioslam_b200 import abi
    eye = np.concatenate([np.eye(3).ravel(), np.zeros(3)])
    half_turn = np.concatenate([np.diag([-1.0, -1.0, 1.0]).ravel(), np.zeros(3)])
    for pose, expect in ((eye, np.pi), (half_turn, 0.0)):
        r, J = po.factor_eval(abi.F_POSE3_COMPASS_PRIOR, [1e-3, 0.0, 1, 0, 0, 1, 0, 0, 0, 0, 1], None, pose)
        assert np.all(np.isfinite(J)) and np.isfinite(r[0]) and abs(abs(r[0]) - expect) < 1e-12
    # coincident points: |v| = 0 -> gradient (1,1,1) for the first point, -(1,1,1) for the second
    r, J = po.factor_eval(abi.F_SEGLEN_MAG, [0.01, 0.4], None, np.array([0.1, 0.2, 0.3, 0.1, 0.2, 0.3]))
    assert np.allclose(J, [[1, 1, 1, -1, -1, -1]]) and abs(r[0] + 0.4) < 1e-15
    r, J = po.factor_eval(abi.F_SEGLEN_DISCREPANCY, [0.002], None, np.array([0.1, 0.2, 0.3, 0.1, 0.2, 0.3, 0, 0, 0, 0, 0, 1.0]))
    assert np.allclose(J, [[1, 1, 1, -1, -1, -1, 0, 0, 1, 0, 0, -1]])
