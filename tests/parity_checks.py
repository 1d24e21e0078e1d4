"""Shared parity assertions: a bioslam_b200.lib.Problem (real CUDA library, or the emulator build in the CPU tests) against
the oracle on identical inputs.  Tolerances: per-factor sums (error, gradient, Hessian blocks) 1e-12 relative -- same
arithmetic, different summation order; linear solves are compared through the residual of the normal equations because
the damped systems have condition numbers up to 1e16."""
import numpy as np
from bioslam_b200 import abi


def assert_linearization_parity(p, o, g, ks=None, tol=1e-11):
    eo, ep = o.error(), p.error()
    assert abs(eo - ep) <= tol * abs(eo), (eo, ep)
    o.linearize(); p.linearize()
    go, gso = o.gradient(); gp, gsp = p.gradient()
    assert np.abs(go - gp).max() <= tol * np.abs(go).max()
    if g.ns:
        assert np.abs(gso - gsp).max() <= tol * np.abs(gso).max()
        So, Sp = o.static_hessian(), p.static_hessian()
        assert np.abs(So - Sp).max() <= tol * np.abs(So).max()
    for k in (ks if ks is not None else sorted({0, 1, g.K // 2, g.K - 2, g.K - 1})):
        if k < 0 or k >= g.K:
            continue
        Do, Oo, Ao = o.hessian_blocks(k); Dp, Op, Ap = p.hessian_blocks(k)
        assert np.abs(Do - Dp).max() <= tol * np.abs(Do).max(), k
        assert np.abs(Oo - Op).max() <= tol * max(np.abs(Oo).max(), 1e-300), k
        if g.ns:
            assert np.abs(Ao - Ap).max() <= tol * np.abs(Ao).max(), k


def normal_eq_residual(o, g, lam, dt, ds):
    """|| (H + lam I) delta + g || / ||g|| evaluated block-wise from the oracle's linearization"""
    K, nt = g.K, g.nt
    gt, gs = o.gradient()
    r_t = lam * dt + gt
    r_s = (lam * ds + gs) if g.ns else np.zeros(0)
    if g.ns:
        r_s = r_s + o.static_hessian() @ ds
    for k in range(K):
        D, O, A = o.hessian_blocks(k)
        r_t[k] += D @ dt[k]
        if k + 1 < K:
            r_t[k] += O @ dt[k + 1]
            r_t[k + 1] += O.T @ dt[k]
        if g.ns:
            r_t[k] += A @ ds
            r_s += A.T @ dt[k]
    num = np.sqrt(np.sum(r_t ** 2) + np.sum(r_s ** 2))
    den = np.sqrt(np.sum(gt ** 2) + np.sum(gs ** 2))
    return num / den


def assert_solve_parity(p, o, g, lams=(1e-3, 1.0), tol=1e-7):
    for lam in lams:
        st_o, dto, dso = o.solve(lam)
        st_p, dtp, dsp = p.solve(lam)
        assert st_o == 0 and st_p == 0
        ro = normal_eq_residual(o, g, lam, dto, dso)
        rp = normal_eq_residual(o, g, lam, dtp, dsp)
        assert rp <= max(10 * ro, tol), (lam, ro, rp)


def assert_lm_parity(p, o, lm, iters, rel=5e-3):
    e0o, e0p = o.lm_begin(lm), p.lm_begin(lm)
    assert abs(e0o - e0p) <= 1e-11 * abs(e0o)
    for it in range(iters):
        so, sp = o.lm_iterate(), p.lm_iterate()
        assert sp.iterations == so.iterations and sp.inner_trials == so.inner_trials, (it, so.inner_trials, sp.inner_trials)
        assert abs(so.error - sp.error) <= rel * abs(so.error), (it, so.error, sp.error)
        assert abs(so.lambda_ - sp.lambda_) <= 1e-12 * so.lambda_


def unpack_sqrt_info(rec, n):
    R = np.zeros((n, n)); off = 70
    for r in range(n):
        R[r, r:] = rec[off:off + n - r]; off += n - r
    return R


def assert_preintegration_parity(a, b, combined):
    """records [n_imu][n_int][190|115]: zeta / bias Jacobians / biasHat / dt to 1e-13 absolute, the information matrix
    R^T R to 1e-11 after symmetric diagonal normalisation (entries of R itself span twelve orders of magnitude)"""
    n = 15 if combined else 9
    assert np.abs(a[..., :70] - b[..., :70]).max() < 1e-13
    for m in range(a.shape[0]):
        for q in range(0, a.shape[1], max(1, a.shape[1] // 7)):
            Ra, Rb = unpack_sqrt_info(a[m, q], n), unpack_sqrt_info(b[m, q], n)
            Ia, Ib = Ra.T @ Ra, Rb.T @ Rb
            s = np.sqrt(np.outer(np.diag(Ia), np.diag(Ia)))
            assert np.abs((Ia - Ib) / s).max() < 1e-11


def assert_local_elimination_parity(p, g, lam, ks, tol=1e-13):
    """the Schur complement of the keyframe-local dofs (kernels_local.cuh) against a 40-digit mpmath evaluation on the blocks the
    device assembled: D_cc + lam I - D_cl (D_ll + lam I)^-1 D_lc cancels eight digits in the gyro-bias block (the angular-velocity
    factor ties bias and angular velocity with information 1.6e8), so a double-precision reference is itself wrong at 1e-10;
    entrywise tolerance relative to the magnitude of the terms that cancel."""
    import mpmath as mp
    mp.mp.dps = 40
    for k in ks:
        out = p.local_elim(k, lam)
        if out is None:
            return False
        ext, Dc, Ac, Sc = out
        D, _, A = p.hessian_blocks(k)
        gt, _ = p.gradient()
        nt = g.nt
        chain = [int(e) for e in ext if e >= 0]
        loc = [i for i in range(nt) if i not in set(chain)]
        n, nl = len(chain), len(loc)
        Dl = mp.matrix(D[np.ix_(loc, loc)].tolist())
        for i in range(nl):
            Dl[i, i] += mp.mpf(lam)
        Dli = mp.inverse(Dl)
        Dcl = mp.matrix(D[np.ix_(chain, loc)].tolist())
        Pl = mp.matrix(np.vstack([A[loc].T, -gt[k][loc][None]]).tolist())           # (ns + 1) x nl: static rows, rhs row
        Pc = np.vstack([A[chain].T, -gt[k][chain][None]])
        cD, cA = Dcl * Dli * Dcl.T, Pl * Dli * Dcl.T
        refD = np.array([[float(mp.mpf(D[chain[i], chain[j]]) + (mp.mpf(lam) if i == j else 0) - cD[i, j]) for j in range(n)] for i in range(n)])
        sD = np.array([[float(mp.sqrt(abs(cD[i, i] * cD[j, j]))) for j in range(n)] for i in range(n)]) + np.abs(D[np.ix_(chain, chain)]) + 1e-300
        assert (np.abs(Dc[:n, :n] - refD) / sD).max() <= tol, (k, (np.abs(Dc[:n, :n] - refD) / sD).max())
        refA = np.array([[float(mp.mpf(Pc[a, j]) - cA[a, j]) for j in range(n)] for a in range(Pc.shape[0])])
        sA = np.abs(np.array([[float(cA[a, j]) for j in range(n)] for a in range(Pc.shape[0])])) + np.abs(Pc) + 1e-9 * np.abs(Pc).max()
        assert (np.abs(Ac[:, :n] - refA) / sA).max() <= 1e-11, (k, (np.abs(Ac[:, :n] - refA) / sA).max())
    return True
