"""Development probe (run under gpurun): parity spot-checks against the oracle and timing sweeps over the solver partition."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bioslam_b200 import abi, lib, synth
import graphs


def log(**kw):
    print(json.dumps(kw), flush=True)


def parity(K_s):
    from oracle import pyoracle as po
    tr = synth.make_trial(seed=3, duration_s=K_s)
    g, tv, sv = graphs.lower_body_graph(tr)
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    o = po.OracleProblem(g, threads=8); o.set_values(tv, sv)
    for chunks in (1, 4, None):
        p = lib.Problem(g, chunks=chunks); p.set_values(tv, sv)
        o.set_values(tv, sv)
        eo, ep = o.error(), p.error()
        o.linearize(); p.linearize()
        go, gso = o.gradient(); gp, gsp = p.gradient()
        Do, Oo, Ao = o.hessian_blocks(1); Dp, Op, Ap = p.hessian_blocks(1)
        st, dto, dso = o.solve(1e-3); st2, dtp, dsp = p.solve(1e-3)
        log(test="parity", K=g.K, chunks=p.chunks, err_rel=abs(eo - ep) / eo, grad=float(np.abs(go - gp).max() / np.abs(go).max()),
            D=float(np.abs(Do - Dp).max() / np.abs(Do).max()), O=float(np.abs(Oo - Op).max() / np.abs(Oo).max()),
            A=float(np.abs(Ao - Ap).max() / np.abs(Ao).max()), solve_st=[st, st2],
            dt=float(np.abs(dto - dtp).max() / np.abs(dto).max()), ds=float(np.abs(dso - dsp).max() / np.abs(dso).max()))
        lm = abi.LmParams.lower_body()
        o.set_values(tv, sv); p.set_values(tv, sv)
        o.lm_begin(lm); p.lm_begin(lm)
        for it in range(5):
            so, sp = o.lm_iterate(), p.lm_iterate()
            log(test="lm", K=g.K, chunks=p.chunks, it=it, o_err=so.error, p_err=sp.error, o_it=so.iterations, p_it=sp.iterations,
                o_tr=so.inner_trials, p_tr=sp.inner_trials, st=sp.status)
        p.close()


def timing(K_s, chunk_list, iters=6):
    tr = synth.make_trial(seed=5, duration_s=K_s)
    t0 = time.time()
    rec = lib.preintegrate(tr.acc, tr.gyro, tr.dt, 10, np.zeros((7, 6)))
    t_pre = time.time() - t0
    g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh))
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    log(test="setup", K=g.K, preint_s=t_pre)
    for chunks in chunk_list:
        p = lib.Problem(g, chunks=chunks); p.set_values(tv, sv)
        p.phase_ms(reset=True)
        lm = abi.LmParams.lower_body()
        p.lm_begin(lm)
        p.lm_iterate()  # warm-up
        p.phase_ms(reset=True)
        t0 = time.time(); trials = 0
        for it in range(iters):
            s = p.lm_iterate(); trials += s.inner_trials
        wall = time.time() - t0
        ph = p.phase_ms()
        log(test="timing", K=g.K, chunks=p.chunks, iters=iters, trials=trials, wall_ms=wall * 1e3, lin_ms_per_iter=ph[0] / iters,
            solve_ms_per_trial=ph[1] / trials, rest_ms_per_trial=ph[2] / trials, err=s.error, lam=s.lambda_)
        p.close()


if __name__ == "__main__":
    what = sys.argv[1:] or ["parity", "t600", "t6000"]
    if "parity" in what:
        parity(4.0)
    if "t600" in what:
        timing(60.0, [1, 8, 16, 24, 36, 48, 74])
    if "t6000" in what:
        timing(600.0, [37, 74, 110, 148])
