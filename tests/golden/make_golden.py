"""Generates tests/golden/*.npz with the CPU oracle (the reference publishes no golden vectors):
seeded inputs -> per-factor residual/Jacobian, and a seeded synthetic trial -> {initial error, per-iteration (error, lambda,
trials), final values}.  Run: python tests/golden/make_golden.py   (commit the .npz files it writes)."""
import os
import subprocess
import sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bioslam_b200 import abi, synth
from oracle import pyoracle as po
import graphs
from test_oracle_factors import rand_factor


def main():
    po.build()
    rng = np.random.default_rng(20261019)
    out = {}
    for ft in sorted(abi.FACTOR_INFO):
        xs, params, consts = rand_factor(rng, ft)
        x = np.concatenate(xs)
        r, J = po.factor_eval(ft, params, consts, x, whiten=False)
        out[f"f{ft}_x"] = x; out[f"f{ft}_params"] = params; out[f"f{ft}_consts"] = consts; out[f"f{ft}_r"] = r; out[f"f{ft}_J"] = J
    np.savez_compressed(os.path.join(HERE, "factors.npz"), **out)
    # LM trace on a 2 s trial (K = 20)
    tr = synth.make_trial(seed=11, duration_s=2.0)
    g, tv, sv = graphs.lower_body_graph(tr)
    o = po.OracleProblem(g); o.set_values(tv, sv)
    lm = abi.LmParams.lower_body()
    e0 = o.lm_begin(lm)
    trace = []
    for _ in range(12):
        st = o.lm_iterate(); trace.append([st.error, st.lambda_, st.inner_trials, st.iterations])
    tvf, svf = o.get_values()
    rec = po.preintegrate(tr.acc, tr.gyro, tr.dt, 10, np.zeros((7, 6)))
    np.savez_compressed(os.path.join(HERE, "lm_trace_seed11_2s.npz"), e0=e0, trace=np.array(trace), tv=tvf, sv=svf,
                        preint_first=rec[:, :2], acc_first=tr.acc[:, :20], gyro_first=tr.gyro[:, :20])
    rev = subprocess.run(["git", "rev-parse", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
    open(os.path.join(HERE, "README.md"), "w").write(
        "Golden fixtures written by make_golden.py with the CPU oracle (seed 20261019 for factors, synth seed 11 for the LM trace).\n"
        f"Oracle revision: {rev}\n")
    print("wrote golden fixtures; e0 =", e0, "final", trace[-1])


if __name__ == "__main__":
    main()
