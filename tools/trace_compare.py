"""K = 6000 (10-minute trial): the CUDA path's LM trace next to the oracle's golden record (tests/golden/converged_600s_seed0.npz):
python tools/trace_compare.py [n_iterations]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bioslam_b200 import abi, lib, synth
import graphs
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 200
d = np.load(os.path.join(ROOT, "tests", "golden", "converged_600s_seed0.npz"))
meta = json.loads(str(d["meta"]))
tr = synth.make_trial(seed=meta["seed"], duration_s=meta["seconds"])
g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n_, bh: lib.preintegrate(a, w, dt, n_, bh))
p = lib.Problem(g); p.set_values(tv, sv)
e0 = p.lm_begin(abi.LmParams.lower_body())
print("K", g.K, "initial cost", e0, "oracle", d["history"][0], "rel", abs(e0 - d["history"][0]) / e0)
first_trial_mismatch = None
for i in range(1, n_it + 1):
    st = p.lm_iterate()
    ot, oc, ol = int(d["inner_trials"][i]), float(d["history"][i]), float(d["lambdas"][i])
    if first_trial_mismatch is None and st.inner_trials != ot: first_trial_mismatch = i
    if i <= 40 or i % 20 == 0 or st.inner_trials != ot:
        print(f"it {i:4d} cuda {st.error:.9f} trials {st.inner_trials} lambda {st.lambda_:g} | oracle {oc:.9f} trials {ot} lambda {ol:g} | rel {abs(st.error - oc) / oc:.2e}")
print("first iteration whose number of lambda trials differs:", first_trial_mismatch)
