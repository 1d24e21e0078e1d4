"""Time-to-converge probe (BASELINE.json metric, second half): python tools/converge_probe.py <seconds of data> <max iterations>"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bioslam_b200 import abi, graphs, lib, synth
dur = float(sys.argv[1]); cap = int(sys.argv[2])
tr = synth.make_trial(seed=0, duration_s=dur)
g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh))
p = lib.Problem(g); p.set_values(tv, sv)
res, hist = p.optimize(abi.LmParams.lower_body(), abi.OuterParams.lower_body(max_iterations=cap))
print(json.dumps({"K": g.K, "levels": p.levels, "iterations": int(res.iterations), "trials": int(res.total_trials), "seconds": res.device_ms * 1e-3,
                  "initial_error": float(res.initial_error), "final_error": float(res.final_error), "converged": bool(res.iterations < cap),
                  "history_every_100": [float(h) for h in hist[::100]]}))
