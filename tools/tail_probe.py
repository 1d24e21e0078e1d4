"""Where does the 10-minute trial really converge?  python tools/tail_probe.py <max iterations> [chunks ...]
Runs the K = 6000 lower-body solve from the reference's initial values with a far stricter exit rule than bioslam's (relative decrease
below 1e-9 for 6 iterations) on several time partitions (0 = the library's own) and prints the cost every 500 iterations: the runs are
different LM trajectories (rounding differs with the partition) and meet bioslam's rule -- 1e-5 relative per iteration -- at different
points of a tail that is still descending; run on, they approach the same minimum."""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bioslam_b200 import abi, graphs, lib, synth
cap = int(sys.argv[1]); parts = [int(a) for a in sys.argv[2:]] or [0]
tr = synth.make_trial(seed=0, duration_s=600.0)
g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh))
for chunks in parts:
    p = lib.Problem(g, chunks=chunks if chunks > 0 else None); p.set_values(tv, sv)
    outer = abi.OuterParams(1.0e-12, 1.0e-9, 1.0e-5, cap, 6, 0, cap + 8)
    res, hist = p.optimize(abi.LmParams.lower_body(), outer)
    hist = np.asarray(hist)
    rel = (hist[:-1] - hist[1:]) / hist[:-1]
    small = rel < 1.0e-5
    run = np.convolve(small.astype(int), np.ones(6, dtype=int), mode="valid")
    stop = int(np.argmax(run == 6)) + 6 if np.any(run == 6) else -1       # outer call at which bioslam's rule would have stopped
    print(json.dumps({"chunks": p.chunks, "levels": p.levels, "outer_calls": int(res.outer_calls), "iterations": int(res.iterations),
                      "trials": int(res.total_trials), "seconds": res.device_ms * 1e-3, "final_error": float(res.final_error),
                      "bioslam_rule_stop_at_call": stop, "cost_there": float(hist[stop]) if stop > 0 else None,
                      "cost_every_500_calls": {str(i): float(hist[i]) for i in range(0, len(hist), 500)}}), flush=True)
    del p
