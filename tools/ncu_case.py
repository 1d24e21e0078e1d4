"""Small single-solve case for ncu captures: python tools/ncu_case.py <seconds> <chunks> [n_solves]
(chunks = 0: the library's own recursive multi-level partition, as in bench.py; chunks > 0: two-level scheme with that many chunks)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bioslam_b200 import abi, lib, synth
import graphs
dur = float(sys.argv[1]); chunks = int(sys.argv[2]); n = int(sys.argv[3]) if len(sys.argv) > 3 else 1
tr = synth.make_trial(seed=5, duration_s=dur)
g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n_, bh: lib.preintegrate(a, w, dt, n_, bh))
tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
p = lib.Problem(g, chunks=chunks if chunks > 0 else None); p.set_values(tv, sv)
p.linearize()
for _ in range(n):
    st, dt, ds = p.solve(1e-3)
print("K", g.K, "chunks", p.chunks, "status", st, "|dt|", float(np.abs(dt).max()))
