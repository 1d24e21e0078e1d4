"""Which partition configurations give bit-identical repeated solves? (development probe)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bioslam_b200 import abi, lib, synth
import graphs
tr = synth.make_trial(seed=5, duration_s=60.0)
g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh))
tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
import itertools
cfgs = [tuple(int(x) for x in a.split(',')) for a in sys.argv[1:]] or [(1, 1), (36, 1), (75, 12), (0, 0)]
for ch, gr in cfgs:
    sols = []
    for rep in range(4):
        p = lib.Problem(g, chunks=ch, groups=gr); p.set_values(tv, sv); p.linearize()
        for inner in range(2):
            st, dt, ds = p.solve(1e-2); sols.append((dt.copy(), ds.copy()))
        cfg = (p.chunks, p.groups)
        p.close()
    ref = sols[0]
    bad = [i for i, s in enumerate(sols) if not (np.array_equal(s[0], ref[0]) and np.array_equal(s[1], ref[1]))]
    msg = ""
    if bad:
        d = np.abs(sols[bad[0]][0] - ref[0]).max(axis=1)
        ks = np.nonzero(d)[0]
        msg = f" first differing run {bad[0]}: {len(ks)} keyframes differ, first {ks[:5]}, max rel {d.max() / np.abs(ref[0]).max():.2e}, ds differ {np.abs(sols[bad[0]][1]-ref[1]).max():.2e}"
    nan = [i for i, s in enumerate(sols) if not np.all(np.isfinite(s[0]))]
    pair_equal = [bool(np.array_equal(sols[2 * i][0], sols[2 * i + 1][0])) for i in range(4)]
    print("partition", cfg, "nondeterministic runs:", bad, "nan runs:", nan, "same-instance pairs equal:", pair_equal, msg, flush=True)
