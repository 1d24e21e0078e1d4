import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bioslam_b200 import abi, lib, synth
import graphs, parity_checks as pc
tr = synth.make_trial(seed=5, duration_s=60.0)
g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh))
tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
ch, gr = int(sys.argv[1]), int(sys.argv[2])
ref = None
for rep in range(10):
    p = lib.Problem(g, chunks=ch, groups=gr); p.set_values(tv, sv); p.linearize()
    gt, gs = p.gradient(); S = p.static_hessian()
    blocks = [p.hessian_blocks(k) for k in (0, 1, 300, 598, 599)]
    st, dt, ds = p.solve(1e-2)
    cur = dict(gt=gt, gs=gs, S=S, D=np.stack([b[0] for b in blocks]), O=np.stack([b[1] for b in blocks]), A=np.stack([b[2] for b in blocks]), dt=dt, ds=ds)
    if ref is None:
        ref = cur
    diffs = {k: float(np.abs(cur[k] - ref[k]).max()) for k in cur}
    print(rep, "status", st, "|dt|max", float(np.abs(dt).max()), {k: v for k, v in diffs.items() if v}, flush=True)
    p.close()
