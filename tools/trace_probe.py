"""LM trace probe (1 or N ranks): [torchrun ...] python tools/trace_probe.py <seconds of data> <iterations>"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bioslam_b200 import abi, graphs, lib, synth
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dur = float(sys.argv[1]); n = int(sys.argv[2])
tr = synth.make_trial(seed=0, duration_s=dur)
g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n_, bh: lib.preintegrate(a, w, dt, n_, bh, device=local))
p = lib.Problem(g, device=local)
if world > 1:
    import torch.distributed as dist
    from bioslam_b200 import sharding
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sharding.attach(p, dist, local)
p.set_values(tv, sv)
e0 = p.lm_begin(abi.LmParams.lower_body())
if rank == 0: print("K", g.K, "levels", p.levels, "e0 %.10e" % e0, flush=True)
for it in range(n):
    st = p.lm_iterate()
    if rank == 0: print(it, "%.10e" % st.error, "%.1e" % st.lambda_, st.inner_trials, st.iterations, st.status, flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
