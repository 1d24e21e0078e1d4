"""BASELINE configs[4] ("batch of independent 60-s subjects, aggregate solves/sec"; SURVEY 8(e): replicas only): B independent
lowerBody problems (seeds 0..B-1, K = 600) on ONE GPU, driven from `streams` host threads (each b200_problem owns a CUDA stream;
ctypes releases the GIL during the calls), every one optimised to the reference's convergence rule.
python tools/batch_probe.py <n_subjects> <concurrent> [max_iterations]"""
import json, os, sys, time, threading, queue
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bioslam_b200 import abi, graphs, lib, synth

B = int(sys.argv[1]); conc = int(sys.argv[2]); max_it = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
lib.load()
pre = lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh)
inputs = []
for seed in range(B):
    tr = synth.make_trial(seed=seed, duration_s=60.0)
    inputs.append(graphs.lower_body_graph(tr, preint=pre))
lm, outer = abi.LmParams.lower_body(), abi.OuterParams.lower_body(max_iterations=max_it)
# warm-up (module load, first-launch costs)
g, tv, sv = inputs[0]
p = lib.Problem(g); p.set_values(tv, sv); p.lm_begin(lm); p.lm_iterate(); p.close()
results = [None] * B
work = queue.Queue()
for i in range(B): work.put(i)

def worker():
    while True:
        try:
            i = work.get_nowait()
        except queue.Empty:
            return
        g, tv, sv = inputs[i]
        p = lib.Problem(g); p.set_values(tv, sv)
        res, hist = p.optimize(lm, outer)
        results[i] = (int(res.iterations), int(res.total_trials), float(res.final_error), float(res.device_ms))
        p.close()

t0 = time.perf_counter()
ths = [threading.Thread(target=worker) for _ in range(conc)]
for t in ths: t.start()
for t in ths: t.join()
wall = time.perf_counter() - t0
print(json.dumps({"subjects": B, "concurrent": conc, "keyframes": inputs[0][0].K, "wall_s": wall, "solves_per_sec": B / wall,
                  "iterations": [r[0] for r in results], "mean_device_s_per_solve": float(np.mean([r[3] for r in results])) * 1e-3,
                  "final_error_range": [min(r[2] for r in results), max(r[2] for r in results)]}))
