"""Per-phase cycle breakdown of the solver sweep (debug build with -DB200_PHASE_CLOCKS): python tools/phase_clocks.py <sec> <chunks>

Build the instrumented library first (from bioslam_b200/):
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -O3 --expt-relaxed-constexpr \
       -DB200_PHASE_CLOCKS -shared -o ../tools/libb200_clk.so csrc/bioslam_b200.cu host/*.cpp -lcudart"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bioslam_b200 import abi, lib, synth
import graphs
SO = os.path.join(ROOT, "tools", "libb200_clk.so")
dur = float(sys.argv[1]); chunks = int(sys.argv[2])
tr = synth.make_trial(seed=5, duration_s=dur)
g, tv, sv = graphs.lower_body_graph(tr, preint=lambda a, w, dt, n_, bh: lib.preintegrate(a, w, dt, n_, bh, so_path=SO))
tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
p = lib.Problem(g, chunks=chunks if chunks > 0 else None, so_path=SO); p.set_values(tv, sv)
p.linearize()
L = lib.load(SO)
out = (C.c_ulonglong * 24)()
p.solve(1e-3); L.b200_debug_phase_clocks(out, 1)
p.solve(1e-3); L.b200_debug_phase_clocks(out, 1)
names = ["A:load D+fill", "A:chol+inv", "A:store W", "B:load panel", "B:y=pW^T", "B:z=yW", "B:fill=zO^T"]
tot = sum(out[:7])
print("K", g.K, "chunks", p.chunks, "levels", p.levels, "cycles per phase of CTA 1 of the level-0 sweep (a chunk with left-separator rows), one solve:")
for n, v in zip(names, out[:7]):
    print("  %-14s %10d  %5.1f%%" % (n, v, 100.0 * v / max(tot, 1)))
for n, i in (("chol:whole phase", 10), ("pivot:potrf16+inv", 8), ("pivot:scale row+diag tile", 9), ("pivot:wait trailing", 11),
             ("worker1:wait W", 12), ("worker1:scale column", 13), ("worker1:wait row", 14), ("worker1:trailing", 15)):
    print("  %-26s %10d  %5.1f%%" % (n, out[i], 100.0 * out[i] / max(tot, 1)))

if out[0] or out[1] or out[2]:
    print("keyframe-pipelined level-0 sweep (kernels_sweep_pipe.cuh), CTA 1, cycles over one solve:")
    g0 = out[0] + out[1] + out[2] + out[3] + out[4]
    for n, i in (("g0: wait [A] (arrow rows of g1)", 0), ("g0: S2 W^T -> tiles, store", 1), ("g0: S1 rows of O", 2), ("g0: S3 load D - fill", 3),
                 ("g0: S3 chol + inverse", 4), ("g1: wait [A] (chol of g0)", 5), ("g1: S2", 16), ("g1: S1 rows of O", 17), ("g1: S3 arrow rows", 6), ("prologue", 7)):
        print("  %-34s %10d  %5.1f%%" % (n, out[i], 100.0 * out[i] / max(g0, 1)))

if out[18] or out[19]:
    print("dense levels (cta_sweep_dense_stream), CTA (0, 0) of every dense sweep of one solve, cycles:")
    tot = sum(out[18:23])
    for n, i in (("load M = D - fill", 18), ("Cholesky (16 warps)", 19), ("W^T by substitution", 20), ("load dense O", 21), ("row tiles (TMEM)", 22)):
        print("  %-26s %10d  %5.1f%%" % (n, out[i], 100.0 * out[i] / max(tot, 1)))
