#!/usr/bin/env python
"""Run the CPU oracle (oracle/, test infrastructure) to the reference's convergence rule on a synthetic 7-IMU trial and write
the golden record the -m gpu parity tests compare the CUDA path with.

  python tools/oracle_converge.py --seconds 60 --seed 0 --out tests/golden/converged_60s_seed0.npz

Per LM iteration: cost, lambda, number of lambda trials (the accept/reject sequence).  At the end: iterations, trials, final
cost, static values, IMU positions and clinical hip/knee angles of every `--stride`-th keyframe, wall time and thread count
(the full-size run is also the measured CPU denominator of the ">= 50x" target).  The exit rule is bioslam's
(reference src/lowerBodyPoseEstimator.cpp:656-667): 6 consecutive iterations with abs decrease < 1e-6 or rel decrease < 1e-5.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--threads", type=int, default=os.cpu_count())
    ap.add_argument("--chunks", type=int, default=16, help="time chunks of the fast CPU solver (results depend on it, not on --threads)")
    ap.add_argument("--plain-solver", action="store_true", help="use the plain dense block sweep instead of oracle/fastsolve.hpp")
    ap.add_argument("--max-iterations", type=int, default=5000)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--out", required=True)
    ap.add_argument("--progress", type=int, default=50)
    ap.add_argument("--checkpoint", type=int, default=0, help="write <out>.partial.npz (history / lambdas / trials so far) every N iterations")
    args = ap.parse_args()

    from bioslam_b200 import abi, synth
    from bioslam_b200 import graphs as bgraphs
    from oracle import pyoracle as po
    import graphs
    po.build()
    trial = synth.make_trial(seed=args.seed, duration_s=args.seconds)
    g, tv, sv = graphs.lower_body_graph(trial)      # oracle preintegration, reference initial values
    o = po.OracleProblem(g, threads=args.threads)
    o.set_solver(not args.plain_solver, args.chunks)
    o.set_values(tv, sv)
    lm, outer = abi.LmParams.lower_body(), abi.OuterParams.lower_body(max_iterations=args.max_iterations)
    t0 = time.perf_counter()
    e0 = o.lm_begin(lm)
    hist, lams, trials = [e0], [lm.lambda_initial], [0]
    consec, cur, status = 0, e0, 0
    while consec < outer.converge_at_consec_small_iter and len(hist) - 1 < args.max_iterations:
        prev = cur
        st = o.lm_iterate()
        cur = st.error
        hist.append(cur); lams.append(st.lambda_); trials.append(st.inner_trials)
        abs_dec, rel_dec = prev - cur, (prev - cur) / prev
        small = abs_dec < outer.abs_err_decrease_limit or rel_dec < outer.rel_err_decrease_limit or cur < outer.abs_err_limit
        consec = consec + 1 if small else 0
        status = st.status
        if args.progress and (len(hist) - 1) % args.progress == 0:
            print(f"it {len(hist) - 1} cost {cur:.10f} lambda {st.lambda_:g} trials {sum(trials)} t {time.perf_counter() - t0:.1f}s", flush=True)
        if args.checkpoint and (len(hist) - 1) % args.checkpoint == 0:
            np.savez_compressed(args.out + ".partial.npz", history=np.array(hist), lambdas=np.array(lams), inner_trials=np.array(trials, dtype=np.int32),
                                wall_s=time.perf_counter() - t0)
        if st.status != 0 and st.iterations == len(hist) - 2:
            break
    wall = time.perf_counter() - t0
    tvf, svf = o.get_values()
    off = g.time_value_offsets()
    pos = np.stack([tvf[:, off[3 * i] + 9:off[3 * i] + 12] for i in range(7)], axis=1)     # [K][7][3]
    rot = np.stack([tvf[:, off[3 * i]:off[3 * i] + 9] for i in range(7)], axis=1)          # [K][7][9]
    ja = po.lower_body_joint_angles(g, tvf, svf, bgraphs.lower_body_joint_angle_desc())  # [K][12] rad
    s = slice(None, None, args.stride)
    meta = {"seconds": args.seconds, "seed": args.seed, "K": int(g.K), "chunks": args.chunks, "threads": args.threads,
            "solver": "plain" if args.plain_solver else "fastsolve", "iterations": len(hist) - 1, "trials": int(sum(trials)),
            "initial_error": e0, "final_error": cur, "final_lambda": lams[-1], "status": int(status), "wall_s": wall,
            "stride": args.stride, "rule": "6 consecutive iterations with abs<1e-6 or rel<1e-5 decrease (reference exit rule)",
            "generator": "tools/oracle_converge.py", "cpu": os.uname().machine}
    np.savez_compressed(args.out, history=np.array(hist), lambdas=np.array(lams), inner_trials=np.array(trials, dtype=np.int32),
                        sv=svf, positions=pos[s], rotations=rot[s], joint_angles=ja[s], meta=json.dumps(meta))
    print(json.dumps(meta))


if __name__ == "__main__":
    main()
