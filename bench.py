#!/usr/bin/env python
"""bench.py -- LM iterations/sec of the 7-IMU lower-body factor-graph solve (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # B200 path (this repo)
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the reference's algorithm on the host cores

A "step" is one LevenbergMarquardtOptimizer::iterate() (one linearization + the lambda trials it needs) of the
synthetic 10-minute 100 Hz 7-IMU trial (K = 6000 keyframes, 756 058 unknowns; BASELINE.json configs[3], the trial the
north_star target is quoted on) starting from the reference's initial values (identity poses, zero velocity/bias,
statics at their prior means).  `value` = iterations/s with everything resident in HBM; `e2e` = the same through the
C ABI with HOST buffers (values uploaded and downloaded every step).  Time axis sharded over N GPUs for N > 1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # name -> (seconds of data, BASELINE.json config, sample rate Hz, samples per keyframe)
    "gait_10min_100hz_7imu": (600.0, "configs[3]", 100.0, 10),
    "gait_60s_100hz_7imu": (60.0, "configs[2]", 100.0, 10),
    "gait_30s_128hz_7imu": (30.0, "configs[1] stand-in (APDM Opal rate, TUG length)", 128.0, 10),
    "gait_10min_100hz_7imu_dense": (600.0, "configs[3], m_numMeasToPreint = 1 (K = 60 000)", 100.0, 1),
}
# executed fp64 work of the level-0 sweep per interior keyframe: n = 106 chain dofs, 59 arrow rows, 106 rows of the
# coupling block and 106 left-separator rows: fused Cholesky + inverse 2/3 n^3, two triangular GEMMs n^2 per dense row (half that
# for the block-sparse rows of O), block-sparse fill
SWEEP_FLOP_PER_KEYFRAME = (2.0 / 3.0) * 106 ** 3 + 2 * ((59 + 106) * 106 ** 2 + 0.5 * 106 * 106 ** 2) + (59 + 2 * 106) * 106 * 16 * 2
# SURVEY.md section 8(d), dense-block model of a whole lambda trial
SOLVE_FLOP_PER_KEYFRAME_ALGORITHMIC = 5.0e6
# algorithmic bytes per keyframe (dense-block model): the sweep kernel reads the compact normal
# equations H (4167 doubles) and writes the block factor F (28539 doubles)
SWEEP_BYTES_PER_KEYFRAME = (4167 + 28539) * 8
TRIAL_BYTES_PER_KEYFRAME = 505472
LINEARIZE_BYTES_PER_KEYFRAME = 45488


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region: NVML in-process every 10 ms (the timed region is a few
    hundred ms, too short for repeated nvidia-smi launches), nvidia-smi as the fallback."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device=0):
        super().__init__(daemon=True)
        self.device, self.samples, self.stop_flag = device, [], False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
            get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            r = int(get(self.handle))
            flags = [bool(r & 0x8), bool(r & 0x40), bool(r & 0x20), bool(r & 0x4)]   # hw, hw thermal, sw thermal, sw power cap
            return [mhz, self.max_mhz] + flags
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return None
        f = [x.strip() for x in out.split(",")]
        return [float(f[0]), float(f[1])] + [v.lower().startswith("active") for v in f[2:6]]

    def run(self):
        while not self.stop_flag:
            try:
                smp = self.sample()
                if smp:
                    self.samples.append(smp)
            except Exception:
                pass
            time.sleep(0.01 if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples: NVML and nvidia-smi unavailable"]}
        mhz = sorted(s[0] for s in self.samples)
        reasons = sorted({n for s in self.samples for n, v in zip(self.NAMES, s[2:6]) if v})
        return {"sm_mhz": mhz[len(mhz) // 2], "sm_max_mhz": self.samples[0][1], "reasons": reasons, "samples": len(mhz),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def build_problem_inputs(workload, preint, seed=0):
    from bioslam_b200 import graphs, synth
    dur, _, rate, n_per = WORKLOADS[workload]
    trial = synth.make_trial(seed=seed, duration_s=dur, rate_hz=rate)
    g, tv, sv = graphs.lower_body_graph(trial, n_per=n_per, preint=preint)
    return trial, g, tv, sv


def host_threads():
    """threads of the CPU arm: every core this process may run on, whatever OMP_NUM_THREADS says (torchrun exports 1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def oracle_problem(args):
    """the CPU port on the FULL workload (same graph, same initial values as the B200 arm), all host threads, the structure-aware
    solver of oracle/fastsolve.hpp (local dofs pre-eliminated, time chunks over threads, blocked AVX2 kernels)"""
    from bioslam_b200 import abi
    from oracle import pyoracle as po
    po.build()
    threads = host_threads()
    _, g, tv, sv = build_problem_inputs(args.workload, lambda a, w, dt, n, bh: po.preintegrate(a, w, dt, n, bh))
    o = po.OracleProblem(g, threads=threads)
    o.set_solver(True, max(16, 2 * threads))
    o.set_values(tv, sv)
    o.lm_begin(abi.LmParams.lower_body())
    return o, g, threads


def run_reference(args):
    """CPU arm: the reference's algorithm -- GTSAM 4.2 Levenberg-Marquardt semantics + bioslam's factors, restated in oracle/ because
    GTSAM / Eigen / Boost are not in this image (cpu_baseline.kind = "port") -- timed on the box's host cores on the SAME
    configuration as the B200 arm: full workload, reference initial values, `--steps` iterate() calls after `--warmup`."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    o, g, threads = oracle_problem(args)
    trials = 0
    for _ in range(args.warmup):
        o.lm_iterate()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = o.lm_iterate(); trials += st.inner_trials
    dt = time.perf_counter() - t0
    value = args.steps / dt
    line = {
        "impl": "reference", "metric": "lm_iterations_per_sec", "value": value, "unit": "iter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "baseline_config": WORKLOADS[args.workload][1], "keyframes": g.K, "unknowns": g.nt * g.K + g.ns,
                   "init": "reference initial values (identity poses, zero vel/bias, statics at prior means)"},
        "lm_trials_per_sec": trials / dt, "lambda_trials": trials, "error_after": st.error,
        "cpu_baseline": {"value": value, "unit": "iter/s", "cores": threads, "kind": "port",
                         "sample": f"the full workload ({g.K} keyframes): {args.steps} LM iterations ({trials} lambda trials) after {args.warmup} warm-up, "
                                   f"oracle/fastsolve.hpp with {max(16, 2 * threads)} time chunks on {threads} threads"},
        "e2e": {"value": value, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline(args):
    """bounded CPU sample for the default line (rank 0, N=1): a few iterate() calls of the FULL workload (10-30 s of CPU work)"""
    o, g, threads = oracle_problem(args)
    o.lm_iterate()
    n, trials = 3, 0
    t0 = time.perf_counter()
    for _ in range(n):
        trials += o.lm_iterate().inner_trials
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "iter/s", "cores": threads, "kind": "port",
            "sample": f"the full workload ({g.K} keyframes): {n} LM iterations ({trials} lambda trials) after 1 warm-up, oracle/fastsolve.hpp on {threads} threads"}


def dmma_peak():
    """fp64 tensor-pipe (DMMA m8n8k4) peak measured live with tools/ubench/dmma_probe (register-only MMA loop on all 148 SMs)"""
    exe = os.path.join(ROOT, "tools", "ubench", "dmma_probe")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        best = 0.0
        for ln in out.splitlines():
            if ln.startswith("r1 dmma registers"):
                best = max(best, float(ln.split("TFLOP/s")[0].split()[-1]))
        if best > 0:
            return best, "tools/ubench/dmma_probe (register-only mma.sync.m8n8k4.f64 loop, this run)"
    except Exception:
        pass
    return 36.8, "fallback: profiles/r01c_dmma_probe.txt (not in MEASURED_PEAKS.json)"


def run_b200(args):
    import torch
    from bioslam_b200 import abi, lib
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bioslam_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    json_fd = 1
    if world > 1:
        # NCCL prints its version banner on the process's stdout (fd 1): keep fd 1 for the ONE JSON line of rank 0 and send
        # everything else that native code prints to stderr
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.load()
    preint = lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh, device=local)
    _, g, tv0, sv0 = build_problem_inputs(args.workload, preint)
    K = g.K
    p = lib.Problem(g, device=local, chunks=args.chunks)
    if world > 1:
        from bioslam_b200 import sharding
        # N GPUs = G lambda groups x T time shards: the groups try consecutive rungs of the LM lambda ladder concurrently
        # (exactly the sequential tryLambda results), each group shards the time axis over its T ranks
        # measured on 8 B200 (profiles/README.md): 4 x 2 136 iter/s, 2 x 4 130, 8 x 1 116; on 4: 2 x 2; on 2: 2 x 1 84 against 1 x 2 70
        lam_groups = args.lambda_groups if args.lambda_groups > 0 else (4 if world % 8 == 0 else (2 if world % 2 == 0 else 1))
        comm = sharding.attach(p, dist, local, lambda_groups=lam_groups) if args.comm == "torch" else sharding.attach_nccl(p, dist, lambda_groups=lam_groups)
    else:
        lam_groups, comm = 1, None
    lm = abi.LmParams.lower_body()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- device-resident arm (production path: lambda ladder on the device, one CUDA graph per trial)
    p.set_option("phase_timing", 0)
    p.set_values(tv0, sv0)
    p.lm_begin(lm)
    for _ in range(args.warmup):
        p.lm_iterate()
    barrier()
    sampler = ClockSampler(local); sampler.start()
    l0, s0 = p.launch_count(), p.host_sync_count()
    p.timer_begin()
    trials = 0
    for _ in range(args.steps):
        st = p.lm_iterate(); trials += st.inner_trials
    ms = p.timer_end()
    barrier()
    launches = p.launch_count() - l0
    syncs = p.host_sync_count() - s0
    clocks = sampler.summary()
    ms = max_over_ranks(ms)
    value = args.steps / (ms * 1e-3)
    err_after = st.error

    # ---------------- attribution pass: the SAME iterations again with per-phase CUDA events on the library's stream (direct
    # launches instead of graphs): launch duration of the dominant kernel for the roofline, time split per phase
    p.set_values(tv0, sv0)
    p.lm_begin(lm)
    p.set_option("phase_timing", 1)
    for _ in range(args.warmup):
        p.lm_iterate()
    p.phase_ms(reset=True)
    p.timer_begin()
    trials_attr = 0
    for _ in range(args.steps):
        trials_attr += p.lm_iterate().inner_trials
    ms_attr = max_over_ranks(p.timer_end())
    ph = p.phase_ms()
    p.set_option("phase_timing", 0)
    comm_ms = comm.take_ms() if hasattr(comm, "take_ms") else None

    # ---------------- end-to-end arm: host buffers in, host buffers out, every step
    # (the host buffers are PINNED: every step uploads the values from them and downloads the new values into them)
    tv = torch.empty(tv0.shape, dtype=torch.float64, pin_memory=True).numpy(); tv[...] = tv0
    sv_buf = torch.empty((max(sv0.size, 1),), dtype=torch.float64, pin_memory=True).numpy(); sv_buf[:sv0.size] = sv0
    sv = sv_buf[:sv0.size]
    p.set_values(tv, sv)
    p.lm_begin(lm)
    for _ in range(args.warmup):
        p.set_values(tv, sv); p.lm_iterate(); p.get_values(out=(tv, sv_buf))
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        p.set_values(tv, sv); st2 = p.lm_iterate(); p.get_values(out=(tv, sv_buf))
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    h2d = tv.nbytes + sv.nbytes
    d2h = tv.nbytes + sv.nbytes + 8 * 10

    # ---------------- time-to-converge (BASELINE.json metric, second half): the reference's own exit rule (6 consecutive
    # small iterations, src/lowerBodyPoseEstimator.cpp:656-663) from the reference's initial values, device-timed
    ttc = None
    if args.converge_max > 0:
        p.set_values(tv0, sv0)
        barrier()
        res, hist = p.optimize(lm, abi.OuterParams.lower_body(max_iterations=args.converge_max))
        barrier()
        ttc = {"seconds": res.device_ms * 1e-3, "iterations": int(res.iterations), "lambda_trials": int(res.total_trials),
               "initial_error": float(res.initial_error), "final_error": float(res.final_error),
               "converged": bool(res.iterations < args.converge_max), "rule": "6 consecutive iterations with abs<1e-6 or rel<1e-5 decrease",
               # the exit rule fires on a still-descending tail (relative decrease per iteration ~ the 1e-5 threshold): the cost AT
               # the stop is defined to ~1e-5 only; costs at fixed iteration counts are the partition-independent comparison
               "error_at_iteration": {str(i): float(hist[i]) for i in (100, 500, 1000, 1500, 2000) if i < len(hist)}}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = measured_peaks()
    P = p.chunks
    sweep_ms = ph[3] / max(ph[4], 1)
    units = K - (P - 1)
    if ph[4] > 0:
        kern = "chain_sweep_kernel<false> (level 0)"
    else:  # single chunk: the top-level sweep (+ its Schur / finish kernels) does everything
        sweep_ms = ph[5] / max(ph[6], 1)
        units = K
        kern = "chain_sweep_kernel (top level)"
    units_rank = units / (world // lam_groups)   # keyframes one launch of one rank sweeps
    achieved_tf = units_rank * SWEEP_FLOP_PER_KEYFRAME / (sweep_ms * 1e-3) / 1e12
    achieved_gbs = units_rank * SWEEP_BYTES_PER_KEYFRAME / (sweep_ms * 1e-3) / 1e9
    tpeak, tpeak_src = dmma_peak()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "sweep_traffic.json")
    ncu = {}
    if os.path.exists(tpath):
        ncu = json.load(open(tpath))
        traffic = ncu.get("dram_bytes_per_launch")
    solve_ms = ph[1] / max(trials_attr, 1)
    ntr = max(ph[6], 1)   # lambda trials solved by this rank's group in the attribution pass (one top-level launch each)
    line = {
        "metric": "lm_iterations_per_sec", "value": value, "unit": "iter/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "baseline_config": WORKLOADS[args.workload][1], "keyframes": K, "unknowns": g.nt * K + g.ns,
                   "chunks": P, "partition_levels": p.levels,
                   "parallelism": f"{lam_groups} lambda group(s) x {world // lam_groups} time shard(s)", "cache": "working set (>2 GB of Hessian/factor blocks at K=6000) exceeds the 126 MB L2; no flush needed",
                   "init": "reference initial values (identity poses, zero vel/bias, statics at prior means)"},
        "lm_trials_per_sec": trials / (ms * 1e-3), "lambda_trials": trials, "error_after": err_after,
        "host_syncs_per_step": syncs / args.steps,
        "attribution_pass": {"what": "the same iterations with per-phase CUDA events (direct launches, no graphs)", "ms_per_step": ms_attr / args.steps,
                             "lambda_trials": trials_attr,
                             "phase_ms_per_step": {"linearize+assemble": ph[0] / args.steps, "solve": ph[1] / args.steps, "model+retract+error": ph[2] / args.steps},
                             "solve_ms_per_trial": solve_ms, "sweep_level0_ms": sweep_ms, "top_level_ms": ph[5] / max(ph[6], 1),
                             "backsub_ms": ph[7] / max(ph[4], 1),
                             "solve_ms_per_trial_split": {"local_elimination": ph[8] / ntr, "level0_sweep": ph[3] / ntr,
                                                          "level0_schur+boundary_exchange+reduce": ph[9] / ntr, "level1": ph[10] / ntr,
                                                          "levels>=2": ph[11] / ntr, "top_level+static_system": ph[5] / ntr,
                                                          "back_substitution+step_exchange": ph[7] / ntr}},
        # N > 1: what does not shrink with the number of time shards (computed redundantly on every rank of a group) and the device time of the
        # collectives, per lambda trial solved by this rank's group (attribution pass, every exchange group bracketed by CUDA events)
        "replicated_ms_per_trial": (ph[11] + ph[5]) / ntr,
        "nccl_ms_per_trial": (comm_ms["device_ms"] / ntr) if comm_ms is not None else 0.0,
        "roofline": {"bound": "tensor", "kernel": kern, "achieved": achieved_tf, "peak": tpeak, "unit": "TFLOP/s", "frac": achieved_tf / tpeak,
                     "traffic": traffic, "peak_source": tpeak_src, "launch_ms": sweep_ms,
                     "flops_per_launch": units_rank * SWEEP_FLOP_PER_KEYFRAME,
                     "hbm": {"achieved": achieved_gbs, "peak": peak, "unit": "GB/s", "frac": achieved_gbs / peak, "peak_source": peak_src,
                             "bytes_per_launch": units_rank * SWEEP_BYTES_PER_KEYFRAME},
                     "algorithmic_solve": {"tflops": units_rank * SOLVE_FLOP_PER_KEYFRAME_ALGORITHMIC / (solve_ms * 1e-3) / 1e12,
                                           "frac": units_rank * SOLVE_FLOP_PER_KEYFRAME_ALGORITHMIC / (solve_ms * 1e-3) / 1e12 / tpeak,
                                           "what": "SURVEY 8(d) dense-block model, 5.0 MFLOP per keyframe and trial, over the whole solve"},
                     "ncu": {k: ncu.get(k) for k in ("dmma_pipe_pct_of_peak", "source") if k in ncu},
                     "note": "fp64 blocks of 106: ~25 flop per algorithmic byte, the kernel is bound by the fp64 tensor pipe (DMMA m8n8k4; tcgen05 has no "
                             "f64 kind) and by the latency of its sequential pivot chain, not by HBM; `achieved` = executed flops of the kernel "
                             "(formula in DESIGN.md section 3) / its launch time measured in this run with CUDA events; `hbm` = SURVEY 8(d) algorithmic bytes "
                             "(H 4167 + F 28539 doubles per keyframe) over the same time, reported because the contract asks for it"},
        "e2e": {"value": args.steps / e2e_s, "unit": "iter/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if comm_ms is not None:
        line["collectives"] = comm_ms
    if ttc is not None:
        line["time_to_converge"] = ttc
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args)
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


def run_batch(args):
    """BASELINE.json configs[4]: a batch of independent synthetic subjects (60 s each, K = 600, seeds 0..B-1), every one optimised to the
    reference's convergence rule (b200_optimize) from the reference's initial values.  SURVEY.md section 8(e): replicas only -- the
    subjects are dealt round-robin to the ranks (one process per GPU), no data-path collective; on a GPU `--concurrent` host threads
    each drive their own b200_problem on its own CUDA stream.  Host buffers in and out per solve (values uploaded, optimised values
    downloaded): `value` and `e2e` are the same measurement.  metric: converged solves per second, whole job."""
    import queue
    import torch
    from bioslam_b200 import abi, graphs, lib, synth
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bioslam_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    json_fd = 1
    if world > 1:
        sys.stdout.flush(); json_fd = os.dup(1); os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib.load()
    B = args.subjects
    mine = list(range(rank, B, world))
    pre = lambda a, w, dt, n, bh: lib.preintegrate(a, w, dt, n, bh, device=local)
    inputs = {i: graphs.lower_body_graph(synth.make_trial(seed=i, duration_s=60.0), preint=pre) for i in mine}
    lm, outer = abi.LmParams.lower_body(), abi.OuterParams.lower_body(max_iterations=args.converge_max if args.converge_max > 0 else 3000)
    g0, tv0, sv0 = next(iter(inputs.values()))
    for _ in range(max(args.warmup, 1)):   # warm-up: module load, graph instantiation, first-launch costs
        p = lib.Problem(g0, device=local); p.set_values(tv0, sv0); p.lm_begin(lm); p.lm_iterate(); p.close()
    results = {}
    work = queue.Queue()
    for i in mine: work.put(i)
    launches = [0]

    def worker():
        torch.cuda.set_device(local)
        while True:
            try:
                i = work.get_nowait()
            except queue.Empty:
                return
            g, tv, sv = inputs[i]
            # --chunks: level-0 partition per problem (upper levels from the cost model).  The library default spreads ONE problem over
            # all 148 SMs; problems that run side by side do better with 148 / concurrent chunks each (longer chains, fewer separators)
            p = lib.Problem(g, device=local, chunks=args.chunks, groups=0 if args.chunks else None); p.set_values(tv, sv)
            res, _ = p.optimize(lm, outer)
            tvf, svf = p.get_values()
            results[i] = (int(res.iterations), int(res.total_trials), float(res.final_error), float(res.device_ms), int(res.status))
            launches[0] += p.launch_count()
            p.close()

    if dist is not None: dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    t0 = time.perf_counter()
    ths = [threading.Thread(target=worker) for _ in range(args.concurrent)]
    for t in ths: t.start()
    for t in ths: t.join()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    clocks = sampler.summary()
    if dist is not None:
        tt = torch.tensor([wall], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); wall = float(tt.item())
        gathered = [None] * world
        dist.all_gather_object(gathered, results)
        results = {k: v for d in gathered for k, v in d.items()}
        dist.barrier()
    if rank == 0:
        g = g0
        its = [results[i][0] for i in range(B)]
        line = {"metric": "converged_solves_per_sec", "value": B / wall, "unit": "solves/s", "n_gpus": world, "steps": B, "warmup": max(args.warmup, 1),
                "ms_per_step": wall / B * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "batch_60s_100hz_7imu", "baseline_config": "configs[4]", "subjects": B, "keyframes": g.K, "unknowns": g.nt * g.K + g.ns,
                           "parallelism": f"replicas only: {world} rank(s) x {args.concurrent} concurrent problems (one CUDA stream each), no collective", "chunks_per_problem": args.chunks,
                           "rule": "6 consecutive iterations with abs<1e-6 or rel<1e-5 decrease (reference exit rule)",
                           "cache": "independent problems of ~0.25 GB each; concurrent working sets exceed L2"},
                "iterations": {"min": min(its), "max": max(its), "mean": float(np.mean(its))},
                "lambda_trials_total": int(sum(results[i][1] for i in range(B))), "all_converged": all(results[i][4] == 0 for i in range(B)),
                "mean_device_s_per_solve": float(np.mean([results[i][3] for i in range(B)])) * 1e-3,
                "final_error_range": [min(results[i][2] for i in range(B)), max(results[i][2] for i in range(B))],
                "e2e": {"value": B / wall, "unit": "solves/s", "h2d_bytes_per_step": int(tv0.nbytes + sv0.nbytes), "d2h_bytes_per_step": int(tv0.nbytes + sv0.nbytes)},
                "gpu_launches": int(launches[0]), "clocks": clocks}
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--lambda-groups", type=int, default=0, help="N > 1: groups of GPUs that try consecutive lambdas concurrently (0 = 2 when N is even)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="gait_10min_100hz_7imu", choices=sorted(WORKLOADS) + ["batch_60s_100hz_7imu"])
    ap.add_argument("--subjects", type=int, default=64, help="batch workload: number of independent subjects (BASELINE configs[4]: 64)")
    ap.add_argument("--concurrent", type=int, default=8, help="batch workload: problems in flight per GPU (host threads, one CUDA stream each)")
    ap.add_argument("--comm", default="nccl", choices=["nccl", "torch"], help="N > 1: NCCL inside the library (default) or torch.distributed callbacks")
    ap.add_argument("--chunks", type=int, default=None, help="solver partition (default: library heuristic)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--converge-max", type=int, default=3000, help="iteration cap of the time-to-converge run (0: skip it)")
    args = ap.parse_args()
    if args.workload == "batch_60s_100hz_7imu":
        if args.impl == "reference":
            raise SystemExit("the CPU arm of the batch workload is the single-subject CPU arm (--workload gait_60s_100hz_7imu) times the subject count")
        run_batch(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
