"""-m "not gpu": the N > 1 path (time axis sharded over ranks) with world_size 2 and 3 over gloo on the CPU, each rank
running the CUDA sources under the emulator (tests/cudaemu, TEST TOOL ONLY): every rank must end up with the same
values as the single-rank run, and with the oracle's accept/reject sequence."""
import os
import subprocess
import sys
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

WORKER = r'''
import os, sys, json
import numpy as np
sys.path.insert(0, %(root)r); sys.path.insert(0, %(here)r)
import torch, torch.distributed as dist
from bioslam_b200 import abi, lib, synth, sharding
import graphs
rank, world = int(sys.argv[1]), int(sys.argv[2])
os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = sys.argv[3]
dist.init_process_group("gloo", rank=rank, world_size=world)
EMU = os.path.join(%(here)r, "cudaemu", "libbioslam_b200_emu.so")
tr = synth.make_trial(seed=3, duration_s=float(sys.argv[4]))
g, tv, sv = graphs.lower_body_graph(tr)
tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
p = lib.Problem(g, so_path=EMU)
lambda_groups = int(sys.argv[6]) if len(sys.argv) > 6 else 1
if len(sys.argv) > 7 and sys.argv[7] == "reference_init":
    g, tv, sv = graphs.lower_body_graph(tr)      # the reference's initial values: the first iterations reject lambda trials
sharding.attach(p, dist, None, lambda_groups=lambda_groups)
p.set_values(tv, sv)
out = {"chunks": p.chunks, "groups": p.groups}
e0 = p.lm_begin(abi.LmParams.lower_body())
trace = []
err_before = e0
for _ in range(int(os.environ.get("B200_TEST_ITERS", "3"))):
    st = p.lm_iterate()
    # linear.error(0) summed over the ranks must be the cost at the linearization point: every factor counted exactly once
    # (a rank also linearizes the keyframe left of its window; its factors belong to the neighbour's cost)
    lin0 = p.last_trial_scalars()[0]
    assert abs(lin0 - err_before) <= 1e-12 * err_before, (lin0, err_before)
    err_before = st.error
    trace.append([st.error, st.lambda_, st.inner_trials, st.iterations])
tv2, sv2 = p.get_values()
np.savez(sys.argv[5] + f".rank{rank}.npz", e0=e0, trace=np.array(trace), tv=tv2, sv=sv2, chunks=p.chunks, groups=p.groups)
dist.destroy_process_group()
'''


def run_world(world, seconds, tmp_path, port, extra=(), tag=""):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT, "here": HERE})
    prefix = str(tmp_path / f"w{world}{tag}")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(world), str(port), str(seconds), prefix, *extra]) for r in range(world)]
    for pr in procs:
        assert pr.wait(timeout=600) == 0
    return [np.load(prefix + f".rank{r}.npz") for r in range(world)]


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_lm_matches_single_rank_and_oracle(world, tmp_path, oracle_lib):
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cudaemu")])
    seconds = 4.0
    outs = run_world(world, seconds, tmp_path, 29500 + world)
    # all ranks agree bit-for-bit (replicated values, redundant top-level solve)
    for o in outs[1:]:
        assert np.array_equal(o["tv"], outs[0]["tv"]) and np.array_equal(o["sv"], outs[0]["sv"])
        assert np.array_equal(o["trace"], outs[0]["trace"])
    assert int(outs[0]["groups"]) % world == 0 and int(outs[0]["chunks"]) % world == 0
    # the single-rank run of the same library on the same inputs: same accept / reject sequence and damping, costs and values to what
    # two partitions of the same damped systems (condition numbers up to 1e16 at lambda <= 1e-5) can agree
    one = run_world(1, seconds, tmp_path, 29520 + world, tag="single")[0]
    assert np.array_equal(one["trace"][:, 1:], outs[0]["trace"][:, 1:])
    assert np.allclose(one["trace"][:, 0], outs[0]["trace"][:, 0], rtol=5e-3, atol=0.0)
    assert float(one["e0"]) == float(outs[0]["e0"]) or abs(float(one["e0"]) - float(outs[0]["e0"])) <= 1e-12 * float(one["e0"])
    # oracle on the same inputs
    from bioslam_b200 import abi, synth
    from oracle import pyoracle as po
    import graphs
    tr = synth.make_trial(seed=3, duration_s=seconds)
    g, tv, sv = graphs.lower_body_graph(tr)
    tv, sv = graphs.perturbed_truth_values(tr, g, tv, sv)
    o = po.OracleProblem(g, threads=4); o.set_values(tv, sv)
    e0 = o.lm_begin(abi.LmParams.lower_body())
    assert abs(e0 - float(outs[0]["e0"])) <= 1e-11 * e0
    for row in outs[0]["trace"]:
        st = o.lm_iterate()
        assert st.inner_trials == int(row[2]) and st.iterations == int(row[3])
        assert abs(st.error - row[0]) <= 5e-3 * st.error


@pytest.mark.parametrize("world,groups", [(2, 2), (4, 2)])
def test_lambda_speculation_equals_sequential_ladder(world, groups, tmp_path):
    """G groups try consecutive rungs of the lambda ladder at once: accepted steps, lambda and trial counts must be those of the
    sequential tryLambda loop run by ONE group (bit-identical: same kernels, same inputs), from the reference's initial values
    where trials are rejected."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "cudaemu")])
    seconds = 2.0
    os.environ["B200_TEST_ITERS"] = "3"
    try:
        spec = run_world(world, seconds, tmp_path, 29600 + world, extra=(str(groups), "reference_init"), tag="spec")
        seq = run_world(world // groups, seconds, tmp_path, 29650 + world, extra=("1", "reference_init"), tag="seq")
    finally:
        del os.environ["B200_TEST_ITERS"]
    trials = [int(row[2]) for row in seq[0]["trace"]]
    assert any(t > 1 for t in trials), "test needs rejected trials (another group's step is adopted)"
    assert any(t % 2 == 1 for t in trials), "test needs an iteration that group 0 wins"
    for o in spec:
        assert np.array_equal(o["trace"], seq[0]["trace"])
        assert np.array_equal(o["tv"], seq[0]["tv"]) and np.array_equal(o["sv"], seq[0]["sv"])
