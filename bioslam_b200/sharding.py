"""Multi-GPU plumbing: one process per GPU, time axis sharded into contiguous windows.

The CUDA library delegates its exchange steps (all-gather of the window-boundary Schur complements / group packages /
steps, all-reduce of the three cost scalars) to host-supplied collectives on DEVICE pointers; here they are
torch.distributed collectives (NCCL over NVLink on the GPU box, gloo in the CPU tests) issued on the library's stream.
"""
import ctypes as C
import numpy as np
import torch


def _view(ptr, nbytes, cuda_device):
    """zero-copy torch view (float64) of raw memory owned by the library"""
    n = nbytes // 8
    if cuda_device is None:
        return torch.from_numpy(np.ctypeslib.as_array((C.c_double * n).from_address(ptr)))

    class _Raw:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(_Raw(), device=torch.device("cuda", cuda_device))


class Comm:
    """what attach() returns: how the collectives run and how long they took (device ms, measured while phase timing is on)"""

    def __init__(self, problem, mode, world, groups):
        self.problem, self.mode, self.world, self.groups = problem, mode, world, groups

    def take_ms(self):
        ms, calls = self.problem.comm_ms(reset=True)
        return {"mode": self.mode, "device_ms": ms, "groups_issued": calls}


def attach_nccl(problem, dist, lambda_groups=1):
    """production path: NCCL inside the library (b200_nccl_init).  torch.distributed only carries the 128-byte unique id."""
    from . import lib
    world_all, rank_all = dist.get_world_size(), dist.get_rank()
    box = [lib.nccl_unique_id() if rank_all == 0 else None]
    dist.broadcast_object_list(box, src=0)
    problem.nccl_init(world_all, rank_all, box[0], lambda_groups)
    return Comm(problem, "in-library NCCL (ncclGroup per exchange step)", world_all, lambda_groups)


def attach(problem, dist, cuda_device, lambda_groups=1):
    """register the collectives for `problem` (bioslam_b200.lib.Problem); cuda_device=None for the CPU (gloo) tests.

    lambda_groups = G > 1 splits the world into G groups of world/G consecutive ranks: each group is one time-sharded solver
    and the groups try consecutive rungs of the LM lambda ladder concurrently (b200_spec_set); every rank must call this."""
    world_all, rank_all = dist.get_world_size(), dist.get_rank()
    G = int(lambda_groups)
    if G < 1 or world_all % G:
        raise ValueError("lambda_groups must divide the world size")
    world = world_all // G
    group_idx, rank = divmod(rank_all, world)
    pg = None
    if G > 1:
        pgs = [dist.new_group(list(range(h * world, (h + 1) * world))) for h in range(G)]   # collective: all ranks create all groups
        pg = pgs[group_idx]

    def on_stream(stream_ptr):
        if cuda_device is None:
            import contextlib
            return contextlib.nullcontext()
        return torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr or 0, device=torch.device("cuda", cuda_device)))

    def make_allgather(n_ranks, my_rank, group):
        def allgather(ctx, buf, bytes_per_rank, stream):
            try:
                t = _view(buf, bytes_per_rank * n_ranks, cuda_device)
                n = bytes_per_rank // 8
                with on_stream(stream):
                    if cuda_device is None:
                        parts = [torch.empty(n, dtype=torch.float64) for _ in range(n_ranks)]
                        dist.all_gather(parts, t[my_rank * n:(my_rank + 1) * n].clone(), group=group)
                        for i, pt in enumerate(parts):
                            t[i * n:(i + 1) * n] = pt
                    else:
                        dist.all_gather_into_tensor(t, t[my_rank * n:(my_rank + 1) * n], group=group)
                return 0
            except Exception as e:  # noqa: BLE001 -- must not propagate through the C ABI
                print("bioslam_b200.sharding allgather failed:", e, flush=True)
                return 1
        return allgather

    def allreduce(ctx, buf, n, stream):
        try:
            t = _view(buf, n * 8, cuda_device)
            with on_stream(stream):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=pg)
            return 0
        except Exception as e:  # noqa: BLE001
            print("bioslam_b200.sharding allreduce failed:", e, flush=True)
            return 1

    problem.comm_set(world, rank, make_allgather(world, rank, pg), allreduce)
    if G > 1:
        problem.spec_set(G, group_idx, make_allgather(world_all, rank_all, None))
    return Comm(problem, "torch.distributed callbacks", world_all, G)
