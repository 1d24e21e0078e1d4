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


def attach(problem, dist, cuda_device):
    """register the collectives for `problem` (bioslam_b200.lib.Problem); cuda_device=None for the CPU (gloo) tests"""
    world, rank = dist.get_world_size(), dist.get_rank()

    def on_stream(stream_ptr):
        if cuda_device is None:
            import contextlib
            return contextlib.nullcontext()
        return torch.cuda.stream(torch.cuda.ExternalStream(stream_ptr or 0, device=torch.device("cuda", cuda_device)))

    def allgather(ctx, buf, bytes_per_rank, stream):
        try:
            t = _view(buf, bytes_per_rank * world, cuda_device)
            n = bytes_per_rank // 8
            with on_stream(stream):
                if cuda_device is None:
                    parts = [torch.empty(n, dtype=torch.float64) for _ in range(world)]
                    dist.all_gather(parts, t[rank * n:(rank + 1) * n].clone())
                    for i, pt in enumerate(parts):
                        t[i * n:(i + 1) * n] = pt
                else:
                    dist.all_gather_into_tensor(t, t[rank * n:(rank + 1) * n])
            return 0
        except Exception as e:  # noqa: BLE001 -- must not propagate through the C ABI
            print("bioslam_b200.sharding allgather failed:", e, flush=True)
            return 1

    def allreduce(ctx, buf, n, stream):
        try:
            t = _view(buf, n * 8, cuda_device)
            with on_stream(stream):
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return 0
        except Exception as e:  # noqa: BLE001
            print("bioslam_b200.sharding allreduce failed:", e, flush=True)
            return 1

    problem.comm_set(world, rank, allgather, allreduce)
    return problem
