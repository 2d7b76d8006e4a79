"""Multi-rank plumbing of the path (one process per GPU, torch.distributed).

Servers shard over ranks by contiguous ranges; accelerator / perf / type tables are replicated.
Analyze needs no communication.  Optimize has ONE exchange step:

  * unlimited mode: each rank solves its own servers; the per-type totals of System.AllocateByType
    (reference pkg/core/system.go:271-300) are partial sums -> one all-reduce of {count[T], cost[T]};
  * limited mode: the greedy assignment is sequential over ALL servers (pkg/solver/greedy.go:107-166),
    so ranks all-gather their (server, accelerator) candidate rows and every rank runs the identical
    solve; totals are again reduced from per-shard partials.

The functions below are backend-agnostic (NCCL on GPUs, gloo in the CPU tests); the device path hands
torch views of the library's own device buffers to the collectives (no staging copies).
"""
import numpy as np
import torch
import torch.distributed as dist

from . import abi


def shard_range(n_servers, rank, world):
    """[first, first+count) of rank: floor(S*g/G) boundaries."""
    first = (n_servers * rank) // world
    last = (n_servers * (rank + 1)) // world
    return first, last - first


class _CudaView:
    """__cuda_array_interface__ wrapper of a raw device pointer owned by the C library."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3}


_TYPESTR = {np.dtype(np.int32): "<i4", np.dtype(np.int64): "<i8", np.dtype(np.float32): "<f4", np.dtype(np.uint8): "|u1"}


def device_tensor(ptr, n, dtype, device):
    return torch.as_tensor(_CudaView(ptr, n, _TYPESTR[np.dtype(dtype)]), device=device)


def allreduce_totals_device(ctx, n_types, device, stream=None):
    """The one collective of the unlimited path, in place on the library's totals buffer."""
    ptr, _ = ctx.type_totals_device()
    cnt = device_tensor(ptr, n_types, np.int64, device)
    cst = device_tensor(ptr + 8 * n_types, n_types, np.float32, device)
    if stream is not None:
        with torch.cuda.stream(stream):
            dist.all_reduce(cnt); dist.all_reduce(cst)
    else:
        dist.all_reduce(cnt); dist.all_reduce(cst)
    return cnt, cst


def allreduce_totals_host(count, cost):
    """Same reduction on host arrays (gloo)."""
    c = torch.from_numpy(np.ascontiguousarray(count)); k = torch.from_numpy(np.ascontiguousarray(cost))
    dist.all_reduce(c); dist.all_reduce(k)
    return c.numpy(), k.numpy()


def gather_pair_rows_host(pairs, feasible, n_servers, n_accels, world):
    """All-gather the S*A candidate records: every rank contributes the rows of its own shard.
    `pairs` is an abi.AllocArrays of full extent whose out-of-shard rows are ignored."""
    rank = dist.get_rank()
    first, count = shard_range(n_servers, rank, world)
    out = abi.AllocArrays(n_servers * n_accels)
    fields = [(name, getattr(pairs, name), getattr(out, name)) for name, _ in abi.ALLOC_FIELDS]
    fe_out = np.zeros(n_servers * n_accels, dtype=np.uint8)
    fields.append(("feasible", feasible, fe_out))
    bounds = [shard_range(n_servers, r, world) for r in range(world)]
    for _, src, dst in fields:
        # shards may differ by one server: one broadcast per owner (backend-agnostic, in place)
        for r, (f, c) in enumerate(bounds):
            part = torch.from_numpy(np.ascontiguousarray(src[f * n_accels:(f + c) * n_accels])) if r == rank \
                else torch.empty(c * n_accels, dtype=torch.from_numpy(src[:1]).dtype)
            if part.numel():
                dist.broadcast(part, src=r)
            dst[f * n_accels:(f + c) * n_accels] = part.numpy()
    return out, fe_out


def gather_pair_rows_device(ctx, n_servers, n_accels, world, device):
    """Limited mode on GPUs: all-gather the candidate rows in place in the library's device arrays
    (wva_pairs_device), then mark them complete (wva_pairs_commit)."""
    ptrs = ctx.pairs_device()
    n = n_servers * n_accels
    bounds = [shard_range(n_servers, r, world) for r in range(world)]
    rank = dist.get_rank()
    for name, (ptr, dt) in ptrs.items():
        full = device_tensor(ptr, n, dt, device)
        for r, (f, c) in enumerate(bounds):          # shards may differ by one server: broadcast per owner, in place
            if c:
                dist.broadcast(full[f * n_accels:(f + c) * n_accels], src=r)
    ctx.pairs_commit()
