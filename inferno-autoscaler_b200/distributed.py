"""Multi-rank plumbing of the path (one process per GPU, torch.distributed).

Servers shard over ranks by contiguous ranges; accelerator / perf / type tables are replicated.
Analyze needs no communication.  Optimize has ONE exchange step:

  * unlimited mode: each rank solves its own servers; the per-type totals of System.AllocateByType
    (reference pkg/core/system.go:271-300) are partial sums -> one all-gather of {count[T], cost[T]} and a sum in rank order (TotalsExchange);
  * limited mode: the greedy assignment is sequential over ALL servers (pkg/solver/greedy.go:107-166),
    so ranks all-gather their (server, accelerator) candidate rows and every rank runs the identical
    solve; totals are again reduced from per-shard partials.

The collective itself lives INSIDE the library (wva_comm_init: ncclAllGather issued from C on the ctx
stream, include/wva_b200.h "multi-GPU inside the library"); `attach_library_comm` only distributes the
128-byte NCCL id over the process group the launcher already made.  The other functions are the
host-driven variant (backend-agnostic: NCCL on GPUs, gloo in the CPU tests) kept for the CPU tests of the
shard / gather / reduction logic and as a cross-check of the in-library exchange; they hand torch views of
the library's own device buffers to the collectives (no staging copies) and enqueue every collective on
the library's stream, so the library's next kernel is ordered after it.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import abi


def attach_library_comm(ctx, device=None):
    """Give `ctx` an NCCL communicator spanning the torch.distributed world: rank 0 creates the id
    (wva_comm_unique_id), the process group broadcasts its 128 bytes, every rank calls wva_comm_init.
    After this wva_allocate_by_type returns global totals and the limited solve gathers by itself."""
    from . import binding
    rank, world = dist.get_rank(), dist.get_world_size()
    if dist.get_backend() == "nccl":
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        t = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(binding.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=0)
        uid = bytes(t.cpu().numpy().tobytes())
    else:
        box = [binding.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    ctx.comm_init(uid, rank, world)
    return rank, world


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


class TotalsExchange:
    """The one collective of the unlimited path: every rank all-gathers its 12*T-byte block of
    partial totals straight from the library's device buffer, then the library sums the blocks in
    rank order (wva_type_totals_merge) back into that buffer.  Views and the gather buffer are
    created once."""

    def __init__(self, ctx, n_types, device):
        # the collective is enqueued on the library's own stream, where the partials were produced
        # and where the merge kernel runs (ordering contract of wva_type_totals_merge)
        self.ctx, self.T = ctx, int(n_types)
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=device)
        self.world = dist.get_world_size()
        ptr, nbytes = ctx.type_totals_device()
        assert nbytes == 12 * self.T
        self.ptr = ptr
        self.mine = device_tensor(ptr, nbytes, np.uint8, device)
        self.gathered = torch.empty(self.world * nbytes, dtype=torch.uint8, device=device)
        self.count = device_tensor(ptr, self.T, np.int64, device)
        self.cost = device_tensor(ptr + 8 * self.T, self.T, np.float32, device)

    def __call__(self, sync=True):
        ptr, _ = self.ctx.type_totals_device()
        assert ptr == self.ptr, "the library moved its totals buffer"
        with torch.cuda.stream(self.stream):
            dist.all_gather_into_tensor(self.gathered, self.mine)
        self.ctx.type_totals_merge(self.gathered.data_ptr(), self.world)
        if sync:
            self.stream.synchronize()
        return self.count, self.cost


def allreduce_totals_device(ctx, n_types, device):
    """One-shot form of TotalsExchange."""
    return TotalsExchange(ctx, n_types, device)()


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
    """Limited mode on GPUs, host-driven variant: all-gather the candidate rows in place in the library's
    device arrays (wva_pairs_device), then mark them complete (wva_pairs_commit).  The collectives run on
    the library's stream (and wva_pairs_commit waits for the device), so the greedy kernels that follow can
    never see half-gathered rows."""
    with torch.cuda.stream(torch.cuda.ExternalStream(ctx.stream(), device=device)):
        _gather_pair_rows_device(ctx, n_servers, n_accels, world, device)
    ctx.pairs_commit()


def _gather_pair_rows_device(ctx, n_servers, n_accels, world, device):
    ptrs = ctx.pairs_device()
    n = n_servers * n_accels
    bounds = [shard_range(n_servers, r, world) for r in range(world)]
    rank = dist.get_rank()
    even = len({c for _, c in bounds}) == 1
    f0, c0 = bounds[rank]
    for name, (ptr, dt) in ptrs.items():
        full = device_tensor(ptr, n, dt, device)
        if even:                                     # equal shards: one in-place all-gather per field
            dist.all_gather_into_tensor(full, full[f0 * n_accels:(f0 + c0) * n_accels])
        else:
            for r, (f, c) in enumerate(bounds):      # shards differ by one server: broadcast per owner, in place
                if c:
                    dist.broadcast(full[f * n_accels:(f + c) * n_accels], src=r)
