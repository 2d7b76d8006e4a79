"""GPU, 2 ranks over NCCL (skipped on boxes with one GPU): servers shard over GPUs, the per-type totals
are all-reduced in place on the library's device buffer, limited mode all-gathers the candidate rows
in place; every rank must reproduce the single-rank oracle."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import wva_import
    wva = wva_import.load()
    from inferno_autoscaler_b200 import binding, distributed as D
    import oracle
    dev = torch.device("cuda", rank)
    img = wva.synth.make_system(61, 4, seed=93, n_types=2, max_pair_batch=128)
    pairs_all, feas_all, _ = oracle.analyze_pairs(img, threads=4)
    acc_u, ch_u = oracle.solve(img, pairs_all, feas_all, unlimited=True)
    wva.synth.set_capacity_from_demand(img, ch_u.acc, ch_u.num_replicas, fraction=0.6)
    ctx = binding.Context(rank)
    first, count = D.shard_range(img.S, rank, world)
    ok = True
    # unlimited: shard-local solve + one all-reduce
    ctx.upload(img); ctx.set_shard(first, count)
    ctx.analyze_pairs(download=False)
    acc, chosen = ctx.solve(unlimited=True)
    ctx.allocate_by_type()
    cnt, cst = D.allreduce_totals_device(ctx, img.T, dev)
    torch.cuda.synchronize()
    w_cnt, w_cst = oracle.allocate_by_type(img, acc_u, ch_u)
    ok = ok and np.array_equal(cnt.cpu().numpy(), w_cnt) and np.allclose(cst.cpu().numpy(), w_cst, rtol=1e-5)
    ok = ok and np.array_equal(acc[first:first + count], acc_u[first:first + count])
    # limited: gather rows in place, identical greedy everywhere
    ctx.upload(img); ctx.set_shard(first, count)
    ctx.analyze_pairs(download=False)
    D.gather_pair_rows_device(ctx, img.S, img.A, world, dev)
    torch.cuda.synchronize()
    acc, chosen = ctx.solve(unlimited=False, policy=wva.abi.POLICY_PRIORITY_ROUND_ROBIN)
    w_acc, w_chosen = oracle.solve(img, pairs_all, feas_all, unlimited=False, policy=wva.abi.POLICY_PRIORITY_ROUND_ROBIN)
    ok = ok and np.array_equal(acc, w_acc) and chosen.equal_bits(w_chosen)[0]
    ctx.allocate_by_type()
    cnt, cst = D.allreduce_totals_device(ctx, img.T, dev)
    torch.cuda.synchronize()
    w_cnt, w_cst = oracle.allocate_by_type(img, w_acc, w_chosen)
    ok = ok and np.array_equal(cnt.cpu().numpy(), w_cnt) and np.allclose(cst.cpu().numpy(), w_cst, rtol=1e-5)
    q.put((rank, bool(ok)))
    ctx.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpus_nccl():
    c = mp.get_context("spawn")
    q = c.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [c.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
