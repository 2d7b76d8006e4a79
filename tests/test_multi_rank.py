"""CPU, world_size 2 over gloo: the multi-rank host logic (shard boundaries, all-gather layout of the
candidate rows, the per-type all-reduce) with the oracle standing in for the per-rank device work.
The GPU counterpart is tests/test_multi_gpu.py (needs 2 B200s)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import wva_import
    wva = wva_import.load()
    from inferno_autoscaler_b200 import distributed as D
    import oracle
    img = wva.synth.make_system(37, 3, seed=91, n_types=2, max_pair_batch=128)
    # capacities that bind (computed identically on every rank from the single-process oracle)
    pairs_all, feas_all, _ = oracle.analyze_pairs(img)
    acc_u, ch_u = oracle.solve(img, pairs_all, feas_all, unlimited=True)
    wva.synth.set_capacity_from_demand(img, ch_u.acc, ch_u.num_replicas, fraction=0.6)
    first, count = D.shard_range(img.S, rank, world)
    # this rank analyses only its shard (rows outside stay zero)
    shard = img.shard(first, count)
    p_sh, f_sh, _ = oracle.analyze_pairs(shard)
    mine = wva.abi.AllocArrays(img.S * img.A)
    fe = np.zeros(img.S * img.A, dtype=np.uint8)
    sl = slice(first * img.A, (first + count) * img.A)
    for name, _ in wva.abi.ALLOC_FIELDS:
        getattr(mine, name)[sl] = getattr(p_sh, name)
    fe[sl] = f_sh
    ok = True
    if mode == "unlimited":
        acc, chosen = oracle.solve(img, mine, fe, unlimited=True)        # rows of other shards are infeasible here
        cnt, cst = oracle.allocate_by_type(img, acc, chosen, first, first + count)
        cnt, cst = D.allreduce_totals_host(cnt, cst)
        w_cnt, w_cst = oracle.allocate_by_type(img, acc_u, ch_u)
        ok = np.array_equal(cnt, w_cnt) and np.allclose(cst, w_cst, rtol=1e-5)
        ok = ok and np.array_equal(acc[first:first + count], acc_u[first:first + count])
    else:
        full, fe_full = D.gather_pair_rows_host(mine, fe, img.S, img.A, world)
        same, field = full.equal_bits(pairs_all)
        ok = same and np.array_equal(fe_full, feas_all)
        acc, chosen = oracle.solve(img, full, fe_full, unlimited=False, policy=wva.abi.POLICY_PRIORITY_EXHAUSTIVE)
        w_acc, w_chosen = oracle.solve(img, pairs_all, feas_all, unlimited=False, policy=wva.abi.POLICY_PRIORITY_EXHAUSTIVE)
        ok = ok and np.array_equal(acc, w_acc) and chosen.equal_bits(w_chosen)[0]
        cnt, cst = oracle.allocate_by_type(img, acc, chosen, first, first + count)
        cnt, cst = D.allreduce_totals_host(cnt, cst)
        w_cnt, w_cst = oracle.allocate_by_type(img, w_acc, w_chosen)
        ok = ok and np.array_equal(cnt, w_cnt) and np.allclose(cst, w_cst, rtol=1e-5)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["unlimited", "limited"])
def test_two_ranks_gloo(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (0 if mode == "unlimited" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_range_partitions(wva):
    from inferno_autoscaler_b200 import distributed as D
    for S in (0, 1, 7, 32, 1000):
        for G in (1, 2, 3, 8):
            spans = [D.shard_range(S, g, G) for g in range(G)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == S
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(G - 1))
