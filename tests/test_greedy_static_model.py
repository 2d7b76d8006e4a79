"""CPU: the static-order form of `allocate` that k_greedy_scan runs (csrc/wva_greedy_scan.cuh) against a literal
restatement of the reference's sorted slice (greedy.go:107-166) on random cases with heavy key ties."""
import os
import random
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import greedy_static_model as gm


def test_static_order_equals_queue():
    rng = random.Random(11)
    for r in range(4000):
        servers, avail = gm.random_case(rng)
        assert gm.queue_allocate(servers, avail) == gm.static_allocate(servers, avail), (servers, avail)


def test_static_order_all_keys_tie():
    rng = random.Random(12)
    for r in range(300):
        S = rng.randint(2, 30)
        servers = [(1, [(0, rng.randrange(2), rng.randint(1, 5), False) for _ in range(rng.randint(1, 4))]) for _ in range(S)]
        avail = [rng.randint(0, 30), rng.randint(0, 30)]
        assert gm.queue_allocate(servers, avail) == gm.static_allocate(servers, avail)


def test_batch_resolution_equals_one_by_one():
    """the warp's lane-parallel resolution of 32 events (register copy of the type's capacity, successes applied in lane
    order) against processing them one by one; counts may be negative (Go's wrapped product) so capacities may grow"""
    rng = random.Random(13)
    for r in range(6000):
        ev, avail, done = gm.random_batch(rng)
        a1, d1 = list(avail), set(done)
        a2, d2 = list(avail), set(done)
        assert (gm.batch_resolve(ev, a1, d1), a1, d1) == (gm.sequential_resolve(ev, a2, d2), a2, d2), (ev, avail, done)


def test_allocate_maximally_batched_equals_sequential():
    """greedy_allocate_maximally serves 32 servers per step (lane = server, forward-only cursor): same takes as
    greedy.go:194-223 one server after the other"""
    rng = random.Random(14)
    for r in range(4000):
        servers, avail = gm.random_maximally(rng)
        assert gm.maximally_sequential(servers, avail) == gm.maximally_batched(servers, avail), (servers, avail)


def test_allocate_equally_round1_batched_equals_sequential():
    rng = random.Random(15)
    for r in range(4000):
        servers, avail = gm.random_maximally(rng)
        assert gm.equally_round1_sequential(servers, avail) == gm.equally_round1_batched(servers, avail), (servers, avail)
