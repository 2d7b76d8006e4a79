"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle on the same
seeded inputs.  Bar: bit-exact on every integer and on every float32 (compared through their bit
patterns).  Run on the B200 box with `pytest -m gpu`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32


def _assert_allocs_equal(got, want, mask=None):
    ok, field = got.equal_bits(want, mask)
    assert ok, "field %s differs" % field


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def test_division_selftest(ctx):
    """div_hoisted(a, b, rcp_refined(b)) == a / b bitwise (the hoisted-reciprocal division of the chain)."""
    for mode in (0, 1, 2):
        assert ctx.selftest_division(1234 + mode, 200_000_000, mode) == 0, mode
    # float32: the hoisted-reciprocal division of the sweep's epilogue (rows share the divisor lambda, pairs share the
    # EffectiveConcurrency denominator) against the plain operator: inside its fast window and on arbitrary bit patterns
    for mode in (3, 4):
        assert ctx.selftest_division(4321 + mode, 2_000_000_000, mode) == 0, mode


def test_queue_analyze_matches_oracle(wva, oracle, ctx):
    rng = np.random.default_rng(11)
    n = 3000
    cfg = np.zeros(n, dtype=wva.abi.QUEUE_CONFIG_DTYPE)
    cfg["max_batch_size"] = rng.integers(1, 200, n)
    cfg["max_queue_size"] = cfg["max_batch_size"] * rng.integers(0, 12, n)
    cfg["alpha"] = rng.uniform(1, 30, n); cfg["beta"] = rng.uniform(0.01, 1, n)
    cfg["gamma"] = rng.uniform(0, 250, n); cfg["delta"] = rng.uniform(0, 0.1, n)
    cfg["avg_input_tokens"] = rng.integers(0, 3000, n); cfg["avg_output_tokens"] = rng.integers(1, 600, n)
    # a few invalid configs and edge cases
    cfg["max_batch_size"][:5] = [0, -1, 1, 1, 2]
    cfg["max_queue_size"][:5] = [4, 4, 0, 1, -1]
    cfg["avg_output_tokens"][5:8] = [0, 1, 1]; cfg["avg_input_tokens"][5:8] = [10, 0, 5]
    _, hi = np.zeros(n, F), np.zeros(n, F)
    # rates spread over (0, 1.1*max] plus non-positive ones
    a = oracle.queue_analyze(cfg, np.full(n, 1e-3, F))[0]
    rates = (a["max_rate"] * rng.uniform(0.0005, 1.1, n)).astype(F)
    rates[10:14] = [0.0, -1.0, 1e-6, 1e9]
    want_m, want_s = oracle.queue_analyze(cfg, rates)
    got_m, got_s = ctx.queue_analyze(cfg, rates)
    assert np.array_equal(got_s, want_s)
    assert got_m.tobytes() == want_m.tobytes()
    assert (want_s == 0).sum() > n // 2


def test_queue_size_matches_oracle(wva, oracle, ctx):
    rng = np.random.default_rng(12)
    n = 600
    cfg = np.zeros(n, dtype=wva.abi.QUEUE_CONFIG_DTYPE)
    cfg["max_batch_size"] = rng.integers(1, 96, n)
    cfg["max_queue_size"] = cfg["max_batch_size"] * 10
    cfg["alpha"] = rng.uniform(1, 30, n); cfg["beta"] = rng.uniform(0.01, 1, n)
    cfg["gamma"] = rng.uniform(0, 250, n); cfg["delta"] = rng.uniform(0, 0.1, n)
    cfg["avg_input_tokens"] = rng.integers(0, 3000, n); cfg["avg_output_tokens"] = rng.integers(1, 600, n)
    targets = np.stack([rng.choice([0.0, 50.0, 500.0, 2000.0], n), rng.choice([0.0, 5.0, 24.0, 80.0, 200.0], n),
                        rng.choice([0.0, 0.0, 100.0], n)], axis=1).astype(F)
    targets[:3] = [[-1, 5, 0], [50, -1, 0], [50, 5, -1]]
    want = oracle.queue_size(cfg, targets)
    got = ctx.queue_size(cfg, targets)
    assert np.array_equal(got[3], want[3])
    assert _bits(got[0]).tobytes() == _bits(want[0]).tobytes()
    assert got[1].tobytes() == want[1].tobytes()
    assert _bits(got[2]).tobytes() == _bits(want[2]).tobytes()
    assert (want[3] == 0).sum() > n // 4


@pytest.mark.parametrize("warp_max", [0, 1 << 22])
@pytest.mark.parametrize("seed,S,A", [(21, 300, 4), (22, 64, 8)])
def test_pairs_match_oracle(wva, oracle, ctx, seed, S, A, warp_max):
    """both pair kernels: one thread per pair (warp_max 0) and one warp per pair with speculative bisection"""
    img = wva.synth.make_system(S, A, seed=seed, n_types=max(1, A // 2), max_pair_batch=512)
    ctx.pairs_set_warp_max(warp_max)
    ctx.upload(img)
    got, gfe = ctx.analyze_pairs()
    ctx.pairs_set_warp_max(1 << 22)
    want, wfe, steps = oracle.analyze_pairs(img, threads=oracle.hardware_threads())
    assert np.array_equal(gfe, wfe)
    _assert_allocs_equal(got, want)
    assert wfe.sum() > S * A // 3
    if warp_max == 0:
        assert ctx.pair_steps() <= steps      # truncation never does more work than the reference


def test_pairs_golden_config1(wva, oracle, ctx):
    """BASELINE config 1 and the survey's derived vectors through the CUDA path."""
    img = wva.synth.config1()
    ctx.upload(img)
    got, fe = ctx.analyze_pairs()
    assert fe[0] == 1 and got.num_replicas[0] == 5 and got.batch_size[0] == 512 and got.cost[0] == F(160.0)
    assert got.itl[0] == F(23.991173) and got.ttft[0] == F(243.6568) and got.rho[0] == F(0.012853656)
    for rpm, mb, rep in [(60.0, 512, 1), (6000.0, 512, 50), (600.0, 8, 7), (600.0, 64, 5), (600.0, 256, 5)]:
        img.srv_arrival_rpm[0] = rpm; img.srv_max_batch[0] = mb
        ctx.upload(img)
        got, fe = ctx.analyze_pairs()
        assert fe[0] == 1 and got.num_replicas[0] == rep
    img.srv_arrival_rpm[0] = 600.0; img.srv_max_batch[0] = 1
    ctx.upload(img)
    assert ctx.analyze_pairs()[1][0] == 0


def test_pairs_iteration_budget(wva, oracle, ctx):
    """a bisection that never meets the 1e-6 tolerance runs all 100 iterations (the survey's H100 case:
    122 Solves); the speculative walk must stop at exactly the reference's iteration."""
    c1 = wva.synth.config1()
    c1.srv_slo_itl[0] = 10.0; c1.srv_slo_ttft[0] = 1000.0
    c1.perf_alpha[0], c1.perf_beta[0], c1.perf_gamma[0], c1.perf_delta[0] = 7.470, 0.044, 15.415, 0.000337
    c1.acc_cost[0] = 100.0; c1.srv_arrival_rpm[0] = 6000.0
    want, wfe, _ = oracle.analyze_pairs(c1)
    for warp_max, pstore in ((0, 0), (1 << 22, 0), (1 << 22, 1)):
        ctx.pairs_set_warp_max(warp_max); ctx.pairs_set_pstore(pstore)
        ctx.upload(c1)
        got, gfe = ctx.analyze_pairs()
        assert np.array_equal(gfe, wfe) and got.num_replicas[0] == 3
        _assert_allocs_equal(got, want)
    ctx.pairs_set_warp_max(1 << 22); ctx.pairs_set_pstore(0)


def test_pairs_overflow_rescale_path(wva, oracle, ctx):
    """Chains that overflow float64 take the materialised-p[] path (mm1modelstatedependent.go:84-89,
    :96-104): tiny beta/delta with a large batch make p[n] grow like N^n/n!."""
    img = wva.synth.make_system(6, 2, seed=5, max_pair_batch=0)
    img.perf_alpha[:] = 20.0; img.perf_beta[:] = 1e-4; img.perf_gamma[:] = 10.0; img.perf_delta[:] = 1e-7
    img.perf_valid[:] = 1; img.srv_target_valid[:] = 1
    img.srv_max_batch[:] = [900, 1500, 2500, 1200, 3000, 2000]
    img.srv_arrival_rpm[:] = [3e5, 6e5, 2e6, 1e3, 5e6, 9e5]
    img.srv_in_tokens[:] = 64; img.srv_out_tokens[:] = 8
    img.srv_slo_itl[:] = 200.0; img.srv_slo_ttft[:] = 2000.0; img.srv_slo_tps[:] = 0.0
    img.srv_keep_acc[:] = 0
    ctx.upload(img)
    got, gfe = ctx.analyze_pairs()
    want, wfe, _ = oracle.analyze_pairs(img)
    assert np.array_equal(gfe, wfe)
    _assert_allocs_equal(got, want)
    assert wfe.sum() >= 6


@pytest.mark.parametrize("mode", [1, 0, 3, 5, 9])   # auto / exact chains only / certified per candidate / thread per row / warp per row
@pytest.mark.parametrize("seed,S,A,R,B", [(31, 6, 3, 8, 48), (32, 3, 2, 5, 70), (33, 2, 1, 64, 33), (35, 4, 2, 6, 130)])
def test_grid_matches_oracle(wva, oracle, ctx, seed, S, A, R, B, mode):
    img = wva.synth.make_system(S, A, seed=seed, zero_load_fraction=0.0)
    img.srv_arrival_rpm[:] = np.maximum(img.srv_arrival_rpm, 30.0)
    ctx.set_certified_tails(mode)
    ctx.upload(img)
    best, cube, status = ctx.analyze_grid(R, B, want_cube=True)
    ctx.set_certified_tails(1)
    w_best, w_cube, w_status, steps = oracle.analyze_grid(img, R, B, threads=oracle.hardware_threads())
    assert np.array_equal(status, w_status)
    assert cube.tobytes() == w_cube.tobytes()
    assert best.tobytes() == w_best.tobytes()
    c = ctx.grid_counters()
    assert c["candidates_ok"] == int(((w_status & 0xfe) == 0).sum())
    assert c["steps_algorithmic"] == steps and c["steps_executed"] <= steps * 1.05   # deferred chains restart from scratch
    assert (w_status & 1).sum() > 0 and (w_best["acc"] >= 0).any()


def test_grid_config1(wva, oracle, ctx):
    """BASELINE config 1: 1 model, 1 accelerator, replicas 1-8, batch 1-256."""
    img = wva.synth.config1()
    ctx.upload(img)
    best, cube, status = ctx.analyze_grid(8, 256, want_cube=True)
    w_best, w_cube, w_status, _ = oracle.analyze_grid(img, 8, 256, threads=oracle.hardware_threads())
    assert np.array_equal(status, w_status) and cube.tobytes() == w_cube.tobytes() and best.tobytes() == w_best.tobytes()
    assert best["acc"][0] == 0 and best["replicas"][0] >= 1


def test_grid_edge_servers(wva, oracle, ctx):
    """zero load, out=1, in=0, missing perf/target, keepAccelerator, negative SLO."""
    img = wva.synth.make_system(8, 2, seed=41)
    img.perf_valid[:] = 1; img.srv_target_valid[:] = 1
    img.srv_arrival_rpm[0] = 0.0
    img.srv_out_tokens[1] = 1; img.srv_in_tokens[1] = 0
    img.srv_in_tokens[2] = 0
    img.perf_valid[3 * 2 + 1] = 0
    img.srv_target_valid[4] = 0
    img.srv_keep_acc[5] = 1; img.srv_cur_acc[5] = 1
    img.srv_slo_itl[6] = -1.0
    img.srv_out_tokens[7] = 0
    w_best, w_cube, w_status, _ = oracle.analyze_grid(img, 6, 40)
    for mode in (1, 0, 3, 5):
        ctx.set_certified_tails(mode)
        ctx.upload(img)
        best, cube, status = ctx.analyze_grid(6, 40, want_cube=True)
        assert np.array_equal(status, w_status) and cube.tobytes() == w_cube.tobytes() and best.tobytes() == w_best.tobytes(), mode
    ctx.set_certified_tails(1)


def _capacity_case(wva, oracle, seed, S, A, T, frac):
    img = wva.synth.make_system(S, A, seed=seed, n_types=T, max_pair_batch=256)
    pairs, feas, _ = oracle.analyze_pairs(img, threads=oracle.hardware_threads())
    acc, chosen = oracle.solve(img, pairs, feas, unlimited=True)
    wva.synth.set_capacity_from_demand(img, chosen.acc, chosen.num_replicas, fraction=frac)
    return img, pairs, feas


def test_solve_unlimited_and_totals(wva, oracle, ctx):
    img, pairs, feas = _capacity_case(wva, oracle, 51, 400, 6, 3, 0.6)
    ctx.upload(img)
    ctx.analyze_pairs(download=False)
    acc, chosen = ctx.solve(unlimited=True)
    w_acc, w_chosen = oracle.solve(img, pairs, feas, unlimited=True)
    assert np.array_equal(acc, w_acc)
    _assert_allocs_equal(chosen, w_chosen)
    count, cost = ctx.allocate_by_type()
    w_count, w_cost = oracle.allocate_by_type(img, w_acc, w_chosen)
    assert np.array_equal(count, w_count) and cost.tobytes() == w_cost.tobytes()
    assert ctx.solution_time_usec() >= 0


@pytest.mark.parametrize("policy", [0, 1, 2, 3])
@pytest.mark.parametrize("delayed", [False, True])
@pytest.mark.parametrize("ranked", [2, 1, 0])
def test_solve_greedy_policies(wva, oracle, ctx, policy, delayed, ranked):
    img, pairs, feas = _capacity_case(wva, oracle, 52, 500, 6, 3, 0.6)
    ctx.upload(img)
    ctx.analyze_pairs(download=False)
    ctx.solve_set_ranked(ranked)
    acc, chosen = ctx.solve(unlimited=False, delayed_best_effort=delayed, policy=policy)
    assert ctx.solve_greedy_path() == ranked + 1
    ctx.solve_set_ranked(2)
    w_acc, w_chosen = oracle.solve(img, pairs, feas, unlimited=False, delayed_best_effort=delayed, policy=policy)
    assert np.array_equal(acc, w_acc)
    _assert_allocs_equal(chosen, w_chosen)
    count, cost = ctx.allocate_by_type()
    w_count, w_cost = oracle.allocate_by_type(img, w_acc, w_chosen)
    assert np.array_equal(count, w_count) and cost.tobytes() == w_cost.tobytes()
    # capacity is respected unless round-robin over-allocation applies (never exceeds capacity either)
    assert (count <= img.type_capacity).all()
    assert (w_acc < 0).any() or policy != 0


def test_solve_greedy_ties(wva, oracle, ctx):
    """identical servers: every comparison in the greedy order ties; canonical order = server index."""
    img = wva.synth.make_system(40, 3, seed=53, n_types=2, one_model_per_server=False, n_models=1)
    for name, _ in wva.abi.SRV_FIELDS:
        arr = getattr(img, name); arr[:] = arr[0]
    img.srv_model[:] = 0; img.perf_valid[:] = 1; img.srv_target_valid[:] = 1
    img.srv_arrival_rpm[:] = 900.0; img.srv_keep_acc[:] = 0; img.srv_cur_acc[:] = -1; img.srv_cur_cost[:] = 0
    img.srv_cur_replicas[:] = 0; img.srv_max_batch[:] = 32; img.srv_slo_tps[:] = 0
    pairs, feas, _ = oracle.analyze_pairs(img)
    assert feas.reshape(img.S, img.A).any(axis=1).all() and feas.sum() >= 2 * img.S
    acc_u, ch_u = oracle.solve(img, pairs, feas, unlimited=True)
    wva.synth.set_capacity_from_demand(img, ch_u.acc, ch_u.num_replicas, fraction=0.45)
    for policy, ranked in ((0, 2), (1, 2), (2, 2), (3, 2), (0, 1), (1, 1), (2, 1), (3, 1), (1, 0), (3, 0)):
        ctx.upload(img)
        ctx.analyze_pairs(download=False)
        ctx.solve_set_ranked(ranked)
        acc, chosen = ctx.solve(unlimited=False, policy=policy)
        path = ctx.solve_greedy_path()
        ctx.solve_set_ranked(2)
        assert path == ranked + 1                    # scan / ranked: every key ties, the shared-group stacks carry the recency order
        w_acc, w_chosen = oracle.solve(img, pairs, feas, unlimited=False, policy=policy)
        assert np.array_equal(acc, w_acc)
        _assert_allocs_equal(chosen, w_chosen)


@pytest.mark.parametrize("A,protos", [(6, 7), (12, 5), (20, 3)])
def test_solve_greedy_partial_ties(wva, oracle, ctx, A, protos):
    """classes of identical servers (prototype i % protos) with different priorities: long shared groups whose stacks the
    static-order scan pops several runs at a time (8, 16 or 32 lanes per run for 6, 12 and 20 accelerators), mixed with
    unique keys; every policy, both best-effort schedules, all three solver paths against the oracle"""
    S = 420
    img = wva.synth.make_system(S, A, seed=91 + A, n_types=3, max_pair_batch=64)
    for name, _ in wva.abi.SRV_FIELDS:
        arr = getattr(img, name)
        for i in range(protos, S):
            if i % 11 != 0:                       # every 11th server keeps its own (unique) keys
                arr[i] = arr[i % protos]
    pairs, feas, _ = oracle.analyze_pairs(img, threads=oracle.hardware_threads())
    acc_u, ch_u = oracle.solve(img, pairs, feas, unlimited=True)
    wva.synth.set_capacity_from_demand(img, ch_u.acc, ch_u.num_replicas, fraction=0.5)
    for policy in (0, 1, 2, 3):
        for delayed in (False, True):
            w_acc, w_chosen = oracle.solve(img, pairs, feas, unlimited=False, delayed_best_effort=delayed, policy=policy)
            for ranked in (2, 1):
                ctx.upload(img)
                ctx.analyze_pairs(download=False)
                ctx.solve_set_ranked(ranked)
                acc, chosen = ctx.solve(unlimited=False, delayed_best_effort=delayed, policy=policy)
                path = ctx.solve_greedy_path()
                st = ctx.solve_stats()
                ctx.solve_set_ranked(2)
                assert path == ranked + 1
                assert np.array_equal(acc, w_acc), (policy, delayed, ranked)
                _assert_allocs_equal(chosen, w_chosen)
                if ranked == 2:
                    assert (st[1] >> 32) > 0       # runs did go through the shared-group stacks


@pytest.mark.parametrize("policy,delayed,ranked", [(1, False, 2), (3, True, 2), (2, False, 2), (0, True, 2), (3, False, 2), (1, False, 1), (3, True, 1), (2, False, 1), (0, True, 1), (3, True, 0), (1, False, 0)])
def test_solve_greedy_large(wva, oracle, ctx, policy, delayed, ranked):
    """12 000 servers: the delayed-best-effort queue no longer fits shared memory (global-memory
    heap), the per-priority groups do.  Candidates come from the CUDA path (pair parity is
    covered above); the assignment is compared with the oracle's."""
    img = wva.synth.make_system(12000, 4, seed=57, n_types=3)
    ctx.upload(img)
    pairs, feas = ctx.analyze_pairs()
    acc_u, ch_u = oracle.solve(img, pairs, feas, unlimited=True)
    wva.synth.set_capacity_from_demand(img, ch_u.acc, ch_u.num_replicas, fraction=0.55)
    ctx.upload(img)
    ctx.analyze_pairs(download=False)
    ctx.solve_set_ranked(ranked)
    acc, chosen = ctx.solve(unlimited=False, delayed_best_effort=delayed, policy=policy)
    path = ctx.solve_greedy_path()
    ctx.solve_set_ranked(2)
    w_acc, w_chosen = oracle.solve(img, pairs, feas, unlimited=False, delayed_best_effort=delayed, policy=policy)
    assert np.array_equal(acc, w_acc)
    _assert_allocs_equal(chosen, w_chosen)
    assert path == ranked + 1
    count, cost = ctx.allocate_by_type()
    assert (count <= img.type_capacity).all()
    assert (acc < 0).any() or policy != 0


def test_overlapped_analyze_equals_separate_calls(wva, oracle, ctx):
    """wva_analyze runs the pair sizing and the candidate sweep concurrently on two streams"""
    img = wva.synth.make_system(40, 3, seed=71)
    ctx.upload(img)
    ctx.analyze(8, 40, want_cube=False)
    pairs, fe = ctx.pairs_fetch()
    best = ctx.grid_fetch()
    want, wfe, _ = oracle.analyze_pairs(img, threads=oracle.hardware_threads())
    w_best, _, _, _ = oracle.analyze_grid(img, 8, 40, want_cube=False, threads=oracle.hardware_threads())
    assert np.array_equal(fe, wfe)
    _assert_allocs_equal(pairs, want)
    assert best.tobytes() == w_best.tobytes()
    acc, chosen = ctx.solve(unlimited=True)
    w_acc, w_chosen = oracle.solve(img, want, wfe, unlimited=True)
    assert np.array_equal(acc, w_acc)


def test_state_errors(wva, ctx):
    from inferno_autoscaler_b200 import binding
    img = wva.synth.make_system(4, 2, seed=1)
    ctx.upload(img)
    with pytest.raises(binding.WvaError) as e:
        ctx.solve(unlimited=True)
    assert e.value.code == wva.abi.ESTATE
    img.srv_priority[0] = 0
    with pytest.raises(binding.WvaError) as e:
        ctx.upload(img)
    assert e.value.code == wva.abi.EINVAL


def test_sharded_pairs_equal_full(wva, oracle, ctx):
    """server shards are independent: shard-wise analysis concatenates to the full result."""
    img = wva.synth.make_system(90, 3, seed=61, max_pair_batch=256)
    ctx.upload(img)
    full, ffe = ctx.analyze_pairs()
    A = img.A
    for first, count in [(0, 30), (30, 45), (75, 15)]:
        ctx.set_shard(first, count)
        part, pfe = ctx.analyze_pairs()
        sl = slice(first * A, (first + count) * A)
        mask = np.zeros(img.S * A, bool); mask[sl] = True
        assert np.array_equal(pfe[sl], ffe[sl])
        _assert_allocs_equal(part, full, mask)


def test_model_solve_sequence_matches_oracle(wva, oracle, ctx):
    """the bare MM1ModelStateDependent of pkg/analyzer (caller-supplied service rates): a sequence of Solve calls on one
    model instance, including the stale-p[0] validity rule (queuemodel.go:30) and invalid calls in the middle"""
    rng = np.random.default_rng(33)
    for K, n_rates in ((40, 7), (1, 1), (300, 300), (120, 200)):
        serv = np.sort(rng.uniform(0.05, 2.0, n_rates)).astype(F)
        lam = np.concatenate([rng.uniform(0.01, 1.5 * serv[-1], 6), [-1.0, 0.3, 0.0, 0.7]]).astype(F)
        mu = np.concatenate([np.ones(7), [0.0, 1.0, 1.0]]).astype(F)
        out, p = ctx.model_solve(K, serv, lam, mu)
        m = oracle.Model(K, serv)
        for i in range(len(lam)):
            want = m.solve(float(lam[i]), float(mu[i]))
            got = dict(zip(oracle._SOLVE_KEYS, out[i]))
            for key in oracle._SOLVE_KEYS:
                assert np.float32(got[key]).tobytes() == np.float32(want[key]).tobytes() or (np.isnan(got[key]) and np.isnan(want[key])), (K, i, key)
        assert p.tobytes() == m.probabilities().tobytes()


@pytest.mark.parametrize("policy", [0, 1, 3])
def test_solve_greedy_nonfinite_values(wva, oracle, ctx, policy):
    """costs that overflow float32 make candidate values Inf / NaN (Inf - Inf in TransitionPenalty): the greedy order is
    cmp.Compare's (NaN lowest; orderFunc's delta == delta test fails for NaN), both solver paths, against the oracle"""
    img = wva.synth.make_system(240, 4, seed=77, n_types=2, max_pair_batch=64)
    img.acc_cost[1] = np.float32(3.0e38)                       # cost = acc.Cost * float32(instances * replicas) -> +Inf beyond 1 unit
    has_cur = img.srv_cur_acc >= 0
    img.srv_cur_cost[has_cur & (np.arange(img.S) % 3 == 0)] = np.float32(np.inf)
    ctx.upload(img)
    pairs, feas = ctx.analyze_pairs()
    o_pairs, o_feas, _ = oracle.analyze_pairs(img, threads=4)
    assert np.array_equal(feas, o_feas) and pairs.equal_bits(o_pairs)[0]
    vals = pairs.value[feas.astype(bool)]
    assert np.isnan(vals).any() and np.isinf(vals).any()
    acc_u, ch_u = oracle.solve(img, o_pairs, o_feas, unlimited=True)
    wva.synth.set_capacity_from_demand(img, ch_u.acc, ch_u.num_replicas, fraction=0.6)
    for ranked in (2, 1, 0):
        ctx.solve_set_ranked(ranked)
        ctx.upload(img)
        ctx.analyze_pairs(download=False)
        acc, chosen = ctx.solve(unlimited=False, policy=policy)
        w_acc, w_chosen = oracle.solve(img, o_pairs, o_feas, unlimited=False, policy=policy)
        assert np.array_equal(acc, w_acc), ranked
        ok, field = chosen.equal_bits(w_chosen)
        assert ok, (ranked, field)
    ctx.solve_set_ranked(2)
