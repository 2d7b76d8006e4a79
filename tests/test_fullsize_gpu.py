"""GPU: BASELINE.json's full sizes, checked through properties that do not need the oracle to redo the
whole job: bit-exact comparison on random samples, sum-of-status accounting, idempotence, shard
invariance, and the defining property of the argmin."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sample_cube_against_oracle(wva, oracle, img, cube, status, R, B, n, seed):
    rng = np.random.default_rng(seed)
    A = img.A
    idx = rng.integers(0, len(cube), n)
    b = idx % B + 1
    r = (idx // B) % R + 1
    pair = idx // (B * R)
    s, a = pair // A, pair % A
    ok_pair = (status[idx] & 0xfe) != wva.abi.CAND_ERR_PAIR
    ok_pair &= (status[idx] & 0xfe) != wva.abi.CAND_ERR_CONFIG
    idx, b, r, s, a = idx[ok_pair], b[ok_pair], r[ok_pair], s[ok_pair], a[ok_pair]
    pi = img.srv_model[s] * A + a
    cfg = np.zeros(len(idx), dtype=wva.abi.QUEUE_CONFIG_DTYPE)
    cfg["max_batch_size"] = b; cfg["max_queue_size"] = 10 * b
    cfg["alpha"] = img.perf_alpha[pi]; cfg["beta"] = img.perf_beta[pi]
    cfg["gamma"] = img.perf_gamma[pi]; cfg["delta"] = img.perf_delta[pi]
    cfg["avg_input_tokens"] = img.srv_in_tokens[s]; cfg["avg_output_tokens"] = img.srv_out_tokens[s]
    tps = img.srv_slo_tps[s]
    total = np.where(tps == 0, img.srv_arrival_rpm[s] / np.float32(60), tps / img.srv_out_tokens[s].astype(np.float32)).astype(np.float32)
    rate = (total / r.astype(np.float32)).astype(np.float32)
    m, st = oracle.queue_analyze(cfg, rate)
    assert np.array_equal(st, status[idx] & 0xfe)
    assert m.tobytes() == cube[idx].tobytes()
    return len(idx)


def test_config2_full_sweep_properties(wva, oracle, ctx):
    """32 servers x 4 accelerators x replicas 1-64 x batch 1-512 = 4 194 304 candidates."""
    img, c = wva.synth.baseline_config(2)
    R, B = c["r_max"], c["b_max"]
    ctx.upload(img)
    best, cube, status = ctx.analyze_grid(R, B, want_cube=True)
    cnt = ctx.grid_counters()
    # accounting: every candidate has exactly one status; the OK count matches the kernel's counter
    assert cnt["candidates_ok"] == int(((status & 0xfe) == 0).sum())
    assert cnt["steps_executed"] < cnt["steps_algorithmic"]
    # random sample, bit-exact against the reference API (oracle)
    assert _sample_cube_against_oracle(wva, oracle, img, cube, status, R, B, 4000, seed=5) > 1000
    # the highest-load candidates (deferred long chains) specifically
    heavy = np.flatnonzero(((status & 0xfe) == 0) & (cube["rho"] > 0.9))
    assert len(heavy) > 100
    # argmin property: the winner is feasible and no feasible candidate of the server has a smaller value
    A = img.A
    feas = (status & 1).reshape(img.S, A, R, B)
    for s in range(img.S):
        if best["acc"][s] < 0:
            assert feas[s].sum() == 0
            continue
        a, r, b = int(best["acc"][s]), int(best["replicas"][s]), int(best["batch"][s])
        assert feas[s, a, r - 1, b - 1] == 1
        rec = cube.reshape(img.S, A, R, B)[s, a, r - 1, b - 1]
        assert rec["avg_token_time"] == best["itl"][s] and rec["rho"] == best["rho"][s]
    # idempotence + independence from the deferral threshold (a pure scheduling knob)
    ctx.grid_set_tail_cap(0)
    best0, cube0, status0 = ctx.analyze_grid(R, B, want_cube=True)
    ctx.grid_set_tail_cap(64)
    best1, _, status1 = ctx.analyze_grid(R, B, want_cube=True)
    ctx.grid_set_tail_cap(192)
    best2, _, status2 = ctx.analyze_grid(R, B, want_cube=True)
    ctx.grid_set_tail_cap(-1)
    assert best2.tobytes() == best.tobytes() and np.array_equal(status2, status)
    assert best0.tobytes() == best.tobytes() and cube0.tobytes() == cube.tobytes() and np.array_equal(status0, status)
    assert best1.tobytes() == best.tobytes() and np.array_equal(status1, status)
    # certified closed-form tails are a pure speed-up: switching them off must not change one bit
    ctx.set_certified_tails(False)
    ctx.upload(img)
    best_x, cube_x, status_x = ctx.analyze_grid(R, B, want_cube=True)
    cnt_x = ctx.grid_counters()
    ctx.set_certified_tails(3)          # certified tails, one thread per candidate
    ctx.upload(img)
    best_y, cube_y, status_y = ctx.analyze_grid(R, B, want_cube=True)
    ctx.set_certified_tails(5)          # certified tails, one thread per row (shared ramp)
    ctx.upload(img)
    best_z, cube_z, status_z = ctx.analyze_grid(R, B, want_cube=True)
    ctx.set_certified_tails(9)          # certified tails, one warp per row (lanes = batch sizes)
    ctx.upload(img)
    best_w, cube_w, status_w = ctx.analyze_grid(R, B, want_cube=True)
    assert best_w.tobytes() == best.tobytes() and cube_w.tobytes() == cube.tobytes() and np.array_equal(status_w, status)
    ctx.set_certified_tails(True)
    ctx.upload(img)
    assert best_x.tobytes() == best.tobytes() and cube_x.tobytes() == cube.tobytes() and np.array_equal(status_x, status)
    assert best_y.tobytes() == best.tobytes() and cube_y.tobytes() == cube.tobytes() and np.array_equal(status_y, status)
    assert best_z.tobytes() == best.tobytes() and cube_z.tobytes() == cube.tobytes() and np.array_equal(status_z, status)
    assert cnt["steps_executed"] < cnt_x["steps_executed"]
    # shard invariance: two shards concatenate to the full result
    half = img.S // 2
    ctx.set_shard(0, half); b0, _, _ = ctx.analyze_grid(R, B)
    ctx.set_shard(half, img.S - half); b1, _, _ = ctx.analyze_grid(R, B)
    assert np.concatenate([b0, b1]).tobytes() == best.tobytes()


def test_config3_pairs_sample(wva, oracle, ctx):
    """1 000 servers x 8 accelerators: all 8 000 pairs on the GPU (both kernels), a random sample of
    servers re-done by the oracle."""
    img, c = wva.synth.baseline_config(3)
    ctx.upload(img)
    got, gfe = ctx.analyze_pairs()
    ctx.pairs_set_warp_max(0)
    ctx.upload(img)
    got_t, gfe_t = ctx.analyze_pairs()
    ctx.pairs_set_warp_max(1 << 22)
    assert np.array_equal(gfe, gfe_t) and got.equal_bits(got_t)[0]
    ctx.set_certified_tails(False)
    ctx.upload(img)
    got_x, gfe_x = ctx.analyze_pairs()
    ctx.set_certified_tails(True)
    ctx.upload(img)
    assert np.array_equal(gfe, gfe_x) and got.equal_bits(got_x)[0]
    ctx.analyze_pairs(download=False)
    rng = np.random.default_rng(3)
    # server 628 holds a pair whose chain spans more than 1000 binades (p[n]/sum subnormal for part of
    # pass 2: the IEEE-division elements of pass2_run)
    pick = np.unique(np.concatenate([rng.choice(img.S, 40, replace=False), [628]]))
    for s in pick:
        sub = img.shard(int(s), 1)
        want, wfe, _ = oracle.analyze_pairs(sub)
        sl = slice(int(s) * img.A, (int(s) + 1) * img.A)
        assert np.array_equal(gfe[sl], wfe)
        for name, dt in wva.abi.ALLOC_FIELDS:
            a, b = getattr(got, name)[sl], getattr(want, name)
            assert a.tobytes() == b.tobytes(), (s, name)
    # solve + totals at this size against the oracle fed with the GPU's candidates (decision step only)
    acc, chosen = ctx.solve(unlimited=True)
    w_acc, w_chosen = oracle.solve(img, got, gfe, unlimited=True)
    assert np.array_equal(acc, w_acc) and chosen.equal_bits(w_chosen)[0]
    cnt, cst = ctx.allocate_by_type()
    w_cnt, w_cst = oracle.allocate_by_type(img, w_acc, w_chosen)
    assert np.array_equal(cnt, w_cnt) and cst.tobytes() == w_cst.tobytes()


def test_config4_greedy_with_caps(wva, oracle, ctx):
    """10 000 servers with per-type capacity caps (BASELINE config 4): the sequential greedy on the device
    against the oracle's, both fed with the GPU's candidate records."""
    img = wva.synth.make_system(10000, 8, seed=4, n_types=4, max_pair_batch=64)
    ctx.upload(img)
    pairs, feas = ctx.analyze_pairs()
    acc_u, ch_u = oracle.solve(img, pairs, feas, unlimited=True)
    wva.synth.set_capacity_from_demand(img, ch_u.acc, ch_u.num_replicas, fraction=0.6)
    for policy in (wva.abi.POLICY_NONE, wva.abi.POLICY_PRIORITY_ROUND_ROBIN):
        ctx.upload(img)
        ctx.analyze_pairs(download=False)
        acc, chosen = ctx.solve(unlimited=False, policy=policy)
        w_acc, w_chosen = oracle.solve(img, pairs, feas, unlimited=False, policy=policy)
        assert np.array_equal(acc, w_acc) and chosen.equal_bits(w_chosen)[0]
        cnt, _ = ctx.allocate_by_type()
        assert (cnt <= img.type_capacity).all()
