"""GPU: incremental updates of the resident image (System.AddServerFromSpec / RemoveServer / SetCountFromSpec /
Model.AddPerfDataFromSpec, reference pkg/core/system.go:99-171, model.go:45-54) -- only the touched rows cross
PCIe, and every result afterwards equals a from-scratch upload of the edited image and the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _one_server(wva, img, s):
    return img.take([s])


def _check_against_oracle(wva, oracle, ctx, img, servers):
    """pairs of the listed servers on the resident image vs the oracle on the host image"""
    got, gfe = ctx.analyze_pairs()
    for s in servers:
        sub = img.take([s])
        want, wfe, _ = oracle.analyze_pairs(sub)
        sl = slice(s * img.A, (s + 1) * img.A)
        assert np.array_equal(gfe[sl], wfe), s
        for name, _ in wva.abi.ALLOC_FIELDS:
            assert getattr(got, name)[sl].tobytes() == getattr(want, name).tobytes(), (s, name)
    return got, gfe


def test_one_server_of_100k_changes_without_full_upload(wva, oracle, ctx):
    """BASELINE config 5's image (100 000 servers x 16 accelerators, ~51 MB): one server's load changes"""
    c = wva.synth.CONFIGS[5]
    img = wva.synth.make_system(c["S"], c["A"], seed=5, n_types=c["T"])
    ctx.upload(img)
    full = ctx.upload_bytes()
    assert full >= img.nbytes() and full < 1.1 * img.nbytes() + (1 << 20)
    s = 54_321
    row = img.take([s])
    row.srv_arrival_rpm[0] = np.float32(1234.5)
    row.srv_in_tokens[0] = 777
    row.srv_out_tokens[0] = 210
    row.srv_max_batch[0] = 0
    ctx.update_servers(s, row)
    assert ctx.upload_bytes() == 54                      # 13 x 4 B + 2 x 1 B: one row of the 15 server arrays
    assert ctx.dims() == (img.S, img.A, img.M, img.T)
    assert img.srv_in_tokens[s] == 777                    # the binding keeps the host image in step
    ctx.set_shard(s - 2, 5)
    _check_against_oracle(wva, oracle, ctx, img, range(s - 2, s + 3))
    # the sweep sees the new row too
    best, _, _ = ctx.analyze_grid(8, 64)
    o_best, _, _, _ = oracle.analyze_grid(img, 8, 64, s0=s - 2, s1=s + 3, want_cube=False, threads=4)
    assert best.tobytes() == o_best.tobytes()


def test_add_remove_servers_models_and_capacity(wva, oracle, ctx):
    img = wva.synth.make_system(300, 6, seed=71, n_types=3)
    ctx.upload(img)
    donor = wva.synth.make_system(8, 6, seed=72, n_types=3)
    # AddModel + AddPerfDataFromSpec for 8 new models, then AddServerFromSpec for 8 servers that use them
    models = donor.take(range(8))                      # carries donor's perf table (8 models x 6 accelerators)
    ctx.update_models(img.M, models)
    assert ctx.upload_bytes() == 8 * 6 * 29
    new = donor.take(range(8))
    new.srv_model[:] = np.arange(300, 308, dtype=np.int32)
    ctx.update_servers(img.S, new)
    assert ctx.dims() == (308, 6, 308, 3) and img.S == 308 and img.M == 308
    _check_against_oracle(wva, oracle, ctx, img, list(range(296, 308)) + [0, 150])
    # RemoveServer: the last row moves into the freed slot
    ctx.remove_server(5)
    assert ctx.dims()[0] == 307 and img.S == 307
    got, gfe = _check_against_oracle(wva, oracle, ctx, img, [4, 5, 6, 306])
    # everything equals a from-scratch upload of the edited image
    ctx.upload(img)
    ref, rfe = ctx.analyze_pairs()
    assert np.array_equal(gfe, rfe) and got.equal_bits(ref)[0]
    # SetCountFromSpec: capacities change, the limited solve follows (candidates stay valid)
    acc_u, ch_u = ctx.solve(unlimited=True)
    demand = wva.synth.set_capacity_from_demand(img, ch_u.acc, ch_u.num_replicas, fraction=0.5)
    ctx.set_capacity(img.type_capacity)
    assert ctx.upload_bytes() == 8 * img.T
    acc, chosen = ctx.solve(unlimited=False, policy=wva.abi.POLICY_PRIORITY_EXHAUSTIVE)
    w_acc, w_chosen = oracle.solve(img, ref, rfe, unlimited=False, policy=wva.abi.POLICY_PRIORITY_EXHAUSTIVE)
    assert np.array_equal(acc, w_acc) and chosen.equal_bits(w_chosen)[0]
    cnt, _ = ctx.allocate_by_type()
    assert (cnt <= img.type_capacity).all() and (demand > img.type_capacity).any()


def test_updates_fail_loudly_when_out_of_rows(wva, ctx):
    from inferno_autoscaler_b200 import binding
    img = wva.synth.make_system(10, 2, seed=73, n_types=2)
    ctx.upload(img)
    big = wva.synth.make_system(200, 2, seed=74, n_types=2)
    big.srv_model[:] = 0
    with pytest.raises(binding.WvaError) as e:
        ctx.update_servers(10, big)                     # 10 + 200 > 10 + 64 spare rows
    assert e.value.code == wva.abi.ECAPACITY
    with pytest.raises(binding.WvaError):
        ctx.update_servers(12, img.take([0]))            # not contiguous with the image
