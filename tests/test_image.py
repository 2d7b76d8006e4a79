"""CPU-only: host-side SetFromSpec equivalent (string interning, class/target resolution)."""
import numpy as np


def _spec():
    return {
        "acceleratorData": {"accelerators": [
            {"name": "A100", "type": "A100", "multiplicity": 1, "cost": 40.0},
            {"name": "G2", "type": "Gaudi", "multiplicity": 2, "cost": 23.0},
            {"name": "A100x2", "type": "A100", "multiplicity": 2, "cost": 80.0}]},
        "modelData": {"models": [
            {"name": "m1", "acc": "A100", "accCount": 0, "maxBatchSize": 8, "atTokens": 128,
             "decodeParms": {"alpha": 20.0, "beta": 0.5}, "prefillParms": {"gamma": 5.0, "delta": 0.01}},
            {"name": "m1", "acc": "G2", "accCount": 2, "maxBatchSize": 16, "atTokens": 128,
             "decodeParms": {"alpha": 25.0, "beta": 0.7}, "prefillParms": {"gamma": 6.0, "delta": 0.02}},
            {"name": "m1", "acc": "nonexistent", "accCount": 1, "maxBatchSize": 16, "atTokens": 128}]},
        "serviceClassData": {"serviceClasses": [
            {"name": "Premium", "priority": 1, "modelTargets": [{"model": "m1", "slo-itl": 24, "slo-ttft": 500}]},
            {"name": "Weird", "priority": 1000, "modelTargets": []}]},
        "serverData": {"servers": [
            {"name": "s1", "class": "Premium", "model": "m1", "minNumReplicas": 1,
             "currentAlloc": {"accelerator": "G2", "numReplicas": 3, "cost": 138.0,
                              "load": {"arrivalRate": 120.0, "avgInTokens": 64, "avgOutTokens": 128}}},
            {"name": "s2", "class": "", "model": "m1", "currentAlloc": {"accelerator": "H100", "load": {}}},
            {"name": "s3", "class": "Weird", "model": "m2", "currentAlloc": {}}]},
        "capacityData": {"count": [{"type": "A100", "count": 16}]},
    }


def test_from_spec_interning(wva):
    img = wva.SystemImage.from_spec(_spec())
    assert (img.S, img.A, img.M, img.T) == (3, 3, 1, 2)
    assert img.acc_names == ["A100", "G2", "A100x2"] and img.type_names == ["A100", "Gaudi"]
    assert list(img.acc_type) == [0, 1, 0] and list(img.type_capacity) == [16, 0]
    assert list(img.perf_valid) == [1, 1, 0]
    assert img.perf_acc_count[0] == 0 and img.perf_acc_count[1] == 2
    # s1: resolved class + target
    assert img.srv_priority[0] == 1 and img.srv_target_valid[0] == 1 and img.srv_slo_itl[0] == np.float32(24)
    assert img.srv_cur_acc[0] == 1 and img.srv_cur_replicas[0] == 3
    # s2: class "" -> "Free" (missing) -> default priority, no target; current accelerator unknown
    assert img.srv_priority[1] == wva.abi.DEFAULT_PRIORITY and img.srv_target_valid[1] == 0
    assert img.srv_cur_acc[1] == wva.abi.ACC_UNKNOWN
    # s3: out-of-range class priority clamps to 100; model unknown
    assert img.srv_priority[2] == 100 and img.srv_model[2] == -1 and img.srv_cur_acc[2] == wva.abi.ACC_NONE


def test_synth_is_deterministic(wva):
    a = wva.synth.make_system(50, 4, seed=3)
    b = wva.synth.make_system(50, 4, seed=3)
    for name, _ in a.ALL_FIELDS:
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    c = wva.synth.make_system(50, 4, seed=4)
    assert not np.array_equal(a.srv_arrival_rpm, c.srv_arrival_rpm)


def test_shard_slices_servers(wva):
    a = wva.synth.make_system(10, 2, seed=1)
    sh = a.shard(3, 4)
    assert sh.S == 4 and np.array_equal(sh.srv_arrival_rpm, a.srv_arrival_rpm[3:7])
    assert np.array_equal(sh.perf_alpha, a.perf_alpha)
