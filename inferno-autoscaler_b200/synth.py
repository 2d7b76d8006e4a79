"""Deterministic synthetic system images for the BASELINE.json configurations.

Distributions follow SURVEY.md 8(d): accelerator costs / multiplicities, per (model,
accelerator) alpha/beta/gamma/delta, three service classes with the SLOs of the reference's
deploy/configmap-serviceclass.yaml, log-uniform arrival rates with a zero-load
sub-population, token counts including the in=0 / out=1 edge cases, random current
allocations (exercise TransitionPenalty) and keepAccelerator servers.  All floats are
drawn in float64 and rounded once to float32, as the reference's JSON decode would.

The generator is shared by the CUDA path, the oracle and bench.py so that every arm sees
bit-identical inputs.  PRNG: numpy PCG64 seeded with the config seed.
"""
import numpy as np

from . import abi
from .image import SystemImage

# (slo_itl, slo_ttft, priority): Premium / Standard / Freemium
SERVICE_CLASSES = [(24.0, 500.0, 1), (80.0, 1000.0, 5), (200.0, 2000.0, 10)]


def make_system(n_servers, n_accels, seed, n_types=None, one_model_per_server=True, n_models=None,
                edge_fraction=0.02, tps_fraction=0.10, zero_load_fraction=0.02, keep_fraction=0.10,
                max_pair_batch=512, out_tokens_min=32):
    """Random system image.  `max_pair_batch` bounds N = maxBatch*atTokens/outTokens of the
    reference sizing path (pkg/core/allocation.go:85) through a server-level batch override, so
    one pathological pair cannot dominate a whole run (N has no upper bound in the reference).
    `max_pair_batch=0, out_tokens_min=1` is SURVEY 8(d)'s generator as written ("un-tamed": out_tokens
    from {1..1024}, no bound on N -- N reaches maxBatch*atTokens = 262 144, K = 11 N = 2.9 M states)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    S, A = int(n_servers), int(n_accels)
    T = int(n_types) if n_types else A
    M = S if one_model_per_server else int(n_models or S)
    img = SystemImage(S, A, M, T)

    # accelerators
    img.acc_cost[:] = np.round(rng.uniform(20.0, 120.0, A), 2).astype(np.float32)
    img.acc_multiplicity[:] = rng.choice(np.array([1, 1, 1, 2, 4, 8], dtype=np.int32), A)
    img.acc_type[:] = np.arange(A, dtype=np.int32) % T
    img.type_capacity[:] = 0

    # model x accelerator perf table
    n = M * A
    img.perf_alpha[:] = rng.uniform(5.0, 30.0, n).astype(np.float32)
    img.perf_beta[:] = rng.uniform(0.02, 0.8, n).astype(np.float32)
    img.perf_gamma[:] = rng.uniform(2.0, 250.0, n).astype(np.float32)
    img.perf_delta[:] = np.exp(rng.uniform(np.log(1e-4), np.log(0.1), n)).astype(np.float32)
    img.perf_acc_count[:] = rng.choice(np.array([1, 2, 4], dtype=np.int32), n)
    img.perf_max_batch[:] = rng.integers(4, 513, n, dtype=np.int32)
    img.perf_at_tokens[:] = rng.choice(np.array([128, 256, 512], dtype=np.int32), n)
    img.perf_valid[:] = (rng.uniform(0, 1, n) >= 0.01).astype(np.uint8)   # 1 % missing perf rows

    # servers
    img.srv_model[:] = np.arange(S, dtype=np.int32) if one_model_per_server else rng.integers(0, M, S, dtype=np.int32)
    cls = rng.integers(0, len(SERVICE_CLASSES), S)
    sc = np.array(SERVICE_CLASSES, dtype=np.float64)
    img.srv_slo_itl[:] = sc[cls, 0].astype(np.float32)
    img.srv_slo_ttft[:] = sc[cls, 1].astype(np.float32)
    img.srv_priority[:] = sc[cls, 2].astype(np.int32)
    tps = rng.uniform(0, 1, S) < tps_fraction
    img.srv_slo_tps[:] = np.where(tps, np.round(rng.uniform(50.0, 5000.0, S), 1), 0.0).astype(np.float32)
    img.srv_target_valid[:] = (rng.uniform(0, 1, S) >= 0.005).astype(np.uint8)
    arrival = np.exp(rng.uniform(np.log(1.0), np.log(60000.0), S))
    arrival[rng.uniform(0, 1, S) < zero_load_fraction] = 0.0
    img.srv_arrival_rpm[:] = arrival.astype(np.float32)
    img.srv_in_tokens[:] = rng.integers(16, 4097, S, dtype=np.int32)
    img.srv_out_tokens[:] = rng.integers(int(out_tokens_min), 1025, S, dtype=np.int32)
    edge = rng.uniform(0, 1, S)
    img.srv_out_tokens[edge < edge_fraction / 2] = 1                      # single output token
    img.srv_in_tokens[(edge >= edge_fraction / 2) & (edge < edge_fraction)] = 0   # decode only
    img.srv_min_replicas[:] = rng.integers(0, 2, S, dtype=np.int32)
    img.srv_max_batch[:] = 0
    over = rng.uniform(0, 1, S) < 0.15
    img.srv_max_batch[over] = rng.integers(1, 257, int(over.sum()), dtype=np.int32)
    img.srv_keep_acc[:] = (rng.uniform(0, 1, S) < keep_fraction).astype(np.uint8)
    has_cur = rng.uniform(0, 1, S) < 0.5
    cur_acc = rng.integers(0, A, S, dtype=np.int32)
    img.srv_cur_acc[:] = np.where(has_cur, cur_acc, abi.ACC_NONE)
    img.srv_cur_replicas[:] = np.where(has_cur, rng.integers(1, 17, S), 0).astype(np.int32)
    img.srv_cur_cost[:] = np.where(has_cur, np.round(rng.uniform(20.0, 2000.0, S), 2), 0.0).astype(np.float32)

    # bound the sizing path's N with a server override where the perf-derived value is too large
    if max_pair_batch:
        pa = img.perf_max_batch.reshape(M, A).astype(np.int64) * img.perf_at_tokens.reshape(M, A).astype(np.int64)
        worst = pa[img.srv_model].max(axis=1) // np.maximum(img.srv_out_tokens.astype(np.int64), 1)
        big = (worst > max_pair_batch) & (img.srv_max_batch == 0)
        img.srv_max_batch[big] = max_pair_batch
    return img


def set_capacity_from_demand(img, chosen_acc_idx, chosen_replicas, fraction=0.6):
    """capacity[type] = ceil(fraction * unconstrained demand) so that the greedy
    solver and the best-effort policies actually bind.  `chosen_acc_idx` is Allocation.accelerator
    of an unlimited solve (-1 = none)."""
    demand = np.zeros(img.T, dtype=np.int64)
    for s in range(img.S):
        a = int(chosen_acc_idx[s])
        if a < 0:
            continue
        cnt = int(img.perf_acc_count[int(img.srv_model[s]) * img.A + a])
        cnt = 1 if cnt <= 0 else cnt
        demand[img.acc_type[a]] += int(chosen_replicas[s]) * cnt * int(img.acc_multiplicity[a])
    img.type_capacity[:] = np.ceil(fraction * demand).astype(np.int64)
    return demand


def config1():
    """BASELINE config 1: Llama-3.1-8B, Premium, on L40S (the reference's own chart values:
    charts/workload-variant-autoscaler/templates/variantautoscaling.yaml, deploy/configmap-serviceclass.yaml)."""
    spec = {
        "acceleratorData": {"accelerators": [{"name": "L40S", "type": "L40S", "multiplicity": 1, "cost": 32.0}]},
        "modelData": {"models": [{"name": "meta/llama-3.1-8b", "acc": "L40S", "accCount": 1, "maxBatchSize": 512,
                                  "atTokens": 128, "decodeParms": {"alpha": 22.619, "beta": 0.181},
                                  "prefillParms": {"gamma": 226.19, "delta": 0.018}}]},
        "serviceClassData": {"serviceClasses": [{"name": "Premium", "priority": 1, "modelTargets": [
            {"model": "meta/llama-3.1-8b", "slo-itl": 24.0, "slo-ttft": 500.0, "slo-tps": 0.0}]}]},
        "serverData": {"servers": [{"name": "llama:default", "class": "Premium", "model": "meta/llama-3.1-8b",
                                    "keepAccelerator": True, "minNumReplicas": 1, "maxBatchSize": 512,
                                    "currentAlloc": {"accelerator": "L40S", "numReplicas": 1, "maxBatch": 512,
                                                     "cost": 32.0, "load": {"arrivalRate": 600.0, "avgInTokens": 128,
                                                                            "avgOutTokens": 128}}}]},
        "capacityData": {"count": [{"type": "L40S", "count": 64}]},
    }
    return SystemImage.from_spec(spec)


# BASELINE.json configs: (S, A, T, r_max, b_max)
CONFIGS = {
    1: dict(S=1, A=1, T=1, r_max=8, b_max=256),
    2: dict(S=32, A=4, T=4, r_max=64, b_max=512),
    3: dict(S=1000, A=8, T=8, r_max=64, b_max=512),
    4: dict(S=10000, A=8, T=4, r_max=64, b_max=512),
    5: dict(S=100000, A=16, T=8, r_max=64, b_max=512),
}


def baseline_config(k, n_servers=None):
    """System image + grid extents of BASELINE.json config k (n_servers overrides S for scaled runs)."""
    c = dict(CONFIGS[k])
    if k == 1:
        return config1(), c
    if n_servers is not None:
        c["S"] = int(n_servers)
    img = make_system(c["S"], c["A"], seed=k, n_types=c["T"])
    return img, c
