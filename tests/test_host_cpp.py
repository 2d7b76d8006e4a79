"""GPU: the C++ host mirror of the reference's Go API (inferno-autoscaler_b200/host) drives the
reconcile call sequence of internal/controller (variantautoscaling_controller.go:143-166) over the
C-ABI; results are checked against the oracle and the reference's integration-test expectations."""
import json
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F = np.float32


def _run():
    import __graft_entry__ as g
    g.build_cuda(); g.build_host()
    exe = os.path.join(ROOT, "inferno-autoscaler_b200", "host", "host_test")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    return {d["scenario"]: d for d in map(json.loads, out.stdout.strip().splitlines())}


def _spec(arrival, out_tok, itl, ttft, two_acc=False):
    accs = [{"name": "A100", "type": "A100", "multiplicity": 1, "cost": 40.0}]
    models = [{"name": "m", "acc": "A100", "accCount": 1, "maxBatchSize": 4, "atTokens": 0,
               "decodeParms": {"alpha": 20.28, "beta": 0.72}, "prefillParms": {"gamma": 0, "delta": 0}}]
    if two_acc:
        accs.append({"name": "H100", "type": "H100", "multiplicity": 1, "cost": 100.0})
        models.append({"name": "m", "acc": "H100", "accCount": 1, "maxBatchSize": 4, "atTokens": 0,
                       "decodeParms": {"alpha": 7.47, "beta": 0.044}, "prefillParms": {"gamma": 0, "delta": 0}})
    return {
        "acceleratorData": {"accelerators": accs}, "modelData": {"models": models},
        "serviceClassData": {"serviceClasses": [{"name": "Premium", "priority": 1, "modelTargets": [
            {"model": "m", "slo-itl": itl, "slo-ttft": ttft}]}]},
        "serverData": {"servers": [{"name": "va:default", "class": "Premium", "model": "m", "keepAccelerator": not two_acc,
                                    "minNumReplicas": 1, "maxBatchSize": 4,
                                    "currentAlloc": {"accelerator": "A100", "numReplicas": 1, "cost": 40.0,
                                                     "load": {"arrivalRate": arrival, "avgInTokens": 20, "avgOutTokens": out_tok}}}]},
        "capacityData": {"count": [{"type": "A100", "count": 10}, {"type": "H100", "count": 8}] if two_acc else []},
    }


def test_reconcile_sequence_through_cpp_host(wva, oracle):
    r = _run()
    # internal/optimizer/optimizer_test.go:333: no load => replicas == minNumReplicas
    assert r["no_load"]["error"] is None and r["no_load"]["optimized"]["va"]["replicas"] == 1
    # :455: 20 req/s => replicas > 1 (43 with the survey's derived vector), keyed by bare va.Name
    so = r["scale_out"]["optimized"]["va"]
    assert so["accelerator"] == "A100" and so["replicas"] == 43
    assert F(so["itl"]) == F(22.345161) and F(so["ttft"]) == F(496.41406) and F(so["cost"]) == F(1720.0)
    assert r["scale_out"]["by_type"]["A100"]["count"] == 43
    # infeasible SLO => CreateAllocation nil => empty solution => the error of optimizer.go:38-40
    assert r["infeasible_slo"]["error"]["code"] == wva.abi.ENOSOLUTION and r["infeasible_slo"]["analyze_allocations"] == 0
    # limited capacity, greedy + PriorityExhaustive against the oracle
    img = wva.SystemImage.from_spec(_spec(1200.0, 200, 80.0, 500.0, two_acc=True))
    pairs, feas, _ = oracle.analyze_pairs(img)
    acc, chosen = oracle.solve(img, pairs, feas, unlimited=False, policy=wva.abi.POLICY_PRIORITY_EXHAUSTIVE)
    lg = r["limited_greedy"]["optimized"]["va"]
    assert r["limited_greedy"]["analyze_allocations"] == int(feas.sum())
    assert lg["accelerator"] == img.acc_names[chosen.acc[0]] and lg["replicas"] == int(chosen.num_replicas[0])
    assert F(lg["cost"]) == chosen.cost[0] and F(lg["itl"]) == chosen.itl[0]
    # Allocation.Scale / ReAllocate (allocation.go:165-207): CreateAllocation on every accelerator, candidate or
    # not (keepAccelerator on: Server.Calculate sized the current accelerator only)
    sr = r["scale_realloc"]
    free = wva.SystemImage.from_spec(_spec(1200.0, 200, 80.0, 500.0, two_acc=True))
    free.srv_keep_acc[:] = 0
    cp, cf, _ = oracle.analyze_pairs(free)
    a100 = free.acc_names.index("A100")
    assert sr["candidates"] == 1 and sr["candidates_after"] == 1 and sr["nil_cases"] == 2
    assert cf[a100] and sr["scale"]["replicas"] == int(cp.num_replicas[a100]) and sr["scale"]["inc"] == int(cp.num_replicas[a100]) - 1
    assert F(sr["scale"]["cost"]) == cp.cost[a100] and F(sr["scale"]["value"]) == cp.cost[a100]      # value = cost (:161)
    best, min_val = None, F(0)
    for a in range(free.A):                      # allocation.go:193-201, map order -> ascending index
        if cf[a] and (min_val == 0 or cp.cost[a] < min_val):
            min_val, best = cp.cost[a], a
    assert sr["realloc"]["accelerator"] == free.acc_names[best] and sr["realloc"]["replicas"] == int(cp.num_replicas[best])
    assert F(sr["realloc"]["cost"]) == cp.cost[best]
    # incremental updates (system.go:99-171): a second server with half the load is added, sized and removed
    inc = r["incremental"]
    two = _spec(1200.0, 200, 80.0, 500.0)
    vb = json.loads(json.dumps(two["serverData"]["servers"][0])); vb["name"] = "vb:default"
    vb["currentAlloc"]["load"]["arrivalRate"] = 600.0
    two["serverData"]["servers"].append(vb)
    img2 = wva.SystemImage.from_spec(two)
    p2, f2, _ = oracle.analyze_pairs(img2)
    a2, c2 = oracle.solve(img2, p2, f2, unlimited=True)
    assert inc["servers_after_add"] == 2 and inc["va_after_add"] == int(c2.num_replicas[0]) == 43
    assert inc["vb_after_add"] == int(c2.num_replicas[1])
    assert inc["removed"] == 1 and inc["removed_again"] == 0
    assert inc["servers_after_remove"] == 1 and inc["va_after_remove"] == 43


def test_delta_uploads_and_solved_state_through_cpp_host(wva, oracle):
    """system.go:99-171 on the resident image: replacing one of 401 servers moves one row (54 B), a capacity
    change moves T counters; ReAllocate after Solve leaves the solved state usable."""
    d = _run()["delta"]
    assert d["servers"] == 401
    assert d["delta_bytes"] == 54 and d["capacity_bytes"] == 16 and d["full_bytes"] > 50 * d["delta_bytes"]
    assert d["replicas_after"] > d["replicas_before"] >= 1
    assert d["realloc"] in ("A100", "H100") and d["by_type_total"] > 0


def test_adapters_in_and_json_out_through_cpp_host(wva, oracle):
    """ConfigMap-shaped strings in (internal/utils/utils.go:108-311), AllocationSolution JSON out
    (pkg/config/types.go:123-143), on the reference's own chart values for Llama-3.1-8B / L40S (BASELINE config 1)."""
    a = _run()["adapters"]
    assert a["accelerators"] == 1                      # the accelerator whose cost does not parse is skipped (utils.go:121-125)
    assert a["err_ok"] == "" and "alpha" in a["err_bad"]
    assert a["in_tokens"] == 128 and a["itl_in"] == 0 and a["min_replicas"] == 1 and a["server_batch"] == 512
    assert a["found"] == 1 and a["missing"] == 0
    sol = a["solution"]["allocations"]["llama:default"]
    # SURVEY 8c derived vector for this fixture: 600 req/min -> N = 512, 5 replicas, cost 160, itl 23.991173, ttft 243.6568
    assert sol["accelerator"] == "L40S" and sol["numReplicas"] == 5 and sol["maxBatch"] == 512
    assert F(sol["cost"]) == F(160.0) and F(sol["itlAverage"]) == F(23.991173) and F(sol["ttftAverage"]) == F(243.6568)
    assert sol["load"] == {"arrivalRate": 600, "avgInTokens": 128, "avgOutTokens": 128}
    opt = a["optimized"]
    assert opt["accelerator"] == "L40S" and opt["numReplicas"] == 5 and opt["maxBatch"] == 512 and F(opt["rho"]) == F(0.012853656)
    # the sweep's winner for the same server against the oracle
    img = wva.synth.config1()
    best, _, _, _ = oracle.analyze_grid(img, 8, 256, want_cube=False)
    sw = opt["sweep"]
    assert (sw["numReplicas"], sw["maxBatch"]) == (int(best["replicas"][0]), int(best["batch"][0]))
    assert F(sw["itlAverage"]) == best["itl"][0] and F(sw["ttftAverage"]) == best["ttft"][0] and F(sw["cost"]) == best["cost"][0]
    # encoding/json float32 formatting: shortest round-trip digits, exponent form below 1e-6 and from 1e21
    assert a["floats"] == ["1e-7", "1e+21", "0.1", "16777216", "-0.000025", "3.4028235e+38"]
