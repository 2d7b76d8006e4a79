"""Pins the CPU oracle against every exact value / known-answer case the reference's own
unit tests hold for the Analyze -> Optimize path.  Each test names the
reference test it restates (file:line relative to the reference tree).  CPU only."""
import math

import numpy as np
import pytest

F = np.float32


def _alloc_test_system(wva, arrival=0.0, ttft=100.0, itl=50.0, tps=0.0, max_batch=0, min_rep=1):
    """setupCompleteTestSystem, pkg/core/allocation_test.go:11-80"""
    spec = {
        "acceleratorData": {"accelerators": [{"name": "test-gpu", "cost": 100.0}]},
        "modelData": {"models": [{"name": "test-model", "acc": "test-gpu", "accCount": 1, "maxBatchSize": 16,
                                  "atTokens": 200, "decodeParms": {"alpha": 5.0, "beta": 2.0},
                                  "prefillParms": {"gamma": 10.0, "delta": 1.5}}]},
        "serviceClassData": {"serviceClasses": [{"name": "default", "priority": 10, "modelTargets": [
            {"model": "test-model", "slo-itl": itl, "slo-ttft": ttft, "slo-tps": tps}]}]},
        "serverData": {"servers": [{"name": "test-server", "class": "default", "model": "test-model",
                                    "minNumReplicas": min_rep, "maxBatchSize": max_batch,
                                    "currentAlloc": {"load": {"arrivalRate": arrival, "avgInTokens": 100,
                                                              "avgOutTokens": 200}}}]},
    }
    return wva.SystemImage.from_spec(spec)


def test_zero_load_getters(wva, oracle):
    """TestAllocation_Getters, pkg/core/allocation_test.go:82-141 (exact float32 pins)."""
    img = _alloc_test_system(wva)
    out, feas, _ = oracle.analyze_pairs(img)
    assert feas[0] == 1
    assert out.acc[0] == 0 and out.num_replicas[0] == 1 and out.batch_size[0] == 16
    assert out.cost[0] == F(100.0)
    # CreateAllocation's value (before Server.Calculate overwrites it with the penalty) equals cost;
    # with an empty current allocation the penalty is 0.1*(0+100)+(100-0):
    assert out.value[0] == F(F(0.1) * F(100.0)) + F(100.0)
    assert out.max_arrv_rate_per_replica[0] == F(0.3298969)
    max_rpm = F(F(out.max_arrv_rate_per_replica[0] * F(1000)) * F(60))      # Allocation.MaxRPM, allocation.go:236-238
    assert max_rpm == F(19793.814)


@pytest.mark.parametrize("b_acc,b_rep,b_cost,want", [
    (0, 2, 100.0, F(0.0)),                                                   # same accelerator same replicas
    (0, 3, 150.0, F(50.0)),                                                  # same accelerator different replicas
    (1, 2, 120.0, F(F(0.1) * F(220.0)) + F(20.0)),                           # different accelerator
])
def test_transition_penalty(oracle, b_acc, b_rep, b_cost, want):
    """TestAllocation_TransitionPenalty, pkg/core/allocation_test.go:238-287"""
    assert oracle.transition_penalty(0, 2, 100.0, b_acc, b_rep, b_cost) == want


@pytest.mark.parametrize("min_rep,max_batch,cost,acc_count,parms,want", [
    (0, 0, 100.0, 1, (5.0, 2.0, 10.0, 1.5), dict(acc=-1, rep=0, batch=0, cost=0.0)),
    (2, 0, 100.0, 1, (5.0, 2.0, 10.0, 1.5), dict(acc=0, rep=2, batch=16, cost=200.0)),
    (1, 8, 50.0, 2, (3.0, 1.0, 8.0, 2.0), dict(acc=0, rep=1, batch=8, cost=100.0)),
])
def test_zero_load_allocation_table(wva, oracle, min_rep, max_batch, cost, acc_count, parms, want):
    """TestZeroLoadAllocation, pkg/core/allocation_test.go:971-1130"""
    alpha, beta, gamma, delta = parms
    spec = {
        "acceleratorData": {"accelerators": [{"name": "test-gpu", "cost": cost}]},
        "modelData": {"models": [{"name": "m", "acc": "test-gpu", "accCount": acc_count, "maxBatchSize": 16,
                                  "atTokens": 200, "decodeParms": {"alpha": alpha, "beta": beta},
                                  "prefillParms": {"gamma": gamma, "delta": delta}}]},
        "serviceClassData": {"serviceClasses": [{"name": "c", "priority": 1, "modelTargets": [
            {"model": "m", "slo-itl": 50.0, "slo-ttft": 100.0}]}]},
        "serverData": {"servers": [{"name": "s", "class": "c", "model": "m", "minNumReplicas": min_rep,
                                    "maxBatchSize": max_batch,
                                    "currentAlloc": {"load": {"arrivalRate": 0, "avgInTokens": 100, "avgOutTokens": 200}}}]},
    }
    img = wva.SystemImage.from_spec(spec)
    out, feas, _ = oracle.analyze_pairs(img)
    assert feas[0] == 1
    assert out.acc[0] == want["acc"] and out.num_replicas[0] == want["rep"] and out.batch_size[0] == want["batch"]
    assert out.cost[0] == F(want["cost"]) and out.rho[0] == 0
    if want["rep"] > 0:
        a, b, g, d = F(alpha), F(beta), F(gamma), F(delta)
        assert out.itl[0] == a + b
        assert out.ttft[0] == g + d
        max_decode = a + b * F(out.batch_size[0])
        assert out.max_arrv_rate_per_replica[0] == F(out.batch_size[0]) / ((g + d) + max_decode)


@pytest.mark.parametrize("in_tok,batch,expected", [(0, 4.0, 0.0), (1000, 1.0, 11.0), (2000, 8.0, 26.0), (500, 2.5, 11.25)])
def test_prefill_time(oracle, in_tok, batch, expected):
    """TestPrefillParms_PrefillTime, pkg/analyzer/queueanalyzer_test.go:226-272"""
    assert abs(float(oracle.prefill_time(10.0, 0.001, in_tok, batch)) - expected) <= 1e-6


@pytest.mark.parametrize("batch,expected", [(1.0, 1.01), (4.0, 1.04), (8.0, 1.08), (2.5, 1.025)])
def test_decode_time(oracle, batch, expected):
    """TestDecodeParms_DecodeTime, pkg/analyzer/queueanalyzer_test.go:274-315"""
    assert abs(float(oracle.decode_time(1.0, 0.01, batch)) - expected) <= 1e-6


@pytest.mark.parametrize("x,v,tol,expected", [
    (1.0, 1.0, 0.01, True), (1.005, 1.0, 0.01, True), (1.02, 1.0, 0.01, False),
    (0.1, 0.0, 0.01, False), (1.0, 1.0, -0.01, True), (0.0, 0.0, 0.01, True)])
def test_within_tolerance(oracle, x, v, tol, expected):
    """TestWithinTolerance, pkg/analyzer/utils_test.go:9-70"""
    assert oracle.within_tolerance(x, v, tol) is expected


@pytest.mark.parametrize("func,lo,hi,y,want_err,ind", [
    (0, 0.0, 10.0, 4.0, False, 0),     # find square root
    (1, 1.0, 5.0, 6.0, False, 0),      # linear, target in range
    (1, 2.0, 5.0, 1.0, False, -1),     # target below range
    (1, 1.0, 3.0, 10.0, False, 1),     # target above range
    (2, 1.0, 5.0, -3.0, False, 0),     # decreasing function
    (1, 5.0, 1.0, 3.0, True, 0),       # invalid range
    (3, 4.0, 6.0, 5.0, True, 0),       # function evaluation error
    (1, 1.0, 5.0, 2.0, False, 0),      # target at boundary
])
def test_binary_search(oracle, func, lo, hi, y, want_err, ind):
    """TestBinarySearch, pkg/analyzer/utils_test.go:72-223"""
    x, got_ind, err = oracle.binary_search_testfunc(func, lo, hi, y)
    assert bool(err) == want_err
    if not want_err:
        assert got_ind == ind
        f = [lambda v: v * v, lambda v: 2 * v, lambda v: -v, lambda v: v][func]
        if ind == 0:
            assert abs(float(f(F(x))) - y) <= 0.1
        if ind == -1:
            assert x == F(lo)
        if ind == 1:
            assert x == F(hi)


def test_state_dependent_solve_sequence(oracle):
    """TestMM1ModelStateDependent_Solve, pkg/analyzer/queuemodel_test.go:325-400: one model, successive
    Solves; lambda=0 valid, negative invalid; rho in [0,1]; Little's law within 1e-4."""
    m = oracle.Model(5, [1.0, 2.0, 3.0])
    for lam, want_valid in [(0.5, True), (1.5, True), (2.8, True), (0.0, True), (-1.0, False)]:
        r = m.solve(lam, 1.0)
        assert bool(r["valid"]) == want_valid
        if want_valid:
            assert r["in_servers"] >= 0 or math.isnan(r["in_servers"])
            assert 0 <= r["rho"] <= 1
            if r["resp"] > 0 and r["throughput"] > 0:
                assert abs(float(r["throughput"] * r["resp"] - r["in_system"])) <= 1e-4
            p = m.probabilities()
            assert abs(p.sum() - 1.0) <= 1e-6


def test_state_dependent_first_solve_needs_k_ge_2(oracle):
    """QueueModel.Solve validity uses the STALE p[0] (queuemodel.go:30, mm1modelstatedependent.go:33-35):
    on a fresh model rho = 1, so K = 1 is invalid on its first Solve and stays invalid."""
    m = oracle.Model(1, [1.0])
    assert not m.solve(0.5)["valid"]
    assert not m.solve(0.5)["valid"]
    m2 = oracle.Model(2, [1.0])
    assert m2.solve(0.5)["valid"]


def test_state_dependent_utilization(oracle):
    """TestMM1ModelStateDependent_UtilizationCalculation, queuemodel_test.go:402-422"""
    m = oracle.Model(4, [2.0, 4.0, 6.0])
    r = m.solve(1.0, 1.0)
    assert r["valid"]
    assert abs(float(r["rho"]) - float(F(1.0) - F(m.probabilities()[0]))) <= 1e-6


def test_mm1k_validity_table(oracle):
    """TestQueueModel_Basic, queuemodel_test.go:9-102: MM1KModel(10); rho<K valid, rho>=K, lambda<0, mu<=0 invalid."""
    assert oracle.mm1k_solve(10, 9.9, 1.0)[0]["valid"]
    assert not oracle.mm1k_solve(10, 11.0, 1.0)[0]["valid"]
    assert not oracle.mm1k_solve(10, -1.0, 1.0)[0]["valid"]
    assert not oracle.mm1k_solve(10, 1.0, 0.0)[0]["valid"]
    assert not oracle.mm1k_solve(10, 1.0, -1.0)[0]["valid"]


def test_mm1k_vs_state_dependent(oracle):
    """TestMM1Models_Comparison, queuemodel_test.go:461-496: closed form vs chain agree within 1e-3."""
    K, rate, lam = 5, 3.0, 1.5
    a, _ = oracle.mm1k_solve(K, lam, rate)
    b = oracle.Model(K, [rate] * K).solve(lam, 1.0)
    assert a["valid"] and b["valid"]
    assert abs(float(a["in_system"] - b["in_system"])) <= 1e-3
    assert abs(float(a["throughput"] - b["throughput"])) <= 1e-3


def _test_analyzer(oracle):
    # testConfig, queueanalyzer_test.go:11-24; request {100, 10}
    return oracle.Analyzer(8, 16, 1.0, 0.01, 10.0, 0.001, 100, 10)


def test_analyze_error_cases(wva, oracle):
    """TestQueueAnalyzer_Analyze, queueanalyzer_test.go:357-446"""
    abi = wva.abi
    qa = _test_analyzer(oracle)
    lo, hi = qa.rate_range()
    assert 0 < lo < hi
    assert qa.analyze(0.0)[0] == abi.CAND_ERR_RATE_LE0
    assert qa.analyze(-1.0)[0] == abi.CAND_ERR_RATE_LE0
    assert qa.analyze(float(hi * F(1.1)))[0] == abi.CAND_ERR_RATE_MAX
    for rate in (lo * F(0.5), (lo + hi) * F(0.5), hi * F(0.9)):
        st, m = qa.analyze(float(rate))
        assert st == abi.CAND_OK
        assert m.throughput >= 0 and m.avg_resp_time >= 0 and m.avg_wait_time >= 0 and m.avg_num_in_serv >= 0
        assert 0 <= m.rho <= 1 and m.avg_prefill_time >= 0 and m.avg_token_time >= 0


@pytest.mark.parametrize("target,want_err", [((50.0, 5.0, 100.0), False), ((0.0, 0.0, 0.0), False),
                                             ((-1.0, 5.0, 100.0), True), ((50.0, -1.0, 100.0), True),
                                             ((50.0, 5.0, -1.0), True)])
def test_size_targets(wva, oracle, target, want_err):
    """TestQueueAnalyzer_Size, queueanalyzer_test.go:448-554"""
    cfg = np.array([(8, 16, 1.0, 0.01, 10.0, 0.001, 100, 10)], dtype=wva.abi.QUEUE_CONFIG_DTYPE)
    rates, metrics, achieved, status = oracle.queue_size(cfg, np.array([target], dtype=np.float32))
    assert bool(status[0]) == want_err
    if not want_err:
        assert (rates[0] >= 0).all() and (achieved[0] >= 0).all()


@pytest.mark.parametrize("cfg_bad", [(0, 16, 100, 10), (-1, 16, 100, 10), (8, -1, 100, 10), (8, 16, -1, 10), (8, 16, 100, 0)])
def test_config_and_request_checks(wva, oracle, cfg_bad):
    """TestConfiguration_Check / TestRequestSize_Check, queueanalyzer_test.go:92-224: N<=0, queue<0,
    in<0, out<1 are rejected by NewQueueAnalyzer."""
    n, q, i, o = cfg_bad
    cfg = np.array([(n, q, 1.0, 0.01, 10.0, 0.001, i, o)], dtype=wva.abi.QUEUE_CONFIG_DTYPE)
    _, status = oracle.queue_analyze(cfg, np.array([1.0], dtype=np.float32))
    assert status[0] == wva.abi.CAND_ERR_CONFIG


@pytest.mark.parametrize("serv_time", [20.0, 50.0, 100.0])
def test_effective_concurrency_bounds(oracle, serv_time):
    """TestEffectiveConcurrency, queueanalyzer_test.go:556-600"""
    n = oracle.effective_concurrency(serv_time, 1.0, 0.01, 10.0, 0.001, 100, 10, 8)
    assert 0 <= n <= 8


def test_create_allocation_nil_table(wva, oracle):
    """TestCreateAllocation, pkg/core/allocation_test.go:579-776 (nil / non-nil table)."""
    # zero load -> non-nil
    assert oracle.analyze_pairs(_alloc_test_system(wva))[1][0] == 1
    # strict targets TTFT 1, ITL 0.1 at 1200 req/min -> nil
    assert oracle.analyze_pairs(_alloc_test_system(wva, arrival=1200.0, ttft=1.0, itl=0.1))[1][0] == 0
    # TPS branch -> non-nil, replicas > 0
    out, feas, _ = oracle.analyze_pairs(_alloc_test_system(wva, arrival=60.0, ttft=2000.0, itl=500.0, tps=2.0))
    assert feas[0] == 1 and out.num_replicas[0] > 0 and out.acc[0] == 0
    # arrival branch
    out, feas, _ = oracle.analyze_pairs(_alloc_test_system(wva, arrival=120.0, ttft=2000.0, itl=500.0))
    assert feas[0] == 1 and out.num_replicas[0] > 0
    # batch override 12
    out, feas, _ = oracle.analyze_pairs(_alloc_test_system(wva, arrival=60.0, ttft=2000.0, itl=500.0, max_batch=12))
    assert feas[0] == 1 and out.batch_size[0] == 12 and out.num_replicas[0] > 0
    # missing perf data / missing target -> nil
    img = _alloc_test_system(wva); img.perf_valid[:] = 0
    assert oracle.analyze_pairs(img)[1][0] == 0
    img = _alloc_test_system(wva); img.srv_target_valid[:] = 0
    assert oracle.analyze_pairs(img)[1][0] == 0


def test_scale_direction(wva, oracle):
    """TestAllocation_Scale, allocation_test.go:778-887: 30 -> 360 req/min increases replicas."""
    lo = oracle.analyze_pairs(_alloc_test_system(wva, arrival=30.0, ttft=2000.0, itl=500.0))[0].num_replicas[0]
    hi = oracle.analyze_pairs(_alloc_test_system(wva, arrival=360.0, ttft=2000.0, itl=500.0))[0].num_replicas[0]
    assert hi - lo > 0


def _optimizer_fixture(wva, arrival_rpm):
    # internal/optimizer/optimizer_test.go:245-457: alpha 20.28, beta 0.72, maxBatch 4, SLO ITL 80 / TTFT 500, cost 40
    spec = {
        "acceleratorData": {"accelerators": [{"name": "A100", "type": "A100", "multiplicity": 1, "cost": 40.0}]},
        "modelData": {"models": [{"name": "m", "acc": "A100", "accCount": 1, "maxBatchSize": 4, "atTokens": 0,
                                  "decodeParms": {"alpha": 20.28, "beta": 0.72}, "prefillParms": {"gamma": 0, "delta": 0}}]},
        "serviceClassData": {"serviceClasses": [{"name": "Premium", "priority": 1, "modelTargets": [
            {"model": "m", "slo-itl": 80.0, "slo-ttft": 500.0}]}]},
        "serverData": {"servers": [{"name": "va:default", "class": "Premium", "model": "m", "keepAccelerator": True,
                                    "minNumReplicas": 1, "maxBatchSize": 4,
                                    "currentAlloc": {"accelerator": "A100", "numReplicas": 1, "cost": 40.0,
                                                     "load": {"arrivalRate": arrival_rpm, "avgInTokens": 20,
                                                              "avgOutTokens": 200 if arrival_rpm else 0}}}]},
    }
    return wva.SystemImage.from_spec(spec)


def test_optimizer_integration_vectors(wva, oracle):
    """internal/optimizer/optimizer_test.go:333 (no load => replicas == 1) and :455 (20 req/s => replicas > 1)."""
    img = _optimizer_fixture(wva, 0.0)
    pairs, feas, _ = oracle.analyze_pairs(img)
    chosen_acc, chosen = oracle.solve(img, pairs, feas, unlimited=True)
    assert chosen_acc[0] == 0 and chosen.num_replicas[0] == 1
    img = _optimizer_fixture(wva, 1200.0)
    pairs, feas, _ = oracle.analyze_pairs(img)
    chosen_acc, chosen = oracle.solve(img, pairs, feas, unlimited=True)
    assert chosen_acc[0] == 0 and chosen.num_replicas[0] > 1


def test_survey_derived_vectors(wva, oracle):
    """Not pinned by the reference: values derived with an independent float32-faithful Python
    restatement during the survey (SURVEY.md 8c "derived smoke vectors").  Two independent
    restatements agreeing bit-for-bit is the strongest evidence available without a Go toolchain."""
    img = _optimizer_fixture(wva, 1200.0)
    out, feas, steps = oracle.analyze_pairs(img)
    assert feas[0] and out.num_replicas[0] == 43 and out.batch_size[0] == 4 and out.cost[0] == F(1720.0)
    assert out.itl[0] == F(22.345161) and out.ttft[0] == F(496.41406) and out.rho[0] == F(0.51705664)
    assert out.max_arrv_rate_per_replica[0] == F(F(0.46595168) / F(1000))
    assert steps == 26 * 2 * 45                       # 26 Solves of K+1 = 45 states
    cases = [(120.0, 0, 2, F(13.86302), F(677.79877), F(0.21459173)), (30.0, 0, 1, None, None, None),
             (360.0, 0, 5, None, None, None), (60.0, 12, 1, None, None, None)]
    for arrival, mb, rep, itl, ttft, rho in cases:
        o, f, _ = oracle.analyze_pairs(_alloc_test_system(wva, arrival=arrival, ttft=2000.0, itl=500.0, max_batch=mb))
        assert f[0] and o.num_replicas[0] == rep
        if itl is not None:
            assert o.itl[0] == itl and o.ttft[0] == ttft and o.rho[0] == rho
    # BASELINE config 1 (Llama-3.1-8B on L40S, Premium): 600 req/min => 5 replicas
    c1 = wva.synth.config1()
    o, f, steps = oracle.analyze_pairs(c1)
    assert f[0] and o.num_replicas[0] == 5 and o.batch_size[0] == 512 and o.cost[0] == F(160.0)
    assert o.itl[0] == F(23.991173) and o.ttft[0] == F(243.6568) and o.rho[0] == F(0.012853656)
    assert steps == 44 * 2 * 5633
    for rpm, mb, rep in [(60.0, 512, 1), (6000.0, 512, 50), (600.0, 8, 7), (600.0, 64, 5), (600.0, 256, 5)]:
        c1.srv_arrival_rpm[0] = rpm; c1.srv_max_batch[0] = mb
        o, f, _ = oracle.analyze_pairs(c1)
        assert f[0] and o.num_replicas[0] == rep, (rpm, mb, o.num_replicas[0])
    c1.srv_arrival_rpm[0] = 600.0; c1.srv_max_batch[0] = 1          # ITL target below bounded region
    assert oracle.analyze_pairs(c1)[1][0] == 0
    # the chart's own SLO (tpot 10 / ttft 1000) is infeasible on L40S ...
    c1 = wva.synth.config1(); c1.srv_slo_itl[0] = 10.0; c1.srv_slo_ttft[0] = 1000.0
    assert oracle.analyze_pairs(c1)[1][0] == 0
    # ... and on H100 needs 1 / 3 replicas, the latter after all 100 bisection iterations (122 Solves)
    c1.perf_alpha[0], c1.perf_beta[0], c1.perf_gamma[0], c1.perf_delta[0] = 7.470, 0.044, 15.415, 0.000337
    c1.acc_cost[0] = 100.0
    o, f, _ = oracle.analyze_pairs(c1)
    assert f[0] and o.num_replicas[0] == 1 and o.itl[0] == F(7.966144) and o.ttft[0] == F(15.9014015)
    c1.srv_arrival_rpm[0] = 6000.0
    o, f, steps = oracle.analyze_pairs(c1)
    assert f[0] and o.num_replicas[0] == 3 and o.itl[0] == F(9.264938) and o.ttft[0] == F(17.174692)
    assert steps == 122 * 2 * 5633
