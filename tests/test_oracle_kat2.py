"""More of the reference's own test cases replayed on the oracle (and, where the product has a host path, on the
product): the MM1KModel behaviour tests, the remaining BinarySearch / Eval* cases, NewQueueAnalyzer's configuration
checks, EffectiveConcurrency's bounds and the SolveUnlimited selection cases.  Citations are reference file:line."""
import math

import numpy as np
import pytest

from workload_variant_autoscaler_b200 import Allocs, Fleet, greedy_solve
from workload_variant_autoscaler_b200._abi import ALLOC_COLUMNS


# ---- pkg/analyzer/queuemodel_test.go: MM1KModel ----------------------------------------------------------------
def test_mm1k_creation(oracle_mod):  # queuemodel_test.go:122-150
    for K in (5, 50, 500, 1):
        m = oracle_mod.MM1K(K)
        assert m.K == K and m.probs().shape == (K + 1,)
        # GetRhoMax = K: rho just below K is valid, rho = K is not (queuemodel.go:31)
        assert m.solve(float(K) * 0.999, 1.0)["is_valid"] == 1 and m.solve(float(K), 1.0)["is_valid"] == 0


def test_mm1k_probability_calculation(oracle_mod):  # queuemodel_test.go:152-222
    m = oracle_mod.MM1K(3)
    for lam, mu in ((0.5, 2.0), (1.5, 2.0), (1.9, 2.0), (2.0, 2.0)):
        st = m.solve(lam, mu)
        assert st["is_valid"]
        p = m.probs()
        assert (p >= 0).all() and abs(p.sum() - 1.0) <= 1e-6
        assert 0 <= st["throughput"] <= lam


def test_mm1k_edge_cases(oracle_mod):  # queuemodel_test.go:224-274
    for K, lam, mu in ((1, 0.5, 1.0), (10, 0.001, 1.0), (10, 1.0, 1000.0)):
        st = oracle_mod.MM1K(K).solve(lam, mu)
        assert st["is_valid"] and st["avg_num_in_system"] >= 0 and st["throughput"] >= 0


def test_mm1k_against_state_dependent_model_with_constant_rates(oracle_mod):  # queuemodel_test.go:461-496
    K, rate, lam = 5, 3.0, 1.5
    a = oracle_mod.MM1K(K).solve(lam, rate)
    b = oracle_mod.StateDependentModel(K, [rate] * K).solve(lam, 1.0)
    assert a["is_valid"] and b["is_valid"]
    assert abs(a["avg_num_in_system"] - b["avg_num_in_system"]) <= 1e-3
    assert abs(a["throughput"] - b["throughput"]) <= 1e-3


def test_mm1k_littles_law(oracle_mod):  # queuemodel_test.go:498-533
    for lam, mu in ((0.5, 2.0), (1.5, 3.0), (2.8, 4.0)):
        st = oracle_mod.MM1K(10).solve(lam, mu)
        assert st["is_valid"]
        assert abs(st["avg_num_in_system"] - st["throughput"] * st["avg_resp_time"]) <= 1e-4


def test_queue_model_basic_statistics(oracle_mod):  # queuemodel_test.go:9-102 (the getters after Solve(1, 2), K = 10)
    st = oracle_mod.MM1K(10).solve(1.0, 2.0)
    assert st["is_valid"] and st["rho"] == 0.5 and st["avg_serv_time"] == 0.5
    assert st["avg_resp_time"] > 0 and st["avg_wait_time"] >= 0 and st["avg_queue_length"] >= 0
    assert st["avg_num_in_system"] == pytest.approx(0.5 / 0.5 - 11 * 0.5 ** 11 / (1 - 0.5 ** 11), rel=1e-5)


# ---- pkg/analyzer/utils_test.go ----------------------------------------------------------------------------------
def test_binary_search_edge_cases(oracle_mod):  # utils_test.go:225-289: none of them is an error
    bs = oracle_mod.binary_search
    assert bs(1.0, 10.0, 5.0, lambda x: 5.0)[0] == 0            # constant function, target matches
    assert bs(1.0, 10.0, 3.0, lambda x: 5.0)[0] == 0            # constant function, target does not match
    assert bs(1.0, 5.0, 5.0, lambda x: 1.0 if x < 3.0 else 10.0)[0] == 0   # step function
    err, x, ind = bs(3.0, 3.0, 6.0, lambda x: 2 * x)            # zero range: the boundary hits the target
    assert err == 0 and x == 3.0 and ind == 0


def test_eval_serv_and_waiting_time_cases(oracle_mod):  # utils_test.go:291-381 (rates 1..5, K = 5)
    m = oracle_mod.StateDependentModel(5, [1.0, 2.0, 3.0, 4.0, 5.0])
    for lam in (0.5, 0.0, 10.0, 0.1, 1.0):
        st = m.solve(lam)
        assert st["is_valid"], lam                                   # EvalServTime / EvalWaitingTime return no error
        if lam > 0:
            assert st["avg_serv_time"] >= 0 and st["avg_wait_time"] >= 0


def test_binary_search_precision(oracle_mod):  # utils_test.go:610-644: y = 2x + 3 on [0, 10], target 9 -> x* = 3
    err, x, ind = oracle_mod.binary_search(0.0, 10.0, 9.0, lambda v: 2 * v + 3)
    assert err == 0 and ind == 0 and abs(x - 3.0) <= 1e-4 and abs((2 * x + 3) - 9.0) <= 1e-4


# ---- pkg/analyzer/queueanalyzer_test.go ------------------------------------------------------------------------
def test_configuration_check(oracle_mod):  # queueanalyzer_test.go:92-176 (the nil-pointer cases have no analogue in C)
    ok = dict(alpha=1.0, beta=0.01, gamma=10.0, delta=0.001, in_tokens=100, out_tokens=10)
    assert oracle_mod.Analyzer(8, 16, **ok) is not None
    for max_batch, max_queue in ((0, 16), (-1, 16), (8, -1)):
        with pytest.raises(ValueError):
            oracle_mod.Analyzer(max_batch, max_queue, **ok)


def test_request_size_check(oracle_mod):  # queueanalyzer_test.go:178-224
    cfg = dict(max_batch=8, max_queue=16, alpha=1.0, beta=0.01, gamma=10.0, delta=0.001)
    for in_tok, out_tok in ((100, 10), (0, 10), (100, 1)):
        assert oracle_mod.Analyzer(in_tokens=in_tok, out_tokens=out_tok, **cfg) is not None
    for in_tok, out_tok in ((-1, 10), (100, 0), (100, -1)):
        with pytest.raises(ValueError):
            oracle_mod.Analyzer(in_tokens=in_tok, out_tokens=out_tok, **cfg)


def test_effective_concurrency_bounds(oracle_mod):  # queueanalyzer_test.go:556-600
    L = oracle_mod.lib()
    for serv in (20.0, 50.0, 100.0):
        v = L.wvao_effective_concurrency(serv, 1.0, 0.01, 10.0, 0.001, 100, 10, 8)
        assert 0.0 <= v <= 8.0
    # the clamp itself (queueanalyzer.go:296-302): far below the base time -> 0, far above -> maxBatchSize
    assert L.wvao_effective_concurrency(0.0, 1.0, 0.01, 10.0, 0.001, 100, 10, 8) == 0.0
    assert L.wvao_effective_concurrency(1e9, 1.0, 0.01, 10.0, 0.001, 100, 10, 8) == 8.0


# ---- pkg/solver/solver_test.go: SolveUnlimited -----------------------------------------------------------------
def _solver_spec(servers, accs=("A100", "H100"), caps=(4, 2)):
    # solver_test.go:283-375: accelerators without a type or cost, perf data without parameters, one service class
    return {
        "acceleratorData": {"accelerators": [{"name": a} for a in accs]},
        "modelData": {"models": [{"name": "llama-7b", "acc": a, "accCount": 1} for a in accs]},
        "capacityData": {"count": [{"type": a, "count": c} for a, c in zip(accs, caps)]},
        "serviceClassData": {"serviceClasses": [{"name": "default", "priority": 1, "modelTargets": [
            {"model": "llama-7b", "slo-itl": 9, "slo-ttft": 1000}]}]},
        "serverData": {"servers": servers},
        "optimizerData": {"optimizer": {"unlimited": True, "saturationPolicy": "None"}},
    }


def _server(name, acc, replicas, keep=True):
    return {"name": name, "class": "default", "model": "llama-7b", "keepAccelerator": keep, "minNumReplicas": 1,
            "maxBatchSize": 512, "currentAlloc": {"accelerator": acc, "numReplicas": replicas}}


def test_solve_unlimited_allocates_servers_with_positive_replicas(oracle_mod):  # solver_test.go:280-425
    f = Fleet.from_spec(_solver_spec([_server("server1", "A100", 2), _server("server2", "H100", 1)]))
    assert f.unlimited
    _, win = oracle_mod.solve(f)
    assert win["feasible"].sum() >= 1
    assert (win["replicas"][win["feasible"].astype(bool)] > 0).all()
    # keepAccelerator: each server stays on its current accelerator (server.go:70-82)
    assert [int(a) for a in win["acc"]] == [0, 1]


def test_solve_unlimited_edge_cases(oracle_mod):  # solver_test.go:427-517
    # no servers at all: nothing to do, no error
    f0 = Fleet.from_spec(_solver_spec([]))
    cand, win = oracle_mod.solve(f0)
    assert win.size == 0
    # a server whose candidate table is empty keeps no allocation (the reference never runs Calculate there)
    f1 = Fleet.from_spec(_solver_spec([_server("test-server", "A100", 1)], accs=("A100",), caps=(2,)))
    empty = np.zeros((1, 1), dtype=oracle_mod.calculate(f1).dtype)
    empty["acc"] = -1
    _, win = oracle_mod.solve(f1, empty)
    assert not win["feasible"][0]


@pytest.mark.parametrize("values,want", [((100.0, 50.0), 1), ((50.0, 100.0), 0), ((7.0, 7.0), 0)])
def test_solve_unlimited_selects_the_minimum_value(oracle_mod, values, want):  # solver_test.go:519-623, 724-833
    f = Fleet.from_spec(_solver_spec([_server("server1", "", 0, keep=False)]))
    cand = oracle_mod.calculate(f)
    assert cand.shape == (1, 2) and cand["feasible"].all(), "both accelerators must be candidates"
    for a, v in enumerate(values):
        cand["value"][0, a] = v
    _, win = oracle_mod.solve(f, cand.copy())
    assert win["feasible"][0] and win["acc"][0] == want and win["value"][0] == min(values)
    # the product's host pass (wva_solve_greedy falls through to the unlimited argmin for an unlimited fleet)
    c = Allocs(2)
    for name, _ in ALLOC_COLUMNS:
        getattr(c, name)[:] = cand.reshape(-1)[name]
    _, win_g = greedy_solve(f, c)
    assert int(win_g.acc[0]) == want and float(win_g.value[0]) == min(values)


# ---- pkg/core/server_test.go: Server.Calculate; pkg/manager/manager_test.go: Manager.Optimize -------------------
def _calc_spec(rate, cur=None, unlimited=True, policy="None", capacity=10):
    # server_test.go:470-513 / manager_test.go:63-131: test-gpu (cost 100), test-model (1 instance, maxBatch 16,
    # atTokens 200, alpha 5 beta 2 gamma 10 delta 1.5), class "default" priority 5, TTFT 2000 / ITL 500
    cur_alloc = {"load": {"arrivalRate": rate, "avgInTokens": 100, "avgOutTokens": 200}}
    cur_alloc.update(cur or {})
    return {
        "acceleratorData": {"accelerators": [{"name": "test-gpu", "type": "gpu", "multiplicity": 1, "cost": 100.0}]},
        "modelData": {"models": [{"name": "test-model", "acc": "test-gpu", "accCount": 1, "maxBatchSize": 16,
                                  "atTokens": 200, "decodeParms": {"alpha": 5.0, "beta": 2.0},
                                  "prefillParms": {"gamma": 10.0, "delta": 1.5}}]},
        "capacityData": {"count": [{"type": "gpu", "count": capacity}]},
        "serviceClassData": {"serviceClasses": [{"name": "default", "priority": 5, "modelTargets": [
            {"model": "test-model", "slo-ttft": 2000.0, "slo-itl": 500.0, "slo-tps": 0.0}]}]},
        "serverData": {"servers": [{"name": "test-server", "model": "test-model", "class": "default",
                                    "minNumReplicas": 1, "currentAlloc": cur_alloc}]},
        "optimizerData": {"optimizer": {"unlimited": unlimited, "saturationPolicy": policy}},
    }


def test_server_calculate_cases(oracle_mod):  # server_test.go:468-614
    # complete system: Calculate creates a candidate allocation
    f = Fleet.from_spec(_calc_spec(60))
    cand = oracle_mod.calculate(f)
    assert cand["feasible"][0, 0] and cand["replicas"][0, 0] >= 1
    # with a current allocation the candidate's value is the transition penalty from it (allocation.go:291-300)
    f = Fleet.from_spec(_calc_spec(60, {"accelerator": "test-gpu", "numReplicas": 2, "maxBatch": 8, "cost": 150.0}))
    c = oracle_mod.calculate(f)[0, 0]
    assert c["feasible"]
    want = oracle_mod.lib().wvao_transition_penalty(f.accel_penalty_factor, 0, 2, 150.0, int(c["acc"]),
                                                    int(c["replicas"]), float(c["cost"]))
    assert np.float32(c["value"]) == np.float32(want)
    # empty system: no accelerators, no candidates
    spec = _calc_spec(60)
    spec["acceleratorData"]["accelerators"] = []
    spec["modelData"]["models"] = []
    f = Fleet.from_spec(spec)
    assert f.n_acc == 0 and oracle_mod.calculate(f).size == 0


@pytest.mark.parametrize("unlimited,policy", [(False, "None"), (True, "PriorityExhaustive")])
def test_manager_optimize_cases(oracle_mod, unlimited, policy):  # manager_test.go:61-281
    f = Fleet.from_spec(_calc_spec(120.0, unlimited=unlimited, policy=policy))
    cand, win = oracle_mod.solve(f)
    assert win["feasible"][0] and win["replicas"][0] > 0 and win["cost"][0] > 0
    # the product's host pass over the same candidate table gives the same solution
    c = Allocs(f.n_servers * f.n_acc)
    for name, _ in ALLOC_COLUMNS:
        getattr(c, name)[:] = oracle_mod.calculate(f).reshape(-1)[name]
    _, win_g = greedy_solve(f, c)
    assert int(win_g.replicas[0]) == int(win["replicas"][0]) and float(win_g.cost[0]) == float(win["cost"][0])
    # what Optimize leaves behind: AllocateByType counts the replicas against the type (system.go:271-300)
    tot = oracle_mod.allocate_by_type(f, win)
    assert tot["present"][0] == 1 and tot["count"][0] == win["replicas"][0] and tot["limit"][0] == 10


def test_manager_optimize_on_an_empty_system(oracle_mod):  # manager_test.go:208-232, 283-332
    for unlimited, policy in ((False, "None"), (True, "PriorityExhaustive")):
        f = Fleet.from_spec({"optimizerData": {"optimizer": {"unlimited": unlimited, "saturationPolicy": policy}}})
        cand, win = oracle_mod.solve(f)
        assert f.n_servers == 0 and win.size == 0 and cand.size == 0
