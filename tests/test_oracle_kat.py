"""The oracle against every known-answer value the reference's own tests hold for this path
(file:line of each reference test is cited at the case)."""
import numpy as np
import pytest

from workload_variant_autoscaler_b200 import Fleet

F = np.float32


def test_prefill_time_kats(oracle_mod):  # pkg/analyzer/queueanalyzer_test.go:226-272
    L = oracle_mod.lib()
    for in_tok, batch, want in ((0, 4.0, 0.0), (1000, 1.0, 11.0), (2000, 8.0, 26.0), (500, 2.5, 11.25)):
        assert abs(L.wvao_prefill_time(10.0, 0.001, in_tok, batch) - want) <= 1e-6


def test_decode_time_kats(oracle_mod):  # queueanalyzer_test.go:274-315
    L = oracle_mod.lib()
    for batch, want in ((1.0, 1.01), (4.0, 1.04), (8.0, 1.08), (2.5, 1.025)):
        assert abs(L.wvao_decode_time(1.0, 0.01, batch) - want) <= 1e-6


def test_within_tolerance_kats(oracle_mod):  # pkg/analyzer/utils_test.go:9-70
    L = oracle_mod.lib()
    cases = ((1.0, 1.0, 0.01, True), (1.005, 1.0, 0.01, True), (1.02, 1.0, 0.01, False), (0.1, 0.0, 0.01, False),
             (1.0, 1.0, -0.01, True), (0.0, 0.0, 0.01, True))
    for x, v, tol, want in cases:
        assert bool(L.wvao_within_tolerance(x, v, tol)) == want


def test_binary_search_kats(oracle_mod):  # utils_test.go:72-223
    bs = oracle_mod.binary_search
    err, x, ind = bs(0.0, 10.0, 4.0, lambda v: F(v) * F(v))
    assert (err, ind) == (0, 0) and abs(x * x - 4.0) <= 0.1
    err, x, ind = bs(1.0, 5.0, 6.0, lambda v: 2 * v)
    assert (err, ind) == (0, 0) and abs(2 * x - 6.0) <= 0.1
    assert bs(2.0, 5.0, 1.0, lambda v: 2 * v)[1:] == (2.0, -1)      # below range -> xMin
    assert bs(1.0, 3.0, 10.0, lambda v: 2 * v)[1:] == (3.0, 1)      # above range -> xMax
    err, x, ind = bs(1.0, 5.0, -3.0, lambda v: -v)
    assert (err, ind) == (0, 0) and abs(-x + 3.0) <= 0.1
    assert bs(5.0, 1.0, 3.0, lambda v: 2 * v)[0] != 0                # invalid range -> error

    def err_fn(v):
        if v > 5.0:
            raise ValueError("x too large")
        return v
    assert bs(4.0, 6.0, 5.0, err_fn)[0] != 0                         # evaluation error
    assert bs(1.0, 5.0, 2.0, lambda v: 2 * v)[1:] == (1.0, 0)        # target at boundary
    err, x, ind = bs(0.0, 10.0, 9.0, lambda v: 2 * v + 3)            # utils_test.go:610-644: x* ~ 3
    assert err == 0 and abs(x - 3.0) <= 1e-3


def _test_system(arrival=0.0, ttft=100.0, itl=50.0, tps=0.0, min_rep=1, srv_max_batch=0):
    # setupCompleteTestSystem: pkg/core/allocation_test.go:11-80
    spec = {
        "acceleratorData": {"accelerators": [{"name": "test-gpu", "cost": 100.0}]},
        "modelData": {"models": [{"name": "test-model", "acc": "test-gpu", "accCount": 1, "maxBatchSize": 16,
                                  "atTokens": 200, "decodeParms": {"alpha": 5.0, "beta": 2.0},
                                  "prefillParms": {"gamma": 10.0, "delta": 1.5}}]},
        "serviceClassData": {"serviceClasses": [{"name": "default", "priority": 10, "modelTargets": [
            {"model": "test-model", "slo-itl": itl, "slo-ttft": ttft, "slo-tps": tps}]}]},
        "serverData": {"servers": [{"name": "test-server", "model": "test-model", "class": "default",
                                    "minNumReplicas": min_rep, "maxBatchSize": srv_max_batch,
                                    "currentAlloc": {"load": {"arrivalRate": arrival, "avgInTokens": 100,
                                                              "avgOutTokens": 200}}}]},
        "optimizerData": {"optimizer": {"unlimited": True}},
    }
    return Fleet.from_spec(spec)


def test_zero_load_allocation_exact(oracle_mod):  # allocation_test.go:82-140
    a = oracle_mod.create_allocation(_test_system(), 0, 0)
    assert (a["feasible"], a["acc"], a["replicas"], a["batch"]) == (1, 0, 1, 16)
    assert F(a["cost"]) == F(100.0) and F(a["value"]) == F(100.0)
    assert F(a["max_rate"]) == F(0.3298969)                       # MaxArrvRatePerReplica, exact ==
    assert F(a["max_rate"]) * F(1000) * F(60) == F(19793.814)     # MaxRPM, exact ==


def test_saturated_thresholds(oracle_mod):  # allocation_test.go:193-236
    a = oracle_mod.create_allocation(_test_system(), 0, 0)
    max_rpm = F(a["max_rate"]) * F(1000) * F(60)
    sat = lambda rate: F(rate) > F(a["replicas"]) * max_rpm  # noqa: E731  (allocation.go:254-256)
    assert [sat(r) for r in (15000.0, 19794.0, 25000.0, 0.0)] == [False, True, True, False]


def test_transition_penalty_kats(oracle_mod):  # allocation_test.go:238-287
    tp = oracle_mod.lib().wvao_transition_penalty
    assert tp(0.1, 0, 2, 100.0, 0, 2, 100.0) == 0.0
    assert tp(0.1, 0, 2, 100.0, 0, 3, 150.0) == 50.0
    assert F(tp(0.1, 0, 2, 100.0, 1, 2, 120.0)) == F(0.1) * F(220.0) + F(20.0)


def test_zero_load_allocation_cases(oracle_mod):  # allocation_test.go:971-1138
    f = _test_system(min_rep=0)
    a = oracle_mod.create_allocation(f, 0, 0)
    assert (a["feasible"], a["acc"], a["replicas"], a["batch"], a["cost"], a["value"]) == (1, -1, 0, 0, 0.0, 0.0)
    f = _test_system(min_rep=2)
    a = oracle_mod.create_allocation(f, 0, 0)
    assert (a["acc"], a["replicas"], a["batch"], a["cost"]) == (0, 2, 16, 200.0)
    assert F(a["itl"]) == F(5.0) + F(2.0) and F(a["ttft"]) == F(10.0) + F(1.5) and a["rho"] == 0.0
    assert F(a["max_rate"]) == F(16) / ((F(10.0) + F(1.5)) + (F(5.0) + F(2.0) * F(16)))
    f = _test_system(min_rep=1, srv_max_batch=8)
    f.acc_cost[0] = 50.0
    f.perf_acc_count[0, 0] = 2
    f.perf_alpha[0, 0], f.perf_beta[0, 0], f.perf_gamma[0, 0], f.perf_delta[0, 0] = 3.0, 1.0, 8.0, 2.0
    a = oracle_mod.create_allocation(f, 0, 0)
    assert (a["acc"], a["replicas"], a["batch"], a["cost"]) == (0, 1, 8, 100.0)
    assert F(a["max_rate"]) == F(8) / ((F(8.0) + F(2.0)) + (F(3.0) + F(1.0) * F(8)))


def test_create_allocation_nil_and_feasible_cases(oracle_mod):  # allocation_test.go:579-776
    ca = oracle_mod.create_allocation
    f = _test_system()
    assert ca(f, 0, 5)["feasible"] == 0 and ca(f, 3, 0)["feasible"] == 0      # nonexistent acc / server
    g = _test_system()
    g.perf_present[:] = 0
    assert ca(g, 0, 0)["feasible"] == 0                                        # no performance data
    g = _test_system()
    g.srv_has_target[:] = 0
    assert ca(g, 0, 0)["feasible"] == 0                                        # no service class target
    assert ca(_test_system(arrival=1200, ttft=1.0, itl=0.1), 0, 0)["feasible"] == 0   # strict SLOs -> nil
    a = ca(_test_system(arrival=60, ttft=2000.0, itl=500.0, tps=2.0), 0, 0)   # TPS branch
    assert a["feasible"] == 1 and a["acc"] == 0 and a["replicas"] > 0
    a = ca(_test_system(arrival=120, ttft=2000.0, itl=500.0), 0, 0)           # arrival-rate branch
    assert a["feasible"] == 1 and a["replicas"] > 0 and a["batch"] == 16
    a = ca(_test_system(arrival=60, ttft=2000.0, itl=500.0, srv_max_batch=12), 0, 0)
    assert a["feasible"] == 1 and a["batch"] == 12                            # batch override


def test_scale_up_increases_replicas(oracle_mod):  # allocation_test.go:778-887
    lo = oracle_mod.create_allocation(_test_system(arrival=30, ttft=2000.0, itl=500.0), 0, 0)
    hi = oracle_mod.create_allocation(_test_system(arrival=360, ttft=2000.0, itl=500.0), 0, 0)
    assert lo["feasible"] and hi["feasible"] and hi["replicas"] - lo["replicas"] > 0


def _a100_fixture(arrival):
    # internal/optimizer/optimizer_test.go:194-201,402-416 with test/utils/unitutils.go:64-115
    spec = {
        "acceleratorData": {"accelerators": [{"name": "A100", "type": "NVIDIA-A100-PCIE-80GB", "multiplicity": 1,
                                              "cost": 40.0}]},
        "modelData": {"models": [{"name": "default/default", "acc": "A100", "accCount": 1, "maxBatchSize": 4,
                                  "decodeParms": {"alpha": 20.28, "beta": 0.72},
                                  "prefillParms": {"gamma": 0.0, "delta": 0.0}}]},
        "serviceClassData": {"serviceClasses": [{"name": "Premium", "priority": 1, "modelTargets": [
            {"model": "default/default", "slo-itl": 80.0, "slo-ttft": 500.0}]}]},
        "serverData": {"servers": [{"name": "va:ns", "model": "default/default", "class": "Premium",
                                    "keepAccelerator": True, "minNumReplicas": 1, "maxBatchSize": 4,
                                    "currentAlloc": {"accelerator": "A100", "numReplicas": 1, "cost": 40.0,
                                                     "load": {"arrivalRate": arrival, "avgInTokens": 20,
                                                              "avgOutTokens": 200}}}]},
        "optimizerData": {"optimizer": {"unlimited": True}},
    }
    return Fleet.from_spec(spec)


def test_integration_fixture_replica_decisions(oracle_mod):  # optimizer_test.go:333,455
    _, win = oracle_mod.solve(_a100_fixture(0.0))
    assert win["feasible"][0] == 1 and win["replicas"][0] == 1
    _, win = oracle_mod.solve(_a100_fixture(1200.0))
    assert win["feasible"][0] == 1 and win["replicas"][0] > 1 and win["acc"][0] == 0


def test_mm1k_validity_table(oracle_mod):  # pkg/analyzer/queuemodel_test.go:9-104
    m = oracle_mod.MM1K(10)
    for lam, mu, valid in ((1.0, 2.0, True), (0.0, 2.0, True), (-1.0, 2.0, False), (1.0, 0.0, False),
                           (1.0, -1.0, False), (9.9, 1.0, True), (11.0, 1.0, False)):
        assert bool(m.solve(lam, mu)["is_valid"]) == valid
    st = m.solve(1.0, 2.0)
    assert st["avg_resp_time"] > 0 and st["avg_serv_time"] > 0 and st["avg_wait_time"] >= 0
    p = oracle_mod.MM1K(3)
    for lam, mu in ((0.5, 2.0), (1.5, 2.0), (1.9, 2.0), (2.0, 2.0)):
        p.solve(lam, mu)
        pr = p.probs()
        assert (pr >= 0).all() and abs(pr.sum() - 1.0) < 1e-6


def test_state_dependent_model_properties(oracle_mod):  # queuemodel_test.go:498-533 (Little's law)
    qa = oracle_mod.Analyzer(8, 80, 1.0, 0.01, 10.0, 0.001, 100, 10)
    rmin, rmax = qa.rate_range()
    for frac in (0.1, 0.5, 0.9):
        lam = (rmin + frac * (rmax - rmin)) / 1000.0
        st = qa.solve(lam)
        assert st["is_valid"]
        p = qa.probs()
        assert (p >= 0).all() and abs(p.sum() - 1.0) < 1e-6
        L = float((np.arange(p.size) * p).sum())
        assert abs(L - st["throughput"] * st["avg_resp_time"]) <= 1e-4 * max(1.0, L)   # Little's law
        assert 0.0 <= st["rho"] <= 1.0


def test_stale_rho_makes_k1_models_invalid(oracle_mod):  # queuemodel.go:31 + mm1modelstatedependent.go:33-35
    qa = oracle_mod.Analyzer(1, 0, 1.0, 0.01, 10.0, 0.001, 100, 10)   # K = 1
    assert not qa.solve(1e-4)["is_valid"]


def test_ttft_itl_monotone_in_rate(oracle_mod):
    qa = oracle_mod.Analyzer(16, 160, 20.58, 0.41, 5.2, 0.1, 128, 128)
    rmin, rmax = qa.rate_range()
    xs = np.linspace(rmin / 1000, rmax / 1000, 25).astype(np.float32)
    ttft = [qa.eval_ttft(float(x))[1] for x in xs]
    itl = [qa.eval_itl(float(x))[1] for x in xs]
    assert all(b >= a - 1e-4 for a, b in zip(ttft, ttft[1:]))
    assert all(b >= a - 1e-4 for a, b in zip(itl, itl[1:]))


# ---- pkg/analyzer/queueanalyzer_test.go behaviour tests, through the oracle ----------------------------
_TEST_CONFIG = dict(max_batch=8, max_queue=16, alpha=1.0, beta=0.01, gamma=10.0, delta=0.001)   # queueanalyzer_test.go:11-24


def _qa(oracle_mod, in_tokens, out_tokens):
    c = _TEST_CONFIG
    return oracle_mod.Analyzer(c["max_batch"], c["max_queue"], c["alpha"], c["beta"], c["gamma"], c["delta"], in_tokens, out_tokens)


def test_new_queue_analyzer_request_sizes(oracle_mod):  # queueanalyzer_test.go:26-90
    for in_tok, out_tok in ((0, 10), (0, 1), (100, 1), (200, 20)):
        assert _qa(oracle_mod, in_tok, out_tok) is not None
    for in_tok, out_tok in ((0, 0), (-1, -1), (50, 0)):
        with pytest.raises(ValueError):
            _qa(oracle_mod, in_tok, out_tok)


def test_build_model_rate_range(oracle_mod):  # queueanalyzer_test.go:317-355
    qa = _qa(oracle_mod, 100, 10)
    rmin, rmax = qa.rate_range()
    assert 0 < rmin < rmax
    assert qa.K == _TEST_CONFIG["max_batch"] + _TEST_CONFIG["max_queue"]
    sr = qa.serv_rate()
    assert sr.shape == (8,) and (np.diff(sr) > 0).all()          # more concurrency, more throughput (beta << alpha)


def test_analyze_rate_cases(oracle_mod):  # queueanalyzer_test.go:357-446
    qa = _qa(oracle_mod, 100, 10)
    rmin, rmax = qa.rate_range()
    f32 = np.float32
    for rate in (0.0, -1.0, float(f32(rmax) * f32(1.1))):
        err, m = qa.analyze(rate)
        assert err != 0 and m is None
    for rate in (float(f32(rmin) * f32(0.5)), float((f32(rmin) + f32(rmax)) * f32(0.5)), float(f32(rmax) * f32(0.9))):
        err, m = qa.analyze(rate)
        assert err == 0
        for k in ("throughput", "avg_resp_time", "avg_wait_time", "avg_num_in_serv", "avg_prefill_time", "avg_token_time"):
            assert m[k] >= 0, k
        assert 0 <= m["rho"] <= 1


def test_size_target_cases(oracle_mod):  # queueanalyzer_test.go:448-554
    qa = _qa(oracle_mod, 100, 10)
    for ttft, itl, tps in ((50.0, 5.0, 100.0), (0.0, 0.0, 0.0)):
        err, rates, metrics, achieved = qa.size(ttft, itl, tps)
        assert err == 0 and metrics is not None
        assert all(v >= 0 for v in rates.values()) and all(v >= 0 for v in achieved.values())
    for ttft, itl, tps in ((-1.0, 5.0, 100.0), (50.0, -1.0, 100.0), (50.0, 5.0, -1.0)):
        assert qa.size(ttft, itl, tps)[0] != 0


def test_eval_ttft_and_itl_cases(oracle_mod):  # pkg/analyzer/utils_test.go:383-519 (N = 4, queue 8)
    qa = oracle_mod.Analyzer(4, 8, 1.0, 0.01, 10.0, 0.001, 100, 10)
    for lam in (0.001, 0.01, 1.0):
        err, ttft = qa.eval_ttft(lam)
        assert err == 0 and ttft >= 10.0            # at least the base prefill time gamma
        err, itl = qa.eval_itl(lam)
        assert err == 0 and itl >= 1.0              # at least alpha


def test_binary_search_with_analyzer_functions(oracle_mod):  # utils_test.go:521-608
    qa = oracle_mod.Analyzer(4, 8, 1.0, 0.01, 10.0, 0.001, 100, 10)
    rmin, rmax = qa.rate_range()
    lmin, lmax = float(np.float32(rmin) / np.float32(1000)), float(np.float32(rmax) / np.float32(1000))

    def ev(fn):
        def f(x):
            err, y = fn(x)
            if err:
                raise ValueError("invalid model")
            return y
        return f
    for target, fn in ((25.0, qa.eval_ttft), (2.0, qa.eval_itl)):
        err, xs, ind = oracle_mod.binary_search(lmin, lmax, target, ev(fn))
        if err == 0 and ind == 0:                    # found inside the bracket: the target is met to 0.1
            e2, y = fn(xs)
            assert e2 == 0 and abs(y - target) <= 0.1


# ---- pkg/analyzer/queuemodel_test.go: MM1ModelStateDependent built from raw rate vectors --------------
def test_state_dependent_solve_cases(oracle_mod):  # queuemodel_test.go:325-400
    m = oracle_mod.StateDependentModel(5, [1.0, 2.0, 3.0])
    for lam, want_valid in ((0.5, True), (1.5, True), (2.8, True), (0.0, True), (-1.0, False)):
        st = m.solve(lam, 1.0)
        assert bool(st["is_valid"]) == want_valid, lam
        if want_valid:
            assert st["avg_num_in_servers"] >= 0 and 0 <= st["rho"] <= 1
            if st["avg_resp_time"] > 0 and st["throughput"] > 0:      # Little's law to 1e-4
                assert abs(np.float32(st["throughput"]) * np.float32(st["avg_resp_time"]) - st["avg_num_in_system"]) <= 1e-4
    # hand check at lambda = 0.5: p = (1, 1/2, 1/8, 1/48, 1/288, 1/1728) / sum
    m.solve(0.5, 1.0)
    w = np.array([1, 1 / 2, 1 / 8, 1 / 48, 1 / 288, 1 / 1728])
    assert np.allclose(m.probs(), w / w.sum(), rtol=1e-15)


def test_state_dependent_utilisation_is_one_minus_p0(oracle_mod):  # queuemodel_test.go:402-422
    m = oracle_mod.StateDependentModel(4, [2.0, 4.0, 6.0])
    st = m.solve(1.0, 1.0)
    assert st["is_valid"] and abs(st["rho"] - np.float32(1.0 - np.float32(m.probs()[0]))) <= 1e-6


def test_state_dependent_service_rate_extension(oracle_mod):  # queuemodel_test.go:424-441: K beyond the rate vector
    m = oracle_mod.StateDependentModel(5, [1.0, 2.0])
    st = m.solve(0.5, 1.0)
    assert st["is_valid"] and st["avg_num_in_system"] >= 0 and st["throughput"] >= 0
    w = np.array([1, 1 / 2, 1 / 8, 1 / 32, 1 / 128, 1 / 512])          # states 2.. are served at the last rate
    assert np.allclose(m.probs(), w / w.sum(), rtol=1e-15)


def test_candidate_accelerators_keep_rule(oracle_mod):  # pkg/core/server_test.go:395-466 (server.go:70-82)
    from workload_variant_autoscaler_b200 import Fleet

    def candidates(keep, cur_acc):
        spec = {
            "acceleratorData": {"accelerators": [{"name": n, "type": n, "multiplicity": 1, "cost": c}
                                                 for n, c in (("gpu-a", 100.0), ("gpu-b", 150.0), ("gpu-c", 80.0))]},
            "modelData": {"models": [{"name": "test-model", "acc": n, "accCount": 1, "maxBatchSize": 8,
                                      "decodeParms": {"alpha": 5.0, "beta": 2.0}, "prefillParms": {"gamma": 10.0, "delta": 1.5}}
                                     for n in ("gpu-a", "gpu-b", "gpu-c")]},
            "serviceClassData": {"serviceClasses": [{"name": "default", "priority": 1, "modelTargets": [
                {"model": "test-model", "slo-itl": 100, "slo-ttft": 1000}]}]},
            "serverData": {"servers": [{"name": "test-server", "model": "test-model", "class": "default",
                                        "keepAccelerator": keep, "minNumReplicas": 1,
                                        "currentAlloc": {"accelerator": cur_acc, "load": {}}}]},
            "optimizerData": {"optimizer": {"unlimited": True}}, "capacityData": {"count": []},
        }
        fleet = Fleet.from_spec(spec)
        cand = oracle_mod.calculate(fleet)[0]                       # zero load: one zero-load allocation per candidate
        return [fleet.acc_names[a] for a in range(fleet.n_acc) if cand["feasible"][a]]

    assert candidates(False, "") == ["gpu-a", "gpu-b", "gpu-c"]            # no keep-accelerator constraint
    assert candidates(True, "") == ["gpu-a", "gpu-b", "gpu-c"]             # keep, but no current accelerator
    assert candidates(True, "gpu-b") == ["gpu-b"]                           # keep the current accelerator
    assert candidates(True, "nonexistent-gpu") == []                        # current accelerator not in the system
