"""Host logic: packing config.SystemSpec (JSON shape) into the flat SoA fleet, sharding, winner (un)packing."""
import numpy as np

from workload_variant_autoscaler_b200 import ACC_NONE, ACC_UNKNOWN, Allocs, Fleet, parallel, synth_fleet


def _greedy_spec():
    # pkg/solver/greedy_test.go:13-208 (setupTestSystemForGreedy), condensed
    return {
        "acceleratorData": {"accelerators": [
            {"name": "A100", "type": "GPU_A100", "multiplicity": 1, "cost": 1.0},
            {"name": "H100", "type": "GPU_H100", "multiplicity": 1, "cost": 2.0}]},
        "modelData": {"models": [
            {"name": "llama-7b", "acc": "A100", "accCount": 1, "maxBatchSize": 16, "atTokens": 100,
             "decodeParms": {"alpha": 10.0, "beta": 2.0}, "prefillParms": {"gamma": 5.0, "delta": 0.1}},
            {"name": "llama-7b", "acc": "H100", "accCount": 1, "maxBatchSize": 32, "atTokens": 100,
             "decodeParms": {"alpha": 8.0, "beta": 1.5}, "prefillParms": {"gamma": 3.0, "delta": 0.08}},
            {"name": "llama-13b", "acc": "A100", "accCount": 2, "maxBatchSize": 8, "atTokens": 150,
             "decodeParms": {"alpha": 15.0, "beta": 3.0}, "prefillParms": {"gamma": 8.0, "delta": 0.15}},
            {"name": "llama-13b", "acc": "H100", "accCount": 1, "maxBatchSize": 16, "atTokens": 150,
             "decodeParms": {"alpha": 12.0, "beta": 2.5}, "prefillParms": {"gamma": 6.0, "delta": 0.12}}]},
        "serviceClassData": {"serviceClasses": [
            {"name": "high-priority", "priority": 1, "modelTargets": [
                {"model": "llama-7b", "slo-itl": 400, "slo-ttft": 2000, "slo-tps": 0}]},
            {"name": "low-priority", "priority": 500, "modelTargets": [
                {"model": "llama-13b", "slo-itl": 800, "slo-ttft": 4000, "slo-tps": 0}]}]},
        "serverData": {"servers": [
            {"name": "s1", "class": "high-priority", "model": "llama-7b", "minNumReplicas": 1,
             "currentAlloc": {"accelerator": "A100", "numReplicas": 1, "cost": 1.0,
                              "load": {"arrivalRate": 60, "avgInTokens": 100, "avgOutTokens": 200}}},
            {"name": "s2", "class": "low-priority", "model": "llama-13b", "keepAccelerator": True,
             "currentAlloc": {"accelerator": "Gaudi", "numReplicas": 2, "cost": 5.0,
                              "load": {"arrivalRate": 30, "avgInTokens": 150, "avgOutTokens": 300}}},
            {"name": "s3", "class": "", "model": "unknown-model",
             "currentAlloc": {"load": {"arrivalRate": 10, "avgInTokens": 10, "avgOutTokens": 10}}}]},
        "optimizerData": {"optimizer": {"unlimited": False, "saturationPolicy": "PriorityRoundRobin"}},
        "capacityData": {"count": [{"type": "GPU_A100", "count": 4}, {"type": "GPU_H100", "count": 2}]},
    }


def test_from_spec_joins():
    f = Fleet.from_spec(_greedy_spec())
    assert (f.n_acc, f.n_models, f.n_servers, f.n_types) == (2, 2, 3, 2)
    assert f.acc_names == ["A100", "H100"] and list(f.type_capacity) == [4, 2]
    assert f.perf_present.all() and f.perf_acc_count[1, 0] == 2 and f.perf_max_batch[0, 1] == 32
    assert list(f.srv_model) == [0, 1, -1]
    # priority outside [1, 100] falls back to the default (serviceclass.go:28-31); unknown class -> 100, no target
    assert list(f.srv_priority) == [1, 100, 100] and list(f.srv_has_target) == [1, 1, 0]
    assert list(f.srv_cur_acc) == [0, ACC_UNKNOWN, ACC_NONE]
    assert not f.unlimited and f.saturation_policy == 2
    assert np.float32(f.srv_slo_itl[0]) == np.float32(400) and f.srv_keep_acc[1] == 1


def test_from_spec_feeds_oracle(oracle_mod):
    f = Fleet.from_spec(_greedy_spec())
    cand = oracle_mod.calculate(f)
    assert cand["feasible"][0].all()            # s1: both accelerators feasible
    assert not cand["feasible"][1].any()        # s2: keepAccelerator with an unknown current accelerator
    assert not cand["feasible"][2].any()        # s3: unknown model


def test_shard_and_gather_roundtrip():
    f = synth_fleet(11, 3, seed=3)
    world = 4
    seen = []
    blocks = []
    for r in range(world):
        sh = f.shard(r, world)
        idx = parallel.shard_indices(f.n_servers, r, world)
        assert np.array_equal(sh.srv_arrival_rpm, f.srv_arrival_rpm[idx])
        assert np.array_equal(sh.perf_alpha, f.perf_alpha)        # tables are replicated
        seen.extend(idx.tolist())
        w = Allocs(sh.n_servers)
        w.feasible[:] = 1
        w.replicas[:] = idx + 1
        w.cost[:] = (idx * 0.5).astype(np.float32)
        blocks.append(parallel.pack_winners(w, (f.n_servers + world - 1) // world))
    assert sorted(seen) == list(range(f.n_servers))
    out = parallel.unpack_winners(np.stack(blocks), f.n_servers, world)
    assert np.array_equal(out.replicas, np.arange(f.n_servers) + 1)
    assert np.array_equal(out.cost, (np.arange(f.n_servers) * 0.5).astype(np.float32))
    assert out.feasible.all()


def test_as_c_cache_follows_the_arrays():
    """Fleet.as_c caches its ctypes struct: in-place edits are seen through the same pointers, a re-attached
    array gets a fresh pointer, and callers receive a private copy they may edit."""
    import ctypes as C
    f = synth_fleet(4, 2, seed=9)
    c1 = f.as_c()
    p1 = C.cast(c1.srv_arrival_rpm, C.c_void_p).value
    f.srv_arrival_rpm[:] = 7.0                                   # in place: same buffer
    c2 = f.as_c()
    assert C.cast(c2.srv_arrival_rpm, C.c_void_p).value == p1 and c2.srv_arrival_rpm[0] == 7.0
    c2.srv_model = None                                          # editing the returned struct ...
    assert bool(f.as_c().srv_model)                              # ... does not poison the cache
    f.srv_arrival_rpm = np.full(4, 3.0, np.float32)              # re-attached: new buffer
    c3 = f.as_c()
    assert C.cast(c3.srv_arrival_rpm, C.c_void_p).value == f.srv_arrival_rpm.ctypes.data
    assert c3.srv_arrival_rpm[0] == 3.0 and c3.n_servers == 4


def test_packer_fixture_repeated_names_last_wins():
    """tests/golden/packer_fixture.json is the fixture the two packers share (Fleet.from_spec here, pack() in
    go/wvab200/wvab200.go): repeated accelerator / server names follow System.SetFromSpec (pkg/core/system.go:99-101,
    151-153: the map entry is replaced), i.e. the last spec wins at the position of the first."""
    import json
    import os
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "packer_fixture.json")))
    f = Fleet.from_spec(fx["spec"])
    # hand-derived from the spec
    assert f.acc_names == ["A100", "H100"] and f.server_names == ["a:ns", "b:ns"]
    assert list(f.acc_cost) == [np.float32(41.5), np.float32(65.0)] and list(f.acc_multiplicity) == [4, 2]
    assert f.type_names == ["GPU_A100_80", "GPU_H100", "TPU"] and list(f.type_capacity) == [0, 7, 3]
    assert list(f.srv_min_replicas) == [2, 0] and list(f.srv_max_batch) == [4, 0]
    assert list(f.srv_cur_acc) == [1, ACC_NONE] and list(f.srv_cur_replicas) == [3, 0]
    assert list(f.srv_arrival_rpm) == [np.float32(600.0), np.float32(0.0)]
    # and the whole SoA as frozen in the fixture (what the Go packer must reproduce)
    for name, want in fx["soa"].items():
        got = getattr(f, name)
        got = got.reshape(-1).tolist() if isinstance(got, np.ndarray) else got
        assert got == want, name
