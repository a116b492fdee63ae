"""C oracle vs the independent numpy restatement (oracle/restate_np.py): bit-for-bit on small
random cases.  Float outputs under load are pinned by no reference test, so two
independently written restatements agreeing is what stands in for the Go binary here."""
import numpy as np
import pytest

from workload_variant_autoscaler_b200 import synth_fleet

F = np.float32


def _same(a, b):
    a, b = F(a), F(b)
    return (np.isnan(a) and np.isnan(b)) or a.view(np.uint32) == b.view(np.uint32)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_create_allocation_cross(oracle_mod, seed):
    from oracle import restate_np as R
    f = synth_fleet(10, 3, seed=seed, max_batch_choices=(1, 2, 4, 8, 16), zero_load_frac=0.2)
    f.srv_min_replicas[::3] = 0
    n_feas = 0
    for s in range(f.n_servers):
        for a in range(f.n_acc):
            o = oracle_mod.create_allocation(f, s, a)
            r = R.create_allocation(f, s, a)
            assert o["feasible"] == r["feasible"], (s, a, o, r)
            if not o["feasible"]:
                continue
            n_feas += 1
            assert (o["acc"], o["replicas"], o["batch"]) == (r["acc"], r["replicas"], r["batch"]), (s, a, o, r)
            for k in ("cost", "itl", "ttft", "rho", "max_rate"):
                assert _same(o[k], r[k]), (s, a, k, o[k], r[k])
    assert n_feas >= 15


def test_analyze_and_solve_cross(oracle_mod):
    from oracle import restate_np as R
    rng = np.random.default_rng(7)
    for _ in range(12):
        N = int(rng.integers(1, 24))
        al, be = F(rng.uniform(5, 25)), F(rng.uniform(0.02, 0.8))
        ga, de = F(rng.uniform(0, 250)), F(np.exp(rng.uniform(np.log(1e-4), np.log(0.1))))
        it, ot = int(rng.integers(0, 2048)), int(rng.integers(1, 512))
        qa = oracle_mod.Analyzer(N, 10 * N, al, be, ga, de, it, ot)
        rq = R.QA(N, 10 * N, al, be, ga, de, it, ot)
        assert np.array_equal(qa.serv_rate().view(np.uint32), rq.sr.view(np.uint32))
        rmin, rmax = qa.rate_range()
        assert _same(rmin, rq.rmin) and _same(rmax, rq.rmax)
        for frac in (0.0, 0.3, 0.7, 1.0):
            rate = F(rmin + F(frac) * (rmax - rmin))
            err, m = qa.analyze(float(rate))
            mr = rq.analyze(rate)
            assert (err == 0) == (mr is not None)
            if mr is None:
                continue
            assert _same(m["throughput"], mr["throughput"]) and _same(m["avg_wait_time"], mr["wait"])
            assert _same(m["avg_prefill_time"], mr["prefill"]) and _same(m["avg_token_time"], mr["itl"])
            assert _same(m["rho"], mr["rho"])
            assert np.array_equal(qa.probs().view(np.uint64), rq.p.view(np.uint64))


def test_overflow_rescale_cross(oracle_mod):
    """N large enough that prod(lambda/servRate) overflows: both restatements must take the
    reference's rescale branches (mm1modelstatedependent.go:84-89) and agree."""
    from oracle import restate_np as R
    args = (800, 8000, 7.47, 0.0001, 1.0, 0.00001, 64, 64)
    qa, rq = oracle_mod.Analyzer(*args), R.QA(*args)
    rmin, rmax = qa.rate_range()
    rate = F(rmax) * F(0.98)
    err, m = qa.analyze(float(rate))
    mr = rq.analyze(rate)
    assert err == 0 and mr is not None
    assert np.array_equal(qa.probs().view(np.uint64), rq.p.view(np.uint64))
    assert _same(m["avg_wait_time"], mr["wait"]) and _same(m["rho"], mr["rho"])
