"""Second, independent restatement of the reference hot path in numpy scalars (TEST ONLY).

Purpose: cross-check the C oracle (oracle/wva_oracle.c) on small cases, because beyond
the reference's own KATs nothing pins float outputs under load.  This
file is written against the Go sources directly and shares no code with the C oracle:
float32 arithmetic is done with ``np.float32`` scalars (each operation rounds to binary32,
no fused multiply-add), the probability vector is a ``float64`` numpy array and the
recurrence is a plain Python loop, exactly as mm1modelstatedependent.go:70-116 reads.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
D = np.float64
EPS = F(0.001)                 # queueanalyzer.go:8
SSF = F(0.1)                   # queueanalyzer.go:11
TOL = F(1e-6)                  # utils.go:8
MAX_IT = 100                   # utils.go:9


def gmin(a, b):
    if np.isnan(a) or np.isnan(b):
        return F(np.nan)
    if a == 0 and b == 0:
        return a if np.signbit(a) else b
    return a if a < b else b


def gmax(a, b):
    if np.isnan(a) or np.isnan(b):
        return F(np.nan)
    if a == 0 and b == 0:
        return b if np.signbit(a) else a
    return a if a > b else b


class QA:
    """QueueAnalyzer + MM1ModelStateDependent (queueanalyzer.go:14-131)."""

    def __init__(self, N, Q, alpha, beta, gamma, delta, in_tok, out_tok):
        self.N, self.Q = N, Q
        self.a, self.b, self.g, self.d = F(alpha), F(beta), F(gamma), F(delta)
        self.it, self.ot = int(in_tok), int(out_tok)
        self.sr = np.zeros(N, np.float32)
        with np.errstate(all="ignore"):
            for n in range(1, N + 1):
                pre = self.prefill(F(n))
                nd = self.ot - 1
                if self.it == 0 and self.ot == 1:
                    nd = 1
                dec = F(nd) * self.decode(F(n))
                self.sr[n - 1] = F(n) / (pre + dec)
            self.rmin = (self.sr[0] * EPS) * F(1000)
            self.rmax = (self.sr[N - 1] * (F(1) - EPS)) * F(1000)
        self.K = N + Q
        self.p = np.zeros(self.K + 1, np.float64)
        self.valid = False

    def prefill(self, batch):  # :257-262
        if self.it == 0:
            return F(0)
        return self.g + (self.d * F(self.it)) * batch

    def decode(self, batch):  # :264-266
        return self.a + self.b * batch

    def effconc(self, serv):  # :296-302
        with np.errstate(all="ignore"):
            tok = F(self.ot - 1)
            num = serv - (self.g + self.a * tok)
            den = (self.d * F(self.it)) + (self.b * tok)
            n = num / den
        return gmin(gmax(n, F(0)), F(self.N))

    def solve(self, lam, mu=F(1)):  # queuemodel.go:27-37
        lam, mu = F(lam), F(mu)
        rho = F(1) - F(self.p[0])
        if rho < 0 or rho >= F(self.K) or lam < 0 or mu <= 0:
            self.valid = False
            return
        self.valid = True
        K, N, p = self.K, self.N, self.p
        with np.errstate(all="ignore"):
            p[0] = 1.0
            scale = D(np.finfo(np.float64).max) / D(K)
            for n in range(K):  # mm1modelstatedependent.go:76-90
                s = D(self.sr[n]) if n < N else D(self.sr[N - 1])
                p[n + 1] = (p[n] * D(lam)) / s
                while p[n + 1] < 0 or math.isinf(p[n + 1]) or math.isnan(p[n + 1]):
                    p[: n + 1] /= scale
                    p[n + 1] = (p[n] * D(lam)) / s
            tot = D(0)
            for n in range(K + 1):  # :93-105
                tot = tot + p[n]
                if tot < 0 or math.isinf(tot):
                    tot = D(0)
                    for i in range(K + 1):
                        p[i] /= scale
                        if i <= n:
                            tot = tot + p[i]
            for n in range(K + 1):  # :108-112
                p[n] = p[n] / tot
            in_sys = D(0)
            in_srv = D(0)
            sp = p[0]
            for i in range(1, K + 1):  # :47-55
                in_sys = in_sys + D(i) * p[i]
                sp = sp + p[i]
                if i == N:
                    in_srv = in_sys + (D(1) - sp) * D(N)
            self.num_serv = F(in_srv)
            num_sys = F(in_sys)
            self.tput = lam * (F(1) - F(p[K]))
            self.resp = num_sys / self.tput
            self.serv = self.num_serv / self.tput
            w = self.resp - self.serv
            self.wait = F(0) if w < 0 else w

    def analyze(self, rate):  # queueanalyzer.go:134-174
        rate = F(rate)
        if rate <= 0 or rate > self.rmax:
            return None
        with np.errstate(all="ignore"):
            self.solve(rate / F(1000))
            if not self.valid:
                return None
            e = self.effconc(self.serv)
            rho = gmin(gmax(self.num_serv / F(self.N), F(0)), F(1))
            return {"throughput": self.tput * F(1000), "wait": self.wait, "prefill": self.prefill(e),
                    "itl": self.decode(e), "rho": rho}

    def eval(self, which, x):  # :270-290
        self.solve(x)
        if not self.valid:
            raise ValueError
        e = self.effconc(self.serv)
        with np.errstate(all="ignore"):
            return self.wait + self.prefill(e) if which == "ttft" else self.decode(e)


def within(x, v, tol):  # utils.go:12-20
    if x == v:
        return True
    if v == 0 or tol < 0:
        return False
    with np.errstate(all="ignore"):
        return abs(D((x - v) / v)) <= D(tol)


def bsearch(xmin, xmax, target, fn):  # utils.go:26-70 -> (x, ind)
    xmin, xmax, target = F(xmin), F(xmax), F(target)
    if xmin > xmax:
        raise ValueError
    ys = []
    for x in (xmin, xmax):
        y = fn(x)
        ys.append(y)
        if within(y, target, TOL):
            return x, 0
    inc = ys[0] < ys[1]
    if (inc and target < ys[0]) or (not inc and target > ys[0]):
        return xmin, -1
    if (inc and target > ys[1]) or (not inc and target < ys[1]):
        return xmax, 1
    xs = F(0)
    for _ in range(MAX_IT):
        xs = F(0.5) * (xmin + xmax)
        y = fn(xs)
        if within(y, target, TOL):
            break
        if (inc and target < y) or (not inc and target > y):
            xmax = xs
        else:
            xmin = xs
    return xs, 0


def size(qa: QA, ttft, itl, tps):  # queueanalyzer.go:185-255 -> metrics dict or None
    ttft, itl, tps = F(ttft), F(itl), F(tps)
    if ttft < 0 or itl < 0 or tps < 0:
        return None
    lmin, lmax = qa.rmin / F(1000), qa.rmax / F(1000)
    try:
        l1 = lmax
        if ttft > 0:
            l1, ind = bsearch(lmin, lmax, ttft, lambda x: qa.eval("ttft", x))
            if ind < 0:
                return None
        l2 = lmax
        if itl > 0:
            l2, ind = bsearch(lmin, lmax, itl, lambda x: qa.eval("itl", x))
            if ind < 0:
                return None
    except ValueError:
        return None
    l3 = lmax * (F(1) - SSF) if tps > 0 else lmax
    lam = gmin(gmin(l1, l2), l3)
    return qa.analyze(lam * F(1000))


def create_allocation(fleet, s, a):
    """allocation.go:27-163 on a Fleet -> dict (feasible, acc, replicas, batch, cost, itl, ttft, rho, max_rate)."""
    nil = {"feasible": 0}
    f = fleet
    if f.srv_arrival_rpm[s] < 0 or f.srv_in_tokens[s] < 0 or f.srv_out_tokens[s] < 0:
        return nil
    m = int(f.srv_model[s])
    if m < 0 or not f.perf_present[m, a] or not f.srv_has_target[s]:
        return nil
    ninst = int(f.perf_acc_count[m, a]) if f.perf_acc_count[m, a] > 0 else 1
    if f.srv_arrival_rpm[s] == 0 or f.srv_out_tokens[s] == 0:  # zeroLoadAllocation :259-288
        nrep = int(f.srv_min_replicas[s])
        if nrep == 0:
            return {"feasible": 1, "acc": -1, "replicas": 0, "batch": 0, "cost": F(0), "itl": F(0), "ttft": F(0),
                    "rho": F(0), "max_rate": F(0)}
        mb = int(f.srv_max_batch[s]) if f.srv_max_batch[s] > 0 else int(f.perf_max_batch[m, a])
        al, be, ga, de = (F(f.perf_alpha[m, a]), F(f.perf_beta[m, a]), F(f.perf_gamma[m, a]), F(f.perf_delta[m, a]))
        pre = ga + de
        return {"feasible": 1, "acc": a, "replicas": nrep, "batch": mb, "cost": F(f.acc_cost[a]) * F(ninst * nrep),
                "itl": al + be, "ttft": pre, "rho": F(0), "max_rate": F(mb) / (pre + (al + be * F(mb)))}
    K = int(f.srv_out_tokens[s])
    N = int(f.srv_max_batch[s]) if f.srv_max_batch[s] > 0 else max(
        int(f.perf_max_batch[m, a]) * int(f.perf_at_tokens[m, a]) // K, 1)
    if K < 1:
        return nil
    qa = QA(N, N * f.max_queue_to_batch_ratio, f.perf_alpha[m, a], f.perf_beta[m, a], f.perf_gamma[m, a],
            f.perf_delta[m, a], f.srv_in_tokens[s], K)
    met = size(qa, f.srv_slo_ttft[s], f.srv_slo_itl[s], f.srv_slo_tps[s])
    if met is None:
        return nil
    rate_star = met["throughput"]
    with np.errstate(all="ignore"):
        total = F(f.srv_arrival_rpm[s]) / F(60) if f.srv_slo_tps[s] == 0 else F(f.srv_slo_tps[s]) / F(K)
        q = D(total) / D(rate_star)
        nrep = int(math.ceil(q)) if math.isfinite(q) else -(1 << 63)
        nrep = max(nrep, int(f.srv_min_replicas[s]))
        cost = F(f.acc_cost[a]) * F(ninst * nrep)
        met2 = qa.analyze(total / F(nrep)) if nrep != 0 else None
    if met2 is None:
        return nil
    return {"feasible": 1, "acc": a, "replicas": nrep, "batch": N, "cost": cost, "itl": met2["itl"],
            "ttft": met2["wait"] + met2["prefill"], "rho": met2["rho"], "max_rate": rate_star / F(1000)}
