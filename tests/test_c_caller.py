"""A compiled C program (tests/c_abi/solve_smoke.c, gcc, linked against libwva_b200.so) drives the C ABI the way
the cgo binding would: no Python, no ctypes between the caller and the library.  CPU: it builds, links and
reports the missing device with its documented exit code.  GPU: its output equals the oracle bit for bit."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi", "solve_smoke.c")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    from workload_variant_autoscaler_b200 import _lib
    _lib.build()
    libdir = os.path.dirname(_lib.SO_PATH)
    out = str(tmp_path_factory.mktemp("c_abi") / "solve_smoke")
    subprocess.run(["gcc", "-O1", "-Wall", "-Wextra", "-Werror", "-std=c11", "-I", os.path.join(ROOT, "include"), SRC,
                    "-o", out, "-L", libdir, "-l:libwva_b200.so", f"-Wl,-rpath,{libdir}"], check=True)
    return out


def test_c_caller_builds_links_and_fails_loudly_without_gpu(exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe, "480", "128", "128"], capture_output=True, text=True)
    assert r.returncode == 3 and r.stdout.startswith("no-device"), (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("load", [(0.0, 0, 278), (60.0, 0, 278), (480.0, 128, 128), (1440.0, 128, 128)])
def test_c_caller_matches_oracle(exe, oracle_mod, load):
    from workload_variant_autoscaler_b200.fleet import config1_fleet
    rpm, it, ot = load
    r = subprocess.run([exe, repr(rpm), str(it), str(ot)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    rows = {ln.split()[0]: ln.split()[1:] for ln in r.stdout.strip().splitlines()}
    f = config1_fleet(rpm, it, ot)
    cand_o, win_o = oracle_mod.solve(f)

    def check(tag, rec):
        got = rows[tag]
        assert [int(x) for x in got[:4]] == [int(rec["feasible"]), int(rec["acc"]), int(rec["replicas"]), int(rec["batch"])]
        for k, name in enumerate(("cost", "value", "itl", "ttft", "rho", "max_rate")):
            assert int(got[4 + k], 16) == int(np.float32(rec[name]).view(np.uint32)), (tag, name)
    check("candidate", cand_o.reshape(-1)[0])
    check("winner", win_o[0])
    t = oracle_mod.allocate_by_type(f, win_o)[0]
    d = oracle_mod.allocation_diffs(f, win_o)[0]
    assert [int(x) for x in rows["type"][:3]] == [int(t["present"]), int(t["count"]), int(t["limit"])]
    assert int(rows["type"][3], 16) == int(np.float32(t["cost"]).view(np.uint32))
    assert [int(x) for x in rows["diff"][:4]] == [int(d["old_acc"]), int(d["new_acc"]), int(d["old_replicas"]),
                                                  int(d["new_replicas"])]
    assert int(rows["diff"][4], 16) == int(np.float32(d["cost_diff"]).view(np.uint32))
    assert int(rows["launches"][0]) > 0
