"""Bit-exact comparison helpers shared by the parity tests."""
import numpy as np

ALLOC_FIELDS = ("feasible", "acc", "replicas", "batch", "cost", "value", "itl", "ttft", "rho", "max_rate")
FLOAT_FIELDS = ("cost", "value", "itl", "ttft", "rho", "max_rate")


def bits(a):
    a = np.ascontiguousarray(a, np.float32)
    b = a.view(np.uint32).copy()
    b[np.isnan(a)] = 0x7FC00000  # any NaN == any NaN (payload is not part of the contract)
    return b


def assert_f32_bits_equal(got, want, what=""):
    gb, wb = bits(got), bits(want)
    bad = np.nonzero(gb != wb)[0]
    if bad.size:
        i = int(bad[0])
        raise AssertionError(f"{what}: {bad.size} of {gb.size} float32 values differ; first at {i}: "
                             f"got {np.float32(got[i])!r} ({gb[i]:#010x}) want {np.float32(want[i])!r} ({wb[i]:#010x})")


def assert_allocs_equal(got, want_struct, what=""):
    """got: product Allocs (SoA); want_struct: oracle structured array (flattened)."""
    want = want_struct.reshape(-1)
    assert got.n == want.size, (got.n, want.size)
    for f in ("feasible", "acc", "replicas", "batch"):
        g = np.asarray(getattr(got, f)).astype(np.int64)
        w = want[f].astype(np.int64)
        bad = np.nonzero(g != w)[0]
        if bad.size:
            i = int(bad[0])
            raise AssertionError(f"{what}: field {f}: {bad.size} of {g.size} differ; first at {i}: got {g[i]} want {w[i]}"
                                 f"\n got  {got.record(i)}\n want {want[i]}")
    for f in FLOAT_FIELDS:
        assert_f32_bits_equal(getattr(got, f), want[f], f"{what}: field {f}")
