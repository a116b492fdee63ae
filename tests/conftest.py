import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def engine():
    import torch  # noqa: F401  (device plumbing only)
    from workload_variant_autoscaler_b200 import Engine, WvaError, _abi
    try:
        eng = Engine(0)
    except WvaError as exc:
        if exc.code == _abi.WVA_ERR_NO_DEVICE:  # a CPU-only box: the GPU tests are skipped, not errors
            pytest.skip("no CUDA device: GPU parity tests need the B200 box")
        raise
    yield eng
    eng.close()
