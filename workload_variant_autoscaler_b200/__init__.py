"""B200-native implementation of the WVA optimizer hot path."""
from ._abi import ACC_NONE, ACC_UNKNOWN, Allocs  # noqa: F401
from .fleet import Fleet, Grid, config2_grid, synth_fleet  # noqa: F401
from .engine import Engine, WvaError, greedy_solve  # noqa: F401,E402
from . import adapters  # noqa: F401,E402
