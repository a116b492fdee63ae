"""Target for ncu: a few grid_solve calls on BASELINE config 2 (see profiles/README.md)."""
import sys
sys.path.insert(0, ".")
from workload_variant_autoscaler_b200 import Engine, config2_grid, synth_fleet
e = Engine(0)
f = synth_fleet(100, 4, seed=42)
g = config2_grid()
for _ in range(3):
    e.grid_solve(f, g)
print("kernel ms", e.last_kernel_ms)
