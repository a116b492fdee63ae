"""Diagnostics (-DWVA_SZCNT build): loop counters of the full-length N = 256 chains over one resolve of the streaming workload."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from workload_variant_autoscaler_b200 import Engine, synth_fleet  # noqa: E402

e = Engine(0)
L = e._L
f = synth_fleet(12500, 8, seed=45, max_batch_choices=(4, 8, 16, 32, 64, 128, 256))
e.upload(f)
e.resolve()
out = (C.c_ulonglong * 24)()
assert L.wva_dbg_szcnt(out, 1) == 0, "library was built without -DWVA_SZCNT"
e.resolve()
L.wva_dbg_szcnt(out, 0)
c = np.array(list(out), dtype=np.float64)
n = c[16]
print("full-length N=256 solves:", int(n), " with a 16-step window:", int(c[17]), " mean nh:", c[18] / n)
print("pass 1 per solve: 16-blocks %.1f  4-blocks %.1f  per-step pairs %.1f  slow %.2f" % (c[0] / n, c[1] / n, c[2] / n, c[3] / n))
print("  16-votes %.1f (mean active lanes %.1f): own clause failed: head %.1f  end %.1f  window %.1f" %
      (c[8] / n, c[9] / max(c[8], 1), c[4] / n, c[5] / n, c[6] / n))
print("pass 2 per solve: 16-blocks %.1f  4-blocks %.1f  per-step pairs %.1f ; own clause failed: end %.1f window %.1f" %
      (c[10] / n, c[11] / n, c[12] / n, c[14] / n, c[15] / n))
