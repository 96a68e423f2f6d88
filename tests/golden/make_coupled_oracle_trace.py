#!/usr/bin/env python
"""Generator of tests/golden/coupled_shipped_events.json (test infrastructure: it runs the oracle).  CPU oracle version of tools/coupled_trace.py (photon loop on the same Philox ids + heat step): slow,
one-off validation of where the shipped configuration first boils / ablates / diverges."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT]
from oracle import oracle as orc  # noqa: E402

n, npk = 80, 125000
stop_at = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
h = orc.HeatOracle(n, 0.03, 0.03, 0.06)
h.init()
o = orc.Oracle(n, n, n, 0.03, 0.03, 0.06)
o.set_optics(0.0, 0.9)
first = {}
t0 = time.time()
for it in range(stop_at):
    if h.scalar("time") > h.scalar("total_time"):
        break
    o.set_rhokap(h.array("rhokap"))
    o.zero_jmean()
    o.seed_philox(95648324, it * npk)
    o.run(npk)
    jm = np.asfortranarray(o.jmean.copy())
    h.scale_jmean(jm, npk)
    h.sim_3d(jm, it)
    h.arrhenius()
    h.setup_thermal_coeff(500.0)
    T = h.array("temp")[1:-1, 1:-1, 1:-1]
    if "boil" not in first and h.array("Q").max() > 0:
        first["boil"] = it
    if "ablate" not in first and (h.array("rhokap")[1:-1, 1:-1, 1:-1] == 0).any():
        first["ablate"] = it
    if not np.isfinite(T).all() or h.scalar("negative_temp"):
        first["diverged"] = it
        break
    if it % 250 == 0 or it in (3999, 4499):
        print(f"iter {it + 1:6d} time {h.scalar('time'):.4f} Tmax {T.max() - 273:9.2f} C boiling {(h.array('Q') > 0).sum():7d} "
              f"tissue>1 {(h.array('tissue') >= 1).sum()} ({time.time() - t0:.0f} s)", flush=True)
print("first events (0-based iteration):", first)
