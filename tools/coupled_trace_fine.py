#!/usr/bin/env python
"""GPU counterpart of tools/coupled_oracle_trace.py: first boil / ablate / divergence iteration."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import tamc  # noqa: E402

n = 80
t = tamc.MCTransport(n, n, n, 0.03, 0.03, 0.06)
t.set_optics(tamc.gridset(0.03, 0.03, 0.06, n, n, n, 680.0)[3], 0.0, 0.9)
t.heat_init()
first, done = {}, 0
marks = {4000, 4500}
while done < 5200:
    step = 1 if done >= 3600 else 100
    try:
        it, _ = t.coupled_loop(125000, 95648324, step)
    except tamc.TamcError:
        first["diverged"] = done
        break
    done += it
    if step == 1 or done % 500 == 0:
        T = t.heat_array("temp")[1:-1, 1:-1, 1:-1]
        q = t.heat_array("Q")
        if "boil" not in first and q.max() > 0:
            first["boil"] = done - 1
        if "ablate" not in first and (t.heat_array("rhokap")[1:-1, 1:-1, 1:-1] == 0).any():
            first["ablate"] = done - 1
        if not np.isfinite(T).all():
            first["diverged"] = done - 1
            break
        if done in marks or done % 500 == 0:
            print(f"iter {done:6d} time {t.heat_scalar('time'):.4f} Tmax {T.max() - 273:9.2f} C boiling {(q > 0).sum():7d} "
                  f"tissue>1 {(t.heat_array('tissue') >= 1).sum()}", flush=True)
print("first events (0-based iteration):", first)
