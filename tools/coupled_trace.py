#!/usr/bin/env python
"""Run the device-resident coupled loop with the shipped parameters in chunks and print the state."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import tamc  # noqa: E402

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 500
n = 80
t = tamc.MCTransport(n, n, n, 0.03, 0.03, 0.06)
t.set_optics(tamc.gridset(0.03, 0.03, 0.06, n, n, n, 680.0)[3], 0.0, 0.9)
t.heat_init()
done = 0
while True:
    try:
        it, pk = t.coupled_loop(125000, 95648324, chunk)
    except tamc.TamcError as e:
        print("stopped:", e)
        it = chunk
        stop = True
    else:
        stop = it < chunk
    done += it
    T = t.heat_array("temp")[1:-1, 1:-1, 1:-1]
    rk = t.heat_array("rhokap")[1:-1, 1:-1, 1:-1]
    q = t.heat_array("Q")
    print(f"iter {done:6d} time {t.heat_scalar('time'):.4f} pwr {t.heat_scalar('pwr'):8.3f} Tmax {np.nanmax(T) - 273:9.2f} C Tmin {np.nanmin(T) - 273:9.2f} C "
          f"finite {np.isfinite(T).all()} ablated {(rk == 0).sum():7d} boiling {(q > 0).sum():7d} tissue>1 {(t.heat_array('tissue') >= 1).sum()}", flush=True)
    if stop:
        break
t.close()
