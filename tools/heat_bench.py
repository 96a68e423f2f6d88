#!/usr/bin/env python
"""Time the device heat / ablation step (tamc_heat_step) and the resident coupled loop."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import tamc  # noqa: E402

# algorithmic bytes per voxel of one tamc_heat_step call with loops = 1
BYTES = {"k_heat_step": 72, "k_post": 128, "k_rule_air": 40}
out = {}
for n, calls in ((80, 400), (200, 60)):
    t = tamc.MCTransport(n, n, n, 0.03, 0.03, 0.06)
    t.set_optics(tamc.gridset(0.03, 0.03, 0.06, n, n, n, 680.0)[3], 0.0, 0.9)
    t.heat_init()
    t.run_async(125000, 1)
    for _ in range(5):
        t.heat_step(125000)
    t.sync()
    t0 = time.perf_counter()
    for _ in range(calls):
        t.heat_step(125000)
    t.sync()
    dt = (time.perf_counter() - t0) / calls
    nbytes = sum(BYTES.values()) * n ** 3
    out[f"heat_step_{n}"] = {"us_per_call": 1e6 * dt, "algorithmic_GBps": nbytes / dt / 1e9, "bytes_per_voxel": sum(BYTES.values())}
    if n == 80:
        t.heat_init()
        t.seek(0)
        t0 = time.perf_counter()
        it, pk = t.coupled_loop(125000, 95648324, 3000)
        dt = time.perf_counter() - t0
        out["coupled_loop_80"] = {"iterations": it, "us_per_iteration": 1e6 * dt / it, "packets_per_s": pk / dt}
    t.close()
print(json.dumps(out))
