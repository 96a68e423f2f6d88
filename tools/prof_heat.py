#!/usr/bin/env python
"""Tiny driver for ncu: a few heat / ablation steps on an n^3 grid (tamc_heat_step)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import tamc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
t = tamc.MCTransport(n, n, n, 0.03, 0.03, 0.06)
t.set_optics(tamc.gridset(0.03, 0.03, 0.06, n, n, n, 680.0)[3], 0.0, 0.9)
t.heat_init()
t.run_async(1000000, 1)
for _ in range(4):
    t.heat_step(1000000)
t.sync()
t.close()
