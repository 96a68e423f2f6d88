#!/usr/bin/env python
"""Tiny driver for ncu: a few MC calls of one workload through the C ABI (no timing claims here)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import tamc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="homog200")
ap.add_argument("--packets", type=int, default=10_000_000)
ap.add_argument("--calls", type=int, default=3)
ap.add_argument("--option", action="append", default=[])
a = ap.parse_args()
c = tamc.configs.CONFIGS[a.workload]
t = tamc.MCTransport(c["n"], c["n"], c["n"], c["xmax"], c["ymax"], c["zmax"])
for kv in a.option:
    k, v = kv.split("=")
    t.set_option(k, int(v))
t.set_optics(c["rhokap"](), c["albedo"], c["hgg"], flags=c["flags"])
for _ in range(a.calls):
    t.run_async(a.packets, 20261017)
    t.sync()
    st = t.get_stats()
print(a.workload, a.packets, {k: st[k] for k in ("kernel_ms", "voxel_steps", "scatters", "packets")})
t.close()
