#!/usr/bin/env python
"""Driver for compute-sanitizer, round 2: the flight kernel (separate arrays / interleaved records / per-voxel optics grids /
aggregated-RED build), the trace probe (k_trace, k_probe_trace), k_column_bound_resident, at small sizes.  No timing claims."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np  # noqa: E402

import tamc  # noqa: E402

npk = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
cfg = tamc.configs.scaled("skin200", 24)
g = cfg["n"]
rk = cfg["rhokap"]()
t = tamc.MCTransport(g, g, g, cfg["xmax"], cfg["ymax"], cfg["zmax"])
t.set_optics(rk, cfg["albedo"], cfg["hgg"], flags=cfg["flags"])
for opts in ({}, {"flight_inter": 1}, {"flight_agg": 1}, {"flight_regs": 2}, {"flight_regs": 4}, {"flight_launch_min": 1}, {"walk_min": 1}, {"walk_min": 32}):
    for k, v in opts.items():
        t.set_option(k, v)
    t.run_async(npk, 5, 0)
    st = t.get_stats()
    assert st["packets"] == npk and t.get_option("form") == 9
    print("flight", opts, "steps", st["voxel_steps"], "scatters", st["scatters"])
    for k in opts:
        t.set_option(k, {"flight_inter": -1, "flight_agg": 0, "flight_regs": 0, "flight_launch_min": 3, "walk_min": 8}[k])
shape = rk.shape
alb = np.full(shape, 0.9, order="F"); alb[:, :, g // 2:] = 0.99
hg = np.full(shape, 0.5, order="F"); hg[:, :, g // 2:] = 0.9
t.set_optics_grids(alb, hg, None)
for inter in (0, 1):
    t.set_option("flight_inter", inter)
    t.run_async(npk, 5, 0)
    print("flight + grids, inter", inter, t.get_stats()["scatters"])
t.set_option("flight_inter", -1)
t.set_optics_grids(None, None, None)
print("trace probe", t.trace_probe(npk, 5)["voxel_steps"])
t.close()
# stub regime, bound from the resident grid without a communicator ("reduce_bound" = 2) in a non-column form
c = tamc.configs.CONFIGS["shipped80"]
t = tamc.MCTransport(80, 80, 80, c["xmax"], c["ymax"], c["zmax"])
t.set_optics(c["rhokap"](), 0.0, 0.9, flags=0)
t.set_option("reduce_bound", 2)
t.run_async(20000, 5, 0)
print("resident bound planes", t.get_option("reduce_planes"), "form", t.get_option("form"))
t.close()
# depth-limited columns-first upload on a grid whose packets go far below the copied planes: k_column_bound fills the
# resident grid from the caller's page-locked array (untiled and regrouped column kernels)
nx, ny, nz = 48, 40, 96
ii, jj, kk = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
rk = np.zeros((nx + 2, ny + 2, nz + 2), order="F")
rk[1:-1, 1:-1, 1:-1] = 100.0 * (1.0 + 0.25 * ((ii + 2 * jj + 3 * kk) % 4))
tamc.pin_host(rk)
for tile, park in ((0, -1), (12, 1)):
    t = tamc.MCTransport(nx, ny, nz, 0.03, 0.03, 0.06)
    for k, v in (("column", 1), ("column_tile", tile), ("column_park", park), ("gather_depth", 5)):
        t.set_option(k, v)
    jm = t.new_jmean()
    tamc.pin_host(jm)
    _, st = t.run_optics(rk, 0.0, 0.9, 4 * npk, 5, out=jm)
    print("deep fill: io_form", t.get_option("io_form"), "form", t.get_option("form"), "bottom exits", st["exits"][4], "depth_hint", t.get_option("depth_hint"))
    assert t.get_option("io_form") == 7
    tamc.unpin_host(jm)
    t.close()
tamc.unpin_host(rk)
