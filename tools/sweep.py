#!/usr/bin/env python
"""Sweep the transport kernel's tuning knobs for one workload; prints packets/s per setting."""
import argparse
import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import tamc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="homog200")
ap.add_argument("--packets", type=int, default=20_000_000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--grid", default="variant=0,1;chunk=0;scatter_min=1,16;block=256;ctas_per_sm=0")
a = ap.parse_args()
c = tamc.configs.CONFIGS[a.workload]
t = tamc.MCTransport(c["n"], c["n"], c["n"], c["xmax"], c["ymax"], c["zmax"])
t.set_optics(c["rhokap"](), c["albedo"], c["hgg"], flags=c["flags"])
axes = [(kv.split("=")[0], [int(x) for x in kv.split("=")[1].split(",")]) for kv in a.grid.split(";")]
names = [k for k, _ in axes]
first = {k: v[0] for k, v in axes}
for combo in itertools.product(*[v for _, v in axes]):
    opts = dict(zip(names, combo))
    if opts.get("variant", 1) == 0 and any(opts.get(k, first.get(k)) != first.get(k) for k in ("chunk", "scatter_min")):
        continue
    try:
        for k, v in opts.items():
            t.set_option(k, v)
        best = 1e30
        for _ in range(a.reps):
            t.run_async(a.packets, 20261017, 0)
            t.sync()
            st = t.get_stats()
            best = min(best, st["kernel_ms"])
        print(f"{a.workload} {opts} kernel_ms={best:.3f} packets/s={a.packets / best * 1e3:.4g} vsteps/s={st['voxel_steps'] / best * 1e3:.4g}", flush=True)
    except tamc.TamcError as e:
        print(opts, "ERR", e, flush=True)
t.close()
