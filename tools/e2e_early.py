"""A/B of the "io_early" option (start the full-grid upload / the zero fill beside the column gather) on the e2e call.
Usage: python tools/e2e_early.py [packets] [calls]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np  # noqa: E402

import tamc  # noqa: E402

packets = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 30
cfg = tamc.configs.CONFIGS["homog200"]
n = cfg["n"]
rk = cfg["rhokap"]()
tamc.pin_host(rk)
t = tamc.MCTransport(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"])
jm = t.new_jmean()
tamc.pin_host(jm)
for rep in range(2):
    for early in (0, 1, 2, 3):
        t.set_option("io_early", early)
        for _ in range(3):
            t.run_optics(rk, cfg["albedo"], cfg["hgg"], packets, 7, flags=cfg["flags"], out=jm)
        walls, parts = [], []
        for _ in range(calls):
            t0 = time.perf_counter()
            _, st = t.run_optics(rk, cfg["albedo"], cfg["hgg"], packets, 7, flags=cfg["flags"], out=jm)
            walls.append(1e3 * (time.perf_counter() - t0))
            parts.append([st[k] for k in ("h2d_ms", "zero_ms", "kernel_ms", "d2h_ms")])
        p = np.mean(parts, axis=0)
        print(f"io_early={early} io_form={t.get_option('io_form')}: wall mean {np.mean(walls):.3f} median {np.median(walls):.3f} min {np.min(walls):.3f} ms; "
              f"h2d {p[0]:.3f} zero {p[1]:.3f} kernel {p[2]:.3f} d2h {p[3]:.3f}", flush=True)
