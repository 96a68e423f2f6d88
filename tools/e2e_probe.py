"""Wall-clock anatomy of the host boundary (tamc_run_optics) on homog200: per-call wall time and the device parts, for
box_io on / off.  Usage: python tools/e2e_probe.py [packets] [calls]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np  # noqa: E402

import tamc  # noqa: E402

packets = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
name = sys.argv[3] if len(sys.argv) > 3 else "homog200"
cfg = tamc.configs.CONFIGS[name]
n = cfg["n"]
rk = cfg["rhokap"]()
tamc.pin_host(rk)
for box_io in (0, -1):
    t = tamc.MCTransport(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"])
    t.set_option("box_io", box_io)
    jm = t.new_jmean()
    tamc.pin_host(jm)
    for _ in range(3):
        t.run_optics(rk, cfg["albedo"], cfg["hgg"], packets, 7, flags=cfg["flags"], out=jm)
    walls, parts = [], []
    for _ in range(calls):
        t0 = time.perf_counter()
        _, st = t.run_optics(rk, cfg["albedo"], cfg["hgg"], packets, 7, flags=cfg["flags"], out=jm)
        walls.append(1e3 * (time.perf_counter() - t0))
        parts.append([st[k] for k in ("h2d_ms", "zero_ms", "kernel_ms", "allreduce_ms", "d2h_ms")])
    p = np.mean(parts, axis=0)
    print(f"{name} box_io={box_io} io_form={t.get_option('io_form')} form={t.get_option('form')}: wall mean {np.mean(walls):.3f} "
          f"min {np.min(walls):.3f} ms; h2d {p[0]:.3f} zero {p[1]:.3f} kernel {p[2]:.3f} ar {p[3]:.3f} d2h {p[4]:.3f} "
          f"sum {p.sum():.3f}; jmean/packet {jm.sum() / packets:.6f}")
    tamc.unpin_host(jm)
    t.close()
