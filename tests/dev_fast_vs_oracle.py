#!/usr/bin/env python
"""Diagnostic: per-packet deviation of the production arithmetic (variant 1) and of the exact one
(variant 2) from the oracle on the same Philox stream."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import tamc  # noqa: E402
from tests.util import make_oracle, make_transport, INT_FIELDS  # noqa: E402

for name, n, npk in (("turbid200", 60, 20000), ("skin200", 100, 20000), ("skin200", 200, 20000)):
    cfg = tamc.configs.scaled(name, n)
    o = make_oracle(cfg)
    o.seed_philox(7, 0)
    want = o.run(npk, records=True)["records"]
    t = make_transport(cfg)
    for variant in (2, 1):
        t.set_option("variant", variant)
        rec, _ = t.run_records(npk, 7, 0)
        flips = np.zeros(npk, bool)
        for f in INT_FIELDS:
            flips |= rec[f] != want[f]
        err = np.zeros(npk)
        for f, s in (("xp", cfg["xmax"]), ("yp", cfg["ymax"]), ("zp", cfg["zmax"]), ("nxp", 1), ("nyp", 1), ("nzp", 1)):
            err = np.maximum(err, np.abs(rec[f] - want[f]) / np.maximum(np.abs(want[f]), s))
        derr = np.abs(rec["deposit"] - want["deposit"]) / np.abs(want["deposit"])
        q = np.quantile(err, [0.5, 0.99, 0.999])
        print(f"{name}{n} variant {variant}: flips {flips.sum()}/{npk}  pos/dir err median {q[0]:.1e} p99 {q[1]:.1e} p99.9 {q[2]:.1e} max {err.max():.1e}"
              f"  frac>1e-6 {np.mean(err > 1e-6):.2e}  deposit err max {derr.max():.1e} nscatt mean {want['nscatt'].mean():.1f}")
    t.close()
