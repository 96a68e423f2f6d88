#!/usr/bin/env python
"""Driver for compute-sanitizer: the column-form kernels (untiled, tiled, regrouped walk) and the overlapped host boundary
(zero-copy gather, k_box_mirror, pitched DMA) at small sizes.  No timing claims."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np  # noqa: E402

import tamc  # noqa: E402

big = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
for name, n in (("shipped80", 80), ("homog200", 63)):
    cfg = tamc.configs.scaled(name, n)
    rk = cfg["rhokap"]()
    tamc.pin_host(rk)
    t = tamc.MCTransport(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"])
    jm = t.new_jmean()
    tamc.pin_host(jm)
    t.set_option("column", 1)
    for split, park in ((0, 0), (12, 0), (12, 1), (4, 1), (23, 1)):
        t.set_option("column_tile", split)
        t.set_option("column_park", park)
        _, st = t.run_optics(rk, 0.0, 0.9, big, 7, out=jm)
        assert t.get_option("io_form") & 3 == 3, t.get_option("io_form")
        assert np.array_equal(jm, t.get_jmean())
        print(name, n, "tile", split, "park", park, "form", t.get_option("form"), "steps", st["voxel_steps"], "sum/packet %.5f" % (jm.sum() / big))
    _, st = t.run(5_000_000 if len(sys.argv) > 2 else big, 7, out=jm)          # long call: pitched DMA download
    assert np.array_equal(jm, t.get_jmean())
    print(name, n, "run: io_form", t.get_option("io_form"), "packets", st["packets"])
    ms, steps = t.roofline_probe(big, 3)
    print(name, n, "probe steps", steps)
    tamc.unpin_host(jm)
    tamc.unpin_host(rk)
    t.close()
