#!/usr/bin/env python
"""Driver for compute-sanitizer: the Gaussian beam and periodic lateral boundaries in every kernel that carries them
(exact, thread-per-packet, pool `ext` build, replay) at small sizes.  No timing claims."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np  # noqa: E402

import tamc  # noqa: E402

n, npk = 20, int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rk = tamc.gridset(0.02, 0.02, 0.2, n, n, n, 100.0)[3]
for sigma in (0.0, 0.015):
    for flags in (1 | 4, 1 | 2 | 4, 0):
        t = tamc.MCTransport(n, n, n, 0.02, 0.02, 0.2)
        t.set_source_co2(0.01)
        if sigma > 0:
            t.set_source_gaussian(sigma)
        t.set_optics(rk, 0.95 if flags & 1 else 0.0, 0.8, n1=1.0, n2=1.38, flags=flags)
        for variant in (0, 1, 2, 3):
            t.set_option("variant", variant)
            t.run_async(npk, 11, 0)
            st = t.get_stats()
            assert st["packets"] == npk and (not (flags & 4) or st["exits"][:4] == [0, 0, 0, 0])
            print("sigma", sigma, "flags", flags, "variant", variant, "form", t.get_option("form"), "steps", st["voxel_steps"])
        rec, _ = t.run_records(npk, 11, 0)
        if not (flags & 2):
            # replay kernel: any uniform draws will do under the sanitizer (fixed-capacity draw lists; a packet that runs
            # out of draws is reported with code 6 and is fine here)
            cap = 4000
            rng = np.random.default_rng(1)
            off = np.arange(npk + 1, dtype=np.int64) * cap
            try:
                rec, _ = t.run_replay(off, rng.random(npk * cap))
                print("  replay ok", int(rec["steps"].sum()))
            except tamc.TamcError as e:
                assert e.code == 6, e
                print("  replay ok (some packets ran out of their", cap, "draws)")
        t.close()

# depth-limited columns-first upload: deep groups read from the caller's page-locked grid (untiled, tiled, regrouped)
nx, ny, nz = 48, 40, 96
rk = np.zeros((nx + 2, ny + 2, nz + 2), order="F")
ii, jj, kk = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
rk[1:-1, 1:-1, 1:-1] = 30.0 * (1.0 + 0.25 * ((ii + 2 * jj + 3 * kk) % 4))
tamc.pin_host(rk)
big = 40 * npk
for tile, park in ((0, 0), (12, 0), (12, 1), (23, 1)):
    t = tamc.MCTransport(nx, ny, nz, 0.03, 0.03, 0.06)
    jm = t.new_jmean()
    tamc.pin_host(jm)
    t.set_option("column", 1)
    t.set_option("column_tile", tile)
    t.set_option("column_park", park)
    for gd in (1, 40, -1, -1):
        t.set_option("gather_depth", gd)
        _, st = t.run_optics(rk, 0.0, 0.9, big, 5, out=jm)
        assert np.array_equal(jm, t.get_jmean()) and st["exits"][4] > 0
        print("deep upload: tile", tile, "park", park, "gather_depth", gd, "io_form", t.get_option("io_form"), "form", t.get_option("form"),
              "depth_hint", t.get_option("depth_hint"), "steps", st["voxel_steps"])
    tamc.unpin_host(jm)
    t.close()
tamc.unpin_host(rk)
