"""Kernel time of one small MC call (the coupled loop's 125 000 packets on the shipped 80^3 grid) per kernel form."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np  # noqa: E402

import tamc  # noqa: E402

cfg = tamc.configs.CONFIGS["shipped80"]
n = cfg["n"]
t = tamc.MCTransport(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"])
t.set_optics(cfg["rhokap"](), cfg["albedo"], cfg["hgg"], flags=cfg["flags"])
for packets in (125_000, 1_000_000):
    for label, opts in (("persistent", dict(variant=1)), ("simple", dict(variant=0)), ("persistent chunk32", dict(variant=1, chunk=32)),
                        ("persistent block128", dict(variant=1, block=128)),
                        ("column", dict(variant=3, column=1, column_tile=0)), ("column resident", dict(variant=3, column=2, column_tile=0)),
                        ("column tile12", dict(variant=3, column=1, column_tile=12, column_park=0)), ("auto", dict(variant=3))):
        for k, v in dict(variant=3, column=-1, column_tile=-1, column_park=-1, chunk=0, block=0).items():
            t.set_option(k, v)
        for k, v in opts.items():
            t.set_option(k, v)
        ms = []
        for _ in range(30):
            t.run_async(packets, 3)
            t.sync()
            st = t.get_stats()
            ms.append(st["kernel_ms"])
        print(f"{packets:8d} {label:22s} kernel {1e3 * np.median(ms[5:]):7.1f} us  zero {1e3 * st['zero_ms']:.1f} us  form {t.get_option('form')} launches {st['gpu_launches']}", flush=True)
t.close()
