"""Column form: launch shape (resident CTAs per SM) and shared-memory tiles (column_tile = 10*ta + tb): kernel ms."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np  # noqa: E402

import tamc  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "homog200"
cfg = tamc.configs.CONFIGS[name]
n = cfg["n"]
t = tamc.MCTransport(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"])
t.set_optics(cfg["rhokap"](), cfg["albedo"], cfg["hgg"], flags=cfg["flags"])


def timed(packets, reps=6):
    ms = []
    for _ in range(reps):
        t.flush_l2()
        t.run_async(packets, 11)
        t.sync()
        ms.append(t.get_stats()["kernel_ms"])
    return float(np.median(ms[1:]))


for packets in (16_000_000, 100_000_000, 100_000_000):
    row = []
    t.set_option("column", 1)
    t.set_option("column_tile", 0)
    for ctas in (0, 2, 3, 4, 5):
        t.set_option("ctas_per_sm", ctas)
        row.append(f"ctas{ctas}:{timed(packets):.3f}")
    t.set_option("ctas_per_sm", 0)
    for split in (1, 11, 15, 25):
        t.set_option("column_tile", split)
        row.append(f"tile{split}:{timed(packets):.3f}")
    t.set_option("column_tile", 0)
    t.set_option("column", 0)
    row.append(f"nocolumn(form {0}):{timed(packets):.3f}")
    row[-1] = row[-1].replace("form 0", f"form {t.get_option('form')}")
    t.set_option("column", 1)
    row.append(f"ctas0:{timed(packets):.3f}")
    print(name, packets, " ".join(row), flush=True)
t.close()
