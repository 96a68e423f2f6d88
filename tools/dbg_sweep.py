import os, sys
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np
mode = sys.argv[1]
if mode == "torch":
    import torch
    torch.cuda.init(); torch.zeros(1, device="cuda")
import tamc
cfg = tamc.configs.CONFIGS["homog200"]
n = cfg["n"]
rk = cfg["rhokap"]()
if mode == "pin":
    tamc.pin_host(rk)
t = tamc.MCTransport(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"])
t.set_optics(rk, cfg["albedo"], cfg["hgg"], flags=cfg["flags"])
def timed(t, packets, seed, flush=True, reps=6):
    ms = []
    for _ in range(reps):
        if flush: t.flush_l2()
        t.run_async(packets, seed)
        if mode == "nosync":
            st = t.get_stats()
        else:
            t.sync(); st = t.get_stats()
        ms.append(st["kernel_ms"])
    return float(np.median(ms[1:]))
t.set_option("column_tile", 0)
if mode == "warm":
    for _ in range(3):
        t.run_async(100_000_000, 5); t.sync()
print(mode, "tile off", timed(t, 100_000_000, 20261017))
t.set_option("column_tile", -1)
print(mode, "tile auto", timed(t, 100_000_000, 20261017), t.get_option("form"))
