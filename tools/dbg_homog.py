import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np
import tamc
import bench
c = tamc.configs.CONFIGS["homog200"]
rk = c["rhokap"]()
rk_b = bench.crater_variant(c, rk, 20, 3)
tamc.pin_host(rk); tamc.pin_host(rk_b)
t = tamc.MCTransport(200, 200, 200, c["xmax"], c["ymax"], c["zmax"])
t.set_optics(rk, c["albedo"], c["hgg"], flags=0)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
for i in range(3):
    t0 = time.time(); t.run_async(per, 1); t.sync(); st = t.get_stats()
    print("resident", i, round(time.time() - t0, 4), st["kernel_ms"], t.get_option("form"), flush=True)
jm = t.new_jmean(); tamc.pin_host(jm)
for i in range(6):
    g = [rk, rk_b][i % 2]
    t0 = time.time(); _, st = t.run_optics(g, c["albedo"], c["hgg"], per, 1, flags=0, out=jm)
    print("e2e", i, round(time.time() - t0, 4), {k: round(st[k], 3) for k in ("kernel_ms", "h2d_ms", "d2h_ms")}, t.get_option("form"), t.get_option("io_form"), t.get_option("depth_hint"), flush=True)
