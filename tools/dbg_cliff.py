"""The cliff behind the depth-limited columns-first upload: homog200, 1e8 packets per tamc_run_optics call,
the opacity drops 16-fold between two calls (a different tissue, not an ablation front)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np
import tamc
c = tamc.configs.CONFIGS["homog200"]
rk_a = c["rhokap"]()
rk_b = np.asfortranarray(rk_a / 16.0)
tamc.pin_host(rk_a); tamc.pin_host(rk_b)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
t = tamc.MCTransport(200, 200, 200, c["xmax"], c["ymax"], c["zmax"])
jm = t.new_jmean(); tamc.pin_host(jm)
ref = tamc.MCTransport(200, 200, 200, c["xmax"], c["ymax"], c["zmax"])
ref.set_option("gather_depth", 0)
jr = ref.new_jmean(); tamc.pin_host(jr)
for i, which in enumerate("AAABBBAAB"):
    g = rk_a if which == "A" else rk_b
    t0 = time.perf_counter(); _, st = t.run_optics(g, c["albedo"], c["hgg"], per, 1 + i, flags=0, out=jm)
    ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter(); ref.run_optics(g, c["albedo"], c["hgg"], per, 1 + i, flags=0, out=jr)
    ms_ref = 1e3 * (time.perf_counter() - t0)
    err = float(np.max(np.abs(jm - jr)) / np.max(jr))
    print(which, i, "ms %.3f (every plane gathered: %.3f)" % (ms, ms_ref), {k: round(st[k], 3) for k in ("kernel_ms", "h2d_ms", "d2h_ms")},
          "form", t.get_option("form"), "io_form", t.get_option("io_form"), "depth_hint", t.get_option("depth_hint"), "max rel diff %.2e" % err, flush=True)
