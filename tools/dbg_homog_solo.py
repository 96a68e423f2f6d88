"""homog200's regrouped column kernel on ONE GPU of a busy box, no communicator: per-call kernel times (device events).
Started eight times side by side (CUDA_VISIBLE_DEVICES = 0..7) it tells a box property from a communicator effect."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import numpy as np
import tamc
c = tamc.configs.CONFIGS["homog200"]
t = tamc.MCTransport(200, 200, 200, c["xmax"], c["ymax"], c["zmax"])
t.set_optics(c["rhokap"](), c["albedo"], c["hgg"], flags=0)
ks = []
t_end = time.time() + float(sys.argv[1]) if len(sys.argv) > 1 else time.time() + 8.0
i = 0
while time.time() < t_end:
    t.run_async(100_000_000, 1, i * 100_000_000); t.sync()
    ks.append(t.get_stats()["kernel_ms"]); i += 1
ks = np.array(ks[3:])
print("gpu", os.environ.get("CUDA_VISIBLE_DEVICES"), "calls", len(ks), "kernel_ms min %.3f median %.3f p95 %.3f max %.3f" % (ks.min(), np.median(ks), np.percentile(ks, 95), ks.max()), "form", t.get_option("form"), flush=True)
