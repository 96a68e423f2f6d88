import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import tamc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
g = int(sys.argv[2]) if len(sys.argv) > 2 else 64
c = tamc.configs.scaled("skin200", g)
t = tamc.MCTransport(g, g, g, c["xmax"], c["ymax"], c["zmax"])
t.set_optics(c["rhokap"](), c["albedo"], c["hgg"], flags=c["flags"])
t0 = time.time()
print(t.trace_probe(n, 1234), time.time() - t0)
t.run_async(n, 1234); t.sync(); print(t.get_stats())
