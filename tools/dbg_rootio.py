import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tissue-ablation-mc_b200")]
import torch, torch.distributed as dist
import tamc, bench
from tamc import dist as tdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
c = tamc.configs.CONFIGS["homog200"]
rk = c["rhokap"](); rk_b = bench.crater_variant(c, rk, 20, 3)
tamc.pin_host(rk); tamc.pin_host(rk_b)
for mode in (0, 1):
    t = tamc.MCTransport(200, 200, 200, c["xmax"], c["ymax"], c["zmax"], device=rank)
    t.set_option("root_io", mode)
    t.set_optics(rk, 0.0, 0.9, flags=0)
    t.comm_init(world, rank, tdist.broadcast_unique_id(tamc.comm_unique_id, dist, torch.device("cuda", rank)))
    jm = t.new_jmean(); tamc.pin_host(jm)
    for i in range(3):
        t.run_optics([rk, rk_b][i % 2], 0.0, 0.9, 100_000_000, 1, flags=0, out=jm)
    dist.barrier(); torch.cuda.synchronize()
    os.environ["TAMC_TRACE"] = "1"
    t0 = time.perf_counter()
    for i in range(4):
        t.run_optics([rk, rk_b][i % 2], 0.0, 0.9, 100_000_000, 1, flags=0, out=jm)
    dt = (time.perf_counter() - t0) / 4
    os.environ.pop("TAMC_TRACE")
    print(f"mode {mode} rank {rank}: {dt*1e3:.3f} ms per call", flush=True)
    tamc.unpin_host(jm); t.close()
dist.barrier(); dist.destroy_process_group()
