#!/usr/bin/env python
"""bench.py -- photon packets/s and voxel-steps/s of the MC transport hot path on N B200s.

A "step" is one MC call (the replacement of /root/reference/src/mcpolar.f90:151-173): clear the
tally, transport P packets per GPU, all-reduce the tally.  Default workload = BASELINE.json
configs[1]: homogeneous 200^3 tissue cube with the reference's extents/optics/source, 1e8 packets
per GPU per step (weak scaling: every rank runs its own 1e8, like the reference's per-rank
`do j = 1, nphotons`).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference ...                     the reference's CPU path (oracle port, all host threads)

One JSON line on stdout (rank 0).  `value` times the device-resident call (inputs already in HBM);
`e2e` times tamc_run_optics (= tamc_set_optics + tamc_run) with pinned HOST buffers, copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "tissue-ablation-mc_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "photon_packets_per_s"
UNIT = "packets/s"
SEED = 20261017
BYTES_PER_VOXEL_STEP = 16          # 8 B rhokap read + 8 B jmean accumulate
HBM_FALLBACK_GBS = 6650.0          # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
FORMS = {0: "k_transport_simple", 1: "k_transport_persistent", 2: "k_transport_exact", 3: "k_transport_pool",
         4: "k_transport_stub_tiled", 5: "k_transport_column", 6: "k_transport_column", 7: "k_transport_column_tiled", 8: "k_transport_column_parked"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="homog200")
    ap.add_argument("--packets", type=int, default=0, help="packets per GPU per step (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary skin200 measurement")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--option", action="append", default=[], help="name=value passed to tamc_set_option")
    return ap.parse_args()


def workload_desc(name, cfg, packets):
    return (f"{name}: {cfg['n']}^3 voxels, xmax/ymax/zmax={cfg['xmax']}/{cfg['ymax']}/{cfg['zmax']} cm, "
            f"albedo={cfg['albedo']:.4g}, hgg={cfg['hgg']}, scatter={'on' if cfg['flags'] & 1 else 'off (shipped stub)'}, "
            f"{packets:.3g} packets per GPU per step")


# ----------------------------------------------------------------------------------------------------
# CPU legs (the only places bench.py touches oracle/)
# ----------------------------------------------------------------------------------------------------
def _oracle_lib_dir():
    """Build the timed CPU baseline (-O3 -march=native -flto, mirroring src/Makefile:3) on this box."""
    from oracle import oracle as orc

    d = tempfile.mkdtemp(prefix="tamc_oracle_")
    try:
        orc.build(fast=True, out_dir=d)
        return d, True
    except Exception:
        return None, False


def cpu_run(cfg, rk, nranks, packets_per_rank, fast_dir):
    from oracle import oracle as orc

    n = cfg["n"]
    return orc.run_ranks(nranks, n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"], rk, cfg["albedo"], cfg["hgg"],
                         packets_per_rank, flags=cfg["flags"], fast=fast_dir is not None, out_dir=fast_dir)


def cpu_threads():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return max(1, min(n, int(os.environ.get("TAMC_CPU_THREADS", "64"))))


def cpu_calibrate(cfg, rk, fast_dir, seconds, threads):
    """Packets per rank so that `threads` ranks take about `seconds` in total."""
    probe = 20000 if cfg["flags"] & 1 else 200000
    r = cpu_run(cfg, rk, 1, probe, fast_dir)
    rate = probe / max(r["seconds"], 1e-6)
    return max(1000, int(rate * seconds * 0.8)), rate


def run_reference(args, cfg, name):
    """--impl reference: the reference's own CPU implementation of the path.  The Fortran cannot be
    built in this image (no Fortran front-end, no MPI), so this is the C restatement in oracle/,
    R emulated MPI ranks on all host threads + the in-memory jmean sum (kind = "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rk = cfg["rhokap"]()
    fast_dir, fast = _oracle_lib_dir()
    threads = cpu_threads()
    per_rank, _ = cpu_calibrate(cfg, rk, fast_dir, 4.0, threads)
    for _ in range(args.warmup):
        cpu_run(cfg, rk, threads, max(1000, per_rank // 8), fast_dir)
    tot_s, tot_p, tot_v = 0.0, 0, 0
    for _ in range(args.steps):
        r = cpu_run(cfg, rk, threads, per_rank, fast_dir)
        tot_s += r["seconds"]
        tot_p += r["stats"]["packets"]
        tot_v += r["stats"]["voxel_steps"]
    value = tot_p / tot_s
    sample = f"{threads} emulated MPI ranks x {per_rank} packets per step (ran2 seeds per mcpolar.f90:97-98) + jmean sum"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_desc(name, cfg, per_rank * threads), "parallelism": f"{threads} host threads"},
        "voxel_steps_per_s": tot_v / tot_s,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "flags": "-O3 -march=native -flto" if fast else "-O2 -ffp-contract=off"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_id):
        self.path = os.path.join(tempfile.gettempdir(), f"tamc_clocks_{os.getpid()}.csv")
        self.f = open(self.path, "w")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_id), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


# ----------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------
def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(name):
    """DRAM bytes per launch of the transport kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


def timed_steps(t, stream, packets, steps, world, dist, dev, torch, flush=True):
    """K MC calls with inputs resident; per-step CUDA events on the library's stream; max over ranks."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sums = {"kernel_ms": 0.0, "allreduce_ms": 0.0, "zero_ms": 0.0, "voxel_steps": 0, "scatters": 0, "launches": 0}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    w0 = time.perf_counter()
    for a, b in ev:
        if flush:
            t.flush_l2(256 << 20)                 # evict L2 between steps, outside the timed events
        a.record(stream)
        t.run_async(packets, SEED)                # ids from the cursor: fresh packets every step
        b.record(stream)
        st = t.get_stats()                        # syncs the stream; per-step counters and device times
        for k in ("kernel_ms", "allreduce_ms", "zero_ms"):
            sums[k] += st[k]
        sums["voxel_steps"] += st["voxel_steps"]
        sums["scatters"] += st["scatters"]
        sums["launches"] += st["gpu_launches"]
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - w0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    red = torch.tensor([dev_ms, sums["kernel_ms"], sums["allreduce_ms"], wall * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([sums["voxel_steps"], sums["scatters"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    red, tot = red.tolist(), tot.tolist()
    return {"dev_ms": red[0], "kernel_ms": red[1], "allreduce_ms": red[2], "wall_ms": red[3],
            "voxel_steps": tot[0], "scatters": tot[1], "launches": sums["launches"],
            "local_kernel_ms": sums["kernel_ms"], "local_voxel_steps": sums["voxel_steps"]}


def run_ours(args, cfg, name):
    import numpy as np
    import torch
    import torch.distributed as dist

    import tamc

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or tamc.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the transport has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)

    n = cfg["n"]
    packets = args.packets or min(cfg["nphotons"], 100_000_000)
    rk = cfg["rhokap"]()
    tamc.pin_host(rk)
    t = tamc.MCTransport(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"], device=local)
    for kv in args.option:
        k, v = kv.split("=")
        t.set_option(k, int(v))
    t.set_optics(rk, cfg["albedo"], cfg["hgg"], flags=cfg["flags"])
    if world > 1:
        from tamc import dist as tdist

        t.comm_init(world, rank, tdist.broadcast_unique_id(tamc.comm_unique_id, dist, dev))
    stream = torch.cuda.ExternalStream(t.stream, device=dev)

    try:
        gpu_id = str(torch.cuda.get_device_properties(local).uuid)
        if not gpu_id.startswith("GPU-"):
            gpu_id = "GPU-" + gpu_id
    except Exception:
        gpu_id = str(local)

    # ---- warm-up, then the timed device-resident steps
    for _ in range(max(args.warmup, 0)):
        t.run_async(packets, SEED)
        t.sync()
    sampler = ClockSampler(gpu_id) if rank == 0 else None
    res = timed_steps(t, stream, packets, args.steps, world, dist, dev, torch)
    total_packets = float(packets) * world * args.steps
    value = total_packets / (res["dev_ms"] * 1e-3)
    vsteps_per_s = res["voxel_steps"] / (res["dev_ms"] * 1e-3)

    # ---- e2e: the reference-facing call with host buffers (upload rhokap, run, download jmean)
    jm = t.new_jmean()
    tamc.pin_host(jm)
    for _ in range(2):
        t.run_optics(rk, cfg["albedo"], cfg["hgg"], packets, SEED, flags=cfg["flags"], out=jm)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0 = time.perf_counter()
    e2e_parts = {"h2d_ms": 0.0, "d2h_ms": 0.0, "kernel_ms": 0.0, "allreduce_ms": 0.0}
    for _ in range(args.steps):
        # H2D of rhokap (as after every setupThermalCoeff) + zero + transport + all-reduce + D2H of jmeanGLOBAL
        _, st = t.run_optics(rk, cfg["albedo"], cfg["hgg"], packets, SEED, flags=cfg["flags"], out=jm)
        for k in e2e_parts:
            e2e_parts[k] += st[k] / args.steps
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e2e_s = torch.tensor([time.perf_counter() - e0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = total_packets / float(e2e_s.item())
    clocks = sampler.stop() if sampler else None      # sampled across both timed regions (device-resident + e2e)
    jm_sum = float(jm.sum())

    # ---- context numbers (rank 0, single GPU semantics)
    peak, peak_src = measured_hbm_peak()
    k_ms = res["local_kernel_ms"] / args.steps
    achieved = BYTES_PER_VOXEL_STEP * (res["local_voxel_steps"] / args.steps) / (k_ms * 1e-3) / 1e9
    form = t.get_option("form")
    kernel_name = FORMS.get(form, "k_transport_persistent")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(name + ":" + kernel_name) or ncu_traffic(name), "peak_source": peak_src, "kernel": kernel_name,
                "kernel_ms": k_ms, "algorithmic_bytes_per_voxel_step": BYTES_PER_VOXEL_STEP,
                "note": "random walk over an L2-resident region with fp64 atomics: instruction-issue / L1TEX-address bound, not HBM "
                        "bound (DESIGN.md section 3); frac is the algorithmic 16 B per voxel-step against the HBM copy peak"}
    if rank == 0 and not (cfg["flags"] & 1):
        try:
            vs_rate = res["local_voxel_steps"] / args.steps / (k_ms * 1e-3)
            t.set_option("probe_form", 0)
            t.roofline_probe(packets, SEED)
            pms, psteps = t.roofline_probe(packets, SEED)
            roofline["probe"] = {"voxel_steps_per_s": psteps / (pms * 1e-3), "ms": pms,
                                 "what": "L2-atomic / grid-lookup roofline of the step-by-step tally: same address stream (column under the beam, "
                                         "geometric step count), one fp64 load of rhokap + one fp64 RED into jmean per voxel-step, no transport arithmetic"}
            roofline["frac_of_probe"] = vs_rate / roofline["probe"]["voxel_steps_per_s"]
            if form in (5, 6, 7, 8):
                t.set_option("probe_form", 1)
                t.roofline_probe(packets, SEED)
                pms, psteps = t.roofline_probe(packets, SEED)
                roofline["probe_column"] = {"voxel_steps_per_s": psteps / (pms * 1e-3), "ms": pms,
                                            "what": "the same for the column form the kernel uses: per packet one 256-bit load of the z-fastest opacity copy per "
                                                    "four voxels + one fp64 RED + at most one u32 RED (forms 7, 8: the top planes in the same shared-memory "
                                                    "tiles, same launch shape), incl. the gather and finish kernels, no transport arithmetic"}
                roofline["frac_of_probe_column"] = vs_rate / roofline["probe_column"]["voxel_steps_per_s"]
            t.set_option("probe_form", -1)
        except Exception as e:  # the probe is context, never fatal
            roofline["probe_error"] = str(e)

    also = None
    if not args.no_also and name != "skin200":
        try:
            c2 = tamc.configs.CONFIGS["skin200"]
            t2 = tamc.MCTransport(c2["n"], c2["n"], c2["n"], c2["xmax"], c2["ymax"], c2["zmax"], device=local)
            for kv in args.option:
                k, v = kv.split("=")
                t2.set_option(k, int(v))
            t2.set_optics(c2["rhokap"](), c2["albedo"], c2["hgg"], flags=c2["flags"])
            if world > 1:
                t2.set_option("reduce", 0)
            p2 = 8_000_000
            s2 = torch.cuda.ExternalStream(t2.stream, device=dev)
            t2.run_async(p2, SEED); t2.sync()
            r2 = timed_steps(t2, s2, p2, 3, world, dist, dev, torch)
            also = {"skin200": {"workload": workload_desc("skin200", c2, p2), "packets_per_s": p2 * world * 3 / (r2["dev_ms"] * 1e-3),
                                "voxel_steps_per_s": r2["voxel_steps"] / (r2["dev_ms"] * 1e-3),
                                "scatters_per_packet": r2["scatters"] / (p2 * world * 3), "ms_per_step": r2["dev_ms"] / 3,
                                "note": "BASELINE config 3 physics (layered skin, albedo 0.98), reduced packet count, no all-reduce"}}
            t2.close()
        except Exception as e:
            also = {"error": str(e)}

    # the first "next" row: device-resident coupled loop with the reference's heat step
    if rank == 0 and not args.no_also:
        try:
            c5 = tamc.configs.CONFIGS["shipped80"]
            t5 = tamc.MCTransport(80, 80, 80, c5["xmax"], c5["ymax"], c5["zmax"], device=local)
            t5.set_optics(c5["rhokap"](), c5["albedo"], c5["hgg"], flags=0)
            t5.heat_init()
            t5.coupled_loop(125000, SEED, 200)
            w0 = time.perf_counter()
            it5, pk5 = t5.coupled_loop(125000, SEED, 2000)
            w5 = time.perf_counter() - w0
            also = dict(also or {})
            also["coupled_loop_shipped80"] = {
                "us_per_iteration": 1e6 * w5 / it5, "packets_per_s": pk5 / w5, "iterations": it5,
                "what": "mcpolar.f90:148-186 resident on the device: MC call of 125 000 packets + heat_sim_3d + arrhenius + setupThermalCoeff per iteration, no PCIe copy (host clock)"}
            t5.close()
        except Exception as e:
            also = dict(also or {})
            also["coupled_loop_shipped80"] = {"error": str(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fast_dir, fast = _oracle_lib_dir()
        threads = cpu_threads()
        per_rank, _ = cpu_calibrate(cfg, rk, fast_dir, args.cpu_seconds, threads)
        r = cpu_run(cfg, rk, threads, per_rank, fast_dir)
        cpu = {"value": r["stats"]["packets"] / r["seconds"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"{threads} emulated MPI ranks x {per_rank} packets of the same workload, ran2 streams, incl. jmean sum; {r['seconds']:.1f} s",
               "voxel_steps_per_s": r["stats"]["voxel_steps"] / r["seconds"],
               "flags": "-O3 -march=native -flto" if fast else "-O2 -ffp-contract=off",
               "note": "C restatement of the Fortran/MPI CPU path (no Fortran compiler in the image)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["dev_ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_desc(name, cfg, packets), "grid": f"{n}^3", "packets_per_gpu_per_step": packets,
                       "parallelism": f"packets partitioned over {world} GPU(s), one Philox stream per packet, one ncclAllReduce(jmean) per step",
                       "l2": "flushed between timed steps (256 MiB device fill outside the timed events)",
                       "rng": "Philox4x32-10, key=seed, counter=(packet id, event)",
                       "options": {k: t.get_option(k) for k in ("variant", "block", "ctas_per_sm", "chunk", "scatter_min", "merge", "min_ctas", "tile", "column")}},
            "voxel_steps_per_s": vsteps_per_s,
            "voxel_steps_per_packet": res["voxel_steps"] / total_packets,
            "breakdown_ms_per_step": {"kernel": res["kernel_ms"] / args.steps, "allreduce": res["allreduce_ms"] / args.steps,
                                      "wall_incl_l2_flush": res["wall_ms"] / args.steps,
                                      "allreduce_planes_of_box": int(t.get_option("reduce_planes"))},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(rk.nbytes), "d2h_bytes_per_step": int(jm.nbytes),
                    "ms_per_step": 1e3 * float(e2e_s.item()) / args.steps, "parts_ms": e2e_parts,
                    "api": "tamc_run_optics(host rhokap -> host jmeanGLOBAL) = tamc_set_optics + tamc_run, pinned host arrays; io_form %d "
                           "(bit0: jmeanGLOBAL written as zero fill beside the kernels + the beam's columns, bit1: the beam's columns of "
                           "rhokap uploaded ahead of the full grid, bit2: ... and only down to the depth the previous call's packets "
                           "reached + margin, deeper planes read from the caller's array on demand; the full grid and the zero fill "
                           "still cross PCIe inside the call); parts_ms h2d/d2h time only the copies not hidden behind the transport" % t.get_option("io_form"),
                    "jmean_sum_per_packet": jm_sum / (packets * world)},
            "gpu_launches": int(res["launches"]), "clocks": clocks, "also": also,
        }
        print(json.dumps(line), flush=True)
    t.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    import tamc

    if args.workload not in tamc.configs.CONFIGS:
        raise SystemExit(f"unknown workload {args.workload}; choose from {sorted(tamc.configs.CONFIGS)}")
    cfg = tamc.configs.CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg, args.workload)
    else:
        run_ours(args, cfg, args.workload)


if __name__ == "__main__":
    main()
