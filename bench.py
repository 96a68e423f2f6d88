#!/usr/bin/env python
"""bench.py -- photon packets/s and voxel-steps/s of the MC transport hot path on N B200s.

A "step" is one MC call (the replacement of /root/reference/src/mcpolar.f90:151-173): clear the
tally, transport the packets, all-reduce the tally over the ranks.

Headline workload = BASELINE.json configs[2], the configuration the north-star target is quoted on:
layered skin 200^3, scatter loop on (albedo 0.98, g 0.9), 1e9 packets per step IN TOTAL, partitioned
by packet id over the N ranks (strong scaling), one ncclAllReduce of the 64 MB tally inside the
timed region of every step.  The other BASELINE configs (homog200 = configs[1], phantom400 =
configs[3], the coupled loop = configs[4]) are measured at reduced size under `also`.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference ...                     the reference's CPU path (oracle port, all host threads)

One JSON line on stdout (rank 0).  `value` times the device-resident call (inputs already in HBM);
`e2e` times tamc_run_optics (= tamc_set_optics + tamc_run) with pinned HOST buffers, copies inside
the timed region, the opacity grid changing from call to call.
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

# NCCL plumbing, before anything initialises NCCL in this process (libtamc sets the same defaults, tamc_api.cu nccl_api()):
# with NVLS / cuMem-backed communicator buffers the one-CTA-per-SM stub-regime kernels measured 10-30 % slower on 2 x B200
os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
os.environ.setdefault("NCCL_CUMEM_ENABLE", "0")

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
         4: "k_transport_stub_tiled", 5: "k_transport_column", 6: "k_transport_column", 7: "k_transport_column_tiled",
         8: "k_transport_column_parked", 9: "k_transport_flight"}


T0 = time.time()


def log(msg):
    """progress on stderr (rank 0 only): where the run is when something takes long"""
    if os.environ.get("RANK", "0") == "0":
        print(f"[bench {time.time() - T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="skin200")
    ap.add_argument("--packets", type=int, default=0, help="packets per step IN TOTAL (default: the config's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary measurements")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--also", default="homog200,phantom400,coupled_calls_shipped80,resident_loop_shipped80",
                    help="comma-separated secondary measurements to run")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--option", action="append", default=[], help="name=value passed to tamc_set_option")
    return ap.parse_args()


def workload_desc(name, cfg, total):
    """The same string in both arms (the driver compares the two lines' config)."""
    return (f"{name}: {cfg['n']}^3 voxels, xmax/ymax/zmax={cfg['xmax']}/{cfg['ymax']}/{cfg['zmax']} cm, "
            f"albedo={cfg['albedo']:.4g}, hgg={cfg['hgg']}, scatter={'on' if cfg['flags'] & 1 else 'off (shipped stub)'}, "
            f"{total:.3g} packets per step in total")


def config_block(name, cfg, total):
    return {"workload": workload_desc(name, cfg, total), "grid": f"{cfg['n']}^3", "packets_per_step": total,
            "partition": "packet ids split evenly over the ranks; one all-reduce of the jmean grid per step",
            "l2": "device arm: L2 flushed between timed steps (256 MiB device fill outside the timed events)"}


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
    probe = 5000 if cfg["flags"] & 1 else 200000
    r = cpu_run(cfg, rk, 1, probe, fast_dir)
    rate = probe / max(r["seconds"], 1e-6)
    return max(500, int(rate * seconds * 0.8)), rate


def run_reference(args, cfg, name):
    """--impl reference: the reference's own CPU implementation of the path.  The Fortran cannot be
    built in this image (no Fortran front-end, no MPI), so this is the C restatement in oracle/,
    R emulated MPI ranks on all host threads + the in-memory jmean sum (kind = "port").  Each step is
    a bounded sample of the workload (about 4 s of all host threads)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    total = args.packets or cfg["nphotons"]
    rk = cfg["rhokap"]()
    fast_dir, fast = _oracle_lib_dir()
    threads = cpu_threads()
    per_rank, _ = cpu_calibrate(cfg, rk, fast_dir, 4.0, threads)
    for _ in range(args.warmup):
        cpu_run(cfg, rk, threads, max(500, per_rank // 8), fast_dir)
    tot_s, tot_p, tot_v = 0.0, 0, 0
    for _ in range(args.steps):
        r = cpu_run(cfg, rk, threads, per_rank, fast_dir)
        tot_s += r["seconds"]
        tot_p += r["stats"]["packets"]
        tot_v += r["stats"]["voxel_steps"]
    value = tot_p / tot_s
    sample = (f"{threads} emulated MPI ranks x {per_rank} packets per step of the same workload "
              f"(ran2 seeds per mcpolar.f90:97-98) + jmean sum; {tot_s:.1f} s over {args.steps} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(name, cfg, total),
        "voxel_steps_per_s": tot_v / tot_s,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "flags": "-O3 -march=native -flto" if fast else "-O2 -ffp-contract=off",
                         "note": "C restatement of the Fortran/MPI CPU path (no Fortran compiler in the image)"},
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
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
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
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); pw.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw))
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


def ncu_traffic(name, voxel_steps_per_launch=None):
    """DRAM bytes per launch of the transport kernel from the committed ncu --set full capture.  A capture taken at a
    smaller packet count than the bench launch is scaled by voxel-steps (the traffic is per voxel visited)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            v = json.load(f).get(name)
    except Exception:
        return None
    if isinstance(v, dict):
        if voxel_steps_per_launch:
            return v["dram_bytes"] * voxel_steps_per_launch / v["voxel_steps"]
        return v["dram_bytes"]
    return v


class Ctx:
    """torch / torch.distributed plumbing of one rank."""

    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce_max(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def reduce_sum(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()


PEER_VARIANT = os.environ.get("TAMC_BENCH_PEER", "0") != "0"      # also.homog200 at N > 1: time the peer-memory box reduce as well


def make_transport(ctx, cfg, rk, options, root_io=0):
    import tamc
    from tamc import dist as tdist

    n = cfg["n"]
    t = tamc.MCTransport(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"], device=ctx.local)
    for kv in options:
        k, v = kv.split("=")
        t.set_option(k, int(v))
    if root_io:
        t.set_option("root_io", 1)                # before comm_init: it shapes the collectives of every call
    t.set_optics(rk, cfg["albedo"], cfg["hgg"], flags=cfg["flags"])
    if ctx.world > 1:
        t.comm_init(ctx.world, ctx.rank, tdist.broadcast_unique_id(tamc.comm_unique_id, ctx.dist, ctx.dev))
    return t


def timed_steps(ctx, t, per_rank, steps, flush=True):
    """K MC calls with inputs resident; per-step CUDA events on the library's stream; max over ranks."""
    torch = ctx.torch
    stream = torch.cuda.ExternalStream(t.stream, device=ctx.dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    sums = {"kernel_ms": 0.0, "allreduce_ms": 0.0, "zero_ms": 0.0, "voxel_steps": 0, "scatters": 0, "launches": 0}
    ctx.barrier()
    w0 = time.perf_counter()
    for a, b in ev:
        if flush:
            t.flush_l2(256 << 20)                 # evict L2 between steps, outside the timed events
        a.record(stream)
        t.run_async(per_rank, SEED)               # ids from the cursor: fresh packets every step, rank r its own slice
        b.record(stream)
        st = t.get_stats()                        # syncs the stream; per-step counters and device times
        for k in ("kernel_ms", "allreduce_ms", "zero_ms"):
            sums[k] += st[k]
        sums["voxel_steps"] += st["voxel_steps"]
        sums["scatters"] += st["scatters"]
        sums["launches"] += st["gpu_launches"]
    ctx.barrier()
    wall = time.perf_counter() - w0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    red = ctx.reduce_max([dev_ms, sums["kernel_ms"], sums["allreduce_ms"], wall * 1e3])
    tot = ctx.reduce_sum([sums["voxel_steps"], sums["scatters"], sums["launches"]])
    per_rank = [0.0] * ctx.world
    per_rank[ctx.rank] = sums["kernel_ms"] / steps
    per_rank = ctx.reduce_sum(per_rank)
    return {"dev_ms": red[0], "kernel_ms": red[1], "allreduce_ms": red[2], "wall_ms": red[3], "kernel_ms_by_rank": per_rank,
            "voxel_steps": tot[0], "scatters": tot[1], "launches": int(tot[2]),
            "local_kernel_ms": sums["kernel_ms"], "local_voxel_steps": sums["voxel_steps"]}


def timed_e2e(ctx, t, cfg, grids, jm, per_rank, steps):
    """K reference-facing calls with HOST buffers: H2D of rhokap (a different grid every call, as after every
    setupThermalCoeff), clear, transport, all-reduce, D2H of jmeanGLOBAL.  Host clock, max over ranks."""
    parts = {"h2d_ms": 0.0, "d2h_ms": 0.0, "kernel_ms": 0.0, "allreduce_ms": 0.0}
    for i in range(2):
        t.run_optics(grids[i % len(grids)], cfg["albedo"], cfg["hgg"], per_rank, SEED, flags=cfg["flags"], out=jm)
    ctx.barrier()
    e0 = time.perf_counter()
    per_call = []
    for i in range(steps):
        c0 = time.perf_counter()
        _, st = t.run_optics(grids[i % len(grids)], cfg["albedo"], cfg["hgg"], per_rank, SEED, flags=cfg["flags"], out=jm)
        per_call.append(time.perf_counter() - c0)
        for k in parts:
            parts[k] += st[k] / steps
    ctx.barrier()
    secs = ctx.reduce_max([time.perf_counter() - e0])[0]
    return secs, parts, per_call


def crater_variant(cfg, rk, radius_vox, depth_vox):
    """A second opacity grid for the e2e loop: an ablated cylinder (rhokap = 0, 3dFD.f90:334-353) under the beam."""
    import numpy as np

    n = cfg["n"]
    out = rk.copy(order="F")
    ii, jj = np.meshgrid(np.arange(1, n + 1), np.arange(1, n + 1), indexing="ij")
    r = np.hypot(ii - 0.5 - n / 2.0, jj - 0.5 - n / 2.0)
    for k in range(n, max(0, n - depth_vox), -1):
        out[1:-1, 1:-1, k][r <= radius_vox] = 0.0
    return out


def parity_check(ctx, t, cfg):
    """Multi-rank correctness inside the bench (the driver's GPU-test box has one GPU): the all-reduced grid of the N ranks,
    each running its own slice of a small id range, against rank 0 running the whole range alone (no reduce)."""
    import numpy as np

    per = 100_000 if cfg["flags"] & 1 else 1_000_000
    base = 7_000_000_000                          # ids far from the ones the timed steps use
    t.seek(base)
    t.run_async(per, SEED)                        # rank r: [base + r*per, base + (r+1)*per), then ncclAllReduce
    red = t.get_jmean()
    st = t.get_stats()
    tot = ctx.reduce_sum([st["packets"], st["voxel_steps"], st["scatters"]])
    out = {"what": f"all-reduced grid of {ctx.world} rank(s) x {per} packets == one rank running the same {ctx.world * per} ids without a reduce",
           "ranks": ctx.world}
    if ctx.rank == 0:
        t.set_option("reduce", 0)
        t.run_async(per * ctx.world, SEED, base)
        one = t.get_jmean()
        s1 = t.get_stats()
        t.set_option("reduce", 1)
        nz = one != 0
        same_support = bool(np.array_equal(red != 0, nz))
        rel = float(np.max(np.abs(red[nz] - one[nz]) / np.abs(one[nz]))) if nz.any() else 0.0
        counters = (int(tot[0]) == s1["packets"] and int(tot[1]) == s1["voxel_steps"] and int(tot[2]) == s1["scatters"])
        out.update(max_rel_diff=rel, same_support=same_support, counters_equal=counters,
                   ok=bool(same_support and counters and rel < 1e-9),
                   status="ok" if (same_support and counters and rel < 1e-9) else "FAILED")
    ctx.barrier()
    t.seek(0)
    return out


def roofline_block(ctx, t, name, res, steps, per_rank):
    peak, peak_src = measured_hbm_peak()
    k_ms = res["local_kernel_ms"] / steps
    vs_rate = res["local_voxel_steps"] / steps / (k_ms * 1e-3)
    achieved = BYTES_PER_VOXEL_STEP * vs_rate / 1e9
    form = t.get_option("form")
    kernel_name = FORMS.get(form, "k_transport_persistent")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(name + ":" + kernel_name, res["local_voxel_steps"] / steps) or ncu_traffic(name),
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture of this kernel "
                                "(profiles/ncu_traffic.json), scaled by voxel-steps to this launch", "peak_source": peak_src,
                "kernel": kernel_name, "kernel_ms": k_ms, "algorithmic_bytes_per_voxel_step": BYTES_PER_VOXEL_STEP,
                "voxel_steps_per_launch": res["local_voxel_steps"] / steps,
                "bound_ncu": "instruction issue / L1TEX (the grids' hot region is L2-resident, DRAM < 1 % busy): see profiles/",
                "note": "frac is the contractual number: algorithmic 16 B per voxel-step against the measured HBM copy peak; "
                        "frac_of_probe compares with the measured grid-lookup / L2-atomic roofline of the same address stream"}
    if ctx.rank == 0:
        try:
            pr = t.trace_probe(min(per_rank, 2_000_000), SEED)
            roofline["probe"] = pr
            roofline["frac_of_probe"] = vs_rate / pr["voxel_steps_per_s"]
        except Exception as e:  # the probe is context, never fatal
            roofline["probe_error"] = str(e)
    return roofline


def also_homog200(ctx, options, steps=10):
    """BASELINE configs[1] (the shipped stub regime at 200^3): 1e8 packets PER GPU per step (weak), every rank its own ids,
    box all-reduce; device-timed and end to end with the opacity grid alternating between two grids."""
    import tamc

    c = tamc.configs.CONFIGS["homog200"]
    rk = c["rhokap"]()
    rk_b = crater_variant(c, rk, 20, 3)
    tamc.pin_host(rk); tamc.pin_host(rk_b)
    t = make_transport(ctx, c, rk, options)
    per = 100_000_000
    for _ in range(3):
        t.run_async(per, SEED); t.sync()
    log("also.homog200: warm")
    r = timed_steps(ctx, t, per, steps)
    log("also.homog200: device-resident steps timed")
    jm = t.new_jmean()
    tamc.pin_host(jm)
    secs, parts, per_call = timed_e2e(ctx, t, c, [rk, rk_b], jm, per, steps)
    log("also.homog200: e2e timed")
    out = {"workload": workload_desc("homog200", c, per * ctx.world) + " (weak: 1e8 per GPU)",
           "packets_per_s": per * ctx.world * steps / (r["dev_ms"] * 1e-3),
           "voxel_steps_per_s": r["voxel_steps"] / (r["dev_ms"] * 1e-3), "ms_per_step": r["dev_ms"] / steps,
           "kernel_ms": r["kernel_ms"] / steps, "kernel_ms_by_rank": r["kernel_ms_by_rank"], "allreduce_ms": r["allreduce_ms"] / steps,
           "kernel": FORMS.get(t.get_option("form")), "allreduce_planes_of_box": int(t.get_option("reduce_planes")),
           "e2e_packets_per_s": per * ctx.world * steps / secs, "e2e_ms_per_step": 1e3 * secs / steps,
           "e2e_ms_by_grid": {"uniform": 1e3 * statistics.mean(per_call[0::2]), "crater": 1e3 * statistics.mean(per_call[1::2])},
           "e2e_parts_ms": parts, "io_form": int(t.get_option("io_form")),
           "note": "e2e alternates two opacity grids (uniform / ablated crater under the beam), so the depth-limited upload cannot coast"}
    if ctx.world > 1:
        # the same end-to-end loop with only rank 0 moving host arrays ("root_io": rhokap broadcast over NVLink, one download)
        t2 = make_transport(ctx, c, rk, options, root_io=1)
        for _ in range(2):
            t2.run_async(per, SEED); t2.sync()
        secs2, parts2, _ = timed_e2e(ctx, t2, c, [rk, rk_b], jm, per, steps)
        out["e2e_root_io_packets_per_s"] = per * ctx.world * steps / secs2
        out["e2e_root_io_ms_per_step"] = 1e3 * secs2 / steps
        out["e2e_root_io_parts_ms"] = parts2
        out["e2e_root_io_form"] = int(t2.get_option("io_form"))
        t2.close()
        log("also.homog200: root_io e2e timed")
    if ctx.world > 1 and PEER_VARIANT:
        # the same device-resident steps with the box summed out of peer memory by the library's own kernel
        # ("peer_reduce", tamc_peer.cuh) instead of by ncclAllReduce
        try:
            t3 = make_transport(ctx, c, rk, list(options) + ["peer_reduce=1"])
            for _ in range(3):
                t3.run_async(per, SEED); t3.sync()
            r3 = timed_steps(ctx, t3, per, steps)
            out["peer_reduce"] = {"ms_per_step": r3["dev_ms"] / steps, "packets_per_s": per * ctx.world * steps / (r3["dev_ms"] * 1e-3),
                                  "kernel_ms": r3["kernel_ms"] / steps, "kernel_ms_by_rank": r3["kernel_ms_by_rank"],
                                  "allreduce_ms": r3["allreduce_ms"] / steps, "peer_state": int(t3.get_option("peer_state")),
                                  "what": "tamc_set_option(peer_reduce, 1): pack + k_peer_box_reduce (every rank sums the ranks' boxes out of "
                                          "peer memory over NVLink, rank order) + unpack; peer_state 1 = in use, -1 = fell back to NCCL"}
            t3.close()
        except Exception as e:
            out["peer_reduce"] = {"error": str(e)}
        log("also.homog200: peer_reduce steps timed")
    if ctx.rank == 0:
        try:
            vs_rate = r["local_voxel_steps"] / steps / (r["local_kernel_ms"] / steps * 1e-3)
            t.set_optics(rk, c["albedo"], c["hgg"], flags=c["flags"])      # the probe runs on the uniform grid the timed steps used
            t.set_option("probe_form", 1)
            t.roofline_probe(per, SEED)
            pms, psteps = t.roofline_probe(per, SEED)
            t.set_option("probe_form", -1)
            out["ratio_to_unregrouped_column_probe"] = vs_rate / (psteps / (pms * 1e-3))
            out["probe_column_ms"] = pms
            out["probe_note"] = ("k_probe_column issues the column form's address stream packet by packet WITHOUT the regrouped walk the "
                                 "transport uses, so it is a reference point, not a ceiling: a ratio above 1 is the gain of regrouping")
        except Exception as e:
            out["probe_error"] = str(e)
    ctx.barrier()
    t.close()
    tamc.unpin_host(rk); tamc.unpin_host(rk_b); tamc.unpin_host(jm)
    return out


def also_phantom400(ctx, options, steps=3):
    """BASELINE configs[3]: 400^3, albedo 0.999 -- grids (1 GB) beyond L2; 2e7 packets per GPU per step (weak), ids
    partitioned over the ranks, the 512 MB tally all-reduced inside the timed events."""
    import tamc

    c = tamc.configs.CONFIGS["phantom400"]
    rk = c["rhokap"]()
    t = make_transport(ctx, c, rk, options)
    del rk
    # per GPU (weak).  Packets live ~190 scatterings on average but the longest of a call ~1e4, so a call has a fixed tail
    # of ~9 ms on top of 12.9 ms per 1e6 packets (measured: 1e6 -> 21.8 ms, 2e6 -> 34.7 ms); SURVEY 8(d) config 4 asks for
    # a call of seconds, 2e7 packets give a quarter of one
    per = 20_000_000
    t.run_async(per, SEED); t.sync()
    r = timed_steps(ctx, t, per, steps)
    out = {"workload": workload_desc("phantom400", c, per * ctx.world) + " (weak: 2e7 per GPU)",
           "packets_per_s": per * ctx.world * steps / (r["dev_ms"] * 1e-3),
           "voxel_steps_per_s": r["voxel_steps"] / (r["dev_ms"] * 1e-3),
           "voxel_steps_per_s_per_gpu": r["voxel_steps"] / (r["dev_ms"] * 1e-3) / ctx.world,
           "scatters_per_packet": r["scatters"] / (per * ctx.world * steps), "ms_per_step": r["dev_ms"] / steps,
           "kernel_ms": r["kernel_ms"] / steps, "allreduce_ms": r["allreduce_ms"] / steps, "kernel": FORMS.get(t.get_option("form"))}
    peak, _ = measured_hbm_peak()
    out["roofline_frac_hbm"] = BYTES_PER_VOXEL_STEP * (r["local_voxel_steps"] / (r["local_kernel_ms"] * 1e-3)) / 1e9 / peak
    t.close()
    return out


def also_coupled_calls(ctx, options, calls=1000):
    """BASELINE configs[4] as SURVEY 8(d) config 5 specifies it: >= 1000 consecutive set_optics + run calls (tamc_run_optics)
    of 125 000 packets per rank on the shipped 80^3 geometry with a scripted crater that changes every call
    (mcpolar.f90:148-186 call pattern; 3dFD.f90:334-353 property update); latency per call on the host clock."""
    import numpy as np
    import tamc

    c = tamc.configs.CONFIGS["shipped80"]
    grids = list(tamc.configs.crater_sequence(80, 16))
    for g in grids:
        tamc.pin_host(g)
    per = 125_000
    jm = np.zeros((80, 80, 80), dtype=np.float64, order="F")
    tamc.pin_host(jm)

    def sequence(t):
        for i in range(20):
            t.run_optics(grids[i % 16], c["albedo"], c["hgg"], per, SEED, flags=0, out=jm)
        ctx.barrier()
        lat, parts = [], {"h2d_ms": 0.0, "kernel_ms": 0.0, "allreduce_ms": 0.0, "d2h_ms": 0.0, "zero_ms": 0.0}
        w0 = time.perf_counter()
        for i in range(calls):
            c0 = time.perf_counter()
            _, st = t.run_optics(grids[i % 16], c["albedo"], c["hgg"], per, SEED, flags=0, out=jm)
            lat.append(time.perf_counter() - c0)
            for k in parts:
                parts[k] += st[k] / calls
        ctx.barrier()
        wall = ctx.reduce_max([time.perf_counter() - w0])[0]
        return np.array(lat) * 1e6, parts, wall

    t = make_transport(ctx, c, grids[0], options)
    lat, parts, wall = sequence(t)
    out = {"calls": calls, "packets_per_call": per * ctx.world, "ranks": ctx.world,
           "mean_us": float(lat.mean()), "p50_us": float(np.percentile(lat, 50)), "p95_us": float(np.percentile(lat, 95)),
           "max_us": float(lat.max()), "packets_per_s": per * ctx.world * calls / wall,
           "breakdown_us": {k[:-3]: 1e3 * v for k, v in parts.items()},
           "what": "tamc_run_optics(host rhokap -> host jmeanGLOBAL) per call, 16 crater grids in rotation (a different grid every call), "
                   "4.4 MB up + 4.1 MB down per rank per call; breakdown = device events of the parts, the rest is launch + sync latency"}
    t.close()
    if ctx.world > 1:
        t2 = make_transport(ctx, c, grids[0], options, root_io=1)
        lat2, parts2, wall2 = sequence(t2)
        out["root_io"] = {"mean_us": float(lat2.mean()), "p50_us": float(np.percentile(lat2, 50)), "p95_us": float(np.percentile(lat2, 95)),
                          "packets_per_s": per * ctx.world * calls / wall2, "breakdown_us": {k[:-3]: 1e3 * v for k, v in parts2.items()},
                          "what": "the same sequence with only rank 0 moving host arrays (tamc_set_option root_io)"}
        t2.close()
    for g in grids:
        tamc.unpin_host(g)
    tamc.unpin_host(jm)
    return out


def also_resident_loop(ctx):
    import tamc

    c5 = tamc.configs.CONFIGS["shipped80"]
    t5 = tamc.MCTransport(80, 80, 80, c5["xmax"], c5["ymax"], c5["zmax"], device=ctx.local)
    t5.set_optics(c5["rhokap"](), c5["albedo"], c5["hgg"], flags=0)
    t5.heat_init()
    t5.coupled_loop(125000, SEED, 200)
    w0 = time.perf_counter()
    it5, pk5 = t5.coupled_loop(125000, SEED, 2000)
    w5 = time.perf_counter() - w0
    t5.close()
    return {"us_per_iteration": 1e6 * w5 / it5, "packets_per_s": pk5 / w5, "iterations": it5,
            "what": "mcpolar.f90:148-186 resident on the device: MC call of 125 000 packets + heat_sim_3d + arrhenius + "
                    "setupThermalCoeff per iteration, no PCIe copy (host clock, rank 0 only)"}


def run_ours(args, cfg, name):
    import torch

    import tamc

    if not torch.cuda.is_available() or tamc.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the transport has no CPU fallback (use --impl reference for the CPU path)")
    ctx = Ctx()
    rank, world = ctx.rank, ctx.world
    if world != args.gpus and rank == 0:
        print(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}; using WORLD_SIZE", file=sys.stderr)

    total = args.packets or cfg["nphotons"]
    per_rank = total // world
    total = per_rank * world
    rk = cfg["rhokap"]()
    rk_b = crater_variant(cfg, rk, 4, 3)
    tamc.pin_host(rk); tamc.pin_host(rk_b)
    t = make_transport(ctx, cfg, rk, args.option)

    try:
        gpu_id = str(torch.cuda.get_device_properties(ctx.local).uuid)
        if not gpu_id.startswith("GPU-"):
            gpu_id = "GPU-" + gpu_id
    except Exception:
        gpu_id = str(ctx.local)

    log(f"set-up done: {name}, {total:.3g} packets per step over {world} rank(s)")
    parity = parity_check(ctx, t, cfg)
    log(f"parity check: {parity.get('status')}")

    # ---- warm-up, then the timed device-resident steps
    for _ in range(max(args.warmup, 0)):
        t.run_async(per_rank, SEED)
        t.sync()
    sampler = ClockSampler(gpu_id) if rank == 0 else None
    res = timed_steps(ctx, t, per_rank, args.steps)
    value = float(total) * args.steps / (res["dev_ms"] * 1e-3)
    log(f"timed steps done: {value:.4g} packets/s")
    vsteps_per_s = res["voxel_steps"] / (res["dev_ms"] * 1e-3)

    # ---- e2e: the reference-facing call with host buffers (upload rhokap, run, download jmean)
    e2e = None
    if not args.no_e2e:
        jm = t.new_jmean()
        tamc.pin_host(jm)
        secs, parts, _ = timed_e2e(ctx, t, cfg, [rk, rk_b], jm, per_rank, args.steps)
        e2e = {"value": float(total) * args.steps / secs, "unit": UNIT, "h2d_bytes_per_step": int(rk.nbytes) * world,
               "d2h_bytes_per_step": int(jm.nbytes) * world, "ms_per_step": 1e3 * secs / args.steps, "parts_ms": parts,
               "api": "tamc_run_optics(host rhokap -> host jmeanGLOBAL) = tamc_set_optics + tamc_run on every rank, pinned host arrays, "
                      "two opacity grids alternating call by call (layered skin / the same with an ablated crater); host clock "
                      "around the K calls, barrier on both sides, max over ranks; bytes are the sum over ranks",
               "jmean_sum_per_packet": float(jm.sum()) / total}
        if world > 1:
            t_root = make_transport(ctx, cfg, rk, args.option, root_io=1)
            t_root.run_async(min(per_rank, 1_000_000), SEED); t_root.sync()
            secs_r, parts_r, _ = timed_e2e(ctx, t_root, cfg, [rk, rk_b], jm, per_rank, args.steps)
            e2e["root_io"] = {"value": float(total) * args.steps / secs_r, "ms_per_step": 1e3 * secs_r / args.steps, "parts_ms": parts_r,
                              "h2d_bytes_per_step": int(rk.nbytes), "d2h_bytes_per_step": int(jm.nbytes),
                              "what": "the same loop with tamc_set_option(root_io, 1): only rank 0's rhokap is uploaded (ncclBroadcast to the "
                                      "other GPUs) and only rank 0's jmeanGLOBAL is written"}
            t_root.close()
    clocks = sampler.stop() if sampler else None      # sampled across both timed regions (device-resident + e2e)
    log("e2e done")

    roofline = roofline_block(ctx, t, name, res, args.steps, per_rank)
    log(f"roofline probe done: {roofline.get('frac_of_probe', roofline.get('probe_error'))}")
    ctx.barrier()

    also = {}
    want_also = set(args.also.split(","))
    if not args.no_also:
        for key, fn in (("homog200", lambda: also_homog200(ctx, args.option)),
                        ("phantom400", lambda: also_phantom400(ctx, args.option)),
                        ("coupled_calls_shipped80", lambda: also_coupled_calls(ctx, args.option))):
            if key not in want_also:
                continue
            try:
                also[key] = fn()
            except Exception as e:
                also[key] = {"error": f"{type(e).__name__}: {e}"}
            log(f"also.{key} done: {str(also[key])[:300]}")
            ctx.barrier()
        if rank == 0 and "resident_loop_shipped80" in want_also:
            try:
                also["resident_loop_shipped80"] = also_resident_loop(ctx)
            except Exception as e:
                also["resident_loop_shipped80"] = {"error": str(e)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fast_dir, fast = _oracle_lib_dir()
        threads = cpu_threads()
        per_cpu, _ = cpu_calibrate(cfg, rk, fast_dir, args.cpu_seconds, threads)
        r = cpu_run(cfg, rk, threads, per_cpu, fast_dir)
        cpu = {"value": r["stats"]["packets"] / r["seconds"], "unit": UNIT, "cores": r["threads"], "kind": "port",
               "sample": f"{threads} emulated MPI ranks x {per_cpu} packets of the same workload, ran2 streams, incl. jmean sum; {r['seconds']:.1f} s",
               "voxel_steps_per_s": r["stats"]["voxel_steps"] / r["seconds"],
               "flags": "-O3 -march=native -flto" if fast else "-O2 -ffp-contract=off",
               "note": "C restatement of the Fortran/MPI CPU path (no Fortran compiler in the image)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["dev_ms"] / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_block(name, cfg, total),
            "run": {"packets_per_gpu_per_step": per_rank,
                    "parallelism": f"packet ids partitioned over {world} GPU(s), one Philox stream per packet, one ncclAllReduce(jmean, "
                                   f"{cfg['n'] ** 3 * 8 / 1e6:.0f} MB) per step inside the timed events",
                    "l2": "flushed between timed steps (256 MiB device fill outside the timed events)",
                    "rng": "Philox4x32-10, key=seed, counter=(packet id, event)",
                    "options": {k: t.get_option(k) for k in ("variant", "block", "ctas_per_sm", "chunk", "scatter_min", "merge", "min_ctas", "tile", "column")}},
            "voxel_steps_per_s": vsteps_per_s,
            "voxel_steps_per_packet": res["voxel_steps"] / (float(total) * args.steps),
            "scatters_per_packet": res["scatters"] / (float(total) * args.steps),
            "breakdown_ms_per_step": {"kernel": res["kernel_ms"] / args.steps, "kernel_by_rank": res["kernel_ms_by_rank"],
                                      "allreduce": res["allreduce_ms"] / args.steps,
                                      "wall_incl_l2_flush": res["wall_ms"] / args.steps},
            "parity_check": parity,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(res["launches"]), "clocks": clocks, "also": also or None,
        }
        print(json.dumps(line), flush=True)
    t.close()
    if world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


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
