"""Two GPUs, one process each (skipped on a single-GPU box): the NCCL all-reduce inside tamc_run gives
every rank the sum of the per-rank tallies == one GPU running all the ids."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tissue-ablation-mc_b200")]
    import torch
    import torch.distributed as dist

    import tamc
    from tamc import dist as tdist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = tamc.configs.scaled("skin200", 64)
    t = tamc.MCTransport(64, 64, 64, cfg["xmax"], cfg["ymax"], cfg["zmax"], device=rank)
    t.set_optics(cfg["rhokap"](), cfg["albedo"], cfg["hgg"], flags=cfg["flags"])
    t.comm_init(world, rank, tdist.broadcast_unique_id(tamc.comm_unique_id, dist))
    jm1, st1 = t.run(20000, 99)          # ids [rank*20000, (rank+1)*20000), cursor -> 40000
    jm2, st2 = t.run(20000, 99)          # ids 40000 + ...
    assert st1["packets"] == 20000 and st1["allreduce_ms"] > 0
    np.save(f"{out}.{rank}.npy", np.stack([jm1, jm2]))
    dist.barrier()
    t.close()
    dist.destroy_process_group()


def test_allreduce_equals_single_gpu(tmp_path):
    import torch

    import tamc
    from tests.util import compare_grids, voxel_tau

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    out = str(tmp_path / "jm")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = np.load(out + ".0.npy"), np.load(out + ".1.npy")
    assert np.array_equal(a, b)                          # every rank holds the same reduced grid
    cfg = tamc.configs.scaled("skin200", 64)
    t = tamc.MCTransport(64, 64, 64, cfg["xmax"], cfg["ymax"], cfg["zmax"], device=0)
    rk = cfg["rhokap"]()
    t.set_optics(rk, cfg["albedo"], cfg["hgg"], flags=cfg["flags"])
    for call in range(2):
        t.run_async(40000, 99, call * 40000)
        compare_grids(a[call], t.get_jmean(), rtol=1e-9, dep_scale=voxel_tau(cfg, rk))
    t.close()


def _coupled_worker(rank, world, port, out):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tissue-ablation-mc_b200")]
    import torch
    import torch.distributed as dist

    import tamc
    from tamc import dist as tdist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    n = 24
    t = tamc.MCTransport(n, n, n, 0.03, 0.03, 0.06, device=rank)
    t.set_optics(tamc.gridset(0.03, 0.03, 0.06, n, n, n, 680.0)[3], 0.0, 0.9)
    t.comm_init(world, rank, tdist.broadcast_unique_id(tamc.comm_unique_id, dist))
    t.heat_init(pulsetype="tophat", power=20.0, energyPerPixel=4000.0, ablateTemp=150.0, loops=2)
    it, pk = t.coupled_loop(10000, 5, 12)             # 10 000 packets per rank per call, tally all-reduced
    assert it == 12 and pk == 12 * 10000 * world
    np.save(f"{out}.{rank}.npy", np.stack([t.heat_array("temp"), t.heat_array("rhokap")]))
    dist.barrier()
    t.close()
    dist.destroy_process_group()


def test_coupled_loop_two_ranks_equals_one_rank_with_all_packets(tmp_path):
    """Every rank repeats the (deterministic) heat step on the all-reduced tally: both replicas stay identical and
    equal one GPU that runs all the packets itself."""
    import torch

    import tamc

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    out = str(tmp_path / "heat")
    mp.spawn(_coupled_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = np.load(out + ".0.npy"), np.load(out + ".1.npy")
    assert np.array_equal(a, b)
    n = 24
    t = tamc.MCTransport(n, n, n, 0.03, 0.03, 0.06, device=0)
    t.set_optics(tamc.gridset(0.03, 0.03, 0.06, n, n, n, 680.0)[3], 0.0, 0.9)
    t.heat_init(pulsetype="tophat", power=20.0, energyPerPixel=4000.0, ablateTemp=150.0, loops=2)
    t.coupled_loop(20000, 5, 12)                      # same ids 0..20000 per call on one GPU
    assert np.allclose(t.heat_array("temp"), a[0], rtol=1e-9, atol=0)
    assert np.array_equal(t.heat_array("rhokap") == 0, a[1] == 0)
    t.close()


def _stub_worker(rank, world, port, out, peer=0):
    import sys
    import time

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "tissue-ablation-mc_b200")]
    import torch
    import torch.distributed as dist

    import tamc
    from tamc import dist as tdist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = tamc.configs.CONFIGS["shipped80"]
    t = tamc.MCTransport(80, 80, 80, cfg["xmax"], cfg["ymax"], cfg["zmax"], device=rank)
    t.set_optics(cfg["rhokap"](), cfg["albedo"], cfg["hgg"], flags=0)
    t.set_option("peer_reduce", peer)                 # the box all-reduce out of peer memory instead of by NCCL (tamc_peer.cuh)
    t.comm_init(world, rank, tdist.broadcast_unique_id(tamc.comm_unique_id, dist))
    grids = []
    for column, box, bound in ((0, 0, 1), (0, -1, 1), (1, -1, 1), (1, 1, 1), (2, 0, 1), (1, -1, 0)):
        t.set_option("column", column)
        t.set_option("box_reduce", box)
        t.set_option("reduce_bound", bound)
        t.seek(0)
        jm, st = t.run(60000, 7)
        assert st["allreduce_ms"] > 0
        # the column form on the z-fastest copy derives how deep a packet can get (tau <= 33 ln 2; 1.02 per voxel here:
        # 23 voxels + one spare) and the all-reduce moves only those planes of the box
        # ... for EVERY kernel form (the bound comes from the gathered copy or from the resident grid): the element count of
        # the all-reduce must not depend on the form a rank happened to run
        want = 0 if box == 0 else (24 if bound else 80)
        assert t.get_option("reduce_planes") == want, (column, box, bound, t.get_option("reduce_planes"))
        grids.append(jm.copy())
    # ranks that differ in packet count, straddling the threshold of the column form (narrow beam: 8 x 148 x columns x tile
    # planes = 4.1e6 packets): rank 0 takes the column form,
    # rank 1 the step-by-step kernel, and both must pass NCCL the same count (explicit id ranges, no overlap)
    for name, value in (("column", -1), ("box_reduce", -1), ("reduce_bound", 1)):
        t.set_option(name, value)
    n0, n1 = 4_400_000, 1_000_000
    t.run_async(n0 if rank == 0 else n1, 7, 0 if rank == 0 else n0)
    mixed = t.get_jmean()
    forms = torch.tensor([t.get_option("form")], device="cuda")
    both = [torch.zeros_like(forms) for _ in range(world)]
    dist.all_gather(both, forms)
    assert int(both[0]) in (5, 7, 8) and int(both[1]) in (1, 4), [int(b) for b in both]
    assert t.get_option("reduce_planes") == 24
    grids.append(mixed.copy())
    if peer:
        # many calls in a row (the two buffer halves alternate), the ranks arriving at different times
        first = None
        for i in range(24):
            if i % 5 == rank:
                time.sleep(0.02)
            t.seek(0)
            jm, st = t.run(60000, 7)
            first = jm.copy() if first is None else first
            assert np.allclose(jm, first, rtol=1e-12, atol=0), i
        with open(f"{out}.{rank}.state", "w") as f:
            f.write(str(t.get_option("peer_state")))
    np.save(f"{out}.{rank}.npy", np.stack(grids))
    dist.barrier()
    t.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("peer", [0, 1])
def test_shipped_regime_box_allreduce_and_column_form_two_ranks(tmp_path, peer):
    """Shipped (stub) regime on two ranks: reducing only the columns under the beam equals reducing the whole grid, for
    the step-by-step kernel and for the column form, and equals one GPU running all the ids.  peer = 1: the box is summed
    out of the other rank's buffer by k_peer_box_reduce ("peer_reduce") instead of by ncclAllReduce."""
    import torch

    import tamc
    from tests.util import compare_grids

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    out = str(tmp_path / "stub")
    mp.spawn(_stub_worker, args=(2, _free_port(), out, peer), nprocs=2, join=True)
    a, b = np.load(out + ".0.npy"), np.load(out + ".1.npy")
    assert np.array_equal(a, b)
    states = [int(open(f"{out}.{r}.state").read()) for r in range(2)] if peer else []
    cfg = tamc.configs.CONFIGS["shipped80"]
    t = tamc.MCTransport(80, 80, 80, cfg["xmax"], cfg["ymax"], cfg["zmax"], device=0)
    t.set_optics(cfg["rhokap"](), cfg["albedo"], cfg["hgg"], flags=0)
    t.run_async(120000, 7, 0)
    one = t.get_jmean()
    t.close()
    compare_grids(a[0], a[1], rtol=1e-12)               # same kernel, box vs whole-grid reduction (atomics order differs)
    for g in a[:-1]:
        compare_grids(g, one, rtol=1e-10)
    t = tamc.MCTransport(80, 80, 80, cfg["xmax"], cfg["ymax"], cfg["zmax"], device=0)
    t.set_optics(cfg["rhokap"](), cfg["albedo"], cfg["hgg"], flags=0)
    t.run_async(5_400_000, 7, 0)
    compare_grids(a[-1], t.get_jmean(), rtol=1e-10)     # ranks of different packet counts / kernel forms
    t.close()
    if peer and states != [1, 1]:
        assert states == [-1, -1], states                # the ranks agree on the fallback
        pytest.skip("CUDA IPC between the two processes is not available here: the box went through ncclAllReduce")
