"""World-size-2 checks of the host-side multi-rank logic on CPU (gloo): id broadcast, packet-range
partition, and -- with the oracle standing in for the device -- that per-rank tallies all-reduced over
the ranks equal one rank running every id (the invariant the NCCL path relies on)."""
import os
import socket

import numpy as np
import pytest


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

    from oracle import oracle as orc
    from tamc import dist as tdist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert tdist.env_rank() == (rank, world, rank)
    uid = tdist.broadcast_unique_id(lambda: bytes((7 * i + 3) % 256 for i in range(128)), dist)
    assert uid == bytes((7 * i + 3) % 256 for i in range(128))

    n, cursor = 3000, 1000
    lo, hi, nxt = tdist.packet_range(rank, world, n, cursor)
    assert (lo, hi, nxt) == (cursor + rank * n, cursor + (rank + 1) * n, cursor + world * n)

    o = orc.Oracle(20, 20, 20, 0.05, 0.05, 0.05)
    o.gridset_uniform(80.0)
    o.set_optics(0.9, 0.8)
    o.set_flags(orc.FLAG_SCATTER)
    o.seed_philox(42, lo)
    st = o.run(n)["stats"]
    t = torch.from_numpy(np.ascontiguousarray(o.jmean.ravel(order="F")).copy())
    dist.all_reduce(t)                                   # stands in for ncclAllReduce / MPI_allREDUCE
    steps = torch.tensor([st["voxel_steps"]], dtype=torch.int64)
    dist.all_reduce(steps)
    if rank == 0:
        np.save(out, t.numpy())
        np.save(out + ".steps.npy", steps.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_sums_to_single_rank(tmp_path):
    import torch.multiprocessing as mp

    from oracle import oracle as orc

    out = str(tmp_path / "sum.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    steps = int(np.load(out + ".steps.npy")[0])
    o = orc.Oracle(20, 20, 20, 0.05, 0.05, 0.05)
    o.gridset_uniform(80.0)
    o.set_optics(0.9, 0.8)
    o.set_flags(orc.FLAG_SCATTER)
    o.seed_philox(42, 1000)
    st = o.run(6000)["stats"]
    want = o.jmean.ravel(order="F")
    assert steps == st["voxel_steps"]
    assert np.array_equal(got != 0, want != 0)
    nz = want != 0
    assert np.abs(got[nz] / want[nz] - 1).max() < 1e-12


def test_packet_ranges_tile_the_id_space():
    from tamc import dist as tdist

    cursor = 0
    seen = []
    for call in range(3):
        for r in range(4):
            lo, hi, nxt = tdist.packet_range(r, 4, 125000, cursor)
            seen.append((lo, hi))
        cursor = nxt
    seen.sort()
    assert seen[0][0] == 0 and seen[-1][1] == 3 * 4 * 125000
    assert all(a[1] == b[0] for a, b in zip(seen, seen[1:]))
