"""root_io (several ranks, tamc_api.cu run_boundary): only rank 0's host arrays are read / written -- its rhokap reaches the
other GPUs over NVLink, only its jmeanGLOBAL is downloaded -- and the result equals the all-ranks mode and one GPU running
all the ids.  Needs 2 GPUs."""
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
    res = {}
    # (a) shipped stub regime, large call: the overlapped column path, shared gather; (b) a small call: plain copies;
    # (c) scatter loop: plain copies + whole-grid all-reduce
    cases = {"stub_big": ("shipped80", 0, 3_000_000), "stub_small": ("shipped80", 0, 40_000), "scatter": ("skin200", 1, 30_000)}
    for key, (name, flags, n) in cases.items():
        cfg = tamc.configs.scaled(name, 64) if name == "skin200" else tamc.configs.CONFIGS[name]
        g = cfg["n"]
        grids = [cfg["rhokap"](), None]
        grids[1] = grids[0].copy(order="F")
        grids[1][1:-1, 1:-1, -4:-1] *= 0.5                  # the opacity changes between the calls
        for mode in (0, 1):
            t = tamc.MCTransport(g, g, g, cfg["xmax"], cfg["ymax"], cfg["zmax"], device=rank)
            t.set_option("root_io", mode)
            t.set_optics(grids[0], cfg["albedo"], cfg["hgg"], flags=cfg["flags"])
            t.comm_init(world, rank, tdist.broadcast_unique_id(tamc.comm_unique_id, dist))
            outs = []
            for call in range(3):
                rk = grids[call % 2]
                if mode == 1 and rank > 0:
                    rk = np.full_like(rk, 1e9)              # never read on ranks > 0: poison it
                tamc.pin_host(rk)
                jm = t.new_jmean()
                tamc.pin_host(jm)
                jm[...] = -7.0
                _, st = t.run_optics(rk, cfg["albedo"], cfg["hgg"], n, 11, flags=cfg["flags"], out=jm)
                outs.append(jm.copy())
                assert st["packets"] == n
                if mode == 1:
                    assert t.get_option("io_form") & 8
                    if rank > 0:
                        assert np.all(jm == -7.0)           # only rank 0's jmeanGLOBAL is written
                tamc.unpin_host(rk); tamc.unpin_host(jm)
            # a call on the RESIDENT grid: every rank's copy must be rank 0's last upload (after an overlapped call the
            # other ranks hold only the beam's columns until the next reader of the resident grid fetches the rest)
            jm4, st4 = t.run(n, 11)
            outs.append(jm4.copy() if (mode == 0 or rank == 0) else outs[-1] * 0)
            res[(key, mode)] = np.stack(outs)
            t.close()
    if rank == 0:
        np.savez(out, **{f"{k}_{m}": v for (k, m), v in res.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_root_io_equals_all_ranks_io(tmp_path):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from tests.util import compare_grids

    out = str(tmp_path / "rootio.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    for key in ("stub_big", "stub_small", "scatter"):
        a, b = z[f"{key}_0"], z[f"{key}_1"]
        assert a.shape == b.shape and a.sum() > 0
        for i in range(a.shape[0]):
            # same packet ids, same grids: the two modes differ by the order of the atomics only
            compare_grids(b[i], a[i], rtol=1e-10)
        assert not np.array_equal(a[0], a[1])               # the second call saw the changed opacity
