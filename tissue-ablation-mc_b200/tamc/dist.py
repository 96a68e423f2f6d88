"""Host-side plumbing for one-process-per-GPU runs (torch.distributed carries the NCCL id and the
barriers; the tally reduction itself is the library's ncclAllReduce, mcpolar.f90:173)."""
from __future__ import annotations

import os


def env_rank():
    """(rank, world, local_rank) from the torchrun / MPI-style environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def packet_range(rank: int, world: int, nphotons: int, cursor: int = 0):
    """Ids rank `rank` runs in an MC call of `nphotons` packets PER RANK starting at `cursor`, and the
    cursor after the call -- the rule tamc_run applies (include/tamc.h): rank r takes
    [cursor + r*n, cursor + (r+1)*n), like the reference's per-rank `do j = 1, nphotons`."""
    first = cursor + rank * nphotons
    return first, first + nphotons, cursor + world * nphotons


def broadcast_unique_id(make_id, dist=None, device=None) -> bytes:
    """Rank 0 calls make_id() (tamc.comm_unique_id); everyone returns the same 128 bytes."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return make_id()
    import torch

    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if dist.get_rank() == 0:
        raw = make_id()
        assert len(raw) == 128
        buf = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
    dist.broadcast(buf, 0)
    return bytes(buf.cpu().numpy().tobytes())
