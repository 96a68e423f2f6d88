"""The north-star's replay criterion taken literally: the CUDA path, fed the reference's ran2 sequence packet by packet,
against what the reference's OWN Fortran text computes (tests/golden/reference_interp.json.gz, made by executing ran2.f /
sourceph.f90 / inttau2.f90 / stokes.f90 / mcpolar.f90:153-169 with oracle/f90interp.py) -- not against the oracle.  The
oracle only supplies the draw list here (its ran2 is bit-identical to the reference's, tests/test_oracle_reference_vectors.py).
Tolerance: the north-star's 1e-6 relative for positions, directions and deposits; voxels, draws and voxel-steps exact."""
import gzip
import json
import os
import struct

import numpy as np
import pytest

from tests.util import make_oracle, make_transport

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-6


def _ref():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "reference_interp.json.gz"), "rt") as f:
        return json.load(f)


def unhex(h):
    return struct.unpack(">d", bytes.fromhex(h))[0]


def check_against_reference(rec, jm, v, cfg, scatter):
    """rec / jm: per-packet records and tally of a run; v: one entry of the reference vectors."""
    rows = v["packets"]
    assert len(rec) == len(rows)
    ext = (cfg["xmax"], cfg["ymax"], cfg["zmax"])
    worst = 0.0
    for q, row in enumerate(rows):
        r = rec[q]
        assert [int(r["xcell"]), int(r["ycell"]), int(r["zcell"])] == row[6:9], (q, r, row)
        assert int(r["ndraws"]) == row[10] and int(r["steps"]) == row[11], (q, r, row)
        if scatter:
            assert int(r["nscatt"]) == row[12] and (int(r["fate"]) == 0) == (row[13] == 1), (q, r, row)
        want = [unhex(h) for h in row[:6]]
        got = [float(r[k]) for k in ("xp", "yp", "zp", "nxp", "nyp", "nzp")]
        for k in range(6):
            scale = max(abs(want[k]), ext[k] if k < 3 else 1.0)
            err = abs(got[k] - want[k]) / scale
            worst = max(worst, err)
            assert err <= RTOL, (q, k, got[k], want[k])
    n = v["grid"]
    want_jm = np.zeros(n, order="F")
    for i, j, k, h in v["jmean"]:
        want_jm[i - 1, j - 1, k - 1] = unhex(h)
    nz = want_jm != 0
    if not scatter:
        assert np.array_equal(jm != 0, nz)                               # straight-down flights: the same voxels, exactly
        scale = np.abs(want_jm[nz])
    else:
        # relative to a voxel's optical depth (~1.3 here), as the other scatter-loop replay tests do: a flight that grazes a
        # voxel leaves a deposit far smaller than the 1e-9-level difference between two correct paths
        scale = np.maximum(np.abs(want_jm[nz]), 1.0)
        assert float(np.max(np.abs(jm[~nz]))) <= RTOL
    gerr = float(np.max(np.abs(jm[nz] - want_jm[nz]) / scale))
    assert gerr <= RTOL, gerr
    return worst, gerr


def _case(v, cfg, rk, scatter, cap):
    o = make_oracle(cfg, rk)
    o.seed_ran2(v["rank"])
    npk = len(v["packets"])
    out = o.run(npk, records=True, draws_cap=npk * cap)                 # the draw list (ran2.f's sequence for this rank)
    assert int(out["offsets"][-1]) == sum(r[10] for r in v["packets"])   # as many draws as the reference consumed
    t = make_transport(cfg, rk)
    rec, jm = t.run_replay(out["offsets"], out["draws"])
    worst, gerr = check_against_reference(rec, jm, v, cfg, scatter)
    t.close()
    return worst, gerr


def test_replay_against_the_reference_text_shipped_loop():
    import tamc

    v = _ref()["shipped"][0]
    cfg = tamc.configs.CONFIGS["shipped80"]
    assert v["grid"] == [80, 80, 80] and unhex(v["kappa"]) == 680.0
    worst, gerr = _case(v, cfg, cfg["rhokap"](), False, 4)
    assert worst < 1e-9 and gerr < 1e-11                                # what the replay actually achieves


def test_replay_against_the_reference_text_scatter_loop():
    v = _ref()["scatter"][0]
    n = v["grid"][0]
    cfg = dict(n=n, xmax=v["extents"][0], ymax=v["extents"][1], zmax=v["extents"][2], albedo=v["mus"] / (v["mus"] + v["mua"]),
               hgg=v["hgg"], flags=1)
    rk = np.zeros((n + 2, n + 2, n + 2), order="F")
    rk[1:-1, 1:-1, 1:-1] = v["mus"] + v["mua"]
    _case(v, cfg, rk, True, 6000)
