"""The overlapped host boundary of the shipped regime (tamc_run / tamc_run_optics, csrc/tamc_api.cu):
zero fill + beam columns down, beam columns ahead of the full grid up.  What lands in the caller's
jmeanGLOBAL must be BIT-IDENTICAL to the tally resident on the device (a full-grid download of the same call), and
equal to the plain sequence tamc_set_optics + tamc_run with full-grid copies ("box_io" = 0) up to the fp64
summation order of the atomics (two runs of the same call never add in the same order; counters are exact).
Stale host content outside the beam's columns must be overwritten with zeros."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import compare_grids, make_oracle

pytestmark = pytest.mark.gpu

SEED = 20261017


def _pinned(tamc, a):
    tamc.pin_host(a)
    return a


def _pair(tamc, cfg, rk, n, dims=None, spot=None):
    """Runs the same call (same packet ids) through the plain and the overlapped boundary; returns both results."""
    nx, ny, nz = dims or (cfg["n"],) * 3
    out = []
    for box_io in (0, -1):
        t = tamc.MCTransport(nx, ny, nz, cfg["xmax"], cfg["ymax"], cfg["zmax"])
        if spot is not None:
            t.set_source_co2(spot)
        t.set_option("box_io", box_io)
        jm = _pinned(tamc, t.new_jmean())
        jm[...] = 7.25                                     # stale content that every byte of the download must replace
        rkp = _pinned(tamc, np.asfortranarray(rk.copy()))
        got, st = t.run_optics(rkp, cfg["albedo"], cfg["hgg"], n, SEED, flags=cfg["flags"], out=jm)
        resident = t.get_jmean().copy()
        io, form = t.get_option("io_form"), t.get_option("form")
        # a second call on the resident grid: the full upload behind the first call must have landed
        jm2 = _pinned(tamc, t.new_jmean())
        jm2[...] = -3.5
        got2, st2 = t.run(n, SEED, out=jm2)
        out.append(dict(jm=got.copy(), st=st, resident=resident, jm2=got2.copy(), st2=st2, io=io, form=form))
        tamc.unpin_host(jm)
        tamc.unpin_host(jm2)
        tamc.unpin_host(rkp)
        t.close()
    return out


@pytest.mark.parametrize("name,n", [("shipped80", 125000), ("homog200", 1500000), ("homog200", 5000000)])
def test_overlapped_boundary_is_bit_identical(name, n):
    import tamc

    cfg = tamc.configs.CONFIGS[name]
    plain, fast = _pair(tamc, cfg, cfg["rhokap"](), n)
    assert plain["io"] == 0
    assert fast["io"] == (3 if n >= (1 << 20) else 1)      # columns-first upload only with the column form
    for key in ("jm", "resident", "jm2"):
        compare_grids(fast[key], plain[key], rtol=1e-11)
    assert np.array_equal(fast["jm"], fast["resident"])          # the download itself: every byte
    assert np.array_equal(plain["jm"], plain["resident"])
    for key in ("packets", "voxel_steps", "absorbed", "exits"):
        assert plain["st"][key] == fast["st"][key] and plain["st2"][key] == fast["st2"][key]
    assert abs(fast["jm"].sum() / n - 1.0) < 5 / np.sqrt(n)
    assert fast["jm2"].sum() > 0 and not np.array_equal(fast["jm"], fast["jm2"])      # the cursor moved on


def test_overlapped_boundary_against_the_oracle_with_a_crater():
    """Heterogeneous grid (ablated crater + water-depleted rim) through the columns-first upload, against the oracle
    on the same Philox stream."""
    import tamc

    cfg = tamc.configs.scaled("homog200", 120)
    rk = list(tamc.configs.crater_sequence(120, 6))[4]
    n = (1 << 20) + 4321
    t = tamc.MCTransport(120, 120, 120, cfg["xmax"], cfg["ymax"], cfg["zmax"])
    t.set_option("column", 1)                 # 51 x 51 columns under the beam: below the auto threshold
    rkp = _pinned(tamc, np.asfortranarray(rk.copy()))
    jm = _pinned(tamc, t.new_jmean())
    jm[...] = 1.0
    got, st = t.run_optics(rkp, 0.0, 0.9, n, SEED, out=jm)
    assert t.get_option("io_form") == 3 and t.get_option("form") == 5
    assert np.array_equal(got, t.get_jmean())
    o = make_oracle(cfg, rk)
    o.zero_jmean()
    o.seed_philox(SEED, 0)
    want = o.run(n)["stats"]
    for key in ("packets", "voxel_steps", "absorbed", "exits"):
        assert st[key] == want[key], key
    compare_grids(got, o.jmean, rtol=1e-10)
    tamc.unpin_host(jm)
    tamc.unpin_host(rkp)
    t.close()


@pytest.mark.parametrize("dims", [(64, 40, 24), (33, 70, 9)])
def test_overlapped_boundary_non_cubic(dims):
    import tamc

    nx, ny, nz = dims
    cfg = dict(xmax=0.04, ymax=0.03, zmax=0.02, albedo=0.0, hgg=0.9, flags=0)
    rk = np.zeros((nx + 2, ny + 2, nz + 2), order="F")
    ii, jj, kk = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    rk[1:-1, 1:-1, 1:-1] = 20.0 + 15.0 * ((ii + 2 * jj + 3 * kk) % 5)
    for n, column in ((40000, -1), (300000, 1)):
        res = []
        for box_io in (0, -1):
            t = tamc.MCTransport(nx, ny, nz, cfg["xmax"], cfg["ymax"], cfg["zmax"])
            t.set_source_co2(0.012)
            t.set_option("box_io", box_io)
            t.set_option("column", column)
            jm = _pinned(tamc, t.new_jmean())
            rkp = _pinned(tamc, np.asfortranarray(rk.copy()))
            got, st = t.run_optics(rkp, 0.0, 0.9, n, SEED, out=jm)
            assert np.array_equal(got, t.get_jmean())
            res.append((got.copy(), st, t.get_option("io_form")))
            tamc.unpin_host(jm)
            tamc.unpin_host(rkp)
            t.close()
        assert res[0][2] == 0 and res[1][2] == (3 if column == 1 else 1)
        compare_grids(res[1][0], res[0][0], rtol=1e-11)
        assert res[0][1]["voxel_steps"] == res[1][1]["voxel_steps"] and res[1][1]["exits"][4] > 0


def test_plain_path_when_not_applicable():
    """Pageable host arrays, a beam wider than half the face, or the scatter loop: full-grid copies, same results."""
    import tamc

    cfg = tamc.configs.CONFIGS["shipped80"]
    rk = cfg["rhokap"]()
    t = tamc.MCTransport(80, 80, 80, cfg["xmax"], cfg["ymax"], cfg["zmax"])
    jm, _ = t.run_optics(rk, 0.0, 0.9, 50000, SEED)                      # pageable
    assert t.get_option("io_form") == 0
    jp = _pinned(tamc, t.new_jmean())
    t.seek(0)
    got, _ = t.run(50000, SEED, out=jp)
    assert t.get_option("io_form") == 1 and np.array_equal(got, t.get_jmean())
    compare_grids(got, jm, rtol=1e-11)
    t.set_source_co2(0.05)                                               # 67 of 80 voxels wide
    t.run(50000, SEED, out=jp)
    assert t.get_option("io_form") == 0
    t.set_source_co2(0.025)
    t.run_optics(None, 0.9, 0.9, 20000, SEED, flags=tamc.SCATTER, out=jp)
    assert t.get_option("io_form") == 0 and jp.sum() > 0
    tamc.unpin_host(jp)
    t.close()


def _deep_grid(nx, ny, nz, kappa=30.0):
    rk = np.zeros((nx + 2, ny + 2, nz + 2), order="F")
    ii, jj, kk = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    rk[1:-1, 1:-1, 1:-1] = kappa * (1.0 + 0.25 * ((ii + 2 * jj + 3 * kk) % 4))
    return rk


@pytest.mark.parametrize("dims,depth,tile,park", [((120, 120, 120), 8, -1, -1), ((120, 120, 120), 40, 12, 1), ((33, 40, 70), 5, 0, -1),
                                                  ((48, 48, 96), 1, 23, 1), ((48, 48, 96), 20, 12, 0)])
def test_depth_limited_upload_reads_deeper_planes_from_the_callers_grid(dims, depth, tile, park):
    """Columns-first upload cut at `gather_depth` planes on a grid whose packets go far deeper (mean free path ~ 8 voxels,
    some leave through the bottom face): k_column_bound carries the deeper planes a packet can reach from the caller's
    page-locked array into the resident grid, where the transport and the finish kernel fetch what was not copied
    (untiled, tiled and regrouped column kernels; the handles are fresh, so a voxel that was not carried over would be
    read as 0).  Same packets, counters and grid as the plain path and as the oracle on the same Philox stream."""
    import tamc

    nx, ny, nz = dims
    ext = dict(xmax=0.03, ymax=0.03, zmax=0.06)
    rk = _deep_grid(nx, ny, nz)
    n = (1 << 20) + 777
    res = []
    for box_io, gd in ((0, 0), (-1, depth), (-1, 0)):
        t = tamc.MCTransport(nx, ny, nz, ext["xmax"], ext["ymax"], ext["zmax"])
        t.set_option("box_io", box_io)
        t.set_option("column", 1)
        t.set_option("column_tile", tile)
        t.set_option("column_park", park)
        t.set_option("gather_depth", gd)
        jm = _pinned(tamc, t.new_jmean())
        jm[...] = 3.0
        rkp = _pinned(tamc, np.asfortranarray(rk.copy()))
        got, st = t.run_optics(rkp, 0.0, 0.9, n, SEED, out=jm)
        assert np.array_equal(got, t.get_jmean())
        res.append((got.copy(), st, t.get_option("io_form"), t.get_option("depth_hint")))
        # the resident grid is complete after the call (the full upload ran beside the transport)
        t.set_option("box_io", 0)
        t.seek(0)
        again, st2 = t.run(n, SEED)
        compare_grids(again, got, rtol=1e-11)
        assert st2["voxel_steps"] == st["voxel_steps"]
        tamc.unpin_host(jm)
        tamc.unpin_host(rkp)
        t.close()
    (plain, st0, io0, _), (cut, st1, io1, dh), (full, st2, io2, _) = res
    assert io0 == 0 and io1 == 7 and io2 == 3
    assert dh > depth and st1["exits"][4] > 0                      # packets did go below the copied planes, some to the bottom
    for key in ("packets", "voxel_steps", "absorbed", "exits"):
        assert st0[key] == st1[key] == st2[key], key
    compare_grids(cut, plain, rtol=1e-11)
    compare_grids(full, plain, rtol=1e-11)
    o = orc.Oracle(nx, ny, nz, ext["xmax"], ext["ymax"], ext["zmax"])
    o.set_rhokap(rk)
    o.set_optics(0.0, 0.9)
    o.seed_philox(SEED, 0)
    want = o.run(n)["stats"]
    assert st1["voxel_steps"] == want["voxel_steps"] and st1["exits"] == want["exits"]
    compare_grids(cut, o.jmean, rtol=1e-10)


def test_depth_limit_follows_the_previous_call():
    """Auto mode: the first tamc_run_optics copies every plane of the beam's columns, the next ones only down to the
    depth the previous call reached plus a margin; when the top 40 planes vanish between two calls (packets suddenly
    40 voxels deeper than the copied planes) the result is still exact -- the deeper planes come from the caller's array
    (through the resident grid, k_column_bound)."""
    import tamc

    cfg = tamc.configs.CONFIGS["shipped80"]
    n = 1_500_000
    t = tamc.MCTransport(80, 80, 80, cfg["xmax"], cfg["ymax"], cfg["zmax"])
    ref = tamc.MCTransport(80, 80, 80, cfg["xmax"], cfg["ymax"], cfg["zmax"])
    ref.set_option("gather_depth", 0)
    for h in (t, ref):
        h.set_option("column", 1)                                       # narrow beam, 1.5e6 packets: below the auto threshold
    grids = [cfg["rhokap"](), cfg["rhokap"](), list(tamc.configs.crater_sequence(80, 8))[7], cfg["rhokap"]()]
    grids[2][1:-1, 1:-1, 41:81] = 0.0                                   # the top 40 planes gone everywhere
    jm, jr = _pinned(tamc, t.new_jmean()), _pinned(tamc, ref.new_jmean())
    forms = []
    for rk in grids:
        rkp = _pinned(tamc, np.asfortranarray(rk.copy()))
        got, st = t.run_optics(rkp, 0.0, 0.9, n, SEED, out=jm)
        want, sr = ref.run_optics(rkp, 0.0, 0.9, n, SEED, out=jr)
        forms.append((t.get_option("io_form"), t.get_option("depth_hint"), ref.get_option("io_form")))
        for key in ("packets", "voxel_steps", "absorbed", "exits"):
            assert st[key] == sr[key], key
        compare_grids(got, want, rtol=1e-11)
        assert np.array_equal(got, t.get_jmean())
        tamc.unpin_host(rkp)
    # (after the deep call the limit would cover the whole column again: plain columns-first upload)
    assert [f[0] for f in forms] == [3, 7, 7, 3] and all(f[2] == 3 for f in forms)
    assert 8 <= forms[0][1] <= 40 and forms[2][1] >= forms[1][1] + 30       # the third call reached far deeper
    tamc.unpin_host(jm)
    tamc.unpin_host(jr)
    t.close()
    ref.close()


def test_an_abrupt_change_of_tissue_does_not_fall_off_a_cliff():
    """homog200, the opacity drops 16-fold between two tamc_run_optics calls, so nearly every packet of the
    next call goes below the planes the depth-limited upload copied.  Measured before k_column_bound filled the resident
    grid: 1 160 ms for that call against 7.9 ms with every plane copied (1e8 packets).  Same result, and the call now
    costs about what the every-plane handle pays (generous factor: host clocks on a shared box)."""
    import time
    import tamc

    c = tamc.configs.CONFIGS["homog200"]
    n = 20_000_000
    rk_a = _pinned(tamc, c["rhokap"]())
    rk_b = _pinned(tamc, np.asfortranarray(rk_a / 16.0))
    t = tamc.MCTransport(200, 200, 200, c["xmax"], c["ymax"], c["zmax"])
    ref = tamc.MCTransport(200, 200, 200, c["xmax"], c["ymax"], c["zmax"])
    ref.set_option("gather_depth", 0)
    jm, jr = _pinned(tamc, t.new_jmean()), _pinned(tamc, ref.new_jmean())
    took = []
    for i, g in enumerate((rk_a, rk_a, rk_a, rk_b)):
        t0 = time.perf_counter()
        got, st = t.run_optics(g, 0.0, 0.9, n, SEED + i, out=jm)
        t1 = time.perf_counter()
        want, sr = ref.run_optics(g, 0.0, 0.9, n, SEED + i, out=jr)
        t2 = time.perf_counter()
        took.append((t1 - t0, t2 - t1, t.get_option("io_form")))
        for key in ("packets", "voxel_steps", "absorbed", "exits"):
            assert st[key] == sr[key], key
        compare_grids(got, want, rtol=1e-11)
    assert took[2][2] == 7 and took[3][2] == 7                       # both under a depth limit; the last one far too shallow
    assert st["exits"][4] > 0                                        # packets reached the bottom face
    assert took[3][0] < 3.0 * took[3][1] + 2e-3, took
    for a in (rk_a, rk_b, jm, jr):
        tamc.unpin_host(a)
    t.close()
    ref.close()
