"""Column form of the shipped (stub) regime (csrc/tamc_column.cuh) through the C ABI.

The column kernels tally each packet's final partial deposit plus a stop count, and reconstruct the full-voxel
crossings as F * dcell * rhokap.  Same Philox streams and the same taurun accumulation as the step-by-step
kernels, so against the oracle run on the same counter-based stream: counters bit-exact, grid to fp64 summation
order (1e-10 relative, identical support).  Integer work (voxel-steps, fates) must be bit-exact."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import compare_grids, make_oracle, make_transport

pytestmark = pytest.mark.gpu

SEED = 20261017


def _check(t, o, n, first=0):
    o.zero_jmean()
    o.seed_philox(SEED, first)
    want = o.run(n)["stats"]
    t.set_option("variant", 1)                       # step-by-step persistent kernel
    t.run_async(n, SEED, first)
    base, sb = t.get_jmean().copy(), t.get_stats()
    t.set_option("variant", 3)
    for column in (1, 2):
        t.set_option("column", column)
        for rep in range(2):                        # twice: the stop counts must be all zero again after a call
            t.run_async(n, SEED, first)
            jm, st = t.get_jmean(), t.get_stats()
            assert st["gpu_launches"] == (3 if column == 1 else 2)
            for key in ("packets", "voxel_steps", "scatters", "absorbed", "exits"):
                assert st[key] == want[key] == sb[key], (key, column, rep)
            compare_grids(jm, o.jmean, rtol=1e-10)
            compare_grids(jm, base, rtol=1e-10)
    # shared-memory tiles for the top planes (10*ta + tb): deposits only, counts only, both, deeper than the grid allows
    t.set_option("column", 1)
    # ... each with the column walk regrouped through the per-warp park queues (default) and without
    for split, park in ((10, 1), (4, 1), (23, 1), (14, 1), (14, 0), (4, 0)):
        t.set_option("column_tile", split)
        t.set_option("column_park", park)
        for rep in range(2):
            t.run_async(n, SEED, first)
            jm, st = t.get_jmean(), t.get_stats()
            for key in ("packets", "voxel_steps", "scatters", "absorbed", "exits"):
                assert st[key] == want[key], (key, split, park, rep)
            compare_grids(jm, o.jmean, rtol=1e-10)
    t.set_option("column_tile", -1)
    t.set_option("column_park", -1)
    t.set_option("column", -1)


def test_column_shipped_and_crater():
    import tamc

    cfg = tamc.configs.CONFIGS["shipped80"]
    t, o = make_transport(cfg), make_oracle(cfg)
    _check(t, o, 125000)
    _check(t, o, 30000, first=(1 << 32) - 10000)      # ids crossing 2^32
    t.close()
    # ablated crater (rhokap == 0 voxels are crossed with zero deposit) with a water-depleted rim
    for rk in list(tamc.configs.crater_sequence(80, 6))[3:6]:
        t, o = make_transport(cfg, rk), make_oracle(cfg, rk)
        _check(t, o, 60000)
        t.close()


@pytest.mark.parametrize("n", [50, 63, 200])
def test_column_homog_grids(n):
    """nzg = 50 and 63 are not multiples of 4 (partial first 256-bit group); 200 is the bench grid."""
    import tamc

    cfg = tamc.configs.scaled("homog200", n)
    t, o = make_transport(cfg), make_oracle(cfg)
    _check(t, o, 80000)
    t.close()


@pytest.mark.parametrize("dims", [(1, 1, 1), (3, 5, 7), (17, 4, 33), (40, 24, 6)])
def test_column_non_cubic_heterogeneous_with_transmission(dims):
    """Heterogeneous opacity, non-cubic boxes, thin enough that a good fraction of the packets crosses every voxel and
    leaves through the bottom face (plane 0 of the stop counts)."""
    import tamc

    nx, ny, nz = dims
    xmax, ymax, zmax = 0.02, 0.03, 0.05
    rk = np.zeros((nx + 2, ny + 2, nz + 2), order="F")
    ii, jj, kk = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    rk[1:-1, 1:-1, 1:-1] = (3.0 + 4.0 * ((ii + 2 * jj + 3 * kk) % 5)) * ((ii + kk) % 7 != 0)     # some voxels empty
    o = orc.Oracle(nx, ny, nz, xmax, ymax, zmax)
    o.set_rhokap(rk)
    o.set_optics(0.0, 0.9)
    o.set_flags(0)
    t = tamc.MCTransport(nx, ny, nz, xmax, ymax, zmax)
    t.set_optics(rk, 0.0, 0.9)
    _check(t, o, 50000)
    st = t.get_stats()
    assert st["exits"][4] > 1000 and st["absorbed"] > 1000
    t.close()


def test_column_is_the_default_for_large_calls_and_matches_the_step_kernel():
    """>= 2^20 packets on the default variant take the column form (3 launches); its grid equals the step-by-step
    kernel's to summation order, the counters exactly."""
    import tamc

    cfg = tamc.configs.CONFIGS["homog200"]
    n = 3_000_000
    t = make_transport(cfg)
    t.run_async(n, SEED, 0)
    jm, st = t.get_jmean().copy(), t.get_stats()
    assert st["gpu_launches"] == 3
    t.set_option("column", 0)
    t.run_async(n, SEED, 0)
    jm0, st0 = t.get_jmean(), t.get_stats()
    assert st0["gpu_launches"] == 1
    for key in ("packets", "voxel_steps", "absorbed", "exits"):
        assert st[key] == st0[key]
    compare_grids(jm, jm0, rtol=1e-10)
    assert abs(jm.sum() / n - 1.0) < 5 / np.sqrt(n)
    t.close()


@pytest.mark.parametrize("name,n,split", [("homog200", 12_000_000, (0, 1)), ("shipped80", 20_000_000, (1, 2))])
def test_auto_tiles_for_calls_that_amortise_the_flush(name, n, split):
    """Large calls take the column form with shared-memory tiles (form 7): one plane of stop counts under a wide beam,
    the top plane of deposits + two planes of counts under a narrow one.  Same counters, grid to summation order; the
    probe follows the same shape."""
    import tamc

    cfg = tamc.configs.CONFIGS[name]
    t = make_transport(cfg)
    t.run_async(n, SEED, 0)
    jm, st = t.get_jmean().copy(), t.get_stats()
    assert t.get_option("form") in (7, 8) and st["gpu_launches"] == 3
    t.set_option("column", 1)
    t.set_option("column_tile", 0)
    t.run_async(n, SEED, 0)
    jm0, st0 = t.get_jmean().copy(), t.get_stats()
    assert t.get_option("form") == 5
    for key in ("packets", "voxel_steps", "absorbed", "exits"):
        assert st[key] == st0[key]
    compare_grids(jm, jm0, rtol=1e-10)
    assert abs(jm.sum() / n - 1.0) < 5 / np.sqrt(n)
    t.set_option("column", -1)
    t.set_option("column_tile", -1)
    ms, steps = t.roofline_probe(n, 3)
    assert ms > 0 and abs(steps / st["voxel_steps"] - 1.0) < 0.02
    t.run_async(1000, SEED, 0)                         # the probe leaves the stop counts clean
    assert t.get_option("form") == 1 and abs(t.get_jmean().sum() / 1000 - 1.0) < 0.2
    t.close()


@pytest.mark.parametrize("dims,ext,spot", [((200, 200, 200), (0.03, 0.03, 0.06), 0.025), ((80, 80, 80), (0.03, 0.03, 0.06), 0.025),
                                           ((4096, 64, 4), (0.5, 0.01, 0.01), 0.019), ((33, 70, 9), (0.04, 0.03, 0.02), 0.012),
                                           ((400, 400, 8), (1.0, 1.0, 1.0), 1.9)])
def test_fp32_first_pass_of_the_launch_voxel_never_disagrees(dims, ext, spot):
    """The column form's launch voxel: fp32 first pass + fp64 redo near voxel edges.  Both passes over 4e9 Philox blocks:
    the first pass never keeps a voxel the fp64 arithmetic would not give, and hands over only a few draws in a thousand
    (more on grids with thousands of voxels per axis, where fp32 resolves a voxel less finely)."""
    import tamc

    t = tamc.MCTransport(*dims, *ext)
    t.set_source_co2(spot)
    n = 4_000_000_000
    fb, bad = t.selfcheck_launch(n, seed=20261017)
    assert bad == 0
    assert 0 < fb < (0.2 if max(dims) > 1000 else 0.02) * n
    t.set_option("launch32", 0)
    fb0, bad0 = t.selfcheck_launch(1_000_000, seed=3)
    assert (fb0, bad0) == (1_000_000, 0)                    # switched off: everything goes to fp64
    t.close()


def test_fp32_first_pass_on_and_off_give_the_same_call():
    import tamc

    cfg = tamc.configs.CONFIGS["homog200"]
    rk = cfg["rhokap"]()
    n = 6_000_000
    res = []
    for on in (1, 0):
        t = tamc.MCTransport(200, 200, 200, cfg["xmax"], cfg["ymax"], cfg["zmax"])
        t.set_option("launch32", on)
        t.set_optics(rk, 0.0, 0.9)
        t.run_async(n, 99, 0)
        res.append((t.get_jmean().copy(), t.get_stats(), t.get_option("form")))
        t.close()
    assert res[0][2] in (5, 7, 8) and res[1][2] == res[0][2]
    for key in ("packets", "voxel_steps", "absorbed", "exits"):
        assert res[0][1][key] == res[1][1][key]
    assert np.array_equal(res[0][0] != 0, res[1][0] != 0)
    nz = res[1][0] != 0
    assert np.abs(res[0][0][nz] / res[1][0][nz] - 1).max() < 1e-11


def test_depth_bound_of_a_call():
    """k_column_bound: a packet's optical depth is at most 33 ln 2 = 22.87 (one 32-bit Philox word), so no packet gets
    deeper than where its column's running optical depth passes that -- the bound the multi-rank all-reduce uses
    ("reduce_bound" = 2 computes it without a communicator).  It must cover the deepest stop actually seen."""
    import tamc

    cfg = tamc.configs.CONFIGS["homog200"]
    t = tamc.MCTransport(200, 200, 200, cfg["xmax"], cfg["ymax"], cfg["zmax"])
    t.set_option("reduce_bound", 2)
    t.set_option("column", 1)
    rk = cfg["rhokap"]()
    t.set_optics(rk, 0.0, 0.9)
    t.run_async(3_000_000, 5, 0)
    st = t.get_stats()
    # 680 / cm x 0.0006 cm = 0.408 per voxel: 23 / 0.408 -> 57 voxels, one spare
    assert t.get_option("reduce_planes") == 58
    assert 20 < t.get_option("depth_hint") <= 57 and st["exits"][4] == 0
    # a crater (rhokap = 0 in the top planes under the beam) pushes the bound down by its depth; a transparent column
    # lifts it to the whole grid
    rk2 = rk.copy()
    rk2[90:112, 90:112, 181:201] = 0.0
    t.set_optics(rk2, 0.0, 0.9)
    t.run_async(3_000_000, 5, 0)
    assert t.get_option("reduce_planes") == 78 and t.get_option("depth_hint") <= 77
    rk2[100, 100, :] = 0.0
    t.set_optics(rk2, 0.0, 0.9)
    t.run_async(3_000_000, 5, 0)
    st = t.get_stats()
    assert t.get_option("reduce_planes") == 200 and st["exits"][4] > 0 and t.get_option("depth_hint") == 200
    t.set_option("reduce_bound", 1)
    t.run_async(3_000_000, 5, 0)
    t.sync()
    assert t.get_option("reduce_planes") == 0                  # single rank, default: not computed
    t.close()
