"""Trace replay: the CUDA path fed the reference's ran2 sequence packet by packet must reproduce the
oracle's photon paths and deposits (north_star: 1e-6 relative, fp64).  Calls go through the C ABI."""
import numpy as np
import pytest

from tests.util import compare_grids, compare_records, make_oracle, make_transport, voxel_tau

pytestmark = pytest.mark.gpu


def _replay_case(cfg, npackets, rank=0, cap_per_packet=4, rhokap=None):
    import tamc  # noqa: F401

    rk = rhokap if rhokap is not None else cfg["rhokap"]()
    o = make_oracle(cfg, rk)
    o.seed_ran2(rank)
    out = o.run(npackets, records=True, draws_cap=npackets * cap_per_packet)
    t = make_transport(cfg, rk)
    rec, jm = t.run_replay(out["offsets"], out["draws"])
    st = t.get_stats()
    scale = {"xp": cfg["xmax"], "yp": cfg["ymax"], "zp": cfg["zmax"], "nxp": 1.0, "nyp": 1.0, "nzp": 1.0}
    worst = compare_records(rec, out["records"], scale=scale)
    if cfg["flags"] & 1:
        gerr = compare_grids(jm, o.jmean, rtol=1e-6, dep_scale=voxel_tau(cfg, rk))
    else:
        gerr = compare_grids(jm, o.jmean)
    assert st["packets"] == npackets
    assert st["voxel_steps"] == out["stats"]["voxel_steps"]
    assert st["scatters"] == out["stats"]["scatters"]
    assert st["absorbed"] == out["stats"]["absorbed"]
    assert st["exits"] == out["stats"]["exits"]
    t.close()
    return worst, gerr


def test_replay_shipped_regime_80():
    import tamc

    worst, gerr = _replay_case(tamc.configs.CONFIGS["shipped80"], 125000)
    assert worst < 1e-9 and gerr < 1e-11


def test_replay_other_rank_seed():
    import tamc

    _replay_case(tamc.configs.CONFIGS["shipped80"], 20000, rank=5)


def test_replay_homogeneous_200_full_grid():
    import tamc

    _replay_case(tamc.configs.CONFIGS["homog200"], 400000)


def test_replay_turbid_scatter_loop():
    import tamc

    cfg = tamc.configs.scaled("turbid200", 60)
    _replay_case(cfg, 4000, rank=2, cap_per_packet=4000)


def test_replay_turbid_200_subset():
    import tamc

    _replay_case(tamc.configs.CONFIGS["turbid200"], 3000, cap_per_packet=6000)


def test_replay_layered_skin():
    import tamc

    cfg = tamc.configs.scaled("skin200", 100)
    _replay_case(cfg, 5000, rank=1, cap_per_packet=4000)


def test_replay_isotropic_branch():
    cfg = dict(n=24, xmax=0.04, ymax=0.04, zmax=0.04, albedo=0.9, hgg=0.0, flags=1)
    rk = np.zeros((26, 26, 26), order="F")
    rk[1:-1, 1:-1, 1:-1] = 60.0
    _replay_case(cfg, 20000, rank=5, cap_per_packet=400, rhokap=rk)


def test_replay_ablated_crater_and_varying_grid():
    import tamc

    rk = next(iter(list(tamc.configs.crater_sequence(80, 6))[5:]))
    cfg = dict(tamc.configs.CONFIGS["shipped80"])
    _replay_case(cfg, 50000, rank=3, rhokap=rk)
    # non-cubic voxels, anisotropic extents, heterogeneous opacity, scatter on
    n = 30
    cfg = dict(n=n, xmax=0.02, ymax=0.035, zmax=0.05, albedo=0.85, hgg=0.7, flags=1)
    ii, jj, kk = np.meshgrid(*[np.arange(1, n + 1)] * 3, indexing="ij")
    rk = np.zeros((n + 2,) * 3, order="F")
    rk[1:-1, 1:-1, 1:-1] = 40.0 + 5.0 * ((ii + 2 * jj + 3 * kk) % 7)
    rk[12:18, 12:18, n - 3:n + 1] = 0.0
    _replay_case(cfg, 20000, rank=7, cap_per_packet=600, rhokap=rk)


def test_replay_thin_slab_exits():
    cfg = dict(n=10, xmax=0.05, ymax=0.05, zmax=0.05, albedo=0.0, hgg=0.9, flags=0)
    rk = np.zeros((12, 12, 12), order="F")
    rk[1:-1, 1:-1, 1:-1] = 20.0
    _replay_case(cfg, 30000, rank=4, rhokap=rk)


def test_replay_empty_and_short_draw_lists():
    import tamc

    cfg = tamc.configs.CONFIGS["shipped80"]
    t = make_transport(cfg)
    rec, jm = t.run_replay(np.zeros(1, dtype=np.int64), np.zeros(0))
    assert rec.size == 0 and not jm.any()
    o = make_oracle(cfg)
    o.seed_ran2(0)
    out = o.run(100, records=True, draws_cap=400)
    off = out["offsets"].copy()
    off[50:] -= 1                                   # packet 49 gets 3 draws instead of 4
    with pytest.raises(tamc.TamcError) as e:
        t.run_replay(off, out["draws"][:-1])
    assert e.value.code == 6                        # TAMC_EREPLAY
    t.close()


def test_replay_full_scale_chunk_homog200():
    """BASELINE config 2(i): one 5e6-packet chunk of the full-scale replay on the 200^3 grid (20e6 ran2 draws
    streamed through the C ABI in 4M-packet pieces)."""
    import tamc

    worst, gerr = _replay_case(tamc.configs.CONFIGS["homog200"], 5_000_000, rank=6)
    assert worst < 1e-9 and gerr < 1e-10


@pytest.mark.slow
def test_replay_full_scale_streamed_1e8_homog200():
    """BASELINE config 2(i) as SURVEY 8(d) states it: the homogeneous 200^3 cube, 1e8 packets, 4e8 ran2 draws (3.2 GB)
    generated sequentially by the oracle -- ONE rank's stream, continued from chunk to chunk -- and streamed through the C
    ABI in chunks of 1e7 packets.  Per chunk: integer fields exact, floating fields within the north-star's 1e-6 (measured
    ~1e-9), counters equal; at the end the sum of the ten device grids against the oracle's accumulated grid."""
    import tamc

    cfg = tamc.configs.CONFIGS["homog200"]
    rk = cfg["rhokap"]()
    o = make_oracle(cfg, rk)
    o.seed_ran2(3)
    t = make_transport(cfg, rk)
    scale = {"xp": cfg["xmax"], "yp": cfg["ymax"], "zp": cfg["zmax"], "nxp": 1.0, "nyp": 1.0, "nzp": 1.0}
    chunk, chunks = 10_000_000, 10
    total = np.zeros((200, 200, 200), order="F")
    steps = 0
    worst = 0.0
    for c in range(chunks):
        out = o.run(chunk, records=True, draws_cap=4 * chunk)          # the ran2 state carries over: one sequential stream
        rec, jm = t.run_replay(out["offsets"], out["draws"])
        st = t.get_stats()
        worst = max(worst, compare_records(rec, out["records"], scale=scale))
        assert st["packets"] == chunk and st["voxel_steps"] == out["stats"]["voxel_steps"]
        assert st["absorbed"] == out["stats"]["absorbed"] == chunk
        total += jm
        steps += st["voxel_steps"]
        del out, rec
    t.close()
    assert worst < 1e-8
    compare_grids(total, o.jmean, rtol=1e-9)                           # the oracle's tally accumulated over the ten chunks
    assert abs(steps / 1e8 - 1 / (1 - np.exp(-680.0 * 0.12 / 200))) < 1e-3
    assert abs(total.sum() / 1e8 - 1.0) < 5e-4                         # E[tau] = 1 per packet


def test_replay_phantom400_subset():
    """BASELINE config 4 (400^3, albedo 0.999, ~750 voxel-steps and ~300 scatterings per packet): a small
    subset replayed on the full-size grid (grids of 0.5 GB each exceed L2)."""
    import tamc

    _replay_case(tamc.configs.CONFIGS["phantom400"], 300, rank=2, cap_per_packet=60000)


@pytest.mark.parametrize("dims", [(1, 1, 1), (3, 5, 7), (17, 4, 33)])
def test_replay_non_cubic_and_degenerate_grids(dims):
    """nxg != nyg != nzg (the reference fixes 80^3 at compile time; the library takes any box) and the one-voxel grid."""
    import tamc
    from oracle import oracle as orc

    nx, ny, nz = dims
    xmax, ymax, zmax = 0.02, 0.03, 0.05
    rk = np.zeros((nx + 2, ny + 2, nz + 2), order="F")
    ii, jj, kk = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
    rk[1:-1, 1:-1, 1:-1] = 15.0 + 4.0 * ((ii + 2 * jj + 3 * kk) % 5)
    npk = 20000
    o = orc.Oracle(nx, ny, nz, xmax, ymax, zmax)
    o.set_rhokap(rk)
    o.set_optics(0.9, 0.6)
    o.set_flags(orc.FLAG_SCATTER)
    o.seed_ran2(9)
    out = o.run(npk, records=True, draws_cap=npk * 400)
    t = tamc.MCTransport(nx, ny, nz, xmax, ymax, zmax)
    t.set_optics(rk, 0.9, 0.6, flags=tamc.SCATTER)
    rec, jm = t.run_replay(out["offsets"], out["draws"])
    scale = {"xp": xmax, "yp": ymax, "zp": zmax, "nxp": 1.0, "nyp": 1.0, "nzp": 1.0}
    compare_records(rec, out["records"], scale=scale)
    cfgish = dict(n=max(dims), xmax=xmax, ymax=ymax, zmax=zmax)
    compare_grids(jm, o.jmean, rtol=1e-6, dep_scale=voxel_tau(cfgish, rk))
    # production kernels on the same box: counters conserve packets, every variant agrees with the others
    grids = []
    for variant in (0, 1, 2, 3):
        t.set_option("variant", variant)
        t.run_async(npk, 3, 0)
        grids.append(t.get_jmean())
        st = t.get_stats()
        assert st["packets"] == npk == st["absorbed"] + sum(st["exits"])
    for g in grids[1:]:
        compare_grids(g, grids[0], rtol=2e-2, dep_scale=voxel_tau(cfgish, rk), sum_rtol=1e-5)
    t.close()
