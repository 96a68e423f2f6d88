"""Per-voxel albedo / hgg / refractive-index grids (tamc_set_optics_grids; EXTENSION, no upstream counterpart: the
reference's optics are rhokap per voxel + the scalars of opt_prop.f90:5).

CPU: the extended oracle -- uniform grids reproduce the scalar run bit for bit; a two-layer albedo grid absorbs where it
should.  GPU: trace replay, the exact / thread-per-packet kernels packet by packet and the production (flight) kernel on
the oracle's own Philox streams; NULL grids leave results bit-identical."""
import numpy as np
import pytest

from tests.util import compare_grids, compare_records, make_oracle, make_transport, voxel_tau

SEED = 20261017


def _cfg(n=40):
    import tamc

    return dict(tamc.configs.scaled("skin200", n))


def _two_layer(cfg, top, bottom, frac=0.2):
    """value `top` in the upper `frac` of the grid (high k), `bottom` below; halo included (never read)."""
    n = cfg["n"]
    a = np.full((n + 2, n + 2, n + 2), bottom, dtype=np.float64, order="F")
    a[:, :, int((1 - frac) * n) + 1:] = top
    return a


def test_oracle_uniform_grids_equal_scalars_bitwise():
    cfg = _cfg(24)
    n = 3000
    o = make_oracle(cfg)
    o.seed_ran2(0)
    a = o.run(n, records=True)
    ja = o.jmean.copy()
    o2 = make_oracle(cfg)
    shape = o2.rhokap.shape
    o2.set_grids(np.full(shape, cfg["albedo"], order="F"), np.full(shape, cfg["hgg"], order="F"), None)
    o2.seed_ran2(0)
    b = o2.run(n, records=True)
    assert np.array_equal(ja, o2.jmean) and a["stats"] == b["stats"]
    for f in a["records"].dtype.names:
        assert np.array_equal(a["records"][f], b["records"][f]), f


def test_oracle_two_layer_albedo_and_isotropic_layer():
    cfg = _cfg(24)
    n = 4000
    # (a) albedo 0 in the top layer: every packet is absorbed at its first interaction if that lies in the top layer
    o = make_oracle(cfg)
    o.set_grids(_two_layer(cfg, 0.0, cfg["albedo"], frac=0.5), None, None)
    o.seed_ran2(0)
    r = o.run(n, records=True)
    rec = r["records"]
    top = rec["zcell"] > 12
    first = rec["nscatt"] == 0
    assert (first & (rec["fate"] == 0)).sum() > 0.9 * n            # optical depth of the top half >> 1
    assert np.all(rec["zcell"][first & (rec["fate"] == 0)] > 12)
    # (b) an isotropic (hgg = 0) grid scatters back more than g = 0.9
    back = []
    for hg in (0.9, 0.0):
        o = make_oracle(cfg)
        o.set_grids(None, np.full(o.rhokap.shape, hg, order="F"), None)
        o.seed_ran2(0)
        back.append(o.run(n)["stats"]["exits"][5])
    assert back[1] > 1.5 * back[0]


@pytest.mark.gpu
def test_gpu_grids_replay_exact_and_production_against_oracle():
    cfg = _cfg(48)
    rk = cfg["rhokap"]()
    alb = _two_layer(cfg, 0.9, 0.995, frac=0.3)
    hgg = _two_layer(cfg, 0.5, 0.9, frac=0.3)
    n = 6000
    scale = {"xp": cfg["xmax"], "yp": cfg["ymax"], "zp": cfg["zmax"], "nxp": 1.0, "nyp": 1.0, "nzp": 1.0}
    # ---- trace replay on the reference's ran2 stream
    o = make_oracle(cfg, rk)
    o.set_grids(alb, hgg, None)
    o.seed_ran2(0)
    want = o.run(3000, records=True, draws_cap=4_000_000)
    t = make_transport(cfg, rk)
    t.set_optics_grids(alb, hgg, None)
    rec, jm = t.run_replay(want["offsets"], want["draws"])
    compare_records(rec, want["records"], rtol=1e-6, scale=scale)
    compare_grids(jm, o.jmean, rtol=1e-6, dep_scale=voxel_tau(cfg, rk))
    # ---- Philox streams: exact arithmetic packet by packet, production kernels on the grid and the counters
    o = make_oracle(cfg, rk)
    o.set_grids(alb, hgg, None)
    o.seed_philox(SEED, 0)
    want = o.run(n, records=True)
    for variant in (2, 0):
        t.set_option("variant", variant)
        rec, jm = t.run_records(n, SEED, 0)
        if variant == 2:
            compare_records(rec, want["records"], rtol=1e-6, scale=scale)
            compare_grids(jm, o.jmean, rtol=1e-6, dep_scale=voxel_tau(cfg, rk))
        else:
            compare_records(rec, want["records"], rtol=1e-4, scale=scale, flip_fraction=1e-3, p99_rtol=5e-6)
            compare_grids(jm, o.jmean, rtol=2e-2, dep_scale=voxel_tau(cfg, rk), sum_rtol=1e-5)
    for variant, form in ((3, 9), (1, 0), (0, 0)):
        t.set_option("variant", variant)
        t.run_async(n, SEED, 0)
        jm, st = t.get_jmean(), t.get_stats()
        assert t.get_option("form") == form, (variant, t.get_option("form"))
        compare_grids(jm, o.jmean, rtol=2e-2, dep_scale=voxel_tau(cfg, rk), sum_rtol=1e-5)
        assert st["scatters"] == want["stats"]["scatters"] and st["absorbed"] == want["stats"]["absorbed"]
        assert st["exits"] == want["stats"]["exits"]
        assert abs(st["voxel_steps"] - want["stats"]["voxel_steps"]) <= 12
    # the grids matter: the scalar run differs
    t.set_optics_grids(None, None, None)
    t.set_option("variant", 3)
    t.run_async(n, SEED, 0)
    plain, st_plain = t.get_jmean(), t.get_stats()
    assert st_plain["scatters"] != want["stats"]["scatters"]
    # ... and uniform grids equal to the scalars reproduce the scalar run (same kernel arithmetic, other build)
    shape = rk.shape
    t.set_optics_grids(np.full(shape, cfg["albedo"], order="F"), np.full(shape, cfg["hgg"], order="F"), None)
    t.run_async(n, SEED, 0)
    uni, st_uni = t.get_jmean(), t.get_stats()
    for k in ("scatters", "absorbed", "exits", "voxel_steps"):
        assert st_uni[k] == st_plain[k], k
    compare_grids(uni, plain, rtol=1e-11)
    t.close()


@pytest.mark.gpu
def test_gpu_refractive_index_grid_with_fresnel():
    """n grid + TAMC_FRESNEL: uniform n grid == scalar n2 packet by packet; a two-region grid against the extended oracle."""
    cfg = _cfg(40)
    cfg["flags"] = cfg["flags"] | 2
    cfg["n1"], cfg["n2"] = 1.0, 1.38
    rk = cfg["rhokap"]()
    n = 5000
    shape = rk.shape
    ngrid = np.full(shape, 1.38, order="F")
    ngrid[: shape[0] // 2, :, :] = 1.5                      # the -x half is optically denser
    o = make_oracle(cfg, rk)
    o.set_grids(None, None, ngrid)
    o.seed_philox(SEED, 0)
    want = o.run(n, records=True)
    scale = {"xp": cfg["xmax"], "yp": cfg["ymax"], "zp": cfg["zmax"], "nxp": 1.0, "nyp": 1.0, "nzp": 1.0}
    t = make_transport(cfg, rk)
    t.set_optics_grids(None, None, ngrid)
    t.set_option("variant", 2)
    rec, jm = t.run_records(n, SEED, 0)
    compare_records(rec, want["records"], rtol=1e-5, scale=scale)
    for variant in (0, 3):
        t.set_option("variant", variant)
        t.run_async(n, SEED, 0)
        st = t.get_stats()
        assert st["specular"] == want["stats"]["specular"] and st["exits"] == want["stats"]["exits"]
        assert st["internal_reflections"] == want["stats"]["internal_reflections"]
    # uniform n grid == scalar n2
    t.set_optics_grids(None, None, np.full(shape, 1.38, order="F"))
    t.set_option("variant", 2)
    a, _ = t.run_records(n, SEED, 0)
    t.set_optics_grids(None, None, None)
    b, _ = t.run_records(n, SEED, 0)
    for f in a.dtype.names:
        assert np.array_equal(a[f], b[f]), f
    t.close()


@pytest.mark.gpu
def test_gpu_grids_argument_validation():
    import tamc

    cfg = _cfg(16)
    t = make_transport(cfg)
    bad = np.full(t.rhokap_shape, 1.5, order="F")
    with pytest.raises(tamc.TamcError) as e:
        t.set_optics_grids(bad, None, None)
    assert e.value.code == 1
    with pytest.raises(ValueError):
        t.set_optics_grids(np.zeros((3, 3, 3)), None, None)
    t.close()
