"""Production (Philox) kernels through the C ABI.

 (1) exact: the oracle run on the same counter-based stream must give the same packets and grid;
 (2) statistical: against the oracle on the reference's own ran2 stream, per-voxel z-scores from
     batch means and chi-square over the grid (north_star: 3 sigma per voxel);
 (3) properties that hold at BASELINE.json's full sizes.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.util import compare_grids, compare_records, make_oracle, make_transport, voxel_tau

pytestmark = pytest.mark.gpu

SEED = 20261017


def _philox_exact(cfg, n, first=0, rhokap=None):
    rk = rhokap if rhokap is not None else cfg["rhokap"]()
    o = make_oracle(cfg, rk)
    o.seed_philox(SEED, first)
    want = o.run(n, records=True)
    t = make_transport(cfg, rk)
    scale = {"xp": cfg["xmax"], "yp": cfg["ymax"], "zp": cfg["zmax"], "nxp": 1.0, "nyp": 1.0, "nzp": 1.0}
    scat = bool(cfg["flags"] & 1)
    # Tolerances.  Exact arithmetic (variant 2): the north_star's 1e-6.  Production arithmetic
    # (tamc_fast.cuh) rotates the direction vector instead of tracking phi, so it does not reproduce one
    # artefact of the reference: `phi -+ TWOPI` with the truncated TWOPI = 6.283185 turns the azimuth by
    # 3.07e-7 rad at every wrap (stokes.f90:102-103).  After tens of scatterings that is ~1e-7..1e-5 of
    # the box size per packet: bounded here at 1e-4 (max) and 5e-6 (99th percentile), with rare edge flips.
    # (with TAMC_FRESNEL a reflected path re-enters voxels it already crossed: a few more sliver deposits per voxel)
    gk_exact = dict(rtol=1e-5 if cfg["flags"] & 2 else 1e-6, dep_scale=voxel_tau(cfg, rk)) if scat else dict(rtol=1e-10)
    # (a packet displaced by 1e-5 of the box changes its chord in a voxel by ~1e-3 of that voxel's optical depth)
    gk_fast = dict(rtol=2e-2, dep_scale=voxel_tau(cfg, rk), sum_rtol=1e-5) if scat else dict(rtol=1e-10)
    for variant in (2, 1):
        t.set_option("variant", variant)
        rec, jm = t.run_records(n, SEED, first)
        if variant == 2 or not scat:
            compare_records(rec, want["records"], rtol=1e-6, scale=scale)
            compare_grids(jm, o.jmean, **gk_exact)
        else:
            compare_records(rec, want["records"], rtol=1e-4, scale=scale, flip_fraction=1e-3, p99_rtol=5e-6)
            compare_grids(jm, o.jmean, **gk_fast)
    # both kernel shapes and both tally policies give the same grid and counters
    for variant in (0, 1, 2, 3):
        for merge in (0, 1):
            for thr in ((32, 1), (0, 16), (64, 32)):
                if variant not in (1, 3) and thr != (32, 1):
                    continue
                t.set_option("block", 128 if (variant == 3 and thr[1] == 16) else 256)
                t.set_option("tile", 3 if thr[1] == 16 else (0 if thr[1] == 1 else -1))   # stub regime: smem tally tile
                t.set_option("variant", variant)
                t.set_option("merge", merge)
                t.set_option("chunk", thr[0])
                t.set_option("scatter_min", thr[1])
                t.run_async(n, SEED, first)
                jm2 = t.get_jmean()
                st = t.get_stats()
                compare_grids(jm2, o.jmean, **(gk_exact if variant == 2 else gk_fast))
                if variant != 2:
                    compare_grids(jm2, jm, **gk_exact)      # device vs device, same arithmetic: summation order only
                assert st["packets"] == n
                slack = 0 if (variant == 2 or not scat) else 2 * max(2, int(1e-3 * n))   # edge flips, see compare_records
                assert abs(st["voxel_steps"] - want["stats"]["voxel_steps"]) <= slack
                assert st["scatters"] == want["stats"]["scatters"]
                assert st["absorbed"] == want["stats"]["absorbed"]
                assert st["exits"] == want["stats"]["exits"]
    t.close()


def test_philox_exact_shipped():
    import tamc

    _philox_exact(tamc.configs.CONFIGS["shipped80"], 125000)


def test_philox_exact_offset_ids_cross_2_32():
    import tamc

    _philox_exact(tamc.configs.CONFIGS["shipped80"], 30000, first=(1 << 32) - 10000)


def test_philox_exact_turbid():
    import tamc

    _philox_exact(tamc.configs.scaled("turbid200", 60), 6000)


def test_philox_exact_skin_and_crater():
    import tamc

    _philox_exact(tamc.configs.scaled("skin200", 100), 6000)
    rk = list(tamc.configs.crater_sequence(80, 5))[4]
    _philox_exact(tamc.configs.CONFIGS["shipped80"], 40000, rhokap=rk)


def test_partition_invariance_and_cursor():
    """Packets keyed by global id: any split over calls / GPUs sums to the same grid."""
    import tamc

    cfg = tamc.configs.scaled("skin200", 64)
    t = make_transport(cfg)
    n = 40000
    t.run_async(n, SEED, 0)
    whole = t.get_jmean()
    parts = np.zeros_like(whole)
    for lo, hi in ((0, 10000), (10000, 10001), (10001, 40000)):
        t.run_async(hi - lo, SEED, lo)
        parts += t.get_jmean()
    compare_grids(parts, whole, rtol=1e-10)
    # tamc_run advances the cursor: two calls of n/2 == one call of n from id 0
    t.seek(0)
    a, sa = t.run(n // 2, SEED)
    b, sb = t.run(n // 2, SEED)
    assert not np.array_equal(a, b)
    compare_grids(a + b, whole, rtol=1e-10)
    assert sa["packets"] == sb["packets"] == n // 2
    assert sa["gpu_launches"] >= 1 and sa["kernel_ms"] > 0
    t.close()


def _batch_grids_gpu(cfg, batches, per_batch):
    t = make_transport(cfg)
    out = []
    for b in range(batches):
        t.run_async(per_batch, SEED, b * per_batch)
        out.append(t.get_jmean().copy())
    t.close()
    return np.stack(out)


def _batch_grids_oracle(cfg, batches, per_batch):
    out = []
    for b in range(batches):
        o = make_oracle(cfg)
        o.seed_ran2(b)                    # one emulated MPI rank per batch (mcpolar.f90:97-98)
        o.run(per_batch)
        out.append(o.jmean.copy())
    return np.stack(out)


def _chi_square(a, b, min_mean):
    """Two sets of batch grids -> per-voxel z from batch means, chi2/dof over well-populated voxels."""
    ma, mb = a.mean(0), b.mean(0)
    va, vb = a.var(0, ddof=1) / a.shape[0], b.var(0, ddof=1) / b.shape[0]
    sel = (ma > min_mean) & (mb > min_mean) & (va + vb > 0)
    z = (ma[sel] - mb[sel]) / np.sqrt(va[sel] + vb[sel])
    return z, float((z ** 2).mean())


@pytest.mark.parametrize("name,n,per_batch,min_mean", [("shipped80", 80, 60000, 20.0), ("skin200", 50, 12000, 15.0)])
def test_chi_square_against_ran2_oracle(name, n, per_batch, min_mean):
    import tamc

    cfg = tamc.configs.scaled(name, n)
    K = 16
    g = _batch_grids_gpu(cfg, K, per_batch)
    c = _batch_grids_oracle(cfg, K, per_batch)
    z, chi2 = _chi_square(g, c, min_mean)
    assert z.size > 200
    # z follows Student-t with ~2(K-1)=30 dof: var = 30/28 = 1.07; chi2/dof within 5 sigma of that
    assert abs(chi2 - 1.07) < 5 * np.sqrt(2.5 / z.size) + 0.05, (chi2, z.size)
    # 3-sigma criterion: the fraction beyond 3 sigma stays near the expected 0.5 % (t_30), never a bulk shift
    assert (np.abs(z) > 3).mean() < 0.012
    assert abs(z.mean()) < 5 / np.sqrt(z.size)
    # total absorbed energy per packet agrees within 4 sigma
    ta, tb = g.sum(axis=(1, 2, 3)), c.sum(axis=(1, 2, 3))
    assert abs(ta.mean() - tb.mean()) < 4 * np.sqrt(ta.var(ddof=1) / K + tb.var(ddof=1) / K)


def test_fractions_reflected_absorbed_transmitted():
    """Totals by fate (north_star: reflected/absorbed/transmitted fractions) vs the ran2 oracle."""
    import tamc

    cfg = tamc.configs.scaled("skin200", 40)
    n = 200000
    t = make_transport(cfg)
    t.run_async(n, SEED, 0)
    st = t.get_stats()
    t.close()
    o = make_oracle(cfg)
    o.seed_ran2(0)
    ref = o.run(n)["stats"]
    assert st["absorbed"] + sum(st["exits"]) == n
    for got, want in [(st["absorbed"], ref["absorbed"])] + list(zip(st["exits"], ref["exits"])):
        p = max(want, 1) / n
        assert abs(got - want) < 5 * np.sqrt(2 * n * p * (1 - p)) + 5
    assert abs(st["voxel_steps"] / ref["voxel_steps"] - 1) < 0.02
    assert abs(st["scatters"] / ref["scatters"] - 1) < 0.02


def test_full_size_homog200_properties():
    """BASELINE config 2 at full grid size, 2e7 packets: properties that need no oracle run."""
    import tamc

    cfg = tamc.configs.CONFIGS["homog200"]
    n = 20_000_000
    t = make_transport(cfg)
    jm, st = t.run(n, SEED)
    t.close()
    tc = 680.0 * (0.12 / 200)
    assert st["packets"] == n and st["absorbed"] == n and sum(st["exits"]) == 0
    assert abs(jm.sum() / n - 1.0) < 5 / np.sqrt(n)                       # E[tau] = 1
    assert abs(st["voxel_steps"] / n - 1 / (1 - np.exp(-tc))) < 5 * 2.5 / np.sqrt(n)
    ii, jj, kk = np.nonzero(jm)
    r = np.hypot(ii + 0.5 - 100.0, jj + 0.5 - 100.0)
    assert r.max() < 0.0125 / (0.06 / 200) + 0.7072                      # under the 0.025 cm disk
    layer = jm.sum(axis=(0, 1))[::-1]
    for k in range(1, 8):
        want = n * np.exp(-(k - 1) * tc) * (1 - np.exp(-tc))
        assert abs(layer[k - 1] / want - 1) < 6 / np.sqrt(want)
    # linearity in packet count: a second, disjoint id range gives a statistically identical grid
    t = make_transport(cfg)
    t.seek(n)
    jm2, _ = t.run(n, SEED)
    t.close()
    top = (jm[:, :, -1] > 0) & (jm2[:, :, -1] > 0)
    z = (jm[:, :, -1][top] - jm2[:, :, -1][top]) / np.sqrt(jm[:, :, -1][top] + jm2[:, :, -1][top])
    assert 0.2 < (z ** 2).mean() < 1.2          # deposit per hit <= tc < 1 -> variance below Poisson


def test_error_behaviour():
    import tamc

    t = tamc.MCTransport(8, 8, 8, 0.1, 0.1, 0.1)
    with pytest.raises(tamc.TamcError) as e:
        t.run(10, 1)
    assert e.value.code == 5                     # TAMC_ESTATE: optics not set
    rk = np.zeros((10, 10, 10), order="F")
    with pytest.raises(tamc.TamcError) as e:
        t.set_optics(rk, 1.5, 0.9)
    assert e.value.code == 1
    with pytest.raises(ValueError):
        t.set_optics(np.zeros((8, 8, 8)), 0.0, 0.9)
    with pytest.raises(tamc.TamcError):
        t.set_source_co2(1.0)                    # spot wider than the face
    t.set_optics(rk, 0.0, 0.9)
    jm, st = t.run(0, 1)
    assert st["packets"] == 0 and not jm.any()
    # rhokap == 0 everywhere: every packet crosses the grid with zero deposit and leaves through -z
    jm, st = t.run(1000, 1)
    assert not jm.any() and st["exits"][4] == 1000 and st["voxel_steps"] == 8000
    t.close()
    with pytest.raises(tamc.TamcError) as e:
        tamc.MCTransport(8, 8, 8, 0.1, 0.1, 0.1, delta=1e-30)
    assert e.value.code == 1
    with pytest.raises(tamc.TamcError) as e:
        tamc.MCTransport(8, 8, 8, 0.1, 0.1, 0.1, device=99)
    assert e.value.code == 2


# ---- EXTENSION (no upstream semantics): Fresnel boundaries, checked against the extended oracle --------------
def _fresnel_cfg(base, n, n1=1.0, n2=1.38):
    import tamc

    cfg = dict(tamc.configs.scaled(base, n))
    cfg["flags"] = cfg["flags"] | 2
    cfg["n1"], cfg["n2"] = n1, n2
    return cfg


def test_fresnel_philox_exact_all_kernels():
    _philox_exact(_fresnel_cfg("skin200", 64), 8000)
    _philox_exact(_fresnel_cfg("turbid200", 40), 4000)
    _philox_exact(_fresnel_cfg("shipped80", 80), 30000)          # stub regime + Fresnel (thread-per-packet kernel)


def test_fresnel_index_matched_is_a_no_op_and_counts():
    import tamc

    base = tamc.configs.scaled("skin200", 48)
    n = 60000
    t0 = make_transport(base)
    t0.set_option("flight", 0)          # the work-queue kernel, whose `ext` build carries the boundary options
    t0.run_async(n, SEED, 0)
    j0, s0 = t0.get_jmean(), t0.get_stats()
    t0.close()
    t1 = make_transport(_fresnel_cfg("skin200", 48, 1.38, 1.38))
    t1.run_async(n, SEED, 0)
    j1, s1 = t1.get_jmean(), t1.get_stats()
    t1.close()
    compare_grids(j1, j0, rtol=1e-10)                             # boundary draws come from their own stream
    assert s1["specular"] == 0 and s1["internal_reflections"] == 0 and s1["exits"] == s0["exits"]
    # air / tissue: specular fraction ((1-1.38)/2.38)^2 = 2.55 %, internal reflections keep more energy inside
    cfg = _fresnel_cfg("skin200", 48)
    t2 = make_transport(cfg)
    t2.run_async(n, SEED, 0)
    s2 = t2.get_stats()
    t2.close()
    r0 = ((1 - 1.38) / 2.38) ** 2
    assert abs(s2["specular"] / n - r0) < 5 * np.sqrt(r0 * (1 - r0) / n)
    assert s2["absorbed"] + sum(s2["exits"]) == n and s2["exits"][5] >= s2["specular"]
    assert s2["internal_reflections"] > 0 and s2["absorbed"] > s0["absorbed"]
    # and the same totals as the extended oracle on its ran2 stream, within counting statistics
    o = make_oracle(cfg)
    o.seed_ran2(0)
    ref = o.run(n)["stats"]
    for got, want in [(s2["absorbed"], ref["absorbed"]), (s2["specular"], ref["specular"])] + list(zip(s2["exits"], ref["exits"])):
        p = max(want, 1) / n
        assert abs(got - want) < 5 * np.sqrt(2 * n * p * (1 - p)) + 5
    assert abs(s2["internal_reflections"] / ref["internal_reflections"] - 1) < 0.05


def test_fresnel_not_replayable():
    import tamc

    t = make_transport(_fresnel_cfg("shipped80", 80))
    with pytest.raises(tamc.TamcError) as e:
        t.run_replay(np.array([0, 4]), np.full(4, 0.5))
    assert e.value.code == 1
    t.close()


def test_chi_square_full_size_layered_skin_200():
    """BASELINE config 3 at its full grid size: 16 x 250 000 packets on the device against 16 emulated MPI ranks x 250 000
    packets of the oracle on the reference's ran2 streams (host threads).  Per-voxel variance from the device batches
    (the same under the null hypothesis); z per voxel, chi-square over the well-populated voxels, 3-sigma fraction."""
    import tamc

    cfg = tamc.configs.CONFIGS["skin200"]
    K, per = 16, 250000
    rk = cfg["rhokap"]()
    ref = orc.run_ranks(K, 200, 200, 200, cfg["xmax"], cfg["ymax"], cfg["zmax"], rk, cfg["albedo"], cfg["hgg"], per,
                        flags=cfg["flags"])
    t = make_transport(cfg, rk)
    s1 = np.zeros((200, 200, 200), order="F")
    s2 = np.zeros_like(s1)
    steps = scat = 0
    for b in range(K):
        t.run_async(per, SEED, b * per)
        j = t.get_jmean()
        st = t.get_stats()
        steps += st["voxel_steps"]
        scat += st["scatters"]
        s1 += j
        s2 += j * j
    t.close()
    mean = s1 / K
    var = (s2 / K - mean * mean) * K / (K - 1)
    sel = (mean > 50.0) & (var > 0)                      # voxels with several hundred deposits per batch
    z = (ref["jmean"][sel] / K - mean[sel]) / np.sqrt(2.0 * var[sel] / K)
    assert z.size > 2000
    chi2 = float((z ** 2).mean())
    # the variance is itself estimated from 16 batches: E[z^2] = (K-1)/(K-3) = 1.15
    assert abs(chi2 - 1.15) < 0.12, (chi2, z.size)
    assert (np.abs(z) > 3).mean() < 0.02 and abs(z.mean()) < 5 / np.sqrt(z.size)
    assert abs(steps / ref["stats"]["voxel_steps"] - 1) < 0.01 and abs(scat / ref["stats"]["scatters"] - 1) < 0.01
    assert abs(s1.sum() / ref["jmean"].sum() - 1) < 0.01


def test_pool_kernel_with_early_opacity_fetch_at_400_cubed():
    """Grids beyond L2 (phantom400: BASELINE config 4).  The work-queue kernel ("flight" = 0) fetches the next voxel's
    opacity one loop pass early there: same production arithmetic as the persistent kernel, so on the same packet ids
    the counters are equal and the grid agrees to summation order."""
    import tamc

    cfg = tamc.configs.CONFIGS["phantom400"]
    t = make_transport(cfg)
    t.set_option("flight", 0)
    n = 4000
    res = {}
    for variant in (3, 1):
        t.set_option("variant", variant)
        t.run_async(n, SEED, 0)
        res[variant] = (t.get_jmean().copy(), t.get_stats(), t.get_option("form"))
    assert res[3][2] == 3 and res[1][2] == 1
    for key in ("packets", "voxel_steps", "scatters", "absorbed", "exits"):
        assert res[3][1][key] == res[1][1][key], key
    a, b = res[3][0], res[1][0]
    assert np.array_equal(a != 0, b != 0)
    nz = b != 0
    assert np.abs(a[nz] - b[nz]).max() <= 1e-9 * np.abs(b[nz]).max()
    assert res[3][1]["scatters"] > 100 * n
    t.close()


def test_flight_kernel_against_oracle_at_400_cubed():
    """BASELINE config 4 at its full grid size: the kernel a production call really runs there (the flight kernel, form 9)
    against the ORACLE on the same Philox streams -- counters by fate exact, voxel-steps within the edge-flip slack, the
    grid within the production tolerance of test_philox_exact (a deposit scale of one voxel's optical depth)."""
    import tamc

    cfg = tamc.configs.CONFIGS["phantom400"]
    rk = cfg["rhokap"]()
    n = 1500
    o = make_oracle(cfg, rk)
    o.seed_philox(SEED, 0)
    want = o.run(n)["stats"]
    t = make_transport(cfg, rk)
    t.run_async(n, SEED, 0)
    jm, st = t.get_jmean(), t.get_stats()
    assert t.get_option("form") == 9
    t.close()
    assert st["packets"] == n and st["scatters"] == want["scatters"] and st["absorbed"] == want["absorbed"]
    assert st["exits"] == want["exits"]
    assert abs(st["voxel_steps"] - want["voxel_steps"]) <= max(4, int(2e-6 * want["voxel_steps"]))
    # 1.1e6 voxel-steps through 190 scatterings per packet: a path that differs from the oracle's in the 9th digit passes a
    # voxel corner on the other side a few times, which moves a sliver of up to a few per cent of a voxel's optical depth
    # between two neighbours (the same effect test_philox_exact bounds at 2e-2 for its shorter walks); the totals agree to 1e-7
    compare_grids(jm, o.jmean, rtol=0.1, dep_scale=voxel_tau(cfg, rk), sum_rtol=1e-7)
