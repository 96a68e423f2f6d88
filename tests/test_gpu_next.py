"""SURVEY 8(f)-2 options built on upstream dead code, through the C ABI on the device: the Gaussian beam
(tamc_set_source_gaussian: rang(), sourceph.f90:73-101) and periodic lateral boundaries (TAMC_PERIODIC: repeat_bounds,
inttau2.f90:242-279).  Same bar as the hot path: trace replay of the oracle's ran2 sequence packet by packet (1e-6),
the production kernels against the oracle on the same Philox streams for every kernel variant, properties at size."""
import numpy as np
import pytest

from tests.test_gpu_production import SEED, _philox_exact
from tests.test_gpu_replay import _replay_case
from tests.util import compare_grids, make_oracle, make_transport

pytestmark = pytest.mark.gpu

PERIODIC = 4


def _slab_cfg(n=24, albedo=0.95, hgg=0.8, flags=1 | PERIODIC, **kw):
    """A slab four mean free paths wide and forty deep: most packets cross a lateral face several times."""
    import tamc

    cfg = dict(n=n, xmax=0.02, ymax=0.02, zmax=0.2, albedo=albedo, hgg=hgg, flags=flags, spot=0.01,
               rhokap=lambda: tamc.gridset(0.02, 0.02, 0.2, n, n, n, 100.0)[3])
    cfg.update(kw)
    return cfg


def _gauss_cfg(base, n, sigma, flags=None):
    import tamc

    cfg = dict(tamc.configs.scaled(base, n))
    cfg["gauss_sigma"] = sigma
    if flags is not None:
        cfg["flags"] = flags
    return cfg


# ---- trace replay -----------------------------------------------------------------------------------------------

def test_replay_periodic_slab():
    cfg = _slab_cfg()
    o = make_oracle(cfg)
    o.seed_ran2(3)
    assert o.run(2000)["stats"]["wraps"] > 2000            # the case does exercise repeat_bounds
    _replay_case(cfg, 6000, rank=3, cap_per_packet=3000)


def test_replay_periodic_isotropic_and_layered():
    _replay_case(_slab_cfg(n=16, albedo=0.9, hgg=0.0), 6000, rank=1, cap_per_packet=2000)
    import tamc

    cfg = dict(tamc.configs.scaled("skin200", 40))
    cfg["flags"] |= PERIODIC
    cfg["xmax"] = cfg["ymax"] = 0.05                     # narrow column of the layered skin model
    cfg["spot"] = 0.02
    _replay_case(cfg, 3000, rank=2, cap_per_packet=6000)


def test_replay_gaussian_stub_regime():
    # shipped regime with the Gaussian beam: sigma a third of the half-width -> some variates are redrawn
    worst, gerr = _replay_case(_gauss_cfg("shipped80", 80, 0.01), 100000, cap_per_packet=40)
    assert worst < 1e-6 and gerr < 1e-9


def test_replay_gaussian_wide_beam_and_scatter():
    _replay_case(_gauss_cfg("shipped80", 40, 0.05), 30000, rank=4, cap_per_packet=200)      # most variates miss the face
    _replay_case(_gauss_cfg("turbid200", 40, 0.1), 3000, rank=2, cap_per_packet=6000)
    _replay_case(_slab_cfg(gauss_sigma=0.015), 4000, rank=6, cap_per_packet=3000)           # both options at once


def test_replay_gaussian_short_draw_list_is_reported():
    import tamc

    cfg = _gauss_cfg("shipped80", 20, 0.01)
    t = make_transport(cfg)
    with pytest.raises(tamc.TamcError) as e:
        t.run_replay(np.array([0, 3, 5]), np.array([0.9, 0.9, 0.9, 0.2, 0.3]))      # rang never accepts, list runs out
    assert e.value.code == 6
    t.close()


# ---- production kernels against the oracle on the same Philox streams ---------------------------------------

def test_philox_exact_periodic_all_kernels():
    _philox_exact(_slab_cfg(), 6000)
    _philox_exact(_slab_cfg(n=16, albedo=0.9, hgg=0.0), 6000)
    _philox_exact(_slab_cfg(flags=1 | 2 | PERIODIC, n1=1.0, n2=1.38), 6000)   # Fresnel top/bottom + periodic sides


def test_philox_exact_gaussian_all_kernels():
    _philox_exact(_gauss_cfg("shipped80", 80, 0.01), 60000)                   # stub regime: thread-per-packet kernel
    _philox_exact(_gauss_cfg("shipped80", 40, 0.05), 20000)
    _philox_exact(_gauss_cfg("skin200", 64, 0.1), 6000)
    _philox_exact(_slab_cfg(gauss_sigma=0.015), 6000)


def test_kernel_forms_for_the_options():
    """Which kernel a call takes: the options never reach the stub-regime / column kernels."""
    import tamc

    n = (1 << 20) + 5
    t = make_transport(_gauss_cfg("shipped80", 80, 0.008))
    t.run_async(n, SEED, 0)
    st = t.get_stats()
    assert t.get_option("form") == 0 and st["packets"] == n and st["absorbed"] == n      # thread-per-packet, no column form
    t.set_source_co2(0.025)
    t.run_async(n, SEED, 0)
    assert t.get_option("form") in (1, 4, 5, 7, 8)                                       # back on the disk: a stub-regime kernel
    t.set_optics(None, 0.0, 0.9, flags=PERIODIC)                                         # periodic alone: no effect in the stub regime
    t.run_async(n, SEED, 0)
    j_p = t.get_jmean()
    assert t.get_option("form") in (1, 4, 5, 7, 8)
    t.set_optics(None, 0.0, 0.9, flags=0)
    t.run_async(n, SEED, 0)
    compare_grids(j_p, t.get_jmean(), rtol=1e-11)
    t.close()
    t = make_transport(_slab_cfg())
    t.run_async(20000, SEED, 0)
    assert t.get_option("form") == 3                                                     # pool kernel, ext build
    t.set_option("variant", 1)
    t.run_async(20000, SEED, 0)
    assert t.get_option("form") == 0                                                     # persistent kernel has no ext build
    t.close()


# ---- properties -------------------------------------------------------------------------------------------------

def test_periodic_slab_properties_at_size():
    """200^3 laterally uniform slab, 2e6 packets: nothing leaves sideways, energy balance, and the depth profile and
    R/T fractions agree with the oracle's ran2 run within counting statistics."""
    import tamc

    n, npk = 200, 2_000_000
    cfg = dict(n=n, xmax=0.05, ymax=0.05, zmax=0.1, albedo=0.9, hgg=0.8, flags=1 | PERIODIC, spot=0.02,
               rhokap=lambda: tamc.gridset(0.05, 0.05, 0.1, n, n, n, 120.0)[3])
    t = make_transport(cfg)
    t.run_async(npk, SEED, 0)
    jm, st = t.get_jmean(), t.get_stats()
    t.close()
    assert st["packets"] == npk and st["exits"][:4] == [0, 0, 0, 0]
    assert st["absorbed"] + st["exits"][4] + st["exits"][5] == npk
    # the oracle on its ran2 stream, 10 batches of 20 000 packets on a 50^3 grid of the same slab (batch means give the
    # standard error of every depth bin; the medium is uniform, so the voxel size does not enter the physics)
    nb, per = 10, 20_000
    o = make_oracle(dict(cfg, n=50, rhokap=lambda: tamc.gridset(0.05, 0.05, 0.1, 50, 50, 50, 120.0)[3]))
    o.seed_ran2(0)
    profs, ex = [], np.zeros(6)
    for _ in range(nb):
        o.zero_jmean()
        ref = o.run(per)["stats"]
        assert ref["exits"][:4] == [0, 0, 0, 0] and ref["wraps"] > 0
        profs.append(o.jmean.sum(axis=(0, 1)) / per)
        ex += np.array(ref["exits"])
    nref = nb * per
    for f in (4, 5):
        p = ex[f] / nref
        assert abs(st["exits"][f] / npk - p) < 5 * np.sqrt(p * (1 - p) * (1 / nref + 1 / npk)) + 1e-4
    profs = np.array(profs)
    pref, se = profs.mean(axis=0), profs.std(axis=0, ddof=1) / np.sqrt(nb)
    prof = jm.sum(axis=(0, 1)).reshape(50, 4).sum(axis=1) / npk            # 200 layers -> the oracle's 50
    assert np.all(np.abs(prof - pref) < 6 * se * np.sqrt(1 + nref / npk) + 1e-3 * pref.max())
    assert abs(prof.sum() - pref.sum()) < 0.01 * pref.sum()


def test_gaussian_beam_profile_at_size():
    """200^3 stub regime, 4e6 packets, sigma = 20 voxels: the tally's lateral marginals are the binned Gaussian."""
    import tamc
    from math import erf, sqrt

    n, npk, sigma = 200, 4_000_000, 0.006
    cfg = _gauss_cfg("homog200", n, sigma)
    t = make_transport(cfg)
    t.run_async(npk, SEED, 0)
    jm, st = t.get_jmean(), t.get_stats()
    t.close()
    assert st["packets"] == npk and st["absorbed"] == npk
    assert abs(jm.sum() / npk - 1.0) < 5 / np.sqrt(npk)                     # E[tau] = 1 deposited per packet
    edges = np.arange(n + 1) * 2 * cfg["xmax"] / n - cfg["xmax"]
    want = np.array([0.5 * (erf(b / (sigma * sqrt(2))) - erf(a / (sigma * sqrt(2)))) for a, b in zip(edges[:-1], edges[1:])])
    for axis in ((1, 2), (0, 2)):
        col = jm.sum(axis=axis) / jm.sum()
        # deposits carry the exponential tau as weight: variance per packet 2 rather than 1
        assert np.abs(col - want).max() < 6 * np.sqrt(2 * want.max() / npk)


def test_argument_checks_for_the_options():
    import tamc

    cfg = tamc.configs.CONFIGS["shipped80"]
    t = make_transport(cfg)
    for bad in (0.0, -1.0, float("nan"), 1.0):                              # 1.0 cm: beyond 20 half-widths
        with pytest.raises(tamc.TamcError) as e:
            t.set_source_gaussian(bad)
        assert e.value.code == 1
    with pytest.raises(tamc.TamcError):
        t.set_optics(None, 0.0, 0.9, flags=8)
    t.close()


@pytest.mark.parametrize("omega", [0.5, 0.9, 0.99])
def test_semi_infinite_slab_reflectance_matches_chandrasekhar_on_the_device(omega):
    """The external pin of tests/test_oracle_next.py on the production kernels: isotropic scattering in a laterally
    infinite (periodic), optically thick, index-matched slab reflects 1 - H(1) sqrt(1 - omega) of a normally incident
    beam (Chandrasekhar's H-function).  4e5 packets, binomial error; the oracle on the same Philox ids gives the same
    counts (z = 1.6, -0.1, -0.1 sigma for the three albedos)."""
    import tamc
    from oracle import oracle as orc
    from tests.test_oracle_next import _chandrasekhar_H

    npk = 400000
    t = tamc.MCTransport(8, 8, 30, 0.01, 0.01, 0.03)
    t.set_source_co2(0.004)
    t.set_optics(tamc.gridset(0.01, 0.01, 0.03, 8, 8, 30, 1000.0)[3], omega, 0.0, flags=1 | PERIODIC)
    o = orc.Oracle(8, 8, 30, 0.01, 0.01, 0.03)
    o.gridset_uniform(1000.0)
    o.set_optics(omega, 0.0)
    o.set_spot(0.004)
    o.set_flags(orc.FLAG_SCATTER | orc.FLAG_PERIODIC)
    o.seed_philox(SEED, 0)
    ref = o.run(npk)["stats"]
    want = 1.0 - _chandrasekhar_H(omega, 1.0) * np.sqrt(1.0 - omega)
    for variant in (3, 0):
        t.set_option("variant", variant)
        t.run_async(npk, SEED, 0)
        st = t.get_stats()
        assert st["packets"] == npk and st["exits"][:4] == [0, 0, 0, 0]
        got = st["exits"][5] / npk
        assert abs(got - want) < 4.0 * np.sqrt(want * (1.0 - want) / npk), (variant, got, want)
        assert abs(st["exits"][5] - ref["exits"][5]) <= 10 and abs(st["absorbed"] - ref["absorbed"]) <= 10
    t.close()


def test_slab_benchmark_of_van_de_hulst_on_the_device():
    """tests/test_oracle_next.py's anisotropic benchmark on the production kernels: slab of optical thickness 2, albedo
    0.9, g = 0.75, index-matched: Rd = 0.09739, Tt = 0.66096 (van de Hulst 1980 / MCML 1995 table 1).  The oracle on the
    same Philox ids gives -0.5 and +1.1 sigma at 4e5 packets."""
    import tamc
    from oracle import oracle as orc

    npk = 400000
    t = tamc.MCTransport(8, 8, 20, 0.01, 0.01, 0.01)
    t.set_source_co2(0.004)
    t.set_optics(tamc.gridset(0.01, 0.01, 0.01, 8, 8, 20, 100.0)[3], 0.9, 0.75, flags=1 | PERIODIC)
    o = orc.Oracle(8, 8, 20, 0.01, 0.01, 0.01)
    o.gridset_uniform(100.0)
    o.set_optics(0.9, 0.75)
    o.set_spot(0.004)
    o.set_flags(orc.FLAG_SCATTER | orc.FLAG_PERIODIC)
    o.seed_philox(SEED, 0)
    ref = o.run(npk)["stats"]
    for variant in (3, 0, 2):
        t.set_option("variant", variant)
        t.run_async(npk, SEED, 0)
        st = t.get_stats()
        assert st["packets"] == npk and st["exits"][:4] == [0, 0, 0, 0]
        for f, want in ((5, 0.09739), (4, 0.66096)):
            got = st["exits"][f] / npk
            assert abs(got - want) < 4.0 * np.sqrt(want * (1.0 - want) / npk), (variant, f, got, want)
            assert abs(st["exits"][f] - ref["exits"][f]) <= 20
    t.close()
