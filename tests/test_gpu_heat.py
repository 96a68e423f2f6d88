"""The heat / ablation step on the device against its oracle, through the C ABI.
Same statements in the same order on both sides and no FMA contraction: everything that does not pass
through exp() must agree bit for bit; the rest to ~1e-14."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

HALO = ("temp", "rhokap", "kappa", "density", "heatcap", "coeff", "alpha")
INNER = ("watercontent", "Q", "tissue")


def _pair(n, xmax=0.03, ymax=0.03, zmax=0.03, **kw):
    import tamc

    t = tamc.MCTransport(n, n, n, xmax, ymax, zmax)
    rk = tamc.gridset(xmax, ymax, zmax, n, n, n, 680.0)[3]
    t.set_optics(rk, 0.0, 0.9)
    d1 = t.heat_init(**kw)
    h = orc.HeatOracle(n, xmax, ymax, zmax)
    d2 = h.init(kappa=680.0, **{k: v for k, v in kw.items() if k != "ablateTemp"})
    assert d1 == d2
    return t, h


def _compare(t, h, exact=True, rtol=1e-10):
    for name in HALO + INNER:
        a, b = t.heat_array(name), h.array(name)
        if exact and name in ("temp", "watercontent", "Q"):
            pass
        scale = np.maximum(np.abs(b), 1e-300)
        err = np.abs(a - b) / scale
        err[a == b] = 0
        assert err.max() <= rtol, f"{name}: max rel err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"
    th_a, th_b = t.heat_array("threstime"), h.threstime()
    assert np.array_equal(th_a != 0, th_b != 0)
    assert np.allclose(th_a, th_b, rtol=1e-13, atol=0)
    for name in ("delt", "time", "total_time", "pulselength", "realPulseLength", "laserOn", "pulseCount", "repetitionCount",
                 "laser_flag", "QVapor"):
        assert t.heat_scalar(name) == h.scalar(name), name


def test_init_parity():
    t, h = _pair(20, zmax=0.06)
    _compare(t, h, rtol=0.0)              # initial state: bit-exact
    assert abs(t.heat_scalar("pwr") - h.get_pwr()) <= 1e-15 * abs(h.get_pwr())
    t.close()


@pytest.mark.parametrize("pulsetype,power,iters", [("tophat", 20.0, 30), ("triangular", 200.0, 50), ("gaussian", 70.0, 12)])
def test_step_by_step_parity_with_boiling_and_ablation(pulsetype, power, iters):
    """Every iteration: same (unscaled) tally into both sides, then scale + heat + Arrhenius + property update.

    Once voxels turn to air the reference's explicit scheme is unstable on this coarse grid (air diffusivity is
    ~150x tissue's at the same time step) and the temperatures overflow within a few calls -- upstream behaviour,
    reproduced by the oracle.  The comparison runs up to the first non-finite temperature, which comes after
    boiling and ablation have both been exercised."""
    n, npk = 24, 20000
    t, h = _pair(n, zmax=0.06, pulsetype=pulsetype, power=power, energyPerPixel=4000.0, ablateTemp=150.0, loops=2)
    boiled = ablated = False
    for it in range(iters):
        t.run_async(npk, 77)                            # tally stays on the device, rhokap is the resident one
        jm = t.get_jmean()                              # unscaled copy for the oracle
        h.scale_jmean(jm, npk)
        h.sim_3d(jm, it)
        h.arrhenius()
        h.setup_thermal_coeff(150.0)
        t.heat_step(npk)
        if not np.isfinite(h.array("temp")).all():
            assert not np.isfinite(t.heat_array("temp")).all()
            break
        _compare(t, h)
        boiled |= h.array("Q").max() > 0
        ablated |= bool((h.array("rhokap")[1:-1, 1:-1, 1:-1] == 0).any())
    if pulsetype != "gaussian":
        assert boiled and ablated                       # the scenario really exercises both branches
        # the transport follows the crater: ablated voxels take no deposit
        rk = t.heat_array("rhokap")[1:-1, 1:-1, 1:-1]
        t.run_async(npk, 77)
        assert t.get_jmean()[rk == 0].sum() == 0.0
    t.close()


def test_coupled_loop_resident_equals_host_driven_loop():
    """tamc_coupled_loop (nothing crosses PCIe) == the same loop driven call by call, and both follow the oracle
    that runs the photon loop on the same Philox stream."""
    import tamc

    n, npk, iters = 24, 20000, 13
    kw = dict(pulsetype="tophat", power=20.0, energyPerPixel=4000.0, ablateTemp=150.0, loops=2)
    t, h = _pair(n, zmax=0.06, **kw)
    done, packets = t.coupled_loop(npk, 5, iters)
    assert done == iters and packets == iters * npk
    # oracle: MC on the same ids, then the heat step
    o = orc.Oracle(n, n, n, 0.03, 0.03, 0.06)
    o.set_optics(0.0, 0.9)
    for it in range(iters):
        o.set_rhokap(h.array("rhokap"))
        o.zero_jmean()
        o.seed_philox(5, it * npk)
        o.run(npk)
        jm = np.asfortranarray(o.jmean.copy())
        h.scale_jmean(jm, npk)
        h.sim_3d(jm, it)
        h.arrhenius()
        h.setup_thermal_coeff(150.0)
    _compare(t, h, rtol=1e-9)             # the tally differs by summation order (1e-13), amplified mildly by 25 steps
    assert (h.array("rhokap")[1:-1, 1:-1, 1:-1] == 0).any()
    t.close()


def test_heat_errors():
    import tamc

    t = tamc.MCTransport(8, 8, 10, 0.03, 0.03, 0.03)
    t.set_optics(np.zeros((10, 10, 12), order="F"), 0.0, 0.9)
    with pytest.raises(tamc.TamcError) as e:
        t.heat_init()
    assert e.value.code == 1                    # needs a cubic grid
    t.close()
    t = tamc.MCTransport(8, 8, 8, 0.03, 0.03, 0.03)
    with pytest.raises(tamc.TamcError) as e:
        t.heat_init()
    assert e.value.code == 5                    # rhokap not resident yet
    with pytest.raises(tamc.TamcError) as e:
        t.heat_step(100)
    assert e.value.code == 5
    t.close()


def test_shipped_configuration_coupled_trace_matches_oracle_fixture():
    """The reference's own run (res/input.params, 80^3, 125 000 packets per call, gaussian pulse) through the
    device-resident loop: temperatures at checkpoints and the iterations of the first boiling voxel, the first
    ablated voxel and the divergence of the explicit scheme equal the CPU oracle's (tests/golden/
    coupled_shipped_events.json, produced by tests/golden/make_coupled_oracle_trace.py in 5.5 CPU-minutes)."""
    import json
    import os

    import tamc

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "coupled_shipped_events.json")))
    n = 80
    t = tamc.MCTransport(n, n, n, 0.03, 0.03, 0.06)
    t.set_optics(tamc.gridset(0.03, 0.03, 0.06, n, n, n, 680.0)[3], 0.0, 0.9)
    t.heat_init()
    assert int(t.heat_scalar("total_time") / t.heat_scalar("delt")) == 13390
    first, done = {}, 0
    while done < 5000:
        step = 1 if done >= 3900 else 100
        try:
            it, _ = t.coupled_loop(125000, 95648324, step)
        except tamc.TamcError:
            first.setdefault("diverged", done)
            break
        done += it
        if str(done) in gold["checkpoints"]:
            tmax = t.heat_array("temp")[1:-1, 1:-1, 1:-1].max() - 273.0
            assert abs(tmax - gold["checkpoints"][str(done)]) < 0.006, (done, tmax)
        if step == 1:
            if "boil" not in first and t.heat_array("Q").max() > 0:
                first["boil"] = done - 1
            if "boil" in first and done > 4600 and "ablate" not in first and (t.heat_array("rhokap")[1:-1, 1:-1, 1:-1] == 0).any():
                first["ablate"] = done - 1
            if done > 4680 and not np.isfinite(t.heat_array("temp")).all():
                first.setdefault("diverged", done - 1)
                break
    assert first == {"boil": gold["first_boil_iteration"], "ablate": gold["first_ablation_iteration"],
                     "diverged": gold["diverged_iteration"]}
    t.close()


@pytest.mark.parametrize("air_fraction,seed", [(0.5, 1), (0.75, 2), (0.9, 3)])
def test_neighbour_rule_single_pass_equals_the_sequential_sweep(air_fraction, seed):
    """The "six neighbours ablated" rule (3dFD.f90:347-353) runs upstream inside the k,j,i sweep; the device evaluates
    it in one data-parallel pass (argument in csrc/tamc_heat.cu).  Hand-made air pockets -- isolated voxels, pairs,
    chains along every axis, voxels against the halo -- against the sequential oracle, several calls in a row."""
    n, npk = 16, 20000
    t, h = _pair(n, zmax=0.06, pulsetype="tophat", power=5.0, energyPerPixel=4000.0, ablateTemp=150.0, loops=1)
    rng = np.random.default_rng(seed)
    rk = h.array("rhokap")                                # writable view of the oracle's array
    hole = rng.uniform(size=(n, n, n)) < air_fraction
    rk[1:-1, 1:-1, 1:-1][hole] = 0.0
    # explicit chains of tissue through an air block, along each axis and ending at the halo
    rk[2:9, 2:9, 2:9] = 0.0
    rk[3:8, 5, 5] = 680.0
    rk[5, 3:8, 3] = 680.0
    rk[3, 3, 3:8] = 680.0
    rk[1, 1, 1] = 680.0
    rk[n, n, n] = 680.0
    t.heat_upload("rhokap", np.asfortranarray(rk.copy()))
    before = rk[1:-1, 1:-1, 1:-1].copy()
    zeroed_by_rule = 0
    for it in range(4):
        t.run_async(npk, 77)
        jm = t.get_jmean()
        h.scale_jmean(jm, npk)
        h.sim_3d(jm, it)
        h.arrhenius()
        h.setup_thermal_coeff(150.0)
        t.heat_step(npk)
        if not np.isfinite(h.array("temp")).all():
            break
        a, b = t.heat_array("rhokap"), h.array("rhokap")
        assert np.array_equal(a == 0, b == 0), f"call {it}: the set of ablated voxels differs"
        _compare(t, h)
        now = b[1:-1, 1:-1, 1:-1]
        zeroed_by_rule += int(((before != 0) & (now == 0)).sum())
        before = now.copy()
    assert zeroed_by_rule > 0                            # the rule really removed isolated tissue
    t.close()
