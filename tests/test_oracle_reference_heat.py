"""The heat / ablation oracle (oracle/heat_oracle.c, the checker of tests/test_gpu_heat.py) against what the reference's OWN
Fortran text computes for one rank: tests/golden/reference_interp_heat.json.gz was produced in the build container by
tests/golden/make_reference_heat_vectors.py, which executes 3dFD.f90 (the time loop of heat_sim_3D, Arrhenius,
setupThermalCoeff, initThermalCoeff's arithmetic, the three getPwr functions), thermalConst_mod.f90 and the driver lines of
mcpolar.f90 with oracle/f90interp.py.  Bit for bit: every array after the first, the middle and the last iteration and at
the end (temperature with its boundary planes, opacity, thermal coefficients, water content, vaporisation energy, damage
integral, threshold times), exact sums and the scalars after every iteration -- through boiling, water loss and ablation,
for the top-hat, triangular and Gaussian pulses.  Nothing here reads /root/reference."""
import gzip
import json
import math
import os
import struct

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "reference_interp_heat.json.gz"), "rt") as f:
        return json.load(f)


def hexf(x):
    return struct.pack(">d", float(x)).hex()


def hexa(a):
    return np.asarray(a, dtype=">f8").tobytes(order="F").hex()


def exact_sum(a):
    return math.fsum(np.asarray(a, dtype=np.float64).ravel().tolist())


ARRAYS = {"temp": "temp", "rhokap": "rhokap", "tissue": "tissue", "kappa": "kappa", "density": "density", "heatcap": "heatcap",
          "coeff": "coeff", "alpha": "alpha", "watercontent": "watercontent", "q": "Q"}
SCALARS = {"time": "time", "delt": "delt", "laseron": "laserOn", "pulsecount": "pulseCount", "repetitioncount": "repetitionCount"}


def _check_snapshot(h, snap, where):
    for key, name in ARRAYS.items():
        assert hexa(h.array(name)) == snap[key], (where, key)
    assert hexa(h.threstime()) == snap["threstime"], (where, "threstime")
    for key, name in SCALARS.items():
        assert hexf(h.scalar(name)) == snap[key], (where, key, h.scalar(name))
    assert int(h.scalar("laser_flag")) == snap["laser_flag"] and int(h.scalar("pulsesDone")) == snap["pulsesdone"], where


@pytest.mark.parametrize("which", [0, 1, 2, 3, 4])
def test_coupled_iterations_bit_for_bit(ref, which):
    c = ref["cases"][which]
    n = c["n"]
    h = orc.HeatOracle(n, *c["extents"])
    h.init(power=c["power"], energyPerPixel=c["energyPerPixel"], total_time=c["total_time_in"], loops=c["loops"],
           repetitionRate_1=c["repetitionRate_1"], pulsesToDo=c["pulsesToDo"], pulsetype=c["pulsetype"])
    for key, name in (("delt", "delt"), ("total_time", "total_time"), ("pulselength", "pulselength"), ("realPulseLength", "realPulseLength"),
                      ("QVapor", "QVapor"), ("volumeVoxel", "volumeVoxel"), ("massVoxel", "massVoxel")):
        assert hexf(h.scalar(name)) == c["init"][key], key
    _check_snapshot(h, c["init"], "init")
    rng = np.random.default_rng(c["seed"])
    jglobal = np.zeros((n, n, n), order="F")
    boiled = ablated = False
    lasers = set()
    for i, step in enumerate(c["steps"]):
        jm = np.asfortranarray(rng.uniform(0.0, 3.0e4, (n, n, n)) * (rng.uniform(size=(n, n, n)) < 0.6))
        assert h.scalar("time") <= h.scalar("total_time")                      # mcpolar.f90:148
        if h.scalar("laser_flag"):                                             # :149
            jglobal[...] = jm                                                  # :173 (one rank)
            h.scale_jmean(jglobal, float(c["nphotons"]))                       # :174
        h.sim_3d(jglobal, i)                                                   # :178
        h.arrhenius()                                                          # :180
        h.setup_thermal_coeff(c["ablateTemp"])                                 # :182
        rk = h.array("rhokap")
        got = {"temp_max": hexf(h.array("temp").max()), "temp_sum": hexf(exact_sum(h.array("temp"))),
               "ablated": int((rk[1:-1, 1:-1, 1:-1] == 0).sum()), "q_sum": hexf(exact_sum(h.array("Q"))),
               "tissue_sum": hexf(exact_sum(h.array("tissue"))), "rhokap_sum": hexf(exact_sum(rk)),
               "time": hexf(h.scalar("time")), "laser_flag": int(h.scalar("laser_flag"))}
        for k, v in got.items():
            assert v == step[k], (i, k, v, step[k])
        if step["digest"] is not None:
            _check_snapshot(h, step["digest"], i)
        boiled |= exact_sum(h.array("Q")) > 0
        ablated |= got["ablated"] > 0
        lasers.add(got["laser_flag"])
    if "error_stop" in c:
        # the reference's run ends here: airThermalCond (thermalConst_mod.f90:20-23) stops on a negative temperature once the
        # explicit scheme diverges in the air voxels.  The oracle has no such stop; its next iteration shows the condition.
        assert c["error_stop"]["where"].startswith("thermalConst_mod.f90:") and c["error_stop"]["iteration"] == len(c["steps"])
        assert boiled and ablated                      # ... after the latent-heat sink and the ablation rule were exercised
        jm = np.asfortranarray(rng.uniform(0.0, 3.0e4, (n, n, n)) * (rng.uniform(size=(n, n, n)) < 0.6))
        if h.scalar("laser_flag"):
            jglobal[...] = jm
            h.scale_jmean(jglobal, float(c["nphotons"]))
        h.sim_3d(jglobal, len(c["steps"]))
        air = h.array("rhokap")[1:-1, 1:-1, 1:-1] == 0
        t = h.array("temp")[1:-1, 1:-1, 1:-1]
        assert (t[air] < 0).any() or not np.isfinite(t).all()
    else:
        _check_snapshot(h, c["final"], "final")
        assert lasers == {0, 1} and int(h.scalar("pulsesDone")) >= 2      # pulses ended and started again


def test_whole_coupled_iteration_bit_for_bit(ref):
    """mcpolar.f90:148-186 as the reference's text runs it for one rank: the photon loop (ran2 stream running on from call to
    call) on the opacity the previous property update left behind -> the scaled tally -> heat step -> Arrhenius -> property
    update, 28 iterations through boiling and ablation until the reference stops.  The two C oracles, coupled the same way."""
    c = ref["coupled"][0]
    n, npk = c["n"], c["nphotons"]
    h = orc.HeatOracle(n, *c["extents"])
    h.init(power=c["power"], energyPerPixel=c["energyPerPixel"], loops=c["loops"], pulsetype=c["pulsetype"])
    _check_snapshot(h, c["init"], "init")
    o = orc.Oracle(n, n, n, *c["extents"])
    o.init_opt1()                                                   # albedo 0, hgg 0.9 (ch_opt.f90:15-23)
    o.seed_ran2(c["rank"])
    jglobal = np.zeros((n, n, n), order="F")
    assert len(c["steps"]) >= 25
    for i, step in enumerate(c["steps"]):
        assert h.scalar("time") <= h.scalar("total_time")
        if h.scalar("laser_flag"):
            o.set_rhokap(h.array("rhokap"))                         # the property update rewrote iarray's rhokap in place
            o.zero_jmean()                                          # :185 of the previous iteration
            st = o.run(npk)["stats"]
            mc = step["mc"]
            assert (st["draws"], st["voxel_steps"]) == (mc["draws"], mc["voxel_steps"]), (i, st, mc)
            assert o.ran2_state()[0] == mc["iseed"] and hexf(exact_sum(o.jmean)) == mc["jmean_sum"], i
            assert int((o.jmean != 0).sum()) == mc["jmean_nonzero"]
            jglobal[...] = o.jmean
            h.scale_jmean(jglobal, float(npk))
        else:
            assert step["mc"] is None
        h.sim_3d(jglobal, i)
        h.arrhenius()
        h.setup_thermal_coeff(c["ablateTemp"])
        rk = h.array("rhokap")
        got = {"temp_max": hexf(h.array("temp").max()), "temp_sum": hexf(exact_sum(h.array("temp"))),
               "ablated": int((rk[1:-1, 1:-1, 1:-1] == 0).sum()), "q_sum": hexf(exact_sum(h.array("Q"))),
               "rhokap_sum": hexf(exact_sum(rk)), "tissue_sum": hexf(exact_sum(h.array("tissue"))),
               "time": hexf(h.scalar("time")), "laser_flag": int(h.scalar("laser_flag"))}
        for k, v in got.items():
            assert v == step[k], (i, k, v, step[k])
        if step["digest"] is not None:
            _check_snapshot(h, step["digest"], i)
    assert got["ablated"] > 0 and exact_sum(h.array("Q")) > 0 and c["error_stop"]["iteration"] == len(c["steps"])
