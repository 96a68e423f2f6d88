"""The coupled iteration on the device against what the reference's OWN Fortran text computes for one rank
(tests/golden/reference_interp_heat.json.gz `coupled`: mcpolar.f90:148-186 -- the photon loop on the opacity the previous
property update left behind, the scaled tally, heat_sim_3D, Arrhenius, setupThermalCoeff -- executed by oracle/f90interp.py).
The MC call of every iteration is a trace replay of the reference's ran2 sequence on the device's own resident opacity; the
heat / ablation step is tamc_heat_step.  28 iterations through boiling and ablation, until the reference stops."""
import gzip
import json
import os
import struct

import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HALO = ("temp", "rhokap", "kappa", "density", "heatcap", "coeff", "alpha")
INNER = {"watercontent": "watercontent", "q": "Q", "tissue": "tissue"}


def unhex(h):
    return struct.unpack(">d", bytes.fromhex(h))[0]


def unhexa(h, shape):
    return np.frombuffer(bytes.fromhex(h), dtype=">f8").astype(np.float64).reshape(shape, order="F")


def _close(got, want, rtol, what):
    err = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
    err[got == want] = 0
    assert err.max() <= rtol, (what, float(err.max()), np.unravel_index(err.argmax(), err.shape))


def drive(dev, c, rtol=1e-10):
    """dev: heat_scalar / heat_array / run_replay / heat_step (the tamc binding).  Everything not passing through exp() is the
    oracle's arithmetic operation for operation (tests/test_gpu_heat.py), so the opacity the replay runs on is the reference's."""
    n, npk = c["n"], c["nphotons"]
    o = orc.Oracle(n, n, n, *c["extents"])
    o.init_opt1()
    o.seed_ran2(c["rank"])
    h3, i3 = (n + 2,) * 3, (n,) * 3
    for i, step in enumerate(c["steps"]):
        assert dev.heat_scalar("time") <= dev.heat_scalar("total_time")          # mcpolar.f90:148
        if dev.heat_scalar("laser_flag"):                                        # :149
            o.set_rhokap(dev.heat_array("rhokap"))
            o.zero_jmean()
            out = o.run(npk, draws_cap=npk * 4)                                   # only the draw list is taken from this run
            mc = step["mc"]
            assert int(out["offsets"][-1]) == mc["draws"] and o.ran2_state()[0] == mc["iseed"], (i, mc)
            rec, jm = dev.run_replay(out["offsets"], out["draws"])                # :151-170 on the device, tally stays resident
            assert int(rec["steps"].sum()) == mc["voxel_steps"] and int((jm != 0).sum()) == mc["jmean_nonzero"], i
            assert abs(float(jm.sum()) - unhex(mc["jmean_sum"])) <= 1e-11 * unhex(mc["jmean_sum"]), i
        dev.heat_step(npk)                                                        # :174-182
        rk = dev.heat_array("rhokap")
        assert int((rk[1:-1, 1:-1, 1:-1] == 0).sum()) == step["ablated"], i
        assert dev.heat_scalar("time") == unhex(step["time"]) and int(dev.heat_scalar("laser_flag")) == step["laser_flag"], i
        tmax = float(dev.heat_array("temp").max())
        assert abs(tmax - unhex(step["temp_max"])) <= rtol * unhex(step["temp_max"]), (i, tmax)
        if step["digest"] is not None:
            d = step["digest"]
            for name in HALO:
                _close(dev.heat_array(name), unhexa(d[name], h3), rtol, (i, name))
            for key, name in INNER.items():
                _close(dev.heat_array(name), unhexa(d[key], i3), rtol, (i, name))
    return step


def test_coupled_loop_on_the_device_against_the_reference_text():
    import tamc

    with gzip.open(os.path.join(ROOT, "tests", "golden", "reference_interp_heat.json.gz"), "rt") as f:
        c = json.load(f)["coupled"][0]
    n = c["n"]
    t = tamc.MCTransport(n, n, n, *c["extents"])
    t.set_optics(tamc.gridset(*c["extents"], n, n, n, 680.0)[3], 0.0, 0.9)
    t.heat_init(pulsetype=c["pulsetype"], power=c["power"], energyPerPixel=c["energyPerPixel"], ablateTemp=c["ablateTemp"],
                loops=c["loops"])
    last = drive(t, c)
    assert len(c["steps"]) >= 25 and last["ablated"] > 0
    t.close()
