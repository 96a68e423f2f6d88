"""The oracles against outputs of the reference's OWN Fortran source.

tests/golden/reference_interp.json.gz was produced in the build container by tests/golden/make_reference_vectors.py, which
executes ran2.f, sourceph.f90, inttau2.f90, stokes.f90, gridset.f90, ch_opt.f90 and statement ranges of mcpolar.f90 -- read
from /root/reference, not restated -- with the Fortran-subset interpreter oracle/f90interp.py (whose own semantics are
tested in tests/test_f90interp.py).  Here the C oracle (the checker of every GPU parity test) and the independent Python
transliteration must reproduce those outputs BIT FOR BIT: every packet's final position, direction, voxel, number of
draws and of voxel-steps, every non-zero voxel of jmean, the generator state afterwards.  Nothing here reads
/root/reference."""
import gzip
import json
import os
import struct

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import pyref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ref():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "reference_interp.json.gz"), "rt") as f:
        return json.load(f)


def unhex(h):
    return struct.unpack(">d", bytes.fromhex(h))[0]


def bits(x):
    return struct.pack(">d", float(x)).hex()


def _dense(sparse, n):
    jm = np.zeros(n, order="F")
    for i, j, k, h in sparse:
        jm[i - 1, j - 1, k - 1] = unhex(h)
    return jm


def _check_packets(rows, rec, scatter=False):
    assert len(rows) == len(rec)
    for q, (row, r) in enumerate(zip(rows, rec)):
        got = [bits(r["xp"]), bits(r["yp"]), bits(r["zp"]), bits(r["nxp"]), bits(r["nyp"]), bits(r["nzp"]),
               int(r["xcell"]), int(r["ycell"]), int(r["zcell"])]
        # (-0.0 and +0.0 are the same direction cosine: `sint * cosp` with sint = 0. carries cosp's sign in both)
        assert got == row[:9], (q, got, row)
        assert int(r["ndraws"]) == row[10] and int(r["steps"]) == row[11], (q, r["ndraws"], r["steps"], row)
        if scatter:
            assert int(r["nscatt"]) == row[12], (q, r["nscatt"], row)
            left = row[6] == -1 or row[7] == -1 or row[8] == -1
            assert (int(r["fate"]) == 0) == (row[13] == 1) and (int(r["fate"]) != 0) == left, (q, r["fate"], row)


@pytest.mark.parametrize("which", [0, 1, 2])
def test_shipped_photon_loop_bit_for_bit(ref, which):
    """mcpolar.f90:153-169 per packet on the shipped configuration (80^3, kappa = 680 /cm from init_opt1), ranks 0, 1, 7."""
    v = ref["shipped"][which]
    n = v["grid"]
    o = orc.Oracle(n[0], n[1], n[2], *v["extents"])
    assert bits(o.init_opt1()) == v["kappa"] and bits(o.delta) == v["delta"]
    o.gridset_uniform(unhex(v["kappa"]))
    o.seed_ran2(v["rank"])
    assert o.ran2_state()[0] == v["seed"]                  # mcpolar.f90:97-98, before the first draw
    assert [bits(o.ran2()) for _ in range(12)] == v["first_draws"]
    o.seed_ran2(v["rank"])
    out = o.run(len(v["packets"]), records=True)
    _check_packets(v["packets"], out["records"])
    want = _dense(v["jmean"], n)
    assert np.array_equal(o.jmean, want)                   # every voxel, every bit
    assert o.ran2_state()[0] == v["iseed_after"]


def _check_pyref(rows, packets, tally, want_sparse, scatter=False):
    for q, (row, r) in enumerate(zip(rows, packets)):
        got = [bits(x) for x in r["pos"]] + [bits(x) for x in r["dir"]] + list(r["cell"])
        assert got == row[:9], (q, got, row)
        assert r["ndraws"] == row[10] and r["steps"] == row[11], (q, r, row)
        if scatter:
            assert r["nscatt"] == row[12] and (r["fate"] == 0) == (row[13] == 1), (q, r, row)
    assert sorted((i, j, k, bits(v)) for (i, j, k), v in tally.items() if v != 0.0) == sorted(tuple(x) for x in want_sparse)


def test_second_transliteration_against_the_reference_outputs(ref):
    """oracle/pyref.py (independently structured pure Python) against the same outputs: the shipped loop and the scatter loop."""
    v = ref["shipped"][1]
    n = v["grid"]
    kappa = unhex(v["kappa"])
    tally, pk = pyref.photon_loop(len(v["packets"]), n[0], n[1], n[2], *v["extents"], lambda i, j, k: kappa, pyref.Ran2(v["rank"]))
    _check_pyref(v["packets"], pk, tally, v["jmean"])
    v = ref["scatter"][1]
    n = v["grid"]
    kappa = v["mus"] + v["mua"]
    tally, pk = pyref.photon_loop(len(v["packets"]), n[0], n[1], n[2], *v["extents"], lambda i, j, k: kappa, pyref.Ran2(v["rank"]),
                                  albedo=v["mus"] / kappa, hgg=v["hgg"], scatter=True)
    _check_pyref(v["packets"], pk, tally, v["jmean"], scatter=True)


@pytest.mark.parametrize("which", [0, 1, 2])
def test_stokes_chain_bit_for_bit(ref, which):
    """stokes.f90 applied again and again (Henyey-Greenstein g = 0.9 and 0.5, isotropic g = 0): all eight photon_vars."""
    v = ref["stokes"][which]
    o = orc.Oracle(80, 80, 80, 0.03, 0.03, 0.06)
    o.init_opt1()
    o.set_optics(0.0, v["hgg"])
    o.seed_ran2(v["rank"])
    got = o.stokes_chain(len(v["rows"]))
    for s, row in enumerate(v["rows"]):
        assert [bits(x) for x in got[s]] == row, (s, [bits(x) for x in got[s]], row)
    assert o.ran2_state()[0] == v["iseed_after"]


@pytest.mark.parametrize("which", [0, 1])
def test_scatter_loop_bit_for_bit(ref, which):
    """SURVEY 3.3's loop (`ran2 < albedo ? stokes : absorbed ; tauint1`) around the reference's own sourcephCO2 / tauint1 /
    stokes in a turbid 1 cm cube (mus 100, mua 1 /cm; g = 0.9 and isotropic): packets of up to hundreds of scatterings."""
    v = ref["scatter"][which]
    n = v["grid"]
    o = orc.Oracle(n[0], n[1], n[2], *v["extents"])
    o.gridset_uniform(v["mus"] + v["mua"])
    albedo = v["mus"] / (v["mus"] + v["mua"])
    assert bits(albedo) == v["albedo"]
    o.set_optics(albedo, v["hgg"])
    o.set_flags(orc.FLAG_SCATTER if hasattr(orc, "FLAG_SCATTER") else 1)
    o.seed_ran2(v["rank"])
    out = o.run(len(v["packets"]), records=True)
    _check_packets(v["packets"], out["records"], scatter=True)
    assert sum(r[12] for r in v["packets"]) > 1000         # thousands of stokes calls behind those final states
    want = _dense(v["jmean"], n)
    assert np.array_equal(o.jmean, want)
    assert o.ran2_state()[0] == v["iseed_after"]


def test_dead_code_the_options_are_built_on(ref):
    """rang (Marsaglia polar method, sourceph.f90:73-101) and repeat_bounds (inttau2.f90:242-279) exist upstream but are never
    called there; the oracle's Gaussian-beam and periodic-boundary options call its restatements of them."""
    v = ref["dead_code"]
    o = orc.Oracle(80, 80, 80, 0.03, 0.03, 0.06)
    o.seed_ran2(v["rank"])
    got = [bits(o.rang(avg, sig)) for avg, sig in [(0.0, 1.0)] * 150 + [(0.25, 0.004)] * 50]
    assert got == v["rang"] and o.ran2_state()[0] == v["iseed_after"]
    assert bits(o.delta) == v["delta"]
    for c in v["repeat_bounds"]:
        cella, cellb, acur, bcur = c["in"]
        rc, ca, cb, xa, xb = o.repeat_bounds(cella, cellb, unhex(acur), unhex(bcur), v["amax"], v["bmax"], v["nag"], v["nbg"], o.delta)
        assert rc == c["status"], c
        if rc == 0:
            assert [ca, cb, bits(xa), bits(xb)] == c["out"], (c, ca, cb, xa, xb)


def test_set_up_routines_and_their_host_side_mirror(ref):
    """gridset.f90 (faces, rhokap with its zero halo), ch_opt.f90 init_opt1 and mcpolar.f90:112 (delta) as the reference's text
    computes them, against the C oracle and against the host-side mirror the binding and the driver shim use (tamc.mcgrid)."""
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tissue-ablation-mc_b200"))
    from tamc import mcgrid

    v = ref["shipped"][0]
    n = v["grid"]
    o = orc.Oracle(n[0], n[1], n[2], *v["extents"])
    xf, yf, zf, rk = mcgrid.gridset(*v["extents"], *n)
    of = o.faces()
    for name, mine, theirs in (("xface", xf, of[0]), ("yface", yf, of[1]), ("zface", zf, of[2])):
        assert [bits(x) for x in mine] == v["faces"][name], name
        assert [bits(x) for x in theirs] == v["faces"][name], name
    opt = mcgrid.init_opt1()
    for k, h in v["optics"].items():
        assert bits(opt[k]) == h, k
    assert bits(mcgrid.delta_for(v["extents"][2], n[2])) == v["delta"]
    assert v["rhokap_interior"] == [v["kappa"]] and unhex(v["rhokap_halo_sum"]) == 0.0
    assert sorted({bits(x) for x in rk[1:-1, 1:-1, 1:-1].ravel()}) == [v["kappa"]] and rk.sum() == rk[1:-1, 1:-1, 1:-1].sum()


def test_shipped_loop_on_an_opacity_with_holes(ref):
    """A transparent shaft (packets leave through the bottom face: find() = -1, tauint1 sets tflag), a shallow crater and a
    water-depleted rim under the beam: the reference's text against both oracles."""
    v = ref["crater"]
    n = v["grid"]
    rk = np.zeros((n[0] + 2, n[1] + 2, n[2] + 2), order="F")
    rk[1:-1, 1:-1, 1:-1] = 680.0
    rk[39:43, 39:43, 1:n[2] + 1] = 0.0
    rk[43:47, 36:47, n[2] - 5:n[2] + 1] = 0.0
    rk[47:49, 36:47, 1:n[2] + 1] = 0.5 * 510.0 + 170.0
    o = orc.Oracle(n[0], n[1], n[2], *v["extents"])
    o.set_rhokap(rk)
    o.set_optics(0.0, 0.9)
    o.seed_ran2(v["rank"])
    out = o.run(len(v["packets"]), records=True)
    _check_packets(v["packets"], out["records"])
    assert v["left_through_the_bottom"] >= 5 and out["stats"]["exits"][4] == v["left_through_the_bottom"]
    assert np.array_equal(o.jmean, _dense(v["jmean"], n)) and o.ran2_state()[0] == v["iseed_after"]
    tally, pk = pyref.photon_loop(len(v["packets"]), n[0], n[1], n[2], *v["extents"], lambda i, j, k: float(rk[i, j, k]),
                                  pyref.Ran2(v["rank"]))
    _check_pyref(v["packets"], pk, tally, v["jmean"])
