"""Shared helpers for the parity tests (test infrastructure; may use the oracle)."""
import numpy as np

from oracle import oracle as orc

INT_FIELDS = ("xcell", "ycell", "zcell", "steps", "nscatt", "ndraws", "fate")
FP_FIELDS = ("xp", "yp", "zp", "nxp", "nyp", "nzp", "deposit")

# north_star: "reproduce individual photon paths and deposited energies to within 1e-6 relative (fp64)"
REPLAY_RTOL = 1e-6


def make_oracle(cfg, rhokap=None):
    n = cfg["n"]
    o = orc.Oracle(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"])
    o.set_rhokap(rhokap if rhokap is not None else cfg["rhokap"]())
    o.set_optics(cfg["albedo"], cfg["hgg"])
    o.set_flags(cfg["flags"])
    if "spot" in cfg:
        o.set_spot(cfg["spot"])
    return o


def make_transport(cfg, rhokap=None, device=0):
    import tamc

    n = cfg["n"]
    t = tamc.MCTransport(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"], device=device)
    if "spot" in cfg:
        t.set_source_co2(cfg["spot"])
    t.set_optics(rhokap if rhokap is not None else cfg["rhokap"](), cfg["albedo"], cfg["hgg"], flags=cfg["flags"])
    return t


def compare_records(got, want, rtol=REPLAY_RTOL, scale=None):
    """Integer fields bit-exact; floating fields within rtol of the oracle (relative to the field's
    natural scale for coordinates that may sit near zero)."""
    assert got.shape == want.shape
    for f in INT_FIELDS:
        bad = np.nonzero(got[f] != want[f])[0]
        assert bad.size == 0, f"{f}: {bad.size} packets differ, first {bad[:5]}: {got[f][bad[:5]]} vs {want[f][bad[:5]]}"
    worst = 0.0
    for f in FP_FIELDS:
        s = np.abs(want[f])
        if scale is not None and f in scale:
            s = np.maximum(s, scale[f])
        err = np.abs(got[f] - want[f]) / np.maximum(s, 1e-300)
        err[(got[f] == want[f])] = 0.0
        worst = max(worst, float(err.max(initial=0.0)))
        assert err.max(initial=0.0) <= rtol, f"{f}: max rel err {err.max():.3e} at packet {err.argmax()}"
    return worst


def compare_grids(got, want, rtol=1e-11):
    """Tally grids: same non-zero support, values equal up to the fp64 summation order."""
    assert got.shape == want.shape
    assert np.array_equal(got != 0, want != 0), "tally support differs"
    nz = want != 0
    if nz.any():
        err = np.abs(got[nz] - want[nz]) / np.abs(want[nz])
        assert err.max() <= rtol, f"tally max rel err {err.max():.3e}"
        return float(err.max())
    return 0.0
