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
    o.set_indices(cfg.get("n1", 1.0), cfg.get("n2", 1.0))
    if "spot" in cfg:
        o.set_spot(cfg["spot"])
    if cfg.get("gauss_sigma", 0.0) > 0.0:
        o.set_source_gaussian(cfg["gauss_sigma"])
    return o


def make_transport(cfg, rhokap=None, device=0):
    import tamc

    n = cfg["n"]
    t = tamc.MCTransport(n, n, n, cfg["xmax"], cfg["ymax"], cfg["zmax"], device=device)
    if "spot" in cfg:
        t.set_source_co2(cfg["spot"])
    if cfg.get("gauss_sigma", 0.0) > 0.0:
        t.set_source_gaussian(cfg["gauss_sigma"])
    t.set_optics(rhokap if rhokap is not None else cfg["rhokap"](), cfg["albedo"], cfg["hgg"], n1=cfg.get("n1", 1.0),
                 n2=cfg.get("n2", 1.0), flags=cfg["flags"])
    return t


def compare_records(got, want, rtol=REPLAY_RTOL, scale=None, flip_fraction=0.0, p99_rtol=None):
    """Integer fields bit-exact; floating fields within rtol of the oracle (relative to the field's
    natural scale for coordinates that may sit near zero).

    flip_fraction > 0 (production arithmetic only): a path that differs from the oracle's in the 9th
    digit can pass on the other side of a voxel edge, which changes the visited-voxel sequence (steps,
    final cell by one) without moving the packet by more than rtol.  At most that fraction of packets
    (and never fewer than 2 allowed) may show such a flip -- by at most one voxel / two steps, with the
    same fate, draws and scatter count -- and they still have to meet the floating-point tolerance."""
    assert got.shape == want.shape
    flipped = np.zeros(got.shape, dtype=bool)
    for f in INT_FIELDS:
        bad = got[f] != want[f]
        if flip_fraction == 0.0 or f in ("nscatt", "ndraws", "fate"):
            idx = np.nonzero(bad)[0]
            assert idx.size == 0, f"{f}: {idx.size} packets differ, first {idx[:5]}: {got[f][idx[:5]]} vs {want[f][idx[:5]]}"
        else:
            lim = 2 if f == "steps" else 1
            assert np.all(np.abs(got[f][bad].astype(np.int64) - want[f][bad]) <= lim), f"{f}: flip larger than {lim}"
            flipped |= bad
    allowed = max(2, int(flip_fraction * got.size))
    assert flipped.sum() <= allowed, f"{flipped.sum()} packets changed their voxel sequence (allowed {allowed})"
    worst = 0.0
    for f in FP_FIELDS:
        s = np.abs(want[f])
        if scale is not None and f in scale:
            s = np.maximum(s, scale[f])
        err = np.abs(got[f] - want[f]) / np.maximum(s, 1e-300)
        err[(got[f] == want[f])] = 0.0
        worst = max(worst, float(err.max(initial=0.0)))
        assert err.max(initial=0.0) <= rtol, f"{f}: max rel err {err.max():.3e} at packet {err.argmax()}"
        if p99_rtol is not None and err.size:
            assert np.quantile(err, 0.99) <= p99_rtol, f"{f}: p99 rel err {np.quantile(err, 0.99):.3e}"
    return worst


def voxel_tau(cfg, rhokap):
    """Natural scale of one deposit: the optical depth of a voxel (largest opacity x smallest edge)."""
    n = cfg["n"]
    return float(np.max(rhokap)) * 2.0 * min(cfg["xmax"], cfg["ymax"], cfg["zmax"]) / n


def compare_grids(got, want, rtol=1e-11, dep_scale=None, sum_rtol=1e-8):
    """Tally grids against the oracle.

    Without scattering every chord is a full voxel edge or the final partial step, the two sides sum the
    same numbers in a different order, and the comparison is purely relative (rtol ~ 1e-11, identical
    support).  With scattering a path that differs in the 9th digit clips voxel corners by slightly
    different chords, so a voxel holding only such slivers can differ by much more than 1e-6 of its own
    (tiny) value; there the error is measured against max(|want|, dep_scale) with dep_scale = the optical
    depth of one voxel, i.e. 1e-6 of the natural size of a deposit (north_star tolerance)."""
    assert got.shape == want.shape
    if dep_scale is None:
        assert np.array_equal(got != 0, want != 0), "tally support differs"
        nz = want != 0
        if not nz.any():
            return 0.0
        err = np.abs(got[nz] - want[nz]) / np.abs(want[nz])
    else:
        err = np.abs(got - want) / np.maximum(np.abs(want), dep_scale)
    assert err.max() <= rtol, f"tally max rel err {err.max():.3e}"
    # the grid totals agree much more tightly than any single voxel
    assert abs(got.sum() - want.sum()) <= sum_rtol * abs(want.sum()) + 1e-300
    return float(err.max())
