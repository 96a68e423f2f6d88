"""The five workloads of BASELINE.json `configs`, made concrete as in SURVEY.md 8(d).

Each entry gives the grid, extents, opacity grid builder, scalar optics and flags.  `rhokap` builders
return the Fortran array (halo included) so it can be handed to tamc_set_optics unchanged.
"""
from __future__ import annotations

import numpy as np

from .mcgrid import gridset, init_opt1

SCATTER = 1


def _uniform(n, xmax, ymax, zmax, kappa):
    return gridset(xmax, ymax, zmax, n, n, n, kappa)[3]


def _layered_skin(n, zmax):
    """Config 3: cube 1 cm; from the top 0-0.01 cm rhokap 350, 0.01-0.2 cm 200, below 120 cm^-1."""
    rk = np.zeros((n + 2, n + 2, n + 2), dtype=np.float64, order="F")
    dz = 2.0 * zmax / n
    depth_top = (n - np.arange(1, n + 1)) * dz          # depth of the top of layer k below the surface
    layer = np.where(depth_top < 0.01 - 1e-12, 350.0, np.where(depth_top < 0.2 - 1e-12, 200.0, 120.0))
    rk[1:-1, 1:-1, 1:-1] = layer[None, None, :]
    return rk


def _crater(n, xmax, ymax, zmax, kappa, radius_vox, depth_vox):
    """Config 5 helper: ablated (rhokap = 0) cylinder under the beam, water-depleted rim."""
    rk = _uniform(n, xmax, ymax, zmax, kappa)
    ii, jj = np.meshgrid(np.arange(1, n + 1), np.arange(1, n + 1), indexing="ij")
    r = np.hypot(ii - 0.5 - n / 2.0, jj - 0.5 - n / 2.0)
    o = init_opt1()
    for k in range(n, max(0, n - depth_vox), -1):
        rk[1:-1, 1:-1, k][r <= radius_vox] = 0.0
        rim = (r > radius_vox) & (r <= radius_vox + 2)
        rk[1:-1, 1:-1, k][rim] = 0.5 * o["mu_water"] + o["mu_protein"]   # w*mu_water + mu_protein, 3dFD.f90:343
    return rk


CONFIGS = {
    # 1: shipped res/input.params + ch_opt.f90:15-23
    "shipped80": dict(n=80, xmax=0.03, ymax=0.03, zmax=0.06, albedo=0.0, hgg=0.9, flags=0,
                      rhokap=lambda: _uniform(80, 0.03, 0.03, 0.06, 680.0), nphotons=125000),
    # 2(i)/(iii): homogeneous 200^3, reference extents/optics/source
    "homog200": dict(n=200, xmax=0.03, ymax=0.03, zmax=0.06, albedo=0.0, hgg=0.9, flags=0,
                     rhokap=lambda: _uniform(200, 0.03, 0.03, 0.06, 680.0), nphotons=100_000_000),
    # 2(ii): turbid replay subset
    "turbid200": dict(n=200, xmax=0.5, ymax=0.5, zmax=0.5, albedo=100.0 / 101.0, hgg=0.9, flags=SCATTER,
                      rhokap=lambda: _uniform(200, 0.5, 0.5, 0.5, 101.0), nphotons=1_000_000),
    # 3: layered skin
    "skin200": dict(n=200, xmax=0.5, ymax=0.5, zmax=0.5, albedo=0.98, hgg=0.9, flags=SCATTER,
                    rhokap=lambda: _layered_skin(200, 0.5), nphotons=1_000_000_000),
    # 4: high-albedo turbid phantom ("stokes" = HG direction update; no polarisation state upstream)
    "phantom400": dict(n=400, xmax=1.0, ymax=1.0, zmax=1.0, albedo=0.999, hgg=0.9, flags=SCATTER,
                       rhokap=lambda: _uniform(400, 1.0, 1.0, 1.0, 100.0), nphotons=10_000_000),
}


def scaled(name, n):
    """Same physics on an n^3 grid (small-grid parity tests that the oracle finishes in seconds)."""
    c = dict(CONFIGS[name])
    if name in ("shipped80", "homog200"):
        c["rhokap"] = lambda: _uniform(n, c["xmax"], c["ymax"], c["zmax"], 680.0)
    elif name == "turbid200":
        c["rhokap"] = lambda: _uniform(n, 0.5, 0.5, 0.5, 101.0)
    elif name == "skin200":
        c["rhokap"] = lambda: _layered_skin(n, 0.5)
    elif name == "phantom400":
        c["rhokap"] = lambda: _uniform(n, 1.0, 1.0, 1.0, 100.0)
    c["n"] = n
    return c


def crater_sequence(n=80, steps=8):
    """Config 5: a scripted crater growing under the beam, one rhokap grid per MC call."""
    for s in range(steps):
        yield _crater(n, 0.03, 0.03, 0.06, 680.0, radius_vox=2 + 1.5 * s, depth_vox=1 + s)
