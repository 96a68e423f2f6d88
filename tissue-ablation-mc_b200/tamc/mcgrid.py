"""Host-side mirror of the reference's set-up routines, same names and argument meaning, so a test
or driver written against the Fortran reads the same here.  These stay on the host in the reference
too (they run once); only their outputs cross the C ABI.

  gridset    /root/reference/src/gridset.f90:6-46   faces + uniform rhokap with a zero halo
  init_opt1  /root/reference/src/ch_opt.f90:7-25    10.6 um tissue optics
  delta_for  /root/reference/src/mcpolar.f90:112
"""
from __future__ import annotations

import numpy as np


def init_opt1():
    hgg = 0.9
    g2 = hgg * hgg
    mu_water, mu_protein = 510.0, 170.0
    mua = mu_water + mu_protein
    mus = 0.0
    kappa = mus + mua
    return {"hgg": hgg, "g2": g2, "mua": mua, "mus": mus, "kappa": kappa, "albedo": mus / kappa,
            "mu_water": mu_water, "mu_protein": mu_protein}


def gridset(xmax, ymax, zmax, nxg, nyg, nzg, kappa=None):
    """Returns (xface, yface, zface, rhokap) with rhokap(0:nxg+1,0:nyg+1,0:nzg+1) Fortran-ordered."""
    if kappa is None:
        kappa = init_opt1()["kappa"]
    xface = np.array([(i - 1) * 2.0 * xmax / nxg for i in range(1, nxg + 2)])
    yface = np.array([(i - 1) * 2.0 * ymax / nyg for i in range(1, nyg + 2)])
    zface = np.array([(i - 1) * 2.0 * zmax / nzg for i in range(1, nzg + 2)])
    rhokap = np.zeros((nxg + 2, nyg + 2, nzg + 2), dtype=np.float64, order="F")
    rhokap[1:-1, 1:-1, 1:-1] = kappa
    return xface, yface, zface, rhokap


def delta_for(zmax, nzg):
    return 1.0e-8 * (2.0 * zmax / nzg)
