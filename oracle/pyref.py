"""Second, independent transliteration of the reference's photon loop -- pure Python, small N only.

TEST INFRASTRUCTURE ONLY.  Written separately from oracle/tamc_oracle.c (different structure: no
module-global photon, closures instead of a state struct, dict-of-tuples tally) so that a slip in one
transliteration shows up as a mismatch against the other.  Python floats are IEEE doubles and
``math`` calls the same libm as the C oracle, so the two must agree bit for bit.

Follows /root/reference/src: ran2.f:1-33, sourceph.f90:7-49, inttau2.f90:7-239, stokes.f90:6-153,
gridset.f90:23-31, mcpolar.f90:97-98,112,151-170.  No upstream golden vectors and no Fortran compiler here: pinned to
what the reference's own source text computes when oracle/f90interp.py executes it (tests/test_oracle_reference_vectors.py),
not to a compiled reference.

Run as a script to regenerate tests/golden/oracle_kat.json.
"""
from __future__ import annotations

import json
import math
import os
import sys

PI = 3.141592      # constants.f90:13 (truncated, kept on purpose)
TWOPI = 6.283185


class Ran2:
    """ran2.f:1-33 with its SAVEd state; seeded as mcpolar.f90:97-98 does for MPI rank `rank_id`."""

    IM1, IM2 = 2147483563, 2147483399
    IA1, IA2 = 40014, 40692
    IQ1, IQ2 = 53668, 52774
    IR1, IR2 = 12211, 3791
    NTAB = 32

    def __init__(self, rank_id: int = 0):
        self.idum = -abs(-95648324 + rank_id)
        self.idum2 = 123456789
        self.iv = [0] * self.NTAB
        self.iy = 0
        self.count = 0

    @staticmethod
    def _schrage(x, a, q, r, m):
        k = x // q          # operands are positive: floor == Fortran integer division
        x = a * (x - k * q) - k * r
        return x + m if x < 0 else x

    def __call__(self) -> float:
        imm1 = self.IM1 - 1
        ndiv = 1 + imm1 // self.NTAB
        if self.idum <= 0:
            self.idum = max(-self.idum, 1)
            self.idum2 = self.idum
            for j in range(self.NTAB + 8, 0, -1):
                self.idum = self._schrage(self.idum, self.IA1, self.IQ1, self.IR1, self.IM1)
                if j <= self.NTAB:
                    self.iv[j - 1] = self.idum
            self.iy = self.iv[0]
        self.idum = self._schrage(self.idum, self.IA1, self.IQ1, self.IR1, self.IM1)
        self.idum2 = self._schrage(self.idum2, self.IA2, self.IQ2, self.IR2, self.IM2)
        j = self.iy // ndiv
        self.iy = self.iv[j] - self.idum2
        self.iv[j] = self.idum
        if self.iy < 1:
            self.iy += imm1
        self.count += 1
        return min((1.0 / self.IM1) * self.iy, 1.0 - 1.2e-7)


def make_faces(n: int, vmax: float):
    """gridset.f90:23-31."""
    return [(i - 1) * 2.0 * vmax / n for i in range(1, n + 2)]


def find(val: float, a) -> int:
    """inttau2.f90:208-239; `a` is a Python list holding the 1-based Fortran array."""
    n = len(a)
    if val == a[0]:
        return 1
    if val == a[n - 1]:
        return n - 1
    if val > a[n - 1] or val < a[0]:
        return -1
    lo, hi = 0, n + 1
    while hi - lo > 1:
        mid = (hi + lo) // 2
        if val >= a[mid - 1]:
            lo = mid
        else:
            hi = mid
    return lo


def rang(avg: float, sigma: float, rng) -> float:
    """sourceph.f90:73-101 (Marsaglia polar method) over ranu, :52-70."""
    while True:
        u = -1.0 + rng() * (1.0 - -1.0)
        s = -1.0 + rng() * (1.0 - -1.0)
        s = s * s + u * u
        if not s >= 1.0:
            break
    return avg + sigma * (u * math.sqrt(-2.0 * math.log(s) / s))


def photon_loop(nphotons, nxg, nyg, nzg, xmax, ymax, zmax, rhokap, rng, albedo=0.0, hgg=0.9,
                scatter=False, spot=250e-4, gauss_sigma=0.0, periodic=False):
    """mcpolar.f90:151-170.  rhokap(i,j,k) is a callable on 1-based interior indices.

    gauss_sigma > 0: launch point from rang() (sourceph.f90:73-101), redrawn while off the top face -- builder-defined
    use of upstream dead code.  periodic: repeat_bounds (inttau2.f90:242-279, dead code upstream) on lateral exits.
    Returns (tally dict {(i,j,k): value}, list of per-packet dicts)."""
    xf, yf, zf = make_faces(nxg, xmax), make_faces(nyg, ymax), make_faces(nzg, zmax)
    delta = 1.0e-8 * (2.0 * zmax / nzg)                     # mcpolar.f90:112
    g2 = hgg * hgg                                          # ch_opt.f90:16
    tally: dict = {}
    packets = []

    for _ in range(nphotons):
        n0 = rng.count
        # ---- sourceph.f90:28-47
        if gauss_sigma > 0.0:
            pos = [0.0, 0.0, zmax - (1.0e-8 * (2.0 * zmax / nzg))]
            for ax, lim in ((0, xmax), (1, ymax)):
                v = rang(0.0, gauss_sigma, rng)
                while not abs(v) < lim:
                    v = rang(0.0, gauss_sigma, rng)
                pos[ax] = v
        else:
            r = rng() * ((spot / 2.0) * (spot / 2.0))
            theta = rng() * TWOPI
            pos = [math.sqrt(r) * math.cos(theta), math.sqrt(r) * math.sin(theta),
                   zmax - (1.0e-8 * (2.0 * zmax / nzg))]
        phi = TWOPI * rng()
        cosp, sinp = math.cos(phi), math.sin(phi)
        sint, cost = 0.0, -1.0
        n = [sint * cosp, sint * sinp, cost]
        cell = [int(nxg * (pos[0] + xmax) / (2.0 * xmax)) + 1,
                int(nyg * (pos[1] + ymax) / (2.0 * ymax)) + 1,
                int(nzg * (pos[2] + zmax) / (2.0 * zmax)) + 1]
        if gauss_sigma > 0.0:
            cell[0], cell[1] = min(cell[0], nxg), min(cell[1], nyg)
        info = {"steps": 0, "deposit": 0.0, "nscatt": 0}
        if periodic:
            info["wraps"] = 0

        def tauint1():
            """inttau2.f90:7-72; returns True when the packet left the grid."""
            cur = [pos[0] + xmax, pos[1] + ymax, pos[2] + zmax]
            faces = (xf, yf, zf)
            taurun = 0.0
            tau = -math.log(rng())
            left = False
            while True:
                # wall_dist, inttau2.f90:75-121
                dist = []
                for ax in range(3):
                    if n[ax] > 0.0:
                        dist.append((faces[ax][cell[ax]] - cur[ax]) / n[ax])          # face(c+1)
                    elif n[ax] < 0.0:
                        dist.append((faces[ax][cell[ax] - 1] - cur[ax]) / n[ax])      # face(c)
                    else:
                        dist.append(100000.0)
                dcell = min(dist[0], dist[1], dist[2])
                hit = None
                for ax in range(3):                       # later axis wins ties
                    if dcell == dist[ax]:
                        hit = ax
                rk = rhokap(cell[0], cell[1], cell[2])
                taucell = dcell * rk
                info["steps"] += 1
                key = (cell[0], cell[1], cell[2])
                if taurun + taucell < tau:
                    taurun = taurun + taucell
                    tally[key] = tally.get(key, 0.0) + dcell * rk
                    info["deposit"] += dcell * rk
                    # update_pos with wall_flag, inttau2.f90:140-170
                    for ax in range(3):
                        if ax == hit:
                            if n[ax] > 0.0:
                                cur[ax] = faces[ax][cell[ax]] + delta
                            elif n[ax] < 0.0:
                                cur[ax] = faces[ax][cell[ax] - 1] - delta
                        else:
                            cur[ax] = cur[ax] + n[ax] * dcell
                    for ax in range(3):                   # update_voxels, inttau2.f90:201-203
                        cell[ax] = find(cur[ax], faces[ax])
                    if periodic:                          # repeat_bounds, inttau2.f90:242-279
                        for ax, lim, ng in ((0, xmax, nxg), (1, ymax, nyg)):
                            if cell[ax] == -1:
                                if cur[ax] < delta:
                                    cur[ax], cell[ax] = 2.0 * lim - delta, ng
                                    info["wraps"] += 1
                                elif cur[ax] > 2.0 * lim - delta:
                                    cur[ax], cell[ax] = delta, 1
                                    info["wraps"] += 1
                else:
                    dcell = (tau - taurun) / rk
                    tally[key] = tally.get(key, 0.0) + dcell * rk
                    info["deposit"] += dcell * rk
                    for ax in range(3):
                        cur[ax] = cur[ax] + n[ax] * dcell
                    break
                if -1 in cell:
                    left = True
                    break
            pos[0], pos[1], pos[2] = cur[0] - xmax, cur[1] - ymax, cur[2] - zmax
            return left

        def stokes():
            """stokes.f90:6-153."""
            nonlocal sint, cost, phi, cosp, sinp
            if hgg == 0.0:
                cost = 2.0 * rng() - 1.0
                sint = 1.0 - cost * cost
                sint = 0.0 if sint <= 0.0 else math.sqrt(sint)
                phi = TWOPI * rng()
                sinp, cosp = math.sin(phi), math.cos(phi)
                n[0], n[1], n[2] = sint * cosp, sint * sinp, cost
                return
            costp, sintp, phip = cost, sint, phi
            q = (1.0 - g2) / (1.0 - hgg + 2.0 * hgg * rng())
            bmu = ((1.0 + g2) - q * q) / (2.0 * hgg)
            cosb2 = bmu * bmu
            if abs(bmu) > 1.0:
                bmu = 1.0 if bmu > 1.0 else -1.0
                cosb2 = 1.0
            sinbt = math.sqrt(1.0 - cosb2)
            ri1 = TWOPI * rng()
            upper = ri1 > PI
            ri = TWOPI - ri1 if upper else ri1
            cosi, sini = math.cos(ri), math.sin(ri)
            if bmu == 1.0 or bmu == -1.0:
                return                                     # goto 100: direction untouched
            cost = costp * bmu + sintp * sinbt * cosi
            cosi2 = 0.0
            if abs(cost) < 1.0:
                sint = abs(math.sqrt(1.0 - cost * cost))
                sini2 = sini * sintp / sint
                bott = sint * sinbt
                cosi2 = costp / bott - cost * bmu / bott
            else:
                sint = 0.0
                sini2 = 0.0
                if cost >= 1.0:
                    cosi2 = -1.0
                if cost <= -1.0:
                    cosi2 = 1.0
            if upper:
                cosdph = -cosi2 * cosi + sini2 * sini * bmu      # stokes.f90:92
            else:
                cosdph = -cosi * cosi2 + sini * sini2 * bmu      # stokes.f90:130
            if abs(cosdph) > 1.0:
                cosdph = 1.0 if cosdph > 1.0 else -1.0
            phi = phip + math.acos(cosdph) if upper else phip - math.acos(cosdph)
            if phi > TWOPI:
                phi = phi - TWOPI
            if phi < 0.0:
                phi = phi + TWOPI
            cosp, sinp = math.cos(phi), math.sin(phi)
            n[0], n[1], n[2] = sint * cosp, sint * sinp, cost

        tflag = tauint1()
        absorbed = False
        while not tflag:
            if not scatter:                                # mcpolar.f90:166-169 stub
                absorbed = True
                break
            if rng() < albedo:
                stokes()
                info["nscatt"] += 1
            else:
                absorbed = True
                break
            tflag = tauint1()

        fate = 0
        if not absorbed:
            for ax in range(3):
                if cell[ax] == -1:
                    fate = 2 * ax + (2 if n[ax] > 0.0 else 1)
                    break
        packets.append({"pos": list(pos), "dir": list(n), "cell": list(cell), "fate": fate,
                        "ndraws": rng.count - n0, **info})
    return tally, packets


def golden():
    """Known answers committed under tests/golden/ (builder-derived, not from a Fortran run)."""
    out = {"source": "oracle/pyref.py (independent Python transliteration of the Fortran)", "ran2": {}}
    for rank in (0, 1, 7):
        g = Ran2(rank)
        out["ran2"][str(rank)] = {"first8": [g() for _ in range(8)], "idum": g.idum, "idum2": g.idum2, "iy": g.iy}
    # shipped regime, rank 0, first 16 packets (res/input.params, ch_opt.f90:17-22)
    tally, pk = photon_loop(16, 80, 80, 80, 0.03, 0.03, 0.06, lambda i, j, k: 680.0, Ran2(0))
    out["shipped_first16"] = pk
    out["shipped_first16_tally"] = [[list(k), v] for k, v in sorted(tally.items())]
    # small turbid cube with the scatter loop on, rank 3
    tally, pk = photon_loop(12, 20, 20, 20, 0.05, 0.05, 0.05, lambda i, j, k: 101.0 if k > 4 else 55.0,
                            Ran2(3), albedo=100.0 / 101.0, hgg=0.9, scatter=True)
    out["turbid_first12"] = pk
    out["turbid_first12_tally"] = [[list(k), v] for k, v in sorted(tally.items())]
    # isotropic branch of stokes (hgg == 0)
    tally, pk = photon_loop(8, 16, 16, 16, 0.04, 0.04, 0.04, lambda i, j, k: 60.0, Ran2(5), albedo=0.9,
                            hgg=0.0, scatter=True)
    out["isotropic_first8"] = pk
    out["isotropic_first8_tally"] = [[list(k), v] for k, v in sorted(tally.items())]
    return out


def golden_next():
    """Known answers for the SURVEY 8(f)-2 options built on upstream dead code (rang, repeat_bounds)."""
    out = {"source": "oracle/pyref.py golden_next() (independent Python transliteration; builder-defined call sites)"}
    # Gaussian beam, shipped stub regime on a coarse grid; sigma comparable to the half-width so redraws happen
    tally, pk = photon_loop(24, 20, 20, 20, 0.03, 0.03, 0.06, lambda i, j, k: 680.0, Ran2(2), gauss_sigma=0.02)
    out["gauss_stub_first24"] = pk
    out["gauss_stub_first24_tally"] = [[list(k), v] for k, v in sorted(tally.items())]
    # periodic lateral boundaries, thin turbid slab (most packets cross a lateral face several times)
    tally, pk = photon_loop(16, 8, 8, 24, 0.01, 0.01, 0.06, lambda i, j, k: 90.0 if k > 6 else 40.0, Ran2(4),
                            albedo=0.95, hgg=0.8, scatter=True, spot=0.01, periodic=True)
    out["periodic_first16"] = pk
    out["periodic_first16_tally"] = [[list(k), v] for k, v in sorted(tally.items())]
    # both at once
    tally, pk = photon_loop(12, 10, 12, 14, 0.02, 0.024, 0.03, lambda i, j, k: 120.0, Ran2(6), albedo=0.9, hgg=0.0,
                            scatter=True, gauss_sigma=0.015, periodic=True)
    out["gauss_periodic_first12"] = pk
    out["gauss_periodic_first12_tally"] = [[list(k), v] for k, v in sorted(tally.items())]
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, "..", "tests", "golden", "oracle_kat.json")
    if len(sys.argv) > 1:
        dst = sys.argv[1]
    with open(dst, "w") as f:
        json.dump(golden(), f, indent=1)
    print("wrote", os.path.normpath(dst))
    dst2 = os.path.join(os.path.dirname(dst), "oracle_kat_next.json")
    with open(dst2, "w") as f:
        json.dump(golden_next(), f, indent=1)
    print("wrote", os.path.normpath(dst2))
