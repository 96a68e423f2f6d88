"""ctypes front-end of the CPU oracle (oracle/tamc_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
``--impl reference`` legs of bench.py.  The product (libtamc.so, the ``tamc`` package) never
imports this module.  Parity: pinned to outputs of the reference's own source text run by oracle/f90interp.py, not to a
compiled reference -- see tamc_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FLAG_SCATTER = 1
FLAG_FRESNEL = 2
FLAG_PERIODIC = 4

RECORD_DTYPE = np.dtype(
    [
        ("xp", "<f8"), ("yp", "<f8"), ("zp", "<f8"),
        ("nxp", "<f8"), ("nyp", "<f8"), ("nzp", "<f8"),
        ("deposit", "<f8"),
        ("xcell", "<i4"), ("ycell", "<i4"), ("zcell", "<i4"),
        ("steps", "<i4"), ("nscatt", "<i4"), ("ndraws", "<i4"),
        ("fate", "<i4"), ("flags", "<i4"),
    ],
    align=True,
)
assert RECORD_DTYPE.itemsize == 88


class Stats(C.Structure):
    _fields_ = [
        ("packets", C.c_int64),
        ("voxel_steps", C.c_int64),
        ("scatters", C.c_int64),
        ("absorbed", C.c_int64),
        ("exits", C.c_int64 * 6),
        ("draws", C.c_int64),
        ("deposit_sum", C.c_double),
        ("specular", C.c_int64),
        ("internal_reflections", C.c_int64),
        ("wraps", C.c_int64),
    ]

    def as_dict(self):
        return {
            "packets": self.packets, "voxel_steps": self.voxel_steps, "scatters": self.scatters,
            "absorbed": self.absorbed, "exits": list(self.exits), "draws": self.draws,
            "deposit_sum": self.deposit_sum, "specular": self.specular,
            "internal_reflections": self.internal_reflections, "wraps": self.wraps,
        }


def build(fast: bool = False, out_dir: str | None = None) -> str:
    """Compile the oracle.  fast=True is the timed CPU baseline (-O3 -march=native -flto)."""
    name = "liboracle_fast.so" if fast else "liboracle.so"
    out_dir = out_dir or _HERE
    out = os.path.join(out_dir, name)
    src = os.path.join(_HERE, "tamc_oracle.c")
    src2 = os.path.join(_HERE, "heat_oracle.c")
    if os.path.exists(out) and os.path.getmtime(out) >= max(
        os.path.getmtime(src), os.path.getmtime(src2), os.path.getmtime(os.path.join(_HERE, "tamc_oracle.h"))
    ):
        return out
    opt = ["-O3", "-march=native", "-flto"] if fast else ["-O2", "-ffp-contract=off"]
    cmd = ["gcc", "-std=c11", "-D_POSIX_C_SOURCE=200809L", "-fPIC", "-shared", "-pthread", *opt,
           "-o", out, src, src2, "-lm"]
    subprocess.run(cmd, check=True, capture_output=True)
    return out


_libs: dict[str, C.CDLL] = {}


def _bind(path: str) -> C.CDLL:
    lib = C.CDLL(path)
    p, d, i, i64 = C.c_void_p, C.c_double, C.c_int, C.c_int64
    lib.orc_create.restype = p
    lib.orc_create.argtypes = [i, i, i, d, d, d]
    lib.orc_destroy.argtypes = [p]
    for f in ("orc_rhokap", "orc_jmean", "orc_xface", "orc_yface", "orc_zface"):
        getattr(lib, f).restype = C.POINTER(d)
        getattr(lib, f).argtypes = [p]
    lib.orc_delta.restype = d
    lib.orc_delta.argtypes = [p]
    lib.orc_gridset_uniform.argtypes = [p, d]
    lib.orc_init_opt1.restype = d
    lib.orc_init_opt1.argtypes = [p]
    lib.orc_set_optics.argtypes = [p, d, d]
    lib.orc_set_spot.argtypes = [p, d]
    lib.orc_set_source_gaussian.argtypes = [p, d]
    lib.orc_set_indices.argtypes = [p, d, d]
    lib.orc_set_flags.argtypes = [p, i]
    lib.orc_set_grids.argtypes = [p, p, p, p]
    lib.orc_zero_jmean.argtypes = [p]
    lib.orc_seed_ran2.argtypes = [p, i]
    lib.orc_seed_philox.argtypes = [p, C.c_uint64, C.c_uint64]
    lib.orc_ran2.restype = d
    lib.orc_ran2.argtypes = [p]
    for f in ("orc_ran2_idum", "orc_ran2_idum2", "orc_ran2_iy"):
        getattr(lib, f).restype = i
        getattr(lib, f).argtypes = [p]
    lib.orc_rang.restype = d
    lib.orc_rang.argtypes = [p, d, d]
    lib.orc_repeat_bounds.restype = i
    lib.orc_repeat_bounds.argtypes = [C.POINTER(i), C.POINTER(i), C.POINTER(d), C.POINTER(d), d, d, i, i, d]
    lib.orc_stokes_chain.restype = None
    lib.orc_stokes_chain.argtypes = [p, i, C.c_void_p]
    lib.orc_find.restype = i
    lib.orc_find.argtypes = [d, C.POINTER(d), i]
    lib.orc_philox4x32_10.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
    lib.orc_run.restype = i
    lib.orc_run.argtypes = [p, i64, p, p, i64, p, C.POINTER(Stats)]
    lib.orc_run_ranks.restype = i
    # heat / ablation step (heat_oracle.c)
    lib.heat_create.restype = p
    lib.heat_create.argtypes = [i, d, d, d]
    lib.heat_destroy.argtypes = [p]
    lib.heat_array.restype = C.POINTER(d)
    lib.heat_array.argtypes = [p, i]
    lib.heat_scalar.restype = d
    lib.heat_scalar.argtypes = [p, i]
    lib.heat_init.restype = d
    lib.heat_init.argtypes = [p, d, d, d, i, d, i, i, d]
    lib.heat_getPwr.restype = d
    lib.heat_getPwr.argtypes = [p]
    lib.heat_scale_jmean.restype = d
    lib.heat_scale_jmean.argtypes = [p, p, d]
    lib.heat_sim_3d.argtypes = [p, p, i]
    lib.heat_arrhenius.argtypes = [p]
    lib.heat_setup_thermal_coeff.argtypes = [p, d]
    lib.orc_run_ranks.argtypes = [i, i, i, i, d, d, d, p, d, d, d, i, i64, p, C.POINTER(Stats), C.POINTER(d)]
    return lib


def load(fast: bool = False, out_dir: str | None = None) -> C.CDLL:
    key = "fast" if fast else "canon"
    if key not in _libs:
        _libs[key] = _bind(build(fast, out_dir))
    return _libs[key]


def philox4x32_10(key, ctr):
    out = (C.c_uint32 * 4)()
    load().orc_philox4x32_10(key[0], key[1], ctr[0], ctr[1], ctr[2], ctr[3], out)
    return list(out)


def find(val: float, faces: np.ndarray) -> int:
    a = np.ascontiguousarray(faces, dtype=np.float64)
    return load().orc_find(float(val), a.ctypes.data_as(C.POINTER(C.c_double)), int(a.size))


class Oracle:
    """One emulated MPI rank: the module state of mcpolar.f90 plus the photon loop."""

    def __init__(self, nxg, nyg, nzg, xmax, ymax, zmax, fast=False):
        self.lib = load(fast)
        self.nxg, self.nyg, self.nzg = int(nxg), int(nyg), int(nzg)
        self.xmax, self.ymax, self.zmax = float(xmax), float(ymax), float(zmax)
        self.h = self.lib.orc_create(self.nxg, self.nyg, self.nzg, self.xmax, self.ymax, self.zmax)
        if not self.h:
            raise MemoryError("orc_create failed")
        self.albedo, self.hgg = 0.0, 0.9

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- arrays (views into the oracle's own memory, Fortran layout) ---------------------------
    @property
    def rhokap(self) -> np.ndarray:
        shape = (self.nxg + 2, self.nyg + 2, self.nzg + 2)
        flat = np.ctypeslib.as_array(self.lib.orc_rhokap(self.h), shape=(int(np.prod(shape)),))
        return flat.reshape(shape, order="F")

    @property
    def jmean(self) -> np.ndarray:
        shape = (self.nxg, self.nyg, self.nzg)
        flat = np.ctypeslib.as_array(self.lib.orc_jmean(self.h), shape=(int(np.prod(shape)),))
        return flat.reshape(shape, order="F")

    def faces(self):
        return tuple(
            np.ctypeslib.as_array(getattr(self.lib, f)(self.h), shape=(n + 1,)).copy()
            for f, n in (("orc_xface", self.nxg), ("orc_yface", self.nyg), ("orc_zface", self.nzg))
        )

    @property
    def delta(self) -> float:
        return self.lib.orc_delta(self.h)

    # -- set-up -------------------------------------------------------------------------------
    def gridset_uniform(self, kappa: float):
        self.lib.orc_gridset_uniform(self.h, float(kappa))

    def init_opt1(self) -> float:
        self.albedo, self.hgg = 0.0, 0.9
        return self.lib.orc_init_opt1(self.h)

    def set_rhokap(self, rhokap_halo: np.ndarray):
        self.rhokap[...] = np.asarray(rhokap_halo, dtype=np.float64).reshape(self.rhokap.shape, order="F")

    def set_optics(self, albedo: float, hgg: float):
        self.albedo, self.hgg = float(albedo), float(hgg)
        self.lib.orc_set_optics(self.h, self.albedo, self.hgg)

    def set_indices(self, n1: float, n2: float):
        self.lib.orc_set_indices(self.h, float(n1), float(n2))

    def set_spot(self, diameter: float):
        self.lib.orc_set_spot(self.h, float(diameter))

    def set_source_gaussian(self, sigma: float):
        """Gaussian beam through rang() (sourceph.f90:73-101); sigma <= 0 = back to the CO2 disk."""
        self.lib.orc_set_source_gaussian(self.h, float(sigma))

    def set_grids(self, albedo=None, hgg=None, n=None):
        """EXTENSION: per-voxel albedo / hgg / refractive index, each shaped like rhokap (halo included) or None."""
        keep = []
        ptrs = []
        for a in (albedo, hgg, n):
            if a is None:
                ptrs.append(None)
            else:
                b = np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel(order="F"))
                assert b.size == self.rhokap.size
                keep.append(b)
                ptrs.append(b.ctypes.data)
        self.lib.orc_set_grids(self.h, *ptrs)

    def set_flags(self, flags: int):
        self.lib.orc_set_flags(self.h, int(flags))

    def zero_jmean(self):
        self.lib.orc_zero_jmean(self.h)

    def seed_ran2(self, rank_id: int = 0):
        self.lib.orc_seed_ran2(self.h, int(rank_id))

    def seed_philox(self, seed: int, first_packet_id: int = 0):
        self.lib.orc_seed_philox(self.h, int(seed) & (2**64 - 1), int(first_packet_id))

    def ran2(self) -> float:
        return self.lib.orc_ran2(self.h)

    def rang(self, avg: float, sigma: float) -> float:
        return self.lib.orc_rang(self.h, avg, sigma)

    def repeat_bounds(self, cella, cellb, acur, bcur, amax, bmax, nag, nbg, delta):
        """inttau2.f90:242-279 -> (status, cella, cellb, acur, bcur); status -1 = the Fortran's error stop."""
        ca, cb, xa, xb = C.c_int(cella), C.c_int(cellb), C.c_double(acur), C.c_double(bcur)
        rc = self.lib.orc_repeat_bounds(C.byref(ca), C.byref(cb), C.byref(xa), C.byref(xb), amax, bmax, nag, nbg, delta)
        return rc, ca.value, cb.value, xa.value, xb.value

    def stokes_chain(self, nsteps: int) -> np.ndarray:
        """sourcephCO2 once, then stokes nsteps times (ran2 stream); rows of nxp nyp nzp cost sint cosp sinp phi."""
        out = np.zeros((int(nsteps), 8), dtype=np.float64)
        self.lib.orc_stokes_chain(self.h, int(nsteps), out.ctypes.data)
        return out

    def ran2_state(self):
        return (self.lib.orc_ran2_idum(self.h), self.lib.orc_ran2_idum2(self.h), self.lib.orc_ran2_iy(self.h))

    # -- the photon loop ----------------------------------------------------------------------
    def run(self, nphotons: int, records: bool = False, draws_cap: int = 0):
        """Run nphotons packets; returns dict(stats, records?, draws?, offsets?)."""
        n = int(nphotons)
        st = Stats()
        rec = np.zeros(n, dtype=RECORD_DTYPE) if records else None
        drw = np.zeros(int(draws_cap), dtype=np.float64) if draws_cap else None
        off = np.zeros(n + 1, dtype=np.int64) if draws_cap else None
        rc = self.lib.orc_run(
            self.h, n,
            rec.ctypes.data if rec is not None else None,
            drw.ctypes.data if drw is not None else None, int(draws_cap),
            off.ctypes.data if off is not None else None, C.byref(st),
        )
        if rc != 0:
            raise RuntimeError("oracle: draw log capacity exceeded")
        out = {"stats": st.as_dict()}
        if rec is not None:
            out["records"] = rec
        if drw is not None:
            out["draws"] = drw[: off[-1]]
            out["offsets"] = off
        return out


def run_ranks(nranks, nxg, nyg, nzg, xmax, ymax, zmax, rhokap_halo, albedo, hgg, nphotons_per_rank,
              spot_diameter=0.0, flags=0, fast=False, out_dir=None):
    """R emulated MPI ranks on host threads + the summed jmean (mcpolar.f90:151-173)."""
    lib = load(fast, out_dir)
    rk = np.ascontiguousarray(np.asarray(rhokap_halo, dtype=np.float64).ravel(order="K"))
    assert rk.size == (nxg + 2) * (nyg + 2) * (nzg + 2)
    jm = np.zeros(nxg * nyg * nzg, dtype=np.float64)
    st = Stats()
    sec = C.c_double(0.0)
    used = lib.orc_run_ranks(int(nranks), nxg, nyg, nzg, xmax, ymax, zmax, rk.ctypes.data, albedo, hgg,
                             spot_diameter, flags, int(nphotons_per_rank), jm.ctypes.data, C.byref(st),
                             C.byref(sec))
    return {"jmean": jm.reshape((nxg, nyg, nzg), order="F"), "stats": st.as_dict(), "seconds": sec.value,
            "threads": used}


PULSETYPES = {"tophat": 0, "gaussian": 1, "triangular": 2}
HEAT_ARRAYS = {"temp": (0, True), "rhokap": (1, True), "kappa": (2, True), "density": (3, True), "heatcap": (4, True),
               "coeff": (5, True), "alpha": (6, True), "watercontent": (7, False), "Q": (8, False), "tissue": (9, False)}
HEAT_SCALARS = {"delt": 0, "time": 1, "total_time": 2, "pulselength": 3, "realPulseLength": 4, "laserOn": 5,
                "pulseCount": 6, "repetitionCount": 7, "laser_flag": 8, "QVapor": 9, "volumeVoxel": 10,
                "pulsesDone": 11, "negative_temp": 12, "massVoxel": 13}


class HeatOracle:
    """The Heat module of 3dFD.f90 for one rank, plus the driver lines around it (mcpolar.f90:123-185)."""

    def __init__(self, n, xmax, ymax, zmax):
        self.lib = load()
        self.n = int(n)
        self.h = self.lib.heat_create(self.n, float(xmax), float(ymax), float(zmax))

    def __del__(self):
        try:
            if self.h:
                self.lib.heat_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def init(self, power=70.0, energyPerPixel=400.0, total_time=2.0, loops=1, repetitionRate_1=1e7, pulsesToDo=1,
             pulsetype="gaussian", kappa=680.0):
        return self.lib.heat_init(self.h, power, energyPerPixel, total_time, loops, repetitionRate_1, pulsesToDo,
                                  PULSETYPES[pulsetype], kappa)

    def array(self, name) -> np.ndarray:
        idx, halo = HEAT_ARRAYS[name]
        m = self.n + 2 if halo else self.n
        flat = np.ctypeslib.as_array(self.lib.heat_array(self.h, idx), shape=(m ** 3,))
        return flat.reshape((m, m, m), order="F")

    def threstime(self) -> np.ndarray:
        flat = np.ctypeslib.as_array(self.lib.heat_array(self.h, 10), shape=(3 * self.n ** 3,))
        return flat.reshape((self.n, self.n, self.n, 3), order="F")

    def scalar(self, name) -> float:
        return self.lib.heat_scalar(self.h, HEAT_SCALARS[name])

    def get_pwr(self) -> float:
        return self.lib.heat_getPwr(self.h)

    def scale_jmean(self, jmean: np.ndarray, nphotons_total: float) -> float:
        assert jmean.flags.f_contiguous and jmean.dtype == np.float64
        return self.lib.heat_scale_jmean(self.h, jmean.ctypes.data, float(nphotons_total))

    def sim_3d(self, jmean: np.ndarray, counter: int):
        assert jmean.flags.f_contiguous and jmean.dtype == np.float64 and jmean.shape == (self.n,) * 3
        self.lib.heat_sim_3d(self.h, jmean.ctypes.data, int(counter))

    def arrhenius(self):
        self.lib.heat_arrhenius(self.h)

    def setup_thermal_coeff(self, ablate_temp: float):
        self.lib.heat_setup_thermal_coeff(self.h, float(ablate_temp))
