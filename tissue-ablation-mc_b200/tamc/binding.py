"""ctypes binding of include/tamc.h.  One MCTransport = one tamc_handle = one GPU = one MPI rank."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

SCATTER = 1
FRESNEL = 2
PERIODIC = 4

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

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


class TamcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tamc error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [
        ("packets", C.c_int64), ("voxel_steps", C.c_int64), ("scatters", C.c_int64), ("absorbed", C.c_int64),
        ("exits", C.c_int64 * 6),
        ("zero_ms", C.c_double), ("kernel_ms", C.c_double), ("allreduce_ms", C.c_double),
        ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("gpu_launches", C.c_int64),
        ("specular", C.c_int64), ("internal_reflections", C.c_int64),
    ]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "exits"}
        d["exits"] = list(self.exits)
        return d


class HeatParams(C.Structure):
    _fields_ = [
        ("power", C.c_double), ("energyPerPixel", C.c_double), ("total_time", C.c_double),
        ("repetitionRate_1", C.c_double), ("ablateTemp", C.c_double),
        ("loops", C.c_int32), ("pulsesToDo", C.c_int32), ("pulsetype", C.c_int32), ("pad_", C.c_int32),
    ]


PULSETYPES = {"tophat": 0, "gaussian": 1, "triangular": 2}
HEAT_ARRAYS = {"temp": 0, "rhokap": 1, "kappa": 2, "density": 3, "heatcap": 4, "coeff": 5, "alpha": 6,
               "watercontent": 7, "Q": 8, "tissue": 9, "threstime": 10, "jmean": 11}
HEAT_SCALARS = {"delt": 0, "time": 1, "total_time": 2, "pulselength": 3, "realPulseLength": 4, "laserOn": 5,
                "pulseCount": 6, "repetitionCount": 7, "laser_flag": 8, "QVapor": 9, "pwr": 10, "counter": 11, "negative_temp": 12}


def lib_path() -> str:
    return os.path.join(_PKG, "libtamc.so")


_lib = None


def lib() -> C.CDLL:
    """Load libtamc.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise TamcError(-1, f"{path} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()'); "
                            "there is no CPU fallback")
    L = C.CDLL(path)
    i, i64, d, p = C.c_int, C.c_int64, C.c_double, C.c_void_p
    sig = {
        "tamc_init": (i, [i, i, i, i, d, d, d, d, C.POINTER(p)]),
        "tamc_finalize": (i, [p]),
        "tamc_set_source_co2": (i, [p, d]),
        "tamc_set_source_gaussian": (i, [p, d]),
        "tamc_set_optics": (i, [p, p, d, d, d, d, i]),
        "tamc_set_optics_grids": (i, [p, p, p, p]),
        "tamc_run": (i, [p, i64, i64, p, C.POINTER(Stats)]),
        "tamc_run_optics": (i, [p, p, d, d, d, d, i, i64, i64, p, C.POINTER(Stats)]),
        "tamc_run_async": (i, [p, i64, i64, i64]),
        "tamc_sync": (i, [p]),
        "tamc_get_jmean": (i, [p, p]),
        "tamc_get_stats": (i, [p, C.POINTER(Stats)]),
        "tamc_seek": (i, [p, i64]),
        "tamc_run_replay": (i, [p, i64, p, p, p, p]),
        "tamc_run_records": (i, [p, i64, i64, i64, p, p]),
        "tamc_comm_unique_id": (i, [p]),
        "tamc_comm_init": (i, [p, i, i, p]),
        "tamc_stream": (p, [p]),
        "tamc_jmean_device": (p, [p]),
        "tamc_rhokap_device": (p, [p]),
        "tamc_pin_host": (i, [p, C.c_uint64]),
        "tamc_unpin_host": (i, [p]),
        "tamc_set_option": (i, [p, C.c_char_p, i64]),
        "tamc_get_option": (i64, [p, C.c_char_p]),
        "tamc_roofline_probe": (i, [p, i64, i64, C.POINTER(d), C.POINTER(i64)]),
        "tamc_selfcheck_launch": (i, [p, i64, i64, C.POINTER(i64), C.POINTER(i64)]),
        "tamc_trace_probe": (i, [p, i64, i64, C.POINTER(d), C.POINTER(i64), C.POINTER(i64)]),
        "tamc_flush_l2": (i, [p, C.c_uint64]),
        "tamc_heat_init": (i, [p, C.POINTER(HeatParams), C.POINTER(d)]),
        "tamc_heat_step": (i, [p, i64]),
        "tamc_coupled_loop": (i, [p, i64, i64, i64, C.POINTER(i64), C.POINTER(i64)]),
        "tamc_heat_array": (i, [p, i, p, i]),
        "tamc_heat_scalar": (i, [p, i, C.POINTER(d)]),
        "tamc_last_error": (C.c_char_p, []),
        "tamc_version": (i, []),
        "tamc_device_count": (i, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype = res
        f.argtypes = args
    _lib = L
    return L


EXPORTS = [
    "tamc_init", "tamc_finalize", "tamc_set_source_co2", "tamc_set_source_gaussian", "tamc_set_optics", "tamc_set_optics_grids", "tamc_run", "tamc_run_optics", "tamc_run_async",
    "tamc_sync", "tamc_get_jmean", "tamc_get_stats", "tamc_seek", "tamc_run_replay", "tamc_run_records",
    "tamc_comm_unique_id", "tamc_comm_init", "tamc_stream", "tamc_jmean_device", "tamc_rhokap_device",
    "tamc_pin_host", "tamc_unpin_host", "tamc_set_option", "tamc_get_option", "tamc_roofline_probe", "tamc_selfcheck_launch", "tamc_trace_probe",
    "tamc_flush_l2", "tamc_last_error", "tamc_version", "tamc_device_count",
    "tamc_heat_init", "tamc_heat_step", "tamc_coupled_loop", "tamc_heat_array", "tamc_heat_scalar",
]


def _ck(rc):
    if rc != 0:
        raise TamcError(rc, lib().tamc_last_error().decode(errors="replace"))


def device_count() -> int:
    return lib().tamc_device_count()


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _ck(lib().tamc_comm_unique_id(buf))
    return buf.raw


def pin_host(a: np.ndarray):
    _ck(lib().tamc_pin_host(a.ctypes.data, a.nbytes))


def unpin_host(a: np.ndarray):
    _ck(lib().tamc_unpin_host(a.ctypes.data))


def delta_default(zmax, nzg):
    return 1.0e-8 * (2.0 * zmax / nzg)  # mcpolar.f90:112


class MCTransport:
    """Device-resident twin of the state the reference's photon loop reads and writes."""

    def __init__(self, nxg, nyg, nzg, xmax, ymax, zmax, delta=None, device=0):
        self.L = lib()
        self.nxg, self.nyg, self.nzg = int(nxg), int(nyg), int(nzg)
        self.xmax, self.ymax, self.zmax = float(xmax), float(ymax), float(zmax)
        self.delta = float(delta) if delta is not None else delta_default(self.zmax, self.nzg)
        self.h = C.c_void_p()
        _ck(self.L.tamc_init(int(device), self.nxg, self.nyg, self.nzg, self.xmax, self.ymax, self.zmax,
                             self.delta, C.byref(self.h)))
        self.device = int(device)

    # -- lifecycle ----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.L.tamc_finalize(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- set-up -------------------------------------------------------------------------------
    @property
    def rhokap_shape(self):
        return (self.nxg + 2, self.nyg + 2, self.nzg + 2)

    @property
    def jmean_shape(self):
        return (self.nxg, self.nyg, self.nzg)

    def set_source_co2(self, spot_diameter_cm: float):
        _ck(self.L.tamc_set_source_co2(self.h, float(spot_diameter_cm)))

    def set_source_gaussian(self, sigma_cm: float):
        """Gaussian beam through rang() (sourceph.f90:73-101); set_source_co2 switches back to the disk."""
        _ck(self.L.tamc_set_source_gaussian(self.h, float(sigma_cm)))

    def set_optics(self, rhokap, albedo, hgg, n1=1.0, n2=1.0, flags=0):
        """rhokap: (nxg+2, nyg+2, nzg+2) Fortran-ordered fp64 with halo, or None to keep the resident grid."""
        ptr = None
        if rhokap is not None:
            rk = np.asarray(rhokap, dtype=np.float64)
            if rk.shape != self.rhokap_shape:
                raise ValueError(f"rhokap must have shape {self.rhokap_shape} (halo included)")
            if not rk.flags.f_contiguous:
                rk = np.asfortranarray(rk)
            self._keep = rk
            ptr = rk.ctypes.data
        _ck(self.L.tamc_set_optics(self.h, ptr, float(albedo), float(hgg), float(n1), float(n2), int(flags)))

    def set_optics_grids(self, albedo=None, hgg=None, n=None):
        """EXTENSION: per-voxel albedo / hgg / refractive index, each shaped like rhokap (halo included) or None = scalar."""
        ptrs, keep = [], []
        for a in (albedo, hgg, n):
            if a is None:
                ptrs.append(None)
                continue
            b = np.asfortranarray(np.asarray(a, dtype=np.float64))
            if b.shape != self.rhokap_shape:
                raise ValueError(f"optics grids must have shape {self.rhokap_shape} (halo included)")
            keep.append(b)
            ptrs.append(b.ctypes.data)
        _ck(self.L.tamc_set_optics_grids(self.h, *ptrs))

    def set_option(self, name: str, value: int):
        _ck(self.L.tamc_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        return self.L.tamc_get_option(self.h, name.encode())

    # -- hot path -----------------------------------------------------------------------------
    def new_jmean(self) -> np.ndarray:
        return np.zeros(self.jmean_shape, dtype=np.float64, order="F")

    def run(self, nphotons, seed, out=None):
        """tamc_run: blocking MC call; returns (jmeanGLOBAL unscaled, stats dict)."""
        jm = out if out is not None else self.new_jmean()
        assert jm.flags.f_contiguous and jm.dtype == np.float64 and jm.shape == self.jmean_shape
        st = Stats()
        _ck(self.L.tamc_run(self.h, int(nphotons), int(seed), jm.ctypes.data, C.byref(st)))
        return jm, st.as_dict()

    def run_optics(self, rhokap, albedo, hgg, nphotons, seed, n1=1.0, n2=1.0, flags=0, out=None):
        """tamc_run_optics: tamc_set_optics + tamc_run in one call (copies overlapped with the transport when the
        arrays are page-locked and the scatter loop is off); returns (jmeanGLOBAL unscaled, stats dict)."""
        ptr = None
        if rhokap is not None:
            rk = np.asarray(rhokap, dtype=np.float64)
            if rk.shape != self.rhokap_shape:
                raise ValueError(f"rhokap must have shape {self.rhokap_shape} (halo included)")
            if not rk.flags.f_contiguous:
                rk = np.asfortranarray(rk)
            self._keep = rk
            ptr = rk.ctypes.data
        jm = out if out is not None else self.new_jmean()
        assert jm.flags.f_contiguous and jm.dtype == np.float64 and jm.shape == self.jmean_shape
        st = Stats()
        _ck(self.L.tamc_run_optics(self.h, ptr, float(albedo), float(hgg), float(n1), float(n2), int(flags),
                                   int(nphotons), int(seed), jm.ctypes.data, C.byref(st)))
        return jm, st.as_dict()

    def run_async(self, nphotons, seed, first_packet_id=-1):
        _ck(self.L.tamc_run_async(self.h, int(nphotons), int(seed), int(first_packet_id)))

    def sync(self):
        _ck(self.L.tamc_sync(self.h))

    def get_jmean(self, out=None):
        jm = out if out is not None else self.new_jmean()
        _ck(self.L.tamc_get_jmean(self.h, jm.ctypes.data))
        return jm

    def get_stats(self):
        st = Stats()
        _ck(self.L.tamc_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def seek(self, next_packet_id):
        _ck(self.L.tamc_seek(self.h, int(next_packet_id)))

    def run_replay(self, offsets, draws, want_records=True, want_jmean=True):
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        drw = np.ascontiguousarray(draws, dtype=np.float64)
        n = off.size - 1
        rec = np.zeros(n, dtype=RECORD_DTYPE) if want_records else None
        jm = self.new_jmean() if want_jmean else None
        _ck(self.L.tamc_run_replay(self.h, n, off.ctypes.data, drw.ctypes.data if drw.size else None,
                                   rec.ctypes.data if rec is not None else None,
                                   jm.ctypes.data if jm is not None else None))
        return rec, jm

    def run_records(self, nphotons, seed, first_packet_id=0):
        rec = np.zeros(int(nphotons), dtype=RECORD_DTYPE)
        jm = self.new_jmean()
        _ck(self.L.tamc_run_records(self.h, int(nphotons), int(seed), int(first_packet_id), rec.ctypes.data,
                                    jm.ctypes.data))
        return rec, jm

    # -- heat / ablation step on the device (3dFD.f90) -----------------------------------------
    def heat_init(self, power=70.0, energyPerPixel=400.0, total_time=2.0, loops=1, repetitionRate_1=1e7,
                  pulsesToDo=1, pulsetype="gaussian", ablateTemp=500.0) -> float:
        """initThermalCoeff + the driver's temperature set-up; returns delt."""
        prm = HeatParams(power, energyPerPixel, total_time, repetitionRate_1, ablateTemp, loops, pulsesToDo,
                         PULSETYPES[pulsetype], 0)
        delt = C.c_double(0)
        _ck(self.L.tamc_heat_init(self.h, C.byref(prm), C.byref(delt)))
        return delt.value

    def heat_step(self, nphotons_times_numproc: int):
        _ck(self.L.tamc_heat_step(self.h, int(nphotons_times_numproc)))

    def coupled_loop(self, nphotons: int, seed: int, max_iterations: int = -1):
        it, pk = C.c_int64(0), C.c_int64(0)
        _ck(self.L.tamc_coupled_loop(self.h, int(nphotons), int(seed), int(max_iterations), C.byref(it), C.byref(pk)))
        return it.value, pk.value

    def heat_array(self, name: str) -> np.ndarray:
        n = self.nxg
        if name == "threstime":
            out = np.zeros((n, n, n, 3), dtype=np.float64, order="F")
        elif name in ("watercontent", "Q", "tissue", "jmean"):
            out = np.zeros((n, n, n), dtype=np.float64, order="F")
        else:
            out = np.zeros((n + 2, n + 2, n + 2), dtype=np.float64, order="F")
        _ck(self.L.tamc_heat_array(self.h, HEAT_ARRAYS[name], out.ctypes.data, 0))
        return out

    def heat_upload(self, name: str, a: np.ndarray):
        a = np.asfortranarray(a, dtype=np.float64)
        _ck(self.L.tamc_heat_array(self.h, HEAT_ARRAYS[name], a.ctypes.data, 1))

    def heat_scalar(self, name: str) -> float:
        v = C.c_double(0)
        _ck(self.L.tamc_heat_scalar(self.h, HEAT_SCALARS[name], C.byref(v)))
        return v.value

    # -- multi-GPU ----------------------------------------------------------------------------
    def comm_init(self, nranks, rank, unique_id: bytes):
        assert len(unique_id) == 128
        _ck(self.L.tamc_comm_init(self.h, int(nranks), int(rank), unique_id))

    # -- residency / measurement --------------------------------------------------------------
    @property
    def stream(self) -> int:
        return int(self.L.tamc_stream(self.h) or 0)

    @property
    def jmean_device(self) -> int:
        return int(self.L.tamc_jmean_device(self.h) or 0)

    def roofline_probe(self, nphotons, seed=1):
        ms, steps = C.c_double(0), C.c_int64(0)
        _ck(self.L.tamc_roofline_probe(self.h, int(nphotons), int(seed), C.byref(ms), C.byref(steps)))
        return ms.value, steps.value

    def trace_probe(self, npackets, seed=1):
        """Scatter-regime roofline probe (tamc_trace_probe): the recorded voxel-index stream replayed as loads + REDs."""
        ms, steps, reds = C.c_double(0), C.c_int64(0), C.c_int64(0)
        _ck(self.L.tamc_trace_probe(self.h, int(npackets), int(seed), C.byref(ms), C.byref(steps), C.byref(reds)))
        return {"ms": ms.value, "voxel_steps": steps.value, "reds": reds.value, "packets": int(npackets),
                "voxel_steps_per_s": steps.value / (ms.value * 1e-3) if ms.value > 0 else 0.0,
                "what": "grid-lookup / L2-atomic roofline of this workload: the voxel-index stream of real packets (recorded once), replayed "
                        "as one 8-byte opacity load per voxel-step + one fp64 RED per voxel left, on the layout the kernel itself uses (separate arrays while the grids fit L2, interleaved records beyond), "
                        "no transport arithmetic"}

    def selfcheck_launch(self, n, seed=1):
        """(draws the fp32 launch-voxel pass handed to fp64, draws where it kept a different voxel -- must be 0)."""
        fb, bad = C.c_int64(0), C.c_int64(0)
        _ck(self.L.tamc_selfcheck_launch(self.h, int(n), int(seed), C.byref(fb), C.byref(bad)))
        return fb.value, bad.value

    def flush_l2(self, nbytes=256 << 20):
        _ck(self.L.tamc_flush_l2(self.h, int(nbytes)))
