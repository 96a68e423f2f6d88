#!/usr/bin/env python
"""Golden vectors for the heat / ablation step from the REFERENCE'S OWN SOURCE TEXT (3dFD.f90, thermalConst_mod.f90 and
the driver lines of mcpolar.f90), executed by oracle/f90interp.py for ONE rank.

    python tests/golden/make_reference_heat_vectors.py        # -> tests/golden/reference_interp_heat.json.gz

Executed from the reference's text (read from /root/reference/src at run time, nothing restated):
  * mcpolar.f90:61-73 (state reset), :123-129 (temperature and its boundary planes), :134-140 (total_time override), :174
    (the scaling of jmeanGLOBAL, with the reference's getPwr), :180-185 (Arrhenius, setupThermalCoeff, counter, jmean = 0);
  * 3dFD.f90 initThermalCoeff :249-251, :263-291 and the realPulseLength line of the selected pulse type (:297 / :300 / :303);
  * 3dFD.f90 heat_sim_3D :45-72, :79-94, :102-110 and the whole time loop :113-215;
  * 3dFD.f90 Arrhenius, setupThermalCoeff, getPwrTopHat / getPwrGaussian / getPwrTriangular and every function of
    thermalConst_mod.f90, as ordinary procedure calls.
What the harness does instead of the reference (each place cites the line it stands for):
  * allocations (subs.f90:57-66, 3dFD.f90:75-78, :254-260): shapes and lower bounds only;
  * MPI for one rank: MPI_allREDUCE (mcpolar.f90:173) and the two mpi_scatter (3dFD.f90:95-99) are copies, MPI_Sendrecv with
    no neighbour (:190-197) leaves the halo planes alone, the two mpi_allgather (:218-222) copy the interior z planes back;
  * `select case(trim(pulsetype))` (:294-307): the harness binds the procedure pointer getPwr and runs the case's one line;
  * the grid size: constants.f90:12's nxg = nyg = nzg = 80 are compile-time parameters; the harness sets them to a small n
    before anything is allocated (the reference's user edits that line to change the grid);
  * Heat's module variable pulsesDone is never initialised upstream (3dFD.f90:12, first use :203): static storage, 0;
  * in the `cases` the tally jmean of every MC call is synthetic (seeded, sparse, the same numbers the test feeds the
    oracle); in `coupled` it is the reference's own photon loop (mcpolar.f90:151-170 with ran2.f, sourceph.f90, inttau2.f90)
    on the opacity the previous property update left behind -- the whole `do while(time <= total_time)` iteration.
"""
import gzip
import json
import os
import struct
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.f90interp import Cell, FArray, Interpreter  # noqa: E402

REF = os.environ.get("TAMC_REFERENCE_SRC", "/root/reference/src")
FILES = ("constants.f90", "photon_vars.f90", "iarray.f90", "opt_prop.f90", "ch_opt.f90", "gridset.f90", "thermalConst_mod.f90",
         "3dFD.f90")
MC, FD = os.path.join(REF, "mcpolar.f90"), os.path.join(REF, "3dFD.f90")
PWR = {"tophat": ("getpwrtophat", 297), "gaussian": ("getpwrgaussian", 300), "triangular": ("getpwrtriangular", 303)}


def hexf(x):
    return struct.pack(">d", float(x)).hex()


def hexa(a):
    """An array as the hex of its column-major binary64 bytes (big-endian)."""
    return np.asarray(a, dtype=">f8").tobytes(order="F").hex()


def exact_sum(a):
    """Correctly rounded sum (math.fsum): the same number whatever the memory order of the array."""
    import math

    return math.fsum(np.asarray(a, dtype=np.float64).ravel().tolist())


def synthetic_jmean(rng, n):
    """What tests/test_oracle_reference_heat.py feeds the oracle: a sparse random tally (path lengths * opacity per voxel)."""
    return np.asfortranarray(rng.uniform(0.0, 3.0e4, (n, n, n)) * (rng.uniform(size=(n, n, n)) < 0.6))


class Machine:
    def __init__(self, n, xmax, ymax, zmax, pulsetype, power, energy, total_time, loops, rep_rate, pulses, ablate, nphotons,
                 extra_files=()):
        it = self.it = Interpreter()
        it.skipped_calls.add("checkallocate")
        for f in FILES + tuple(extra_files):
            it.load(os.path.join(REF, f))
        for k in ("nxg", "nyg", "nzg"):
            it.var("constants", k).set(n)                       # constants.f90:12 (see the header)
        self.n = n
        h = n + 2
        # subs.f90:57-66 (alloc_array) and 3dFD.f90:254-260 (initThermalCoeff)
        for name, shape, lb in (("xface", (n + 1,), None), ("yface", (n + 1,), None), ("zface", (n + 1,), None),
                                ("rhokap", (h, h, h), (0, 0, 0)), ("jmean", (n, n, n), None), ("jmeanglobal", (n, n, n), None),
                                ("tissue", (n, n, n), None), ("temp", (h, h, h), (0, 0, 0)), ("threstime", (n, n, n, 3), None)):
            it.allocate("iarray", name, shape, lb)
        for name in ("coeff", "alpha", "kappa", "density", "heatcap"):
            it.allocate("heat", name, (h, h, h), (0, 0, 0))
        it.allocate("heat", "q", (n, n, n))
        it.allocate("heat", "watercontent", (n, n, n))
        it.var("heat", "pulsesdone").set(0)                     # never initialised upstream: static storage
        # the main program's frame: its own variables (mcpolar.f90:26-35) + what it uses (:6-9, :19-21)
        fr = self.fr = {"n": Cell("i"), "counter": Cell("i"), "id": Cell("i", 0), "numproc": Cell("i", 1), "nphotons": Cell("i", nphotons),
                        "xmax": Cell("r", xmax), "ymax": Cell("r", ymax), "zmax": Cell("r", zmax), "ablatetemp": Cell("r", ablate)}
        for mod in ("constants", "iarray", "opt_prop"):
            for k, v in it.modules[mod].vars.items():
                if v is not None:
                    fr[k] = v
        for k in ("power", "delt", "energyperpixel", "laser_flag", "laseron", "loops", "pulsecount", "pulsestodo", "pulseflag",
                  "repetitioncount", "repetitionrate_1", "time", "total_time", "realpulselength", "pulselength"):
            fr[k] = it.var("heat", k)                            # mcpolar.f90:19-21  use Heat, only : ...
        # res/input.params (mcpolar.f90:79-94)
        fr["total_time"].set(total_time); fr["loops"].set(loops); fr["repetitionrate_1"].set(rep_rate); fr["power"].set(power)
        fr["energyperpixel"].set(energy); fr["pulsestodo"].set(pulses); fr["pulsetype"].set(pulsetype)
        it.call("init_opt1", [])                                  # mcpolar.f90:101
        it.call("gridset", [fr["xmax"], fr["ymax"], fr["zmax"], fr["id"]])   # :109
        it.run_block(MC, 61, 73, fr)                              # N, tissue, ThresTime, time ... counter
        it.run_block(MC, 123, 129, fr)                            # temp and its six boundary planes
        self.init_thermal_coeff(pulsetype)                        # :130
        it.run_block(MC, 134, 140, fr)                            # total_time override

    def heat_frame(self):
        """Host association of a procedure contained in module Heat."""
        return {k: v for k, v in self.it.modules["heat"].vars.items() if v is not None}

    def init_thermal_coeff(self, pulsetype):
        it, fr = self.it, self.fr
        f = self.heat_frame()
        for mod in ("thermalconstants",):                         # 3dFD.f90:235  use thermalConstants
            f.update({k: v for k, v in it.modules[mod].vars.items() if v is not None})
        for k in ("nxg", "nyg", "nzg", "spotsperrow", "spotspercol", "pulsetype"):   # :236
            f[k] = it.var("constants", k)
        f.update({"delt": fr["delt"], "numpoints": fr["n"], "xmax": fr["xmax"], "ymax": fr["ymax"], "zmax": fr["zmax"],
                  "numproc": fr["numproc"]})                      # the dummies
        for k in ("densitytmp", "alphatmp", "kappatmp", "heatcaptmp", "constd"):     # :246
            f[k] = Cell("r")
        it.run_block(FD, 249, 251, f)
        it.run_block(FD, 263, 291, f)
        name, line = PWR[pulsetype]
        it.alias["getpwr"] = name                                 # getPwr => getPwr...
        it.run_block(FD, line, line, f)

    def heat_sim_3d(self):
        """call heat_sim_3d(jmeanGLOBAL, temp, N, id, numproc, new_comm, right, left, counter)  -- mcpolar.f90:178"""
        it, fr, n = self.it, self.fr, self.n
        f = self.heat_frame()
        f["qvapor"] = it.var("thermalconstants", "qvapor")        # 3dFD.f90:25
        f.update({"jmean": fr["jmeanglobal"], "temp": fr["temp"], "numpoints": fr["n"], "id": fr["id"], "numproc": fr["numproc"],
                  "counter": fr["counter"]})
        for k in ("u_xx", "u_yy", "u_zz", "tempincrease", "kappaminhalf", "kappaplushalf", "energyincrease", "heatcapminhalf",
                  "densityplushalf", "densityminhalf", "heatcapplushalf", "a", "b", "d"):          # :38-39
            f[k] = Cell("r")
        for k in ("i", "j", "k", "p", "size_x", "size_y", "size_z", "xi", "yi", "zi", "xf", "yf", "zf", "n", "tag", "zsta", "zfin"):
            f[k] = Cell("i")                                      # :41
        it.run_block(FD, 45, 72, f)
        zi, zf = f["zi"].v, f["zf"].v
        f["t0"] = FArray("r", (n + 2, n + 2, zf - zi + 3), (0, 0, zi - 1))       # :75-78
        f["tn"] = FArray("r", (n + 2, n + 2, zf - zi + 3), (0, 0, zi - 1))
        f["jtmp"] = FArray("r", (n, n, zf - zi + 1), (1, 1, zi))
        f["qtmp"] = FArray("r", (n, n, zf - zi + 1), (1, 1, zi))
        it.run_block(FD, 79, 94, f)
        f["jtmp"].a[...] = f["jmean"].a[:, :, zi - 1:zf]         # :95-96  mpi_scatter, one rank
        f["qtmp"].a[...] = f["q"].a[:, :, zi - 1:zf]             # :98-99
        it.run_block(FD, 102, 110, f)
        it.run_block(FD, 113, 215, f)                             # the time loop
        fr["temp"].a[:, :, zi:zf + 1] = f["t0"].a[:, :, 1:zf - zi + 2]           # :218-219  mpi_allgather, one rank
        f["q"].a[:, :, zi - 1:zf] = f["qtmp"].a                  # :221-222

    def iteration(self, jmean):
        """One pass of `do while(time <= total_time)`, mcpolar.f90:148-186; returns False when the loop condition fails."""
        it, fr = self.it, self.fr
        if not fr["time"].v <= fr["total_time"].v:
            return False
        scale = None
        if fr["laser_flag"].v:                                    # :149
            fr["jmean"].a[...] = jmean                            # the MC call's tally (:151-170), synthetic here
            fr["jmeanglobal"].a[...] = fr["jmean"].a              # :173  MPI_allREDUCE, one rank
            before = fr["jmeanglobal"].a.copy()
            it.run_block(MC, 174, 174, fr)
            nz = before != 0
            scale = float((fr["jmeanglobal"].a[nz] / before[nz])[0]) if nz.any() else None
        self.heat_sim_3d()                                        # :178
        it.run_block(MC, 180, 185, fr)
        return True

    def snapshot(self):
        it, fr = self.it, self.fr
        out = {k: hexa(fr[k].a) for k in ("temp", "rhokap", "tissue", "threstime")}
        out.update({k: hexa(it.var("heat", k).a) for k in ("kappa", "density", "heatcap", "coeff", "alpha", "watercontent", "q")})
        out.update({k: hexf(it.var("heat", k).v) for k in ("time", "delt", "laseron", "pulsecount", "repetitioncount")})
        out["laser_flag"] = int(it.var("heat", "laser_flag").v)
        out["pulsesdone"] = it.var("heat", "pulsesdone").v
        return out


TRANSPORT_FILES = ("ran2.f", "sourceph.f90", "inttau2.f90")


def run_coupled_case(n, npackets, pulsetype, power, energy, loops, iterations, ablate=150.0, rank=0, extents=(0.03, 0.03, 0.06)):
    """mcpolar.f90:148-186 as it stands: the photon loop (:151-170, the reference's ran2 stream running on from call to call)
    on the opacity the last property update left behind, the tally scaled (:174) into the heat step, Arrhenius, the property
    update -- every iteration, until the reference stops."""
    m = Machine(n, *extents, pulsetype, power, energy, 2.0, loops, 1e7, 1, ablate, npackets,
                extra_files=TRANSPORT_FILES)
    it, fr = m.it, m.fr
    for k, t in (("iseed", "i"), ("delta", "r"), ("nscatt", "r"), ("tflag", "l"), ("xcell", "i"), ("ycell", "i"), ("zcell", "i"), ("j", "i")):
        fr[k] = Cell(t)
    for k, v in it.modules["photon_vars"].vars.items():
        fr[k] = v
    fr["id"].set(rank)
    it.run_block(MC, 97, 98, fr)                                  # seed rule
    it.run_block(MC, 112, 113, fr)                                # delta
    ran2, wall = it.procs["ran2"], it.procs["wall_dist"]
    h = lambda k: it.var("heat", k).v
    case = {"n": n, "extents": list(extents), "pulsetype": pulsetype, "power": power, "energyPerPixel": energy, "loops": loops,
            "ablateTemp": ablate, "nphotons": npackets, "rank": rank, "init": m.snapshot(), "steps": []}
    for i in range(iterations):
        if not fr["time"].v <= fr["total_time"].v:               # :148
            break
        mc = None
        try:
            if fr["laser_flag"].v:                                # :149
                d0, s0 = ran2.calls, wall.calls
                it.run_block(MC, 151, 170, fr)                    # do j = 1, nphotons ... end do
                mc = {"draws": ran2.calls - d0, "voxel_steps": wall.calls - s0, "iseed": fr["iseed"].v,
                      "jmean_sum": hexf(exact_sum(fr["jmean"].a)), "jmean_nonzero": int((fr["jmean"].a != 0).sum())}
                fr["jmeanglobal"].a[...] = fr["jmean"].a          # :173  MPI_allREDUCE, one rank
                it.run_block(MC, 174, 174, fr)
            m.heat_sim_3d()                                       # :178
            it.run_block(MC, 180, 185, fr)
        except Exception as e:
            if "ERROR STOP" not in str(e):
                raise
            case["error_stop"] = {"iteration": i, "where": str(e).split(" at ")[-1].replace(REF + "/", "")}
            break
        snap = m.snapshot()
        rk = fr["rhokap"].a
        case["steps"].append({"mc": mc, "digest": snap if i % 9 == 0 else None, "temp_max": hexf(fr["temp"].a.max()),
                              "temp_sum": hexf(exact_sum(fr["temp"].a)), "ablated": int((rk[1:-1, 1:-1, 1:-1] == 0).sum()),
                              "q_sum": hexf(exact_sum(it.var("heat", "q").a)), "rhokap_sum": hexf(exact_sum(rk)),
                              "tissue_sum": hexf(exact_sum(fr["tissue"].a)), "time": snap["time"], "laser_flag": snap["laser_flag"]})
    case["last"] = m.snapshot() if "error_stop" not in case else None
    print("coupled", pulsetype, n, "iterations", len(case["steps"]), case.get("error_stop"), flush=True)
    return case


def run_case(n, pulsetype, power, energy, loops, iterations, seed, ablate=150.0, total_time=2.0, rep_rate=1e7, pulses=1,
             nphotons=1000, extents=(0.03, 0.03, 0.06)):
    m = Machine(n, *extents, pulsetype, power, energy, total_time, loops, rep_rate, pulses, ablate, nphotons)
    h = lambda k: m.it.var("heat", k).v
    case = {"n": n, "extents": list(extents), "pulsetype": pulsetype, "power": power, "energyPerPixel": energy, "loops": loops,
            "total_time_in": total_time, "repetitionRate_1": rep_rate, "pulsesToDo": pulses, "ablateTemp": ablate,
            "nphotons": nphotons, "seed": seed,
            "init": {"delt": hexf(h("delt")), "total_time": hexf(h("total_time")), "pulselength": hexf(h("pulselength")),
                     "realPulseLength": hexf(h("realpulselength")), "QVapor": hexf(m.it.var("thermalconstants", "qvapor").v),
                     "volumeVoxel": hexf(h("volumevoxel")), "massVoxel": hexf(h("massvoxel")), **m.snapshot()},
            "steps": []}
    rng = np.random.default_rng(seed)
    for i in range(iterations):
        jm = synthetic_jmean(rng, n)
        try:
            if not m.iteration(jm):
                break
        except Exception as e:
            # thermalConst_mod.f90:20-23: airThermalCond stops the program on a negative temperature -- the explicit scheme
            # has diverged in an air voxel, and the reference's run ends here
            if "ERROR STOP" not in str(e):
                raise
            case["error_stop"] = {"iteration": i, "where": str(e).split(" at ")[-1].replace(REF + "/", "")}
            break
        snap = m.snapshot()
        rk = m.fr["rhokap"].a
        case["steps"].append({"digest": snap if i in (0, iterations // 2, iterations - 1) else None,
                              "temp_max": hexf(m.fr["temp"].a.max()), "temp_sum": hexf(exact_sum(m.fr["temp"].a)),
                              "ablated": int((rk[1:-1, 1:-1, 1:-1] == 0).sum()), "q_sum": hexf(exact_sum(m.it.var("heat", "q").a)),
                              "tissue_sum": hexf(exact_sum(m.fr["tissue"].a)), "rhokap_sum": hexf(exact_sum(rk)),
                              "time": snap["time"], "laser_flag": snap["laser_flag"]})
    if "error_stop" not in case:
        case["final"] = m.snapshot()
    print(pulsetype, n, "iterations", len(case["steps"]), case.get("error_stop"), flush=True)
    return case


def main():
    t0 = time.time()
    out = {"what": "the heat / ablation step as the reference's own Fortran text computes it for one rank, executed by oracle/f90interp.py "
                   "(tests/golden/make_reference_heat_vectors.py)",
           "arrays": "hex of big-endian binary64, column-major; temp / rhokap / kappa / density / heatcap / coeff / alpha with halo"}
    # strengths chosen so that the runs pass through boiling, water loss and ablation before the explicit scheme diverges
    # in the air voxels (a few iterations after the first ablation on grids this small) and the reference stops
    out["cases"] = [run_case(5, "tophat", 40.0, 4000.0, 2, 60, 5, ablate=105.0, nphotons=1000000),
                    run_case(5, "tophat", 40.0, 4000.0, 2, 60, 5, ablate=120.0, nphotons=600000),
                    run_case(5, "triangular", 300.0, 4000.0, 2, 60, 5, ablate=110.0, nphotons=3000000),
                    run_case(5, "gaussian", 70.0, 400.0, 2, 60, 5, ablate=110.0, nphotons=200000),
                    # three short pulses with pauses, no boiling: the laser on / off bookkeeping of 3dFD.f90:199-214
                    run_case(6, "tophat", 70.0, 40.0, 1, 70, 11, ablate=150.0, nphotons=5000000, rep_rate=0.1, pulses=3,
                             extents=(0.03, 0.03, 0.03))]
    out["coupled"] = [run_coupled_case(8, 150, "tophat", 8.0, 400.0, 1, 40)]
    print("cases done", round(time.time() - t0, 1), flush=True)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_interp_heat.json.gz")
    with gzip.GzipFile(path, "wb", compresslevel=9, mtime=0) as g:
        g.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
