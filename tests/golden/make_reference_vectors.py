#!/usr/bin/env python
"""Golden vectors from the REFERENCE'S OWN SOURCE TEXT, executed here by oracle/f90interp.py.

Run in the build container (reads /root/reference/src, which does not travel; the vectors it writes are committed):

    python tests/golden/make_reference_vectors.py            # -> tests/golden/reference_interp.json.gz

What is executed, all of it read from the reference at run time and none of it restated:
  * ran2.f (the generator), with mcpolar.f90:97-98 (the per-rank seed rule) run as a statement range;
  * subs.f90:57-62's allocations are mirrored by the harness (shapes and lower bounds only: `allocate` is not interpreted);
  * ch_opt.f90 init_opt1, gridset.f90 gridset (faces + uniform rhokap), mcpolar.f90:112 (delta);
  * mcpolar.f90:153-169 -- the body of the photon loop: sourcephCO2 -> tauint1 -> the stub -- once per packet;
  * stokes.f90 (compiled upstream but never called): a chain of direction updates, Henyey-Greenstein and isotropic;
  * rang / ranu (sourceph.f90:52-101) and repeat_bounds (inttau2.f90:242-279), dead code upstream, as unit calls;
  * the scatter loop SURVEY 3.3 specifies around those routines (`ran2 < albedo ? stokes : absorbed ; tauint1`): its
    three lines are the harness's, every routine it calls is the reference's.
The harness counts calls (ran2 = draws, wall_dist = voxel-steps) and reads module variables; it never computes physics.
"""
import json
import os
import struct
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.f90interp import Cell, Interpreter  # noqa: E402

REF = os.environ.get("TAMC_REFERENCE_SRC", "/root/reference/src")
FILES = ("ran2.f", "constants.f90", "photon_vars.f90", "iarray.f90", "opt_prop.f90", "ch_opt.f90", "gridset.f90",
         "sourceph.f90", "inttau2.f90", "stokes.f90")


def hexf(x):
    """binary64 as 16 hex digits: the vectors are compared bit for bit."""
    return struct.pack(">d", float(x)).hex()


def machine(xmax, ymax, zmax, rank, kappa=None):
    """The reference's modules loaded and set up as mcpolar.f90 does before its photon loop; returns (interpreter, frame)."""
    it = Interpreter()
    for f in FILES:
        it.load(os.path.join(REF, f))
    n = [it.var("constants", k).v for k in ("nxg", "nyg", "nzg")]
    # subs.f90:57-62 (alloc_array): xface(nxg+1) ..., rhokap(0:nxg+1, 0:nyg+1, 0:nzg+1), jmean(nxg, nyg, nzg)
    it.allocate("iarray", "xface", (n[0] + 1,))
    it.allocate("iarray", "yface", (n[1] + 1,))
    it.allocate("iarray", "zface", (n[2] + 1,))
    it.allocate("iarray", "rhokap", (n[0] + 2, n[1] + 2, n[2] + 2), (0, 0, 0))
    it.allocate("iarray", "jmean", (n[0], n[1], n[2]))
    # the variables of the main program that the interpreted statement ranges name (mcpolar.f90:26-35)
    fr = {"xmax": Cell("r", xmax), "ymax": Cell("r", ymax), "zmax": Cell("r", zmax), "id": Cell("i", rank), "iseed": Cell("i"),
          "delta": Cell("r"), "nscatt": Cell("r"), "tflag": Cell("l"), "xcell": Cell("i"), "ycell": Cell("i"), "zcell": Cell("i"),
          "j": Cell("i"), "nphotons": Cell("i")}
    for mod in ("constants", "photon_vars", "iarray", "opt_prop"):        # mcpolar.f90:6-9 `use`
        for k, v in it.modules[mod].vars.items():
            if v is not None:
                fr[k] = v
    mc = os.path.join(REF, "mcpolar.f90")
    it.run_block(mc, 97, 98, fr)                                           # seed rule
    it.call("init_opt1", [])                                               # mcpolar.f90:101
    it.call("gridset", [fr["xmax"], fr["ymax"], fr["zmax"], fr["id"]])    # mcpolar.f90:109
    it.run_block(mc, 112, 113, fr)                                         # delta, nscatt
    if kappa is not None:
        # a turbid medium for the scatter loop -- test INPUT, set where 3dFD.f90:334-353 sets it in the coupled run: the
        # optical properties in opt_prop and the opacity of every voxel (the halo keeps gridset's 0.)
        for k, v in kappa.items():
            it.var("opt_prop", k).set(v)
        it.var("iarray", "rhokap").a[1:-1, 1:-1, 1:-1] = kappa["kappa"]
    return it, fr, n


def packet_row(it, fr, draws, steps, extra=()):
    g = lambda m, k: it.var(m, k).v
    return [hexf(g("photon_vars", "xp")), hexf(g("photon_vars", "yp")), hexf(g("photon_vars", "zp")),
            hexf(g("photon_vars", "nxp")), hexf(g("photon_vars", "nyp")), hexf(g("photon_vars", "nzp")),
            fr["xcell"].v, fr["ycell"].v, fr["zcell"].v, int(fr["tflag"].v), draws, steps, *extra]


def sparse(jm):
    import numpy as np

    idx = np.argwhere(jm != 0.0)
    return [[int(i) + 1, int(j) + 1, int(k) + 1, hexf(jm[i, j, k])] for i, j, k in idx]


def shipped(rank, npackets):
    """The shipped configuration (res/input.params extents, init_opt1's optics): the photon loop body per packet."""
    it, fr, n = machine(0.03, 0.03, 0.06, rank)
    seed0 = fr["iseed"].v
    ran2, wall = it.procs["ran2"], it.procs["wall_dist"]
    probe = Cell("i", seed0)
    first_draws = [hexf(it.call("ran2", [probe], want_result=True)) for _ in range(12)]
    # a fresh machine for the packets: ran2 keeps state in SAVE'd locals
    it, fr, n = machine(0.03, 0.03, 0.06, rank)
    ran2, wall = it.procs["ran2"], it.procs["wall_dist"]
    mc = os.path.join(REF, "mcpolar.f90")
    rows = []
    for _ in range(npackets):
        d0, s0 = ran2.calls, wall.calls
        it.run_block(mc, 153, 169, fr)                                     # tflag = F; sourcephCO2; tauint1; the stub
        rows.append(packet_row(it, fr, ran2.calls - d0, wall.calls - s0))
    faces = {k: [hexf(x) for x in it.var("iarray", k).a] for k in ("xface", "yface", "zface")}
    rk = it.var("iarray", "rhokap").a
    optics = {k: hexf(it.var("opt_prop", k).v) for k in ("hgg", "g2", "mua", "mus", "kappa", "albedo", "mu_water", "mu_protein")}
    return {"rank": rank, "seed": seed0, "delta": hexf(fr["delta"].v), "kappa": hexf(it.var("opt_prop", "kappa").v),
            "faces": faces, "optics": optics,
            "rhokap_interior": sorted({hexf(x) for x in rk[1:-1, 1:-1, 1:-1].ravel()}),
            "rhokap_halo_sum": hexf(float(rk.sum() - rk[1:-1, 1:-1, 1:-1].sum())),
            "grid": n, "extents": [0.03, 0.03, 0.06], "first_draws": first_draws, "packets": rows,
            "jmean": sparse(it.var("iarray", "jmean").a), "iseed_after": fr["iseed"].v}


def crater(rank, npackets):
    """The shipped loop on an opacity with holes -- test INPUT, written where 3dFD.f90:334-353 writes it: a transparent shaft
    under the beam axis (packets cross every voxel of their column and leave through the bottom face: find() = -1, tflag set
    by tauint1), a shallow crater beside it, a water-depleted rim."""
    it, fr, n = machine(0.03, 0.03, 0.06, rank)
    rk = it.var("iarray", "rhokap").a                      # stored 0-based with the halo: rk[i, j, k] = rhokap(i, j, k)
    rk[39:43, 39:43, 1:n[2] + 1] = 0.0                     # shaft: i, j = 39..42, every k
    rk[43:47, 36:47, n[2] - 5:n[2] + 1] = 0.0              # crater six voxels deep
    rk[47:49, 36:47, 1:n[2] + 1] = 0.5 * 510.0 + 170.0     # rim: w * mu_water + mu_protein (3dFD.f90:343)
    ran2, wall = it.procs["ran2"], it.procs["wall_dist"]
    mc = os.path.join(REF, "mcpolar.f90")
    rows = []
    for _ in range(npackets):
        d0, s0 = ran2.calls, wall.calls
        it.run_block(mc, 153, 169, fr)
        rows.append(packet_row(it, fr, ran2.calls - d0, wall.calls - s0))
    return {"rank": rank, "grid": n, "extents": [0.03, 0.03, 0.06], "packets": rows, "jmean": sparse(it.var("iarray", "jmean").a),
            "iseed_after": fr["iseed"].v, "left_through_the_bottom": sum(1 for r in rows if r[8] == -1)}


def stokes_chain(hgg, nsteps, rank=0):
    """stokes.f90 applied again and again to the direction sourcephCO2 leaves behind."""
    it, fr, n = machine(0.03, 0.03, 0.06, rank)
    it.var("opt_prop", "hgg").set(hgg)
    it.var("opt_prop", "g2").set(hgg * hgg)          # ch_opt.f90:18 (`g2 = hgg**2.`; the product is the same binary64)
    it.call("sourcephco2", [fr[k] for k in ("xmax", "ymax", "zmax", "xcell", "ycell", "zcell", "iseed")])
    pv = lambda k: it.var("photon_vars", k).v
    rows = []
    for _ in range(nsteps):
        it.call("stokes", [fr["iseed"]])
        rows.append([hexf(pv(k)) for k in ("nxp", "nyp", "nzp", "cost", "sint", "cosp", "sinp", "phi")])
    return {"hgg": hgg, "rank": rank, "rows": rows, "iseed_after": fr["iseed"].v}


def dead_code(rank=4):
    """rang (sourceph.f90:73-101) and repeat_bounds (inttau2.f90:242-279): defined upstream, never called there; the
    oracle's Gaussian-beam and periodic-boundary options are built on them."""
    it, fr, n = machine(0.03, 0.03, 0.06, rank)
    rang = [hexf(it.call("rang", [Cell("r", avg), Cell("r", sig), fr["iseed"]], want_result=True))
            for avg, sig in [(0.0, 1.0)] * 150 + [(0.25, 0.004)] * 50]
    delta = fr["delta"].v
    cases = []
    amax, bmax, nag, nbg = 0.03, 0.05, 80, 64
    for cella, cellb, acur, bcur in [(-1, 7, delta / 2, 0.01), (-1, 7, 2 * amax - delta / 2, 0.01), (5, -1, 0.02, delta / 4),
                                     (5, -1, 0.02, 2 * bmax), (-1, -1, -1e-12, 2 * bmax + 1e-12), (3, 4, 0.01, 0.02),
                                     (-1, 2, 0.015, 0.02), (2, -1, 0.015, 0.02)]:
        ca, cb, xa, xb = Cell("i", cella), Cell("i", cellb), Cell("r", acur), Cell("r", bcur)
        try:
            it.call("repeat_bounds", [ca, cb, xa, xb, Cell("r", amax), Cell("r", bmax), Cell("i", nag), Cell("i", nbg), fr["delta"]])
            status = 0
        except Exception as e:                       # 'Error in Repeat_bounds...' ; error stop 0
            assert "ERROR STOP" in str(e), e
            status = -1
        cases.append({"in": [cella, cellb, hexf(acur), hexf(bcur)], "status": status,
                      "out": [ca.v, cb.v, hexf(xa.v), hexf(xb.v)]})
    return {"rank": rank, "rang": rang, "iseed_after": fr["iseed"].v, "repeat_bounds": cases, "delta": hexf(delta),
            "amax": amax, "bmax": bmax, "nag": nag, "nbg": nbg}


def scatter_loop(rank, npackets, extent, mus, mua, hgg):
    """SURVEY 3.3's loop around the reference's routines in a turbid cube (the oracle's `turbid` configuration)."""
    kappa = mus + mua
    it, fr, n = machine(extent, extent, extent, rank, kappa={"mus": mus, "mua": mua, "kappa": kappa, "albedo": mus / kappa,
                                                             "hgg": hgg, "g2": hgg * hgg})
    albedo = it.var("opt_prop", "albedo").v
    ran2, wall, stokes = it.procs["ran2"], it.procs["wall_dist"], it.procs["stokes"]
    args_src = [fr[k] for k in ("xmax", "ymax", "zmax", "xcell", "ycell", "zcell", "iseed")]
    args_tau = [fr[k] for k in ("xmax", "ymax", "zmax", "xcell", "ycell", "zcell", "tflag", "iseed", "delta")]
    rows = []
    for _ in range(npackets):
        d0, s0, c0 = ran2.calls, wall.calls, stokes.calls
        fr["tflag"].set(False)                                             # mcpolar.f90:153
        it.call("sourcephco2", args_src)                                   # :160
        it.call("tauint1", args_tau)                                       # :163
        absorbed = 0
        while not fr["tflag"].v:                                           # :166, body per SURVEY 3.3
            if it.call("ran2", [fr["iseed"]], want_result=True) < albedo:
                it.call("stokes", [fr["iseed"]])
                it.call("tauint1", args_tau)
            else:
                absorbed = 1
                break
        rows.append(packet_row(it, fr, ran2.calls - d0, wall.calls - s0, (stokes.calls - c0, absorbed)))
    return {"rank": rank, "grid": n, "extents": [extent] * 3, "mus": mus, "mua": mua, "hgg": hgg, "albedo": hexf(albedo),
            "packets": rows, "jmean": sparse(it.var("iarray", "jmean").a), "iseed_after": fr["iseed"].v}


def main():
    t0 = time.time()
    out = {"what": "outputs of the reference's own Fortran source executed by oracle/f90interp.py (tests/golden/make_reference_vectors.py)",
           "reference_files": list(FILES) + ["mcpolar.f90:97-98,112-113,153-169"],
           "row": "xp yp zp nxp nyp nzp (binary64 hex) xcell ycell zcell tflag draws voxel_steps [scatterings absorbed]"}
    out["shipped"] = [shipped(0, 1500), shipped(1, 300), shipped(7, 300)]
    print("shipped done", round(time.time() - t0, 1), flush=True)
    out["crater"] = crater(2, 400)
    out["stokes"] = [stokes_chain(0.9, 400), stokes_chain(0.0, 100), stokes_chain(0.5, 200, rank=3)]
    print("stokes done", round(time.time() - t0, 1), flush=True)
    out["dead_code"] = dead_code()
    out["scatter"] = [scatter_loop(0, 150, 0.5, 100.0, 1.0, 0.9), scatter_loop(2, 60, 0.5, 100.0, 1.0, 0.0)]
    print("scatter done", round(time.time() - t0, 1), flush=True)
    import gzip

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_interp.json.gz")
    with gzip.GzipFile(path, "wb", compresslevel=9, mtime=0) as g:         # mtime 0: the same vectors give the same bytes
        g.write(json.dumps(out, separators=(",", ":")).encode())
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
