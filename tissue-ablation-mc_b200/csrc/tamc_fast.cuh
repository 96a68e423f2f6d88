// tamc_fast.cuh -- production arithmetic of the transport (fp64), used by every Philox kernel.
//
// Same physics and the same mapping from draws to paths as tamc_transport.cuh (the statement-by-
// statement restatement that the trace-replay kernel uses), reorganised so the per-voxel-step and
// per-event instruction counts are small:
//   * wall distances multiply by per-event reciprocals of the direction cosines instead of three
//     fp64 divisions per step (inttau2.f90:75-121);
//   * after a wall crossing only the crossed axis is re-indexed (cell +- 1, exit = index leaves
//     1..n) instead of re-deriving all three indices from the position (inttau2.f90:190-239);
//     the snapped coordinate `face +- delta` lies in that neighbour by construction;
//   * the new direction after a scattering is the same rotation as stokes.f90:40-148 written on the
//     direction vector: u' = cos(T) u - sin(T) (cos(i1) e_theta + sin(i1) e_phi), where T is the
//     scattering angle and i1 = TWOPI*xi; this is what the spherical-triangle formulas evaluate
//     (cost' = costp*bmu + sintp*sinbt*cos(i1); phi' = phip -+ acos(cosdph)), without the acos, the
//     second sincos and four of the divisions.  The bmu == +-1 "goto 100" no-op is kept.
// Results differ from the exact path by fp64 rounding only; tests/test_gpu_production.py checks the
// kernels built on this header packet by packet (1e-6 relative) against the oracle run on the same
// Philox stream, and statistically against the oracle on the reference's ran2 stream.
#pragma once

#include "tamc_math.cuh"
#include "tamc_transport.cuh"

namespace tamc {

constexpr double kInvPi = 0.31830988618379067154;   // 1/pi (the true pi: only converts radians for sincospi)

// Philox4x32-10 (the same function as philox4x32_10) with the round keys read from the kernel-parameter bank
// (DevGrid::rk, filled per call by launch_transport): a LOP3 takes them as a direct operand instead of
// re-deriving key + r*W in every block.
__device__ __forceinline__ uint4 philox_block(const DevGrid &g, uint32_t id_lo, uint32_t id_hi, uint32_t blk)
{
    uint4 c = make_uint4(id_lo, id_hi, blk, 0u);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ g.rk[2 * r], lo1, hi0 ^ c.w ^ g.rk[2 * r + 1], lo0);
    }
    return c;
}
__device__ __forceinline__ uint4 philox_block(const DevGrid &g, PhiloxRng &rng) { return philox_block(g, rng.id_lo, rng.id_hi, rng.blk++); }

// u32_to_unit in one fused operation: x*2^-32 + 2^-33 is exact, the same number as (x + 0.5)*2^-32.
__device__ __forceinline__ double unit_fast(uint32_t x) { return __fma_rn((double)x, 1.0 / 4294967296.0, 1.0 / 8589934592.0); }

struct FastPhoton {
    double xcur, ycur, zcur;      // shifted frame, inttau2.f90:24-26
    double nxp, nyp, nzp;         // direction cosines (nzp = cost)
    double inx, iny, inz;         // reciprocals (0 where the cosine is 0)
    double sint, cosp, sinp;      // photon_vars sint, cos(phi), sin(phi)
    double tau, taurun;
    int celli, cellj, cellk;      // 1-based voxel
    int ridx, jidx;               // linear indices into rhokap (halo layout) and jmean
    int dflags;                   // bit0-2: cosine < 0 (x,y,z); bit3-5: cosine == 0
    double rk;                    // kAhead kernels: rhokap of the voxel the packet is in, fetched when it entered
};

__device__ __forceinline__ void set_direction(FastPhoton &p)
{
    int f = 0;
    f |= (p.nxp < 0.) ? 1 : 0;
    f |= (p.nyp < 0.) ? 2 : 0;
    f |= (p.nzp < 0.) ? 4 : 0;
    f |= (p.nxp == 0.) ? 8 : 0;
    f |= (p.nyp == 0.) ? 16 : 0;
    f |= (p.nzp == 0.) ? 32 : 0;
    p.dflags = f;
    if ((f & 56) == 0) {
        // one division for the three reciprocals (the products cannot under/overflow for unit vectors
        // whose smallest component is far above 1e-100)
        const double xy = p.nxp * p.nyp;
        const double r = __drcp_rn(xy * p.nzp);
        p.inz = xy * r;
        const double rz = r * p.nzp;
        p.inx = p.nyp * rz;
        p.iny = p.nxp * rz;
    } else {
        p.inx = (f & 8) ? 0. : 1. / p.nxp;
        p.iny = (f & 16) ? 0. : 1. / p.nyp;
        p.inz = (f & 32) ? 0. : 1. / p.nzp;
    }
}

struct LaunchConsts {
    double zcur0;     // zp0 + zmax
    int cellk0;       // int(nzg*(zp0+zmax)/(2.*zmax))+1, sourceph.f90:47
};

// What a launch produces; small enough to park in shared memory until a lane is free.
struct Launched {
    double xcur, ycur, tau, cosp, sinp;
    int cells;        // celli | cellj << 16 (tamc_init bounds the grid at 4096 per axis)
    int ridx, jidx;   // linear indices of the launch voxel in rhokap / jmean
};

// sourceph.f90:28-31,45-46 in the production arithmetic: the launch point (shifted frame) and its voxel.
__device__ __forceinline__ void launch_point(const DevGrid &g, uint32_t wx, uint32_t wy, double &xcur, double &ycur, int &celli, int &cellj)
{
    const double r = unit_fast(wx) * g.spot_r2;
    const double theta = unit_fast(wy) * kTWOPI;
    double s, c;
    fm::sincospi_0_2(theta * kInvPi, &s, &c);
    const double sr = (r > 1e-280) ? fm::sqrt_normal(r) : sqrt(r);   // (spot diameter 0: r = 0)
    xcur = sr * c + g.xmax;
    ycur = sr * s + g.ymax;
    celli = (int)(xcur * g.inv_dx) + 1;
    cellj = (int)(ycur * g.inv_dy) + 1;
}

// The launch VOXEL alone -- all the column form needs of the launch point (a straight-down flight never leaves its
// column and its deposits do not depend on where in the column it flies).  fp32 first pass with hardware sqrt / sin /
// cos; launch_point() decides whenever the fp32 result lies within eps of a voxel edge, so the voxel is ALWAYS the one
// launch_point() gives.  Error of the fp32 pass in voxel units, R = spot radius in voxels, n = voxels per axis:
//   u (x + 0.5) 2^-32 in fp32: relative 1.2e-7; r = u * spot_r2: 2.4e-7; sqrt.approx: 1.2e-7 on top of half of that
//     -> sr relative 2.4e-7;
//   angle a = u * 6.283185f: relative 2.4e-7 -> <= 1.5e-6 rad; minus 2 pi (fp32) above pi: 2e-6 rad in all;
//   sin.approx / cos.approx on [-pi, pi]: absolute 2^-20 = 9.5e-7 (PTX ISA: 2^-20.5 in the primary range)
//     -> sr * cos: absolute <= R * 3.3e-6; times inv_dx in fp32: + R * 6e-8; the fma rounds X <= n to 6e-8 * n;
//   X - floor(X) and the comparison with 0.5: exact / 3e-8.
// Bound: R * 3.5e-6 + n * 6e-8 + 1e-7; DevGrid::half_* holds 0.5 - 4 x that bound (tamc_api.cu).  For homog200
// (R = 41.7 voxels, n = 200) eps = 6.3e-4: one packet in 400 is redone in fp64.  tamc_selfcheck_launch() runs both
// passes over billions of draws and counts disagreements (tests/test_gpu_column.py: zero).
__device__ __forceinline__ bool launch_voxel_fp32(const DevGrid &g, uint32_t wx, uint32_t wy, int &celli, int &cellj)
{
    const float u0 = fmaf((float)wx, 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (x + 0.5) 2^-32
    const float u1 = fmaf((float)wy, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    float sr, sn, cs;
    const float r = u0 * g.spot_r2_f;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sr) : "f"(r));
    float a = u1 * 6.283185f;                               // the reference's truncated TWOPI (constants.f90:13)
    a = a > 3.14159265f ? a - 6.28318531f : a;              // the same angle, in [-pi, pi]
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(a));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(a));
    const float X = fmaf(sr * cs, g.inv_dx_f, g.x0_f), Y = fmaf(sr * sn, g.inv_dy_f, g.y0_f);
    const float Xi = floorf(X), Yi = floorf(Y);
    celli = (int)Xi + 1;
    cellj = (int)Yi + 1;
    return fabsf((X - Xi) - 0.5f) < g.half_x && fabsf((Y - Yi) - 0.5f) < g.half_y;
}

struct LaunchedColumn {
    double tau;
    int cells;        // celli | cellj << 16
    int jidx;         // linear index of the launch voxel in jmean
};

// Launch of a stub-regime packet in the column form: voxel + optical depth (sourceph.f90:28-47, inttau2.f90:36).
__device__ __forceinline__ LaunchedColumn launch_column(const DevGrid &g, const uint4 w)
{
    LaunchedColumn L;
    int celli, cellj;
    if (!launch_voxel_fp32(g, w.x, w.y, celli, cellj)) {
        double xcur, ycur;
        launch_point(g, w.x, w.y, xcur, ycur, celli, cellj);
    }
    L.cells = celli | (cellj << 16);
    L.jidx = (celli - 1) + g.nxg * ((cellj - 1) + g.nyg * (g.cellk0 - 1));
    L.tau = fm::neglog_u32(w.w);
    return L;
}

// sourceph.f90:28-47 + inttau2.f90:36.  w = one Philox block = the (r, theta, phi, tau) draws in the reference's order.
__device__ __forceinline__ Launched launch_fast(const DevGrid &g, const uint4 w, bool need_azimuth)
{
    Launched L;
    int celli, cellj;
    launch_point(g, w.x, w.y, L.xcur, L.ycur, celli, cellj);
    L.cells = celli | (cellj << 16);
    L.ridx = celli + g.sx * (cellj + (g.nyg + 2) * g.cellk0);
    L.jidx = (celli - 1) + g.nxg * ((cellj - 1) + g.nyg * (g.cellk0 - 1));
    L.cosp = 1.;
    L.sinp = 0.;
    if (need_azimuth) fm::sincospi_0_2(kTWOPI * unit_fast(w.z) * kInvPi, &L.sinp, &L.cosp);   // phi is first used by the first scattering
    L.tau = fm::neglog_u32(w.w);
    return L;
}

// Gaussian beam (tamc_set_source_gaussian; oracle: sourcephGauss): rang()'s polar method on the source stream
// (counter word 3 = 2), each variate redrawn while it misses the top face; phi and tau are words 2 and 3 of the
// packet's block 0 `w`, exactly where the disk source takes them.  Not a hot path: plain fp64 log / sqrt.
__device__ __forceinline__ double rang_fast(const DevGrid &g, uint32_t id_lo, uint32_t id_hi, int &ns, double sigma)
{
    const uint2 key = make_uint2(g.rk[0], g.rk[1]);
    double u, s;
    do {
        u = -1. + source_draw(key, id_lo, id_hi, ns) * 2.;
        s = -1. + source_draw(key, id_lo, id_hi, ns) * 2.;
        s = s * s + u * u;
    } while (s >= 1.);
    return sigma * (u * sqrt(-2. * log(s) / s));
}

__device__ __forceinline__ Launched launch_fast_gauss(const DevGrid &g, uint32_t id_lo, uint32_t id_hi, const uint4 w, bool need_azimuth)
{
    Launched L;
    int ns = 0;
    double xp, yp;
    do { xp = rang_fast(g, id_lo, id_hi, ns, g.gauss_sigma); } while (!(fabs(xp) < g.xmax));
    do { yp = rang_fast(g, id_lo, id_hi, ns, g.gauss_sigma); } while (!(fabs(yp) < g.ymax));
    L.xcur = xp + g.xmax;
    L.ycur = yp + g.ymax;
    const int celli = min(g.nxg, (int)(L.xcur * g.inv_dx) + 1);
    const int cellj = min(g.nyg, (int)(L.ycur * g.inv_dy) + 1);
    L.cells = celli | (cellj << 16);
    L.ridx = celli + g.sx * (cellj + (g.nyg + 2) * g.cellk0);
    L.jidx = (celli - 1) + g.nxg * ((cellj - 1) + g.nyg * (g.cellk0 - 1));
    L.cosp = 1.;
    L.sinp = 0.;
    if (need_azimuth) fm::sincospi_0_2(kTWOPI * unit_fast(w.z) * kInvPi, &L.sinp, &L.cosp);
    L.tau = fm::neglog_u32(w.w);
    return L;
}

// Launch of packet (id_lo, id_hi) from whichever source the handle has; consumes block 0 of the packet's main stream.
__device__ __forceinline__ Launched launch_any(const DevGrid &g, uint32_t id_lo, uint32_t id_hi, bool need_azimuth)
{
    const uint4 w = philox_block(g, id_lo, id_hi, 0u);
    if (g.gauss_sigma > 0.) return launch_fast_gauss(g, id_lo, id_hi, w, need_azimuth);
    return launch_fast(g, w, need_azimuth);
}

__device__ __forceinline__ void adopt(const DevGrid &g, const LaunchConsts &lc, FastPhoton &p, const Launched &L)
{
    p.xcur = L.xcur; p.ycur = L.ycur; p.zcur = lc.zcur0;
    p.nxp = 0.; p.nyp = 0.; p.nzp = -1.;          // sint = 0, cost = -1 (sourceph.f90:37-42)
    p.inx = 0.; p.iny = 0.; p.inz = -1.;
    p.dflags = 4 | 8 | 16;
    p.sint = 0.; p.cosp = L.cosp; p.sinp = L.sinp;
    p.tau = L.tau; p.taurun = 0.;
    p.celli = L.cells & 0xffff; p.cellj = L.cells >> 16; p.cellk = lc.cellk0;
    p.ridx = L.ridx;
    p.jidx = L.jidx;
}

__device__ __forceinline__ double flip_sign(double v, int neg)
{
    return __hiloint2double(__double2hiint(v) ^ (neg << 31), __double2loint(v));
}

// One voxel-step, inttau2.f90:37-63, written without divergent branches: the wall-crossing step
// (inttau2.f90:42-48) and the final partial step (:50-55) share one instruction stream and differ
// only through selects, so a warp whose lanes cross different faces -- or end their flight -- in the
// same iteration still issues the step once.
// kNeedPos = false (shipped stub regime: the packet ends at its first interaction) drops the final
// position update and its division: the last deposit dcell*rhokap = ((tau-taurun)/rhokap)*rhokap is
// tallied as tau-taurun.
// kAhead: the opacity is read from p.rk, loaded as soon as the packet's voxel was known (on adoption / at the end of
// the previous step), so the load's latency runs beside the loop's bookkeeping instead of at the head of the step --
// what matters when the grids exceed L2 (400^3: ncu long-scoreboard 3.5 per issue); a scattering that leaves the
// packet in its voxel re-uses the value.
template <bool kNeedPos, class Tally, bool kAhead = false>
__device__ __forceinline__ int voxel_step_fast(const DevGrid &g, const double *xf, const double *yf, const double *zf,
                                               FastPhoton &p, Tally &tally)
{
    const int f = p.dflags;
    const int negx = f & 1, negy = (f >> 1) & 1, negz = (f >> 2) & 1;
    const double fx = xf[p.celli - negx], fy = yf[p.cellj - negy], fz = zf[p.cellk - negz];
    const double dx = (f & 8) ? 100000. : (fx - p.xcur) * p.inx;
    const double dy = (f & 16) ? 100000. : (fy - p.ycur) * p.iny;
    const double dz = (f & 32) ? 100000. : (fz - p.zcur) * p.inz;
    const double dwall = fmin(fmin(dx, dy), dz);
    const double rk = kAhead ? p.rk : __ldg(g.rhokap + p.ridx);
    const double taucell = dwall * rk;
    const bool wall = p.taurun + taucell < p.tau;                    // inttau2.f90:42
    const double rest = p.tau - p.taurun;

    if (!kNeedPos) {
        tally.add(p, wall ? taucell : rest);
        if (!wall) return STEP_INTERACT;
    }
    double dcell = dwall;
    if (kNeedPos) {
        const double dpart = rest * __drcp_rn(rk);                    // inttau2.f90:51 (unused when rk == 0: wall is true)
        dcell = wall ? dwall : dpart;
        tally.add(p, wall ? taucell : dpart * rk);
    }
    p.taurun += taucell;
    // which face: the later axis wins ties (inttau2.f90:116-118); none on the final partial step
    const bool hz = wall & (dwall == dz);
    const bool hy = wall & !hz & (dwall == dy);
    const bool hx = wall & !hz & !hy;
    const double xn = p.xcur + p.nxp * dcell, yn = p.ycur + p.nyp * dcell, zn = p.zcur + p.nzp * dcell;
    // face -+ delta (inttau2.f90:140-170): the sign of delta follows the direction, flipped in the sign bit
    const double sxd = flip_sign(g.delta, negx), syd = flip_sign(g.delta, negy), szd = flip_sign(g.delta, negz);
    const double xs = fx + sxd, ys = fy + syd, zs = fz + szd;
    p.xcur = hx ? xs : xn;
    p.ycur = hy ? ys : yn;
    p.zcur = hz ? zs : zn;
    const int sx = hx ? 1 - 2 * negx : 0, sy = hy ? 1 - 2 * negy : 0, sz = hz ? 1 - 2 * negz : 0;
    p.celli += sx;
    p.cellj += sy;
    p.cellk += sz;
    p.ridx += sx + sy * g.sx + sz * (int)g.sxy;
    p.jidx += sx + sy * g.nxg + sz * (g.nxg * g.nyg);
    if (kAhead && wall) p.rk = __ldg(g.rhokap + p.ridx);          // the halo keeps this in bounds when the packet has left
    if (!wall) return STEP_INTERACT;
    const bool out = ((unsigned)(p.celli - 1) >= (unsigned)g.nxg) | ((unsigned)(p.cellj - 1) >= (unsigned)g.nyg) |
                     ((unsigned)(p.cellk - 1) >= (unsigned)g.nzg);
    return out ? STEP_EXIT : STEP_WALL;
}

// stokes.f90:6-153 as a rotation of the direction vector (see the header comment).
// u1 -> stokes.f90:24/:48, u2 -> :32/:64, tau = -log(u3) -> the next tauint1 draw (inttau2.f90:36).
template <bool kSetDir = true>
__device__ __forceinline__ void scatter_dir(const DevGrid &g, FastPhoton &p, double u1, double u2);
template <bool kSetDir = true>
__device__ __forceinline__ void scatter_dir(FastPhoton &p, double u1, double u2, double hgg, const ScatterConsts &sc);

// The constants of the Henyey-Greenstein draw for one value of g (per-voxel hgg: formed per event; otherwise DevGrid::sc)
__device__ __forceinline__ ScatterConsts make_scatter_consts(double hgg)
{
    ScatterConsts sc;
    const double g2 = hgg * hgg;
    sc.one_m_g2 = 1. - g2;
    sc.one_p_g2 = 1. + g2;
    sc.one_m_g = 1. - hgg;
    sc.two_g = 2. * hgg;
    sc.inv_two_g = (hgg != 0.) ? 1. / (2. * hgg) : 0.;
    return sc;
}

// albedo and Henyey-Greenstein constants of an interaction in the voxel with halo-layout index v
__device__ __forceinline__ void voxel_optics(const DevGrid &g, long long v, double &albedo, double &hgg, ScatterConsts &sc)
{
    albedo = g.albedo_g ? __ldg(g.albedo_g + v) : g.albedo;
    hgg = g.hgg;
    sc = g.sc;
    if (g.hgg_g) {
        hgg = __ldg(g.hgg_g + v);
        sc = make_scatter_consts(hgg);
    }
}

__device__ __forceinline__ void scatter_fast(const DevGrid &g, FastPhoton &p, double u1, double u2, double tau, double hgg,
                                             const ScatterConsts &sc)
{
    p.taurun = 0.;
    p.tau = tau;
    p.xcur = (p.xcur - g.xmax) + g.xmax;
    p.ycur = (p.ycur - g.ymax) + g.ymax;
    p.zcur = (p.zcur - g.zmax) + g.zmax;
    scatter_dir<true>(p, u1, u2, hgg, sc);
}

__device__ __forceinline__ void scatter_fast(const DevGrid &g, FastPhoton &p, double u1, double u2, double tau)
{
    p.taurun = 0.;
    p.tau = tau;
    // the centred position round trip of inttau2.f90:65-67 / :24-26
    p.xcur = (p.xcur - g.xmax) + g.xmax;
    p.ycur = (p.ycur - g.ymax) + g.ymax;
    p.zcur = (p.zcur - g.zmax) + g.zmax;
    scatter_dir<true>(g, p, u1, u2);
}

// The direction part of a scattering event: reads and writes nxp,nyp,nzp, sint,cosp,sinp, the
// reciprocals and dflags of `p` (the last two only with kSetDir); position and optical depths are untouched.
template <bool kSetDir>
__device__ __forceinline__ void scatter_dir(const DevGrid &g, FastPhoton &p, double u1, double u2)
{
    scatter_dir<kSetDir>(p, u1, u2, g.hgg, g.sc);
}

template <bool kSetDir>
__device__ __forceinline__ void scatter_dir(FastPhoton &p, double u1, double u2, double hgg, const ScatterConsts &sc)
{
    if (hgg == 0.0) {                                     // isotropic, stokes.f90:23-38
        const double cost = 2. * u1 - 1.;
        const double s2 = 1. - cost * cost;
        p.sint = (s2 <= 0.) ? 0. : sqrt(s2);
        fm::sincospi_0_2(kTWOPI * u2 * kInvPi, &p.sinp, &p.cosp);
        p.nxp = p.sint * p.cosp;
        p.nyp = p.sint * p.sinp;
        p.nzp = cost;
        if (kSetDir) set_direction(p);
        return;
    }
    const double q = sc.one_m_g2 * __drcp_rn(sc.one_m_g + sc.two_g * u1);   // stokes.f90:48
    double bmu = (sc.one_p_g2 - q * q) * sc.inv_two_g;
    bmu = fmin(1., fmax(-1., bmu));
    if (bmu == 1. || bmu == -1.) return;                               // goto 100, stokes.f90:71-77
    const double sinbt = sqrt(1. - bmu * bmu);
    // i1 = TWOPI*xi; beyond PI the reference works with i3 = TWOPI - i1 and adds instead of subtracts
    // (stokes.f90:66-68,101).  With the truncated constants i3 is not exactly -i1 (mod 2 pi), so the
    // same i3 is formed here and only the sign of its sine is flipped.
    const double ri1 = kTWOPI * u2;
    const bool upper = ri1 > kPI;
    double si, ci;
    fm::sincospi_0_2((upper ? kTWOPI - ri1 : ri1) * kInvPi, &si, &ci);   // sin/cos of the angle, argument reduced exactly
    si = upper ? -si : si;

    const double costp = p.nzp, sintp = p.sint;
    const double a = sinbt * ci * costp, b = sinbt * si;
    double uz = costp * bmu + sintp * sinbt * ci;                      // stokes.f90:79 / :117
    const double ux = bmu * p.nxp - (a * p.cosp - b * p.sinp);
    const double uy = bmu * p.nyp - (a * p.sinp + b * p.cosp);
    uz = fmin(1., fmax(-1., uz));
    const double h2 = ux * ux + uy * uy;
    if (h2 > 0.) {
        const double ih = rsqrt(h2);
        p.cosp = ux * ih;
        p.sinp = uy * ih;
    }
    p.sint = sqrt(1. - uz * uz);                                        // stokes.f90:81 / :119
    p.nxp = p.sint * p.cosp;                                            // stokes.f90:143-148
    p.nyp = p.sint * p.sinp;
    p.nzp = uz;
    if (kSetDir) set_direction(p);
}

// TAMC_FRESNEL on the production arithmetic.  After STEP_EXIT exactly one index is out of range (only the
// crossed axis is re-indexed).  Returns true when the packet was reflected back into the grid.
__device__ __forceinline__ bool fresnel_reflect_fast(const DevGrid &g, const double *xf, const double *yf, const double *zf,
                                                     FastPhoton &p, uint2 key, uint32_t id_lo, uint32_t id_hi, int &nb)
{
    const int a = ((unsigned)(p.celli - 1) >= (unsigned)g.nxg) ? 0 : (((unsigned)(p.cellj - 1) >= (unsigned)g.nyg) ? 1 : 2);
    const double na = a == 0 ? p.nxp : (a == 1 ? p.nyp : p.nzp);
    const int back = (na > 0.) ? -1 : 1;             // undo the index step of the crossing
    double n_in = g.n2;
    if (g.n_g) n_in = __ldg(g.n_g + (p.ridx + back * (a == 0 ? 1 : (a == 1 ? g.sx : (int)g.sxy))));   // the voxel being left
    if (!(boundary_draw(key, id_lo, id_hi, nb) < fresnel_reflectance(n_in, g.n1, fabs(na)))) return false;
    if (a == 0) {
        p.xcur = (na > 0.) ? xf[g.nxg] - g.delta : xf[0] + g.delta;
        p.celli += back; p.ridx += back; p.jidx += back;
        p.nxp = -p.nxp; p.inx = -p.inx; p.cosp = -p.cosp; p.dflags ^= 1;
    } else if (a == 1) {
        p.ycur = (na > 0.) ? yf[g.nyg] - g.delta : yf[0] + g.delta;
        p.cellj += back; p.ridx += back * g.sx; p.jidx += back * g.nxg;
        p.nyp = -p.nyp; p.iny = -p.iny; p.sinp = -p.sinp; p.dflags ^= 2;
    } else {
        p.zcur = (na > 0.) ? zf[g.nzg] - g.delta : zf[0] + g.delta;
        p.cellk += back; p.ridx += back * (int)g.sxy; p.jidx += back * (g.nxg * g.nyg);
        p.nzp = -p.nzp; p.inz = -p.inz; p.dflags ^= 4;
    }
    return true;
}

// TAMC_PERIODIC on the production arithmetic: repeat_bounds (inttau2.f90:242-279).  After STEP_EXIT exactly one
// index is out of range; a lateral one re-enters on the opposite side at `delta` / `2*max - delta`.
__device__ __forceinline__ bool periodic_wrap_fast(const DevGrid &g, FastPhoton &p)
{
    if ((unsigned)(p.celli - 1) >= (unsigned)g.nxg) {
        const bool low = p.celli < 1;                     // left through -x: acur < delta
        p.xcur = low ? 2. * g.xmax - g.delta : g.delta;
        const int d = low ? g.nxg : -g.nxg;
        p.celli += d; p.ridx += d; p.jidx += d;
        return true;
    }
    if ((unsigned)(p.cellj - 1) >= (unsigned)g.nyg) {
        const bool low = p.cellj < 1;
        p.ycur = low ? 2. * g.ymax - g.delta : g.delta;
        const int d = low ? g.nyg : -g.nyg;
        p.cellj += d; p.ridx += d * g.sx; p.jidx += d * g.nxg;
        return true;
    }
    return false;
}

// What happens to a packet that a wall crossing took out of the grid, when boundary options are on:
// 0 = it leaves, 1 = Fresnel-reflected back in, 2 = re-entered through the opposite lateral face.
__device__ __forceinline__ int boundary_fast(const DevGrid &g, const double *xf, const double *yf, const double *zf,
                                             FastPhoton &p, uint2 key, uint32_t id_lo, uint32_t id_hi, int &nb)
{
    if ((g.flags & TAMC_PERIODIC) && periodic_wrap_fast(g, p)) return 2;
    if ((g.flags & TAMC_FRESNEL) && fresnel_reflect_fast(g, xf, yf, zf, p, key, id_lo, id_hi, nb)) return 1;
    return 0;
}

// TAMC_FRESNEL: probability of the specular reflection at launch, ((n1-n2)/(n1+n2))^2 with the launch voxel's index
__device__ __forceinline__ double specular_r0sq(const DevGrid &g, int ridx)
{
    if (!g.n_g) return g.r0sq;
    const double n_in = __ldg(g.n_g + ridx);
    const double r = (g.n1 - n_in) / (g.n1 + n_in);
    return r * r;
}

__device__ __forceinline__ int exit_face_fast(const FastPhoton &p, const DevGrid &g)
{
    if (p.celli < 1 || p.celli > g.nxg) return p.nxp > 0. ? 2 : 1;
    if (p.cellj < 1 || p.cellj > g.nyg) return p.nyp > 0. ? 4 : 3;
    if (p.cellk < 1 || p.cellk > g.nzg) return p.nzp > 0. ? 6 : 5;
    return 0;
}

// Counters kept in shared memory, one set per warp, updated when a packet ends.  Used by the
// scattering kernel, where packets end rarely (once per tens of events) and registers are scarce.
struct WarpCounters {
    unsigned long long *w;      // CNT_N slots of this warp
    __device__ __forceinline__ void clear()
    {
        for (int i = threadIdx.x & 31; i < CNT_N; i += 32) w[i] = 0ull;
        __syncwarp();
    }
    __device__ __forceinline__ void death(int f, int nsteps, int nscatt, bool err)
    {
        atomicAdd(w + CNT_PACKETS, 1ull);
        atomicAdd(w + CNT_STEPS, (unsigned long long)nsteps);
        atomicAdd(w + CNT_SCATTERS, (unsigned long long)nscatt);
        atomicAdd(w + (f == 0 ? CNT_ABSORBED : CNT_EXIT0 + f - 1), 1ull);
        if (err) atomicAdd(w + CNT_ERRORS, 1ull);
    }
    __device__ __forceinline__ void note(int slot) { atomicAdd(w + slot, 1ull); }
    __device__ __forceinline__ void commit(unsigned long long *g) const
    {
        __syncwarp();
        const int i = threadIdx.x & 31;
        if (i < CNT_N && i != CNT_WORK && w[i]) atomicAdd(g + i, w[i]);
    }
};

// Tally policies on 32-bit voxel indices (tamc_init bounds the grid so they fit).
struct DirectTally32 {
    double *jm;
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void add(const FastPhoton &p, double v)
    {
        if (v != 0.) atomicAdd(jm + p.jidx, v);           // RED.E.ADD.F64
    }
    __device__ __forceinline__ void flush() {}
};

struct MergeTally32 {
    double *jm;
    int pidx;
    double pval;
    __device__ __forceinline__ void begin() { pidx = -1; pval = 0.; }
    __device__ __forceinline__ void add(const FastPhoton &p, double v)
    {
        const int idx = p.jidx;
        if (idx == pidx) { pval += v; return; }
        if (pval != 0.) atomicAdd(jm + pidx, pval);
        pidx = idx;
        pval = v;
    }
    __device__ __forceinline__ void flush()
    {
        if (pval != 0.) atomicAdd(jm + pidx, pval);
        pidx = -1;
        pval = 0.;
    }
};

// Stub regime (straight-down flights): the top `layers` planes of the tally under the beam's bounding box
// are privatised in shared memory, one tile per CTA, and flushed with one RED per non-zero entry at the end
// of the kernel.  Two thirds of all deposits land there, so the global fp64 REDs -- which share the L1TEX
// address throughput with the rhokap loads, the limit of this regime -- drop by that much.  fp64 atomicAdd
// on shared memory is a compare-and-swap loop (ATOMS.CAST.SPIN): a few instructions, rarely contended
// (tens of thousands of tile entries per CTA).
struct TileGeom {
    int i0, j0;        // first voxel (1-based) of the tile in x and y
    int tw, th;        // tile extent in x and y
    int layers;        // planes from the top face: k = nzg, nzg-1, ...
};

struct TiledTally32 {
    double *jm;
    double *tile;
    TileGeom t;
    int nzg;
    __device__ __forceinline__ void begin() {}
    __device__ __forceinline__ void add(const FastPhoton &p, double v)
    {
        const unsigned di = (unsigned)(p.celli - t.i0), dj = (unsigned)(p.cellj - t.j0), dk = (unsigned)(nzg - p.cellk);
        if (v == 0.) return;
        if (di < (unsigned)t.tw && dj < (unsigned)t.th && dk < (unsigned)t.layers) atomicAdd(tile + (di + t.tw * (dj + t.th * dk)), v);
        else atomicAdd(jm + p.jidx, v);
    }
    __device__ __forceinline__ void flush() {}
};

}  // namespace tamc
