// tamc_transport.cuh -- device-side photon transport for sm_100a.
//
// One packet = one GPU thread.  The arithmetic is the reference's, statement for statement, in
// fp64 (the reference is fp64 throughout: src/Makefile:3 -freal-4-real-8); what is new is the
// execution model: counter-based Philox instead of the sequential ran2, face tables in shared
// memory, position->voxel by multiply+fix-up instead of three bisections, fp64 RED atomics into
// the per-GPU tally.  Included by two translation units: tamc_kernels.cu (production, FMA
// contraction allowed) and tamc_replay.cu (trace replay, compiled -fmad=false).
//
// Reference lines restated here (all under /root/reference/src):
//   sourceph.f90:28-47  launch()          inttau2.f90:75-121  wall distance in voxel_step()
//   inttau2.f90:36-63   voxel_step()      inttau2.f90:124-187 position update in voxel_step()
//   inttau2.f90:208-239 find_cell()       stokes.f90:6-153    stokes()
//   mcpolar.f90:151-170 transport_packet() / the persistent kernel in tamc_kernels.cu
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/tamc.h"

namespace tamc {

// constants.f90:13 -- 7-digit truncations held in doubles; replay parity needs exactly these.
constexpr double kPI = 3.141592;
constexpr double kTWOPI = 6.283185;

// A packet that takes more voxel-steps than this is killed and counted as an error: guards the GPU
// against a caller-supplied delta below the ulp of a face coordinate (the reference would spin).
constexpr int kMaxStepsPerPacket = 1 << 26;

enum { CNT_PACKETS = 0, CNT_STEPS, CNT_SCATTERS, CNT_ABSORBED, CNT_EXIT0, CNT_ERRORS = 10, CNT_OVERFLOW = 11, CNT_WORK = 12,
       CNT_SPECULAR = 13, CNT_REFLECT = 14,
       CNT_DEPTH = 15,   // column form: planes from the top face down to the deepest stop of the call (k_column_finish)
       CNT_N = 16 };

// Constants of the Henyey-Greenstein draw (stokes.f90:48), formed once on the host.
struct ScatterConsts {
    double one_m_g2, one_p_g2, one_m_g, two_g, inv_two_g;
};

struct DevGrid {
    int nxg, nyg, nzg;
    int sx;               // rhokap stride in j: nxg+2
    long long sxy;        // rhokap stride in k: (nxg+2)*(nyg+2)
    double xmax, ymax, zmax, delta;
    double spot_r2;       // (spotSize/2.)**2, sourceph.f90:28
    double zp0;           // zmax-(1.e-8*(2.*zmax/nzg)), sourceph.f90:32
    double inv_dx, inv_dy, inv_dz;  // nxg/(2 xmax) ...: first guess of the voxel index only
    double albedo, hgg, g2;
    double zcur0;         // zp0 + zmax: every packet starts at this height (inttau2.f90:26)
    int cellk0;           // int(nzg*(zp0+zmax)/(2.*zmax))+1, sourceph.f90:47
    int flags;
    double n1, n2, r0sq;  // TAMC_FRESNEL: indices outside / inside the grid, ((n1-n2)/(n1+n2))**2
    double gauss_sigma;   // > 0: Gaussian beam through rang() (sourceph.f90:73-101) instead of the CO2 disk
    // fp32 first pass of the launch voxel in the column form (launch_cells, tamc_fast.cuh): spot_r2, inv_dx, inv_dy in
    // fp32, the centre of the face in voxel units (nxg/2, nyg/2: exact), and 0.5 - eps per axis, eps = the proven bound
    // of the fp32 pass in voxel units (x 4); a launch point closer than eps to a voxel edge is redone in fp64.
    // half < 0 switches the fp32 pass off.
    float spot_r2_f, inv_dx_f, inv_dy_f, x0_f, y0_f, half_x, half_y;
    ScatterConsts sc;
    // flight kernel (tamc_flight.cuh): voxel edges 2*max/n, the same minus delta, and the distance from the launch height
    // down to the bottom face of the launch voxel -- formed once on the host so the kernel reads them from the parameter bank
    double fwx, fwy, fwz, wx, wy, wz, ez0;
    const double *rhokap; // (0:nxg+1,0:nyg+1,0:nzg+1) column-major, as uploaded
    double *jmean;        // (nxg,nyg,nzg) column-major
    const double *faces;  // xface(1:nxg+1) | yface(1:nyg+1) | zface(1:nzg+1)
    // EXTENSION (tamc_set_optics_grids; upstream's opt_prop.f90:5 holds scalars): optional per-voxel albedo / hgg /
    // refractive index in rhokap's layout, read-only through L2 next to rhokap; null = the scalar above
    const double *albedo_g, *hgg_g, *n_g;
    uint32_t rk[20];      // Philox round keys of this call's seed: key + r*(0x9E3779B9, 0xBB67AE85), r = 0..9 (production kernels)
};

// Bounding box of the beam's footprint on the top face (shipped regime: every flight stays inside these columns).
struct ColGeom {
    int i0, j0;        // first voxel (1-based) of the box in x and y
    int tw, th;        // its extent
    int nzp;           // column length of the z-fastest opacity copy: nzg rounded up to a multiple of 4
    // Columns-first upload over PCIe with a depth limit (tamc_run_optics, LaunchCfg::gather_depth): only planes kz = k-1
    // >= kz_lo (a multiple of 32) were copied into the z-fastest / box copies; the rare packet that goes deeper reads
    // `deep` (reference layout, strides deep_sx / deep_sxy): the caller's page-locked grid for k_column_bound, which copies
    // the planes such a packet can reach into the resident grid, and that resident grid for the kernels after it.
    // kz_lo = 0: everything copied.
    int kz_lo;
    int deep_sx;
    long long deep_sxy;
    const double *deep;
};

// ---------------------------------------------------------------------------------------------
// RNG: Philox4x32-10, key = run seed, counter = (packet id lo, hi, draw block, 0).
// One block = 4 uniforms = exactly what one event consumes (launch: r, theta, phi, tau;
// interaction: albedo test, HG cosine, azimuth, tau).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint2 key, uint4 c)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ key.x, lo1, hi0 ^ c.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return c;
}

// (x + 0.5) * 2^-32: exact in fp64, strictly inside (0,1) like ran2's output (ran2.f:31).
__device__ __forceinline__ double u32_to_unit(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }

__device__ __forceinline__ double source_draw(uint2 key, uint32_t id_lo, uint32_t id_hi, int &ns);

struct PhiloxRng {
    static constexpr bool kHasBoundary = true;
    uint2 key;
    uint32_t id_lo, id_hi, blk;
    __device__ __forceinline__ void seed(uint64_t s, uint64_t packet)
    {
        key = make_uint2((uint32_t)s, (uint32_t)(s >> 32));
        id_lo = (uint32_t)packet;
        id_hi = (uint32_t)(packet >> 32);
        blk = 0;
    }
    __device__ __forceinline__ void block(double u[4])
    {
        const uint4 r = philox4x32_10(key, make_uint4(id_lo, id_hi, blk++, 0u));
        u[0] = u32_to_unit(r.x); u[1] = u32_to_unit(r.y); u[2] = u32_to_unit(r.z); u[3] = u32_to_unit(r.w);
    }
    // Gaussian source: rang()'s draws come from the source stream; phi and tau stay words 2 and 3 of block 0
    static constexpr bool kSequential = false;
    __device__ __forceinline__ double src(int &ns) { return source_draw(key, id_lo, id_hi, ns); }
    __device__ __forceinline__ bool exhausted() const { return false; }
    __device__ __forceinline__ void launch_tail(double &uphi, double &utau)
    {
        double u[4];
        block(u);
        uphi = u[2];
        utau = u[3];
    }
};

// Boundary decisions (TAMC_FRESNEL) draw from a stream of their own: counter word 3 = 1, draw nb of the packet.
__device__ __forceinline__ double boundary_draw(uint2 key, uint32_t id_lo, uint32_t id_hi, int &nb)
{
    const uint4 r = philox4x32_10(key, make_uint4(id_lo, id_hi, (uint32_t)(nb >> 2), 1u));
    const int l = nb & 3;
    ++nb;
    return u32_to_unit(l == 0 ? r.x : (l == 1 ? r.y : (l == 2 ? r.z : r.w)));
}

// The Gaussian source (tamc_set_source_gaussian) draws from a stream of its own as well: counter word 3 = 2.  The polar
// method consumes a variable number of draws; the packet's main stream keeps its layout of one block per event.
__device__ __forceinline__ double source_draw(uint2 key, uint32_t id_lo, uint32_t id_hi, int &ns)
{
    const uint4 r = philox4x32_10(key, make_uint4(id_lo, id_hi, (uint32_t)(ns >> 2), 2u));
    const int l = ns & 3;
    ++ns;
    return u32_to_unit(l == 0 ? r.x : (l == 1 ? r.y : (l == 2 ? r.z : r.w)));
}

// Unpolarised Fresnel reflectance going from index n_in to n_out at cosine of incidence ci (1 beyond the
// critical angle).  Builder-defined extension: the reference has no boundary optics.
__device__ __forceinline__ double fresnel_reflectance(double n_in, double n_out, double ci)
{
    const double ratio = n_in / n_out;
    const double si2 = (ratio * ratio) * (1. - ci * ci);
    if (si2 >= 1.) return 1.;
    const double ct = sqrt(1. - si2);
    const double rs = (n_in * ci - n_out * ct) / (n_in * ci + n_out * ct);
    const double rp = (n_in * ct - n_out * ci) / (n_in * ct + n_out * ci);
    return 0.5 * (rs * rs + rp * rp);
}

// Replay: the packet's slice of the reference ran2 sequence, consumed in the reference's order.
struct ReplayRng {
    static constexpr bool kHasBoundary = false;   // the reference has no boundary optics to replay
    const double *p;
    long long pos, end;
    __device__ __forceinline__ void block(double u[4])
    {
#pragma unroll
        for (int i = 0; i < 4; ++i) u[i] = (pos + i < end) ? p[pos + i] : 0.5;
        pos += 4;
    }
    // Gaussian source: the reference's single sequential stream -- rang()'s draws, then phi, then tau
    static constexpr bool kSequential = true;
    __device__ __forceinline__ double src(int &ns)
    {
        ++ns;
        const double v = (pos < end) ? p[pos] : 0.75;
        ++pos;
        return v;
    }
    __device__ __forceinline__ bool exhausted() const { return pos >= end; }
    __device__ __forceinline__ void launch_tail(double &uphi, double &utau)
    {
        int n = 0;
        uphi = src(n);
        utau = src(n);
    }
};

// ---------------------------------------------------------------------------------------------
// Tally policies.  fp64 atomicAdd with the result unused compiles to RED.E.ADD.F64 (fire and
// forget, resolved in L2).  A zero deposit (ablated voxel, rhokap == 0) is skipped: jmean += 0.
// ---------------------------------------------------------------------------------------------
struct DirectTally {
    double *jm;
    double packet_sum;
    __device__ __forceinline__ void begin() { packet_sum = 0.; }
    __device__ __forceinline__ void add(long long idx, double v)
    {
        packet_sum += v;
        if (v != 0.) atomicAdd(jm + idx, v);
    }
    __device__ __forceinline__ void flush() {}
};

// Consecutive deposits into the same voxel (the partial step that ends at a scattering site and the
// first step after it) are merged in registers and written once.
struct MergeTally {
    double *jm;
    double packet_sum;
    long long pidx;
    double pval;
    __device__ __forceinline__ void begin() { packet_sum = 0.; pidx = -1; pval = 0.; }
    __device__ __forceinline__ void add(long long idx, double v)
    {
        packet_sum += v;
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

// ---------------------------------------------------------------------------------------------
// Packet state: photon_vars.f90:11 (xp.. are kept in the shifted frame of inttau2.f90:24-26).
// ---------------------------------------------------------------------------------------------
struct Photon {
    double xcur, ycur, zcur;
    double nxp, nyp, nzp;
    double cost, sint, phi;
    double tau, taurun;
    int celli, cellj, cellk;
};

enum { STEP_WALL = 0, STEP_INTERACT = 1, STEP_EXIT = 2 };

// inttau2.f90:208-239 `find`: same result as the bisection over a(1:n) -- 1 when val == a(1), n-1
// when val == a(n), -1 outside, otherwise the largest lo with a(lo) <= val -- reached from an
// arithmetic first guess plus a fix-up against the very same face values.  `a` is 0-based here.
__device__ __forceinline__ int find_cell(double val, const double *a, int n, double inv_w)
{
    if (val < a[0] || val > a[n - 1]) return -1;
    if (val == a[n - 1]) return n - 1;
    int lo = (int)(val * inv_w);
    lo = max(0, min(lo, n - 2));
    while (a[lo] > val) --lo;
    while (a[lo + 1] <= val) ++lo;
    return lo + 1;
}

// sourceph.f90:28-47.  u[0..2] = the three ran2 draws in call order, u[3] = tauint1's draw
// (inttau2.f90:36).  sint = 0, so nxp = sint*cosp and nyp = sint*sinp are (signed) zeros and
// cos(phi)/sin(phi) are not needed; phi itself is kept for the first scattering.
__device__ __forceinline__ void launch(const DevGrid &g, Photon &p, const double u[4])
{
    const double r = u[0] * g.spot_r2;
    const double theta = u[1] * kTWOPI;
    double s, c;
    sincos(theta, &s, &c);
    const double sr = sqrt(r);
    const double xp = sr * c;
    const double yp = sr * s;
    const double zp = g.zp0;
    p.phi = kTWOPI * u[2];
    p.sint = 0.;
    p.cost = -1.;
    p.nxp = 0.;
    p.nyp = 0.;
    p.nzp = -1.;
    p.celli = (int)((double)g.nxg * (xp + g.xmax) / (2. * g.xmax)) + 1;
    p.cellj = (int)((double)g.nyg * (yp + g.ymax) / (2. * g.ymax)) + 1;
    p.cellk = (int)((double)g.nzg * (zp + g.zmax) / (2. * g.zmax)) + 1;
    // inttau2.f90:24-26,36
    p.xcur = xp + g.xmax;
    p.ycur = yp + g.ymax;
    p.zcur = zp + g.zmax;
    p.taurun = 0.;
    p.tau = -log(u[3]);
}

// sourceph.f90:52-70 ranu + :73-101 rang (Marsaglia polar method; the pair's first variate is used), statement for
// statement.  A replayed packet whose draw list runs out leaves the loop (the record is flagged by the caller).
template <class Rng>
__device__ __forceinline__ double rang(Rng &rng, int &ns, double avg, double sigma)
{
    double u = 0., s = 1.;
    while (s >= 1. && !rng.exhausted()) {
        u = -1. + rng.src(ns) * (1. - -1.);
        s = -1. + rng.src(ns) * (1. - -1.);
        s = s * s + u * u;
    }
    const double tmp = u * sqrt(-2. * log(s) / s);
    return avg + sigma * tmp;
}

// Builder-defined launch on top of rang() (the reference defines rang and never calls it; oracle: sourcephGauss):
// xp, yp ~ N(0, sigma), each redrawn while it misses the top face; the rest is sourceph.f90:32-47.  Returns the number
// of draws the launch consumed from the packet's main stream (replay: all of them; Philox: block 0 = 4).
template <class Rng>
__device__ __forceinline__ int launch_gauss(const DevGrid &g, Photon &p, Rng &rng)
{
    int ns = 0;
    double xp, yp;
    do { xp = rang(rng, ns, 0., g.gauss_sigma); } while (!(fabs(xp) < g.xmax) && !rng.exhausted());
    do { yp = rang(rng, ns, 0., g.gauss_sigma); } while (!(fabs(yp) < g.ymax) && !rng.exhausted());
    if (!(fabs(xp) < g.xmax)) xp = 0.;          // only a replayed packet that ran out of draws
    if (!(fabs(yp) < g.ymax)) yp = 0.;
    const double zp = g.zp0;
    double uphi, utau;
    rng.launch_tail(uphi, utau);
    p.phi = kTWOPI * uphi;
    p.sint = 0.;
    p.cost = -1.;
    p.nxp = 0.;
    p.nyp = 0.;
    p.nzp = -1.;
    p.celli = min(g.nxg, (int)((double)g.nxg * (xp + g.xmax) / (2. * g.xmax)) + 1);
    p.cellj = min(g.nyg, (int)((double)g.nyg * (yp + g.ymax) / (2. * g.ymax)) + 1);
    p.cellk = (int)((double)g.nzg * (zp + g.zmax) / (2. * g.zmax)) + 1;
    p.xcur = xp + g.xmax;
    p.ycur = yp + g.ymax;
    p.zcur = zp + g.zmax;
    p.taurun = 0.;
    p.tau = -log(utau);
    return Rng::kSequential ? ns + 2 : 4;
}

// inttau2.f90:242-279 repeat_bounds on the exact arithmetic (TAMC_PERIODIC): a packet that left through a lateral
// face re-enters on the opposite side.  Where the Fortran would print 'Error in Repeat_bounds...' the index stays -1.
__device__ __forceinline__ void repeat_bounds(int &cella, int &cellb, double &acur, double &bcur, double amax, double bmax,
                                              int nag, int nbg, double delta)
{
    if (cella == -1) {
        if (acur < delta) { acur = 2. * amax - delta; cella = nag; }
        else if (acur > 2. * amax - delta) { acur = delta; cella = 1; }
    }
    if (cellb == -1) {
        if (bcur < delta) { bcur = 2. * bmax - delta; cellb = nbg; }
        else if (bcur > 2. * bmax - delta) { bcur = delta; cellb = 1; }
    }
}

// tauint1 returns the centred position and the next call shifts it again (inttau2.f90:65-67 then
// :24-26): (cur - max) + max is not always cur in fp64, so the round trip is replayed.
__device__ __forceinline__ void recentre(const DevGrid &g, Photon &p)
{
    p.xcur = (p.xcur - g.xmax) + g.xmax;
    p.ycur = (p.ycur - g.ymax) + g.ymax;
    p.zcur = (p.zcur - g.zmax) + g.zmax;
}

// One pass of the loop body inttau2.f90:37-63 = one voxel-step.
template <class Tally>
__device__ __forceinline__ int voxel_step(const DevGrid &g, const double *xf, const double *yf, const double *zf,
                                          Photon &p, Tally &tally)
{
    // wall_dist, inttau2.f90:75-121 (face(c+1) is xf[c], face(c) is xf[c-1])
    double dx, dy, dz;
    if (p.nxp > 0.) dx = (xf[p.celli] - p.xcur) / p.nxp;
    else if (p.nxp < 0.) dx = (xf[p.celli - 1] - p.xcur) / p.nxp;
    else dx = 100000.;
    if (p.nyp > 0.) dy = (yf[p.cellj] - p.ycur) / p.nyp;
    else if (p.nyp < 0.) dy = (yf[p.cellj - 1] - p.ycur) / p.nyp;
    else dy = 100000.;
    if (p.nzp > 0.) dz = (zf[p.cellk] - p.zcur) / p.nzp;
    else if (p.nzp < 0.) dz = (zf[p.cellk - 1] - p.zcur) / p.nzp;
    else dz = 100000.;
    double dcell = fmin(fmin(dx, dy), dz);
    const int dir = (dcell == dz) ? 2 : ((dcell == dy) ? 1 : 0);   // later axis wins ties, :116-118

    const double rk = __ldg(g.rhokap + ((long long)p.celli + (long long)g.sx * p.cellj + g.sxy * p.cellk));
    const double taucell = dcell * rk;
    const long long jidx = (long long)(p.celli - 1) + (long long)g.nxg * ((long long)(p.cellj - 1) + (long long)g.nyg * (p.cellk - 1));

    if (p.taurun + taucell < p.tau) {
        p.taurun = p.taurun + taucell;
        tally.add(jidx, taucell);                                  // dcell*rhokap, inttau2.f90:46
        // update_pos(wall_flag=.TRUE.), inttau2.f90:140-170
        if (dir == 0) {
            p.xcur = (p.nxp > 0.) ? xf[p.celli] + g.delta : xf[p.celli - 1] - g.delta;
            p.ycur = p.ycur + p.nyp * dcell;
            p.zcur = p.zcur + p.nzp * dcell;
        } else if (dir == 1) {
            p.xcur = p.xcur + p.nxp * dcell;
            p.ycur = (p.nyp > 0.) ? yf[p.cellj] + g.delta : yf[p.cellj - 1] - g.delta;
            p.zcur = p.zcur + p.nzp * dcell;
        } else {
            p.xcur = p.xcur + p.nxp * dcell;
            p.ycur = p.ycur + p.nyp * dcell;
            p.zcur = (p.nzp > 0.) ? zf[p.cellk] + g.delta : zf[p.cellk - 1] - g.delta;
        }
        // update_voxels, inttau2.f90:201-203
        p.celli = find_cell(p.xcur, xf, g.nxg + 1, g.inv_dx);
        p.cellj = find_cell(p.ycur, yf, g.nyg + 1, g.inv_dy);
        p.cellk = find_cell(p.zcur, zf, g.nzg + 1, g.inv_dz);
        return (p.celli == -1 || p.cellj == -1 || p.cellk == -1) ? STEP_EXIT : STEP_WALL;   // :58-61
    }
    // inttau2.f90:51-55
    dcell = (p.tau - p.taurun) / rk;
    tally.add(jidx, dcell * rk);
    p.xcur = p.xcur + p.nxp * dcell;
    p.ycur = p.ycur + p.nyp * dcell;
    p.zcur = p.zcur + p.nzp * dcell;
    return STEP_INTERACT;
}

// stokes.f90:6-153: new direction after a scattering event (no polarisation state exists).
// u1 feeds stokes.f90:24 / :48, u2 feeds :32 / :64.
__device__ __forceinline__ void stokes(const DevGrid &g, Photon &p, double u1, double u2, double hgg, double g2)
{
    double sinp, cosp;
    if (hgg == 0.0) {
        p.cost = 2. * u1 - 1.;
        double s2 = 1. - p.cost * p.cost;
        p.sint = (s2 <= 0.) ? 0. : sqrt(s2);
        p.phi = kTWOPI * u2;
        sincos(p.phi, &sinp, &cosp);
        p.nxp = p.sint * cosp;
        p.nyp = p.sint * sinp;
        p.nzp = p.cost;
        return;
    }
    const double costp = p.cost, sintp = p.sint, phip = p.phi;
    const double q = (1. - g2) / (1. - hgg + 2. * hgg * u1);
    double bmu = ((1. + g2) - q * q) / (2. * hgg);
    double cosb2 = bmu * bmu;
    if (fabs(bmu) > 1.) {
        bmu = (bmu > 1.) ? 1. : -1.;
        cosb2 = 1.;
    }
    const double sinbt = sqrt(1. - cosb2);
    const double ri1 = kTWOPI * u2;
    const bool upper = ri1 > kPI;
    const double ri = upper ? kTWOPI - ri1 : ri1;     // ri3 (:67) or ri1 (:107)
    double sini, cosi;
    sincos(ri, &sini, &cosi);
    if (bmu == 1. || bmu == -1.) return;              // goto 100: direction left untouched (:71-77)

    p.cost = costp * bmu + sintp * sinbt * cosi;
    double sini2, cosi2 = 0.;
    if (fabs(p.cost) < 1.) {
        p.sint = fabs(sqrt(1. - p.cost * p.cost));
        sini2 = sini * sintp / p.sint;
        const double bott = p.sint * sinbt;
        cosi2 = costp / bott - p.cost * bmu / bott;
    } else {
        p.sint = 0.;
        sini2 = 0.;
        if (p.cost >= 1.) cosi2 = -1.;
        if (p.cost <= -1.) cosi2 = 1.;
    }
    double cosdph = -(cosi2 * cosi) + sini2 * sini * bmu;
    if (fabs(cosdph) > 1.) cosdph = (cosdph > 1.) ? 1. : -1.;
    p.phi = upper ? phip + acos(cosdph) : phip - acos(cosdph);
    if (p.phi > kTWOPI) p.phi = p.phi - kTWOPI;
    if (p.phi < 0.) p.phi = p.phi + kTWOPI;
    sincos(p.phi, &sinp, &cosp);
    p.nxp = p.sint * cosp;
    p.nyp = p.sint * sinp;
    p.nzp = p.cost;
}

// TAMC_FRESNEL on the exact arithmetic: the packet stands just outside the grid on exactly one axis; decide
// reflection and, if reflected, put it back like the extended oracle does (position face -+ delta inside, last
// voxel on that axis, normal cosine flipped, phi rebuilt from the flipped azimuth).
__device__ __forceinline__ bool fresnel_reflect_exact(const DevGrid &g, const double *xf, const double *yf, const double *zf,
                                                      Photon &p, uint2 key, uint32_t id_lo, uint32_t id_hi, int &nb)
{
    const int out = (p.celli == -1) + (p.cellj == -1) + (p.cellk == -1);
    if (out != 1) return false;
    const int a = (p.celli == -1) ? 0 : ((p.cellj == -1) ? 1 : 2);
    const double na = a == 0 ? p.nxp : (a == 1 ? p.nyp : p.nzp);
    double n_in = g.n2;
    if (g.n_g) {                                  // the voxel the packet is leaving (tamc_set_optics_grids)
        const int li = a == 0 ? ((na > 0.) ? g.nxg : 1) : p.celli, lj = a == 1 ? ((na > 0.) ? g.nyg : 1) : p.cellj,
                  lk = a == 2 ? ((na > 0.) ? g.nzg : 1) : p.cellk;
        n_in = __ldg(g.n_g + ((long long)li + (long long)g.sx * lj + g.sxy * lk));
    }
    if (!(boundary_draw(key, id_lo, id_hi, nb) < fresnel_reflectance(n_in, g.n1, fabs(na)))) return false;
    double s, c;
    if (a == 0) {
        p.xcur = (na > 0.) ? xf[g.nxg] - g.delta : xf[0] + g.delta;
        p.celli = (na > 0.) ? g.nxg : 1;
        p.nxp = -p.nxp;
        sincos(p.phi, &s, &c);
        p.phi = atan2(s, -c);
    } else if (a == 1) {
        p.ycur = (na > 0.) ? yf[g.nyg] - g.delta : yf[0] + g.delta;
        p.cellj = (na > 0.) ? g.nyg : 1;
        p.nyp = -p.nyp;
        sincos(p.phi, &s, &c);
        p.phi = atan2(-s, c);
    } else {
        p.zcur = (na > 0.) ? zf[g.nzg] - g.delta : zf[0] + g.delta;
        p.cellk = (na > 0.) ? g.nzg : 1;
        p.nzp = -p.nzp;
        p.cost = -p.cost;
    }
    return true;
}

__device__ __forceinline__ int exit_face(const Photon &p)
{
    if (p.celli == -1) return p.nxp > 0. ? 2 : 1;
    if (p.cellj == -1) return p.nyp > 0. ? 4 : 3;
    if (p.cellk == -1) return p.nzp > 0. ? 6 : 5;
    return 0;
}

// Per-thread counters, folded into the global counters once per warp at kernel end.
struct Counters {
    unsigned long long steps, scatters;
    unsigned int packets, absorbed, errors, overflow, specular, reflections;
    unsigned int exits[6];
    __device__ __forceinline__ void clear()
    {
        steps = scatters = 0ull;
        packets = absorbed = errors = overflow = specular = reflections = 0u;
#pragma unroll
        for (int f = 0; f < 6; ++f) exits[f] = 0u;
    }
    __device__ __forceinline__ void fate(int f)
    {
        packets++;
        if (f == 0) {
            absorbed++;
        } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) exits[i] += (f == i + 1);
        }
    }
    __device__ __forceinline__ void note(int slot)
    {
        if (slot == CNT_SPECULAR) specular++;
        else reflections++;
    }
    __device__ __forceinline__ void death(int f, int nsteps, int nscatt, bool err)
    {
        steps += (unsigned long long)nsteps;
        scatters += (unsigned long long)nscatt;
        errors += err;
        fate(f);
    }
    __device__ __forceinline__ void commit(unsigned long long *g) const
    {
        unsigned long long v[14] = {packets, steps, scatters, absorbed, exits[0], exits[1], exits[2],
                                    exits[3], exits[4], exits[5], errors, overflow, specular, reflections};
#pragma unroll
        for (int i = 0; i < 14; ++i) {
            unsigned long long x = v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
            const int slot = i < 12 ? i : i + 1;            // CNT_WORK sits between CNT_OVERFLOW and CNT_SPECULAR
            if ((threadIdx.x & 31) == 0 && x) atomicAdd(g + slot, x);
        }
    }
};

// mcpolar.f90:153-169 for one packet: launch, tauint1, then either the shipped stub (:166-169:
// the packet ends at its first interaction) or, with TAMC_SCATTER, the albedo test + stokes loop.
template <class Rng, class Tally, bool kRecord>
__device__ __forceinline__ void transport_packet(const DevGrid &g, const double *xf, const double *yf, const double *zf,
                                                 Rng &rng, Tally &tally, Counters &cnt, tamc_packet_record *rec,
                                                 long long draws_available)
{
    double u[4];
    Photon p;
    int ndraws = 4, steps = 0, nscatt = 0, fate = 0;
    if (g.gauss_sigma > 0.) {
        ndraws = launch_gauss(g, p, rng);
    } else {
        rng.block(u);
        launch(g, p, u);
    }
    tally.begin();
    int nb = 0;
    bool specular = false;
    if constexpr (Rng::kHasBoundary) {
        double r0sq = g.r0sq;
        if ((g.flags & TAMC_FRESNEL) && g.n_g) {
            const double n_in = __ldg(g.n_g + ((long long)p.celli + (long long)g.sx * p.cellj + g.sxy * p.cellk));
            r0sq = ((g.n1 - n_in) / (g.n1 + n_in)) * ((g.n1 - n_in) / (g.n1 + n_in));
        }
        if ((g.flags & TAMC_FRESNEL) && boundary_draw(rng.key, rng.id_lo, rng.id_hi, nb) < r0sq) {
            specular = true;                                       // reflected at the top surface before entering
            fate = 6;
            ndraws = 3;                                            // the optical depth was never drawn
            p.nzp = 1.;
            p.cellk = -1;
            cnt.specular++;
        }
    }
    while (!specular) {
        int r = voxel_step(g, xf, yf, zf, p, tally);
        ++steps;
        if (r == STEP_EXIT && (g.flags & TAMC_PERIODIC) && (p.celli == -1 || p.cellj == -1)) {
            repeat_bounds(p.celli, p.cellj, p.xcur, p.ycur, g.xmax, g.ymax, g.nxg, g.nyg, g.delta);
            if (p.celli != -1 && p.cellj != -1 && p.cellk != -1) r = STEP_WALL;      // re-entered: the flight goes on
        }
        if constexpr (Rng::kHasBoundary) {
            if (r == STEP_EXIT && (g.flags & TAMC_FRESNEL) &&
                fresnel_reflect_exact(g, xf, yf, zf, p, rng.key, rng.id_lo, rng.id_hi, nb)) {
                cnt.reflections++;
                r = STEP_WALL;
            }
        }
        if (r == STEP_WALL) {
            if (steps >= kMaxStepsPerPacket) { cnt.errors++; break; }
            continue;
        }
        if (r == STEP_EXIT) {
            fate = exit_face(p);
            break;
        }
        if (!(g.flags & TAMC_SCATTER)) break;                      // stub: tflag = .true.; exit
        rng.block(u);
        double albedo = g.albedo, hgg = g.hgg, g2 = g.g2;
        if (g.albedo_g || g.hgg_g) {                               // the voxel of the interaction (tamc_set_optics_grids)
            const long long v = (long long)p.celli + (long long)g.sx * p.cellj + g.sxy * p.cellk;
            if (g.albedo_g) albedo = __ldg(g.albedo_g + v);
            if (g.hgg_g) { hgg = __ldg(g.hgg_g + v); g2 = hgg * hgg; }
        }
        if (u[0] < albedo) {
            stokes(g, p, u[1], u[2], hgg, g2);
            ++nscatt;
            ndraws += 4;
            recentre(g, p);
            p.taurun = 0.;
            p.tau = -log(u[3]);
        } else {
            ndraws += 1;
            break;                                                 // absorbed
        }
    }
    tally.flush();
    cnt.steps += (unsigned long long)steps;
    cnt.scatters += (unsigned long long)nscatt;
    cnt.fate(fate);
    if (kRecord) {
        const bool over = ndraws > draws_available;
        cnt.overflow += over;
        rec->xp = p.xcur - g.xmax; rec->yp = p.ycur - g.ymax; rec->zp = p.zcur - g.zmax;
        rec->nxp = p.nxp; rec->nyp = p.nyp; rec->nzp = p.nzp;
        rec->deposit = tally.packet_sum;
        rec->xcell = p.celli; rec->ycell = p.cellj; rec->zcell = p.cellk;
        rec->steps = steps; rec->nscatt = nscatt; rec->ndraws = ndraws;
        rec->fate = fate; rec->flags = over ? 1 : 0;
    }
}

// Copies the three face tables into shared memory; returns pointers to each.
__device__ __forceinline__ void stage_faces(const DevGrid &g, double *smem, const double *&xf, const double *&yf,
                                            const double *&zf)
{
    const int n = g.nxg + g.nyg + g.nzg + 3;
    for (int i = threadIdx.x; i < n; i += blockDim.x) smem[i] = g.faces[i];
    __syncthreads();
    xf = smem;
    yf = smem + (g.nxg + 1);
    zf = yf + (g.nyg + 1);
}

}  // namespace tamc
