// tamc_flight.cuh -- scatter-regime transport (TAMC_SCATTER), one packet per lane, flight by flight.
//
// Replaces the scatter loop the driver's shell implies (mcpolar.f90:165-169 around tauint1, inttau2.f90:7-72,
// and stokes.f90:6-153) for the default production path.  The work-queue kernel of round 1
// (tamc_pool.cuh) was issue-bound with 40 warp instructions per scattering event, 40 % of them queue management
// (profiles/r02a_skin200_pool_ncu_summary.txt).  What is different here (DESIGN.md 3b has the measurements):
//
// 1. The voxel walk of one flight (tauint1's loop, inttau2.f90:37-63) is a 3-D DDA on the RAY PARAMETER.  At the
//    start of a flight the three distances to the next x / y / z face are formed as wall_dist does
//    (inttau2.f90:75-121, per-event reciprocals as in tamc_fast.cuh); after that a crossing of axis a only adds the
//    constant dt_a = (w_a - delta) * |1/n_a| to that axis' entry.  This is the reference's geometry, not an
//    approximation of it: update_pos (inttau2.f90:140-170) puts the crossed coordinate at `face +- delta`, i.e.
//    delta INSIDE the next voxel, so the next wall on that axis is a full voxel edge minus delta away, while the
//    other two coordinates advance by n * dcell and keep their distances.  (Faces are (i-1)*2*max/n,
//    gridset.f90:23-31, so w_a is constant up to the rounding of the face table, ~1e-13 of an edge.)  A voxel-step
//    is then: min of three, one subtraction, one product, one compare, predicated updates -- no face look-ups, no
//    position update, no division.
//
// 2. Between flights a packet is described RELATIVE TO ITS VOXEL: the voxel index, per axis the number of crossings
//    left before the grid ends (r*), and per axis the distance e_a to the face ahead -- (t_a - t_end) * |n_a| at the
//    end of a flight, `edge - e_a` when the new direction looks at the other face.  Absolute positions and the
//    face tables are touched once per packet, at launch.  (The centred round trip of inttau2.f90:65-67 / :24-26
//    is a rounding of the absolute position and has no counterpart here; the deviation is of the order of the
//    production arithmetic's, DESIGN.md "Two arithmetics".)
//
// 3. No shared-memory pool.  Every lane keeps its packet in registers and the warp alternates between an EVENT
//    phase (end of flight, albedo test + stokes rotation + next optical depth -- or, for a lane whose packet
//    ended, the launch of a new one -- all sharing one Philox block, one log and one sincos) and a WALK phase
//    that steps every walking lane until fewer than `walk_min` lanes are still in flight.  Flights are short
//    (2.4 voxel-steps per scattering in the layered-skin grid), so letting the few long flights run on while the
//    rest of the warp waits is cheaper than moving packets through queues.
//
// 4. The walk step is one block of predicated PTX (walk_step): state updated in place, nothing copied at a
//    reconvergence point; the opacity of the voxel two crossings ahead is requested with cp.async into a per-lane
//    shared-memory slot and picked up two steps later (cp.async.wait_group 1).
//
// Memory layout: while the grids fit L2 the kernel reads a compact copy of the opacities (k_rk_compact: no halo,
// the tally's own index) and tallies into g.jmean -- separate arrays on purpose: under a narrow beam the REDs into
// the few voxels below it queue up in their L2 slices, and loads of the same sectors would wait behind them
// (measured: 2x).  Beyond L2 (400^3) one 16-byte record per voxel, vox[idx] = {rhokap, jmean} (k_vox_pack /
// k_vox_unpack), halves the sectors per visited voxel.  The deposit of the partial step that ends a flight stays in
// a register (`pend`): the packet is still in that voxel after the scattering, so it is added to the first deposit
// of the next flight and both go out as one RED when the packet leaves the voxel (or is absorbed).
//
// Same Philox streams (one block per event, counter = (packet id, event index)) and the same production
// arithmetic for the launch and the rotation as tamc_fast.cuh; tests/test_gpu_production.py checks the grid and
// the counters of this kernel against the oracle on the same streams, and statistically on the ran2 streams.
#pragma once

#include "tamc_fast.cuh"

namespace tamc {

constexpr double kFar = 1e300;

// |v| is a normal number far from both ends of the exponent range (2^-1007 .. 2^993): one integer compare on the high
// word, no 64-bit immediate -- the guard in front of the range-restricted reciprocal / roots of tamc_math.cuh
__device__ __forceinline__ bool mid_range(double v)
{
    return (unsigned)((__double2hiint(v) & 0x7fffffff) - 0x01000000) < 0x7d000000u;
}       // ray parameter of a face that is never reached (direction cosine 0)

// FL_EVENT + 4: the flight ended in a step that held its voxel's opacity in the second register (walk_step<1>);
// FL_EXITED + axis: the packet left the grid through a face of that axis
enum { FL_DEAD = 0, FL_EVENT = 1, FL_WALK = 2, FL_EVENT_B = 5, FL_EXITED = 8 };

// vox[idx] = {rhokap, 0} from the resident grid (halo layout); one thread per voxel, x fastest.
__global__ void __launch_bounds__(256) k_vox_pack(const DevGrid g, double2 *__restrict__ vox)
{
    const long long nvox = (long long)g.nxg * g.nyg * g.nzg;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += stride) {
        const int ci = (int)(i % g.nxg);
        const long long r = i / g.nxg;
        const int cj = (int)(r % g.nyg), ck = (int)(r / g.nyg);
        vox[i] = make_double2(__ldg(g.rhokap + ((ci + 1) + (long long)g.sx * (cj + 1) + g.sxy * (ck + 1))), 0.);
    }
}

// rkc[idx] = rhokap(i,j,k): the opacities without the halo, in the tally's own index
__global__ void __launch_bounds__(256) k_rk_compact(const DevGrid g, double *__restrict__ rkc)
{
    const long long nvox = (long long)g.nxg * g.nyg * g.nzg;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += stride) {
        const int ci = (int)(i % g.nxg);
        const long long r = i / g.nxg;
        const int cj = (int)(r % g.nyg), ck = (int)(r / g.nyg);
        rkc[i] = __ldg(g.rhokap + ((ci + 1) + (long long)g.sx * (cj + 1) + g.sxy * (ck + 1)));
    }
}

// per-voxel albedo / hgg (tamc_set_optics_grids) in the tally's own index, for the flight kernel's event phase
__global__ void __launch_bounds__(256) k_optics_compact(const DevGrid g, double *__restrict__ albc, double *__restrict__ hggc)
{
    const long long nvox = (long long)g.nxg * g.nyg * g.nzg;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += stride) {
        const int ci = (int)(i % g.nxg);
        const long long r = i / g.nxg;
        const int cj = (int)(r % g.nyg), ck = (int)(r / g.nyg);
        const long long v = (ci + 1) + (long long)g.sx * (cj + 1) + g.sxy * (ck + 1);
        if (albc) albc[i] = __ldg(g.albedo_g + v);
        if (hggc) hggc[i] = __ldg(g.hgg_g + v);
    }
}

// jmean(i,j,k) = vox[idx].y
__global__ void __launch_bounds__(256) k_vox_unpack(const DevGrid g, const double2 *__restrict__ vox)
{
    const long long nvox = (long long)g.nxg * g.nyg * g.nzg;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvox; i += stride) g.jmean[i] = vox[i].y;
}

// The direction part of a scattering event with the sine / cosine of the azimuthal rotation angle already formed
// (the event phase shares that sincos with the launch): stokes.f90:40-148 as a rotation of the direction vector,
// exactly scatter_dir() of tamc_fast.cuh.  u1 -> stokes.f90:24 / :48.  (si, ci) = sin / cos of ri1 = TWOPI*u2
// (isotropic: of phi = TWOPI*u2, stokes.f90:32).
__device__ __forceinline__ void scatter_rotate(double hgg, const ScatterConsts &sc, double &nzp, double &sint, double &cosp, double &sinp,
                                               double u1, double si, double ci)
{
    if (hgg == 0.0) {                                     // isotropic, stokes.f90:23-38
        const double cost = 2. * u1 - 1.;
        const double s2 = 1. - cost * cost;
        sint = (s2 <= 0.) ? 0. : sqrt(s2);
        sinp = si;
        cosp = ci;
        nzp = cost;
        return;
    }
    // (1 - g + 2 g u lies in [1 - |g|, 1 + |g|], 1 - bmu^2 in (0, 1] once bmu = +-1 has left: normal numbers, so the
    // range-restricted reciprocal / roots of tamc_math.cuh apply: ~1 ulp, a third of the library versions' instructions)
    const double q = sc.one_m_g2 * fm::rcp_normal(sc.one_m_g + sc.two_g * u1);   // stokes.f90:48
    double bmu = (sc.one_p_g2 - q * q) * sc.inv_two_g;
    bmu = fmin(1., fmax(-1., bmu));
    if (bmu == 1. || bmu == -1.) return;                               // goto 100, stokes.f90:71-77
    const double s2b = 1. - bmu * bmu;
    const double sinbt = mid_range(s2b) ? fm::sqrt_normal(s2b) : sqrt(s2b);
    const double costp = nzp, sintp = sint;
    const double nxp = sintp * cosp, nyp = sintp * sinp;
    const double a = sinbt * ci * costp, b = sinbt * si;
    double uz = costp * bmu + sintp * sinbt * ci;                      // stokes.f90:79 / :117
    const double ux = bmu * nxp - (a * cosp - b * sinp);
    const double uy = bmu * nyp - (a * sinp + b * cosp);
    uz = fmin(1., fmax(-1., uz));
    // sint = sqrt(1 - cost^2) (stokes.f90:81 / :119) is the length of the lateral part (ux, uy) of the unit vector: one
    // reciprocal root gives it and the new azimuth
    const double h2 = ux * ux + uy * uy;
    if (h2 > 0.) {
        const double ih = mid_range(h2) ? fm::rsqrt_normal(h2) : rsqrt(h2);
        cosp = ux * ih;
        sinp = uy * ih;
        sint = h2 * ih;
    } else {
        sint = 0.;
    }
    nzp = uz;
}

// One voxel-step (inttau2.f90:37-63) of every lane in flight, as straight-line predicated PTX: every state variable is
// updated in place, so a lane that is not walking passes through untouched and nothing is copied at a reconvergence
// point (a compiler-generated copy of a prefetched opacity waits for the load right there).
//   taucell = (tmin - tcur) * rcur                      inttau2.f90:39-40
//   wall    = walking && taucell < taul                 :42   (taurun + taucell < tau)
//   wall:   RED(tally[idx], pend + taucell), pend = 0, taul -= taucell, tcur = tmin,           :43-46
//           t[ax] += dt[ax], r[ax] -= 1, idx += sn                                              :48 (update_pos: face +- delta)
//           rn < 0 ? mode = EXITED + ax                                                         :57-61
//           the crossing after this one: tmin = min3(t) (later axis wins ties, :116-118), ax, sn, rn
//           slot <- opacity[idx + sn] (cp.async) if that crossing stays inside: the opacity two voxels ahead
//   else walking: mode = EVENT                                                                  :50-55 -> event phase
// Opacity pipeline: each lane owns two 8-byte slots of shared memory.  A step first waits for the copy issued two steps
// ago (cp.async.wait_group 1: the one issued in the previous step may still be in flight), reads its slot -- the opacity
// of the voxel the packet is in -- and at its end asks for the opacity two voxels ahead into the same slot; the caller
// alternates the slots (walk_step<0>(slot0), walk_step<1>(slot1)).  All lanes in flight cross exactly one face per step,
// so the parity is the same for the whole warp.  (A register prefetch cannot do this: ptxas puts both loads on one
// scoreboard, and a wait for the older one waits for the newer one too.)  Conditional fp64 updates are written as fma
// with a 1.0 / 0.0 factor (exact), which is one instruction where a predicated fp64 add becomes an add and two selects.
// kAgg (north-star: "warp-aggregated (shuffle-reduced) atomics"): lanes of the warp that tally into the SAME voxel in
// this step are found with match.any, their deposits summed through shuffles, and one lane issues the RED.  Measured
// (profiles/README.md, round 2): no gain on the layered-skin grid -- the deposits of a step are spread over 20-30
// different voxels per warp and the L2 atomic unit is not the bound -- so the default build issues one RED per lane.
__device__ __forceinline__ void aggregated_red(double *addr, double v, bool pv)
{
    const unsigned act = __ballot_sync(0xffffffffu, pv);
    if (!pv) return;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned grp = __match_any_sync(act, (unsigned long long)addr);
    const int most = __reduce_max_sync(act, __popc(grp));
    double sum = v;
    for (int k = 1; k < most; ++k) {
        // the k-th other member of this lane's group, or the lane itself (adds nothing) when the group is smaller
        const unsigned others = grp & ~(1u << lane);
        const unsigned src = __fns(others, 0, k);
        const double o = __shfl_sync(act, v, src < 32u ? src : lane);
        if (src < 32u) sum += o;
    }
    if ((unsigned)(__ffs(grp) - 1) == lane) atomicAdd(addr, sum);
}

template <int kParity, bool kAgg = false>
__device__ __forceinline__ void walk_step(double &tx, double &ty, double &tz, double &tcur, double &tmin, double &taul, double &pend,
                                          double &rcur, int &idx, int &rx, int &ry, int &rz, int &ax, int &sn, int &rn,
                                          int &steps, int &mode, double dtx, double dty, double dtz, int sax, int say, int saz,
                                          const double *rkb, unsigned slot, double *jmb, int stride)
{
    if (kAgg) {
        // the deposit of this step, formed as in the block below (which then skips its own RED), through the aggregated RED
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        const bool pw = mode == FL_WALK;
        double rc = rcur;
        if (pw) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(rc) : "r"(slot) : "memory");
        const double tc = (tmin - tcur) * rc;
        const bool wall = pw && tc < taul;
        const double v = pend + tc;
        aggregated_red(jmb + (size_t)idx * (size_t)(stride >> 3), v, wall && v != 0.);
    }
    asm volatile(
        "{\n\t"
        ".reg .pred pw, pwall, pnw, pev, pv, pz, py, px, pout, pin, pxy, pmz, pld;\n\t"
        ".reg .f64 tc, v, m, d, fw, fz, fy, fx;\n\t"
        ".reg .b32 hw, hz, hy, hx, zero;\n\t"
        ".reg .s32 r1, s1, i2, mo;\n\t"
        ".reg .u64 ad, an;\n\t"
        "mov.b32 zero, 0;\n\t"
        "setp.eq.s32 pw, %16, 2;\n\t"
        // the opacity of this voxel: asked for two steps ago, landed in this lane's slot of shared memory
        "cp.async.wait_group 1;\n\t"
        "@pw ld.shared.f64 %7, [%24];\n\t"
        "sub.f64 d, %4, %3;\n\t"
        "mul.f64 tc, d, %7;\n\t"
        "setp.lt.and.f64 pwall, tc, %5, pw;\n\t"
        "mad.wide.s32 ad, %8, %26, %25;\n\t"
        "@pwall add.s32 %8, %8, %13;\n\t"
        "setp.ge.and.s32 pin, %14, 0, pwall;\n\t"
        "setp.lt.and.s32 pout, %14, 0, pwall;\n\t"
        // the faces crossed by this step, then the crossing after the next one and the request for the opacity behind it
        "setp.eq.and.s32 pz, %12, 2, pwall;\n\t"
        "setp.eq.and.s32 py, %12, 1, pwall;\n\t"
        "setp.eq.and.s32 px, %12, 0, pwall;\n\t"
        "selp.b32 hz, 0x3ff00000, 0, pz;\n\t"
        "selp.b32 hy, 0x3ff00000, 0, py;\n\t"
        "selp.b32 hx, 0x3ff00000, 0, px;\n\t"
        "mov.b64 fz, {zero, hz};\n\t"
        "mov.b64 fy, {zero, hy};\n\t"
        "mov.b64 fx, {zero, hx};\n\t"
        "fma.rn.f64 %2, %19, fz, %2;\n\t"
        "fma.rn.f64 %1, %18, fy, %1;\n\t"
        "fma.rn.f64 %0, %17, fx, %0;\n\t"
        "@pz add.s32 %11, %11, -1;\n\t"
        "@py add.s32 %10, %10, -1;\n\t"
        "@px add.s32 %9, %9, -1;\n\t"
        "add.s32 mo, %12, 8;\n\t"
        "@pout mov.s32 %16, mo;\n\t"
        "setp.lt.f64 pxy, %0, %1;\n\t"
        "selp.f64 m, %0, %1, pxy;\n\t"
        "setp.lt.f64 pmz, m, %2;\n\t"
        "selp.s32 r1, %9, %10, pxy;\n\t"
        "selp.s32 r1, r1, %11, pmz;\n\t"
        "add.s32 r1, r1, -1;\n\t"
        "selp.s32 s1, %20, %21, pxy;\n\t"
        "selp.s32 s1, s1, %22, pmz;\n\t"
        "setp.ge.and.s32 pld, r1, 0, pin;\n\t"
        "add.s32 i2, %8, s1;\n\t"
        "mad.wide.s32 an, i2, %26, %23;\n\t"
        "@pld cp.async.ca.shared.global [%24], [an], 8;\n\t"
        "cp.async.commit_group;\n\t"
        // the rest of the step
        "not.pred pnw, pwall;\n\t"
        "and.pred pev, pw, pnw;\n\t"
        "@pw add.s32 %15, %15, 1;\n\t"
        "@pev mov.s32 %16, 1;\n\t"
        "add.f64 v, %6, tc;\n\t"
        "setp.neu.and.f64 pv, v, 0d0000000000000000, pwall;\n\t"
        "setp.eq.and.s32 pv, %27, 0, pv;\n\t"
        "@pv red.global.add.f64 [ad], v;\n\t"
        "selp.b32 hw, 0x3ff00000, 0, pwall;\n\t"
        "mov.b64 fw, {zero, hw};\n\t"
        "neg.f64 v, %6;\n\t"
        "fma.rn.f64 %6, v, fw, %6;\n\t"
        "neg.f64 v, tc;\n\t"
        "fma.rn.f64 %5, v, fw, %5;\n\t"
        "fma.rn.f64 %3, d, fw, %3;\n\t"
        "selp.f64 %4, m, %2, pmz;\n\t"
        "selp.s32 %12, 0, 1, pxy;\n\t"
        "selp.s32 %12, %12, 2, pmz;\n\t"
        "@pwall mov.s32 %14, r1;\n\t"
        "@pwall mov.s32 %13, s1;\n\t"
        "}"
        : "+d"(tx), "+d"(ty), "+d"(tz), "+d"(tcur), "+d"(tmin), "+d"(taul), "+d"(pend), "+d"(rcur), "+r"(idx), "+r"(rx),
          "+r"(ry), "+r"(rz), "+r"(ax), "+r"(sn), "+r"(rn), "+r"(steps), "+r"(mode)
        : "d"(dtx), "d"(dty), "d"(dtz), "r"(sax), "r"(say), "r"(saz), "l"(rkb), "r"(slot), "l"(jmb), "r"(stride), "r"(kAgg ? 1 : 0)
        : "memory");
}

template <int kBlock, int kMinCtas, bool kInter, bool kGrids = false, bool kAgg = false>
__global__ void __launch_bounds__(kBlock, kMinCtas) k_transport_flight(const DevGrid g, double *__restrict__ vox, long long n,
                                                                   uint64_t first_id, int chunk, int walk_min,
                                                                   unsigned long long *__restrict__ cnt,
                                                                   const double *__restrict__ albc = nullptr,
                                                                   const double *__restrict__ hggc = nullptr, int launch_min = 1)
{
    extern __shared__ double s_faces[];
    const double *xf, *yf, *zf;
    stage_faces(g, s_faces, xf, yf, zf);
    const int nfaces = g.nxg + g.nyg + g.nzg + 3;
    unsigned long long *wc = reinterpret_cast<unsigned long long *>(s_faces + nfaces) + (threadIdx.x >> 5) * CNT_N;
    // two opacity slots per lane (walk_step): the asynchronous copies of the walk land here
    double *slots = reinterpret_cast<double *>(reinterpret_cast<unsigned long long *>(s_faces + nfaces) + (kBlock >> 5) * CNT_N);
    unsigned slot0 = (unsigned)__cvta_generic_to_shared(slots + threadIdx.x);
    asm volatile("" : "+r"(slot0));
    unsigned slot1 = (unsigned)__cvta_generic_to_shared(slots + kBlock + threadIdx.x);
    asm volatile("" : "+r"(slot1));          // keep both addresses in registers: re-deriving them costs three instructions per step

    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    if (lane < CNT_N) wc[lane] = 0ull;
    __syncwarp();

    // kInter: vox = interleaved records {rhokap, jmean} (16 bytes per voxel; grids beyond L2: one sector per visited voxel);
    // otherwise vox = a compact copy of the opacities (no halo, the tally's own index) and the tally is g.jmean -- under a
    // narrow beam the REDs into the few voxels below it queue up in their L2 slices, and loads of the same sectors would
    // wait behind them
    const double *rkb = vox;
    double *jmb = kInter ? vox + 1 : g.jmean;
    constexpr int ws = kInter ? 2 : 1;                    // doubles per voxel record
    const int nxy = g.nxg * g.nyg;
    // voxel edges; w* = the edge minus the snap: the distance to the next face on an axis right after crossing one (header)
    const double fwx = g.fwx, fwy = g.fwy, fwz = g.fwz, wx = g.wx, wy = g.wy, wz = g.wz;
    const double ez0 = g.ez0;                              // launch: distance down to the bottom face of the launch voxel

    // ---- the packet of this lane.  Position is held relative to the voxel: during a flight through (t*, dt*), between
    // flights as e* = distance to the face AHEAD on each axis (ahead = the side the stride sa* points to).
    int mode = FL_DEAD;
    double nzp = -1., sint = 0., cosp = 1., sinp = 0.;
    double tx = kFar, ty = kFar, tz = kFar, dtx = 0., dty = 0., dtz = 0.;   // a direction cosine of 0: t = kFar, dt stashes e
    double tcur = 0., tmin = 0., taul = 0.;             // ray parameter now / at the next crossing; optical depth left
    double pend = 0., rkc = 0.;                         // pending deposit; opacity of this voxel
    int ax = 2, sn = 0, rn = 0;                         // the next crossing: axis, index stride, crossings left on that axis after it
    int idx = 0, sax = 1, say = 1, saz = 1, rx = 0, ry = 0, rz = 0;
    int steps = 0, ns = 0;
    uint32_t id_lo = 0u, id_hi = 0u;
    // warp-uniform: the chunk of packet ids this warp owns
    long long next = 0, end = 0;
    bool exhausted = false;
    // per-lane accumulators, folded into the warp's counters at the end
    unsigned long long acc_steps = 0ull, acc_scat = 0ull;
    unsigned int acc_pk = 0u, acc_abs = 0u;

    for (;;) {
        // =====================================================================================================
        // EVENT phase: every lane that is not in flight
        // =====================================================================================================
        // (a) packets that left the grid (inttau2.f90:57-61): the crossed axis and its direction give the face
        if (mode >= FL_EXITED) {
            const int xa = mode - FL_EXITED;
            const int f = xa == 0 ? (sax > 0 ? 2 : 1) : (xa == 1 ? (say > 0 ? 4 : 3) : (saz > 0 ? 6 : 5));
            atomicAdd(&wc[CNT_EXIT0 + f - 1], 1ull);
            acc_steps += (unsigned long long)steps;
            acc_scat += (unsigned long long)ns;
            acc_pk++;
            mode = FL_DEAD;
        }
        // (b) packet ids for the lanes without a packet
        const unsigned deadm = __ballot_sync(full, mode == FL_DEAD);
        bool fresh = false;
        // launches are batched: the launch path (a second sincos, a root) runs for the whole warp whenever one lane is
        // fresh, so lanes without a packet wait until `launch_min` of them can launch together -- unless nothing else is
        // left to do in this warp
        const bool launch_now = __popc(deadm) >= launch_min || __ballot_sync(full, mode != FL_DEAD) == 0u;
        if (deadm && !exhausted && launch_now) {
            if (next >= end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(cnt + CNT_WORK, (unsigned long long)chunk);
                base = __shfl_sync(full, base, 0);
                if ((long long)base >= n) exhausted = true;
                else { next = (long long)base; end = min(next + chunk, n); }
            }
            if (!exhausted) {
                const int rank = __popc(deadm & lt_mask);
                const long long avail = end - next;
                if (mode == FL_DEAD && rank < avail) {
                    const uint64_t gid = first_id + (uint64_t)(next + rank);
                    id_lo = (uint32_t)gid;
                    id_hi = (uint32_t)(gid >> 32);
                    ns = 0;
                    fresh = true;
                }
                next += min((long long)__popc(deadm), avail);
            }
        }
        // (c) one Philox block, one sincos, one log per lane: the scattering event or the launch
        const bool at_event = mode == FL_EVENT || mode == FL_EVENT_B;
        if (at_event || fresh) {
            const uint4 r = philox_block(g, id_lo, id_hi, fresh ? 0u : (uint32_t)ns + 1u);
            // azimuth: launch phi = TWOPI*u (sourceph.f90:34); scattering ri1 = TWOPI*u, beyond PI the reference works with
            // ri3 = TWOPI - ri1 (stokes.f90:66-68; with the truncated constants not exactly -ri1 mod 2 pi: same ri3 here)
            const double ri1 = kTWOPI * unit_fast(r.z);
            // kGrids: albedo / hgg of the voxel the interaction happens in (tamc_set_optics_grids), else the scalars
            double albedo = g.albedo, hgg = g.hgg;
            ScatterConsts sc = g.sc;
            if (kGrids && !fresh) {
                if (albc) albedo = __ldg(albc + idx);
                if (hggc) { hgg = __ldg(hggc + idx); sc = make_scatter_consts(hgg); }
            }
            const bool upper = !fresh && hgg != 0.0 && ri1 > kPI;
            double si, co;
            fm::sincospi_0_2((upper ? kTWOPI - ri1 : ri1) * kInvPi, &si, &co);
            si = upper ? -si : si;
            bool alive = true;
            if (fresh) {
                // sourceph.f90:28-47: straight down from the disk; the flight needs no reciprocals
                double x, y;
                int ci, cj;
                launch_point(g, r.x, r.y, x, y, ci, cj);
                nzp = -1.; sint = 0.;                                   // sourceph.f90:37-42
                cosp = co; sinp = si;
                pend = 0.;
                steps = 0;
                idx = (ci - 1) + g.nxg * ((cj - 1) + g.nyg * (g.cellk0 - 1));
                rkc = __ldg(rkb + (size_t)idx * ws);
                tx = kFar; dtx = xf[ci] - x;                            // cosine 0: never crossed; e stashed in dt
                ty = kFar; dty = yf[cj] - y;
                tz = ez0; dtz = wz;
                sax = 1; say = g.nxg; saz = -nxy;
                rx = g.nxg - ci; ry = g.nyg - cj; rz = g.cellk0 - 1;
            } else {
                // ---- end of the flight (inttau2.f90:50-55): where the packet stands in its voxel
                pend += taul;                                           // dcell*rhokap = ((tau-taurun)/rhokap)*rhokap
                const double tend = tcur + taul * (mid_range(rkc) ? fm::rcp_normal(rkc) : __drcp_rn(rkc));   // inttau2.f90:51
                const double ex = tx < kFar ? (tx - tend) * fabs(sint * cosp) : dtx;
                const double ey = ty < kFar ? (ty - tend) * fabs(sint * sinp) : dty;
                const double ez = tz < kFar ? (tz - tend) * fabs(nzp) : dtz;
                if (unit_fast(r.x) < albedo) {                          // draw < albedo ? stokes : absorbed
                    scatter_rotate(hgg, sc, nzp, sint, cosp, sinp, unit_fast(r.y), si, co);
                    ++ns;
                    // ---- start of the next flight: wall_dist (inttau2.f90:75-121) from the in-voxel distances
                    const double nxp = sint * cosp, nyp = sint * sinp;
                    // an axis whose cosine changed sign now looks at the opposite face (a cosine of exactly 0 keeps its face)
                    const bool fx = (nxp < 0.) != (sax < 0) && nxp != 0., fy = (nyp < 0.) != (say < 0) && nyp != 0.,
                               fz = (nzp < 0.) != (saz < 0) && nzp != 0.;
                    const double ax_ = fx ? fwx - ex : ex, ay_ = fy ? fwy - ey : ey, az_ = fz ? fwz - ez : ez;
                    rx = fx ? (g.nxg - 1) - rx : rx;
                    ry = fy ? (g.nyg - 1) - ry : ry;
                    rz = fz ? (g.nzg - 1) - rz : rz;
                    sax = fx ? -sax : sax;
                    say = fy ? -say : say;
                    saz = fz ? -saz : saz;
                    const double xy = nxp * nyp;
                    const double xyz = xy * nzp;
                    if (mid_range(xyz)) {
                        // the common case, no cosine is 0: one reciprocal for the three (set_direction, tamc_fast.cuh)
                        const double rr = fm::rcp_normal(xyz);
                        const double inz = fabs(xy * rr);
                        const double rzz = rr * nzp;
                        const double inx = fabs(nyp * rzz), iny = fabs(nxp * rzz);
                        tx = ax_ * inx; dtx = wx * inx;
                        ty = ay_ * iny; dty = wy * iny;
                        tz = az_ * inz; dtz = wz * inz;
                    } else {
                        // a cosine of 0 (or a product of cosines beyond the range of the fast reciprocal): axis by axis
                        const bool zx = nxp == 0., zy = nyp == 0., zzr = nzp == 0.;
                        const double inx = zx ? 0. : fabs(1. / nxp), iny = zy ? 0. : fabs(1. / nyp), inz = zzr ? 0. : fabs(1. / nzp);
                        tx = zx ? kFar : ax_ * inx; dtx = zx ? ax_ : wx * inx;
                        ty = zy ? kFar : ay_ * iny; dty = zy ? ay_ : wy * iny;
                        tz = zzr ? kFar : az_ * inz; dtz = zzr ? az_ : wz * inz;
                    }
                } else {
                    if (pend != 0.) atomicAdd(jmb + (size_t)idx * ws, pend);
                    acc_steps += (unsigned long long)steps;
                    acc_scat += (unsigned long long)ns;
                    acc_pk++;
                    acc_abs++;
                    alive = false;
                    mode = FL_DEAD;
                }
            }
            if (alive) {
                // the first crossing of the flight, and the opacity behind it on its way while the log runs
                const bool pxy = tx < ty;                               // ties: the later axis wins (inttau2.f90:116-118)
                const double m = pxy ? tx : ty;
                const bool pmz = m < tz;
                tmin = pmz ? m : tz;
                ax = pmz ? (pxy ? 0 : 1) : 2;
                rn = (pmz ? (pxy ? rx : ry) : rz) - 1;                  // < 0: that crossing leaves the grid
                sn = pmz ? (pxy ? sax : say) : saz;
                // slot 0 <- this voxel's opacity, slot 1 <- (asynchronously) the one behind the first face.  Copies of the
                // flight that just ended may still be on their way into these slots: let them land first.
                asm volatile("cp.async.wait_all;" ::: "memory");
                slots[threadIdx.x] = rkc;
                {
                    const double *nx_ = rkb + (size_t)(idx + sn) * ws;
                    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\t@q cp.async.ca.shared.global [%0], [%1], 8;\n\tcp.async.commit_group;\n\t}"
                                 ::"r"(slot1), "l"(nx_), "r"((int)(rn >= 0)) : "memory");
                }
                taul = fm::neglog_u32(r.w);                             // inttau2.f90:36
                tcur = 0.;
                mode = FL_WALK;
            }
        }

        // =====================================================================================================
        // WALK phase: one voxel-step (inttau2.f90:37-63) per pass for every lane in flight.  Software-pipelined: the
        // crossing of this pass (tmin, ax) was found in the previous one, and the opacity behind it is already on its way.
        // =====================================================================================================
        for (;;) {
            const unsigned wm = __ballot_sync(full, mode == FL_WALK);
            if (wm == 0u) break;
            if (__popc(wm) < walk_min) {
                // hand over to the event phase only if it has something to do
                const unsigned dm = __ballot_sync(full, mode == FL_DEAD || mode >= FL_EXITED);
                const unsigned em = __ballot_sync(full, mode == FL_EVENT) | ((!exhausted && __popc(dm) >= launch_min) ? dm : 0u);
                if (em) break;
            }
            walk_step<0, kAgg>(tx, ty, tz, tcur, tmin, taul, pend, rkc, idx, rx, ry, rz, ax, sn, rn, steps, mode, dtx, dty, dtz, sax, say, saz, rkb, slot0, jmb, 8 * ws);
            walk_step<1, kAgg>(tx, ty, tz, tcur, tmin, taul, pend, rkc, idx, rx, ry, rz, ax, sn, rn, steps, mode, dtx, dty, dtz, sax, say, saz, rkb, slot1, jmb, 8 * ws);
        }
        if (exhausted && __ballot_sync(full, mode != FL_DEAD) == 0u) break;
    }

    // fold the per-lane accumulators into the warp's counters, then into the global ones
    unsigned long long v[4] = {acc_pk, acc_steps, acc_scat, acc_abs};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unsigned long long s = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(full, s, o);
        if (lane == 0) wc[i] += s;
    }
    __syncwarp();
    if (lane < CNT_N && lane != CNT_WORK && wc[lane]) atomicAdd(cnt + lane, wc[lane]);
}

}  // namespace tamc
