// tamc_pool.cuh -- variant 3: the scattering transport regrouped through per-warp work queues.
//
// In the scatter loop (mcpolar.f90:165-169 shell + stokes.f90) a packet alternates
// between two very different pieces of code: a short voxel-step repeated a random number of times,
// and one long scattering event (Philox block, HG cosine, rotation, log).  With one packet per lane
// the lanes of a warp are out of phase, so either piece runs with half the warp masked off.
//
// Here every warp owns a pool of 64 packets in shared memory and two queues of slot numbers:
//   walk queue      packets ready to walk (fresh launches and packets that just scattered)
//   interact queue  packets stopped at an interaction site
// A lane walks one packet held in registers.  When it reaches an interaction it parks the packet in
// its slot, pushes the slot on the interact queue and immediately pops another packet from the walk
// queue, so the voxel-step keeps (nearly) all 32 lanes busy.  When 32 packets wait in the interact
// queue the warp runs ONE scattering pass with all 32 lanes active and pushes the survivors back on
// the walk queue.  With 64 packets per warp the two queues work as a double buffer: the walk queue
// runs dry exactly when the interact queue holds a full warp's worth.
// Launches fill free slots 32 at a time, again with all lanes active.
//
// Same production arithmetic as variants 0/1 (tamc_fast.cuh); the result depends on the schedule
// only through the fp64 summation order of the tally.
#pragma once

#include "tamc_fast.cuh"

namespace tamc {

constexpr int kPool = 64;

// One parked packet: seven 16-byte groups, each moved with a single 128-bit shared-memory access.  With the
// lanes of a warp touching unrelated slots, 16-byte accesses keep the bank conflicts of the pool traffic
// close to the bandwidth minimum (the slot stride of 28 words spreads consecutive slots over the bank groups).
struct alignas(16) PoolSlot {
    double2 pos_xy;                          // g0: position (shifted frame)
    double2 pz_pval;                         // g1: z position; pending deposit
    double2 nz_st;                           // g2: cost, sint
    double2 cp_sp;                           // g3: cos(phi), sin(phi)
    int4 tau_ns_nb;                          // g4: tau (lo, hi), scatter count, boundary draws used
    int4 pidx_steps_cells;                   // g5: pending-deposit voxel, voxel-steps, celli | cellj << 16, cellk
    uint4 id;                                // g6: packet id (lo, hi)
};
static_assert(sizeof(PoolSlot) == 112, "seven 16-byte groups");

struct alignas(16) WarpPool {
    PoolSlot slot[kPool];
    unsigned char wq[kPool], iq[kPool], fq[kPool];    // walk / interact / free queues (stacks of slot numbers)
    unsigned char pad[64];
    unsigned long long cnt[CNT_N];
};

// kFresnel is the `ext` build: it compiles in everything outside the shipped path -- boundary optics (TAMC_FRESNEL),
// periodic lateral boundaries (TAMC_PERIODIC), the Gaussian beam -- selected at run time; the default build carries none of it.
template <int kBlock, int kMinCtas, bool kFresnel, bool kAhead = false>
__global__ void __launch_bounds__(kBlock, kMinCtas) k_transport_pool(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                                                  int chunk, int scatter_min,
                                                                  unsigned long long *__restrict__ cnt)
{
    // kAhead: opacity of the next voxel fetched one loop pass early (voxel_step_fast).  Never with kFresnel: a reflection
    // moves the packet back into the grid after the step.
    static_assert(!(kAhead && kFresnel), "kAhead and kFresnel are exclusive");

    extern __shared__ double s_faces[];
    const double *xf, *yf, *zf;
    stage_faces(g, s_faces, xf, yf, zf);
    const int nfaces = (g.nxg + g.nyg + g.nzg + 3 + 1) & ~1;           // keep the pools 16-byte aligned
    WarpPool &P = reinterpret_cast<WarpPool *>(s_faces + nfaces)[threadIdx.x >> 5];

    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const LaunchConsts lc{g.zcur0, g.cellk0};
    const int launch_min = 16;

    // queues: everything free at the start
    P.fq[lane] = (unsigned char)lane;
    P.fq[lane + 32] = (unsigned char)(lane + 32);
    if (lane < CNT_N) P.cnt[lane] = 0ull;
    __syncwarp();
    int nw = 0, ni = 0, nf = kPool;          // queue depths (warp-uniform)
    long long next = 0, end = 0;             // ids of the chunk this warp owns
    bool exhausted = false;

    // the packet this lane is walking
    bool walking = false;
    int slot = 0;
    FastPhoton p;
    MergeTally32 tally;
    tally.jm = g.jmean;
    tally.begin();
    int steps = 0, nb = 0, nscat_reg = 0;
    bool doa = false;                        // dead on arrival: specularly reflected at the surface (TAMC_FRESNEL)
    constexpr bool fresnel = kFresnel;
    // per-lane accumulators, folded into the warp's counters at the end
    unsigned long long acc_steps = 0ull, acc_scat = 0ull;
    unsigned int acc_pk = 0u, acc_abs = 0u;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));

    for (;;) {
        const unsigned wmask = __ballot_sync(full, walking);
        const int nwalking = __popc(wmask);
        const int nempty = 32 - nwalking;

        // ---- (1) scattering pass: a full warp's worth waits, or the pool is draining
        if (ni >= 32 || (ni > 0 && nw == 0 && (ni >= scatter_min || nwalking == 0))) {
            const int k = min(ni, 32);
            int s = -1;
            bool survive = false, absorbed = false;
            if (lane < k) {
                s = P.iq[ni - 1 - lane];
                PoolSlot &S = P.slot[s];
                int4 g4 = S.tau_ns_nb;
                const int ns = g4.z;
                const uint4 id = S.id;
                const uint4 r = philox_block(g, id.x, id.y, (uint32_t)ns + 1u);
                if (unit_fast(r.x) < g.albedo) {            // draw < albedo ? stokes : absorbed
                    FastPhoton q;
                    const double2 a = S.nz_st, b = S.cp_sp;
                    q.nzp = a.x; q.sint = a.y; q.cosp = b.x; q.sinp = b.y;
                    q.nxp = q.sint * q.cosp; q.nyp = q.sint * q.sinp;
                    // the position stays parked in the slot; the walker re-derives reciprocals and indices
                    scatter_dir<false>(g, q, unit_fast(r.y), unit_fast(r.z));
                    S.nz_st = make_double2(q.nzp, q.sint);
                    S.cp_sp = make_double2(q.cosp, q.sinp);
                    const double tau = fm::neglog_u32(r.w);
                    g4.x = __double2loint(tau); g4.y = __double2hiint(tau); g4.z = ns + 1;
                    S.tau_ns_nb = g4;
                    survive = true;
                } else {
                    absorbed = true;
                    const double pv = S.pz_pval.y;
                    const int4 g5 = S.pidx_steps_cells;
                    if (pv != 0.) atomicAdd(g.jmean + g5.x, pv);
                    acc_steps += (unsigned long long)g5.y;
                    acc_scat += (unsigned long long)ns;
                    acc_pk++;
                    acc_abs++;
                }
            }
            ni -= k;
            const unsigned sm = __ballot_sync(full, survive), am = __ballot_sync(full, absorbed);
            if (survive) P.wq[nw + __popc(sm & lt_mask)] = (unsigned char)s;
            if (absorbed) P.fq[nf + __popc(am & lt_mask)] = (unsigned char)s;
            nw += __popc(sm);
            nf += __popc(am);
            __syncwarp();
        }

        // ---- (2) launch into free slots when the walk queue cannot feed the empty lanes
        // (batched: at least `launch_min` free slots, unless lanes would otherwise starve)
        if (!exhausted && nf > 0 && (nf >= launch_min ? nw < nempty : (nw == 0 && ni < scatter_min && nempty > 0))) {
            if (next >= end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(cnt + CNT_WORK, (unsigned long long)chunk);
                base = __shfl_sync(full, base, 0);
                if ((long long)base >= n) exhausted = true;
                else { next = (long long)base; end = min(next + chunk, n); }
            }
            if (!exhausted) {
                const int k = (int)min((long long)min(nf, 32), end - next);
                if (lane < k) {
                    const int s = P.fq[nf - 1 - lane];
                    const uint64_t gid = first_id + (uint64_t)(next + lane);
                    const Launched L = kFresnel ? launch_any(g, (uint32_t)gid, (uint32_t)(gid >> 32), true)
                                                : launch_fast(g, philox_block(g, (uint32_t)gid, (uint32_t)(gid >> 32), 0u), true);
                    PoolSlot &S = P.slot[s];
                    S.pos_xy = make_double2(L.xcur, L.ycur);
                    S.pz_pval = make_double2(lc.zcur0, 0.);
                    S.nz_st = make_double2(-1., 0.);                                       // sourceph.f90:37-42
                    S.cp_sp = make_double2(L.cosp, L.sinp);
                    S.tau_ns_nb = make_int4(__double2loint(L.tau), __double2hiint(L.tau), 0, 0);
                    S.pidx_steps_cells = make_int4(-1, 0, L.cells, lc.cellk0);
                    S.id = make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), 0u, 0u);
                    P.wq[nw + lane] = (unsigned char)s;
                }
                nf -= k;
                nw += k;
                next += k;
                __syncwarp();
            }
        }

        // ---- (3) empty lanes take a packet from the walk queue
        if (nempty > 0 && nw > 0) {
            const int rank = __popc(~wmask & lt_mask);
            const int k = min(nempty, nw);
            if (!walking && rank < k) {
                slot = P.wq[nw - 1 - rank];
                const PoolSlot &S = P.slot[slot];
                const double2 g0 = S.pos_xy, g1 = S.pz_pval, g2 = S.nz_st, g3 = S.cp_sp;
                const int4 g4 = S.tau_ns_nb, g5 = S.pidx_steps_cells;
                p.xcur = g0.x; p.ycur = g0.y; p.zcur = g1.x;
                p.nzp = g2.x;
                p.nxp = g2.y * g3.x; p.nyp = g2.y * g3.y;                              // stokes.f90:143-148
                if (kFresnel) { p.sint = g2.y; p.cosp = g3.x; p.sinp = g3.y; }
                set_direction(p);                                                     // reciprocals + sign flags
                p.tau = __hiloint2double(g4.y, g4.x); p.taurun = 0.;
                nscat_reg = g4.z;
                p.celli = g5.z & 0xffff; p.cellj = g5.z >> 16; p.cellk = g5.w;
                p.ridx = p.celli + g.sx * (p.cellj + (g.nyg + 2) * p.cellk);
                p.jidx = (p.celli - 1) + g.nxg * ((p.cellj - 1) + g.nyg * (p.cellk - 1));
                if (kAhead) p.rk = __ldg(g.rhokap + p.ridx);
                tally.pidx = g5.x; tally.pval = g1.y;
                steps = g5.y;
                if (kFresnel) nb = g4.w;
                walking = true;
                if (fresnel && (g.flags & TAMC_FRESNEL) && nb == 0 && boundary_draw(key, S.id.x, S.id.y, nb) < specular_r0sq(g, p.ridx)) {
                    walking = false;                      // fresh packet reflected at the top surface: never enters
                    doa = true;
                }
            }
            nw -= k;
            __syncwarp();
        } else if (nwalking == 0 && nw == 0 && ni == 0 && exhausted) {
            break;
        }

        // ---- (4) one voxel-step for every walking lane; park or retire the packet when the flight ends
        bool park = false, retire = false;
        if (doa) {
            doa = false;
            retire = true;
            acc_pk++;
            atomicAdd(&P.cnt[CNT_EXIT0 + 5], 1ull);
            atomicAdd(&P.cnt[CNT_SPECULAR], 1ull);
        } else if (walking) {
            int r = voxel_step_fast<true, decltype(tally), kAhead>(g, xf, yf, zf, p, tally);
            ++steps;
            if (fresnel && r == STEP_EXIT) {
                const int b = boundary_fast(g, xf, yf, zf, p, key, P.slot[slot].id.x, P.slot[slot].id.y, nb);
                if (b == 1) atomicAdd(&P.cnt[CNT_REFLECT], 1ull);
                if (b) r = STEP_WALL;                 // reflected or re-entered: the flight goes on
            }
            if (r == STEP_INTERACT) {
                // the centred-position round trip of inttau2.f90:65-67 / :24-26
                PoolSlot &S = P.slot[slot];
                S.pos_xy = make_double2((p.xcur - g.xmax) + g.xmax, (p.ycur - g.ymax) + g.ymax);
                S.pz_pval = make_double2((p.zcur - g.zmax) + g.zmax, tally.pval);
                S.pidx_steps_cells = make_int4(tally.pidx, steps, p.celli | (p.cellj << 16), p.cellk);
                if (kFresnel) {                   // a reflection may have turned the packet around since it was adopted
                    S.tau_ns_nb = make_int4(0, 0, nscat_reg, nb);
                    S.nz_st = make_double2(p.nzp, p.sint);
                    S.cp_sp = make_double2(p.cosp, p.sinp);
                }
                park = true;
                walking = false;
            } else if (r == STEP_EXIT || steps >= kMaxStepsPerPacket) {
                tally.flush();
                acc_steps += (unsigned long long)steps;
                acc_scat += (unsigned long long)nscat_reg;
                acc_pk++;
                if (r == STEP_EXIT) atomicAdd(&P.cnt[CNT_EXIT0 + exit_face_fast(p, g) - 1], 1ull);
                else { atomicAdd(&P.cnt[CNT_ERRORS], 1ull); acc_abs++; }
                retire = true;
                walking = false;
            }
        }
        const unsigned pm = __ballot_sync(full, park), rm = __ballot_sync(full, retire);
        if (pm | rm) {
            if (park) P.iq[ni + __popc(pm & lt_mask)] = (unsigned char)slot;
            if (retire) P.fq[nf + __popc(rm & lt_mask)] = (unsigned char)slot;
            ni += __popc(pm);
            nf += __popc(rm);
            __syncwarp();
        }
    }

    // fold the per-lane accumulators into the warp's counters, then into the global ones
    unsigned long long v[4] = {acc_pk, acc_steps, acc_scat, acc_abs};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        unsigned long long x = v[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(full, x, o);
        if (lane == 0) P.cnt[i] += x;
    }
    __syncwarp();
    if (lane < CNT_N && lane != CNT_WORK && P.cnt[lane]) atomicAdd(cnt + lane, P.cnt[lane]);
}

}  // namespace tamc
