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

struct WarpPool {
    double px[kPool], py[kPool], pz[kPool];           // position (shifted frame)
    double nz[kPool], st[kPool], cp[kPool], sp[kPool]; // cost, sint, cos(phi), sin(phi)
    double tau[kPool], pval[kPool];                   // optical depth to the next interaction; pending deposit
    int cells[kPool], cellk[kPool];                   // celli | cellj << 16 ; cellk
    int pidx[kPool];
    int steps[kPool], nscat[kPool], nbnd[kPool];
    unsigned int idlo[kPool], idhi[kPool];
    unsigned char wq[kPool], iq[kPool], fq[kPool];    // walk / interact / free queues (stacks of slot numbers)
    unsigned char pad[64];
    unsigned long long cnt[CNT_N];
};

// kFresnel compiles the boundary-optics extension in (TAMC_FRESNEL); the default build carries none of it.
template <int kBlock, int kMinCtas, bool kFresnel>
__global__ void __launch_bounds__(kBlock, kMinCtas) k_transport_pool(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                                                  int chunk, int scatter_min,
                                                                  unsigned long long *__restrict__ cnt)
{
    extern __shared__ double s_faces[];
    const double *xf, *yf, *zf;
    stage_faces(g, s_faces, xf, yf, zf);
    const int nfaces = g.nxg + g.nyg + g.nzg + 3;
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
    int steps = 0, nb = 0;
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
                const int ns = P.nscat[s];
                const uint4 r = philox4x32_10(key, make_uint4(P.idlo[s], P.idhi[s], (uint32_t)ns + 1u, 0u));
                if (u32_to_unit(r.x) < g.albedo) {            // draw < albedo ? stokes : absorbed
                    FastPhoton q;
                    q.nzp = P.nz[s]; q.sint = P.st[s]; q.cosp = P.cp[s]; q.sinp = P.sp[s];
                    q.nxp = q.sint * q.cosp; q.nyp = q.sint * q.sinp;
                    // the position stays parked in the slot; the walker re-derives reciprocals and indices
                    scatter_dir<false>(g, q, u32_to_unit(r.y), u32_to_unit(r.z));
                    P.nz[s] = q.nzp; P.st[s] = q.sint; P.cp[s] = q.cosp; P.sp[s] = q.sinp;
                    P.tau[s] = -log(u32_to_unit(r.w));
                    P.nscat[s] = ns + 1;
                    survive = true;
                } else {
                    absorbed = true;
                    const double pv = P.pval[s];
                    if (pv != 0.) atomicAdd(g.jmean + P.pidx[s], pv);
                    acc_steps += (unsigned long long)P.steps[s];
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
                    PhiloxRng lr;
                    lr.seed(seed, gid);
                    double u[4];
                    lr.block(u);
                    const Launched L = launch_fast(g, u, true);
                    P.px[s] = L.xcur; P.py[s] = L.ycur; P.pz[s] = lc.zcur0;
                    P.nz[s] = -1.; P.st[s] = 0.; P.cp[s] = L.cosp; P.sp[s] = L.sinp;   // sourceph.f90:37-42
                    P.tau[s] = L.tau; P.pval[s] = 0.; P.pidx[s] = -1;
                    P.cells[s] = L.cells; P.cellk[s] = lc.cellk0;
                    P.steps[s] = 0; P.nscat[s] = 0; P.nbnd[s] = 0;
                    P.idlo[s] = (uint32_t)gid; P.idhi[s] = (uint32_t)(gid >> 32);
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
                p.xcur = P.px[slot]; p.ycur = P.py[slot]; p.zcur = P.pz[slot];
                p.nzp = P.nz[slot];
                const double st = P.st[slot];
                p.nxp = st * P.cp[slot]; p.nyp = st * P.sp[slot];                      // stokes.f90:143-148
                if (kFresnel) { p.sint = st; p.cosp = P.cp[slot]; p.sinp = P.sp[slot]; }
                set_direction(p);                                                     // reciprocals + sign flags
                p.tau = P.tau[slot]; p.taurun = 0.;
                const int c = P.cells[slot];
                p.celli = c & 0xffff; p.cellj = c >> 16; p.cellk = P.cellk[slot];
                p.ridx = p.celli + g.sx * (p.cellj + (g.nyg + 2) * p.cellk);
                p.jidx = (p.celli - 1) + g.nxg * ((p.cellj - 1) + g.nyg * (p.cellk - 1));
                tally.pidx = P.pidx[slot]; tally.pval = P.pval[slot];
                steps = P.steps[slot];
                if (kFresnel) nb = P.nbnd[slot];
                walking = true;
                if (fresnel && nb == 0 && boundary_draw(key, P.idlo[slot], P.idhi[slot], nb) < g.r0sq) {
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
            int r = voxel_step_fast<true>(g, xf, yf, zf, p, tally);
            ++steps;
            if (fresnel && r == STEP_EXIT && fresnel_reflect_fast(g, xf, yf, zf, p, key, P.idlo[slot], P.idhi[slot], nb)) {
                atomicAdd(&P.cnt[CNT_REFLECT], 1ull);
                r = STEP_WALL;
            }
            if (r == STEP_INTERACT) {
                // the centred-position round trip of inttau2.f90:65-67 / :24-26
                P.px[slot] = (p.xcur - g.xmax) + g.xmax;
                P.py[slot] = (p.ycur - g.ymax) + g.ymax;
                P.pz[slot] = (p.zcur - g.zmax) + g.zmax;
                P.cells[slot] = p.celli | (p.cellj << 16); P.cellk[slot] = p.cellk;
                P.pidx[slot] = tally.pidx; P.pval[slot] = tally.pval;
                P.steps[slot] = steps;
                if (kFresnel) {
                    P.nbnd[slot] = nb;            // a reflection may have turned the packet around since it was adopted
                    P.nz[slot] = p.nzp; P.cp[slot] = p.cosp; P.sp[slot] = p.sinp;
                }
                park = true;
                walking = false;
            } else if (r == STEP_EXIT || steps >= kMaxStepsPerPacket) {
                tally.flush();
                acc_steps += (unsigned long long)steps;
                acc_scat += (unsigned long long)P.nscat[slot];
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
