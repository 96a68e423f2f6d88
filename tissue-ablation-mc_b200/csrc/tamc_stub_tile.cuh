// tamc_stub_tile.cuh -- shipped (stub) regime with the hot top of the tally privatised in shared memory.
//
// In the shipped regime (mcpolar.f90:166-169: a packet ends at its first interaction) every flight is
// straight down from the beam disk, so the deposits concentrate in the top few planes under the disk's
// bounding box: plane k (from the top) takes exp(-(k-1) tau_c) (1 - exp(-tau_c)) of them.  The plain
// persistent kernel sits on the L1TEX address throughput shared by the rhokap loads and the jmean REDs
// (profiles/); here one CTA per SM (1024 threads) keeps those planes in a shared-memory tile and flushes
// it once, which removes most of the global REDs.  Same arithmetic and ids as variant 1 (kScatter=false).
#pragma once

#include "tamc_fast.cuh"

namespace tamc {

struct MiniReservoir {          // launched packets parked per warp (24 B each; indices are re-derived on adoption)
    double xcur[32], ycur[32], tau[32];
};

__global__ void __launch_bounds__(1024, 1) k_transport_stub_tiled(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                                                  int chunk, const TileGeom tg,
                                                                  unsigned long long *__restrict__ cnt)
{
    extern __shared__ double s_faces[];
    const double *xf, *yf, *zf;
    stage_faces(g, s_faces, xf, yf, zf);
    const int nfaces = g.nxg + g.nyg + g.nzg + 3;
    double *tile = s_faces + nfaces;
    const int tile_elems = tg.tw * tg.th * tg.layers;
    MiniReservoir &R = reinterpret_cast<MiniReservoir *>(tile + tile_elems)[threadIdx.x >> 5];
    for (int e = threadIdx.x; e < tile_elems; e += blockDim.x) tile[e] = 0.;
    __syncthreads();

    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const LaunchConsts lc{g.zcur0, g.cellk0};

    Counters c;
    c.clear();
    TiledTally32 tally;
    tally.jm = g.jmean;
    tally.tile = tile;
    tally.t = tg;
    tally.nzg = g.nzg;
    FastPhoton p;
    bool walking = false;
    int steps = 0;
    int count = 0;
    long long next = 0, end = 0;
    bool exhausted = false;

    for (;;) {
        const unsigned idle = __ballot_sync(full, !walking);
        const int nidle = __popc(idle);
        // ---- reservoir empty and lanes idle: launch 32 packets with every lane active
        if (nidle > 0 && count == 0 && !exhausted) {
            if (next >= end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(cnt + CNT_WORK, (unsigned long long)chunk);
                base = __shfl_sync(full, base, 0);
                if ((long long)base >= n) exhausted = true;
                else { next = (long long)base; end = min(next + chunk, n); }
            }
            if (!exhausted) {
                const long long id = next + lane;
                if (id < end) {
                    const uint64_t gid = first_id + (uint64_t)id;
                    const Launched L = launch_fast(g, philox_block(g, (uint32_t)gid, (uint32_t)(gid >> 32), 0u), false);
                    R.xcur[lane] = L.xcur; R.ycur[lane] = L.ycur; R.tau[lane] = L.tau;
                }
                count = (int)min((long long)32, end - next);
                next += count;
                __syncwarp();
            }
        }
        // ---- idle lanes adopt parked packets
        if (nidle && count) {
            const int rank = __popc(idle & lt_mask);
            if (!walking && rank < count) {
                const int s = count - 1 - rank;
                Launched L;
                L.xcur = R.xcur[s]; L.ycur = R.ycur[s]; L.tau = R.tau[s];
                L.cosp = 1.; L.sinp = 0.;
                const int celli = (int)(L.xcur * g.inv_dx) + 1, cellj = (int)(L.ycur * g.inv_dy) + 1;   // as launch_fast
                L.cells = celli | (cellj << 16);
                L.ridx = celli + g.sx * (cellj + (g.nyg + 2) * g.cellk0);
                L.jidx = (celli - 1) + g.nxg * ((cellj - 1) + g.nyg * (g.cellk0 - 1));
                adopt(g, lc, p, L);
                steps = 0;
                walking = true;
            }
            count -= min(nidle, count);
            __syncwarp();
        } else if (nidle == 32 && count == 0 && exhausted) {
            break;
        }
        // ---- one voxel-step for every walking lane
        if (walking) {
            const int r = voxel_step_fast<false>(g, xf, yf, zf, p, tally);
            ++steps;
            if (r != STEP_WALL || steps >= kMaxStepsPerPacket) {
                c.death(r == STEP_EXIT ? exit_face_fast(p, g) : 0, steps, 0, r == STEP_WALL);
                walking = false;
            }
        }
    }
    c.commit(cnt);

    // ---- flush the tile: one RED per non-zero entry
    __syncthreads();
    for (int e = threadIdx.x; e < tile_elems; e += blockDim.x) {
        const double v = tile[e];
        if (v != 0.) {
            const int di = e % tg.tw, r2 = e / tg.tw, dj = r2 % tg.th, dk = r2 / tg.th;
            const int i = tg.i0 + di, j = tg.j0 + dj, k = g.nzg - dk;
            atomicAdd(g.jmean + ((i - 1) + g.nxg * ((j - 1) + g.nyg * (k - 1))), v);
        }
    }
}

}  // namespace tamc
