// tamc_kernels.cu -- production transport kernels (Philox streams) for sm_100a.
//
// Replaces the per-rank `do j = 1, nphotons` of /root/reference/src/mcpolar.f90:151-170.  Two
// execution shapes over the same transport code (tamc_transport.cuh):
//   variant 0  one packet per thread, grid-stride over packet ids (also the records path);
//   variant 1  persistent warps: each warp owns a contiguous id range and refills idle lanes in
//              place, so lanes whose packet died early do not wait for the longest walk in the warp;
//              walking and scattering are phased so the divergent scattering code runs for many
//              lanes at once.
// The path is a random walk over an fp64 grid with fp64 atomics: no dense contraction, no tensor
// cores.
#include "tamc_internal.h"

namespace tamc {

// ---------------------------------------------------------------------------------------------
// variant 0
// ---------------------------------------------------------------------------------------------
template <class Tally, bool kRecord>
__global__ void __launch_bounds__(256) k_transport_simple(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                                          unsigned long long *__restrict__ cnt,
                                                          tamc_packet_record *__restrict__ rec)
{
    extern __shared__ double s_faces[];
    const double *xf, *yf, *zf;
    stage_faces(g, s_faces, xf, yf, zf);

    Counters c;
    c.clear();
    Tally tally;
    tally.jm = g.jmean;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        PhiloxRng rng;
        rng.seed(seed, first_id + (uint64_t)i);
        transport_packet<PhiloxRng, Tally, kRecord>(g, xf, yf, zf, rng, tally, c, kRecord ? rec + i : nullptr,
                                                    0x7fffffffffffffffll);
    }
    c.commit(cnt);
}

// ---------------------------------------------------------------------------------------------
// variant 1: persistent warps with in-place refill
// ---------------------------------------------------------------------------------------------
enum { LANE_IDLE = 0, LANE_WALK = 1, LANE_INTERACT = 2 };

template <class Tally>
__global__ void __launch_bounds__(256) k_transport_persistent(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                                              int refill_min, int scatter_min,
                                                              unsigned long long *__restrict__ cnt)
{
    extern __shared__ double s_faces[];
    const double *xf, *yf, *zf;
    stage_faces(g, s_faces, xf, yf, zf);

    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    // contiguous id range of this warp (n < 2^47, nwarps < 2^16: no overflow)
    long long next = (n / nwarps) * warp + min(n % nwarps, warp);
    const long long end = next + n / nwarps + (warp < n % nwarps ? 1 : 0);
    const bool scatter_on = (g.flags & TAMC_SCATTER) != 0;

    Counters c;
    c.clear();
    Tally tally;
    tally.jm = g.jmean;
    tally.begin();
    Photon p;
    PhiloxRng rng;
    int mode = LANE_IDLE, steps = 0, nscatt = 0;
    double u[4];

    for (;;) {
        const unsigned idle = __ballot_sync(full, mode == LANE_IDLE);
        const int nidle = __popc(idle);
        // ---- refill idle lanes from the warp's id range
        if (next < end && (nidle >= refill_min || nidle == 32)) {
            const long long avail = end - next;
            const int rank = __popc(idle & lt_mask);
            if (mode == LANE_IDLE && rank < avail) {
                rng.seed(seed, first_id + (uint64_t)(next + rank));
                rng.block(u);
                launch(g, p, u);
                tally.begin();
                steps = 0;
                nscatt = 0;
                mode = LANE_WALK;
            }
            next += min((long long)nidle, avail);
        }
        // ---- scattering phase for the lanes waiting at an interaction site
        const unsigned waiting = __ballot_sync(full, mode == LANE_INTERACT);
        const unsigned walking = __ballot_sync(full, mode == LANE_WALK);
        if (waiting && (__popc(waiting) >= scatter_min || walking == 0u)) {
            if (mode == LANE_INTERACT) {
                rng.block(u);
                if (u[0] < g.albedo) {
                    stokes(g, p, u[1], u[2]);
                    ++nscatt;
                    recentre(g, p);
                    p.taurun = 0.;
                    p.tau = -log(u[3]);
                    mode = LANE_WALK;
                } else {
                    tally.flush();
                    c.steps += (unsigned long long)steps;
                    c.scatters += (unsigned long long)nscatt;
                    c.fate(0);
                    mode = LANE_IDLE;
                }
            }
        } else if (walking == 0u && waiting == 0u && next >= end) {
            break;   // nothing in flight and the id range is exhausted
        }
        // ---- one voxel-step for every walking lane
        if (mode == LANE_WALK) {
            const int r = voxel_step(g, xf, yf, zf, p, tally);
            ++steps;
            if (r == STEP_INTERACT && scatter_on) {
                mode = LANE_INTERACT;
            } else if (r != STEP_WALL || steps >= kMaxStepsPerPacket) {
                tally.flush();
                c.steps += (unsigned long long)steps;
                c.scatters += (unsigned long long)nscatt;
                c.errors += (r == STEP_WALL);
                c.fate(r == STEP_EXIT ? exit_face(p) : 0);
                mode = LANE_IDLE;
            }
        }
    }
    c.commit(cnt);
}

// ---------------------------------------------------------------------------------------------
// roofline probe: the tally / grid address stream of straight-down packets and nothing else.
// Per packet: a column under the beam disk (fp32 rejection sampling), then voxel after voxel from
// the top face, one fp64 load of rhokap and one fp64 RED into jmean per voxel, continuing with
// probability exp(-rhokap*dz) decided by comparing raw Philox words against a threshold.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_probe(const DevGrid g, long long n, uint64_t seed, float disk_r_vox,
                                               unsigned long long *__restrict__ cnt)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long steps = 0;
    const float cx = 0.5f * g.nxg, cy = 0.5f * g.nyg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
        uint32_t blk = 0;
        uint4 r = philox4x32_10(key, make_uint4((uint32_t)i, (uint32_t)(i >> 32), blk++, 1u));
        float x = (r.x * 2.3283064e-10f) * 2.f - 1.f, y = (r.y * 2.3283064e-10f) * 2.f - 1.f;
        if (x * x + y * y > 1.f) { x *= 0.70710678f; y *= 0.70710678f; }
        int ci = min(g.nxg, max(1, (int)(cx + x * disk_r_vox) + 1));
        int cj = min(g.nyg, max(1, (int)(cy + y * disk_r_vox) + 1));
        int ck = g.nzg;
        uint32_t w = r.z;
        int used = 3;
        for (;;) {
            const double rk = __ldg(g.rhokap + ((long long)ci + (long long)g.sx * cj + g.sxy * ck));
            const long long jidx = (long long)(ci - 1) + (long long)g.nxg * ((long long)(cj - 1) + (long long)g.nyg * (ck - 1));
            atomicAdd(g.jmean + jidx, rk);
            ++steps;
            // continue with probability exp(-rk*dz); dz = 2 zmax / nzg
            const float pc = __expf(-(float)rk * (float)(2. * g.zmax / g.nzg));
            if (w * 2.3283064e-10f >= pc || --ck < 1) break;
            if (used == 3) { w = r.w; used = 4; }
            else { r = philox4x32_10(key, make_uint4((uint32_t)i, (uint32_t)(i >> 32), blk++, 1u)); w = r.x; used = 1; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(cnt + CNT_STEPS, steps);
}

__global__ void __launch_bounds__(256) k_fill(double *__restrict__ p, size_t n, double v)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
static inline size_t faces_bytes(const DevGrid &g) { return sizeof(double) * (size_t)(g.nxg + g.nyg + g.nzg + 3); }

template <class K>
static long long resident_ctas(K kernel, const LaunchCfg &cfg, size_t smem)
{
    int per_sm = cfg.ctas_per_sm;
    if (per_sm <= 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, cfg.block, smem) != cudaSuccess || per_sm < 1) {
            cudaGetLastError();
            per_sm = 1;
        }
    }
    return (long long)per_sm * cfg.num_sms;
}

template <class K, class... Args>
static cudaError_t launch_sized(K kernel, const LaunchCfg &cfg, size_t smem, long long n, cudaStream_t s, Args... args)
{
    // grid = min(resident CTAs on the whole chip, CTAs needed to give every thread one packet)
    const long long resident = resident_ctas(kernel, cfg, smem);
    const long long want = (n + cfg.block - 1) / cfg.block;
    const int grid = (int)(want < resident ? want : resident);
    kernel<<<grid, cfg.block, smem, s>>>(args...);
    return cudaGetLastError();
}

cudaError_t launch_transport(const DevGrid &g, const LaunchCfg &cfg, long long n, uint64_t seed, uint64_t first_id,
                             unsigned long long *d_cnt, tamc_packet_record *d_rec, cudaStream_t s, int *launches)
{
    if (n <= 0) return cudaSuccess;
    const bool merge = cfg.merge < 0 ? (g.flags & TAMC_SCATTER) != 0 : cfg.merge != 0;
    const size_t smem = faces_bytes(g);
    if (launches) *launches += 1;

    if (d_rec) {
        if (merge) return launch_sized(k_transport_simple<MergeTally, true>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, d_rec);
        return launch_sized(k_transport_simple<DirectTally, true>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, d_rec);
    }
    if (cfg.variant == 0) {
        tamc_packet_record *none = nullptr;
        if (merge) return launch_sized(k_transport_simple<MergeTally, false>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, none);
        return launch_sized(k_transport_simple<DirectTally, false>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, none);
    }
    if (merge)
        return launch_sized(k_transport_persistent<MergeTally>, cfg, smem, n, s, g, n, seed, first_id, cfg.refill_min, cfg.scatter_min, d_cnt);
    return launch_sized(k_transport_persistent<DirectTally>, cfg, smem, n, s, g, n, seed, first_id, cfg.refill_min, cfg.scatter_min, d_cnt);
}

cudaError_t launch_probe(const DevGrid &g, const LaunchCfg &cfg, long long n, uint64_t seed, unsigned long long *d_cnt,
                         cudaStream_t s)
{
    const float disk_r_vox = (float)(sqrt(g.spot_r2) * g.inv_dx);
    return launch_sized(k_probe, cfg, 0, n, s, g, n, seed, disk_r_vox, d_cnt);
}

cudaError_t launch_fill(double *p, size_t n, double v, int num_sms, cudaStream_t s)
{
    k_fill<<<num_sms * 8, 256, 0, s>>>(p, n, v);
    return cudaGetLastError();
}

}  // namespace tamc
