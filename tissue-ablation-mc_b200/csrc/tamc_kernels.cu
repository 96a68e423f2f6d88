// tamc_kernels.cu -- production transport kernels (Philox streams) for sm_100a.
//
// Replaces the per-rank `do j = 1, nphotons` of /root/reference/src/mcpolar.f90:151-170.
//   variant 1 (default)  persistent warps.  A warp claims chunks of packet ids from one global
//              counter, launches 32 packets at a time with all lanes active into a small shared-
//              memory reservoir, and every lane whose packet has ended adopts the next one from
//              the reservoir in place -- so neither the launch code nor the walk runs with a
//              mostly-empty warp.  Walking and scattering are phased: lanes that reached an
//              interaction site wait until `scatter_min` of them can run the scattering code
//              together.
//   variant 0  one packet per thread, grid-stride (also the per-packet records path).
//   variant 2  variant 0 on the statement-by-statement arithmetic of tamc_transport.cuh
//              (device-side cross-check of the production arithmetic in tamc_fast.cuh).
// The path is a random walk over an fp64 grid with fp64 atomics: no dense contraction, no tensor
// cores.
#include <type_traits>

#include "tamc_fast.cuh"
#include "tamc_internal.h"
#include "tamc_column.cuh"
#include "tamc_pool.cuh"
#include "tamc_flight.cuh"
#include "tamc_stub_tile.cuh"

namespace tamc {

// Adds the packet's deposit sum for the records path.
template <class Base>
struct SumTally : Base {
    double packet_sum;
    __device__ __forceinline__ void begin() { packet_sum = 0.; Base::begin(); }
    __device__ __forceinline__ void add(const FastPhoton &p, double v) { packet_sum += v; Base::add(p, v); }
};

// ---------------------------------------------------------------------------------------------
// variant 0: one packet per thread (production arithmetic), optional per-packet records
// ---------------------------------------------------------------------------------------------
template <class Tally, bool kRecord>
__global__ void __launch_bounds__(256) k_transport_simple(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                                          unsigned long long *__restrict__ cnt,
                                                          tamc_packet_record *__restrict__ rec)
{
    extern __shared__ double s_faces[];
    const double *xf, *yf, *zf;
    stage_faces(g, s_faces, xf, yf, zf);
    const bool scatter_on = (g.flags & TAMC_SCATTER) != 0;
    const bool fresnel = (g.flags & TAMC_FRESNEL) != 0;
    const bool bounds = (g.flags & (TAMC_FRESNEL | TAMC_PERIODIC)) != 0;
    const LaunchConsts lc{g.zcur0, g.cellk0};

    Counters c;
    c.clear();
    SumTally<Tally> tally;
    tally.jm = g.jmean;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        PhiloxRng rng;
        rng.seed(seed, first_id + (uint64_t)i);
        FastPhoton p;
        adopt(g, lc, p, launch_any(g, rng.id_lo, rng.id_hi, scatter_on));
        rng.blk = 1;                                              // block 0 went into the launch
        tally.begin();
        int steps = 0, nscatt = 0, fate = 0, ndraws = 4, nb = 0;
        bool specular = false;
        if (fresnel && boundary_draw(rng.key, rng.id_lo, rng.id_hi, nb) < specular_r0sq(g, p.ridx)) {
            specular = true;                                      // reflected at the top surface before entering
            fate = 6;
            ndraws = 3;
            p.nzp = 1.;
            p.cellk = g.nzg + 1;
            c.note(CNT_SPECULAR);
        }
        while (!specular) {
            int r = voxel_step_fast<true>(g, xf, yf, zf, p, tally);
            ++steps;
            if (bounds && r == STEP_EXIT) {
                const int b = boundary_fast(g, xf, yf, zf, p, rng.key, rng.id_lo, rng.id_hi, nb);
                if (b == 1) c.note(CNT_REFLECT);
                if (b) r = STEP_WALL;                             // reflected or re-entered: the flight goes on
            }
            if (r == STEP_WALL) {
                if (steps >= kMaxStepsPerPacket) { c.errors++; break; }
                continue;
            }
            if (r == STEP_EXIT) {
                fate = exit_face_fast(p, g);
                break;
            }
            if (!scatter_on) break;                               // mcpolar.f90:166-169 stub
            const uint4 w = philox_block(g, rng);
            double albedo, hgg;
            ScatterConsts sc;
            voxel_optics(g, p.ridx, albedo, hgg, sc);             // scalars, or the voxel's own (tamc_set_optics_grids)
            if (unit_fast(w.x) < albedo) {
                scatter_fast(g, p, unit_fast(w.y), unit_fast(w.z), fm::neglog_u32(w.w), hgg, sc);
                ++nscatt;
                ndraws += 4;
            } else {
                ndraws += 1;
                break;
            }
        }
        tally.flush();
        c.steps += (unsigned long long)steps;
        c.scatters += (unsigned long long)nscatt;
        c.fate(fate);
        if (kRecord) {
            tamc_packet_record *r = rec + i;
            r->xp = p.xcur - g.xmax; r->yp = p.ycur - g.ymax; r->zp = p.zcur - g.zmax;
            r->nxp = p.nxp; r->nyp = p.nyp; r->nzp = p.nzp;
            r->deposit = tally.packet_sum;
            r->xcell = (p.celli < 1 || p.celli > g.nxg) ? -1 : p.celli;
            r->ycell = (p.cellj < 1 || p.cellj > g.nyg) ? -1 : p.cellj;
            r->zcell = (p.cellk < 1 || p.cellk > g.nzg) ? -1 : p.cellk;
            r->steps = steps; r->nscatt = nscatt; r->ndraws = ndraws; r->fate = fate; r->flags = 0;
        }
    }
    c.commit(cnt);
}

// ---------------------------------------------------------------------------------------------
// variant 2: one packet per thread on the exact (replay) arithmetic
// ---------------------------------------------------------------------------------------------
template <class Tally, bool kRecord>
__global__ void __launch_bounds__(256) k_transport_exact(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
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
// variant 1: persistent warps, chunked ids, launch reservoir, phased scattering
// ---------------------------------------------------------------------------------------------
enum { LANE_IDLE = 0, LANE_WALK = 1, LANE_INTERACT = 2 };
constexpr int kResv = 64;   // reservoir slots per warp: < 32 left over + one generation of 32

struct WarpReservoir {
    double xcur[kResv], ycur[kResv], tau[kResv], cosp[kResv], sinp[kResv];
    unsigned int id_lo[kResv], id_hi[kResv];
    int cells[kResv], ridx[kResv], jidx[kResv];
};

// kScatter = false is the shipped regime (mcpolar.f90:166-169 stub): no scattering phase, no azimuth,
// no position after the final partial step -- the compiler drops that state and its registers.
// With scattering the per-packet state is large; the counters then live in shared memory (one set per
// warp, touched only when a packet ends) and kMinCtas selects the register budget the kernel is built for.
template <class Tally, bool kScatter, int kMinCtas>
__global__ void __launch_bounds__(256, kMinCtas) k_transport_persistent(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                                              int chunk, int scatter_min,
                                                              unsigned long long *__restrict__ cnt)
{
    extern __shared__ double s_faces[];
    const double *xf, *yf, *zf;
    stage_faces(g, s_faces, xf, yf, zf);
    const int nfaces = g.nxg + g.nyg + g.nzg + 3;
    WarpReservoir &R = reinterpret_cast<WarpReservoir *>(s_faces + nfaces)[threadIdx.x >> 5];
    unsigned long long *s_cnt = reinterpret_cast<unsigned long long *>(
        reinterpret_cast<WarpReservoir *>(s_faces + nfaces) + (blockDim.x >> 5)) + (threadIdx.x >> 5) * CNT_N;

    const unsigned full = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    const LaunchConsts lc{g.zcur0, g.cellk0};

    typename std::conditional<kScatter, WarpCounters, Counters>::type c;
    if constexpr (kScatter) c.w = s_cnt;
    c.clear();
    Tally tally;
    tally.jm = g.jmean;
    tally.begin();
    FastPhoton p;
    PhiloxRng rng;
    rng.key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    int mode = LANE_IDLE, steps = 0, nscatt = 0, nb = 0;
    const bool fresnel = kScatter && (g.flags & TAMC_FRESNEL) != 0;
    int count = 0;                       // launched packets parked in the reservoir (warp-uniform)
    long long next = 0, end = 0;         // ids of the chunk this warp currently owns
    bool exhausted = false;

    for (;;) {
        const unsigned idle = __ballot_sync(full, mode == LANE_IDLE);
        const int nidle = __popc(idle);
        // ---- more idle lanes than parked packets: launch another 32 with every lane active
        if (nidle > count && !exhausted) {
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
                    const Launched L = launch_fast(g, philox_block(g, (uint32_t)gid, (uint32_t)(gid >> 32), 0u), kScatter);
                    const int s = count + (int)lane;
                    R.xcur[s] = L.xcur; R.ycur[s] = L.ycur; R.tau[s] = L.tau;
                    if (kScatter) {
                        R.cosp[s] = L.cosp; R.sinp[s] = L.sinp;
                        R.id_lo[s] = (uint32_t)gid; R.id_hi[s] = (uint32_t)(gid >> 32);
                    }
                    R.cells[s] = L.cells; R.ridx[s] = L.ridx; R.jidx[s] = L.jidx;
                }
                const int ngen = (int)min((long long)32, end - next);
                count += ngen;
                next += ngen;
                __syncwarp();
            }
        }
        // ---- idle lanes adopt parked packets
        if (nidle && count) {
            const int rank = __popc(idle & lt_mask);
            if (mode == LANE_IDLE && rank < count) {
                const int s = count - 1 - rank;
                Launched L;
                L.xcur = R.xcur[s]; L.ycur = R.ycur[s]; L.tau = R.tau[s];
                L.cosp = 1.; L.sinp = 0.;
                if (kScatter) {
                    L.cosp = R.cosp[s]; L.sinp = R.sinp[s];
                    rng.id_lo = R.id_lo[s]; rng.id_hi = R.id_hi[s]; rng.blk = 1;   // block 0 went into the launch
                }
                L.cells = R.cells[s]; L.ridx = R.ridx[s]; L.jidx = R.jidx[s];
                adopt(g, lc, p, L);
                tally.begin();
                steps = 0;
                nscatt = 0;
                nb = 0;
                mode = LANE_WALK;
                if (fresnel && boundary_draw(rng.key, rng.id_lo, rng.id_hi, nb) < specular_r0sq(g, p.ridx)) {
                    c.note(CNT_SPECULAR);                 // reflected at the top surface: never enters
                    c.death(6, 0, 0, false);
                    mode = LANE_IDLE;
                }
            }
            count -= min(nidle, count);
            __syncwarp();
        }
        // ---- scattering phase for the lanes waiting at an interaction site
        const unsigned waiting = kScatter ? __ballot_sync(full, mode == LANE_INTERACT) : 0u;
        const unsigned walking = __ballot_sync(full, mode == LANE_WALK);
        if (kScatter && waiting && (__popc(waiting) >= scatter_min || walking == 0u)) {
            if (mode == LANE_INTERACT) {
                const uint4 w = philox_block(g, rng);
                if (unit_fast(w.x) < g.albedo) {          // draw < albedo ? stokes : absorbed
                    scatter_fast(g, p, unit_fast(w.y), unit_fast(w.z), fm::neglog_u32(w.w));
                    ++nscatt;
                    mode = LANE_WALK;
                } else {
                    tally.flush();
                    c.death(0, steps, nscatt, false);
                    mode = LANE_IDLE;
                }
            }
        } else if (walking == 0u && waiting == 0u && count == 0 && exhausted) {
            break;
        }
        // ---- one voxel-step for every walking lane
        if (mode == LANE_WALK) {
            int r = voxel_step_fast<kScatter>(g, xf, yf, zf, p, tally);
            ++steps;
            if (fresnel && r == STEP_EXIT && fresnel_reflect_fast(g, xf, yf, zf, p, rng.key, rng.id_lo, rng.id_hi, nb)) {
                c.note(CNT_REFLECT);
                r = STEP_WALL;
            }
            if (kScatter && r == STEP_INTERACT) {
                mode = LANE_INTERACT;
            } else if (r != STEP_WALL || steps >= kMaxStepsPerPacket) {
                tally.flush();
                c.death(r == STEP_EXIT ? exit_face_fast(p, g) : 0, steps, nscatt, r == STEP_WALL);
                mode = LANE_IDLE;
            }
        }
    }
    c.commit(cnt);
}

// ---------------------------------------------------------------------------------------------
// roofline probe: the grid / tally address stream of straight-down packets and as little else as
// possible -- the "L2-atomic / grid-lookup roofline" the transport kernel is compared with.
// Per packet: one 64-bit hash gives a column under the beam disk (fp32 polar sampling with fast
// intrinsics) and a geometric number of voxel-steps with continuation probability exp(-rhokap*dz)
// (one fp32 log); then voxel after voxel from the top face: one fp64 load of rhokap and one fp64
// RED into jmean per step, nothing more.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

__global__ void __launch_bounds__(256) k_probe(const DevGrid g, long long n, uint64_t seed, float disk_r_vox,
                                               unsigned long long *__restrict__ cnt)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long steps = 0;
    double keep = 0.;
    const float cx = 0.5f * g.nxg, cy = 0.5f * g.nyg;
    const double rk0 = __ldg(g.rhokap + ((long long)(g.nxg / 2) + (long long)g.sx * (g.nyg / 2) + g.sxy * g.nzg));
    const float inv_logp = -1.f / ((float)rk0 * (float)(2. * g.zmax / g.nzg));      // 1 / log(P(continue))
    const int plane_r = (int)g.sxy, plane_j = g.nxg * g.nyg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
        const float u0 = ((uint32_t)h & 0xffffffu) * 5.9604645e-8f, u1 = ((uint32_t)(h >> 24) & 0xffffffu) * 5.9604645e-8f;
        const float u2 = ((uint32_t)(h >> 40) + 0.5f) * 5.9604645e-8f;
        const float rr = disk_r_vox * __fsqrt_rn(u0);
        float sn, cs;
        __sincosf(6.2831853f * u1, &sn, &cs);
        const int ci = min(g.nxg, max(1, (int)(cx + rr * cs) + 1));
        const int cj = min(g.nyg, max(1, (int)(cy + rr * sn) + 1));
        int nst = 1 + (int)(__logf(u2) * inv_logp);                                   // geometric
        nst = min(nst, g.nzg);
        int ridx = ci + g.sx * (cj + (g.nyg + 2) * g.nzg);
        int jidx = (ci - 1) + g.nxg * ((cj - 1) + g.nyg * (g.nzg - 1));
        steps += (unsigned long long)nst;
        for (int k = 0; k < nst; ++k) {
            keep += __ldg(g.rhokap + ridx);
            atomicAdd(g.jmean + jidx, 1.0);
            ridx -= plane_r;
            jidx -= plane_j;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(cnt + CNT_STEPS, steps);
    if (keep == 123.456) cnt[CNT_N - 1] = 1ull;   // keeps the loads alive; never true for opacity sums
}

// The same for the column form (tamc_column.cuh): per packet one 256-bit load of the z-fastest opacity copy per
// group of four voxels crossed, one fp64 RED (the partial deposit) and, unless the packet stops in the top plane, one
// u32 RED (the stop count) -- and as little else as possible.
template <bool kTiled>
__global__ void __launch_bounds__(kTiled ? 1024 : 256) k_probe_column(const DevGrid g, long long n, uint64_t seed, float disk_r_vox, const ColGeom cg,
                                                      const double *__restrict__ rkT, unsigned int *__restrict__ stops,
                                                      unsigned long long *__restrict__ cnt, int ta, int tb)
{
    extern __shared__ double s_tiles[];
    const int cols = cg.tw * cg.th;
    double *s_dep = s_tiles;
    unsigned int *s_stop = reinterpret_cast<unsigned int *>(s_dep + (size_t)ta * cols);
    if (kTiled) {
        for (int i = threadIdx.x; i < ta * cols; i += blockDim.x) s_dep[i] = 0.;
        for (int i = threadIdx.x; i < tb * cols; i += blockDim.x) s_stop[i] = 0u;
        __syncthreads();
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long steps = 0;
    double keep = 0.;
    const float cx = 0.5f * g.nxg, cy = 0.5f * g.nyg;
    const double rk0 = __ldg(g.rhokap + ((long long)(g.nxg / 2) + (long long)g.sx * (g.nyg / 2) + g.sxy * g.nzg));
    const float inv_logp = -1.f / ((float)rk0 * (float)(2. * g.zmax / g.nzg));
    const int plane = g.nxg * g.nyg;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t h = mix64(seed + 0x9E3779B97F4A7C15ull * (uint64_t)(i + 1));
        const float u0 = ((uint32_t)h & 0xffffffu) * 5.9604645e-8f, u1 = ((uint32_t)(h >> 24) & 0xffffffu) * 5.9604645e-8f;
        const float u2 = ((uint32_t)(h >> 40) + 0.5f) * 5.9604645e-8f;
        const float rr = disk_r_vox * __fsqrt_rn(u0);
        float sn, cs;
        __sincosf(6.2831853f * u1, &sn, &cs);
        const int ci = min(cg.i0 + cg.tw - 1, max(cg.i0, (int)(cx + rr * cs) + 1));
        const int cj = min(cg.j0 + cg.th - 1, max(cg.j0, (int)(cy + rr * sn) + 1));
        int nst = 1 + (int)(__logf(u2) * inv_logp);                                   // geometric
        nst = min(nst, g.nzg);
        steps += (unsigned long long)nst;
        const int col_id = (cj - cg.j0) * cg.tw + (ci - cg.i0);
        const double *col = rkT + (size_t)col_id * cg.nzp;
        const int kstop = g.nzg - nst + 1;
        for (int gb = (g.nzg - 1) & ~3; gb >= ((kstop - 1) & ~3); gb -= 4) {
            double a, b, c, d;
            ldg256(col + gb, a, b, c, d);
            keep += a + d;
        }
        const int j = (ci - 1) + g.nxg * ((cj - 1) + g.nyg * (kstop - 1));
        const int d = nst - 1;
        if (kTiled && d < ta) smem_add_f64(s_dep + d * cols + col_id, 1.0);
        else atomicAdd(g.jmean + j, 1.0);
        if (kTiled && d >= 1 && d <= tb) atomicAdd(s_stop + (d - 1) * cols + col_id, 1u);
        else if (kstop < g.nzg) atomicAdd(stops + (j + plane), 1u);
    }
    if (kTiled) {
        __syncthreads();
        for (int i = threadIdx.x; i < ta * cols; i += blockDim.x) {
            const double v = s_dep[i];
            if (v != 0.) {
                const int d = i / cols, c = i - d * cols, dj = c / cg.tw, di = c - dj * cg.tw;
                atomicAdd(g.jmean + ((cg.i0 - 1 + di) + g.nxg * ((cg.j0 - 1 + dj) + g.nyg * (g.nzg - d - 1))), v);
            }
        }
        for (int i = threadIdx.x; i < tb * cols; i += blockDim.x) {
            const unsigned int v = s_stop[i];
            if (v) {
                const int d = i / cols + 1, c = i - (d - 1) * cols, dj = c / cg.tw, di = c - dj * cg.tw;
                atomicAdd(stops + ((cg.i0 - 1 + di) + g.nxg * ((cg.j0 - 1 + dj) + g.nyg * (g.nzg - d))), v);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) steps += __shfl_xor_sync(0xffffffffu, steps, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(cnt + CNT_STEPS, steps);
    if (keep == 123.456) cnt[CNT_N - 1] = 1ull;
}

// ---------------------------------------------------------------------------------------------
// roofline probe of the scatter regime: the REAL voxel-index stream of a sample of packets, recorded once
// (k_trace: production arithmetic, one packet per thread; pass 1 counts the voxel-steps per packet, pass 2 writes
// the jmean index of every step) and then replayed as memory operations only (k_probe_trace): per voxel-step one
// 8-byte load of the voxel record's opacity and, when the packet leaves the voxel, one fp64 RED into its tally --
// the access pattern of the flight kernel (tamc_flight.cuh) on the same interleaved layout, with no transport
// arithmetic.  The only extra traffic is the index stream itself, read 16 bytes (four steps) at a time.
// ---------------------------------------------------------------------------------------------
struct TraceTally {
    int *out;            // null in the counting pass
    long long pos;
    int count;
    __device__ __forceinline__ void begin() { count = 0; }
    __device__ __forceinline__ void add(const FastPhoton &p, double)
    {
        if (out) out[pos + count] = p.jidx;
        ++count;
    }
    __device__ __forceinline__ void flush() {}
};

__global__ void __launch_bounds__(256) k_trace(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                               const long long *__restrict__ off, int *__restrict__ trace, int *__restrict__ counts)
{
    extern __shared__ double s_faces[];
    const double *xf, *yf, *zf;
    stage_faces(g, s_faces, xf, yf, zf);
    const bool scatter_on = (g.flags & TAMC_SCATTER) != 0;
    const LaunchConsts lc{g.zcur0, g.cellk0};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        PhiloxRng rng;
        rng.seed(seed, first_id + (uint64_t)i);
        FastPhoton p;
        adopt(g, lc, p, launch_fast(g, philox_block(g, rng.id_lo, rng.id_hi, 0u), scatter_on));
        rng.blk = 1;
        TraceTally tally;
        tally.out = trace;
        tally.pos = trace ? off[i] : 0;
        tally.begin();
        for (;;) {
            const int r = voxel_step_fast<true>(g, xf, yf, zf, p, tally);
            if (r == STEP_WALL) continue;
            if (r == STEP_EXIT || !scatter_on) break;
            const uint4 w = philox_block(g, rng);
            if (!(unit_fast(w.x) < g.albedo)) break;
            scatter_fast(g, p, unit_fast(w.y), unit_fast(w.z), fm::neglog_u32(w.w));
        }
        if (!trace) counts[i] = tally.count;
        else for (int k = tally.count; k & 3; ++k) trace[tally.pos + k] = -1;       // pad to a multiple of four steps
    }
}

template <bool kInter>
__global__ void __launch_bounds__(256) k_probe_trace(const double *__restrict__ rkb, double *__restrict__ jmb, long long n,
                                                     const long long *__restrict__ off, const int *__restrict__ trace,
                                                     unsigned long long *__restrict__ cnt)
{
    constexpr int ws = kInter ? 2 : 1;            // doubles per voxel record: the flight kernel's two layouts (tamc_flight.cuh)
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    long long pos = 0, end = 0;           // this lane's slice of the index stream
    long long next = 0, last = 0;         // warp-uniform: packets this warp owns
    bool exhausted = false;
    unsigned long long steps = 0ull, reds = 0ull;
    double keep = 0.;
    int cur = -1;
    for (;;) {
        // lanes whose packet is finished take the next one (64 packets per claim)
        const unsigned idle = __ballot_sync(full, pos >= end);
        if (idle) {
            if (cur >= 0 && pos >= end) { atomicAdd(jmb + (size_t)cur * ws, 1.0); ++reds; cur = -1; }
            if (!exhausted) {
                if (next >= last) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(cnt + CNT_WORK, 64ull);
                    base = __shfl_sync(full, base, 0);
                    if ((long long)base >= n) exhausted = true;
                    else { next = (long long)base; last = min(next + 64, n); }
                }
                if (!exhausted) {
                    const int rank = __popc(idle & lt_mask);
                    if (pos >= end && next + rank < last) { pos = off[next + rank]; end = off[next + rank + 1]; }
                    next += min((long long)__popc(idle), last - next);
                }
            } else if (idle == full) {
                break;
            }
        }
        if (pos < end) {
            const int4 q = *reinterpret_cast<const int4 *>(trace + pos);            // four voxel-steps of this packet
            pos += 4;
            const int ix[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int idx = ix[k];
                if (idx < 0) continue;                                              // padding
                if (idx != cur) {                                                   // the packet left voxel `cur`
                    if (cur >= 0) { atomicAdd(jmb + (size_t)cur * ws, 1.0); ++reds; }
                    cur = idx;
                }
                keep += __ldg(rkb + (size_t)idx * ws);
                ++steps;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        steps += __shfl_xor_sync(full, steps, o);
        reds += __shfl_xor_sync(full, reds, o);
    }
    if (lane == 0) { atomicAdd(cnt + CNT_STEPS, steps); atomicAdd(cnt + CNT_SCATTERS, reds); }
    if (keep == 123.456) cnt[CNT_N - 1] = 1ull;   // keeps the loads alive
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

// grid = min(CTAs resident on the whole chip, CTAs needed to give every thread one packet)
template <class K, class... Args>
static cudaError_t launch_sized(K kernel, const LaunchCfg &cfg, size_t smem, long long n, cudaStream_t s, Args... args)
{
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const long long resident = resident_ctas(kernel, cfg, smem);
    const long long want = (n + cfg.block - 1) / cfg.block;
    const int grid = (int)(want < resident ? want : resident);
    kernel<<<grid, cfg.block, smem, s>>>(args...);
    return cudaGetLastError();
}

bool flight_interleaved(const LaunchCfg &cfg, size_t nvox);

// Column form of the shipped regime (tamc_column.cuh): gather the beam's columns, transport, add the full-crossing term.
// Bounding box of the launch voxels: launch_fast computes celli = int(xcur*inv_dx) + 1 with |xcur - xmax| <= R, and
// both the product and the truncation are monotonic, so the box below contains every launch voxel.
bool beam_box(const DevGrid &g, ColGeom &cg)
{
    if (g.gauss_sigma > 0.) return false;          // Gaussian beam: launch points anywhere on the top face
    const double R = sqrt(g.spot_r2);
    cg.i0 = max(1, (int)((g.xmax - R) * g.inv_dx) + 1);
    cg.j0 = max(1, (int)((g.ymax - R) * g.inv_dy) + 1);
    cg.tw = min(g.nxg, (int)((g.xmax + R) * g.inv_dx) + 1) - cg.i0 + 1;
    cg.th = min(g.nyg, (int)((g.ymax + R) * g.inv_dy) + 1) - cg.j0 + 1;
    cg.nzp = (g.nzg + 3) & ~3;
    cg.kz_lo = 0;
    cg.deep_sx = 0;
    cg.deep_sxy = 0;
    cg.deep = nullptr;
    return cg.tw >= 1 && cg.th >= 1;
}

cudaError_t launch_box_copy(const DevGrid &g, const ColGeom &cg, double *dense, bool unpack, int num_sms, cudaStream_t s, int kz0)
{
    if (unpack) k_box_copy<true><<<num_sms * 8, 256, 0, s>>>(g, cg, dense, kz0);
    else k_box_copy<false><<<num_sms * 8, 256, 0, s>>>(g, cg, dense, kz0);
    return cudaGetLastError();
}

cudaError_t launch_box_mirror(const DevGrid &g, const ColGeom &cg, double *dst, int num_sms, cudaStream_t s)
{
    k_box_mirror<<<num_sms * 8, 256, 0, s>>>(g, cg, dst);
    return cudaGetLastError();
}

cudaError_t launch_peer_box_reduce(const PeerSet &ps, double *out, size_t cnt, int nranks, int rank, unsigned long long call,
                                   unsigned int *err, int num_sms, cudaStream_t s)
{
    // a few MB: enough CTAs to keep (nranks - 1) NVLink reads per thread in flight on every SM, no more (each CTA polls the flags)
    const size_t pairs = (cnt + 1) >> 1;
    size_t want = (pairs + 255) / 256;
    const int grid = (int)(want < (size_t)num_sms * 4 ? (want ? want : 1) : (size_t)num_sms * 4);
    const unsigned long long timeout_ns = 20ull * 1000000000ull;      // a rank that never arrives: give up, report, do not hang the GPU
    if (nranks == 2) k_peer_box_reduce<2><<<grid, 256, 0, s>>>(ps, out, cnt, nranks, rank, call, timeout_ns, err);
    else if (nranks == 4) k_peer_box_reduce<4><<<grid, 256, 0, s>>>(ps, out, cnt, nranks, rank, call, timeout_ns, err);
    else if (nranks == 8) k_peer_box_reduce<8><<<grid, 256, 0, s>>>(ps, out, cnt, nranks, rank, call, timeout_ns, err);
    else k_peer_box_reduce<0><<<grid, 256, 0, s>>>(ps, out, cnt, nranks, rank, call, timeout_ns, err);
    return cudaGetLastError();
}

// the workspace of the column form (+ the z-fastest copy of the opacities under the box)
static cudaError_t column_setup(const DevGrid &g, ColumnWorkspace *ws, bool gather, cudaStream_t s, ColGeom &cg, int gather_planes = 0)
{
    if (!beam_box(g, cg)) return cudaErrorInvalidValue;
    const size_t nstops = (size_t)g.nxg * g.nyg * g.nzg;
    if (ws->stops_elems < nstops) {
        cudaFree(ws->stops);
        ws->stops = nullptr;
        ws->stops_elems = 0;
        cudaError_t e = cudaMalloc(&ws->stops, nstops * sizeof(unsigned int));
        if (e == cudaSuccess) e = cudaMemsetAsync(ws->stops, 0, nstops * sizeof(unsigned int), s);
        if (e != cudaSuccess) return e;
        ws->stops_elems = nstops;
    }
    if (gather) {
        const size_t nrk = (size_t)cg.tw * cg.th * cg.nzp;
        if (ws->rkT_elems < nrk) {
            cudaFree(ws->rkT);
            ws->rkT = nullptr;
            ws->rkT_elems = 0;
            cudaError_t e = cudaMalloc(&ws->rkT, nrk * sizeof(double));
            if (e != cudaSuccess) return e;
            ws->rkT_elems = nrk;
        }
        // columns-first upload over PCIe: only the planes the packets are expected to reach (gather_planes from the top
        // face, whole 32-plane tiles); anything deeper is read from the caller's grid where it is needed
        if (ws->gather_src && ws->box_rk && gather_planes > 0 && gather_planes < g.nzg && ws->share_gather == 0) {
            cg.kz_lo = (g.nzg - gather_planes) & ~31;
            cg.deep = ws->gather_src;
            cg.deep_sx = g.sx;
            cg.deep_sxy = g.sxy;
        }
        const dim3 gg((cg.tw + 31) / 32, cg.th, (cg.nzp - cg.kz_lo + 31) / 32);
        if (ws->share_gather == 2 && ws->box_rk) {
            // root_io, rank > 0: both copies of the columns arrive from rank 0 below
        } else if (ws->gather_src && ws->box_rk) {
            if (ws->ev_gather0) cudaEventRecord(ws->ev_gather0, s);
            k_column_gather<true><<<gg, 256, 0, s>>>(g, cg, ws->gather_src, ws->rkT, ws->box_rk);
            if (ws->ev_gather1) cudaEventRecord(ws->ev_gather1, s);
        } else {
            k_column_gather<false><<<gg, 256, 0, s>>>(g, cg, g.rhokap, ws->rkT, nullptr);
        }
        if (ws->share_gather && ws->box_rk && ws->share_fn) {
            cudaError_t e = ws->share_fn(ws->share_comm, ws->rkT, nrk, s);
            if (e == cudaSuccess) e = ws->share_fn(ws->share_comm, ws->box_rk, (size_t)cg.tw * cg.th * g.nzg, s);
            if (e != cudaSuccess) return e;
        }
    }
    return cudaGetLastError();
}

// Plan of the column form for a call of n packets: whether to use it, and the shared-memory tile split (tamc_column.cuh:
// ta planes of partial deposits, tb planes of stop counts per CTA).  Measured rules (profiles/README.md, tools/tile_sweep.py):
//   * wide beam (> 4096 columns): the global REDs are not contended; one plane of stop counts and one 1024-thread CTA
//     per SM is the fastest shape (deeper tiles only add flush work);
//   * narrow beam: the same few thousand tally addresses are hit so often that the L2 atomic unit serialises -- keep the
//     top plane of the deposits and two planes of counts in shared memory;
//   * tiles only when the call amortises their flush (148 CTAs x columns x planes REDs), the column form itself only for
//     calls that pay for its two small extra kernels (>= 2^20 packets).
struct ColumnPlan {
    bool use;
    int ta, tb;
};

static ColumnPlan column_plan(const DevGrid &g, const LaunchCfg &cfg, long long n)
{
    ColumnPlan p{false, 0, 0};
    ColGeom cg;
    if (cfg.variant != 3 || (g.flags & (TAMC_SCATTER | TAMC_FRESNEL)) || cfg.column == 0 || n <= 0 || !beam_box(g, cg)) return p;
    const long long cols = (long long)cg.tw * cg.th;
    const bool wide = cols > 4096;
    if (cfg.column != 2 && cfg.column_tile != 0) {
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        const long long avail = (long long)optin - 8ll * cg.nzp - 1024 - 32ll * (long long)sizeof(ParkQueue);
        int ta = wide ? 0 : 1, tb = wide ? 1 : 2;
        if (cfg.column_tile > 0) { ta = cfg.column_tile / 10; tb = cfg.column_tile % 10; }       // forced split
        ta = ta < g.nzg ? ta : g.nzg;
        tb = tb < g.nzg ? tb : g.nzg;
        const bool fits = cols * (8ll * ta + 4ll * tb) <= avail;
        const bool pays = n >= 8ll * cfg.num_sms * cols * (ta + tb);
        if (ta + tb > 0 && fits && (cfg.column_tile > 0 || pays)) { p.ta = ta; p.tb = tb; }
    }
    p.use = cfg.column > 0 || (n >= (1ll << 20) && (wide || p.ta + p.tb > 0));
    return p;
}

// regrouped column walk (k_transport_column_parked) for a tiled call? -- LaunchCfg::column_park
static bool column_parked(const LaunchCfg &cfg) { return cfg.column_park > 0 || (cfg.column_park < 0 && cfg.steps_hint >= 2.25); }

bool column_gather_selected(const DevGrid &g, const LaunchCfg &cfg, long long n)
{
    return column_plan(g, cfg, n).use && cfg.column != 2;
}

// The depth bound of this call for the all-reduce (k_column_bound), handed to the host through one int in mapped
// page-locked memory.  from_copy: walk the z-fastest copy the
// gather just made (the resident grid may still be on its way over PCIe); otherwise walk the resident grid.
// fill / to_host: see k_column_bound (depth-limited upload: the deep planes a packet can read go into the resident grid).
// Returns whether the kernel was enqueued.
static bool enqueue_bound(const DevGrid &g, const ColGeom &cg, ColumnWorkspace *ws, cudaStream_t s, int *launches, bool from_copy,
                          double *fill = nullptr, bool to_host = true)
{
    if (!ws->h_bound) {
        if (cudaHostAlloc((void **)&ws->h_bound, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void **)&ws->d_bound, ws->h_bound, 0) != cudaSuccess ||
            cudaMalloc((void **)&ws->bound_scratch, 2 * sizeof(int)) != cudaSuccess ||
            cudaMemsetAsync(ws->bound_scratch, 0, 2 * sizeof(int), s) != cudaSuccess ||
            cudaEventCreateWithFlags(&ws->ev_bound, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ws->ev_gathered, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            if (ws->h_bound) cudaFreeHost(ws->h_bound);
            ws->h_bound = nullptr;
        }
    }
    if (!ws->h_bound) return false;
    if (to_host) *ws->h_bound = 0;
    // On the transport's own stream, AHEAD of the transport (10-15 us): the host waits for this answer before it can
    // enqueue the reduction, and a kernel on a side stream does not get onto the SMs while the transport's one-CTA-per-SM
    // launch holds them -- measured (TAMC_TRACE): the host then sat in that wait until the transport had finished, and
    // everything it enqueues afterwards (the copies that should run BESIDE the transport) started 1.8 ms late.
    cudaStream_t sb = s;
    const int cols = cg.tw * cg.th;
    if (from_copy)
        k_column_bound<<<(cols + 255) / 256, 256, sizeof(double) * (size_t)cg.nzp, sb>>>(g, cg, (const double *)ws->rkT, ws->bound_scratch,
                                                                                      reinterpret_cast<unsigned int *>(ws->bound_scratch + 1),
                                                                                      to_host ? ws->d_bound : nullptr, fill);
    else
        k_column_bound_resident<<<(cols + 255) / 256, 256, sizeof(double) * (size_t)cg.nzp, sb>>>(g, cg, ws->bound_scratch,
                                                                                               reinterpret_cast<unsigned int *>(ws->bound_scratch + 1), ws->d_bound);
    if (to_host) {
        cudaEventRecord(ws->ev_bound, sb);
        ws->bound_pending = true;
    }
    if (launches) *launches += 1;
    return true;
}

// Column form of the shipped regime (tamc_column.cuh): gather the beam's columns, transport, add the full-crossing term.
static cudaError_t launch_column(const DevGrid &g, const LaunchCfg &cfg, long long n, uint64_t seed, uint64_t first_id,
                                 unsigned long long *d_cnt, cudaStream_t s, int *launches, ColumnWorkspace *ws, bool gather,
                                 const ColumnPlan &plan)
{
    ColGeom cg;
    int planes = 0;                                    // depth limit of a columns-first upload over PCIe (0 = every plane)
    if (cfg.gather_depth > 0) planes = cfg.gather_depth;
    else if (cfg.gather_depth < 0 && cfg.depth_hint > 0) planes = cfg.depth_hint + (cfg.depth_hint / 2 > 16 ? cfg.depth_hint / 2 : 16);
    cudaError_t e0 = column_setup(g, ws, gather, s, cg, planes);
    if (e0 != cudaSuccess) return e0;
    ws->last_kz_lo = cg.kz_lo;
    if (launches && gather) *launches += 1;
    ws->bound_pending = false;
    const bool deep = cg.kz_lo > 0;          // depth-limited columns-first upload: the builds that can read below the copied planes
    if (gather && (cfg.want_bound || deep)) {
        // depth-limited: the same walk also copies the deeper planes a packet of this call can reach into the resident grid,
        // and the kernels below take their deep reads from there (not from the caller's array over PCIe)
        double *fill = deep ? const_cast<double *>(g.rhokap) : nullptr;
        if (enqueue_bound(g, cg, ws, s, launches, true, fill, cfg.want_bound) && deep) cg.deep = g.rhokap;
    } else if (cfg.want_bound) enqueue_bound(g, cg, ws, s, launches, false);
    const size_t smem = sizeof(double) * (size_t)cg.nzp;
    LaunchCfg c2 = cfg;
    c2.block = 256;
    cudaError_t e;
    const int ta = gather ? plan.ta : 0, tb = gather ? plan.tb : 0;
    if (ta + tb > 0) {
        const size_t tsmem = smem + (size_t)cg.tw * cg.th * (8 * (size_t)ta + 4 * (size_t)tb);
        const long long want = (n + 1023) / 1024;
        const int grid = (int)(want < cfg.num_sms ? want : cfg.num_sms);
        if (column_parked(cfg)) {
            // tiles, then (8-byte aligned) one ParkQueue per warp
            const size_t psmem = smem + 8 * ((size_t)cg.tw * cg.th * (size_t)ta + (((size_t)cg.tw * cg.th * (size_t)tb + 1) >> 1)) + 32 * sizeof(ParkQueue);
            auto kern = deep ? k_transport_column_parked<true> : k_transport_column_parked<false>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
            if (e != cudaSuccess) return e;
            kern<<<grid, 1024, psmem, s>>>(g, n, seed, first_id, cg, (const double *)ws->rkT, ws->stops, d_cnt, ta, tb);
        } else {
            auto kern = deep ? k_transport_column_tiled<true> : k_transport_column_tiled<false>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
            if (e != cudaSuccess) return e;
            kern<<<grid, 1024, tsmem, s>>>(g, n, seed, first_id, cg, (const double *)ws->rkT, ws->stops, d_cnt, ta, tb);
        }
        e = cudaGetLastError();
    } else
    if (!gather) e = launch_sized(k_transport_column<false, 4>, c2, smem, n, s, g, n, seed, first_id, cg, (const double *)nullptr, ws->stops, d_cnt);
    else if (deep) e = launch_sized(k_transport_column<true, 4, true>, c2, smem, n, s, g, n, seed, first_id, cg, (const double *)ws->rkT, ws->stops, d_cnt);
    else if (cfg.min_ctas == 2) e = launch_sized(k_transport_column<true, 6>, c2, smem, n, s, g, n, seed, first_id, cg, (const double *)ws->rkT, ws->stops, d_cnt);
    else e = launch_sized(k_transport_column<true, 4>, c2, smem, n, s, g, n, seed, first_id, cg, (const double *)ws->rkT, ws->stops, d_cnt);
    if (e != cudaSuccess) return e;
    DevGrid gf = g;
    if (gather && ws->box_rk && (ws->gather_src || ws->share_gather == 2)) {
        // rhokap(i,j,k) = box[(i-i0) + tw*((j-j0) + th*(k-1))]: the resident index expression i + sx*j + sxy*k with
        // sx = tw, sxy = tw*th and the origin moved
        gf.sx = cg.tw;
        gf.sxy = (long long)cg.tw * cg.th;
        gf.rhokap = ws->box_rk - ((long long)cg.i0 + (long long)cg.tw * cg.j0 + gf.sxy);
    }
    if (deep) k_column_finish<true><<<(cg.tw * cg.th + 31) / 32, 32 * kFinishChunks, smem, s>>>(gf, cg, ws->stops, d_cnt);
    else k_column_finish<false><<<(cg.tw * cg.th + 31) / 32, 32 * kFinishChunks, smem, s>>>(gf, cg, ws->stops, d_cnt);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

cudaError_t launch_transport(const DevGrid &g_in, const LaunchCfg &cfg_in, long long n, uint64_t seed, uint64_t first_id,
                             unsigned long long *d_cnt, tamc_packet_record *d_rec, cudaStream_t s, int *launches,
                             ColumnWorkspace *ws, int *form)
{
    if (n <= 0) return cudaSuccess;
    DevGrid g = g_in;
    for (int r = 0; r < 10; ++r) {          // Philox4x32 key schedule (Random123): key + r * (W32_0, W32_1)
        g.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        g.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    LaunchCfg cfg = cfg_in;
    const bool pool = cfg.variant == 3 && (g.flags & TAMC_SCATTER) && !d_rec;
    const bool auto_block = cfg.block <= 0;
    if (auto_block) cfg.block = pool ? 128 : 256;
    const bool merge = cfg.merge < 0 ? (g.flags & TAMC_SCATTER) != 0 : cfg.merge != 0;
    const size_t smem = faces_bytes(g);
    // shipped regime, default variant: the column form once the call is large enough to pay for its two small extra kernels
    const ColumnPlan plan = (ws && !d_rec) ? column_plan(g, cfg, n) : ColumnPlan{false, 0, 0};
    if (plan.use) {
        const bool parked = plan.ta + plan.tb > 0 && cfg.column != 2 && column_parked(cfg);
        if (form) *form = cfg.column == 2 ? FORM_COLUMN_RESIDENT : (parked ? FORM_COLUMN_PARKED : (plan.ta + plan.tb > 0 ? FORM_COLUMN_TILED : FORM_COLUMN));
        return launch_column(g, cfg, n, seed, first_id, d_cnt, s, launches, ws, cfg.column != 2, plan);
    }
    if (launches) *launches += 1;
    tamc_packet_record *none = nullptr;
    if (ws && !d_rec) {
        ws->bound_pending = false;
        ColGeom cgb;
        // the all-reduce's depth bound must not depend on the kernel form (ranks may differ in packet count): same rule here
        if (cfg.want_bound && !(g.flags & (TAMC_SCATTER | TAMC_FRESNEL)) && beam_box(g, cgb)) enqueue_bound(g, cgb, ws, s, launches, false);
    }

    // Options outside the shipped path (Fresnel boundaries, periodic lateral boundaries, Gaussian beam) are compiled into
    // the thread-per-packet kernels and the `ext` build of the pool kernel only; the stub-regime and persistent kernels
    // stay as they are.  Periodic boundaries alone change nothing in the stub regime (straight-down flights).
    const bool scat = (g.flags & TAMC_SCATTER) != 0, gauss = g.gauss_sigma > 0.;
    // per-voxel albedo / hgg (tamc_set_optics_grids) live in the exact, thread-per-packet and flight kernels only
    const bool vox_optics = scat && (g.albedo_g || g.hgg_g);
    const bool flight_ok = pool && ws && cfg.flight != 0 && !((g.flags & (TAMC_FRESNEL | TAMC_PERIODIC)) != 0 || gauss);
    const bool simple_ext = (scat ? (!pool && (gauss || (g.flags & TAMC_PERIODIC))) : (gauss || (g.flags & TAMC_FRESNEL))) ||
                            (vox_optics && cfg.variant != 2 && !flight_ok);
    if (form) *form = cfg.variant == 2 ? FORM_EXACT : ((d_rec || cfg.variant == 0 || simple_ext) ? FORM_SIMPLE : (pool ? FORM_POOL : FORM_PERSISTENT));
    if (cfg.variant == 2) {
        if (d_rec) {
            if (merge) return launch_sized(k_transport_exact<MergeTally, true>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, d_rec);
            return launch_sized(k_transport_exact<DirectTally, true>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, d_rec);
        }
        if (merge) return launch_sized(k_transport_exact<MergeTally, false>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, none);
        return launch_sized(k_transport_exact<DirectTally, false>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, none);
    }
    if (d_rec) {
        if (merge) return launch_sized(k_transport_simple<MergeTally32, true>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, d_rec);
        return launch_sized(k_transport_simple<DirectTally32, true>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, d_rec);
    }
    // e.g. Fresnel boundaries without the scatter loop: the stub kernels keep no packet id, use the thread-per-packet kernel
    if (cfg.variant == 0 || simple_ext) {
        if (merge) return launch_sized(k_transport_simple<MergeTally32, false>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, none);
        return launch_sized(k_transport_simple<DirectTally32, false>, cfg, smem, n, s, g, n, seed, first_id, d_cnt, none);
    }
    if (pool && !simple_ext) {
        const bool ext = (g.flags & (TAMC_FRESNEL | TAMC_PERIODIC)) != 0 || gauss;      // options only the `ext` pool build carries
        if (ws && cfg.flight != 0 && !ext) {
            // flight kernel (tamc_flight.cuh): interleaved voxel records built before, tally written back after
            const size_t nvox = (size_t)g.nxg * g.nyg * g.nzg;
            if (ws->vox_elems < nvox) {
                cudaFree(ws->vox);
                ws->vox = nullptr;
                ws->vox_elems = 0;
                cudaError_t e = cudaMalloc(&ws->vox, nvox * sizeof(double2));
                if (e != cudaSuccess) return e;
                ws->vox_elems = nvox;
            }
            // interleaved {opacity, tally} records once the grids exceed L2 (one sector per visited voxel instead of two);
            // while they fit, separate arrays: the REDs under a narrow beam must not share sectors with the loads
            const bool inter = flight_interleaved(cfg, nvox);
            double *voxd = reinterpret_cast<double *>(ws->vox);
            if (inter) k_vox_pack<<<cfg.num_sms * 8, 256, 0, s>>>(g, ws->vox);
            else k_rk_compact<<<cfg.num_sms * 8, 256, 0, s>>>(g, voxd);
            const int chunk = cfg.chunk > 0 ? cfg.chunk : 64;
            const int walk_min = cfg.walk_min < 1 ? 1 : (cfg.walk_min > 32 ? 32 : cfg.walk_min);
            const int launch_min = cfg.flight_launch_min < 1 ? 1 : (cfg.flight_launch_min > 16 ? 16 : cfg.flight_launch_min);
            const size_t fsmem = ((smem + 7) & ~(size_t)7) + (size_t)(256 / 32) * CNT_N * sizeof(unsigned long long) + 2 * 256 * sizeof(double);
            LaunchCfg c2 = cfg;
            c2.block = 256;
            cudaError_t e;
            const int regs = cfg.flight_regs ? cfg.flight_regs : (inter ? 2 : 3);
            if (g.albedo_g || g.hgg_g) {
                // per-voxel albedo / hgg: compact copies in the tally's index beside the opacities, the kGrids build
                if (ws->optc_elems < 2 * nvox) {
                    cudaFree(ws->optc);
                    ws->optc = nullptr;
                    ws->optc_elems = 0;
                    e = cudaMalloc(&ws->optc, 2 * nvox * sizeof(double));
                    if (e != cudaSuccess) return e;
                    ws->optc_elems = 2 * nvox;
                }
                double *albc = g.albedo_g ? ws->optc : nullptr, *hggc = g.hgg_g ? ws->optc + nvox : nullptr;
                k_optics_compact<<<cfg.num_sms * 8, 256, 0, s>>>(g, albc, hggc);
                if (launches) *launches += 1;
                if (inter) e = launch_sized(k_transport_flight<256, 2, true, true>, c2, fsmem, n, s, g, voxd, n, first_id, chunk, walk_min, d_cnt, (const double *)albc, (const double *)hggc, launch_min);
                else e = launch_sized(k_transport_flight<256, 2, false, true>, c2, fsmem, n, s, g, voxd, n, first_id, chunk, walk_min, d_cnt, (const double *)albc, (const double *)hggc, launch_min);
            } else if (inter) {
                if (regs == 2) e = launch_sized(k_transport_flight<256, 2, true>, c2, fsmem, n, s, g, voxd, n, first_id, chunk, walk_min, d_cnt, (const double *)nullptr, (const double *)nullptr, launch_min);
                else if (regs == 4) e = launch_sized(k_transport_flight<256, 4, true>, c2, fsmem, n, s, g, voxd, n, first_id, chunk, walk_min, d_cnt, (const double *)nullptr, (const double *)nullptr, launch_min);
                else e = launch_sized(k_transport_flight<256, 3, true>, c2, fsmem, n, s, g, voxd, n, first_id, chunk, walk_min, d_cnt, (const double *)nullptr, (const double *)nullptr, launch_min);
            } else if (cfg.flight_agg > 0) {
                e = launch_sized(k_transport_flight<256, 3, false, false, true>, c2, fsmem, n, s, g, voxd, n, first_id, chunk, walk_min, d_cnt, (const double *)nullptr, (const double *)nullptr, launch_min);
            } else {
                if (regs == 2) e = launch_sized(k_transport_flight<256, 2, false>, c2, fsmem, n, s, g, voxd, n, first_id, chunk, walk_min, d_cnt, (const double *)nullptr, (const double *)nullptr, launch_min);
                else if (regs == 4) e = launch_sized(k_transport_flight<256, 4, false>, c2, fsmem, n, s, g, voxd, n, first_id, chunk, walk_min, d_cnt, (const double *)nullptr, (const double *)nullptr, launch_min);
                else e = launch_sized(k_transport_flight<256, 3, false>, c2, fsmem, n, s, g, voxd, n, first_id, chunk, walk_min, d_cnt, (const double *)nullptr, (const double *)nullptr, launch_min);
            }
            if (e != cudaSuccess) return e;
            if (inter) k_vox_unpack<<<cfg.num_sms * 8, 256, 0, s>>>(g, ws->vox);
            if (launches) *launches += inter ? 2 : 1;
            if (form) *form = FORM_FLIGHT;
            return cudaGetLastError();
        }
        // work-queue regrouping: faces + one 64-packet pool per warp in shared memory
        int chunk = cfg.chunk > 0 ? cfg.chunk : 64;
        // grids beyond L2 (400^3: 1 GB): the walk waits on DRAM for every opacity (ncu: long scoreboard 3.5 per issue) --
        // fetch it one loop pass early (kAhead), and 256-thread CTAs (measured on phantom400: 128 x 5: 58.7 ms per 2e6
        // packets, with kAhead 54.5; 256 x 2: 51.6; when the grids sit in L2 kAhead costs 1 %, so only there)
        int dev = 0, l2 = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
        const bool beyond_l2 = 8. * ((double)g.sxy * (g.nzg + 2) + (double)g.nxg * g.nyg * g.nzg) > 2. * (double)l2;
        const bool fres = (g.flags & (TAMC_FRESNEL | TAMC_PERIODIC)) != 0 || gauss;      // the `ext` build
        if (auto_block ? !beyond_l2 : cfg.block <= 128) {       // auto: 128 threads, 5 CTAs per SM while the grids fit L2
            LaunchCfg c2 = cfg;
            c2.block = 128;
            const size_t qsmem = ((smem + 15) & ~(size_t)15) + 4 * sizeof(WarpPool);
            if (fres) return launch_sized(k_transport_pool<128, 5, true>, c2, qsmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
            if (beyond_l2) return launch_sized(k_transport_pool<128, 5, false, true>, c2, qsmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
            return launch_sized(k_transport_pool<128, 5, false>, c2, qsmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
        }
        LaunchCfg c2 = cfg;
        c2.block = 256;
        const size_t qsmem = ((smem + 15) & ~(size_t)15) + 8 * sizeof(WarpPool);
        if (fres) return launch_sized(k_transport_pool<256, 2, true>, c2, qsmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
        if (beyond_l2) return launch_sized(k_transport_pool<256, 2, false, true>, c2, qsmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
        return launch_sized(k_transport_pool<256, 2, false>, c2, qsmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
    }
    // shipped (stub) regime with enough packets to pay for zeroing and flushing a tile per SM: privatise the
    // top planes of the tally under the beam in shared memory
    if (!(g.flags & TAMC_SCATTER) && cfg.tile != 0 && (cfg.tile > 0 || n >= (1ll << 21))) {
        const double R = sqrt(g.spot_r2);
        TileGeom tg;
        tg.i0 = max(1, (int)((g.xmax - R) * g.inv_dx) + 1);
        tg.j0 = max(1, (int)((g.ymax - R) * g.inv_dy) + 1);
        tg.tw = min(g.nxg, (int)((g.xmax + R) * g.inv_dx) + 1) - tg.i0 + 1;
        tg.th = min(g.nyg, (int)((g.ymax + R) * g.inv_dy) + 1) - tg.j0 + 1;
        int dev = 0, optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        const long long avail = (long long)optin - (long long)smem - 32 * (long long)sizeof(MiniReservoir) - 1024;
        const long long plane = (long long)tg.tw * tg.th * (long long)sizeof(double);
        long long layers = plane > 0 ? avail / plane : 0;
        layers = min(layers, (long long)min(g.nzg, cfg.tile > 0 ? cfg.tile : 8));
        // auto: only narrow beams (few thousand columns), where the same tally addresses are hit so often that
        // the L2 atomic unit serialises; on a wide footprint (homog200: 7 500 columns) the tile is neutral.
        const bool wanted = cfg.tile > 0 || (long long)tg.tw * tg.th <= 4096;
        if (wanted && tg.tw > 0 && tg.th > 0 && layers >= 1) {
            tg.layers = (int)layers;
            const size_t tsmem = smem + (size_t)(plane * layers) + 32 * sizeof(MiniReservoir);
            int chunk = cfg.chunk > 0 ? cfg.chunk : 1024;
            cudaError_t e = cudaFuncSetAttribute(k_transport_stub_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
            if (e != cudaSuccess) return e;
            if (form) *form = FORM_TILE;
            k_transport_stub_tiled<<<cfg.num_sms, 1024, tsmem, s>>>(g, n, seed, first_id, chunk, tg, d_cnt);
            return cudaGetLastError();
        }
    }
    // persistent warps: faces + one reservoir per warp in shared memory
    const size_t psmem = smem + (size_t)(cfg.block / 32) * (sizeof(WarpReservoir) + CNT_N * sizeof(unsigned long long));
    int chunk = cfg.chunk;
    if (chunk <= 0) {
        // aim for >= 8 chunks per resident warp so the tail is short, within [32, 1024] ids per claim
        const long long warps = (long long)cfg.num_sms * 24;
        long long c = n / (warps * 8);
        c = c < 32 ? 32 : (c > 1024 ? 1024 : c);
        chunk = (int)(c / 32 * 32);
    }
    if (g.flags & TAMC_SCATTER) {
        if (cfg.min_ctas >= 3) {
            if (merge) return launch_sized(k_transport_persistent<MergeTally32, true, 3>, cfg, psmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
            return launch_sized(k_transport_persistent<DirectTally32, true, 3>, cfg, psmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
        }
        if (merge) return launch_sized(k_transport_persistent<MergeTally32, true, 2>, cfg, psmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
        return launch_sized(k_transport_persistent<DirectTally32, true, 2>, cfg, psmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
    }
    if (merge) return launch_sized(k_transport_persistent<MergeTally32, false, 4>, cfg, psmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
    return launch_sized(k_transport_persistent<DirectTally32, false, 4>, cfg, psmem, n, s, g, n, seed, first_id, chunk, cfg.scatter_min, d_cnt);
}

// Both passes of the column form's launch voxel (launch_voxel_fp32 / launch_point, tamc_fast.cuh) over n Philox blocks:
// out[0] = draws the fp32 pass handed to fp64, out[1] = draws where it kept a voxel that differs from fp64's (must be 0).
__global__ void __launch_bounds__(256) k_selfcheck_launch(const DevGrid g, long long n, uint64_t first_id, unsigned long long *__restrict__ out)
{
    unsigned long long fb = 0ull, bad = 0ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t gid = first_id + (uint64_t)i;
        const uint4 w = philox_block(g, (uint32_t)gid, (uint32_t)(gid >> 32), 0u);
        int a, b, c, d;
        double x, y;
        const bool ok = launch_voxel_fp32(g, w.x, w.y, a, b);
        launch_point(g, w.x, w.y, x, y, c, d);
        fb += !ok;
        bad += ok && (a != c || b != d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        fb += __shfl_xor_sync(0xffffffffu, fb, o);
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (fb) atomicAdd(out, fb);
        if (bad) atomicAdd(out + 1, bad);
    }
}

cudaError_t launch_selfcheck(const DevGrid &g_in, long long n, uint64_t seed, uint64_t first_id, unsigned long long *d_out, int num_sms, cudaStream_t s)
{
    DevGrid g = g_in;
    for (int r = 0; r < 10; ++r) {
        g.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        g.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    k_selfcheck_launch<<<num_sms * 8, 256, 0, s>>>(g, n, first_id, d_out);
    return cudaGetLastError();
}

cudaError_t launch_probe(const DevGrid &g, const LaunchCfg &cfg, long long n, uint64_t seed, unsigned long long *d_cnt,
                         cudaStream_t s, ColumnWorkspace *ws, int probe_form)
{
    const float disk_r_vox = (float)(sqrt(g.spot_r2) * g.inv_dx);
    LaunchCfg c2 = cfg;
    c2.block = 256;
    LaunchCfg cp = cfg;
    if (probe_form == 1 && cp.column <= 0) cp.column = 1;
    const ColumnPlan plan = ws ? column_plan(g, cp, n) : ColumnPlan{false, 0, 0};
    const bool column = probe_form < 0 ? plan.use : probe_form == 1;
    if (column && ws) {
        ColGeom cg;
        cudaError_t e = column_setup(g, ws, true, s, cg);
        if (e != cudaSuccess) return e;
        if (plan.ta + plan.tb > 0) {       // the transport's shape: one 1024-thread CTA per SM with its tiles
            const size_t tsmem = (size_t)cg.tw * cg.th * (8 * (size_t)plan.ta + 4 * (size_t)plan.tb);
            e = cudaFuncSetAttribute(k_probe_column<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem);
            if (e != cudaSuccess) return e;
            const long long want = (n + 1023) / 1024;
            k_probe_column<true><<<(int)(want < cfg.num_sms ? want : cfg.num_sms), 1024, tsmem, s>>>(
                g, n, seed, disk_r_vox, cg, (const double *)ws->rkT, ws->stops, d_cnt, plan.ta, plan.tb);
            e = cudaGetLastError();
        } else {
            e = launch_sized(k_probe_column<false>, c2, 0, n, s, g, n, seed, disk_r_vox, cg, (const double *)ws->rkT, ws->stops, d_cnt, 0, 0);
        }
        if (e != cudaSuccess) return e;
        const size_t smem = sizeof(double) * (size_t)cg.nzp;
        k_column_finish<false><<<(cg.tw * cg.th + 31) / 32, 32 * kFinishChunks, smem, s>>>(g, cg, ws->stops, d_cnt);
        return cudaGetLastError();
    }
    return launch_sized(k_probe, c2, 0, n, s, g, n, seed, disk_r_vox, d_cnt);
}

// record / replay of the voxel-index stream (see k_trace)
cudaError_t launch_trace(const DevGrid &g_in, long long n, uint64_t seed, uint64_t first_id, const long long *d_off, int *d_trace,
                         int *d_counts, int num_sms, cudaStream_t s)
{
    DevGrid g = g_in;
    for (int r = 0; r < 10; ++r) {
        g.rk[2 * r] = (uint32_t)seed + (uint32_t)r * 0x9E3779B9u;
        g.rk[2 * r + 1] = (uint32_t)(seed >> 32) + (uint32_t)r * 0xBB67AE85u;
    }
    const long long want = (n + 255) / 256;
    const int grid = (int)(want < (long long)num_sms * 8 ? want : (long long)num_sms * 8);
    k_trace<<<grid, 256, faces_bytes(g), s>>>(g, n, seed, first_id, d_off, d_trace, d_counts);
    return cudaGetLastError();
}

// the layout rule of the flight kernel: interleaved {opacity, tally} records once the grids exceed L2
bool flight_interleaved(const LaunchCfg &cfg, size_t nvox)
{
    if (cfg.flight_inter >= 0) return cfg.flight_inter != 0;
    int dev = 0, l2 = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev);
    return 16. * (double)nvox > 2. * (double)l2;
}

cudaError_t launch_probe_trace(const DevGrid &g, const LaunchCfg &cfg, double2 *vox, long long n, const long long *d_off, const int *d_trace,
                               unsigned long long *d_cnt, int num_sms, cudaStream_t s, bool pack)
{
    const bool inter = flight_interleaved(cfg, (size_t)g.nxg * g.nyg * g.nzg);
    double *voxd = reinterpret_cast<double *>(vox);
    if (pack) {
        if (inter) k_vox_pack<<<num_sms * 8, 256, 0, s>>>(g, vox);
        else {
            k_rk_compact<<<num_sms * 8, 256, 0, s>>>(g, voxd);
            cudaMemsetAsync(g.jmean, 0, sizeof(double) * (size_t)g.nxg * g.nyg * g.nzg, s);
        }
        return cudaGetLastError();
    }
    if (inter) k_probe_trace<true><<<num_sms * 6, 256, 0, s>>>(voxd, voxd + 1, n, d_off, d_trace, d_cnt);
    else k_probe_trace<false><<<num_sms * 6, 256, 0, s>>>(voxd, g.jmean, n, d_off, d_trace, d_cnt);
    return cudaGetLastError();
}

cudaError_t launch_fill(double *p, size_t n, double v, int num_sms, cudaStream_t s)
{
    k_fill<<<num_sms * 8, 256, 0, s>>>(p, n, v);
    return cudaGetLastError();
}

}  // namespace tamc
