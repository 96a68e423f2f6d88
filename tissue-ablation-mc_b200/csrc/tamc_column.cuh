// tamc_column.cuh -- shipped (stub) regime, column form: variant 3's kernel when the scatter loop is off.
//
// In the shipped regime (mcpolar.f90:166-169: a packet ends at its first interaction) sourcephCO2
// (sourceph.f90:37-42) sends every packet straight down (nxp = nyp = 0, nzp = -1), so a flight never leaves
// the voxel column it was launched into and tauint1 (inttau2.f90:37-63) degenerates to a walk down that
// column.  Two consequences, both exact:
//
//   1. The wall distance of a FULL crossing of voxel k is the same number for every packet: the packet
//      enters at zface(k+1) - delta (update_pos, inttau2.f90:163-166; the launch voxel is entered at zp0)
//      and leaves through zface(k), so dcell(k) = -(zface(k) - (zface(k+1) - delta)) and the deposit
//      dcell(k)*rhokap(i,j,k) (inttau2.f90:46) depends on the voxel only.  The tally of all full crossings
//      is therefore  F(i,j,k) * dcell(k) * rhokap(i,j,k)  with F = the number of packets of the column that
//      stopped below k.  The transport kernel records where each packet stopped (one u32 RED) and tallies
//      only the final partial deposit tau - taurun (inttau2.f90:51-53, one fp64 RED); k_column_finish turns
//      the stop counts into F by a running sum up the column and adds the full-crossing term.  Per packet
//      that is 2 atomics instead of one per voxel-step.
//   2. The opacities a packet reads are consecutive in z.  A z-fastest copy of the columns under the beam's
//      bounding box (k_column_gather, refreshed every MC call: ~10 MB at 200^3) lets one 256-bit load
//      (LDG.E.256) fetch four voxel-steps' worth of rhokap.
//
// The plain persistent kernel sits on the L1TEX address throughput (one uncoalesced rhokap load + one jmean
// RED per voxel-step, profiles/); this form issues ~1.3 loads + <= 2 REDs per PACKET.  Same Philox streams,
// same launch arithmetic (launch_fast), same taurun accumulation order, hence the same stop voxel and the
// same partial deposit bit for bit; the grid differs from the step-by-step tally only by fp64 summation
// order (F*d instead of d+d+...+d).
#pragma once

#include "tamc_fast.cuh"

namespace tamc {

// z-fastest copy of the bounding-box columns: rkT[(dj*tw + di)*nzp + (k-1)] = rhokap(i0+di, j0+dj, k).
// 32 x 32 (x, z) tiles through shared memory: reads coalesced along x, writes coalesced along z.
// `src` is the opacity grid in the reference's layout (halo included): the resident copy, or -- tamc_run_optics, the
// columns-first upload -- the caller's page-locked host array read over PCIe (zero-copy; every voxel of the box is read
// exactly once, in 256-byte row segments).  kBox: also keep the tile in x-fastest order, box[di + tw*(dj + th*(k-1))],
// for k_column_finish, so nothing on this call's critical path waits for the full-grid upload.
template <bool kBox>
__global__ void __launch_bounds__(256) k_column_gather(const DevGrid g, const ColGeom cg, const double *__restrict__ src,
                                                       double *__restrict__ rkT, double *__restrict__ box)
{
    __shared__ double tile[32][33];
    const int dj = blockIdx.y, di0 = blockIdx.x * 32, kz0 = cg.kz_lo + blockIdx.z * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long rowj = (long long)g.sx * (cg.j0 + dj);
    double v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {                                // all four loads in flight before the first use
        const int kz = kz0 + 8 * r + ty, di = di0 + tx;          // kz = k - 1
        v[r] = 0.;
        if (di < cg.tw && kz < g.nzg) v[r] = src[(cg.i0 + di) + rowj + g.sxy * (kz + 1)];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int kz = kz0 + 8 * r + ty, di = di0 + tx;
        tile[8 * r + ty][tx] = v[r];
        if (kBox && di < cg.tw && kz < g.nzg) box[di + (size_t)cg.tw * (dj + (size_t)cg.th * kz)] = v[r];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        const int di = di0 + r + ty, kz = kz0 + tx;
        if (di < cg.tw && kz < cg.nzp) rkT[((size_t)dj * cg.tw + di) * cg.nzp + kz] = tile[tx][r + ty];
    }
}

__device__ __forceinline__ void ldg256(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

// The opacities of the four-voxel group kz = gb .. gb+3 of column col_id: one 256-bit load of the z-fastest copy, or --
// below the depth the columns-first upload copied (ColGeom::kz_lo) -- four loads from the caller's grid over PCIe.
template <bool kDeep>
__device__ __forceinline__ void load_group(const ColGeom &cg, const double *__restrict__ col, int col_id, int gb,
                                           double &r0, double &r1, double &r2, double &r3)
{
    if (!kDeep || gb >= cg.kz_lo) {
        ldg256(col + gb, r0, r1, r2, r3);           // col = the column's slice of the z-fastest copy
        return;
    }
    const int dj = col_id / cg.tw, di = col_id - dj * cg.tw;
    const double *p = cg.deep + ((long long)(cg.i0 + di) + (long long)cg.deep_sx * (cg.j0 + dj) + cg.deep_sxy * (gb + 1));
    r0 = p[0]; r1 = p[cg.deep_sxy]; r2 = p[2 * cg.deep_sxy]; r3 = p[3 * cg.deep_sxy];    // gb + 3 < kz_lo <= nzg
}

// dcell(k) for k = 1..nzp into shared memory (0 past the launch plane / the grid): the wall distance
// voxel_step_fast computes for a straight-down flight, fmin(fmin(100000., 100000.), (fz - zcur) * inz) with inz = -1.
__device__ __forceinline__ void stage_column_steps(const DevGrid &g, int nzp, double *s_dz)
{
    const double *zf = g.faces + (g.nxg + 1) + (g.nyg + 1);      // zface(1:nzg+1), 0-based here: zf[k-1] = zface(k)
    for (int k = 1 + threadIdx.x; k <= nzp; k += blockDim.x) {
        double d = 0.;
        if (k <= g.cellk0 && k <= g.nzg) {
            const double zcur = (k == g.cellk0) ? g.zcur0 : zf[k] - g.delta;
            d = fmin(100000., (zf[k - 1] - zcur) * -1.);
        }
        s_dz[k - 1] = d;
    }
    __syncthreads();
}

// One packet per thread, grid-stride.  kGather: read the z-fastest copy with 256-bit loads; otherwise walk the
// resident grid itself (one 8-byte load per voxel-step).
template <bool kGather, int kMinCtas, bool kDeep = false>
__global__ void __launch_bounds__(256, kMinCtas) k_transport_column(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                                             const ColGeom cg, const double *__restrict__ rkT,
                                                             unsigned int *__restrict__ stops,
                                                             unsigned long long *__restrict__ cnt)
{
    extern __shared__ double s_dz[];
    stage_column_steps(g, cg.nzp, s_dz);
    const int plane = g.nxg * g.nyg;
    const int k0 = g.cellk0;
    unsigned long long steps = 0ull;
    unsigned int packets = 0u, absorbed = 0u, bottom = 0u;

    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t gid = first_id + (uint64_t)i;
        const LaunchedColumn L = launch_column(g, philox_block(g, (uint32_t)gid, (uint32_t)(gid >> 32), 0u));
        const double tau = L.tau;
        double taurun = 0.;
        int kstop = 0;                                   // voxel of the interaction; 0 = left through the bottom face
        if (kGather) {
            const int di = (L.cells & 0xffff) - cg.i0, dj = (L.cells >> 16) - cg.j0;
            const int col_id = dj * cg.tw + di;
            const double *col = rkT + (size_t)col_id * cg.nzp;
            int idx = k0 - 1;                            // k - 1 of the voxel the packet is in
            while (idx >= 0) {
                const int gb = idx & ~3;
                double r0, r1, r2, r3;
                load_group<kDeep>(cg, col, col_id, gb, r0, r1, r2, r3);
                const double2 da = *reinterpret_cast<const double2 *>(s_dz + gb), db = *reinterpret_cast<const double2 *>(s_dz + gb + 2);
                // dcell*rhokap, inttau2.f90:40 -- products and sums rounded separately (no FMA contraction), as the oracle does
                const double tc3 = __dmul_rn(db.y, r3), tc2 = __dmul_rn(db.x, r2), tc1 = __dmul_rn(da.y, r1), tc0 = __dmul_rn(da.x, r0);
                // the four voxel-steps of the group without a divergent branch: running sums in the packet's order
                // (slot 3 = the highest voxel first), then the first slot whose sum reaches tau.  Slots above the launch
                // voxel (first group when nzg is not a multiple of 4) hold dcell = 0 and rhokap = 0: taurun + 0 < tau passes.
                const double t3 = __dadd_rn(taurun, tc3), t2 = __dadd_rn(t3, tc2), t1 = __dadd_rn(t2, tc1), t0 = __dadd_rn(t1, tc0);
                const bool p3 = t3 < tau, p2 = t2 < tau, p1 = t1 < tau, p0 = t0 < tau;               // inttau2.f90:42
                const int nstop = !p3 ? 4 : (!p2 ? 3 : (!p1 ? 2 : (!p0 ? 1 : 0)));
                const double before = !p3 ? taurun : (!p2 ? t3 : (!p1 ? t2 : t1));
                if (nstop) { kstop = gb + nstop; taurun = before; break; }
                taurun = t0;
                idx = gb - 1;
            }
        } else {
            int ridx = (L.cells & 0xffff) + g.sx * ((L.cells >> 16) + (g.nyg + 2) * k0);
            for (int k = k0; k >= 1; --k) {
                const double taucell = __dmul_rn(s_dz[k - 1], __ldg(g.rhokap + ridx));
                const double t = __dadd_rn(taurun, taucell);
                if (t < tau) taurun = t; else { kstop = k; break; }
                ridx -= (int)g.sxy;
            }
        }
        ++packets;
        if (kstop) {
            const double rest = tau - taurun;            // inttau2.f90:51-53: dcell*rhokap = ((tau-taurun)/rhokap)*rhokap
            const int j = L.jidx - (k0 - kstop) * plane;
            if (rest != 0.) atomicAdd(g.jmean + j, rest);
            // the stop count only feeds voxels above kstop: nothing to record for a stop in the top plane
            if (kstop < g.nzg) atomicAdd(stops + (j + plane), 1u);
            steps += (unsigned long long)(k0 - kstop + 1);
            ++absorbed;
        } else {
            atomicAdd(stops + (L.jidx - (k0 - 1) * plane), 1u);      // plane 0 of the counts: crossed every voxel
            steps += (unsigned long long)k0;
            ++bottom;
        }
    }
    unsigned long long v[4] = {packets, steps, absorbed, bottom};
    const int slot[4] = {CNT_PACKETS, CNT_STEPS, CNT_ABSORBED, CNT_EXIT0 + 4};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned long long x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(cnt + slot[q], x);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Column form with the top planes of the two tallies privatised in shared memory.
//
// k_transport_column sits on the L1TEX address throughput: per packet two uncoalesced global REDs (partial deposit,
// stop count) and ~1.3 uncoalesced 256-bit loads.  Most packets stop within a few voxels of the surface
// (P(depth d) = q^d (1-q), q = exp(-rhokap*dz)), so one CTA per SM keeps, for every column under the beam,
//   * the partial deposits of the top `ta` planes (fp64; shared memory has no fp64 add, so a CAS loop -- a CTA's
//     packets fall on thousands of columns, contention is negligible), and
//   * the stop counts of the `tb` planes below the launch plane (u32, native shared-memory atomics),
// and flushes them with coalesced REDs at the end.  A shared-memory atomic on 32 random addresses costs a few bank
// conflict cycles instead of 32 address cycles.  Same arithmetic, same result up to fp64 summation order.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void smem_add_f64(double *p, double v)
{
    unsigned long long *a = reinterpret_cast<unsigned long long *>(p);
    unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(a), assumed;
    do {
        assumed = old;
        old = atomicCAS(a, assumed, (unsigned long long)__double_as_longlong(__longlong_as_double((long long)assumed) + v));
    } while (old != assumed);
}

template <bool kDeep>
__global__ void __launch_bounds__(1024, 1) k_transport_column_tiled(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                                                    const ColGeom cg, const double *__restrict__ rkT,
                                                                    unsigned int *__restrict__ stops,
                                                                    unsigned long long *__restrict__ cnt, int ta, int tb)
{
    extern __shared__ double s_dz[];
    const int cols = cg.tw * cg.th;
    double *s_dep = s_dz + cg.nzp;                                             // [ta][cols]
    unsigned int *s_stop = reinterpret_cast<unsigned int *>(s_dep + (size_t)ta * cols);   // [tb][cols], depth 1 first
    for (int i = threadIdx.x; i < ta * cols; i += blockDim.x) s_dep[i] = 0.;
    for (int i = threadIdx.x; i < tb * cols; i += blockDim.x) s_stop[i] = 0u;
    stage_column_steps(g, cg.nzp, s_dz);                                        // ends with __syncthreads()
    const int plane = g.nxg * g.nyg;
    const int k0 = g.cellk0;
    unsigned long long steps = 0ull;
    unsigned int packets = 0u, absorbed = 0u, bottom = 0u;

    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t gid = first_id + (uint64_t)i;
        const LaunchedColumn L = launch_column(g, philox_block(g, (uint32_t)gid, (uint32_t)(gid >> 32), 0u));
        const double tau = L.tau;
        double taurun = 0.;
        int kstop = 0;
        const int di = (L.cells & 0xffff) - cg.i0, dj = (L.cells >> 16) - cg.j0;
        const int col_id = dj * cg.tw + di;
        const double *col = rkT + (size_t)col_id * cg.nzp;
        int idx = k0 - 1;
        while (idx >= 0) {                                                     // as k_transport_column<true>
            const int gb = idx & ~3;
            double r0, r1, r2, r3;
            load_group<kDeep>(cg, col, col_id, gb, r0, r1, r2, r3);
            const double2 da = *reinterpret_cast<const double2 *>(s_dz + gb), db = *reinterpret_cast<const double2 *>(s_dz + gb + 2);
            const double tc3 = __dmul_rn(db.y, r3), tc2 = __dmul_rn(db.x, r2), tc1 = __dmul_rn(da.y, r1), tc0 = __dmul_rn(da.x, r0);
            const double t3 = __dadd_rn(taurun, tc3), t2 = __dadd_rn(t3, tc2), t1 = __dadd_rn(t2, tc1), t0 = __dadd_rn(t1, tc0);
            const bool p3 = t3 < tau, p2 = t2 < tau, p1 = t1 < tau, p0 = t0 < tau;
            const int nstop = !p3 ? 4 : (!p2 ? 3 : (!p1 ? 2 : (!p0 ? 1 : 0)));
            const double before = !p3 ? taurun : (!p2 ? t3 : (!p1 ? t2 : t1));
            if (nstop) { kstop = gb + nstop; taurun = before; break; }
            taurun = t0;
            idx = gb - 1;
        }
        ++packets;
        if (kstop) {
            const double rest = tau - taurun;
            const int d = k0 - kstop;                                          // depth below the launch plane
            const int j = L.jidx - d * plane;
            if (rest != 0.) {
                if (d < ta) smem_add_f64(s_dep + d * cols + col_id, rest);
                else atomicAdd(g.jmean + j, rest);
            }
            if (d >= 1 && d <= tb) atomicAdd(s_stop + (d - 1) * cols + col_id, 1u);
            else if (kstop < g.nzg) atomicAdd(stops + (j + plane), 1u);
            steps += (unsigned long long)(d + 1);
            ++absorbed;
        } else {
            atomicAdd(stops + (L.jidx - (k0 - 1) * plane), 1u);
            steps += (unsigned long long)k0;
            ++bottom;
        }
    }
    __syncthreads();
    // flush: lanes along x, coalesced REDs; k = k0 - depth
    for (int i = threadIdx.x; i < ta * cols; i += blockDim.x) {
        const double v = s_dep[i];
        if (v != 0.) {
            const int d = i / cols, c = i - d * cols, dj = c / cg.tw, di = c - dj * cg.tw;
            atomicAdd(g.jmean + ((cg.i0 - 1 + di) + g.nxg * ((cg.j0 - 1 + dj) + g.nyg * (k0 - d - 1))), v);
        }
    }
    for (int i = threadIdx.x; i < tb * cols; i += blockDim.x) {
        const unsigned int v = s_stop[i];
        if (v) {
            const int d = i / cols + 1, c = i - (d - 1) * cols, dj = c / cg.tw, di = c - dj * cg.tw;
            atomicAdd(stops + ((cg.i0 - 1 + di) + g.nxg * ((cg.j0 - 1 + dj) + g.nyg * (k0 - d))), v);
        }
    }
    unsigned long long v[4] = {packets, steps, absorbed, bottom};
    const int slot[4] = {CNT_PACKETS, CNT_STEPS, CNT_ABSORBED, CNT_EXIT0 + 4};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned long long x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(cnt + slot[q], x);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// The same with the column walk regrouped (default for tiled calls).
//
// In k_transport_column(_tiled) a lane loops over four-voxel groups until its packet stops; most packets stop in the
// first group (1 - q^4), so the later passes of that loop run with a handful of lanes (ncu: 13 of 32 on average, and
// the loop is half of all instructions).  Here every warp does ONE group per packet and parks the unfinished packets
// (column, next group, tau, taurun, tally index: 28 bytes) in a 64-entry queue of its own in shared memory; whenever 32
// are parked the warp takes them out and gives each its next group with all lanes busy.  The launch arithmetic (the
// other half of the instructions) always runs with 32 lanes, as before.  Same per-packet arithmetic in the same order.
// ---------------------------------------------------------------------------------------------------------------
struct ParkQueue {                 // one per warp
    double tau[64], taurun[64];
    int col[64], idx[64], jidx[64];
};

// one four-voxel group of a packet's walk (inttau2.f90:37-63 for a straight-down flight); returns the stop voxel
// (1-based k), 0 = left through the bottom face, -1 = goes on with the group below (idx, taurun advanced)
template <bool kDeep>
__device__ __forceinline__ int column_group(const ColGeom &cg, const double *__restrict__ col, int col_id, const double *s_dz, int &idx,
                                            double tau, double &taurun)
{
    const int gb = idx & ~3;
    double r0, r1, r2, r3;
    load_group<kDeep>(cg, col, col_id, gb, r0, r1, r2, r3);
    const double2 da = *reinterpret_cast<const double2 *>(s_dz + gb), db = *reinterpret_cast<const double2 *>(s_dz + gb + 2);
    const double tc3 = __dmul_rn(db.y, r3), tc2 = __dmul_rn(db.x, r2), tc1 = __dmul_rn(da.y, r1), tc0 = __dmul_rn(da.x, r0);
    const double t3 = __dadd_rn(taurun, tc3), t2 = __dadd_rn(t3, tc2), t1 = __dadd_rn(t2, tc1), t0 = __dadd_rn(t1, tc0);
    const bool p3 = t3 < tau, p2 = t2 < tau, p1 = t1 < tau, p0 = t0 < tau;
    const int nstop = !p3 ? 4 : (!p2 ? 3 : (!p1 ? 2 : (!p0 ? 1 : 0)));
    if (nstop) {
        taurun = !p3 ? taurun : (!p2 ? t3 : (!p1 ? t2 : t1));
        return gb + nstop;
    }
    taurun = t0;
    idx = gb - 1;
    return idx >= 0 ? -1 : 0;
}

template <bool kDeep>
__global__ void __launch_bounds__(1024, 1) k_transport_column_parked(const DevGrid g, long long n, uint64_t seed, uint64_t first_id,
                                                                     const ColGeom cg, const double *__restrict__ rkT,
                                                                     unsigned int *__restrict__ stops,
                                                                     unsigned long long *__restrict__ cnt, int ta, int tb)
{
    extern __shared__ double s_dz[];
    const int cols = cg.tw * cg.th;
    double *s_dep = s_dz + cg.nzp;                                             // [ta][cols]
    unsigned int *s_stop = reinterpret_cast<unsigned int *>(s_dep + (size_t)ta * cols);   // [tb][cols], depth 1 first
    ParkQueue *queues = reinterpret_cast<ParkQueue *>(s_dep + (size_t)ta * cols + (((size_t)tb * cols + 1) >> 1));
    for (int i = threadIdx.x; i < ta * cols; i += blockDim.x) s_dep[i] = 0.;
    for (int i = threadIdx.x; i < tb * cols; i += blockDim.x) s_stop[i] = 0u;
    stage_column_steps(g, cg.nzp, s_dz);                                        // ends with __syncthreads()
    const int plane = g.nxg * g.nyg;
    const int k0 = g.cellk0;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    ParkQueue &Q = queues[threadIdx.x >> 5];
    int qn = 0;                                                                 // parked packets of this warp (warp-uniform)
    unsigned long long steps = 0ull;
    unsigned int packets = 0u, absorbed = 0u, bottom = 0u;

    // a packet that stopped in voxel kstop (> 0) or left through the bottom face (0): the two tallies
    auto tally = [&](int kstop, int col_id, int jidx, double tau, double taurun) {
        if (kstop) {
            const double rest = tau - taurun;                                  // inttau2.f90:51-53
            const int d = k0 - kstop;
            const int j = jidx - d * plane;
            if (rest != 0.) {
                if (d < ta) smem_add_f64(s_dep + d * cols + col_id, rest);
                else atomicAdd(g.jmean + j, rest);
            }
            if (d >= 1 && d <= tb) atomicAdd(s_stop + (d - 1) * cols + col_id, 1u);
            else if (kstop < g.nzg) atomicAdd(stops + (j + plane), 1u);
            steps += (unsigned long long)(d + 1);
            ++absorbed;
        } else {
            atomicAdd(stops + (jidx - (k0 - 1) * plane), 1u);
            steps += (unsigned long long)k0;
            ++bottom;
        }
    };
    // one group for the packet in this lane's registers; parks it if it goes on
    auto advance = [&](bool live, int col_id, int idx, int jidx, double tau, double taurun) {
        int r = 1;
        if (live) {
            r = column_group<kDeep>(cg, rkT + (size_t)col_id * cg.nzp, col_id, s_dz, idx, tau, taurun);
            if (r >= 0) tally(r, col_id, jidx, tau, taurun);
        }
        const bool park = live && r < 0;
        const unsigned pm = __ballot_sync(0xffffffffu, park);
        if (park) {
            const int s = qn + __popc(pm & lt_mask);
            Q.tau[s] = tau; Q.taurun[s] = taurun; Q.col[s] = col_id; Q.idx[s] = idx; Q.jidx[s] = jidx;
        }
        qn += __popc(pm);
        __syncwarp();
    };

    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long wbase = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); wbase < n; wbase += stride) {
        const long long i = wbase + lane;
        const bool live = i < n;
        const uint64_t gid = first_id + (uint64_t)i;
        const LaunchedColumn L = launch_column(g, philox_block(g, (uint32_t)gid, (uint32_t)(gid >> 32), 0u));
        const int di = (L.cells & 0xffff) - cg.i0, dj = (L.cells >> 16) - cg.j0;
        if (live) ++packets;
        advance(live, dj * cg.tw + di, k0 - 1, L.jidx, L.tau, 0.);
        while (qn >= 32) {                                                     // a full warp of parked packets: their next group
            const int s = qn - 32 + lane;
            const double tau = Q.tau[s], taurun = Q.taurun[s];
            const int col_id = Q.col[s], idx = Q.idx[s], jidx = Q.jidx[s];
            qn -= 32;
            __syncwarp();
            advance(true, col_id, idx, jidx, tau, taurun);
        }
    }
    // what is still parked when the ids run out: to completion, one packet per lane
    while (qn > 0) {
        const int take = qn < 32 ? qn : 32;
        const int s = qn - take + lane;
        const bool live = lane < take;
        double tau = 0., taurun = 0.;
        int col_id = 0, idx = 0, jidx = 0;
        if (live) { tau = Q.tau[s]; taurun = Q.taurun[s]; col_id = Q.col[s]; idx = Q.idx[s]; jidx = Q.jidx[s]; }
        qn -= take;
        __syncwarp();
        if (live) {
            int r;
            do r = column_group<kDeep>(cg, rkT + (size_t)col_id * cg.nzp, col_id, s_dz, idx, tau, taurun); while (r < 0);
            tally(r, col_id, jidx, tau, taurun);
        }
    }
    __syncthreads();
    // flush: lanes along x, coalesced REDs; k = k0 - depth
    for (int i = threadIdx.x; i < ta * cols; i += blockDim.x) {
        const double v = s_dep[i];
        if (v != 0.) {
            const int d = i / cols, c = i - d * cols, dj = c / cg.tw, di = c - dj * cg.tw;
            atomicAdd(g.jmean + ((cg.i0 - 1 + di) + g.nxg * ((cg.j0 - 1 + dj) + g.nyg * (k0 - d - 1))), v);
        }
    }
    for (int i = threadIdx.x; i < tb * cols; i += blockDim.x) {
        const unsigned int v = s_stop[i];
        if (v) {
            const int d = i / cols + 1, c = i - (d - 1) * cols, dj = c / cg.tw, di = c - dj * cg.tw;
            atomicAdd(stops + ((cg.i0 - 1 + di) + g.nxg * ((cg.j0 - 1 + dj) + g.nyg * (k0 - d))), v);
        }
    }
    unsigned long long v[4] = {packets, steps, absorbed, bottom};
    const int slot[4] = {CNT_PACKETS, CNT_STEPS, CNT_ABSORBED, CNT_EXIT0 + 4};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unsigned long long x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(cnt + slot[q], x);
    }
}

// Full-crossing deposits.  F(i,j,k) = packets of the column that stopped below voxel k = the running sum of the stop
// counts up the column (plane k of `stops` holds the packets that stopped in voxel k, plane 0 those that left through
// the bottom face); jmean(i,j,k) += F * dcell(k) * rhokap(i,j,k).  A CTA takes 32 columns (lanes along x: coalesced)
// times kFinishChunks z-chunks: every thread first sums the counts of its chunk, a scan over the chunks of a column
// gives each thread its starting F, then it walks its chunk.  Clears the counts it consumed, so the array is all zero
// again for the next call.
constexpr int kFinishChunks = 32;

template <bool kDeep>
__global__ void __launch_bounds__(32 * kFinishChunks) k_column_finish(const DevGrid g, const ColGeom cg, unsigned int *__restrict__ stops,
                                                                      unsigned long long *__restrict__ cnt)
{
    extern __shared__ double s_dz[];
    __shared__ unsigned long long s_sum[kFinishChunks][32];
    stage_column_steps(g, cg.nzp, s_dz);
    const int lane = threadIdx.x & 31, chunk = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    const bool live = t < cg.tw * cg.th;
    const int i = cg.i0 + (live ? t % cg.tw : 0), j = cg.j0 + (live ? t / cg.tw : 0);
    const int plane = g.nxg * g.nyg;
    const int c0 = (i - 1) + g.nxg * (j - 1);
    const long long r0 = (long long)i + (long long)g.sx * j;
    const long long rdeep = (long long)i + (long long)cg.deep_sx * j;          // the same voxel in the caller's grid (ColGeom::deep)
    const int len = (g.nzg + kFinishChunks - 1) / kFinishChunks;
    const int klo = chunk * len + 1, khi = min(g.nzg, klo + len - 1);      // voxels of this chunk

    // counts of planes klo .. khi (the top plane nzg is never recorded); chunk 0 also takes plane 0
    unsigned long long a = 0ull;
    if (live) {
        if (chunk == 0) a = stops[c0];
        for (int k = klo; k <= min(khi, g.nzg - 1); ++k) a += stops[c0 + k * plane];
    }
    s_sum[chunk][lane] = a;
    __syncthreads();
    int depth = 0;                                                          // planes from the top face down to the deepest stop seen here
    if (live) {
        // F of the first voxel of the chunk: plane 0 + every plane below klo
        // (s_sum[c] covers planes c*len+1 .. (c+1)*len, plus plane 0 for c = 0: together exactly the planes <= klo - 1)
        unsigned long long F = 0ull;
        if (chunk == 0) {
            F = stops[c0];
            if (F) { stops[c0] = 0u; depth = g.nzg; }
        } else {
            for (int c = 0; c < chunk; ++c) F += s_sum[c][lane];
        }
        for (int k = klo; k <= khi; ++k) {
            const unsigned int c = (k < g.nzg) ? stops[c0 + k * plane] : 0u;
            if (F) {
                const double rk = (!kDeep || k - 1 >= cg.kz_lo) ? __ldg(g.rhokap + r0 + g.sxy * k) : cg.deep[rdeep + cg.deep_sxy * k];
                const double d = (double)F * (s_dz[k - 1] * rk);
                if (d != 0.) g.jmean[c0 + (k - 1) * plane] += d;
            }
            if (c) { F += c; stops[c0 + k * plane] = 0u; depth = max(depth, g.nzg - k + 1); }
        }
    }
    // feeds the depth limit of the next columns-first upload (LaunchCfg::depth_hint)
    depth = __reduce_max_sync(0xffffffffu, depth);
    if (lane == 0 && depth > 0) atomicMax(cnt + CNT_DEPTH, (unsigned long long)depth);
}

// How deep can a packet of this call get?  The optical depth is tau = -log((x + 0.5) 2^-32) for a 32-bit x
// (inttau2.f90:36 on one Philox word), so tau <= 33 ln 2 = 22.874 for EVERY packet, and a straight-down flight stops no
// later than where its column's running optical depth passes that.  One thread per column of the z-fastest copy walks
// down from the launch plane with the kernel's own chords until the sum passes 23; *out receives the largest number of
// planes (from the top face, one spare) over the columns, nzg when some column never gets there.  Every rank holds the
// same grid and gets the same number: the all-reduce (mcpolar.f90:173) moves only those planes of the box.
//
// Depth-limited columns-first upload (ColGeom::kz_lo > 0): a column whose copied planes do not reach optical depth 23
// goes on in the caller's page-locked grid `cg.deep`, and -- `fill` = the resident grid -- keeps what it reads there, in
// whole four-plane groups (the unit load_group fetches) down to the group in which the column reaches 23.  That is every
// voxel of the column a packet of this call can read: the transport kernels then take their rare deep reads from the
// resident grid in HBM instead of one PCIe round trip per packet and group (measured, homog200 / 1e8 packets, opacity
// dropping 16-fold between two calls: 1 160 ms for that call before).  The full-grid upload that follows
// beside the transport writes the same values to the same addresses.  out == nullptr: only the fill is wanted.
__global__ void __launch_bounds__(256) k_column_bound(const DevGrid g, const ColGeom cg, const double *__restrict__ rkT, int *__restrict__ acc_max,
                                                      unsigned int *__restrict__ done, int *__restrict__ out, double *__restrict__ fill)
{
    extern __shared__ double s_dz[];
    stage_column_steps(g, cg.nzp, s_dz);                               // ends with __syncthreads()
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    int planes = 1;
    if (c < cg.tw * cg.th) {
        const double *col = rkT + (size_t)c * cg.nzp;
        double acc = 0.;
        planes = g.nzg;
        int kz = g.cellk0 - 1;
        bool found = false;
        // four planes per 256-bit load, from the launch plane down (slots above it hold chord 0)
        for (int gb = kz & ~3; gb >= cg.kz_lo && !found; gb -= 4) {
            double r0, r1, r2, r3;
            ldg256(col + gb, r0, r1, r2, r3);
            const double r[4] = {r0, r1, r2, r3};
#pragma unroll
            for (int q = 3; q >= 0; --q) {
                if (!found && gb + q <= g.cellk0 - 1) {
                    acc += s_dz[gb + q] * r[q];
                    kz = gb + q;
                    if (acc >= 23.0) { planes = min(g.nzg, g.nzg - kz + 1); found = true; }
                }
            }
        }
        if (!found && cg.kz_lo > 0 && cg.deep) {
            // (depth-limited upload and a column that needs more than the copied planes: go on in the caller's grid, so the
            // answer never depends on how much this rank happened to copy)
            const int dj = c / cg.tw, di = c - dj * cg.tw;
            const long long at = (long long)(cg.i0 + di) + (long long)cg.deep_sx * (cg.j0 + dj);
            const double *q = cg.deep + at;
            double *w = fill ? fill + at : nullptr;                     // the resident grid has the caller's layout
            for (int gb = cg.kz_lo - 4; gb >= 0 && !found; gb -= 4) {   // kz_lo is a multiple of 32
                double r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = q[cg.deep_sxy * (gb + j + 1)];
                if (w) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) w[cg.deep_sxy * (gb + j + 1)] = r[j];
                }
#pragma unroll
                for (int j = 3; j >= 0; --j) {
                    if (!found) {
                        acc += s_dz[gb + j] * r[j];
                        if (acc >= 23.0) { planes = min(g.nzg, g.nzg - (gb + j) + 1); found = true; }
                    }
                }
            }
        }
    }
    planes = __reduce_max_sync(0xffffffffu, planes);
    if ((threadIdx.x & 31) == 0) atomicMax(acc_max, planes);
    // the last block to finish hands the answer to the host (a plain store into page-locked memory) and resets the scratch
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        const int deepest = atomicMax(acc_max, 0);
        if (out) *out = deepest;
        *acc_max = 0;
        *done = 0u;
        __threadfence_system();
    }
}

// The same bound from the RESIDENT grid (reference layout), for calls that do not take the column form (small calls,
// "column" = 0 / 2): the number of planes the all-reduce moves must not depend on which kernel a rank happened to run
// -- ranks may differ in packet count -- only on the grid and the options every rank shares.  One thread per column,
// neighbouring threads neighbouring x, so every plane is read in coalesced row segments.
__global__ void __launch_bounds__(256) k_column_bound_resident(const DevGrid g, const ColGeom cg, int *__restrict__ acc_max,
                                                               unsigned int *__restrict__ done, int *__restrict__ out)
{
    extern __shared__ double s_dz[];
    stage_column_steps(g, cg.nzp, s_dz);                               // ends with __syncthreads()
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    int planes = 1;
    if (c < cg.tw * cg.th) {
        const int dj = c / cg.tw, di = c - dj * cg.tw;
        const double *q = g.rhokap + ((long long)(cg.i0 + di) + (long long)g.sx * (cg.j0 + dj));
        double acc = 0.;
        planes = g.nzg;
        for (int kz = g.cellk0 - 1; kz >= 0; --kz) {
            acc += s_dz[kz] * q[g.sxy * (kz + 1)];
            if (acc >= 23.0) { planes = min(g.nzg, g.nzg - kz + 1); break; }
        }
    }
    planes = __reduce_max_sync(0xffffffffu, planes);
    if ((threadIdx.x & 31) == 0) atomicMax(acc_max, planes);
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        *out = atomicMax(acc_max, 0);
        *acc_max = 0;
        *done = 0u;
        __threadfence_system();
    }
}

// The tally under the beam's bounding box <-> a dense (tw, th, nzg) buffer.  In the shipped regime every deposit lies
// in those columns, so the all-reduce (mcpolar.f90:173) only has to move them: 18 % of the grid for the reference's
// 0.025 cm spot on a 0.06 cm face.  kUnpack = false: box -> dense; true: dense -> box.
// kz0: only the planes kz >= kz0 travel (k_column_bound: nothing of this call lies deeper).
template <bool kUnpack>
__global__ void __launch_bounds__(256) k_box_copy(const DevGrid g, const ColGeom cg, double *__restrict__ dense, int kz0)
{
    const size_t total = (size_t)cg.tw * cg.th * (g.nzg - kz0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const int di = (int)(t % cg.tw);
        const size_t r = t / cg.tw;
        const int dj = (int)(r % cg.th), kz = kz0 + (int)(r / cg.th);
        const size_t j = (size_t)(cg.i0 - 1 + di) + (size_t)g.nxg * ((size_t)(cg.j0 - 1 + dj) + (size_t)g.nyg * kz);
        if (kUnpack) g.jmean[j] = dense[t];
        else dense[t] = g.jmean[j];
    }
}

// The tally under the beam's bounding box -> the same voxels of `dst`, an array in the tally's own layout: the caller's
// page-locked jmeanGLOBAL, written over PCIe (posted writes, 256-byte row segments).  The rest of jmeanGLOBAL is zero
// fill that travelled while the transport ran (tamc_api.cu) -- and that fill covers the box as well, so a row of the
// box that holds only zeros (every plane below the deepest deposit: most of the box in the shipped regime) is not sent.
__global__ void __launch_bounds__(256) k_box_mirror(const DevGrid g, const ColGeom cg, double *__restrict__ dst)
{
    const int rows = cg.th * g.nzg;
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < rows; r += nwarps) {
        const int dj = r % cg.th, kz = r / cg.th;
        const size_t row = (size_t)(cg.i0 - 1) + (size_t)g.nxg * ((size_t)(cg.j0 - 1 + dj) + (size_t)g.nyg * kz);
        bool any = false;
        for (int di = lane; di < cg.tw; di += 32) any |= (g.jmean[row + di] != 0.);
        if (!__any_sync(0xffffffffu, any)) continue;
        for (int di = lane; di < cg.tw; di += 32) dst[row + di] = g.jmean[row + di];
    }
}

}  // namespace tamc
