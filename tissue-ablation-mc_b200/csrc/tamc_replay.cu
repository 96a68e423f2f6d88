// tamc_replay.cu -- trace-replay kernel.  Compiled with -fmad=false so every product and sum is
// rounded separately, like the oracle (-ffp-contract=off) and like a non-FMA build of the Fortran.
//
// Packet p consumes draws[off[p] .. off[p+1]) -- its slice of the reference's sequential ran2
// stream (ran2.f:1-33) -- in the order the reference would have called ran2 (sourceph.f90:28,29,34;
// inttau2.f90:36; albedo test; stokes.f90:48,64), and writes a per-packet record that the tests
// compare with the oracle to 1e-6 relative (BASELINE.json north_star).
#include "tamc_internal.h"

namespace tamc {

__global__ void __launch_bounds__(256) k_transport_replay(const DevGrid g, long long n, const long long *__restrict__ off,
                                                          const double *__restrict__ draws,
                                                          unsigned long long *__restrict__ cnt,
                                                          tamc_packet_record *__restrict__ rec)
{
    extern __shared__ double s_faces[];
    const double *xf, *yf, *zf;
    stage_faces(g, s_faces, xf, yf, zf);

    Counters c;
    c.clear();
    DirectTally tally;
    tally.jm = g.jmean;
    tamc_packet_record scratch;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        ReplayRng rng;
        rng.p = draws;
        rng.pos = off[i];
        rng.end = off[i + 1];
        transport_packet<ReplayRng, DirectTally, true>(g, xf, yf, zf, rng, tally, c, rec ? rec + i : &scratch,
                                                       rng.end - rng.pos);
    }
    c.commit(cnt);
}

cudaError_t launch_replay(const DevGrid &g, long long n, const long long *d_off, const double *d_draws,
                          unsigned long long *d_cnt, tamc_packet_record *d_rec, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    const int block = 128;
    const long long want = (n + block - 1) / block;
    const int grid = (int)(want < 148 * 16 ? want : 148 * 16);
    const size_t smem = sizeof(double) * (size_t)(g.nxg + g.nyg + g.nzg + 3);
    k_transport_replay<<<grid, block, smem, s>>>(g, n, d_off, d_draws, d_cnt, d_rec);
    return cudaGetLastError();
}

}  // namespace tamc
