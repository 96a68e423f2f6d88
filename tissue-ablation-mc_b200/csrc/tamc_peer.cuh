// tamc_peer.cuh -- the stub regime's box all-reduce summed straight out of peer memory (option "peer_reduce").
//
// mcpolar.f90:173 (MPI_allREDUCE of jmean) in the shipped regime moves only the tally under the beam's bounding box,
// down to the depth bound of the call: a few MB.  At that size an NCCL all-reduce over eight ranks is all latency, and
// its kernel's channels either slow the one-CTA-per-SM transport kernel or make the reduction slow (measured).
// Here every rank packs its box into a buffer of its own that the other ranks of the node have mapped (CUDA IPC handles
// exchanged once, tamc_api.cu), tells them so with one flag store each, waits for their flags, and sums the nranks
// buffers in rank order -- the same order on every rank, so every rank holds bit-identical sums, as after MPI's
// all-reduce.  NVSwitch gives every GPU full bandwidth to every peer: (nranks - 1) x box bytes of reads per rank.
//
// Buffers alternate between two halves by call parity: a rank can only be one call ahead of the slowest reader of its
// buffer (it passes the flag wait of call e + 1 only after every rank has finished the sum of call e), so the half
// it packs for call e + 2 is no longer being read.  Flags only ever grow (the call number), nothing is reset.
#pragma once

#include <cstdint>

namespace tamc {

constexpr int kPeerMaxRanks = 16;
constexpr size_t kPeerFlagBytes = 4096;        // head of every rank's allocation: the flag words below
// flag words (unsigned long long) of a rank's allocation: [s] = last call rank s has packed, [16 + s] = last call whose sum
// rank s has finished (read by tamc_finalize before the buffer is freed), [32] = finished-block counter of the own kernel
constexpr int kPeerDone = 16, kPeerCounter = 32;

struct PeerSet {
    const double *buf[kPeerMaxRanks];            // this call's half of every rank's buffer (own included)
    unsigned long long *flags[kPeerMaxRanks];    // every rank's flag words
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double2 ld_peer(const double *p)
{
    double2 v;
    asm volatile("ld.volatile.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// out[i] = sum over ranks r = 0 .. nranks-1 of buf[r][i], i < cnt (cnt even-padded buffers: pairs are always readable).
// *err (mapped host memory) is set when a peer's flag does not arrive within timeout_ns: the sums are then not formed.
template <int kRanks>
__global__ void __launch_bounds__(256) k_peer_box_reduce(const PeerSet ps, double *__restrict__ out, size_t cnt, int nranks, int rank,
                                                         unsigned long long call, unsigned long long timeout_ns, unsigned int *err)
{
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            // the pack kernel ran before this one on the stream: its writes are ordered before the flag stores
            __threadfence_system();
            for (int r = 0; r < nranks; ++r)
                if (r != rank) st_release_sys(ps.flags[r] + rank, call);
        }
        int ok = 1;
        const unsigned long long t0 = global_ns();
        for (int r = 0; r < nranks && ok; ++r) {
            if (r == rank) continue;
            while (ld_acquire_sys(ps.flags[rank] + r) < call) {
                if (global_ns() - t0 > timeout_ns) { ok = 0; *err = 1u; __threadfence_system(); break; }
                __nanosleep(200);
            }
        }
        s_ok = ok;
    }
    __syncthreads();
    if (s_ok) {
        const size_t pairs = (cnt + 1) >> 1;
        const size_t stride = (size_t)gridDim.x * blockDim.x;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += stride) {
            double2 v[kRanks > 0 ? kRanks : 1];
            double2 acc = make_double2(0., 0.);
            if (kRanks > 0) {
#pragma unroll
                for (int r = 0; r < kRanks; ++r) v[r] = ld_peer(ps.buf[r] + 2 * i);          // all loads in flight, then rank order
#pragma unroll
                for (int r = 0; r < kRanks; ++r) { acc.x += v[r].x; acc.y += v[r].y; }
            } else {
                for (int r = 0; r < nranks; ++r) { const double2 w = ld_peer(ps.buf[r] + 2 * i); acc.x += w.x; acc.y += w.y; }
            }
            if (2 * i + 1 < cnt) *reinterpret_cast<double2 *>(out + 2 * i) = acc;
            else out[2 * i] = acc.x;
        }
    }
    // the last block to finish tells the other ranks that this rank no longer reads their buffers of this call
    __shared__ bool s_last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(ps.flags[rank] + kPeerCounter, 1ull) == (unsigned long long)gridDim.x - 1ull;
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        ps.flags[rank][kPeerCounter] = 0ull;
        __threadfence_system();
        for (int r = 0; r < nranks; ++r)
            if (r != rank) st_release_sys(ps.flags[r] + kPeerDone + rank, call);
    }
}

}  // namespace tamc
