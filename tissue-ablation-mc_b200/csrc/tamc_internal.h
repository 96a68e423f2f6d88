// tamc_internal.h -- launcher prototypes shared by the C-ABI layer and the kernel translation units.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

#include "tamc_transport.cuh"
#include "tamc_peer.cuh"

namespace tamc {

struct LaunchCfg {
    int variant;        // 0 = thread-per-packet grid-stride, 1 = persistent warps (default), 2 = variant 0 on the exact arithmetic
    int block;          // threads per CTA
    int ctas_per_sm;    // resident CTAs per SM the grid is sized for (0 = ask the occupancy API)
    int num_sms;
    int chunk;          // persistent: packet ids a warp claims per atomic (0 = auto)
    int scatter_min;    // persistent: run the scattering phase once this many lanes wait for it
    int merge;          // merge consecutive same-voxel deposits in registers (-1 = auto: on with TAMC_SCATTER)
    int min_ctas;       // scattering kernel: the __launch_bounds__ min-CTAs-per-SM build to use (2 or 3)
    int tile;           // stub regime: shared-memory tally tile; -1 = auto, 0 = off, k > 0 = at most k planes
    int column;         // stub regime, column form (tamc_column.cuh): -1 = auto (variant 3, >= 2^20 packets), 0 = off,
                        // 1 = on with the z-fastest copy of the beam's columns, 2 = on, reading the resident grid
    int column_tile;    // column form: shared-memory tiles for the top planes of the deposits / stop counts; -1 = auto
                        // (column_plan in tamc_kernels.cu), 0 = off, 10*ta + tb = force that split
    int column_park;    // tiled column form: regroup the column walk through per-warp queues; 1 = on, 0 = off, -1 = auto:
                        // on when the previous stub-regime call on this handle averaged >= 2.25 voxel-steps per packet
                        // (measured: +7 % at 2.98 steps per packet, -11 % at 1.56, where nearly every packet stops in
                        // its first four-voxel group and there is nothing to regroup)
    double steps_hint;  // voxel-steps per packet of the previous stub-regime call (0 = none yet)
    int gather_depth;   // columns-first upload (tamc_run_optics): planes below the top face that k_column_gather copies over
                        // PCIe ahead of the transport; -1 = auto: depth_hint + max(16, depth_hint / 2), 0 = all, > 0 = that many
                        // (rounded up to whole 32-plane tiles).  A packet that goes deeper reads the caller's grid directly.
    int depth_hint;     // planes from the top face to the deepest stop of the previous column-form call (0 = unknown)
    int want_bound;     // set per call by the C-ABI layer: compute the depth bound of this call (k_column_bound) for the all-reduce
    int flight;         // scatter loop: the flight kernel (tamc_flight.cuh) instead of the work-queue kernel; -1 = auto (on unless
                        // Fresnel / periodic boundaries / the Gaussian beam are selected), 0 = off, 1 = on
    int walk_min;       // flight kernel: the walk phase hands over to the event phase once fewer lanes than this are in flight
    int flight_regs;    // flight kernel: 0 = auto, 2 / 3 / 4 = the 256-thread build for that many CTAs per SM
    int flight_agg;     // flight kernel: 1 = the build with warp-aggregated REDs (match.any + shuffles); default 0 (measured: no gain)
    int flight_launch_min;  // flight kernel: lanes without a packet wait until this many can launch together (1 = at once;
                        // default 3: measured 53.2 / 51.9 / 51.6 / 51.9 / 53.1 ms per 4e7 skin200 packets for 1 / 2 / 3 / 4 / 6)
    int flight_inter;   // flight kernel: interleaved {opacity, tally} voxel records; -1 = auto (grids beyond L2), 0 = off, 1 = on
};

// Which kernel an MC call ran ("form" read-only option of tamc_get_option)
enum { FORM_SIMPLE = 0, FORM_PERSISTENT = 1, FORM_EXACT = 2, FORM_POOL = 3, FORM_TILE = 4, FORM_COLUMN = 5, FORM_COLUMN_RESIDENT = 6, FORM_COLUMN_TILED = 7, FORM_COLUMN_PARKED = 8, FORM_FLIGHT = 9 };

// Device buffers of the column form, owned by the handle and grown on demand by launch_transport.
struct ColumnWorkspace {
    unsigned int *stops = nullptr;   // nxg*nyg*nzg stop counts (plane 0: packets that left through the bottom face); all zero between calls
    size_t stops_elems = 0;
    double *rkT = nullptr;           // z-fastest copy of rhokap under the beam's bounding box
    size_t rkT_elems = 0;
    double *dense = nullptr;         // (tw, th, nzg) staging of the tally under the box for the all-reduce
    size_t dense_elems = 0;
    // columns-first upload (tamc_run_optics), set for one call: k_column_gather reads `gather_src` (device-visible
    // address of the caller's page-locked rhokap) instead of the resident grid, keeps an x-fastest copy of the box in
    // `box_rk` for k_column_finish, and is bracketed by the two events
    const double *gather_src = nullptr;
    double *box_rk = nullptr;
    cudaEvent_t ev_gather0 = nullptr, ev_gather1 = nullptr;
    int last_kz_lo = 0;              // read-back: first plane (k - 1) the last gather copied (> 0: depth-limited)
    // root_io (tamc_api.cu): 1 = this rank gathers every plane and broadcasts both copies of the columns, 2 = this rank
    // receives them instead of gathering; 0 = every rank gathers from its own host array
    int share_gather = 0;
    void *share_comm = nullptr;      // ncclComm_t of the handle
    cudaError_t (*share_fn)(void *comm, double *buf, size_t count, cudaStream_t s) = nullptr;   // broadcast from rank 0 (set by tamc_api.cu)
    // depth bound of the call (k_column_bound): one int in mapped page-locked memory, written by the kernel, read by the
    // host once ev_bound has passed -- while the transport is still running
    int *h_bound = nullptr, *d_bound = nullptr;
    int *bound_scratch = nullptr;    // device: running maximum + finished-block counter of k_column_bound (both zero between calls)
    cudaEvent_t ev_bound = nullptr, ev_gathered = nullptr;
    cudaStream_t s_side = nullptr;   // a side stream, if the handle has one: the bound is computed beside the transport
    bool bound_pending = false;
    // flight kernel (tamc_flight.cuh): the interleaved {rhokap, jmean} voxel records, rebuilt per MC call
    double2 *vox = nullptr;
    size_t vox_elems = 0;
    double *optc = nullptr;          // compact copies of the per-voxel albedo | hgg grids (tamc_set_optics_grids)
    size_t optc_elems = 0;
};

// shipped regime: the columns every deposit lies in, and the copy between them and a dense buffer
bool beam_box(const DevGrid &g, ColGeom &cg);
// true when launch_transport would run the column form on the z-fastest copy (the only form that reads rhokap
// solely through k_column_gather / k_column_finish)
bool column_gather_selected(const DevGrid &g, const LaunchCfg &cfg, long long n);
cudaError_t launch_box_copy(const DevGrid &g, const ColGeom &cg, double *dense, bool unpack, int num_sms, cudaStream_t s, int kz0 = 0);
cudaError_t launch_box_mirror(const DevGrid &g, const ColGeom &cg, double *dst, int num_sms, cudaStream_t s);
// "peer_reduce" (tamc_peer.cuh): out[i] = sum over the ranks of their packed boxes, read from peer memory; *err = mapped host word
cudaError_t launch_peer_box_reduce(const PeerSet &ps, double *out, size_t cnt, int nranks, int rank, unsigned long long call,
                                   unsigned int *err, int num_sms, cudaStream_t s);

// production transport (Philox).  d_rec may be null; when non-null the thread-per-packet kernel is used.
// ws may be null (no column form).
cudaError_t launch_transport(const DevGrid &g, const LaunchCfg &cfg, long long n, uint64_t seed, uint64_t first_id,
                             unsigned long long *d_cnt, tamc_packet_record *d_rec, cudaStream_t s, int *launches,
                             ColumnWorkspace *ws = nullptr, int *form = nullptr);

// trace replay (tamc_replay.cu, compiled with -fmad=false)
cudaError_t launch_replay(const DevGrid &g, long long n, const long long *d_off, const double *d_draws,
                          unsigned long long *d_cnt, tamc_packet_record *d_rec, cudaStream_t s);

// access-pattern-only probe (probe_form: -1 = the form launch_transport would pick, 0 = per-voxel-step stream,
// 1 = column form) and utilities
cudaError_t launch_probe(const DevGrid &g, const LaunchCfg &cfg, long long n, uint64_t seed, unsigned long long *d_cnt,
                         cudaStream_t s, ColumnWorkspace *ws, int probe_form);
cudaError_t launch_selfcheck(const DevGrid &g, long long n, uint64_t seed, uint64_t first_id, unsigned long long *d_out, int num_sms,
                             cudaStream_t s);
cudaError_t launch_fill(double *p, size_t n, double v, int num_sms, cudaStream_t s);
// scatter-regime roofline probe: record the voxel-index stream of n packets (d_trace == nullptr: count the steps per
// packet into d_counts), replay it as loads + REDs on the interleaved voxel records (pack = true: build them instead)
cudaError_t launch_trace(const DevGrid &g, long long n, uint64_t seed, uint64_t first_id, const long long *d_off, int *d_trace,
                         int *d_counts, int num_sms, cudaStream_t s);
cudaError_t launch_probe_trace(const DevGrid &g, const LaunchCfg &cfg, double2 *vox, long long n, const long long *d_off, const int *d_trace,
                               unsigned long long *d_cnt, int num_sms, cudaStream_t s, bool pack);

}  // namespace tamc
