// tamc_context.h -- the state behind a tamc_handle, shared by the C-ABI translation units.
#pragma once

#include <nccl.h>

#include <cstdint>
#include <string>

#include "tamc_internal.h"

struct tamc_heat;   // device-resident heat / ablation state (tamc_heat.cu)

enum { EV_ZERO0 = 0, EV_K0, EV_K1, EV_AR1, EV_H0, EV_H1, EV_D0, EV_D1, EV_FORK, EV_UP, EV_DN, EV_N };

struct tamc_context {
    // (members are documented where they are used in tamc_api.cu)
    int device = 0;
    int num_sms = 148;
    int nxg = 0, nyg = 0, nzg = 0;
    double xmax = 0, ymax = 0, zmax = 0, delta = 0;
    double spot = 250e-4;               // sourceph.f90:23
    double gauss_sigma = 0.;            // > 0: Gaussian beam (tamc_set_source_gaussian) instead of the CO2 disk
    double albedo = 0, hgg = 0.9, n1 = 1, n2 = 1;
    int flags = 0;
    bool optics_set = false;

    size_t n_rhokap = 0, n_jmean = 0;
    double *d_rhokap = nullptr, *d_jmean = nullptr, *d_faces = nullptr, *d_flush = nullptr;
    double *d_albedo_g = nullptr, *d_hgg_g = nullptr, *d_n_g = nullptr;   // tamc_set_optics_grids: per-voxel optics, null = scalar
    size_t flush_elems = 0;
    unsigned long long *d_cnt = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[EV_N] = {};

    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
    int64_t cursor = 0;

    tamc::LaunchCfg cfg{3, 0, 0, 148, 0, 20, -1, 3, -1, -1, -1, -1, 0., -1, 0, 0, -1, 8, 0, 0, 3, -1};
    tamc::ColumnWorkspace colws;
    int reduce = 1;
    int reduce_bound = 1;   // all-reduce only the planes k_column_bound proves reachable (column form; 0 = every plane of the box)
    int reduce_planes = 0;  // read-only: planes of the box the last all-reduce moved (0 = no box reduce)
    int box_reduce = -1;    // shipped regime: all-reduce only the columns under the beam (-1 = auto, 0 = off, 1 = on)
    int form = -1;          // FORM_* of the last MC call
    int launch32 = 1;       // column form: fp32 first pass for the launch voxel (exact: redone in fp64 near voxel edges); 0 = off
    int probe_form = -1;    // tamc_roofline_probe: -1 = match the transport, 0 = per-voxel-step stream, 1 = column form

    // overlapped boundary copies of the shipped regime (tamc_run / tamc_run_optics, tamc_api.cu)
    int box_io = -1;        // -1 = auto, 0 = off: plain full-grid copies in sequence
    int io_early = 0;       // tamc_run_optics, columns-first upload: bit0 = the full-grid upload, bit1 = the zero fill start beside
                            // the column gather instead of behind it
    int root_io = 0;        // several ranks: 1 = only rank 0's host arrays are read / written (broadcast + root download); set before tamc_comm_init
    bool resident_behind = false;   // root_io: ranks > 0 hold only the beam's columns of the last uploaded grid (sync_resident, tamc_api.cu)
    int *d_path = nullptr;  // root_io: the copy path rank 0 chose for this call, broadcast to the other ranks
    // "peer_reduce": the box all-reduce of the stub regime summed out of peer memory instead of by NCCL (tamc_peer.cuh)
    int peer_reduce = 0;        // 0 = NCCL, 1 = own kernel; set on every rank before tamc_comm_init
    int peer_state = 0;         // 0 = not set up yet, 1 = ready, -1 = unavailable on this node / in this process layout (NCCL is used)
    void *peer_base = nullptr;  // own allocation: flag words + two buffer halves; exported to the other ranks
    void *peer_open[16] = {};   // the other ranks' allocations (cudaIpcOpenMemHandle)
    size_t peer_elems = 0;      // doubles per buffer half (even)
    unsigned long long peer_call = 0;       // calls reduced this way so far (the flag value of the next one is peer_call + 1)
    unsigned int *h_peer_err = nullptr, *d_peer_err = nullptr;   // mapped host word: a peer's flag did not arrive in time
    int io_form = 0;        // read-only: bit0 = the last tamc_run downloaded zero fill + beam columns, bit1 = the last
                            // tamc_run_optics uploaded the beam columns ahead of the grid, bit2 = ... and only down to the
                            // depth the previous call's packets reached (+ margin)
    cudaStream_t s_up = nullptr, s_dn = nullptr;
    double *d_zero = nullptr;           // n_jmean zeros: the tally outside the beam's columns
    double *d_box_rk = nullptr;         // (tw, th, nzg+2) opacities under the beam, uploaded ahead of the full grid
    size_t box_rk_elems = 0;

    tamc_heat *heat = nullptr;

    // bookkeeping of the last call
    int64_t last_launches = 0;
    bool timed_reduce = false, timed_h2d = false, timed_d2h = false, ran = false;
};


// helpers defined in tamc_api.cu
int tamc_fail_(int code, const std::string &msg);
int tamc_check_(tamc_handle h);
tamc::DevGrid tamc_make_grid_(const tamc_context *c);
int tamc_sync_resident_(tamc_handle h);
// defined in tamc_heat.cu
void tamc_heat_release_(tamc_context *c);
