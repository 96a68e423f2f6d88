/*
 * tamc.h -- C ABI of libtamc.so: the B200-native photon Monte-Carlo transport that replaces the
 * per-rank photon loop and the jmean reduction of lewisfish/Tissue-Ablation-MC
 * (/root/reference/src/mcpolar.f90:151-173) behind the reference's own call sites.
 *
 * The reference has no plugin/FFI layer: its boundary is a source-level cut in the main program.
 * State crosses it through module globals -- iarray::rhokap, xface/yface/zface (iarray.f90:8-9),
 * opt_prop::albedo,hgg,g2,n1,n2 (opt_prop.f90:5), constants::nxg,nyg,nzg (constants.f90:12), the
 * locals nphotons,xmax,ymax,zmax,delta,iseed,id,numproc -- and comes back as iarray::jmeanGLOBAL.
 * Each entry point below names the reference lines it stands in for.  The Fortran binding
 * (iso_c_binding) and the patched call site are in tissue-ablation-mc_b200/fortran/ and
 * INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a TAMC_E*
 * code (message via tamc_last_error()); nothing throws or exits across the ABI; arrays are
 * Fortran column-major fp64 exactly as the reference allocates them (subs.f90:62-68); the library
 * never keeps a host pointer past the call that received it; there is no CPU fallback -- without a
 * CUDA device every compute entry point fails with TAMC_ENODEVICE.
 */
#ifndef TAMC_H
#define TAMC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAMC_VERSION 105

enum {
    TAMC_OK = 0,
    TAMC_EINVAL = 1,     /* bad argument */
    TAMC_ENODEVICE = 2,  /* no CUDA device / device ordinal out of range */
    TAMC_ECUDA = 3,      /* CUDA runtime error (see tamc_last_error) */
    TAMC_ENCCL = 4,      /* NCCL missing or failed */
    TAMC_ESTATE = 5,     /* call order: optics not set, communicator not initialised, ... */
    TAMC_EREPLAY = 6     /* a replayed packet ran out of draws */
};

/* tamc_set_optics flags */
enum {
    TAMC_SCATTER = 1,    /* run the albedo test + stokes() loop the driver's shell implies
                            (mcpolar.f90:165-169, stokes.f90:6-153) instead of the shipped stub */
    TAMC_FRESNEL = 2,    /* EXTENSION, no upstream semantics (the reference reads n1, n2 and never uses them,
                            mcpolar.f90:84-85, inttau2.f90:125): specular reflection at launch with probability
                            ((n1-n2)/(n1+n2))^2, and unpolarised Fresnel reflection / escape (n2 inside, n1 outside,
                            total internal reflection past the critical angle) at the six outer faces of the grid */
    TAMC_PERIODIC = 4    /* periodic lateral boundaries: repeat_bounds (inttau2.f90:242-279) applied where tauint1 finds
                            xcell or ycell == -1 after a wall crossing (inttau2.f90:57-61) -- the packet re-enters at
                            `delta` / `2*max - delta` on the opposite side and the optical-depth integration goes on;
                            only the top and bottom faces end a flight.  Upstream defines the routine and never calls
                            it, so the call site is this library's.  No effect in the shipped stub
                            regime, whose packets fly straight down.  Replayable (it draws nothing). */
};

typedef struct tamc_context *tamc_handle;

/* One record per packet (replay / validation).  88 bytes, no padding surprises. */
typedef struct {
    double xp, yp, zp;            /* final position, grid-centred (photon_vars.f90:11) */
    double nxp, nyp, nzp;         /* final direction cosines */
    double deposit;               /* sum of the packet's jmean increments (inttau2.f90:46,53) */
    int32_t xcell, ycell, zcell;  /* final voxel, 1-based; -1 = outside (inttau2.f90:208-239) */
    int32_t steps;                /* voxel-steps: passes of the loop body inttau2.f90:37-63 */
    int32_t nscatt;               /* scattering events (mcpolar.f90:28 nscatt) */
    int32_t ndraws;               /* uniform draws consumed */
    int32_t fate;                 /* 0 absorbed/interaction; 1..6 left through -x,+x,-y,+y,-z,+z */
    int32_t flags;                /* bit0: draw list exhausted */
} tamc_packet_record;

/* Counters and device timings of the most recent MC call on this handle (this rank only). */
typedef struct {
    int64_t packets;
    int64_t voxel_steps;
    int64_t scatters;
    int64_t absorbed;             /* ended by interaction (stub) or analog absorption */
    int64_t exits[6];             /* -x,+x,-y,+y,-z(bottom: transmitted),+z(top: reflected) */
    double zero_ms;               /* tally clear */
    double kernel_ms;             /* transport kernel, CUDA events on the library's stream */
    double allreduce_ms;          /* ncclAllReduce of the tally (0 without a communicator) */
    double h2d_ms, d2h_ms;        /* rhokap upload / jmean download inside the last calls */
    int64_t gpu_launches;         /* kernels launched by the library for this call */
    int64_t specular;             /* TAMC_FRESNEL: reflected at the top surface before entering (also in exits[5]) */
    int64_t internal_reflections; /* TAMC_FRESNEL: reflections back into the grid at an outer face */
} tamc_stats;

/* ---- lifecycle ------------------------------------------------------------------------------- */

/* Device-side twin of alloc_array + gridset's face arrays (subs.f90:44-77, gridset.f90:23-31) and
 * of the scalars the loop reads (mcpolar.f90:86-88 xmax..zmax, :112 delta).  `device` is the CUDA
 * ordinal this handle (one MPI rank / one process) drives. */
int tamc_init(int device, int nxg, int nyg, int nzg, double xmax, double ymax, double zmax, double delta,
              tamc_handle *out);
/* Releases device buffers, streams, events and the communicator. */
int tamc_finalize(tamc_handle h);

/* sourceph.f90:23 spotSize (cm).  Default 250d-4. */
int tamc_set_source_co2(tamc_handle h, double spot_diameter_cm);

/* Gaussian beam in place of the CO2 disk: xp = rang(0., sigma), yp = rang(0., sigma) with sourceph.f90:73-101's
 * Marsaglia polar method over ranu (:52-70), each variate redrawn while it misses the top face (|xp| >= xmax);
 * everything else as sourcephCO2 (sourceph.f90:32-47).  Upstream defines rang and never calls it, so this launch is
 * the library's own, checked against the oracle's sourcephGauss.  Production runs take rang's
 * draws from a Philox stream of their own (counter word 3 = 2); trace replay consumes them from the packet's draw
 * list in the reference's order (rang's pairs for x, for y, then phi, then tau).  sigma_cm must be positive;
 * tamc_set_source_co2 switches back to the disk.  The column / tile forms and the beam-box copies of the shipped
 * regime need the disk's bounded footprint and are not used with this source. */
int tamc_set_source_gaussian(tamc_handle h, double sigma_cm);

/* Uploads iarray::rhokap exactly as Fortran holds it -- (0:nxg+1,0:nyg+1,0:nzg+1), column-major,
 * halo included, i.e. pass rhokap(0,0,0) -- plus opt_prop's scalars.  Called once after gridset
 * (gridset.f90:33-45, ch_opt.f90:15-23) and again after every setupThermalCoeff (3dFD.f90:312-361).
 * rhokap == NULL keeps the resident grid and only updates the scalars.  n1/n2 are accepted and
 * stored (the reference reads them, mcpolar.f90:84-85, and never uses them). */
int tamc_set_optics(tamc_handle h, const double *rhokap, double albedo, double hgg, double n1, double n2,
                    int flags);

/* EXTENSION -- no upstream counterpart: the reference's optics are the per-voxel opacity rhokap (iarray.f90:9) plus the
 * SCALARS albedo, hgg, n1, n2 of opt_prop.f90:5.  Optional per-voxel grids, each laid out exactly like rhokap
 * ((0:nxg+1,0:nyg+1,0:nzg+1), column-major, halo included and never read) or NULL to keep the scalar of tamc_set_optics:
 *   albedo  the albedo test of an interaction (`draw < albedo ? stokes : absorbed`) uses the value of the
 *           voxel the interaction happens in;
 *   hgg     ... and the Henyey-Greenstein draw of stokes.f90:48 the anisotropy of that voxel (hgg == 0: the isotropic branch);
 *   n       with TAMC_FRESNEL the inside index at an outer face of the grid is the index of the voxel the packet leaves,
 *           and the specular reflection at launch uses the launch voxel's; index changes BETWEEN voxels do not refract.
 * The grids stay resident (read-only through L2, beside rhokap) until the next call; all three NULL switches them off,
 * and a call with NULL grids gives bit-identical results to a library that never heard of them.  Trace replay, the exact
 * and thread-per-packet kernels and the production (flight) kernel honour albedo / hgg grids; calls with boundary options
 * take the thread-per-packet kernel.  Checked against the extended oracle (orc_set_grids) on the same streams. */
int tamc_set_optics_grids(tamc_handle h, const double *albedo, const double *hgg, const double *n);

/* ---- the hot path ------------------------------------------------------------------------------ */

/* Replaces mcpolar.f90:151-173: runs `nphotons` packets ON THIS RANK (the reference's per-rank
 * `do j = 1, nphotons`), sums the tally over all ranks of the communicator (the MPI_allREDUCE) and
 * writes the UNSCALED sum to jmean_global (nxg*nyg*nzg, column-major); line :174's scaling stays
 * in the driver.  Packets draw from Philox4x32-10 keyed by `seed` with the global packet id as
 * counter; ids are taken from the handle's cursor, which advances by nranks*nphotons per call so
 * repeated calls (the ablation loop) never reuse a stream.  Blocking.  stats may be NULL. */
int tamc_run(tamc_handle h, int64_t nphotons, int64_t seed, double *jmean_global, tamc_stats *stats);

/* One pass of the ablation loop's MC side in a single call: tamc_set_optics(rhokap, ...) followed by
 * tamc_run(...), i.e. mcpolar.f90:151-173 as it runs after every setupThermalCoeff (3dFD.f90:312-361).
 * Same arguments, same results in jmean_global and on the device as the two calls; rhokap == NULL keeps
 * the resident grid.  Having both arrays in one call lets the library overlap the PCIe copies with the
 * transport when they are page-locked (tamc_pin_host) and the scatter loop is off: the tally is zero
 * outside the columns under the beam, so jmean_global is written as a zero fill that runs beside the
 * kernels plus a pitched copy of those columns, and the opacities of those columns are uploaded ahead
 * of the full grid, which follows on a second stream.  From the second such call on, the columns go up only down to the
 * depth the previous call's packets reached plus a margin ("gather_depth": -1 auto, 0 = every plane, n = n planes; the
 * deeper planes a packet of the call can still reach -- optical depth at most 33 ln 2 -- are fetched ahead of the
 * transport by the depth-bound kernel, so the result never depends on the limit), and the download skips the rows that
 * hold only zeros.
 * Options "box_io" (-1 auto, 0 = plain copies in sequence) and read-only "io_form" (bit0: columns-only download, bit1:
 * columns-first upload, bit2: depth-limited), "io_early" (bit0 / bit1: start the full-grid upload / the zero fill beside the
 * column gather instead of behind it; measured slower, default 0) and "depth_hint" (planes from the top face to the deepest stop of the last
 * column-form call).
 * tamc_run alone uses the same download.  In that mode stats->h2d_ms / d2h_ms time only the column copies
 * that are not hidden behind the transport. */
int tamc_run_optics(tamc_handle h, const double *rhokap, double albedo, double hgg, double n1, double n2, int flags,
                    int64_t nphotons, int64_t seed, double *jmean_global, tamc_stats *stats);

/* Same work, split so a driver can overlap it or keep the tally on the device:
 * enqueue tally clear + transport + all-reduce on the handle's stream; first_packet_id < 0 takes
 * (and advances) the cursor, otherwise this rank runs ids [first_packet_id, first_packet_id+n). */
int tamc_run_async(tamc_handle h, int64_t nphotons, int64_t seed, int64_t first_packet_id);
int tamc_sync(tamc_handle h);
int tamc_get_jmean(tamc_handle h, double *jmean_global);
int tamc_get_stats(tamc_handle h, tamc_stats *stats);
int tamc_seek(tamc_handle h, int64_t next_packet_id);

/* Trace replay (validation): packet p consumes draws[draw_offsets[p] .. draw_offsets[p+1]) in the
 * order the reference would call ran2 (sourceph.f90:28,29,34; inttau2.f90:36; scatter loop: albedo
 * test, stokes.f90:48, :64).  fp64, no FMA contraction.  records and jmean are host arrays
 * (npackets records; nxg*nyg*nzg doubles, overwritten); either may be NULL.  No all-reduce. */
int tamc_run_replay(tamc_handle h, int64_t npackets, const int64_t *draw_offsets, const double *draws,
                    tamc_packet_record *records, double *jmean);

/* Production kernel with per-packet records (validation of the Philox path against the oracle
 * running the same counter-based stream).  records: nphotons host entries. */
int tamc_run_records(tamc_handle h, int64_t nphotons, int64_t seed, int64_t first_packet_id,
                     tamc_packet_record *records, double *jmean);

/* ---- multi-GPU: one process (MPI rank) per GPU ------------------------------------------------- */

/* Rank 0 fills a 128-byte id, the driver broadcasts it (MPI_Bcast / torch.distributed), every
 * rank calls tamc_comm_init.  After that tamc_run's reduction is one ncclAllReduce(ncclDouble,
 * ncclSum) of the tally over NVLink on the handle's stream -- mcpolar.f90:173. */
int tamc_comm_unique_id(void *id128);
int tamc_comm_init(tamc_handle h, int nranks, int rank, const void *id128);

/* ---- device residency, tuning, measurement ----------------------------------------------------- */

void *tamc_stream(tamc_handle h);          /* cudaStream_t the library launches on */
double *tamc_jmean_device(tamc_handle h);  /* device tally, nxg*nyg*nzg fp64 */
double *tamc_rhokap_device(tamc_handle h); /* device opacity grid with halo */
/* Page-lock a caller-owned host array once so uploads/downloads run at PCIe speed. */
int tamc_pin_host(void *ptr, uint64_t bytes);
int tamc_unpin_host(void *ptr);
/* Tuning knobs (every default is the measured best; profiles/README.md):
 *   "variant"       0 thread-per-packet, 1 persistent warps, 2 exact arithmetic, 3 = default: with the scatter loop the
 *                   flight kernel (or the work-queue kernel, "flight" = 0); in the shipped stub regime the column form
 *                   for large calls, persistent warps otherwise
 *   "block" (0 = auto), "ctas_per_sm", "chunk", "scatter_min", "merge", "min_ctas"
 *   "flight"        scatter loop: -1 = auto (flight kernel unless Fresnel / periodic boundaries / the Gaussian beam are
 *                   selected), 0 = work-queue kernel, 1 = flight kernel; "walk_min" (1..32, default 8), "flight_regs"
 *                   (0 = auto, 2 / 3 / 4 CTAs per SM), "flight_launch_min" (default 3), "flight_inter" (interleaved
 *                   {opacity, tally} records: -1 = auto, beyond L2 only), "flight_agg" (1 = warp-aggregated REDs; slower)
 *   "tile" / "column"   stub regime: -1 = auto, 0 = off, > 0 = force; "column_tile" (shared-memory tiles of the column
 *                   form: -1 = auto, 0 = off, 10*ta + tb = ta planes of deposits and tb planes of stop counts);
 *                   "column_park" (regrouped column walk of a tiled call: -1 = auto from the previous call's voxel-steps
 *                   per packet, 0 = off, 1 = on); "launch32" (fp32 first pass of the launch voxel, 1 = on)
 *   "gather_depth"  tamc_run_optics, columns-first upload: planes below the top face copied ahead of the transport
 *                   (-1 = auto from the previous call, 0 = all); deeper planes a packet can reach follow through
 *                   k_column_bound, so the result never depends on it; "io_early"; "box_io" (-1 = auto, 0 = plain copies)
 *   "reduce" (0 = skip the all-reduce), "box_reduce" (-1 = auto, 0 = all-reduce the whole grid), "reduce_bound" (all-reduce
 *                   only the planes of the box that a packet of the call can reach -- the optical depth of a packet is at
 *                   most 33 ln 2, so every rank derives the same bound from its copy of the grid; 1 = on, 0 = every
 *                   plane, 2 = compute it even without a communicator)
 *   "root_io"       several ranks: 1 = only rank 0's host arrays are read / written; before tamc_comm_init
 *   "peer_reduce"   several ranks on one node: 1 = the stub regime's box all-reduce is summed straight out of the other
 *                   ranks' buffers (CUDA IPC, NVLink) by the library's own kernel instead of ncclAllReduce; falls back to
 *                   NCCL (every rank alike) where the buffers cannot be mapped; before tamc_comm_init; default 0
 *   "probe_form"    tamc_roofline_probe: -1 = the form the transport would take, 0 = per-voxel-step address stream, 1 =
 *                   column-form address stream
 * Read-only: "form" = the kernel the last MC call ran (0 thread-per-packet, 1 persistent, 2 exact, 3 work-queue pool,
 * 4 tile, 5 column, 6 column on the resident grid, 7 column with shared-memory tiles, 8 the same with the regrouped walk,
 * 9 flight kernel), "io_form" (see tamc_run_optics), "reduce_planes" (planes of the box the last all-reduce moved),
 * "depth_hint", "peer_state" (1 = peer_reduce in use, -1 = not available here). */
int tamc_set_option(tamc_handle h, const char *name, int64_t value);
int64_t tamc_get_option(tamc_handle h, const char *name);
/* Access-pattern-only kernel: the tally/grid address stream of `nphotons` straight-down packets
 * with no transport arithmetic (per voxel-step one rhokap load + one jmean RED; in the column form
 * per packet one 256-bit load per four voxels + at most two REDs); ms receives its device time,
 * steps the voxel-steps it stands for. */
int tamc_roofline_probe(tamc_handle h, int64_t nphotons, int64_t seed, double *ms, int64_t *steps);
/* Roofline probe of the scatter regime: records the voxel-index stream of `npackets` real packets of the current optics
 * (production arithmetic, disk source, no boundary options) and replays it as memory operations only -- per voxel-step one
 * 8-byte load of the voxel's opacity, per voxel left one fp64 RED into its tally, on the interleaved {rhokap, jmean}
 * records the flight kernel uses -- i.e. the grid-lookup / L2-atomic rate this address stream admits with no transport
 * arithmetic at all.  ms = device time of the replay (best of 3), steps / reds = loads / REDs it issued. */
int tamc_trace_probe(tamc_handle h, int64_t npackets, int64_t seed, double *ms, int64_t *steps, int64_t *reds);
/* The column form takes a packet's launch voxel from an fp32 first pass (hardware sqrt / sin / cos) and redoes it in the
 * production fp64 arithmetic (sourceph.f90:28-31,45-46) whenever the point lies within a proven error bound of a voxel
 * edge, so the voxel is always the fp64 one ("launch32" = 0 switches the first pass off).  This runs both passes over
 * the first n Philox blocks of `seed` and counts the draws handed to fp64 and the draws where the fp32 pass kept a
 * voxel that differs (must be 0). */
int tamc_selfcheck_launch(tamc_handle h, int64_t n, int64_t seed, int64_t *fallbacks, int64_t *mismatches);
/* Writes >= bytes of device memory to evict L2 between timed steps. */
int tamc_flush_l2(tamc_handle h, uint64_t bytes);

/* ---- next to the hot path: the heat / ablation step on the device ------- */

/* The run-time parameters the Heat module takes from res/input.params (mcpolar.f90:85-94). */
typedef struct {
    double power;             /* W */
    double energyPerPixel;    /* mJ */
    double total_time;        /* s (overridden for gaussian pulses, mcpolar.f90:134-137) */
    double repetitionRate_1;  /* Hz */
    double ablateTemp;        /* C */
    int32_t loops;            /* heat sub-steps per MC call */
    int32_t pulsesToDo;
    int32_t pulsetype;        /* 0 tophat, 1 gaussian, 2 triangular (3dFD.f90:294-307) */
    int32_t pad_;
} tamc_heat_params;

enum {   /* tamc_heat_array ids; arrays are Fortran column-major fp64 as the reference allocates them */
    TAMC_HEAT_TEMP = 0, TAMC_HEAT_RHOKAP, TAMC_HEAT_KAPPA, TAMC_HEAT_DENSITY, TAMC_HEAT_HEATCAP, TAMC_HEAT_COEFF,
    TAMC_HEAT_ALPHA,          /* the seven above: (0:n+1)^3 */
    TAMC_HEAT_WATER, TAMC_HEAT_Q, TAMC_HEAT_TISSUE,   /* n^3 */
    TAMC_HEAT_THRESTIME,      /* (n,n,n,3) */
    TAMC_HEAT_JMEAN           /* n^3: the resident, unscaled tally of the last MC call */
};
enum {   /* tamc_heat_scalar ids */
    TAMC_HEAT_S_DELT = 0, TAMC_HEAT_S_TIME, TAMC_HEAT_S_TOTAL_TIME, TAMC_HEAT_S_PULSELENGTH, TAMC_HEAT_S_REALPULSELENGTH,
    TAMC_HEAT_S_LASERON, TAMC_HEAT_S_PULSECOUNT, TAMC_HEAT_S_REPETITIONCOUNT, TAMC_HEAT_S_LASER_FLAG, TAMC_HEAT_S_QVAPOR,
    TAMC_HEAT_S_PWR, TAMC_HEAT_S_COUNTER,
    TAMC_HEAT_S_NEGATIVE_TEMP /* 1 once a temperature went negative: where the reference calls mpi_abort (3dFD.f90:179-182) */
};

/* initThermalCoeff (3dFD.f90:233-309) + the driver's temperature boundary set-up and total_time
 * override (mcpolar.f90:65-71,123-140), on device arrays.  Needs nxg = nyg = nzg and a resident
 * rhokap (tamc_set_optics).  delt receives the time step (may be NULL). */
int tamc_heat_init(tamc_handle h, const tamc_heat_params *p, double *delt);
/* mcpolar.f90:174-182 on the resident arrays: the tally left by the last MC call, scaled on the fly, heat_sim_3D
 * (3dFD.f90:21-230, single-rank semantics), Arrhenius (:424-466), setupThermalCoeff (:312-361, which
 * rewrites the resident rhokap for the next MC call). */
/* Asynchronous (enqueued on the handle's stream, nothing is read back): a driver that steps with tamc_run_async +
 * tamc_heat_step must poll tamc_heat_scalar(TAMC_HEAT_S_NEGATIVE_TEMP) for the reference's abort condition
 * (tamc_coupled_loop does, and returns an error).  The step swaps the resident opacity buffer with its successor: a pointer
 * obtained from tamc_rhokap_device() before the call refers to the PREVIOUS opacity afterwards -- query it again. */
int tamc_heat_step(tamc_handle h, int64_t nphotons_times_numproc);
/* The whole `do while(time <= total_time)` loop (mcpolar.f90:148-186): MC call, all-reduce, heat step,
 * property update, with nothing crossing PCIe.  max_iterations < 0 = run to total_time. */
int tamc_coupled_loop(tamc_handle h, int64_t nphotons, int64_t seed, int64_t max_iterations, int64_t *iterations_done,
                      int64_t *packets_done);
/* Download (upload = 0) or upload (upload = 1) one of the heat arrays / read one of its scalars. */
int tamc_heat_array(tamc_handle h, int which, double *host, int upload);
int tamc_heat_scalar(tamc_handle h, int which, double *out);

const char *tamc_last_error(void);
int tamc_version(void);
int tamc_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
