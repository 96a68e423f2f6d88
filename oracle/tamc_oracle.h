/*
 * tamc_oracle.h -- CPU oracle for the photon Monte-Carlo hot path of
 * lewisfish/Tissue-Ablation-MC.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it, and only as the checker / CPU baseline.
 *
 * PARITY: NOT PINNED BY A COMPILED REFERENCE -- the reference ships no tests, golden vectors or fixtures, and it cannot be
 * compiled here (Fortran + mpi_f08, no Fortran front-end or MPI in the image).  This file is a plain-C, fp64,
 * statement-by-statement restatement of the Fortran sources (every function cites the file:line it follows;
 * -freal-4-real-8 promotion semantics from src/Makefile:3, no FMA contraction).  What pins it:
 *   - outputs of the reference's OWN SOURCE TEXT (ran2.f, sourceph.f90, inttau2.f90, stokes.f90, gridset.f90, ch_opt.f90,
 *     statement ranges of mcpolar.f90) executed by a Fortran-subset interpreter (oracle/f90interp.py, itself tested on
 *     known-answer snippets); the vectors are committed with the script that made them
 *     (tests/golden/make_reference_vectors.py -> reference_interp.json.gz) and this oracle reproduces them bit for bit:
 *     2 100 packets of the shipped loop on three ranks, 700 stokes steps, 210 packets of the scatter loop, every voxel of
 *     jmean, the generator state (tests/test_oracle_reference_vectors.py).  An interpreter written for the purpose is not
 *     gfortran: its assumptions are listed in its header;
 *   - the hand-evaluated ran2 / first-packet known answers of SURVEY.md section 8(c) and the Numerical Recipes sequence,
 *     an independent Python transliteration (oracle/pyref.py), the analytic invariants of the shipped regime and, for the
 *     scatter loop, Chandrasekhar's semi-infinite-slab reflectance and van de Hulst's slab (tests/test_oracle_*.py).
 */
#ifndef TAMC_ORACLE_H
#define TAMC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-packet record; binary-identical to tamc_packet_record in include/tamc.h. */
typedef struct {
    double xp, yp, zp;      /* final position, grid-centred coordinates (photon_vars.f90:11) */
    double nxp, nyp, nzp;   /* final direction cosines */
    double deposit;         /* sum of this packet's jmean increments (inttau2.f90:46,53) */
    int32_t xcell, ycell, zcell; /* final voxel, 1-based, -1 = outside grid */
    int32_t steps;          /* voxel-steps = passes of the loop body inttau2.f90:37-63 */
    int32_t nscatt;         /* scattering events */
    int32_t ndraws;         /* uniform draws consumed */
    int32_t fate;           /* 0 = interaction/absorbed, 1..6 = left through -x,+x,-y,+y,-z,+z */
    int32_t flags;          /* bit0: replay draw list exhausted (device only) */
} orc_packet_record;

typedef struct {
    int64_t packets;
    int64_t voxel_steps;
    int64_t scatters;
    int64_t absorbed;       /* packets that ended by interaction (stub) or analog absorption */
    int64_t exits[6];       /* -x,+x,-y,+y,-z,+z */
    int64_t draws;
    double  deposit_sum;
    int64_t specular;       /* ORC_FLAG_FRESNEL: reflected at the top surface before entering (also counted in exits[5]) */
    int64_t internal_reflections;
    int64_t wraps;          /* ORC_FLAG_PERIODIC: lateral re-entries (repeat_bounds) */
} orc_stats;

typedef struct orc_state orc_state;

enum { ORC_RNG_RAN2 = 0, ORC_RNG_PHILOX = 1 };
enum { ORC_FLAG_SCATTER = 1, ORC_FLAG_FRESNEL = 2,
       ORC_FLAG_PERIODIC = 4 /* repeat_bounds (inttau2.f90:242-279, never called upstream) applied to lateral exits */ };

/* Allocate module state for an nxg*nyg*nzg grid and build the face arrays
 * (gridset.f90:23-31); rhokap and jmean start at zero. */
orc_state *orc_create(int nxg, int nyg, int nzg, double xmax, double ymax, double zmax);
void orc_destroy(orc_state *o);

/* constants / opt_prop / iarray accessors */
double *orc_rhokap(orc_state *o);     /* (0:nxg+1,0:nyg+1,0:nzg+1) column-major, with halo */
double *orc_jmean(orc_state *o);      /* (1:nxg,1:nyg,1:nzg) column-major */
double *orc_xface(orc_state *o);
double *orc_yface(orc_state *o);
double *orc_zface(orc_state *o);
double orc_delta(const orc_state *o); /* mcpolar.f90:112 */

/* gridset.f90:33-45: rhokap = 0 everywhere, = kappa in the interior */
void orc_gridset_uniform(orc_state *o, double kappa);
/* ch_opt.f90:15-23 shipped optics: hgg .9, g2, mua 680, mus 0, kappa, albedo; returns kappa */
double orc_init_opt1(orc_state *o);
void orc_set_optics(orc_state *o, double albedo, double hgg);
/* Builder-defined extension (no upstream semantics): refractive indices outside / inside the grid
 * for ORC_FLAG_FRESNEL -- specular reflection at launch, Fresnel reflection or escape at the six outer faces. */
void orc_set_indices(orc_state *o, double n1, double n2);
/* EXTENSION: per-voxel albedo / hgg / refractive index, rhokap's layout; NULL = the scalar (tamc_set_optics_grids) */
void orc_set_grids(orc_state *o, const double *albedo, const double *hgg, const double *n);
void orc_set_spot(orc_state *o, double spot_diameter);   /* sourceph.f90:23, default 250d-4 */
/* Gaussian beam built on rang() (sourceph.f90:73-101: Marsaglia polar method over ranu, :52-70; dead code
 * upstream, so the launch that uses it is builder-defined): xp = rang(0, sigma), yp = rang(0, sigma), each
 * redrawn while it falls off the top face (|xp| >= xmax); everything else as sourcephCO2.  sigma <= 0
 * switches back to the CO2 disk. */
void orc_set_source_gaussian(orc_state *o, double sigma);
void orc_set_flags(orc_state *o, int flags);
void orc_zero_jmean(orc_state *o);

/* RNG selection.  RAN2: the reference generator, seeded per mcpolar.f90:97-98 with rank id.
 * PHILOX: the builder-defined counter-based stream the device uses in production
 * (key = seed, counter = (packet id, draw index / 4)); lets the device result be checked
 * packet-for-packet instead of only statistically. */
void orc_seed_ran2(orc_state *o, int id);
void orc_seed_philox(orc_state *o, uint64_t seed, uint64_t first_packet_id);

/* single-function entry points (unit tests) */
double orc_ran2(orc_state *o);
int orc_ran2_idum(const orc_state *o);
int orc_ran2_idum2(const orc_state *o);
int orc_ran2_iy(const orc_state *o);
int orc_find(double val, const double *a, int n);
/* sourcephCO2 once (sourceph.f90:7-49), then stokes (stokes.f90:6-153) nsteps times on the ran2 stream;
 * out[8*s .. 8*s+7] = nxp nyp nzp cost sint cosp sinp phi after step s. */
void orc_stokes_chain(orc_state *o, int nsteps, double *out);
/* rang (sourceph.f90:73-101) on the ran2 stream; repeat_bounds (inttau2.f90:242-279): returns 0, or -1 where the
 * Fortran stops with 'Error in Repeat_bounds...'.  Both are dead code upstream; the oracle's options call them. */
double orc_rang(orc_state *o, double avg, double sigma);
int orc_repeat_bounds(int *cella, int *cellb, double *acur, double *bcur, double amax, double bmax, int nag, int nbg,
                      double delta);
void orc_philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                       uint32_t out[4]);

/* The photon loop, mcpolar.f90:151-170 (+ the scatter loop of SURVEY 3.3 when ORC_FLAG_SCATTER).
 * Adds to jmean.  records (nullable): nphotons entries.  draws (nullable): receives every uniform
 * draw in order, up to draw_cap; offsets (nullable): nphotons+1 entries, draw index where each
 * packet starts.  Returns 0, or -1 if draw_cap was exceeded (run still completes). */
int orc_run(orc_state *o, int64_t nphotons, orc_packet_record *records, double *draws, int64_t draw_cap,
            int64_t *offsets, orc_stats *stats);

/* R emulated MPI ranks (host threads, private state and jmean, ran2 seeded with id = rank),
 * followed by the in-memory sum standing in for MPI_allREDUCE (mcpolar.f90:173).
 * rhokap_halo: (nxg+2)(nyg+2)(nzg+2); jmean_global: nxg*nyg*nzg, overwritten.
 * seconds (nullable): wall time of photon loops + sum.  Returns threads actually used. */
int orc_run_ranks(int nranks, int nxg, int nyg, int nzg, double xmax, double ymax, double zmax,
                  const double *rhokap_halo, double albedo, double hgg, double spot_diameter, int flags,
                  int64_t nphotons_per_rank, double *jmean_global, orc_stats *stats, double *seconds);

#ifdef __cplusplus
}
#endif
#endif
