/*
 * tamc_oracle.c -- plain-C fp64 restatement of the reference's photon Monte-Carlo hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see tamc_oracle.h).  PARITY NOT PINNED BY A COMPILED REFERENCE: no golden
 * vectors exist upstream and the Fortran cannot be built here; pinned instead, bit for bit, by outputs of the
 * reference's own source text executed by oracle/f90interp.py (tests/golden/reference_interp.json.gz), by the KATs in
 * tests/golden/ and by oracle/pyref.py.
 *
 * Conventions that matter for trace-replay parity:
 *  - every Fortran `real` is a double (src/Makefile:3, -freal-4-real-8), literals included;
 *  - PI / TWOPI are the truncated 7-digit constants of src/constants.f90:13;
 *  - expressions are evaluated left to right exactly as written; build with -ffp-contract=off;
 *  - arrays keep the Fortran layout: rhokap(0:nxg+1,0:nyg+1,0:nzg+1), jmean(nxg,nyg,nzg),
 *    faces 1-based (src/iarray.f90:8-10, src/subs.f90:62-68).
 * Variable names follow the Fortran so the two can be read side by side.
 */
#include "tamc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <pthread.h>
#include <unistd.h>

/* src/constants.f90:13 -- truncated on purpose */
static const double PI = 3.141592;
static const double TWOPI = 6.283185;

struct orc_state {
    /* constants.f90:12 (compile-time parameters upstream, run-time here) */
    int nxg, nyg, nzg;
    /* mcpolar.f90 locals */
    double xmax, ymax, zmax, delta;
    int iseed;
    /* iarray.f90:8-10 */
    double *xface, *yface, *zface;
    double *rhokap, *jmean;
    /* opt_prop.f90:5 */
    double mua, mus, g2, hgg, kappa, albedo, mu_water, mu_protein, n1, n2;
    /* EXTENSION (no upstream counterpart: opt_prop.f90:5 holds scalars): optional per-voxel albedo / hgg / refractive
     * index, laid out like rhokap (0:nxg+1,0:nyg+1,0:nzg+1); NULL = the scalar above */
    double *albedo_g, *hgg_g, *n_g;
    /* photon_vars.f90:11 */
    double xp, yp, zp, nxp, nyp, nzp, sint, cost, sinp, cosp, phi;
    /* ran2.f:8-9 SAVEd state */
    int iv[32], iy, idum2;
    /* sourceph.f90:23 */
    double spotSize;
    double gauss_sigma;         /* > 0: Gaussian beam through rang() instead of the CO2 disk */
    /* oracle plumbing */
    int flags;
    int rng_mode;
    uint64_t ph_seed, ph_packet;
    uint32_t ph_block[4];
    int64_t ph_block_idx;       /* which block of the packet's main stream ph_block holds (-1 = none) */
    int64_t pkt_sdraws;         /* source-stream draws consumed by the current packet (Philox: counter word 3 = 2) */
    int64_t pkt_draws;          /* draws consumed by the current packet */
    double *draw_log;
    int64_t draw_cap, draw_n;
    int draw_overflow;
    /* per-packet counters */
    int32_t steps, nscatt;
    double deposit;
    int64_t pkt_bdraws;         /* boundary-stream draws consumed by the current packet */
    int64_t internal_reflections;
    int64_t wraps;              /* ORC_FLAG_PERIODIC: lateral re-entries */
};

#define RHOKAP(o, i, j, k) ((o)->rhokap[(size_t)(i) + (size_t)((o)->nxg + 2) * ((size_t)(j) + (size_t)((o)->nyg + 2) * (size_t)(k))])
#define HALO(o, i, j, k) ((size_t)(i) + (size_t)((o)->nxg + 2) * ((size_t)(j) + (size_t)((o)->nyg + 2) * (size_t)(k)))
#define JMEAN(o, i, j, k) ((o)->jmean[(size_t)((i)-1) + (size_t)(o)->nxg * ((size_t)((j)-1) + (size_t)(o)->nyg * (size_t)((k)-1))])
#define XFACE(o, i) ((o)->xface[(i)-1])
#define YFACE(o, i) ((o)->yface[(i)-1])
#define ZFACE(o, i) ((o)->zface[(i)-1])

/* ------------------------------------------------------------------ set-up */

orc_state *orc_create(int nxg, int nyg, int nzg, double xmax, double ymax, double zmax)
{
    orc_state *o = (orc_state *)calloc(1, sizeof(*o));
    int i;
    if (!o) return NULL;
    o->nxg = nxg; o->nyg = nyg; o->nzg = nzg;
    o->xmax = xmax; o->ymax = ymax; o->zmax = zmax;
    o->xface = (double *)calloc((size_t)nxg + 1, sizeof(double));
    o->yface = (double *)calloc((size_t)nyg + 1, sizeof(double));
    o->zface = (double *)calloc((size_t)nzg + 1, sizeof(double));
    o->rhokap = (double *)calloc((size_t)(nxg + 2) * (nyg + 2) * (nzg + 2), sizeof(double));
    o->jmean = (double *)calloc((size_t)nxg * nyg * nzg, sizeof(double));
    if (!o->xface || !o->yface || !o->zface || !o->rhokap || !o->jmean) { orc_destroy(o); return NULL; }
    /* gridset.f90:23-31: face(i) = (i-1) * 2. * max/n, evaluated left to right */
    for (i = 1; i <= nxg + 1; i++) XFACE(o, i) = (double)(i - 1) * 2. * xmax / (double)nxg;
    for (i = 1; i <= nyg + 1; i++) YFACE(o, i) = (double)(i - 1) * 2. * ymax / (double)nyg;
    for (i = 1; i <= nzg + 1; i++) ZFACE(o, i) = (double)(i - 1) * 2. * zmax / (double)nzg;
    /* mcpolar.f90:112 */
    o->delta = 1.e-8 * (2. * zmax / (double)nzg);
    o->spotSize = 250e-4; /* sourceph.f90:23 */
    o->idum2 = 123456789; /* ran2.f:9 DATA */
    o->iy = 0;
    o->iseed = -95648324;
    o->rng_mode = ORC_RNG_RAN2;
    o->n1 = 1.; o->n2 = 1.;
    orc_init_opt1(o);
    return o;
}

void orc_destroy(orc_state *o)
{
    if (!o) return;
    free(o->xface); free(o->yface); free(o->zface); free(o->rhokap); free(o->jmean);
    free(o->albedo_g); free(o->hgg_g); free(o->n_g);
    free(o);
}

double *orc_rhokap(orc_state *o) { return o->rhokap; }
double *orc_jmean(orc_state *o) { return o->jmean; }
double *orc_xface(orc_state *o) { return o->xface; }
double *orc_yface(orc_state *o) { return o->yface; }
double *orc_zface(orc_state *o) { return o->zface; }
double orc_delta(const orc_state *o) { return o->delta; }

/* gridset.f90:33-45 */
void orc_gridset_uniform(orc_state *o, double kappa)
{
    int i, j, k;
    memset(o->rhokap, 0, sizeof(double) * (size_t)(o->nxg + 2) * (o->nyg + 2) * (o->nzg + 2));
    for (i = 1; i <= o->nxg; i++)
        for (j = 1; j <= o->nyg; j++)
            for (k = 1; k <= o->nzg; k++) RHOKAP(o, i, j, k) = kappa;
}

/* ch_opt.f90:15-23 */
double orc_init_opt1(orc_state *o)
{
    o->hgg = 0.9;
    o->g2 = o->hgg * o->hgg; /* hgg**2. */
    o->mu_water = 510.;
    o->mu_protein = 170.;
    o->mua = o->mu_water + o->mu_protein;
    o->mus = 0.;
    o->kappa = o->mus + o->mua;
    o->albedo = o->mus / o->kappa;
    return o->kappa;
}

void orc_set_optics(orc_state *o, double albedo, double hgg)
{
    o->albedo = albedo;
    o->hgg = hgg;
    o->g2 = hgg * hgg; /* ch_opt.f90:16 */
}

void orc_set_spot(orc_state *o, double d) { o->spotSize = d; }
void orc_set_source_gaussian(orc_state *o, double sigma) { o->gauss_sigma = sigma > 0. ? sigma : 0.; }
void orc_set_indices(orc_state *o, double n1, double n2) { o->n1 = n1; o->n2 = n2; }

/* EXTENSION: per-voxel albedo / hgg / refractive index (each NULL = keep the scalar; all NULL = grids off).  Builder-defined
 * semantics, the library's tamc_set_optics_grids: the albedo test and the Henyey-Greenstein draw of an interaction take the
 * values of the voxel the interaction happens in; with ORC_FLAG_FRESNEL the inside index at an outer face is the index of
 * the voxel the packet leaves (and of the launch voxel for the specular reflection).  Index changes BETWEEN voxels do not
 * refract (documented limitation). */
void orc_set_grids(orc_state *o, const double *albedo, const double *hgg, const double *n)
{
    const size_t nh = (size_t)(o->nxg + 2) * (size_t)(o->nyg + 2) * (size_t)(o->nzg + 2);
    const double *src[3] = {albedo, hgg, n};
    double **dst[3] = {&o->albedo_g, &o->hgg_g, &o->n_g};
    int i;
    for (i = 0; i < 3; i++) {
        free(*dst[i]);
        *dst[i] = NULL;
        if (src[i]) {
            *dst[i] = (double *)malloc(nh * sizeof(double));
            memcpy(*dst[i], src[i], nh * sizeof(double));
        }
    }
}
void orc_set_flags(orc_state *o, int flags) { o->flags = flags; }
void orc_zero_jmean(orc_state *o) { memset(o->jmean, 0, sizeof(double) * (size_t)o->nxg * o->nyg * o->nzg); }

/* ------------------------------------------------------------------ RNGs */

/* mcpolar.f90:97-98, plus a fresh process image for ran2's SAVEd variables (ran2.f:9) */
void orc_seed_ran2(orc_state *o, int id)
{
    int iseed = -95648324 + id;
    iseed = -abs(iseed);
    o->iseed = iseed;
    o->idum2 = 123456789;
    memset(o->iv, 0, sizeof(o->iv));
    o->iy = 0;
    o->rng_mode = ORC_RNG_RAN2;
}

void orc_seed_philox(orc_state *o, uint64_t seed, uint64_t first_packet_id)
{
    o->ph_seed = seed;
    o->ph_packet = first_packet_id;
    o->rng_mode = ORC_RNG_PHILOX;
}

/* ran2.f:1-33 */
static double ran2(orc_state *o, int *idum)
{
    enum { IM1 = 2147483563, IM2 = 2147483399, IMM1 = IM1 - 1, IA1 = 40014, IA2 = 40692, IQ1 = 53668,
           IQ2 = 52774, IR1 = 12211, IR2 = 3791, NTAB = 32, NDIV = 1 + IMM1 / NTAB };
    const double AM = 1. / (double)IM1, EPS = 1.2e-7, RNMX = 1. - EPS;
    int j, k;
    double r;

    if (*idum <= 0) {
        *idum = (-(*idum) > 1) ? -(*idum) : 1;
        o->idum2 = *idum;
        for (j = NTAB + 8; j >= 1; j--) {
            k = *idum / IQ1;
            *idum = IA1 * (*idum - k * IQ1) - k * IR1;
            if (*idum < 0) *idum = *idum + IM1;
            if (j <= NTAB) o->iv[j - 1] = *idum;
        }
        o->iy = o->iv[0];
    }
    k = *idum / IQ1;
    *idum = IA1 * (*idum - k * IQ1) - k * IR1;
    if (*idum < 0) *idum = *idum + IM1;
    k = o->idum2 / IQ2;
    o->idum2 = IA2 * (o->idum2 - k * IQ2) - k * IR2;
    if (o->idum2 < 0) o->idum2 = o->idum2 + IM2;
    j = 1 + o->iy / NDIV;
    o->iy = o->iv[j - 1] - o->idum2;
    o->iv[j - 1] = *idum;
    if (o->iy < 1) o->iy = o->iy + IMM1;
    r = AM * (double)o->iy;
    return (r < RNMX) ? r : RNMX;
}

/* Philox4x32-10 (Salmon et al., SC'11): the device's production generator, restated. */
void orc_philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                       uint32_t out[4])
{
    int r;
    for (r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* one uniform draw from whichever generator is selected; logged for replay */
static double draw(orc_state *o)
{
    double r;
    if (o->rng_mode == ORC_RNG_RAN2) {
        r = ran2(o, &o->iseed);
    } else {
        int lane = (int)(o->pkt_draws & 3);
        if (o->ph_block_idx != (o->pkt_draws >> 2)) {
            o->ph_block_idx = o->pkt_draws >> 2;
            orc_philox4x32_10((uint32_t)o->ph_seed, (uint32_t)(o->ph_seed >> 32), (uint32_t)o->ph_packet,
                              (uint32_t)(o->ph_packet >> 32), (uint32_t)o->ph_block_idx, 0u, o->ph_block);
        }
        r = ((double)o->ph_block[lane] + 0.5) * (1.0 / 4294967296.0);
    }
    o->pkt_draws++;
    if (o->draw_log) {
        if (o->draw_n < o->draw_cap) o->draw_log[o->draw_n] = r;
        else o->draw_overflow = 1;
    }
    o->draw_n++;
    return r;
}

/* ---- builder-defined extension: Fresnel boundaries (ORC_FLAG_FRESNEL) -------------------------
 * The reference reads n1, n2 (mcpolar.f90:84-85) and never uses them; the only trace of boundary optics is a
 * comment (inttau2.f90:125).  What the north-star asks for is specified here, default off:
 *   - at launch the packet is specularly reflected with probability ((n1-n2)/(n1+n2))^2 (normal incidence);
 *   - when a wall crossing would take the packet out of the grid through a face, it is reflected back with the
 *     unpolarised Fresnel reflectance for n2 -> n1 at its angle of incidence (total internal reflection
 *     beyond the critical angle): the normal direction cosine changes sign, the position is snapped to
 *     `face -+ delta` INSIDE the grid, the optical-depth integration simply continues.
 * Boundary decisions draw from their own stream (ran2: the same sequential generator; Philox: counter word 3
 * = 1), so switching the flag on with n1 == n2 changes nothing in the packets' paths. */
static double draw_boundary(orc_state *o)
{
    if (o->rng_mode == ORC_RNG_RAN2) return ran2(o, &o->iseed);
    {
        uint32_t b[4];
        orc_philox4x32_10((uint32_t)o->ph_seed, (uint32_t)(o->ph_seed >> 32), (uint32_t)o->ph_packet,
                          (uint32_t)(o->ph_packet >> 32), (uint32_t)(o->pkt_bdraws >> 2), 1u, b);
        return ((double)b[o->pkt_bdraws++ & 3] + 0.5) * (1.0 / 4294967296.0);
    }
}

static double fresnel_reflectance(double n_in, double n_out, double ci)
{
    const double ratio = n_in / n_out;
    const double si2 = (ratio * ratio) * (1. - ci * ci);
    double ct, rs, rp;
    if (si2 >= 1.) return 1.;
    ct = sqrt(1. - si2);
    rs = (n_in * ci - n_out * ct) / (n_in * ci + n_out * ct);
    rp = (n_in * ct - n_out * ci) / (n_in * ct + n_out * ci);
    return 0.5 * (rs * rs + rp * rp);
}

double orc_ran2(orc_state *o) { return ran2(o, &o->iseed); }
int orc_ran2_idum(const orc_state *o) { return o->iseed; }
int orc_ran2_idum2(const orc_state *o) { return o->idum2; }
int orc_ran2_iy(const orc_state *o) { return o->iy; }

/* ------------------------------------------------------------------ source */

/* sourceph.f90:7-49 */
static void sourcephCO2(orc_state *o, double xmax, double ymax, double zmax, int *xcell, int *ycell, int *zcell)
{
    double theta, r;
    const double spotSize = o->spotSize;

    r = draw(o) * ((spotSize / 2.) * (spotSize / 2.));
    theta = draw(o) * TWOPI;
    o->xp = sqrt(r) * cos(theta);
    o->yp = sqrt(r) * sin(theta);
    o->zp = zmax - (1.e-8 * (2. * zmax / (double)o->nzg));

    o->phi = TWOPI * draw(o);
    o->cosp = cos(o->phi);
    o->sinp = sin(o->phi);
    o->sint = 0.;
    o->cost = -1.;

    o->nxp = o->sint * o->cosp;
    o->nyp = o->sint * o->sinp;
    o->nzp = o->cost;

    *xcell = (int)((double)o->nxg * (o->xp + xmax) / (2. * xmax)) + 1;
    *ycell = (int)((double)o->nyg * (o->yp + ymax) / (2. * ymax)) + 1;
    *zcell = (int)((double)o->nzg * (o->zp + zmax) / (2. * zmax)) + 1;
}

/* Draws of the Gaussian source.  ran2: the one sequential generator (logged, so replay sees them in order).
 * Philox: a stream of their own (counter word 3 = 2, draw index / 4 in word 2) because the polar method consumes a
 * variable number of draws, and the packet's main stream keeps its fixed layout of one block per event. */
static double draw_source(orc_state *o)
{
    if (o->rng_mode == ORC_RNG_RAN2) return draw(o);
    {
        uint32_t b[4];
        orc_philox4x32_10((uint32_t)o->ph_seed, (uint32_t)(o->ph_seed >> 32), (uint32_t)o->ph_packet,
                          (uint32_t)(o->ph_packet >> 32), (uint32_t)(o->pkt_sdraws >> 2), 2u, b);
        return ((double)b[o->pkt_sdraws++ & 3] + 0.5) * (1.0 / 4294967296.0);
    }
}

/* sourceph.f90:52-70 */
static double ranu(orc_state *o, double a, double b)
{
    return a + draw_source(o) * (b - a);
}

/* sourceph.f90:73-101 (Marsaglia polar method; only the first variate of the pair is used) */
static double rang(orc_state *o, double avg, double sigma)
{
    double u = 0., s, tmp;

    s = 1.;
    while (s >= 1.) {
        u = ranu(o, -1., 1.);
        s = ranu(o, -1., 1.);
        s = s * s + u * u;       /* s**2. + u**2. */
    }
    tmp = u * sqrt(-2. * log(s) / s);

    return avg + sigma * tmp;
}

/* Builder-defined launch on top of rang() -- the reference defines rang and never calls it.  Everything but the
 * entry point is sourcephCO2 (sourceph.f90:32-47); a variate that misses the top face is redrawn, so every packet
 * enters the grid and the cell formula stays in range. */
static void sourcephGauss(orc_state *o, double xmax, double ymax, double zmax, int *xcell, int *ycell, int *zcell)
{
    const double sigma = o->gauss_sigma;

    do { o->xp = rang(o, 0., sigma); } while (!(fabs(o->xp) < xmax));
    do { o->yp = rang(o, 0., sigma); } while (!(fabs(o->yp) < ymax));
    o->zp = zmax - (1.e-8 * (2. * zmax / (double)o->nzg));

    if (o->rng_mode == ORC_RNG_PHILOX) o->pkt_draws = 2;   /* phi and tau stay words 2 and 3 of the packet's block 0 */
    o->phi = TWOPI * draw(o);
    o->cosp = cos(o->phi);
    o->sinp = sin(o->phi);
    o->sint = 0.;
    o->cost = -1.;

    o->nxp = o->sint * o->cosp;
    o->nyp = o->sint * o->sinp;
    o->nzp = o->cost;

    *xcell = (int)((double)o->nxg * (o->xp + xmax) / (2. * xmax)) + 1;
    *ycell = (int)((double)o->nyg * (o->yp + ymax) / (2. * ymax)) + 1;
    *zcell = (int)((double)o->nzg * (o->zp + zmax) / (2. * zmax)) + 1;
    if (*xcell > o->nxg) *xcell = o->nxg;   /* (xp + xmax) rounded up to 2 xmax: one ulp from the edge */
    if (*ycell > o->nyg) *ycell = o->nyg;
}

/* ------------------------------------------------------------------ tauint1 and helpers */

/* inttau2.f90:208-239 */
int orc_find(double val, const double *a, int n)
{
    int lo = 0, hi = n + 1, mid;
    if (val == a[0]) return 1;
    if (val == a[n - 1]) return n - 1;
    if (val > a[n - 1] || val < a[0]) return -1;
    for (;;) {
        if (hi - lo <= 1) break;
        mid = (hi + lo) / 2;
        if (val >= a[mid - 1]) lo = mid;
        else hi = mid;
    }
    return lo;
}

/* inttau2.f90:75-121 */
static double wall_dist(orc_state *o, int celli, int cellj, int cellk, double xcur, double ycur, double zcur, int dir[3])
{
    double dx = 0., dy = 0., dz = 0., wd;

    if (o->nxp > 0.) dx = (XFACE(o, celli + 1) - xcur) / o->nxp;
    else if (o->nxp < 0.) dx = (XFACE(o, celli) - xcur) / o->nxp;
    else if (o->nxp == 0.) dx = 100000.;

    if (o->nyp > 0.) dy = (YFACE(o, cellj + 1) - ycur) / o->nyp;
    else if (o->nyp < 0.) dy = (YFACE(o, cellj) - ycur) / o->nyp;
    else if (o->nyp == 0.) dy = 100000.;

    if (o->nzp > 0.) dz = (ZFACE(o, cellk + 1) - zcur) / o->nzp;
    else if (o->nzp < 0.) dz = (ZFACE(o, cellk) - zcur) / o->nzp;
    else if (o->nzp == 0.) dz = 100000.;

    wd = dx < dy ? dx : dy;       /* min(dx,dy,dz) */
    wd = wd < dz ? wd : dz;
    /* later axis wins ties: inttau2.f90:116-118 */
    if (wd == dx) { dir[0] = 1; dir[1] = 0; dir[2] = 0; }
    if (wd == dy) { dir[0] = 0; dir[1] = 1; dir[2] = 0; }
    if (wd == dz) { dir[0] = 0; dir[1] = 0; dir[2] = 1; }
    return wd;
}

/* inttau2.f90:190-205 */
static void update_voxels(orc_state *o, double xcur, double ycur, double zcur, int *celli, int *cellj, int *cellk)
{
    *celli = orc_find(xcur, o->xface, o->nxg + 1);
    *cellj = orc_find(ycur, o->yface, o->nyg + 1);
    *cellk = orc_find(zcur, o->zface, o->nzg + 1);
}

/* inttau2.f90:124-187 */
static void update_pos(orc_state *o, double *xcur, double *ycur, double *zcur, int *celli, int *cellj, int *cellk,
                       double dcell, int wall_flag, const int dir[3], double delta)
{
    if (wall_flag) {
        if (dir[0]) {
            if (o->nxp > 0.) *xcur = XFACE(o, *celli + 1) + delta;
            else if (o->nxp < 0.) *xcur = XFACE(o, *celli) - delta;
            *ycur = *ycur + o->nyp * dcell;
            *zcur = *zcur + o->nzp * dcell;
        } else if (dir[1]) {
            *xcur = *xcur + o->nxp * dcell;
            if (o->nyp > 0.) *ycur = YFACE(o, *cellj + 1) + delta;
            else if (o->nyp < 0.) *ycur = YFACE(o, *cellj) - delta;
            *zcur = *zcur + o->nzp * dcell;
        } else if (dir[2]) {
            *xcur = *xcur + o->nxp * dcell;
            *ycur = *ycur + o->nyp * dcell;
            if (o->nzp > 0.) *zcur = ZFACE(o, *cellk + 1) + delta;
            else if (o->nzp < 0.) *zcur = ZFACE(o, *cellk) - delta;
        }
    } else {
        *xcur = *xcur + o->nxp * dcell;
        *ycur = *ycur + o->nyp * dcell;
        *zcur = *zcur + o->nzp * dcell;
    }
    if (wall_flag) update_voxels(o, *xcur, *ycur, *zcur, celli, cellj, cellk);
}

/* inttau2.f90:242-279.  Returns 0, or -1 where the Fortran prints 'Error in Repeat_bounds...' and stops. */
static int repeat_bounds(int *cella, int *cellb, double *acur, double *bcur, double amax, double bmax, int nag, int nbg,
                         double delta)
{
    if (*cella == -1) {
        if (*acur < delta) {
            *acur = 2. * amax - delta;
            *cella = nag;
        } else if (*acur > 2. * amax - delta) {
            *acur = delta;
            *cella = 1;
        } else {
            return -1;
        }
    }
    if (*cellb == -1) {
        if (*bcur < delta) {
            *bcur = 2. * bmax - delta;
            *cellb = nbg;
        } else if (*bcur > 2. * bmax - delta) {
            *bcur = delta;
            *cellb = 1;
        } else {
            return -1;
        }
    }
    return 0;
}

/* inttau2.f90:7-72 */
static void tauint1(orc_state *o, double xmax, double ymax, double zmax, int *xcell, int *ycell, int *zcell,
                    int *tflag, double delta)
{
    double tau, taurun, taucell, xcur, ycur, zcur, d, dcell;
    int celli, cellj, cellk;
    int dir[3];

    xcur = o->xp + xmax;
    ycur = o->yp + ymax;
    zcur = o->zp + zmax;

    celli = *xcell;
    cellj = *ycell;
    cellk = *zcell;

    taurun = 0.;
    d = 0.;

    tau = -log(draw(o));
    for (;;) {
        dir[0] = dir[1] = dir[2] = 0;
        dcell = wall_dist(o, celli, cellj, cellk, xcur, ycur, zcur, dir);
        taucell = dcell * RHOKAP(o, celli, cellj, cellk);
        o->steps++;

        if (taurun + taucell < tau) {
            taurun = taurun + taucell;
            d = d + dcell;
            JMEAN(o, celli, cellj, cellk) = JMEAN(o, celli, cellj, cellk) + dcell * RHOKAP(o, celli, cellj, cellk);
            o->deposit += dcell * RHOKAP(o, celli, cellj, cellk);
            {
                const int pi = celli, pj = cellj, pk = cellk;
                update_pos(o, &xcur, &ycur, &zcur, &celli, &cellj, &cellk, dcell, 1, dir, delta);
                /* ORC_FLAG_PERIODIC: the call site repeat_bounds never got upstream -- a packet that left through a
                 * lateral face re-enters on the opposite side and the optical-depth integration continues */
                if ((o->flags & ORC_FLAG_PERIODIC) && (celli == -1 || cellj == -1)) {
                    if (repeat_bounds(&celli, &cellj, &xcur, &ycur, xmax, ymax, o->nxg, o->nyg, delta) == 0) o->wraps++;
                }
                if ((o->flags & ORC_FLAG_FRESNEL) && (celli == -1 || cellj == -1 || cellk == -1)) {
                    /* extension: the crossed face is an outer face of the grid */
                    const int a = dir[0] ? 0 : (dir[1] ? 1 : 2);
                    const double na = a == 0 ? o->nxp : (a == 1 ? o->nyp : o->nzp);
                    const int only = (a == 0 && celli == -1 && cellj != -1 && cellk != -1) ||
                                     (a == 1 && cellj == -1 && celli != -1 && cellk != -1) ||
                                     (a == 2 && cellk == -1 && celli != -1 && cellj != -1);
                    const double n_in = o->n_g ? o->n_g[HALO(o, pi, pj, pk)] : o->n2;
                    if (only && draw_boundary(o) < fresnel_reflectance(n_in, o->n1, fabs(na))) {
                        o->internal_reflections++;
                        if (a == 0) {
                            xcur = (na > 0.) ? XFACE(o, pi + 1) - delta : XFACE(o, pi) + delta;
                            celli = pi;
                            o->nxp = -o->nxp; o->cosp = -o->cosp;
                        } else if (a == 1) {
                            ycur = (na > 0.) ? YFACE(o, pj + 1) - delta : YFACE(o, pj) + delta;
                            cellj = pj;
                            o->nyp = -o->nyp; o->sinp = -o->sinp;
                        } else {
                            zcur = (na > 0.) ? ZFACE(o, pk + 1) - delta : ZFACE(o, pk) + delta;
                            cellk = pk;
                            o->nzp = -o->nzp; o->cost = -o->cost;
                        }
                        if (a != 2) o->phi = atan2(o->sinp, o->cosp);   /* keeps (cost, sint, phi) consistent for stokes */
                    }
                }
            }
        } else {
            dcell = (tau - taurun) / RHOKAP(o, celli, cellj, cellk);
            d = d + dcell;
            JMEAN(o, celli, cellj, cellk) = JMEAN(o, celli, cellj, cellk) + dcell * RHOKAP(o, celli, cellj, cellk);
            o->deposit += dcell * RHOKAP(o, celli, cellj, cellk);
            update_pos(o, &xcur, &ycur, &zcur, &celli, &cellj, &cellk, dcell, 0, dir, delta);
            break;
        }

        if (celli == -1 || cellj == -1 || cellk == -1) {
            *tflag = 1;
            break;
        }
    }
    (void)d;

    o->xp = xcur - xmax;
    o->yp = ycur - ymax;
    o->zp = zcur - zmax;
    *xcell = celli;
    *ycell = cellj;
    *zcell = cellk;
}

/* ------------------------------------------------------------------ scattering */

/* stokes.f90:6-153.  Not compiled upstream; restated because the north-star path
 * names it.  No polarisation state exists: it only rotates the direction. */
static void stokes(orc_state *o)
{
    double costp, sintp, phip, bmu, ri1, ri3, cosi3, sini3;
    double cosb2, sinbt, cosi2 = 0., sini1, cosi1, sini2, bott, cosdph, t;

    if (o->hgg == 0.0) {
        /* stokes.f90:23-38 */
        o->cost = 2. * draw(o) - 1.;
        o->sint = (1. - o->cost * o->cost);
        if (o->sint <= 0.) o->sint = 0.;
        else o->sint = sqrt(o->sint);

        o->phi = TWOPI * draw(o);
        o->sinp = sin(o->phi);
        o->cosp = cos(o->phi);

        o->nxp = o->sint * o->cosp;
        o->nyp = o->sint * o->sinp;
        o->nzp = o->cost;
        return;
    }

    costp = o->cost;
    sintp = o->sint;
    phip = o->phi;

    /* stokes.f90:48 */
    t = (1. - o->g2) / (1. - o->hgg + 2. * o->hgg * draw(o));
    bmu = ((1. + o->g2) - t * t) / (2. * o->hgg);
    cosb2 = bmu * bmu;

    if (fabs(bmu) > 1.) {
        if (bmu > 1.) { bmu = 1.; cosb2 = 1.; }
        else { bmu = -1.; cosb2 = 1.; }
    }
    sinbt = sqrt(1. - cosb2);
    ri1 = TWOPI * draw(o);

    if (ri1 > PI) {
        ri3 = TWOPI - ri1;
        cosi3 = cos(ri3);
        sini3 = sin(ri3);

        if (bmu == 1.) return;      /* goto 100, stokes.f90:71-77 */
        if (bmu == -1.) return;

        o->cost = costp * bmu + sintp * sinbt * cosi3;
        if (fabs(o->cost) < 1.) {
            o->sint = fabs(sqrt(1. - o->cost * o->cost));
            sini2 = sini3 * sintp / o->sint;
            bott = o->sint * sinbt;
            cosi2 = costp / bott - o->cost * bmu / bott;
        } else {
            o->sint = 0.;
            sini2 = 0.;
            if (o->cost >= 1.) cosi2 = -1.;
            if (o->cost <= -1.) cosi2 = 1.;
        }

        cosdph = -cosi2 * cosi3 + sini2 * sini3 * bmu;
        if (fabs(cosdph) > 1.) {
            if (cosdph > 1.) cosdph = 1.;
            else cosdph = -1.;
        }

        o->phi = phip + acos(cosdph);
        if (o->phi > TWOPI) o->phi = o->phi - TWOPI;
        if (o->phi < 0.) o->phi = o->phi + TWOPI;
    } else {
        cosi1 = cos(ri1);
        sini1 = sin(ri1);
        if (bmu == 1.) return;      /* goto 100, stokes.f90:109-115 */
        if (bmu == -1.) return;

        o->cost = costp * bmu + sintp * sinbt * cosi1;
        if (fabs(o->cost) < 1.) {
            o->sint = fabs(sqrt(1. - o->cost * o->cost));
            sini2 = sini1 * sintp / o->sint;
            bott = o->sint * sinbt;
            cosi2 = costp / bott - o->cost * bmu / bott;
        } else {
            o->sint = 0.;
            sini2 = 0.;
            if (o->cost >= 1.) cosi2 = -1.;
            if (o->cost <= -1.) cosi2 = 1.;
        }

        cosdph = -cosi1 * cosi2 + sini1 * sini2 * bmu;
        if (fabs(cosdph) > 1.) {
            if (cosdph > 1.) cosdph = 1.;
            else cosdph = -1.;
        }
        o->phi = phip - acos(cosdph);
        if (o->phi > TWOPI) o->phi = o->phi - TWOPI;
        if (o->phi < 0.) o->phi = o->phi + TWOPI;
    }

    o->cosp = cos(o->phi);
    o->sinp = sin(o->phi);

    o->nxp = o->sint * o->cosp;
    o->nyp = o->sint * o->sinp;
    o->nzp = o->cost;
}

/* ------------------------------------------------------------------ the photon loop */

double orc_rang(orc_state *o, double avg, double sigma) { return rang(o, avg, sigma); }
int orc_repeat_bounds(int *cella, int *cellb, double *acur, double *bcur, double amax, double bmax, int nag, int nbg,
                      double delta)
{
    return repeat_bounds(cella, cellb, acur, bcur, amax, bmax, nag, nbg, delta);
}

void orc_stokes_chain(orc_state *o, int nsteps, double *out)
{
    int xcell, ycell, zcell;
    sourcephCO2(o, o->xmax, o->ymax, o->zmax, &xcell, &ycell, &zcell);
    for (int s = 0; s < nsteps; ++s) {
        stokes(o);
        double *q = out + 8 * (size_t)s;
        q[0] = o->nxp; q[1] = o->nyp; q[2] = o->nzp; q[3] = o->cost;
        q[4] = o->sint; q[5] = o->cosp; q[6] = o->sinp; q[7] = o->phi;
    }
}

static int exit_face(const orc_state *o, int xcell, int ycell, int zcell)
{
    if (xcell == -1) return o->nxp > 0. ? 2 : 1;
    if (ycell == -1) return o->nyp > 0. ? 4 : 3;
    if (zcell == -1) return o->nzp > 0. ? 6 : 5;
    return 0;
}

/* mcpolar.f90:151-170; with ORC_FLAG_SCATTER the stubbed inner loop (:166-169) is replaced by the
 * loop its shell, `albedo` (ch_opt.f90:23) and stokes() imply -- SURVEY.md 3.3:
 *     draw < albedo ? stokes : absorbed ; tauint1                                          */
int orc_run(orc_state *o, int64_t nphotons, orc_packet_record *records, double *draws, int64_t draw_cap,
            int64_t *offsets, orc_stats *stats)
{
    int64_t j;
    int xcell, ycell, zcell, tflag;
    orc_stats st;
    memset(&st, 0, sizeof(st));

    o->draw_log = draws;
    o->draw_cap = draws ? draw_cap : 0;
    o->draw_n = 0;
    o->draw_overflow = 0;

    o->internal_reflections = 0;
    o->wraps = 0;
    for (j = 1; j <= nphotons; j++) {
        int absorbed = 0, specular = 0;
        tflag = 0;
        o->pkt_draws = 0;
        o->steps = 0;
        o->nscatt = 0;
        o->deposit = 0.;
        if (offsets) offsets[j - 1] = o->draw_n;

        o->pkt_bdraws = 0;
        o->pkt_sdraws = 0;
        o->ph_block_idx = -1;
        if (o->gauss_sigma > 0.) sourcephGauss(o, o->xmax, o->ymax, o->zmax, &xcell, &ycell, &zcell);
        else sourcephCO2(o, o->xmax, o->ymax, o->zmax, &xcell, &ycell, &zcell);
        if (o->flags & ORC_FLAG_FRESNEL) {
            /* extension: specular reflection at the top surface, normal incidence */
            const double n_in = o->n_g ? o->n_g[HALO(o, xcell, ycell, zcell)] : o->n2;
            const double r0 = (o->n1 - n_in) / (o->n1 + n_in);
            if (draw_boundary(o) < r0 * r0) {
                specular = 1;
                tflag = 1;
                o->nzp = 1.;        /* leaves upwards */
                zcell = -1;
            }
        }
        if (!tflag) tauint1(o, o->xmax, o->ymax, o->zmax, &xcell, &ycell, &zcell, &tflag, o->delta);

        while (!tflag) {
            if (!(o->flags & ORC_FLAG_SCATTER)) { /* mcpolar.f90:167-168 */
                tflag = 1;
                absorbed = 1;
                break;
            }
            if (o->hgg_g) {                 /* the voxel of the interaction (tauint1 leaves its indices in x/y/zcell) */
                o->hgg = o->hgg_g[HALO(o, xcell, ycell, zcell)];
                o->g2 = o->hgg * o->hgg;
            }
            if (draw(o) < (o->albedo_g ? o->albedo_g[HALO(o, xcell, ycell, zcell)] : o->albedo)) {
                stokes(o);
                o->nscatt++;
            } else {
                tflag = 1;
                absorbed = 1;
                break;
            }
            tauint1(o, o->xmax, o->ymax, o->zmax, &xcell, &ycell, &zcell, &tflag, o->delta);
        }

        {
            int fate = absorbed ? 0 : exit_face(o, xcell, ycell, zcell);
            st.packets++;
            st.voxel_steps += o->steps;
            st.scatters += o->nscatt;
            st.draws += o->pkt_draws;
            st.deposit_sum += o->deposit;
            if (fate == 0) st.absorbed++;
            else st.exits[fate - 1]++;
            st.specular += specular;
            if (records) {
                orc_packet_record *r = &records[j - 1];
                r->xp = o->xp; r->yp = o->yp; r->zp = o->zp;
                r->nxp = o->nxp; r->nyp = o->nyp; r->nzp = o->nzp;
                r->deposit = o->deposit;
                r->xcell = xcell; r->ycell = ycell; r->zcell = zcell;
                r->steps = o->steps; r->nscatt = o->nscatt; r->ndraws = (int32_t)o->pkt_draws;
                r->fate = fate; r->flags = 0;
            }
        }
        if (o->rng_mode == ORC_RNG_PHILOX) o->ph_packet++;
    }
    if (offsets) offsets[nphotons] = o->draw_n;
    st.internal_reflections = o->internal_reflections;
    st.wraps = o->wraps;
    if (stats) *stats = st;
    o->draw_log = NULL;
    return o->draw_overflow ? -1 : 0;
}

/* ------------------------------------------------------------------ emulated MPI ranks */

static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

typedef struct {
    orc_state **ranks;
    orc_stats *rstats;
    int nranks, nthreads, tid;
    int64_t nphotons;
} rank_job;

static void *rank_worker(void *arg)
{
    rank_job *jb = (rank_job *)arg;
    int r;
    for (r = jb->tid; r < jb->nranks; r += jb->nthreads)
        orc_run(jb->ranks[r], jb->nphotons, NULL, NULL, 0, NULL, &jb->rstats[r]);
    return NULL;
}

int orc_run_ranks(int nranks, int nxg, int nyg, int nzg, double xmax, double ymax, double zmax,
                  const double *rhokap_halo, double albedo, double hgg, double spot_diameter, int flags,
                  int64_t nphotons_per_rank, double *jmean_global, orc_stats *stats, double *seconds)
{
    const size_t nh = (size_t)(nxg + 2) * (nyg + 2) * (nzg + 2), nj = (size_t)nxg * nyg * nzg;
    orc_state **ranks = (orc_state **)calloc((size_t)nranks, sizeof(*ranks));
    orc_stats *rstats = (orc_stats *)calloc((size_t)nranks, sizeof(*rstats));
    int r, t, used;
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    double t0, t1;
    size_t v;
    pthread_t *th;
    rank_job *jobs;

    for (r = 0; r < nranks; r++) {
        ranks[r] = orc_create(nxg, nyg, nzg, xmax, ymax, zmax);
        memcpy(ranks[r]->rhokap, rhokap_halo, nh * sizeof(double));
        orc_set_optics(ranks[r], albedo, hgg);
        if (spot_diameter > 0.) orc_set_spot(ranks[r], spot_diameter);
        orc_set_flags(ranks[r], flags);
        orc_seed_ran2(ranks[r], r); /* mcpolar.f90:97-98 */
    }
    used = (int)(ncpu < 1 ? 1 : ncpu);
    if (used > nranks) used = nranks;
    th = (pthread_t *)calloc((size_t)used, sizeof(*th));
    jobs = (rank_job *)calloc((size_t)used, sizeof(*jobs));

    t0 = now_s();
    for (t = 0; t < used; t++) {
        jobs[t].ranks = ranks; jobs[t].rstats = rstats; jobs[t].nranks = nranks;
        jobs[t].nthreads = used; jobs[t].tid = t; jobs[t].nphotons = nphotons_per_rank;
        if (t > 0) pthread_create(&th[t], NULL, rank_worker, &jobs[t]);
    }
    rank_worker(&jobs[0]);
    for (t = 1; t < used; t++) pthread_join(th[t], NULL);
    /* mcpolar.f90:173: MPI_allREDUCE(jmean -> jmeanGLOBAL, SUM) */
    memset(jmean_global, 0, nj * sizeof(double));
    for (r = 0; r < nranks; r++)
        for (v = 0; v < nj; v++) jmean_global[v] += ranks[r]->jmean[v];
    t1 = now_s();
    if (seconds) *seconds = t1 - t0;
    if (stats) {
        int f;
        memset(stats, 0, sizeof(*stats));
        for (r = 0; r < nranks; r++) {
            stats->packets += rstats[r].packets;
            stats->voxel_steps += rstats[r].voxel_steps;
            stats->scatters += rstats[r].scatters;
            stats->absorbed += rstats[r].absorbed;
            stats->draws += rstats[r].draws;
            stats->deposit_sum += rstats[r].deposit_sum;
            stats->specular += rstats[r].specular;
            stats->internal_reflections += rstats[r].internal_reflections;
            stats->wraps += rstats[r].wraps;
            for (f = 0; f < 6; f++) stats->exits[f] += rstats[r].exits[f];
        }
    }
    for (r = 0; r < nranks; r++) orc_destroy(ranks[r]);
    free(ranks); free(rstats); free(th); free(jobs);
    return used;
}
