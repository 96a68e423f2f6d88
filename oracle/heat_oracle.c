/*
 * heat_oracle.c -- plain-C fp64 restatement of the reference's heat / ablation step, the caller on
 * the far side of the photon-transport hot path:
 *
 *   thermalConst_mod.f90:1-88   material laws
 *   3dFD.f90:233-309            initThermalCoeff
 *   3dFD.f90:21-230             heat_sim_3D, single-rank path (numproc = 1: the send/recv, scatter,
 *                               halo Sendrecv to MPI_PROC_NULL and allgather are identities)
 *   3dFD.f90:312-361            setupThermalCoeff (rewrites rhokap)
 *   3dFD.f90:365-421            getPwr{Gaussian,TopHat,Triangular}
 *   3dFD.f90:424-466            Arrhenius
 *   mcpolar.f90:123-140,174     temperature boundary set-up, total_time override, jmean scaling
 *
 * TEST INFRASTRUCTURE ONLY, like tamc_oracle.c.  No upstream vectors and no Fortran compiler here: pinned, bit for bit,
 * to what the reference's own text computes for one rank when oracle/f90interp.py executes it (the time loop of
 * heat_sim_3D, Arrhenius, setupThermalCoeff, initThermalCoeff's arithmetic, the getPwr functions, thermalConst_mod.f90,
 * the driver lines of mcpolar.f90; tests/golden/reference_interp_heat.json.gz, tests/test_oracle_reference_heat.py) --
 * not to a compiled reference.  Variable names follow the Fortran; every `real` is a double
 * (-freal-4-real-8).  Quirks kept: dx,dy,dz use numpoints+2 while volumeVoxel uses nxg (:249-251,
 * :287); the negative-temperature check looks at the INPUT temp (:179); pulsesDone starts at 0
 * (uninitialised upstream); the "six neighbours ablated" rule reads neighbours in sweep order, i.e.
 * already-updated values behind the sweep and previous-call values ahead of it (:347-349).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n;                                 /* nxg = nyg = nzg = numpoints */
    double xmax, ymax, zmax;
    /* Heat module scalars (3dFD.f90:7-12) */
    double pulseCount, repetitionCount, time, laserOn, total_time, repetitionRate_1, energyPerPixel;
    double Power, pulselength, delt, realPulseLength;
    double dx, dy, dz, massVoxel, volumeVoxel;
    int laser_flag, pulseFlag, loops, pulsesToDo, pulsesDone, loops_left;
    int pulsetype;                         /* 0 tophat, 1 gaussian, 2 triangular */
    /* thermalConstants */
    double skinDensityInit, QVapor;
    /* arrays: (0:n+1)^3 */
    double *coeff, *kappa, *density, *heatcap, *alpha, *temp, *rhokap;
    /* arrays: (1:n)^3 */
    double *WaterContent, *Q, *tissue, *ThresTime; /* ThresTime(n,n,n,3) */
    int negative_temp;                     /* set instead of mpi_abort (3dFD.f90:179-182) */
} heat_state;

#define H3(s, i, j, k) ((size_t)(i) + (size_t)((s)->n + 2) * ((size_t)(j) + (size_t)((s)->n + 2) * (size_t)(k)))
#define I3(s, i, j, k) ((size_t)((i)-1) + (size_t)(s)->n * ((size_t)((j)-1) + (size_t)(s)->n * (size_t)((k)-1)))

/* ---- thermalConst_mod.f90 ---------------------------------------------------------------------- */
static const double airHeatCap = 1.006e3, lw = 2256.e3;
static const double waterContentInit = .75, proteinContent = 1. - .75;

/* (thermalConst_mod.f90:20-23 stops the program when T < 0 -- the explicit scheme has diverged in an air voxel; the oracle
 * goes on with whatever exp() returns, and tests/test_oracle_reference_heat.py checks that condition at the iteration where
 * the reference's own text stops) */
static double airThermalCond(double T)
{
    const double a = -0.188521, b = 0.000367259, c = 0.212453;
    return a * exp(-b * (T - 273.15)) + c;
}
static double airDensity(double T) { return 101.325e3 / (287.058 * T); }
static double getWaterContent(const heat_state *s, double Qcurrent, double watercurrent)
{
    double v = waterContentInit - waterContentInit * (Qcurrent / s->QVapor);
    v = v < waterContentInit ? v : waterContentInit;
    v = v < watercurrent ? v : watercurrent;
    return v > 0.0 ? v : 0.0;
}
static double getSkinDensity(double w) { return 1000. / (w + 0.649 * proteinContent); }
static double getSkinHeatCap(double w) { return 1000. * (4.2 * w + 1.09 * proteinContent); }
static double getSkinThermalCond(double w, double rho) { return rho * (6.28e-4 * w + 1.17e-4 * proteinContent); }

/* ---- power functions, 3dFD.f90:365-421 ---------------------------------------------------------- */
double heat_getPwr(heat_state *s)
{
    if (s->pulsetype == 1) {
        const double fact = (2. * sqrt(2. * log(2.)));
        const double mu = fact * s->pulselength;
        const double sig = s->pulselength / fact;
        return s->Power * exp(-((s->time - mu) * (s->time - mu)) / (2. * (sig * sig)));
    }
    if (s->pulsetype == 0) return s->laser_flag ? s->Power : 0.;
    {
        const double m = s->Power / s->pulselength, c = 2. * s->Power;
        double p;
        if (!s->laser_flag) return 0.;
        if (s->pulseFlag) {
            p = -m * s->time + c;
            return p < 0. ? 0. : p;
        }
        if (s->time >= s->pulselength) {
            s->pulseFlag = 1;
            p = -m * s->time + c;
            return p < 0. ? 0. : p;
        }
        return m * s->time;
    }
}

/* ---- set-up ------------------------------------------------------------------------------------- */
heat_state *heat_create(int n, double xmax, double ymax, double zmax)
{
    heat_state *s = (heat_state *)calloc(1, sizeof(*s));
    const size_t nh = (size_t)(n + 2) * (n + 2) * (n + 2), ni = (size_t)n * n * n;
    s->n = n; s->xmax = xmax; s->ymax = ymax; s->zmax = zmax;
    s->coeff = calloc(nh, 8); s->kappa = calloc(nh, 8); s->density = calloc(nh, 8); s->heatcap = calloc(nh, 8);
    s->alpha = calloc(nh, 8); s->temp = calloc(nh, 8); s->rhokap = calloc(nh, 8);
    s->WaterContent = calloc(ni, 8); s->Q = calloc(ni, 8); s->tissue = calloc(ni, 8); s->ThresTime = calloc(3 * ni, 8);
    return s;
}

void heat_destroy(heat_state *s)
{
    if (!s) return;
    free(s->coeff); free(s->kappa); free(s->density); free(s->heatcap); free(s->alpha); free(s->temp); free(s->rhokap);
    free(s->WaterContent); free(s->Q); free(s->tissue); free(s->ThresTime);
    free(s);
}

double *heat_array(heat_state *s, int which)
{
    switch (which) {
    case 0: return s->temp;  case 1: return s->rhokap; case 2: return s->kappa; case 3: return s->density;
    case 4: return s->heatcap; case 5: return s->coeff; case 6: return s->alpha; case 7: return s->WaterContent;
    case 8: return s->Q; case 9: return s->tissue; case 10: return s->ThresTime;
    }
    return NULL;
}

double heat_scalar(const heat_state *s, int which)
{
    switch (which) {
    case 0: return s->delt; case 1: return s->time; case 2: return s->total_time; case 3: return s->pulselength;
    case 4: return s->realPulseLength; case 5: return s->laserOn; case 6: return s->pulseCount;
    case 7: return s->repetitionCount; case 8: return (double)s->laser_flag; case 9: return s->QVapor;
    case 10: return s->volumeVoxel; case 11: return (double)s->pulsesDone; case 12: return (double)s->negative_temp;
    case 13: return s->massVoxel;
    }
    return 0.;
}

/* mcpolar.f90:65-71,123-140 + initThermalCoeff (3dFD.f90:233-309).  rhokap_kappa = opt_prop::kappa
 * for gridset.f90:33-45.  Returns delt. */
double heat_init(heat_state *s, double power, double energyPerPixel, double total_time, int loops,
                 double repetitionRate_1, int pulsesToDo, int pulsetype, double rhokap_kappa)
{
    const int n = s->n;
    const size_t nh = (size_t)(n + 2) * (n + 2) * (n + 2), ni = (size_t)n * n * n;
    const int spotsPerRow = 9, spotsPerCol = 9;            /* constants.f90:12 */
    double densitytmp, alphatmp, kappatmp, heatCaptmp, constd;
    int i, j, k;
    size_t v;

    s->Power = power; s->energyPerPixel = energyPerPixel; s->total_time = total_time; s->loops = loops;
    s->repetitionRate_1 = repetitionRate_1; s->pulsesToDo = pulsesToDo; s->pulsetype = pulsetype;
    /* mcpolar.f90:65-71 */
    s->time = 0.; s->pulseCount = 0.; s->repetitionCount = 0.; s->laserOn = 1.; s->laser_flag = 1; s->pulseFlag = 0;
    s->pulsesDone = 0; s->negative_temp = 0;
    memset(s->tissue, 0, ni * 8); memset(s->ThresTime, 0, 3 * ni * 8);
    /* gridset.f90:33-45 */
    memset(s->rhokap, 0, nh * 8);
    for (k = 1; k <= n; k++) for (j = 1; j <= n; j++) for (i = 1; i <= n; i++) s->rhokap[H3(s, i, j, k)] = rhokap_kappa;
    /* mcpolar.f90:123-129: later assignments win on shared edges */
    for (v = 0; v < nh; v++) s->temp[v] = 5. + 273.;
    for (k = 0; k <= n + 1; k++) for (j = 0; j <= n + 1; j++) { s->temp[H3(s, n + 1, j, k)] = 5. + 273.; s->temp[H3(s, 0, j, k)] = 5. + 273.; }
    for (k = 0; k <= n + 1; k++) for (i = 0; i <= n + 1; i++) { s->temp[H3(s, i, 0, k)] = 5. + 273.; s->temp[H3(s, i, n + 1, k)] = 5. + 273.; }
    for (j = 0; j <= n + 1; j++) for (i = 0; i <= n + 1; i++) { s->temp[H3(s, i, j, 0)] = 25. + 273.; s->temp[H3(s, i, j, n + 1)] = 25. + 273.; }

    /* initThermalCoeff, 3dFD.f90:249-293 */
    s->dx = (2. * s->xmax * 1.e-2) / ((double)n + 2.);
    s->dy = (2. * s->ymax * 1.e-2) / ((double)n + 2.);
    s->dz = (2. * s->zmax * 1.e-2) / ((double)n + 2.);
    memset(s->Q, 0, ni * 8);
    s->skinDensityInit = getSkinDensity(waterContentInit);
    for (v = 0; v < ni; v++) s->WaterContent[v] = waterContentInit;
    heatCaptmp = getSkinHeatCap(waterContentInit);
    densitytmp = getSkinDensity(waterContentInit);
    kappatmp = getSkinThermalCond(waterContentInit, densitytmp);
    alphatmp = kappatmp / (densitytmp * getSkinHeatCap(waterContentInit));
    for (v = 0; v < nh; v++) s->alpha[v] = alphatmp;
    for (j = 0; j <= n + 1; j++) for (i = 0; i <= n + 1; i++)
        s->alpha[H3(s, i, j, n + 1)] = airThermalCond(25. + 273.) / (airDensity(25. + 273.) * airHeatCap);
    for (v = 0; v < nh; v++) s->kappa[v] = airThermalCond(25. + 273.);
    for (k = 1; k <= n; k++) for (j = 1; j <= n; j++) for (i = 1; i <= n; i++)
        s->kappa[H3(s, i, j, k)] = getSkinThermalCond(waterContentInit, densitytmp);
    for (v = 0; v < nh; v++) { s->density[v] = densitytmp; s->heatcap[v] = heatCaptmp; }
    constd = (1. / (s->dx * s->dx)) + (1. / (s->dy * s->dy)) + (1. / (s->dz * s->dz));
    s->delt = 1. / (1. * alphatmp * constd);
    memset(s->coeff, 0, nh * 8);
    for (k = 1; k <= n; k++) for (j = 1; j <= n; j++) for (i = 1; i <= n; i++)
        s->coeff[H3(s, i, j, k)] = alphatmp * s->delt / kappatmp;
    s->pulselength = (energyPerPixel * 1.e-3 * (double)(spotsPerRow * spotsPerCol)) / power;
    s->volumeVoxel = (2. * s->xmax * 1.e-2 / n) * (2. * s->ymax * 1.e-2 / n) * (2. * s->zmax * 1.e-2 / n);
    s->massVoxel = densitytmp * s->volumeVoxel;
    s->QVapor = lw * s->massVoxel;
    if (pulsetype == 0) s->realPulseLength = s->pulselength;
    else if (pulsetype == 1) s->realPulseLength = 20000. * s->pulselength;
    else s->realPulseLength = 2. * s->pulselength;
    /* mcpolar.f90:134-140 */
    if (pulsetype == 1) {
        s->total_time = 2. * s->pulselength * (2. * sqrt(2. * log(2.)));
        s->realPulseLength = s->total_time;
    } else if ((int)(s->total_time / s->delt) <= (int)(s->realPulseLength / s->delt)) {
        s->total_time = s->delt * (s->realPulseLength / s->delt + 2000.);
    }
    return s->delt;
}

/* mcpolar.f90:174: jmeanGLOBAL *= (getPwr()/81)/(nphotons*numproc*Vvoxel); returns the factor */
double heat_scale_jmean(heat_state *s, double *jmeanGLOBAL, double nphotons_times_numproc)
{
    const int n = s->n;
    const double f = (heat_getPwr(s) / 81.) / (nphotons_times_numproc * (2. * s->xmax * 1.e-2 / n) *
                                               (2. * s->ymax * 1.e-2 / n) * (2. * s->zmax * 1.e-2 / n));
    size_t v, ni = (size_t)n * n * n;
    for (v = 0; v < ni; v++) jmeanGLOBAL[v] = jmeanGLOBAL[v] * f;
    return f;
}

/* heat_sim_3D, 3dFD.f90:21-230, numproc = 1 */
void heat_sim_3d(heat_state *s, const double *jmean, int counter)
{
    const int n = s->n;
    const size_t nh = (size_t)(n + 2) * (n + 2) * (n + 2);
    double *t0 = (double *)malloc(nh * 8), *tn = (double *)malloc(nh * 8);
    const double *kappa = s->kappa, *density = s->density, *heatcap = s->heatcap;
    const double dx = s->dx, dy = s->dy, dz = s->dz;
    int i, j, k, p;

    memcpy(t0, s->temp, nh * 8);      /* t0(:,:,:) = temp(:,:,0:zf+1) */
    memcpy(tn, t0, nh * 8);
    if (s->pulselength < s->delt) s->delt = s->pulselength / 100.;        /* :101-104 */
    s->loops_left = (int)(s->total_time / ((double)s->loops * s->delt)) - counter;

    for (p = 1; p <= s->loops; p++) {
        for (k = 1; k <= n; k++)
            for (j = 1; j <= n; j++)
                for (i = 1; i <= n; i++) {
                    double kappaPlusHalf, kappaMinHalf, densityPlusHalf, densityMinHalf, heatcapPlusHalf, heatcapMinHalf;
                    double a, b, d, u_xx, u_yy, u_zz, tempIncrease, energyIncrease;
                    const size_t c = H3(s, i, j, k), qi = I3(s, i, j, k);

                    kappaPlusHalf = .5 * (kappa[c] + kappa[H3(s, i, j, k + 1)]);
                    kappaMinHalf = .5 * (kappa[c] + kappa[H3(s, i, j, k - 1)]);
                    densityPlusHalf = .5 * (density[c] + density[H3(s, i, j, k + 1)]);
                    densityMinHalf = .5 * (density[c] + density[H3(s, i, j, k - 1)]);
                    heatcapPlusHalf = .5 * (heatcap[c] + heatcap[H3(s, i, j, k + 1)]);
                    heatcapMinHalf = .5 * (heatcap[c] + heatcap[H3(s, i, j, k - 1)]);
                    a = 0.5 * (kappaMinHalf / (densityMinHalf * heatcapMinHalf)) * (1. / (dz * dz));
                    d = 0.5 * (kappaPlusHalf / (densityPlusHalf * heatcapPlusHalf)) * (1. / (dz * dz));
                    b = 0.5 * (a + d);
                    u_zz = a * t0[H3(s, i, j, k - 1)] - 2. * b * t0[c] + d * t0[H3(s, i, j, k + 1)];

                    kappaPlusHalf = 0.5 * (kappa[c] + kappa[H3(s, i, j + 1, k)]);
                    kappaMinHalf = 0.5 * (kappa[c] + kappa[H3(s, i, j - 1, k)]);
                    densityPlusHalf = 0.5 * (density[c] + density[H3(s, i, j + 1, k)]);
                    densityMinHalf = 0.5 * (density[c] + density[H3(s, i, j - 1, k)]);
                    heatcapPlusHalf = 0.5 * (heatcap[c] + heatcap[H3(s, i, j + 1, k)]);
                    heatcapMinHalf = 0.5 * (heatcap[c] + heatcap[H3(s, i, j - 1, k)]);
                    a = 0.5 * (kappaMinHalf / (densityMinHalf * heatcapMinHalf)) * (1. / (dy * dy));
                    d = 0.5 * (kappaPlusHalf / (densityPlusHalf * heatcapPlusHalf)) * (1. / (dy * dy));
                    b = 0.5 * (a + d);
                    u_yy = a * t0[H3(s, i, j - 1, k)] - 2. * b * t0[c] + d * t0[H3(s, i, j + 1, k)];

                    kappaPlusHalf = .5 * (kappa[c] + kappa[H3(s, i + 1, j, k)]);
                    kappaMinHalf = .5 * (kappa[c] + kappa[H3(s, i - 1, j, k)]);
                    densityPlusHalf = .5 * (density[c] + density[H3(s, i + 1, j, k)]);
                    densityMinHalf = .5 * (density[c] + density[H3(s, i - 1, j, k)]);
                    heatcapPlusHalf = .5 * (heatcap[c] + heatcap[H3(s, i + 1, j, k)]);
                    heatcapMinHalf = .5 * (heatcap[c] + heatcap[H3(s, i - 1, j, k)]);
                    a = 0.5 * (kappaMinHalf / (densityMinHalf * heatcapMinHalf)) * (1. / (dx * dx));
                    d = 0.5 * (kappaPlusHalf / (densityPlusHalf * heatcapPlusHalf)) * (1. / (dx * dx));
                    b = 0.5 * (a + d);
                    u_xx = a * t0[H3(s, i - 1, j, k)] - 2. * b * t0[c] + d * t0[H3(s, i + 1, j, k)];

                    tempIncrease = s->delt * (u_xx + u_yy + u_zz);
                    energyIncrease = s->laserOn * jmean[qi] * s->delt * s->volumeVoxel + heatcap[c] * s->massVoxel * tempIncrease;
                    if (tn[c] >= 100. + 273. && s->Q[qi] < s->QVapor) {                       /* boil water */
                        if (energyIncrease > 0.) {
                            const double q = s->Q[qi] + energyIncrease;
                            s->Q[qi] = q < s->QVapor ? q : s->QVapor;
                            tn[c] = 100. + 273.;
                        } else {
                            tn[c] = tn[c] + tempIncrease + s->laserOn * s->coeff[c] * jmean[qi];
                        }
                    } else {
                        tn[c] = tn[c] + tempIncrease + s->laserOn * s->coeff[c] * jmean[qi];
                        if (s->temp[c] < 0.) s->negative_temp = 1;                             /* :179-182 */
                    }
                }
        memcpy(t0, tn, nh * 8);                                                               /* t0 = tn */
        /* halo Sendrecv with MPI_PROC_NULL neighbours: nothing moves (:190-197) */
        if (s->pulseCount >= s->realPulseLength && s->laser_flag) {                           /* :199-211 */
            s->laser_flag = 0; s->laserOn = 0.; s->pulseCount = 0.; s->pulsesDone = s->pulsesDone + 1; s->repetitionCount = 0.;
        } else if (s->repetitionCount >= s->repetitionRate_1 && !s->laser_flag && s->pulsesDone < s->pulsesToDo) {
            s->laser_flag = 1; s->laserOn = 1.; s->pulseCount = 0.; s->repetitionCount = 0.;
        }
        s->pulseCount = s->pulseCount + s->delt;
        s->repetitionCount = s->repetitionCount + s->delt;
        s->time = s->time + s->delt;
    }
    /* allgather: temp(:,:,zi:zf) = t0(:,:,zi:zf), x/y halo columns included (:218-219) */
    for (k = 1; k <= n; k++)
        memcpy(s->temp + H3(s, 0, 0, k), t0 + H3(s, 0, 0, k), (size_t)(n + 2) * (n + 2) * 8);
    free(t0);
    free(tn);
}

/* Arrhenius, 3dFD.f90:424-466, called with zi = 1, zf = numpoints (mcpolar.f90:180) */
void heat_arrhenius(heat_state *s)
{
    const int n = s->n;
    const size_t ni = (size_t)n * n * n;
    const double A = 3.1e98, dE = 6.3e5, R = 8.314, first = .53, second = 1., third = 10000.;
    int x, y, z;
    for (z = 1; z <= n; z++)
        for (y = 1; y <= n; y++)
            for (x = 1; x <= n; x++) {
                const size_t c = H3(s, x, y, z), q = I3(s, x, y, z);
                const double T = s->temp[c];
                if (T >= 43. + 273. && T < 100. + 273. && s->rhokap[c] >= 0.)
                    s->tissue[q] = s->tissue[q] + s->delt * A * exp(-dE / (R * T));
                if (s->ThresTime[q] == 0. && s->tissue[q] >= first) s->ThresTime[q] = s->time;
                else if (s->ThresTime[q + ni] == 0. && s->tissue[q] >= second) s->ThresTime[q + ni] = s->time;
                else if (s->ThresTime[q + 2 * ni] == 0. && s->tissue[q] >= third) s->ThresTime[q + 2 * ni] = s->time;
            }
}

/* setupThermalCoeff, 3dFD.f90:312-361 */
void heat_setup_thermal_coeff(heat_state *s, double ablateTemp)
{
    const int n = s->n;
    const double mu_water = 510., mu_protein = 170.;        /* ch_opt.f90:17-18 */
    int i, j, k;
    size_t v, ni = (size_t)n * n * n;
    for (v = 0; v < ni; v++) s->WaterContent[v] = getWaterContent(s, s->Q[v], s->WaterContent[v]);   /* :327 */

    for (k = 1; k <= n; k++)
        for (j = 1; j <= n; j++)
            for (i = 1; i <= n; i++) {
                const size_t c = H3(s, i, j, k), q = I3(s, i, j, k);
                double summ;
                if (s->temp[c] >= ablateTemp + 273.) {
                    s->rhokap[c] = 0.;
                } else if (s->rhokap[c] > 0.) {
                    s->density[c] = getSkinDensity(s->WaterContent[q]);
                    s->rhokap[c] = s->WaterContent[q] * mu_water + mu_protein;
                    s->heatcap[c] = getSkinHeatCap(s->WaterContent[q]);
                    s->kappa[c] = getSkinThermalCond(s->WaterContent[q], s->density[c]);
                    s->coeff[c] = s->delt / (s->density[c] * s->heatcap[c]);
                }
                summ = s->rhokap[H3(s, i, j, k + 1)] + s->rhokap[H3(s, i, j + 1, k)] + s->rhokap[H3(s, i + 1, j, k)] +
                       s->rhokap[H3(s, i, j, k - 1)] + s->rhokap[H3(s, i, j - 1, k)] + s->rhokap[H3(s, i - 1, j, k)];
                if (summ == 0.) s->rhokap[c] = 0.;
                if (s->rhokap[c] <= 0.01) {
                    s->density[c] = airDensity(s->temp[c]);
                    s->heatcap[c] = 1.006e3;
                    s->rhokap[c] = 0.;
                    s->kappa[c] = airThermalCond(s->temp[c]);
                    s->alpha[c] = s->kappa[c] / (s->density[c] * s->heatcap[c]);
                    s->coeff[c] = s->delt / (airDensity(s->temp[c]) * s->heatcap[c]);
                }
            }
}
