// tamc_heat.cu -- the caller on the far side of the transport hot path, on the device: the explicit
// 3-D finite-difference heat step, the Arrhenius damage integral and the thermal / optical property
// update that rewrites rhokap.  With these on the GPU the coupled
// ablation loop of /root/reference/src/mcpolar.f90:148-186 runs with jmean, temp and rhokap resident
// in HBM: no PCIe copy per MC call.
//
//   (in k_heat_step)   mcpolar.f90:174        jmeanGLOBAL *= (getPwr()/81)/(nphotons*numproc*Vvoxel)
//   k_heat_step        3dFD.f90:113-186       FTCS 7-point stencil with variable kappa/rho/c, boiling sink
//   k_post             3dFD.f90:424-466 + :327-345   Arrhenius + the voxel-local half of setupThermalCoeff
//   k_rule_air         3dFD.f90:347-357       the sweep-order "six neighbours ablated" rule (first pass of a
//                                             fixed point, see k_rule) + air properties; k_rule / k_air finish
//                                             the rare cases that need more passes
// Single-rank semantics (numproc = 1) of heat_sim_3D; with several ranks every GPU repeats the same
// deterministic step on its own replica (the tally it consumes is already all-reduced), so no exchange
// is needed.  Compiled with -fmad=false: the arithmetic is the oracle's, operation for operation
// (the CPU restatement the tests check against); results differ only through exp() in the air conductivity and the damage rate.
//
// Each kernel is a streaming pass over the grid: HBM-bound, one thread per voxel, x fastest (coalesced).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "tamc_context.h"

using namespace tamc;

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return tamc_fail_(TAMC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));            \
    } while (0)

// ------------------------------------------------------------------------------------------------
// state: the Heat module's variables (3dFD.f90:7-12) + the arrays the driver owns (iarray.f90:11-12)
// ------------------------------------------------------------------------------------------------
struct tamc_heat {
    int n = 0;
    size_t nh = 0, ni = 0;
    // host-side scalars, advanced exactly as the reference does
    double pulseCount = 0, repetitionCount = 0, time = 0, laserOn = 1, total_time = 0, repetitionRate_1 = 0, energyPerPixel = 0;
    double Power = 0, pulselength = 0, delt = 0, realPulseLength = 0;
    double dx = 0, dy = 0, dz = 0, massVoxel = 0, volumeVoxel = 0, QVapor = 0, ablateTemp = 0, jscale = 0;
    bool laser_flag = true, pulseFlag = false;
    int loops = 1, pulsesToDo = 1, pulsesDone = 0, pulsetype = 1, counter = 0;
    // device arrays
    double *coeff = nullptr, *kappa = nullptr, *density = nullptr, *heatcap = nullptr, *alpha = nullptr;   // (0:n+1)^3
    double *temp = nullptr, *tn = nullptr, *cand = nullptr, *cand2 = nullptr, *rk_new = nullptr;  // (0:n+1)^3
    double *water = nullptr, *Q = nullptr, *tissue = nullptr, *thres = nullptr;                              // n^3 (thres: 3 n^3)
    int *flags = nullptr;                                                                                     // [1] negative temperature (sticky)
};

namespace {

constexpr double kAirHeatCap = 1.006e3, kLw = 2256.e3, kWaterInit = .75, kProtein = 1. - .75;

__host__ __device__ inline double airThermalCond(double T) { return -0.188521 * exp(-0.000367259 * (T - 273.15)) + 0.212453; }
__host__ __device__ inline double airDensity(double T) { return 101.325e3 / (287.058 * T); }
__host__ __device__ inline double skinDensity(double w) { return 1000. / (w + 0.649 * kProtein); }
__host__ __device__ inline double skinHeatCap(double w) { return 1000. * (4.2 * w + 1.09 * kProtein); }
__host__ __device__ inline double skinThermalCond(double w, double rho) { return rho * (6.28e-4 * w + 1.17e-4 * kProtein); }

__device__ __forceinline__ size_t h3(int n, int i, int j, int k) { return (size_t)i + (size_t)(n + 2) * ((size_t)j + (size_t)(n + 2) * (size_t)k); }

// thread -> interior voxel (i fastest); returns false past the end
__device__ __forceinline__ bool voxel_of_thread(int n, int &i, int &j, int &k)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * n * n) return false;
    i = (int)(t % n) + 1;
    j = (int)((t / n) % n) + 1;
    k = (int)(t / ((size_t)n * n)) + 1;
    return true;
}

// 3dFD.f90:113-186, one sub-step: reads t0, writes tn and Q
__global__ void __launch_bounds__(256) k_heat_step(int n, const double *__restrict__ t0, double *__restrict__ tn,
                                                   const double *__restrict__ kappa, const double *__restrict__ density,
                                                   const double *__restrict__ heatcap, const double *__restrict__ coeff,
                                                   const double *__restrict__ jmean, double *__restrict__ Q,
                                                   double inv_dx2, double inv_dy2, double inv_dz2, double delt, double laserOn,
                                                   double volumeVoxel, double massVoxel, double QVapor, double jscale,
                                                   int *__restrict__ flags)
{
    int i, j, k;
    if (!voxel_of_thread(n, i, j, k)) return;
    const size_t c = h3(n, i, j, k), sx = 1, sy = (size_t)(n + 2), sz = (size_t)(n + 2) * (n + 2);
    const size_t qi = (size_t)(i - 1) + (size_t)n * ((size_t)(j - 1) + (size_t)n * (size_t)(k - 1));
    const double kc = kappa[c], dc = density[c], hc = heatcap[c], tc = t0[c];

    // inv_h2 = 1.d0/dz**2 etc. (3dFD.f90:131-132), formed once on the host: the same IEEE quotient
    auto second_derivative = [&](size_t s, double inv_h2) {
        const double kappaPlusHalf = .5 * (kc + kappa[c + s]), kappaMinHalf = .5 * (kc + kappa[c - s]);
        const double densityPlusHalf = .5 * (dc + density[c + s]), densityMinHalf = .5 * (dc + density[c - s]);
        const double heatcapPlusHalf = .5 * (hc + heatcap[c + s]), heatcapMinHalf = .5 * (hc + heatcap[c - s]);
        const double a = 0.5 * (kappaMinHalf / (densityMinHalf * heatcapMinHalf)) * inv_h2;
        const double d = 0.5 * (kappaPlusHalf / (densityPlusHalf * heatcapPlusHalf)) * inv_h2;
        const double b = 0.5 * (a + d);
        return a * t0[c - s] - 2. * b * tc + d * t0[c + s];
    };
    const double u_zz = second_derivative(sz, inv_dz2);
    const double u_yy = second_derivative(sy, inv_dy2);
    const double u_xx = second_derivative(sx, inv_dx2);

    const double jv = jmean[qi] * jscale;          // mcpolar.f90:174 applied on the fly: the resident tally stays unscaled
    const double tempIncrease = delt * (u_xx + u_yy + u_zz);
    const double energyIncrease = laserOn * jv * delt * volumeVoxel + hc * massVoxel * tempIncrease;
    const double q = Q[qi];
    double out;
    if (tc >= 100. + 273. && q < QVapor) {                      // boil water, :168-175
        if (energyIncrease > 0.) {
            const double qn = q + energyIncrease;
            Q[qi] = qn < QVapor ? qn : QVapor;
            out = 100. + 273.;
        } else {
            out = tc + tempIncrease + laserOn * coeff[c] * jv;
        }
    } else {
        out = tc + tempIncrease + laserOn * coeff[c] * jv;
        if (tc < 0.) flags[1] = 1;                                // :179-182 (mpi_abort upstream)
    }
    tn[c] = out;
}

// The "remove tissue whose six neighbours are ablated" rule (3dFD.f90:347-353) is evaluated upstream
// inside the k,j,i sweep, so a voxel sees FINAL values of the neighbours behind the sweep (i-1, j-1, k-1)
// and PREVIOUS-CALL values of those ahead (i+1, j+1, k+1):
//     final(v) = clamp(rule(local(v), final(behind), old(ahead))).
// One data-parallel pass that reads clamp(local) for the neighbours behind gives exactly that sequential result.
// Proof: the two can differ at a voxel Y only if a neighbour X behind it has final(X) = 0 != clamp(local(X)), i.e. X
// was zeroed by the rule itself.  That needs all six terms of X's sum to vanish, among them old(Y) -- Y is one of
// X's neighbours AHEAD.  But opacity never grows back: old(Y) = 0 makes local(Y) = 0 (the property update only
// rescales rhokap > 0, 3dFD.f90:334-345), so final(Y) = 0 whatever X became.  Hence no host round trip and no
// iteration: the whole time loop is enqueued ahead of the GPU.  (tests/test_gpu_heat.py compares the opacity with the
// sequential oracle after every call, through boiling, ablation and hand-made air pockets.)

// ---- fused passes (the common path of tamc_heat_step) --------------------------------------------------------
// k_post = Arrhenius + the voxel-local half of setupThermalCoeff + clamp: one read of temp / rhokap / Q per voxel.
__global__ void __launch_bounds__(256) k_post(int n, const double *__restrict__ temp, const double *__restrict__ rhokap,
                                              double *__restrict__ tissue, double *__restrict__ thres, double *__restrict__ cand,
                                              double *__restrict__ cur, double *__restrict__ water, const double *__restrict__ Q,
                                              double *__restrict__ density, double *__restrict__ heatcap, double *__restrict__ kappa,
                                              double *__restrict__ coeff, double QVapor, double ablateTemp, double delt, double time)
{
    int i, j, k;
    if (!voxel_of_thread(n, i, j, k)) return;
    const size_t c = h3(n, i, j, k), ni = (size_t)n * n * n;
    const size_t q = (size_t)(i - 1) + (size_t)n * ((size_t)(j - 1) + (size_t)n * (size_t)(k - 1));
    const double T = temp[c];
    double rk = rhokap[c];
    // Arrhenius, 3dFD.f90:447-460 (reads the opacity of the previous property update, like upstream)
    double ts = tissue[q];
    if (T >= 43. + 273. && T < 100. + 273. && rk >= 0.) {
        ts = ts + delt * 3.1e98 * exp(-6.3e5 / (8.314 * T));
        tissue[q] = ts;
    }
    if (thres[q] == 0. && ts >= .53) thres[q] = time;
    else if (thres[q + ni] == 0. && ts >= 1.) thres[q + ni] = time;
    else if (thres[q + 2 * ni] == 0. && ts >= 10000.) thres[q + 2 * ni] = time;
    // setupThermalCoeff, 3dFD.f90:327-345
    double w = kWaterInit - kWaterInit * (Q[q] / QVapor);
    w = w < kWaterInit ? w : kWaterInit;
    const double wc = water[q];
    w = w < wc ? w : wc;
    w = w > 0.0 ? w : 0.0;
    water[q] = w;
    if (T >= ablateTemp + 273.) {
        rk = 0.;
    } else if (rk > 0.) {
        const double rho = skinDensity(w);
        density[c] = rho;
        rk = w * 510. + 170.;
        const double hcap = skinHeatCap(w);
        heatcap[c] = hcap;
        kappa[c] = skinThermalCond(w, rho);
        coeff[c] = delt / (rho * hcap);
    }
    cand[c] = rk;
    cur[c] = rk <= 0.01 ? 0. : rk;
}

// k_rule_air = the neighbour rule (see above) + the air-property block, writing the new opacity to a second buffer
// (the rule still needs last call's values ahead of the sweep).
__global__ void __launch_bounds__(256) k_rule_air(int n, const double *__restrict__ local, const double *__restrict__ cur,
                                                  const double *__restrict__ old,
                                                  const double *__restrict__ temp, double *__restrict__ rk_new,
                                                  double *__restrict__ density, double *__restrict__ heatcap,
                                                  double *__restrict__ kappa, double *__restrict__ alpha, double *__restrict__ coeff,
                                                  double delt)
{
    int i, j, k;
    if (!voxel_of_thread(n, i, j, k)) return;
    const size_t c = h3(n, i, j, k), sy = (size_t)(n + 2), sz = (size_t)(n + 2) * (n + 2);
    double v = local[c];
    const double summ = old[c + sz] + old[c + sy] + old[c + 1] + cur[c - sz] + cur[c - sy] + cur[c - 1];
    if (summ == 0.) v = 0.;
    if (v <= 0.01) v = 0.;
    rk_new[c] = v;
    if (v <= 0.01) {
        const double T = temp[c];
        const double rho = airDensity(T);
        density[c] = rho;
        heatcap[c] = 1.006e3;
        const double kap = airThermalCond(T);
        kappa[c] = kap;
        alpha[c] = kap / (rho * 1.006e3);
        coeff[c] = delt / (airDensity(T) * 1.006e3);
    }
}

inline int blocks_for(size_t n) { return (int)((n + 255) / 256); }

// 3dFD.f90:365-421
double get_pwr(tamc_heat *s)
{
    if (s->pulsetype == 1) {
        const double fact = (2. * sqrt(2. * log(2.)));
        const double mu = fact * s->pulselength, sig = s->pulselength / fact;
        return s->Power * exp(-((s->time - mu) * (s->time - mu)) / (2. * (sig * sig)));
    }
    if (s->pulsetype == 0) return s->laser_flag ? s->Power : 0.;
    const double m = s->Power / s->pulselength, c = 2. * s->Power;
    if (!s->laser_flag) return 0.;
    if (s->pulseFlag || s->time >= s->pulselength) {
        s->pulseFlag = true;
        const double p = -m * s->time + c;
        return p < 0. ? 0. : p;
    }
    return m * s->time;
}

}  // namespace

void tamc_heat_release_(tamc_context *c)
{
    tamc_heat *s = c->heat;
    if (!s) return;
    double *arrs[] = {s->coeff, s->kappa, s->density, s->heatcap, s->alpha, s->temp, s->tn, s->cand, s->cand2, s->rk_new,
                      s->water, s->Q, s->tissue, s->thres};
    for (double *p : arrs) cudaFree(p);
    cudaFree(s->flags);
    delete s;
    c->heat = nullptr;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int tamc_heat_init(tamc_handle h, const tamc_heat_params *p, double *delt_out)
{
    if (int rc = tamc_check_(h)) return rc;
    if (!p) return tamc_fail_(TAMC_EINVAL, "tamc_heat_init: null parameters");
    if (int rc = tamc_sync_resident_(h)) return rc;
    if (h->nxg != h->nyg || h->nxg != h->nzg)
        return tamc_fail_(TAMC_EINVAL, "tamc_heat_init: the heat solver assumes nxg = nyg = nzg (numpoints, mcpolar.f90:61)");
    if (!h->optics_set) return tamc_fail_(TAMC_ESTATE, "tamc_heat_init: call tamc_set_optics first (rhokap must be resident)");
    if (p->pulsetype < 0 || p->pulsetype > 2 || !(p->power > 0) || p->loops < 1)
        return tamc_fail_(TAMC_EINVAL, "tamc_heat_init: bad pulsetype / power / loops");
    tamc_heat_release_(h);
    tamc_heat *s = new tamc_heat();
    h->heat = s;
    const int n = h->nxg;
    s->n = n;
    s->nh = (size_t)(n + 2) * (n + 2) * (n + 2);
    s->ni = (size_t)n * n * n;
    s->Power = p->power; s->energyPerPixel = p->energyPerPixel; s->total_time = p->total_time; s->loops = p->loops;
    s->repetitionRate_1 = p->repetitionRate_1; s->pulsesToDo = p->pulsesToDo; s->pulsetype = p->pulsetype;
    s->ablateTemp = p->ablateTemp;

    // initThermalCoeff, 3dFD.f90:249-293 (same statements as the CPU restatement used by the tests)
    s->dx = (2. * h->xmax * 1.e-2) / ((double)n + 2.);
    s->dy = (2. * h->ymax * 1.e-2) / ((double)n + 2.);
    s->dz = (2. * h->zmax * 1.e-2) / ((double)n + 2.);
    const double heatCaptmp = skinHeatCap(kWaterInit), densitytmp = skinDensity(kWaterInit);
    const double kappatmp = skinThermalCond(kWaterInit, densitytmp);
    const double alphatmp = kappatmp / (densitytmp * skinHeatCap(kWaterInit));
    const double constd = (1. / (s->dx * s->dx)) + (1. / (s->dy * s->dy)) + (1. / (s->dz * s->dz));
    s->delt = 1. / (1. * alphatmp * constd);
    s->pulselength = (p->energyPerPixel * 1.e-3 * (double)(9 * 9)) / p->power;          // spotsPerRow*spotsPerCol, constants.f90:12
    s->volumeVoxel = (2. * h->xmax * 1.e-2 / n) * (2. * h->ymax * 1.e-2 / n) * (2. * h->zmax * 1.e-2 / n);
    s->massVoxel = densitytmp * s->volumeVoxel;
    s->QVapor = kLw * s->massVoxel;
    s->realPulseLength = p->pulsetype == 0 ? s->pulselength : (p->pulsetype == 1 ? 20000. * s->pulselength : 2. * s->pulselength);
    if (p->pulsetype == 1) {                                                              // mcpolar.f90:134-140
        s->total_time = 2. * s->pulselength * (2. * sqrt(2. * log(2.)));
        s->realPulseLength = s->total_time;
    } else if ((int)(s->total_time / s->delt) <= (int)(s->realPulseLength / s->delt)) {
        s->total_time = s->delt * (s->realPulseLength / s->delt + 2000.);
    }

    // host images of the initial arrays, then one upload each
    std::vector<double> alpha(s->nh, alphatmp), kappa(s->nh, airThermalCond(25. + 273.)), density(s->nh, densitytmp),
        heatcap(s->nh, heatCaptmp), coeff(s->nh, 0.), temp(s->nh, 5. + 273.);
    auto H = [n](int i, int j, int k) { return (size_t)i + (size_t)(n + 2) * ((size_t)j + (size_t)(n + 2) * (size_t)k); };
    const double alpha_air = airThermalCond(25. + 273.) / (airDensity(25. + 273.) * kAirHeatCap);
    for (int j = 0; j <= n + 1; ++j)
        for (int i = 0; i <= n + 1; ++i) alpha[H(i, j, n + 1)] = alpha_air;
    for (int k = 1; k <= n; ++k)
        for (int j = 1; j <= n; ++j)
            for (int i = 1; i <= n; ++i) {
                kappa[H(i, j, k)] = skinThermalCond(kWaterInit, densitytmp);
                coeff[H(i, j, k)] = alphatmp * s->delt / kappatmp;
            }
    // mcpolar.f90:123-129 (z faces last)
    for (int j = 0; j <= n + 1; ++j)
        for (int i = 0; i <= n + 1; ++i) { temp[H(i, j, 0)] = 25. + 273.; temp[H(i, j, n + 1)] = 25. + 273.; }

    auto up = [&](double **d, const std::vector<double> &v) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, v.size() * sizeof(double));
        return e != cudaSuccess ? e : cudaMemcpy(*d, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice);
    };
    CU(up(&s->alpha, alpha)); CU(up(&s->kappa, kappa)); CU(up(&s->density, density)); CU(up(&s->heatcap, heatcap));
    CU(up(&s->coeff, coeff)); CU(up(&s->temp, temp)); CU(up(&s->tn, temp));
    CU(cudaMalloc(&s->cand, s->nh * sizeof(double))); CU(cudaMalloc(&s->cand2, s->nh * sizeof(double)));
    CU(cudaMemset(s->cand, 0, s->nh * sizeof(double))); CU(cudaMemset(s->cand2, 0, s->nh * sizeof(double)));
    CU(cudaMalloc(&s->rk_new, s->nh * sizeof(double)));
    CU(cudaMemset(s->rk_new, 0, s->nh * sizeof(double)));
    std::vector<double> water(s->ni, kWaterInit);
    CU(up(&s->water, water));
    CU(cudaMalloc(&s->Q, s->ni * sizeof(double))); CU(cudaMemset(s->Q, 0, s->ni * sizeof(double)));
    CU(cudaMalloc(&s->tissue, s->ni * sizeof(double))); CU(cudaMemset(s->tissue, 0, s->ni * sizeof(double)));
    CU(cudaMalloc(&s->thres, 3 * s->ni * sizeof(double))); CU(cudaMemset(s->thres, 0, 3 * s->ni * sizeof(double)));
    CU(cudaMalloc(&s->flags, 4 * sizeof(int))); CU(cudaMemset(s->flags, 0, 4 * sizeof(int)));
    if (delt_out) *delt_out = s->delt;
    return TAMC_OK;
}

static int need_heat(tamc_handle h, const char *who)
{
    if (int rc = tamc_check_(h)) return rc;
    if (!h->heat) return tamc_fail_(TAMC_ESTATE, std::string(who) + ": tamc_heat_init has not been called");
    return tamc_sync_resident_(h);          // root_io: the heat step reads the resident opacity grid on every rank
}

// mcpolar.f90:174 followed by :178-182: scale the resident tally, heat_sim_3d, arrhenius, setupThermalCoeff.
extern "C" int tamc_heat_step(tamc_handle h, int64_t nphotons_times_numproc)
{
    if (int rc = need_heat(h, "tamc_heat_step")) return rc;
    tamc_heat *s = h->heat;
    const int n = s->n;
    cudaStream_t st = h->stream;
    const int gi = blocks_for(s->ni);

    // mcpolar.f90:149,174: jmeanGLOBAL is rescaled only while the laser is on; afterwards the array keeps its last
    // scaled values (and laserOn = 0 multiplies them away).  The factor is applied inside the stencil kernel.
    if (s->laser_flag) {
        if (nphotons_times_numproc <= 0) return tamc_fail_(TAMC_EINVAL, "tamc_heat_step: packet count must be positive");
        s->jscale = (get_pwr(s) / 81.) / ((double)nphotons_times_numproc * (2. * h->xmax * 1.e-2 / n) *
                                          (2. * h->ymax * 1.e-2 / n) * (2. * h->zmax * 1.e-2 / n));
    }
    // heat_sim_3D, 3dFD.f90:101-214
    if (s->pulselength < s->delt) s->delt = s->pulselength / 100.;
    for (int p = 1; p <= s->loops; ++p) {
        k_heat_step<<<gi, 256, 0, st>>>(n, s->temp, s->tn, s->kappa, s->density, s->heatcap, s->coeff, h->d_jmean, s->Q, 1. / (s->dx * s->dx),
                                        1. / (s->dy * s->dy), 1. / (s->dz * s->dz), s->delt, s->laserOn, s->volumeVoxel, s->massVoxel, s->QVapor, s->jscale, s->flags);
        std::swap(s->temp, s->tn);                                                       // t0 = tn (both hold the same halo)
        if (s->pulseCount >= s->realPulseLength && s->laser_flag) {                      // :199-211
            s->laser_flag = false; s->laserOn = 0.; s->pulseCount = 0.; s->pulsesDone += 1; s->repetitionCount = 0.;
        } else if (s->repetitionCount >= s->repetitionRate_1 && !s->laser_flag && s->pulsesDone < s->pulsesToDo) {
            s->laser_flag = true; s->laserOn = 1.; s->pulseCount = 0.; s->repetitionCount = 0.;
        }
        s->pulseCount += s->delt;
        s->repetitionCount += s->delt;
        s->time += s->delt;
    }
    // arrhenius(temp, delt, tissue, ThresTime, 1, N, N) and setupThermalCoeff(temp, N, ablateTemp), mcpolar.f90:180-182,
    // in two fused passes; the new opacity goes to a second buffer that then becomes the resident rhokap
    double *cur = s->cand2;                        // clamp(local); its halo stays 0 like rhokap's
    k_post<<<gi, 256, 0, st>>>(n, s->temp, h->d_rhokap, s->tissue, s->thres, s->cand, cur, s->water, s->Q, s->density, s->heatcap,
                               s->kappa, s->coeff, s->QVapor, s->ablateTemp, s->delt, s->time);
    k_rule_air<<<gi, 256, 0, st>>>(n, s->cand, cur, h->d_rhokap, s->temp, s->rk_new, s->density, s->heatcap, s->kappa,
                                   s->alpha, s->coeff, s->delt);
    std::swap(h->d_rhokap, s->rk_new);              // both buffers keep a zero halo
    CU(cudaGetLastError());
    s->counter += 1;
    return TAMC_OK;
}

// do while(time <= total_time) ... end do, mcpolar.f90:148-186, everything resident on the device
extern "C" int tamc_coupled_loop(tamc_handle h, int64_t nphotons, int64_t seed, int64_t max_iterations, int64_t *iterations_done,
                                 int64_t *packets_done)
{
    if (int rc = need_heat(h, "tamc_coupled_loop")) return rc;
    tamc_heat *s = h->heat;
    int64_t it = 0, pk = 0;
    while (s->time <= s->total_time && (max_iterations < 0 || it < max_iterations)) {
        if (s->laser_flag) {
            if (int rc = tamc_run_async(h, nphotons, seed, -1)) return rc;             // replaces :151-173
            pk += nphotons * h->nranks;
        }
        if (int rc = tamc_heat_step(h, nphotons * (int64_t)h->nranks)) return rc;      // :174-185
        ++it;
    }
    CU(cudaStreamSynchronize(h->stream));
    int flags[2] = {0, 0};
    CU(cudaMemcpy(flags, s->flags, sizeof(flags), cudaMemcpyDeviceToHost));
    if (iterations_done) *iterations_done = it;
    if (packets_done) *packets_done = pk;
    if (flags[1]) return tamc_fail_(TAMC_EINVAL, "heat step: negative temperature (the reference aborts here, 3dFD.f90:179-182)");
    return TAMC_OK;
}

extern "C" int tamc_heat_array(tamc_handle h, int which, double *host, int upload)
{
    if (int rc = need_heat(h, "tamc_heat_array")) return rc;
    tamc_heat *s = h->heat;
    if (!host) return tamc_fail_(TAMC_EINVAL, "tamc_heat_array: null host pointer");
    double *d = nullptr;
    size_t cnt = s->nh;
    switch (which) {
    case TAMC_HEAT_TEMP: d = s->temp; break;
    case TAMC_HEAT_RHOKAP: d = h->d_rhokap; break;
    case TAMC_HEAT_KAPPA: d = s->kappa; break;
    case TAMC_HEAT_DENSITY: d = s->density; break;
    case TAMC_HEAT_HEATCAP: d = s->heatcap; break;
    case TAMC_HEAT_COEFF: d = s->coeff; break;
    case TAMC_HEAT_ALPHA: d = s->alpha; break;
    case TAMC_HEAT_WATER: d = s->water; cnt = s->ni; break;
    case TAMC_HEAT_Q: d = s->Q; cnt = s->ni; break;
    case TAMC_HEAT_TISSUE: d = s->tissue; cnt = s->ni; break;
    case TAMC_HEAT_THRESTIME: d = s->thres; cnt = 3 * s->ni; break;
    case TAMC_HEAT_JMEAN: d = h->d_jmean; cnt = s->ni; break;
    default: return tamc_fail_(TAMC_EINVAL, "tamc_heat_array: unknown array id");
    }
    CU(cudaStreamSynchronize(h->stream));
    if (upload) {
        CU(cudaMemcpy(d, host, cnt * sizeof(double), cudaMemcpyHostToDevice));
        if (which == TAMC_HEAT_TEMP) CU(cudaMemcpy(s->tn, host, cnt * sizeof(double), cudaMemcpyHostToDevice));
    } else {
        CU(cudaMemcpy(host, d, cnt * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return TAMC_OK;
}

extern "C" int tamc_heat_scalar(tamc_handle h, int which, double *out)
{
    if (int rc = need_heat(h, "tamc_heat_scalar")) return rc;
    tamc_heat *s = h->heat;
    if (!out) return tamc_fail_(TAMC_EINVAL, "tamc_heat_scalar: null destination");
    switch (which) {
    case TAMC_HEAT_S_DELT: *out = s->delt; break;
    case TAMC_HEAT_S_TIME: *out = s->time; break;
    case TAMC_HEAT_S_TOTAL_TIME: *out = s->total_time; break;
    case TAMC_HEAT_S_PULSELENGTH: *out = s->pulselength; break;
    case TAMC_HEAT_S_REALPULSELENGTH: *out = s->realPulseLength; break;
    case TAMC_HEAT_S_LASERON: *out = s->laserOn; break;
    case TAMC_HEAT_S_PULSECOUNT: *out = s->pulseCount; break;
    case TAMC_HEAT_S_REPETITIONCOUNT: *out = s->repetitionCount; break;
    case TAMC_HEAT_S_LASER_FLAG: *out = s->laser_flag ? 1. : 0.; break;
    case TAMC_HEAT_S_QVAPOR: *out = s->QVapor; break;
    case TAMC_HEAT_S_PWR: *out = get_pwr(s); break;
    case TAMC_HEAT_S_COUNTER: *out = (double)s->counter; break;
    case TAMC_HEAT_S_NEGATIVE_TEMP: {
        // the sticky flag of the stencil kernel: the condition on which the reference calls mpi_abort (3dFD.f90:179-182)
        int flags[2] = {0, 0};
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaMemcpy(flags, s->flags, sizeof(flags), cudaMemcpyDeviceToHost));
        *out = flags[1] ? 1. : 0.;
        break;
    }
    default: return tamc_fail_(TAMC_EINVAL, "tamc_heat_scalar: unknown scalar id");
    }
    return TAMC_OK;
}
