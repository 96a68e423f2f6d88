// mcgrid_shim.cpp -- host driver that reproduces the call sequence of the reference's main program
// (/root/reference/src/mcpolar.f90) around the GPU transport, in C++ because this image has no
// Fortran compiler.  It is the stand-in for the patched mcpolar.f90 (fortran/mcpolar_patch.f90):
//
//   read input.params (mcpolar.f90:79-94)  ->  init_opt1 (ch_opt.f90:15-23)  ->  gridset
//   (gridset.f90:23-45)  ->  delta (mcpolar.f90:112)  ->  loop { MC call (replaces :151-173) ;
//   scale jmeanGLOBAL (:174) ; property update that rewrites rhokap (stand-in for heat_sim_3d +
//   setupThermalCoeff, 3dFD.f90:312-361) }  ->  write jmean in writer.f90's raw-stream format.
//
// The 3-D heat solver is out of scope; its effect on the optical grid is
// scripted: an ablation crater (rhokap = 0) that grows under the beam with a water-depleted rim
// (rhokap = w*mu_water + mu_protein), plus the reference's "remove tissue whose six neighbours are
// ablated" rule.  That is BASELINE.json config 5: many small MC calls, each preceded by a re-upload
// of rhokap, latency per call reported with its breakdown.
//
// With --resident the heat step is the real one (3dFD.f90 on the device, tamc_heat_*) and the whole
// `do while(time <= total_time)` loop runs through tamc_coupled_loop with jmean, temp and rhokap resident
// in HBM; --calls then bounds the number of loop iterations (-1 = run to total_time, 13 390 for the
// shipped parameters).
//
//   usage: mcgrid_shim [--params FILE] [--calls N] [--nxg N] [--scatter] [--resident] [--out DIR] [--device D]
//                      [--gaussian SIGMA_CM] [--periodic]     (the rang / repeat_bounds options, default off)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/tamc.h"

namespace {

struct Params {                 // res/input.params, in file order (mcpolar.f90:79-94)
    long long nphotons = 125000;
    double xmax = 0.03, ymax = 0.03, zmax = 0.06, n1 = 1.0, n2 = 1.38, total_time = 2.0;
    int loops = 1;
    double repetitionRate_1 = 1e7, power = 70, energyPerPixel = 400, ablateTemp = 500;
    int pulsesToDo = 1;
    std::string pulsetype = "gaussian";
};

// Fortran list-directed read of one value per record: first token, and a '/' ends the record
// (so the shipped "10000000./1." reads as 10000000.).
std::string first_token(const std::string &line)
{
    std::istringstream is(line);
    std::string tok;
    is >> tok;
    const size_t slash = tok.find('/');
    if (slash != std::string::npos) tok = tok.substr(0, slash);
    return tok;
}

bool read_params(const std::string &path, Params &p)
{
    std::ifstream f(path);
    if (!f) return false;
    std::vector<std::string> v;
    std::string line;
    while (std::getline(f, line)) v.push_back(first_token(line));
    if (v.size() < 14) return false;
    p.nphotons = std::atoll(v[0].c_str());
    p.xmax = std::atof(v[1].c_str()); p.ymax = std::atof(v[2].c_str()); p.zmax = std::atof(v[3].c_str());
    p.n1 = std::atof(v[4].c_str()); p.n2 = std::atof(v[5].c_str()); p.total_time = std::atof(v[6].c_str());
    p.loops = std::atoi(v[7].c_str());
    p.repetitionRate_1 = std::atof(v[8].c_str()); p.power = std::atof(v[9].c_str());
    p.energyPerPixel = std::atof(v[10].c_str()); p.ablateTemp = std::atof(v[11].c_str());
    p.pulsesToDo = std::atoi(v[12].c_str());
    p.pulsetype = v[13];
    return true;
}

struct Optics { double hgg, g2, mu_water, mu_protein, mua, mus, kappa, albedo; };

Optics init_opt1()               // ch_opt.f90:15-23
{
    Optics o;
    o.hgg = 0.9; o.g2 = o.hgg * o.hgg;
    o.mu_water = 510.; o.mu_protein = 170.;
    o.mua = o.mu_water + o.mu_protein;
    o.mus = 0.;
    o.kappa = o.mus + o.mua;
    o.albedo = o.mus / o.kappa;
    return o;
}

struct Grid {
    int nxg, nyg, nzg;
    std::vector<double> rhokap;  // (0:nxg+1,0:nyg+1,0:nzg+1) column-major
    double &rk(int i, int j, int k) { return rhokap[(size_t)i + (size_t)(nxg + 2) * ((size_t)j + (size_t)(nyg + 2) * k)]; }
};

void gridset(Grid &g, double kappa)   // gridset.f90:33-45 (faces live in the library)
{
    g.rhokap.assign((size_t)(g.nxg + 2) * (g.nyg + 2) * (g.nzg + 2), 0.);
    for (int k = 1; k <= g.nzg; ++k)
        for (int j = 1; j <= g.nyg; ++j)
            for (int i = 1; i <= g.nxg; ++i) g.rk(i, j, k) = kappa;
}

// Stand-in for heat_sim_3d + setupThermalCoeff (3dFD.f90:312-361): crater radius/depth grow with the
// call index; inside -> ablated, a two-voxel rim loses water, isolated tissue is removed.
void scripted_property_update(Grid &g, const Optics &o, int call, int ncalls)
{
    const double frac = (double)(call + 1) / ncalls;
    const double radius = 1.0 + 0.2 * g.nxg * frac;        // voxels
    const int depth = 1 + (int)(0.25 * g.nzg * frac);
    for (int k = g.nzg; k > g.nzg - depth && k >= 1; --k)
        for (int j = 1; j <= g.nyg; ++j)
            for (int i = 1; i <= g.nxg; ++i) {
                const double r = std::hypot(i - 0.5 - g.nxg / 2.0, j - 0.5 - g.nyg / 2.0);
                const double shrink = radius * (1.0 - 0.5 * (g.nzg - k) / (double)depth);
                if (r <= shrink) g.rk(i, j, k) = 0.;                                   // ablated, :334-335
                else if (r <= shrink + 2. && g.rk(i, j, k) > 0.)
                    g.rk(i, j, k) = 0.5 * o.mu_water + o.mu_protein;                  // water loss, :343
            }
    for (int k = 1; k <= g.nzg; ++k)                                                   // :347-349
        for (int j = 1; j <= g.nyg; ++j)
            for (int i = 1; i <= g.nxg; ++i) {
                const double summ = g.rk(i, j, k + 1) + g.rk(i, j + 1, k) + g.rk(i + 1, j, k) + g.rk(i, j, k - 1) +
                                    g.rk(i, j - 1, k) + g.rk(i - 1, j, k);
                if (summ == 0.) g.rk(i, j, k) = 0.;
            }
}

std::string fstr(double v, int len)   // utils.f90 str(real, len): leading characters of the value
{
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.10f", v);
    return std::string(buf).substr(0, (size_t)len);
}

void die(const char *what, int rc)
{
    std::fprintf(stderr, "mcgrid_shim: %s failed (%d): %s\n", what, rc, tamc_last_error());
    std::exit(1);
}

}  // namespace

int main(int argc, char **argv)
{
    std::string params_path, out_dir;
    int calls = 200, nxg = 80, device = 0;
    bool scatter = false, resident = false, periodic = false;
    double gauss_sigma = 0.;
    for (int a = 1; a < argc; ++a) {
        const std::string s = argv[a];
        auto next = [&](const char *name) -> const char * {
            if (a + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", name); std::exit(2); }
            return argv[++a];
        };
        if (s == "--params") params_path = next("--params");
        else if (s == "--calls") calls = std::atoi(next("--calls"));
        else if (s == "--nxg") nxg = std::atoi(next("--nxg"));
        else if (s == "--device") device = std::atoi(next("--device"));
        else if (s == "--out") out_dir = next("--out");
        else if (s == "--scatter") scatter = true;
        else if (s == "--resident") resident = true;
        else if (s == "--periodic") periodic = true;
        else if (s == "--gaussian") gauss_sigma = std::atof(next("--gaussian"));
        else { std::fprintf(stderr, "unknown argument %s\n", s.c_str()); return 2; }
    }
    Params P;
    if (!params_path.empty() && !read_params(params_path, P)) {
        std::fprintf(stderr, "cannot read %s (14 list-directed records expected)\n", params_path.c_str());
        return 2;
    }
    const Optics o = init_opt1();
    Grid g{nxg, nxg, nxg, {}};
    gridset(g, o.kappa);
    const double delta = 1.e-8 * (2. * P.zmax / g.nzg);                               // mcpolar.f90:112
    const long long numproc = 1;
    std::printf("# of photons to run %lld per call, %d calls, grid %d^3\n", P.nphotons * numproc, calls, nxg);

    tamc_handle h = nullptr;
    int rc = tamc_init(device, g.nxg, g.nyg, g.nzg, P.xmax, P.ymax, P.zmax, delta, &h);
    if (rc) die("tamc_init", rc);
    if (gauss_sigma > 0. && (rc = tamc_set_source_gaussian(h, gauss_sigma))) die("tamc_set_source_gaussian", rc);
    const int xflags = periodic ? TAMC_PERIODIC : 0;
    std::vector<double> jmeanGLOBAL((size_t)g.nxg * g.nyg * g.nzg, 0.);
    tamc_pin_host(g.rhokap.data(), g.rhokap.size() * sizeof(double));
    tamc_pin_host(jmeanGLOBAL.data(), jmeanGLOBAL.size() * sizeof(double));

    if (resident) {
        // the reference's loop with its real heat / ablation step, nothing crossing PCIe per iteration
        rc = tamc_set_optics(h, g.rhokap.data(), o.albedo, o.hgg, P.n1, P.n2, xflags);
        if (rc) die("tamc_set_optics", rc);
        tamc_heat_params hp{};
        hp.power = P.power; hp.energyPerPixel = P.energyPerPixel; hp.total_time = P.total_time;
        hp.repetitionRate_1 = P.repetitionRate_1; hp.ablateTemp = P.ablateTemp; hp.loops = P.loops; hp.pulsesToDo = P.pulsesToDo;
        hp.pulsetype = P.pulsetype == "tophat" ? 0 : (P.pulsetype == "gaussian" ? 1 : 2);
        double delt = 0., total = 0.;
        rc = tamc_heat_init(h, &hp, &delt);
        if (rc) die("tamc_heat_init", rc);
        tamc_heat_scalar(h, TAMC_HEAT_S_TOTAL_TIME, &total);
        std::printf("delt %.6e s, total_time %.5f s -> %d loop iterations\n", delt, total, (int)(total / delt));
        const auto t0 = std::chrono::steady_clock::now();
        int64_t iters = 0, pk = 0;
        rc = tamc_coupled_loop(h, P.nphotons, 95648324, calls, &iters, &pk);
        if (rc) die("tamc_coupled_loop", rc);
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        double tsim = 0.;
        tamc_heat_scalar(h, TAMC_HEAT_S_TIME, &tsim);
        std::vector<double> temp((size_t)(nxg + 2) * (nxg + 2) * (nxg + 2));
        tamc_heat_array(h, TAMC_HEAT_TEMP, temp.data(), 0);
        tamc_heat_array(h, TAMC_HEAT_RHOKAP, g.rhokap.data(), 0);
        double tmax = 0.; long long ablated = 0;
        for (int k = 1; k <= nxg; ++k) for (int j = 1; j <= nxg; ++j) for (int i = 1; i <= nxg; ++i) {
            const size_t c = (size_t)i + (size_t)(nxg + 2) * ((size_t)j + (size_t)(nxg + 2) * k);
            if (temp[c] > tmax) tmax = temp[c];
            if (g.rhokap[c] == 0.) ++ablated;
        }
        std::printf("resident coupled loop: %lld iterations, %lld packets in %.3f s -> %.1f us per iteration (MC call + heat + "
                    "Arrhenius + property update), simulated time %.5f s, max temp %.1f C, ablated voxels %lld\n",
                    (long long)iters, (long long)pk, sec, 1e6 * sec / (double)(iters > 0 ? iters : 1), tsim, tmax - 273., ablated);
        if (!out_dir.empty()) {
            // mcpolar.f90:193-206 + writer.f90:6-81: the eight raw fp64 stream files, reference names and contents
            const size_t ni = (size_t)nxg * nxg * nxg;
            std::vector<double> tissue(ni), water(ni), thres(3 * ni), inner(ni);
            tamc_heat_array(h, TAMC_HEAT_TISSUE, tissue.data(), 0);
            tamc_heat_array(h, TAMC_HEAT_WATER, water.data(), 0);
            tamc_heat_array(h, TAMC_HEAT_THRESTIME, thres.data(), 0);
            tamc_heat_array(h, TAMC_HEAT_JMEAN, jmeanGLOBAL.data(), 0);
            // the file holds the scaled jmeanGLOBAL of the last MC call (mcpolar.f90:174)
            double pwr = 0.;
            tamc_heat_scalar(h, TAMC_HEAT_S_PWR, &pwr);
            const double vox = (2. * P.xmax * 1e-2 / nxg) * (2. * P.ymax * 1e-2 / nxg) * (2. * P.zmax * 1e-2 / nxg);
            for (double &v : jmeanGLOBAL) v *= (pwr / 81.) / ((double)P.nphotons * numproc * vox);
            // delete damage info about the ablation crater (mcpolar.f90:196-205)
            for (int k = 1; k <= nxg; ++k) for (int j = 1; j <= nxg; ++j) for (int i = 1; i <= nxg; ++i) {
                const size_t c = (size_t)i + (size_t)(nxg + 2) * ((size_t)j + (size_t)(nxg + 2) * k);
                const size_t q = (size_t)(i - 1) + (size_t)nxg * ((size_t)(j - 1) + (size_t)nxg * (k - 1));
                if (g.rhokap[c] <= 0.1) tissue[q] = -1.;
                inner[q] = g.rhokap[c];
            }
            for (double &v : temp) v -= 273.;
            const std::string tail = "w-" + std::to_string(nxg) + "-" + fstr(P.ablateTemp, 3) + "-" + fstr((int)P.energyPerPixel, 3) +
                                     "-" + fstr(P.xmax, 5) + "-" + fstr(P.ymax, 5) + "-" + fstr(P.zmax, 5) + ".dat";
            auto dump = [&](const std::string &stem, const double *d, size_t cnt) {
                const std::string name = out_dir + "/" + stem + std::to_string((int)P.power) + tail;
                FILE *f = std::fopen(name.c_str(), "wb");
                if (!f) { std::fprintf(stderr, "cannot write %s\n", name.c_str()); return; }
                std::fwrite(d, sizeof(double), cnt, f);
                std::fclose(f);
                std::printf("wrote %s\n", name.c_str());
            };
            dump("jmean-t", jmeanGLOBAL.data(), ni);
            dump("rhokap-t", inner.data(), ni);                     // rhokap(1:nxg,1:nyg,1:nzg)
            dump("temp-t", temp.data(), temp.size());               // temp - 273, halo included
            dump("water-t", water.data(), ni);
            dump("tissue-t", tissue.data(), ni);
            dump("time-t-1-", thres.data(), ni);
            dump("time-t-2-", thres.data() + ni, ni);
            dump("time-t-3-", thres.data() + 2 * ni, ni);
        }
        tamc_unpin_host(g.rhokap.data());
        tamc_unpin_host(jmeanGLOBAL.data());
        tamc_finalize(h);
        return 0;
    }

    std::vector<double> wall_ms, kernel_ms, h2d_ms, d2h_ms;
    long long packets = 0, vsteps = 0;
    double absorbed_energy = 0.;
    for (int c = 0; c < calls; ++c) {
        const auto t0 = std::chrono::steady_clock::now();
        tamc_stats st;
        // upload of the rewritten rhokap + mcpolar.f90:151-173, in one call so the copies overlap the transport
        rc = tamc_run_optics(h, g.rhokap.data(), scatter ? 0.9 : o.albedo, o.hgg, P.n1, P.n2, (scatter ? TAMC_SCATTER : 0) | xflags,
                             P.nphotons, 95648324, jmeanGLOBAL.data(), &st);
        if (rc) die("tamc_run_optics", rc);
        const auto t1 = std::chrono::steady_clock::now();
        // mcpolar.f90:174 -- getPwr()/81 is the per-spot power; a constant 1 W stands in for the pulse shape
        const double vox = (2. * P.xmax * 1e-2 / g.nxg) * (2. * P.ymax * 1e-2 / g.nyg) * (2. * P.zmax * 1e-2 / g.nzg);
        const double scale = (1.0 / 81.) / ((double)P.nphotons * numproc * vox);
        double sum = 0.;
        for (double &v : jmeanGLOBAL) { sum += v; v *= scale; }
        absorbed_energy += sum / (double)P.nphotons;
        wall_ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
        kernel_ms.push_back(st.kernel_ms); h2d_ms.push_back(st.h2d_ms); d2h_ms.push_back(st.d2h_ms);
        packets += st.packets; vsteps += st.voxel_steps;
        scripted_property_update(g, o, c, calls);                                       // rewrites rhokap for the next call
    }
    auto stat = [](std::vector<double> v, double &mean, double &p95) {
        mean = 0.; for (double x : v) mean += x; mean /= v.size();
        std::sort(v.begin(), v.end()); p95 = v[(size_t)(0.95 * (v.size() - 1))];
    };
    double m, p;
    stat(wall_ms, m, p);
    std::printf("MC call latency (tamc_run_optics: upload + run + download, host clock): mean %.3f ms, p95 %.3f ms\n", m, p);
    const double wall_mean = m;
    stat(h2d_ms, m, p); std::printf("  rhokap H2D   mean %.3f ms\n", m);
    stat(kernel_ms, m, p); std::printf("  transport    mean %.3f ms\n", m);
    stat(d2h_ms, m, p); std::printf("  jmean D2H    mean %.3f ms\n", m);
    std::printf("packets %lld, voxel-steps %lld, %.4g packets/s end to end, mean absorbed optical depth per packet %.4f\n",
                packets, vsteps, packets / (wall_mean * 1e-3 * calls), absorbed_energy / calls);

    if (!out_dir.empty()) {
        // writer.f90:23-28: raw little-endian fp64 stream, jmean-t<power>w-<nzg>-<ablateTemp>-<energy>-<xmax>-<ymax>-<zmax>.dat
        const std::string name = out_dir + "/jmean-t" + std::to_string((int)P.power) + "w-" + std::to_string(g.nzg) + "-" +
                                 fstr(P.ablateTemp, 3) + "-" + fstr((int)P.energyPerPixel, 3) + "-" + fstr(P.xmax, 5) + "-" +
                                 fstr(P.ymax, 5) + "-" + fstr(P.zmax, 5) + ".dat";
        FILE *f = std::fopen(name.c_str(), "wb");
        if (f) {
            std::fwrite(jmeanGLOBAL.data(), sizeof(double), jmeanGLOBAL.size(), f);
            std::fclose(f);
            std::printf("wrote %s\n", name.c_str());
        }
    }
    tamc_unpin_host(g.rhokap.data());
    tamc_unpin_host(jmeanGLOBAL.data());
    tamc_finalize(h);
    return 0;
}
