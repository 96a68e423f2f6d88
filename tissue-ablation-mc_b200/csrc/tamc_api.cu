// tamc_api.cu -- the C ABI of libtamc.so (include/tamc.h): device residency of the grids, launch of
// the transport kernels, the tally all-reduce over NCCL, and the host<->device copies at the
// boundary with the reference's driver (/root/reference/src/mcpolar.f90:151-173).
//
// NCCL is bound at run time (dlopen) so a single-GPU caller needs no NCCL at all and a process
// that already loaded an NCCL (e.g. through torch) shares that copy.
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tamc_internal.h"

using namespace tamc;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(TAMC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL, bound lazily
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void *so = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};

static NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.so ? &api : nullptr;
    tried = true;
    // Measured on 2 x B200 (profiles/README.md, round 2): once a communicator with NVLS (NVLink SHARP multicast) and
    // cuMem-backed buffers exists in the process, the one-CTA-per-SM stub-regime kernels run 10-30 % slower with
    // rank-to-rank jitter (k_transport_column_parked 1.30 -> 1.43-1.71 ms per 1e8 packets); with both off they run at
    // their single-process speed.  The tally all-reduce is a few MB to 512 MB once per MC call and does not need either,
    // so they are switched off unless the caller's environment says otherwise (read by NCCL at its first use in the process).
    setenv("NCCL_NVLS_ENABLE", "0", 0);
    setenv("NCCL_CUMEM_ENABLE", "0", 0);
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        api.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api.so) break;
    }
    if (!api.so) return nullptr;
#define BIND(f) api.f = reinterpret_cast<decltype(api.f)>(dlsym(api.so, "nccl" #f))
    BIND(GetUniqueId); BIND(CommInitRank); BIND(AllReduce); BIND(Broadcast); BIND(CommDestroy); BIND(GetErrorString); BIND(GetVersion);
#undef BIND
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.Broadcast || !api.CommDestroy) {
        dlclose(api.so);
        api.so = nullptr;
        return nullptr;
    }
    return &api;
}

#define NC(call)                                                                                         \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess)                                                                           \
            return fail(TAMC_ENCCL, std::string(#call) + ": " +                                           \
                                        (nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r_) : "nccl error")); \
    } while (0)

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
#include "tamc_context.h"

static int check(tamc_handle h)
{
    if (!h) return fail(TAMC_EINVAL, "null handle");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(TAMC_ECUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return TAMC_OK;
}

static DevGrid make_grid(const tamc_context *c)
{
    DevGrid g{};
    g.nxg = c->nxg; g.nyg = c->nyg; g.nzg = c->nzg;
    g.sx = c->nxg + 2;
    g.sxy = (long long)(c->nxg + 2) * (c->nyg + 2);
    g.xmax = c->xmax; g.ymax = c->ymax; g.zmax = c->zmax; g.delta = c->delta;
    g.spot_r2 = (c->spot / 2.) * (c->spot / 2.);                       // sourceph.f90:28
    g.zp0 = c->zmax - (1.e-8 * (2. * c->zmax / (double)c->nzg));       // sourceph.f90:32
    g.inv_dx = (double)c->nxg / (2. * c->xmax);
    g.inv_dy = (double)c->nyg / (2. * c->ymax);
    g.inv_dz = (double)c->nzg / (2. * c->zmax);
    g.albedo = c->albedo; g.hgg = c->hgg; g.g2 = c->hgg * c->hgg;      // ch_opt.f90:16
    g.zcur0 = g.zp0 + c->zmax;
    g.cellk0 = (int)((double)c->nzg * (g.zp0 + c->zmax) / (2. * c->zmax)) + 1;
    g.flags = c->flags;
    g.n1 = c->n1; g.n2 = c->n2;
    g.r0sq = ((c->n1 - c->n2) / (c->n1 + c->n2)) * ((c->n1 - c->n2) / (c->n1 + c->n2));
    g.gauss_sigma = c->gauss_sigma;
    {   // launch_cells (tamc_fast.cuh): error bound of the fp32 pass in voxel units, see the derivation there
        const double R = c->spot / 2.;
        const double ex = 4. * (3.5e-6 * R * g.inv_dx + 6.0e-8 * (double)c->nxg + 1e-7);
        const double ey = 4. * (3.5e-6 * R * g.inv_dy + 6.0e-8 * (double)c->nyg + 1e-7);
        g.spot_r2_f = (float)g.spot_r2;
        g.inv_dx_f = (float)g.inv_dx;
        g.inv_dy_f = (float)g.inv_dy;
        g.x0_f = 0.5f * (float)c->nxg;
        g.y0_f = 0.5f * (float)c->nyg;
        g.half_x = (ex < 0.125 && c->launch32) ? (float)(0.5 - ex) : -1.f;
        g.half_y = (ey < 0.125 && c->launch32) ? (float)(0.5 - ey) : -1.f;
    }
    g.fwx = 2. * c->xmax / (double)c->nxg; g.fwy = 2. * c->ymax / (double)c->nyg; g.fwz = 2. * c->zmax / (double)c->nzg;
    g.wx = g.fwx - c->delta; g.wy = g.fwy - c->delta; g.wz = g.fwz - c->delta;
    g.ez0 = g.zcur0 - (double)(g.cellk0 - 1) * 2. * c->zmax / (double)c->nzg;      // zface(cellk0), gridset.f90:29-31
    g.sc.one_m_g2 = 1. - g.g2;                                         // stokes.f90:48
    g.sc.one_p_g2 = 1. + g.g2;
    g.sc.one_m_g = 1. - g.hgg;
    g.sc.two_g = 2. * g.hgg;
    g.sc.inv_two_g = (g.hgg != 0.) ? 1. / (2. * g.hgg) : 0.;
    g.rhokap = c->d_rhokap; g.jmean = c->d_jmean; g.faces = c->d_faces;
    g.albedo_g = c->d_albedo_g; g.hgg_g = c->d_hgg_g; g.n_g = c->d_n_g;
    return g;
}

int tamc_fail_(int code, const std::string &msg) { return fail(code, msg); }
int tamc_check_(tamc_handle h) { return check(h); }
DevGrid tamc_make_grid_(const tamc_context *c) { return make_grid(c); }

static bool host_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// ------------------------------------------------------------------------------------------------
// lifecycle
// ------------------------------------------------------------------------------------------------
extern "C" const char *tamc_last_error(void) { return g_err.c_str(); }
extern "C" int tamc_version(void) { return TAMC_VERSION; }

extern "C" int tamc_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int tamc_init(int device, int nxg, int nyg, int nzg, double xmax, double ymax, double zmax, double delta,
                         tamc_handle *out)
{
    if (!out) return fail(TAMC_EINVAL, "tamc_init: out is null");
    *out = nullptr;
    if (nxg < 1 || nyg < 1 || nzg < 1 || nxg > 4096 || nyg > 4096 || nzg > 4096)
        return fail(TAMC_EINVAL, "tamc_init: grid dimensions must be in [1,4096]");
    if ((double)(nxg + 2) * (nyg + 2) * (nzg + 2) >= 2147483648.)
        return fail(TAMC_EINVAL, "tamc_init: grid too large, (nxg+2)(nyg+2)(nzg+2) must stay below 2^31 voxels");
    if (!(xmax > 0) || !(ymax > 0) || !(zmax > 0) || !(delta > 0))
        return fail(TAMC_EINVAL, "tamc_init: xmax, ymax, zmax and delta must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(TAMC_ENODEVICE, "tamc_init: no CUDA device (libtamc has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(TAMC_ENODEVICE, "tamc_init: device ordinal out of range");
    // the snap `face -+ delta` must move the position off the face (inttau2.f90:142-166)
    if (2. * zmax + delta == 2. * zmax || 2. * xmax + delta == 2. * xmax || 2. * ymax + delta == 2. * ymax)
        return fail(TAMC_EINVAL, "tamc_init: delta is below the fp64 resolution of the face coordinates");
    // ... and land in the adjacent voxel: the production kernels re-index the crossed axis by +-1 (the reference re-derives
    // the voxel from the position, inttau2.f90:190-239), which is the same thing only while delta is a small part of an edge
    {
        const double wmin = fmin(fmin(2. * xmax / nxg, 2. * ymax / nyg), 2. * zmax / nzg);
        if (!(delta < 1e-3 * wmin))
            return fail(TAMC_EINVAL, "tamc_init: delta must stay below 1e-3 of the smallest voxel edge (the reference uses 1e-8 of it, mcpolar.f90:112)");
    }

    CU(cudaSetDevice(device));
    tamc_context *c = new tamc_context();
    c->device = device;
    c->nxg = nxg; c->nyg = nyg; c->nzg = nzg;
    c->xmax = xmax; c->ymax = ymax; c->zmax = zmax; c->delta = delta;
    c->n_rhokap = (size_t)(nxg + 2) * (nyg + 2) * (nzg + 2);
    c->n_jmean = (size_t)nxg * nyg * nzg;
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    c->num_sms = prop.multiProcessorCount;
    c->cfg.num_sms = c->num_sms;

    // gridset.f90:23-31: face(i) = (i-1) * 2. * max/n, left to right
    std::vector<double> faces;
    faces.reserve((size_t)nxg + nyg + nzg + 3);
    for (int i = 1; i <= nxg + 1; ++i) faces.push_back((double)(i - 1) * 2. * xmax / (double)nxg);
    for (int i = 1; i <= nyg + 1; ++i) faces.push_back((double)(i - 1) * 2. * ymax / (double)nyg);
    for (int i = 1; i <= nzg + 1; ++i) faces.push_back((double)(i - 1) * 2. * zmax / (double)nzg);

    auto cleanup = [&](int code) { tamc_finalize(c); return code; };
#define CUI(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            fail(TAMC_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                        \
            return cleanup(TAMC_ECUDA);                                                                  \
        }                                                                                                \
    } while (0)
    CUI(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < EV_N; ++i) CUI(cudaEventCreate(&c->ev[i]));
    CUI(cudaMalloc(&c->d_rhokap, c->n_rhokap * sizeof(double)));
    // the tally and, right behind it, the call's counters: one clear per MC call covers both
    CUI(cudaMalloc(&c->d_jmean, (c->n_jmean + CNT_N) * sizeof(double)));
    c->d_cnt = reinterpret_cast<unsigned long long *>(c->d_jmean + c->n_jmean);
    CUI(cudaMalloc(&c->d_faces, faces.size() * sizeof(double)));
    CUI(cudaMemcpy(c->d_faces, faces.data(), faces.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUI(cudaMemset(c->d_rhokap, 0, c->n_rhokap * sizeof(double)));
    CUI(cudaMemset(c->d_jmean, 0, c->n_jmean * sizeof(double)));
    CUI(cudaMemset(c->d_cnt, 0, CNT_N * sizeof(unsigned long long)));
#undef CUI
    *out = c;
    return TAMC_OK;
}

// "peer_reduce": give the buffers back.  The other ranks read this rank's buffer in their own k_peer_box_reduce, which may
// still be running when this rank's stream has drained: wait (bounded) until each of them has flagged the last call done.
static void peer_release(tamc_handle h)
{
    if (h->peer_state == 1 && h->peer_base) {
        unsigned long long flags[tamc::kPeerCounter];
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            if (cudaMemcpy(flags, h->peer_base, sizeof(flags), cudaMemcpyDeviceToHost) != cudaSuccess) break;
            bool done = true;
            for (int r = 0; r < h->nranks; ++r)
                if (r != h->rank && flags[tamc::kPeerDone + r] < h->peer_call) done = false;
            if (done || std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 2.0) break;
        }
    }
    for (int r = 0; r < tamc::kPeerMaxRanks; ++r)
        if (h->peer_open[r]) { cudaIpcCloseMemHandle(h->peer_open[r]); h->peer_open[r] = nullptr; }
    if (h->peer_base) { cudaFree(h->peer_base); h->peer_base = nullptr; }
    if (h->h_peer_err) { cudaFreeHost(h->h_peer_err); h->h_peer_err = nullptr; h->d_peer_err = nullptr; }
    h->peer_state = 0;
    cudaGetLastError();
}

extern "C" int tamc_finalize(tamc_handle h)
{
    if (!h) return TAMC_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    peer_release(h);
    if (h->comm && nccl_api()) nccl_api()->CommDestroy(h->comm);
    tamc_heat_release_(h);
    cudaFree(h->d_rhokap); cudaFree(h->d_jmean); cudaFree(h->d_faces); cudaFree(h->d_flush);
    cudaFree(h->d_albedo_g); cudaFree(h->d_hgg_g); cudaFree(h->d_n_g); cudaFree(h->colws.optc);
    cudaFree(h->colws.stops); cudaFree(h->colws.rkT); cudaFree(h->colws.dense); cudaFree(h->colws.vox);
    if (h->colws.s_side) { cudaStreamSynchronize(h->colws.s_side); cudaStreamDestroy(h->colws.s_side); }
    if (h->colws.h_bound) cudaFreeHost(h->colws.h_bound);
    cudaFree(h->colws.bound_scratch);
    if (h->colws.ev_bound) cudaEventDestroy(h->colws.ev_bound);
    if (h->colws.ev_gathered) cudaEventDestroy(h->colws.ev_gathered);
    if (h->s_up) { cudaStreamSynchronize(h->s_up); cudaStreamDestroy(h->s_up); }
    if (h->s_dn) { cudaStreamSynchronize(h->s_dn); cudaStreamDestroy(h->s_dn); }
    cudaFree(h->d_zero); cudaFree(h->d_box_rk); cudaFree(h->d_path);
    for (int i = 0; i < EV_N; ++i)
        if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    cudaGetLastError();
    delete h;
    return TAMC_OK;
}

extern "C" int tamc_set_source_co2(tamc_handle h, double spot_diameter_cm)
{
    if (int rc = check(h)) return rc;
    if (!(spot_diameter_cm >= 0)) return fail(TAMC_EINVAL, "tamc_set_source_co2: negative spot diameter");
    // the launch formula (sourceph.f90:45-46) indexes outside the grid if the spot overhangs it
    if (spot_diameter_cm / 2. >= h->xmax || spot_diameter_cm / 2. >= h->ymax)
        return fail(TAMC_EINVAL, "tamc_set_source_co2: spot does not fit on the top face");
    h->spot = spot_diameter_cm;
    h->gauss_sigma = 0.;
    return TAMC_OK;
}

extern "C" int tamc_set_source_gaussian(tamc_handle h, double sigma_cm)
{
    if (int rc = check(h)) return rc;
    if (!(sigma_cm > 0.) || !(sigma_cm < 1e300)) return fail(TAMC_EINVAL, "tamc_set_source_gaussian: sigma must be positive and finite");
    // each variate is redrawn while it misses the top face: keep the acceptance sane (sigma = 20 half-widths accepts ~4 %)
    if (sigma_cm > 20. * h->xmax || sigma_cm > 20. * h->ymax)
        return fail(TAMC_EINVAL, "tamc_set_source_gaussian: sigma beyond 20 half-widths of the top face");
    h->gauss_sigma = sigma_cm;
    return TAMC_OK;
}

static int check_optics(tamc_handle h, const double *rhokap, double albedo, double hgg, double n1, double n2, int flags, const char *who)
{
    const std::string w(who);
    if (!(albedo >= 0. && albedo <= 1.)) return fail(TAMC_EINVAL, w + ": albedo must be in [0,1]");
    if (!(hgg > -1. && hgg < 1.)) return fail(TAMC_EINVAL, w + ": hgg must be in (-1,1)");
    if (flags & ~(TAMC_SCATTER | TAMC_FRESNEL | TAMC_PERIODIC)) return fail(TAMC_EINVAL, w + ": unknown flag bits");
    if ((flags & TAMC_FRESNEL) && !(n1 > 0. && n2 > 0.)) return fail(TAMC_EINVAL, w + ": TAMC_FRESNEL needs positive n1, n2");
    if (!rhokap && !h->optics_set) return fail(TAMC_ESTATE, w + ": first call needs the rhokap grid");
    return TAMC_OK;
}

// root_io: after an overlapped columns-first call only rank 0 holds the new opacity grid in full (the other GPUs got the
// beam's columns, all that call read).  The next use of the RESIDENT grid -- a call without a new rhokap, a kernel form
// that walks the resident grid, the heat step -- first brings the other ranks up to date over NVLink.  The flag is set and
// cleared by the same calls on every rank, so the broadcast stays collective.
static int sync_resident(tamc_handle h)
{
    if (!h->resident_behind) return TAMC_OK;
    h->resident_behind = false;
    if (h->comm && h->nranks > 1)
        NC(nccl_api()->Broadcast(h->d_rhokap, h->d_rhokap, h->n_rhokap, ncclDouble, 0, h->comm, h->stream));
    return TAMC_OK;
}
int tamc_sync_resident_(tamc_handle h) { return sync_resident(h); }

// full-grid upload on the handle's stream, not synchronised
static int enqueue_upload(tamc_handle h, const double *rhokap)
{
    CU(cudaEventRecord(h->ev[EV_H0], h->stream));
    // stream-ordered for pageable memory too (the runtime stages it and returns once the caller's buffer is free again): a
    // legacy-stream cudaMemcpy is not ordered against the handle's non-blocking stream
    CU(cudaMemcpyAsync(h->d_rhokap, rhokap, h->n_rhokap * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    if (!host_is_pinned(rhokap)) CU(cudaStreamSynchronize(h->stream));
    CU(cudaEventRecord(h->ev[EV_H1], h->stream));
    h->timed_h2d = true;
    return TAMC_OK;
}

extern "C" int tamc_set_optics(tamc_handle h, const double *rhokap, double albedo, double hgg, double n1, double n2,
                               int flags)
{
    if (int rc = check(h)) return rc;
    if (int rc = check_optics(h, rhokap, albedo, hgg, n1, n2, flags, "tamc_set_optics")) return rc;
    h->timed_h2d = false;
    if (rhokap) {
        const bool rooted = h->root_io && h->comm && h->nranks > 1;
        if (!rooted || h->rank == 0) { if (int rc = enqueue_upload(h, rhokap)) return rc; }
        if (rooted) NC(nccl_api()->Broadcast(h->d_rhokap, h->d_rhokap, h->n_rhokap, ncclDouble, 0, h->comm, h->stream));
        h->resident_behind = false;
        // the caller may rewrite rhokap as soon as this returns (no host pointer is kept past the call)
        CU(cudaStreamSynchronize(h->stream));
    }
    h->albedo = albedo; h->hgg = hgg; h->n1 = n1; h->n2 = n2; h->flags = flags;
    h->optics_set = true;
    return TAMC_OK;
}

// EXTENSION (no upstream counterpart: opt_prop.f90:5 holds scalars): per-voxel albedo / hgg / refractive index.
extern "C" int tamc_set_optics_grids(tamc_handle h, const double *albedo, const double *hgg, const double *n)
{
    if (int rc = check(h)) return rc;
    const double *src[3] = {albedo, hgg, n};
    double **dst[3] = {&h->d_albedo_g, &h->d_hgg_g, &h->d_n_g};
    const char *what[3] = {"albedo must be in [0,1]", "hgg must be in (-1,1)", "refractive index must be positive"};
    for (int i = 0; i < 3; ++i) {
        if (!src[i]) continue;
        // the halo is never read; the interior has to be a valid optical property
        for (int k = 1; k <= h->nzg; ++k)
            for (int j = 1; j <= h->nyg; ++j) {
                const double *row = src[i] + (size_t)(h->nxg + 2) * ((size_t)j + (size_t)(h->nyg + 2) * (size_t)k);
                for (int ii = 1; ii <= h->nxg; ++ii) {
                    const double v = row[ii];
                    const bool ok = i == 0 ? (v >= 0. && v <= 1.) : (i == 1 ? (v > -1. && v < 1.) : (v > 0. && v < 1e300));
                    if (!ok) return fail(TAMC_EINVAL, std::string("tamc_set_optics_grids: ") + what[i]);
                }
            }
    }
    CU(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 3; ++i) {
        if (!src[i]) {
            cudaFree(*dst[i]);
            *dst[i] = nullptr;
            continue;
        }
        if (!*dst[i]) CU(cudaMalloc(dst[i], h->n_rhokap * sizeof(double)));
        CU(cudaMemcpyAsync(*dst[i], src[i], h->n_rhokap * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return TAMC_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU
// ------------------------------------------------------------------------------------------------
extern "C" int tamc_comm_unique_id(void *id128)
{
    if (!id128) return fail(TAMC_EINVAL, "tamc_comm_unique_id: null buffer");
    NcclApi *n = nccl_api();
    if (!n) return fail(TAMC_ENCCL, "libnccl.so.2 not found");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    NC(n->GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return TAMC_OK;
}

extern "C" int tamc_comm_init(tamc_handle h, int nranks, int rank, const void *id128)
{
    if (int rc = check(h)) return rc;
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) return fail(TAMC_EINVAL, "tamc_comm_init: bad rank/size/id");
    if (h->comm) return fail(TAMC_ESTATE, "tamc_comm_init: communicator already initialised");
    NcclApi *n = nccl_api();
    if (!n) return fail(TAMC_ENCCL, "libnccl.so.2 not found");
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NC(n->CommInitRank(&h->comm, nranks, id, rank));
    h->nranks = nranks;
    h->rank = rank;
    // Every rank must describe the same grid and the same reduction (mcpolar.f90:173 sums arrays of one size): compare a
    // fingerprint of what shapes the all-reduce -- the max over the ranks of each word and of its negative must coincide.
    if (nranks > 1) {
        const double fp[16] = {(double)h->nxg, (double)h->nyg, (double)h->nzg, h->xmax, h->ymax, h->zmax, h->delta, h->spot,
                               h->gauss_sigma, (double)h->reduce, (double)h->box_reduce, (double)h->reduce_bound,
                               (double)TAMC_VERSION, (double)h->root_io, (double)h->peer_reduce, 0.};
        double *d_fp = nullptr;
        CU(cudaMalloc(&d_fp, 32 * sizeof(double)));
        double both[32];
        for (int i = 0; i < 16; ++i) { both[i] = fp[i]; both[16 + i] = -fp[i]; }
        CU(cudaMemcpyAsync(d_fp, both, sizeof(both), cudaMemcpyHostToDevice, h->stream));
        const ncclResult_t r = n->AllReduce(d_fp, d_fp, 32, ncclDouble, ncclMax, h->comm, h->stream);
        cudaError_t e = cudaMemcpyAsync(both, d_fp, sizeof(both), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        cudaFree(d_fp);
        if (r != ncclSuccess) return fail(TAMC_ENCCL, "tamc_comm_init: fingerprint all-reduce failed");
        if (e != cudaSuccess) return fail(TAMC_ECUDA, std::string("tamc_comm_init: ") + cudaGetErrorString(e));
        for (int i = 0; i < 16; ++i)
            if (both[i] != -both[16 + i])
                return fail(TAMC_ESTATE, "tamc_comm_init: the ranks disagree on the grid, the source or the reduction options (word " +
                                             std::to_string(i) + " of the fingerprint): every rank must set them identically before tamc_comm_init");
    }
    return TAMC_OK;
}

// ------------------------------------------------------------------------------------------------
// the hot path
// ------------------------------------------------------------------------------------------------
// "peer_reduce": one-time set-up of the buffers the ranks read from each other (tamc_peer.cuh).  Collective: every rank
// reaches it in the same call (the option is part of tamc_comm_init's fingerprint, the element count is the same on
// every rank).  Any rank that cannot export or map a buffer -- ranks in one process, no peer access, another node -- and
// all ranks agree to stay with ncclAllReduce (peer_state = -1).
static int peer_setup(tamc_handle h, size_t elems_needed)
{
    h->peer_state = -1;
    const int nr = h->nranks;
    if (nr > tamc::kPeerMaxRanks) return TAMC_OK;
    NcclApi *n = nccl_api();
    int mine_ok = 1;
    const size_t elems = (elems_needed + 1) & ~(size_t)1;
    const size_t bytes = tamc::kPeerFlagBytes + 2 * elems * sizeof(double);
    cudaIpcMemHandle_t hm;
    memset(&hm, 0, sizeof(hm));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the handle table below carries 16 words per rank");
    if (cudaMalloc(&h->peer_base, bytes) != cudaSuccess) { cudaGetLastError(); h->peer_base = nullptr; mine_ok = 0; }
    if (mine_ok && cudaMemsetAsync(h->peer_base, 0, tamc::kPeerFlagBytes, h->stream) != cudaSuccess) { cudaGetLastError(); mine_ok = 0; }
    if (mine_ok && cudaIpcGetMemHandle(&hm, h->peer_base) != cudaSuccess) { cudaGetLastError(); mine_ok = 0; }
    if (mine_ok && !h->h_peer_err) {
        if (cudaHostAlloc((void **)&h->h_peer_err, sizeof(unsigned int), cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer((void **)&h->d_peer_err, h->h_peer_err, 0) != cudaSuccess) {
            cudaGetLastError();
            mine_ok = 0;
        } else
            *h->h_peer_err = 0u;
    }
    // every rank's handle (16 words) + its "fine so far": an all-reduce(sum) of a table that is zero outside the own row
    constexpr int W = 17;
    int table[tamc::kPeerMaxRanks * W];
    memset(table, 0, sizeof(table));
    memcpy(&table[h->rank * W], &hm, sizeof(hm));
    table[h->rank * W + 16] = mine_ok;
    int *d_tab = nullptr;
    CU(cudaMalloc(&d_tab, sizeof(table)));
    CU(cudaMemcpyAsync(d_tab, table, sizeof(table), cudaMemcpyHostToDevice, h->stream));
    ncclResult_t r = n->AllReduce(d_tab, d_tab, tamc::kPeerMaxRanks * W, ncclInt, ncclSum, h->comm, h->stream);
    cudaError_t e = cudaMemcpyAsync(table, d_tab, sizeof(table), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (r != ncclSuccess || e != cudaSuccess) { cudaFree(d_tab); return fail(TAMC_ENCCL, "peer_reduce: exchanging the buffer handles failed"); }
    int all_ok = 1;
    for (int q = 0; q < nr; ++q) all_ok &= table[q * W + 16] == 1;
    if (all_ok) {
        for (int q = 0; q < nr && all_ok; ++q) {
            if (q == h->rank) continue;
            cudaIpcMemHandle_t hq;
            memcpy(&hq, &table[q * W], sizeof(hq));
            if (cudaIpcOpenMemHandle(&h->peer_open[q], hq, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                h->peer_open[q] = nullptr;
                all_ok = 0;
            }
        }
    }
    // ... and whether every rank could map every other rank's buffer
    int agreed = all_ok;
    CU(cudaMemcpyAsync(d_tab, &agreed, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    r = n->AllReduce(d_tab, d_tab, 1, ncclInt, ncclMin, h->comm, h->stream);
    e = cudaMemcpyAsync(&agreed, d_tab, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_tab);
    if (r != ncclSuccess || e != cudaSuccess) return fail(TAMC_ENCCL, "peer_reduce: agreeing on the buffer mappings failed");
    if (agreed != 1) {
        peer_release(h);
        h->peer_state = -1;
        return TAMC_OK;
    }
    h->peer_elems = elems;
    h->peer_call = 0;
    h->peer_state = 1;
    return TAMC_OK;
}

static int enqueue_reduce(tamc_handle h)
{
    h->timed_reduce = false;
    // the depth bound of this call (k_column_bound), if one was asked for: read as soon as that small kernel has run --
    // the transport is still going.  Always consumed here, so the next call's kernel never meets an unread answer.
    int bound = 0;
    if (h->colws.bound_pending) {
        CU(cudaEventSynchronize(h->colws.ev_bound));
        bound = *(volatile int *)h->colws.h_bound;
        h->colws.bound_pending = false;
    }
    h->reduce_planes = bound;
    if (h->comm && h->reduce && h->nranks > 1) {
        // mcpolar.f90:173: MPI_allREDUCE(jmean, jmeanGLOBAL, nxg*nyg*nzg, MPI_DOUBLE_PRECISION, MPI_SUM)
        // Shipped regime (no scatter loop): every flight is straight down, so the tally is zero outside the columns under
        // the beam's bounding box on every rank -- reduce only those, and only down to the depth bound of the grid.  The
        // element count must be the same on every rank: it depends on the optics flags, the grid, the spot and the options
        // "reduce" / "box_reduce" / "reduce_bound" (checked against the other ranks by tamc_comm_init and to be kept equal
        // afterwards), and NOT on the packet count or the kernel form a rank ran: the bound is computed for every form
        // (k_column_bound on the gathered copy, k_column_bound_resident otherwise) whenever the box reduce is on.
        const DevGrid g = make_grid(h);
        ColGeom cg;
        const bool box = h->box_reduce != 0 && !(h->flags & (TAMC_SCATTER | TAMC_FRESNEL)) && beam_box(g, cg) &&
                         (h->box_reduce > 0 || 2 * (size_t)cg.tw * cg.th <= (size_t)h->nxg * h->nyg);
        h->reduce_planes = 0;
        if (box) {
            // how many planes below the top face can hold anything (the same number on every rank: same grid)
            int planes = h->nzg;
            if (bound > 0 && bound < planes) planes = bound;
            const int kz0 = h->nzg - planes;
            h->reduce_planes = planes;
            const size_t cnt = (size_t)cg.tw * cg.th * (size_t)planes;
            if (h->colws.dense_elems < cnt) {
                cudaFree(h->colws.dense);
                h->colws.dense = nullptr;
                h->colws.dense_elems = 0;
                CU(cudaMalloc(&h->colws.dense, cnt * sizeof(double)));
                h->colws.dense_elems = cnt;
            }
            // "peer_reduce": the few MB of the box summed straight out of the other ranks' buffers (tamc_peer.cuh) -- no NCCL
            // kernel between two transport kernels; buffers sized once for every plane of the box
            if (h->peer_reduce && h->peer_state == 0) {
                if (int rc = peer_setup(h, (size_t)cg.tw * cg.th * (size_t)h->nzg)) return rc;
            }
            if (h->peer_reduce && h->peer_state == 1 && cnt <= h->peer_elems) {
                const unsigned long long call = ++h->peer_call;
                const size_t half = (size_t)(call & 1ull) * h->peer_elems;
                tamc::PeerSet ps{};
                for (int q = 0; q < h->nranks; ++q) {
                    char *base = (char *)(q == h->rank ? h->peer_base : h->peer_open[q]);
                    ps.flags[q] = reinterpret_cast<unsigned long long *>(base);
                    ps.buf[q] = reinterpret_cast<const double *>(base + tamc::kPeerFlagBytes) + half;
                }
                CU(launch_box_copy(g, cg, const_cast<double *>(ps.buf[h->rank]), false, h->num_sms, h->stream, kz0));
                CU(launch_peer_box_reduce(ps, h->colws.dense, cnt, h->nranks, h->rank, call, h->d_peer_err, h->num_sms, h->stream));
            } else {
                CU(launch_box_copy(g, cg, h->colws.dense, false, h->num_sms, h->stream, kz0));
                NC(nccl_api()->AllReduce(h->colws.dense, h->colws.dense, cnt, ncclDouble, ncclSum, h->comm, h->stream));
            }
            CU(launch_box_copy(g, cg, h->colws.dense, true, h->num_sms, h->stream, kz0));
        } else {
            NC(nccl_api()->AllReduce(h->d_jmean, h->d_jmean, h->n_jmean, ncclDouble, ncclSum, h->comm, h->stream));
        }
        h->timed_reduce = true;
    }
    CU(cudaEventRecord(h->ev[EV_AR1], h->stream));
    return TAMC_OK;
}

// tally clear + transport + all-reduce on the handle's stream
static int enqueue_mc(tamc_handle h, int64_t nphotons, int64_t seed, int64_t first_packet_id)
{
    if (!h->optics_set) return fail(TAMC_ESTATE, "tamc_run: tamc_set_optics has not been called");
    if (nphotons < 0 || nphotons > ((int64_t)1 << 46)) return fail(TAMC_EINVAL, "tamc_run: nphotons out of range");
    int64_t first = first_packet_id;
    if (first < 0) {
        first = h->cursor + (int64_t)h->rank * nphotons;     // rank r runs its own nphotons (mcpolar.f90:151)
        h->cursor += (int64_t)h->nranks * nphotons;
    }
    const DevGrid g = make_grid(h);
    int launches = 0;
    CU(cudaEventRecord(h->ev[EV_ZERO0], h->stream));
    static_assert(sizeof(unsigned long long) == sizeof(double), "the counters sit behind the tally in one allocation");
    CU(cudaMemsetAsync(h->d_jmean, 0, (h->n_jmean + CNT_N) * sizeof(double), h->stream));   // zarray / jmean = 0. (mcpolar.f90:185) + counters
    CU(cudaEventRecord(h->ev[EV_K0], h->stream));
    LaunchCfg cfg = h->cfg;
    // ("reduce_bound" = 2: also without a communicator, for diagnostics -- the answer lands in "reduce_planes")
    const bool stub = !(h->flags & (TAMC_SCATTER | TAMC_FRESNEL));
    cfg.want_bound = (stub && ((h->comm && h->reduce && h->nranks > 1 && h->box_reduce != 0 && h->reduce_bound) || h->reduce_bound == 2)) ? 1 : 0;
    h->colws.bound_pending = false;
    if (cfg.want_bound && !h->colws.s_side) {       // the bound kernel runs beside the transport, on a stream of its own
        if (cudaStreamCreateWithFlags(&h->colws.s_side, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); h->colws.s_side = nullptr; }
    }
    CU(launch_transport(g, cfg, nphotons, (uint64_t)seed, (uint64_t)first, h->d_cnt, nullptr, h->stream, &launches, &h->colws, &h->form));
    CU(cudaEventRecord(h->ev[EV_K1], h->stream));
    if (int rc = enqueue_reduce(h)) return rc;
    h->colws.bound_pending = false;
    h->last_launches = launches;
    h->timed_d2h = false;
    h->ran = true;
    return TAMC_OK;
}

extern "C" int tamc_run_async(tamc_handle h, int64_t nphotons, int64_t seed, int64_t first_packet_id)
{
    if (int rc = check(h)) return rc;
    if (int rc = sync_resident(h)) return rc;
    return enqueue_mc(h, nphotons, seed, first_packet_id);
}

extern "C" int tamc_sync(tamc_handle h)
{
    if (int rc = check(h)) return rc;
    CU(cudaStreamSynchronize(h->stream));
    return TAMC_OK;
}

extern "C" int tamc_get_jmean(tamc_handle h, double *jmean_global)
{
    if (int rc = check(h)) return rc;
    if (!jmean_global) return fail(TAMC_EINVAL, "tamc_get_jmean: null destination");
    CU(cudaEventRecord(h->ev[EV_D0], h->stream));
    CU(cudaMemcpyAsync(jmean_global, h->d_jmean, h->n_jmean * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaEventRecord(h->ev[EV_D1], h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->timed_d2h = true;
    return TAMC_OK;
}

extern "C" int tamc_get_stats(tamc_handle h, tamc_stats *st)
{
    if (int rc = check(h)) return rc;
    if (!st) return fail(TAMC_EINVAL, "tamc_get_stats: null destination");
    memset(st, 0, sizeof(*st));
    if (!h->ran) return TAMC_OK;
    CU(cudaStreamSynchronize(h->stream));
    unsigned long long cnt[CNT_N];
    CU(cudaMemcpy(cnt, h->d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
    st->packets = (int64_t)cnt[CNT_PACKETS];
    st->voxel_steps = (int64_t)cnt[CNT_STEPS];
    // feeds the launch plan of the next stub-regime call (LaunchCfg::column_park)
    if (!(h->flags & (TAMC_SCATTER | TAMC_FRESNEL)) && cnt[CNT_PACKETS] >= 10000)
        h->cfg.steps_hint = (double)cnt[CNT_STEPS] / (double)cnt[CNT_PACKETS];
    // ... and the depth limit of the next columns-first upload (LaunchCfg::gather_depth); only the column form measures it
    const bool column_form = h->form == FORM_COLUMN || h->form == FORM_COLUMN_TILED || h->form == FORM_COLUMN_PARKED;
    h->cfg.depth_hint = (column_form && cnt[CNT_PACKETS] >= 10000) ? (int)cnt[CNT_DEPTH] : 0;
    st->scatters = (int64_t)cnt[CNT_SCATTERS];
    st->absorbed = (int64_t)cnt[CNT_ABSORBED];
    for (int f = 0; f < 6; ++f) st->exits[f] = (int64_t)cnt[CNT_EXIT0 + f];
    st->specular = (int64_t)cnt[CNT_SPECULAR];
    st->internal_reflections = (int64_t)cnt[CNT_REFLECT];
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev[EV_ZERO0], h->ev[EV_K0])); st->zero_ms = ms;
    CU(cudaEventElapsedTime(&ms, h->ev[EV_K0], h->ev[EV_K1])); st->kernel_ms = ms;
    if (h->timed_reduce) { CU(cudaEventElapsedTime(&ms, h->ev[EV_K1], h->ev[EV_AR1])); st->allreduce_ms = ms; }
    if (h->timed_h2d) { CU(cudaEventElapsedTime(&ms, h->ev[EV_H0], h->ev[EV_H1])); st->h2d_ms = ms; }
    if (h->timed_h2d && (h->io_form & 2)) st->kernel_ms -= st->h2d_ms;     // k_column_gather over PCIe sits inside K0..K1
    if (h->timed_d2h) { CU(cudaEventElapsedTime(&ms, h->ev[EV_D0], h->ev[EV_D1])); st->d2h_ms = ms; }
    st->gpu_launches = h->last_launches;
    if (h->h_peer_err && *(volatile unsigned int *)h->h_peer_err) {
        *h->h_peer_err = 0u;
        return fail(TAMC_ENCCL, "peer_reduce: another rank's box did not arrive within 20 s; the tally of this call is not reduced");
    }
    if (cnt[CNT_ERRORS])
        return fail(TAMC_EINVAL, "transport: " + std::to_string(cnt[CNT_ERRORS]) + " packet(s) exceeded the voxel-step cap");
    return TAMC_OK;
}

extern "C" int tamc_seek(tamc_handle h, int64_t next_packet_id)
{
    if (!h || next_packet_id < 0) return fail(TAMC_EINVAL, "tamc_seek: bad argument");
    h->cursor = next_packet_id;
    return TAMC_OK;
}

// ------------------------------------------------------------------------------------------------
// The boundary with the host driver, overlapped (shipped regime).
//
// Without the scatter loop every flight is straight down (sourceph.f90:37-42, mcpolar.f90:166-169), so
//   * the tally is zero outside the columns under the beam's bounding box -- jmeanGLOBAL is written as a zero fill
//     (device zeros -> host, a DMA on its own stream WHILE the transport runs) followed by the box columns after the
//     all-reduce (k_box_mirror stores them straight into the caller's array, skipping rows of the box that hold only
//     zeros; long calls on grids whose deposits reach deep use one pitched 3-D DMA whose descriptors are submitted while
//     the transport runs): the same bytes in host memory as the plain full-grid download;
//   * the column form reads the opacities of those columns only -- in tamc_run_optics k_column_gather fetches them
//     straight from the caller's array (each voxel once) and the DMA of the full grid, which later calls and the heat
//     step read, follows on a second stream beside the transport.
// The kernels access page-locked host memory over PCIe (zero-copy) in 256-byte row segments: no per-slice DMA
// descriptors to submit (a pitched 3-D cudaMemcpy of the box costs ~0.3 ms of host time at 200^3) ahead of the transport.
// PCIe is full duplex, so the step costs  box up + transport + box down  instead of  grid up + transport + grid down.
// Needs page-locked host arrays (tamc_pin_host); anything else takes the plain sequential path.
// ------------------------------------------------------------------------------------------------
// root_io: rank 0's copies of the beam's columns to the other GPUs (called by the column set-up, tamc_kernels.cu)
static cudaError_t share_broadcast(void *comm, double *buf, size_t count, cudaStream_t s)
{
    NcclApi *n = nccl_api();
    if (!n || !comm) return cudaErrorInvalidValue;
    return n->Broadcast(buf, buf, count, ncclDouble, 0, (ncclComm_t)comm, s) == ncclSuccess ? cudaSuccess : cudaErrorUnknown;
}

static int side_streams(tamc_handle h)
{
    if (!h->s_up) CU(cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
    if (!h->s_dn) CU(cudaStreamCreateWithFlags(&h->s_dn, cudaStreamNonBlocking));
    if (!h->d_zero) {
        CU(cudaMalloc(&h->d_zero, h->n_jmean * sizeof(double)));
        CU(cudaMemsetAsync(h->d_zero, 0, h->n_jmean * sizeof(double), h->s_dn));
    }
    return TAMC_OK;
}

static bool box_io_wanted(tamc_handle h, int flags, ColGeom &cg)
{
    if (h->box_io == 0 || (flags & (TAMC_SCATTER | TAMC_FRESNEL))) return false;
    const DevGrid g = make_grid(h);
    return beam_box(g, cg) && 2 * (size_t)cg.tw * cg.th <= (size_t)h->nxg * h->nyg;
}

// TAMC_TRACE=1: host-clock marks of one boundary call on stderr (where a call spends its wall time)
struct CallTrace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    char buf[512];
    int len = 0;
    CallTrace() : on(getenv("TAMC_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what)
    {
        if (!on) return;
        const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
        len += snprintf(buf + len, sizeof(buf) - (size_t)len, " %s=%.0f", what, us);
    }
    void flush(int rank) { if (on) fprintf(stderr, "[tamc trace r%d]%s\n", rank, buf); }
};

static int run_boundary(tamc_handle h, const double *rhokap, int64_t nphotons, int64_t seed, double *jmean_global, tamc_stats *stats)
{
    CallTrace tr;
    ColGeom cg{};
    const DevGrid g = make_grid(h);
    // device-visible addresses of the caller's page-locked arrays (cudaHostRegister / cudaHostAlloc memory is mapped)
    double *jm_dev = nullptr;
    const double *rk_dev = nullptr;
    // "root_io" (several ranks): only rank 0 touches host arrays.  Its rhokap goes up once and reaches the other GPUs over
    // NVLink (ncclBroadcast), and only its jmeanGLOBAL is written -- the reference's ranks all hold identical copies of both
    // (3dFD.f90:312-361 runs on every rank, and only rank 0's jmeanGLOBAL is read, :95), so nothing is lost and the host's
    // PCIe / memory system carries 130 MB per call instead of nranks x 130 MB.  Every rank must make the same call with the
    // same `rhokap != NULL`; on ranks > 0 the contents of rhokap are never read and jmean_global may be NULL.
    const bool rooted = h->root_io && h->comm && h->nranks > 1;
    const bool root = !rooted || h->rank == 0;
    bool box_dn = box_io_wanted(h, h->flags, cg) && (root ? host_is_pinned(jmean_global) : true);
    if (box_dn && root && cudaHostGetDevicePointer((void **)&jm_dev, jmean_global, 0) != cudaSuccess) { cudaGetLastError(); box_dn = false; }
    bool box_up = rhokap && box_dn && (root ? host_is_pinned(rhokap) : true) && column_gather_selected(g, h->cfg, nphotons);
    if (box_up && root && cudaHostGetDevicePointer((void **)&rk_dev, const_cast<double *>(rhokap), 0) != cudaSuccess) { cudaGetLastError(); box_up = false; }
    if (rooted) {
        // the ranks must take the same path (the collectives below differ): rank 0 decides, from what its host arrays allow
        int path[2] = {box_dn ? 1 : 0, box_up ? 1 : 0};
        if (!h->d_path) CU(cudaMalloc(&h->d_path, 2 * sizeof(int)));
        CU(cudaMemcpyAsync(h->d_path, path, sizeof(path), cudaMemcpyHostToDevice, h->stream));
        NC(nccl_api()->Broadcast(h->d_path, h->d_path, 2, ncclInt, 0, h->comm, h->stream));
        CU(cudaMemcpyAsync(path, h->d_path, sizeof(path), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        box_dn = path[0] != 0;
        box_up = path[1] != 0;
        tr.mark("path");
    }
    h->io_form = (box_dn ? 1 : 0) | (box_up ? 2 : 0) | (rooted ? 8 : 0);
    if (rhokap) h->timed_h2d = false;   // otherwise keep the upload time of the tamc_set_optics before this call

    if (box_dn && root) {
        if (int rc = side_streams(h)) return rc;
        // nothing of this call may overtake work already queued on the handle's stream
        CU(cudaEventRecord(h->ev[EV_FORK], h->stream));
        CU(cudaStreamWaitEvent(h->s_dn, h->ev[EV_FORK], 0));
        if (!box_up) {      // with a columns-first upload the fill starts behind it (below): measured, the two slow each other
            CU(cudaMemcpyAsync(jmean_global, h->d_zero, h->n_jmean * sizeof(double), cudaMemcpyDeviceToHost, h->s_dn));
            CU(cudaEventRecord(h->ev[EV_DN], h->s_dn));
        }
    }
    if (box_up) {
        const size_t need = (size_t)cg.tw * cg.th * (size_t)h->nzg;
        if (h->box_rk_elems < need) {
            cudaFree(h->d_box_rk);
            h->d_box_rk = nullptr;
            h->box_rk_elems = 0;
            CU(cudaMalloc(&h->d_box_rk, need * sizeof(double)));
            h->box_rk_elems = need;
        }
        // k_column_gather reads the beam's columns straight from the caller's array and k_column_finish the copy it keeps
        h->colws.gather_src = rk_dev;
        h->colws.box_rk = h->d_box_rk;
        h->colws.ev_gather0 = h->ev[EV_H0];
        h->colws.ev_gather1 = h->ev[EV_H1];
        // root_io: rank 0 gathers every plane (the other ranks have no host array to fall back on below a depth limit) and
        // hands both copies of the columns to the other GPUs over NVLink; those skip their gather
        h->colws.share_gather = rooted ? (root ? 1 : 2) : 0;
        h->colws.share_comm = rooted ? (void *)h->comm : nullptr;
        h->colws.share_fn = share_broadcast;
        // "io_early": start the full-grid upload (bit0) / the zero fill (bit1) at once, beside the gather, instead of behind it
        if (root && (h->io_early & 1)) {
            CU(cudaStreamWaitEvent(h->s_up, h->ev[EV_FORK], 0));
            CU(cudaMemcpyAsync(h->d_rhokap, rhokap, h->n_rhokap * sizeof(double), cudaMemcpyHostToDevice, h->s_up));
            CU(cudaEventRecord(h->ev[EV_UP], h->s_up));
        }
        if (root && (h->io_early & 2)) {
            CU(cudaMemcpyAsync(jmean_global, h->d_zero, h->n_jmean * sizeof(double), cudaMemcpyDeviceToHost, h->s_dn));
            CU(cudaEventRecord(h->ev[EV_DN], h->s_dn));
        }
    } else if (rhokap) {
        if (root) { if (int rc = enqueue_upload(h, rhokap)) return rc; }
        if (rooted) NC(nccl_api()->Broadcast(h->d_rhokap, h->d_rhokap, h->n_rhokap, ncclDouble, 0, h->comm, h->stream));
        h->resident_behind = false;
    } else {
        if (int rc = sync_resident(h)) return rc;
    }

    tr.mark("pre");
    const int rc_mc = enqueue_mc(h, nphotons, seed, -1);
    tr.mark("mc_enqueued");
    if (box_up && h->colws.last_kz_lo > 0) h->io_form |= 4;     // the gather stopped at the depth the last call's packets reached
    h->colws.gather_src = nullptr;
    h->colws.box_rk = nullptr;
    h->colws.ev_gather0 = h->colws.ev_gather1 = nullptr;
    h->colws.share_gather = 0;
    h->colws.share_comm = nullptr;
    if (rc_mc) {
        cudaStreamSynchronize(h->stream);
        if (box_dn) cudaStreamSynchronize(h->s_dn);
        if (box_up) cudaStreamSynchronize(h->s_up);
        return rc_mc;
    }
    if (box_up && root) {
        // the full grid, for every later reader of the resident rhokap: behind the column upload so the two do not share
        // the link, beside the transport (which no longer reads the resident grid in this call)
        h->timed_h2d = true;
        if (!(h->io_early & 2)) {
            CU(cudaStreamWaitEvent(h->s_dn, h->ev[EV_H1], 0));
            CU(cudaMemcpyAsync(jmean_global, h->d_zero, h->n_jmean * sizeof(double), cudaMemcpyDeviceToHost, h->s_dn));
            CU(cudaEventRecord(h->ev[EV_DN], h->s_dn));
        }
        if (!(h->io_early & 1)) {
            CU(cudaStreamWaitEvent(h->s_up, h->ev[EV_H1], 0));
            CU(cudaMemcpyAsync(h->d_rhokap, rhokap, h->n_rhokap * sizeof(double), cudaMemcpyHostToDevice, h->s_up));
            CU(cudaEventRecord(h->ev[EV_UP], h->s_up));
        }
    }

    if (box_up && rooted) h->resident_behind = true;      // the other ranks' resident grid: brought up to date when next read
    if (!root) {
        CU(cudaStreamSynchronize(h->stream));
        h->timed_d2h = false;
        tr.mark("synced");
    } else
    if (box_dn) {
        CU(cudaStreamWaitEvent(h->stream, h->ev[EV_DN], 0));          // the zero fill lands first
        CU(cudaEventRecord(h->ev[EV_D0], h->stream));
        // Long call on a grid whose deposits reach deep: one pitched 3-D DMA (52 GB/s; its per-slice descriptors are
        // submitted while the transport runs).  Otherwise posted writes from a kernel (41 GB/s, nothing to submit per
        // slice) that skips the all-zero rows of the box -- the zero fill above already covers them; in the shipped regime
        // the tally dies out e-fold per mean free path, so most planes of the box hold nothing (homog200, 1e8 packets:
        // ~40 of 200 planes).  "Deep" is judged from the previous stub-regime call's voxel-steps per packet.
        const bool shallow = h->cfg.steps_hint > 0. && 16. * h->cfg.steps_hint < (double)h->nzg;
        const bool dma = nphotons >= ((int64_t)1 << 22) && !shallow;
        if (dma) {
            cudaMemcpy3DParms p{};
            p.srcPtr = make_cudaPitchedPtr(h->d_jmean, (size_t)h->nxg * sizeof(double), (size_t)h->nxg, (size_t)h->nyg);
            p.srcPos = make_cudaPos((size_t)(cg.i0 - 1) * sizeof(double), (size_t)(cg.j0 - 1), 0);
            p.dstPtr = make_cudaPitchedPtr(jmean_global, (size_t)h->nxg * sizeof(double), (size_t)h->nxg, (size_t)h->nyg);
            p.dstPos = p.srcPos;
            p.extent = make_cudaExtent((size_t)cg.tw * sizeof(double), (size_t)cg.th, (size_t)h->nzg);
            p.kind = cudaMemcpyDeviceToHost;
            CU(cudaMemcpy3DAsync(&p, h->stream));
        } else
            CU(launch_box_mirror(g, cg, jm_dev, h->num_sms, h->stream));
        CU(cudaEventRecord(h->ev[EV_D1], h->stream));
        if (box_up) CU(cudaStreamWaitEvent(h->stream, h->ev[EV_UP], 0));
        tr.mark("dn_enqueued");
        CU(cudaStreamSynchronize(h->stream));
        tr.mark("synced");
        if (tr.on && box_up) {
            float a = 0.f, b = 0.f;
            cudaEventSynchronize(h->ev[EV_UP]); cudaEventSynchronize(h->ev[EV_DN]);
            cudaEventElapsedTime(&a, h->ev[EV_FORK], h->ev[EV_UP]); cudaEventElapsedTime(&b, h->ev[EV_FORK], h->ev[EV_DN]);
            tr.len += snprintf(tr.buf + tr.len, sizeof(tr.buf) - (size_t)tr.len, " up_done_at=%.0f dn_done_at=%.0f", a * 1e3, b * 1e3);
            cudaEventElapsedTime(&a, h->ev[EV_FORK], h->ev[EV_K0]); cudaEventElapsedTime(&b, h->ev[EV_FORK], h->ev[EV_K1]);
            tr.len += snprintf(tr.buf + tr.len, sizeof(tr.buf) - (size_t)tr.len, " k0=%.0f k1=%.0f", a * 1e3, b * 1e3);
            cudaEventElapsedTime(&a, h->ev[EV_FORK], h->ev[EV_AR1]); cudaEventElapsedTime(&b, h->ev[EV_FORK], h->ev[EV_D1]);
            tr.len += snprintf(tr.buf + tr.len, sizeof(tr.buf) - (size_t)tr.len, " ar1=%.0f d1=%.0f", a * 1e3, b * 1e3);
        }
        h->timed_d2h = true;
        if (!dma) h->last_launches += 1;
    } else {
        if (int rc = tamc_get_jmean(h, jmean_global)) return rc;
    }
    tr.mark("end");
    tr.flush(h->rank);
    if (stats) return tamc_get_stats(h, stats);
    tamc_stats tmp;
    return tamc_get_stats(h, &tmp);   // surfaces transport errors even when the caller wants no stats
}

extern "C" int tamc_run(tamc_handle h, int64_t nphotons, int64_t seed, double *jmean_global, tamc_stats *stats)
{
    if (int rc = check(h)) return rc;
    if (!jmean_global && !(h->root_io && h->comm && h->rank > 0)) return fail(TAMC_EINVAL, "tamc_run: jmean_global is null");
    if (!h->optics_set) return fail(TAMC_ESTATE, "tamc_run: tamc_set_optics has not been called");
    if (nphotons < 0 || nphotons > ((int64_t)1 << 46)) return fail(TAMC_EINVAL, "tamc_run: nphotons out of range");
    return run_boundary(h, nullptr, nphotons, seed, jmean_global, stats);
}

extern "C" int tamc_run_optics(tamc_handle h, const double *rhokap, double albedo, double hgg, double n1, double n2, int flags,
                               int64_t nphotons, int64_t seed, double *jmean_global, tamc_stats *stats)
{
    if (int rc = check(h)) return rc;
    if (!jmean_global && !(h->root_io && h->comm && h->rank > 0)) return fail(TAMC_EINVAL, "tamc_run_optics: jmean_global is null");
    if (int rc = check_optics(h, rhokap, albedo, hgg, n1, n2, flags, "tamc_run_optics")) return rc;
    if (nphotons < 0 || nphotons > ((int64_t)1 << 46)) return fail(TAMC_EINVAL, "tamc_run_optics: nphotons out of range");
    h->albedo = albedo; h->hgg = hgg; h->n1 = n1; h->n2 = n2; h->flags = flags;
    h->optics_set = true;
    return run_boundary(h, rhokap, nphotons, seed, jmean_global, stats);
}

// ------------------------------------------------------------------------------------------------
// validation paths
// ------------------------------------------------------------------------------------------------
extern "C" int tamc_run_replay(tamc_handle h, int64_t npackets, const int64_t *draw_offsets, const double *draws,
                               tamc_packet_record *records, double *jmean)
{
    if (int rc = check(h)) return rc;
    if (!h->optics_set) return fail(TAMC_ESTATE, "tamc_run_replay: tamc_set_optics has not been called");
    if (h->flags & TAMC_FRESNEL) return fail(TAMC_EINVAL, "tamc_run_replay: the reference has no boundary optics to replay (TAMC_FRESNEL set)");
    if (npackets < 0 || (npackets > 0 && (!draw_offsets || !draws))) return fail(TAMC_EINVAL, "tamc_run_replay: bad arguments");
    const DevGrid g = make_grid(h);
    CU(cudaMemsetAsync(h->d_jmean, 0, h->n_jmean * sizeof(double), h->stream));
    CU(cudaMemsetAsync(h->d_cnt, 0, CNT_N * sizeof(unsigned long long), h->stream));
    CU(cudaEventRecord(h->ev[EV_ZERO0], h->stream));
    CU(cudaEventRecord(h->ev[EV_K0], h->stream));

    // chunks bounded in packets and in draws so the staging buffers stay modest
    const int64_t max_pk = 1 << 22, max_dr = (int64_t)1 << 27;
    long long *d_off = nullptr;
    double *d_draws = nullptr;
    tamc_packet_record *d_rec = nullptr;
    int64_t cap_dr = 0, cap_pk = 0;
    std::vector<long long> rebased;
    int rc = TAMC_OK;
    int launches = 0;
    for (int64_t p0 = 0; p0 < npackets && rc == TAMC_OK;) {
        int64_t p1 = p0;
        while (p1 < npackets && p1 - p0 < max_pk && (p1 == p0 || draw_offsets[p1 + 1] - draw_offsets[p0] <= max_dr)) ++p1;
        const int64_t npk = p1 - p0, ndr = draw_offsets[p1] - draw_offsets[p0];
        if (ndr < 0) { rc = fail(TAMC_EINVAL, "tamc_run_replay: draw_offsets must be non-decreasing"); break; }
        auto cu = [&](cudaError_t e, const char *what) {
            if (e != cudaSuccess && rc == TAMC_OK) rc = fail(TAMC_ECUDA, std::string(what) + ": " + cudaGetErrorString(e));
        };
        if (npk > cap_pk) {
            cudaFree(d_off); cudaFree(d_rec);
            cap_pk = npk;
            cu(cudaMalloc(&d_off, (size_t)(cap_pk + 1) * sizeof(long long)), "cudaMalloc offsets");
            cu(cudaMalloc(&d_rec, (size_t)cap_pk * sizeof(tamc_packet_record)), "cudaMalloc records");
        }
        if (ndr > cap_dr) {
            cudaFree(d_draws);
            cap_dr = ndr;
            cu(cudaMalloc(&d_draws, (size_t)(cap_dr > 0 ? cap_dr : 1) * sizeof(double)), "cudaMalloc draws");
        }
        if (rc != TAMC_OK) break;
        rebased.resize((size_t)npk + 1);
        for (int64_t i = 0; i <= npk; ++i) rebased[(size_t)i] = draw_offsets[p0 + i] - draw_offsets[p0];
        cu(cudaMemcpyAsync(d_off, rebased.data(), (size_t)(npk + 1) * sizeof(long long), cudaMemcpyHostToDevice, h->stream), "H2D offsets");
        cu(cudaMemcpyAsync(d_draws, draws + draw_offsets[p0], (size_t)ndr * sizeof(double), cudaMemcpyHostToDevice, h->stream), "H2D draws");
        cu(launch_replay(g, npk, d_off, d_draws, h->d_cnt, d_rec, h->stream), "replay launch");
        ++launches;
        if (records)
            cu(cudaMemcpyAsync(records + p0, d_rec, (size_t)npk * sizeof(tamc_packet_record), cudaMemcpyDeviceToHost, h->stream), "D2H records");
        cu(cudaStreamSynchronize(h->stream), "replay sync");
        p0 = p1;
    }
    cudaFree(d_off); cudaFree(d_draws); cudaFree(d_rec);
    if (rc != TAMC_OK) return rc;
    CU(cudaEventRecord(h->ev[EV_K1], h->stream));
    CU(cudaEventRecord(h->ev[EV_AR1], h->stream));
    h->timed_reduce = false; h->timed_d2h = false; h->ran = true; h->last_launches = launches;
    if (jmean) CU(cudaMemcpy(jmean, h->d_jmean, h->n_jmean * sizeof(double), cudaMemcpyDeviceToHost));
    unsigned long long cnt[CNT_N];
    CU(cudaMemcpy(cnt, h->d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
    if (cnt[CNT_OVERFLOW]) return fail(TAMC_EREPLAY, std::to_string(cnt[CNT_OVERFLOW]) + " packet(s) ran out of replay draws");
    return TAMC_OK;
}

extern "C" int tamc_run_records(tamc_handle h, int64_t nphotons, int64_t seed, int64_t first_packet_id,
                                tamc_packet_record *records, double *jmean)
{
    if (int rc = check(h)) return rc;
    if (!h->optics_set) return fail(TAMC_ESTATE, "tamc_run_records: tamc_set_optics has not been called");
    if (nphotons < 0 || first_packet_id < 0 || !records) return fail(TAMC_EINVAL, "tamc_run_records: bad arguments");
    const DevGrid g = make_grid(h);
    CU(cudaMemsetAsync(h->d_jmean, 0, h->n_jmean * sizeof(double), h->stream));
    CU(cudaMemsetAsync(h->d_cnt, 0, CNT_N * sizeof(unsigned long long), h->stream));
    CU(cudaEventRecord(h->ev[EV_ZERO0], h->stream));
    CU(cudaEventRecord(h->ev[EV_K0], h->stream));
    const int64_t chunk = 1 << 22;
    tamc_packet_record *d_rec = nullptr;
    CU(cudaMalloc(&d_rec, (size_t)(nphotons < chunk ? (nphotons > 0 ? nphotons : 1) : chunk) * sizeof(tamc_packet_record)));
    int launches = 0;
    int rc = TAMC_OK;
    for (int64_t p0 = 0; p0 < nphotons && rc == TAMC_OK; p0 += chunk) {
        const int64_t npk = nphotons - p0 < chunk ? nphotons - p0 : chunk;
        cudaError_t e = launch_transport(g, h->cfg, npk, (uint64_t)seed, (uint64_t)(first_packet_id + p0), h->d_cnt, d_rec, h->stream, &launches);
        if (e == cudaSuccess) e = cudaMemcpyAsync(records + p0, d_rec, (size_t)npk * sizeof(tamc_packet_record), cudaMemcpyDeviceToHost, h->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
        if (e != cudaSuccess) rc = fail(TAMC_ECUDA, std::string("tamc_run_records: ") + cudaGetErrorString(e));
    }
    cudaFree(d_rec);
    if (rc != TAMC_OK) return rc;
    CU(cudaEventRecord(h->ev[EV_K1], h->stream));
    CU(cudaEventRecord(h->ev[EV_AR1], h->stream));
    h->timed_reduce = false; h->timed_d2h = false; h->ran = true; h->last_launches = launches;
    if (jmean) CU(cudaMemcpy(jmean, h->d_jmean, h->n_jmean * sizeof(double), cudaMemcpyDeviceToHost));
    return TAMC_OK;
}

// ------------------------------------------------------------------------------------------------
// residency, tuning, measurement
// ------------------------------------------------------------------------------------------------
extern "C" void *tamc_stream(tamc_handle h) { return h ? (void *)h->stream : nullptr; }
extern "C" double *tamc_jmean_device(tamc_handle h) { return h ? h->d_jmean : nullptr; }
extern "C" double *tamc_rhokap_device(tamc_handle h) { return h ? h->d_rhokap : nullptr; }

extern "C" int tamc_pin_host(void *ptr, uint64_t bytes)
{
    if (!ptr || !bytes) return fail(TAMC_EINVAL, "tamc_pin_host: bad argument");
    CU(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
    return TAMC_OK;
}

extern "C" int tamc_unpin_host(void *ptr)
{
    if (!ptr) return fail(TAMC_EINVAL, "tamc_unpin_host: null pointer");
    CU(cudaHostUnregister(ptr));
    return TAMC_OK;
}

static int *option_slot(tamc_handle h, const char *name)
{
    if (!h || !name) return nullptr;
    if (!strcmp(name, "variant")) return &h->cfg.variant;
    if (!strcmp(name, "block")) return &h->cfg.block;
    if (!strcmp(name, "ctas_per_sm")) return &h->cfg.ctas_per_sm;
    if (!strcmp(name, "chunk")) return &h->cfg.chunk;
    if (!strcmp(name, "scatter_min")) return &h->cfg.scatter_min;
    if (!strcmp(name, "merge")) return &h->cfg.merge;
    if (!strcmp(name, "min_ctas")) return &h->cfg.min_ctas;
    if (!strcmp(name, "tile")) return &h->cfg.tile;
    if (!strcmp(name, "column")) return &h->cfg.column;
    if (!strcmp(name, "column_tile")) return &h->cfg.column_tile;
    if (!strcmp(name, "column_park")) return &h->cfg.column_park;
    if (!strcmp(name, "flight")) return &h->cfg.flight;
    if (!strcmp(name, "walk_min")) return &h->cfg.walk_min;
    if (!strcmp(name, "flight_regs")) return &h->cfg.flight_regs;
    if (!strcmp(name, "flight_inter")) return &h->cfg.flight_inter;
    if (!strcmp(name, "flight_agg")) return &h->cfg.flight_agg;
    if (!strcmp(name, "flight_launch_min")) return &h->cfg.flight_launch_min;
    if (!strcmp(name, "launch32")) return &h->launch32;
    if (!strcmp(name, "io_early")) return &h->io_early;
    if (!strcmp(name, "gather_depth")) return &h->cfg.gather_depth;
    if (!strcmp(name, "depth_hint")) return &h->cfg.depth_hint;
    if (!strcmp(name, "reduce")) return &h->reduce;
    if (!strcmp(name, "probe_form")) return &h->probe_form;
    if (!strcmp(name, "box_reduce")) return &h->box_reduce;
    if (!strcmp(name, "reduce_bound")) return &h->reduce_bound;
    if (!strcmp(name, "reduce_planes")) return &h->reduce_planes;
    if (!strcmp(name, "form")) return &h->form;
    if (!strcmp(name, "box_io")) return &h->box_io;
    if (!strcmp(name, "root_io")) return &h->root_io;
    if (!strcmp(name, "peer_reduce")) return &h->peer_reduce;
    if (!strcmp(name, "peer_state")) return &h->peer_state;
    if (!strcmp(name, "io_form")) return &h->io_form;
    return nullptr;
}

extern "C" int tamc_set_option(tamc_handle h, const char *name, int64_t value)
{
    int *slot = option_slot(h, name);
    if (!slot) return fail(TAMC_EINVAL, std::string("tamc_set_option: unknown option ") + (name ? name : "(null)"));
    if (slot == &h->form) return fail(TAMC_EINVAL, "form is read-only: the kernel the last MC call ran");
    if (slot == &h->io_form) return fail(TAMC_EINVAL, "io_form is read-only: how the last tamc_run moved its arrays");
    if (slot == &h->reduce_planes) return fail(TAMC_EINVAL, "reduce_planes is read-only: planes of the box the last all-reduce moved");
    if (slot == &h->cfg.depth_hint) return fail(TAMC_EINVAL, "depth_hint is read-only: planes from the top face to the deepest stop of the last column-form call");
    if (slot == &h->cfg.gather_depth && (value < -1 || value > 4096)) return fail(TAMC_EINVAL, "gather_depth must be -1 (auto), 0 (all planes) or a number of planes");
    if (slot == &h->probe_form && (value < -1 || value > 1)) return fail(TAMC_EINVAL, "probe_form must be -1, 0 or 1");
    if (slot == &h->cfg.block && (value < 0 || value > 256 || value % 32)) return fail(TAMC_EINVAL, "block must be 0 (auto) or a multiple of 32 up to 256");
    if (slot == &h->cfg.variant && (value < 0 || value > 3)) return fail(TAMC_EINVAL, "variant must be 0..3");
    if (slot == &h->root_io && (value < 0 || value > 1)) return fail(TAMC_EINVAL, "root_io must be 0 or 1");
    if (slot == &h->peer_state) return fail(TAMC_EINVAL, "peer_state is read-only: 1 = the box all-reduce runs out of peer memory, -1 = not available here (NCCL)");
    if (slot == &h->peer_reduce && (value < 0 || value > 1)) return fail(TAMC_EINVAL, "peer_reduce must be 0 or 1");
    if (slot == &h->peer_reduce && h->comm) return fail(TAMC_ESTATE, "peer_reduce shapes the collectives of every call: set it on every rank before tamc_comm_init");
    if (slot == &h->root_io && h->comm) return fail(TAMC_ESTATE, "root_io shapes the collectives of every call: set it on every rank before tamc_comm_init");
    if (slot == &h->cfg.flight && (value < -1 || value > 1)) return fail(TAMC_EINVAL, "flight must be -1 (auto), 0 or 1");
    if (slot == &h->cfg.walk_min && (value < 1 || value > 32)) return fail(TAMC_EINVAL, "walk_min must be in [1,32]");
    if (slot == &h->cfg.flight_regs && value != 0 && (value < 2 || value > 4)) return fail(TAMC_EINVAL, "flight_regs must be 0, 2, 3 or 4");
    if (slot == &h->cfg.scatter_min && (value < 1 || value > 32)) return fail(TAMC_EINVAL, "scatter_min must be in [1,32]");
    if (slot == &h->cfg.chunk && (value < 0 || value > 65536 || value % 32)) return fail(TAMC_EINVAL, "chunk must be 0 (auto) or a multiple of 32 up to 65536");
    if (slot == &h->cfg.min_ctas && (value < 2 || value > 3)) return fail(TAMC_EINVAL, "min_ctas must be 2 or 3");
    if (slot == &h->cfg.column && (value < -1 || value > 2)) return fail(TAMC_EINVAL, "column must be -1 (auto), 0 (off), 1 or 2");
    if (slot == &h->cfg.column_tile && (value < -1 || value > 99)) return fail(TAMC_EINVAL, "column_tile must be -1 (auto), 0 (off) or 10*ta + tb");
    if (slot == &h->cfg.ctas_per_sm && (value < 0 || value > 32)) return fail(TAMC_EINVAL, "ctas_per_sm must be in [0,32]");
    *slot = (int)value;
    return TAMC_OK;
}

extern "C" int64_t tamc_get_option(tamc_handle h, const char *name)
{
    int *slot = option_slot(h, name);
    return slot ? *slot : -1;
}

extern "C" int tamc_selfcheck_launch(tamc_handle h, int64_t n, int64_t seed, int64_t *fallbacks, int64_t *mismatches)
{
    if (int rc = check(h)) return rc;
    if (n < 0 || !fallbacks || !mismatches) return fail(TAMC_EINVAL, "tamc_selfcheck_launch: bad arguments");
    const DevGrid g = make_grid(h);
    unsigned long long *d_out = nullptr, out[2] = {0ull, 0ull};
    CU(cudaMalloc(&d_out, sizeof(out)));
    cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(out), h->stream);
    if (e == cudaSuccess && n > 0) e = launch_selfcheck(g, n, (uint64_t)seed, 0, d_out, h->num_sms, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d_out, sizeof(out), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(TAMC_ECUDA, std::string("tamc_selfcheck_launch: ") + cudaGetErrorString(e));
    *fallbacks = (int64_t)out[0];
    *mismatches = (int64_t)out[1];
    return TAMC_OK;
}

extern "C" int tamc_roofline_probe(tamc_handle h, int64_t nphotons, int64_t seed, double *ms, int64_t *steps)
{
    if (int rc = check(h)) return rc;
    if (!h->optics_set) return fail(TAMC_ESTATE, "tamc_roofline_probe: tamc_set_optics has not been called");
    const DevGrid g = make_grid(h);
    {   // the probe draws its step counts from the opacity at the centre of the top face: it has to be positive
        double rk0 = 0.;
        const size_t at = (size_t)(h->nxg / 2) + (size_t)g.sx * (size_t)(h->nyg / 2) + (size_t)g.sxy * (size_t)h->nzg;
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaMemcpy(&rk0, h->d_rhokap + at, sizeof(double), cudaMemcpyDeviceToHost));
        if (!(rk0 > 0.) || !(rk0 < 1e300))
            return fail(TAMC_EINVAL, "tamc_roofline_probe: the resident grid has no positive opacity at the centre of the top face "
                                     "(the probe's geometric step count is drawn from it)");
    }
    CU(cudaMemsetAsync(h->d_jmean, 0, h->n_jmean * sizeof(double), h->stream));
    CU(cudaMemsetAsync(h->d_cnt, 0, CNT_N * sizeof(unsigned long long), h->stream));
    CU(cudaEventRecord(h->ev[EV_K0], h->stream));
    CU(launch_probe(g, h->cfg, nphotons, (uint64_t)seed, h->d_cnt, h->stream, &h->colws, h->probe_form));
    CU(cudaEventRecord(h->ev[EV_K1], h->stream));
    CU(cudaStreamSynchronize(h->stream));
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, h->ev[EV_K0], h->ev[EV_K1]));
    unsigned long long cnt[CNT_N];
    CU(cudaMemcpy(cnt, h->d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
    if (ms) *ms = t;
    if (steps) *steps = (int64_t)cnt[CNT_STEPS];
    h->ran = false;
    return TAMC_OK;
}

extern "C" int tamc_trace_probe(tamc_handle h, int64_t npackets, int64_t seed, double *ms, int64_t *steps, int64_t *reds)
{
    if (int rc = check(h)) return rc;
    if (!h->optics_set) return fail(TAMC_ESTATE, "tamc_trace_probe: tamc_set_optics has not been called");
    if (npackets < 1 || npackets > ((int64_t)1 << 26)) return fail(TAMC_EINVAL, "tamc_trace_probe: npackets must be in [1, 2^26]");
    if ((h->flags & (TAMC_FRESNEL | TAMC_PERIODIC)) || h->gauss_sigma > 0.)
        return fail(TAMC_EINVAL, "tamc_trace_probe: records the disk source without boundary options only");
    const DevGrid g = make_grid(h);
    int *d_counts = nullptr, *d_trace = nullptr;
    long long *d_off = nullptr;
    int rc = TAMC_OK;
    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == TAMC_OK) rc = fail(TAMC_ECUDA, std::string("tamc_trace_probe: ") + what + ": " + cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    std::vector<int> counts((size_t)npackets);
    std::vector<long long> off((size_t)npackets + 1);
    const size_t nvox = h->n_jmean;
    // ids far behind anything a run uses, so the probe never replays a production stream's packets twice in a bench
    const uint64_t first = (uint64_t)1 << 50;
    do {
        if (!cu(cudaMalloc(&d_counts, (size_t)npackets * sizeof(int)), "cudaMalloc counts")) break;
        if (!cu(cudaMalloc(&d_off, ((size_t)npackets + 1) * sizeof(long long)), "cudaMalloc offsets")) break;
        if (!cu(launch_trace(g, npackets, (uint64_t)seed, first, nullptr, nullptr, d_counts, h->num_sms, h->stream), "count pass")) break;
        if (!cu(cudaMemcpyAsync(counts.data(), d_counts, (size_t)npackets * sizeof(int), cudaMemcpyDeviceToHost, h->stream), "D2H counts")) break;
        if (!cu(cudaStreamSynchronize(h->stream), "sync")) break;
        long long tot = 0;
        for (int64_t i = 0; i < npackets; ++i) { off[(size_t)i] = tot; tot += (counts[(size_t)i] + 3) & ~3; }
        off[(size_t)npackets] = tot;
        if (!cu(cudaMalloc(&d_trace, (size_t)(tot > 0 ? tot : 4) * sizeof(int)), "cudaMalloc trace")) break;
        if (!cu(cudaMemcpyAsync(d_off, off.data(), off.size() * sizeof(long long), cudaMemcpyHostToDevice, h->stream), "H2D offsets")) break;
        if (!cu(launch_trace(g, npackets, (uint64_t)seed, first, d_off, d_trace, d_counts, h->num_sms, h->stream), "record pass")) break;
        if (h->colws.vox_elems < nvox) {
            cudaFree(h->colws.vox);
            h->colws.vox = nullptr;
            h->colws.vox_elems = 0;
            if (!cu(cudaMalloc(&h->colws.vox, nvox * sizeof(double2)), "cudaMalloc vox")) break;
            h->colws.vox_elems = nvox;
        }
        float best = 0.f;
        unsigned long long cnt[CNT_N] = {};
        for (int rep = 0; rep < 3; ++rep) {           // first pass warms the caches the way consecutive MC calls do
            cu(launch_probe_trace(g, h->cfg, h->colws.vox, npackets, d_off, d_trace, h->d_cnt, h->num_sms, h->stream, true), "pack");
            cu(cudaMemsetAsync(h->d_cnt, 0, CNT_N * sizeof(unsigned long long), h->stream), "clear counters");
            cu(cudaEventRecord(h->ev[EV_K0], h->stream), "event");
            cu(launch_probe_trace(g, h->cfg, h->colws.vox, npackets, d_off, d_trace, h->d_cnt, h->num_sms, h->stream, false), "replay");
            cu(cudaEventRecord(h->ev[EV_K1], h->stream), "event");
            cu(cudaStreamSynchronize(h->stream), "sync");
            if (rc != TAMC_OK) break;
            float t = 0.f;
            cu(cudaEventElapsedTime(&t, h->ev[EV_K0], h->ev[EV_K1]), "elapsed");
            if (rep == 0 || t < best) best = t;
            cu(cudaMemcpy(cnt, h->d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost), "D2H counters");
        }
        if (ms) *ms = best;
        if (steps) *steps = (int64_t)cnt[CNT_STEPS];
        if (reds) *reds = (int64_t)cnt[CNT_SCATTERS];
    } while (false);
    cudaFree(d_counts); cudaFree(d_off); cudaFree(d_trace);
    h->ran = false;
    return rc;
}

extern "C" int tamc_flush_l2(tamc_handle h, uint64_t bytes)
{
    if (int rc = check(h)) return rc;
    const size_t n = (size_t)(bytes + 7) / 8;
    if (n > h->flush_elems) {
        cudaFree(h->d_flush);
        h->d_flush = nullptr;
        h->flush_elems = 0;
        CU(cudaMalloc(&h->d_flush, n * sizeof(double)));
        h->flush_elems = n;
    }
    CU(launch_fill(h->d_flush, n, 0., h->num_sms, h->stream));
    return TAMC_OK;
}
