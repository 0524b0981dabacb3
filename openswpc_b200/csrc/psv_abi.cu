// psv_abi.cu -- C ABI of include/swpcpsv_b200.h: device state of one swpc_psv rank, kernel launches, halo exchange.
// All reference citations are file:line under /root/reference/src/swpc_psv.
#include "../../include/swpcpsv_b200.h"
#include "psv_kernels.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

using namespace swpc;

static thread_local std::string g_perr;
extern "C" const char *swpcpsv_last_error(void) { return g_perr.c_str(); }
extern "C" const char *swpcpsv_version(void) { return "swpcpsv_b200 0.1 (reference: OpenSWPC 25.05.2 swpc_psv)"; }
static int fail(const std::string &m) { g_perr = m; return 1; }

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            char b_[512];                                                                                \
            snprintf(b_, sizeof(b_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return fail(b_);                                                                             \
        }                                                                                                \
    } while (0)

// NCCL, loaded lazily (single-GPU use has no NCCL dependency)
struct PsvNccl {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static PsvNccl g_nc;
static int nccl_load() {
    if (g_nc.lib) return 0;
    for (const char *n : {"libnccl.so.2", "libnccl.so"}) {
        g_nc.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nc.lib) break;
    }
    if (!g_nc.lib) return fail(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define SYM(f, s)                                                    \
    *(void **)(&g_nc.f) = dlsym(g_nc.lib, s);                        \
    if (!g_nc.f) return fail(std::string("libnccl: missing symbol ") + s);
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv"); SYM(AllReduce, "ncclAllReduce");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return 0;
}
#define NK(call)                                                                                                  \
    do {                                                                                                          \
        ncclResult_t r_ = (call);                                                                                 \
        if (r_ != ncclSuccess) {                                                                                  \
            char b_[512];                                                                                         \
            snprintf(b_, sizeof(b_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nc.GetErrorString(r_));       \
            return fail(b_);                                                                                      \
        }                                                                                                         \
    } while (0)

struct swpcpsv_handle {
    swpcpsv_grid g{};
    int dev = 0, nxp = 0, NZP = 0, NXM = 0, nzm_h = 0, fb = 8, nm = 0;
    long long ncell = 0;
    void *Fall = nullptr;                 // Vx Vz Sxx Szz Sxz, contiguous
    void *F[5] = {};
    float *R = nullptr;                   // 3*nm arrays
    float *Mall = nullptr, *med[5] = {};  // rho lam mu taup taus
    int4 *band = nullptr;
    int *kbeg_a = nullptr, *kob = nullptr;
    std::vector<int> h_kbeg_a;
    long long *aoff = nullptr, naux = 0;
    float *aux = nullptr;
    float4 *g4[4] = {};                   // gxc gxe gzc gze
    float *cg[4] = {};                    // gx_c gx_b gz_c gz_b
    bool absorber_ready = false, medium_ready = false;
    double r40[4][2] = {};                // x40 x41 z40 z41
    float r20x = 0.f, r20z = 0.f;
    float c1[MAXNM] = {}, c2[MAXNM] = {}, d1[MAXNM] = {}, d2 = 0.f;
    double dt_dxz = 0, w40x = 0, w40z = 0, w41x = 0, w41z = 0;
    long long cells_interior = 0, cells_absorber = 0;
    // sources
    int nsrc = 0, stf = 3, bf_mode = 0;
    float tbeg = 0.f;
    int *src_ik = nullptr;
    double *src_mo = nullptr, *src_m3 = nullptr;
    float *src_prm = nullptr;
    // stations
    int nst = 0, ntdec_w = 0, ntw = 0, sw[4] = {1, 0, 0, 0};
    int *st_ik = nullptr;
    float *wav[4] = {}, *wav_acc = nullptr;
    float M0 = 1.f, UC = 1e-12f;
    bool pw_mode = false;   // plane-wave mode: edge extrapolation ahead of the PML updates
    // snapshots (m_snap.f90)
    swpcpsv_snap_cfg snap{};
    bool snap_on = false;
    float *snap_buf[3] = {nullptr, nullptr, nullptr};   // ps, v, u: (2, nzs, nxs) over the whole snapshot grid
    float *snap_tmp = nullptr;
    size_t snap_tmp_n = 0;
    unsigned int *vmax_d = nullptr;
    // halo: 0 = +x (ip), 1 = -x (im)
    void *sbuf[2] = {}, *rbuf[2] = {};
    int nbr[2] = {-1, -1};
    ncclComm_t comm = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int tk = 256, ilen = 64, pf = 1;
    long long launches = 0;
    bool ktiming = false;
    std::vector<cudaEvent_t> kev[2][2];
    size_t kev_used[2] = {0, 0};
};

template <typename F>
static PsvParams<F> make_params(const swpcpsv_handle *h) {
    PsvParams<F> p{};
    const swpcpsv_grid &g = h->g;
    p.nz = g.nz; p.nxp = h->nxp; p.NZP = h->NZP; p.NXM = h->NXM; p.ncell = h->ncell;
    p.li0_k = g.ibeg_k - g.ibeg; p.li1_k = g.iend_k - g.ibeg; p.k1_k = g.kend_k;
    p.abc = g.abc_type;
    p.Vx = (F *)h->F[0]; p.Vz = (F *)h->F[1]; p.Sxx = (F *)h->F[2]; p.Szz = (F *)h->F[3]; p.Sxz = (F *)h->F[4];
    p.R = h->R;
    p.rho = h->med[0]; p.lam = h->med[1]; p.mu = h->med[2]; p.taup = h->med[3]; p.taus = h->med[4];
    p.band = h->band; p.kbeg_a = h->kbeg_a; p.kob = h->kob;
    p.aoff = h->aoff; p.aux = h->aux; p.naux = h->naux;
    p.gxc = h->g4[0]; p.gxe = h->g4[1]; p.gzc = h->g4[2]; p.gze = h->g4[3];
    p.cgx_c = h->cg[0]; p.cgx_b = h->cg[1]; p.cgz_c = h->cg[2]; p.cgz_b = h->cg[3];
    for (int o = 0; o < 2; o++) { p.r40x[o] = (F)h->r40[0][o]; p.r41x[o] = (F)h->r40[1][o]; p.r40z[o] = (F)h->r40[2][o]; p.r41z[o] = (F)h->r40[3][o]; }
    p.r20x = h->r20x; p.r20z = h->r20z;
    for (int m = 0; m < MAXNM; m++) { p.c1[m] = h->c1[m]; p.c2[m] = h->c2[m]; p.d1[m] = h->d1[m]; }
    p.d2 = h->d2; p.dt = g.dt;
    return p;
}

// kernel__setup m_kernel.f90:45-67, r20 of m_absorb_p.f90:68-69, dt_dxz of m_source.f90:252, r40/r41 of m_wav.f90:127-130
template <typename F>
static void setup_coefs(swpcpsv_handle *h, const float *ts) {
    const swpcpsv_grid &g = h->g;
    const F d[2] = {(F)g.dx, (F)g.dz};
    for (int a = 0; a < 2; a++) {
        const F rc40 = (F)17.0 / (F)16.0 / d[a], rc41 = (F)1.0 / (F)48.0 / d[a];
        const F rd40 = -(F)1.0 / (F)16.0 / d[a], rd41 = -(F)1.0 / (F)48.0 / d[a];
        h->r40[2 * a][0] = (double)(F)(rc40 + (-1) * rd40);
        h->r40[2 * a][1] = (double)(F)(rc40 + (1) * rd40);
        h->r40[2 * a + 1][0] = (double)(F)(rc41 + (-1) * rd41);
        h->r40[2 * a + 1][1] = (double)(F)(rc41 + (1) * rd41);
    }
    h->r20x = (float)((F)1.0f / d[0]);
    h->r20z = (float)((F)1.0f / d[1]);
    h->w40x = (double)((F)9.0 / (F)8.0 / d[0]); h->w40z = (double)((F)9.0 / (F)8.0 / d[1]);
    h->w41x = (double)((F)1.0 / (F)24.0 / d[0]); h->w41z = (double)((F)1.0 / (F)24.0 / d[1]);
    const float dt = g.dt;
    const int nm = g.nm;
    h->d2 = 0.0f;
    if (nm > 0) {
        float sum = 0.0f;
        for (int m = 0; m < nm; m++) {
            h->c1[m] = (2 * ts[m] - dt) / (2 * ts[m] + dt);
            h->c2[m] = (2) / (2 * ts[m] + dt) / nm;
            h->d1[m] = 2 * ts[m] / (2 * ts[m] - dt);
            sum += dt / (2 * ts[m] - dt);
        }
        h->d2 = sum / nm;
    }
    h->dt_dxz = (double)((F)dt / ((F)g.dx * (F)g.dz));
}

static int create_state(swpcpsv_handle *h, const swpcpsv_grid *g, const float *ts);

extern "C" int swpcpsv_create(const swpcpsv_grid *g, const float *ts, swpcpsv_handle **out) {
    if (!g || !out) return fail("swpcpsv_create: null argument");
    *out = nullptr;
    if (g->field_bytes != 8 && g->field_bytes != 4) return fail("field_bytes must be 8 (MP=DP) or 4 (MP=SP)");
    if (g->nm < 0 || g->nm > MAXNM) return fail("nm must be 0..3");
    if (g->abc_type != SWPCPSV_ABC_PML && g->abc_type != SWPCPSV_ABC_CERJAN) return fail("abc_type must be 1 (pml) or 2 (cerjan)");
    if (g->nm > 0 && !ts) return fail("ts[nm] required when nm > 0");
    if (g->iend < g->ibeg || g->nz < 1) return fail("empty subdomain");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(std::string("no CUDA device available (the swpcpsv_b200 path has no CPU fallback): ") + cudaGetErrorString(e));
    swpcpsv_handle *h = new swpcpsv_handle();
    h->g = *g;
    h->dev = g->device >= 0 ? g->device : (g->myid % ndev);
    const int rc = create_state(h, g, ts);
    if (rc) {   // e.g. out of device memory half way: give everything back, keep the message
        const std::string msg = swpcpsv_last_error();
        swpcpsv_destroy(h);
        cudaGetLastError();   // a failed cudaMalloc is not sticky, but it stays the "last error" until it is read
        return fail(msg);
    }
    *out = h;
    return 0;
}

static int create_state(swpcpsv_handle *h, const swpcpsv_grid *g, const float *ts) {
    CK(cudaSetDevice(h->dev));
    h->nxp = g->iend - g->ibeg + 1;
    h->NXM = h->nxp + 2 * HALO + g->ipad;
    h->nzm_h = g->nz + 6 + g->kpad;
    h->NZP = ((g->nz + g->kpad + KOFF + 3) + 31) / 32 * 32;
    h->ncell = (long long)h->NZP * h->NXM;
    h->fb = g->field_bytes;
    h->nm = g->nm;
    if (h->fb == 8) setup_coefs<double>(h, ts);
    else setup_coefs<float>(h, ts);
    CK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    CK(cudaMalloc(&h->Fall, (size_t)h->ncell * h->fb * 5));
    CK(cudaMemsetAsync(h->Fall, 0, (size_t)h->ncell * h->fb * 5, h->st));
    for (int a = 0; a < 5; a++) h->F[a] = (char *)h->Fall + (size_t)a * h->ncell * h->fb;
    if (h->nm > 0) {
        CK(cudaMalloc(&h->R, (size_t)h->ncell * 3 * h->nm * sizeof(float)));
        CK(cudaMemsetAsync(h->R, 0, (size_t)h->ncell * 3 * h->nm * sizeof(float), h->st));
    }
    CK(cudaMalloc(&h->Mall, (size_t)h->ncell * sizeof(float) * 5));
    CK(cudaMemsetAsync(h->Mall, 0, (size_t)h->ncell * sizeof(float) * 5, h->st));
    for (int a = 0; a < 5; a++) h->med[a] = h->Mall + (size_t)a * h->ncell;
    CK(cudaMalloc(&h->band, (size_t)h->NXM * sizeof(int4)));
    CK(cudaMalloc(&h->kbeg_a, (size_t)h->NXM * sizeof(int)));
    CK(cudaMalloc(&h->kob, (size_t)h->NXM * sizeof(int)));
    CK(cudaMemsetAsync(h->band, 0, (size_t)h->NXM * sizeof(int4), h->st));
    CK(cudaMemsetAsync(h->kob, 0, (size_t)h->NXM * sizeof(int), h->st));
    h->h_kbeg_a.resize(h->NXM);   // m_global.f90:259-266
    for (int mi = 0; mi < h->NXM; mi++) {
        const int i = g->ibeg - HALO + mi;
        h->h_kbeg_a[mi] = (i <= g->na || g->nx - g->na + 1 <= i) ? 1 : g->nz - g->na + 1;
    }
    CK(cudaMemcpyAsync(h->kbeg_a, h->h_kbeg_a.data(), (size_t)h->NXM * sizeof(int), cudaMemcpyHostToDevice, h->st));
    CK(cudaMalloc(&h->vmax_d, 2 * sizeof(unsigned int)));
    const size_t bsz = (size_t)3 * g->nz * h->fb;   // m_global.f90:214-215
    for (int f = 0; f < 2; f++) {
        CK(cudaMalloc(&h->sbuf[f], bsz)); CK(cudaMalloc(&h->rbuf[f], bsz));
        CK(cudaMemsetAsync(h->sbuf[f], 0, bsz, h->st)); CK(cudaMemsetAsync(h->rbuf[f], 0, bsz, h->st));
    }
    h->nbr[0] = (g->myid + 1 < g->nproc_x) ? g->myid + 1 : -1;   // itbl, m_global.f90:436-454
    h->nbr[1] = (g->myid - 1 >= 0) ? g->myid - 1 : -1;
    // cell census for the roofline (PML: absorber cells = k >= kbeg_a(i); interior = the kernel box)
    for (int li = 0; li < h->nxp; li++) {
        const int kb = g->abc_type == SWPCPSV_ABC_PML ? h->h_kbeg_a[li + HALO] : g->nz + 1;
        h->cells_absorber += g->nz - kb + 1;
        const int i = g->ibeg + li;
        if (i >= g->ibeg_k && i <= g->iend_k) h->cells_interior += std::min(g->kend_k, kb - 1);
    }
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int swpcpsv_destroy(swpcpsv_handle *h) {
    if (!h) return 0;
    cudaSetDevice(h->dev);
    cudaDeviceSynchronize();
    if (h->comm && g_nc.CommDestroy) g_nc.CommDestroy(h->comm);
    cudaFree(h->Fall); cudaFree(h->R); cudaFree(h->Mall); cudaFree(h->band); cudaFree(h->kbeg_a); cudaFree(h->kob);
    cudaFree(h->aoff); cudaFree(h->aux);
    for (int a = 0; a < 4; a++) { cudaFree(h->g4[a]); cudaFree(h->cg[a]); cudaFree(h->wav[a]); }
    cudaFree(h->src_ik); cudaFree(h->src_mo); cudaFree(h->src_m3); cudaFree(h->src_prm);
    cudaFree(h->st_ik); cudaFree(h->wav_acc); cudaFree(h->vmax_d);
    for (int q = 0; q < 3; q++) cudaFree(h->snap_buf[q]);
    cudaFree(h->snap_tmp);
    for (int f = 0; f < 2; f++) { cudaFree(h->sbuf[f]); cudaFree(h->rbuf[f]); }
    for (int w = 0; w < 2; w++) for (int b = 0; b < 2; b++) for (cudaEvent_t ev : h->kev[w][b]) cudaEventDestroy(ev);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return 0;
}

// host (reference layout, k from -2) <-> device (padded) copy of one 2-D array
static int copy2d(swpcpsv_handle *h, void *dev, const void *host_c, void *host_m, size_t elem, bool to_device) {
    char *d = (char *)dev + (size_t)(KOFF - 3) * elem;
    if (to_device)
        CK(cudaMemcpy2DAsync(d, (size_t)h->NZP * elem, host_c, (size_t)h->nzm_h * elem, (size_t)h->nzm_h * elem, (size_t)h->NXM,
                             cudaMemcpyHostToDevice, h->st));
    else
        CK(cudaMemcpy2DAsync(host_m, (size_t)h->nzm_h * elem, d, (size_t)h->NZP * elem, (size_t)h->nzm_h * elem, (size_t)h->NXM,
                             cudaMemcpyDeviceToHost, h->st));
    return 0;
}

extern "C" int swpcpsv_upload_medium(swpcpsv_handle *h, const float *rho, const float *lam, const float *mu, const float *taup,
                                     const float *taus, const int32_t *kfs, const int32_t *kob, const int32_t *kfs_top,
                                     const int32_t *kfs_bot, const int32_t *kob_top, const int32_t *kob_bot, const int32_t *kbeg_a) {
    if (!h) return fail("null handle");
    if (!rho || !lam || !mu || !kob || !kfs_top || !kfs_bot || !kob_top || !kob_bot) return fail("swpcpsv_upload_medium: null array");
    if (h->nm > 0 && (!taup || !taus)) return fail("swpcpsv_upload_medium: taup / taus required when nm > 0");
    (void)kfs;
    CK(cudaSetDevice(h->dev));
    const float *src[5] = {rho, lam, mu, taup, taus};
    for (int a = 0; a < 5; a++)
        if (src[a] && copy2d(h, h->med[a], src[a], nullptr, sizeof(float), true)) return 1;
    std::vector<int4> band(h->NXM);
    for (int mi = 0; mi < h->NXM; mi++) band[mi] = make_int4(kfs_top[mi], kfs_bot[mi], kob_top[mi], kob_bot[mi]);
    CK(cudaMemcpyAsync(h->band, band.data(), band.size() * sizeof(int4), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->kob, kob, (size_t)h->NXM * sizeof(int), cudaMemcpyHostToDevice, h->st));
    if (kbeg_a)
        for (int mi = 0; mi < h->NXM; mi++)
            if (kbeg_a[mi] != h->h_kbeg_a[mi]) return fail("swpcpsv_upload_medium: kbeg_a differs from m_global.f90:259-266 for this grid");
    CK(cudaStreamSynchronize(h->st));
    h->medium_ready = true;
    return 0;
}

extern "C" int swpcpsv_upload_fields(swpcpsv_handle *h, const void *Vx, const void *Vz, const void *Sxx, const void *Szz, const void *Sxz) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    const void *src[5] = {Vx, Vz, Sxx, Szz, Sxz};
    for (int a = 0; a < 5; a++)
        if (src[a] && copy2d(h, h->F[a], src[a], nullptr, (size_t)h->fb, true)) return 1;
    CK(cudaStreamSynchronize(h->st));
    return 0;
}
extern "C" int swpcpsv_download_fields(swpcpsv_handle *h, void *Vx, void *Vz, void *Sxx, void *Szz, void *Sxz) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    void *dst[5] = {Vx, Vz, Sxx, Szz, Sxz};
    for (int a = 0; a < 5; a++)
        if (dst[a] && copy2d(h, h->F[a], nullptr, dst[a], (size_t)h->fb, false)) return 1;
    CK(cudaStreamSynchronize(h->st));
    return 0;
}
extern "C" int swpcpsv_download_memvars(swpcpsv_handle *h, float *Rxx, float *Rzz, float *Rxz) {
    if (!h) return fail("null handle");
    if (h->nm == 0) return 0;
    CK(cudaSetDevice(h->dev));
    float *dst[3] = {Rxx, Rzz, Rxz};
    const size_t nh = (size_t)h->nzm_h * h->NXM;
    std::vector<float> tmp(nh);
    for (int c = 0; c < 3; c++) {
        if (!dst[c]) continue;
        for (int m = 0; m < h->nm; m++) {
            if (copy2d(h, h->R + (size_t)(c * h->nm + m) * h->ncell, nullptr, tmp.data(), sizeof(float), false)) return 1;
            CK(cudaStreamSynchronize(h->st));
            for (size_t q = 0; q < nh; q++) dst[c][q * h->nm + m] = tmp[q];
        }
    }
    return 0;
}
extern "C" int swpcpsv_zero_state(swpcpsv_handle *h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    CK(cudaMemsetAsync(h->Fall, 0, (size_t)h->ncell * h->fb * 5, h->st));
    if (h->R) CK(cudaMemsetAsync(h->R, 0, (size_t)h->ncell * 3 * h->nm * sizeof(float), h->st));
    if (h->aux) CK(cudaMemsetAsync(h->aux, 0, (size_t)h->naux * 8 * sizeof(float), h->st));
    if (h->wav_acc) CK(cudaMemsetAsync(h->wav_acc, 0, (size_t)h->nst * 5 * sizeof(float), h->st));
    return 0;
}

extern "C" int swpcpsv_setup_pml(swpcpsv_handle *h, const float *gxc, const float *gxe, const float *gzc, const float *gze) {
    if (!h) return fail("null handle");
    if (h->g.abc_type != SWPCPSV_ABC_PML) return fail("swpcpsv_setup_pml: abc_type is not pml");
    if (!gxc || !gxe || !gzc || !gze) return fail("swpcpsv_setup_pml: null profile");
    CK(cudaSetDevice(h->dev));
    const float *src[4] = {gxc, gxe, gzc, gze};
    const size_t n[4] = {(size_t)h->nxp, (size_t)h->nxp, (size_t)h->g.nz, (size_t)h->g.nz};
    for (int a = 0; a < 4; a++) {
        if (!h->g4[a]) CK(cudaMalloc(&h->g4[a], n[a] * sizeof(float4)));
        CK(cudaMemcpyAsync(h->g4[a], src[a], n[a] * sizeof(float4), cudaMemcpyHostToDevice, h->st));
    }
    // ADE arrays for absorber cells only; element k of column li sits at aoff(li) + k - kbeg_a, with aoff = (kbeg_a-1) mod 32
    // so that lane (k-1) mod 32 owns it, as in the field arrays
    std::vector<long long> aoff(h->nxp);
    long long off = 0;
    for (int li = 0; li < h->nxp; li++) {
        const int kb = h->h_kbeg_a[li + HALO];
        off = (off + 31) / 32 * 32 + (kb - 1) % 32;
        aoff[li] = off;
        off += h->g.nz - kb + 1;
    }
    h->naux = (off + 31) / 32 * 32;
    if (!h->aoff) CK(cudaMalloc(&h->aoff, (size_t)h->nxp * sizeof(long long)));
    CK(cudaMemcpyAsync(h->aoff, aoff.data(), (size_t)h->nxp * sizeof(long long), cudaMemcpyHostToDevice, h->st));
    cudaFree(h->aux);
    CK(cudaMalloc(&h->aux, (size_t)h->naux * 8 * sizeof(float)));
    CK(cudaMemsetAsync(h->aux, 0, (size_t)h->naux * 8 * sizeof(float), h->st));
    CK(cudaStreamSynchronize(h->st));
    h->absorber_ready = true;
    return 0;
}

extern "C" int swpcpsv_setup_cerjan(swpcpsv_handle *h, const float *gx_c, const float *gx_b, const float *gz_c, const float *gz_b) {
    if (!h) return fail("null handle");
    if (h->g.abc_type != SWPCPSV_ABC_CERJAN) return fail("swpcpsv_setup_cerjan: abc_type is not cerjan");
    if (!gx_c || !gx_b || !gz_c || !gz_b) return fail("swpcpsv_setup_cerjan: null vector");
    CK(cudaSetDevice(h->dev));
    const float *src[4] = {gx_c, gx_b, gz_c, gz_b};
    for (int a = 0; a < 4; a++) {
        const bool isx = a < 2;
        const size_t nd = isx ? (size_t)h->NXM : (size_t)h->NZP;
        std::vector<float> v(nd, 1.0f);
        if (isx) for (int q = 0; q < h->NXM; q++) v[q] = src[a][q];
        else for (int q = 0; q < h->nzm_h; q++) v[q + KOFF - 3] = src[a][q];   // host k = -2 + q -> device k + KOFF - 1
        if (!h->cg[a]) CK(cudaMalloc(&h->cg[a], nd * sizeof(float)));
        CK(cudaMemcpyAsync(h->cg[a], v.data(), nd * sizeof(float), cudaMemcpyHostToDevice, h->st));
        CK(cudaStreamSynchronize(h->st));
    }
    h->absorber_ready = true;
    return 0;
}

static int stf_code(const char *s) {
    if (!s) return 3;
    if (!strcmp(s, "boxcar")) return 0;
    if (!strcmp(s, "triangle")) return 1;
    if (!strcmp(s, "herrmann")) return 2;
    if (!strcmp(s, "kupper")) return 3;
    if (!strcmp(s, "cosine") || !strcmp(s, "scosine")) return 4;
    if (!strcmp(s, "texp")) return 5;
    return 3;   // momentrate's default branch (src/shared/m_fdtool.f90:494)
}

extern "C" int swpcpsv_set_sources(swpcpsv_handle *h, int32_t nsrc, const int32_t *isrc, const int32_t *ksrc, const double *mo,
                                   const double *mxx, const double *mzz, const double *mxz, const float *srcprm, const char *stftype,
                                   int32_t bf_mode, float tbeg) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    h->nsrc = nsrc; h->bf_mode = bf_mode; h->tbeg = tbeg; h->stf = stf_code(stftype);
    cudaFree(h->src_ik); cudaFree(h->src_mo); cudaFree(h->src_m3); cudaFree(h->src_prm);
    h->src_ik = nullptr; h->src_mo = nullptr; h->src_m3 = nullptr; h->src_prm = nullptr;
    if (nsrc <= 0) return 0;
    if (!isrc || !ksrc || !srcprm || !mxx || !mzz || (!bf_mode && (!mo || !mxz))) return fail("swpcpsv_set_sources: null array");
    std::vector<int> ik(2 * (size_t)nsrc);
    std::vector<double> m3(3 * (size_t)nsrc), mo_(nsrc);
    for (int i = 0; i < nsrc; i++) {
        const int mi = isrc[i] - h->g.ibeg + HALO;
        // the stencil writes (ii, ii-1) x (kk, kk-1): the sleeve rule ibeg-2 <= is <= iend+3 of m_source.f90:177-178 keeps it inside the memory box
        if (mi - 1 < 0 || mi >= h->NXM || ksrc[i] - 1 < -2 || ksrc[i] > h->g.nz + 3) return fail("swpcpsv_set_sources: source outside this rank's memory box");
        ik[2 * i] = mi; ik[2 * i + 1] = ksrc[i];
        m3[3 * i] = mxx[i]; m3[3 * i + 1] = mzz[i]; m3[3 * i + 2] = (bf_mode || !mxz) ? 0.0 : mxz[i];
        mo_[i] = (bf_mode || !mo) ? 0.0 : mo[i];
    }
    CK(cudaMalloc(&h->src_ik, ik.size() * sizeof(int))); CK(cudaMalloc(&h->src_mo, mo_.size() * sizeof(double)));
    CK(cudaMalloc(&h->src_m3, m3.size() * sizeof(double))); CK(cudaMalloc(&h->src_prm, 2 * (size_t)nsrc * sizeof(float)));
    CK(cudaMemcpyAsync(h->src_ik, ik.data(), ik.size() * sizeof(int), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->src_mo, mo_.data(), mo_.size() * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->src_m3, m3.data(), m3.size() * sizeof(double), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->src_prm, srcprm, 2 * (size_t)nsrc * sizeof(float), cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int swpcpsv_set_stations(swpcpsv_handle *h, int32_t nst, const int32_t *ist, const int32_t *kst, int32_t ntdec_w, int32_t ntw,
                                    float M0, float UC, int32_t sw_v, int32_t sw_u, int32_t sw_stress, int32_t sw_strain) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    h->nst = nst; h->ntdec_w = ntdec_w; h->ntw = ntw; h->M0 = M0; h->UC = UC;
    h->sw[0] = sw_v; h->sw[1] = sw_u; h->sw[2] = sw_stress; h->sw[3] = sw_strain;
    cudaFree(h->st_ik); cudaFree(h->wav_acc);
    h->st_ik = nullptr; h->wav_acc = nullptr;
    for (int a = 0; a < 4; a++) { cudaFree(h->wav[a]); h->wav[a] = nullptr; }
    if (nst <= 0) return 0;
    if (!ist || !kst || ntdec_w <= 0 || ntw <= 0) return fail("swpcpsv_set_stations: bad arguments");
    std::vector<int> ik(2 * (size_t)nst);
    for (int i = 0; i < nst; i++) {
        if (ist[i] < h->g.ibeg || ist[i] > h->g.iend || kst[i] < 1 || kst[i] > h->g.nz) return fail("swpcpsv_set_stations: station outside the owned box");
        ik[2 * i] = ist[i] - h->g.ibeg + HALO; ik[2 * i + 1] = kst[i];
    }
    CK(cudaMalloc(&h->st_ik, ik.size() * sizeof(int)));
    CK(cudaMemcpyAsync(h->st_ik, ik.data(), ik.size() * sizeof(int), cudaMemcpyHostToDevice, h->st));
    for (int a = 0; a < 4; a++) {
        if (!h->sw[a]) continue;
        const size_t n = (size_t)ntw * (a < 2 ? 2 : 3) * nst * sizeof(float);
        CK(cudaMalloc(&h->wav[a], n));
        CK(cudaMemsetAsync(h->wav[a], 0, n, h->st));
    }
    CK(cudaMalloc(&h->wav_acc, (size_t)nst * 5 * sizeof(float)));
    CK(cudaMemsetAsync(h->wav_acc, 0, (size_t)nst * 5 * sizeof(float), h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int swpcpsv_get_wav(swpcpsv_handle *h, int32_t which, float *out) {
    if (!h || !out) return fail("null argument");
    if (which < 0 || which > 3) return fail("swpcpsv_get_wav: which must be 0..3");
    if (h->nst <= 0 || !h->wav[which]) return 0;
    CK(cudaSetDevice(h->dev));
    CK(cudaMemcpyAsync(out, h->wav[which], (size_t)h->ntw * (which < 2 ? 2 : 3) * h->nst * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

static int ready(swpcpsv_handle *h) {
    if (!h) return fail("null handle");
    if (!h->medium_ready) return fail("swpcpsv: upload_medium has not been called");
    if (!h->absorber_ready) return fail("swpcpsv: setup_pml / setup_cerjan has not been called");
    if (cudaSetDevice(h->dev) != cudaSuccess) return fail("cudaSetDevice failed");
    return 0;
}

static int ktime(swpcpsv_handle *h, int w, int b) {
    if (!h->ktiming) return 0;
    if (b == 0 && h->kev_used[w] >= h->kev[w][0].size())
        for (int q = 0; q < 2; q++) { cudaEvent_t ev; CK(cudaEventCreate(&ev)); h->kev[w][q].push_back(ev); }
    CK(cudaEventRecord(h->kev[w][b][h->kev_used[w]], h->st));
    if (b == 1) h->kev_used[w]++;
    return 0;
}

template <typename F, int NM, bool STRESS>
static int launch_phase(swpcpsv_handle *h, const PsvParams<F> &p, int phase) {
    const int tk = std::min(h->tk, (h->g.nz + 31) / 32 * 32);
    dim3 blk((unsigned)tk), grd((unsigned)((h->g.nz + tk - 1) / tk), (unsigned)((h->nxp + h->ilen - 1) / h->ilen));
    psv_sweep<F, NM, STRESS><<<grd, blk, 0, h->st>>>(p, phase, 0, h->nxp - 1, h->ilen, h->pf, 1);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <typename F, bool STRESS>
static int launch_sweep(swpcpsv_handle *h, int phase) {
    const PsvParams<F> p = make_params<F>(h);
    switch (h->nm) {
    case 0: return launch_phase<F, 0, STRESS>(h, p, phase);
    case 1: return launch_phase<F, 1, STRESS>(h, p, phase);
    case 2: return launch_phase<F, 2, STRESS>(h, p, phase);
    default: return launch_phase<F, 3, STRESS>(h, p, phase);
    }
}

// plane-wave mode + PML: absorb_p__update_stress extrapolates the velocities, absorb_p__update_vel the stresses, into
// column 0 / nx+1 on the outer ranks ahead of the update (the interior kernel never reads those columns)
static int pw_edges(swpcpsv_handle *h, bool stress_fields) {
    if (!h->pw_mode || h->g.abc_type != SWPCPSV_ABC_PML) return 0;
    const int dstL = h->g.myid == 0 ? HALO - 1 : -1, dstR = h->g.myid == h->g.nproc_x - 1 ? HALO + h->nxp : -1;
    if (dstL < 0 && dstR < 0) return 0;
    dim3 blk(128), grd((unsigned)((h->g.nz + 127) / 128), 2);
    if (h->fb == 8) {
        const PsvParams<double> p = make_params<double>(h);
        if (stress_fields) psv_pw_edge_kernel<double><<<grd, blk, 0, h->st>>>(p.Sxx, p.Szz, p.Sxz, h->g.nz, h->NZP, dstL, dstR);
        else psv_pw_edge_kernel<double><<<grd, blk, 0, h->st>>>(p.Vx, p.Vz, (double *)nullptr, h->g.nz, h->NZP, dstL, dstR);
    } else {
        const PsvParams<float> p = make_params<float>(h);
        if (stress_fields) psv_pw_edge_kernel<float><<<grd, blk, 0, h->st>>>(p.Sxx, p.Szz, p.Sxz, h->g.nz, h->NZP, dstL, dstR);
        else psv_pw_edge_kernel<float><<<grd, blk, 0, h->st>>>(p.Vx, p.Vz, (float *)nullptr, h->g.nz, h->NZP, dstL, dstR);
    }
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int swpcpsv_update_stress(swpcpsv_handle *h) {
    if (ready(h)) return 1;
    if (ktime(h, 0, 0)) return 1;
    if (pw_edges(h, false)) return 1;
    if (h->fb == 8 ? launch_sweep<double, true>(h, PSV_FUSED) : launch_sweep<float, true>(h, PSV_FUSED)) return 1;
    return ktime(h, 0, 1);
}

template <typename F>
static int launch_source(swpcpsv_handle *h, int it, bool body) {
    if (h->nsrc <= 0) return 0;
    PsvSrc s{};
    s.nsrc = h->nsrc; s.ik = h->src_ik; s.mo = h->src_mo; s.m3 = h->src_m3; s.prm = h->src_prm; s.stf = h->stf; s.dt_dxz = h->dt_dxz;
    // m_source.f90:568 tbeg + (it-0.5)*dt ; :610 tbeg + it*dt  (default-real arithmetic)
    s.t = body ? h->tbeg + (float)it * h->g.dt : h->tbeg + ((float)it - 0.5f) * h->g.dt;
    const int nb = (h->nsrc + 127) / 128;
    if (body) psv_bodyforce_kernel<F><<<nb, 128, 0, h->st>>>(make_params<F>(h), s);
    else psv_stressglut_kernel<F><<<nb, 128, 0, h->st>>>(make_params<F>(h), s);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int swpcpsv_stressglut(swpcpsv_handle *h, int32_t it) {
    if (ready(h)) return 1;
    if (h->bf_mode) return 0;
    return h->fb == 8 ? launch_source<double>(h, it, false) : launch_source<float>(h, it, false);
}

extern "C" int swpcpsv_update_vel(swpcpsv_handle *h, int32_t it) {
    if (ready(h)) return 1;
    if (ktime(h, 1, 0)) return 1;
    if (pw_edges(h, true)) return 1;
    if (!h->bf_mode || h->nsrc <= 0) {
        if (h->fb == 8 ? launch_sweep<double, false>(h, PSV_FUSED) : launch_sweep<float, false>(h, PSV_FUSED)) return 1;
    } else {   // main.f90:108-110: update_vel -> bodyforce -> absorb_vel
        if (h->fb == 8 ? launch_sweep<double, false>(h, PSV_INTERIOR) : launch_sweep<float, false>(h, PSV_INTERIOR)) return 1;
        if (h->fb == 8 ? launch_source<double>(h, it, true) : launch_source<float>(h, it, true)) return 1;
        if (h->fb == 8 ? launch_sweep<double, false>(h, PSV_ABSORBER) : launch_sweep<float, false>(h, PSV_ABSORBER)) return 1;
    }
    return ktime(h, 1, 1);
}

extern "C" int swpcpsv_wav_store(swpcpsv_handle *h, int32_t it) {
    if (ready(h)) return 1;
    if (h->nst <= 0) return 0;
    PsvWav w{};
    w.nst = h->nst; w.ntw = h->ntw;
    w.sample = ((it - 1) % h->ntdec_w == 0) ? 1 : 0;
    w.itw = (it - 1) / h->ntdec_w + 1;
    if (w.itw > h->ntw) w.sample = 0;
    w.sw_v = h->sw[0]; w.sw_u = h->sw[1]; w.sw_stress = h->sw[2]; w.sw_strain = h->sw[3];
    if (!w.sample && !w.sw_u && !w.sw_strain) return 0;
    w.ik = h->st_ik; w.wav_v = h->wav[0]; w.wav_u = h->wav[1]; w.wav_s = h->wav[2]; w.wav_e = h->wav[3]; w.acc = h->wav_acc;
    w.M0 = h->M0; w.UC = h->UC; w.r40x = h->w40x; w.r40z = h->w40z; w.r41x = h->w41x; w.r41z = h->w41z;
    const int nb = (h->nst + 127) / 128;
    if (h->fb == 8) psv_wav_kernel<double><<<nb, 128, 0, h->st>>>(make_params<double>(h), w);
    else psv_wav_kernel<float><<<nb, 128, 0, h->st>>>(make_params<float>(h), w);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int swpcpsv_vmax(swpcpsv_handle *h, float out[2]) {
    if (ready(h)) return 1;
    const swpcpsv_grid &g = h->g;
    out[0] = out[1] = 0.0f;
    if (g.iend_k < g.ibeg_k) return 0;
    CK(cudaMemsetAsync(h->vmax_d, 0, 2 * sizeof(unsigned int), h->st));
    const int n = g.iend_k - g.ibeg_k + 1;
    const int nb = std::min((n + 255) / 256, 592);
    if (h->fb == 8) psv_vmax_kernel<double><<<nb, 256, 0, h->st>>>(make_params<double>(h), g.ibeg_k - g.ibeg, g.iend_k - g.ibeg, h->vmax_d);
    else psv_vmax_kernel<float><<<nb, 256, 0, h->st>>>(make_params<float>(h), g.ibeg_k - g.ibeg, g.iend_k - g.ibeg, h->vmax_d);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, h->vmax_d, 2 * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}
extern "C" int swpcpsv_vmax_global(swpcpsv_handle *h, float out[2]) {
    if (swpcpsv_vmax(h, out)) return 1;
    if (!h->comm) return 0;
    CK(cudaMemcpyAsync(h->vmax_d, out, 2 * sizeof(float), cudaMemcpyHostToDevice, h->st));
    NK(g_nc.AllReduce(h->vmax_d, h->vmax_d, 2, ncclFloat, ncclMax, h->comm, h->st));
    CK(cudaMemcpyAsync(out, h->vmax_d, 2 * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int swpcpsv_sync(swpcpsv_handle *h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    CK(cudaStreamSynchronize(h->st));
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// halo columns: global__comm_stress m_global.f90:366-418, global__comm_vel :312-364.  dir 0 = towards idx+1 (ip).
// send[0]: sbuf_ip, send[1]: sbuf_im; recv[0]: rbuf_ip -> columns iend+1.., recv[1]: rbuf_im -> columns ibeg-2..
struct PsvLists { PsvCols send[2], recv[2]; };
static PsvLists col_lists(const swpcpsv_handle *h, int which) {
    PsvLists L{};
    const int b = HALO, e = HALO + h->nxp - 1;   // memory columns of ibeg and iend
    void *Vx = h->F[0], *Vz = h->F[1], *Sxx = h->F[2], *Sxz = h->F[4];
    if (which == 1) {
        L.send[0] = {{Vx, Vx, Vz}, {e - 1, e, e}};
        L.send[1] = {{Vx, Vz, Vz}, {b, b, b + 1}};
        L.recv[1] = {{Vx, Vx, Vz}, {b - 2, b - 1, b - 1}};
        L.recv[0] = {{Vx, Vz, Vz}, {e + 1, e + 1, e + 2}};
    } else {
        L.send[0] = {{Sxx, Sxz, Sxz}, {e, e - 1, e}};
        L.send[1] = {{Sxx, Sxx, Sxz}, {b, b + 1, b}};
        L.recv[1] = {{Sxx, Sxz, Sxz}, {b - 1, b - 2, b - 1}};
        L.recv[0] = {{Sxx, Sxx, Sxz}, {e + 1, e + 2, e + 1}};
    }
    return L;
}

template <typename F>
static int launch_halo(swpcpsv_handle *h, const PsvLists &L, bool pack) {
    const int nz = h->g.nz;
    for (int f = 0; f < 2; f++) {
        const bool outer = h->nbr[f] < 0;
        if (outer && pack) continue;
        // an MPI_PROC_NULL side: the reference still unpacks its never-written receive buffer (m_global.f90:352-359), taken as zeros
        dim3 blk(128), grd((unsigned)((nz + 127) / 128), 3);
        F *buf = outer ? nullptr : (F *)(pack ? h->sbuf[f] : h->rbuf[f]);
        if (pack) psv_halo_kernel<F, true><<<grd, blk, 0, h->st>>>(nz, h->NZP, L.send[f], buf);
        else psv_halo_kernel<F, false><<<grd, blk, 0, h->st>>>(nz, h->NZP, L.recv[f], buf);
        h->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

static int comm_exchange(swpcpsv_handle *h, int which) {
    if (ready(h)) return 1;
    const PsvLists L = col_lists(h, which);
    const bool any = h->nbr[0] >= 0 || h->nbr[1] >= 0;
    if (!any) return h->fb == 8 ? launch_halo<double>(h, L, false) : launch_halo<float>(h, L, false);
    if (!h->comm) return fail("swpcpsv_comm_*: this rank has neighbours but swpcpsv_comm_init was not called");
    if (h->fb == 8 ? launch_halo<double>(h, L, true) : launch_halo<float>(h, L, true)) return 1;
    const ncclDataType_t ty = h->fb == 8 ? ncclDouble : ncclFloat;
    const size_t cnt = (size_t)3 * h->g.nz;
    NK(g_nc.GroupStart());
    for (int f = 0; f < 2; f++) {
        if (h->nbr[f] < 0) continue;
        NK(g_nc.Send(h->sbuf[f], cnt, ty, h->nbr[f], h->comm, h->st));
        NK(g_nc.Recv(h->rbuf[f], cnt, ty, h->nbr[f], h->comm, h->st));
    }
    NK(g_nc.GroupEnd());
    return h->fb == 8 ? launch_halo<double>(h, L, false) : launch_halo<float>(h, L, false);
}
extern "C" int swpcpsv_comm_stress(swpcpsv_handle *h) { return comm_exchange(h, 0); }
extern "C" int swpcpsv_comm_vel(swpcpsv_handle *h) { return comm_exchange(h, 1); }

extern "C" int swpcpsv_nccl_unique_id(char id[128]) {
    if (nccl_load()) return 1;
    ncclUniqueId u;
    NK(g_nc.GetUniqueId(&u));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, 128);
    return 0;
}
extern "C" int swpcpsv_comm_init(swpcpsv_handle *h, const char id[128], int32_t nranks, int32_t rank) {
    if (!h) return fail("null handle");
    if (nccl_load()) return 1;
    if (nranks != h->g.nproc_x) return fail("swpcpsv_comm_init: nranks != nproc_x (assert, m_global.f90:199)");
    if (rank != h->g.myid) return fail("swpcpsv_comm_init: rank != myid");
    CK(cudaSetDevice(h->dev));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    NK(g_nc.CommInitRank(&h->comm, nranks, u, rank));
    return 0;
}
extern "C" int swpcpsv_comm_local(swpcpsv_handle **hs, int32_t n, int32_t which) {
    if (!hs || n <= 0) return fail("swpcpsv_comm_local: no handles");
    for (int q = 0; q < n; q++) {
        if (ready(hs[q])) return 1;
        if (hs[q]->g.myid != q) return fail("swpcpsv_comm_local: handles must be ordered by myid");
        const PsvLists L = col_lists(hs[q], which);
        if (hs[q]->fb == 8 ? launch_halo<double>(hs[q], L, true) : launch_halo<float>(hs[q], L, true)) return 1;
    }
    for (int q = 0; q < n; q++) { CK(cudaSetDevice(hs[q]->dev)); CK(cudaStreamSynchronize(hs[q]->st)); }
    for (int q = 0; q < n; q++) {
        swpcpsv_handle *h = hs[q];
        for (int f = 0; f < 2; f++) {
            if (h->nbr[f] < 0) continue;
            if (h->nbr[f] >= n) return fail("swpcpsv_comm_local: neighbour not in the handle list");
            // on the receiver's own non-blocking stream, ahead of its unpack (a default-stream D2D cudaMemcpy orders against neither)
            CK(cudaMemcpyAsync(hs[h->nbr[f]]->rbuf[1 - f], h->sbuf[f], (size_t)3 * h->g.nz * h->fb, cudaMemcpyDefault, hs[h->nbr[f]]->st));
        }
    }
    for (int q = 0; q < n; q++) {
        const PsvLists L = col_lists(hs[q], which);
        CK(cudaSetDevice(hs[q]->dev));
        if (hs[q]->fb == 8 ? launch_halo<double>(hs[q], L, false) : launch_halo<float>(hs[q], L, false)) return 1;
    }
    for (int q = 0; q < n; q++) { CK(cudaSetDevice(hs[q]->dev)); CK(cudaStreamSynchronize(hs[q]->st)); }
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// snapshots, m_snap.f90
extern "C" int swpcpsv_snap_setup(swpcpsv_handle *h, const swpcpsv_snap_cfg *cfg) {
    if (!h || !cfg) return fail("null argument");
    CK(cudaSetDevice(h->dev));
    for (int q = 0; q < 3; q++) { cudaFree(h->snap_buf[q]); h->snap_buf[q] = nullptr; }
    h->snap = *cfg;
    h->snap_on = cfg->sw_ps || cfg->sw_v || cfg->sw_u;
    if (!h->snap_on) return 0;
    if (cfg->idec < 1 || cfg->kdec < 1 || cfg->ntdec_s < 1 || cfg->nxs < 1 || cfg->nzs < 1) return fail("swpcpsv_snap_setup: bad decimation / size");
    if (cfg->is1 >= cfg->is0 && (cfg->is0 * cfg->idec - cfg->idec / 2 < h->g.ibeg || cfg->is1 * cfg->idec - cfg->idec / 2 > h->g.iend || cfg->is1 > cfg->nxs))
        return fail("swpcpsv_snap_setup: is0..is1 outside of the owned columns (m_snap.f90:111-112)");
    if (cfg->ks1 >= cfg->ks0 && (cfg->ks0 * cfg->kdec - cfg->kdec / 2 < 1 || cfg->ks1 * cfg->kdec - cfg->kdec / 2 > h->g.nz || cfg->ks1 > cfg->nzs))
        return fail("swpcpsv_snap_setup: ks0..ks1 outside of 1..nz (m_snap.f90:113-114)");
    const int sw[3] = {cfg->sw_ps, cfg->sw_v, cfg->sw_u};
    const size_t n = (size_t)2 * cfg->nxs * cfg->nzs * sizeof(float);
    for (int q = 0; q < 3; q++)
        if (sw[q]) { CK(cudaMalloc(&h->snap_buf[q], n)); CK(cudaMemsetAsync(h->snap_buf[q], 0, n, h->st)); }
    return 0;
}

extern "C" int swpcpsv_snap_step(swpcpsv_handle *h, int32_t it) {
    if (ready(h)) return 1;
    if (!h->snap_on) return 0;
    const swpcpsv_snap_cfg &c = h->snap;
    const bool sample = (it - 1) % c.ntdec_s == 0;
    if (!(c.sw_u || sample) || c.is1 < c.is0 || c.ks1 < c.ks0) return 0;
    PsvSnap g{};
    g.idec = c.idec; g.kdec = c.kdec; g.nxs = c.nxs; g.nzs = c.nzs; g.is0 = c.is0; g.is1 = c.is1; g.ks0 = c.ks0; g.ks1 = c.ks1; g.ibeg = h->g.ibeg;
    g.do_ps = c.sw_ps && sample; g.do_v = c.sw_v && sample; g.do_u = c.sw_u;
    g.UC = c.UC; g.M0 = c.M0;
    if (h->fb == 8) { g.r20x = 1.0 / h->g.dx; g.r20z = 1.0 / h->g.dz; }
    else { g.r20x = (double)(1.0f / (float)h->g.dx); g.r20z = (double)(1.0f / (float)h->g.dz); }
    g.buf_ps = h->snap_buf[0]; g.buf_v = h->snap_buf[1]; g.buf_u = h->snap_buf[2];
    dim3 blk(128), grd((unsigned)((c.ks1 - c.ks0 + 1 + 127) / 128), (unsigned)(c.is1 - c.is0 + 1));
    if (h->fb == 8) psv_snap_kernel<double><<<grd, blk, 0, h->st>>>(make_params<double>(h), g);
    else psv_snap_kernel<float><<<grd, blk, 0, h->st>>>(make_params<float>(h), g);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

// mpi_reduce(SUM) onto the I/O rank (m_snap.f90:395-417, :500-507): every rank calls; `out` is filled on `root` only
static int snap_reduce_to_host(swpcpsv_handle *h, const float *dev_src, size_t n, int root, float *out) {
    const float *src = dev_src;
    if (h->comm) {
        if (h->snap_tmp_n < n) {
            cudaFree(h->snap_tmp);
            CK(cudaMalloc(&h->snap_tmp, n * sizeof(float)));
            h->snap_tmp_n = n;
        }
        NK(g_nc.AllReduce(dev_src, h->snap_tmp, n, ncclFloat, ncclSum, h->comm, h->st));
        src = h->snap_tmp;
    }
    if ((!h->comm || h->g.myid == root) && out) CK(cudaMemcpyAsync(out, src, n * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}
extern "C" int swpcpsv_snap_fetch(swpcpsv_handle *h, int32_t product, int32_t root, float *out) {
    if (!h) return fail("null handle");
    if (product < 0 || product > 2 || !h->snap_buf[product]) return fail("swpcpsv_snap_fetch: product not enabled (swpcpsv_snap_setup)");
    CK(cudaSetDevice(h->dev));
    return snap_reduce_to_host(h, h->snap_buf[product], (size_t)2 * h->snap.nxs * h->snap.nzs, root, out);
}
extern "C" int swpcpsv_reduce_sum(swpcpsv_handle *h, float *buf, int64_t n, int32_t root) {
    if (!h || !buf || n < 0) return fail("swpcpsv_reduce_sum: bad argument");
    if (!h->comm || n == 0) return 0;
    CK(cudaSetDevice(h->dev));
    float *tmp = nullptr;
    CK(cudaMalloc(&tmp, (size_t)n * sizeof(float)));
    CK(cudaMemcpyAsync(tmp, buf, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, h->st));
    const int rc = snap_reduce_to_host(h, tmp, (size_t)n, root, buf);
    cudaFree(tmp);
    return rc;
}

extern "C" int swpcpsv_step(swpcpsv_handle *h, int32_t it) {   // main.f90:99-111
    if (swpcpsv_snap_step(h, it)) return 1;
    if (swpcpsv_wav_store(h, it)) return 1;
    if (swpcpsv_update_stress(h)) return 1;
    if (swpcpsv_stressglut(h, it)) return 1;
    if (swpcpsv_comm_stress(h)) return 1;
    if (swpcpsv_update_vel(h, it)) return 1;
    if (swpcpsv_comm_vel(h)) return 1;
    return 0;
}
extern "C" int swpcpsv_run(swpcpsv_handle *h, int32_t it0, int32_t it1) {
    for (int it = it0; it <= it1; it++)
        if (swpcpsv_step(h, it)) return 1;
    return 0;
}

extern "C" int swpcpsv_timer_start(swpcpsv_handle *h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    CK(cudaEventRecord(h->ev0, h->st));
    return 0;
}
extern "C" int swpcpsv_timer_stop(swpcpsv_handle *h, float *ms) {
    if (!h || !ms) return fail("null argument");
    CK(cudaSetDevice(h->dev));
    CK(cudaEventRecord(h->ev1, h->st));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return 0;
}

extern "C" int swpcpsv_set_option(swpcpsv_handle *h, const char *key, int32_t value) {
    if (!h || !key) return fail("null argument");
    if (!strcmp(key, "tk")) { if (value < 32 || value > 256 || value % 32) return fail("tk must be a multiple of 32 in 32..256"); h->tk = value; }
    else if (!strcmp(key, "ilen")) { if (value < 1) return fail("ilen must be >= 1"); h->ilen = value; }
    else if (!strcmp(key, "pf")) { if (value < 0 || value > 8) return fail("pf must be 0..8"); h->pf = value; }
    else if (!strcmp(key, "pw_mode")) h->pw_mode = value != 0;
    else if (!strcmp(key, "kernel_timing")) { h->ktiming = value != 0; h->kev_used[0] = h->kev_used[1] = 0; }
    else return fail(std::string("unknown option ") + key);
    return 0;
}
extern "C" int swpcpsv_get_info(swpcpsv_handle *h, const char *key, double *value) {
    if (!h || !key || !value) return fail("null argument");
    if (!strcmp(key, "launches")) *value = (double)h->launches;
    else if (!strcmp(key, "NZP")) *value = h->NZP;
    else if (!strcmp(key, "NXM")) *value = h->NXM;
    else if (!strcmp(key, "naux")) *value = (double)h->naux;
    else if (!strcmp(key, "device")) *value = h->dev;
    else if (!strcmp(key, "cells_interior")) *value = (double)h->cells_interior;
    else if (!strcmp(key, "cells_absorber")) *value = (double)h->cells_absorber;
    else if (!strcmp(key, "state_bytes")) *value = (double)h->ncell * (5.0 * h->fb + 5 * 4 + 3.0 * h->nm * 4) + (double)h->naux * 8 * 4;
    else if (!strcmp(key, "ms_stress") || !strcmp(key, "ms_vel") || !strcmp(key, "n_stress") || !strcmp(key, "n_vel")) {
        const int w = strstr(key, "stress") ? 0 : 1;
        if (key[0] == 'n') { *value = (double)h->kev_used[w]; return 0; }
        CK(cudaSetDevice(h->dev));
        CK(cudaStreamSynchronize(h->st));
        double sum = 0;
        for (size_t q = 0; q < h->kev_used[w]; q++) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, h->kev[w][0][q], h->kev[w][1][q]));
            sum += ms;
        }
        *value = h->kev_used[w] ? sum / (double)h->kev_used[w] : 0.0;
    } else return fail(std::string("unknown info key ") + key);
    return 0;
}
