// abi.cu -- the C ABI declared in include/swpc3d_b200.h: device state, uploads, kernel launches and
// the halo exchange (NCCL send/recv or single-process emulation).  No torch types, no CPU fallback:
// every entry point fails loudly when CUDA is not usable.
#include "../../include/swpc3d_b200.h"
#include "kernels.cuh"
#include "stress_tma.cuh"
#include "pml_tma.cuh"
#include "bottom_tma.cuh"
#include "snap.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <memory>
#include <chrono>
#include <functional>
#include <thread>
#include <string>
#include <vector>

using namespace swpc;

static thread_local std::string g_err;
extern "C" const char *swpc3d_last_error(void) { return g_err.c_str(); }
extern "C" const char *swpc3d_version(void) { return "swpc3d_b200 0.1 (reference: OpenSWPC 25.05.2 swpc_3d)"; }

// development aid: SWPC3D_TRACE=1 prints where the host is (stderr), for diagnosing hangs
static bool g_trace = getenv("SWPC3D_TRACE") != nullptr;
#define TRACE(...) do { if (g_trace) { fprintf(stderr, "[swpc3d %d] ", (int)getpid()); fprintf(stderr, __VA_ARGS__); fprintf(stderr, "\n"); fflush(stderr); } } while (0)

static int fail(const std::string &m) {
    g_err = m;
    return 1;
}
#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            char b_[512];                                                                                \
            snprintf(b_, sizeof(b_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return fail(b_);                                                                             \
        }                                                                                                \
    } while (0)

// ------------------------------------------------------------------------------------------------
// NCCL, loaded lazily so that single-GPU use has no NCCL dependency
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t *) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *) = nullptr;   // optional (NCCL >= 2.18)
};
static NcclApi g_nccl;
static int nccl_load() {
    if (g_nccl.lib) return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) return fail(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define SYM(f, s)                                                         \
    *(void **)(&g_nccl.f) = dlsym(g_nccl.lib, s);                         \
    if (!g_nccl.f) return fail(std::string("libnccl: missing symbol ") + s);
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(AllReduce, "ncclAllReduce");
    SYM(Reduce, "ncclReduce");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
    SYM(CommGetAsyncError, "ncclCommGetAsyncError");
    SYM(CommAbort, "ncclCommAbort");
#undef SYM
    *(void **)(&g_nccl.CommSplit) = dlsym(g_nccl.lib, "ncclCommSplit");
    return 0;
}
#define NK(call)                                                                                       \
    do {                                                                                               \
        ncclResult_t r_ = (call);                                                                      \
        if (r_ != ncclSuccess) {                                                                       \
            char b_[512];                                                                              \
            snprintf(b_, sizeof(b_), "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
            return fail(b_);                                                                           \
        }                                                                                              \
    } while (0)

// ------------------------------------------------------------------------------------------------
struct swpc3d_handle {
    swpc3d_grid g{};
    int dev = 0;
    int nxp = 0, nyp = 0, NZP = 0, NXM = 0, NYM = 0;
    int nzm_h = 0;                         // host k extent (nz + 6 + kpad)
    long long ncell = 0;
    int fb = 8;                            // field bytes
    int nm = 0;
    void *Fall = nullptr;                  // 9 fields, contiguous
    float *Mall = nullptr;                 // 5 medium arrays, contiguous: rho mu lam taup taus
    void *F[9] = {};                       // Vx Vy Vz Sxx Syy Szz Syz Sxz Sxy
    float *R = nullptr;
    float *med[5] = {};                    // rho lam mu taup taus
    int4 *band = nullptr;
    int *kbeg_a = nullptr, *kob = nullptr, *kfs = nullptr;
    std::vector<int> h_kbeg_a;
    std::vector<long long> h_aoff;         // host copy of aoff (per owned column)
    // ADE columns of the bottom rows (kbeg_a > 1) are PACKED: stride = their 20-odd elements rounded up to 4, not to 32 -- a tile's aux
    // box is then one contiguous run of whole 64-byte pieces instead of 80 used bytes per 128-byte line.  (SWPC3D_AUX_PACK=0: the
    // round-1 layout that keeps lane = (k-1) mod 32, for comparison.)
    bool aux_pack = true;
    std::vector<struct PmlPlan *> pml[2];  // [stress | velocity]: TMA-staged absorber shell (pml_tma.cuh), one plan per region swept (whole box, core, boundary slabs)
    struct TmaPlan *tplan[2] = {};         // [whole | core region]: work lists of the persistent interior stress kernel (stress_tma_p)
    int tma_persist = 0;                   // option "tma_persist": 1 = stress_tma_p (persistent blocks, ticket counter) instead of one block per 16-plane chunk; measured slower
    int tma_pl = 64;                       // option "tma_pl": planes per work item of the persistent kernel
    std::vector<struct BotPlan *> bot[2];  // [stress | velocity]: whole-line bottom tiles (bottom_tma.cuh), one plan per region swept
    // options "bottom_tma" / "bot_jl".  Off by default: built, bit-exact, and measured at 1024 x 1024 x 512 -- stress sweep 25.71 -> 25.60 ms,
    // velocity sweep 10.62 -> 11.49 ms (the tile's 256 active cells per plane cannot hide the latency of a plane)
    int use_bot = 0, bot_jl = 82;
    int use_pml = 1;                       // option "pml_tma": 0 = the whole shell with sweep_direct
    int pml_jl = 32, pml_jl_bottom = 41;   // planes per work item (targets; evened out over the region)
    int l2hint = 0;                        // option "l2hint": L2 eviction policy of stress_tma's TMA loads (1: centre boxes evict-first, 2: + halo boxes evict-last)
    int l2promo = 2, l2promo_halo = 2;     // options "l2promo" / "l2promo_halo": L2 promotion of stress_tma's centre / halo boxes (0 none, 1 64 B, 2 128 B, 3 256 B)
    int pml_promo = 1, pml_promo_b = 1;    // options "pml_promo" / "pml_promo_bottom": the same for pml_tma's wall / bottom items
    int pml_promo_aux_b = -1;              // option "pml_promo_aux_bottom": the ADE map of the bottom items alone (-1: as pml_promo_bottom)
    long long *aoff = nullptr;
    float *aux = nullptr;
    long long naux = 0;
    float4 *g4[6] = {};                    // gxc gxe gyc gye gzc gze
    float *cg[6] = {};                     // cerjan gx_c gx_b gy_c gy_b gz_c gz_b
    bool absorber_ready = false, medium_ready = false;
    // coefficients
    double r40[6][2] = {};                 // x40 x41 y40 y41 z40 z41, in F precision but stored as double
    double r20[3] = {};
    float c1[MAXNM] = {}, c2[MAXNM] = {}, d1[MAXNM] = {}, d2 = 0.f;
    // sources
    int nsrc = 0, stf = 3, bf_mode = 0;
    float tbeg = 0.f;
    int *src_ijk = nullptr;
    double *src_mo = nullptr, *src_mij = nullptr;
    float *src_prm = nullptr, *src_stime = nullptr;
    // host-evaluated source-time values of a step, 17..256 sources: pinned host ring -> device ring (a copy from pageable memory would
    // synchronise the host with the launch stream at every step)
    static constexpr int STIME_RING = 32;
    float *stime_pin = nullptr;
    cudaEvent_t stime_ev[STIME_RING] = {};
    unsigned int stime_n = 0;
    float *stime_cur = nullptr;
    std::vector<float> h_prm;
    double dt_dxyz = 0;
    // stations
    int nst = 0, ntdec_w = 0, ntw = 0;
    int *st_ijk = nullptr;
    float *wav = nullptr;                  // velocity traces
    float *wav_u = nullptr, *wav_s = nullptr, *wav_e = nullptr, *wav_acc = nullptr;
    int sw_v = 1, sw_u = 0, sw_stress = 0, sw_strain = 0;
    float M0 = 1.f, UC = 1e-15f;
    // Green's-function mode (m_green.f90)
    int ng = 0, g_ncmp = 6, g_bforce = 0, g_ntdec_w = 1, g_ntw = 0, g_stf = 3, g_is_src = 0, g_src[3] = {0, 0, 0};
    float g_f[3] = {0, 0, 0}, g_trise = 1.0f, g_dt_dxyz = 0.0f, g_tbeg = 0.0f;
    int *g_ijk = nullptr;
    float *g_acc = nullptr, *g_gf = nullptr;
    unsigned int *vmax_d = nullptr;
    float *vmax_h = nullptr;               // pinned: a device-to-host copy into pageable memory blocks the host until the stream gets there,
                                           // i.e. for ever behind an exchange whose peer is gone -- before any time-out could fire
    // snapshots
    swpc3d_snap_cfg snap{};
    bool snap_on = false;
    float *snap_buf[15] = {}, *snap_max[15] = {}, *snap_tmp = nullptr;
    size_t snap_tmp_n = 0;
    // asynchronous fetch (swpc3d_snap_fetch_begin / _end): device staging copy, reduce target, double-buffered pinned host memory
    cudaStream_t ss = nullptr;
    float *snap_stage[15] = {}, *snap_red[15] = {}, *snap_pin[15][2] = {};
    cudaEvent_t snap_ev_staged[15] = {}, snap_ev_done[15][2] = {};
    bool snap_inflight[15] = {};
    int snap_last_slot[15] = {};
    ncclComm_t comm_io = nullptr;          // a communicator of its own for the snapshot reductions (they run beside the halo exchange)
    // halo
    void *sbuf[4] = {}, *rbuf[4] = {};     // 0: +x (ip) 1: -x (im) 2: +y (jp) 3: -y (jm)
    int nbr[4] = {-1, -1, -1, -1};
    ncclComm_t comm = nullptr;
    int comm_rank = -1, comm_size = 0;
    cudaEvent_t nccl_ev[4] = {};           // the last four NCCL exchanges: the host never runs more than four exchanges ahead of the device,
    unsigned int nccl_n = 0;               // so a peer that stopped is noticed here (a full NCCL work queue would block ncclGroupEnd for ever)
    int comm_timeout_s = 1800;             // option "comm_timeout_s": a host-side wait on a stream that carries NCCL work gives up after this
    bool comm_dead = false;                // the communicator was aborted (peer failure / timeout): every later exchange fails at once
    // peer-to-peer exchange (one node): a face's planes are stored straight into the neighbour's receive buffer (halo_push / halo_pull)
    int use_p2p = 1;                       // option "p2p": 0 = pack + ncclSend/ncclRecv + unpack
    bool p2p_ok = false;
    char *p2p_base = nullptr;              // my block: receive buffers [family][parity][face], flag words [family][face], block counters
    size_t p2p_off[2][2][4] = {};
    size_t p2p_flag_off = 0, p2p_count_off = 0, p2p_bytes = 0;
    void *p2p_peer[4] = {};                // the neighbours' blocks, opened through CUDA IPC
    size_t p2p_peer_off[4][2][2] = {};     // [face][family][parity]: where my message lands in the neighbour's block
    size_t p2p_peer_flag[4][2] = {};       // [face][family]: the neighbour's flag word for messages from me
    unsigned int p2p_seq[2] = {0, 0};      // exchanges done per family
    // streams
    cudaStream_t st = nullptr;
    cudaStream_t side[5] = {};             // absorber-shell boxes run beside the TMA interior kernel
    cudaEvent_t ev_fork = nullptr, ev_join[5] = {};
    int use_side = 1;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // tuning
    int tk = 64, ti = 4, jlen = 16, pf = 1;
    int tma_shift = 1;              // option "tma_shift": shift stress_tma's partial last k-tile up instead of masking absorber rows
    int use_tma = 1, tma_jl = 16;   // use_tma: 0 off, 1 stress sweep only (default: measured fastest), 2 stress + velocity sweeps
    bool tma_ready = false, tma_ok = false;
    TmaMaps tmaps{};
    TmaMapsVel vmaps{};
    bool vtma_ok = false;
    int variant = 1;
    int use_ring = 1, ring_jlen = 32, ring_pf = 2;   // vel_ring: register-pipelined interior velocity sweep
    int ring_pair = 1;                               // option "ring_pair": float32 fields use vel_ring2 (two cells per thread)
    int flat_bottom = 1;                             // flat thread numbering for the bottom absorber slab (Box3::flat)
    // boundary-first overlap of the halo exchange (swpc3d_step): the two outermost owned planes towards every neighbour are
    // swept first, then pack / NCCL / unpack run on `cs` while the core of the subdomain is swept on `st`
    cudaStream_t cs = nullptr;
    cudaStream_t cur = nullptr;   // set while the boundary slabs are launched: the sweeps' kernels go to this stream, no side streams
    cudaEvent_t ev_b = nullptr, ev_c = nullptr;
    int overlap = 1;       // 0: the reference's fully exposed order
    int slab_x = 8;        // option "slab_x": columns of the x boundary slabs (>= 2; 8 = one tile column of stress_tma)
    int slab_tiled = 1;    // option "slab_tiled": 0 = the slabs with sweep_direct
    int split_test = 0;    // testing aid: split the sweeps as if all four faces had neighbours, without any exchange
    int pw_mode = 0;   // plane-wave mode: edge extrapolation ahead of the PML sweeps
    int zero_outer = 0;   // re-zero the outer halo planes at every exchange (see launch_halo)
    long long launches = 0;
    // per-kernel CUDA-event timing of the two sweeps (option "kernel_timing"): event pairs recorded on the launch
    // stream, read back after a synchronisation by swpc3d_get_info("ms_stress" / "ms_vel")
    bool ktiming = false;
    std::vector<cudaEvent_t> kev[2][2];   // [stress|vel][begin|end]
    size_t kev_used[2] = {0, 0};
    std::vector<cudaEvent_t> cev[2];      // halo exchange (pack + NCCL + unpack) [begin|end]
    size_t cev_used = 0;
    double halo_bytes = 0.0;              // bytes this rank sent since kernel_timing was switched on
};

// Host-side wait on a stream.  Without a communicator this is cudaStreamSynchronize.  With one, the stream may carry NCCL
// work that never completes when a peer has died or diverged (it left the time loop through the divergence abort of
// m_report.f90:144-151, say): poll the stream and ncclCommGetAsyncError instead, and turn an asynchronous NCCL error or a
// timeout (option "comm_timeout_s") into the reference's clean abort -- ncclCommAbort, an error message, a non-zero return
// -- where a blocking synchronise would hang for ever.
// the peer-to-peer exchange's own failure flag: halo_wait gave up on a neighbour that never pushed
static int p2p_check(swpc3d_handle *h) {
    if (!h->p2p_ok || !h->p2p_base) return 0;
    unsigned int err = 0;
    if (cudaMemcpy(&err, (unsigned int *)(h->p2p_base + h->p2p_count_off) + 16, sizeof(err), cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    if (!err) return 0;
    char b[256];
    snprintf(b, sizeof(b), "halo exchange aborted on rank %d: the neighbour across face %u sent nothing for %d s (it died or left the time loop)",
             h->comm_rank, err - 1, h->comm_timeout_s);
    h->comm_dead = true;
    return fail(b);
}
template <typename Query>
static int poll_wait(swpc3d_handle *h, Query query) {
    const auto t0 = std::chrono::steady_clock::now();
    for (long long spin = 0;; spin++) {
        const cudaError_t e = query();
        if (e == cudaSuccess) return p2p_check(h);
        if (e != cudaErrorNotReady) {
            char b[256];
            snprintf(b, sizeof(b), "poll_wait: %s", cudaGetErrorString(e));
            return fail(b);
        }
        if ((spin & 255) == 255) {
            ncclResult_t ar = ncclSuccess;
            const ncclResult_t qr = (h->comm && g_nccl.CommGetAsyncError) ? g_nccl.CommGetAsyncError(h->comm, &ar) : ncclSuccess;
            const double waited = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            const bool timeout = h->comm_timeout_s > 0 && waited > (double)h->comm_timeout_s;
            if (qr != ncclSuccess || (ar != ncclSuccess && ar != ncclInProgress) || timeout) {
                TRACE("poll_wait: giving up after %.1f s (qr %d ar %d): aborting the communicators", waited, (int)qr, (int)ar);
                char b[384];
                snprintf(b, sizeof(b), "halo exchange aborted on rank %d: %s (waited %.1f s); communicator destroyed with ncclCommAbort",
                         h->comm_rank, timeout ? "time-out, a neighbour rank stopped exchanging" : g_nccl.GetErrorString(qr != ncclSuccess ? qr : ar), waited);
                // ncclCommAbort can itself block while a kernel of the communicator sits on the device waiting for the lost peer:
                // abort from a helper thread and give it a few seconds; the error is returned to the caller either way
                {
                    ncclComm_t c0 = h->comm, c1 = h->comm_io;
                    auto abort_fn = g_nccl.CommAbort;
                    auto done = std::make_shared<std::atomic<int>>(0);
                    std::thread([c0, c1, abort_fn, done]() {
                        if (abort_fn) {
                            if (c0) abort_fn(c0);
                            if (c1) abort_fn(c1);
                        }
                        done->store(1);
                    }).detach();
                    for (int q = 0; q < 100 && !done->load(); q++) std::this_thread::sleep_for(std::chrono::milliseconds(50));
                    TRACE("poll_wait: abort %s", done->load() ? "completed" : "still running in its helper thread");
                }
                h->comm = nullptr;
                h->comm_io = nullptr;
                h->comm_dead = true;
                TRACE("poll_wait: communicators aborted");
                return fail(b);
            }
            if (waited > 2e-3) std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
    }
}
static int stream_wait(swpc3d_handle *h, cudaStream_t st) {
    if (!h->comm) {
        CK(cudaStreamSynchronize(st));
        return 0;
    }
    return poll_wait(h, [st]() { return cudaStreamQuery(st); });
}
static int event_wait(swpc3d_handle *h, cudaEvent_t ev) {
    if (!h->comm) {
        CK(cudaEventSynchronize(ev));
        return 0;
    }
    return poll_wait(h, [ev]() { return cudaEventQuery(ev); });
}
#define WAIT(h_, st_)                        \
    do {                                     \
        if (stream_wait((h_), (st_))) return 1; \
    } while (0)

static void pml_drop(swpc3d_handle *h);
static void tplan_drop(swpc3d_handle *h);
static void bot_drop(swpc3d_handle *h);

// the aux allocation of a column whose first element is k = kb: `lead` unused elements, then nz - kb + 1 values, `klen` elements in all
static inline int aux_lead(const swpc3d_handle *h, int kb) { return (h->aux_pack && kb > 1) ? (kb - 1) % 4 : (kb - 1) & 31; }
static inline long long aux_klen(const swpc3d_handle *h, int kb) {
    const long long n = aux_lead(h, kb) + (h->g.nz - kb + 1);
    return (h->aux_pack && kb > 1) ? (n + 3) / 4 * 4 : (n + 31) / 32 * 32;
}
static inline cudaStream_t main_stream(const swpc3d_handle *h) { return h->cur ? h->cur : h->st; }
static inline bool side_streams(const swpc3d_handle *h) { return h->use_side && !h->cur; }
static inline long long col_of(const swpc3d_handle *h, int mi, int mj) { return (long long)mi + (long long)h->NXM * mj; }

template <typename F>
static KParams<F> make_params(const swpc3d_handle *h) {
    KParams<F> p{};
    const swpc3d_grid &g = h->g;
    p.nz = g.nz; p.nxp = h->nxp; p.nyp = h->nyp;
    p.NZP = h->NZP; p.NXM = h->NXM; p.NYM = h->NYM;
    p.SI = h->NZP; p.SJ = (long long)h->NZP * h->NXM;
    p.ncell = h->ncell;
    p.li0_k = g.ibeg_k - g.ibeg; p.li1_k = g.iend_k - g.ibeg;
    p.lj0_k = g.jbeg_k - g.jbeg; p.lj1_k = g.jend_k - g.jbeg;
    p.k1_k = g.kend_k;
    p.abc = g.abc_type;
    p.Vx = (F *)h->F[0]; p.Vy = (F *)h->F[1]; p.Vz = (F *)h->F[2];
    p.Sxx = (F *)h->F[3]; p.Syy = (F *)h->F[4]; p.Szz = (F *)h->F[5];
    p.Syz = (F *)h->F[6]; p.Sxz = (F *)h->F[7]; p.Sxy = (F *)h->F[8];
    p.R = h->R;
    p.rho = h->med[0]; p.lam = h->med[1]; p.mu = h->med[2]; p.taup = h->med[3]; p.taus = h->med[4];
    p.band = h->band; p.kbeg_a = h->kbeg_a; p.kob = h->kob;
    p.aoff = h->aoff; p.aux = h->aux; p.naux = h->naux;
    p.gxc = h->g4[0]; p.gxe = h->g4[1]; p.gyc = h->g4[2]; p.gye = h->g4[3]; p.gzc = h->g4[4]; p.gze = h->g4[5];
    p.cgx_c = h->cg[0]; p.cgx_b = h->cg[1]; p.cgy_c = h->cg[2]; p.cgy_b = h->cg[3]; p.cgz_c = h->cg[4]; p.cgz_b = h->cg[5];
    for (int o = 0; o < 2; o++) {
        p.r40x[o] = (F)h->r40[0][o]; p.r41x[o] = (F)h->r40[1][o];
        p.r40y[o] = (F)h->r40[2][o]; p.r41y[o] = (F)h->r40[3][o];
        p.r40z[o] = (F)h->r40[4][o]; p.r41z[o] = (F)h->r40[5][o];
    }
    p.r20x = (F)h->r20[0]; p.r20y = (F)h->r20[1]; p.r20z = (F)h->r20[2];
    for (int m = 0; m < MAXNM; m++) { p.c1[m] = h->c1[m]; p.c2[m] = h->c2[m]; p.d1[m] = h->d1[m]; }
    p.d2 = h->d2;
    p.dt = g.dt;
    return p;
}

// kernel__setup coefficients (m_kernel.f90:43-67) in the kind F, and r20 (m_absorb_p.f90:70-72)
template <typename F>
static void setup_coefs(swpc3d_handle *h, const float *ts) {
    const swpc3d_grid &g = h->g;
    const F d[3] = {(F)g.dx, (F)g.dy, (F)g.dz};
    for (int a = 0; a < 3; a++) {
        const F rc40 = (F)17.0 / (F)16.0 / d[a], rc41 = (F)1.0 / (F)48.0 / d[a];
        const F rd40 = -(F)1.0 / (F)16.0 / d[a], rd41 = -(F)1.0 / (F)48.0 / d[a];
        h->r40[2 * a][0] = (double)(F)(rc40 + (-1) * rd40);
        h->r40[2 * a][1] = (double)(F)(rc40 + (1) * rd40);
        h->r40[2 * a + 1][0] = (double)(F)(rc41 + (-1) * rd41);
        h->r40[2 * a + 1][1] = (double)(F)(rc41 + (1) * rd41);
        h->r20[a] = (double)(F)((F)1.0f / d[a]);
    }
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
    h->dt_dxyz = (double)((F)dt / ((F)g.dx * (F)g.dy * (F)g.dz));   // m_source.f90:306
}

static int create_state(swpc3d_handle *h, const swpc3d_grid *g, const float *ts);

extern "C" int swpc3d_create(const swpc3d_grid *g, const float *ts, swpc3d_handle **out) {
    if (!g || !out) return fail("swpc3d_create: null argument");
    *out = nullptr;
    if (g->field_bytes != 8 && g->field_bytes != 4) return fail("field_bytes must be 8 (MP=DP) or 4 (MP=SP)");
    if (g->nm < 0 || g->nm > MAXNM) return fail("nm must be 0..3");
    if (g->abc_type != SWPC3D_ABC_PML && g->abc_type != SWPC3D_ABC_CERJAN) return fail("abc_type must be 1 (pml) or 2 (cerjan)");
    if (g->nm > 0 && !ts) return fail("ts[nm] required when nm > 0");
    if (g->iend < g->ibeg || g->jend < g->jbeg || g->nz < 1) return fail("empty subdomain");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(std::string("no CUDA device available (the swpc3d_b200 path has no CPU fallback): ") + cudaGetErrorString(e));
    swpc3d_handle *h = new swpc3d_handle();
    h->g = *g;
    h->dev = g->device >= 0 ? g->device : (g->myid % ndev);   // m_global.f90:208-214
    if (const char *e = getenv("SWPC3D_AUX_PACK")) h->aux_pack = atoi(e) != 0;
    const int rc = create_state(h, g, ts);
    if (rc) {   // e.g. out of device memory half way: give everything back, keep the message
        const std::string msg = swpc3d_last_error();
        swpc3d_destroy(h);
        cudaGetLastError();   // a failed cudaMalloc is not sticky, but it stays the "last error" until it is read
        return fail(msg);
    }
    *out = h;
    return 0;
}

static int create_state(swpc3d_handle *h, const swpc3d_grid *g, const float *ts) {
    CK(cudaSetDevice(h->dev));
    h->nxp = g->iend - g->ibeg + 1;
    h->nyp = g->jend - g->jbeg + 1;
    h->NXM = h->nxp + 2 * HALO + g->ipad;
    h->NYM = h->nyp + 2 * HALO + g->jpad;
    h->nzm_h = g->nz + 6 + g->kpad;
    h->NZP = ((g->nz + g->kpad + KOFF + 3) + 31) / 32 * 32;
    h->ncell = (long long)h->NZP * h->NXM * h->NYM;
    h->fb = g->field_bytes;
    h->nm = g->nm;
    if (h->fb == 8) setup_coefs<double>(h, ts);
    else setup_coefs<float>(h, ts);
    CK(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking));
    for (int q = 0; q < 5; q++) {
        CK(cudaStreamCreateWithFlags(&h->side[q], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->ev_join[q], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    {   // the exchange stream outranks the sweeps: its short kernels (slabs, pack, NCCL, unpack) take the next free SM slots
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&h->cs, cudaStreamNonBlocking, hi));
    }
    CK(cudaEventCreateWithFlags(&h->ev_b, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&h->ev_c, cudaEventDisableTiming));
    CK(cudaEventCreate(&h->ev0));
    CK(cudaEventCreate(&h->ev1));
    // the nine fields live in ONE allocation (field f at f*ncell) so that a single 4-D TMA tensor covers them
    CK(cudaMalloc(&h->Fall, (size_t)h->ncell * h->fb * 9));
    CK(cudaMemsetAsync(h->Fall, 0, (size_t)h->ncell * h->fb * 9, h->st));
    {   // device slot order: Vx Vy Vz | Sxx Szz Sxz | Syy Syz Sxy  (the velocity sweep needs the last three as a j-ring,
        // the middle three only in-plane; F[] keeps the reference's order Vx Vy Vz Sxx Syy Szz Syz Sxz Sxy)
        const int slot[9] = {0, 1, 2, 3, 6, 4, 7, 5, 8};
        for (int a = 0; a < 9; a++) h->F[a] = (char *)h->Fall + (size_t)slot[a] * h->ncell * h->fb;
    }
    if (h->nm > 0) {
        CK(cudaMalloc(&h->R, (size_t)h->ncell * 6 * h->nm * sizeof(float)));
        CK(cudaMemsetAsync(h->R, 0, (size_t)h->ncell * 6 * h->nm * sizeof(float), h->st));
    }
    // medium: one allocation, order rho, mu, lam, taup, taus (lam..taus consecutive for one TMA box)
    CK(cudaMalloc(&h->Mall, (size_t)h->ncell * sizeof(float) * 5));
    CK(cudaMemsetAsync(h->Mall, 0, (size_t)h->ncell * sizeof(float) * 5, h->st));
    {
        const int slot[5] = {0, 2, 1, 3, 4};   // med[] order is rho lam mu taup taus
        for (int a = 0; a < 5; a++) h->med[a] = h->Mall + (size_t)slot[a] * h->ncell;
    }
    const size_t n2 = (size_t)h->NXM * h->NYM;
    CK(cudaMalloc(&h->band, n2 * sizeof(int4)));
    CK(cudaMalloc(&h->kbeg_a, n2 * sizeof(int)));
    CK(cudaMalloc(&h->kob, n2 * sizeof(int)));
    CK(cudaMalloc(&h->kfs, n2 * sizeof(int)));
    CK(cudaMemsetAsync(h->kfs, 0, n2 * sizeof(int), h->st));
    CK(cudaMemsetAsync(h->band, 0, n2 * sizeof(int4), h->st));
    CK(cudaMemsetAsync(h->kob, 0, n2 * sizeof(int), h->st));
    // kbeg_a: m_global.f90:334-343
    h->h_kbeg_a.resize(n2);
    for (int mj = 0; mj < h->NYM; mj++)
        for (int mi = 0; mi < h->NXM; mi++) {
            const int i = g->ibeg - HALO + mi, j = g->jbeg - HALO + mj;
            const bool wall = (i <= g->na || g->nx - g->na + 1 <= i || j <= g->na || g->ny - g->na + 1 <= j);
            h->h_kbeg_a[(size_t)mi + (size_t)h->NXM * mj] = wall ? 1 : g->nz - g->na + 1;
        }
    CK(cudaMemcpyAsync(h->kbeg_a, h->h_kbeg_a.data(), n2 * sizeof(int), cudaMemcpyHostToDevice, h->st));
    CK(cudaMalloc(&h->vmax_d, 3 * sizeof(unsigned int)));
    CK(cudaMallocHost(&h->vmax_h, 3 * sizeof(float)));
    // halo buffers: 5 planes per face, m_global.f90:251-258
    const size_t isz = (size_t)5 * h->nyp * g->nz * h->fb, jsz = (size_t)5 * h->nxp * g->nz * h->fb;
    for (int f = 0; f < 4; f++) {
        const size_t sz = f < 2 ? isz : jsz;
        CK(cudaMalloc(&h->sbuf[f], sz));
        CK(cudaMalloc(&h->rbuf[f], sz));
        CK(cudaMemsetAsync(h->sbuf[f], 0, sz, h->st));
        CK(cudaMemsetAsync(h->rbuf[f], 0, sz, h->st));
    }
    // neighbour table: itbl, m_global.f90:624-642
    const int idx = g->myid % g->nproc_x, idy = g->myid / g->nproc_x;
    h->nbr[0] = (idx + 1 < g->nproc_x) ? g->myid + 1 : -1;
    h->nbr[1] = (idx - 1 >= 0) ? g->myid - 1 : -1;
    h->nbr[2] = (idy + 1 < g->nproc_y) ? g->myid + g->nproc_x : -1;
    h->nbr[3] = (idy - 1 >= 0) ? g->myid - g->nproc_x : -1;
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int swpc3d_destroy(swpc3d_handle *h) {
    if (!h) return 0;
    TRACE("destroy");
    cudaSetDevice(h->dev);
    if (h->comm_dead) {   // after a communication failure kernels may still sit on the device waiting for the lost peer: every
        delete h;         // cudaFree would wait for them.  The caller is about to stop (m_report.f90:144-151); the context goes with the process.
        return 0;
    }
    cudaDeviceSynchronize();
    for (int f = 0; f < 4; f++)
        if (h->p2p_peer[f]) cudaIpcCloseMemHandle(h->p2p_peer[f]);
    for (cudaEvent_t e : h->nccl_ev)
        if (e) cudaEventDestroy(e);
    cudaFree(h->p2p_base);
    if (h->comm_io && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm_io);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (int q = 0; q < 15; q++) {
        cudaFree(h->snap_stage[q]); cudaFree(h->snap_red[q]);
        if (h->snap_ev_staged[q]) cudaEventDestroy(h->snap_ev_staged[q]);
        for (int b = 0; b < 2; b++) {
            if (h->snap_pin[q][b]) cudaFreeHost(h->snap_pin[q][b]);
            if (h->snap_ev_done[q][b]) cudaEventDestroy(h->snap_ev_done[q][b]);
        }
    }
    if (h->ss) cudaStreamDestroy(h->ss);
    cudaFree(h->Fall);
    cudaFree(h->R);
    cudaFree(h->Mall);
    pml_drop(h);
    tplan_drop(h);
    bot_drop(h);
    cudaFree(h->band); cudaFree(h->kbeg_a); cudaFree(h->kob); cudaFree(h->kfs); cudaFree(h->snap_tmp);
    for (int q = 0; q < 15; q++) { cudaFree(h->snap_buf[q]); cudaFree(h->snap_max[q]); } cudaFree(h->aoff); cudaFree(h->aux);
    for (int a = 0; a < 6; a++) { cudaFree(h->g4[a]); cudaFree(h->cg[a]); }
    cudaFree(h->src_ijk); cudaFree(h->src_mo); cudaFree(h->src_mij); cudaFree(h->src_prm); cudaFree(h->src_stime);
    cudaFree(h->g_ijk); cudaFree(h->g_acc); cudaFree(h->g_gf);
    cudaFree(h->st_ijk); cudaFree(h->wav); cudaFree(h->wav_u); cudaFree(h->wav_s); cudaFree(h->wav_e); cudaFree(h->wav_acc); cudaFree(h->vmax_d);
    if (h->vmax_h) cudaFreeHost(h->vmax_h);
    if (h->stime_pin) cudaFreeHost(h->stime_pin);
    for (cudaEvent_t e : h->stime_ev)
        if (e) cudaEventDestroy(e);
    for (int f = 0; f < 4; f++) { cudaFree(h->sbuf[f]); cudaFree(h->rbuf[f]); }
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (int q = 0; q < 5; q++) {
        if (h->side[q]) cudaStreamDestroy(h->side[q]);
        if (h->ev_join[q]) cudaEventDestroy(h->ev_join[q]);
    }
    for (auto &per_sweep : h->kev)            // kernel_timing stopwatches
        for (auto &side : per_sweep)
            for (cudaEvent_t e : side) cudaEventDestroy(e);
    for (auto &side : h->cev)
        for (cudaEvent_t e : side) cudaEventDestroy(e);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_b) cudaEventDestroy(h->ev_b);
    if (h->ev_c) cudaEventDestroy(h->ev_c);
    if (h->cs) cudaStreamDestroy(h->cs);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
    return 0;
}

// host (reference layout, k from -2) <-> device (padded) copy of one 3-D array
static int copy3d(swpc3d_handle *h, void *dev, const void *host_c, void *host_m, size_t elem, bool to_device) {
    // host row = one column: nzm_h elements starting at k = -2  <->  device index KOFF-1-2 = KOFF-3
    const size_t rows = (size_t)h->NXM * h->NYM;
    char *d = (char *)dev + (size_t)(KOFF - 3) * elem;
    if (to_device)
        CK(cudaMemcpy2DAsync(d, (size_t)h->NZP * elem, host_c, (size_t)h->nzm_h * elem, (size_t)h->nzm_h * elem, rows,
                             cudaMemcpyHostToDevice, h->st));
    else
        CK(cudaMemcpy2DAsync(host_m, (size_t)h->nzm_h * elem, d, (size_t)h->NZP * elem, (size_t)h->nzm_h * elem, rows,
                             cudaMemcpyDeviceToHost, h->st));
    return 0;
}

extern "C" int swpc3d_upload_medium(swpc3d_handle *h, const float *rho, const float *lam, const float *mu, const float *taup,
                                    const float *taus, const int32_t *kfs, const int32_t *kob, const int32_t *kfs_top,
                                    const int32_t *kfs_bot, const int32_t *kob_top, const int32_t *kob_bot,
                                    const int32_t *kbeg_a) {
    if (!h) return fail("null handle");
    if (!rho || !lam || !mu || !taup || !taus || !kob || !kfs_top || !kfs_bot || !kob_top || !kob_bot)
        return fail("swpc3d_upload_medium: null array");
    CK(cudaSetDevice(h->dev));
    const float *src[5] = {rho, lam, mu, taup, taus};
    for (int a = 0; a < 5; a++)
        if (copy3d(h, h->med[a], src[a], nullptr, sizeof(float), true)) return 1;
    const size_t n2 = (size_t)h->NXM * h->NYM;
    std::vector<int4> band(n2);
    for (size_t n = 0; n < n2; n++) band[n] = make_int4(kfs_top[n], kfs_bot[n], kob_top[n], kob_bot[n]);
    // the reference defines the bands on owned (i,j) only (m_medium.f90:376-386); make the rest inert
    for (int mj = 0; mj < h->NYM; mj++)
        for (int mi = 0; mi < h->NXM; mi++)
            if (mi < HALO || mi >= HALO + h->nxp || mj < HALO || mj >= HALO + h->nyp) band[(size_t)mi + (size_t)h->NXM * mj] = make_int4(1, 1, 1, 1);
    CK(cudaMemcpyAsync(h->band, band.data(), n2 * sizeof(int4), cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->kob, kob, n2 * sizeof(int), cudaMemcpyHostToDevice, h->st));
    if (kfs) CK(cudaMemcpyAsync(h->kfs, kfs, n2 * sizeof(int), cudaMemcpyHostToDevice, h->st));
    if (kbeg_a) {
        for (int lj = 0; lj < h->nyp; lj++)
            for (int li = 0; li < h->nxp; li++) {
                const size_t n = (size_t)(li + HALO) + (size_t)h->NXM * (lj + HALO);
                if (kbeg_a[n] != h->h_kbeg_a[n]) return fail("swpc3d_upload_medium: kbeg_a differs from m_global.f90:334-343");
            }
    }
    CK(cudaStreamSynchronize(h->st));
    h->medium_ready = true;
    return 0;
}

extern "C" int swpc3d_upload_fields(swpc3d_handle *h, const void *Vx, const void *Vy, const void *Vz, const void *Sxx,
                                    const void *Syy, const void *Szz, const void *Syz, const void *Sxz, const void *Sxy) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    const void *src[9] = {Vx, Vy, Vz, Sxx, Syy, Szz, Syz, Sxz, Sxy};
    for (int a = 0; a < 9; a++)
        if (src[a] && copy3d(h, h->F[a], src[a], nullptr, (size_t)h->fb, true)) return 1;
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int swpc3d_download_fields(swpc3d_handle *h, void *Vx, void *Vy, void *Vz, void *Sxx, void *Syy, void *Szz,
                                      void *Syz, void *Sxz, void *Sxy) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    void *dst[9] = {Vx, Vy, Vz, Sxx, Syy, Szz, Syz, Sxz, Sxy};
    for (int a = 0; a < 9; a++)
        if (dst[a] && copy3d(h, h->F[a], nullptr, dst[a], (size_t)h->fb, false)) return 1;
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

static void snap_dims(const swpc3d_snap_cfg &c, int product, int &n1, int &n2, int &nvar);

extern "C" int swpc3d_zero_state(swpc3d_handle *h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    for (int a = 0; a < 9; a++) CK(cudaMemsetAsync(h->F[a], 0, (size_t)h->ncell * h->fb, h->st));
    if (h->R) CK(cudaMemsetAsync(h->R, 0, (size_t)h->ncell * 6 * h->nm * sizeof(float), h->st));
    if (h->aux) CK(cudaMemsetAsync(h->aux, 0, (size_t)h->naux * 18 * sizeof(float), h->st));
    if (h->wav) CK(cudaMemsetAsync(h->wav, 0, (size_t)h->ntw * 3 * h->nst * sizeof(float), h->st));
    if (h->wav_acc) CK(cudaMemsetAsync(h->wav_acc, 0, (size_t)9 * h->nst * sizeof(float), h->st));
    // every accumulating product: displacement / stress / strain traces, Green's-function sums, snapshot slices and maxima
    const size_t n3 = (size_t)h->ntw * 3 * h->nst * sizeof(float);
    if (h->wav_u) CK(cudaMemsetAsync(h->wav_u, 0, n3, h->st));
    if (h->wav_s) CK(cudaMemsetAsync(h->wav_s, 0, 2 * n3, h->st));
    if (h->wav_e) CK(cudaMemsetAsync(h->wav_e, 0, 2 * n3, h->st));
    if (h->g_acc) CK(cudaMemsetAsync(h->g_acc, 0, (size_t)12 * h->ng * sizeof(float), h->st));
    if (h->g_gf) CK(cudaMemsetAsync(h->g_gf, 0, (size_t)h->g_ntw * h->g_ncmp * h->ng * sizeof(float), h->st));
    for (int q = 0; q < 15; q++) {
        if (!h->snap_buf[q]) continue;
        int n1, n2, nvar;
        snap_dims(h->snap, q, n1, n2, nvar);
        CK(cudaMemsetAsync(h->snap_buf[q], 0, (size_t)n1 * n2 * nvar * sizeof(float), h->st));
        if (h->snap_max[q]) CK(cudaMemsetAsync(h->snap_max[q], 0, (size_t)n1 * n2 * 3 * sizeof(float), h->st));
    }
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int swpc3d_setup_pml(swpc3d_handle *h, const float *gxc, const float *gxe, const float *gyc, const float *gye,
                                const float *gzc, const float *gze) {
    if (!h) return fail("null handle");
    if (h->g.abc_type != SWPC3D_ABC_PML) return fail("swpc3d_setup_pml: abc_type is not pml");
    if (!gxc || !gxe || !gyc || !gye || !gzc || !gze) return fail("swpc3d_setup_pml: null profile");
    CK(cudaSetDevice(h->dev));
    const float *src[6] = {gxc, gxe, gyc, gye, gzc, gze};
    const int len[6] = {h->nxp, h->nxp, h->nyp, h->nyp, h->g.nz, h->g.nz};
    for (int a = 0; a < 6; a++) {
        if (!h->g4[a]) CK(cudaMalloc(&h->g4[a], (size_t)len[a] * sizeof(float4)));
        CK(cudaMemcpyAsync(h->g4[a], src[a], (size_t)len[a] * sizeof(float4), cudaMemcpyHostToDevice, h->st));
    }
    // shell-only ADE storage: column (i,j) holds k = kbeg_a(i,j)..nz, columns start on 32-element boundaries
    std::vector<long long> aoff((size_t)h->nxp * h->nyp);
    long long off = 0;
    for (int lj = 0; lj < h->nyp; lj++)
        for (int li = 0; li < h->nxp; li++) {
            const int kb = h->h_kbeg_a[(size_t)(li + HALO) + (size_t)h->NXM * (lj + HALO)];
            const int len_k = h->g.nz - kb + 1;
            // wall columns keep their lane phase: element k sits at off + (k - kb) with off = 32*q + ((kb-1) & 31); the short columns
            // of the bottom rows are packed (aux_lead / aux_klen)
            (void)len_k;
            aoff[(size_t)li + (size_t)h->nxp * lj] = off + aux_lead(h, kb);
            off += aux_klen(h, kb);
        }
    h->naux = off;
    h->h_aoff = aoff;
    pml_drop(h);
    bot_drop(h);
    if (h->aoff) cudaFree(h->aoff);
    if (h->aux) cudaFree(h->aux);
    h->aoff = nullptr; h->aux = nullptr;
    CK(cudaMalloc(&h->aoff, aoff.size() * sizeof(long long)));
    CK(cudaMemcpyAsync(h->aoff, aoff.data(), aoff.size() * sizeof(long long), cudaMemcpyHostToDevice, h->st));
    CK(cudaMalloc(&h->aux, (size_t)std::max<long long>(h->naux, 1) * 18 * sizeof(float)));
    CK(cudaMemsetAsync(h->aux, 0, (size_t)std::max<long long>(h->naux, 1) * 18 * sizeof(float), h->st));
    CK(cudaStreamSynchronize(h->st));
    h->absorber_ready = true;
    return 0;
}

extern "C" int swpc3d_setup_cerjan(swpc3d_handle *h, const float *gx_c, const float *gx_b, const float *gy_c, const float *gy_b,
                                   const float *gz_c, const float *gz_b) {
    if (!h) return fail("null handle");
    if (h->g.abc_type != SWPC3D_ABC_CERJAN) return fail("swpc3d_setup_cerjan: abc_type is not cerjan");
    if (!gx_c || !gx_b || !gy_c || !gy_b || !gz_c || !gz_b) return fail("swpc3d_setup_cerjan: null vector");
    CK(cudaSetDevice(h->dev));
    const float *src[6] = {gx_c, gx_b, gy_c, gy_b, gz_c, gz_b};
    for (int a = 0; a < 6; a++) {
        const bool isz = a >= 4;
        const int len_d = a < 2 ? h->NXM : (a < 4 ? h->NYM : h->NZP);
        std::vector<float> tmp((size_t)len_d, 1.0f);
        if (isz) {   // host vector covers k = -2 .. nz+3+kpad  ->  device index k + KOFF - 1
            for (int q = 0; q < h->nzm_h; q++) tmp[(size_t)(q - 2 + KOFF - 1)] = src[a][q];
        } else {
            for (int q = 0; q < len_d; q++) tmp[(size_t)q] = src[a][q];
        }
        if (!h->cg[a]) CK(cudaMalloc(&h->cg[a], (size_t)len_d * sizeof(float)));
        CK(cudaMemcpy(h->cg[a], tmp.data(), (size_t)len_d * sizeof(float), cudaMemcpyHostToDevice));
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
    if (!strcmp(s, "cosine")) return 4;
    if (!strcmp(s, "texp")) return 5;
    return 3;   // default branch of momentrate, m_fdtool.f90:494
}

extern "C" int swpc3d_set_sources(swpc3d_handle *h, int32_t nsrc, const int32_t *isrc, const int32_t *jsrc, const int32_t *ksrc,
                                  const double *mo, const double *mxx, const double *myy, const double *mzz, const double *myz,
                                  const double *mxz, const double *mxy, const float *srcprm, const char *stftype, int32_t bf_mode,
                                  float tbeg) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    cudaFree(h->src_ijk); cudaFree(h->src_mo); cudaFree(h->src_mij); cudaFree(h->src_prm); cudaFree(h->src_stime);
    h->src_ijk = nullptr; h->src_mo = nullptr; h->src_mij = nullptr; h->src_prm = nullptr; h->src_stime = nullptr;
    h->nsrc = nsrc; h->bf_mode = bf_mode; h->stf = stf_code(stftype); h->tbeg = tbeg;
    if (nsrc <= 0) return 0;
    std::vector<int> ijk(3 * (size_t)nsrc);
    std::vector<double> vmo((size_t)nsrc), mij(6 * (size_t)nsrc);
    for (int i = 0; i < nsrc; i++) {
        const int mi = isrc[i] - h->g.ibeg + HALO, mj = jsrc[i] - h->g.jbeg + HALO;
        // a source stencil that reaches an outer halo plane: the reference re-zeroes that plane at every exchange
        if (isrc[i] <= 1 || isrc[i] >= h->g.nx || jsrc[i] <= 1 || jsrc[i] >= h->g.ny) h->zero_outer = 1;
        // the reference keeps sources in the sleeve [ibeg-2, iend+3] (m_source.f90:209-211); the 4-node shear
        // stencil then touches mi-1 >= 0
        if (mi < 1 || mi >= h->NXM || mj < 1 || mj >= h->NYM || ksrc[i] < 1 - 2 + 1 || ksrc[i] > h->g.nz + 3)
            return fail("swpc3d_set_sources: source outside of the subdomain sleeve");
        ijk[3 * i] = mi; ijk[3 * i + 1] = mj; ijk[3 * i + 2] = ksrc[i];
        vmo[i] = mo ? mo[i] : 0.0;
        mij[6 * i] = mxx ? mxx[i] : 0; mij[6 * i + 1] = myy ? myy[i] : 0; mij[6 * i + 2] = mzz ? mzz[i] : 0;
        mij[6 * i + 3] = myz ? myz[i] : 0; mij[6 * i + 4] = mxz ? mxz[i] : 0; mij[6 * i + 5] = mxy ? mxy[i] : 0;
    }
    h->h_prm.assign(srcprm, srcprm + 2 * (size_t)nsrc);
    CK(cudaMalloc(&h->src_ijk, ijk.size() * sizeof(int)));
    CK(cudaMalloc(&h->src_mo, vmo.size() * sizeof(double)));
    CK(cudaMalloc(&h->src_mij, mij.size() * sizeof(double)));
    CK(cudaMalloc(&h->src_prm, h->h_prm.size() * sizeof(float)));
    CK(cudaMalloc(&h->src_stime, (size_t)swpc3d_handle::STIME_RING * 256 * sizeof(float)));
    if (!h->stime_pin) CK(cudaMallocHost(&h->stime_pin, (size_t)swpc3d_handle::STIME_RING * 256 * sizeof(float)));
    CK(cudaMemcpy(h->src_ijk, ijk.data(), ijk.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->src_mo, vmo.data(), vmo.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->src_mij, mij.data(), mij.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(h->src_prm, h->h_prm.data(), h->h_prm.size() * sizeof(float), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int swpc3d_set_stations(swpc3d_handle *h, int32_t nst, const int32_t *ist, const int32_t *jst, const int32_t *kst,
                                   int32_t ntdec_w, int32_t ntw, float M0, float UC) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    cudaFree(h->st_ijk); cudaFree(h->wav);
    h->st_ijk = nullptr; h->wav = nullptr;
    h->nst = nst; h->ntdec_w = ntdec_w; h->ntw = ntw; h->M0 = M0; h->UC = UC;
    if (nst <= 0 || ntw <= 0) return 0;
    std::vector<int> ijk(3 * (size_t)nst);
    for (int i = 0; i < nst; i++) {
        const int mi = ist[i] - h->g.ibeg + HALO, mj = jst[i] - h->g.jbeg + HALO;
        if (mi < HALO || mi >= HALO + h->nxp || mj < HALO || mj >= HALO + h->nyp || kst[i] < 1 || kst[i] > h->g.nz)
            return fail("swpc3d_set_stations: station outside of the owned box (m_wav.f90:203)");
        ijk[3 * i] = mi; ijk[3 * i + 1] = mj; ijk[3 * i + 2] = kst[i];
    }
    CK(cudaMalloc(&h->st_ijk, ijk.size() * sizeof(int)));
    CK(cudaMemcpy(h->st_ijk, ijk.data(), ijk.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->wav, (size_t)ntw * 3 * nst * sizeof(float)));
    CK(cudaMemset(h->wav, 0, (size_t)ntw * 3 * nst * sizeof(float)));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// sweeps
template <typename F, bool STRESS>
static int launch_direct_box(swpc3d_handle *h, const KParams<F> &p, const Box3 &b, cudaStream_t st = nullptr) {
    if (!st) st = main_stream(h);
    if (b.k1 < b.k0 || b.li1 < b.li0 || b.lj1 < b.lj0) return 0;
    dim3 blk((unsigned)h->tk, (unsigned)h->ti, 1);
    const int jlen = std::max(1, h->jlen);
    dim3 grd((unsigned)((b.k1 - b.k0 + 1 + h->tk - 1) / h->tk), (unsigned)((b.li1 - b.li0 + 1 + h->ti - 1) / h->ti),
             (unsigned)((b.lj1 - b.lj0 + 1 + jlen - 1) / jlen));
    if (b.flat) {   // threads numbered over the (k, i) cells of the box
        const long long cells = (long long)(b.k1 - b.k0 + 1) * (b.li1 - b.li0 + 1), per = (long long)h->tk * h->ti;
        grd.x = (unsigned)((cells + per - 1) / per);
        grd.y = 1;
    }
    switch (h->nm) {
    case 0: sweep_direct<F, 0, STRESS><<<grd, blk, 0, st>>>(p, b, jlen, h->pf); break;
    case 1: sweep_direct<F, 1, STRESS><<<grd, blk, 0, st>>>(p, b, jlen, h->pf); break;
    case 2: sweep_direct<F, 2, STRESS><<<grd, blk, 0, st>>>(p, b, jlen, h->pf); break;
    default: sweep_direct<F, 3, STRESS><<<grd, blk, 0, st>>>(p, b, jlen, h->pf); break;
    }
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

// TMA tensor maps over the contiguous field / memory-variable / medium allocations (4-D: k, i, j, array)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}
static CUtensorMapL2promotion l2promo_of(int v) {
    return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
}
static bool make_map(swpc3d_handle *h, CUtensorMap *m, void *base, int elem, int narr, int bk, int bi, int barr, int promo = 2) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)h->NZP, (cuuint64_t)h->NXM, (cuuint64_t)h->NYM, (cuuint64_t)narr};
    const cuuint64_t strides[3] = {(cuuint64_t)h->NZP * elem, (cuuint64_t)h->NZP * h->NXM * elem, (cuuint64_t)h->ncell * elem};
    const cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)bi, 1u, (cuuint32_t)barr};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapDataType dt = elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return enc(m, dt, 4, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               l2promo_of(promo), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename F, int NM>
static int tma_prepare(swpc3d_handle *h) {
    using C = TmaCfg<F, NM>;
    h->tma_ready = true;
    h->tma_ok = false;
    if (C::SMEM > 227 * 1024) return 0;
    const int pc = h->l2promo, ph = h->l2promo_halo;
    bool ok = make_map(h, &h->tmaps.S, h->Fall, sizeof(F), 9, C::TK, C::TI, 6, pc) && make_map(h, &h->tmaps.V, h->Fall, sizeof(F), 9, C::VK, C::VI, 3, ph) &&
              make_map(h, &h->tmaps.M, h->Mall, 4, 5, C::TK, C::TI, C::NMED, pc) && make_map(h, &h->tmaps.Mu, h->Mall, 4, 5, C::MUK, C::MUI, 1, ph);
    if (NM > 0) ok = ok && make_map(h, &h->tmaps.R, h->R, 4, 6 * NM, C::TK, C::TI, 6 * NM, pc);
    if (!ok) return 0;
    if (cudaFuncSetAttribute(stress_tma<F, NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(stress_tma_p<F, NM>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    h->tma_ok = true;
    using CV = TmaCfgVel<F>;
    h->vtma_ok = CV::SMEM <= 227 * 1024 && make_map(h, &h->vmaps.Vv, h->Fall, sizeof(F), 9, CV::TK, CV::TI, 3) &&
                 make_map(h, &h->vmaps.Sh, h->Fall, sizeof(F), 9, CV::SK, CV::SI_, 3) && make_map(h, &h->vmaps.Rho, h->Mall, 4, 5, CV::RK, CV::RI, 1);
    if (h->vtma_ok && cudaFuncSetAttribute(vel_tma<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV::SMEM) != cudaSuccess) {
        cudaGetLastError();
        h->vtma_ok = false;
    }
    return 0;
}

// A sweep covers a rectangle of owned columns (local 0-based, inclusive): the whole subdomain, or -- boundary-first
// overlap -- its core (everything but the two outermost planes towards each neighbour) and the boundary slabs.
struct Region { int li0, li1, lj0, lj1; };
static Region whole_region(const swpc3d_handle *h) { return Region{0, h->nxp - 1, 0, h->nyp - 1}; }
static bool face_split(const swpc3d_handle *h, int f) { return h->split_test || h->nbr[f] >= 0; }
// Width of the boundary slab towards face f.  The exchange carries the two outermost planes; the x slabs are made one tile
// column (8 columns) wide, so that the slab and what is left for the core are both swept by the tiled kernels at full
// efficiency (a 2-column slab is 0.4 % of the cells but cost 1 ms per step with sweep_direct next to the core sweep).
static int slab_width(const swpc3d_handle *h, int f) {
    if (f >= 2) return 2;
    return (h->slab_x >= 2 && h->nxp >= 4 * h->slab_x) ? h->slab_x : 2;
}
static Region core_region(const swpc3d_handle *h) {   // faces: 0 +x, 1 -x, 2 +y, 3 -y; planes sent: 2 per face (m_global.f90:416-443, 527-553)
    Region r = whole_region(h);
    if (face_split(h, 1)) r.li0 += slab_width(h, 1);
    if (face_split(h, 0)) r.li1 -= slab_width(h, 0);
    if (face_split(h, 3)) r.lj0 += slab_width(h, 3);
    if (face_split(h, 2)) r.lj1 -= slab_width(h, 2);
    return r;
}
// the boundary slabs: x slabs over every owned j, y slabs over the core's i range only (each cell exactly once)
static int boundary_boxes(const swpc3d_handle *h, Box3 out[4]) {
    const Region c = core_region(h);
    const int nz = h->g.nz;
    int n = 0;
    if (c.li0 > 0) out[n++] = Box3{1, nz, 0, c.li0 - 1, 0, h->nyp - 1, 0};
    if (c.li1 < h->nxp - 1) out[n++] = Box3{1, nz, c.li1 + 1, h->nxp - 1, 0, h->nyp - 1, 0};
    if (c.lj0 > 0) out[n++] = Box3{1, nz, c.li0, c.li1, 0, c.lj0 - 1, 0};
    if (c.lj1 < h->nyp - 1) out[n++] = Box3{1, nz, c.li0, c.li1, c.lj1 + 1, h->nyp - 1, 0};
    return n;
}

// The box of whole TK x TI tiles inside (interior kernel box) x (region) that stress_tma handles; empty if TMA is off.
template <typename F, int NM>
static Box3 tma_box(const swpc3d_handle *h, const Region &rg) {
    using C = TmaCfg<F, NM>;
    Box3 b{1, 0, 0, -1, 0, -1, 0};
    if (!h->use_tma || !h->tma_ok) return b;
    const swpc3d_grid &g = h->g;
    const int nkt = (g.kend_k + C::TK - 1) / C::TK;                     // interior k is 1..kend_k; the last tile may be partial
    const int li0 = std::max(g.ibeg_k - g.ibeg, rg.li0), li1 = std::min(g.iend_k - g.ibeg, rg.li1);
    const int lj0 = std::max(g.jbeg_k - g.jbeg, rg.lj0), lj1 = std::min(g.jend_k - g.jbeg, rg.lj1);
    if (nkt < 1 || li1 < li0 || lj1 < lj0) return b;
    // (the last k-tile may reach past the padded column and the last i-tile past the box: TMA zero-fills what is out of bounds,
    // and those lanes are masked)
    b.k0 = 1; b.k1 = nkt * C::TK; b.li0 = li0; b.li1 = li1; b.lj0 = lj0; b.lj1 = lj1;
    return b;
}


// ------------------------------------------------------------------------------------------------
// The absorber shell of a region, split into work items of pml_tma (tiles whose active cells are all PML cells and whose
// ADE columns are regularly spaced in the aux arrays) and the boxes that stay with sweep_direct (ragged interior columns,
// regions thinner than the pipeline is deep, rows that do not fit the bottom box).
// The kernel classes (columns x rows of a tile): 0 = wall slabs, 1 = the bottom rows (one box of 20 rows), 2 = wall slabs whose
// width is a multiple of 10 but not of 8 (the reference's default absorber is 20 columns thick)
constexpr int PML_NCLS = 3;
template <int CLS> struct PmlClass;
template <> struct PmlClass<0> { static constexpr int TI = 8, BK = 32; };
template <> struct PmlClass<1> { static constexpr int TI = 16, BK = 20; };
template <> struct PmlClass<2> { static constexpr int TI = 10, BK = 32; };
static const int PML_TI[PML_NCLS] = {8, 16, 10}, PML_BK[PML_NCLS] = {32, 20, 32};

struct PmlPlan {
    Region rg{};
    Box3 cols{};                         // the interior columns whose bottom rows this plan covers (stress: the TMA box; velocity: the kernel box)
    int n_items[PML_NCLS] = {};
    PmlItem *d_items[PML_NCLS] = {};
    unsigned int *d_ticket = nullptr;    // one counter per kernel class
    unsigned int base[PML_NCLS] = {};
    PmlMaps maps[PML_NCLS]{};
    std::vector<Box3> direct;
    int jl = 0, jlb = 0;
    bool skip_bottom = false;
};
static void pml_drop(swpc3d_handle *h) {
    if (h->pml[0].empty() && h->pml[1].empty()) return;
    cudaSetDevice(h->dev);
    cudaDeviceSynchronize();
    for (auto &list : h->pml) {
        for (PmlPlan *pl : list) {
            for (int c = 0; c < PML_NCLS; c++) cudaFree(pl->d_items[c]);
            cudaFree(pl->d_ticket);
            delete pl;
        }
        list.clear();
    }
}

static bool make_map_generic(CUtensorMap *m, void *base, int elem, const cuuint64_t dims[4], const cuuint64_t strides_bytes[3], const cuuint32_t box[4],
                             int promo) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const cuuint32_t es[4] = {1, 1, 1, 1};
    const CUtensorMapDataType dt = elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    return enc(m, dt, 4, base, dims, strides_bytes, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               l2promo_of(promo), CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// split `planes` into chunks of about `target` planes, none shorter than `minlen`; returns the chunk count (0: too thin)
static int even_chunks(int planes, int target, int minlen) {
    if (planes < minlen) return 0;
    int n = std::max(1, (planes + target / 2) / std::max(1, target));
    while (n > 1 && planes / n < minlen) n--;
    return n;
}

template <typename F, bool STRESS, int CLS>
static cudaError_t pml_set_smem() {
    using C = PmlCfg<F, STRESS, PmlClass<CLS>::TI, PmlClass<CLS>::BK>;
    return cudaFuncSetAttribute(pml_tma<F, STRESS, PmlClass<CLS>::TI, PmlClass<CLS>::BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
}
template <typename F, bool STRESS, int CLS>
static int pml_ns() { return PmlCfg<F, STRESS, PmlClass<CLS>::TI, PmlClass<CLS>::BK>::NS; }

template <typename F, bool STRESS>
static int pml_build(swpc3d_handle *h, const Region &rg, const Box3 &cols, bool skip_bottom, PmlPlan *pl) {
    const swpc3d_grid &g = h->g;
    const int nz = g.nz, nxp = h->nxp;
    pl->rg = rg; pl->cols = cols; pl->jl = h->pml_jl; pl->jlb = h->pml_jl_bottom; pl->skip_bottom = skip_bottom;
    pl->direct.clear();
    for (int c = 0; c < PML_NCLS; c++) pl->n_items[c] = 0;
    const int ki0 = std::max(g.ibeg_k - g.ibeg, rg.li0), ki1 = std::min(g.iend_k - g.ibeg, rg.li1);
    const int kj0 = std::max(g.jbeg_k - g.jbeg, rg.lj0), kj1 = std::min(g.jend_k - g.jbeg, rg.lj1);
    std::vector<PmlItem> items[PML_NCLS];
    auto aoff = [&](int li, int lj) { return h->h_aoff[(size_t)li + (size_t)nxp * lj]; };
    auto kba = [&](int li, int lj) { return h->h_kbeg_a[(size_t)(li + HALO) + (size_t)h->NXM * (lj + HALO)]; };
    const bool usable = h->use_pml && g.abc_type == SWPC3D_ABC_PML && get_encode() && !h->h_aoff.empty() && ki1 >= ki0 && kj1 >= kj0;
    const int nstage[PML_NCLS] = {pml_ns<F, STRESS, 0>(), pml_ns<F, STRESS, 1>(), pml_ns<F, STRESS, 2>()};
    int nmap[PML_NCLS] = {};
    // one region: columns [a0,a1] x planes [b0,b1], rows kb..nz (kb = 1: wall columns)
    auto add_region = [&](int a0, int a1, int b0, int b1, int kb, bool bottom, bool flat) {
        if (a1 < a0 || b1 < b0 || kb > nz) return;   // empty: nothing to do
        const Box3 whole{kb, nz, a0, a1, b0, b1, 0, flat ? 1 : 0};
        const int ncols = a1 - a0 + 1, nrows_j = b1 - b0 + 1;
        int cls = bottom ? 1 : 0;
        if (!bottom && (ncols + 9) / 10 * 10 < (ncols + 7) / 8 * 8) cls = 2;   // fewer idle columns with 10-wide tiles
        const int TI = PML_TI[cls], BK = PML_BK[cls], NS = nstage[cls];
        bool ok = usable && nmap[cls] < PML_NMAP;
        const int nch = ok ? even_chunks(nrows_j, bottom ? h->pml_jl_bottom : h->pml_jl, std::max(NS, 2)) : 0;
        if (nch < 1) ok = false;
        const int shift = (kb - 1) % 4;                  // the boxes start at kb - shift (16-byte aligned for float arrays)
        if (ok && bottom && shift + (nz - kb + 1) > BK) ok = false;
        long long asi = 0, asj = 0;
        const int phase = aux_lead(h, kb);
        const long long klen = aux_klen(h, kb);
        if (ok) {   // every column starts at kb and the columns are regularly spaced in the aux arrays
            asi = ncols > 1 ? aoff(a0 + 1, b0) - aoff(a0, b0) : klen;
            asj = nrows_j > 1 ? aoff(a0, b0 + 1) - aoff(a0, b0) : asi * ncols;
            for (int lj = b0; lj <= b1 && ok; lj++)
                for (int li = a0; li <= a1; li++)
                    if (kba(li, lj) != kb || aoff(li, lj) != aoff(a0, b0) + (long long)(li - a0) * asi + (long long)(lj - b0) * asj) { ok = false; break; }
            if (asi < klen || asj < asi * ncols || asi % 4 || asj % 4) ok = false;
        }
        if (!ok) { pl->direct.push_back(whole); return; }
        const int im = nmap[cls]++;
        {   // aux tensor of the region: (k within the column allocation, column, plane, array)
            const cuuint64_t dims[4] = {(cuuint64_t)klen, (cuuint64_t)ncols, (cuuint64_t)nrows_j, 18};
            const cuuint64_t st[3] = {(cuuint64_t)asi * 4, (cuuint64_t)asj * 4, (cuuint64_t)h->naux * 4};
            const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TI, 1, 9};
            if (!make_map_generic(&pl->maps[cls].aux[im], h->aux + (aoff(a0, b0) - phase), 4, dims, st, box,
                                  bottom ? (h->pml_promo_aux_b >= 0 ? h->pml_promo_aux_b : h->pml_promo_b) : h->pml_promo)) {
                nmap[cls]--;
                pl->direct.push_back(whole);
                return;
            }
        }
        // items: k-tile fastest, then column tile, then plane chunk -- blocks that run side by side work on neighbouring tiles
        const int nkt = bottom ? 1 : (nz + BK - 1) / BK;
        for (int c = 0; c < nch; c++) {
            const int j0 = b0 + (int)((long long)nrows_j * c / nch), j1 = b0 + (int)((long long)nrows_j * (c + 1) / nch) - 1;
            for (int it = 0; it * TI < ncols; it++)
                for (int kt = 0; kt < nkt; kt++) {
                    PmlItem I{};
                    if (bottom) { I.k0 = kb - shift; I.r0 = shift; I.r1 = shift + (nz - kb); I.ak = phase - shift; }
                    else { I.k0 = 1 + kt * BK; I.r0 = 0; I.r1 = std::min(BK - 1, nz - I.k0); I.ak = I.k0 - 1; }
                    I.li0 = a0 + it * TI; I.ncol = std::min(TI, a1 - I.li0 + 1);
                    I.lj0 = j0; I.nsteps = j1 - j0 + 1;
                    I.amap = im; I.ai = I.li0 - a0; I.aj = j0 - b0;
                    I.asi = (int)asi; I.asj = asj;
                    I.aux0 = aoff(a0, b0) + (I.k0 - kb) + (long long)I.ai * asi + (long long)I.aj * asj;
                    items[cls].push_back(I);
                }
        }
    };
    if (ki1 < ki0 || kj1 < kj0) {   // a subdomain entirely inside the absorber
        pl->direct.push_back(Box3{1, nz, rg.li0, rg.li1, rg.lj0, rg.lj1, 0, 0});
    } else {
        add_region(rg.li0, rg.li1, rg.lj0, kj0 - 1, 1, false, false);
        add_region(rg.li0, rg.li1, kj1 + 1, rg.lj1, 1, false, false);
        add_region(rg.li0, ki0 - 1, kj0, kj1, 1, false, h->flat_bottom != 0);
        add_region(ki1 + 1, rg.li1, kj0, kj1, 1, false, h->flat_bottom != 0);
        // interior columns that are not under `cols` (ragged tiles of the stress sweep): all their rows with sweep_direct
        if (cols.li0 > ki0) pl->direct.push_back(Box3{1, nz, ki0, cols.li0 - 1, kj0, kj1, 0, h->flat_bottom});
        if (cols.li1 < ki1) pl->direct.push_back(Box3{1, nz, cols.li1 + 1, ki1, kj0, kj1, 0, h->flat_bottom});
        if (cols.lj0 > kj0) pl->direct.push_back(Box3{1, nz, cols.li0, cols.li1, kj0, cols.lj0 - 1, 0, 0});
        if (cols.lj1 < kj1) pl->direct.push_back(Box3{1, nz, cols.li0, cols.li1, cols.lj1 + 1, kj1, 0, 0});
        if (!skip_bottom) add_region(cols.li0, cols.li1, cols.lj0, cols.lj1, g.kend_k + 1, true, h->flat_bottom != 0);
    }
    CK(cudaMalloc(&pl->d_ticket, PML_NCLS * sizeof(unsigned int)));
    CK(cudaMemset(pl->d_ticket, 0, PML_NCLS * sizeof(unsigned int)));
    // field / medium maps of the kernel classes in use
    for (int cls = 0; cls < PML_NCLS; cls++) {
        if (items[cls].empty()) continue;
        const int TI = PML_TI[cls], BK = PML_BK[cls], pr = cls == 1 ? h->pml_promo_b : h->pml_promo;
        PmlMaps &M = pl->maps[cls];
        const bool ok = make_map(h, &M.C, h->Fall, sizeof(F), 9, BK, TI, STRESS ? 6 : 3, pr) && make_map(h, &M.H, h->Fall, sizeof(F), 9, BK + 8, TI + 2, 3, pr) &&
                        make_map(h, &M.M1, h->Mall, 4, 5, BK, TI, 1, pr) && make_map(h, &M.Mh, h->Mall, 4, 5, BK + 4, TI + 1, 1, pr);
        const cudaError_t e = !ok ? cudaErrorUnknown : cls == 0 ? pml_set_smem<F, STRESS, 0>() : cls == 1 ? pml_set_smem<F, STRESS, 1>() : pml_set_smem<F, STRESS, 2>();
        if (!ok || e != cudaSuccess) {
            cudaGetLastError();
            return fail("pml_tma: cannot encode the tensor maps / set the shared-memory size");
        }
        CK(cudaMalloc(&pl->d_items[cls], items[cls].size() * sizeof(PmlItem)));
        CK(cudaMemcpy(pl->d_items[cls], items[cls].data(), items[cls].size() * sizeof(PmlItem), cudaMemcpyHostToDevice));
        pl->n_items[cls] = (int)items[cls].size();
    }
    return 0;
}

template <typename F, bool STRESS, int CLS>
static void pml_launch_class(const KParams<F> &p, PmlPlan *pl, const PmlGeom &gm, int grid, cudaStream_t st) {
    using C = PmlCfg<F, STRESS, PmlClass<CLS>::TI, PmlClass<CLS>::BK>;
    pml_tma<F, STRESS, PmlClass<CLS>::TI, PmlClass<CLS>::BK><<<grid, C::THREADS, C::SMEM, st>>>(p, pl->maps[CLS], pl->d_items[CLS], pl->d_ticket + CLS,
                                                                                                    pl->base[CLS], gm);
    pl->base[CLS] += (unsigned int)(gm.nitems + grid);   // every block takes one ticket past the end
}

// the absorber shell of region rg: pml_tma work lists + the boxes left to sweep_direct, all beside the interior kernel
template <typename F, bool STRESS>
static int launch_shell(swpc3d_handle *h, const KParams<F> &p, const Region &rg, const Box3 &cols, bool skip_bottom = false,
                        const std::function<int(cudaStream_t)> &extra = nullptr) {
    auto same = [](const Box3 &a, const Box3 &b) { return a.li0 == b.li0 && a.li1 == b.li1 && a.lj0 == b.lj0 && a.lj1 == b.lj1; };
    std::vector<PmlPlan *> &list = h->pml[STRESS ? 0 : 1];
    PmlPlan *pl = nullptr;
    for (size_t q = 0; q < list.size(); q++) {
        PmlPlan *c = list[q];
        if (!(c->rg.li0 == rg.li0 && c->rg.li1 == rg.li1 && c->rg.lj0 == rg.lj0 && c->rg.lj1 == rg.lj1)) continue;
        if (same(c->cols, cols) && c->jl == h->pml_jl && c->jlb == h->pml_jl_bottom && c->skip_bottom == skip_bottom) { pl = c; break; }
        CK(cudaDeviceSynchronize());   // the same region with other tiling options: rebuild
        for (int k = 0; k < PML_NCLS; k++) cudaFree(c->d_items[k]);
        cudaFree(c->d_ticket);
        delete c;
        list.erase(list.begin() + (long)q);
        break;
    }
    if (!pl) {
        pl = new PmlPlan();
        if (pml_build<F, STRESS>(h, rg, cols, skip_bottom, pl)) { delete pl; return 1; }
        list.push_back(pl);
    }
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->dev);
    int qs = 0;   // side stream round robin
    auto fork = [&](cudaStream_t &st) -> int {
        st = main_stream(h);
        if (side_streams(h)) {
            st = h->side[qs % 5];
            if (qs < 5) CK(cudaStreamWaitEvent(st, h->ev_fork, 0));
        }
        return 0;
    };
    auto join = [&](cudaStream_t st) -> int {
        if (side_streams(h)) {
            CK(cudaEventRecord(h->ev_join[qs % 5], st));
            CK(cudaStreamWaitEvent(h->st, h->ev_join[qs % 5], 0));
            qs++;
        }
        return 0;
    };
    PmlGeom gm{};
    if (STRESS) { gm.c_first = 3; gm.h_first = 0; gm.sa_first = 0; gm.m1_index = 2; gm.mh_index = 1; gm.a_first = 0; }
    else { gm.c_first = 0; gm.h_first = 6; gm.sa_first = 3; gm.m1_index = 2; gm.mh_index = 0; gm.a_first = 9; }
    if (extra) {   // the whole-line bottom tiles of this region (bottom_tma)
        cudaStream_t st;
        if (fork(st)) return 1;
        if (extra(st)) return 1;
        if (join(st)) return 1;
    }
    for (int cls = 0; cls < PML_NCLS; cls++) {
        if (!pl->n_items[cls]) continue;
        cudaStream_t st;
        if (fork(st)) return 1;
        gm.nitems = pl->n_items[cls];
        const int grid = std::min(gm.nitems, nsm);
        if (cls == 0) pml_launch_class<F, STRESS, 0>(p, pl, gm, grid, st);
        else if (cls == 1) pml_launch_class<F, STRESS, 1>(p, pl, gm, grid, st);
        else pml_launch_class<F, STRESS, 2>(p, pl, gm, grid, st);
        h->launches++;
        CK(cudaGetLastError());
        if (join(st)) return 1;
    }
    for (const Box3 &b : pl->direct) {
        if (b.k1 < b.k0 || b.li1 < b.li0 || b.lj1 < b.lj0) continue;
        cudaStream_t st;
        if (fork(st)) return 1;
        if (launch_direct_box<F, STRESS>(h, p, b, st)) return 1;
        if (join(st)) return 1;
    }
    return 0;
}

// Work items of stress_tma_p for the tile box t, in ticket order: k-tile fastest, then i-tile, then j-chunk of about tma_pl planes
struct TmaPlan {
    Box3 t{};
    int shift = 0, pl = 0, grid = 0, nitems = 0;
    TmaItem *d_items = nullptr;
    unsigned int *d_ticket = nullptr;
    unsigned int base = 0;     // tickets handed out by the launches so far (each launch takes nitems + grid)
};
static void tplan_drop(swpc3d_handle *h) {
    for (TmaPlan *&tp : h->tplan) {
        if (!tp) continue;
        cudaSetDevice(h->dev);
        cudaDeviceSynchronize();
        cudaFree(tp->d_items);
        cudaFree(tp->d_ticket);
        delete tp;
        tp = nullptr;
    }
}
static int tplan_build(swpc3d_handle *h, const Box3 &t, int TK, int TI, int blocks_per_sm, int k1_k, TmaPlan *tp) {
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->dev);
    nsm *= std::max(1, blocks_per_sm);
    const int nkt = (t.k1 - t.k0 + 1) / TK, nit = (t.li1 - t.li0 + TI) / TI, P = t.lj1 - t.lj0 + 1;
    const int nch = std::max(1, (P + h->tma_pl / 2) / std::max(1, h->tma_pl));
    std::vector<TmaItem> items;
    for (int c = 0; c < nch; c++) {
        const int j0 = t.lj0 + (int)((long long)P * c / nch), j1 = t.lj0 + (int)((long long)P * (c + 1) / nch) - 1;
        for (int it = 0; it < nit; it++)
            for (int kt = 0; kt < nkt; kt++) {
                TmaItem I{};
                const int k0t = 1 + kt * TK;
                // a last tile that would reach below the interior box is shifted up to the first k = 1 mod 4 that still covers kend_k
                I.k0 = (k0t + TK - 1 > k1_k && k1_k >= TK && h->tma_shift) ? ((k1_k - TK + 3) / 4) * 4 + 1 : k0t;
                I.kown = k0t;
                I.li0 = t.li0 + it * TI; I.lj0 = j0; I.nsteps = j1 - j0 + 1;
                items.push_back(I);
            }
    }
    tp->t = t; tp->shift = h->tma_shift; tp->pl = h->tma_pl; tp->nitems = (int)items.size();
    tp->grid = std::min(nsm, tp->nitems);
    tp->base = 0;
    CK(cudaMalloc(&tp->d_items, items.size() * sizeof(TmaItem)));
    CK(cudaMalloc(&tp->d_ticket, sizeof(unsigned int)));
    CK(cudaMemcpy(tp->d_items, items.data(), items.size() * sizeof(TmaItem), cudaMemcpyHostToDevice));
    CK(cudaMemset(tp->d_ticket, 0, sizeof(unsigned int)));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// bottom_tma: the rows k = nz-31 .. nz under the interior columns `cols` of a region, as whole 32-row tiles
struct BotPlan {
    Box3 cols{};
    int jl = 0, nitems = 0, grid = 0;
    BotItem *d_items = nullptr;
    unsigned int *d_ticket = nullptr;
    unsigned int base = 0;
    BotMaps maps{};
    BotGeom g{};
    bool ok = false;
};
static void bot_drop(swpc3d_handle *h) {
    if (h->bot[0].empty() && h->bot[1].empty()) return;
    cudaSetDevice(h->dev);
    cudaDeviceSynchronize();
    for (auto &list : h->bot) {
        for (BotPlan *b : list) { cudaFree(b->d_items); cudaFree(b->d_ticket); delete b; }
        list.clear();
    }
}
// can the bottom rows of this run be swept as whole-line tiles at all?  (geometry only; the per-region checks are in bot_build)
static bool bot_possible(const swpc3d_handle *h, bool stress) {
    const swpc3d_grid &g = h->g;
    if (!h->use_bot || !h->use_pml || g.abc_type != SWPC3D_ABC_PML || !get_encode() || h->h_aoff.empty()) return false;
    if (!h->use_tma || !h->tma_ok || (!stress && !h->use_ring)) return false;
    if (g.nz % 4 != 0 || g.nz < 64) return false;
    const int k0 = g.nz - 31, RI = g.kend_k - k0 + 1;
    if (RI < 1 || RI > 31 || g.kend_k != g.nz - g.na) return false;
    const int nint32 = (RI * 8 + 31) / 32 * 32, npml = (32 - RI) * 8;
    return (stress ? 2 : 1) * nint32 + npml <= 512;
}
template <typename F, int NM, bool STRESS>
static int bot_build(swpc3d_handle *h, const Box3 &cols, BotPlan *bp) {
    using C = BotCfg<F, NM, STRESS>;
    const swpc3d_grid &g = h->g;
    bp->cols = cols; bp->jl = h->bot_jl; bp->ok = false;
    const int nz = g.nz, nxp = h->nxp, kb = g.kend_k + 1;
    auto aoff = [&](int li, int lj) { return h->h_aoff[(size_t)li + (size_t)nxp * lj]; };
    auto kba = [&](int li, int lj) { return h->h_kbeg_a[(size_t)(li + HALO) + (size_t)h->NXM * (lj + HALO)]; };
    const int a0 = cols.li0, a1 = cols.li1, b0 = cols.lj0, b1 = cols.lj1;
    if (a1 < a0 || b1 < b0 || C::SMEM > 227 * 1024) return 0;
    const int ncols = a1 - a0 + 1, nrows_j = b1 - b0 + 1;
    const int phase = aux_lead(h, kb);
    const long long klen = aux_klen(h, kb);
    const long long asi = ncols > 1 ? aoff(a0 + 1, b0) - aoff(a0, b0) : klen;
    const long long asj = nrows_j > 1 ? aoff(a0, b0 + 1) - aoff(a0, b0) : asi * ncols;
    for (int lj = b0; lj <= b1; lj++)
        for (int li = a0; li <= a1; li++)
            if (kba(li, lj) != kb || aoff(li, lj) != aoff(a0, b0) + (long long)(li - a0) * asi + (long long)(lj - b0) * asj) return 0;
    if (asi < klen || asj < asi * ncols || asi % 4 || asj % 4) return 0;
    BotGeom &G = bp->g;
    G = BotGeom{};
    G.k0 = nz - 31; G.RI = g.kend_k - G.k0 + 1;
    G.RIB = (G.RI + 3) / 4 * 4; G.RIA = G.RI / 4 * 4; G.NA = 32 - G.RIA;
    G.ak = phase - (G.RI - G.RIA);
    if (G.ak < 0) return 0;
    G.asi = (int)asi; G.asj = asj;
    if (STRESS) { G.c_first = 3; G.h_first = 0; G.sa_first = 0; G.m_first = 2; G.mh_index = 1; G.a_first = 0; }
    else { G.c_first = 0; G.h_first = 6; G.sa_first = 3; G.m_first = 2; G.mh_index = 0; G.a_first = 9; }
    G.nint = G.RI * 8; G.nint32 = (G.nint + 31) / 32 * 32;
    // tensor maps
    const int pr = h->pml_promo;
    BotMaps &M = bp->maps;
    bool ok = make_map(h, &M.C, h->Fall, sizeof(F), 9, 32, 8, STRESS ? 6 : 3, pr) && make_map(h, &M.H, h->Fall, sizeof(F), 9, 40, 12, 3, pr) &&
              make_map(h, &M.Mh, h->Mall, 4, 5, 36, 9, 1, pr);
    if (STRESS) ok = ok && make_map(h, &M.M, h->Mall, 4, 5, 32, 8, C::NMED, pr);
    if (STRESS && NM > 0) ok = ok && make_map(h, &M.R, h->R, 4, 6 * NM, G.RIB, 8, 6 * NM, pr);
    {
        const cuuint64_t dims[4] = {(cuuint64_t)klen, (cuuint64_t)ncols, (cuuint64_t)nrows_j, 18};
        const cuuint64_t st[3] = {(cuuint64_t)asi * 4, (cuuint64_t)asj * 4, (cuuint64_t)h->naux * 4};
        const cuuint32_t box[4] = {(cuuint32_t)G.NA, 8, 1, 9};
        ok = ok && make_map_generic(&M.aux, h->aux + (aoff(a0, b0) - phase), 4, dims, st, box, pr);
    }
    if (!ok || cudaFuncSetAttribute(bottom_tma<F, NM, STRESS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    // items: tile column fastest, then runs of about bot_jl planes
    const int nch = std::max(1, (nrows_j + h->bot_jl / 2) / std::max(1, h->bot_jl));
    std::vector<BotItem> items;
    for (int c = 0; c < nch; c++) {
        const int j0 = b0 + (int)((long long)nrows_j * c / nch), j1 = b0 + (int)((long long)nrows_j * (c + 1) / nch) - 1;
        for (int it = 0; it * 8 < ncols; it++) {
            BotItem I{};
            I.li0 = a0 + it * 8; I.ncol = std::min(8, a1 - I.li0 + 1);
            I.lj0 = j0; I.nsteps = j1 - j0 + 1;
            I.ai = I.li0 - a0; I.aj = j0 - b0;
            I.aux0 = aoff(a0, b0) + (G.RIA - G.RI) + (long long)I.ai * asi + (long long)I.aj * asj;
            items.push_back(I);
        }
    }
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, h->dev);
    bp->nitems = (int)items.size();
    G.nitems = bp->nitems;
    bp->grid = std::min(nsm, bp->nitems);
    CK(cudaMalloc(&bp->d_items, items.size() * sizeof(BotItem)));
    CK(cudaMalloc(&bp->d_ticket, sizeof(unsigned int)));
    CK(cudaMemcpy(bp->d_items, items.data(), items.size() * sizeof(BotItem), cudaMemcpyHostToDevice));
    CK(cudaMemset(bp->d_ticket, 0, sizeof(unsigned int)));
    bp->ok = true;
    return 0;
}
// the plan of region columns `cols` (built on first use); nullptr when the bottom rows cannot be swept this way
template <typename F, int NM, bool STRESS>
static BotPlan *bot_plan(swpc3d_handle *h, const Box3 &cols) {
    if (!bot_possible(h, STRESS)) return nullptr;
    auto same = [](const Box3 &a, const Box3 &b) { return a.li0 == b.li0 && a.li1 == b.li1 && a.lj0 == b.lj0 && a.lj1 == b.lj1; };
    std::vector<BotPlan *> &list = h->bot[STRESS ? 0 : 1];
    for (BotPlan *b : list)
        if (same(b->cols, cols) && b->jl == h->bot_jl) return b->ok ? b : nullptr;
    BotPlan *bp = new BotPlan();
    if (bot_build<F, NM, STRESS>(h, cols, bp)) { bp->ok = false; cudaGetLastError(); }
    list.push_back(bp);
    return bp->ok ? bp : nullptr;
}
template <typename F, int NM, bool STRESS>
static int bot_launch(swpc3d_handle *h, const KParams<F> &p, BotPlan *bp, cudaStream_t st) {
    using C = BotCfg<F, NM, STRESS>;
    bottom_tma<F, NM, STRESS><<<bp->grid, C::THREADS, C::SMEM, st>>>(p, bp->maps, bp->d_items, bp->d_ticket, bp->base, bp->g);
    bp->base += (unsigned int)(bp->nitems + bp->grid);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <typename F, int NM>
static int launch_stress_nm(swpc3d_handle *h, const KParams<F> &p0, const Region &rg) {
    using C = TmaCfg<F, NM>;
    if (!h->tma_ready) tma_prepare<F, NM>(h);
    Box3 t = tma_box<F, NM>(h, rg);
    const Box3 all{1, h->g.nz, rg.li0, rg.li1, rg.lj0, rg.lj1, 0};
    if (t.k1 < t.k0) return launch_direct_box<F, true>(h, p0, all);
    // the bottom 32 rows of the interior columns as whole-line tiles (bottom_tma): the interior kernel then stops at k = nz - 32
    KParams<F> p = p0;
    BotPlan *bp = bot_plan<F, NM, true>(h, t);
    if (bp) {
        p.k1_k = h->g.nz - 32;
        t.k1 = (p.k1_k + C::TK - 1) / C::TK * C::TK;
    }
    TmaGeom g{};
    g.li0 = t.li0; g.li1 = t.li1; g.lj0 = t.lj0; g.lj1 = t.lj1; g.jl = std::max(1, h->tma_jl); g.m_first = 2; g.mu_index = 1;
    g.shift_last = h->tma_shift;
    g.l2hint = h->l2hint;
    dim3 grd((unsigned)((t.k1 - t.k0 + 1) / C::TK), (unsigned)((t.li1 - t.li0 + C::TI) / C::TI), (unsigned)((t.lj1 - t.lj0 + 1 + g.jl - 1) / g.jl));
    // complement of the TMA box inside the owned box: the absorber shell (four slabs of wall columns, the bottom rows under the
    // interior columns).  All launches touch disjoint cells and only read V, so they are issued on side streams next to the
    // interior kernel.
    if (side_streams(h)) CK(cudaEventRecord(h->ev_fork, h->st));
    if (h->tma_persist && !h->cur) {
        const Region w = whole_region(h);
        const int ri = (rg.li0 == w.li0 && rg.li1 == w.li1 && rg.lj0 == w.lj0 && rg.lj1 == w.lj1) ? 0 : 1;
        TmaPlan *&tp = h->tplan[ri];
        auto same = [](const Box3 &a, const Box3 &b) { return a.k1 == b.k1 && a.li0 == b.li0 && a.li1 == b.li1 && a.lj0 == b.lj0 && a.lj1 == b.lj1; };
        if (tp && !(same(tp->t, t) && tp->shift == h->tma_shift && tp->pl == h->tma_pl)) {
            CK(cudaDeviceSynchronize());
            cudaFree(tp->d_items); cudaFree(tp->d_ticket);
            delete tp;
            tp = nullptr;
        }
        if (!tp) {
            tp = new TmaPlan();
            int bps = 1;   // resident blocks per SM (2 for the elastic instantiation)
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, stress_tma_p<F, NM>, C::THREADS, C::SMEM) != cudaSuccess) { cudaGetLastError(); bps = 1; }
            if (tplan_build(h, t, C::TK, C::TI, bps, p.k1_k, tp)) return 1;
        }
        stress_tma_p<F, NM><<<tp->grid, C::THREADS, C::SMEM, main_stream(h)>>>(p, h->tmaps, tp->d_items, tp->nitems, tp->d_ticket, tp->base, g);
        tp->base += (unsigned int)(tp->nitems + tp->grid);
    } else {
        stress_tma<F, NM><<<grd, C::THREADS, C::SMEM, main_stream(h)>>>(p, h->tmaps, g);
    }
    h->launches++;
    CK(cudaGetLastError());
    if (bp) return launch_shell<F, true>(h, p0, rg, t, true, [&](cudaStream_t st) { return bot_launch<F, NM, true>(h, p0, bp, st); });
    return launch_shell<F, true>(h, p0, rg, t);
}
template <typename F, int NM>
static int launch_vel_nm(swpc3d_handle *h, const KParams<F> &p, const Region &rg) {
    using C = TmaCfg<F, NM>;
    using CV = TmaCfgVel<F>;
    if (!h->tma_ready) tma_prepare<F, NM>(h);
    const Box3 t = tma_box<F, NM>(h, rg);
    const Box3 all{1, h->g.nz, rg.li0, rg.li1, rg.lj0, rg.lj1, 0};
    const int ki0 = std::max(h->g.ibeg_k - h->g.ibeg, rg.li0), ki1 = std::min(h->g.iend_k - h->g.ibeg, rg.li1);
    const int kj0 = std::max(h->g.jbeg_k - h->g.jbeg, rg.lj0), kj1 = std::min(h->g.jend_k - h->g.jbeg, rg.lj1);
    if (h->use_ring && ki1 >= ki0 && kj1 >= kj0 && h->tk * h->ti <= 256) {
        // interior kernel box with the register-ring kernel; the absorber shell (all PML cells: two j slabs, two i slabs,
        // the bottom k slab) with the direct kernel on side streams.  Every launch writes disjoint cells and only reads S.
        const swpc3d_grid &gg = h->g;
        Box3 in{1, gg.kend_k, ki0, ki1, kj0, kj1, 0};
        BotPlan *bp = bot_plan<F, NM, false>(h, in);   // the bottom 32 rows as whole-line tiles: the ring kernel stops at k = nz - 32
        const Box3 cols = in;
        if (bp) in.k1 = gg.nz - 32;
        const int jlen = std::max(1, h->ring_jlen);
        dim3 blk((unsigned)h->tk, (unsigned)h->ti, 1);
        dim3 grd((unsigned)((in.k1 + h->tk - 1) / h->tk), (unsigned)((in.li1 - in.li0 + 1 + h->ti - 1) / h->ti), (unsigned)((in.lj1 - in.lj0 + 1 + jlen - 1) / jlen));
        if (side_streams(h)) CK(cudaEventRecord(h->ev_fork, h->st));
        bool pair = false;
        if ((sizeof(F) == 4 && h->ring_pair) || (sizeof(F) == 8 && h->ring_pair >= 2)) {   // two cells per thread, 64 / 128-bit loads
            pair = true;
            grd.x = (unsigned)((in.k1 + 2 * h->tk - 1) / (2 * h->tk));
            vel_ring2<F><<<grd, blk, 0, main_stream(h)>>>(p, in, jlen, h->ring_pf);
        }
        if (!pair) vel_ring<F><<<grd, blk, 0, main_stream(h)>>>(p, in, jlen, h->ring_pf);
        h->launches++;
        CK(cudaGetLastError());
        if (bp) return launch_shell<F, false>(h, p, rg, cols, true, [&](cudaStream_t st) { return bot_launch<F, NM, false>(h, p, bp, st); });
        return launch_shell<F, false>(h, p, rg, cols);
    }
    if (t.k1 < t.k0 || !h->vtma_ok || h->use_tma < 2) return launch_direct_box<F, false>(h, p, all);
    TmaGeom g{};
    g.li0 = t.li0; g.li1 = t.li1; g.lj0 = t.lj0; g.lj1 = t.lj1; g.jl = std::max(1, h->tma_jl);
    dim3 grd((unsigned)((t.k1 - t.k0 + 1) / CV::TK), (unsigned)((t.li1 - t.li0 + CV::TI) / CV::TI), (unsigned)((t.lj1 - t.lj0 + 1 + g.jl - 1) / g.jl));
    const int kE = (h->g.kend_k / C::TK) * C::TK + 1;
    const Box3 boxes[5] = {Box3{1, h->g.nz, rg.li0, rg.li1, rg.lj0, t.lj0 - 1, 0}, Box3{1, h->g.nz, rg.li0, rg.li1, t.lj1 + 1, rg.lj1, 0},
                           Box3{1, h->g.nz, rg.li0, t.li0 - 1, t.lj0, t.lj1, 0}, Box3{1, h->g.nz, t.li1 + 1, rg.li1, t.lj0, t.lj1, 0},
                           Box3{kE, h->g.nz, t.li0, t.li1, t.lj0, t.lj1, 1}};
    if (side_streams(h)) CK(cudaEventRecord(h->ev_fork, h->st));
    vel_tma<F><<<grd, CV::THREADS, CV::SMEM, main_stream(h)>>>(p, h->vmaps, g);
    h->launches++;
    CK(cudaGetLastError());
    for (int q = 0; q < 5; q++) {
        const Box3 &b = boxes[q];
        if (b.k1 < b.k0 || b.li1 < b.li0 || b.lj1 < b.lj0) continue;
        if (side_streams(h)) {
            CK(cudaStreamWaitEvent(h->side[q], h->ev_fork, 0));
            if (launch_direct_box<F, false>(h, p, b, h->side[q])) return 1;
            CK(cudaEventRecord(h->ev_join[q], h->side[q]));
            CK(cudaStreamWaitEvent(h->st, h->ev_join[q], 0));
        } else if (launch_direct_box<F, false>(h, p, b)) return 1;
    }
    return 0;
}

// part 0: the whole subdomain; boundary-first overlap: 3 = prologue on the launch stream (opens the timing bracket,
// plane-wave edges), 1 = the boundary slabs on the exchange stream, 2 = the core on the launch stream (closes the bracket)
template <typename F, bool STRESS>
static int launch_sweep(swpc3d_handle *h, int part = 0) {
    const KParams<F> p = make_params<F>(h);
    if (h->tk * h->ti > 256) return fail("tk*ti must be <= 256 (launch bounds)");
    const int w = STRESS ? 0 : 1;
    const bool timed = h->ktiming && h->kev_used[w] < 4096;
    if (part == 1) {
        Box3 bb[4];
        const int nb = boundary_boxes(h, bb);
        int rc = 0;
        if (h->slab_tiled) {   // every slab is a region of its own for the tiled kernels, on the exchange stream, one after the other
            h->cur = h->cs;
            for (int q = 0; q < nb && !rc; q++) {
                const Region r{bb[q].li0, bb[q].li1, bb[q].lj0, bb[q].lj1};
                if (STRESS) {
                    switch (h->nm) {
                    case 0: rc = launch_stress_nm<F, 0>(h, p, r); break;
                    case 1: rc = launch_stress_nm<F, 1>(h, p, r); break;
                    case 2: rc = launch_stress_nm<F, 2>(h, p, r); break;
                    default: rc = launch_stress_nm<F, 3>(h, p, r); break;
                    }
                } else {
                    switch (h->nm) {
                    case 0: rc = launch_vel_nm<F, 0>(h, p, r); break;
                    case 1: rc = launch_vel_nm<F, 1>(h, p, r); break;
                    case 2: rc = launch_vel_nm<F, 2>(h, p, r); break;
                    default: rc = launch_vel_nm<F, 3>(h, p, r); break;
                    }
                }
            }
            h->cur = nullptr;
            return rc;
        }
        // sweep_direct: x slabs 2 columns wide get a (128 k) x (2 i) block so that every lane is busy
        const int tk0 = h->tk, ti0 = h->ti;
        for (int q = 0; q < nb && !rc; q++) {
            if (bb[q].li1 - bb[q].li0 + 1 <= 2) { h->tk = 128; h->ti = 2; }
            rc = launch_direct_box<F, STRESS>(h, p, bb[q], h->cs);
            h->tk = tk0; h->ti = ti0;
        }
        return rc;
    }
    if (timed && part != 2) {
        if (h->kev_used[w] == h->kev[w][0].size()) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            h->kev[w][0].push_back(a);
            h->kev[w][1].push_back(b);
        }
        CK(cudaEventRecord(h->kev[w][0][h->kev_used[w]], h->st));
    }
    if (h->pw_mode && h->g.abc_type == SWPC3D_ABC_PML && part != 2) {   // absorb_p__update_stress extrapolates V, absorb_p__update_vel the stresses
        const swpc3d_grid &g = h->g;
        const int idx = g.myid % g.nproc_x, idy = g.myid / g.nproc_x;
        PwEdges e{};
        for (int q = 0; q < 4; q++) e.dst[q] = -1;
        const int mx = HALO + h->nxp, my = HALO + h->nyp;   // memory index of iend+1 / jend+1
        if (idx == 0) { e.dst[0] = HALO - 1; e.s1[0] = HALO; e.s2[0] = HALO + 1; e.nrow[0] = h->nyp; }
        if (idx == g.nproc_x - 1) { e.dst[1] = mx; e.s1[1] = mx - 1; e.s2[1] = mx - 2; e.nrow[1] = h->nyp; }
        if (idy == 0) { e.dst[2] = HALO - 1; e.s1[2] = HALO; e.s2[2] = HALO + 1; e.nrow[2] = h->nxp; }
        if (idy == g.nproc_y - 1) { e.dst[3] = my; e.s1[3] = my - 1; e.s2[3] = my - 2; e.nrow[3] = h->nxp; }
        dim3 blk(32, 8, 1), grd((unsigned)((g.nz + 31) / 32), (unsigned)((std::max(h->nxp, h->nyp) + 7) / 8), 4);
        F *base = (F *)h->Fall + (STRESS ? 0 : 3) * h->ncell;
        pw_edge_kernel<F><<<grd, blk, 0, h->st>>>(base, h->ncell, STRESS ? 3 : 6, g.nz, h->NZP, h->NXM, e);
        h->launches++;
        CK(cudaGetLastError());
    }
    int rc = 0;
    if (part == 3) return 0;
    const Region rg = part == 2 ? core_region(h) : whole_region(h);
    if (rg.li1 < rg.li0 || rg.lj1 < rg.lj0) rc = 0;
    else if (STRESS) {
        switch (h->nm) {
        case 0: rc = launch_stress_nm<F, 0>(h, p, rg); break;
        case 1: rc = launch_stress_nm<F, 1>(h, p, rg); break;
        case 2: rc = launch_stress_nm<F, 2>(h, p, rg); break;
        default: rc = launch_stress_nm<F, 3>(h, p, rg); break;
        }
    } else {
        switch (h->nm) {
        case 0: rc = launch_vel_nm<F, 0>(h, p, rg); break;
        case 1: rc = launch_vel_nm<F, 1>(h, p, rg); break;
        case 2: rc = launch_vel_nm<F, 2>(h, p, rg); break;
        default: rc = launch_vel_nm<F, 3>(h, p, rg); break;
        }
    }
    if (rc) return rc;
    if (timed) {
        CK(cudaEventRecord(h->kev[w][1][h->kev_used[w]], h->st));
        h->kev_used[w]++;
    }
    return 0;
}

static int ready(swpc3d_handle *h) {
    if (!h) return fail("null handle");
    if (!h->medium_ready) return fail("swpc3d: medium not uploaded (swpc3d_upload_medium)");
    if (!h->absorber_ready) return fail("swpc3d: absorber not set up (swpc3d_setup_pml / swpc3d_setup_cerjan)");
    CK(cudaSetDevice(h->dev));
    return 0;
}

extern "C" int swpc3d_update_stress(swpc3d_handle *h) {
    if (ready(h)) return 1;
    return h->fb == 8 ? launch_sweep<double, true>(h) : launch_sweep<float, true>(h);
}
extern "C" int swpc3d_update_vel(swpc3d_handle *h) {
    if (ready(h)) return 1;
    return h->fb == 8 ? launch_sweep<double, false>(h) : launch_sweep<float, false>(h);
}

// host evaluation of the moment-rate functions (same formulas as kernels.cuh / m_fdtool.f90:339-497) so that small
// source sets get a libm-evaluated value
static float momentrate_host(float t, int stf, float ts, float tr) {
    const double PI = 3.14159265358979323846;
    switch (stf) {
    case 0: return (ts <= t && t <= ts + tr) ? 1.0f / tr : 0.0f;
    case 1:
        if (ts <= t && t <= ts + tr / 2) return 4 * (t - ts) / (tr * tr);
        if (ts + tr / 2 < t && t <= ts + tr) return -4 * (t - ts - tr) / (tr * tr);
        return 0.0f;
    case 2: {
        const float t1 = ts + tr / 4, t2 = ts + 3 * tr / 4, tr3 = tr * tr * tr;
        if (ts <= t && t < t1) return 16 * ((t - ts) * (t - ts)) / tr3;
        if (t1 <= t && t < t2) return -2 * (8 * (t * t + tr * ts + ts * ts - t * tr - 2 * t * ts) + tr * tr) / tr3;
        if (t2 <= t && t <= ts + tr) return 16 * ((ts + tr - t) * (ts + tr - t)) / tr3;
        return 0.0f;
    }
    case 4: return (ts <= t && t <= ts + tr) ? (float)((1 - cos(2 * PI * (double)(t - ts) / (double)tr)) / (double)tr) : 0.0f;
    case 5:
        if (ts <= t) {
            const float tt = t - ts;
            return (float)((2 * PI) * (2 * PI) * (double)tt / (double)(tr * tr) * exp(-2 * PI * (double)tt / (double)tr));
        }
        return 0.0f;
    default:
        if (ts <= t && t <= ts + tr) {
            const double s = sin(PI * (double)(t - ts) / (double)tr);
            return (float)(3 * PI * (s * s * s) / (double)(4 * tr));
        }
        return 0.0f;
    }
}

// phase 0: upload of the host-evaluated moment rates + every target cell, on the launch stream; 3: that upload only;
// 1: targets outside the core box, on `stream`; 2: targets inside it (1 and 2 rely on an earlier phase-3 call)
template <typename F>
static int launch_source(swpc3d_handle *h, int it, bool body, int phase = 0, cudaStream_t stream = nullptr) {
    if (!stream) stream = h->st;
    SrcParams s{};
    s.phase = phase;
    const Region cr = core_region(h);
    s.c_li0 = cr.li0; s.c_li1 = cr.li1; s.c_lj0 = cr.lj0; s.c_lj1 = cr.lj1;
    s.nsrc = h->nsrc; s.ijk = h->src_ijk; s.mo = h->src_mo; s.mij = h->src_mij; s.prm = h->src_prm;
    s.stf = h->stf; s.dt_dxyz = h->dt_dxyz;
    const float dt = h->g.dt;
    s.t = body ? h->tbeg + it * dt : h->tbeg + ((float)it - 0.5f) * dt;   // m_source.f90:873 / :800
    s.stime = nullptr;
    s.nstv = 0;
    if (h->nsrc <= 16) {          // a handful of sources: the values ride in the kernel parameters
        s.nstv = h->nsrc;
        for (int i = 0; i < h->nsrc; i++) s.stv[i] = momentrate_host(s.t, h->stf, h->h_prm[2 * i], h->h_prm[2 * i + 1]);
    } else if (h->nsrc <= 256) {  // pinned ring -> device ring, asynchronous; a slot is reused STIME_RING steps later, once its copy is done
        if (phase == 0 || phase == 3) {
            const unsigned int slot = h->stime_n % swpc3d_handle::STIME_RING;
            cudaEvent_t &ev = h->stime_ev[slot];
            if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            else if (event_wait(h, ev)) return 1;
            float *st = h->stime_pin + (size_t)slot * 256;
            for (int i = 0; i < h->nsrc; i++) st[i] = momentrate_host(s.t, h->stf, h->h_prm[2 * i], h->h_prm[2 * i + 1]);
            h->stime_cur = h->src_stime + (size_t)slot * 256;
            CK(cudaMemcpyAsync(h->stime_cur, st, (size_t)h->nsrc * sizeof(float), cudaMemcpyHostToDevice, h->st));
            CK(cudaEventRecord(ev, h->st));
            h->stime_n++;
        }
        s.stime = h->stime_cur;
    }
    if (phase == 3) return 0;
    const KParams<F> p = make_params<F>(h);
    const int nb = (h->nsrc + 127) / 128;
    if (body) bodyforce_kernel<F><<<nb, 128, 0, stream>>>(p, s);
    else stressglut_kernel<F><<<nb, 128, 0, stream>>>(p, s);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int swpc3d_stressglut(swpc3d_handle *h, int32_t it) {
    if (ready(h)) return 1;
    if (h->bf_mode || h->nsrc <= 0) return 0;
    return h->fb == 8 ? launch_source<double>(h, it, false) : launch_source<float>(h, it, false);
}
extern "C" int swpc3d_bodyforce(swpc3d_handle *h, int32_t it) {
    if (ready(h)) return 1;
    if (!h->bf_mode || h->nsrc <= 0) return 0;
    return h->fb == 8 ? launch_source<double>(h, it, true) : launch_source<float>(h, it, true);
}

extern "C" int swpc3d_set_wav_products(swpc3d_handle *h, int32_t sw_v, int32_t sw_u, int32_t sw_stress, int32_t sw_strain) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    h->sw_v = sw_v; h->sw_u = sw_u; h->sw_stress = sw_stress; h->sw_strain = sw_strain;
    cudaFree(h->wav_u); cudaFree(h->wav_s); cudaFree(h->wav_e); cudaFree(h->wav_acc);
    h->wav_u = h->wav_s = h->wav_e = h->wav_acc = nullptr;
    if (h->nst <= 0 || h->ntw <= 0) return 0;
    const size_t n3 = (size_t)h->ntw * 3 * h->nst * sizeof(float), n6 = 2 * n3;
    if (sw_u) { CK(cudaMalloc(&h->wav_u, n3)); CK(cudaMemset(h->wav_u, 0, n3)); }
    if (sw_stress) { CK(cudaMalloc(&h->wav_s, n6)); CK(cudaMemset(h->wav_s, 0, n6)); }
    if (sw_strain) { CK(cudaMalloc(&h->wav_e, n6)); CK(cudaMemset(h->wav_e, 0, n6)); }
    CK(cudaMalloc(&h->wav_acc, (size_t)9 * h->nst * sizeof(float)));
    CK(cudaMemset(h->wav_acc, 0, (size_t)9 * h->nst * sizeof(float)));
    return 0;
}

extern "C" int swpc3d_wav_store(swpc3d_handle *h, int32_t it) {
    if (ready(h)) return 1;
    if (h->nst <= 0 || h->ntdec_w <= 0) return 0;
    const bool sample = (it - 1) % h->ntdec_w == 0 && (it - 1) / h->ntdec_w + 1 <= h->ntw;
    const bool accum = (h->sw_u || h->sw_strain) && h->wav_acc;
    if (!sample && !accum) return 0;
    WavParams w{};
    w.nst = h->nst; w.ntw = h->ntw; w.itw = (it - 1) / h->ntdec_w + 1; w.sample = sample ? 1 : 0;
    w.sw_v = h->sw_v && h->wav; w.sw_u = h->sw_u && h->wav_u; w.sw_stress = h->sw_stress && h->wav_s; w.sw_strain = h->sw_strain && h->wav_e;
    w.ijk = h->st_ijk; w.wav_v = h->wav; w.wav_u = h->wav_u; w.wav_s = h->wav_s; w.wav_e = h->wav_e; w.acc = h->wav_acc;
    w.M0 = h->M0; w.UC = h->UC;
    const double d[3] = {h->g.dx, h->g.dy, h->g.dz};
    for (int a = 0; a < 3; a++) {
        if (h->fb == 8) { w.r40[a] = 9.0 / 8.0 / d[a]; w.r41[a] = 1.0 / 24.0 / d[a]; }
        else { w.r40[a] = (double)(9.0f / 8.0f / (float)d[a]); w.r41[a] = (double)(1.0f / 24.0f / (float)d[a]); }
    }
    const int nb = (h->nst + 127) / 128;
    if (h->fb == 8) wav_store_kernel<double><<<nb, 128, 0, h->st>>>(make_params<double>(h), w);
    else wav_store_kernel<float><<<nb, 128, 0, h->st>>>(make_params<float>(h), w);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Green's-function mode (m_green.f90)
extern "C" int swpc3d_set_green(swpc3d_handle *h, int32_t ng, const int32_t *ig, const int32_t *jg, const int32_t *kg, int32_t bforce,
                                int32_t is_src, int32_t isrc, int32_t jsrc, int32_t ksrc, float fx1, float fy1, float fz1, float trise,
                                const char *stftype, int32_t ntdec_w, int32_t ntw, float tbeg) {
    if (!h) return fail("null handle");
    if (ntdec_w < 1 || ntw < 0 || ng < 0) return fail("swpc3d_set_green: bad ntdec_w / ntw / ng");
    CK(cudaSetDevice(h->dev));
    cudaFree(h->g_ijk); cudaFree(h->g_acc); cudaFree(h->g_gf);
    h->g_ijk = nullptr; h->g_acc = h->g_gf = nullptr;
    h->ng = ng; h->g_bforce = bforce != 0; h->g_ncmp = bforce ? 9 : 6; h->g_ntdec_w = ntdec_w; h->g_ntw = ntw;
    h->g_stf = stf_code(stftype); h->g_trise = trise; h->g_tbeg = tbeg;
    h->g_f[0] = fx1; h->g_f[1] = fy1; h->g_f[2] = fz1;
    h->g_dt_dxyz = (float)((double)h->g.dt / (h->g.dx * h->g.dy * h->g.dz));   // real(dt / (dx*dy*dz)), m_green.f90:156
    h->g_is_src = 0;
    if (is_src) {   // redefined is_src (:185-186): the pseudo source may sit one cell into the +x / +y halo
        const int mi = isrc - h->g.ibeg + HALO, mj = jsrc - h->g.jbeg + HALO;
        if (mi < HALO || mi > HALO + h->nxp || mj < HALO || mj > HALO + h->nyp || ksrc < 1 || ksrc > h->g.nz)
            return fail("swpc3d_set_green: pseudo source outside ibeg..iend+1 x jbeg..jend+1 x kbeg..kend (m_green.f90:185-186)");
        h->g_is_src = 1; h->g_src[0] = mi; h->g_src[1] = mj; h->g_src[2] = ksrc;
    }
    if (ng == 0 || ntw == 0) return 0;
    std::vector<int> ijk(3 * (size_t)ng);
    for (int i = 0; i < ng; i++) {
        const int mi = ig[i] - h->g.ibeg + HALO, mj = jg[i] - h->g.jbeg + HALO;
        if (mi < HALO || mi >= HALO + h->nxp || mj < HALO || mj >= HALO + h->nyp || kg[i] < 1 || kg[i] > h->g.nz)
            return fail("swpc3d_set_green: grid point outside of the owned box (m_green.f90:227-228)");
        ijk[3 * i] = mi; ijk[3 * i + 1] = mj; ijk[3 * i + 2] = kg[i];
    }
    CK(cudaMalloc(&h->g_ijk, ijk.size() * sizeof(int)));
    CK(cudaMemcpy(h->g_ijk, ijk.data(), ijk.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&h->g_acc, (size_t)12 * ng * sizeof(float)));
    CK(cudaMemset(h->g_acc, 0, (size_t)12 * ng * sizeof(float)));
    CK(cudaMalloc(&h->g_gf, (size_t)ntw * h->g_ncmp * ng * sizeof(float)));
    CK(cudaMemset(h->g_gf, 0, (size_t)ntw * h->g_ncmp * ng * sizeof(float)));
    return 0;
}

extern "C" int swpc3d_green_store(swpc3d_handle *h, int32_t it) {
    if (ready(h)) return 1;
    if (h->ng <= 0 || !h->g_gf) return 0;
    GreenParams g{};
    g.ng = h->ng; g.ncmp = h->g_ncmp; g.ntw = h->g_ntw; g.bforce = h->g_bforce;
    g.itw = (it - 1) / h->g_ntdec_w + 1;
    g.sample = ((it - 1) % h->g_ntdec_w == 0 && g.itw <= h->g_ntw) ? 1 : 0;
    g.ijk = h->g_ijk; g.acc = h->g_acc; g.gf = h->g_gf;
    const double d[3] = {h->g.dx, h->g.dy, h->g.dz};
    for (int a = 0; a < 3; a++) {
        if (h->fb == 8) { g.r40[a] = 9.0 / 8.0 / d[a]; g.r41[a] = 1.0 / 24.0 / d[a]; }
        else { g.r40[a] = (double)(9.0f / 8.0f / (float)d[a]); g.r41[a] = (double)(1.0f / 24.0f / (float)d[a]); }
    }
    const int nb = (h->ng + 127) / 128;
    if (h->fb == 8) green_store_kernel<double><<<nb, 128, 0, h->st>>>(make_params<double>(h), g);
    else green_store_kernel<float><<<nb, 128, 0, h->st>>>(make_params<float>(h), g);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int swpc3d_green_source(swpc3d_handle *h, int32_t it) {
    if (ready(h)) return 1;
    if (!h->g_is_src) return 0;
    const float stf = momentrate_host(h->g_tbeg + it * h->g.dt, h->g_stf, 0.0f, h->g_trise);   // green_tbeg = 0 (:62)
    const float fx = h->g_f[0] * h->g_dt_dxyz * stf, fy = h->g_f[1] * h->g_dt_dxyz * stf, fz = h->g_f[2] * h->g_dt_dxyz * stf;
    if (h->fb == 8) green_source_kernel<double><<<1, 32, 0, h->st>>>(make_params<double>(h), h->g_src[0], h->g_src[1], h->g_src[2], fx, fy, fz);
    else green_source_kernel<float><<<1, 32, 0, h->st>>>(make_params<float>(h), h->g_src[0], h->g_src[1], h->g_src[2], fx, fy, fz);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int swpc3d_get_green(swpc3d_handle *h, float *gf) {
    if (!h || !gf) return fail("null argument");
    if (h->ng <= 0 || !h->g_gf) return 0;
    CK(cudaSetDevice(h->dev));
    CK(cudaMemcpyAsync(gf, h->g_gf, (size_t)h->g_ntw * h->g_ncmp * h->ng * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int swpc3d_get_wav_product(swpc3d_handle *h, int32_t which, float *out) {
    if (!h || !out) return fail("null argument");
    const float *src = which == 0 ? h->wav : which == 1 ? h->wav_u : which == 2 ? h->wav_s : which == 3 ? h->wav_e : nullptr;
    if (!src) return fail("swpc3d_get_wav_product: product not enabled (swpc3d_set_wav_products)");
    CK(cudaSetDevice(h->dev));
    CK(cudaMemcpyAsync(out, src, (size_t)h->ntw * (which < 2 ? 3 : 6) * h->nst * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    return 0;
}

extern "C" int swpc3d_vmax(swpc3d_handle *h, float out[3]) {
    if (ready(h)) return 1;
    TRACE("vmax");
    const swpc3d_grid &g = h->g;
    const int margin = 5;
    const int i0 = std::max(g.na + margin + 1, g.ibeg_k), i1 = std::min(g.nx - g.na - margin, g.iend_k);
    const int j0 = std::max(g.na + margin + 1, g.jbeg_k), j1 = std::min(g.ny - g.na - margin, g.jend_k);
    out[0] = out[1] = out[2] = 0.0f;
    CK(cudaMemsetAsync(h->vmax_d, 0, 3 * sizeof(unsigned int), h->st));   // (also what swpc3d_vmax_global reduces for a rank without such cells)
    if (i1 < i0 || j1 < j0) return 0;
    const long long n = (long long)(i1 - i0 + 1) * (j1 - j0 + 1);
    const int nb = (int)std::min<long long>((n + 255) / 256, 1184);
    if (h->fb == 8) vmax_kernel<double><<<nb, 256, 0, h->st>>>(make_params<double>(h), i0 - g.ibeg, i1 - g.ibeg, j0 - g.jbeg, j1 - g.jbeg, h->vmax_d);
    else vmax_kernel<float><<<nb, 256, 0, h->st>>>(make_params<float>(h), i0 - g.ibeg, i1 - g.ibeg, j0 - g.jbeg, j1 - g.jbeg, h->vmax_d);
    h->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->vmax_h, h->vmax_d, 3 * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    WAIT(h, h->st);
    out[0] = h->vmax_h[0]; out[1] = h->vmax_h[1]; out[2] = h->vmax_h[2];
    return 0;
}

extern "C" int swpc3d_vmax_global(swpc3d_handle *h, float out[3]) {
    if (swpc3d_vmax(h, out)) return 1;
    if (!h->comm) return 0;
    // non-negative floats: the max of the values is the max of their bit patterns, reduce as float (vmax_d still holds this rank's maxima)
    NK(g_nccl.AllReduce(h->vmax_d, h->vmax_d, 3, ncclFloat, ncclMax, h->comm, h->st));
    CK(cudaMemcpyAsync(h->vmax_h, h->vmax_d, 3 * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    WAIT(h, h->st);
    out[0] = h->vmax_h[0]; out[1] = h->vmax_h[1]; out[2] = h->vmax_h[2];
    return 0;
}

extern "C" int swpc3d_get_wav(swpc3d_handle *h, float *wav_vel) {
    if (!h) return fail("null handle");
    if (h->nst <= 0 || h->ntw <= 0) return 0;
    CK(cudaSetDevice(h->dev));
    CK(cudaMemcpyAsync(wav_vel, h->wav, (size_t)h->ntw * 3 * h->nst * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    WAIT(h, h->st);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// snapshots (m_snap.f90)
static void snap_dims(const swpc3d_snap_cfg &c, int product, int &n1, int &n2, int &nvar) {
    const int sec = product / 3, typ = product % 3;
    n1 = sec == SEC_YZ ? c.nys : c.nxs;
    n2 = (sec == SEC_XZ || sec == SEC_YZ) ? c.nzs : c.nys;
    nvar = typ == TYP_PS ? 4 : 3;
}

extern "C" int swpc3d_snap_setup(swpc3d_handle *h, const swpc3d_snap_cfg *cfg) {
    if (!h || !cfg) return fail("null argument");
    CK(cudaSetDevice(h->dev));
    h->snap = *cfg;
    h->snap_on = false;
    for (int q = 0; q < 15; q++) {
        cudaFree(h->snap_buf[q]); cudaFree(h->snap_max[q]);
        h->snap_buf[q] = h->snap_max[q] = nullptr;
        if (!cfg->sw[q]) continue;
        int n1, n2, nvar;
        snap_dims(*cfg, q, n1, n2, nvar);
        const size_t bytes = (size_t)n1 * n2 * nvar * sizeof(float);
        CK(cudaMalloc(&h->snap_buf[q], bytes));
        CK(cudaMemset(h->snap_buf[q], 0, bytes));
        const int sec = q / 3, typ = q % 3;
        if ((sec == SEC_FS || sec == SEC_OB) && typ != TYP_PS) {
            CK(cudaMalloc(&h->snap_max[q], (size_t)n1 * n2 * 3 * sizeof(float)));
            CK(cudaMemset(h->snap_max[q], 0, (size_t)n1 * n2 * 3 * sizeof(float)));
        }
        h->snap_on = true;
    }
    return 0;
}

template <typename F>
static int snap_launch(swpc3d_handle *h, int product) {
    const swpc3d_snap_cfg &c = h->snap;
    const int sec = product / 3, typ = product % 3;
    // ranks that do not hold the section keep their zero buffer (m_snap.f90:1527 "if (idy /= idy_xz) return")
    if (sec == SEC_XZ && !(h->g.jbeg <= c.j0_xz && c.j0_xz <= h->g.jend)) return 0;
    if (sec == SEC_YZ && !(h->g.ibeg <= c.i0_yz && c.i0_yz <= h->g.iend)) return 0;
    SnapGeom g{};
    g.idec = c.idec; g.jdec = c.jdec; g.kdec = c.kdec; g.nxs = c.nxs; g.nys = c.nys; g.nzs = c.nzs;
    g.is0 = c.is0; g.is1 = c.is1; g.js0 = c.js0; g.js1 = c.js1; g.ks0 = c.ks0; g.ks1 = c.ks1;
    g.k0_xy = c.k0_xy; g.i0_yz = c.i0_yz; g.j0_xz = c.j0_xz; g.ibeg = h->g.ibeg; g.jbeg = h->g.jbeg;
    g.UC = c.UC; g.M0 = c.M0; g.kfs = h->kfs; g.kob = h->kob;
    const int a0 = sec == SEC_YZ ? c.js0 : c.is0, a1 = sec == SEC_YZ ? c.js1 : c.is1;
    const int b0 = (sec == SEC_XZ || sec == SEC_YZ) ? c.ks0 : c.js0, b1 = (sec == SEC_XZ || sec == SEC_YZ) ? c.ks1 : c.js1;
    if (a1 < a0 || b1 < b0) return 0;
    dim3 blk(32, 8, 1), grd((unsigned)((a1 - a0 + 1 + 31) / 32), (unsigned)((b1 - b0 + 1 + 7) / 8), 1);
    snap_kernel<F><<<grd, blk, 0, h->st>>>(make_params<F>(h), g, sec, typ, h->snap_buf[product], h->snap_max[product]);
    h->launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int swpc3d_snap_step(swpc3d_handle *h, int32_t it) {
    if (ready(h)) return 1;
    if (!h->snap_on) return 0;
    const bool out = h->snap.ntdec_s > 0 && (it - 1) % h->snap.ntdec_s == 0;
    for (int q = 0; q < 15; q++) {
        if (!h->snap_buf[q]) continue;
        const int sec = q / 3, typ = q % 3;
        const bool every = typ == TYP_U || (typ == TYP_V && (sec == SEC_FS || sec == SEC_OB));
        if (!(every || out)) continue;
        if (h->fb == 8 ? snap_launch<double>(h, q) : snap_launch<float>(h, q)) return 1;
    }
    return 0;
}

static int snap_fetch_impl(swpc3d_handle *h, const float *src, size_t n, int root, float *out) {
    if (!src) return fail("swpc3d_snap_fetch: product not enabled");
    CK(cudaSetDevice(h->dev));
    const float *from = src;
    if (h->comm) {
        if (h->snap_tmp_n < n) {
            cudaFree(h->snap_tmp);
            h->snap_tmp = nullptr; h->snap_tmp_n = 0;
            CK(cudaMalloc(&h->snap_tmp, n * sizeof(float)));
            h->snap_tmp_n = n;
        }
        NK(g_nccl.Reduce(src, h->snap_tmp, n, ncclFloat, ncclSum, root, h->comm, h->st));
        from = h->snap_tmp;
    }
    if (out && (!h->comm || h->g.myid == root)) CK(cudaMemcpyAsync(out, from, n * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    WAIT(h, h->st);
    return 0;
}

extern "C" int swpc3d_snap_fetch(swpc3d_handle *h, int32_t product, int32_t root, float *out) {
    if (!h || product < 0 || product >= 15) return fail("swpc3d_snap_fetch: bad argument");
    int n1, n2, nvar;
    snap_dims(h->snap, product, n1, n2, nvar);
    return snap_fetch_impl(h, h->snap_buf[product], (size_t)n1 * n2 * nvar, root, out);
}
extern "C" int swpc3d_snap_fetch_max(swpc3d_handle *h, int32_t product, int32_t root, float *out) {
    if (!h || product < 0 || product >= 15) return fail("swpc3d_snap_fetch_max: bad argument");
    int n1, n2, nvar;
    snap_dims(h->snap, product, n1, n2, nvar);
    return snap_fetch_impl(h, h->snap_max[product], (size_t)n1 * n2 * 3, root, out);
}
// The asynchronous pair (m_snap.f90:1057-1064: `mpi_wait` of the previous record / `mpi_ireduce` of the current one): _begin
// copies the slice buffer aside on the launch stream (device to device, microseconds) and lets a third stream reduce it onto
// the root and copy it into pinned host memory [slot]; _end waits for that copy and returns the host pointer.  The sweeps go on
// meanwhile; a slot may be reused once its _end has returned.
extern "C" int swpc3d_snap_fetch_begin(swpc3d_handle *h, int32_t product, int32_t root, int32_t slot) {
    if (!h || product < 0 || product >= 15 || slot < 0 || slot > 1) return fail("swpc3d_snap_fetch_begin: bad argument");
    if (!h->snap_buf[product]) return fail("swpc3d_snap_fetch_begin: product not enabled");
    CK(cudaSetDevice(h->dev));
    const int q = product;
    int n1, n2, nvar;
    snap_dims(h->snap, q, n1, n2, nvar);
    const size_t n = (size_t)n1 * n2 * nvar, bytes = n * sizeof(float);
    const bool is_root = !h->comm || h->g.myid == root;
    if (!h->ss) CK(cudaStreamCreateWithFlags(&h->ss, cudaStreamNonBlocking));
    if (!h->snap_stage[q]) {
        CK(cudaMalloc(&h->snap_stage[q], bytes));
        CK(cudaEventCreateWithFlags(&h->snap_ev_staged[q], cudaEventDisableTiming));
        for (int b = 0; b < 2; b++) CK(cudaEventCreateWithFlags(&h->snap_ev_done[q][b], cudaEventDisableTiming));
    }
    if (is_root && !h->snap_pin[q][slot]) CK(cudaMallocHost(&h->snap_pin[q][slot], bytes));
    if (h->comm && is_root && !h->snap_red[q]) CK(cudaMalloc(&h->snap_red[q], bytes));
    // the staging copy is read by the previous record's reduce / copy on `ss`: overwrite it only after those
    if (h->snap_inflight[q]) CK(cudaStreamWaitEvent(h->st, h->snap_ev_done[q][h->snap_last_slot[q]], 0));
    CK(cudaMemcpyAsync(h->snap_stage[q], h->snap_buf[q], bytes, cudaMemcpyDeviceToDevice, h->st));
    CK(cudaEventRecord(h->snap_ev_staged[q], h->st));
    CK(cudaStreamWaitEvent(h->ss, h->snap_ev_staged[q], 0));
    const float *from = h->snap_stage[q];
    if (h->comm) {
        ncclComm_t c = h->comm_io ? h->comm_io : h->comm;
        NK(g_nccl.Reduce(h->snap_stage[q], is_root ? h->snap_red[q] : h->snap_stage[q], n, ncclFloat, ncclSum, root, c, h->ss));
        from = h->snap_red[q];
    }
    if (is_root) CK(cudaMemcpyAsync(h->snap_pin[q][slot], from, bytes, cudaMemcpyDeviceToHost, h->ss));
    CK(cudaEventRecord(h->snap_ev_done[q][slot], h->ss));
    h->snap_inflight[q] = true;
    h->snap_last_slot[q] = slot;
    return 0;
}
extern "C" int swpc3d_snap_fetch_end(swpc3d_handle *h, int32_t product, int32_t slot, const float **data) {
    if (!h || product < 0 || product >= 15 || slot < 0 || slot > 1) return fail("swpc3d_snap_fetch_end: bad argument");
    if (!h->snap_ev_done[product][slot]) return fail("swpc3d_snap_fetch_end: no fetch was begun for this product");
    CK(cudaSetDevice(h->dev));
    if (event_wait(h, h->snap_ev_done[product][slot])) return 1;
    if (data) *data = h->snap_pin[product][slot];   // NULL on the ranks that are not the root of this product
    return 0;
}
extern "C" int swpc3d_reduce_sum(swpc3d_handle *h, float *buf, int64_t n, int32_t root) {
    if (!h || !buf || n < 0) return fail("swpc3d_reduce_sum: bad argument");
    if (!h->comm || n == 0) return 0;
    CK(cudaSetDevice(h->dev));
    if (h->snap_tmp_n < (size_t)n) {
        cudaFree(h->snap_tmp);
        h->snap_tmp = nullptr; h->snap_tmp_n = 0;
        CK(cudaMalloc(&h->snap_tmp, (size_t)n * sizeof(float)));
        h->snap_tmp_n = (size_t)n;
    }
    CK(cudaMemcpyAsync(h->snap_tmp, buf, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, h->st));
    NK(g_nccl.Reduce(h->snap_tmp, h->snap_tmp, (size_t)n, ncclFloat, ncclSum, root, h->comm, h->st));
    if (h->g.myid == root) CK(cudaMemcpyAsync(buf, h->snap_tmp, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, h->st));
    WAIT(h, h->st);
    return 0;
}

extern "C" int swpc3d_sync(swpc3d_handle *h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    WAIT(h, h->st);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// halo exchange.  Plane lists of m_global.f90:416-443/461-486 (velocity) and :527-553/573-599 (stress);
// mi/mj are memory-box indices: ibeg <-> HALO, iend <-> HALO+nxp-1.
struct FaceLists { PlaneList send[4], recv[4]; };

static FaceLists face_lists(const swpc3d_handle *h, int which) {
    FaceLists L{};
    const int ib = HALO, ie = HALO + h->nxp - 1, jb = HALO, je = HALO + h->nyp - 1;
    void *const *F = h->F;
    enum { Vx = 0, Vy, Vz, Sxx, Syy, Szz, Syz, Sxz, Sxy };
    auto set = [](PlaneList &pl, std::initializer_list<std::pair<void *, int>> v) {
        pl.n = 0;
        for (auto &e : v) { pl.field[pl.n] = e.first; pl.m[pl.n] = e.second; pl.n++; }
    };
    if (which == 1) {   // velocity
        set(L.send[0], {{F[Vx], ie - 1}, {F[Vx], ie}, {F[Vy], ie}, {F[Vz], ie}});
        set(L.send[1], {{F[Vx], ib}, {F[Vy], ib}, {F[Vy], ib + 1}, {F[Vz], ib}, {F[Vz], ib + 1}});
        set(L.send[2], {{F[Vx], je}, {F[Vy], je - 1}, {F[Vy], je}, {F[Vz], je}});
        set(L.send[3], {{F[Vx], jb}, {F[Vx], jb + 1}, {F[Vy], jb}, {F[Vz], jb}, {F[Vz], jb + 1}});
        set(L.recv[1], {{F[Vx], ib - 2}, {F[Vx], ib - 1}, {F[Vy], ib - 1}, {F[Vz], ib - 1}});                      // from -x
        set(L.recv[0], {{F[Vx], ie + 1}, {F[Vy], ie + 1}, {F[Vy], ie + 2}, {F[Vz], ie + 1}, {F[Vz], ie + 2}});     // from +x
        set(L.recv[3], {{F[Vx], jb - 1}, {F[Vy], jb - 2}, {F[Vy], jb - 1}, {F[Vz], jb - 1}});                      // from -y
        set(L.recv[2], {{F[Vx], je + 1}, {F[Vx], je + 2}, {F[Vy], je + 1}, {F[Vz], je + 1}, {F[Vz], je + 2}});     // from +y
    } else {            // stress
        set(L.send[0], {{F[Sxx], ie}, {F[Sxy], ie - 1}, {F[Sxy], ie}, {F[Sxz], ie - 1}, {F[Sxz], ie}});
        set(L.send[1], {{F[Sxx], ib}, {F[Sxx], ib + 1}, {F[Sxy], ib}, {F[Sxz], ib}});
        set(L.send[2], {{F[Syy], je}, {F[Sxy], je - 1}, {F[Sxy], je}, {F[Syz], je - 1}, {F[Syz], je}});
        set(L.send[3], {{F[Syy], jb}, {F[Syy], jb + 1}, {F[Sxy], jb}, {F[Syz], jb}});
        set(L.recv[1], {{F[Sxx], ib - 1}, {F[Sxy], ib - 2}, {F[Sxy], ib - 1}, {F[Sxz], ib - 2}, {F[Sxz], ib - 1}});
        set(L.recv[0], {{F[Sxx], ie + 1}, {F[Sxx], ie + 2}, {F[Sxy], ie + 1}, {F[Sxz], ie + 1}});
        set(L.recv[3], {{F[Syy], jb - 1}, {F[Sxy], jb - 2}, {F[Sxy], jb - 1}, {F[Syz], jb - 2}, {F[Syz], jb - 1}});
        set(L.recv[2], {{F[Syy], je + 1}, {F[Syy], je + 2}, {F[Sxy], je + 1}, {F[Syz], je + 1}});
    }
    return L;
}

template <typename F>
static int launch_halo(swpc3d_handle *h, const FaceLists &L, bool pack, cudaStream_t st_, bool outer_only = false) {
    const int nz = h->g.nz;
    for (int f = 0; f < 4; f++) {
        if (outer_only && h->nbr[f] >= 0) continue;
        // A face without a neighbour: the reference still unpacks its (all-zero) receive buffer into the outer halo
        // planes (m_global.f90:458-488).  Those planes are zero anyway unless something else wrote them:
        // the plane-wave initial condition / edge extrapolation, or a source stencil at the model edge.
        const bool outer = h->nbr[f] < 0;
        if (outer && (pack || !h->zero_outer)) continue;
        const bool xface = f < 2;
        const int nline = xface ? h->nyp : h->nxp;
        const PlaneList &pl = pack ? L.send[f] : L.recv[f];
        dim3 blk(128, 1, 1), grd((unsigned)((nz + 127) / 128), (unsigned)nline, (unsigned)pl.n);
        F *buf = outer ? nullptr : (F *)(pack ? h->sbuf[f] : h->rbuf[f]);
        if (xface) {
            if (pack) halo_kernel<F, true, true><<<grd, blk, 0, st_>>>(nz, nline, h->NZP, h->NXM, pl, buf);
            else halo_kernel<F, true, false><<<grd, blk, 0, st_>>>(nz, nline, h->NZP, h->NXM, pl, buf);
        } else {
            if (pack) halo_kernel<F, false, true><<<grd, blk, 0, st_>>>(nz, nline, h->NZP, h->NXM, pl, buf);
            else halo_kernel<F, false, false><<<grd, blk, 0, st_>>>(nz, nline, h->NZP, h->NXM, pl, buf);
        }
        h->launches++;
        CK(cudaGetLastError());
    }
    return 0;
}

static size_t face_count(const swpc3d_handle *h, const PlaneList &pl, int f) {
    return (size_t)pl.n * (size_t)(f < 2 ? h->nyp : h->nxp) * (size_t)h->g.nz;
}

static bool has_neighbour(const swpc3d_handle *h) {
    for (int f = 0; f < 4; f++)
        if (h->nbr[f] >= 0) return true;
    return false;
}

// ---- peer-to-peer form of the exchange.  What the neighbours tell each other once, at swpc3d_comm_init (over NCCL, 200 bytes per
// face): the CUDA IPC handle of their block and where in it the messages from this rank go.
struct P2pInfo {
    cudaIpcMemHandle_t handle;
    unsigned long long off[2][2];   // [family][parity]: receive buffer for the messages that cross this face
    unsigned long long flag[2];     // [family]
    unsigned long long ok;
};
static int p2p_setup(swpc3d_handle *h) {
    h->p2p_ok = false;
    const bool any = has_neighbour(h);
    int ok = 1;
    // my block
    size_t off = 0;
    const size_t isz = (size_t)5 * h->nyp * h->g.nz * h->fb, jsz = (size_t)5 * h->nxp * h->g.nz * h->fb;
    for (int fam = 0; fam < 2; fam++)
        for (int par = 0; par < 2; par++)
            for (int f = 0; f < 4; f++) {
                h->p2p_off[fam][par][f] = off;
                if (h->nbr[f] >= 0) off += ((f < 2 ? isz : jsz) + 255) / 256 * 256;
            }
    h->p2p_flag_off = off; off += 256;
    h->p2p_count_off = off; off += 256;
    h->p2p_bytes = off;
    if (cudaMalloc(&h->p2p_base, off) != cudaSuccess) { cudaGetLastError(); h->p2p_base = nullptr; ok = 0; }
    if (h->p2p_base) CK(cudaMemset(h->p2p_base, 0, off));
    P2pInfo mine[4], theirs[4];
    memset(mine, 0, sizeof(mine));
    memset(theirs, 0, sizeof(theirs));
    const int opp[4] = {1, 0, 3, 2};
    cudaIpcMemHandle_t hd;
    memset(&hd, 0, sizeof(hd));
    if (ok && cudaIpcGetMemHandle(&hd, h->p2p_base) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    for (int f = 0; f < 4; f++) {   // what the neighbour across face f needs: my buffers / flags for the messages that arrive through face f
        mine[f].handle = hd;
        for (int fam = 0; fam < 2; fam++) {
            for (int par = 0; par < 2; par++) mine[f].off[fam][par] = h->p2p_off[fam][par][f];
            mine[f].flag[fam] = h->p2p_flag_off + (size_t)(fam * 4 + f) * sizeof(unsigned int);
        }
        mine[f].ok = (unsigned long long)ok;
    }
    (void)opp;
    if (any) {
        P2pInfo *d = nullptr;
        CK(cudaMalloc(&d, 8 * sizeof(P2pInfo)));
        CK(cudaMemcpy(d, mine, sizeof(mine), cudaMemcpyHostToDevice));
        NK(g_nccl.GroupStart());
        for (int f = 0; f < 4; f++) {
            if (h->nbr[f] < 0) continue;
            NK(g_nccl.Send(d + f, sizeof(P2pInfo), ncclChar, h->nbr[f], h->comm, h->st));
            NK(g_nccl.Recv(d + 4 + f, sizeof(P2pInfo), ncclChar, h->nbr[f], h->comm, h->st));
        }
        NK(g_nccl.GroupEnd());
        CK(cudaMemcpyAsync(theirs, d + 4, sizeof(theirs), cudaMemcpyDeviceToHost, h->st));
        CK(cudaStreamSynchronize(h->st));
        cudaFree(d);
        for (int f = 0; f < 4 && ok; f++) {
            if (h->nbr[f] < 0) continue;
            if (!theirs[f].ok) { ok = 0; break; }
            // (a neighbour on two faces would be the same process on both: open its block once)
            void *ptr = nullptr;
            for (int e = 0; e < f; e++)
                if (h->nbr[e] == h->nbr[f]) ptr = h->p2p_peer[e];
            if (!ptr && cudaIpcOpenMemHandle(&ptr, theirs[f].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            h->p2p_peer[f] = ptr;
            for (int fam = 0; fam < 2; fam++) {
                for (int par = 0; par < 2; par++) h->p2p_peer_off[f][fam][par] = (size_t)theirs[f].off[fam][par];
                h->p2p_peer_flag[f][fam] = (size_t)theirs[f].flag[fam];
            }
        }
    }
    // every rank takes the same path: peer-to-peer only if it works everywhere
    int *dok = nullptr;
    CK(cudaMalloc(&dok, sizeof(int)));
    CK(cudaMemcpy(dok, &ok, sizeof(int), cudaMemcpyHostToDevice));
    NK(g_nccl.AllReduce(dok, dok, 1, ncclInt, ncclMin, h->comm, h->st));
    CK(cudaMemcpyAsync(&ok, dok, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CK(cudaStreamSynchronize(h->st));
    cudaFree(dok);
    h->p2p_ok = ok != 0;
    return 0;
}

// one exchange, peer to peer: ONE push launch for all faces, a one-warp wait, ONE pull launch
template <typename F>
static int p2p_exchange(swpc3d_handle *h, const FaceLists &L, int which, cudaStream_t st_) {
    const int nz = h->g.nz, fam = which;
    const unsigned int seq = ++h->p2p_seq[fam];
    const int par = (int)(seq & 1u);
    constexpr int VEC = 16 / (int)sizeof(F);
    P2pFaces push{}, pull{};
    int z = 0, maxline = 0;
    const unsigned int *fl[4];
    for (int f = 0; f < 4; f++) {
        const bool on = h->nbr[f] >= 0;
        const int nline = on ? (f < 2 ? h->nyp : h->nxp) : 0;
        push.pl[f] = L.send[f]; pull.pl[f] = L.recv[f];
        if (!on) { push.pl[f].n = 0; pull.pl[f].n = 0; }
        push.nline[f] = pull.nline[f] = nline;
        z += on ? L.send[f].n : 0;      // (send and receive lists of a face pair up: 4 planes one way, 5 the other -- see below)
        push.zend[f] = z;
        maxline = std::max(maxline, nline);
        push.buf[f] = on ? (void *)((char *)h->p2p_peer[f] + h->p2p_peer_off[f][fam][par]) : nullptr;
        push.flag[f] = on ? (unsigned int *)((char *)h->p2p_peer[f] + h->p2p_peer_flag[f][fam]) : nullptr;
        push.count[f] = (unsigned int *)(h->p2p_base + h->p2p_count_off) + (fam * 4 + f);
        pull.buf[f] = on ? (void *)(h->p2p_base + h->p2p_off[fam][par][f]) : nullptr;
        fl[f] = on ? (const unsigned int *)(h->p2p_base + h->p2p_flag_off) + (fam * 4 + f) : nullptr;
        if (on) h->halo_bytes += (double)(face_count(h, L.send[f], f) * (size_t)h->fb);
    }
    int zr = 0;
    for (int f = 0; f < 4; f++) { zr += pull.pl[f].n; pull.zend[f] = zr; }
    dim3 blk(128, 1, 1);
    const unsigned gx = (unsigned)((nz + 128 * VEC - 1) / (128 * VEC));
    halo_p2p<F, true><<<dim3(gx, (unsigned)maxline, (unsigned)z), blk, 0, st_>>>(nz, h->NZP, h->NXM, push, seq);
    h->launches++;
    CK(cudaGetLastError());
    unsigned int *err = (unsigned int *)(h->p2p_base + h->p2p_count_off) + 16;   // (behind the eight block counters)
    const unsigned long long tmo = h->comm_timeout_s > 0 ? (unsigned long long)h->comm_timeout_s * 1000000000ull : 0ull;
    halo_wait<<<1, 32, 0, st_>>>(fl[0], fl[1], fl[2], fl[3], seq, tmo, err);
    h->launches++;
    CK(cudaGetLastError());
    halo_p2p<F, false><<<dim3(gx, (unsigned)maxline, (unsigned)zr), blk, 0, st_>>>(nz, h->NZP, h->NXM, pull, seq);
    h->launches++;
    CK(cudaGetLastError());
    if (h->zero_outer && launch_halo<F>(h, L, false, st_, true)) return 1;   // faces without a neighbour: the reference unpacks its zero receive buffers there
    return 0;
}

// pack -> ncclSend/ncclRecv with up to four neighbours -> unpack, all on `st_` (the launch stream, or the exchange
// stream of the boundary-first overlap); with kernel_timing the exchange is bracketed by an event pair (halo phase)
static int comm_exchange(swpc3d_handle *h, int which, cudaStream_t st_ = nullptr) {
    if (ready(h)) return 1;
    if (!st_) st_ = h->st;
    const bool any = has_neighbour(h);
    if (!any && !h->zero_outer) return 0;   // all neighbours MPI_PROC_NULL: outer halos keep their zeros
    const FaceLists L = face_lists(h, which);
    if (!any) return h->fb == 8 ? launch_halo<double>(h, L, false, st_) : launch_halo<float>(h, L, false, st_);
    if (h->comm_dead) return fail("swpc3d_comm_*: the communicator was aborted after a peer failure");
    if (!h->comm) return fail("swpc3d_comm_*: this rank has neighbours but swpc3d_comm_init was not called");
    const bool timed = h->ktiming && h->cev_used < 4096;
    if (timed) {
        if (h->cev_used == h->cev[0].size()) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            h->cev[0].push_back(a);
            h->cev[1].push_back(b);
        }
        CK(cudaEventRecord(h->cev[0][h->cev_used], st_));
    }
    if (h->p2p_ok && h->use_p2p) {
        if (h->fb == 8 ? p2p_exchange<double>(h, L, which, st_) : p2p_exchange<float>(h, L, which, st_)) return 1;
        if (timed) {
            CK(cudaEventRecord(h->cev[1][h->cev_used], st_));
            h->cev_used++;
        }
        return 0;
    }
    TRACE("comm_exchange which=%d n=%u", which, h->nccl_n);
    {   // bound the host's run-ahead: exchange n is only issued once exchange n-4 has completed (polled with the failure checks)
        cudaEvent_t &ev = h->nccl_ev[h->nccl_n & 3];
        if (!ev) CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        else if (event_wait(h, ev)) return 1;
    }
    if (h->fb == 8 ? launch_halo<double>(h, L, true, st_) : launch_halo<float>(h, L, true, st_)) return 1;
    const ncclDataType_t ty = h->fb == 8 ? ncclDouble : ncclFloat;
    NK(g_nccl.GroupStart());
    for (int f = 0; f < 4; f++) {
        if (h->nbr[f] < 0) continue;
        NK(g_nccl.Send(h->sbuf[f], face_count(h, L.send[f], f), ty, h->nbr[f], h->comm, st_));
        NK(g_nccl.Recv(h->rbuf[f], face_count(h, L.recv[f], f), ty, h->nbr[f], h->comm, st_));
        h->halo_bytes += (double)(face_count(h, L.send[f], f) * (size_t)h->fb);
    }
    NK(g_nccl.GroupEnd());
    TRACE("comm_exchange group issued");
    if (h->fb == 8 ? launch_halo<double>(h, L, false, st_) : launch_halo<float>(h, L, false, st_)) return 1;
    CK(cudaEventRecord(h->nccl_ev[h->nccl_n & 3], st_));
    h->nccl_n++;
    if (timed) {
        CK(cudaEventRecord(h->cev[1][h->cev_used], st_));
        h->cev_used++;
    }
    return 0;
}

extern "C" int swpc3d_comm_stress(swpc3d_handle *h) { return comm_exchange(h, 0); }
extern "C" int swpc3d_comm_vel(swpc3d_handle *h) { return comm_exchange(h, 1); }

extern "C" int swpc3d_nccl_unique_id(char id[128]) {
    if (nccl_load()) return 1;
    ncclUniqueId u;
    NK(g_nccl.GetUniqueId(&u));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, 128);
    return 0;
}

extern "C" int swpc3d_comm_init(swpc3d_handle *h, const char id[128], int32_t nranks, int32_t rank) {
    if (!h) return fail("null handle");
    if (nccl_load()) return 1;
    if (nranks != h->g.nproc_x * h->g.nproc_y) return fail("swpc3d_comm_init: nranks != nproc_x*nproc_y (assert, m_global.f90:232)");
    if (rank != h->g.myid) return fail("swpc3d_comm_init: rank != myid");
    CK(cudaSetDevice(h->dev));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    NK(g_nccl.CommInitRank(&h->comm, nranks, u, rank));
    h->comm_rank = rank;
    h->comm_size = nranks;
    // snapshot reductions get a communicator of their own: they are issued on another stream, beside the halo exchange
    // and it is held to ONE channel (one CTA): a snapshot slice is a few MB, and every CTA of a collective kernel takes an SM away
    // from the DRAM-bound sweeps while it waits for its peers (with NCCL's default channel count a reduction cost 1 ms of step time)
    if (g_nccl.CommSplit) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = 1;
        if (g_nccl.CommSplit(h->comm, 0, rank, &h->comm_io, &cfg) != ncclSuccess) h->comm_io = nullptr;
    }
    return p2p_setup(h);
}

// single-process emulation: all handles on GPUs of this process, ordered by myid
extern "C" int swpc3d_comm_local(swpc3d_handle **hs, int32_t n, int32_t which) {
    if (!hs || n <= 0) return fail("swpc3d_comm_local: no handles");
    for (int q = 0; q < n; q++) {
        if (ready(hs[q])) return 1;
        if (hs[q]->g.myid != q) return fail("swpc3d_comm_local: handles must be ordered by myid");
        const FaceLists L = face_lists(hs[q], which);
        if (hs[q]->fb == 8 ? launch_halo<double>(hs[q], L, true, hs[q]->st) : launch_halo<float>(hs[q], L, true, hs[q]->st)) return 1;
    }
    for (int q = 0; q < n; q++) { CK(cudaSetDevice(hs[q]->dev)); CK(cudaStreamSynchronize(hs[q]->st)); }
    const int opp[4] = {1, 0, 3, 2};
    for (int q = 0; q < n; q++) {
        swpc3d_handle *h = hs[q];
        const FaceLists L = face_lists(h, which);
        for (int f = 0; f < 4; f++) {
            if (h->nbr[f] < 0) continue;
            if (h->nbr[f] >= n) return fail("swpc3d_comm_local: neighbour not in the handle list");
            swpc3d_handle *o = hs[h->nbr[f]];
            // on the receiver's own (non-blocking) stream, ahead of its unpack kernels: a device-to-device cudaMemcpy on the
            // legacy default stream neither blocks the host nor orders against non-blocking streams
            CK(cudaMemcpyAsync(o->rbuf[opp[f]], h->sbuf[f], face_count(h, L.send[f], f) * (size_t)h->fb, cudaMemcpyDefault, o->st));
        }
    }
    for (int q = 0; q < n; q++) {
        const FaceLists L = face_lists(hs[q], which);
        CK(cudaSetDevice(hs[q]->dev));
        if (hs[q]->fb == 8 ? launch_halo<double>(hs[q], L, false, hs[q]->st) : launch_halo<float>(hs[q], L, false, hs[q]->st)) return 1;
    }
    for (int q = 0; q < n; q++) { CK(cudaSetDevice(hs[q]->dev)); CK(cudaStreamSynchronize(hs[q]->st)); }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// One field family of the boundary-first schedule: [exchange stream: boundary slabs, source terms that land outside the
// core, pack, NCCL, unpack] beside [launch stream: core sweep, source terms inside the core] -> join.
// The arithmetic per cell and the order sweep -> source -> exchange per cell are the reference's, so the results are
// bit-identical to the exposed order.
template <typename F, bool STRESS>
static int family_overlapped(swpc3d_handle *h, int it) {
    const bool src = h->nsrc > 0 && (STRESS ? !h->bf_mode : h->bf_mode);
    if (launch_sweep<F, STRESS>(h, 3)) return 1;
    if (src && launch_source<F>(h, it, !STRESS, 3)) return 1;
    CK(cudaEventRecord(h->ev_b, h->st));
    CK(cudaStreamWaitEvent(h->cs, h->ev_b, 0));
    if (launch_sweep<F, STRESS>(h, 1)) return 1;
    if (src && launch_source<F>(h, it, !STRESS, 1, h->cs)) return 1;
    if (!h->split_test && comm_exchange(h, STRESS ? 0 : 1, h->cs)) return 1;
    CK(cudaEventRecord(h->ev_c, h->cs));
    if (launch_sweep<F, STRESS>(h, 2)) return 1;
    if (src && launch_source<F>(h, it, !STRESS, 2)) return 1;
    CK(cudaStreamWaitEvent(h->st, h->ev_c, 0));
    return 0;
}

extern "C" int swpc3d_step(swpc3d_handle *h, int32_t it) {
    // main.f90:119-139
    if (swpc3d_green_store(h, it)) return 1;
    if (swpc3d_wav_store(h, it)) return 1;
    return swpc3d_advance(h, it);
}

// main.f90:126-138: everything of iteration `it` after the sampling calls (green__store, wav__store, snap__write)
extern "C" int swpc3d_advance(swpc3d_handle *h, int32_t it) {
    // (green__source writes V cells next to the subdomain edge between the sweep and the exchange: exposed order then)
    if (((h->overlap && has_neighbour(h) && h->comm) || h->split_test) && !h->g_is_src) {
        if (ready(h)) return 1;
        if (h->fb == 8 ? family_overlapped<double, true>(h, it) : family_overlapped<float, true>(h, it)) return 1;
        if (h->fb == 8 ? family_overlapped<double, false>(h, it) : family_overlapped<float, false>(h, it)) return 1;
        return 0;
    }
    if (swpc3d_update_stress(h)) return 1;
    if (swpc3d_stressglut(h, it)) return 1;
    if (swpc3d_comm_stress(h)) return 1;
    if (swpc3d_update_vel(h)) return 1;
    if (swpc3d_bodyforce(h, it)) return 1;
    if (swpc3d_green_source(h, it)) return 1;
    if (swpc3d_comm_vel(h)) return 1;
    return 0;
}

extern "C" int swpc3d_run(swpc3d_handle *h, int32_t it0, int32_t it1) {
    for (int it = it0; it <= it1; it++)
        if (swpc3d_step(h, it)) return 1;
    return 0;
}

extern "C" int swpc3d_timer_start(swpc3d_handle *h) {
    if (!h) return fail("null handle");
    CK(cudaSetDevice(h->dev));
    CK(cudaEventRecord(h->ev0, h->st));
    return 0;
}
extern "C" int swpc3d_timer_stop(swpc3d_handle *h, float *ms) {
    if (!h || !ms) return fail("null argument");
    CK(cudaSetDevice(h->dev));
    CK(cudaEventRecord(h->ev1, h->st));
    WAIT(h, h->st);
    CK(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return 0;
}

extern "C" int swpc3d_set_option(swpc3d_handle *h, const char *key, int32_t value) {
    if (!h || !key) return fail("null argument");
    if (!strcmp(key, "tk")) { if (value < 32 || value > 1024 || value % 32) return fail("tk must be a multiple of 32 in 32..1024"); h->tk = value; }
    else if (!strcmp(key, "ti")) { if (value < 1 || value > 32) return fail("ti must be 1..32"); h->ti = value; }
    else if (!strcmp(key, "jlen")) { if (value < 1) return fail("jlen must be >= 1"); h->jlen = value; }
    else if (!strcmp(key, "pf")) { if (value < 0 || value > 8) return fail("pf must be 0..8"); h->pf = value; }
    else if (!strcmp(key, "variant")) h->variant = value;
    else if (!strcmp(key, "pw_mode")) { h->pw_mode = value != 0; if (value) h->zero_outer = 1; }
    else if (!strcmp(key, "zero_outer_halo")) h->zero_outer = value != 0;
    else if (!strcmp(key, "tma")) h->use_tma = value;
    else if (!strcmp(key, "tma_shift")) h->tma_shift = value != 0;
    else if (!strcmp(key, "tma_persist")) h->tma_persist = value != 0;
    else if (!strcmp(key, "tma_pl")) { if (value < 1) return fail("tma_pl must be >= 1"); h->tma_pl = value; }
    else if (!strcmp(key, "vel_ring")) h->use_ring = value;
    else if (!strcmp(key, "flat_bottom")) h->flat_bottom = value != 0;
    else if (!strcmp(key, "ring_pair")) h->ring_pair = value;
    else if (!strcmp(key, "overlap")) h->overlap = value != 0;
    else if (!strcmp(key, "split_test")) h->split_test = value != 0;
    else if (!strcmp(key, "slab_x")) { if (value < 2) return fail("slab_x must be >= 2"); h->slab_x = value; }
    else if (!strcmp(key, "slab_tiled")) h->slab_tiled = value != 0;
    else if (!strcmp(key, "ring_jlen")) { if (value < 1) return fail("ring_jlen must be >= 1"); h->ring_jlen = value; }
    else if (!strcmp(key, "ring_pf")) { if (value < 0 || value > 8) return fail("ring_pf must be 0..8"); h->ring_pf = value; }
    else if (!strcmp(key, "side_streams")) h->use_side = value;
    else if (!strcmp(key, "comm_timeout_s")) h->comm_timeout_s = value;
    else if (!strcmp(key, "p2p")) h->use_p2p = value != 0;
    else if (!strcmp(key, "pml_tma")) { h->use_pml = value; pml_drop(h); bot_drop(h); }
    else if (!strcmp(key, "bottom_tma")) { h->use_bot = value != 0; pml_drop(h); bot_drop(h); }
    else if (!strcmp(key, "bot_jl")) { if (value < 1) return fail("bot_jl must be >= 1"); h->bot_jl = value; bot_drop(h); }
    else if (!strcmp(key, "l2hint")) h->l2hint = value;
    else if (!strcmp(key, "l2promo")) { h->l2promo = value; h->tma_ready = false; }
    else if (!strcmp(key, "l2promo_halo")) { h->l2promo_halo = value; h->tma_ready = false; }
    else if (!strcmp(key, "pml_promo")) { h->pml_promo = value; pml_drop(h); }
    else if (!strcmp(key, "pml_promo_bottom")) { h->pml_promo_b = value; pml_drop(h); }
    else if (!strcmp(key, "pml_promo_aux_bottom")) { h->pml_promo_aux_b = value; pml_drop(h); }
    else if (!strcmp(key, "pml_jl")) { if (value < 1) return fail("pml_jl must be >= 1"); h->pml_jl = value; pml_drop(h); }
    else if (!strcmp(key, "pml_jl_bottom")) { if (value < 1) return fail("pml_jl_bottom must be >= 1"); h->pml_jl_bottom = value; pml_drop(h); }
    else if (!strcmp(key, "tma_jl")) { if (value < 1) return fail("tma_jl must be >= 1"); h->tma_jl = value; }
    else if (!strcmp(key, "kernel_timing")) { h->ktiming = value != 0; h->kev_used[0] = h->kev_used[1] = 0; h->cev_used = 0; h->halo_bytes = 0.0; }
    else return fail(std::string("unknown option ") + key);
    return 0;
}

extern "C" int swpc3d_get_info(swpc3d_handle *h, const char *key, double *value) {
    if (!h || !key || !value) return fail("null argument");
    if (!strcmp(key, "launches")) *value = (double)h->launches;
    else if (!strcmp(key, "NZP")) *value = h->NZP;
    else if (!strcmp(key, "NXM")) *value = h->NXM;
    else if (!strcmp(key, "NYM")) *value = h->NYM;
    else if (!strcmp(key, "naux")) *value = (double)h->naux;
    else if (!strcmp(key, "device")) *value = h->dev;
    else if (!strcmp(key, "tma_ok")) *value = h->tma_ok ? 1.0 : 0.0;
    else if (!strcmp(key, "p2p_ok")) *value = (h->p2p_ok && h->use_p2p) ? 1.0 : 0.0;
    else if (!strcmp(key, "overlapped")) *value = (h->overlap && has_neighbour(h) && h->comm) ? 1.0 : 0.0;
    else if (!strcmp(key, "bottom_items") || !strcmp(key, "bottom_items_vel")) {
        double n = 0;
        for (const BotPlan *b : h->bot[strstr(key, "_vel") ? 1 : 0]) n += b->ok ? b->nitems : 0;
        *value = n;
    }
    else if (!strncmp(key, "pml_", 4)) {   // the shell plan of the last whole-region (core-region: "_core" suffix) sweeps
        const int w = strstr(key, "_vel") ? 1 : 0;
        const bool core = strstr(key, "_core") != nullptr;
        const Region wr = whole_region(h);
        const PmlPlan *pl = nullptr;
        for (const PmlPlan *c : h->pml[w]) {
            const bool whole = c->rg.li0 == wr.li0 && c->rg.li1 == wr.li1 && c->rg.lj0 == wr.lj0 && c->rg.lj1 == wr.lj1;
            if (whole != core) { pl = c; break; }
        }
        if (!strncmp(key, "pml_items_walls", 15)) *value = pl ? pl->n_items[0] + pl->n_items[2] : 0;
        else if (!strncmp(key, "pml_items_bottom", 16)) *value = pl ? pl->n_items[1] : 0;
        else if (!strncmp(key, "pml_direct_boxes", 16)) *value = pl ? (double)pl->direct.size() : 0;
        else return fail(std::string("unknown info key ") + key);
    }
    else if (!strcmp(key, "ms_stress") || !strcmp(key, "ms_vel") || !strcmp(key, "n_stress") || !strcmp(key, "n_vel")) {
        const int w = strstr(key, "stress") ? 0 : 1;
        if (key[0] == 'n') { *value = (double)h->kev_used[w]; return 0; }
        CK(cudaSetDevice(h->dev));
        WAIT(h, h->st);
        double sum = 0;
        for (size_t q = 0; q < h->kev_used[w]; q++) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, h->kev[w][0][q], h->kev[w][1][q]));
            sum += ms;
        }
        *value = h->kev_used[w] ? sum / (double)h->kev_used[w] : 0.0;   // average launch duration [ms]
    }
    else if (!strcmp(key, "ms_halo") || !strcmp(key, "n_halo") || !strcmp(key, "halo_bytes")) {
        // the halo phase: pack + ncclSend/Recv + unpack of one field family, CUDA events on the stream it ran on
        if (key[0] == 'n') { *value = (double)h->cev_used; return 0; }
        if (key[0] == 'h') { *value = h->halo_bytes; return 0; }
        CK(cudaSetDevice(h->dev));
        WAIT(h, h->st);
        WAIT(h, h->cs);
        double sum = 0;
        for (size_t q = 0; q < h->cev_used; q++) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, h->cev[0][q], h->cev[1][q]));
            sum += ms;
        }
        *value = h->cev_used ? sum / (double)h->cev_used : 0.0;
    }
    else if (!strcmp(key, "device_bytes")) {
        double b = (double)h->ncell * (9.0 * h->fb + 5 * 4 + 6.0 * h->nm * 4) + (double)h->naux * 18 * 4;
        *value = b;
    } else return fail(std::string("unknown info key ") + key);
    return 0;
}
