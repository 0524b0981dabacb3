/*
 * swpc3d_b200.h -- C ABI of the B200-native swpc_3d time-stepping path.
 *
 * The reference (OpenSWPC 25.05.2, Fortran) has no FFI for this path: its seam is the set of public,
 * argument-less module procedures that src/swpc_3d/main.f90:64-78,119-143 calls, all state living in
 * m_global's public allocatables.  Each entry point below replaces one of those procedures (cited
 * as file:line under /root/reference); a Fortran host binds them with ISO_C_BINDING
 * (openswpc_b200/fortran/m_swpc3d_b200.f90).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; swpc3d_last_error() gives the text.
 *     (The Fortran side turns non-zero into the reference's convention: assert -> message + stop,
 *     src/shared/m_debug.f90:206-221.)
 *   - host arrays use the REFERENCE layout and extents: 3-D arrays are (k,i,j), k fastest, over
 *     (kbeg_m:kend_m, ibeg_m:iend_m, jbeg_m:jend_m) with kbeg_m=-2, kend_m=nz+3+kpad, ibeg_m=ibeg-3,
 *     iend_m=iend+3+ipad, ... (m_global.f90:295-300); 2-D maps are (i,j) over (ibeg_m:iend_m,
 *     jbeg_m:jend_m).  The library copies at upload and owns all device memory afterwards.
 *   - field arrays (V*, S*) are `field_bytes`-wide reals: 8 = real(MP=DP) (reference default,
 *     m_global.f90:30), 4 = real(MP=SP).  Everything else is float / int32 as in the reference.
 *   - one host thread per handle; calls are stream-ordered and asynchronous except the get_* /
 *     vmax / download calls.  One GPU per rank.
 */
#ifndef SWPC3D_B200_H
#define SWPC3D_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swpc3d_handle swpc3d_handle;

#define SWPC3D_ABC_PML 1
#define SWPC3D_ABC_CERJAN 2

/* grid / decomposition description == the integers of m_global.f90:52-86 for this rank */
typedef struct {
    int32_t nx, ny, nz;                 /* global grid                         m_global.f90:52  */
    int32_t nproc_x, nproc_y, myid;     /* 2-D decomposition, rank = idy*nproc_x+idx  :234-237  */
    int32_t ibeg, iend, jbeg, jend;     /* owned range (global, 1-based)              :275-288  */
    int32_t ipad, jpad, kpad;           /* host-array padding                         :295-300  */
    int32_t ibeg_k, iend_k, jbeg_k, jend_k, kbeg_k, kend_k; /* interior kernel box    :348-376  */
    int32_t na;                         /* absorber thickness                         :82       */
    int32_t nm;                         /* number of relaxation mechanisms (NM, :31); 0..3      */
    int32_t abc_type;                   /* SWPC3D_ABC_PML | SWPC3D_ABC_CERJAN         :87       */
    int32_t field_bytes;                /* 8 (MP=DP) or 4 (MP=SP)                      :30       */
    int32_t device;                     /* CUDA device ordinal; <0: myid mod ngpus    :208-214  */
    int32_t reserved;
    double dx, dy, dz;                  /* real(MP)                                   :54       */
    float dt;                           /* real(SP)                                   :55       */
    float reserved_f;
} swpc3d_grid;

const char *swpc3d_last_error(void);
const char *swpc3d_version(void);

/* memory_allocate (m_kernel.f90:376-400) + kernel__setup coefficient table (:43-67).  ts[nm] are the
 * relaxation times from visco_set_relaxtime (m_fdtool.f90:691-727); ignored when nm == 0. */
int swpc3d_create(const swpc3d_grid *g, const float *ts, swpc3d_handle **out);
int swpc3d_destroy(swpc3d_handle *h);

/* `!$acc enter data copyin(rho, lam, mu, taup, taus, kfs.., kbeg_a)` of main.f90:100-113 */
int swpc3d_upload_medium(swpc3d_handle *h, const float *rho, const float *lam, const float *mu, const float *taup,
                         const float *taus, const int32_t *kfs, const int32_t *kob, const int32_t *kfs_top,
                         const int32_t *kfs_bot, const int32_t *kob_top, const int32_t *kob_bot, const int32_t *kbeg_a);

/* `!$acc enter data copyin(Vx..Sxy)` of main.f90:80-90 (only needed for non-zero initial fields);
 * a NULL pointer leaves that field untouched.  `field_bytes`-wide elements. */
int swpc3d_upload_fields(swpc3d_handle *h, const void *Vx, const void *Vy, const void *Vz, const void *Sxx,
                         const void *Syy, const void *Szz, const void *Syz, const void *Sxz, const void *Sxy);
int swpc3d_download_fields(swpc3d_handle *h, void *Vx, void *Vy, void *Vz, void *Sxx, void *Syy, void *Szz, void *Syz,
                           void *Sxz, void *Sxy);
int swpc3d_zero_state(swpc3d_handle *h);

/* absorb_p__setup (m_absorb_p.f90:60-124): damping profiles g(1:4, n) over the OWNED ranges
 * (gx*: ibeg:iend, gy*: jbeg:jend, gz*: 1:nz); the 18 ADE arrays are allocated (shell only) here. */
int swpc3d_setup_pml(swpc3d_handle *h, const float *gxc, const float *gxe, const float *gyc, const float *gye,
                     const float *gzc, const float *gze);
/* absorb_c__setup (m_absorb_c.f90:27-111): sponge vectors over the memory ranges (ibeg_m:iend_m ...) */
int swpc3d_setup_cerjan(swpc3d_handle *h, const float *gx_c, const float *gx_b, const float *gy_c, const float *gy_b,
                        const float *gz_c, const float *gz_b);

/* source__setup's device copy-in (m_source.f90:309-310).  Moment mode: mo (already divided by M0),
 * mij are double (real(MP) there); body-force mode (bf_mode != 0): fx,fy,fz in mxx,myy,mzz.
 * srcprm is (2,nsrc); stftype one of boxcar triangle herrmann kupper cosine texp. */
int swpc3d_set_sources(swpc3d_handle *h, int32_t nsrc, const int32_t *isrc, const int32_t *jsrc, const int32_t *ksrc,
                       const double *mo, const double *mxx, const double *myy, const double *mzz, const double *myz,
                       const double *mxz, const double *mxy, const float *srcprm, const char *stftype, int32_t bf_mode,
                       float tbeg);
/* wav__setup's device copy-in (m_wav.f90:134-135).  scale = M0*UC (m_wav.f90:529) */
int swpc3d_set_stations(swpc3d_handle *h, int32_t nst, const int32_t *ist, const int32_t *jst, const int32_t *kst,
                        int32_t ntdec_w, int32_t ntw, float M0, float UC);

/* which station products wav__store keeps (m_wav.f90:67-70 sw_wav_v/u/stress/strain); call after swpc3d_set_stations.
 * Default: velocity only. */
int swpc3d_set_wav_products(swpc3d_handle *h, int32_t sw_v, int32_t sw_u, int32_t sw_stress, int32_t sw_strain);
/* `!$acc update self(wav_*)` m_wav.f90:672-675; which: 0 velocity (ntw,3,nst) [nm/s], 1 displacement (ntw,3,nst) [nm],
 * 2 stress (ntw,6,nst) [Pa], 3 strain (ntw,6,nst) */
int swpc3d_get_wav_product(swpc3d_handle *h, int32_t which, float *out);

/* Green's-function mode, m_green.f90.  green__setup's device copy-in (:351): the grid points of this rank (global 1-based
 * indices), the pseudo source (a station; is_src as redefined at :185-186, i.e. inside ibeg..iend+1 x jbeg..jend+1), the unit
 * force direction fx1 fy1 fz1 (:146-154), green_trise, stftype, ntdec_w and ntw.  bforce = green_bforce (9 instead of 6 traces). */
int swpc3d_set_green(swpc3d_handle *h, int32_t ng, const int32_t *ig, const int32_t *jg, const int32_t *kg, int32_t bforce,
                     int32_t is_src, int32_t isrc, int32_t jsrc, int32_t ksrc, float fx1, float fy1, float fz1, float trise,
                     const char *stftype, int32_t ntdec_w, int32_t ntw, float tbeg);
int swpc3d_green_store(swpc3d_handle *h, int32_t it);    /* green__store  m_green.f90:357-551 */
int swpc3d_green_source(swpc3d_handle *h, int32_t it);   /* green__source m_green.f90:606-649 */
/* `!$acc update self(gf)` m_green.f90:562: gf(ntw, ncmp*ng), ntw fastest, sign as stored (green__export flips z) */
int swpc3d_get_green(swpc3d_handle *h, float *gf);

/* the hot path, one call per reference subroutine */
int swpc3d_update_stress(swpc3d_handle *h);            /* kernel__update_stress m_kernel.f90:142 + absorb__update_stress m_absorb.f90:60 (fused) */
int swpc3d_stressglut(swpc3d_handle *h, int32_t it);   /* source__stressglut    m_source.f90:776 */
int swpc3d_comm_stress(swpc3d_handle *h);              /* global__comm_stress   m_global.f90:500 */
int swpc3d_update_vel(swpc3d_handle *h);               /* kernel__update_vel m_kernel.f90:75 + absorb__update_vel m_absorb.f90:43 (fused) */
int swpc3d_bodyforce(swpc3d_handle *h, int32_t it);    /* source__bodyforce     m_source.f90:850 */
int swpc3d_comm_vel(swpc3d_handle *h);                 /* global__comm_vel      m_global.f90:391 */
int swpc3d_wav_store(swpc3d_handle *h, int32_t it);    /* wav__store (velocity) m_wav.f90:515-539 */
/* one whole iteration of main.f90:119-139 (green_store, wav_store, stress, glut, comm, vel, bodyforce, green_source, comm);
 * with neighbours the exchange runs boundary-first on a second stream beside the core sweeps (option "overlap", default 1) */
int swpc3d_step(swpc3d_handle *h, int32_t it);
/* the part of swpc3d_step after the sampling calls (main.f90:126-138: stress .. comm_vel, overlapped like swpc3d_step), for
 * hosts that put their own snap__write between wav__store and the sweeps */
int swpc3d_advance(swpc3d_handle *h, int32_t it);
/* it = it0..it1 without returning to the host in between */
int swpc3d_run(swpc3d_handle *h, int32_t it0, int32_t it1);
int swpc3d_sync(swpc3d_handle *h);

/* kernel__vmax (m_kernel.f90:350-374): this rank's max |V| at k = kob(i,j)+1, unscaled */
int swpc3d_vmax(swpc3d_handle *h, float out[3]);
/* kernel__vmax + mpi_reduce(MAX) of report__progress (m_report.f90:138-140): NCCL max over all ranks when a
 * communicator is attached, else identical to swpc3d_vmax */
int swpc3d_vmax_global(swpc3d_handle *h, float out[3]);
/* `!$acc update self(wav_vel)` (m_wav.f90:672): (ntw,3,nst) floats */
int swpc3d_get_wav(swpc3d_handle *h, float *wav_vel);

/* ---- snapshots (m_snap.f90): decimated 2-D slices gathered on the device.  product = section*3 + type with
 * section 0 xy, 1 xz, 2 yz, 3 fs, 4 ob and type 0 ps (div, rot_x, rot_y, rot_z), 1 v (Vx,Vy,Vz), 2 u (Ux,Uy,Uz). */
typedef struct {
    int32_t idec, jdec, kdec, ntdec_s;          /* m_snap.f90:116-119 */
    int32_t nxs, nys, nzs;                      /* :123-125 */
    int32_t is0, is1, js0, js1, ks0, ks1;       /* :143-149, this rank's part of the slices */
    int32_t k0_xy, i0_yz, j0_xz;                /* :152-154 */
    int32_t sw[15];                             /* switches xy_ps xy_v xy_u xz_ps ... ob_u */
    float M0, UC;
} swpc3d_snap_cfg;
int swpc3d_snap_setup(swpc3d_handle *h, const swpc3d_snap_cfg *cfg);
/* snap__write(it) device part (m_snap.f90:919-948): displacement accumulation and fs/ob running maxima every step,
 * slice evaluation when mod(it-1, ntdec_s) == 0.  Call at the top of iteration it, after swpc3d_wav_store. */
int swpc3d_snap_step(swpc3d_handle *h, int32_t it);
/* the slice buffer buf(n1,n2,nvar) summed over ranks to `root` (mpi_reduce SUM of m_snap.f90:1064; NCCL when a communicator
 * is attached); out is written on the root only.  _max: max-V/H/A arrays of the fs/ob v and u products (:2295-2348). */
int swpc3d_snap_fetch(swpc3d_handle *h, int32_t product, int32_t root, float *out);
int swpc3d_snap_fetch_max(swpc3d_handle *h, int32_t product, int32_t root, float *out);
/* The same reduction without stopping the time loop -- the reference overlaps it with the sweeps: `mpi_ireduce` of the current
 * record, `mpi_wait` before the next one (m_snap.f90:1057-1064).  _begin (stream-ordered, returns at once) sets the slice
 * buffer aside and starts reduce + device-to-host copy into pinned buffer `slot` (0 or 1) on a stream of its own; _end waits
 * for it and returns the host pointer (valid until the next _begin of that product and slot; NULL on non-root ranks).
 * _end may be called from another host thread than the one that drives the time loop. */
int swpc3d_snap_fetch_begin(swpc3d_handle *h, int32_t product, int32_t root, int32_t slot);
int swpc3d_snap_fetch_end(swpc3d_handle *h, int32_t product, int32_t slot, const float **data);
/* sum-reduce a host float buffer over all ranks to root (setup-time medium slices of the snapshot headers, :600-607) */
int swpc3d_reduce_sum(swpc3d_handle *h, float *buf, int64_t n, int32_t root);

/* multi-GPU: NCCL send/recv replaces the MPI p2p of m_global.f90:408-456, 517-567.  The 128-byte
 * unique id is created on one rank and broadcast by the host (MPI_Bcast in a Fortran host,
 * torch.distributed in the Python host). */
int swpc3d_nccl_unique_id(char id[128]);
int swpc3d_comm_init(swpc3d_handle *h, const char id[128], int32_t nranks, int32_t rank);
/* single-process emulation of the exchange for ranks living on one GPU (tests): handles[] ordered by myid */
int swpc3d_comm_local(swpc3d_handle **handles, int32_t n, int32_t which /* 0 = stress, 1 = velocity */);

/* CUDA-event stopwatch on the stream the kernels are launched on (replaces m_pwatch's wall-clock blocks,
 * src/shared/m_pwatch.f90:60-143): start records an event; stop records, synchronises and returns the ms. */
int swpc3d_timer_start(swpc3d_handle *h);
int swpc3d_timer_stop(swpc3d_handle *h, float *ms);

/* tuning / introspection */
int swpc3d_set_option(swpc3d_handle *h, const char *key, int32_t value);
int swpc3d_get_info(swpc3d_handle *h, const char *key, double *value);

#ifdef __cplusplus
}
#endif
#endif
