/*
 * swpcpsv_b200.h -- C ABI of the B200-native swpc_psv (2-D P-SV) time-stepping path.
 *
 * Same contract as swpc3d_b200.h: the reference (OpenSWPC 25.05.2, src/swpc_psv) has no FFI, its seam is the set of
 * argument-less module procedures called from src/swpc_psv/main.f90:64-78, 95-113; each entry point below replaces one of
 * them (cited as file:line under /root/reference/src/swpc_psv).
 *
 * Conventions
 *   - every function returns 0 on success; swpcpsv_last_error() gives the text of the last failure.
 *   - host arrays use the REFERENCE layout: 2-D arrays are (k,i), k fastest, over (kbeg_m:kend_m, ibeg_m:iend_m) with
 *     kbeg_m = -2, kend_m = nz+3+kpad, ibeg_m = ibeg-3, iend_m = iend+3+ipad (m_global.f90:244-247); 1-D maps are over
 *     (ibeg_m:iend_m).  The library copies at upload and owns all device memory afterwards.
 *   - V*, S* are `field_bytes`-wide reals (8: MP = DP, m_global.f90:29; 4: MP = SP); everything else float / int32.
 *   - one host thread per handle; calls are stream-ordered and asynchronous except get_* / vmax / download.
 */
#ifndef SWPCPSV_B200_H
#define SWPCPSV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swpcpsv_handle swpcpsv_handle;

#define SWPCPSV_ABC_PML 1
#define SWPCPSV_ABC_CERJAN 2

/* the integers of m_global.f90:41-69 for this rank */
typedef struct {
    int32_t nx, nz;                     /* global grid                                 m_global.f90:41      */
    int32_t nproc_x, myid;              /* 1-D decomposition along x, idx = myid       :196-208, :452       */
    int32_t ibeg, iend;                 /* owned columns (global, 1-based)             :232-238             */
    int32_t ipad, kpad;                 /* host-array padding                          :244-247             */
    int32_t ibeg_k, iend_k, kend_k;     /* interior kernel box (kbeg_k = 1)            :270-292             */
    int32_t na;                         /* absorber thickness                          :63                  */
    int32_t nm;                         /* relaxation mechanisms (NM, :30); 0..3                            */
    int32_t abc_type;                   /* SWPCPSV_ABC_PML | SWPCPSV_ABC_CERJAN        :67                  */
    int32_t field_bytes;                /* 8 (MP=DP) or 4 (MP=SP)                      :29                  */
    int32_t device;                     /* CUDA device ordinal; <0: myid mod ngpus     :178-184             */
    double dx, dz;                      /* real(MP)                                    :43                  */
    float dt;                           /* real(SP)                                    :44                  */
    float reserved_f;
} swpcpsv_grid;

/* snap__setup (m_snap.f90:78-164): decimation, snapshot grid size, this rank's region is0..is1 x ks0..ks1 (1-based snapshot
 * indices, :111-114), the three product switches xz_ps / xz_v / xz_u, and the output scaling */
typedef struct {
    int32_t idec, kdec, ntdec_s, nxs, nzs, is0, is1, ks0, ks1, sw_ps, sw_v, sw_u;
    float M0, UC;
} swpcpsv_snap_cfg;

const char *swpcpsv_last_error(void);
const char *swpcpsv_version(void);

/* memory_allocate (m_kernel.f90:329-343) + kernel__setup coefficients (:45-67); ts[nm] from visco_set_relaxtime */
int swpcpsv_create(const swpcpsv_grid *g, const float *ts, swpcpsv_handle **out);
int swpcpsv_destroy(swpcpsv_handle *h);

/* `!$acc enter data copyin(rho, lam, mu, taup, taus, kfs.., kbeg_a)` of main.f90:80-93 */
int swpcpsv_upload_medium(swpcpsv_handle *h, const float *rho, const float *lam, const float *mu, const float *taup,
                          const float *taus, const int32_t *kfs, const int32_t *kob, const int32_t *kfs_top,
                          const int32_t *kfs_bot, const int32_t *kob_top, const int32_t *kob_bot, const int32_t *kbeg_a);
/* `!$acc enter data copyin(Vx, Vz, Sxx, Szz, Sxz)` (main.f90:80-84); NULL leaves a field untouched */
int swpcpsv_upload_fields(swpcpsv_handle *h, const void *Vx, const void *Vz, const void *Sxx, const void *Szz, const void *Sxz);
int swpcpsv_download_fields(swpcpsv_handle *h, void *Vx, void *Vz, void *Sxx, void *Szz, void *Sxz);
/* memory variables in the reference layout (m, k, i) over the memory box (m_kernel.f90:337-339); NULL skips */
int swpcpsv_download_memvars(swpcpsv_handle *h, float *Rxx, float *Rzz, float *Rxz);
int swpcpsv_zero_state(swpcpsv_handle *h);

/* absorb_p__setup (m_absorb_p.f90:57-101): g(1:4, ibeg:iend) and g(1:4, 1:nz); the 8 ADE arrays (absorber cells only) */
int swpcpsv_setup_pml(swpcpsv_handle *h, const float *gxc, const float *gxe, const float *gzc, const float *gze);
/* absorb_c__setup (m_absorb_c.f90:28-96): sponge vectors over (ibeg_m:iend_m) and (kbeg_m:kend_m) */
int swpcpsv_setup_cerjan(swpcpsv_handle *h, const float *gx_c, const float *gx_b, const float *gz_c, const float *gz_b);

/* source__setup's device copy-in (m_source.f90:254).  Moment mode: mo (already / M0), mxx, mzz, mxz (real(MP));
 * body-force mode (bf_mode != 0): fx, fz in mxx, mzz.  srcprm is (2, nsrc). */
int swpcpsv_set_sources(swpcpsv_handle *h, int32_t nsrc, const int32_t *isrc, const int32_t *ksrc, const double *mo,
                        const double *mxx, const double *mzz, const double *mxz, const float *srcprm, const char *stftype,
                        int32_t bf_mode, float tbeg);
/* wav__setup's device copy-in (m_wav.f90:137-138) and product switches (:63-66) */
int swpcpsv_set_stations(swpcpsv_handle *h, int32_t nst, const int32_t *ist, const int32_t *kst, int32_t ntdec_w, int32_t ntw,
                         float M0, float UC, int32_t sw_v, int32_t sw_u, int32_t sw_stress, int32_t sw_strain);
/* `!$acc update self(wav_*)` m_wav.f90:321-324; which: 0 velocity (ntw,2,nst) [nm/s], 1 displacement (ntw,2,nst) [nm],
 * 2 stress (ntw,3,nst) [Pa], 3 strain (ntw,3,nst) */
int swpcpsv_get_wav(swpcpsv_handle *h, int32_t which, float *out);

/* the hot path, main.f90:95-113 */
int swpcpsv_update_stress(swpcpsv_handle *h);            /* kernel__update_stress m_kernel.f90:142 + absorb__update_stress m_absorb.f90:60 (one fused sweep) */
int swpcpsv_stressglut(swpcpsv_handle *h, int32_t it);   /* source__stressglut    m_source.f90:550 */
int swpcpsv_comm_stress(swpcpsv_handle *h);              /* global__comm_stress   m_global.f90:366 */
/* kernel__update_vel m_kernel.f90:76 -> source__bodyforce m_source.f90:591 -> absorb__update_vel m_absorb.f90:43, in the
 * reference's order (main.f90:108-110).  Without body forces the three are one fused sweep; in bf_mode the interior sweep,
 * the force injection and the absorber sweep are three launches so that every cell sees the reference's summation order. */
int swpcpsv_update_vel(swpcpsv_handle *h, int32_t it);
int swpcpsv_comm_vel(swpcpsv_handle *h);                 /* global__comm_vel      m_global.f90:312 */
int swpcpsv_wav_store(swpcpsv_handle *h, int32_t it);    /* wav__store            m_wav.f90:143-306 */
int swpcpsv_step(swpcpsv_handle *h, int32_t it);         /* one iteration (snap_step, wav_store, stress .. comm_vel; without report) */

/* snapshots (m_snap.f90).  swpcpsv_snap_step = the device part of snap__write(it) (:435-650): displacement accumulation
 * every step, ps / v slices when mod(it-1, ntdec_s) == 0.  swpcpsv_snap_fetch = mpi_reduce(SUM) of buf(nxs, nzs, 2) onto the
 * I/O rank (product 0 ps, 1 v, 2 u; :395-417, :500-507): every rank calls it, `out` is filled on `root`.
 * swpcpsv_reduce_sum does the same for a host array (the medium slices of newfile_xz, :167-269). */
int swpcpsv_snap_setup(swpcpsv_handle *h, const swpcpsv_snap_cfg *cfg);
int swpcpsv_snap_step(swpcpsv_handle *h, int32_t it);
int swpcpsv_snap_fetch(swpcpsv_handle *h, int32_t product, int32_t root, float *out);
int swpcpsv_reduce_sum(swpcpsv_handle *h, float *buf, int64_t n, int32_t root);
int swpcpsv_run(swpcpsv_handle *h, int32_t it0, int32_t it1);
int swpcpsv_sync(swpcpsv_handle *h);

/* kernel__vmax (m_kernel.f90:313-327): this rank's max |Vx|, |Vz| at k = kob(i)+1, unscaled; _global adds the
 * mpi_reduce(MAX) of report__progress (m_report.f90:136-137) over NCCL when a communicator is attached */
int swpcpsv_vmax(swpcpsv_handle *h, float out[2]);
int swpcpsv_vmax_global(swpcpsv_handle *h, float out[2]);

/* multi-GPU: NCCL send/recv replaces the MPI p2p of m_global.f90:325-346, 379-400 (3 columns of nz per neighbour) */
int swpcpsv_nccl_unique_id(char id[128]);
int swpcpsv_comm_init(swpcpsv_handle *h, const char id[128], int32_t nranks, int32_t rank);
/* single-process emulation of the exchange for ranks living on one GPU (tests): handles[] ordered by myid */
int swpcpsv_comm_local(swpcpsv_handle **handles, int32_t n, int32_t which /* 0 = stress, 1 = velocity */);

/* CUDA-event stopwatch on the launch stream (m_pwatch replacement) */
int swpcpsv_timer_start(swpcpsv_handle *h);
int swpcpsv_timer_stop(swpcpsv_handle *h, float *ms);

/* tuning / introspection: options tk, ilen, pf, kernel_timing; infos launches, state_bytes, cells_interior,
 * cells_absorber, ms_stress, ms_vel */
int swpcpsv_set_option(swpcpsv_handle *h, const char *key, int32_t value);
int swpcpsv_get_info(swpcpsv_handle *h, const char *key, double *value);

#ifdef __cplusplus
}
#endif
#endif
