/* psv.h -- CPU restatement of OpenSWPC's swpc_psv (2-D P-SV) setup chain and time step.
 *
 * TEST INFRASTRUCTURE, like the rest of oracle/: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it; the product (openswpc_b200/csrc/psv) never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no swpc_psv output, golden vector or test, and cannot be built here (no Fortran
 * compiler, MPI or netCDF).  This file follows src/swpc_psv/ *.f90 line by line (kinds, evaluation order, loop bounds,
 * quirks) and is itself checked only through self-consistency tests (decomposition independence, symmetry).
 *
 * Quirks kept (Q-psv):
 *   1. surface_detection assigns kfs_top twice and never kfs_bot (m_medium.f90:281-282); kfs_bot keeps its allocation
 *      value, taken as 0 here (fresh pages).
 *   2. the halo receive buffers are never initialised (m_global.f90:222-223) and are unpacked on outer ranks too
 *      (:312-319); taken as 0 here, i.e. the outer halo columns are zeroed at every exchange.
 *   3. body force is applied between the interior and the absorber velocity updates (main.f90:108-110).
 */
#ifndef ORACLE_PSV_H
#define ORACLE_PSV_H

#include "ora.h"

typedef struct psv_sim psv_sim;

psv_sim *psv_create(const char *inf_path, const char *base_dir, int nm, int nproc_x, int nt);
psv_sim *psv_create_from_text(const char *text, const char *base_dir, int nm, int nproc_x, int nt);
void psv_destroy(psv_sim *s);
const char *psv_last_error(void);

void psv_set_exedate(psv_sim *s, int exedate, int tz_minutes);
void psv_step(psv_sim *s, int it);                         /* main.f90:97-113 */
int psv_run(psv_sim *s, int it0, int it1, float *vm, int nvm); /* + report__progress amplitudes (vx, vz) every ntdec_r */
void psv_vmax(psv_sim *s, float out[2]);

int psv_nranks(const psv_sim *s);
/* which: 0 ibeg 1 iend 2 ibeg_k 3 iend_k 4 kbeg_k 5 kend_k 6 nsrc 7 nst 8 nzm 9 nxm 10 ibeg_m 11 kbeg_m 12 kbeg_min 13 nxp */
int psv_rank_int(const psv_sim *s, int q, int which);
/* which: 0 vmin 1 vmax 2 fmax 3 fcut 4 M0 5 UC 6 zeta 7 d2 8 dt 9 xbeg 10 zbeg 11 dx 12 dz 13 evlo 14 evla 15.. ts[m] 23.. c1 31.. c2 39.. d1 47 tbeg 48 r20x 49 r20z */
double psv_cfg_value(const psv_sim *s, int which);
/* which: 0 nx 1 nz 2 nt 3 na 4 nm 5 nproc_x 6 ntw 7 ntdec_w 8 ntdec_r 9 bf_mode 10 pw_mode 11 sw_v 12 sw_u 13 sw_stress 14 sw_strain */
int psv_cfg_int(const psv_sim *s, int which);
const char *psv_cfg_str(const psv_sim *s, int which);     /* 0 title 1 odir 2 abc_type 3 stftype */

int psv_get_field(const psv_sim *s, int q, const char *name, double *out);   /* Vx Vz Sxx Szz Sxz rho lam mu taup taus: (nxm, nzm) */
int psv_set_field(psv_sim *s, int q, const char *name, const double *in);         /* tests: replace a field / medium array */
void psv_redetect_surface(psv_sim *s);                                          /* ... and re-run surface_detection */
int psv_get_memvar(const psv_sim *s, int q, const char *name, float *out);      /* Rxx Rzz Rxz: (nm, nzm, nxm) */
int psv_get_map(const psv_sim *s, int q, const char *name, int *out);        /* kfs kob kfs_top kfs_bot kob_top kob_bot kbeg_a: (nxm) */
int psv_get_profile(const psv_sim *s, int q, const char *name, float *out);  /* gxc gxe (4,nxp) gzc gze (4,nz) gx_c gx_b (nxm) gz_c gz_b (nzm) */
int psv_get_sources(const psv_sim *s, int q, int *ik, double *val);          /* ik (2,nsrc); val (6,nsrc): mo mxx mzz mxz t0 tr  (bf: fx fz in mxx mzz) */
int psv_get_stations(const psv_sim *s, int q, int *ik, char *names9);
int psv_get_wav(const psv_sim *s, int q, int prod, float *out);              /* prod 0 v 1 u (ntw,2,nst); 2 stress 3 strain (ntw,3,nst) */
/* snapshots (m_snap.f90): info = idec kdec ntdec_s nxs nzs sw_ps sw_v sw_u; products 0 ps (div, rot) 1 v 2 u; records (2, nzs, nxs) */
int psv_snap_info(const psv_sim *s, int *info);
int psv_snap_coords(const psv_sim *s, float *x, float *z);
int psv_snap_nrec(const psv_sim *s, int p);
int psv_snap_rec(const psv_sim *s, int p, int rec, float *out, int *it0);
int psv_snap_medium(const psv_sim *s, int which, float *out);
int psv_write_sac(psv_sim *s, const char *odir);                             /* wav_format = sac; returns the number of files */

#endif
