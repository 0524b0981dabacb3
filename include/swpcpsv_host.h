/*
 * swpcpsv_host.h -- host-side mirror of swpc_psv's driver (src/swpc_psv/main.f90) for hosts without Fortran: reads
 * input.inf, runs the reference's setup chain for ONE rank on the CPU (setup only: global__setup/setup2, medium__setup,
 * kernel__setup, source__setup, absorb__setup, wav__setup, report__setup), attaches a GPU through swpcpsv_b200.h, steps and
 * writes the waveform files.  The time loop itself never runs on the CPU.  Citations: file:line under
 * /root/reference/src/swpc_psv.
 *
 * Scope of this build: every vmodel_type but the compile-time 'user' plug-in (uni, lhm, lgm, uni_rmed, lhm_rmed, lgm_rmed, grd,
 * grd_rmed -- random-media sections and GMT grids in the netCDF classic container), benchmark_mode, stabilize_pml;
 * stf_format {xy,ll}{m0,mw}{ij,dc} and body forces; PML and Cerjan; station products v / u / stress / strain in
 * sac | csf | tar_st | tar_node containers; snapshots (m_snap.f90: xz_ps / xz_v / xz_u, netcdf or native); plane-wave mode.
 */
#ifndef SWPCPSV_HOST_H
#define SWPCPSV_HOST_H

#include "swpcpsv_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swpcpsv_host swpcpsv_host;

/* main.f90:55-78 for rank `myid`.  nm: NM of m_global.f90:30; overrides <= 0 keep the file's values; field_bytes 8 / 4. */
int swpcpsv_host_create(const char *inf_path, const char *base_dir, int32_t nm, int32_t myid, int32_t nproc_x, int32_t nt,
                        int32_t field_bytes, swpcpsv_host **out);
int swpcpsv_host_create_from_text(const char *inf_text, const char *base_dir, int32_t nm, int32_t myid, int32_t nproc_x, int32_t nt,
                                  int32_t field_bytes, swpcpsv_host **out);
int swpcpsv_host_destroy(swpcpsv_host *h);
const char *swpcpsv_host_last_error(void);

/* scalars by name: nx nz nt na nm nproc_x myid ibeg iend nxp ibeg_k iend_k kend_k nsrc nst ntw ntdec_w ntdec_r bf_mode |
 * dx dz dt xbeg zbeg tbeg vmin vmax vmin_local vmax_local fmax fcut M0 UC zeta d2 c r */
int swpcpsv_host_get_int(swpcpsv_host *h, const char *name, int32_t *v);
int swpcpsv_host_get_double(swpcpsv_host *h, const char *name, double *v);
/* mpi_allreduce of m_medium.f90:315-316 is the caller's job: reduce vmin_local / vmax_local and set them here */
int swpcpsv_host_set_minmax(swpcpsv_host *h, float vmin, float vmax);
int swpcpsv_host_set_exedate(swpcpsv_host *h, int32_t exedate, int32_t tz_minutes);

/* arrays by name (reference layout).  float: rho lam mu taup taus (k,i over the memory box); gxc gxe gzc gze (4,n);
 * gx_c gx_b gz_c gz_b; ts; srcprm (2,nsrc).  int32: kfs kob kfs_top kfs_bot kob_top kob_bot kbeg_a; src_ik (2,nsrc);
 * st_ik (2,nst).  double: mo; m3 (3,nsrc: mxx mzz mxz | fx fz 0).  Returns the element count in *n; copies min(cap, count). */
int swpcpsv_host_get_array(swpcpsv_host *h, const char *name, void *out, int64_t cap, int64_t *n);
int swpcpsv_host_station_name(swpcpsv_host *h, int32_t i, char *buf9);

/* `!$acc enter data` of main.f90:80-93 */
int swpcpsv_host_attach_device(swpcpsv_host *h, int32_t device);
swpcpsv_handle *swpcpsv_host_handle(swpcpsv_host *h);

/* main.f90:95-113 for it = it0..it1; every ntdec_r steps the max amplitudes (NCCL max when a communicator is attached) and,
 * on rank 0 with verbose != 0, the progress line of m_report.f90:164-167.  vm: 2 floats per report. */
int swpcpsv_host_run(swpcpsv_host *h, int32_t it0, int32_t it1, int32_t verbose, float *vm, int32_t nvm, int32_t *nrec);
/* wav__write (m_wav.f90:308-420): <odir>/wav/<title>.psv.<stnm>.<cmp>.sac and the csf / tar containers */
int swpcpsv_host_write_wav(swpcpsv_host *h, const char *odir, int32_t *nfiles);
/* snapshots (m_snap.f90): create <odir>/<title>.psv.xz.<ps|v|u>.<nc|snp> on the I/O ranks (after attach_device / comm init and
 * before the first swpcpsv_host_run, which then writes one record every ntdec_s steps); close flushes and closes. */
int swpcpsv_host_snap_open(swpcpsv_host *h, const char *odir);
int swpcpsv_host_snap_close(swpcpsv_host *h);
/* report__setup banner (m_report.f90:52-100) to stderr; fails when the stability condition is violated */
int swpcpsv_host_banner(swpcpsv_host *h);

#ifdef __cplusplus
}
#endif
#endif
