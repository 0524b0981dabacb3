/*
 * swpc3d_host.h -- host-side mirror of swpc_3d's driver (src/swpc_3d/main.f90) for hosts that have no
 * Fortran: reads input.inf, runs the reference's setup chain for ONE rank on the CPU (setup only:
 * global__setup/setup2, medium__setup, kernel__setup, source__setup, absorb__setup, wav__setup), attaches a
 * GPU through the kernel ABI of swpc3d_b200.h, steps, and writes SAC files.  The time loop itself never runs
 * on the CPU.
 *
 * A Fortran host does not need this layer: it keeps its own setup modules and binds swpc3d_b200.h directly
 * (INTEGRATION.md).
 */
#ifndef SWPC3D_HOST_H
#define SWPC3D_HOST_H

#include "swpc3d_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swpc3d_host swpc3d_host;

/* main.f90:55-78 for rank `myid`.  base_dir resolves relative file names of the parameter file (the reference
 * resolves against the cwd).  nm: NM of m_global.f90:31.  Overrides <= 0 keep the file's values.
 * field_bytes: 8 (MP=DP) / 4 (MP=SP).  No GPU is touched. */
int swpc3d_host_create(const char *inf_path, const char *base_dir, int32_t nm, int32_t myid, int32_t nproc_x,
                       int32_t nproc_y, int32_t nt, int32_t field_bytes, swpc3d_host **out);
int swpc3d_host_create_from_text(const char *inf_text, const char *base_dir, int32_t nm, int32_t myid, int32_t nproc_x,
                                 int32_t nproc_y, int32_t nt, int32_t field_bytes, swpc3d_host **out);
int swpc3d_host_destroy(swpc3d_host *h);
const char *swpc3d_host_last_error(void);

/* scalars by name: nx ny nz nt na nm nproc_x nproc_y myid ibeg iend jbeg jend nxp nyp ibeg_k iend_k jbeg_k jend_k
 * kbeg_k kend_k nsrc nst ntw ntdec_w ntdec_r bf_mode | dx dy dz dt xbeg ybeg zbeg tbeg vmin vmax vmin_local vmax_local
 * fmax fcut M0 UC zeta d2 c r */
int swpc3d_host_get_int(swpc3d_host *h, const char *name, int32_t *v);
int swpc3d_host_get_double(swpc3d_host *h, const char *name, double *v);
int swpc3d_host_get_string(swpc3d_host *h, const char *name, char *buf, int32_t cap);
/* mpi_allreduce of m_medium.f90:424-425 is the caller's job: reduce vmin_local/vmax_local and set them here */
int swpc3d_host_set_minmax(swpc3d_host *h, float vmin, float vmax);
int swpc3d_host_set_exedate(swpc3d_host *h, int32_t exedate, int32_t tz_minutes);

/* arrays by name (reference layout).  float: rho lam mu taup taus (k,i,j over the memory box); gxc gxe gyc gye gzc
 * gze (4,n); gx_c gx_b gy_c gy_b gz_c gz_b; ts c1 c2 d1; srcprm (2,nsrc).  int32: kfs kob kfs_top kfs_bot kob_top
 * kob_bot kbeg_a (i,j over the memory box); src_ijk (3,nsrc); st_ijk (3,nst).  double: mo; mij (6,nsrc).
 * Returns the element count in *n; copies min(cap, count) elements when out != NULL. */
int swpc3d_host_get_array(swpc3d_host *h, const char *name, void *out, int64_t cap, int64_t *n);
int swpc3d_host_station_name(swpc3d_host *h, int32_t i, char *buf9);

/* `!$acc enter data` of main.f90:80-113: create the device state and upload everything */
int swpc3d_host_attach_device(swpc3d_host *h, int32_t device);
swpc3d_handle *swpc3d_host_handle(swpc3d_host *h);

/* the time loop main.f90:119-139 for it = it0..it1.  Every ntdec_r steps the local max amplitudes are reduced
 * (NCCL max when a communicator is attached) and, on rank 0, a progress line in the reference's format
 * (m_report.f90:176-180) goes to stderr when verbose != 0.  vm (3 floats per report, capacity nvm reports) may be NULL. */
int swpc3d_host_run(swpc3d_host *h, int32_t it0, int32_t it1, int32_t verbose, float *vm, int32_t nvm, int32_t *nrec);
/* wav__write (m_wav.f90:658-792), SAC format: <odir>/wav/<title>.3d.<stnm>.<cmp>.sac ; returns file count in *nfiles */
int swpc3d_host_write_sac(swpc3d_host *h, const char *odir, int32_t *nfiles);
/* pwatch__report (m_pwatch.f90:146-195, main.f90:148-154): <odir>/<title>.tim for this rank when stopwatch_mode is on (the
 * default), in the reference's table layout; the phases are the ones the library's CUDA-event stopwatches bracket */
int swpc3d_host_write_tim(swpc3d_host *h, const char *odir);
/* Green's-function mode (green_mode = .true., m_green.f90).  The pseudo source is a station: its owner rank finds it with
 * wav__stquery and the reference broadcasts indices and coordinates (m_green.f90:161-183).  A multi-rank host does the
 * same: query every rank, hand the owner's answer to all of them (single-rank runs need neither call).
 * swpc3d_host_write_green = green__export (:553-604): wav_format 'sac' -> <odir>/green/<stnm>/<title>__<gid>__<stnm>__<cmp>__<mij>__.sac,
 * 'csf' -> one container per rank.  Extra get_int names: green_mode ng green_ncmp green_ntw; arrays: green_ijk (3,ng) green_gid green_gf. */
int swpc3d_host_green_query(swpc3d_host *h, int32_t *found, int32_t ijk[3], float xyz[3], float lonlat[2]);
int swpc3d_host_green_set_source(swpc3d_host *h, const int32_t ijk[3], const float xyz[3], const float lonlat[2]);
int swpc3d_host_write_green(swpc3d_host *h, const char *odir, int32_t *nfiles);
/* snapshots (m_snap.f90, snp_format = 'netcdf'): create <odir>/<title>.3d.<xy|xz|yz|fs|ob>.<ps|v|u>.nc on the I/O ranks
 * (call after attach_device / comm init and before the first swpc3d_host_run); swpc3d_host_run then writes one record every
 * ntdec_s steps; swpc3d_host_snap_close flushes the running maxima (max-V/H/A) and closes (snap__closefiles). */
int swpc3d_host_snap_open(swpc3d_host *h, const char *odir);
int swpc3d_host_snap_close(swpc3d_host *h);
int swpc3d_host_nc_selftest(const char *path);   /* writes a tiny CDF-1 file with the in-tree writer (tests) */
/* report__setup banner values (m_report.f90:71-105) to stderr */
int swpc3d_host_banner(swpc3d_host *h);

#ifdef __cplusplus
}
#endif
#endif
