/*
 * oracle/ora.h -- CPU restatement of OpenSWPC swpc_3d's time-stepping path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load
 * or run it, and only as the checker / the CPU baseline.  The product path (openswpc_b200/)
 * never links or imports it.
 *
 * The reference (Fortran 2008 + MPI + OpenMP, version 25.05.2) cannot be compiled in this image
 * (no gfortran / MPI / netCDF), so this is a restatement in C99 that keeps
 *   - the declared type of every temporary (real(SP) -> float, real(MP) -> ora_mp),
 *   - the expression order of the reference (compiled with -ffp-contract=off, no fast-math),
 *   - the per-rank array extents, loop ranges and halo plane lists,
 * with MPI ranks emulated in one process (one ora_rank per emulated rank).
 *
 * Parity pinning: checked against the reference's only known answers, example/example.out
 * (header values c, r, vmin, vmax, fmax and the 20 max-amplitude triplets); see tests/.
 *
 * All citations are file:line in /root/reference (OpenSWPC 25.05.2).
 */
#ifndef ORA_H
#define ORA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* m_global.f90:30  "integer, parameter :: MP = DP  !< DP for mixed, SP for single precisions" */
#ifdef ORA_MP_SP
typedef float ora_mp;
#else
typedef double ora_mp;
#endif

#define ORA_NBD 9        /* m_global.f90:32 */
#define ORA_MAXNM 8
#define ORA_STRLEN 256

/* ---------------------------------------------------------------- ini file (m_readini.f90) */
typedef struct {
    int nlines;
    char **lines;
    int strict_mode;
    int verbose;
} ora_ini;

ora_ini *ora_ini_open(const char *path);
ora_ini *ora_ini_from_text(const char *text);
void ora_ini_close(ora_ini *ini);
/* returns 1 when the key was found, 0 when the default was used */
int ora_readini_c(const ora_ini *ini, const char *key, char *var, const char *def);
int ora_readini_d(const ora_ini *ini, const char *key, double *var, double def);
int ora_readini_s(const ora_ini *ini, const char *key, float *var, float def);
int ora_readini_i(const ora_ini *ini, const char *key, int *var, int def);
int ora_readini_l(const ora_ini *ini, const char *key, int *var, int def);

/* ---------------------------------------------------------------- m_fdtool / m_std helpers */
int ora_x2i(float x, float xbeg, float dx);                 /* m_fdtool.f90:600-636 */
float ora_i2x(int i, float xbeg, float dx);                 /* m_fdtool.f90:639-675 */
float ora_momentrate(float t, const char *stftype, const float *srcprm); /* m_fdtool.f90:472-497 */
void ora_visco_set_relaxtime(int nm, float *ts, float fmin, float fmax); /* m_fdtool.f90:691-727 */
float ora_visco_constq_zeta(int nm, float fmin, float fmax, const float *ts); /* :756-812 */
void ora_fdm_stable_dt(float dx, float dy, float dz, float vmax, float *dt); /* :81-96 */
float ora_moment_magnitude(float m0);                       /* m_fdtool.f90:281-293 */
float ora_deg2rad(float deg);                                  /* m_std.f90:132-139 */
float ora_powi_sp(float x, int m);                             /* real ** integer (libgcc __powisf2) */
float ora_seismic_moment(float mw);                         /* m_fdtool.f90:296-304 */
void ora_sdr2moment(float strike, float dip, float rake, float *mxx, float *myy, float *mzz,
                    float *myz, float *mxz, float *mxy);    /* m_fdtool.f90:307-336 */
float ora_seawater_vel(float z, int use_munk);              /* m_seawater.f90:34-47 */
void ora_geomap_g2c(float lon, float lat, float lon0, float lat0, float phi, float *x, float *y);
void ora_geomap_c2g(float x, float y, float lon0, float lat0, float phi, float *lon, float *lat);
void ora_decomp1d(int n, int nproc, int proc, int *np, int *beg, int *end); /* m_global.f90:234-288 */
void ora_damping_profile(float x, float H, float xbeg0, float xend0, int na, float fcut, float dt,
                         float g[4]);                       /* m_absorb_p.f90:533-573 */

/* ---------------------------------------------------------------- configuration (global) */
typedef struct {
    /* m_global.f90:124-177 */
    int benchmark_mode;
    char title[ORA_STRLEN];
    int nproc_x, nproc_y, nproc;
    int nx, ny, nz, nt;
    int ipad, jpad, kpad;
    char odir[ORA_STRLEN];
    double dx, dy, dz;           /* real(MP) in the reference; read as double */
    float dt;
    int na;
    float xbeg, ybeg, zbeg, tbeg;
    float xend, yend, zend, tend;
    float clon, clat, phi;
    char abc_type[16];
    int nm;                      /* compile-time NM in the reference (m_global.f90:31) */
    /* medium (m_medium.f90:79-84) */
    float fq_min, fq_max, fq_ref;
    char vmodel_type[16];
    float vcut;
    /* source (m_source.f90:67-101) */
    int pw_mode, green_mode, bf_mode;
    char fn_stf[ORA_STRLEN];
    char stftype[16];
    char stf_format[8];
    char sdep_fit[8];
    int earth_flattening;
    /* wav (m_wav.f90:66-74) */
    int ntdec_w;
    int sw_wav_v, sw_wav_u, sw_wav_stress, sw_wav_strain;
    char wav_format[16];
    char st_format[8];
    char fn_stloc[ORA_STRLEN];
    /* report */
    int ntdec_r;
    /* derived / reduced over ranks */
    float vmin, vmax, fmax, fcut;
    float M0, UC;
    float ts[ORA_MAXNM];
    float zeta;
    /* kernel coefficients m_kernel.f90:43-67 */
    ora_mp rc40x, rc41x, rc40y, rc41y, rc40z, rc41z;
    ora_mp rd40x, rd41x, rd40y, rd41y, rd40z, rd41z;
    float c1[ORA_MAXNM], c2[ORA_MAXNM], d1[ORA_MAXNM], d2;
    ora_mp r20x, r20y, r20z;     /* m_absorb_p.f90:70-72 */
    ora_mp dt_dxyz;              /* m_source.f90:306 */
    /* hypocentre bookkeeping for SAC headers */
    float evlo, evla, evdp, mxx0, myy0, mzz0, myz0, mxz0, mxy0, fx0, fy0, fz0, otim, sx0, sy0;
    int exedate;                 /* unix time */
    int tz_minutes;              /* date_and_time values(4) */
    int ntw;
} ora_cfg;

/* ---------------------------------------------------------------- one emulated MPI rank */
typedef struct {
    int myid, idx, idy;
    int nxp, nyp;
    int ibeg, iend, jbeg, jend, kbeg, kend;
    int ibeg_m, iend_m, jbeg_m, jend_m, kbeg_m, kend_m;
    int ibeg_k, iend_k, jbeg_k, jend_k, kbeg_k, kend_k;
    int nzm, nxm, nym;            /* extents of the _m box */
    size_t ncell_m;
    /* fields (k,i,j) over the _m box */
    ora_mp *Vx, *Vy, *Vz, *Sxx, *Syy, *Szz, *Syz, *Sxz, *Sxy;
    /* memory variables (m,k,i,j) over the _k box */
    float *Rxx, *Ryy, *Rzz, *Ryz, *Rxz, *Rxy;
    int nzk, nxk, nyk;
    /* medium */
    float *rho, *lam, *mu, *taup, *taus;
    int *kfs, *kob, *kfs_top, *kfs_bot, *kob_top, *kob_bot, *kbeg_a;   /* (i,j) over _m */
    int psv;                      /* 1: a swpc_psv section seen as a one-plane rank by the model builders (ora_models.c) */
    float *bddep;                 /* (i,j,0:NBD) */
    float *xc, *yc, *zc;
    /* PML */
    float *gxc, *gxe, *gyc, *gye, *gzc, *gze;   /* (4, n) */
    int64_t *aoff;                /* (i,j) over owned cells: offset of column into aux arrays */
    int64_t naux;
    float *axVx, *ayVx, *azVx, *axVy, *ayVy, *azVy, *axVz, *ayVz, *azVz;
    float *axSxx, *aySxy, *azSxz, *axSxy, *aySyy, *azSyz, *axSxz, *aySyz, *azSzz;
    /* Cerjan */
    float *gx_c, *gx_b, *gy_c, *gy_b, *gz_c, *gz_b;
    /* sources owned (incl. sleeve) */
    int nsrc;
    int *isrc, *jsrc, *ksrc;
    float *sx, *sy, *sz;
    float *srcprm;                /* (2,nsrc) */
    ora_mp *mo, *mxx, *myy, *mzz, *myz, *mxz, *mxy, *fx, *fy, *fz;
    /* stations owned */
    int nst;
    int *ist, *jst, *kst;
    float *xst, *yst, *zst, *stlo, *stla;
    char (*stnm)[9];
    float *wav_vel;               /* (ntw,3,nst) */
    float *wav_disp;              /* (ntw,3,nst) */
    float *wav_stress, *wav_strain; /* (ntw,6,nst) */
    float *ux, *uy, *uz;          /* running displacement at stations, m_wav.f90:41 */
    float *exx, *eyy, *ezz, *eyz, *exz, *exy;   /* running strain, m_wav.f90:42 */
    /* halo buffers m_global.f90:251-258 */
    ora_mp *sbuf_ip, *sbuf_im, *sbuf_jp, *sbuf_jm, *rbuf_ip, *rbuf_im, *rbuf_jp, *rbuf_jm;
} ora_rank;

/* ---------------------------------------------------------------- Green's function mode (m_green.f90, ora_green.c) */
typedef struct {
    int ng;                       /* grid points of the list file owned by this rank (m_green.f90:216-280) */
    int *ig, *jg, *kg, *gid;
    float *xg, *yg, *zg, *lon, *lat;
    float *gf;                    /* (ntw, ncmp*ng) */
    float *acc;                   /* 12 running sums per point: dxUx dxUy dxUz dyUx dyUy dyUz dzUx dzUy dzUz Ux Uy Uz */
    int is_src;                   /* redefined is_src, m_green.f90:185-186 */
} ora_green_rank;

typedef struct {
    char stnm[9];
    char cmp;                     /* 'x' 'y' 'z' */
    float trise, maxdist;
    int bforce, ncmp;
    int isrc, jsrc, ksrc;
    float xsrc, ysrc, zsrc, evlo0, evla0;
    float fx1, fy1, fz1;
    float dt_dxyz;                /* real(SP) in m_green.f90:32 */
    int ntdec_w, ntw;
    char stftype[16], wav_format[16];
    ora_mp r40x, r40y, r40z, r41x, r41y, r41z;
    ora_green_rank *r;            /* one per emulated rank */
} ora_green;

typedef struct {
    ora_cfg cfg;
    int nranks;
    ora_rank *r;
    int *itbl;                    /* (-1:nproc_x, -1:nproc_y), -1 == MPI_PROC_NULL */
    void *snap;                   /* snapshot state (ora_snap.c) */
    ora_green *green;             /* NULL unless green_mode */
    char errmsg[512];
} ora_sim;

/* flat index helpers */
static inline size_t ora_idx3(const ora_rank *r, int k, int i, int j) {
    return (size_t)(k - r->kbeg_m) +
           (size_t)r->nzm * ((size_t)(i - r->ibeg_m) + (size_t)r->nxm * (size_t)(j - r->jbeg_m));
}
static inline size_t ora_idx2(const ora_rank *r, int i, int j) {
    return (size_t)(i - r->ibeg_m) + (size_t)r->nxm * (size_t)(j - r->jbeg_m);
}

/* ---------------------------------------------------------------- heavier model builders (ora_models.c) */
int ora_rdrmed2d(int ib, int ie, int kb, int ke, const char *fn, float *vol, char *err, size_t cap);
int ora_rdrmed3d(int ib, int ie, int jb, int je, int kb, int ke, const char *fn, float *vol, char *err, size_t cap);
int ora_vmodel_lgm(const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs, char *err, size_t cap);
int ora_vmodel_uni_rmed(const ora_cfg *c, const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs, char *err, size_t cap);
int ora_vmodel_lhm_rmed(const ora_cfg *c, const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs, char *err, size_t cap);
int ora_vmodel_lgm_rmed(const ora_cfg *c, const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs, char *err, size_t cap);
int ora_vmodel_grd(const ora_cfg *c, const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs, int with_rmed, char *err, size_t cap);
void ora_stabilize_absorber(const ora_cfg *c, ora_rank *r);

/* ---------------------------------------------------------------- life cycle / driver API */
/* base_dir: directory against which relative file names in the ini are resolved (the reference
 * resolves against the cwd of the run; tests pass the reference's top directory or a tmp dir). */
ora_sim *ora_create(const char *inf_path, const char *base_dir, int nm, int nproc_x_override,
                    int nproc_y_override, int nt_override);
ora_sim *ora_create_from_text(const char *inf_text, const char *base_dir, int nm,
                              int nproc_x_override, int nproc_y_override, int nt_override);
void ora_destroy(ora_sim *s);
const char *ora_last_error(void);

/* one iteration of main.f90:119-139 (without report/wav) */
void ora_update_stress(ora_sim *s);        /* kernel__update_stress + absorb__update_stress */
void ora_stressglut(ora_sim *s, int it);
void ora_comm_stress(ora_sim *s);
void ora_update_vel(ora_sim *s);           /* kernel__update_vel + absorb__update_vel */
void ora_bodyforce(ora_sim *s, int it);
void ora_comm_vel(ora_sim *s);
void ora_wav_store(ora_sim *s, int it);
void ora_vmax(ora_sim *s, float out[3]);   /* kernel__vmax + MPI max, times UC*M0 (m_report.f90:155) */
void ora_step(ora_sim *s, int it);         /* wav_store + everything above in main.f90 order */
/* run it=it0..it1 inclusive, recording vmax triplets at the top of every it with mod(it,ntdec_r)==0
 * into vm (3 floats each, capacity nvm); returns number recorded */
int ora_run(ora_sim *s, int it0, int it1, float *vm, int nvm);

/* accessors for tests (ctypes) */
int ora_nranks(const ora_sim *s);
const ora_cfg *ora_get_cfg(const ora_sim *s);
/* ints: 0 ibeg 1 iend 2 jbeg 3 jend 4 nxp 5 nyp 6 ibeg_k 7 iend_k 8 jbeg_k 9 jend_k 10 kbeg_k 11 kend_k
 *       12 nsrc 13 nst 14 idx 15 idy 16 nzm 17 nxm 18 nym 19 ibeg_m 20 jbeg_m 21 kbeg_m */
int ora_rank_int(const ora_sim *s, int rank, int what);
/* name: Vx..Sxy (ora_mp -> double), rho lam mu taup taus (float -> double); full _m box, (k,i,j) */
int ora_get_field(const ora_sim *s, int rank, const char *name, double *out);
int ora_set_field(ora_sim *s, int rank, const char *name, const double *in);
/* name: kfs kob kfs_top kfs_bot kob_top kob_bot kbeg_a ; (i,j) over _m box */
int ora_get_map(const ora_sim *s, int rank, const char *name, int *out);
/* gather the owned cells of every rank into a global (k=1..nz, i=1..nx, j=1..ny) array */
int ora_gather_field(const ora_sim *s, const char *name, double *out);
int ora_get_sources(const ora_sim *s, int rank, int *ijk /*3*nsrc*/, double *mo /*nsrc*/);
int ora_get_source_details(const ora_sim *s, int rank, double *mij /*6*nsrc*/, float *srcprm /*2*nsrc*/);
int ora_get_stations(const ora_sim *s, int rank, int *ijk /*3*nst*/, char *names /*9*nst*/);
int ora_get_wav(const ora_sim *s, int rank, float *out /*(ntw,3,nst)*/);
/* which: 0 velocity (ntw,3,nst), 1 displacement (ntw,3,nst), 2 stress (ntw,6,nst), 3 strain (ntw,6,nst) */
int ora_get_wav_product(const ora_sim *s, int rank, int which, float *out);
/* PML profile g?c/g?e: name gxc gxe gyc gye gzc gze -> (4,n) floats */
int ora_get_profile(const ora_sim *s, int rank, const char *name, float *out);
/* snapshots (ora_snap.c) */
int ora_snap_setup(ora_sim *s, const ora_ini *ini);
void ora_snap_write(ora_sim *s, int it);
void ora_snap_free(ora_sim *s);
int ora_snap_info(const ora_sim *s, int *info);
int ora_snap_coords(const ora_sim *s, float *x, float *y, float *z);
int ora_snap_nrec(const ora_sim *s, int q);
int ora_snap_rec(const ora_sim *s, int q, int rec, float *out, int *it0);
int ora_snap_max(const ora_sim *s, int q, float *out);
int ora_snap_medium(const ora_sim *s, int q, int which, float *out);
/* Green's function mode (m_green.f90): setup after wav_setup, store / source inside ora_step */
int ora_green_setup(ora_sim *s, const ora_ini *ini, const char *base, char *err, size_t cap);
void ora_green_store(ora_sim *s, int it);
void ora_green_source(ora_sim *s, int it);
void ora_green_free(ora_sim *s);
/* ints: 0 ng(rank) 1 ncmp 2 isrc 3 jsrc 4 ksrc 5 is_src(rank) 6 ntw */
int ora_green_int(const ora_sim *s, int rank, int what);
int ora_green_points(const ora_sim *s, int rank, int *ijk /*3*ng*/, int *gid /*ng*/);
int ora_get_green(const ora_sim *s, int rank, float *out /*(ncmp*ng, ntw) C order*/);
/* green__export (m_green.f90:553-604), wav_format = 'sac': <odir>/green/<stnm>/<title>__<gid>__<stnm>__<cmp>__<mij>__.sac */
int ora_write_green_sac(const ora_sim *s, const char *odir);
/* SAC output of all stations of all ranks (m_wav.f90:658-792); returns number of files */
int ora_write_sac(const ora_sim *s, const char *odir);

#ifdef __cplusplus
}
#endif
#endif
