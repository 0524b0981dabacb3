/*
 * oracle/psv.c -- CPU restatement of OpenSWPC's swpc_psv (2-D P-SV): setup chain + time step + halo exchange with the
 * MPI ranks emulated in one process + SAC writer.  TEST INFRASTRUCTURE ONLY (see psv.h / ora.h): nothing in the
 * product links or calls it.
 *
 * Follows src/swpc_psv/ of OpenSWPC 25.05.2 (citations are file:line under /root/reference/src/swpc_psv unless another
 * directory is named):
 *   main.f90:64-78    setup order: global__setup -> global__setup2 -> medium__setup -> kernel__setup -> source__setup ->
 *                     absorb__setup -> (snap__setup) -> wav__setup -> report__setup
 *   main.f90:95-113   time loop: report -> snap -> wav__store -> update_stress -> absorb_stress -> stressglut ->
 *                     comm_stress -> update_vel -> bodyforce -> absorb_vel -> comm_vel
 * Every temporary keeps its declared kind (real(SP) -> float, real(MP) -> ora_mp) and every expression the reference's
 * association; built with -ffp-contract=off and no fast-math.  PARITY UNPINNED: the reference holds no swpc_psv output.
 *
 * Scope: every vmodel_type but the compile-time 'user' plug-in (uni, lhm here; lgm, *_rmed, grd, grd_rmed through the `psv`
 * branches of ora_models.c), benchmark_mode, stabilize_pml, all moment / body-force source formats of m_source.f90 except the
 * slip-based ones, PML and Cerjan absorbers, station products v / u / stress / strain, SAC files, plane-wave mode, snapshots.
 */
#include "psv.h"

#include <complex.h>
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

double ora_r_earth(void);
double ora_pi(void);
float ora_rad2deg_s(float rad);

#define FLT_EPS 1.1920929e-07f /* epsilon(1.0) */
#define PSV_MAXNM 8

static char g_err[512] = "";
const char *psv_last_error(void) { return g_err; }
static void set_err(const char *m) {
    strncpy(g_err, m, sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
}
static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "[oracle/psv] out of memory\n"); abort(); }
    return p;
}
static void resolve_path(const char *base, const char *fn, char *out, size_t cap) {
    if (fn[0] == '/' || !base || !base[0]) snprintf(out, cap, "%s", fn);
    else snprintf(out, cap, "%s/%s", base, fn);
}
static int is_blank(const char *s) {
    for (; *s; s++) if (!isspace((unsigned char)*s)) return 0;
    return 1;
}
static int parse_sp(const char *line, float *v, int maxn) {   /* list-directed reals */
    int n = 0;
    const char *p = line;
    while (*p && n < maxn) {
        while (*p == ' ' || *p == '\t' || *p == ',') p++;
        if (!*p || *p == '\n' || *p == '\r') break;
        char tok[64];
        int l = 0;
        while (*p && !isspace((unsigned char)*p) && *p != ',' && l < 63) {
            char ch = *p++;
            tok[l++] = (ch == 'd' || ch == 'D') ? 'e' : ch;
        }
        tok[l] = 0;
        char *e;
        float x = strtof(tok, &e);
        if (e == tok) break;
        v[n++] = x;
    }
    return n;
}

/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    int benchmark_mode;
    char title[ORA_STRLEN], odir[ORA_STRLEN];
    int nproc_x, nx, nz, nt, ipad, kpad, na, nm;
    double dx, dz;
    float dt, xbeg, zbeg, tbeg, xend, zend, clon, clat, phi;
    char abc_type[16];
    float fq_min, fq_max, fq_ref, vcut;
    int pw_mode, bf_mode, earth_flattening;
    char stftype[16], stf_format[8], sdep_fit[8], fn_stf[ORA_STRLEN];
    int ntdec_w, ntdec_r, ntw, sw_v, sw_u, sw_stress, sw_strain;
    char wav_format[16], st_format[8], fn_stloc[ORA_STRLEN];
    float vmin, vmax, fmax, fcut, M0, UC, zeta;
    float ts[PSV_MAXNM], c1[PSV_MAXNM], c2[PSV_MAXNM], d1[PSV_MAXNM], d2;
    ora_mp rc40x, rc41x, rc40z, rc41z, rd40x, rd41x, rd40z, rd41z;
    float r20x, r20z;            /* m_absorb_p.f90:54 real(SP) */
    ora_mp dt_dxz;               /* m_source.f90:36 */
    ora_mp w40x, w40z, w41x, w41z; /* m_wav.f90:49, :127-130 */
    float evlo, evla, evdp, mxx0, mzz0, mxz0, fx0, fz0, otim, sx0, sy0;
    int exedate, tz_minutes;
} psv_cfg;

typedef struct {
    int myid, idx, nxp;
    int ibeg, iend, kbeg, kend, ibeg_m, iend_m, kbeg_m, kend_m, ibeg_k, iend_k, kbeg_k, kend_k;
    int nzm, nxm;
    size_t ncell;
    ora_mp *Vx, *Vz, *Sxx, *Szz, *Sxz;
    float *Rxx, *Rzz, *Rxz;           /* (m, k, i) over the memory box, m_kernel.f90:337-339 */
    float *rho, *lam, *mu, *taup, *taus;
    int *kfs, *kob, *kfs_top, *kfs_bot, *kob_top, *kob_bot, *kbeg_a;
    float *bddep;                     /* (i, 0:NBD) */
    float *xc, *zc;
    float *gxc, *gxe, *gzc, *gze;     /* (4, ibeg:iend) / (4, kbeg:kend) */
    int kbeg_min;
    float *aux[8];                    /* axVx azVx axVz azVz axSxx azSxz axSxz azSzz : (kbeg_min:kend, ibeg:iend) */
    float *gx_c, *gx_b, *gz_c, *gz_b;
    int nsrc;
    int *isrc, *ksrc;
    float *sx, *sz, *srcprm;
    ora_mp *mo, *mxx, *mzz, *mxz, *fx, *fz;
    int nst;
    int *ist, *kst;
    float *xst, *zst, *stlo, *stla;
    char (*stnm)[9];
    float *wav[4];                    /* v u (ntw,2,nst); stress strain (ntw,3,nst) */
    float *ux, *uz, *exx, *ezz, *exz;
    ora_mp *sbuf_ip, *sbuf_im, *rbuf_ip, *rbuf_im;
} psv_rank;

/* snapshots, m_snap.f90: three products over the decimated xz plane, each with two components */
typedef struct {
    int sw[3];                        /* xz_ps, xz_v, xz_u */
    int idec, kdec, ntdec_s, nxs, nzs;
    char snp_format[8];
    float *xsnp, *zsnp;
    float *medium;                    /* (3, nzs, nxs): rho lambda mu, summed over ranks */
    float *buf_u;                     /* (2, nzs, nxs): running displacement, m_snap.f90:588-612 */
    int nrec[3], cap[3];
    int *it0[3];
    float *rec[3];                    /* records (nrec, 2, nzs, nxs) of the netCDF files / native streams */
} psv_snap;

struct psv_sim {
    psv_cfg cfg;
    int nranks;
    psv_rank *r;
    psv_snap snap;
};

static inline size_t IX(const psv_rank *r, int k, int i) { return (size_t)(k - r->kbeg_m) + (size_t)r->nzm * (size_t)(i - r->ibeg_m); }
static inline size_t AX(const psv_rank *r, int k, int i) { return (size_t)(k - r->kbeg_min) + (size_t)(r->kend - r->kbeg_min + 1) * (size_t)(i - r->ibeg); }

/* ------------------------------------------------------------------------------------------------------------ */
/* m_global.f90:116-153 global__readprm, :183-185                                                                 */
static void global_setup(psv_cfg *c, const ora_ini *ini) {
    ora_readini_l(ini, "benchmark_mode", &c->benchmark_mode, 0);
    ora_readini_c(ini, "title", c->title, "swpc_psv");
    ora_readini_i(ini, "nproc_x", &c->nproc_x, 1);
    ora_readini_i(ini, "nx", &c->nx, 256);
    ora_readini_i(ini, "nz", &c->nz, 256);
    ora_readini_i(ini, "nt", &c->nt, 1000);
    ora_readini_i(ini, "ipad", &c->ipad, 0);
    ora_readini_i(ini, "kpad", &c->kpad, 0);
    ora_readini_c(ini, "odir", c->odir, "./out");
    if (c->benchmark_mode) {
        c->dx = 0.5f; c->dz = 0.5f; c->dt = 0.04f; c->na = 20;
        c->xbeg = -((float)c->nx / 2.0f * (float)c->dx);
        c->zbeg = -30 * (float)c->dz;
        c->tbeg = 0.0f; c->clon = 139.7604f; c->clat = 35.7182f; c->phi = 0.0f;
        strcpy(c->abc_type, "pml");
    } else {
        ora_readini_d(ini, "dx", &c->dx, 0.5);
        ora_readini_d(ini, "dz", &c->dz, 0.5);
        ora_readini_s(ini, "dt", &c->dt, 0.01f);
        ora_readini_i(ini, "na", &c->na, 20);
        ora_readini_s(ini, "xbeg", &c->xbeg, -(float)(c->nx / 2) * (float)c->dx);
        ora_readini_s(ini, "zbeg", &c->zbeg, -30 * (float)c->dz);
        ora_readini_s(ini, "tbeg", &c->tbeg, 0.0f);
        ora_readini_s(ini, "clon", &c->clon, 139.7604f);
        ora_readini_s(ini, "clat", &c->clat, 35.7182f);
        ora_readini_s(ini, "phi", &c->phi, 0.0f);
        char abc[ORA_STRLEN];
        ora_readini_c(ini, "abc_type", abc, "pml");
        strncpy(c->abc_type, abc, sizeof(c->abc_type) - 1);
    }
    c->xend = c->xbeg + c->nx * (float)c->dx;
    c->zend = c->zbeg + c->nz * (float)c->dz;
    c->UC = 1e-12f; /* m_global.f90:28  UC = 10.0**(-12) */
}

/* m_global.f90:189-310 global__setup2 for rank myid */
static void rank_geometry(const psv_cfg *c, psv_rank *r, int myid) {
    memset(r, 0, sizeof(*r));
    r->myid = myid;
    r->idx = myid % c->nproc_x;   /* set_mpi_table :452 */
    const int nx = c->nx, np = c->nproc_x, mx = nx % np, proc_x = myid;
    /* :201-208 -- note the "+ 1" where the ibeg/iend block below has "- 1"; nxp is not used by the hot path */
    r->nxp = (proc_x <= np - mx + 1) ? (nx - mx) / np : (nx - mx) / np + 1;
    if (proc_x <= np - mx - 1) {   /* :232-238 */
        r->ibeg = proc_x * (nx - mx) / np + 1;
        r->iend = (proc_x + 1) * (nx - mx) / np;
    } else {
        r->ibeg = proc_x * ((nx - mx) / np + 1) - (np - mx) + 1;
        r->iend = (proc_x + 1) * ((nx - mx) / np + 1) - (np - mx);
    }
    r->kbeg = 1;
    r->kend = c->nz;
    r->ibeg_m = r->ibeg - 3; r->iend_m = r->iend + 3 + c->ipad;
    r->kbeg_m = r->kbeg - 3; r->kend_m = r->kend + 3 + c->kpad;
    r->nzm = r->kend_m - r->kbeg_m + 1;
    r->nxm = r->iend_m - r->ibeg_m + 1;
    r->ncell = (size_t)r->nzm * r->nxm;
    r->xc = (float *)xcalloc((size_t)r->nxm, sizeof(float));
    r->zc = (float *)xcalloc((size_t)r->nzm, sizeof(float));
    for (int i = r->ibeg_m; i <= r->iend_m; i++) r->xc[i - r->ibeg_m] = ora_i2x(i, c->xbeg, (float)c->dx);
    for (int k = r->kbeg_m; k <= r->kend_m; k++) r->zc[k - r->kbeg_m] = ora_i2x(k, c->zbeg, (float)c->dz);
    r->kbeg_a = (int *)xcalloc((size_t)r->nxm, sizeof(int));   /* :259-266 */
    for (int i = r->ibeg_m; i <= r->iend_m; i++)
        r->kbeg_a[i - r->ibeg_m] = (i <= c->na || nx - c->na + 1 <= i) ? r->kbeg : r->kend - c->na + 1;
    r->ibeg_k = r->ibeg; r->iend_k = r->iend; r->kbeg_k = r->kbeg; r->kend_k = r->kend;   /* :270-292 */
    if (!strcmp(c->abc_type, "pml")) {
        const int na = c->na;
        if (r->iend <= na) r->ibeg_k = r->iend + 1;
        else if (r->ibeg <= na) r->ibeg_k = na + 1;
        if (r->ibeg >= nx - na + 1) r->iend_k = r->ibeg - 1;
        else if (r->iend >= nx - na + 1) r->iend_k = nx - na;
        r->kend_k = c->nz - na;
    }
    const size_t nb = (size_t)3 * c->nz;   /* :214-215 */
    r->sbuf_ip = (ora_mp *)xcalloc(nb, sizeof(ora_mp)); r->sbuf_im = (ora_mp *)xcalloc(nb, sizeof(ora_mp));
    r->rbuf_ip = (ora_mp *)xcalloc(nb, sizeof(ora_mp)); r->rbuf_im = (ora_mp *)xcalloc(nb, sizeof(ora_mp));
}

/* ------------------------------------------------------------------------------------------------------------ */
/* m_vmodel_uni.f90:21-138 */
static void vmodel_uni(const ora_ini *ini, psv_rank *r, float *qp, float *qs) {
    float vp0, vs0, rho0, qp0, qs0, topo0;
    int use_munk, ef;
    ora_readini_s(ini, "vp0", &vp0, 5.0f);
    ora_readini_s(ini, "vs0", &vs0, vp0 / sqrtf(3.0f));
    ora_readini_s(ini, "rho0", &rho0, 2.7f);
    ora_readini_s(ini, "qp0", &qp0, 1000000.0f);
    ora_readini_s(ini, "qs0", &qs0, 1000000.0f);
    ora_readini_s(ini, "topo0", &topo0, 0.0f);
    ora_readini_l(ini, "munk_profile", &use_munk, 0);
    ora_readini_l(ini, "earth_flattening", &ef, 0);
    const double RE = ora_r_earth();
    for (int i = r->ibeg_m; i <= r->iend_m; i++) {
        r->bddep[i - r->ibeg_m] = topo0;
        for (int k = r->kbeg_m; k <= r->kend_m; k++) {
            const float zc = r->zc[k - r->kbeg_m];
            float zs = zc, Cv = 1.0f;
            if (ef) { zs = (float)(RE - RE * exp(-(double)zc / RE)); Cv = (float)exp((double)zc / RE); }
            const size_t n = IX(r, k, i);
            float vp1, vs1;
            if (zs > topo0) {
                vp1 = Cv * vp0; vs1 = Cv * vs0;
                r->rho[n] = rho0; r->mu[n] = rho0 * vs1 * vs1; r->lam[n] = rho0 * (vp1 * vp1 - 2 * vs1 * vs1);
                qp[n] = qp0; qs[n] = qs0;
            } else if (zc > 0.0f) {
                vp1 = Cv * ora_seawater_vel(zs, use_munk); vs1 = 0.0f;
                r->rho[n] = 1.0f; r->mu[n] = r->rho[n] * vs1 * vs1; r->lam[n] = r->rho[n] * (vp1 * vp1 - 2 * vs1 * vs1);
                qp[n] = 1000000.0f; qs[n] = 1000000.0f;
            } else {
                vp1 = 0.0f; vs1 = 0.0f;
                r->rho[n] = 0.001f; r->mu[n] = r->rho[n] * vs1 * vs1; r->lam[n] = r->rho[n] * (vp1 * vp1 - 2 * vs1 * vs1);
                qp[n] = 10.0f; qs[n] = 10.0f;
            }
        }
    }
    for (int b = 1; b <= ORA_NBD; b++)
        for (int i = 0; i < r->nxm; i++) r->bddep[(size_t)b * r->nxm + i] = -9999.0f;
}

/* m_vmodel_lhm.f90:22-155 */
static int vmodel_lhm(const ora_ini *ini, const char *base, psv_rank *r, float vcut, float *qp, float *qs) {
    char fn[ORA_STRLEN], path[2 * ORA_STRLEN];
    int use_munk, ef;
    ora_readini_c(ini, "fn_lhm", fn, "");
    ora_readini_l(ini, "munk_profile", &use_munk, 0);
    ora_readini_l(ini, "earth_flattening", &ef, 0);
    resolve_path(base, fn, path, sizeof(path));
    FILE *fp = fopen(path, "r");
    if (!fp) { char m[700]; snprintf(m, sizeof(m), "vmodel_lhm: cannot open %s", path); set_err(m); return -1; }
    float depth[256], rho0[256], vp0[256], vs0[256], qp0[256], qs0[256];
    int nl = 0;
    char line[512];
    while (fgets(line, sizeof(line), fp) && nl < 256) {
        char *p = line;
        while (*p == ' ' || *p == '\t') p++;
        if (is_blank(p) || *p == '#') continue;
        float v[6];
        if (parse_sp(p, v, 6) < 6) continue;
        depth[nl] = v[0]; rho0[nl] = v[1]; vp0[nl] = v[2]; vs0[nl] = v[3]; qp0[nl] = v[4]; qs0[nl] = v[5];
        nl++;
    }
    fclose(fp);
    for (int l = nl - 2; l >= 0; l--)   /* :85-94 */
        if ((vp0[l] < vcut || vs0[l] < vcut) && (vp0[l] > 0 && vs0[l] > 0)) {
            vp0[l] = vp0[l + 1]; vs0[l] = vs0[l + 1]; rho0[l] = rho0[l + 1]; qp0[l] = qp0[l + 1]; qs0[l] = qs0[l + 1];
        }
    for (int i = 0; i < r->nxm; i++) r->bddep[i] = depth[0];
    const double RE = ora_r_earth();
    for (int k = r->kbeg_m; k <= r->kend_m; k++) {
        const float zc = r->zc[k - r->kbeg_m];
        float zs = zc, Cv = 1.0f;
        if (ef) { zs = (float)(RE - RE * exp(-(double)zc / RE)); Cv = (float)exp((double)zc / RE); }
        float rho1 = 0, vp1 = 0, vs1 = 0, qp1 = 0, qs1 = 0;
        if (zs < depth[0]) {
            if (zs < 0.0f) { rho1 = 0.001f; vp1 = 0.0f; vs1 = 0.0f; qp1 = 10.0f; qs1 = 10.0f; }
            else { rho1 = 1.0f; vp1 = Cv * ora_seawater_vel(zs, use_munk); vs1 = 0.0f; qp1 = 1000000.0f; qs1 = 1000000.0f; }
        } else {
            for (int l = 0; l < nl; l++)
                if (zs >= depth[l]) { rho1 = rho0[l]; vp1 = Cv * vp0[l]; vs1 = Cv * vs0[l]; qp1 = qp0[l]; qs1 = qs0[l]; }
        }
        const float mu1 = rho1 * vs1 * vs1, lam1 = rho1 * (vp1 * vp1 - 2 * vs1 * vs1);
        for (int i = r->ibeg_m; i <= r->iend_m; i++) {
            const size_t n = IX(r, k, i);
            r->rho[n] = rho1; r->mu[n] = mu1; r->lam[n] = lam1; qp[n] = qp1; qs[n] = qs1;
        }
    }
    for (int b = 1; b <= ORA_NBD; b++)
        for (int i = 0; i < r->nxm; i++) r->bddep[(size_t)b * r->nxm + i] = -9999.0f;
    return 0;
}

/* m_medium.f90:260-292 surface_detection.  kfs_top is assigned twice and kfs_bot never (:281-282): kfs_bot keeps its
 * allocation value, taken as 0 (Q-psv 1 in psv.h). */
static void surface_detection(psv_rank *r) {
    for (int i = 0; i < r->nxm; i++) { r->kfs[i] = r->kbeg - 1; r->kob[i] = r->kbeg - 1; }
    for (int i = r->ibeg - 1; i <= r->iend + 2; i++)
        for (int k = r->kbeg; k <= r->kend - 1; k++) {
            const size_t n0 = IX(r, k, i), n1 = IX(r, k + 1, i);
            if (fabsf(r->mu[n0]) < FLT_EPS && fabsf(r->mu[n1]) > FLT_EPS) r->kob[i - r->ibeg_m] = k;
            if (fabsf(r->lam[n0]) < FLT_EPS && fabsf(r->lam[n1]) > FLT_EPS) r->kfs[i - r->ibeg_m] = k;
        }
    for (int i = r->ibeg; i <= r->iend; i++) {
        int fmin_ = 1 << 30, fmax_ = -(1 << 30), omin = 1 << 30, omax = -(1 << 30);
        for (int ii = i - 2; ii <= i + 3; ii++) {
            const int a = r->kfs[ii - r->ibeg_m], b = r->kob[ii - r->ibeg_m];
            if (a < fmin_) fmin_ = a;
            if (a > fmax_) fmax_ = a;
            if (b < omin) omin = b;
            if (b > omax) omax = b;
        }
        const int q = i - r->ibeg_m;
        r->kfs_top[q] = (fmin_ - 2 > r->kbeg) ? fmin_ - 2 : r->kbeg;
        r->kfs_top[q] = (fmax_ + 2 < r->kend) ? fmax_ + 2 : r->kend;   /* sic */
        r->kob_top[q] = (omin - 2 > r->kbeg) ? omin - 2 : r->kbeg;
        r->kob_bot[q] = (omax + 2 < r->kend) ? omax + 2 : r->kend;
    }
}

/* m_medium.f90:36-205 */
static int medium_setup(psv_sim *s, psv_rank *r, const ora_ini *ini, const char *base, float *vmin1, float *vmax1) {
    psv_cfg *c = &s->cfg;
    const size_t nc = r->ncell;
    r->rho = (float *)xcalloc(nc, sizeof(float)); r->lam = (float *)xcalloc(nc, sizeof(float)); r->mu = (float *)xcalloc(nc, sizeof(float));
    r->taup = (float *)xcalloc(nc, sizeof(float)); r->taus = (float *)xcalloc(nc, sizeof(float));
    r->kfs = (int *)xcalloc((size_t)r->nxm, sizeof(int)); r->kob = (int *)xcalloc((size_t)r->nxm, sizeof(int));
    r->kfs_top = (int *)xcalloc((size_t)r->nxm, sizeof(int)); r->kfs_bot = (int *)xcalloc((size_t)r->nxm, sizeof(int));
    r->kob_top = (int *)xcalloc((size_t)r->nxm, sizeof(int)); r->kob_bot = (int *)xcalloc((size_t)r->nxm, sizeof(int));
    r->bddep = (float *)xcalloc((size_t)r->nxm * (ORA_NBD + 1), sizeof(float));
    const int nm = c->nm, na = c->na, nx = c->nx, nz = c->nz;
    if (c->benchmark_mode) {   /* :53-72 */
        c->fq_min = 0.05f; c->fq_max = 5.0f; c->fq_ref = 1.0f;
        for (int k = r->kbeg_m; k <= r->kend_m; k++) {
            const float zc = r->zc[k - r->kbeg_m];
            float rr, mm, ll;
            if (zc < 0.0f) { rr = 0.001f; mm = 0.0f; ll = 0.0f; }
            else { rr = 2.7f; mm = 2.7f * 3.5f * 3.5f; ll = 2.7f * 3.5f * 3.5f; }
            for (int i = r->ibeg_m; i <= r->iend_m; i++) {
                const size_t n = IX(r, k, i);
                r->rho[n] = rr; r->mu[n] = mm; r->lam[n] = ll; r->taup[n] = 1e10f; r->taus[n] = 1e10f;
            }
        }
    } else {
        ora_readini_s(ini, "fq_min", &c->fq_min, 0.05f);
        ora_readini_s(ini, "fq_max", &c->fq_max, 5.00f);
        ora_readini_s(ini, "fq_ref", &c->fq_ref, 1.00f);
        char vt[ORA_STRLEN];
        ora_readini_c(ini, "vmodel_type", vt, "uni");
        ora_readini_s(ini, "vcut", &c->vcut, 0.0f);
        if (!strcmp(vt, "uni")) vmodel_uni(ini, r, r->taup, r->taus);
        else if (!strcmp(vt, "lhm")) { if (vmodel_lhm(ini, base, r, c->vcut, r->taup, r->taus)) return -1; }
        else if (!strcmp(vt, "lgm") || !strcmp(vt, "uni_rmed") || !strcmp(vt, "lhm_rmed") || !strcmp(vt, "lgm_rmed") || !strcmp(vt, "grd") || !strcmp(vt, "grd_rmed")) {
            /* swpc_psv/m_vmodel_{lgm,uni_rmed,lhm_rmed,lgm_rmed}.f90 are the 3-D builders with the y axis removed and a handful
             * of differing expressions (ora_models.c, the `psv` branches): run them on a one-plane view of this section */
            ora_cfg c3;
            ora_rank r3;
            memset(&c3, 0, sizeof(c3));
            memset(&r3, 0, sizeof(r3));
            c3.nx = c->nx; c3.ny = 0; c3.nz = c->nz; c3.dx = c->dx; c3.dy = c->dx; c3.dz = c->dz; c3.dt = c->dt; c3.na = c->na;
            c3.xbeg = c->xbeg; c3.zbeg = c->zbeg; c3.clon = c->clon; c3.clat = c->clat; c3.phi = c->phi;
            r3.psv = 1;
            r3.ibeg_m = r->ibeg_m; r3.iend_m = r->iend_m; r3.jbeg_m = 1; r3.jend_m = 1; r3.kbeg_m = r->kbeg_m; r3.kend_m = r->kend_m;
            r3.nzm = r->nzm; r3.nxm = r->nxm; r3.nym = 1; r3.ncell_m = r->ncell;
            r3.rho = r->rho; r3.lam = r->lam; r3.mu = r->mu; r3.bddep = r->bddep; r3.zc = r->zc; r3.xc = r->xc;
            char m[600] = "";
            int rc;
            if (!strcmp(vt, "lgm")) rc = ora_vmodel_lgm(ini, base, &r3, c->vcut, r->taup, r->taus, m, sizeof(m));
            else if (!strcmp(vt, "uni_rmed")) rc = ora_vmodel_uni_rmed(&c3, ini, base, &r3, c->vcut, r->taup, r->taus, m, sizeof(m));
            else if (!strcmp(vt, "lhm_rmed")) rc = ora_vmodel_lhm_rmed(&c3, ini, base, &r3, c->vcut, r->taup, r->taus, m, sizeof(m));
            else if (!strcmp(vt, "lgm_rmed")) rc = ora_vmodel_lgm_rmed(&c3, ini, base, &r3, c->vcut, r->taup, r->taus, m, sizeof(m));
            else rc = ora_vmodel_grd(&c3, ini, base, &r3, c->vcut, r->taup, r->taus, !strcmp(vt, "grd_rmed"), m, sizeof(m));
            if (rc) { set_err(m); return -1; }
        }
        else { char m[400]; snprintf(m, sizeof(m), "swpc_psv vmodel_type '%s' is outside the restated scope (everything but the compile-time 'user' plug-in)", vt); set_err(m); return -1; }
    }
#define CP5(dst, src) do { r->rho[dst] = r->rho[src]; r->lam[dst] = r->lam[src]; r->mu[dst] = r->mu[src]; \
                           r->taup[dst] = r->taup[src]; r->taus[dst] = r->taus[src]; } while (0)
    /* :118-148 homogenize the absorber (only where the source column lives in this rank's memory box) */
    if (na + 1 >= r->ibeg_m && na + 1 <= r->iend_m)
        for (int i = r->ibeg_m; i <= na; i++)
            for (int k = r->kbeg_m; k <= r->kend_m; k++) CP5(IX(r, k, i), IX(r, k, na + 1));
    if (nx - na >= r->ibeg_m && nx - na <= r->iend_m)
        for (int i = (nx - na + 1 > r->ibeg_m ? nx - na + 1 : r->ibeg_m); i <= r->iend_m; i++)
            for (int k = r->kbeg_m; k <= r->kend_m; k++) CP5(IX(r, k, i), IX(r, k, nx - na));
    for (int i = r->ibeg_m; i <= r->iend_m; i++)
        for (int k = nz - na + 1; k <= r->kend_m; k++) CP5(IX(r, k, i), IX(r, nz - na, i));
#undef CP5
    ora_visco_set_relaxtime(nm, c->ts, c->fq_min, c->fq_max);   /* :151-152 */
    float zeta = ora_visco_constq_zeta(nm, c->fq_min, c->fq_max, c->ts);
    if (c->benchmark_mode) zeta = 0.0f;
    c->zeta = zeta;
    for (size_t n = 0; n < nc; n++) { r->taup[n] = nm * zeta / r->taup[n]; r->taus[n] = nm * zeta / r->taus[n]; }
    if (nm > 0) {   /* relaxed_medium :181-204 + visco_chi (src/shared/m_fdtool.f90:730-753) */
        const float omega = (float)(2 * ora_pi() * (double)c->fq_ref);
        float complex cc = 0.0f;
        for (int im = 0; im < nm; im++) {
            double complex w = I * (double)omega * (double)c->ts[im];
            double complex q = w / (1.0 - w);
            cc = cc + (float complex)q;
        }
        cc = (crealf(cc) / (float)nm) + (cimagf(cc) / (float)nm) * I;
        for (size_t n = 0; n < nc; n++) {
            const float rho_beta2 = r->mu[n], rho_alpha2 = r->lam[n] + 2 * r->mu[n];
            const float complex zs = 1.0f - cc * r->taus[n], zp = 1.0f - cc * r->taup[n];
            const float chi_mu = 1.0f / crealf(1.0f / csqrtf(zs)), chi_lam = 1.0f / crealf(1.0f / csqrtf(zp));
            r->mu[n] = rho_beta2 / (chi_mu * chi_mu);
            r->lam[n] = rho_alpha2 / (chi_lam * chi_lam) - 2 * r->mu[n];
        }
    }
    surface_detection(r);
    float vmx = -1.0f, vmn = 1e30f;   /* velocity_minmax :294-318 */
    for (int i = r->ibeg; i <= r->iend; i++)
        for (int k = r->kfs[i - r->ibeg_m] + 1; k <= r->kend; k++) {
            const size_t n = IX(r, k, i);
            const float vp = sqrtf((r->lam[n] + 2 * r->mu[n]) / r->rho[n]), vs = sqrtf(r->mu[n] / r->rho[n]);
            if (vp > vmx) vmx = vp;
            if (vs < FLT_EPS) continue;
            if (vs < vmn) vmn = vs;
        }
    *vmin1 = vmn; *vmax1 = vmx;
    return 0;
}

/* m_medium.f90:309-366 stabilize_absorber, called after velocity_minmax's all-reduce (:169-174): vmax is the global maximum */
static void stabilize_absorber(const psv_cfg *c, psv_rank *r) {
    const float vmin_pml = c->vmax * 0.4f;   /* V_DYNAMIC_RANGE */
    const int LV_THICK = 20;
    for (int i = r->ibeg - 1; i <= r->iend + 1; i++) {
        int ktop = 1 << 30;
        for (int ii = i - 2; ii <= i + 2; ii++) if (r->kbeg_a[ii - r->ibeg_m] < ktop) ktop = r->kbeg_a[ii - r->ibeg_m];
        int k = ktop;
        while (k <= r->kend) {
            if (r->lam[IX(r, k, i)] < r->lam[IX(r, k - 1, i)] || r->mu[IX(r, k, i)] < r->mu[IX(r, k - 1, i)]) {
                int k2;
                for (k2 = k + 1; k2 <= r->kend; k2++)   /* the bottom of the low-velocity layer */
                    if (r->lam[IX(r, k2, i)] > r->lam[IX(r, k2 - 1, i)] || r->mu[IX(r, k2, i)] > r->mu[IX(r, k2 - 1, i)]) break;
                if (k2 - k <= LV_THICK) {
                    const size_t m = IX(r, k - 1, i);
                    for (int q = k; q <= k2 - 1; q++) {
                        const size_t n = IX(r, q, i);
                        r->rho[n] = r->rho[m]; r->lam[n] = r->lam[m]; r->mu[n] = r->mu[m]; r->taup[n] = r->taup[m]; r->taus[n] = r->taus[m];
                    }
                    k = k2 - 1;
                }
            }
            k = k + 1;
        }
    }
    for (int i = r->ibeg - 1; i <= r->iend + 1; i++) {
        int ktop = 1 << 30;
        for (int ii = i - 2; ii <= i + 2; ii++) if (r->kbeg_a[ii - r->ibeg_m] < ktop) ktop = r->kbeg_a[ii - r->ibeg_m];
        for (int k = ktop; k <= r->kend; k++) {
            const size_t n = IX(r, k, i);
            float vs = sqrtf(r->mu[n] / r->rho[n]);
            if (vs < FLT_EPS) continue;
            if (vs < vmin_pml) {
                vs = vmin_pml;
                const float vp = vs * sqrtf(3.0f);
                r->lam[n] = r->rho[n] * (vp * vp - 2 * (vs * vs));
                r->mu[n] = r->rho[n] * (vs * vs);
            }
        }
    }
}

/* m_kernel.f90:30-74 + memory_allocate :329-343 */
static void kernel_setup(psv_sim *s) {
    psv_cfg *c = &s->cfg;
    const ora_mp dx = (ora_mp)c->dx, dz = (ora_mp)c->dz;
    c->rc40x = (ora_mp)17.0 / (ora_mp)16.0 / dx; c->rc40z = (ora_mp)17.0 / (ora_mp)16.0 / dz;
    c->rc41x = (ora_mp)1.0 / (ora_mp)48.0 / dx;  c->rc41z = (ora_mp)1.0 / (ora_mp)48.0 / dz;
    c->rd40x = -(ora_mp)1.0 / (ora_mp)16.0 / dx; c->rd40z = -(ora_mp)1.0 / (ora_mp)16.0 / dz;
    c->rd41x = -(ora_mp)1.0 / (ora_mp)48.0 / dx; c->rd41z = -(ora_mp)1.0 / (ora_mp)48.0 / dz;
    const int nm = c->nm;
    const float dt = c->dt;
    c->d2 = 0.0f;
    if (nm > 0) {
        for (int m = 0; m < nm; m++) { c->c1[m] = (2 * c->ts[m] - dt) / (2 * c->ts[m] + dt); c->c2[m] = (2) / (2 * c->ts[m] + dt) / nm; }
        float sum = 0.0f;
        for (int m = 0; m < nm; m++) sum += dt / (2 * c->ts[m] - dt);
        c->d2 = sum / nm;
        for (int m = 0; m < nm; m++) c->d1[m] = 2 * c->ts[m] / (2 * c->ts[m] - dt);
    }
    for (int q = 0; q < s->nranks; q++) {
        psv_rank *r = &s->r[q];
        const size_t nc = r->ncell;
        r->Vx = (ora_mp *)xcalloc(nc, sizeof(ora_mp)); r->Vz = (ora_mp *)xcalloc(nc, sizeof(ora_mp));
        r->Sxx = (ora_mp *)xcalloc(nc, sizeof(ora_mp)); r->Szz = (ora_mp *)xcalloc(nc, sizeof(ora_mp)); r->Sxz = (ora_mp *)xcalloc(nc, sizeof(ora_mp));
        if (nm > 0) {
            r->Rxx = (float *)xcalloc(nc * nm, sizeof(float)); r->Rzz = (float *)xcalloc(nc * nm, sizeof(float)); r->Rxz = (float *)xcalloc(nc * nm, sizeof(float));
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* pw_setup, m_source.f90:656-785: plane P / SV wave as the initial condition over the whole memory box.  pw_strike and
 * pw_rake are read but then forced to 90 degrees (:684-686).                                                        */
static int pw_setup(psv_sim *s, const ora_ini *ini) {
    psv_cfg *c = &s->cfg;
    float pw_ztop, pw_zlen, strike, dip, rake;
    char ps[ORA_STRLEN], tmp[ORA_STRLEN];
    ora_readini_s(ini, "pw_ztop", &pw_ztop, 1e30f);
    if (!(pw_ztop < c->zend)) { set_err("assert: pw_ztop < zend (m_source.f90:672)"); return -1; }
    ora_readini_s(ini, "pw_zlen", &pw_zlen, -1.0f);
    if (!(pw_zlen > 0.0f)) { set_err("assert: pw_zlen > 0 (m_source.f90:675)"); return -1; }
    ora_readini_c(ini, "pw_ps", ps, "");
    const int is_p = (ps[0] == 'p' || ps[0] == 'P'), is_s = (ps[0] == 's' || ps[0] == 'S');
    if (!(is_p || is_s) || ps[1]) { set_err("assert: pw_ps must be p or s (m_source.f90:678)"); return -1; }
    ora_readini_s(ini, "pw_strike", &strike, 0.0f);
    ora_readini_s(ini, "pw_dip", &dip, 0.0f);
    ora_readini_s(ini, "pw_rake", &rake, 0.0f);
    strike = ora_deg2rad(90.0f); dip = ora_deg2rad(dip); rake = ora_deg2rad(90.0f);
    ora_readini_c(ini, "stftype", tmp, "kupper");
    snprintf(c->stftype, sizeof(c->stftype), "%.15s", tmp);
    const float sd = sinf(dip), cd = cosf(dip), sf = sinf(strike), cf = cosf(strike), sl = sinf(rake), cl = cosf(rake);
    const float c2d = cosf(2 * dip);
    const float prm[2] = {0.0f, pw_zlen};
    const float dt = c->dt;
    const ora_mp dx = (ora_mp)c->dx, dz = (ora_mp)c->dz;
    const char *st = c->stftype;
    float fcut = 0.0f;
    for (int q = 0; q < s->nranks; q++) {
        psv_rank *r = &s->r[q];
        for (int i = r->ibeg_m; i <= r->iend_m; i++)
            for (int k = r->kbeg_m; k <= r->kend_m; k++) {
                const size_t n = IX(r, k, i);
                const float la0 = r->lam[n], mu0 = r->mu[n];
                const float v = is_p ? sqrtf((la0 + 2 * mu0) / r->rho[n]) : sqrtf(mu0 / r->rho[n]);
                if (v < FLT_EPS) continue;
                const float x0 = (float)(c->xbeg + (i - 0.5f) * dx);
                const float z0 = (float)(c->zbeg + (k - 0.5f) * dz - pw_ztop);
                const float x1 = (float)(x0 + dx / 2.0f), z1 = (float)(z0 + dz / 2.0f);
                const float stf_ii = ora_momentrate(sd * sf * x0 + cd * z0, st, prm);
                const float stf_vx = ora_momentrate(sd * sf * x1 + cd * z0 + dt / 2.0f * v, st, prm);
                const float stf_vz = ora_momentrate(sd * sf * x0 + cd * z1 + dt / 2.0f * v, st, prm);
                const float stf_xz = ora_momentrate(sd * sf * x1 + cd * z1, st, prm);
                if (is_p) {
                    r->Vx[n] = -sd * sf * stf_vx;
                    r->Vz[n] = -cd * stf_vz;
                    r->Sxx[n] = -(la0 + 2 * mu0 * sd * sd * sf * sf) * stf_ii / v;
                    r->Szz[n] = -(la0 + 2 * mu0 * cd * cd) * stf_ii / v;
                    r->Sxz[n] = -2 * mu0 * sd * cd * sf * stf_xz / v;
                } else {
                    r->Vx[n] = (cl * cf + sl * cd * sf) * stf_vx;
                    r->Vz[n] = -sl * sd * stf_vz;
                    r->Sxx[n] = 2 * mu0 * sd * sf * (cl * cf + sl * cd * sf) * stf_ii / v;
                    r->Szz[n] = -2 * mu0 * cd * sl * sd * stf_ii / v;
                    r->Sxz[n] = mu0 * (cl * cd * cf + sl * c2d * sf) * stf_xz / v;
                }
            }
        /* wavelength condition :764-783 (MPI_MAX over the ranks) */
        const int i = ora_x2i((c->xbeg + c->xend) / 2, c->xbeg, (float)c->dx), k = ora_x2i(pw_ztop, c->zbeg, (float)c->dz);
        if (r->ibeg <= i && i <= r->iend) {
            const size_t n = IX(r, k, i);
            const float v = is_p ? sqrtf((r->lam[n] + 2 * r->mu[n]) / r->rho[n]) : sqrtf(r->mu[n] / r->rho[n]);
            if (v / pw_zlen > fcut) fcut = v / pw_zlen;
        }
    }
    c->fcut = fcut;
    c->fmax = fcut * 2.0f;
    return 0;
}

/* Horizontal zero-derivative boundary of the plane-wave mode: linear extrapolation into the first column outside the
 * model on the outer ranks, ahead of the PML update (m_absorb_p.f90:114-155 stresses, :287-326 velocities)           */
static void pw_edges(const psv_cfg *c, psv_rank *r, int nproc_x, int stress_fields) {
    ora_mp *f[3] = {stress_fields ? r->Sxx : r->Vx, stress_fields ? r->Szz : r->Vz, stress_fields ? r->Sxz : NULL};
    const int nf = stress_fields ? 3 : 2;
    for (int side = 0; side < 2; side++) {
        if (side == 0 && r->idx != 0) continue;
        if (side == 1 && r->idx != nproc_x - 1) continue;
        const int dst = side == 0 ? 0 : c->nx + 1, s1 = side == 0 ? 1 : c->nx, s2 = side == 0 ? 2 : c->nx - 1;
        for (int k = r->kbeg_a[s1 - r->ibeg_m]; k <= r->kend; k++)
            for (int q = 0; q < nf; q++) f[q][IX(r, k, dst)] = 2 * f[q][IX(r, k, s1)] - f[q][IX(r, k, s2)];
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* m_source.f90:41-258 (+ source__grid_moment :260-474, source__grid_bodyforce :476-548)                           */
static int source_setup(psv_sim *s, const ora_ini *ini, const char *base) {
    psv_cfg *c = &s->cfg;
    ora_readini_l(ini, "pw_mode", &c->pw_mode, 0);
    if (c->pw_mode && !c->benchmark_mode) { /* :68-76 */
        if (pw_setup(s, ini)) return -1;
        for (int q = 0; q < s->nranks; q++) s->r[q].nsrc = 0;
        c->M0 = 1.0f / c->UC;
        return 0;
    }
    ora_readini_l(ini, "bf_mode", &c->bf_mode, 0);
    char tmp[ORA_STRLEN];
    ora_readini_c(ini, "fn_stf", c->fn_stf, "");
    ora_readini_c(ini, "stftype", tmp, "kupper");
    strncpy(c->stftype, tmp, sizeof(c->stftype) - 1);
    ora_readini_c(ini, "stf_format", tmp, "xym0ij");
    strncpy(c->stf_format, tmp, sizeof(c->stf_format) - 1);
    ora_readini_c(ini, "sdep_fit", tmp, "asis");
    strncpy(c->sdep_fit, tmp, sizeof(c->sdep_fit) - 1);
    ora_readini_l(ini, "earth_flattening", &c->earth_flattening, 0);
    if (!strcmp(c->stftype, "scosine")) strcpy(c->stftype, "cosine");

    int cap = 16, ns = 0;
    float *sx = (float *)malloc(sizeof(float) * cap), *sz = (float *)malloc(sizeof(float) * cap), *p1 = (float *)malloc(sizeof(float) * cap),
          *p2 = (float *)malloc(sizeof(float) * cap), *mo = (float *)malloc(sizeof(float) * cap), *m3 = (float *)malloc(sizeof(float) * 3 * cap);
#define GROW() if (ns == cap) { cap *= 2; sx = (float *)realloc(sx, sizeof(float) * cap); sz = (float *)realloc(sz, sizeof(float) * cap); \
        p1 = (float *)realloc(p1, sizeof(float) * cap); p2 = (float *)realloc(p2, sizeof(float) * cap); mo = (float *)realloc(mo, sizeof(float) * cap); \
        m3 = (float *)realloc(m3, sizeof(float) * 3 * cap); }
    int rc = 0;
    if (c->benchmark_mode) {   /* :99-104, :131-139 */
        strcpy(c->stftype, "kupper");
        c->bf_mode = 0;
        ns = 1;
        sx[0] = 0.0f; sz[0] = 5.0f; mo[0] = 1e15f;
        m3[0] = 1 / sqrtf(2.0f); m3[1] = 1 / sqrtf(2.0f); m3[2] = 0.0f;
        p1[0] = 0.1f; p2[0] = 2.0f;
    } else {
        char path[2 * ORA_STRLEN];
        resolve_path(base, c->fn_stf, path, sizeof(path));
        FILE *fp = fopen(path, "r");
        if (!fp) { char m[700]; snprintf(m, sizeof(m), "source__setup: cannot open %s", path); set_err(m); rc = -1; goto done; }
        char line[1024];
        const char *fmt = c->stf_format;
        while (fgets(line, sizeof(line), fp)) {
            char *p = line;
            while (*p == ' ' || *p == '\t') p++;
            if (*p == '#' || is_blank(p)) continue;
            GROW();
            float v[16], sy = 0.0f;
            const int nv = parse_sp(p, v, 16);
            float *M = &m3[3 * ns];   /* mxx mzz mxz | fx fz */
            const int is_ll = (fmt[0] == 'l' && fmt[1] == 'l'), is_xy = (fmt[0] == 'x' && fmt[1] == 'y');
            if (c->bf_mode) {   /* :513-528  x y z tbeg trise fx fy fz */
                if (nv < 8 || !(is_ll || is_xy)) { set_err("source file: bad body-force record / invalid source type"); rc = -1; fclose(fp); goto done; }
                if (is_xy) sx[ns] = v[0];
                else ora_geomap_g2c(v[0], v[1], c->clon, c->clat, c->phi, &sx[ns], &sy);
                sz[ns] = v[2]; p1[ns] = v[3]; p2[ns] = v[4]; M[0] = v[5]; M[1] = v[7]; M[2] = 0.0f; mo[ns] = 0.0f;
                if (ns == 0) {   /* :536-543 */
                    ora_geomap_c2g(sx[0], 0.0f, c->clon, c->clat, c->phi, &c->evlo, &c->evla);
                    c->evdp = sz[0]; c->otim = p1[0]; c->fx0 = M[0]; c->fz0 = M[1];
                }
                ns++;
                continue;
            }
            const char *kind = fmt + 2;
            if (!(is_ll || is_xy) || !(!strcmp(kind, "m0ij") || !strcmp(kind, "m0dc") || !strcmp(kind, "mwij") || !strcmp(kind, "mwdc"))) {
                char m[200]; snprintf(m, sizeof(m), "swpc_psv stf_format '%s' is outside the restated scope", fmt); set_err(m); rc = -1; fclose(fp); goto done;
            }
            const int need = (kind[2] == 'i') ? 12 : 9;
            if (nv < need) { set_err("source file: bad moment record"); rc = -1; fclose(fp); goto done; }
            if (is_xy) { sx[ns] = v[0]; sy = v[1]; }
            else ora_geomap_g2c(v[0], v[1], c->clon, c->clat, c->phi, &sx[ns], &sy);
            sz[ns] = v[2]; p1[ns] = v[3]; p2[ns] = v[4];
            mo[ns] = (kind[1] == '0') ? v[5] : ora_seismic_moment(v[5]);
            if (kind[2] == 'i') { M[0] = v[6]; M[1] = v[8]; M[2] = v[10]; }   /* mxx myy mzz myz mxz mxy -> mxx, mzz, mxz (:307-309) */
            else { float d1, d2, d3; ora_sdr2moment(v[6] - c->phi, v[7], v[8], &M[0], &d1, &M[1], &d2, &M[2], &d3); }
            if (ns == 0) {   /* :454-466 */
                ora_geomap_c2g(sx[0], sy, c->clon, c->clat, c->phi, &c->evlo, &c->evla);
                c->sx0 = sx[0]; c->sy0 = sy; c->evdp = sz[0]; c->mxx0 = M[0]; c->mzz0 = M[1]; c->mxz0 = M[2]; c->otim = p1[0];
            }
            ns++;
        }
        fclose(fp);
    }
    if (c->earth_flattening) {   /* :144-148 */
        const double RE = ora_r_earth();
        for (int k = 0; k < ns; k++) sz[k] = -(float)(RE * log((RE - (double)sz[k]) / RE));
    }
    if (c->bf_mode) {   /* :151-157 */
        float sum = 0.0f;
        for (int i = 0; i < ns; i++) sum += m3[3 * i] * m3[3 * i] + m3[3 * i + 1] * m3[3 * i + 1];
        c->M0 = sqrtf(sum);
        c->UC = c->UC * 1000;
    } else {
        float sum = 0.0f;
        for (int i = 0; i < ns; i++) sum += mo[i];
        c->M0 = sum;
    }
    c->fcut = 0.0f;   /* :161-165 */
    for (int i = 0; i < ns; i++) { const float f = 1 / p2[i]; if (f > c->fcut) c->fcut = f; }
    c->fmax = 2 * c->fcut;
    c->dt_dxz = (ora_mp)c->dt / ((ora_mp)c->dx * (ora_mp)c->dz);   /* :252 */
    for (int q = 0; q < s->nranks; q++) {
        psv_rank *r = &s->r[q];
        int n = 0;
        for (int i = 0; i < ns; i++) {   /* :168-184 */
            const int is = ora_x2i(sx[i], c->xbeg, (float)c->dx), ks = ora_x2i(sz[i], c->zbeg, (float)c->dz);
            if (r->ibeg - 2 <= is && is <= r->iend + 3 && r->kbeg - 2 <= ks && ks <= r->kend + 3) n++;
        }
        r->nsrc = n;
        r->isrc = (int *)xcalloc(n, sizeof(int)); r->ksrc = (int *)xcalloc(n, sizeof(int));
        r->sx = (float *)xcalloc(n, sizeof(float)); r->sz = (float *)xcalloc(n, sizeof(float)); r->srcprm = (float *)xcalloc(2 * (size_t)n, sizeof(float));
        r->mo = (ora_mp *)xcalloc(n, sizeof(ora_mp)); r->mxx = (ora_mp *)xcalloc(n, sizeof(ora_mp)); r->mzz = (ora_mp *)xcalloc(n, sizeof(ora_mp));
        r->mxz = (ora_mp *)xcalloc(n, sizeof(ora_mp)); r->fx = (ora_mp *)xcalloc(n, sizeof(ora_mp)); r->fz = (ora_mp *)xcalloc(n, sizeof(ora_mp));
        int nn = 0;
        for (int i = 0; i < ns; i++) {
            const int is = ora_x2i(sx[i], c->xbeg, (float)c->dx), ks = ora_x2i(sz[i], c->zbeg, (float)c->dz);
            if (!(r->ibeg - 2 <= is && is <= r->iend + 3 && r->kbeg - 2 <= ks && ks <= r->kend + 3)) continue;
            r->isrc[nn] = is; r->ksrc[nn] = ks; r->sx[nn] = sx[i]; r->sz[nn] = sz[i];
            if (c->bf_mode) { r->fx[nn] = m3[3 * i]; r->fz[nn] = m3[3 * i + 1]; }
            else { r->mo[nn] = mo[i]; r->mxx[nn] = m3[3 * i]; r->mzz[nn] = m3[3 * i + 1]; r->mxz[nn] = m3[3 * i + 2]; }
            r->srcprm[2 * nn] = p1[i]; r->srcprm[2 * nn + 1] = p2[i];
            if (c->sdep_fit[0] == 'b' && c->sdep_fit[1] == 'd' && isdigit((unsigned char)c->sdep_fit[2])) {   /* :219-227 */
                r->sz[nn] = r->bddep[(size_t)(c->sdep_fit[2] - '0') * r->nxm + (is - r->ibeg_m)];
                r->ksrc[nn] = ora_x2i(r->sz[nn], c->zbeg, (float)c->dz);
            }
            if (!(c->xbeg <= r->sx[nn] && r->sx[nn] <= c->xend && c->zbeg <= r->sz[nn] && r->sz[nn] <= c->zend)) {   /* :233-236 */
                set_err("source__setup: assert failed, source outside of the model space");
                rc = -1;
            }
            nn++;
        }
        for (int i = 0; i < n; i++) {   /* :243-249 */
            if (c->bf_mode) { r->fx[i] = r->fx[i] / c->M0; r->fz[i] = r->fz[i] / c->M0; }
            else r->mo[i] = r->mo[i] / c->M0;
        }
    }
done:
    free(sx); free(sz); free(p1); free(p2); free(mo); free(m3);
    return rc;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* m_absorb_p.f90:57-101 / m_absorb_c.f90:28-96 */
static int absorb_setup(psv_sim *s) {
    psv_cfg *c = &s->cfg;
    const float dx = (float)c->dx, dz = (float)c->dz;
    const int na = c->na, nx = c->nx, nz = c->nz;
    if (!strcmp(c->abc_type, "pml")) {
        c->r20x = (float)(1.0f / c->dx);   /* real(1.0 / dx) */
        c->r20z = (float)(1.0f / c->dz);
        const float hx = (float)(na * c->dx), hz = (float)(na * c->dz);
        for (int q = 0; q < s->nranks; q++) {
            psv_rank *r = &s->r[q];
            const int nxo = r->iend - r->ibeg + 1;
            r->gxc = (float *)xcalloc(4 * (size_t)nxo, sizeof(float)); r->gxe = (float *)xcalloc(4 * (size_t)nxo, sizeof(float));
            r->gzc = (float *)xcalloc(4 * (size_t)nz, sizeof(float)); r->gze = (float *)xcalloc(4 * (size_t)nz, sizeof(float));
            for (int i = r->ibeg; i <= r->iend; i++) {
                const float x = r->xc[i - r->ibeg_m];
                ora_damping_profile(x, hx, c->xbeg, c->xend, na, c->fcut, c->dt, &r->gxc[4 * (i - r->ibeg)]);
                ora_damping_profile(x + dx / 2.0f, hx, c->xbeg, c->xend, na, c->fcut, c->dt, &r->gxe[4 * (i - r->ibeg)]);
            }
            for (int k = r->kbeg; k <= r->kend; k++) {
                const float z = r->zc[k - r->kbeg_m];
                ora_damping_profile(z, hz, c->zbeg, c->zend, na, c->fcut, c->dt, &r->gzc[4 * (k - r->kbeg)]);
                ora_damping_profile(z + dz / 2.0f, hz, c->zbeg, c->zend, na, c->fcut, c->dt, &r->gze[4 * (k - r->kbeg)]);
            }
            int kmin = 1 << 30;   /* :88  minval(kbeg_a(:)) over the memory box */
            for (int i = 0; i < r->nxm; i++) if (r->kbeg_a[i] < kmin) kmin = r->kbeg_a[i];
            r->kbeg_min = kmin;
            const size_t na_ = (size_t)(r->kend - kmin + 1) * nxo;
            for (int a = 0; a < 8; a++) r->aux[a] = (float *)xcalloc(na_, sizeof(float));
        }
    } else if (!strcmp(c->abc_type, "cerjan")) {
        const float alpha = 0.09f;
        const float Lx = (float)(na * c->dx), Lz = (float)(na * c->dz);
#define SQ(x) ((x) * (x))
        for (int q = 0; q < s->nranks; q++) {
            psv_rank *r = &s->r[q];
            r->gx_c = (float *)xcalloc((size_t)r->nxm, sizeof(float)); r->gx_b = (float *)xcalloc((size_t)r->nxm, sizeof(float));
            r->gz_c = (float *)xcalloc((size_t)r->nzm, sizeof(float)); r->gz_b = (float *)xcalloc((size_t)r->nzm, sizeof(float));
            for (int i = 0; i < r->nxm; i++) r->gx_c[i] = r->gx_b[i] = 1.0f;
            for (int k = 0; k < r->nzm; k++) r->gz_c[k] = r->gz_b[k] = 1.0f;
            for (int i = r->ibeg; i <= r->iend; i++) {
                float *gc = &r->gx_c[i - r->ibeg_m], *gb = &r->gx_b[i - r->ibeg_m];
                if (i <= na) {
                    *gc = expf(-(alpha * SQ(1.0f - (ora_i2x(i, 0.0f, dx)) / Lx)));
                    *gb = expf(-(alpha * SQ(1.0f - ((ora_i2x(i, 0.0f, dx) + dx / 2)) / Lx)));
                } else if (i >= nx - na + 1) {
                    *gc = expf(-(alpha * SQ(1.0f - (ora_i2x(i, nx * dx, -dx) + dx / 2) / Lx)));
                    *gb = expf(-(alpha * SQ(1.0f - ((ora_i2x(i, nx * dx, -dx))) / Lx)));
                }
            }
            for (int k = r->kbeg; k <= r->kend; k++) {
                float *gc = &r->gz_c[k - r->kbeg_m], *gb = &r->gz_b[k - r->kbeg_m];
                if (k <= na) { *gc = 1.0f; *gb = 1.0f; }
                else if (k >= nz - na + 1) {
                    *gc = expf(-(alpha * SQ(1.0f - (ora_i2x(k, nz * dz, -dz) + dz / 2) / Lz)));
                    *gb = expf(-(alpha * SQ(1.0f - ((ora_i2x(k, nz * dz, -dz))) / Lz)));
                }
            }
        }
#undef SQ
    } else { set_err("absorb__setup: unknown abc_type (assert(.false.), m_absorb.f90:37)"); return -1; }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* m_wav.f90:53-141 wav__setup + set_stinfo :433-571 */
static int wav_setup(psv_sim *s, const ora_ini *ini, const char *base) {
    psv_cfg *c = &s->cfg;
    char tmp[ORA_STRLEN];
    ora_readini_i(ini, "ntdec_w", &c->ntdec_w, 10);
    ora_readini_l(ini, "sw_wav_v", &c->sw_v, 0);
    ora_readini_l(ini, "sw_wav_u", &c->sw_u, 0);
    ora_readini_l(ini, "sw_wav_stress", &c->sw_stress, 0);
    ora_readini_l(ini, "sw_wav_strain", &c->sw_strain, 0);
    ora_readini_c(ini, "wav_format", tmp, "sac");
    strncpy(c->wav_format, tmp, sizeof(c->wav_format) - 1);
    ora_readini_c(ini, "st_format", tmp, "xy");
    strncpy(c->st_format, tmp, sizeof(c->st_format) - 1);
    ora_readini_c(ini, "fn_stloc", c->fn_stloc, "");
    if (!(c->sw_v || c->sw_u || c->sw_stress || c->sw_strain)) return 0;
    c->ntw = (int)floorf((float)(c->nt - 1) / (float)c->ntdec_w + 1.0f);
    const ora_mp dxm = (ora_mp)c->dx, dzm = (ora_mp)c->dz;
    c->w40x = (ora_mp)9.0 / (ora_mp)8.0 / dxm; c->w40z = (ora_mp)9.0 / (ora_mp)8.0 / dzm;
    c->w41x = (ora_mp)1.0 / (ora_mp)24.0 / dxm; c->w41z = (ora_mp)1.0 / (ora_mp)24.0 / dzm;
    char path[2 * ORA_STRLEN];
    resolve_path(base, c->fn_stloc, path, sizeof(path));
    FILE *fp = fopen(path, "r");
    if (!fp) return 0;
    char line[1024];
    const float dx = (float)c->dx, dz = (float)c->dz;
    while (fgets(line, sizeof(line), fp)) {
        char *p = line;
        while (*p == ' ' || *p == '\t') p++;
        if (*p == '#' || is_blank(p)) continue;
        float a, b, zst;
        char stnm[64] = "", zsw[64] = "";
        if (sscanf(p, "%f %f %f %63s %63s", &a, &b, &zst, stnm, zsw) < 5) continue;
        stnm[8] = 0; zsw[3] = 0;
        float xst, stlo, stla, dum;
        if (!strcmp(c->st_format, "xy")) { xst = a; ora_geomap_c2g(xst, 0.0f, c->clon, c->clat, c->phi, &stlo, &stla); }
        else if (!strcmp(c->st_format, "ll")) { stlo = a; stla = b; ora_geomap_g2c(stlo, stla, c->clon, c->clat, c->phi, &xst, &dum); }
        else { set_err("unknown st_format"); fclose(fp); return -1; }
        const int ist = ora_x2i(xst, c->xbeg, dx), kst = ora_x2i(zst, c->zbeg, dz);
        if (!(ora_i2x(1, c->xbeg, dx) < xst && xst < ora_i2x(c->nx, c->xbeg, dx) && 1 < kst && kst < c->nz)) continue;   /* :497-498 */
        for (int q = 0; q < s->nranks; q++) {
            psv_rank *r = &s->r[q];
            if (!(r->ibeg <= ist && ist <= r->iend)) continue;
            int k;
            if (!strcmp(zsw, "dep")) k = ora_x2i(zst, c->zbeg, dz);
            else if (!strcmp(zsw, "fsb")) k = r->kfs[ist - r->ibeg_m] + 1;
            else if (!strcmp(zsw, "obb")) k = r->kob[ist - r->ibeg_m] + 1;
            else if (!strcmp(zsw, "oba")) k = r->kob[ist - r->ibeg_m] - 1;
            else if (zsw[0] == 'b' && zsw[1] == 'd' && isdigit((unsigned char)zsw[2]))
                k = ora_x2i(r->bddep[(size_t)(zsw[2] - '0') * r->nxm + (ist - r->ibeg_m)], c->zbeg, dz);
            else k = ora_x2i(zst, c->zbeg, dz);
            if (k > r->kend) k = r->kend - 1;
            if (k < r->kbeg) k = r->kbeg + 1;
            const int n = r->nst++;
            r->ist = (int *)realloc(r->ist, sizeof(int) * r->nst); r->kst = (int *)realloc(r->kst, sizeof(int) * r->nst);
            r->xst = (float *)realloc(r->xst, sizeof(float) * r->nst); r->zst = (float *)realloc(r->zst, sizeof(float) * r->nst);
            r->stlo = (float *)realloc(r->stlo, sizeof(float) * r->nst); r->stla = (float *)realloc(r->stla, sizeof(float) * r->nst);
            r->stnm = (char(*)[9])realloc(r->stnm, 9 * (size_t)r->nst);
            r->ist[n] = ist; r->kst[n] = k; r->xst[n] = xst; r->zst[n] = zst; r->stlo[n] = stlo; r->stla[n] = stla;
            memset(r->stnm[n], 0, 9);
            strncpy(r->stnm[n], stnm, 8);
        }
    }
    fclose(fp);
    for (int q = 0; q < s->nranks; q++) {
        psv_rank *r = &s->r[q];
        if (r->nst <= 0) continue;
        const size_t n2 = (size_t)c->ntw * 2 * r->nst, n3 = (size_t)c->ntw * 3 * r->nst;
        if (c->sw_v) r->wav[0] = (float *)xcalloc(n2, sizeof(float));
        if (c->sw_u) { r->wav[1] = (float *)xcalloc(n2, sizeof(float)); r->ux = (float *)xcalloc((size_t)r->nst, sizeof(float)); r->uz = (float *)xcalloc((size_t)r->nst, sizeof(float)); }
        if (c->sw_stress) r->wav[2] = (float *)xcalloc(n3, sizeof(float));
        if (c->sw_strain) {
            r->wav[3] = (float *)xcalloc(n3, sizeof(float));
            r->exx = (float *)xcalloc((size_t)r->nst, sizeof(float)); r->ezz = (float *)xcalloc((size_t)r->nst, sizeof(float)); r->exz = (float *)xcalloc((size_t)r->nst, sizeof(float));
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* ------------------------------------------------------------------------------------------------------------ */
/* snap__setup m_snap.f90:78-164 (+ the medium slices of newfile_xz[_nc] :167-269)                                  */
static void snap_region(const psv_snap *sn, const psv_rank *r, int *is0, int *is1, int *ks0, int *ks1) {
    *is0 = (int)ceilf((float)(r->ibeg + sn->idec / 2) / (float)sn->idec);
    *is1 = (int)floorf((float)(r->iend + sn->idec / 2) / (float)sn->idec);
    *ks0 = (int)ceilf((float)(r->kbeg + sn->kdec / 2) / (float)sn->kdec);
    *ks1 = (int)floorf((float)(r->kend + sn->kdec / 2) / (float)sn->kdec);
}
static void snap_setup(psv_sim *s, const ora_ini *ini) {
    psv_snap *sn = &s->snap;
    const psv_cfg *c = &s->cfg;
    ora_readini_l(ini, "xz_ps%sw", &sn->sw[0], 0);
    ora_readini_l(ini, "xz_v%sw", &sn->sw[1], 0);
    ora_readini_l(ini, "xz_u%sw", &sn->sw[2], 0);
    ora_readini_i(ini, "idec", &sn->idec, 1);
    ora_readini_i(ini, "kdec", &sn->kdec, 1);
    ora_readini_i(ini, "ntdec_s", &sn->ntdec_s, 10);
    char tmp[ORA_STRLEN];
    ora_readini_c(ini, "snp_format", tmp, "native");
    snprintf(sn->snp_format, sizeof(sn->snp_format), "%.7s", tmp);
    sn->nxs = (c->nx + (sn->idec / 2)) / sn->idec;
    sn->nzs = (c->nz + (sn->kdec / 2)) / sn->kdec;
    sn->xsnp = (float *)xcalloc((size_t)sn->nxs, sizeof(float));
    sn->zsnp = (float *)xcalloc((size_t)sn->nzs, sizeof(float));
    for (int i = 1; i <= sn->nxs; i++) sn->xsnp[i - 1] = ora_i2x(i * sn->idec - (sn->idec / 2), c->xbeg, (float)c->dx);
    for (int k = 1; k <= sn->nzs; k++) sn->zsnp[k - 1] = ora_i2x(k * sn->kdec - (sn->kdec / 2), c->zbeg, (float)c->dz);
    const size_t n2 = (size_t)sn->nxs * sn->nzs;
    sn->medium = (float *)xcalloc(3 * n2, sizeof(float));
    sn->buf_u = (float *)xcalloc(2 * n2, sizeof(float));
    for (int q = 0; q < s->nranks; q++) {
        const psv_rank *r = &s->r[q];
        int is0, is1, ks0, ks1;
        snap_region(sn, r, &is0, &is1, &ks0, &ks1);
        for (int i = is0; i <= is1; i++)
            for (int k = ks0; k <= ks1; k++) {
                const int ii = i * sn->idec - sn->idec / 2, kk = k * sn->kdec - sn->kdec / 2;
                const size_t o = (size_t)(i - 1) + (size_t)sn->nxs * (size_t)(k - 1);
                sn->medium[o] += r->rho[IX(r, kk, ii)];
                sn->medium[n2 + o] += r->lam[IX(r, kk, ii)];
                sn->medium[2 * n2 + o] += r->mu[IX(r, kk, ii)];
            }
    }
}

/* snap__write m_snap.f90:419-650, called at the top of the iteration (main.f90:99); the mpi_reduce(SUM) onto the I/O
 * rank is the disjoint fill below.  A record is kept for every it with mod(it-1, ntdec_s) == 0. */
static void snap_write(psv_sim *s, int it) {
    psv_snap *sn = &s->snap;
    const psv_cfg *c = &s->cfg;
    if (!(sn->sw[0] || sn->sw[1] || sn->sw[2])) return;
    const size_t n2 = (size_t)sn->nxs * sn->nzs;
    const int sample = (it - 1) % sn->ntdec_s == 0;
    const ora_mp r20x = (ora_mp)1.0 / (ora_mp)c->dx, r20z = (ora_mp)1.0 / (ora_mp)c->dz;
    float *out[3] = {NULL, NULL, NULL};
    for (int p = 0; p < 3; p++) {
        if (!sn->sw[p] || !sample) continue;
        if (sn->nrec[p] == sn->cap[p]) {
            sn->cap[p] = sn->cap[p] ? 2 * sn->cap[p] : 16;
            sn->rec[p] = (float *)realloc(sn->rec[p], sizeof(float) * 2 * n2 * (size_t)sn->cap[p]);
            sn->it0[p] = (int *)realloc(sn->it0[p], sizeof(int) * (size_t)sn->cap[p]);
        }
        out[p] = sn->rec[p] + 2 * n2 * (size_t)sn->nrec[p];
        memset(out[p], 0, sizeof(float) * 2 * n2);
        sn->it0[p][sn->nrec[p]++] = it;
    }
    for (int q = 0; q < s->nranks; q++) {
        const psv_rank *r = &s->r[q];
        int is0, is1, ks0, ks1;
        snap_region(sn, r, &is0, &is1, &ks0, &ks1);
        for (int ii = is0; ii <= is1; ii++)
            for (int kk = ks0; kk <= ks1; kk++) {
                const int k = kk * sn->kdec - sn->kdec / 2, i = ii * sn->idec - sn->idec / 2;
                const size_t o = (size_t)(ii - 1) + (size_t)sn->nxs * (size_t)(kk - 1);
                if (out[0]) { /* :468-494 */
                    float div = (float)((r->Vx[IX(r, k, i)] - r->Vx[IX(r, k, i - 1)]) * r20x + (r->Vz[IX(r, k, i)] - r->Vz[IX(r, k - 1, i)]) * r20z);
                    float rot = (float)((r->Vx[IX(r, k + 1, i)] - r->Vx[IX(r, k, i)]) * r20z - (r->Vz[IX(r, k, i + 1)] - r->Vz[IX(r, k, i)]) * r20x);
                    const float nnn = r->mu[IX(r, k, i)], pnn = r->mu[IX(r, k + 1, i)], npn = r->mu[IX(r, k, i + 1)], ppn = r->mu[IX(r, k + 1, i + 1)];
                    const float mu_xz = 4 * nnn * pnn * npn * ppn / (nnn * pnn * npn + nnn * pnn * ppn + nnn * npn * ppn + pnn * npn * ppn + FLT_EPS);
                    const float lam0 = r->lam[IX(r, k, i)];
                    div = div * lam0 / (fabsf(lam0) + FLT_EPS);
                    rot = rot * mu_xz / fabsf(mu_xz + FLT_EPS);
                    out[0][o] = div * c->UC * c->M0 * 1e-3f;   /* buf = buf * UC * M0 * 1e-3 */
                    out[0][n2 + o] = rot * c->UC * c->M0 * 1e-3f;
                }
                if (out[1]) { /* :550-551 */
                    out[1][o] = (float)(r->Vx[IX(r, k, i)] * c->UC * c->M0);
                    out[1][n2 + o] = (float)(r->Vz[IX(r, k, i)] * c->UC * c->M0);
                }
                if (sn->sw[2]) { /* every step, :609-610 */
                    sn->buf_u[o] = (float)(sn->buf_u[o] + r->Vx[IX(r, k, i)] * c->UC * c->M0 * c->dt);
                    sn->buf_u[n2 + o] = (float)(sn->buf_u[n2 + o] + r->Vz[IX(r, k, i)] * c->UC * c->M0 * c->dt);
                }
            }
    }
    if (out[2]) memcpy(out[2], sn->buf_u, sizeof(float) * 2 * n2);
}

static psv_sim *create_from_ini(ora_ini *ini, const char *base, int nm, int npx, int nt) {
    psv_sim *s = (psv_sim *)xcalloc(1, sizeof(psv_sim));
    psv_cfg *c = &s->cfg;
    global_setup(c, ini);
    if (nm < 0 || nm > PSV_MAXNM) { set_err("nm out of range"); free(s); return NULL; }
    c->nm = nm;
    if (npx > 0) c->nproc_x = npx;
    if (nt > 0) c->nt = nt;
    c->exedate = (int)time(NULL);
    s->nranks = c->nproc_x;
    s->r = (psv_rank *)xcalloc((size_t)s->nranks, sizeof(psv_rank));
    for (int q = 0; q < s->nranks; q++) rank_geometry(c, &s->r[q], q);
    float vmin = 1e30f, vmax = -1.0f;
    for (int q = 0; q < s->nranks; q++) {
        float a, b;
        if (medium_setup(s, &s->r[q], ini, base, &a, &b)) { psv_destroy(s); return NULL; }
        if (a < vmin) vmin = a;
        if (b > vmax) vmax = b;
    }
    c->vmin = vmin; c->vmax = vmax;
    int stab;
    ora_readini_l(ini, "stabilize_pml", &stab, 0);
    if (stab) for (int q = 0; q < s->nranks; q++) stabilize_absorber(c, &s->r[q]);
    kernel_setup(s);
    if (source_setup(s, ini, base) || absorb_setup(s) || wav_setup(s, ini, base)) { psv_destroy(s); return NULL; }
    ora_readini_i(ini, "ntdec_r", &c->ntdec_r, 10);
    snap_setup(s, ini);
    return s;
}
psv_sim *psv_create(const char *inf_path, const char *base, int nm, int npx, int nt) {
    ora_ini *ini = ora_ini_open(inf_path);
    if (!ini) { set_err("cannot open the parameter file"); return NULL; }
    psv_sim *s = create_from_ini(ini, base, nm, npx, nt);
    ora_ini_close(ini);
    return s;
}
psv_sim *psv_create_from_text(const char *text, const char *base, int nm, int npx, int nt) {
    ora_ini *ini = ora_ini_from_text(text);
    psv_sim *s = create_from_ini(ini, base, nm, npx, nt);
    ora_ini_close(ini);
    return s;
}
void psv_destroy(psv_sim *s) {
    if (!s) return;
    for (int q = 0; q < s->nranks; q++) {
        psv_rank *r = &s->r[q];
        void *ptrs[] = {r->Vx, r->Vz, r->Sxx, r->Szz, r->Sxz, r->Rxx, r->Rzz, r->Rxz, r->rho, r->lam, r->mu, r->taup, r->taus, r->kfs, r->kob,
                        r->kfs_top, r->kfs_bot, r->kob_top, r->kob_bot, r->kbeg_a, r->bddep, r->xc, r->zc, r->gxc, r->gxe, r->gzc, r->gze,
                        r->gx_c, r->gx_b, r->gz_c, r->gz_b, r->isrc, r->ksrc, r->sx, r->sz, r->srcprm, r->mo, r->mxx, r->mzz, r->mxz, r->fx, r->fz,
                        r->ist, r->kst, r->xst, r->zst, r->stlo, r->stla, r->stnm, r->ux, r->uz, r->exx, r->ezz, r->exz, r->sbuf_ip, r->sbuf_im,
                        r->rbuf_ip, r->rbuf_im};
        for (size_t a = 0; a < sizeof(ptrs) / sizeof(ptrs[0]); a++) free(ptrs[a]);
        for (int a = 0; a < 8; a++) free(r->aux[a]);
        for (int a = 0; a < 4; a++) free(r->wav[a]);
    }
    free(s->r);
    free(s->snap.xsnp); free(s->snap.zsnp); free(s->snap.medium); free(s->snap.buf_u);
    for (int p = 0; p < 3; p++) { free(s->snap.rec[p]); free(s->snap.it0[p]); }
    free(s);
}
void psv_set_exedate(psv_sim *s, int exedate, int tz_minutes) { s->cfg.exedate = exedate; s->cfg.tz_minutes = tz_minutes; }

/* ------------------------------------------------------------------------------------------------------------ */
/* the time step                                                                                                   */
static inline int imax(int a, int b) { return a > b ? a : b; }
/* m_kernel.f90:104-105: isign = sign(1, max((k-kfs_top)(kfs_bot-k), (k-kob_top)(kob_bot-k))) */
static inline int fd_sign(const psv_rank *r, int k, int i) {
    const int q = i - r->ibeg_m;
    return imax((k - r->kfs_top[q]) * (r->kfs_bot[q] - k), (k - r->kob_top[q]) * (r->kob_bot[q] - k)) >= 0 ? 1 : -1;
}
static inline float mu_harm(float a, float b, float cc, float d) {
    return 4 * a * b * cc * d / (a * b * cc + a * b * d + a * cc * d + b * cc * d + FLT_EPS);
}

/* m_kernel.f90:76-140 */
static void kernel_update_vel(const psv_cfg *c, psv_rank *r) {
    const float dt = c->dt;
    const size_t SI = (size_t)r->nzm;
#pragma omp parallel for schedule(static, 1)
    for (int i = r->ibeg_k; i <= r->iend_k; i++)
        for (int k = r->kbeg_k; k <= r->kend_k; k++) {
            const int isign = fd_sign(r, k, i);
            const ora_mp re40x = c->rc40x + isign * c->rd40x, re41x = c->rc41x + isign * c->rd41x;
            const ora_mp re40z = c->rc40z + isign * c->rd40z, re41z = c->rc41z + isign * c->rd41z;
            const size_t n = IX(r, k, i);
            const ora_mp *Sxx = r->Sxx, *Szz = r->Szz, *Sxz = r->Sxz;
            const ora_mp dxSxx = (Sxx[n + SI] - Sxx[n]) * re40x - (Sxx[n + 2 * SI] - Sxx[n - SI]) * re41x;
            const ora_mp dzSzz = (Szz[n + 1] - Szz[n]) * re40z - (Szz[n + 2] - Szz[n - 1]) * re41z;
            const ora_mp dxSxz = (Sxz[n] - Sxz[n - SI]) * re40x - (Sxz[n + SI] - Sxz[n - 2 * SI]) * re41x;
            const ora_mp dzSxz = (Sxz[n] - Sxz[n - 1]) * re40z - (Sxz[n + 1] - Sxz[n - 2]) * re41z;
            const float bx = 2.0f / (r->rho[n] + r->rho[n + SI]);
            const float bz = 2.0f / (r->rho[n] + r->rho[n + 1]);
            r->Vx[n] = r->Vx[n] + bx * (dxSxx + dzSxz) * dt;
            r->Vz[n] = r->Vz[n] + bz * (dxSxz + dzSzz) * dt;
        }
}

/* m_kernel.f90:142-311 */
static void kernel_update_stress(const psv_cfg *c, psv_rank *r) {
    const float dt = c->dt, d2 = c->d2;
    const int nm = c->nm;
    const size_t SI = (size_t)r->nzm;
#pragma omp parallel for schedule(static, 1)
    for (int i = r->ibeg_k; i <= r->iend_k; i++)
        for (int k = r->kbeg_k; k <= r->kend_k; k++) {
            const int isign = fd_sign(r, k, i);
            const ora_mp re40x = c->rc40x + isign * c->rd40x, re41x = c->rc41x + isign * c->rd41x;
            const ora_mp re40z = c->rc40z + isign * c->rd40z, re41z = c->rc41z + isign * c->rd41z;
            const size_t n = IX(r, k, i);
            const ora_mp *Vx = r->Vx, *Vz = r->Vz;
            const ora_mp dxVx = (Vx[n] - Vx[n - SI]) * re40x - (Vx[n + SI] - Vx[n - 2 * SI]) * re41x;
            const ora_mp dzVz = (Vz[n] - Vz[n - 1]) * re40z - (Vz[n + 1] - Vz[n - 2]) * re41z;
            const float mu2 = 2 * r->mu[n];
            const float lam2mu = r->lam[n] + mu2;
            const float taup1 = r->taup[n], taus1 = r->taus[n];
            const float d2v2 = (float)(dxVx + dzVz);
            const float f_Rxx = (float)(lam2mu * taup1 * d2v2 - mu2 * taus1 * dzVz);
            const float f_Rzz = (float)(lam2mu * taup1 * d2v2 - mu2 * taus1 * dxVx);
            float Rxx_n = 0.0f, Rzz_n = 0.0f;
            for (int m = 0; m < nm; m++) {
                float *Rxx = &r->Rxx[(size_t)nm * n + m], *Rzz = &r->Rzz[(size_t)nm * n + m];
                *Rxx = c->c1[m] * *Rxx - c->c2[m] * f_Rxx * dt;
                *Rzz = c->c1[m] * *Rzz - c->c2[m] * f_Rzz * dt;
                Rxx_n = Rxx_n + c->d1[m] * *Rxx;
                Rzz_n = Rzz_n + c->d1[m] * *Rzz;
            }
            const float taup_plus1 = 1 + taup1 * (1 + d2), taus_plus1 = 1 + taus1 * (1 + d2);
            r->Sxx[n] = r->Sxx[n] + (lam2mu * taup_plus1 * d2v2 - mu2 * taus_plus1 * dzVz + Rxx_n) * dt;
            r->Szz[n] = r->Szz[n] + (lam2mu * taup_plus1 * d2v2 - mu2 * taus_plus1 * dxVx + Rzz_n) * dt;
        }
#pragma omp parallel for schedule(static, 1)
    for (int i = r->ibeg_k; i <= r->iend_k; i++)
        for (int k = r->kbeg_k; k <= r->kend_k; k++) {
            const int isign = fd_sign(r, k, i);
            const ora_mp re40x = c->rc40x + isign * c->rd40x, re41x = c->rc41x + isign * c->rd41x;
            const ora_mp re40z = c->rc40z + isign * c->rd40z, re41z = c->rc41z + isign * c->rd41z;
            const size_t n = IX(r, k, i);
            const ora_mp *Vx = r->Vx, *Vz = r->Vz;
            const ora_mp dxVz = (Vz[n + SI] - Vz[n]) * re40x - (Vz[n + 2 * SI] - Vz[n - SI]) * re41x;
            const ora_mp dzVx = (Vx[n + 1] - Vx[n]) * re40z - (Vx[n + 2] - Vx[n - 1]) * re41z;
            const float taus1 = r->taus[n];
            const float mu_xz = mu_harm(r->mu[n], r->mu[n + 1], r->mu[n + SI], r->mu[n + 1 + SI]);
            const float dxVz_dzVx = (float)(dxVz + dzVx);   /* real(SP), :155 */
            const float f_Rxz = mu_xz * taus1 * dxVz_dzVx;
            float Rxz_n = 0.0f;
            for (int m = 0; m < nm; m++) {
                float *Rxz = &r->Rxz[(size_t)nm * n + m];
                *Rxz = c->c1[m] * *Rxz - c->c2[m] * f_Rxz * dt;
                Rxz_n = Rxz_n + c->d1[m] * *Rxz;
            }
            const float taus_plus1 = 1 + taus1 * (1 + d2);
            r->Sxz[n] = r->Sxz[n] + (mu_xz * taus_plus1 * dxVz_dzVx + Rxz_n) * dt;
        }
}

/* m_absorb_p.f90:103-201 (plane-wave edges :112-155 outside the scope) */
static void absorb_p_update_vel(const psv_cfg *c, psv_rank *r) {
    const float dt = c->dt, r20x = c->r20x, r20z = c->r20z;
    const size_t SI = (size_t)r->nzm;
    float *axSxx = r->aux[4], *azSxz = r->aux[5], *axSxz = r->aux[6], *azSzz = r->aux[7];
#pragma omp parallel for schedule(dynamic)
    for (int i = r->ibeg; i <= r->iend; i++) {
        const float *gxc = &r->gxc[4 * (i - r->ibeg)], *gxe = &r->gxe[4 * (i - r->ibeg)];
        for (int k = r->kbeg_a[i - r->ibeg_m]; k <= r->kend; k++) {
            const float *gzc = &r->gzc[4 * (k - r->kbeg)], *gze = &r->gze[4 * (k - r->kbeg)];
            const size_t n = IX(r, k, i), a = AX(r, k, i);
            const ora_mp dxSxx = (r->Sxx[n + SI] - r->Sxx[n]) * r20x;
            const ora_mp dzSzz = (r->Szz[n + 1] - r->Szz[n]) * r20z;
            const ora_mp dxSxz = (r->Sxz[n] - r->Sxz[n - SI]) * r20x;
            const ora_mp dzSxz = (r->Sxz[n] - r->Sxz[n - 1]) * r20z;
            const float bx = 2.0f / (r->rho[n] + r->rho[n + SI]);
            const float bz = 2.0f / (r->rho[n] + r->rho[n + 1]);
            r->Vx[n] = r->Vx[n] + bx * (gxe[0] * dxSxx + gzc[0] * dzSxz + gxe[1] * axSxx[a] + gzc[1] * azSxz[a]) * dt;
            r->Vz[n] = r->Vz[n] + bz * (gxc[0] * dxSxz + gze[0] * dzSzz + gxc[1] * axSxz[a] + gze[1] * azSzz[a]) * dt;
            axSxx[a] = (float)(gxe[2] * axSxx[a] + gxe[3] * dxSxx * dt);
            azSxz[a] = (float)(gzc[2] * azSxz[a] + gzc[3] * dzSxz * dt);
            axSxz[a] = (float)(gxc[2] * axSxz[a] + gxc[3] * dxSxz * dt);
            azSzz[a] = (float)(gze[2] * azSzz[a] + gze[3] * dzSzz * dt);
        }
    }
}

/* m_absorb_p.f90:266-400 */
static void absorb_p_update_stress(const psv_cfg *c, psv_rank *r) {
    const float dt = c->dt, r20x = c->r20x, r20z = c->r20z;
    const size_t SI = (size_t)r->nzm;
    float *axVx = r->aux[0], *azVx = r->aux[1], *axVz = r->aux[2], *azVz = r->aux[3];
#pragma omp parallel for schedule(dynamic)
    for (int i = r->ibeg; i <= r->iend; i++) {
        const float *gxc = &r->gxc[4 * (i - r->ibeg)], *gxe = &r->gxe[4 * (i - r->ibeg)];
        for (int k = r->kbeg_a[i - r->ibeg_m]; k <= r->kend; k++) {
            const float *gzc = &r->gzc[4 * (k - r->kbeg)];
            const size_t n = IX(r, k, i), a = AX(r, k, i);
            const ora_mp dxVx = (r->Vx[n] - r->Vx[n - SI]) * r20x;
            const ora_mp dzVz = (r->Vz[n] - r->Vz[n - 1]) * r20z;
            const float lam2mu_R = r->lam[n] + 2 * r->mu[n];
            const float lam_R = lam2mu_R - 2 * r->mu[n];
            const float dxVx_ade = (float)(gxc[0] * dxVx + gxc[1] * axVx[a]);
            const float dzVz_ade = (float)(gzc[0] * dzVz + gzc[1] * azVz[a]);
            r->Sxx[n] = r->Sxx[n] + (lam2mu_R * dxVx_ade + lam_R * dzVz_ade) * dt;
            r->Szz[n] = r->Szz[n] + (lam2mu_R * dzVz_ade + lam_R * dxVx_ade) * dt;
            axVx[a] = (float)(gxc[2] * axVx[a] + gxc[3] * dxVx * dt);
            azVz[a] = (float)(gzc[2] * azVz[a] + gzc[3] * dzVz * dt);
        }
        for (int k = r->kbeg_a[i - r->ibeg_m]; k <= r->kend; k++) {
            const float *gze = &r->gze[4 * (k - r->kbeg)];
            const size_t n = IX(r, k, i), a = AX(r, k, i);
            const ora_mp dzVx = (r->Vx[n + 1] - r->Vx[n]) * r20z;
            const ora_mp dxVz = (r->Vz[n + SI] - r->Vz[n]) * r20x;
            const float muxz = mu_harm(r->mu[n], r->mu[n + 1], r->mu[n + SI], r->mu[n + 1 + SI]);
            r->Sxz[n] = r->Sxz[n] + muxz * (gxe[0] * dxVz + gze[0] * dzVx + gxe[1] * axVz[a] + gze[1] * azVx[a]) * dt;
            azVx[a] = (float)(gze[2] * azVx[a] + gze[3] * dzVx * dt);
            axVz[a] = (float)(gxe[2] * axVz[a] + gxe[3] * dxVz * dt);
        }
    }
}

/* m_absorb_c.f90:98-125 / :127-151 */
static void absorb_c_update_stress(psv_rank *r) {
#pragma omp parallel for
    for (int i = r->ibeg; i <= r->iend; i++)
        for (int k = r->kbeg; k <= r->kend; k++) {
            const float gc = r->gx_c[i - r->ibeg_m] * r->gz_c[k - r->kbeg_m], gb = r->gx_b[i - r->ibeg_m] * r->gz_b[k - r->kbeg_m];
            const size_t n = IX(r, k, i);
            r->Sxx[n] = r->Sxx[n] * gc; r->Szz[n] = r->Szz[n] * gc; r->Sxz[n] = r->Sxz[n] * gb;
        }
}
static void absorb_c_update_vel(psv_rank *r) {
#pragma omp parallel for
    for (int i = r->ibeg; i <= r->iend; i++)
        for (int k = r->kbeg; k <= r->kend; k++) {
            const size_t n = IX(r, k, i);
            r->Vx[n] = r->Vx[n] * r->gx_b[i - r->ibeg_m] * r->gz_c[k - r->kbeg_m];
            r->Vz[n] = r->Vz[n] * r->gx_c[i - r->ibeg_m] * r->gz_b[k - r->kbeg_m];
        }
}

/* m_source.f90:550-589 */
static void stressglut(const psv_cfg *c, psv_rank *r, int it) {
    if (c->bf_mode) return;
    const size_t SI = (size_t)r->nzm;
    for (int i = 0; i < r->nsrc; i++) {
        const float t = c->tbeg + ((float)it - 0.5f) * c->dt;
        const ora_mp stime = ora_momentrate(t, c->stftype, &r->srcprm[2 * i]);
        const ora_mp sdrop = r->mo[i] * stime * c->dt_dxz;
        const size_t n = IX(r, r->ksrc[i], r->isrc[i]);
        r->Sxx[n] = r->Sxx[n] - r->mxx[i] * sdrop;
        r->Szz[n] = r->Szz[n] - r->mzz[i] * sdrop;
        r->Sxz[n] = r->Sxz[n] - r->mxz[i] * sdrop / 4;
        r->Sxz[n - 1] = r->Sxz[n - 1] - r->mxz[i] * sdrop / 4;
        r->Sxz[n - SI] = r->Sxz[n - SI] - r->mxz[i] * sdrop / 4;
        r->Sxz[n - 1 - SI] = r->Sxz[n - 1 - SI] - r->mxz[i] * sdrop / 4;
    }
}
/* m_source.f90:591-629 */
static void bodyforce(const psv_cfg *c, psv_rank *r, int it) {
    if (!c->bf_mode) return;
    const size_t SI = (size_t)r->nzm;
    for (int i = 0; i < r->nsrc; i++) {
        const float t = c->tbeg + it * c->dt;
        const ora_mp stime = ora_momentrate(t, c->stftype, &r->srcprm[2 * i]);
        const size_t n = IX(r, r->ksrc[i], r->isrc[i]);
        const float rho = r->rho[n];
        r->Vx[n] = r->Vx[n] + r->fx[i] / rho * stime * c->dt_dxz / 2;
        r->Vx[n - SI] = r->Vx[n - SI] + r->fx[i] / rho * stime * c->dt_dxz / 2;
        r->Vz[n] = r->Vz[n] + r->fz[i] / rho * stime * c->dt_dxz / 2;
        r->Vz[n - 1] = r->Vz[n - 1] + r->fz[i] / rho * stime * c->dt_dxz / 2;
    }
}

/* m_global.f90:312-364 / :366-418.  which = 0 stress, 1 velocity.  MPI_PROC_NULL neighbours: the receive buffer is never
 * written, taken as zeros (Q-psv 2), and unpacked all the same. */
static void comm(psv_sim *s, int which) {
    const int nz = s->cfg.nz;
    for (int q = 0; q < s->nranks; q++) {
        psv_rank *r = &s->r[q];
        const size_t a1 = IX(r, 1, r->iend - 1), a0 = IX(r, 1, r->iend), b0 = IX(r, 1, r->ibeg), b1 = IX(r, 1, r->ibeg + 1);
        for (int k = 0; k < nz; k++) {
            if (which == 1) {
                r->sbuf_ip[k] = r->Vx[a1 + k]; r->sbuf_ip[nz + k] = r->Vx[a0 + k]; r->sbuf_ip[2 * nz + k] = r->Vz[a0 + k];
                r->sbuf_im[k] = r->Vx[b0 + k]; r->sbuf_im[nz + k] = r->Vz[b0 + k]; r->sbuf_im[2 * nz + k] = r->Vz[b1 + k];
            } else {
                r->sbuf_ip[k] = r->Sxx[a0 + k]; r->sbuf_ip[nz + k] = r->Sxz[a1 + k]; r->sbuf_ip[2 * nz + k] = r->Sxz[a0 + k];
                r->sbuf_im[k] = r->Sxx[b0 + k]; r->sbuf_im[nz + k] = r->Sxx[b1 + k]; r->sbuf_im[2 * nz + k] = r->Sxz[b0 + k];
            }
        }
    }
    for (int q = 0; q < s->nranks; q++) {
        psv_rank *r = &s->r[q];
        /* rbuf_ip <- sbuf_im of idx+1 ; rbuf_im <- sbuf_ip of idx-1 */
        if (q + 1 < s->nranks) memcpy(r->rbuf_ip, s->r[q + 1].sbuf_im, sizeof(ora_mp) * 3 * (size_t)nz);
        else memset(r->rbuf_ip, 0, sizeof(ora_mp) * 3 * (size_t)nz);
        if (q > 0) memcpy(r->rbuf_im, s->r[q - 1].sbuf_ip, sizeof(ora_mp) * 3 * (size_t)nz);
        else memset(r->rbuf_im, 0, sizeof(ora_mp) * 3 * (size_t)nz);
        const size_t m2 = IX(r, 1, r->ibeg - 2), m1 = IX(r, 1, r->ibeg - 1), p1 = IX(r, 1, r->iend + 1), p2 = IX(r, 1, r->iend + 2);
        for (int k = 0; k < nz; k++) {
            if (which == 1) {
                r->Vx[m2 + k] = r->rbuf_im[k]; r->Vx[m1 + k] = r->rbuf_im[nz + k]; r->Vz[m1 + k] = r->rbuf_im[2 * nz + k];
                r->Vx[p1 + k] = r->rbuf_ip[k]; r->Vz[p1 + k] = r->rbuf_ip[nz + k]; r->Vz[p2 + k] = r->rbuf_ip[2 * nz + k];
            } else {
                r->Sxx[m1 + k] = r->rbuf_im[k]; r->Sxz[m2 + k] = r->rbuf_im[nz + k]; r->Sxz[m1 + k] = r->rbuf_im[2 * nz + k];
                r->Sxx[p1 + k] = r->rbuf_ip[k]; r->Sxx[p2 + k] = r->rbuf_ip[nz + k]; r->Sxz[p1 + k] = r->rbuf_ip[2 * nz + k];
            }
        }
    }
}

/* m_wav.f90:143-306 */
static void wav_store(psv_sim *s, int it) {
    const psv_cfg *c = &s->cfg;
    const float dt = c->dt, M0 = c->M0, UC = c->UC;
    for (int q = 0; q < s->nranks; q++) {
        psv_rank *r = &s->r[q];
        if (r->nst == 0) continue;
        const size_t SI = (size_t)r->nzm;
        const ora_mp *Vx = r->Vx, *Vz = r->Vz;
        if (c->sw_u)
            for (int n = 0; n < r->nst; n++) {
                const size_t p = IX(r, r->kst[n], r->ist[n]);
                r->ux[n] = r->ux[n] + (float)(Vx[p] + Vx[p - SI]) * 0.5f * dt;
                r->uz[n] = r->uz[n] - (float)(Vz[p] + Vz[p - 1]) * 0.5f * dt;
            }
        if (c->sw_strain)
            for (int n = 0; n < r->nst; n++) {
                const size_t p = IX(r, r->kst[n], r->ist[n]);
                const ora_mp r40x = c->w40x, r40z = c->w40z, r41x = c->w41x, r41z = c->w41z;
                const ora_mp dxVx = (Vx[p] - Vx[p - SI]) * r40x - (Vx[p + SI] - Vx[p - 2 * SI]) * r41x;
                const ora_mp dzVz = (Vz[p] - Vz[p - 1]) * r40z - (Vz[p + 1] - Vz[p - 2]) * r41z;
                const ora_mp dxVz = ((Vz[p + SI] - Vz[p]) * r40x - (Vz[p + 2 * SI] - Vz[p - SI]) * r41x
                                     + (Vz[p - 1 + SI] - Vz[p - 1]) * r40x - (Vz[p - 1 + 2 * SI] - Vz[p - 1 - SI]) * r41x
                                     + (Vz[p] - Vz[p - SI]) * r40x - (Vz[p + SI] - Vz[p - 2 * SI]) * r41x
                                     + (Vz[p - 1] - Vz[p - 1 - SI]) * r40x - (Vz[p - 1 + SI] - Vz[p - 1 - 2 * SI]) * r41x) / 4.0f;
                const ora_mp dzVx = ((Vx[p + 1] - Vx[p]) * r40z - (Vx[p + 2] - Vx[p - 1]) * r41z
                                     + (Vx[p + 1 - SI] - Vx[p - SI]) * r40z - (Vx[p + 2 - SI] - Vx[p - 1 - SI]) * r41z
                                     + (Vx[p] - Vx[p - 1]) * r40z - (Vx[p + 1] - Vx[p - 2]) * r41z
                                     + (Vx[p - SI] - Vx[p - 1 - SI]) * r40z - (Vx[p + 1 - SI] - Vx[p - 2 - SI]) * r41z) / 4.0f;
                r->exx[n] = r->exx[n] + (float)(dxVx) * dt;
                r->ezz[n] = r->ezz[n] + (float)(dzVz) * dt;
                r->exz[n] = r->exz[n] + (float)(dxVz + dzVx) / 2.0f * dt;
            }
        if ((it - 1) % c->ntdec_w != 0) continue;
        const int itw = (it - 1) / c->ntdec_w + 1;
        const size_t ntw = (size_t)c->ntw;
        for (int n = 0; n < r->nst; n++) {
            const size_t p = IX(r, r->kst[n], r->ist[n]);
            if (c->sw_v) {
                float *o = r->wav[0] + ntw * 2 * n + (itw - 1);
                o[0] = (float)(Vx[p] + Vx[p - SI]) / 2.0f * M0 * UC * 1e9f;
                o[ntw] = -(float)(Vz[p] + Vz[p - 1]) / 2.0f * M0 * UC * 1e9f;
            }
            if (c->sw_u) {
                float *o = r->wav[1] + ntw * 2 * n + (itw - 1);
                o[0] = r->ux[n] * M0 * UC * 1e9f;
                o[ntw] = r->uz[n] * M0 * UC * 1e9f;
            }
            if (c->sw_stress) {
                float *o = r->wav[2] + ntw * 3 * n + (itw - 1);
                o[0] = (float)(r->Sxx[p]) * M0 * UC * 1e6f;
                o[ntw] = (float)(r->Szz[p]) * M0 * UC * 1e6f;
                o[2 * ntw] = (float)(r->Sxz[p] + r->Sxz[p - SI] + r->Sxz[p - 1] + r->Sxz[p - 1 - SI]) / 4.0f * M0 * UC * 1e6f;
            }
            if (c->sw_strain) {
                float *o = r->wav[3] + ntw * 3 * n + (itw - 1);
                o[0] = r->exx[n] * M0 * UC * 1e-3f;
                o[ntw] = r->ezz[n] * M0 * UC * 1e-3f;
                o[2 * ntw] = r->exz[n] * M0 * UC * 1e-3f;
            }
        }
    }
}

/* m_kernel.f90:313-327 + m_report.f90:133-145 */
void psv_vmax(psv_sim *s, float out[2]) {
    float xa = 0.0f, za = 0.0f;
    for (int q = 0; q < s->nranks; q++) {
        const psv_rank *r = &s->r[q];
        float xm = 0.0f, zm = 0.0f;
        for (int i = r->ibeg_k; i <= r->iend_k; i++) {
            const size_t n = IX(r, r->kob[i - r->ibeg_m] + 1, i);
            const float ax = fabsf((float)r->Vx[n]), az = fabsf((float)r->Vz[n]);
            if (ax > xm) xm = ax;
            if (az > zm) zm = az;
        }
        if (q == 0 || xm > xa) xa = xm;
        if (q == 0 || zm > za) za = zm;
    }
    out[0] = xa * s->cfg.UC * s->cfg.M0;
    out[1] = za * s->cfg.UC * s->cfg.M0;
}

void psv_step(psv_sim *s, int it) {
    const psv_cfg *c = &s->cfg;
    const int pml = !strcmp(c->abc_type, "pml");
    snap_write(s, it);
    wav_store(s, it);
    for (int q = 0; q < s->nranks; q++) {
        psv_rank *r = &s->r[q];
        kernel_update_stress(c, r);
        if (pml && c->pw_mode) pw_edges(c, r, c->nproc_x, 0);   /* absorb_p__update_stress extrapolates the velocities */
        if (pml) absorb_p_update_stress(c, r); else absorb_c_update_stress(r);
        stressglut(c, r, it);
    }
    comm(s, 0);
    for (int q = 0; q < s->nranks; q++) {
        psv_rank *r = &s->r[q];
        kernel_update_vel(c, r);
        bodyforce(c, r, it);
        if (pml && c->pw_mode) pw_edges(c, r, c->nproc_x, 1);   /* absorb_p__update_vel extrapolates the stresses */
        if (pml) absorb_p_update_vel(c, r); else absorb_c_update_vel(r);
    }
    comm(s, 1);
}

int psv_run(psv_sim *s, int it0, int it1, float *vm, int nvm) {
    int rec = 0;
    for (int it = it0; it <= it1; it++) {
        if (it % s->cfg.ntdec_r == 0 && vm && rec < nvm) { psv_vmax(s, vm + 2 * rec); rec++; }
        psv_step(s, it);
    }
    return rec;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* snapshot accessors                                                                                              */
int psv_snap_info(const psv_sim *s, int *info) {   /* idec kdec ntdec_s nxs nzs sw_ps sw_v sw_u */
    const psv_snap *sn = &s->snap;
    info[0] = sn->idec; info[1] = sn->kdec; info[2] = sn->ntdec_s; info[3] = sn->nxs; info[4] = sn->nzs;
    for (int p = 0; p < 3; p++) info[5 + p] = sn->sw[p];
    return 0;
}
int psv_snap_coords(const psv_sim *s, float *x, float *z) {
    memcpy(x, s->snap.xsnp, sizeof(float) * (size_t)s->snap.nxs);
    memcpy(z, s->snap.zsnp, sizeof(float) * (size_t)s->snap.nzs);
    return 0;
}
int psv_snap_nrec(const psv_sim *s, int p) { return (p >= 0 && p < 3) ? s->snap.nrec[p] : -1; }
int psv_snap_rec(const psv_sim *s, int p, int rec, float *out, int *it0) {   /* out (2, nzs, nxs) */
    const psv_snap *sn = &s->snap;
    if (p < 0 || p > 2 || rec < 0 || rec >= sn->nrec[p]) return -1;
    const size_t n = 2 * (size_t)sn->nxs * sn->nzs;
    memcpy(out, sn->rec[p] + n * (size_t)rec, sizeof(float) * n);
    *it0 = sn->it0[p][rec];
    return 0;
}
int psv_snap_medium(const psv_sim *s, int which, float *out) {   /* 0 rho 1 lambda 2 mu: (nzs, nxs) */
    const size_t n2 = (size_t)s->snap.nxs * s->snap.nzs;
    memcpy(out, s->snap.medium + n2 * (size_t)which, sizeof(float) * n2);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* accessors                                                                                                       */
int psv_nranks(const psv_sim *s) { return s->nranks; }
int psv_rank_int(const psv_sim *s, int q, int which) {
    const psv_rank *r = &s->r[q];
    const int v[] = {r->ibeg, r->iend, r->ibeg_k, r->iend_k, r->kbeg_k, r->kend_k, r->nsrc, r->nst, r->nzm, r->nxm, r->ibeg_m, r->kbeg_m, r->kbeg_min, r->nxp};
    return (which >= 0 && which < (int)(sizeof(v) / sizeof(v[0]))) ? v[which] : -1;
}
double psv_cfg_value(const psv_sim *s, int which) {
    const psv_cfg *c = &s->cfg;
    if (which >= 15 && which < 15 + PSV_MAXNM) return c->ts[which - 15];
    if (which >= 23 && which < 23 + PSV_MAXNM) return c->c1[which - 23];
    if (which >= 31 && which < 31 + PSV_MAXNM) return c->c2[which - 31];
    if (which >= 39 && which < 39 + PSV_MAXNM) return c->d1[which - 39];
    const double v[] = {c->vmin, c->vmax, c->fmax, c->fcut, c->M0, c->UC, c->zeta, c->d2, c->dt, c->xbeg, c->zbeg, c->dx, c->dz, c->evlo, c->evla};
    if (which >= 0 && which < 15) return v[which];
    if (which == 47) return c->tbeg;
    if (which == 48) return c->r20x;
    if (which == 49) return c->r20z;
    return 0.0;
}
int psv_cfg_int(const psv_sim *s, int which) {
    const psv_cfg *c = &s->cfg;
    const int v[] = {c->nx, c->nz, c->nt, c->na, c->nm, c->nproc_x, c->ntw, c->ntdec_w, c->ntdec_r, c->bf_mode, c->pw_mode, c->sw_v, c->sw_u, c->sw_stress, c->sw_strain};
    return (which >= 0 && which < (int)(sizeof(v) / sizeof(v[0]))) ? v[which] : -1;
}
const char *psv_cfg_str(const psv_sim *s, int which) {
    const psv_cfg *c = &s->cfg;
    return which == 0 ? c->title : which == 1 ? c->odir : which == 2 ? c->abc_type : which == 3 ? c->stftype : "";
}
static int field_ptr(const psv_rank *r, const char *name, ora_mp **mp, float **sp) {
    *mp = NULL; *sp = NULL;
    if (!strcmp(name, "Vx")) *mp = r->Vx; else if (!strcmp(name, "Vz")) *mp = r->Vz; else if (!strcmp(name, "Sxx")) *mp = r->Sxx;
    else if (!strcmp(name, "Szz")) *mp = r->Szz; else if (!strcmp(name, "Sxz")) *mp = r->Sxz; else if (!strcmp(name, "rho")) *sp = r->rho;
    else if (!strcmp(name, "lam")) *sp = r->lam; else if (!strcmp(name, "mu")) *sp = r->mu; else if (!strcmp(name, "taup")) *sp = r->taup;
    else if (!strcmp(name, "taus")) *sp = r->taus; else return -1;
    return 0;
}
int psv_get_field(const psv_sim *s, int q, const char *name, double *out) {
    const psv_rank *r = &s->r[q];
    ora_mp *mp; float *sp;
    if (field_ptr(r, name, &mp, &sp)) return -1;
    for (size_t n = 0; n < r->ncell; n++) out[n] = mp ? (double)mp[n] : (double)sp[n];
    return 0;
}
int psv_set_field(psv_sim *s, int q, const char *name, const double *in) {
    psv_rank *r = &s->r[q];
    ora_mp *mp; float *sp;
    if (field_ptr(r, name, &mp, &sp)) return -1;
    for (size_t n = 0; n < r->ncell; n++) { if (mp) mp[n] = (ora_mp)in[n]; else sp[n] = (float)in[n]; }
    return 0;
}
/* re-run surface_detection after a test replaced the medium */
void psv_redetect_surface(psv_sim *s) { for (int q = 0; q < s->nranks; q++) surface_detection(&s->r[q]); }
int psv_get_memvar(const psv_sim *s, int q, const char *name, float *out) {   /* Rxx Rzz Rxz: (m, k, i) */
    const psv_rank *r = &s->r[q];
    const float *p = !strcmp(name, "Rxx") ? r->Rxx : !strcmp(name, "Rzz") ? r->Rzz : !strcmp(name, "Rxz") ? r->Rxz : NULL;
    if (!p) return -1;
    memcpy(out, p, sizeof(float) * r->ncell * (size_t)s->cfg.nm);
    return 0;
}
static int *map_ptr(const psv_rank *r, const char *name) {
    return !strcmp(name, "kfs") ? r->kfs : !strcmp(name, "kob") ? r->kob : !strcmp(name, "kfs_top") ? r->kfs_top : !strcmp(name, "kfs_bot") ? r->kfs_bot
         : !strcmp(name, "kob_top") ? r->kob_top : !strcmp(name, "kob_bot") ? r->kob_bot : !strcmp(name, "kbeg_a") ? r->kbeg_a : NULL;
}
int psv_get_map(const psv_sim *s, int q, const char *name, int *out) {
    const psv_rank *r = &s->r[q];
    const int *p = map_ptr(r, name);
    if (!p) return -1;
    memcpy(out, p, sizeof(int) * (size_t)r->nxm);
    return 0;
}
int psv_get_profile(const psv_sim *s, int q, const char *name, float *out) {
    const psv_rank *r = &s->r[q];
    const int nxo = r->iend - r->ibeg + 1, nz = s->cfg.nz;
    const float *p = NULL; size_t n = 0;
    if (!strcmp(name, "gxc")) { p = r->gxc; n = 4 * (size_t)nxo; } else if (!strcmp(name, "gxe")) { p = r->gxe; n = 4 * (size_t)nxo; }
    else if (!strcmp(name, "gzc")) { p = r->gzc; n = 4 * (size_t)nz; } else if (!strcmp(name, "gze")) { p = r->gze; n = 4 * (size_t)nz; }
    else if (!strcmp(name, "gx_c")) { p = r->gx_c; n = (size_t)r->nxm; } else if (!strcmp(name, "gx_b")) { p = r->gx_b; n = (size_t)r->nxm; }
    else if (!strcmp(name, "gz_c")) { p = r->gz_c; n = (size_t)r->nzm; } else if (!strcmp(name, "gz_b")) { p = r->gz_b; n = (size_t)r->nzm; }
    if (!p) return -1;
    memcpy(out, p, sizeof(float) * n);
    return (int)n;
}
int psv_get_sources(const psv_sim *s, int q, int *ik, double *val) {
    const psv_rank *r = &s->r[q];
    for (int i = 0; i < r->nsrc; i++) {
        ik[2 * i] = r->isrc[i]; ik[2 * i + 1] = r->ksrc[i];
        val[6 * i] = (double)r->mo[i];
        val[6 * i + 1] = s->cfg.bf_mode ? (double)r->fx[i] : (double)r->mxx[i];
        val[6 * i + 2] = s->cfg.bf_mode ? (double)r->fz[i] : (double)r->mzz[i];
        val[6 * i + 3] = (double)r->mxz[i];
        val[6 * i + 4] = r->srcprm[2 * i]; val[6 * i + 5] = r->srcprm[2 * i + 1];
    }
    return r->nsrc;
}
int psv_get_stations(const psv_sim *s, int q, int *ik, char *names9) {
    const psv_rank *r = &s->r[q];
    for (int i = 0; i < r->nst; i++) { ik[2 * i] = r->ist[i]; ik[2 * i + 1] = r->kst[i]; memcpy(names9 + 9 * i, r->stnm[i], 9); }
    return r->nst;
}
int psv_get_wav(const psv_sim *s, int q, int prod, float *out) {
    const psv_rank *r = &s->r[q];
    if (prod < 0 || prod > 3 || !r->wav[prod]) return 0;
    const size_t n = (size_t)s->cfg.ntw * (prod < 2 ? 2 : 3) * r->nst;
    memcpy(out, r->wav[prod], sizeof(float) * n);
    return (int)n;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* SAC: m_wav.f90:573-694 (set_sac_header, initialize_sac_header, export_wav__sac) + src/shared/m_sac.f90:314-449    */
typedef struct { float f[70]; int32_t i[35]; int32_t l[5]; char a[192]; } sac_raw;
static void put8(char *dst, const char *src, int n) {
    const int l = (int)strlen(src);
    for (int q = 0; q < n; q++) dst[q] = (q < l) ? src[q] : ' ';
}
static const char *cmpnm(int prod, int cmp) {
    static const char *nm[4][3] = {{"Vx", "Vz", ""}, {"Ux", "Uz", ""}, {"Sxx", "Szz", "Sxz"}, {"Exx", "Ezz", "Exz"}};
    return nm[prod][cmp];
}
static void sac_header(const psv_cfg *c, const psv_rank *r, int n, int prod, int cmp, sac_raw *h) {
    for (int q = 0; q < 70; q++) h->f[q] = -12345.0f;
    for (int q = 0; q < 35; q++) h->i[q] = -12345;
    for (int q = 0; q < 5; q++) h->l[q] = 0;
    for (int q = 0; q < 24; q++) put8(h->a + 8 * q, "-12345", 8);
    put8(h->a + 8, "-12345", 16);
    const double delta = (double)(c->ntdec_w * c->dt);
    h->f[0] = (float)((int)(delta * 1e7)) / 1e7f;
    h->f[5] = c->tbeg;
    h->f[7] = c->otim;
    h->f[31] = r->stla[n]; h->f[32] = r->stlo[n];
    h->f[34] = r->zst[n] * 1000;
    h->f[35] = c->evla; h->f[36] = c->evlo; h->f[38] = c->evdp;
    h->f[39] = ora_moment_magnitude(c->M0);
    if (c->bf_mode) { h->f[40] = c->fx0; h->f[42] = c->fz0; }
    else { h->f[40] = c->mxx0; h->f[42] = c->mzz0; h->f[44] = c->mxz0; }
    h->f[46] = c->clon; h->f[47] = c->clat; h->f[48] = c->phi;
    const float dd = c->sx0 - r->xst[n];
    h->f[50] = sqrtf(dd * dd);
    h->f[51] = ora_rad2deg_s(atan2f(0.0f, r->xst[n] - c->sx0));
    h->f[52] = ora_rad2deg_s(atan2f(0.0f, c->sx0 - r->xst[n]));
    if (prod < 2) { h->f[58] = 90.0f; h->f[57] = (cmp == 0) ? 0.0f + c->phi : 0.0f; }   /* :583-584, :593-594 */
    time_t tt = (time_t)c->exedate + (time_t)c->tz_minutes * 60;
    struct tm g;
    gmtime_r(&tt, &g);
    h->i[0] = g.tm_year + 1900; h->i[1] = g.tm_yday + 1; h->i[2] = g.tm_hour; h->i[3] = g.tm_min; h->i[4] = g.tm_sec; h->i[5] = 0;
    h->i[6] = 6; h->i[9] = c->ntw; h->i[15] = 1;
    h->i[16] = prod == 0 ? 7 : prod == 1 ? 6 : 5;
    h->l[0] = 1; h->l[1] = 0; h->l[2] = 1; h->l[3] = 0; h->l[4] = 0;
    put8(h->a + 0, r->stnm[n], 8);
    char t16[17];
    const char *t = c->title;
    while (*t == ' ') t++;
    strncpy(t16, t, 16);
    t16[16] = 0;
    put8(h->a + 8, t16, 16);
    put8(h->a + 8 * 20, cmpnm(prod, cmp), 8);
}
static void mkdir_p(const char *path) {
    char tmp[1024];
    snprintf(tmp, sizeof(tmp), "%s", path);
    for (char *p = tmp + 1; *p; p++) if (*p == '/') { *p = 0; mkdir(tmp, 0777); *p = '/'; }
    mkdir(tmp, 0777);
}
int psv_write_sac(psv_sim *s, const char *odir) {
    const psv_cfg *c = &s->cfg;
    char dir[1024];
    snprintf(dir, sizeof(dir), "%s/wav", odir);
    mkdir_p(dir);
    int nfiles = 0;
    for (int q = 0; q < s->nranks; q++) {
        const psv_rank *r = &s->r[q];
        for (int n = 0; n < r->nst; n++)
            for (int prod = 0; prod < 4; prod++) {
                if (!r->wav[prod]) continue;
                const int ncmp = prod < 2 ? 2 : 3;
                for (int cmp = 0; cmp < ncmp; cmp++) {
                    sac_raw h;
                    sac_header(c, r, n, prod, cmp, &h);
                    char fn[1400];
                    snprintf(fn, sizeof(fn), "%s/%s.psv.%s.%s.sac", dir, c->title, r->stnm[n], cmpnm(prod, cmp));
                    FILE *fp = fopen(fn, "wb");
                    if (!fp) return -1;
                    fwrite(h.f, 4, 70, fp); fwrite(h.i, 4, 35, fp); fwrite(h.l, 4, 5, fp); fwrite(h.a, 1, 192, fp);
                    fwrite(r->wav[prod] + (size_t)c->ntw * ncmp * n + (size_t)c->ntw * cmp, 4, (size_t)c->ntw, fp);
                    fclose(fp);
                    nfiles++;
                }
            }
    }
    return nfiles;
}
