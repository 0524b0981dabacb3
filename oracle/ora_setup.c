/*
 * oracle/ora_setup.c -- restatement of swpc_3d's setup chain (TEST INFRASTRUCTURE, see ora.h).
 *
 * Order follows src/swpc_3d/main.f90:64-78:
 *   global__setup (m_global.f90:180-216, readprm :124-177) -> global__setup2 (:219-388) ->
 *   medium__setup (m_medium.f90:36-227) -> kernel__setup (m_kernel.f90:30-73) ->
 *   source__setup (m_source.f90:41-314) -> absorb__setup (m_absorb_p.f90:60-124 /
 *   m_absorb_c.f90:27-111) -> wav__setup (m_wav.f90:54-271).
 * MPI ranks are emulated: every ora_rank is set up exactly as the rank with that myid would be.
 */
#include "ora.h"

#include <complex.h>
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

double ora_r_earth(void);
double ora_pi(void);

static char g_err[512] = "";
const char *ora_last_error(void) { return g_err; }
static void set_err(const char *msg) {
    strncpy(g_err, msg, sizeof(g_err) - 1);
    g_err[sizeof(g_err) - 1] = 0;
}

#define FLT_EPS 1.1920929e-07f /* epsilon(1.0) */

static void resolve_path(const char *base, const char *fn, char *out, size_t cap) {
    if (fn[0] == '/' || !base || !base[0])
        snprintf(out, cap, "%s", fn);
    else
        snprintf(out, cap, "%s/%s", base, fn);
}

static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) {
        fprintf(stderr, "[oracle] out of memory (%zu x %zu)\n", n, sz);
        abort();
    }
    return p;
}

/* ------------------------------------------------------------------------------------------ */
/* m_global.f90:124-177 global__readprm, :196-206 derived values                               */
static void global_setup(ora_cfg *c, const ora_ini *ini) {
    ora_readini_l(ini, "benchmark_mode", &c->benchmark_mode, 0);
    ora_readini_c(ini, "title", c->title, "swpc3d");
    ora_readini_i(ini, "nproc_x", &c->nproc_x, 1);
    ora_readini_i(ini, "nproc_y", &c->nproc_y, 2);
    ora_readini_i(ini, "nx", &c->nx, 256);
    ora_readini_i(ini, "ny", &c->ny, 256);
    ora_readini_i(ini, "nz", &c->nz, 256);
    ora_readini_i(ini, "nt", &c->nt, 1000);
    ora_readini_i(ini, "ipad", &c->ipad, 0);
    ora_readini_i(ini, "jpad", &c->jpad, 0);
    ora_readini_i(ini, "kpad", &c->kpad, 0);
    ora_readini_c(ini, "odir", c->odir, "./out");

    if (c->benchmark_mode) {
        c->dx = 0.5f; /* "dx = 0.5": a default-real literal assigned to real(MP) */
        c->dy = 0.5f;
        c->dz = 0.5f;
        c->dt = 0.04f;
        c->na = 20;
        c->xbeg = -((float)c->nx / 2.0f * (float)c->dx);
        c->ybeg = -((float)c->ny / 2.0f * (float)c->dy);
        c->zbeg = -30 * (float)c->dz;
        c->tbeg = 0.0f;
        c->clon = 139.7604f;
        c->clat = 35.7182f;
        c->phi = 0.0f;
        strcpy(c->abc_type, "pml");
    } else {
        ora_readini_d(ini, "dx", &c->dx, 0.5);
        ora_readini_d(ini, "dy", &c->dy, 0.5);
        ora_readini_d(ini, "dz", &c->dz, 0.5);
        ora_readini_s(ini, "dt", &c->dt, 0.01f);
        ora_readini_i(ini, "na", &c->na, 20);
        ora_readini_s(ini, "xbeg", &c->xbeg, -(float)(c->nx / 2) * (float)c->dx);
        ora_readini_s(ini, "ybeg", &c->ybeg, -(float)(c->ny / 2) * (float)c->dy);
        ora_readini_s(ini, "zbeg", &c->zbeg, -30 * (float)c->dz);
        ora_readini_s(ini, "tbeg", &c->tbeg, 0.0f);
        ora_readini_s(ini, "clon", &c->clon, 139.7604f);
        ora_readini_s(ini, "clat", &c->clat, 35.7182f);
        ora_readini_s(ini, "phi", &c->phi, 0.0f);
        char abc[ORA_STRLEN];
        ora_readini_c(ini, "abc_type", abc, "pml");
        strncpy(c->abc_type, abc, sizeof(c->abc_type) - 1);
    }
    c->nproc = c->nproc_x * c->nproc_y;
    c->xend = c->xbeg + c->nx * (float)c->dx;
    c->yend = c->ybeg + c->ny * (float)c->dy;
    c->zend = c->zbeg + c->nz * (float)c->dz;
    c->tend = c->tbeg + c->nt * c->dt;
    c->UC = 1e-15f; /* m_global.f90:29 */
}

/* m_global.f90:219-388 global__setup2 for the rank with this myid */
static void rank_geometry(const ora_cfg *c, ora_rank *r, int myid) {
    memset(r, 0, sizeof(*r));
    r->myid = myid;
    int proc_x = myid % c->nproc_x;
    int proc_y = myid / c->nproc_x;
    r->idx = proc_x;
    r->idy = proc_y;
    ora_decomp1d(c->nx, c->nproc_x, proc_x, &r->nxp, &r->ibeg, &r->iend);
    ora_decomp1d(c->ny, c->nproc_y, proc_y, &r->nyp, &r->jbeg, &r->jend);
    r->kbeg = 1;
    r->kend = c->nz;
    r->ibeg_m = r->ibeg - 3;
    r->iend_m = r->iend + 3 + c->ipad;
    r->jbeg_m = r->jbeg - 3;
    r->jend_m = r->jend + 3 + c->jpad;
    r->kbeg_m = r->kbeg - 3;
    r->kend_m = r->kend + 3 + c->kpad;
    r->nzm = r->kend_m - r->kbeg_m + 1;
    r->nxm = r->iend_m - r->ibeg_m + 1;
    r->nym = r->jend_m - r->jbeg_m + 1;
    r->ncell_m = (size_t)r->nzm * (size_t)r->nxm * (size_t)r->nym;

    r->xc = (float *)xcalloc((size_t)r->nxm, sizeof(float));
    r->yc = (float *)xcalloc((size_t)r->nym, sizeof(float));
    r->zc = (float *)xcalloc((size_t)r->nzm, sizeof(float));
    for (int i = r->ibeg_m; i <= r->iend_m; i++) r->xc[i - r->ibeg_m] = ora_i2x(i, c->xbeg, (float)c->dx);
    for (int j = r->jbeg_m; j <= r->jend_m; j++) r->yc[j - r->jbeg_m] = ora_i2x(j, c->ybeg, (float)c->dy);
    for (int k = r->kbeg_m; k <= r->kend_m; k++) r->zc[k - r->kbeg_m] = ora_i2x(k, c->zbeg, (float)c->dz);

    /* m_global.f90:334-343 */
    r->kbeg_a = (int *)xcalloc((size_t)r->nxm * r->nym, sizeof(int));
    for (int j = r->jbeg_m; j <= r->jend_m; j++)
        for (int i = r->ibeg_m; i <= r->iend_m; i++) {
            if (i <= c->na || c->nx - c->na + 1 <= i || j <= c->na || c->ny - c->na + 1 <= j)
                r->kbeg_a[ora_idx2(r, i, j)] = r->kbeg;
            else
                r->kbeg_a[ora_idx2(r, i, j)] = r->kend - c->na + 1;
        }

    /* m_global.f90:348-376 */
    r->ibeg_k = r->ibeg; r->iend_k = r->iend;
    r->jbeg_k = r->jbeg; r->jend_k = r->jend;
    r->kbeg_k = r->kbeg; r->kend_k = r->kend;
    if (!strcmp(c->abc_type, "pml")) {
        int na = c->na, nx = c->nx, ny = c->ny;
        if (r->iend <= na) r->ibeg_k = r->iend + 1;
        else if (r->ibeg <= na) r->ibeg_k = na + 1;
        if (r->ibeg >= nx - na + 1) r->iend_k = r->ibeg - 1;
        else if (r->iend >= nx - na + 1) r->iend_k = nx - na;
        if (r->jend <= na) r->jbeg_k = r->jend + 1;
        else if (r->jbeg <= na) r->jbeg_k = na + 1;
        if (r->jbeg >= ny - na + 1) r->jend_k = r->jbeg - 1;
        else if (r->jend >= ny - na + 1) r->jend_k = ny - na;
        r->kend_k = c->nz - na;
    }

    /* m_global.f90:251-258 */
    size_t isz = (size_t)5 * r->nyp * c->nz, jsz = (size_t)5 * r->nxp * c->nz;
    r->sbuf_ip = (ora_mp *)xcalloc(isz, sizeof(ora_mp));
    r->sbuf_im = (ora_mp *)xcalloc(isz, sizeof(ora_mp));
    r->rbuf_ip = (ora_mp *)xcalloc(isz, sizeof(ora_mp));
    r->rbuf_im = (ora_mp *)xcalloc(isz, sizeof(ora_mp));
    r->sbuf_jp = (ora_mp *)xcalloc(jsz, sizeof(ora_mp));
    r->sbuf_jm = (ora_mp *)xcalloc(jsz, sizeof(ora_mp));
    r->rbuf_jp = (ora_mp *)xcalloc(jsz, sizeof(ora_mp));
    r->rbuf_jm = (ora_mp *)xcalloc(jsz, sizeof(ora_mp));
}

/* ------------------------------------------------------------------------------------------ */
/* velocity models                                                                             */

/* m_vmodel_uni.f90:21-147 */
static void vmodel_uni(const ora_ini *ini, ora_rank *r, float vcut, float *qp, float *qs) {
    (void)vcut;
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
    double RE = ora_r_earth();
    for (int j = r->jbeg_m; j <= r->jend_m; j++)
        for (int i = r->ibeg_m; i <= r->iend_m; i++) {
            float bd0 = topo0;
            r->bddep[ora_idx2(r, i, j)] = bd0;
            for (int k = r->kbeg_m; k <= r->kend_m; k++) {
                float zc = r->zc[k - r->kbeg_m];
                float zs = zc, Cv = 1.0f;
                if (ef) {
                    zs = (float)(RE - RE * exp(-(double)zc / RE));
                    Cv = (float)exp((double)zc / RE);
                }
                size_t n = ora_idx3(r, k, i, j);
                float vp1, vs1;
                if (zs > bd0) {
                    vp1 = Cv * vp0;
                    vs1 = Cv * vs0;
                    r->rho[n] = rho0;
                    r->mu[n] = r->rho[n] * vs1 * vs1;
                    r->lam[n] = r->rho[n] * (vp1 * vp1 - 2 * vs1 * vs1);
                    qp[n] = qp0;
                    qs[n] = qs0;
                } else if (zc > 0.0f) {
                    vp1 = Cv * ora_seawater_vel(zs, use_munk);
                    vs1 = 0.0f;
                    r->rho[n] = 1.0f;
                    r->mu[n] = r->rho[n] * vs1 * vs1;
                    r->lam[n] = r->rho[n] * (vp1 * vp1 - 2 * vs1 * vs1);
                    qp[n] = 1000000.0f;
                    qs[n] = 1000000.0f;
                } else {
                    vp1 = 0.0f;
                    vs1 = 0.0f;
                    r->rho[n] = 0.001f;
                    r->mu[n] = r->rho[n] * vs1 * vs1;
                    r->lam[n] = r->rho[n] * (vp1 * vp1 - 2 * vs1 * vs1);
                    qp[n] = 10.0f;
                    qs[n] = 10.0f;
                }
            }
        }
    size_t n2 = (size_t)r->nxm * r->nym;
    for (int b = 1; b <= ORA_NBD; b++)
        for (size_t n = 0; n < n2; n++) r->bddep[(size_t)b * n2 + n] = -9999.0f;
}

/* read whitespace/comma separated reals from a data line */
static int parse_floats(const char *line, double *v, int maxn) {
    int n = 0;
    const char *p = line;
    while (*p && n < maxn) {
        while (*p == ' ' || *p == '\t' || *p == ',') p++;
        if (!*p) break;
        char tok[64];
        int l = 0;
        while (*p && *p != ' ' && *p != '\t' && *p != ',' && l < 63) {
            char ch = *p++;
            if (ch == 'd' || ch == 'D') ch = 'e';
            tok[l++] = ch;
        }
        tok[l] = 0;
        char *e;
        double x = strtod(tok, &e);
        if (e == tok) break;
        /* keep float parsing exact for SP targets: the caller re-parses with strtof when needed */
        v[n++] = x;
    }
    return n;
}
static int parse_floats_sp(const char *line, float *v, int maxn) {
    int n = 0;
    const char *p = line;
    while (*p && n < maxn) {
        while (*p == ' ' || *p == '\t' || *p == ',') p++;
        if (!*p) break;
        char tok[64];
        int l = 0;
        while (*p && *p != ' ' && *p != '\t' && *p != ',' && l < 63) {
            char ch = *p++;
            if (ch == 'd' || ch == 'D') ch = 'e';
            tok[l++] = ch;
        }
        tok[l] = 0;
        char *e;
        float x = strtof(tok, &e);
        if (e == tok) break;
        v[n++] = x;
    }
    return n;
}

static int is_blank(const char *s) {
    while (*s) {
        if (!isspace((unsigned char)*s)) return 0;
        s++;
    }
    return 1;
}

/* m_vmodel_lhm.f90:21-168 */
static int vmodel_lhm(const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs) {
    char fn[ORA_STRLEN], path[2 * ORA_STRLEN];
    int use_munk, ef;
    ora_readini_c(ini, "fn_lhm", fn, "");
    ora_readini_l(ini, "munk_profile", &use_munk, 0);
    ora_readini_l(ini, "earth_flattening", &ef, 0);
    resolve_path(base, fn, path, sizeof(path));
    FILE *fp = fopen(path, "r");
    if (!fp) {
        char m[600];
        snprintf(m, sizeof(m), "vmodel_lhm: cannot open %s", path);
        set_err(m);
        return -1;
    }
    float depth[256], rho0[256], vp0[256], vs0[256], qp0[256], qs0[256];
    int nl = 0;
    char line[512];
    while (fgets(line, sizeof(line), fp) && nl < 256) {
        char *p = line;
        while (*p == ' ' || *p == '\t') p++;
        if (is_blank(p) || *p == '#') continue;
        float v[6];
        if (parse_floats_sp(p, v, 6) < 6) continue;
        depth[nl] = v[0]; rho0[nl] = v[1]; vp0[nl] = v[2]; vs0[nl] = v[3]; qp0[nl] = v[4]; qs0[nl] = v[5];
        nl++;
    }
    fclose(fp);
    /* velocity cut-off :89-98 */
    for (int l = nl - 2; l >= 0; l--) {
        if ((vp0[l] < vcut || vs0[l] < vcut) && (vp0[l] > 0 && vs0[l] > 0)) {
            vp0[l] = vp0[l + 1]; vs0[l] = vs0[l + 1]; rho0[l] = rho0[l + 1];
            qp0[l] = qp0[l + 1]; qs0[l] = qs0[l + 1];
        }
    }
    size_t n2 = (size_t)r->nxm * r->nym;
    for (size_t n = 0; n < n2; n++) r->bddep[n] = depth[0];
    double RE = ora_r_earth();
    for (int k = r->kbeg_m; k <= r->kend_m; k++) {
        float zc = r->zc[k - r->kbeg_m];
        float zs = zc, Cv = 1.0f;
        if (ef) {
            zs = (float)(RE - RE * exp(-(double)zc / RE));
            Cv = (float)exp((double)zc / RE);
        }
        float rho1, mu1, lam1, qp1, qs1;
        if (zs < depth[0]) {
            if (zs < 0.0f) {
                rho1 = 0.001f; mu1 = 0.0f; lam1 = 0.0f; qp1 = 10.0f; qs1 = 10.0f;
            } else {
                float vp1 = Cv * ora_seawater_vel(zc, use_munk);
                rho1 = 1.0f; mu1 = 0.0f; lam1 = 1.0f * vp1 * vp1; qp1 = 1000000.0f; qs1 = 1000000.0f;
            }
        } else {
            float rr = 0, vp1 = 0, vs1 = 0;
            qp1 = qs1 = 0;
            for (int l = 0; l < nl; l++)
                if (zs >= depth[l]) {
                    rr = rho0[l]; vp1 = Cv * vp0[l]; vs1 = Cv * vs0[l]; qp1 = qp0[l]; qs1 = qs0[l];
                }
            rho1 = rr;
            mu1 = rr * vs1 * vs1;
            lam1 = rr * (vp1 * vp1 - 2 * vs1 * vs1);
        }
        for (int j = r->jbeg_m; j <= r->jend_m; j++)
            for (int i = r->ibeg_m; i <= r->iend_m; i++) {
                size_t n = ora_idx3(r, k, i, j);
                r->rho[n] = rho1; r->mu[n] = mu1; r->lam[n] = lam1; qp[n] = qp1; qs[n] = qs1;
            }
    }
    for (int b = 1; b <= ORA_NBD; b++)
        for (size_t n = 0; n < n2; n++) r->bddep[(size_t)b * n2 + n] = -9999.0f;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* m_medium.f90:36-227                                                                         */
static int medium_setup(ora_sim *s, ora_rank *r, const ora_ini *ini, const char *base, float *vmin1,
                        float *vmax1) {
    ora_cfg *c = &s->cfg;
    size_t nc = r->ncell_m, n2 = (size_t)r->nxm * r->nym;
    r->rho = (float *)xcalloc(nc, sizeof(float));
    r->lam = (float *)xcalloc(nc, sizeof(float));
    r->mu = (float *)xcalloc(nc, sizeof(float));
    r->taup = (float *)xcalloc(nc, sizeof(float));
    r->taus = (float *)xcalloc(nc, sizeof(float));
    r->kfs = (int *)xcalloc(n2, sizeof(int));
    r->kob = (int *)xcalloc(n2, sizeof(int));
    r->kfs_top = (int *)xcalloc(n2, sizeof(int));
    r->kfs_bot = (int *)xcalloc(n2, sizeof(int));
    r->kob_top = (int *)xcalloc(n2, sizeof(int));
    r->kob_bot = (int *)xcalloc(n2, sizeof(int));
    r->bddep = (float *)xcalloc(n2 * (ORA_NBD + 1), sizeof(float));
    int nm = c->nm, na = c->na, nx = c->nx, ny = c->ny, nz = c->nz;

    if (c->benchmark_mode) { /* :55-74 */
        c->fq_min = 0.05f; c->fq_max = 5.0f; c->fq_ref = 1.0f;
        for (int k = r->kbeg_m; k <= r->kend_m; k++) {
            float zc = r->zc[k - r->kbeg_m];
            float rr, mm, ll;
            if (zc < 0.0f) { rr = 0.001f; mm = 0.0f; ll = 0.0f; }
            else { rr = 2.7f; mm = 2.7f * 3.5f * 3.5f; ll = 2.7f * 3.5f * 3.5f; }
            for (int j = r->jbeg_m; j <= r->jend_m; j++)
                for (int i = r->ibeg_m; i <= r->iend_m; i++) {
                    size_t n = ora_idx3(r, k, i, j);
                    r->rho[n] = rr; r->mu[n] = mm; r->lam[n] = ll;
                    r->taup[n] = 1e10f; r->taus[n] = 1e10f;
                }
        }
    } else {
        ora_readini_s(ini, "fq_min", &c->fq_min, 0.05f);
        ora_readini_s(ini, "fq_max", &c->fq_max, 5.00f);
        ora_readini_s(ini, "fq_ref", &c->fq_ref, 1.00f);
        char vt[ORA_STRLEN];
        ora_readini_c(ini, "vmodel_type", vt, "uni");
        strncpy(c->vmodel_type, vt, sizeof(c->vmodel_type) - 1);
        ora_readini_s(ini, "vcut", &c->vcut, 0.0f);
        if (!strcmp(vt, "uni")) vmodel_uni(ini, r, c->vcut, r->taup, r->taus);
        else if (!strcmp(vt, "lhm")) { if (vmodel_lhm(ini, base, r, c->vcut, r->taup, r->taus)) return -1; }
        else if (!strcmp(vt, "lgm") || !strcmp(vt, "uni_rmed") || !strcmp(vt, "lhm_rmed") || !strcmp(vt, "lgm_rmed") || !strcmp(vt, "grd") || !strcmp(vt, "grd_rmed")) { /* ora_models.c */
            char m[700] = "";
            int rc;
            if (!strcmp(vt, "lgm")) rc = ora_vmodel_lgm(ini, base, r, c->vcut, r->taup, r->taus, m, sizeof(m));
            else if (!strcmp(vt, "uni_rmed")) rc = ora_vmodel_uni_rmed(c, ini, base, r, c->vcut, r->taup, r->taus, m, sizeof(m));
            else if (!strcmp(vt, "lhm_rmed")) rc = ora_vmodel_lhm_rmed(c, ini, base, r, c->vcut, r->taup, r->taus, m, sizeof(m));
            else if (!strcmp(vt, "grd")) rc = ora_vmodel_grd(c, ini, base, r, c->vcut, r->taup, r->taus, 0, m, sizeof(m));
            else if (!strcmp(vt, "grd_rmed")) rc = ora_vmodel_grd(c, ini, base, r, c->vcut, r->taup, r->taus, 1, m, sizeof(m));
            else rc = ora_vmodel_lgm_rmed(c, ini, base, r, c->vcut, r->taup, r->taus, m, sizeof(m));
            if (rc) { set_err(m); return -1; }
        } else {
            char m[300];
            snprintf(m, sizeof(m), "vmodel_type '%s' is not restated ('user' is a compile-time plug-in)", vt);
            set_err(m);
            return -1;
        }
    }

#define CP5(dst, src) do { r->rho[dst] = r->rho[src]; r->lam[dst] = r->lam[src]; r->mu[dst] = r->mu[src]; \
                           r->taup[dst] = r->taup[src]; r->taus[dst] = r->taus[src]; } while (0)
    /* homogenize absorber region :124-190 (x, then y, then z) */
    for (int i = r->ibeg_m; i <= na; i++)
        for (int j = r->jbeg_m; j <= r->jend_m; j++)
            for (int k = r->kbeg_m; k <= r->kend_m; k++) CP5(ora_idx3(r, k, i, j), ora_idx3(r, k, na + 1, j));
    for (int i = nx - na + 1; i <= r->iend_m; i++)
        for (int j = r->jbeg_m; j <= r->jend_m; j++)
            for (int k = r->kbeg_m; k <= r->kend_m; k++) CP5(ora_idx3(r, k, i, j), ora_idx3(r, k, nx - na, j));
    for (int j = r->jbeg_m; j <= na; j++)
        for (int i = r->ibeg_m; i <= r->iend_m; i++)
            for (int k = r->kbeg_m; k <= r->kend_m; k++) CP5(ora_idx3(r, k, i, j), ora_idx3(r, k, i, na + 1));
    for (int j = ny - na + 1; j <= r->jend_m; j++)
        for (int i = r->ibeg_m; i <= r->iend_m; i++)
            for (int k = r->kbeg_m; k <= r->kend_m; k++) CP5(ora_idx3(r, k, i, j), ora_idx3(r, k, i, ny - na));
    for (int j = r->jbeg_m; j <= r->jend_m; j++)
        for (int i = r->ibeg_m; i <= r->iend_m; i++)
            for (int k = nz - na + 1; k <= r->kend_m; k++) CP5(ora_idx3(r, k, i, j), ora_idx3(r, nz - na, i, j));
#undef CP5

    /* tau-method :193-207 */
    ora_visco_set_relaxtime(nm, c->ts, c->fq_min, c->fq_max);
    float zeta = ora_visco_constq_zeta(nm, c->fq_min, c->fq_max, c->ts);
    if (c->benchmark_mode) zeta = 0.0f;
    c->zeta = zeta;
    for (size_t n = 0; n < nc; n++) {
        r->taup[n] = nm * zeta / r->taup[n];
        r->taus[n] = nm * zeta / r->taus[n];
    }

    /* relaxed_medium :231-271 */
    if (nm > 0) {
        float omega = (float)(2 * ora_pi() * (double)c->fq_ref);
        float complex cc = 0.0f;
        for (int im = 0; im < nm; im++) {
            double complex w = I * (double)omega * (double)c->ts[im];
            double complex q = w / (1.0 - w);
            cc = cc + (float complex)q;
        }
        cc = (crealf(cc) / (float)nm) + (cimagf(cc) / (float)nm) * I;
        for (size_t n = 0; n < nc; n++) {
            float rho_beta2 = r->mu[n];
            float rho_alpha2 = r->lam[n] + 2 * r->mu[n];
            float complex zs = 1.0f - cc * r->taus[n];
            float complex zp = 1.0f - cc * r->taup[n];
            float chi_mu = 1.0f / crealf(1.0f / csqrtf(zs));
            float chi_lam = 1.0f / crealf(1.0f / csqrtf(zp));
            r->mu[n] = rho_beta2 / (chi_mu * chi_mu);
            r->lam[n] = rho_alpha2 / (chi_lam * chi_lam) - 2 * r->mu[n];
        }
    }

    /* surface_detection :339-394 */
    for (size_t n = 0; n < n2; n++) { r->kfs[n] = r->kbeg - 1; r->kob[n] = r->kbeg - 1; }
    for (int j = r->jbeg - 1; j <= r->jend + 2; j++)
        for (int i = r->ibeg - 1; i <= r->iend + 2; i++)
            for (int k = r->kbeg; k <= r->kend - 1; k++) {
                size_t n0 = ora_idx3(r, k, i, j), n1 = ora_idx3(r, k + 1, i, j);
                if (fabsf(r->mu[n0]) < FLT_EPS && fabsf(r->mu[n1]) > FLT_EPS) r->kob[ora_idx2(r, i, j)] = k;
                if (fabsf(r->lam[n0]) < FLT_EPS && fabsf(r->lam[n1]) > FLT_EPS) r->kfs[ora_idx2(r, i, j)] = k;
            }
    for (int j = r->jbeg; j <= r->jend; j++)
        for (int i = r->ibeg; i <= r->iend; i++) {
            int fmin_ = 1 << 30, fmax_ = -(1 << 30), omin = 1 << 30, omax = -(1 << 30);
            for (int jj = j - 2; jj <= j + 3; jj++)
                for (int ii = i - 2; ii <= i + 3; ii++) {
                    int a = r->kfs[ora_idx2(r, ii, jj)], b = r->kob[ora_idx2(r, ii, jj)];
                    if (a < fmin_) fmin_ = a;
                    if (a > fmax_) fmax_ = a;
                    if (b < omin) omin = b;
                    if (b > omax) omax = b;
                }
            size_t n = ora_idx2(r, i, j);
            r->kfs_top[n] = (fmin_ - 2 > r->kbeg) ? fmin_ - 2 : r->kbeg;
            r->kfs_bot[n] = (fmax_ + 2 < r->kend) ? fmax_ + 2 : r->kend;
            r->kob_top[n] = (omin - 2 > r->kbeg) ? omin - 2 : r->kbeg;
            r->kob_bot[n] = (omax + 2 < r->kend) ? omax + 2 : r->kend;
        }

    /* velocity_minmax :396-427 (the allreduce is done by the caller) */
    float vmx = -1.0f, vmn = 1e30f;
    for (int j = r->jbeg; j <= r->jend; j++)
        for (int i = r->ibeg; i <= r->iend; i++)
            for (int k = r->kfs[ora_idx2(r, i, j)] + 1; k <= r->kend; k++) {
                size_t n = ora_idx3(r, k, i, j);
                float vp = sqrtf((r->lam[n] + 2 * r->mu[n]) / r->rho[n]);
                float vs = sqrtf(r->mu[n] / r->rho[n]);
                if (vp > vmx) vmx = vp;
                if (vs < FLT_EPS) continue;
                if (vs < vmn) vmn = vs;
            }
    *vmin1 = vmn;
    *vmax1 = vmx;

    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* m_kernel.f90:30-73 + memory_allocate :376-400                                               */
static void kernel_setup(ora_sim *s) {
    ora_cfg *c = &s->cfg;
    ora_mp dx = (ora_mp)c->dx, dy = (ora_mp)c->dy, dz = (ora_mp)c->dz;
    c->rc40x = (ora_mp)17.0 / (ora_mp)16.0 / dx; c->rc40y = (ora_mp)17.0 / (ora_mp)16.0 / dy; c->rc40z = (ora_mp)17.0 / (ora_mp)16.0 / dz;
    c->rc41x = (ora_mp)1.0 / (ora_mp)48.0 / dx;  c->rc41y = (ora_mp)1.0 / (ora_mp)48.0 / dy;  c->rc41z = (ora_mp)1.0 / (ora_mp)48.0 / dz;
    c->rd40x = -(ora_mp)1.0 / (ora_mp)16.0 / dx; c->rd40y = -(ora_mp)1.0 / (ora_mp)16.0 / dy; c->rd40z = -(ora_mp)1.0 / (ora_mp)16.0 / dz;
    c->rd41x = -(ora_mp)1.0 / (ora_mp)48.0 / dx; c->rd41y = -(ora_mp)1.0 / (ora_mp)48.0 / dy; c->rd41z = -(ora_mp)1.0 / (ora_mp)48.0 / dz;
    int nm = c->nm;
    float dt = c->dt;
    c->d2 = 0.0f; /* Q5: unset when nm == 0 in the reference; taup=taus=0 there so it never matters */
    if (nm > 0) {
        for (int m = 0; m < nm; m++) {
            c->c1[m] = (2 * c->ts[m] - dt) / (2 * c->ts[m] + dt);
            c->c2[m] = (2) / (2 * c->ts[m] + dt) / nm;
        }
        float sum = 0.0f;
        for (int m = 0; m < nm; m++) sum += dt / (2 * c->ts[m] - dt);
        c->d2 = sum / nm;
        for (int m = 0; m < nm; m++) c->d1[m] = 2 * c->ts[m] / (2 * c->ts[m] - dt);
    }
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        size_t nc = r->ncell_m;
        r->Vx = (ora_mp *)xcalloc(nc, sizeof(ora_mp)); r->Vy = (ora_mp *)xcalloc(nc, sizeof(ora_mp));
        r->Vz = (ora_mp *)xcalloc(nc, sizeof(ora_mp));
        r->Sxx = (ora_mp *)xcalloc(nc, sizeof(ora_mp)); r->Syy = (ora_mp *)xcalloc(nc, sizeof(ora_mp));
        r->Szz = (ora_mp *)xcalloc(nc, sizeof(ora_mp)); r->Syz = (ora_mp *)xcalloc(nc, sizeof(ora_mp));
        r->Sxz = (ora_mp *)xcalloc(nc, sizeof(ora_mp)); r->Sxy = (ora_mp *)xcalloc(nc, sizeof(ora_mp));
        r->nzk = r->kend_k - r->kbeg_k + 1; if (r->nzk < 0) r->nzk = 0;
        r->nxk = r->iend_k - r->ibeg_k + 1; if (r->nxk < 0) r->nxk = 0;
        r->nyk = r->jend_k - r->jbeg_k + 1; if (r->nyk < 0) r->nyk = 0;
        if (nm > 0) {
            size_t nr = (size_t)nm * r->nzk * r->nxk * r->nyk;
            r->Rxx = (float *)xcalloc(nr, sizeof(float)); r->Ryy = (float *)xcalloc(nr, sizeof(float));
            r->Rzz = (float *)xcalloc(nr, sizeof(float)); r->Ryz = (float *)xcalloc(nr, sizeof(float));
            r->Rxz = (float *)xcalloc(nr, sizeof(float)); r->Rxy = (float *)xcalloc(nr, sizeof(float));
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* pw_setup  m_source.f90:316-466: plane-wave initial condition over the whole memory box; no source grid afterwards */
static int pw_setup(ora_sim *s, const ora_ini *ini) {
    ora_cfg *c = &s->cfg;
    float pw_ztop, pw_zlen, strike, dip, rake;
    char ps[ORA_STRLEN], tmp[ORA_STRLEN];
    ora_readini_s(ini, "pw_ztop", &pw_ztop, 1e30f);
    if (!(pw_ztop < c->zend)) { set_err("assert: pw_ztop < zend (m_source.f90:332)"); return -1; }
    ora_readini_s(ini, "pw_zlen", &pw_zlen, -1.0f);
    if (!(pw_zlen > 0.0f)) { set_err("assert: pw_zlen > 0 (m_source.f90:335)"); return -1; }
    ora_readini_c(ini, "pw_ps", ps, "");
    const int is_p = (ps[0] == 'p' || ps[0] == 'P'), is_s = (ps[0] == 's' || ps[0] == 'S');
    if (!(is_p || is_s) || ps[1]) { set_err("assert: pw_ps must be p or s (m_source.f90:338)"); return -1; }
    ora_readini_s(ini, "pw_strike", &strike, 0.0f);
    ora_readini_s(ini, "pw_dip", &dip, 0.0f);
    ora_readini_s(ini, "pw_rake", &rake, 0.0f);
    strike = ora_deg2rad(strike); dip = ora_deg2rad(dip); rake = ora_deg2rad(rake);
    ora_readini_c(ini, "stftype", tmp, "kupper");
    strncpy(c->stftype, tmp, sizeof(c->stftype) - 1);
    const float sd = sinf(dip), cd = cosf(dip), sf = sinf(strike), cf = cosf(strike), sl = sinf(rake), cl = cosf(rake);
    const float c2d = cosf(2 * dip), c2f = cosf(2 * strike);
    const float prm[2] = {0.0f, pw_zlen};
    const float dt = c->dt;
    const ora_mp dx = c->dx, dy = c->dy, dz = c->dz;
    const char *st = c->stftype;
    float fcut = 0.0f;
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        for (int j = r->jbeg_m; j <= r->jend_m; j++)
            for (int i = r->ibeg_m; i <= r->iend_m; i++)
                for (int k = r->kbeg_m; k <= r->kend_m; k++) {
                    const size_t n = ora_idx3(r, k, i, j);
                    const float la0 = r->lam[n], mu0 = r->mu[n];
                    const float v = is_p ? sqrtf((la0 + 2 * mu0) / r->rho[n]) : sqrtf(mu0 / r->rho[n]);
                    if (v < FLT_EPS) continue;
                    const float x0 = (float)(c->xbeg + (i - 0.5f) * dx), y0 = (float)(c->ybeg + (j - 0.5f) * dy);
                    const float z0 = (float)(c->zbeg + (k - 0.5f) * dz - pw_ztop);
                    const float x1 = (float)(x0 + dx / 2.0f), y1 = (float)(y0 + dy / 2.0f), z1 = (float)(z0 + dz / 2.0f);
                    const float a = sd * sf, b = sd * cf;
                    const float stf_ii = ora_momentrate(a * x0 - b * y0 + cd * z0, st, prm);
                    const float stf_vx = ora_momentrate(a * x1 - b * y0 + cd * z0 + dt / 2.0f * v, st, prm);
                    const float stf_vy = ora_momentrate(a * x0 - b * y1 + cd * z0 + dt / 2.0f * v, st, prm);
                    const float stf_vz = ora_momentrate(a * x0 - b * y0 + cd * z1 + dt / 2.0f * v, st, prm);
                    const float stf_yz = ora_momentrate(a * x0 - b * y1 + cd * z1, st, prm);
                    const float stf_xz = ora_momentrate(a * x1 - b * y0 + cd * z1, st, prm);
                    const float stf_xy = ora_momentrate(a * x1 - b * y1 + cd * z0, st, prm);
                    if (is_p) {
                        r->Vx[n] = -sd * sf * stf_vx;
                        r->Vy[n] = sd * cf * stf_vy;
                        r->Vz[n] = -cd * stf_vz;
                        r->Sxx[n] = -(la0 + 2 * mu0 * sd * sd * sf * sf) * stf_ii / v;
                        r->Syy[n] = -(la0 + 2 * mu0 * sd * sd * cf * cf) * stf_ii / v;
                        r->Szz[n] = -(la0 + 2 * mu0 * cd * cd) * stf_ii / v;
                        r->Syz[n] = 2 * mu0 * sd * cd * cf * stf_yz / v;
                        r->Sxz[n] = -(2 * mu0 * sd * cd * sf * stf_xz / v);
                        r->Sxy[n] = 2 * mu0 * sd * cd * sf * stf_xy / v;
                    } else {
                        r->Vx[n] = (cl * cf + sl * cd * sf) * stf_vx;
                        r->Vy[n] = (cl * sf - sl * cd * cf) * stf_vy;
                        r->Vz[n] = -sl * sd * stf_vz;
                        r->Sxx[n] = 2 * mu0 * sd * sf * (cl * cf + sl * cd * sf) * stf_ii / v;
                        r->Syy[n] = -(2 * mu0 * sd * cf * (cl * sf - sl * cd * cf) * stf_ii / v);
                        r->Szz[n] = -(2 * mu0 * cd * sl * sd * stf_ii / v);
                        r->Syz[n] = mu0 * (cl * cd * sf - sl * c2d * cf) * stf_yz / v;
                        r->Sxz[n] = mu0 * (cl * cd * cf + sl * c2d * sf) * stf_xz / v;
                        r->Sxy[n] = -(mu0 * (cl * sd * c2f + 2 * sl * sd * cd * sf * cf) * stf_xy / v);
                    }
                }
        /* wavelength condition :445-464 (MPI_MAX over the ranks) */
        const int i = ora_x2i((c->xbeg + c->xend) / 2, c->xbeg, (float)c->dx), j = ora_x2i((c->ybeg + c->yend) / 2, c->ybeg, (float)c->dy);
        const int k = ora_x2i(pw_ztop, c->zbeg, (float)c->dz);
        if (r->ibeg <= i && i <= r->iend && r->jbeg <= j && j <= r->jend) {
            const size_t n = ora_idx3(r, k, i, j);
            const float v = is_p ? sqrtf((r->lam[n] + 2 * r->mu[n]) / r->rho[n]) : sqrtf(r->mu[n] / r->rho[n]);
            if (v / pw_zlen > fcut) fcut = v / pw_zlen;
        }
        r->nsrc = 0;
    }
    c->fcut = fcut;
    c->fmax = fcut * 2.0f;
    c->M0 = 1.0f / c->UC; /* fictitious scalar moment for output :77 */
    c->dt_dxyz = (ora_mp)c->dt / ((ora_mp)c->dx * (ora_mp)c->dy * (ora_mp)c->dz);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* m_source.f90:41-314 (+ grid_moment :468-700, grid_bodyforce :702-774)                        */
static int source_setup(ora_sim *s, const ora_ini *ini, const char *base) {
    ora_cfg *c = &s->cfg;
    ora_readini_l(ini, "pw_mode", &c->pw_mode, 0);
    ora_readini_l(ini, "green_mode", &c->green_mode, 0);
    ora_readini_l(ini, "bf_mode", &c->bf_mode, 0);
    if (c->pw_mode && c->green_mode) { set_err("assert: pw_mode and green_mode are exclusive (m_source.f90:70)"); return -1; }
    if (c->pw_mode && !c->benchmark_mode) return pw_setup(s, ini); /* :73-82 */
    if (c->green_mode && !c->benchmark_mode) { /* :84-91: no regular source grid; M0 / fmax come from green__setup */
        c->pw_mode = 0;
        return 0;
    }
    char tmp[ORA_STRLEN];
    ora_readini_c(ini, "fn_stf", c->fn_stf, "");
    ora_readini_c(ini, "stftype", tmp, "kupper");
    strncpy(c->stftype, tmp, sizeof(c->stftype) - 1);
    if (!strcmp(c->stftype, "scosine")) strcpy(c->stftype, "cosine");
    ora_readini_c(ini, "stf_format", tmp, "xym0ij");
    strncpy(c->stf_format, tmp, sizeof(c->stf_format) - 1);
    ora_readini_c(ini, "sdep_fit", tmp, "asis");
    strncpy(c->sdep_fit, tmp, sizeof(c->sdep_fit) - 1);
    ora_readini_l(ini, "earth_flattening", &c->earth_flattening, 0);

    int cap = 16, ns = 0;
    float *sx = (float *)malloc(sizeof(float) * cap), *sy = (float *)malloc(sizeof(float) * cap),
          *sz = (float *)malloc(sizeof(float) * cap), *p1 = (float *)malloc(sizeof(float) * cap),
          *p2 = (float *)malloc(sizeof(float) * cap), *mo = (float *)malloc(sizeof(float) * cap);
    float *m6 = (float *)malloc(sizeof(float) * 6 * cap); /* mxx myy mzz myz mxz mxy | fx fy fz */
#define GROW()                                                                                         \
    if (ns == cap) {                                                                                   \
        cap *= 2;                                                                                      \
        sx = (float *)realloc(sx, sizeof(float) * cap); sy = (float *)realloc(sy, sizeof(float) * cap); \
        sz = (float *)realloc(sz, sizeof(float) * cap); p1 = (float *)realloc(p1, sizeof(float) * cap); \
        p2 = (float *)realloc(p2, sizeof(float) * cap); mo = (float *)realloc(mo, sizeof(float) * cap); \
        m6 = (float *)realloc(m6, sizeof(float) * 6 * cap);                                            \
    }

    if (c->benchmark_mode) { /* :115-166 */
        strcpy(c->stftype, "kupper");
        c->bf_mode = 0;
        c->pw_mode = 0;
        ns = 1;
        sx[0] = 0.0f; sy[0] = 0.0f; sz[0] = 5.0f; mo[0] = 1e15f;
        m6[0] = m6[1] = m6[2] = 1 / sqrtf(3.0f);
        m6[3] = m6[4] = m6[5] = 0.0f;
        p1[0] = 0.1f; p2[0] = 2.0f;
        c->evlo = c->clon; c->evla = c->clat; c->evdp = sz[0];
        c->mxx0 = m6[0]; c->myy0 = m6[1]; c->mzz0 = m6[2]; c->myz0 = m6[3]; c->mxz0 = m6[4]; c->mxy0 = m6[5];
    } else {
        char path[2 * ORA_STRLEN];
        resolve_path(base, c->fn_stf, path, sizeof(path));
        FILE *fp = fopen(path, "r");
        if (!fp) {
            char m[600];
            snprintf(m, sizeof(m), "source__setup: cannot open %s", path);
            set_err(m);
            return -1;
        }
        char line[1024];
        const char *fmt = c->stf_format;
        while (fgets(line, sizeof(line), fp)) {
            char *p = line;
            while (*p == ' ' || *p == '\t') p++;
            if (*p == '#' || is_blank(p)) continue;
            GROW();
            float v[16];
            int nv = parse_floats_sp(p, v, 16);
            float *M = &m6[6 * ns];
            if (c->bf_mode) { /* :741-758 */
                if (nv < 8) { set_err("source file: bad body-force record"); fclose(fp); return -1; }
                if (fmt[0] == 'x' && fmt[1] == 'y') { sx[ns] = v[0]; sy[ns] = v[1]; }
                else if (fmt[0] == 'l' && fmt[1] == 'l') ora_geomap_g2c(v[0], v[1], c->clon, c->clat, c->phi, &sx[ns], &sy[ns]);
                else { set_err("invalid source type"); fclose(fp); return -1; }
                sz[ns] = v[2]; p1[ns] = v[3]; p2[ns] = v[4];
                M[0] = v[5]; M[1] = v[6]; M[2] = v[7]; M[3] = M[4] = M[5] = 0;
                mo[ns] = 0;
                if (ns == 0) {
                    ora_geomap_c2g(sx[0], sy[0], c->clon, c->clat, c->phi, &c->evlo, &c->evla);
                    c->evdp = sz[0]; c->fx0 = M[0]; c->fy0 = M[1]; c->fz0 = M[2]; c->otim = p1[0];
                }
                ns++;
                continue;
            }
            int is_ll = (fmt[0] == 'l' && fmt[1] == 'l');
            int is_xy = (fmt[0] == 'x' && fmt[1] == 'y');
            const char *kind = fmt + 2; /* m0ij m0dc mwij mwdc dsdc */
            if (!strcmp(fmt, "psmeca")) { /* :641-666  lon lat z mzz mxx myy mxz myz mxy iex (dyn-cm, GMT psmeca order) */
                if (nv < 10) { set_err("source file: bad psmeca record"); fclose(fp); return -1; }
                sz[ns] = v[2];
                M[2] = v[3]; M[0] = v[4]; M[1] = v[5]; M[4] = v[6]; M[3] = -v[7]; M[5] = -v[8];
                int iex = (int)v[9];
                ora_geomap_g2c(v[0], v[1], c->clon, c->clat, c->phi, &sx[ns], &sy[ns]);
                float M0tmp = sqrtf(M[0] * M[0] + M[1] * M[1] + M[2] * M[2] + 2 * (M[4] * M[4] + M[3] * M[3] + M[5] * M[5])) / sqrtf(2.0f);
                mo[ns] = M0tmp * ora_powi_sp(10.0f, iex);
                p1[ns] = 0.0f;
                p2[ns] = (float)((double)(2 * 1.05f * 1e-8f) * pow((double)mo[ns], 1.0 / 3.0));
                mo[ns] = mo[ns] * 1e-7f;
                M[0] = M[0] / M0tmp; M[1] = M[1] / M0tmp; M[2] = M[2] / M0tmp; M[4] = M[4] / M0tmp; M[3] = M[3] / M0tmp; M[5] = M[5] / M0tmp;
            } else if ((is_ll || is_xy) && !strcmp(kind, "dsdc")) { /* :585-639  x y z tbeg trise D S strike dip rake */
                if (nv < 10) { set_err("source file: bad dsdc record"); fclose(fp); return -1; }
                if (is_xy) { sx[ns] = v[0]; sy[ns] = v[1]; }
                else ora_geomap_g2c(v[0], v[1], c->clon, c->clat, c->phi, &sx[ns], &sy[ns]);
                sz[ns] = v[2]; p1[ns] = v[3]; p2[ns] = v[4];
                ora_sdr2moment(v[7] - c->phi, v[8], v[9], &M[0], &M[1], &M[2], &M[3], &M[4], &M[5]);
                int is0 = ora_x2i(sx[ns], c->xbeg, (float)c->dx), js0 = ora_x2i(sy[ns], c->ybeg, (float)c->dy), ks0;
                if (c->earth_flattening) ks0 = ora_x2i((float)(-ora_r_earth() * log((ora_r_earth() - (double)sz[ns]) / ora_r_earth())), c->zbeg, (float)c->dz);
                else ks0 = ora_x2i(sz[ns], c->zbeg, (float)c->dz);
                float best = 0.0f; /* mpi_allreduce(MAX) over the ranks :691-696 */
                for (int q = 0; q < s->nranks; q++) {
                    const ora_rank *r = &s->r[q];
                    float m = 0.0f;
                    if (r->ibeg - 2 <= is0 && is0 <= r->iend + 3 && r->jbeg - 2 <= js0 && js0 <= r->jend + 3 && r->kbeg - 2 <= ks0 && ks0 <= r->kend + 3)
                        m = (1e9f * r->mu[ora_idx3(r, ks0, is0, js0)]) * v[5] * v[6];
                    if (q == 0 || m > best) best = m;
                }
                mo[ns] = best;
            } else {
            if (!(is_ll || is_xy) || !(!strcmp(kind, "m0ij") || !strcmp(kind, "m0dc") || !strcmp(kind, "mwij") || !strcmp(kind, "mwdc"))) {
                set_err("invalid source type"); fclose(fp); return -1; /* :668-670 */
            }
            int need = (kind[2] == 'i') ? 12 : 9;
            if (nv < need) { set_err("source file: bad moment record"); fclose(fp); return -1; }
            if (is_xy) { sx[ns] = v[0]; sy[ns] = v[1]; }
            else ora_geomap_g2c(v[0], v[1], c->clon, c->clat, c->phi, &sx[ns], &sy[ns]);
            sz[ns] = v[2]; p1[ns] = v[3]; p2[ns] = v[4];
            mo[ns] = (kind[1] == '0') ? v[5] : ora_seismic_moment(v[5]);
            if (kind[2] == 'i') { for (int q = 0; q < 6; q++) M[q] = v[6 + q]; }
            else ora_sdr2moment(v[6] - c->phi, v[7], v[8], &M[0], &M[1], &M[2], &M[3], &M[4], &M[5]);
            }
            if (ns == 0) { /* :675-687 */
                ora_geomap_c2g(sx[0], sy[0], c->clon, c->clat, c->phi, &c->evlo, &c->evla);
                c->sx0 = sx[0]; c->sy0 = sy[0]; c->evdp = sz[0];
                c->mxx0 = M[0]; c->myy0 = M[1]; c->mzz0 = M[2]; c->myz0 = M[3]; c->mxz0 = M[4]; c->mxy0 = M[5];
                c->otim = p1[0];
            }
            ns++;
        }
        fclose(fp);
    }

    if (c->earth_flattening) { /* :176-180 */
        double RE = ora_r_earth();
        for (int k = 0; k < ns; k++) sz[k] = -(float)(RE * log((RE - (double)sz[k]) / RE));
    }
    c->fcut = 0.0f; /* :184-188 */
    for (int i = 0; i < ns; i++) { float f = 1 / p2[i]; if (f > c->fcut) c->fcut = f; }
    c->fmax = 2 * c->fcut;
    if (c->bf_mode) { /* :191-198 */
        float sum = 0.0f;
        for (int i = 0; i < ns; i++) sum += m6[6 * i] * m6[6 * i] + m6[6 * i + 1] * m6[6 * i + 1] + m6[6 * i + 2] * m6[6 * i + 2];
        c->M0 = sqrtf(sum);
        c->UC = c->UC * 1000;
    } else {
        float sum = 0.0f;
        for (int i = 0; i < ns; i++) sum += mo[i];
        c->M0 = sum;
    }
    c->dt_dxyz = (ora_mp)c->dt / ((ora_mp)c->dx * (ora_mp)c->dy * (ora_mp)c->dz); /* :306 */

    int *isg = (int *)malloc(sizeof(int) * (ns + 1)), *jsg = (int *)malloc(sizeof(int) * (ns + 1)),
        *ksg = (int *)malloc(sizeof(int) * (ns + 1));
    for (int i = 0; i < ns; i++) { /* :204-206 */
        isg[i] = ora_x2i(sx[i], c->xbeg, (float)c->dx);
        jsg[i] = ora_x2i(sy[i], c->ybeg, (float)c->dy);
        ksg[i] = ora_x2i(sz[i], c->zbeg, (float)c->dz);
    }
    int rc = 0;
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        int n = 0;
        for (int i = 0; i < ns; i++)
            if (r->ibeg - 2 <= isg[i] && isg[i] <= r->iend + 3 && r->jbeg - 2 <= jsg[i] && jsg[i] <= r->jend + 3 &&
                r->kbeg - 2 <= ksg[i] && ksg[i] <= r->kend + 3) n++;
        r->nsrc = n;
        r->isrc = (int *)xcalloc(n, sizeof(int)); r->jsrc = (int *)xcalloc(n, sizeof(int)); r->ksrc = (int *)xcalloc(n, sizeof(int));
        r->sx = (float *)xcalloc(n, sizeof(float)); r->sy = (float *)xcalloc(n, sizeof(float)); r->sz = (float *)xcalloc(n, sizeof(float));
        r->srcprm = (float *)xcalloc(2 * (size_t)n, sizeof(float));
        r->mo = (ora_mp *)xcalloc(n, sizeof(ora_mp));
        r->mxx = (ora_mp *)xcalloc(n, sizeof(ora_mp)); r->myy = (ora_mp *)xcalloc(n, sizeof(ora_mp)); r->mzz = (ora_mp *)xcalloc(n, sizeof(ora_mp));
        r->myz = (ora_mp *)xcalloc(n, sizeof(ora_mp)); r->mxz = (ora_mp *)xcalloc(n, sizeof(ora_mp)); r->mxy = (ora_mp *)xcalloc(n, sizeof(ora_mp));
        r->fx = (ora_mp *)xcalloc(n, sizeof(ora_mp)); r->fy = (ora_mp *)xcalloc(n, sizeof(ora_mp)); r->fz = (ora_mp *)xcalloc(n, sizeof(ora_mp));
        int nn = 0;
        for (int i = 0; i < ns; i++) {
            if (!(r->ibeg - 2 <= isg[i] && isg[i] <= r->iend + 3 && r->jbeg - 2 <= jsg[i] && jsg[i] <= r->jend + 3 &&
                  r->kbeg - 2 <= ksg[i] && ksg[i] <= r->kend + 3)) continue;
            r->isrc[nn] = isg[i]; r->jsrc[nn] = jsg[i]; r->ksrc[nn] = ksg[i];
            r->sx[nn] = sx[i]; r->sy[nn] = sy[i]; r->sz[nn] = sz[i];
            const float *M = &m6[6 * i];
            if (c->bf_mode) { r->fx[nn] = M[0]; r->fy[nn] = M[1]; r->fz[nn] = M[2]; }
            else {
                r->mo[nn] = mo[i];
                r->mxx[nn] = M[0]; r->myy[nn] = M[1]; r->mzz[nn] = M[2]; r->myz[nn] = M[3]; r->mxz[nn] = M[4]; r->mxy[nn] = M[5];
            }
            r->srcprm[2 * nn] = p1[i]; r->srcprm[2 * nn + 1] = p2[i];
            /* sdep_fit = bd0..bd9 :263-271 */
            if (c->sdep_fit[0] == 'b' && c->sdep_fit[1] == 'd' && isdigit((unsigned char)c->sdep_fit[2])) {
                int b = c->sdep_fit[2] - '0';
                size_t n2 = (size_t)r->nxm * r->nym;
                r->sz[nn] = r->bddep[(size_t)b * n2 + ora_idx2(r, r->isrc[nn], r->jsrc[nn])];
                r->ksrc[nn] = ora_x2i(r->sz[nn], c->zbeg, (float)c->dz);
            }
            /* :277-281 */
            if (!(c->xbeg <= r->sx[nn] && r->sx[nn] <= c->xend && c->ybeg <= r->sy[nn] && r->sy[nn] <= c->yend &&
                  c->zbeg <= r->sz[nn] && r->sz[nn] <= c->zend)) {
                set_err("source__setup: assert failed, source outside of the model space");
                rc = -1;
            }
            nn++;
        }
        /* :297-303 */
        for (int i = 0; i < n; i++) {
            if (c->bf_mode) { r->fx[i] = r->fx[i] / c->M0; r->fy[i] = r->fy[i] / c->M0; r->fz[i] = r->fz[i] / c->M0; }
            else r->mo[i] = r->mo[i] / c->M0;
        }
    }
    free(isg); free(jsg); free(ksg);
    free(sx); free(sy); free(sz); free(p1); free(p2); free(mo); free(m6);
    return rc;
}

/* ------------------------------------------------------------------------------------------ */
/* m_absorb_p.f90:60-124  /  m_absorb_c.f90:27-111                                             */
static int absorb_setup(ora_sim *s) {
    ora_cfg *c = &s->cfg;
    float dx = (float)c->dx, dy = (float)c->dy, dz = (float)c->dz;
    if (!strcmp(c->abc_type, "pml")) {
        c->r20x = (ora_mp)(1.0f / c->dx); /* "r20x = 1.0 / dx": SP literal / real(MP) */
        c->r20y = (ora_mp)(1.0f / c->dy);
        c->r20z = (ora_mp)(1.0f / c->dz);
        float hx = c->na * dx, hy = c->na * dy, hz = c->na * dz;
        for (int q = 0; q < s->nranks; q++) {
            ora_rank *r = &s->r[q];
            r->gxc = (float *)xcalloc(4 * (size_t)r->nxp, sizeof(float)); r->gxe = (float *)xcalloc(4 * (size_t)r->nxp, sizeof(float));
            r->gyc = (float *)xcalloc(4 * (size_t)r->nyp, sizeof(float)); r->gye = (float *)xcalloc(4 * (size_t)r->nyp, sizeof(float));
            r->gzc = (float *)xcalloc(4 * (size_t)c->nz, sizeof(float)); r->gze = (float *)xcalloc(4 * (size_t)c->nz, sizeof(float));
            for (int i = r->ibeg; i <= r->iend; i++) {
                float x = r->xc[i - r->ibeg_m];
                ora_damping_profile(x, hx, c->xbeg, c->xend, c->na, c->fcut, c->dt, &r->gxc[4 * (i - r->ibeg)]);
                ora_damping_profile(x + dx / 2.0f, hx, c->xbeg, c->xend, c->na, c->fcut, c->dt, &r->gxe[4 * (i - r->ibeg)]);
            }
            for (int j = r->jbeg; j <= r->jend; j++) {
                float y = r->yc[j - r->jbeg_m];
                ora_damping_profile(y, hy, c->ybeg, c->yend, c->na, c->fcut, c->dt, &r->gyc[4 * (j - r->jbeg)]);
                ora_damping_profile(y + dy / 2.0f, hy, c->ybeg, c->yend, c->na, c->fcut, c->dt, &r->gye[4 * (j - r->jbeg)]);
            }
            for (int k = r->kbeg; k <= r->kend; k++) {
                float z = r->zc[k - r->kbeg_m];
                ora_damping_profile(z, hz, c->zbeg, c->zend, c->na, c->fcut, c->dt, &r->gzc[4 * (k - r->kbeg)]);
                ora_damping_profile(z + dz / 2.0f, hz, c->zbeg, c->zend, c->na, c->fcut, c->dt, &r->gze[4 * (k - r->kbeg)]);
            }
            /* the reference allocates 18 full (kbeg_min:kend, ibeg:iend, jbeg:jend) arrays (:99-116);
             * only cells with k >= kbeg_a(i,j) are ever touched, so the oracle stores that shell only */
            r->aoff = (int64_t *)xcalloc((size_t)r->nxp * r->nyp, sizeof(int64_t));
            int64_t off = 0;
            for (int j = r->jbeg; j <= r->jend; j++)
                for (int i = r->ibeg; i <= r->iend; i++) {
                    r->aoff[(size_t)(i - r->ibeg) + (size_t)r->nxp * (j - r->jbeg)] = off;
                    off += r->kend - r->kbeg_a[ora_idx2(r, i, j)] + 1;
                }
            r->naux = off;
            float **aux[18] = {&r->axVx, &r->ayVx, &r->azVx, &r->axVy, &r->ayVy, &r->azVy, &r->axVz, &r->ayVz, &r->azVz,
                               &r->axSxx, &r->aySxy, &r->azSxz, &r->axSxy, &r->aySyy, &r->azSyz, &r->axSxz, &r->aySyz, &r->azSzz};
            for (int a = 0; a < 18; a++) *aux[a] = (float *)xcalloc((size_t)off, sizeof(float));
        }
    } else if (!strcmp(c->abc_type, "cerjan")) {
        const float alpha = 0.09f;
        float Lx = c->na * dx, Ly = c->na * dy, Lz = c->na * dz;
        int na = c->na, nx = c->nx, ny = c->ny, nz = c->nz;
        for (int q = 0; q < s->nranks; q++) {
            ora_rank *r = &s->r[q];
            r->gx_c = (float *)xcalloc((size_t)r->nxm, sizeof(float)); r->gx_b = (float *)xcalloc((size_t)r->nxm, sizeof(float));
            r->gy_c = (float *)xcalloc((size_t)r->nym, sizeof(float)); r->gy_b = (float *)xcalloc((size_t)r->nym, sizeof(float));
            r->gz_c = (float *)xcalloc((size_t)r->nzm, sizeof(float)); r->gz_b = (float *)xcalloc((size_t)r->nzm, sizeof(float));
            for (int i = 0; i < r->nxm; i++) r->gx_c[i] = r->gx_b[i] = 1.0f;
            for (int j = 0; j < r->nym; j++) r->gy_c[j] = r->gy_b[j] = 1.0f;
            for (int k = 0; k < r->nzm; k++) r->gz_c[k] = r->gz_b[k] = 1.0f;
#define SQ(x) ((x) * (x))
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
            for (int j = r->jbeg; j <= r->jend; j++) {
                float *gc = &r->gy_c[j - r->jbeg_m], *gb = &r->gy_b[j - r->jbeg_m];
                if (j <= na) {
                    *gc = expf(-(alpha * SQ(1.0f - (ora_i2x(j, 0.0f, dy)) / Ly)));
                    *gb = expf(-(alpha * SQ(1.0f - ((ora_i2x(j, 0.0f, dy) + dy / 2)) / Ly)));
                } else if (j >= ny - na + 1) {
                    *gc = expf(-(alpha * SQ(1.0f - (ora_i2x(j, ny * dy, -dy) + dy / 2) / Ly)));
                    *gb = expf(-(alpha * SQ(1.0f - ((ora_i2x(j, ny * dy, -dy))) / Ly)));
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
#undef SQ
        }
    } else {
        set_err("absorb__setup: unknown abc_type (assert(.false.), m_absorb.f90:37)");
        return -1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* m_wav.f90:54-271                                                                             */
static int wav_setup(ora_sim *s, const ora_ini *ini, const char *base) {
    ora_cfg *c = &s->cfg;
    char tmp[ORA_STRLEN];
    ora_readini_i(ini, "ntdec_w", &c->ntdec_w, 10);
    ora_readini_l(ini, "sw_wav_v", &c->sw_wav_v, 0);
    ora_readini_l(ini, "sw_wav_u", &c->sw_wav_u, 0);
    ora_readini_l(ini, "sw_wav_stress", &c->sw_wav_stress, 0);
    ora_readini_l(ini, "sw_wav_strain", &c->sw_wav_strain, 0);
    ora_readini_c(ini, "wav_format", tmp, "sac");
    strncpy(c->wav_format, tmp, sizeof(c->wav_format) - 1);
    ora_readini_c(ini, "st_format", tmp, "xy");
    strncpy(c->st_format, tmp, sizeof(c->st_format) - 1);
    ora_readini_c(ini, "fn_stloc", c->fn_stloc, "");
    {   /* m_wav.f90:76-86: Green's-function mode keeps the station table (wav__stquery) but records nothing */
        int gm;
        ora_readini_l(ini, "green_mode", &gm, 0);
        if (gm) c->sw_wav_v = c->sw_wav_u = c->sw_wav_stress = c->sw_wav_strain = 0;
        else if (!(c->sw_wav_v || c->sw_wav_u || c->sw_wav_stress || c->sw_wav_strain)) return 0;
    }
    c->ntw = (int)floorf((float)(c->nt - 1) / (float)c->ntdec_w + 1.0f); /* :89 */

    char path[2 * ORA_STRLEN];
    resolve_path(base, c->fn_stloc, path, sizeof(path));
    FILE *fp = fopen(path, "r");
    if (!fp) return 0; /* 'no station location file found' :157-161 */
    char line[1024];
    float dx = (float)c->dx, dy = (float)c->dy, dz = (float)c->dz;
    while (fgets(line, sizeof(line), fp)) {
        char *p = line;
        while (*p == ' ' || *p == '\t') p++;
        if (*p == '#' || is_blank(p)) continue;
        float a, b, zst;
        char stnm[64] = "", zsw[64] = "";
        if (sscanf(p, "%f %f %f %63s %63s", &a, &b, &zst, stnm, zsw) < 5) continue;
        stnm[8] = 0; /* character(8) */
        zsw[3] = 0;  /* character(3) */
        float xst, yst, stlo, stla;
        if (!strcmp(c->st_format, "xy")) {
            xst = a; yst = b;
            ora_geomap_c2g(xst, yst, c->clon, c->clat, c->phi, &stlo, &stla);
        } else if (!strcmp(c->st_format, "ll")) {
            stlo = a; stla = b;
            ora_geomap_g2c(stlo, stla, c->clon, c->clat, c->phi, &xst, &yst);
        } else { set_err("unknown st_format"); fclose(fp); return -1; }
        int ist = ora_x2i(xst, c->xbeg, dx), jst = ora_x2i(yst, c->ybeg, dy), kst = ora_x2i(zst, c->zbeg, dz);
        if (!(ora_i2x(1, c->xbeg, dx) < xst && xst < ora_i2x(c->nx, c->xbeg, dx) && ora_i2x(1, c->ybeg, dy) < yst &&
              yst < ora_i2x(c->ny, c->ybeg, dy) && 1 < kst && kst < c->nz)) continue; /* :197-199 */
        for (int q = 0; q < s->nranks; q++) {
            ora_rank *r = &s->r[q];
            if (!(r->ibeg <= ist && ist <= r->iend && r->jbeg <= jst && jst <= r->jend)) continue;
            int k = kst;
            size_t n2 = (size_t)r->nxm * r->nym;
            if (!strcmp(zsw, "dep")) k = ora_x2i(zst, c->zbeg, dz);
            else if (!strcmp(zsw, "fsb")) k = r->kfs[ora_idx2(r, ist, jst)] + 1;
            else if (!strcmp(zsw, "obb")) k = r->kob[ora_idx2(r, ist, jst)] + 1;
            else if (!strcmp(zsw, "oba")) k = r->kob[ora_idx2(r, ist, jst)] - 1;
            else if (zsw[0] == 'b' && zsw[1] == 'd' && isdigit((unsigned char)zsw[2]))
                k = ora_x2i(r->bddep[(size_t)(zsw[2] - '0') * n2 + ora_idx2(r, ist, jst)], c->zbeg, dz);
            else k = ora_x2i(zst, c->zbeg, dz);
            if (k > r->kend) k = r->kend - 1;
            if (k < r->kbeg) k = r->kbeg + 1;
            int n = r->nst++;
            r->ist = (int *)realloc(r->ist, sizeof(int) * r->nst); r->jst = (int *)realloc(r->jst, sizeof(int) * r->nst);
            r->kst = (int *)realloc(r->kst, sizeof(int) * r->nst);
            r->xst = (float *)realloc(r->xst, sizeof(float) * r->nst); r->yst = (float *)realloc(r->yst, sizeof(float) * r->nst);
            r->zst = (float *)realloc(r->zst, sizeof(float) * r->nst);
            r->stlo = (float *)realloc(r->stlo, sizeof(float) * r->nst); r->stla = (float *)realloc(r->stla, sizeof(float) * r->nst);
            r->stnm = (char(*)[9])realloc(r->stnm, 9 * (size_t)r->nst);
            r->ist[n] = ist; r->jst[n] = jst; r->kst[n] = k;
            r->xst[n] = xst; r->yst[n] = yst; r->zst[n] = zst; r->stlo[n] = stlo; r->stla[n] = stla;
            memset(r->stnm[n], 0, 9);
            strncpy(r->stnm[n], stnm, 8);
        }
    }
    fclose(fp);
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        if (r->nst <= 0) continue;
        size_t n3 = (size_t)c->ntw * 3 * r->nst, n6 = (size_t)c->ntw * 6 * r->nst;
        if (c->sw_wav_v) r->wav_vel = (float *)xcalloc(n3, sizeof(float));
        if (c->sw_wav_u) {
            r->wav_disp = (float *)xcalloc(n3, sizeof(float));
            r->ux = (float *)xcalloc((size_t)r->nst, sizeof(float)); r->uy = (float *)xcalloc((size_t)r->nst, sizeof(float));
            r->uz = (float *)xcalloc((size_t)r->nst, sizeof(float));
        }
        if (c->sw_wav_stress) r->wav_stress = (float *)xcalloc(n6, sizeof(float));
        if (c->sw_wav_strain) {
            r->wav_strain = (float *)xcalloc(n6, sizeof(float));
            r->exx = (float *)xcalloc((size_t)r->nst, sizeof(float)); r->eyy = (float *)xcalloc((size_t)r->nst, sizeof(float));
            r->ezz = (float *)xcalloc((size_t)r->nst, sizeof(float)); r->eyz = (float *)xcalloc((size_t)r->nst, sizeof(float));
            r->exz = (float *)xcalloc((size_t)r->nst, sizeof(float)); r->exy = (float *)xcalloc((size_t)r->nst, sizeof(float));
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
static ora_sim *create_from_ini(ora_ini *ini, const char *base_dir, int nm, int npx, int npy, int nt_override) {
    if (nm < 0 || nm > ORA_MAXNM) { set_err("nm out of range"); return NULL; }
    ora_sim *s = (ora_sim *)xcalloc(1, sizeof(ora_sim));
    ora_cfg *c = &s->cfg;
    int strict = 0;
    ora_readini_l(ini, "strict_mode", &strict, 0); /* main.f90:60-61 */
    ini->strict_mode = strict;
    c->nm = nm;
    global_setup(c, ini);
    if (npx > 0) c->nproc_x = npx;
    if (npy > 0) c->nproc_y = npy;
    if (nt_override > 0) { c->nt = nt_override; c->tend = c->tbeg + c->nt * c->dt; }
    c->nproc = c->nproc_x * c->nproc_y;
    c->exedate = (int)time(NULL);
    {
        time_t now = time(NULL);
        struct tm lt;
        localtime_r(&now, &lt);
        c->tz_minutes = (int)(lt.tm_gmtoff / 60);
    }
    s->nranks = c->nproc;
    s->r = (ora_rank *)xcalloc((size_t)s->nranks, sizeof(ora_rank));
    /* m_global.f90:624-642 */
    int tw = c->nproc_x + 2;
    s->itbl = (int *)xcalloc((size_t)tw * (c->nproc_y + 2), sizeof(int));
    for (int i = 0; i < tw * (c->nproc_y + 2); i++) s->itbl[i] = -1;
    for (int i = 0; i < c->nproc; i++) s->itbl[(i % c->nproc_x + 1) + tw * (i / c->nproc_x + 1)] = i;

    float vmin = 1e30f, vmax = -1.0f;
    for (int q = 0; q < s->nranks; q++) {
        rank_geometry(c, &s->r[q], q);
        float a, b;
        if (medium_setup(s, &s->r[q], ini, base_dir, &a, &b)) { ora_destroy(s); return NULL; }
        if (a < vmin) vmin = a;
        if (b > vmax) vmax = b;
    }
    c->vmin = vmin; /* mpi_allreduce m_medium.f90:424-425 */
    c->vmax = vmax;
    {   /* m_medium.f90:218-222: after velocity_minmax, with the global vmax */
        int stab;
        ora_readini_l(ini, "stabilize_pml", &stab, 0);
        if (stab)
            for (int q = 0; q < s->nranks; q++) ora_stabilize_absorber(c, &s->r[q]);
    }
    kernel_setup(s);
    if (source_setup(s, ini, base_dir)) { ora_destroy(s); return NULL; }
    if (absorb_setup(s)) { ora_destroy(s); return NULL; }
    ora_snap_setup(s, ini); /* main.f90:76 */
    if (wav_setup(s, ini, base_dir)) { ora_destroy(s); return NULL; }
    {   /* main.f90:78 */
        char m[600] = "";
        if (ora_green_setup(s, ini, base_dir, m, sizeof(m))) { set_err(m); ora_destroy(s); return NULL; }
    }
    ora_readini_i(ini, "ntdec_r", &c->ntdec_r, 10); /* m_report.f90:47 */
    return s;
}

ora_sim *ora_create(const char *inf_path, const char *base_dir, int nm, int npx, int npy, int nt_override) {
    ora_ini *ini = ora_ini_open(inf_path);
    if (!ini) {
        char m[600];
        snprintf(m, sizeof(m), "cannot open parameter file %s", inf_path);
        set_err(m);
        return NULL;
    }
    ora_sim *s = create_from_ini(ini, base_dir, nm, npx, npy, nt_override);
    ora_ini_close(ini);
    return s;
}

ora_sim *ora_create_from_text(const char *inf_text, const char *base_dir, int nm, int npx, int npy, int nt_override) {
    ora_ini *ini = ora_ini_from_text(inf_text);
    ora_sim *s = create_from_ini(ini, base_dir, nm, npx, npy, nt_override);
    ora_ini_close(ini);
    return s;
}

void ora_destroy(ora_sim *s) {
    if (!s) return;
    ora_snap_free(s);
    ora_green_free(s);
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        void *ptrs[] = {r->Vx, r->Vy, r->Vz, r->Sxx, r->Syy, r->Szz, r->Syz, r->Sxz, r->Sxy, r->Rxx, r->Ryy, r->Rzz, r->Ryz,
                        r->Rxz, r->Rxy, r->rho, r->lam, r->mu, r->taup, r->taus, r->kfs, r->kob, r->kfs_top, r->kfs_bot,
                        r->kob_top, r->kob_bot, r->kbeg_a, r->bddep, r->xc, r->yc, r->zc, r->gxc, r->gxe, r->gyc, r->gye,
                        r->gzc, r->gze, r->aoff, r->axVx, r->ayVx, r->azVx, r->axVy, r->ayVy, r->azVy, r->axVz, r->ayVz,
                        r->azVz, r->axSxx, r->aySxy, r->azSxz, r->axSxy, r->aySyy, r->azSyz, r->axSxz, r->aySyz, r->azSzz,
                        r->gx_c, r->gx_b, r->gy_c, r->gy_b, r->gz_c, r->gz_b, r->isrc, r->jsrc, r->ksrc, r->sx, r->sy, r->sz,
                        r->srcprm, r->mo, r->mxx, r->myy, r->mzz, r->myz, r->mxz, r->mxy, r->fx, r->fy, r->fz, r->ist, r->jst,
                        r->kst, r->xst, r->yst, r->zst, r->stlo, r->stla, r->stnm, r->wav_vel, r->wav_disp, r->wav_stress, r->wav_strain, r->ux, r->uy, r->uz, r->exx, r->eyy, r->ezz, r->eyz, r->exz, r->exy, r->sbuf_ip, r->sbuf_im,
                        r->sbuf_jp, r->sbuf_jm, r->rbuf_ip, r->rbuf_im, r->rbuf_jp, r->rbuf_jm};
        for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(ptrs[i]);
    }
    free(s->r);
    free(s->itbl);
    free(s);
}
