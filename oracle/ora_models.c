/*
 * oracle/ora_models.c -- TEST INFRASTRUCTURE (see ora.h): the heavier velocity-model builders of swpc_3d
 * (SURVEY 8f-4): the linear-gradient model, the random-media variants of the uni / lhm / lgm models, the
 * random-media volume reader and the PML stabiliser.
 *
 *   vmodel_lgm        src/swpc_3d/m_vmodel_lgm.f90:20-175
 *   vmodel_uni_rmed   src/swpc_3d/m_vmodel_uni_rmed.f90:22-170
 *   vmodel_lhm_rmed   src/swpc_3d/m_vmodel_lhm_rmed.f90:22-244
 *   vmodel_lgm_rmed   src/swpc_3d/m_vmodel_lgm_rmed.f90:22-258
 *   rdrmed__3d        src/shared/m_rdrmed.f90:73-134    (netCDF classic volume written by tools/gen_rmed3d.f90:91-135)
 *   vcheck            src/shared/m_fdtool.f90:15-48
 *   independent_list  src/shared/m_fdtool.f90:815-855
 *   stabilize_absorber src/swpc_3d/m_medium.f90:273-337
 *
 * The reference reads the random-media volume through the netCDF-Fortran library, which this image does not
 * have; gen_rmed3d creates the file with NF90_CLOBBER, i.e. the netCDF *classic* format, whose published layout
 * (CDF-1 / CDF-2 header, big-endian, non-record variables stored contiguously in definition order) is restated in
 * nc_open_classic() below.
 */
#include "ora.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

double ora_r_earth(void);
#define FLT_EPS 1.1920929e-07f /* epsilon(1.0) */

/* ------------------------------------------------------------------------------------------ */
/* netCDF classic reader (just enough for rdrmed__3d)                                          */
typedef struct {
    int ndims;
    long long dimlen[16];
    int nvars;
    struct { char name[128]; int ndims; int dimid[8]; int type; long long begin; long long vsize; } var[32];
    unsigned char *buf;
    size_t len;
} nc_file;

static unsigned long long be_u(const unsigned char *p, int n) {
    unsigned long long v = 0;
    for (int i = 0; i < n; i++) v = (v << 8) | p[i];
    return v;
}

static void copy_name(char *dst, const unsigned char *src, unsigned n) {
    if (!dst) return;
    size_t c = n < 127 ? n : 127;
    memcpy(dst, src, c);
    dst[c] = 0;
}

static void nc_close(nc_file *f) {
    if (!f) return;
    free(f->buf);
    free(f);
}

static nc_file *nc_open_classic(const char *path, char *err, size_t cap) {
    FILE *fp = fopen(path, "rb");
    if (!fp) { snprintf(err, cap, "cannot open %s", path); return NULL; }
    fseek(fp, 0, SEEK_END);
    long long len = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    nc_file *f = (nc_file *)calloc(1, sizeof(nc_file));
    f->buf = (unsigned char *)malloc((size_t)len);
    f->len = (size_t)len;
    if (fread(f->buf, 1, (size_t)len, fp) != (size_t)len) { fclose(fp); nc_close(f); snprintf(err, cap, "short read of %s", path); return NULL; }
    fclose(fp);
    const unsigned char *b = f->buf;
    if (len < 32 || b[0] != 'C' || b[1] != 'D' || b[2] != 'F' || (b[3] != 1 && b[3] != 2)) {
        nc_close(f);
        snprintf(err, cap, "%s is not a netCDF classic (CDF-1/2) file", path);
        return NULL;
    }
    const int offw = b[3] == 2 ? 8 : 4;
    size_t p = 8; /* magic + numrecs */
#define U4() (p += 4, (unsigned)be_u(b + p - 4, 4))
#define SKIPNAME(dst) do { unsigned nl_ = U4(); copy_name(dst, b + p, nl_); p += (nl_ + 3u) & ~3u; } while (0)
    /* dim_list */
    unsigned tag = U4(), n = U4();
    if (tag == 0x0A) {
        f->ndims = (int)n;
        for (unsigned i = 0; i < n && i < 16; i++) { SKIPNAME((char *)NULL); f->dimlen[i] = U4(); }
    }
    static const int tsz[7] = {0, 1, 1, 2, 4, 4, 8};
    /* gatt_list */
    tag = U4(); n = U4();
    if (tag == 0x0C)
        for (unsigned i = 0; i < n; i++) {
            SKIPNAME((char *)NULL);
            unsigned ty = U4(), ne = U4();
            p += ((size_t)ne * tsz[ty < 7 ? ty : 0] + 3u) & ~(size_t)3u;
        }
    /* var_list */
    tag = U4(); n = U4();
    if (tag == 0x0B) {
        f->nvars = (int)(n < 32 ? n : 32);
        for (unsigned i = 0; i < n; i++) {
            char nm[128];
            SKIPNAME(nm);
            unsigned nd = U4();
            int dimid[8] = {0};
            for (unsigned d = 0; d < nd; d++) { unsigned id = U4(); if (d < 8) dimid[d] = (int)id; }
            unsigned vt = U4(), vn = U4();
            if (vt == 0x0C)
                for (unsigned a = 0; a < vn; a++) {
                    SKIPNAME((char *)NULL);
                    unsigned ty = U4(), ne = U4();
                    p += ((size_t)ne * tsz[ty < 7 ? ty : 0] + 3u) & ~(size_t)3u;
                }
            unsigned type = U4(), vsize = U4();
            long long begin = (long long)be_u(b + p, offw);
            p += (size_t)offw;
            if (i < 32) {
                strncpy(f->var[i].name, nm, 127);
                f->var[i].ndims = (int)nd;
                memcpy(f->var[i].dimid, dimid, sizeof(dimid));
                f->var[i].type = (int)type;
                f->var[i].begin = begin;
                f->var[i].vsize = vsize;
            }
        }
    }
#undef U4
#undef SKIPNAME
    return f;
}

/* m_rdrmed.f90:73-134.  vol is (kb:ke, ib:ie, jb:je), k fastest. */
int ora_rdrmed3d(int ib, int ie, int jb, int je, int kb, int ke, const char *fn, float *vol, char *err, size_t cap) {
    nc_file *f = nc_open_classic(fn, err, cap);
    if (!f) return -1;
    if (f->ndims < 3 || f->nvars < 4 || f->var[3].type != 5 /* NC_FLOAT */) {
        nc_close(f);
        snprintf(err, cap, "%s: expected dimensions x,y,z and a 4th variable of type float (gen_rmed3d.f90:91-126)", fn);
        return -1;
    }
    const long long nxc = f->dimlen[0], nyc = f->dimlen[1], nzc = f->dimlen[2]; /* dimension ids 1,2,3 */
    const long long begin = f->var[3].begin;
    if (begin + nxc * nyc * nzc * 4 > (long long)f->len) { nc_close(f); snprintf(err, cap, "%s: truncated volume", fn); return -1; }
    const int nk = ke - kb + 1, ni = ie - ib + 1;
#define VOL(k, i, j) vol[(size_t)((k) - kb) + (size_t)nk * ((size_t)((i) - ib) + (size_t)ni * (size_t)((j) - jb))]
    const int ktop = ke < nzc ? ke : (int)nzc;
    for (int k = kb; k <= ktop; k++) {
        const long long kk = k <= 0 ? k + nzc : k;
        const unsigned char *hh = f->buf + begin + (kk - 1) * nxc * nyc * 4;
        for (int j = jb; j <= je; j++) {
            long long jj = j % nyc;
            if (jj <= 0) jj += nyc;
            for (int i = ib; i <= ie; i++) {
                long long ii = i % nxc;
                if (ii <= 0) ii += nxc;
                unsigned u = (unsigned)be_u(hh + ((ii - 1) + nxc * (jj - 1)) * 4, 4);
                float v;
                memcpy(&v, &u, 4);
                VOL(k, i, j) = v;
            }
        }
    }
    for (long long k = nzc + 1; k <= ke; k++) { /* bottom cyclic part :128-131 */
        const long long kk = k % nzc;
        for (int j = jb; j <= je; j++)
            for (int i = ib; i <= ie; i++) VOL(k, i, j) = VOL(kk, i, j);
    }
#undef VOL
    nc_close(f);
    return 0;
}

/* m_rdrmed.f90:19-70 (swpc_psv).  vol is (kb:ke, ib:ie), k fastest; the file of tools/gen_rmed2d.f90:87-123 has dimensions
 * x, z and the section as its 3rd variable, x fastest. */
int ora_rdrmed2d(int ib, int ie, int kb, int ke, const char *fn, float *vol, char *err, size_t cap) {
    nc_file *f = nc_open_classic(fn, err, cap);
    if (!f) return -1;
    if (f->ndims < 2 || f->nvars < 3 || f->var[2].type != 5 /* NC_FLOAT */) {
        nc_close(f);
        snprintf(err, cap, "%s: expected dimensions x,z and a 3rd variable of type float (gen_rmed2d.f90:87-123)", fn);
        return -1;
    }
    const long long nxc = f->dimlen[0], nzc = f->dimlen[1];
    const long long begin = f->var[2].begin;
    if (begin + nxc * nzc * 4 > (long long)f->len) { nc_close(f); snprintf(err, cap, "%s: truncated section", fn); return -1; }
    const int nk = ke - kb + 1;
#define VOL(k, i) vol[(size_t)((k) - kb) + (size_t)nk * (size_t)((i) - ib)]
    const int ktop = ke < nzc ? ke : (int)nzc;
    for (int k = kb; k <= ktop; k++) {
        const long long kk = k <= 0 ? k + nzc : k;
        const unsigned char *hh = f->buf + begin + (kk - 1) * nxc * 4;
        for (int i = ib; i <= ie; i++) {
            long long ii = i % nxc;
            if (ii <= 0) ii += nxc;
            unsigned u = (unsigned)be_u(hh + (ii - 1) * 4, 4);
            float v;
            memcpy(&v, &u, 4);
            VOL(k, i) = v;
        }
    }
    for (long long k = nzc + 1; k <= ke; k++) { /* bottom cyclic part :64-68 (here with the kk <= 0 wrap the 3-D reader lacks) */
        long long kk = k % nzc;
        if (kk <= 0) kk += nzc;
        for (int i = ib; i <= ie; i++) VOL(k, i) = VOL(kk, i);
    }
#undef VOL
    nc_close(f);
    return 0;
}

/* one random-media volume over the memory box of r: the 3-D reader, or the section reader for a swpc_psv rank */
static int read_volume(const ora_rank *r, const char *path, float *xi, char *err, size_t cap) {
    if (r->psv) return ora_rdrmed2d(r->ibeg_m, r->iend_m, r->kbeg_m, r->kend_m, path, xi, err, cap);
    return ora_rdrmed3d(r->ibeg_m, r->iend_m, r->jbeg_m, r->jend_m, r->kbeg_m, r->kend_m, path, xi, err, cap);
}

/* ------------------------------------------------------------------------------------------ */
static int is_blank(const char *s) {
    while (*s) {
        if (!isspace((unsigned char)*s)) return 0;
        s++;
    }
    return 1;
}

static void join(const char *base, const char *fn, char *out, size_t cap) {
    if (fn[0] == '/' || !base || !base[0]) snprintf(out, cap, "%s", fn);
    else snprintf(out, cap, "%s/%s", base, fn);
}

typedef struct {
    int nl;
    float depth[256], rho0[256], vp0[256], vs0[256], qp0[256], qs0[256];
    char fn_rmed[256][ORA_STRLEN];
} layers;

/* layer table "depth rho vp vs Qp Qs [rmed-file]" with the velocity cut-off of m_vmodel_lgm.f90:92-101 */
static int read_layers(const char *path, int with_rmed, float vcut, layers *L, char *err, size_t cap) {
    FILE *fp = fopen(path, "r");
    if (!fp) { snprintf(err, cap, "cannot open layer file %s", path); return -1; }
    char line[1024];
    L->nl = 0;
    while (fgets(line, sizeof(line), fp) && L->nl < 256) {
        char *p = line;
        while (*p == ' ' || *p == '\t') p++;
        if (is_blank(p) || *p == '#') continue;
        for (char *q = p; *q; q++) if (*q == ',') *q = ' ';
        float v[6];
        char name[ORA_STRLEN] = "";
        int off = 0, got = 0;
        for (int c = 0; c < 6; c++) {
            char tok[64];
            int adv = 0;
            if (sscanf(p + off, " %63s%n", tok, &adv) != 1) break;
            for (char *t = tok; *t; t++) if (*t == 'd' || *t == 'D') *t = 'e';
            v[c] = strtof(tok, NULL);
            off += adv;
            got++;
        }
        if (got < 6) continue;
        if (with_rmed) {
            if (sscanf(p + off, " %255s", name) != 1) name[0] = 0;
            size_t ln = strlen(name); /* list-directed character read: quotes delimit */
            if (ln >= 2 && (name[0] == '\'' || name[0] == '"') && name[ln - 1] == name[0]) { memmove(name, name + 1, ln - 2); name[ln - 2] = 0; }
        }
        int l = L->nl++;
        L->depth[l] = v[0]; L->rho0[l] = v[1]; L->vp0[l] = v[2]; L->vs0[l] = v[3]; L->qp0[l] = v[4]; L->qs0[l] = v[5];
        snprintf(L->fn_rmed[l], ORA_STRLEN, "%s", name);
    }
    fclose(fp);
    if (L->nl == 0) { snprintf(err, cap, "no layer in %s", path); return -1; }
    for (int l = L->nl - 2; l >= 0; l--)
        if ((L->vp0[l] < vcut || L->vs0[l] < vcut) && (L->vp0[l] > 0 && L->vs0[l] > 0)) {
            L->vp0[l] = L->vp0[l + 1]; L->vs0[l] = L->vs0[l + 1]; L->rho0[l] = L->rho0[l + 1];
            L->qp0[l] = L->qp0[l + 1]; L->qs0[l] = L->qs0[l + 1];
        }
    return 0;
}

static void zs_cv(const ora_rank *r, int ef, int k, float *zs, float *Cv) {
    const double RE = ora_r_earth();
    const float zc = r->zc[k - r->kbeg_m];
    if (ef) {
        *zs = (float)(RE - RE * exp(-(double)zc / RE));
        *Cv = (float)exp((double)zc / RE);
    } else {
        *zs = zc;
        *Cv = 1.0f;
    }
}

static void fill_plane(ora_rank *r, int k, float rho1, float vp1, float vs1, float qp1, float qs1, float *qp, float *qs) {
    for (int j = r->jbeg_m; j <= r->jend_m; j++)
        for (int i = r->ibeg_m; i <= r->iend_m; i++) {
            size_t n = ora_idx3(r, k, i, j);
            r->rho[n] = rho1;
            r->mu[n] = rho1 * vs1 * vs1;
            r->lam[n] = rho1 * (vp1 * vp1 - 2 * vs1 * vs1);
            qp[n] = qp1;
            qs[n] = qs1;
        }
}

static void bd_dummy(ora_rank *r, float bd0) {
    size_t n2 = (size_t)r->nxm * r->nym;
    for (size_t n = 0; n < n2; n++) r->bddep[n] = bd0;
    for (int b = 1; b <= ORA_NBD; b++)
        for (size_t n = 0; n < n2; n++) r->bddep[(size_t)b * n2 + n] = -9999.0f;
}

/* m_vmodel_lgm.f90:20-175 */
int ora_vmodel_lgm(const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs, char *err, size_t cap) {
    char fn[ORA_STRLEN], path[2 * ORA_STRLEN];
    int use_munk, ef;
    ora_readini_c(ini, "fn_lhm", fn, "");
    ora_readini_l(ini, "munk_profile", &use_munk, 0);
    ora_readini_l(ini, "earth_flattening", &ef, 0);
    join(base, fn, path, sizeof(path));
    layers *L = (layers *)calloc(1, sizeof(layers));
    if (read_layers(path, 0, vcut, L, err, cap)) { free(L); return -1; }
    const int nl = L->nl;
    for (int k = r->kbeg_m; k <= r->kend_m; k++) {
        float zs, Cv, rho1, vp1, vs1, qp1, qs1;
        zs_cv(r, ef, k, &zs, &Cv);
        const float zc = r->zc[k - r->kbeg_m];
        if (zs < L->depth[0]) {
            if (zs < 0.0f) { rho1 = 0.001f; vp1 = 0.0f; vs1 = 0.0f; qp1 = 10.0f; qs1 = 10.0f; }
            /* swpc_psv/m_vmodel_lgm.f90:110 evaluates the sea-water profile at the spherical depth zs, the 3-D code at zc (:123) */
            else { rho1 = 1.0f; vp1 = Cv * ora_seawater_vel(r->psv ? zs : zc, use_munk); vs1 = 0.0f; qp1 = 1000000.0f; qs1 = 1000000.0f; }
        } else {
            rho1 = L->rho0[nl - 1]; vp1 = Cv * L->vp0[nl - 1]; vs1 = Cv * L->vs0[nl - 1]; qp1 = L->qp0[nl - 1]; qs1 = L->qs0[nl - 1];
            for (int l = 0; l < nl - 1; l++)
                if (L->depth[l] <= zs && zs < L->depth[l + 1]) {
                    const float dd = L->depth[l + 1] - L->depth[l], dz = zs - L->depth[l];
                    if (r->psv) { /* swpc_psv/m_vmodel_lgm.f90:128-132: (b - a) * dz / dd, not (b - a) / dd * dz */
                        rho1 = L->rho0[l] + (L->rho0[l + 1] - L->rho0[l]) * dz / dd;
                        vp1 = Cv * (L->vp0[l] + (L->vp0[l + 1] - L->vp0[l]) * dz / dd);
                        vs1 = Cv * (L->vs0[l] + (L->vs0[l + 1] - L->vs0[l]) * dz / dd);
                        qp1 = L->qp0[l] + (L->qp0[l + 1] - L->qp0[l]) * dz / dd;
                        qs1 = L->qs0[l] + (L->qs0[l + 1] - L->qs0[l]) * dz / dd;
                        break;
                    }
                    rho1 = L->rho0[l] + (L->rho0[l + 1] - L->rho0[l]) / dd * dz;
                    vp1 = Cv * (L->vp0[l] + (L->vp0[l + 1] - L->vp0[l]) / dd * dz);
                    vs1 = Cv * (L->vs0[l] + (L->vs0[l + 1] - L->vs0[l]) / dd * dz);
                    qp1 = L->qp0[l] + (L->qp0[l + 1] - L->qp0[l]) / dd * dz;
                    qs1 = L->qs0[l] + (L->qs0[l + 1] - L->qs0[l]) / dd * dz;
                    break;
                }
        }
        fill_plane(r, k, rho1, vp1, vs1, qp1, qs1, qp, qs);
    }
    bd_dummy(r, L->depth[0]);
    free(L);
    return 0;
}

/* m_fdtool.f90:15-48 */
static void vcheck(float *vp, float *vs, float *rho, float xi, float vmin, float vmax, float rhomin) {
    float gamma = *vp / *vs;
    if (gamma < FLT_EPS) gamma = sqrtf(3.0f);
    if (*vp > vmax || *vs > vmax) {
        const float xi2 = (1 + xi) * vmax / *vp - 1;
        *vp = vmax;
        *vs = vmax / gamma;
        *rho = *rho * (1 + 0.8f * xi2) / (1 + 0.8f * xi);
    }
    if (*vp < vmin || *vs < vmin) {
        *vs = vmin;
        *vp = vmin * gamma;
    }
    if (*rho < rhomin) *rho = rhomin;
}

/* vmax of the random-media models: cc * dh / dt, m_vmodel_uni_rmed.f90:79-83 (dx, dy, dz are real(MP)) */
static float rmed_vmax(const ora_cfg *c) {
    /* swpc_psv (ny = 0 in the cfg the P-SV oracle builds): dh = 1 / sqrt(1/dx**2 + 1/dz**2), swpc_psv/m_vmodel_uni_rmed.f90:80 */
    const float dh = c->ny == 0 ? (float)(1.0 / sqrt(1.0 / (c->dx * c->dx) + 1.0 / (c->dz * c->dz)))
                                : (float)(1.0 / sqrt(1.0 / (c->dx * c->dx) + 1.0 / (c->dy * c->dy) + 1.0 / (c->dz * c->dz)));
    const float cc = 6.0f / 7.0f;
    return cc * dh / c->dt;
}

static int file_exists(const char *p) {
    FILE *fp = fopen(p, "rb");
    if (!fp) return 0;
    fclose(fp);
    return 1;
}

/* m_vmodel_uni_rmed.f90:22-170 */
int ora_vmodel_uni_rmed(const ora_cfg *c, const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs, char *err, size_t cap) {
    float vp0, vs0, rho0, qp0, qs0, topo0, rhomin;
    int use_munk, ef;
    ora_readini_s(ini, "vp0", &vp0, 5.0f);
    ora_readini_s(ini, "vs0", &vs0, vp0 / sqrtf(3.0f));
    ora_readini_s(ini, "rho0", &rho0, 2.7f);
    ora_readini_s(ini, "qp0", &qp0, 1000000.0f);
    ora_readini_s(ini, "qs0", &qs0, 1000000.0f);
    ora_readini_s(ini, "topo0", &topo0, 0.0f);
    ora_readini_s(ini, "rhomin", &rhomin, 1.0f);
    ora_readini_l(ini, "munk_profile", &use_munk, 0);
    ora_readini_l(ini, "earth_flattening", &ef, 0);
    const float vmin = vcut, vmax = rmed_vmax(c);
    char dir[ORA_STRLEN], fn[ORA_STRLEN], rel[2 * ORA_STRLEN + 2], path[3 * ORA_STRLEN];
    ora_readini_c(ini, "dir_rmed", dir, "");
    ora_readini_c(ini, "fn_rmed0", fn, "");
    snprintf(rel, sizeof(rel), "%s/%s", dir, fn);
    join(base, rel, path, sizeof(path));
    float *xi = (float *)calloc(r->ncell_m, sizeof(float));
    if (file_exists(path) && read_volume(r, path, xi, err, cap)) { free(xi); return -1; }
    for (int j = r->jbeg_m; j <= r->jend_m; j++)
        for (int i = r->ibeg_m; i <= r->iend_m; i++)
            for (int k = r->kbeg_m; k <= r->kend_m; k++) {
                float zs, Cv;
                zs_cv(r, ef, k, &zs, &Cv);
                const float zc = r->zc[k - r->kbeg_m];
                size_t n = ora_idx3(r, k, i, j);
                if (zs > topo0) {
                    float rho2 = (1.0f + 0.8f * xi[n]) * rho0;
                    float vp2 = (1.0f + xi[n]) * Cv * vp0;
                    float vs2 = (1.0f + xi[n]) * Cv * vs0;
                    vcheck(&vp2, &vs2, &rho2, xi[n], vmin, vmax, rhomin);
                    r->rho[n] = rho2;
                    r->mu[n] = r->rho[n] * vs2 * vs2;
                    r->lam[n] = r->rho[n] * (vp2 * vp2 - 2 * vs2 * vs2);
                    qp[n] = qp0;
                    qs[n] = qs0;
                } else if (zs > 0.0f) {
                    const float vp1 = Cv * ora_seawater_vel(zc, use_munk), vs1 = 0.0f;
                    r->rho[n] = 1.0f;
                    r->mu[n] = r->rho[n] * vs1 * vs1;
                    r->lam[n] = r->rho[n] * (vp1 * vp1 - 2 * vs1 * vs1);
                    qp[n] = 1000000.0f;
                    qs[n] = 1000000.0f;
                } else {
                    const float vp1 = 0.0f, vs1 = 0.0f;
                    r->rho[n] = 0.001f;
                    r->mu[n] = r->rho[n] * vs1 * vs1;
                    r->lam[n] = r->rho[n] * (vp1 * vp1 - 2 * vs1 * vs1);
                    qp[n] = 10.0f;
                    qs[n] = 10.0f;
                }
            }
    free(xi);
    bd_dummy(r, topo0);
    return 0;
}

/* m_fdtool.f90:815-855 + the volume reads of m_vmodel_lhm_rmed.f90:129-141 */
static int read_rmed_set(const ora_ini *ini, const char *base, const layers *L, ora_rank *r, int *tbl, float **xi_out, char *err, size_t cap) {
    char dir[ORA_STRLEN];
    ora_readini_c(ini, "dir_rmed", dir, "");
    char (*uniq)[2 * ORA_STRLEN + 2] = calloc((size_t)L->nl, sizeof(*uniq));
    int nind = 0;
    for (int l = 0; l < L->nl; l++) {
        char full[2 * ORA_STRLEN + 2];
        snprintf(full, sizeof(full), "%s/%s", dir, L->fn_rmed[l]);
        int found = -1;
        for (int q = 0; q < nind; q++) if (!strcmp(uniq[q], full)) { found = q; break; }
        if (found < 0) { found = nind; strcpy(uniq[nind++], full); }
        tbl[l] = found;
    }
    float *xi = (float *)calloc(r->ncell_m * (size_t)nind, sizeof(float));
    for (int q = 0; q < nind; q++) {
        char path[3 * ORA_STRLEN];
        join(base, uniq[q], path, sizeof(path));
        if (file_exists(path) && read_volume(r, path, xi + r->ncell_m * (size_t)q, err, cap)) {
            free(xi);
            free(uniq);
            return -1;
        }
    }
    free(uniq);
    *xi_out = xi;
    return 0;
}

/* m_vmodel_lhm_rmed.f90:22-244 */
int ora_vmodel_lhm_rmed(const ora_cfg *c, const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs, char *err, size_t cap) {
    char fn[ORA_STRLEN], path[2 * ORA_STRLEN];
    int use_munk, ef;
    float rhomin;
    ora_readini_c(ini, "fn_lhm_rmed", fn, "");
    ora_readini_s(ini, "rhomin", &rhomin, 1.0f);
    ora_readini_l(ini, "munk_profile", &use_munk, 0);
    ora_readini_l(ini, "earth_flattening", &ef, 0);
    join(base, fn, path, sizeof(path));
    layers *L = (layers *)calloc(1, sizeof(layers));
    if (read_layers(path, 1, vcut, L, err, cap)) { free(L); return -1; }
    const float vmin = vcut, vmax = rmed_vmax(c);
    int tbl[256];
    float *xi = NULL;
    if (read_rmed_set(ini, base, L, r, tbl, &xi, err, cap)) { free(L); return -1; }
    for (int k = r->kbeg_m; k <= r->kend_m; k++) {
        float zs, Cv;
        zs_cv(r, ef, k, &zs, &Cv);
        const float zc = r->zc[k - r->kbeg_m];
        /* swpc_psv/m_vmodel_lhm_rmed.f90:148-179 tests the grid depth zc where the 3-D code tests the spherical depth zs */
        if (r->psv) zs = zc;
        if (zs < L->depth[0]) {
            if (zs < 0.0f) {
                for (int j = r->jbeg_m; j <= r->jend_m; j++)
                    for (int i = r->ibeg_m; i <= r->iend_m; i++) {
                        size_t n = ora_idx3(r, k, i, j);
                        r->rho[n] = 0.001f; r->mu[n] = 0.0f; r->lam[n] = 0.0f; qp[n] = 10.0f; qs[n] = 10.0f;
                    }
            } else {
                const float vp1 = Cv * ora_seawater_vel(zc, use_munk);
                for (int j = r->jbeg_m; j <= r->jend_m; j++)
                    for (int i = r->ibeg_m; i <= r->iend_m; i++) {
                        size_t n = ora_idx3(r, k, i, j);
                        r->rho[n] = 1.0f; r->mu[n] = 0.0f; r->lam[n] = 1.0f * vp1 * vp1; qp[n] = 1000000.0f; qs[n] = 1000000.0f;
                    }
            }
            continue;
        }
        for (int j = r->jbeg_m; j <= r->jend_m; j++)
            for (int i = r->ibeg_m; i <= r->iend_m; i++) {
                size_t n = ora_idx3(r, k, i, j);
                float rho1 = 0, vp1 = 0, vs1 = 0, qp1 = 0, qs1 = 0;
                for (int l = 0; l < L->nl; l++)
                    if (zs >= L->depth[l]) {
                        const float x = xi[r->ncell_m * (size_t)tbl[l] + n];
                        rho1 = L->rho0[l] * (1 + 0.8f * x);
                        vp1 = Cv * L->vp0[l] * (1 + x);
                        vs1 = Cv * L->vs0[l] * (1 + x);
                        if (L->vp0[l] > 0 && L->vs0[l] > 0) vcheck(&vp1, &vs1, &rho1, x, vmin, vmax, rhomin);
                        qp1 = L->qp0[l];
                        qs1 = L->qs0[l];
                    }
                r->rho[n] = rho1;
                r->mu[n] = rho1 * vs1 * vs1;
                r->lam[n] = rho1 * (vp1 * vp1 - 2 * vs1 * vs1);
                qp[n] = qp1;
                qs[n] = qs1;
            }
    }
    free(xi);
    bd_dummy(r, L->depth[0]);
    free(L);
    return 0;
}

/* m_vmodel_lgm_rmed.f90:22-258.  Quirk kept: inside the (i,j) loops the reference assigns the whole plane
 * `rho(k, i0:i1, j0:j1) = rho1` (:236-240), so every plane ends up laterally uniform with the value computed at the
 * LAST (i1, j1) of the rank's memory box. */
int ora_vmodel_lgm_rmed(const ora_cfg *c, const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs, char *err, size_t cap) {
    char fn[ORA_STRLEN], path[2 * ORA_STRLEN];
    int use_munk, ef;
    float rhomin;
    ora_readini_c(ini, "fn_lhm_rmed", fn, "");
    ora_readini_s(ini, "rhomin", &rhomin, 1.0f);
    ora_readini_l(ini, "munk_profile", &use_munk, 0);
    ora_readini_l(ini, "earth_flattening", &ef, 0);
    join(base, fn, path, sizeof(path));
    layers *L = (layers *)calloc(1, sizeof(layers));
    if (read_layers(path, 1, vcut, L, err, cap)) { free(L); return -1; }
    const float vmin = vcut, vmax = rmed_vmax(c);
    int tbl[256];
    float *xi = NULL;
    if (read_rmed_set(ini, base, L, r, tbl, &xi, err, cap)) { free(L); return -1; }
    const int nl = L->nl;
    for (int k = r->kbeg_m; k <= r->kend_m; k++) {
        float zs, Cv, rho1, vp1, vs1, qp1, qs1;
        zs_cv(r, ef, k, &zs, &Cv);
        const float zc = r->zc[k - r->kbeg_m];
        if (zs < L->depth[0]) {
            if (zs < 0.0f) { rho1 = 0.001f; vp1 = 0.0f; vs1 = 0.0f; qp1 = 10.0f; qs1 = 10.0f; }
            else { rho1 = 1.0f; vp1 = Cv * ora_seawater_vel(zc, use_munk); vs1 = 0.0f; qp1 = 1000000.0f; qs1 = 1000000.0f; }
        } else {
            const size_t n = ora_idx3(r, k, r->iend_m, r->jend_m);
            const float xl = xi[r->ncell_m * (size_t)tbl[nl - 1] + n];
            rho1 = L->rho0[nl - 1] * (1 + 0.8f * xl);
            vp1 = Cv * L->vp0[nl - 1] * (1 + xl);
            vs1 = Cv * L->vs0[nl - 1] * (1 + xl);
            qp1 = L->qp0[nl - 1];
            qs1 = L->qs0[nl - 1];
            for (int l = 0; l < nl - 1; l++)
                if (L->depth[l] <= zs && zs < L->depth[l + 1]) {
                    const float dd = L->depth[l + 1] - L->depth[l], dz = zs - L->depth[l];
                    rho1 = L->rho0[l] + (L->rho0[l + 1] - L->rho0[l]) / dd * dz;
                    vp1 = Cv * (L->vp0[l] + (L->vp0[l + 1] - L->vp0[l]) / dd * dz);
                    vs1 = Cv * (L->vs0[l] + (L->vs0[l + 1] - L->vs0[l]) / dd * dz);
                    qp1 = L->qp0[l] + (L->qp0[l + 1] - L->qp0[l]) / dd * dz;
                    qs1 = L->qs0[l] + (L->qs0[l + 1] - L->qs0[l]) / dd * dz;
                    const float x = xi[r->ncell_m * (size_t)tbl[l] + n];
                    rho1 = rho1 * (1 + 0.8f * x);
                    vp1 = vp1 * (1 + x);
                    vs1 = vs1 * (1 + x);
                    if (L->vp0[l] > 0 && L->vs0[l] > 0) vcheck(&vp1, &vs1, &rho1, x, vmin, vmax, rhomin);
                    break;
                }
        }
        fill_plane(r, k, rho1, vp1, vs1, qp1, qs1, qp, qs);
    }
    free(xi);
    bd_dummy(r, L->depth[0]);
    free(L);
    return 0;
}

/* m_medium.f90:273-337; vmax is the GLOBAL maximum (velocity_minmax :396-427 runs before it) */
void ora_stabilize_absorber(const ora_cfg *c, ora_rank *r) {
    const float V_DYNAMIC_RANGE = 0.4f;
    const int LV_THICK = 20;
    const float vmin_pml = c->vmax * V_DYNAMIC_RANGE;
    for (int j = r->jbeg - 1; j <= r->jend + 1; j++)
        for (int i = r->ibeg - 1; i <= r->iend + 1; i++) {
            int k = 1 << 30;
            for (int jj = j - 2; jj <= j + 2; jj++)
                for (int ii = i - 2; ii <= i + 2; ii++)
                    if (r->kbeg_a[ora_idx2(r, ii, jj)] < k) k = r->kbeg_a[ora_idx2(r, ii, jj)];
            while (k <= r->kend) {
                size_t n = ora_idx3(r, k, i, j), m = ora_idx3(r, k - 1, i, j);
                if (r->lam[n] < r->lam[m] || r->mu[n] < r->mu[m]) {
                    int k2;
                    for (k2 = k + 1; k2 <= r->kend; k2++) {
                        size_t a = ora_idx3(r, k2, i, j), b = ora_idx3(r, k2 - 1, i, j);
                        if (r->lam[a] > r->lam[b] || r->mu[a] > r->mu[b]) break;
                    }
                    if (k2 - k <= LV_THICK) {
                        r->rho[n] = r->rho[m]; r->lam[n] = r->lam[m]; r->mu[n] = r->mu[m]; r->taup[n] = r->taup[m]; r->taus[n] = r->taus[m];
                        k = k2 - 1;
                    }
                }
                k = k + 1;
            }
        }
    for (int j = r->jbeg - 1; j <= r->jend + 1; j++)
        for (int i = r->ibeg - 1; i <= r->iend + 1; i++) {
            int k0 = 1 << 30;
            for (int jj = j - 2; jj <= j + 2; jj++)
                for (int ii = i - 2; ii <= i + 2; ii++)
                    if (r->kbeg_a[ora_idx2(r, ii, jj)] < k0) k0 = r->kbeg_a[ora_idx2(r, ii, jj)];
            for (int k = k0; k <= r->kend; k++) {
                size_t n = ora_idx3(r, k, i, j);
                float vp = sqrtf((r->lam[n] + 2 * r->mu[n]) / r->rho[n]);
                float vs = sqrtf(r->mu[n] / r->rho[n]);
                if (vs < FLT_EPS) continue;
                const float gamma = sqrtf(3.0f);
                if (vs < vmin_pml) {
                    vs = vmin_pml;
                    vp = vs * gamma;
                    r->lam[n] = r->rho[n] * (vp * vp - 2 * (vs * vs));
                    r->mu[n] = r->rho[n] * (vs * vs);
                }
            }
        }
}

/* ------------------------------------------------------------------------------------------ */
/* vmodel_grd, src/swpc_3d/m_vmodel_grd.f90:28-303: layered 3-D model whose interfaces are GMT grids in geographic
 * coordinates, interpolated bicubically (src/shared/m_bicubic.f90) onto the FDM columns.  The reference reads the grids
 * through the netCDF library; this restatement (and the product's host) reads netCDF *classic* files only -- grids in
 * GMT's default netCDF-4 container have to be converted first (e.g. `gmt grdconvert in.grd out.grd=cf`, `nccopy -k classic`).
 * The summation order inside bicubic__coef's matmul is not defined by the Fortran standard; plain left-to-right sums are
 * used here and in the product. */
static double nc_value(const nc_file *f, int v, long long idx) {
    const unsigned char *p = f->buf + f->var[v].begin;
    switch (f->var[v].type) {
    case 5: { unsigned u = (unsigned)be_u(p + 4 * idx, 4); float x; memcpy(&x, &u, 4); return (double)x; }
    case 6: { unsigned long long u = be_u(p + 8 * idx, 8); double x; memcpy(&x, &u, 8); return x; }
    case 4: return (double)(int)be_u(p + 4 * idx, 4);
    case 3: return (double)(short)be_u(p + 2 * idx, 2);
    default: return 0.0;
    }
}

typedef struct {
    int nx, ny;
    double x0, y0, dx, dy;
    double *f, *fx, *fy, *fxy;
    int ii0, jj0, first;
    double aa[4][4]; /* aa[j][i] = aa(i, j) */
} bicubic;

static void bc_diffx(int nx, int ny, double dx, const double *ff, double *out) { /* m_bicubic.f90:291-313 */
    for (int j = 0; j < ny; j++) {
        for (int i = 1; i < nx - 1; i++) out[i + nx * j] = (ff[i + 1 + nx * j] - ff[i - 1 + nx * j]) / (2 * dx);
        out[nx * j] = (ff[1 + nx * j] - ff[nx * j]) / dx;
        out[nx - 1 + nx * j] = (ff[nx - 1 + nx * j] - ff[nx - 2 + nx * j]) / dx;
    }
}
static void bc_diffy(int nx, int ny, double dy, const double *ff, double *out) { /* :316-338 */
    for (int i = 0; i < nx; i++) {
        for (int j = 1; j < ny - 1; j++) out[i + nx * j] = (ff[i + nx * (j + 1)] - ff[i + nx * (j - 1)]) / (2 * dy);
        out[i] = (ff[i + nx] - ff[i]) / dy;
        out[i + nx * (ny - 1)] = (ff[i + nx * (ny - 1)] - ff[i + nx * (ny - 2)]) / dy;
    }
}
static void bc_init(bicubic *b, int nx, int ny, double x0, double y0, double dx, double dy, const double *dat) { /* :72-110 */
    const size_t n = (size_t)nx * ny;
    b->nx = nx; b->ny = ny; b->x0 = x0; b->y0 = y0; b->dx = dx; b->dy = dy;
    b->f = (double *)malloc(4 * n * sizeof(double));
    b->fx = b->f + n; b->fy = b->fx + n; b->fxy = b->fy + n;
    memcpy(b->f, dat, n * sizeof(double));
    bc_diffx(nx, ny, dx, b->f, b->fx);
    bc_diffy(nx, ny, dy, b->f, b->fy);
    bc_diffx(nx, ny, dx, b->fy, b->fxy);
    for (size_t q = 0; q < n; q++) { b->fx[q] = b->fx[q] * dx; b->fy[q] = b->fy[q] * dy; b->fxy[q] = b->fxy[q] * dx * dy; }
    b->first = 1; b->ii0 = b->jj0 = 0;
}
static void bc_coef(bicubic *b, int ii, int jj) { /* :231-288; ii, jj 1-based */
    static const double mat[16][16] = {
        {1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
        {-3, 3, 0, 0, -2, -1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {2, -2, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
        {0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0},
        {0, 0, 0, 0, 0, 0, 0, 0, -3, 3, 0, 0, -2, -1, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 2, -2, 0, 0, 1, 1, 0, 0},
        {-3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0, 0, 0, 0, 0}, {0, 0, 0, 0, -3, 0, 3, 0, 0, 0, 0, 0, -2, 0, -1, 0},
        {9, -9, -9, 9, 6, 3, -6, -3, 6, -6, 3, -3, 4, 2, 2, 1}, {-6, 6, 6, -6, -3, -3, 3, 3, -4, 4, -2, 2, -2, -2, -1, -1},
        {2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 2, 0, -2, 0, 0, 0, 0, 0, 1, 0, 1, 0},
        {-6, 6, 6, -6, -4, -2, 4, 2, -3, 3, -3, 3, -2, -1, -2, -1}, {4, -4, -4, 4, 2, 2, -2, -2, 2, -2, 2, -2, 1, 1, 1, 1}};
    const int nx = b->nx, i0 = ii - 1, j0 = jj - 1;
    double xx[16];
    const double *src[4] = {b->f, b->fx, b->fy, b->fxy};
    for (int q = 0; q < 4; q++) {
        xx[4 * q] = src[q][i0 + nx * j0]; xx[4 * q + 1] = src[q][i0 + 1 + nx * j0];
        xx[4 * q + 2] = src[q][i0 + nx * (j0 + 1)]; xx[4 * q + 3] = src[q][i0 + 1 + nx * (j0 + 1)];
    }
    for (int r = 0; r < 16; r++) {
        double acc = 0.0;
        for (int l = 0; l < 16; l++) acc = acc + mat[r][l] * xx[l];
        b->aa[r / 4][r % 4] = acc; /* aa(0:3, r/4) = rows 4*(r/4)+1 .. +4 */
    }
}
static double bc_interp(bicubic *b, double xi, double yi) { /* :141-228, no default value */
    const double x0 = b->x0, y0 = b->y0, dx = b->dx, dy = b->dy;
    const int nx = b->nx, ny = b->ny;
    int ii = (int)floor((xi - x0) / dx) + 1, jj = (int)floor((yi - y0) / dy) + 1;
    double xi2 = xi, yi2 = yi;
    if (ii < 1 || ii > nx - 1 || jj <= 0 || jj > ny - 1) {
        if (ii < 1) { ii = 1; xi2 = x0; }
        if (ii > nx - 1) { ii = nx - 1; xi2 = x0 + (nx - 1) * dx; }
        if (jj < 1) { jj = 1; yi2 = y0; }
        if (jj > ny - 1) { jj = ny - 1; yi2 = y0 + (ny - 1) * dy; }
    }
    if (b->ii0 != ii || b->jj0 != jj || b->first) bc_coef(b, ii, jj);
    const double xd = (xi2 - (x0 + (ii - 1) * dx)) / dx, yd = (yi2 - (y0 + (jj - 1) * dy)) / dy;
    const double xda[4] = {1.0, xd, xd * xd, xd * xd * xd}, yda[4] = {1.0, yd, yd * yd, yd * yd * yd};
    double v = 0.0;
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++) v = v + b->aa[j][i] * xda[i] * yda[j];
    b->first = 0; b->ii0 = ii; b->jj0 = jj;
    return v;
}

/* with_rmed: vmodel_grd_rmed (m_vmodel_grd_rmed.f90:28-388) -- the same layers, each perturbed by a random-media volume that is
 * indexed by the depth below a reference interface (reflyr) */
int ora_vmodel_grd(const ora_cfg *c, const ora_ini *ini, const char *base, ora_rank *r, float vcut, float *qp, float *qs, int with_rmed, char *err, size_t cap) {
    char fn_lst[ORA_STRLEN], dir_grd[ORA_STRLEN], path[3 * ORA_STRLEN];
    int is_ocean, is_flatten, use_munk, ef, node_grd;
    ora_readini_c(ini, with_rmed ? "fn_grdlst_rmed" : "fn_grdlst", fn_lst, ".");
    ora_readini_c(ini, "dir_grd", dir_grd, ".");
    ora_readini_i(ini, "node_grd", &node_grd, 0);
    ora_readini_l(ini, "is_ocean", &is_ocean, 1);
    ora_readini_l(ini, "topo_flatten", &is_flatten, 0);
    if (is_flatten) is_ocean = 1;
    ora_readini_l(ini, "munk_profile", &use_munk, 0);
    ora_readini_l(ini, "earth_flattening", &ef, 0);
    const double RE = ora_r_earth();
    const int nk = r->nzm;
    float *Cv = (float *)malloc(sizeof(float) * (size_t)nk);
    for (int k = r->kbeg_m; k <= r->kend_m; k++) Cv[k - r->kbeg_m] = ef ? (float)exp((double)r->zc[k - r->kbeg_m] / RE) : 1.0f;
    /* air, then ocean :88-133 */
    for (int j = r->jbeg_m; j <= r->jend_m; j++)
        for (int i = r->ibeg_m; i <= r->iend_m; i++)
            for (int k = r->kbeg_m; k <= r->kend_m; k++) {
                const size_t n = ora_idx3(r, k, i, j);
                const float zc = r->zc[k - r->kbeg_m];
                float vp0 = 0.0f, vs0 = 0.0f, rho0 = 0.001f, qp0 = 1.0f, qs0 = 1.0f;
                if (is_ocean && !(zc < 0)) { vp0 = Cv[k - r->kbeg_m] * ora_seawater_vel(zc, use_munk); rho0 = 1.0f; qp0 = 1000000.0f; qs0 = 1000000.0f; }
                r->rho[n] = rho0; r->lam[n] = rho0 * (vp0 * vp0 - 2 * vs0 * vs0); r->mu[n] = rho0 * vs0 * vs0; qp[n] = qp0; qs[n] = qs0;
            }
    /* geographic location of every column, clamped to the inner edge of the absorber :137-151 */
    const float dx = (float)c->dx, dy = (float)c->dy, dz = (float)c->dz;
    const float x_AB = ora_i2x(c->na + 1, c->xbeg, dx), x_AE = ora_i2x(c->nx - c->na, c->xbeg, dx);
    const float y_AB = ora_i2x(c->na + 1, c->ybeg, dy), y_AE = ora_i2x(c->ny - c->na, c->ybeg, dy);
    const size_t n2 = (size_t)r->nxm * r->nym;
    float *glon = (float *)malloc(sizeof(float) * n2), *glat = (float *)malloc(sizeof(float) * n2);
    for (int j = r->jbeg_m; j <= r->jend_m; j++)
        for (int i = r->ibeg_m; i <= r->iend_m; i++) {
            if (r->psv) { /* swpc_psv/m_vmodel_grd.f90:127-129: the section runs along y = 0 and is not clamped to the absorber edge */
                ora_geomap_c2g(r->xc[i - r->ibeg_m], 0.0f, c->clon, c->clat, c->phi, &glon[ora_idx2(r, i, j)], &glat[ora_idx2(r, i, j)]);
                continue;
            }
            const float xc = r->xc[i - r->ibeg_m], yc = r->yc[j - r->jbeg_m];
            const float xx = fminf(fmaxf(xc, x_AB), x_AE), yy = fminf(fmaxf(yc, y_AB), y_AE);
            ora_geomap_c2g(xx, yy, c->clon, c->clat, c->phi, &glon[ora_idx2(r, i, j)], &glat[ora_idx2(r, i, j)]);
        }
    /* layer list :154-173 */
    join(base, fn_lst, path, sizeof(path));
    FILE *fp = fopen(path, "r");
    if (!fp) { snprintf(err, cap, "vmodel_grd: cannot open the layer list %s", path); free(Cv); free(glon); free(glat); return -1; }
    int ngrd = 0;
    static char fn_grd[64][ORA_STRLEN];
    float rho1[64], vp1[64], vs1[64], qp1[64], qs1[64];
    int pid[64], reflyr[64];
    layers *LR = with_rmed ? (layers *)calloc(1, sizeof(layers)) : NULL;   /* only fn_rmed[] / nl are used */
    char line[1024];
    while (fgets(line, sizeof(line), fp) && ngrd < 64) {
        char *p = line;
        while (*p == ' ' || *p == '\t') p++;
        if (*p == '#' || is_blank(p)) continue;
        for (char *t = p; *t; t++) if (*t == ',') *t = ' ';
        char name[ORA_STRLEN], rname[ORA_STRLEN] = "";
        reflyr[ngrd] = 0;
        const int got = sscanf(p, " %255s %f %f %f %f %f %d %255s %d", name, &rho1[ngrd], &vp1[ngrd], &vs1[ngrd], &qp1[ngrd], &qs1[ngrd], &pid[ngrd], rname, &reflyr[ngrd]);
        if (got < (with_rmed ? 9 : 7)) continue;
        size_t ln = strlen(name);
        if (ln >= 2 && (name[0] == '\'' || name[0] == '"') && name[ln - 1] == name[0]) { memmove(name, name + 1, ln - 2); name[ln - 2] = 0; }
        if (with_rmed) {
            ln = strlen(rname);
            if (ln >= 2 && (rname[0] == '\'' || rname[0] == '"') && rname[ln - 1] == rname[0]) { memmove(rname, rname + 1, ln - 2); rname[ln - 2] = 0; }
            snprintf(LR->fn_rmed[ngrd], ORA_STRLEN, "%s", rname);
        }
        snprintf(fn_grd[ngrd], ORA_STRLEN, "%.120s/%.120s", dir_grd, name);
        ngrd++;
    }
    fclose(fp);
    if (ngrd == 0) { snprintf(err, cap, "vmodel_grd: no layer in the list %s", path); free(LR); free(Cv); free(glon); free(glat); return -1; }
    for (int n = ngrd - 2; n >= 0; n--) {
        /* swpc_psv/m_vmodel_grd_rmed.f90:185 drops the "(vp1 > 0 .and. vs1 > 0)" clause the other three variants have */
        const int positive = (r->psv && with_rmed) ? 1 : (vp1[n] > 0 && vs1[n] > 0);
        if ((vp1[n] < vcut || vs1[n] < vcut) && positive) { vp1[n] = vp1[n + 1]; vs1[n] = vs1[n + 1]; rho1[n] = rho1[n + 1]; qp1[n] = qp1[n + 1]; qs1[n] = qs1[n + 1]; }
    }
    int tbl[64];
    float *xi = NULL, rhomin = 1.0f;
    const float vmin = vcut, vmax = rmed_vmax(c);
    if (with_rmed) {
        ora_readini_s(ini, "rhomin", &rhomin, 1.0f);
        LR->nl = ngrd;
        for (int n = 0; n < ngrd; n++)
            if (!(0 <= reflyr[n] && reflyr[n] <= ngrd)) { snprintf(err, cap, "assert: 0 <= reflyr <= ngrd (m_vmodel_grd_rmed.f90:208)"); free(LR); free(Cv); free(glon); free(glat); return -1; }
        if (read_rmed_set(ini, base, LR, r, tbl, &xi, err, cap)) { free(LR); free(Cv); free(glon); free(glat); return -1; }
    }
    int *kgrd = (int *)malloc(sizeof(int) * n2 * (size_t)(ngrd + 1));   /* kgrd(0:ngrd, i, j) */
    for (size_t q = 0; q < n2; q++) kgrd[q] = r->kbeg_m - 1;
    const int ktopo = ora_x2i(0.0f - dz / 2, c->zbeg, dz);
    size_t nbd = n2;
    for (int b = 1; b <= ORA_NBD; b++)
        for (size_t q = 0; q < nbd; q++) r->bddep[(size_t)b * n2 + q] = 0.0f;   /* bd is intent(out): only bd(:,:,0) and bd(:,:,pid) are set */
    int rc = 0;
    for (int n = 1; n <= ngrd && !rc; n++) {
        char full[4 * ORA_STRLEN];
        join(base, fn_grd[n - 1], full, sizeof(full));
        nc_file *f = nc_open_classic(full, err, cap);
        if (!f) { rc = -1; break; }
        if (f->ndims != 2 || f->nvars < 3) { snprintf(err, cap, "%s: expected a 2-D grid (x, y, z)", full); nc_close(f); rc = -1; break; }
        const int nlon = (int)f->dimlen[0], nlat = (int)f->dimlen[1];
        double *dep = (double *)malloc(sizeof(double) * (size_t)nlon * nlat);
        for (long long q = 0; q < (long long)nlon * nlat; q++) dep[q] = nc_value(f, 2, q) / 1000;   /* m -> km */
        const double lon0 = nc_value(f, 0, 0), lonN = nc_value(f, 0, nlon - 1), lat0 = nc_value(f, 1, 0), latN = nc_value(f, 1, nlat - 1);
        const double dlon = (lonN - lon0) / (nlon - 1), dlat = (latN - lat0) / (nlat - 1);
        nc_close(f);
        bicubic bc;
        bc_init(&bc, nlon, nlat, lon0, lat0, dlon, dlat, dep);
        free(dep);
        for (int j = r->jbeg_m; j <= r->jend_m; j++)
            for (int i = r->ibeg_m; i <= r->iend_m; i++) {
                const size_t q = ora_idx2(r, i, j);
                float zgrd = (float)bc_interp(&bc, (double)glon[q], (double)glat[q]);
                if (ef) zgrd = (float)(-RE * log((RE - (double)zgrd) / RE));
                if (n == 1) r->bddep[q] = zgrd;
                if (is_flatten) zgrd = zgrd - r->bddep[q];
                int kg = ora_x2i(zgrd - dz / 2, c->zbeg, dz);
                if (kg < kgrd[(size_t)(n - 1) * n2 + q]) kg = kgrd[(size_t)(n - 1) * n2 + q];
                if (n == 1 && zgrd > 0 && kg < ktopo + 2) kg = ktopo + 2;   /* sea column thickness >= 2 */
                kgrd[(size_t)n * n2 + q] = kg;
                if (pid[n - 1] > 0 && pid[n - 1] <= ORA_NBD) r->bddep[(size_t)pid[n - 1] * n2 + q] = zgrd;
            }
        free(bc.f);
    }
    if (!rc)
        for (int j = r->jbeg_m; j <= r->jend_m; j++)
            for (int i = r->ibeg_m; i <= r->iend_m; i++)
                for (int n = 1; n <= ngrd; n++)
                    for (int k = kgrd[(size_t)n * n2 + ora_idx2(r, i, j)] + 1; k <= r->kend_m; k++) {
                        if (k < r->kbeg_m) continue;
                        const size_t m = ora_idx3(r, k, i, j);
                        const float cv = Cv[k - r->kbeg_m];
                        if (with_rmed) { /* m_vmodel_grd_rmed.f90:340-366 */
                            int kk = k - kgrd[(size_t)reflyr[n - 1] * n2 + ora_idx2(r, i, j)] + 1;   /* relative depth index */
                            if (kk < r->kbeg_m) kk = kk + c->nz;
                            if (kk > r->kend_m) kk = kk - r->kend_m;
                            if (!r->psv && !(vp1[n - 1] < vmax && vs1[n - 1] < vmax)) { /* (no such assertion in swpc_psv) */ snprintf(err, cap, "assert: background velocity exceeds the stability limit (m_vmodel_grd_rmed.f90:342-343)"); rc = -1; goto done; }
                            if (kk < r->kbeg_m || kk > r->kend_m) { snprintf(err, cap, "vmodel_grd_rmed: relative depth index out of the volume"); rc = -1; goto done; }
                            const float x = xi[r->ncell_m * (size_t)tbl[n - 1] + ora_idx3(r, kk, i, j)];
                            float vp2 = cv * vp1[n - 1] * (1.0f + x), vs2 = cv * vs1[n - 1] * (1.0f + x), rho2 = rho1[n - 1] * (1.0f + 0.8f * x);
                            if (vp1[n - 1] > 0 && vs1[n - 1] > 0) vcheck(&vp2, &vs2, &rho2, x, vmin, vmax, rhomin);
                            r->rho[m] = rho2;
                            r->lam[m] = rho2 * (vp2 * vp2 - 2 * vs2 * vs2);
                            r->mu[m] = rho2 * vs2 * vs2;
                        } else {
                            r->rho[m] = rho1[n - 1];
                            r->lam[m] = rho1[n - 1] * (cv * cv) * (vp1[n - 1] * vp1[n - 1] - 2 * vs1[n - 1] * vs1[n - 1]);
                            r->mu[m] = rho1[n - 1] * (cv * cv) * vs1[n - 1] * vs1[n - 1];
                        }
                        qp[m] = qp1[n - 1];
                        qs[m] = qs1[n - 1];
                    }
done:
    free(kgrd); free(glon); free(glat); free(Cv); free(xi); free(LR);
    (void)node_grd;
    return rc;
}
