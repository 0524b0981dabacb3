/*
 * oracle/ora_api.c -- accessors used by the tests (ctypes) and the SAC writer
 * (TEST INFRASTRUCTURE, see ora.h).
 *
 * SAC output restates m_wav.f90:273-395 (header fill), :782-792 (file name) and
 * src/shared/m_sac.f90:267-311 (wsac_s), :314-449 (sac__whdr), :452-562 (sac__init).
 */
#include "ora.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

float ora_rad2deg_s(float rad);

int ora_nranks(const ora_sim *s) { return s->nranks; }
const ora_cfg *ora_get_cfg(const ora_sim *s) { return &s->cfg; }

int ora_rank_int(const ora_sim *s, int rank, int what) {
    const ora_rank *r = &s->r[rank];
    switch (what) {
    case 0: return r->ibeg;
    case 1: return r->iend;
    case 2: return r->jbeg;
    case 3: return r->jend;
    case 4: return r->nxp;
    case 5: return r->nyp;
    case 6: return r->ibeg_k;
    case 7: return r->iend_k;
    case 8: return r->jbeg_k;
    case 9: return r->jend_k;
    case 10: return r->kbeg_k;
    case 11: return r->kend_k;
    case 12: return r->nsrc;
    case 13: return r->nst;
    case 14: return r->idx;
    case 15: return r->idy;
    case 16: return r->nzm;
    case 17: return r->nxm;
    case 18: return r->nym;
    case 19: return r->ibeg_m;
    case 20: return r->jbeg_m;
    case 21: return r->kbeg_m;
    default: return -999999;
    }
}

/* scalar config values for ctypes without mirroring the struct:
 * 0 vmin 1 vmax 2 fmax 3 fcut 4 M0 5 UC 6 zeta 7 d2 8 dt 9 xbeg 10 ybeg 11 zbeg 12 dx 13 dy 14 dz
 * 15 c (stability) 16 r (wavelength) 17..24 ts[0..7] 25.. c1 33.. c2 41.. d1 */
double ora_cfg_value(const ora_sim *s, int what) {
    const ora_cfg *c = &s->cfg;
    if (what >= 17 && what < 25) return c->ts[what - 17];
    if (what >= 25 && what < 33) return c->c1[what - 25];
    if (what >= 33 && what < 41) return c->c2[what - 33];
    if (what >= 41 && what < 49) return c->d1[what - 41];
    switch (what) {
    case 0: return c->vmin;
    case 1: return c->vmax;
    case 2: return c->fmax;
    case 3: return c->fcut;
    case 4: return c->M0;
    case 5: return c->UC;
    case 6: return c->zeta;
    case 7: return c->d2;
    case 8: return c->dt;
    case 9: return c->xbeg;
    case 10: return c->ybeg;
    case 11: return c->zbeg;
    case 12: return c->dx;
    case 13: return c->dy;
    case 14: return c->dz;
    case 15: { /* m_fdtool.f90:99-113 via m_report.f90:71 */
        float dt2;
        ora_fdm_stable_dt((float)c->dx, (float)c->dy, (float)c->dz, c->vmax, &dt2);
        return c->dt / dt2;
    }
    case 16: { /* m_fdtool.f90:116-134 */
        float dh = fmaxf(fmaxf((float)c->dx, (float)c->dy), (float)c->dz);
        float lambda_min = c->vmin / c->fmax;
        return lambda_min / dh;
    }
    default: return NAN;
    }
}
/* 0 nx 1 ny 2 nz 3 nt 4 na 5 nm 6 nproc_x 7 nproc_y 8 ntw 9 ntdec_w 10 ntdec_r 11 bf_mode */
int ora_cfg_int(const ora_sim *s, int what) {
    const ora_cfg *c = &s->cfg;
    switch (what) {
    case 0: return c->nx;
    case 1: return c->ny;
    case 2: return c->nz;
    case 3: return c->nt;
    case 4: return c->na;
    case 5: return c->nm;
    case 6: return c->nproc_x;
    case 7: return c->nproc_y;
    case 8: return c->ntw;
    case 9: return c->ntdec_w;
    case 10: return c->ntdec_r;
    case 11: return c->bf_mode;
    default: return -999999;
    }
}
const char *ora_cfg_str(const ora_sim *s, int what) {
    const ora_cfg *c = &s->cfg;
    switch (what) {
    case 0: return c->title;
    case 1: return c->odir;
    case 2: return c->abc_type;
    case 3: return c->stftype;
    case 4: return c->vmodel_type;
    default: return "";
    }
}
void ora_set_exedate(ora_sim *s, int exedate, int tz_minutes) {
    s->cfg.exedate = exedate;
    s->cfg.tz_minutes = tz_minutes;
}

static int field_ptr(const ora_rank *r, const char *name, ora_mp **mp, float **sp) {
    *mp = NULL;
    *sp = NULL;
    if (!strcmp(name, "Vx")) *mp = r->Vx;
    else if (!strcmp(name, "Vy")) *mp = r->Vy;
    else if (!strcmp(name, "Vz")) *mp = r->Vz;
    else if (!strcmp(name, "Sxx")) *mp = r->Sxx;
    else if (!strcmp(name, "Syy")) *mp = r->Syy;
    else if (!strcmp(name, "Szz")) *mp = r->Szz;
    else if (!strcmp(name, "Syz")) *mp = r->Syz;
    else if (!strcmp(name, "Sxz")) *mp = r->Sxz;
    else if (!strcmp(name, "Sxy")) *mp = r->Sxy;
    else if (!strcmp(name, "rho")) *sp = r->rho;
    else if (!strcmp(name, "lam")) *sp = r->lam;
    else if (!strcmp(name, "mu")) *sp = r->mu;
    else if (!strcmp(name, "taup")) *sp = r->taup;
    else if (!strcmp(name, "taus")) *sp = r->taus;
    else return -1;
    return 0;
}

int ora_get_field(const ora_sim *s, int rank, const char *name, double *out) {
    const ora_rank *r = &s->r[rank];
    ora_mp *mp;
    float *sp;
    if (field_ptr(r, name, &mp, &sp)) return -1;
    for (size_t n = 0; n < r->ncell_m; n++) out[n] = mp ? (double)mp[n] : (double)sp[n];
    return 0;
}

int ora_set_field(ora_sim *s, int rank, const char *name, const double *in) {
    ora_rank *r = &s->r[rank];
    ora_mp *mp;
    float *sp;
    if (field_ptr(r, name, &mp, &sp)) return -1;
    for (size_t n = 0; n < r->ncell_m; n++) {
        if (mp) mp[n] = (ora_mp)in[n];
        else sp[n] = (float)in[n];
    }
    return 0;
}

int ora_get_map(const ora_sim *s, int rank, const char *name, int *out) {
    const ora_rank *r = &s->r[rank];
    const int *p = NULL;
    if (!strcmp(name, "kfs")) p = r->kfs;
    else if (!strcmp(name, "kob")) p = r->kob;
    else if (!strcmp(name, "kfs_top")) p = r->kfs_top;
    else if (!strcmp(name, "kfs_bot")) p = r->kfs_bot;
    else if (!strcmp(name, "kob_top")) p = r->kob_top;
    else if (!strcmp(name, "kob_bot")) p = r->kob_bot;
    else if (!strcmp(name, "kbeg_a")) p = r->kbeg_a;
    else return -1;
    memcpy(out, p, sizeof(int) * (size_t)r->nxm * r->nym);
    return 0;
}

int ora_gather_field(const ora_sim *s, const char *name, double *out) {
    const ora_cfg *c = &s->cfg;
    for (int q = 0; q < s->nranks; q++) {
        const ora_rank *r = &s->r[q];
        ora_mp *mp;
        float *sp;
        if (field_ptr(r, name, &mp, &sp)) return -1;
        for (int j = r->jbeg; j <= r->jend; j++)
            for (int i = r->ibeg; i <= r->iend; i++)
                for (int k = 1; k <= c->nz; k++) {
                    size_t n = ora_idx3(r, k, i, j);
                    size_t g = (size_t)(k - 1) + (size_t)c->nz * ((size_t)(i - 1) + (size_t)c->nx * (size_t)(j - 1));
                    out[g] = mp ? (double)mp[n] : (double)sp[n];
                }
    }
    return 0;
}

int ora_get_sources(const ora_sim *s, int rank, int *ijk, double *mo) {
    const ora_rank *r = &s->r[rank];
    for (int i = 0; i < r->nsrc; i++) {
        ijk[3 * i] = r->isrc[i];
        ijk[3 * i + 1] = r->jsrc[i];
        ijk[3 * i + 2] = r->ksrc[i];
        if (mo) mo[i] = (double)r->mo[i];
    }
    return r->nsrc;
}

/* mij (6,nsrc): mxx myy mzz myz mxz mxy (bf_mode: fx fy fz 0 0 0); srcprm (2,nsrc) */
int ora_get_source_details(const ora_sim *s, int rank, double *mij, float *srcprm) {
    const ora_rank *r = &s->r[rank];
    for (int i = 0; i < r->nsrc; i++) {
        if (s->cfg.bf_mode) {
            mij[6 * i] = (double)r->fx[i]; mij[6 * i + 1] = (double)r->fy[i]; mij[6 * i + 2] = (double)r->fz[i];
            mij[6 * i + 3] = mij[6 * i + 4] = mij[6 * i + 5] = 0.0;
        } else {
            mij[6 * i] = (double)r->mxx[i]; mij[6 * i + 1] = (double)r->myy[i]; mij[6 * i + 2] = (double)r->mzz[i];
            mij[6 * i + 3] = (double)r->myz[i]; mij[6 * i + 4] = (double)r->mxz[i]; mij[6 * i + 5] = (double)r->mxy[i];
        }
        srcprm[2 * i] = r->srcprm[2 * i];
        srcprm[2 * i + 1] = r->srcprm[2 * i + 1];
    }
    return r->nsrc;
}

int ora_get_stations(const ora_sim *s, int rank, int *ijk, char *names) {
    const ora_rank *r = &s->r[rank];
    for (int i = 0; i < r->nst; i++) {
        ijk[3 * i] = r->ist[i];
        ijk[3 * i + 1] = r->jst[i];
        ijk[3 * i + 2] = r->kst[i];
        if (names) memcpy(names + 9 * i, r->stnm[i], 9);
    }
    return r->nst;
}

int ora_get_wav(const ora_sim *s, int rank, float *out) {
    const ora_rank *r = &s->r[rank];
    if (!r->wav_vel) return 0;
    memcpy(out, r->wav_vel, sizeof(float) * (size_t)s->cfg.ntw * 3 * r->nst);
    return r->nst;
}

int ora_get_wav_product(const ora_sim *s, int rank, int which, float *out) {
    const ora_rank *r = &s->r[rank];
    const float *src = which == 0 ? r->wav_vel : which == 1 ? r->wav_disp : which == 2 ? r->wav_stress : r->wav_strain;
    if (!src) return 0;
    memcpy(out, src, sizeof(float) * (size_t)s->cfg.ntw * (which < 2 ? 3 : 6) * r->nst);
    return r->nst;
}

int ora_get_profile(const ora_sim *s, int rank, const char *name, float *out) {
    const ora_rank *r = &s->r[rank];
    const float *p = NULL;
    int n = 0;
    if (!strcmp(name, "gxc")) { p = r->gxc; n = 4 * r->nxp; }
    else if (!strcmp(name, "gxe")) { p = r->gxe; n = 4 * r->nxp; }
    else if (!strcmp(name, "gyc")) { p = r->gyc; n = 4 * r->nyp; }
    else if (!strcmp(name, "gye")) { p = r->gye; n = 4 * r->nyp; }
    else if (!strcmp(name, "gzc")) { p = r->gzc; n = 4 * s->cfg.nz; }
    else if (!strcmp(name, "gze")) { p = r->gze; n = 4 * s->cfg.nz; }
    else if (!strcmp(name, "gx_c")) { p = r->gx_c; n = r->nxm; }
    else if (!strcmp(name, "gx_b")) { p = r->gx_b; n = r->nxm; }
    else if (!strcmp(name, "gy_c")) { p = r->gy_c; n = r->nym; }
    else if (!strcmp(name, "gy_b")) { p = r->gy_b; n = r->nym; }
    else if (!strcmp(name, "gz_c")) { p = r->gz_c; n = r->nzm; }
    else if (!strcmp(name, "gz_b")) { p = r->gz_b; n = r->nzm; }
    if (!p) return -1;
    memcpy(out, p, sizeof(float) * (size_t)n);
    return n;
}

/* ------------------------------------------------------------------------------------------ */
/* SAC                                                                                          */
typedef struct {
    float f[70];
    int32_t i[35];
    int32_t l[5];
    char a[192];
} sac_raw;

static void put8(char *dst, const char *src, int n) {
    int l = (int)strlen(src);
    for (int q = 0; q < n; q++) dst[q] = (q < l) ? src[q] : ' ';
}

static const char *ora_cmpnm(int prod, int cmp) {
    static const char *nm[4][6] = {{"Vx", "Vy", "Vz", "", "", ""}, {"Ux", "Uy", "Uz", "", "", ""},
                                   {"Sxx", "Syy", "Szz", "Syz", "Sxz", "Sxy"}, {"Exx", "Eyy", "Ezz", "Eyz", "Exz", "Exy"}};
    return nm[prod][cmp];
}

/* prod: 0 velocity, 1 displacement, 2 stress, 3 strain (m_wav.f90:280-334) */
static void sac_header(const ora_cfg *c, const ora_rank *r, int n, int prod, int cmp, sac_raw *h) {
    /* sac__whdr initial fill, m_sac.f90:330-337 */
    for (int q = 0; q < 70; q++) h->f[q] = -12345.0f;
    for (int q = 0; q < 35; q++) h->i[q] = -12345;
    for (int q = 0; q < 5; q++) h->l[q] = 0;
    for (int q = 0; q < 24; q++) put8(h->a + 8 * q, "-12345", 8);
    /* every k* field of sac__init is '-12345' (m_sac.f90:517-540); kevnm is 16 chars */
    put8(h->a + 8, "-12345", 16);

    /* header values, m_wav.f90:346-393 */
    double delta = (double)(c->ntdec_w * c->dt);
    h->f[0] = (float)((int)(delta * 1e7)) / 1e7f; /* m_sac.f90:339 */
    h->f[5] = c->tbeg;                             /* b */
    h->f[7] = c->otim;                             /* o */
    h->f[31] = r->stla[n];
    h->f[32] = r->stlo[n];
    h->f[34] = r->zst[n] * 1000;                   /* stdp [m] */
    h->f[35] = c->evla;
    h->f[36] = c->evlo;
    h->f[38] = c->evdp;
    h->f[39] = ora_moment_magnitude(c->M0);
    if (c->bf_mode) {
        h->f[40] = c->fx0; h->f[41] = c->fy0; h->f[42] = c->fz0;
    } else {
        h->f[40] = c->mxx0; h->f[41] = c->myy0; h->f[42] = c->mzz0;
        h->f[43] = c->myz0; h->f[44] = c->mxz0; h->f[45] = c->mxy0;
    }
    h->f[46] = c->clon; h->f[47] = c->clat; h->f[48] = c->phi;
    float ddx = c->sx0 - r->xst[n], ddy = c->sy0 - r->yst[n];
    h->f[50] = sqrtf(ddx * ddx + ddy * ddy);
    h->f[51] = ora_rad2deg_s(atan2f(r->yst[n] - c->sy0, r->xst[n] - c->sx0));
    h->f[52] = ora_rad2deg_s(atan2f(c->sy0 - r->yst[n], c->sx0 - r->xst[n]));
    if (prod < 2) {
        h->f[58] = 90.0f; /* cmpinc */
        h->f[57] = (cmp == 0) ? 0.0f + c->phi : (cmp == 1) ? 90.0f + c->phi : 0.0f; /* cmpaz m_wav.f90:286-288, :297-299 */
    }

    /* daytim__localtime(exedate) m_daytim.f90:246-314 : local time = utc + values(4) minutes */
    time_t tt = (time_t)c->exedate + (time_t)c->tz_minutes * 60;
    struct tm g;
    gmtime_r(&tt, &g);
    h->i[0] = g.tm_year + 1900;
    h->i[1] = g.tm_yday + 1;
    h->i[2] = g.tm_hour;
    h->i[3] = g.tm_min;
    h->i[4] = g.tm_sec;
    h->i[5] = 0;
    h->i[6] = 6;          /* nvhdr */
    h->i[9] = c->ntw;     /* npts */
    h->i[15] = 1;         /* iftype */
    h->i[16] = prod == 0 ? 7 : prod == 1 ? 6 : 5; /* idep m_wav.f90:290, :301, :316, :331 */
    /* ievtyp, iuser0-7 stay -12345 */
    h->l[0] = 1;          /* leven */
    h->l[1] = 0;          /* lpspol */
    h->l[2] = 1;          /* lovrok */
    h->l[3] = 0;          /* lcalda m_wav.f90:390 */
    h->l[4] = 0;          /* luser0 */

    put8(h->a + 0, r->stnm[n], 8);
    {
        char t16[17];
        const char *t = c->title;
        while (*t == ' ') t++;
        strncpy(t16, t, 16);
        t16[16] = 0;
        put8(h->a + 8, t16, 16);
    }
    put8(h->a + 8 * 20, ora_cmpnm(prod, cmp), 8); /* kcmpnm is words 151-152 -> 8-byte slot index 20 */
}

static void mkdir_p(const char *path) {
    char tmp[1024];
    snprintf(tmp, sizeof(tmp), "%s", path);
    for (char *p = tmp + 1; *p; p++)
        if (*p == '/') {
            *p = 0;
            mkdir(tmp, 0777);
            *p = '/';
        }
    mkdir(tmp, 0777);
}

int ora_write_sac(const ora_sim *s, const char *odir) {
    const ora_cfg *c = &s->cfg;
    if (!(c->sw_wav_v || c->sw_wav_u || c->sw_wav_stress || c->sw_wav_strain)) return 0;
    char dir[1024];
    snprintf(dir, sizeof(dir), "%s/wav", odir);
    mkdir_p(dir);
    int nfiles = 0;
    for (int q = 0; q < s->nranks; q++) {
        const ora_rank *r = &s->r[q];
        for (int n = 0; n < r->nst; n++)
            for (int prod = 0; prod < 4; prod++) { /* m_wav.f90:679-705 */
                const float *src = prod == 0 ? r->wav_vel : prod == 1 ? r->wav_disp : prod == 2 ? r->wav_stress : r->wav_strain;
                const int ncmp = prod < 2 ? 3 : 6;
                if (!src) continue;
                for (int cmp = 0; cmp < ncmp; cmp++) {
                    sac_raw h;
                    sac_header(c, r, n, prod, cmp, &h);
                    char fn[1400];
                    snprintf(fn, sizeof(fn), "%s/%s.3d.%s.%s.sac", dir, c->title, r->stnm[n], ora_cmpnm(prod, cmp));
                    FILE *fp = fopen(fn, "wb");
                    if (!fp) return -1;
                    fwrite(h.f, 4, 70, fp);
                    fwrite(h.i, 4, 35, fp);
                    fwrite(h.l, 4, 5, fp);
                    fwrite(h.a, 1, 192, fp);
                    fwrite(src + (size_t)c->ntw * ncmp * n + (size_t)c->ntw * cmp, 4, (size_t)c->ntw, fp);
                    fclose(fp);
                    nfiles++;
                }
            }
    }
    return nfiles;
}

/* ------------------------------------------------------------------------------------------ */
/* green__export, wav_format = 'sac' (m_green.f90:553-573); headers of green__setup :282-349 on sac__init defaults */
int ora_write_green_sac(const ora_sim *s, const char *odir) {
    const ora_green *g = s->green;
    if (!g) return 0;
    const ora_cfg *c = &s->cfg;
    static const char *cmpn[9] = {"mxx", "myy", "mzz", "myz", "mxz", "mxy", "fx_", "fy_", "fz_"};
    char dir[1200];
    snprintf(dir, sizeof(dir), "%s/green/%s", odir, g->stnm);
    mkdir_p(dir);
    int nfiles = 0;
    const float dx = (float)c->dx, dy = (float)c->dy, dz = (float)c->dz;
    for (int q = 0; q < s->nranks; q++) {
        const ora_green_rank *gr = &g->r[q];
        for (int i = 0; i < gr->ng; i++)
            for (int j = 0; j < g->ncmp; j++) {
                sac_raw h;
                for (int w = 0; w < 70; w++) h.f[w] = -12345.0f;
                for (int w = 0; w < 35; w++) h.i[w] = -12345;
                for (int w = 0; w < 24; w++) put8(h.a + 8 * w, "-12345", 8);
                put8(h.a + 8, "-12345", 16);
                const double delta = (double)(g->ntdec_w * c->dt);
                h.f[0] = (float)((int)(delta * 1e7)) / 1e7f;
                h.f[5] = c->tbeg;
                h.f[31] = g->evla0; h.f[32] = g->evlo0; h.f[34] = g->zsrc * 1000;   /* station = pseudo source */
                h.f[35] = gr->lat[i]; h.f[36] = gr->lon[i]; h.f[38] = gr->zg[i];
                h.f[40] = gr->xg[i]; h.f[41] = gr->yg[i]; h.f[42] = gr->zg[i];
                h.f[43] = ora_i2x(gr->ig[i], c->xbeg, dx); h.f[44] = ora_i2x(gr->jg[i], c->ybeg, dy); h.f[45] = ora_i2x(gr->kg[i], c->zbeg, dz);
                h.f[46] = c->clon; h.f[47] = c->clat; h.f[48] = c->phi;
                char kc[9];
                if (g->cmp == 'x') { snprintf(kc, sizeof(kc), "G_Vx_%s", cmpn[j]); h.f[58] = 90.0f; h.f[57] = 0.0f + c->phi; }
                else if (g->cmp == 'y') { snprintf(kc, sizeof(kc), "G_Vy_%s", cmpn[j]); h.f[58] = 90.0f; h.f[57] = 90.0f + c->phi; }
                else { snprintf(kc, sizeof(kc), "G_Vz_%s", cmpn[j]); h.f[58] = 0.0f; h.f[57] = 0.0f; }
                time_t tt = (time_t)c->exedate + (time_t)c->tz_minutes * 60;
                struct tm gm;
                gmtime_r(&tt, &gm);
                h.i[0] = gm.tm_year + 1900; h.i[1] = gm.tm_yday + 1; h.i[2] = gm.tm_hour; h.i[3] = gm.tm_min; h.i[4] = gm.tm_sec; h.i[5] = 0;
                h.i[6] = 6; h.i[9] = g->ntw; h.i[15] = 1; h.i[16] = j < 6 ? 7 : 6;
                h.l[0] = 1; h.l[1] = 0; h.l[2] = 1; h.l[3] = 0; h.l[4] = 0;
                char cid8[16];
                snprintf(cid8, sizeof(cid8), "%08d", gr->gid[i]);
                put8(h.a + 0, g->stnm, 8);
                put8(h.a + 8, cid8, 16);
                put8(h.a + 8 * 20, kc, 8);
                char fn[1600];
                snprintf(fn, sizeof(fn), "%s/%s__%s__%s__%c__%s__.sac", dir, c->title, cid8, g->stnm, g->cmp, cmpn[j]);
                FILE *fp = fopen(fn, "wb");
                if (!fp) return -1;
                fwrite(h.f, 4, 70, fp); fwrite(h.i, 4, 35, fp); fwrite(h.l, 4, 5, fp); fwrite(h.a, 1, 192, fp);
                const float *src = gr->gf + (size_t)g->ntw * (g->ncmp * (size_t)i + j);
                for (int t = 0; t < g->ntw; t++) {   /* positive upward for the z component :563-565 */
                    const float v = g->cmp == 'z' ? -src[t] : src[t];
                    fwrite(&v, 4, 1, fp);
                }
                fclose(fp);
                nfiles++;
            }
    }
    return nfiles;
}
