/*
 * oracle/ora_green.c -- TEST INFRASTRUCTURE (see ora.h): reciprocal Green's-function mode of swpc_3d,
 * src/swpc_3d/m_green.f90.
 *
 *   green__setup   :67-355   pseudo source = a station (wav__stquery m_wav.f90:627-656), list of Green's-function grid
 *                            points (xyz / llz), ownership ibeg<=ii<=iend & jbeg<=jj<=jend & kob(ii,jj)<=kk<=kend
 *   green__store   :357-551  nine displacement-gradient sums (4th-order differences averaged onto the normal-stress node)
 *                            every step, + three displacement sums with green_bforce; sampled every ntdec_w steps
 *   green__source  :606-649  body force at the pseudo source, stf at tbeg + it*dt
 *   green__export  :553-604  (SAC files: ora_api.c)
 *
 * In this mode source__setup returns before fcut / fmax are computed (m_source.f90:84-91) while absorb__setup, which uses
 * fcut, runs before green__setup (main.f90:74-78): the reference's fcut is then the module variable's initial storage.
 * It is taken as 0 here (static storage of gfortran / nvfortran).
 */
#include "ora.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static int blank(const char *s) {
    while (*s) {
        if (*s != ' ' && *s != '\t' && *s != '\r' && *s != '\n') return 0;
        s++;
    }
    return 1;
}

void ora_green_free(ora_sim *s) {
    ora_green *g = s->green;
    if (!g) return;
    for (int q = 0; q < s->nranks; q++) {
        ora_green_rank *r = &g->r[q];
        free(r->ig); free(r->jg); free(r->kg); free(r->gid); free(r->xg); free(r->yg); free(r->zg); free(r->lon); free(r->lat);
        free(r->gf); free(r->acc);
    }
    free(g->r);
    free(g);
    s->green = NULL;
}

int ora_green_setup(ora_sim *s, const ora_ini *ini, const char *base, char *err, size_t cap) {
    ora_cfg *c = &s->cfg;
    if (c->benchmark_mode) { c->green_mode = 0; return 0; }
    ora_readini_l(ini, "green_mode", &c->green_mode, 0);
    if (!c->green_mode) return 0;
    ora_green *g = (ora_green *)calloc(1, sizeof(ora_green));
    g->r = (ora_green_rank *)calloc((size_t)s->nranks, sizeof(ora_green_rank));
    s->green = g;
    const ora_mp C40 = (ora_mp)9.0 / (ora_mp)8.0, C41 = (ora_mp)1.0 / (ora_mp)24.0;
    g->r40x = C40 / (ora_mp)c->dx; g->r40y = C40 / (ora_mp)c->dy; g->r40z = C40 / (ora_mp)c->dz;
    g->r41x = C41 / (ora_mp)c->dx; g->r41y = C41 / (ora_mp)c->dy; g->r41z = C41 / (ora_mp)c->dz;
    char tmp[ORA_STRLEN], fn_glst[ORA_STRLEN], fmt[ORA_STRLEN];
    ora_readini_c(ini, "green_stnm", tmp, "");
    snprintf(g->stnm, sizeof(g->stnm), "%.8s", tmp);
    ora_readini_c(ini, "green_cmp", tmp, "");
    g->cmp = tmp[0];
    ora_readini_s(ini, "green_trise", &g->trise, 1.0f);
    ora_readini_l(ini, "green_bforce", &g->bforce, 0);
    ora_readini_s(ini, "green_maxdist", &g->maxdist, 1e30f);
    if (!(g->maxdist > 0.0f)) { snprintf(err, cap, "assert: green_maxdist > 0 (m_green.f90:124)"); return -1; }
    c->M0 = 1;                      /* :126 */
    c->fmax = 2.0f / g->trise;      /* :127 */
    ora_readini_c(ini, "fn_glst", fn_glst, "");
    ora_readini_c(ini, "green_fmt", fmt, "xyz");
    if (strcmp(fmt, "xyz") && strcmp(fmt, "llz")) { snprintf(err, cap, "assert: green_fmt is 'xyz' or 'llz' (m_green.f90:135)"); return -1; }
    ora_readini_i(ini, "ntdec_w", &g->ntdec_w, 10);
    ora_readini_c(ini, "stftype", tmp, "kupper");
    snprintf(g->stftype, sizeof(g->stftype), "%.15s", tmp);
    if (!strcmp(g->stftype, "scosine")) strcpy(g->stftype, "cosine");
    ora_readini_c(ini, "wav_format", tmp, "sac");
    snprintf(g->wav_format, sizeof(g->wav_format), "%.15s", tmp);
    switch (g->cmp) {
    case 'x': g->fx1 = 1.0f; break;
    case 'y': g->fy1 = 1.0f; break;
    case 'z': g->fz1 = 1.0f; break;
    default: snprintf(err, cap, "no matching green_cmp (m_green.f90:151-153)"); return -1;
    }
    g->dt_dxyz = (float)((double)c->dt / (c->dx * c->dy * c->dz));   /* real(dt / (dx*dy*dz)) :156 */
    /* wav__stquery on every rank + broadcast from the owner :161-183 */
    int found = 0;
    for (int q = 0; q < s->nranks && !found; q++) {
        const ora_rank *r = &s->r[q];
        for (int n = 0; n < r->nst; n++)
            if (!strcmp(g->stnm, r->stnm[n])) {
                g->isrc = r->ist[n]; g->jsrc = r->jst[n]; g->ksrc = r->kst[n];
                g->xsrc = r->xst[n]; g->ysrc = r->yst[n]; g->zsrc = r->zst[n]; g->evlo0 = r->stlo[n]; g->evla0 = r->stla[n];
                found = 1;
                break;
            }
    }
    if (!found) { snprintf(err, cap, "assert: station '%s' (green_stnm) is not inside the model (m_green.f90:165)", g->stnm); return -1; }
    g->ntw = (int)floorf((float)(c->nt - 1) / (float)g->ntdec_w + 1.0f);
    g->ncmp = g->bforce ? 9 : 6;
    for (int q = 0; q < s->nranks; q++) {
        const ora_rank *r = &s->r[q];
        g->r[q].is_src = (r->ibeg <= g->isrc && g->isrc <= r->iend + 1) && (r->jbeg <= g->jsrc && g->jsrc <= r->jend + 1) &&
                         (r->kbeg <= g->ksrc && g->ksrc <= r->kend);
    }
    /* list file :188-280 */
    char path[2 * ORA_STRLEN + 2];
    if (fn_glst[0] == '/' || !base || !base[0]) snprintf(path, sizeof(path), "%s", fn_glst);
    else snprintf(path, sizeof(path), "%s/%s", base, fn_glst);
    FILE *fp = fopen(path, "r");
    if (!fp) { snprintf(err, cap, "assert: fn_glst %s exists (m_green.f90:131-132)", path); return -1; }
    char line[1024];
    const float dx = (float)c->dx, dy = (float)c->dy, dz = (float)c->dz;
    while (fgets(line, sizeof(line), fp)) {
        char *p = line;
        while (*p == ' ' || *p == '\t') p++;
        if (*p == '#' || blank(p)) continue;
        for (char *t = p; *t; t++) if (*t == ',') *t = ' ';
        float a, b, zg0, xg0, yg0, lo, la;
        int gid0;
        if (sscanf(p, "%f %f %f %d", &a, &b, &zg0, &gid0) < 4) { fclose(fp); snprintf(err, cap, "assert: bad line in fn_glst"); return -1; }
        if (!strcmp(fmt, "xyz")) { xg0 = a; yg0 = b; ora_geomap_c2g(xg0, yg0, c->clon, c->clat, c->phi, &lo, &la); }
        else { lo = a; la = b; ora_geomap_g2c(lo, la, c->clon, c->clat, c->phi, &xg0, &yg0); }
        if (!(0 <= gid0 && gid0 <= 99999999)) { fclose(fp); snprintf(err, cap, "assert: 0 <= gid <= 99999999"); return -1; }
        const float ddx = xg0 - g->xsrc, ddy = yg0 - g->ysrc;
        const float dd = sqrtf(ddx * ddx + ddy * ddy);
        if (dd > g->maxdist) continue;
        const int ii = ora_x2i(xg0, c->xbeg, dx), jj = ora_x2i(yg0, c->ybeg, dy), kk = ora_x2i(zg0, c->zbeg, dz);
        for (int q = 0; q < s->nranks; q++) {
            const ora_rank *r = &s->r[q];
            if (!(r->ibeg <= ii && ii <= r->iend && r->jbeg <= jj && jj <= r->jend)) continue;
            if (!(r->kob[ora_idx2(r, ii, jj)] <= kk && kk <= r->kend)) continue;   /* only the solid part */
            ora_green_rank *gr = &g->r[q];
            const int n = gr->ng++;
#define GROW(ptr, T) gr->ptr = (T *)realloc(gr->ptr, sizeof(T) * (size_t)gr->ng)
            GROW(ig, int); GROW(jg, int); GROW(kg, int); GROW(gid, int);
            GROW(xg, float); GROW(yg, float); GROW(zg, float); GROW(lon, float); GROW(lat, float);
#undef GROW
            gr->ig[n] = ii; gr->jg[n] = jj; gr->kg[n] = kk; gr->gid[n] = gid0;
            gr->xg[n] = xg0; gr->yg[n] = yg0; gr->zg[n] = zg0; gr->lon[n] = lo; gr->lat[n] = la;
        }
    }
    fclose(fp);
    for (int q = 0; q < s->nranks; q++) {
        ora_green_rank *gr = &g->r[q];
        gr->gf = (float *)calloc((size_t)g->ntw * g->ncmp * (gr->ng > 0 ? gr->ng : 1), sizeof(float));
        gr->acc = (float *)calloc((size_t)12 * (gr->ng > 0 ? gr->ng : 1), sizeof(float));
    }
    return 0;
}

/* m_green.f90:357-551 */
void ora_green_store(ora_sim *s, int it) {
    ora_green *g = s->green;
    if (!g) return;
    const ora_cfg *c = &s->cfg;
    const float dt = c->dt;
    const float UC_BF = 1e-12f, UC_DERIV = 1e-15f;   /* 10.0**(-12), 10.0**(-15) :30-31 */
    const ora_mp r40x = g->r40x, r40y = g->r40y, r40z = g->r40z, r41x = g->r41x, r41y = g->r41y, r41z = g->r41z;
    for (int q = 0; q < s->nranks; q++) {
        const ora_rank *r = &s->r[q];
        ora_green_rank *gr = &g->r[q];
        for (int i = 0; i < gr->ng; i++) {
            const int ii = gr->ig[i], jj = gr->jg[i], kk = gr->kg[i];
#define VX(k, i_, j) r->Vx[ora_idx3(r, k, i_, j)]
#define VY(k, i_, j) r->Vy[ora_idx3(r, k, i_, j)]
#define VZ(k, i_, j) r->Vz[ora_idx3(r, k, i_, j)]
            const ora_mp dxVx = (VX(kk, ii, jj) - VX(kk, ii - 1, jj)) * r40x - (VX(kk, ii + 1, jj) - VX(kk, ii - 2, jj)) * r41x;
            const ora_mp dyVy = (VY(kk, ii, jj) - VY(kk, ii, jj - 1)) * r40y - (VY(kk, ii, jj + 1) - VY(kk, ii, jj - 2)) * r41y;
            const ora_mp dzVz = (VZ(kk, ii, jj) - VZ(kk - 1, ii, jj)) * r40z - (VZ(kk + 1, ii, jj) - VZ(kk - 2, ii, jj)) * r41z;
            const ora_mp dxVy1 = (VY(kk, ii + 1, jj) - VY(kk, ii, jj)) * r40x - (VY(kk, ii + 2, jj) - VY(kk, ii - 1, jj)) * r41x;
            const ora_mp dxVy2 = (VY(kk, ii + 1, jj - 1) - VY(kk, ii, jj - 1)) * r40x - (VY(kk, ii + 2, jj - 1) - VY(kk, ii - 1, jj - 1)) * r41x;
            const ora_mp dxVy3 = (VY(kk, ii, jj) - VY(kk, ii - 1, jj)) * r40x - (VY(kk, ii + 1, jj) - VY(kk, ii - 2, jj)) * r41x;
            const ora_mp dxVy4 = (VY(kk, ii, jj - 1) - VY(kk, ii - 1, jj - 1)) * r40x - (VY(kk, ii + 1, jj - 1) - VY(kk, ii - 2, jj - 1)) * r41x;
            const ora_mp dxVz1 = (VZ(kk, ii + 1, jj) - VZ(kk, ii, jj)) * r40x - (VZ(kk, ii + 2, jj) - VZ(kk, ii - 1, jj)) * r41x;
            const ora_mp dxVz2 = (VZ(kk - 1, ii + 1, jj) - VZ(kk - 1, ii, jj)) * r40x - (VZ(kk - 1, ii + 2, jj) - VZ(kk - 1, ii - 1, jj)) * r41x;
            const ora_mp dxVz3 = (VZ(kk, ii, jj) - VZ(kk, ii - 1, jj)) * r40x - (VZ(kk, ii + 1, jj) - VZ(kk, ii - 2, jj)) * r41x;
            const ora_mp dxVz4 = (VZ(kk - 1, ii, jj) - VZ(kk - 1, ii - 1, jj)) * r40x - (VZ(kk - 1, ii + 1, jj) - VZ(kk - 1, ii - 2, jj)) * r41x;
            const ora_mp dyVx1 = (VX(kk, ii, jj + 1) - VX(kk, ii, jj)) * r40y - (VX(kk, ii, jj + 2) - VX(kk, ii, jj - 1)) * r41y;
            const ora_mp dyVx2 = (VX(kk, ii - 1, jj + 1) - VX(kk, ii - 1, jj)) * r40y - (VX(kk, ii - 1, jj + 2) - VX(kk, ii - 1, jj - 1)) * r41y;
            const ora_mp dyVx3 = (VX(kk, ii, jj) - VX(kk, ii, jj - 1)) * r40y - (VX(kk, ii, jj + 1) - VX(kk, ii, jj - 2)) * r41y;
            const ora_mp dyVx4 = (VX(kk, ii - 1, jj) - VX(kk, ii - 1, jj - 1)) * r40y - (VX(kk, ii - 1, jj + 1) - VX(kk, ii - 1, jj - 2)) * r41y;
            const ora_mp dyVz1 = (VZ(kk, ii, jj + 1) - VZ(kk, ii, jj)) * r40y - (VZ(kk, ii, jj + 2) - VZ(kk, ii, jj - 1)) * r41y;
            const ora_mp dyVz2 = (VZ(kk - 1, ii, jj + 1) - VZ(kk - 1, ii, jj)) * r40y - (VZ(kk - 1, ii, jj + 2) - VZ(kk - 1, ii, jj - 1)) * r41y;
            const ora_mp dyVz3 = (VZ(kk, ii, jj) - VZ(kk, ii, jj - 1)) * r40y - (VZ(kk, ii, jj + 1) - VZ(kk, ii, jj - 2)) * r41y;
            const ora_mp dyVz4 = (VZ(kk - 1, ii, jj) - VZ(kk - 1, ii, jj - 1)) * r40y - (VZ(kk - 1, ii, jj + 1) - VZ(kk - 1, ii, jj - 2)) * r41y;
            const ora_mp dzVx1 = (VX(kk + 1, ii, jj) - VX(kk, ii, jj)) * r40z - (VX(kk + 2, ii, jj) - VX(kk - 1, ii, jj)) * r41z;
            const ora_mp dzVx2 = (VX(kk + 1, ii - 1, jj) - VX(kk, ii - 1, jj)) * r40z - (VX(kk + 2, ii - 1, jj) - VX(kk - 1, ii - 1, jj)) * r41z;
            const ora_mp dzVx3 = (VX(kk, ii, jj) - VX(kk - 1, ii, jj)) * r40z - (VX(kk + 1, ii, jj) - VX(kk - 2, ii, jj)) * r41z;
            const ora_mp dzVx4 = (VX(kk, ii - 1, jj) - VX(kk - 1, ii - 1, jj)) * r40z - (VX(kk + 1, ii - 1, jj) - VX(kk - 2, ii - 1, jj)) * r41z;
            const ora_mp dzVy1 = (VY(kk + 1, ii, jj) - VY(kk, ii, jj)) * r40z - (VY(kk + 2, ii, jj) - VY(kk - 1, ii, jj)) * r41z;
            const ora_mp dzVy2 = (VY(kk + 1, ii, jj - 1) - VY(kk, ii, jj - 1)) * r40z - (VY(kk + 2, ii, jj - 1) - VY(kk - 1, ii, jj - 1)) * r41z;
            const ora_mp dzVy3 = (VY(kk, ii, jj) - VY(kk - 1, ii, jj)) * r40z - (VY(kk + 1, ii, jj) - VY(kk - 2, ii, jj)) * r41z;
            const ora_mp dzVy4 = (VY(kk, ii, jj - 1) - VY(kk - 1, ii, jj - 1)) * r40z - (VY(kk + 1, ii, jj - 1) - VY(kk - 2, ii, jj - 1)) * r41z;
            /* `* 0.25`: default-real constant times real(MP) */
            const ora_mp dxVy = (dxVy1 + dxVy2 + dxVy3 + dxVy4) * 0.25f;
            const ora_mp dxVz = (dxVz1 + dxVz2 + dxVz3 + dxVz4) * 0.25f;
            const ora_mp dyVx = (dyVx1 + dyVx2 + dyVx3 + dyVx4) * 0.25f;
            const ora_mp dyVz = (dyVz1 + dyVz2 + dyVz3 + dyVz4) * 0.25f;
            const ora_mp dzVx = (dzVx1 + dzVx2 + dzVx3 + dzVx4) * 0.25f;
            const ora_mp dzVy = (dzVy1 + dzVy2 + dzVy3 + dzVy4) * 0.25f;
            float *a = gr->acc + 12 * (size_t)i;
            a[0] = a[0] + (float)(dxVx * dt); a[1] = a[1] + (float)(dxVy * dt); a[2] = a[2] + (float)(dxVz * dt);
            a[3] = a[3] + (float)(dyVx * dt); a[4] = a[4] + (float)(dyVy * dt); a[5] = a[5] + (float)(dyVz * dt);
            a[6] = a[6] + (float)(dzVx * dt); a[7] = a[7] + (float)(dzVy * dt); a[8] = a[8] + (float)(dzVz * dt);
            if (g->bforce) {
                a[9] = a[9] + 0.5f * (float)(VX(kk, ii, jj) + VX(kk, ii - 1, jj)) * dt;
                a[10] = a[10] + 0.5f * (float)(VY(kk, ii, jj) + VY(kk, ii, jj - 1)) * dt;
                a[11] = a[11] + 0.5f * (float)(VZ(kk, ii, jj) + VZ(kk - 1, ii, jj)) * dt;
            }
#undef VX
#undef VY
#undef VZ
        }
        if ((it - 1) % g->ntdec_w == 0) {
            const int itw = (it - 1) / g->ntdec_w + 1;
            if (itw > g->ntw) continue;
            for (int i = 0; i < gr->ng; i++) {
                const float *a = gr->acc + 12 * (size_t)i;
                float *o = gr->gf + (size_t)g->ntw * g->ncmp * i + (itw - 1);   /* gf(itw, (i-1)*ncmp + j) */
                const size_t nt_ = (size_t)g->ntw;
                o[0 * nt_] = a[0] * UC_DERIV * 1e9f;
                o[1 * nt_] = a[4] * UC_DERIV * 1e9f;
                o[2 * nt_] = a[8] * UC_DERIV * 1e9f;
                o[3 * nt_] = (a[5] + a[7]) * UC_DERIV * 1e9f;
                o[4 * nt_] = (a[2] + a[6]) * UC_DERIV * 1e9f;
                o[5 * nt_] = (a[1] + a[3]) * UC_DERIV * 1e9f;
                if (g->bforce) {
                    o[6 * nt_] = a[9] * UC_BF * 1e9f;
                    o[7 * nt_] = a[10] * UC_BF * 1e9f;
                    o[8 * nt_] = a[11] * UC_BF * 1e9f;
                }
            }
        }
    }
}

/* m_green.f90:606-649 */
void ora_green_source(ora_sim *s, int it) {
    ora_green *g = s->green;
    if (!g) return;
    const ora_cfg *c = &s->cfg;
    const float prm[2] = {0.0f /* green_tbeg */, g->trise};
    const float stf = ora_momentrate(c->tbeg + it * c->dt, g->stftype, prm);
    const float fx = g->fx1 * g->dt_dxyz * stf, fy = g->fy1 * g->dt_dxyz * stf, fz = g->fz1 * g->dt_dxyz * stf;
    const int k = g->ksrc, i = g->isrc, j = g->jsrc;
    for (int q = 0; q < s->nranks; q++) {
        if (!g->r[q].is_src) continue;
        ora_rank *r = &s->r[q];
#define RHO(k_, i_, j_) r->rho[ora_idx3(r, k_, i_, j_)]
        r->Vx[ora_idx3(r, k, i, j)] = r->Vx[ora_idx3(r, k, i, j)] + (2.0f / (RHO(k, i, j) + RHO(k, i + 1, j))) * fx / 2;
        r->Vx[ora_idx3(r, k, i - 1, j)] = r->Vx[ora_idx3(r, k, i - 1, j)] + (2.0f / (RHO(k, i, j) + RHO(k, i - 1, j))) * fx / 2;
        r->Vy[ora_idx3(r, k, i, j)] = r->Vy[ora_idx3(r, k, i, j)] + (2.0f / (RHO(k, i, j) + RHO(k, i, j + 1))) * fy / 2;
        r->Vy[ora_idx3(r, k, i, j - 1)] = r->Vy[ora_idx3(r, k, i, j - 1)] + (2.0f / (RHO(k, i, j) + RHO(k, i, j - 1))) * fy / 2;
        r->Vz[ora_idx3(r, k, i, j)] = r->Vz[ora_idx3(r, k, i, j)] + (2.0f / (RHO(k, i, j) + RHO(k + 1, i, j))) * fz / 2;
        r->Vz[ora_idx3(r, k - 1, i, j)] = r->Vz[ora_idx3(r, k - 1, i, j)] + (2.0f / (RHO(k, i, j) + RHO(k - 1, i, j))) * fz / 2;
#undef RHO
    }
}

int ora_green_int(const ora_sim *s, int rank, int what) {
    const ora_green *g = s->green;
    if (!g || rank < 0 || rank >= s->nranks) return -1;
    switch (what) {
    case 0: return g->r[rank].ng;
    case 1: return g->ncmp;
    case 2: return g->isrc;
    case 3: return g->jsrc;
    case 4: return g->ksrc;
    case 5: return g->r[rank].is_src;
    case 6: return g->ntw;
    }
    return -1;
}

int ora_green_points(const ora_sim *s, int rank, int *ijk, int *gid) {
    const ora_green *g = s->green;
    if (!g) return -1;
    const ora_green_rank *gr = &g->r[rank];
    for (int i = 0; i < gr->ng; i++) {
        ijk[3 * i] = gr->ig[i]; ijk[3 * i + 1] = gr->jg[i]; ijk[3 * i + 2] = gr->kg[i];
        gid[i] = gr->gid[i];
    }
    return gr->ng;
}

int ora_get_green(const ora_sim *s, int rank, float *out) {
    const ora_green *g = s->green;
    if (!g) return -1;
    const ora_green_rank *gr = &g->r[rank];
    memcpy(out, gr->gf, sizeof(float) * (size_t)g->ntw * g->ncmp * gr->ng);
    return gr->ng;
}
