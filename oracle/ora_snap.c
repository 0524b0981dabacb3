/*
 * oracle/ora_snap.c -- restatement of the snapshot products of src/swpc_3d/m_snap.f90 (TEST INFRASTRUCTURE, see ora.h).
 *
 * 15 products = sections xy / xz / yz / fs / ob  x  types ps / v / u; product id = section*3 + type with section
 * 0 xy, 1 xz, 2 yz, 3 fs, 4 ob and type 0 ps, 1 v, 2 u.  The MPI sum-reduce to the I/O rank (m_snap.f90:1064) is a
 * plain fill of the global (n1, n2, nvar) buffer: every rank contributes only its own disjoint part, zeros elsewhere.
 * Records are kept in memory (one per output step) instead of being written to netCDF: the file format is checked
 * on the product side with an independent reader (scipy.io.netcdf_file).
 */
#include "ora.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FLT_EPS 1.1920929e-07f

typedef struct {
    int on;
    int n1, n2, nvar;
    float *buf;      /* current slice / displacement accumulator (n1,n2,nvar) */
    float *maxv;     /* (n1,n2,3) for fs/ob v and u */
    int nrec, cap;
    float **rec;
    int *rec_it;
} ora_prod;

typedef struct ora_snapstate {
    int idec, jdec, kdec, ntdec_s, nxs, nys, nzs, k0_xy, i0_yz, j0_xz;
    float z0_xy, x0_yz, y0_xz;
    float *xsnp, *ysnp, *zsnp;
    ora_prod p[15];
} ora_snapstate;

static ora_snapstate *snap_of(const ora_sim *s) { return (ora_snapstate *)s->snap; }

static int ceil_div_f(int a, int d) { return (int)ceilf((float)a / (float)d); }
static int floor_div_f(int a, int d) { return (int)floorf((float)a / (float)d); }

/* m_snap.f90:95-322 (keys :105-127, sizes :122-125, coordinates :128-140, positions :152-154) */
int ora_snap_setup(ora_sim *s, const ora_ini *ini) {
    const ora_cfg *c = &s->cfg;
    ora_snapstate *st = (ora_snapstate *)calloc(1, sizeof(ora_snapstate));
    s->snap = st;
    static const char *keys[15] = {"xy_ps%sw", "xy_v%sw", "xy_u%sw", "xz_ps%sw", "xz_v%sw", "xz_u%sw", "yz_ps%sw", "yz_v%sw", "yz_u%sw",
                                   "fs_ps%sw", "fs_v%sw", "fs_u%sw", "ob_ps%sw", "ob_v%sw", "ob_u%sw"};
    int any = 0;
    for (int q = 0; q < 15; q++) { ora_readini_l(ini, keys[q], &st->p[q].on, 0); any |= st->p[q].on; }
    float zd = fminf(10.0f, c->zend); zd = fmaxf(zd, c->zbeg);
    float xd = fminf(0.0f, c->xend); xd = fmaxf(xd, c->xbeg);
    float yd = fminf(0.0f, c->yend); yd = fmaxf(yd, c->ybeg);
    ora_readini_s(ini, "z0_xy", &st->z0_xy, zd);
    ora_readini_s(ini, "x0_yz", &st->x0_yz, xd);
    ora_readini_s(ini, "y0_xz", &st->y0_xz, yd);
    ora_readini_i(ini, "idec", &st->idec, 1);
    ora_readini_i(ini, "jdec", &st->jdec, 1);
    ora_readini_i(ini, "kdec", &st->kdec, 1);
    ora_readini_i(ini, "ntdec_s", &st->ntdec_s, 10);
    st->nxs = (c->nx + (st->idec / 2)) / st->idec;
    st->nys = (c->ny + (st->jdec / 2)) / st->jdec;
    st->nzs = (c->nz + (st->kdec / 2)) / st->kdec;
    st->xsnp = (float *)calloc((size_t)st->nxs, sizeof(float));
    st->ysnp = (float *)calloc((size_t)st->nys, sizeof(float));
    st->zsnp = (float *)calloc((size_t)st->nzs, sizeof(float));
    for (int i = 1; i <= st->nxs; i++) st->xsnp[i - 1] = ora_i2x(i * st->idec - (st->idec / 2), c->xbeg, (float)c->dx);
    for (int j = 1; j <= st->nys; j++) st->ysnp[j - 1] = ora_i2x(j * st->jdec - (st->jdec / 2), c->ybeg, (float)c->dy);
    for (int k = 1; k <= st->nzs; k++) st->zsnp[k - 1] = ora_i2x(k * st->kdec - (st->kdec / 2), c->zbeg, (float)c->dz);
    st->k0_xy = ora_x2i(st->z0_xy, c->zbeg, (float)c->dz);
    st->i0_yz = ora_x2i(st->x0_yz, c->xbeg, (float)c->dx);
    st->j0_xz = ora_x2i(st->y0_xz, c->ybeg, (float)c->dy);
    for (int q = 0; q < 15; q++) {
        ora_prod *p = &st->p[q];
        int sec = q / 3, typ = q % 3;
        p->n1 = sec == 2 ? st->nys : st->nxs;
        p->n2 = (sec == 1 || sec == 2) ? st->nzs : st->nys;
        p->nvar = typ == 0 ? 4 : 3;
        if (!p->on) continue;
        p->buf = (float *)calloc((size_t)p->n1 * p->n2 * p->nvar, sizeof(float));
        if ((sec == 3 || sec == 4) && typ != 0) p->maxv = (float *)calloc((size_t)p->n1 * p->n2 * 3, sizeof(float));
    }
    return any;
}

static void rank_ranges(const ora_snapstate *st, const ora_rank *r, int *is0, int *is1, int *js0, int *js1, int *ks0, int *ks1) {
    /* m_snap.f90:143-149 */
    *is0 = ceil_div_f(r->ibeg + st->idec / 2, st->idec); *is1 = floor_div_f(r->iend + st->idec / 2, st->idec);
    *js0 = ceil_div_f(r->jbeg + st->jdec / 2, st->jdec); *js1 = floor_div_f(r->jend + st->jdec / 2, st->jdec);
    *ks0 = ceil_div_f(r->kbeg + st->kdec / 2, st->kdec); *ks1 = floor_div_f(r->kend + st->kdec / 2, st->kdec);
}

/* evaluate / accumulate one product over all ranks into the global buffer */
static void eval_product(ora_sim *s, int q) {
    const ora_cfg *c = &s->cfg;
    ora_snapstate *st = snap_of(s);
    ora_prod *p = &st->p[q];
    const int sec = q / 3, typ = q % 3;
    const size_t plane = (size_t)p->n1 * p->n2;
    const ora_mp r20x = (ora_mp)1.0 / (ora_mp)c->dx, r20y = (ora_mp)1.0 / (ora_mp)c->dy, r20z = (ora_mp)1.0 / (ora_mp)c->dz; /* :281-283 */
    for (int rk = 0; rk < s->nranks; rk++) {
        const ora_rank *r = &s->r[rk];
        int is0, is1, js0, js1, ks0, ks1;
        rank_ranges(st, r, &is0, &is1, &js0, &js1, &ks0, &ks1);
        if (sec == 1 && !(r->jbeg <= st->j0_xz && st->j0_xz <= r->jend)) continue;
        if (sec == 2 && !(r->ibeg <= st->i0_yz && st->i0_yz <= r->iend)) continue;
        const int a0 = sec == 2 ? js0 : is0, a1 = sec == 2 ? js1 : is1;
        const int b0 = (sec == 1 || sec == 2) ? ks0 : js0, b1 = (sec == 1 || sec == 2) ? ks1 : js1;
        const ptrdiff_t si = r->nzm, sj = (ptrdiff_t)r->nzm * r->nxm;
        for (int b = b0; b <= b1; b++)
            for (int a = a0; a <= a1; a++) {
                int i, j, k;
                if (sec == 1) { i = a * st->idec - st->idec / 2; j = st->j0_xz; k = b * st->kdec - st->kdec / 2; }
                else if (sec == 2) { i = st->i0_yz; j = a * st->jdec - st->jdec / 2; k = b * st->kdec - st->kdec / 2; }
                else { i = a * st->idec - st->idec / 2; j = b * st->jdec - st->jdec / 2; k = st->k0_xy; }
                if (sec == 3) k = r->kfs[ora_idx2(r, i, j)] + 1;
                if (sec == 4) k = r->kob[ora_idx2(r, i, j)] + 1;
                const ptrdiff_t n = (ptrdiff_t)ora_idx3(r, k, i, j);
                float *o = p->buf + (size_t)(a - 1) + (size_t)p->n1 * (size_t)(b - 1);
                if (typ == 1) {
                    o[0] = (float)(r->Vx[n] * c->UC * c->M0);
                    o[plane] = (float)(r->Vy[n] * c->UC * c->M0);
                    o[2 * plane] = (float)(r->Vz[n] * c->UC * c->M0);
                } else if (typ == 2) {
                    o[0] = (float)(o[0] + r->Vx[n] * c->UC * c->M0 * c->dt);
                    o[plane] = (float)(o[plane] + r->Vy[n] * c->UC * c->M0 * c->dt);
                    o[2 * plane] = (float)(o[2 * plane] + r->Vz[n] * c->UC * c->M0 * c->dt);
                } else {
                    const ora_mp *Vx = r->Vx, *Vy = r->Vy, *Vz = r->Vz;
                    float div = (float)((Vx[n] - Vx[n - si]) * r20x + (Vy[n] - Vy[n - sj]) * r20y + (Vz[n] - Vz[n - 1]) * r20z);
                    float rot_x = (float)((Vz[n + sj] - Vz[n]) * r20y - (Vy[n + 1] - Vy[n]) * r20z);
                    float rot_y = (float)((Vx[n + 1] - Vx[n]) * r20z - (Vz[n + si] - Vz[n]) * r20x);
                    float rot_z = (float)((Vy[n + si] - Vy[n]) * r20x - (Vx[n + sj] - Vx[n]) * r20y);
                    const float lam = r->lam[n];
                    div = div * lam / fabsf(lam + FLT_EPS);
#ifdef ORA_MP_SP
                    rot_x = rot_x * fabsf(r->Syz[n]) / fabsf(r->Syz[n] + FLT_EPS);
                    rot_y = rot_y * fabsf(r->Sxz[n]) / fabsf(r->Sxz[n] + FLT_EPS);
                    rot_z = rot_z * fabsf(r->Sxy[n]) / fabsf(r->Sxy[n] + FLT_EPS);
#else
                    rot_x = (float)(rot_x * fabs(r->Syz[n]) / fabs(r->Syz[n] + FLT_EPS));
                    rot_y = (float)(rot_y * fabs(r->Sxz[n]) / fabs(r->Sxz[n] + FLT_EPS));
                    rot_z = (float)(rot_z * fabs(r->Sxy[n]) / fabs(r->Sxy[n] + FLT_EPS));
#endif
                    o[0] = div * c->UC * c->M0 * 1e-3f;
                    o[plane] = rot_x * c->UC * c->M0 * 1e-3f;
                    o[2 * plane] = rot_y * c->UC * c->M0 * 1e-3f;
                    o[3 * plane] = rot_z * c->UC * c->M0 * 1e-3f;
                }
                if (p->maxv && typ != 0) { /* :1704-1706 */
                    float *m = p->maxv + (size_t)(a - 1) + (size_t)p->n1 * (size_t)(b - 1);
                    float m1 = fmaxf(m[0], fabsf(o[2 * plane]));
                    float m2 = fmaxf(m[plane], sqrtf(o[0] * o[0] + o[plane] * o[plane]));
                    m[0] = m1;
                    m[plane] = m2;
                    m[2 * plane] = sqrtf(m1 * m1 + m2 * m2);
                }
            }
    }
}

/* snap__write(it) m_snap.f90:919-948; netCDF records carry the state at the top of iteration it0 with t = it0*dt (:961-966) */
void ora_snap_write(ora_sim *s, int it) {
    ora_snapstate *st = snap_of(s);
    if (!st) return;
    const int out = st->ntdec_s > 0 && (it - 1) % st->ntdec_s == 0;
    for (int q = 0; q < 15; q++) {
        ora_prod *p = &st->p[q];
        if (!p->on) continue;
        const int sec = q / 3, typ = q % 3;
        const int every = typ == 2 || (typ == 1 && (sec == 3 || sec == 4));
        if (!(every || out)) continue;
        eval_product(s, q);
        if (!out) continue;
        if (p->nrec == p->cap) {
            p->cap = p->cap ? 2 * p->cap : 16;
            p->rec = (float **)realloc(p->rec, sizeof(float *) * (size_t)p->cap);
            p->rec_it = (int *)realloc(p->rec_it, sizeof(int) * (size_t)p->cap);
        }
        size_t n = (size_t)p->n1 * p->n2 * p->nvar;
        p->rec[p->nrec] = (float *)malloc(n * sizeof(float));
        memcpy(p->rec[p->nrec], p->buf, n * sizeof(float));
        p->rec_it[p->nrec] = it;
        p->nrec++;
    }
}

void ora_snap_free(ora_sim *s) {
    ora_snapstate *st = snap_of(s);
    if (!st) return;
    for (int q = 0; q < 15; q++) {
        ora_prod *p = &st->p[q];
        for (int n = 0; n < p->nrec; n++) free(p->rec[n]);
        free(p->rec); free(p->rec_it); free(p->buf); free(p->maxv);
    }
    free(st->xsnp); free(st->ysnp); free(st->zsnp);
    free(st);
    s->snap = NULL;
}

/* ---- accessors: info[0..12] = idec jdec kdec ntdec_s nxs nys nzs k0_xy i0_yz j0_xz; coords */
int ora_snap_info(const ora_sim *s, int *info) {
    const ora_snapstate *st = snap_of(s);
    if (!st) return -1;
    int v[10] = {st->idec, st->jdec, st->kdec, st->ntdec_s, st->nxs, st->nys, st->nzs, st->k0_xy, st->i0_yz, st->j0_xz};
    memcpy(info, v, sizeof(v));
    return 0;
}
int ora_snap_coords(const ora_sim *s, float *x, float *y, float *z) {
    const ora_snapstate *st = snap_of(s);
    if (!st) return -1;
    memcpy(x, st->xsnp, sizeof(float) * (size_t)st->nxs);
    memcpy(y, st->ysnp, sizeof(float) * (size_t)st->nys);
    memcpy(z, st->zsnp, sizeof(float) * (size_t)st->nzs);
    return 0;
}
int ora_snap_nrec(const ora_sim *s, int q) { return snap_of(s) ? snap_of(s)->p[q].nrec : 0; }
int ora_snap_rec(const ora_sim *s, int q, int rec, float *out, int *it0) {
    const ora_prod *p = &snap_of(s)->p[q];
    if (rec < 0 || rec >= p->nrec) return -1;
    memcpy(out, p->rec[rec], sizeof(float) * (size_t)p->n1 * p->n2 * p->nvar);
    if (it0) *it0 = p->rec_it[rec];
    return 0;
}
int ora_snap_max(const ora_sim *s, int q, float *out) {
    const ora_prod *p = &snap_of(s)->p[q];
    if (!p->maxv) return -1;
    memcpy(out, p->maxv, sizeof(float) * (size_t)p->n1 * p->n2 * 3);
    return 0;
}
/* header medium slices (newfile_*_nc, m_snap.f90:475-845): which = 0 rho 1 lambda 2 mu 3 topo 4 lon 5 lat (3-5: xy/fs/ob only) */
int ora_snap_medium(const ora_sim *s, int q, int which, float *out) {
    const ora_cfg *c = &s->cfg;
    const ora_snapstate *st = snap_of(s);
    const ora_prod *p = &st->p[q];
    const int sec = q / 3;
    memset(out, 0, sizeof(float) * (size_t)p->n1 * p->n2);
    if (which > 2 && (sec == 1 || sec == 2)) return -1;
    for (int rk = 0; rk < s->nranks; rk++) {
        const ora_rank *r = &s->r[rk];
        int is0, is1, js0, js1, ks0, ks1;
        rank_ranges(st, r, &is0, &is1, &js0, &js1, &ks0, &ks1);
        if (sec == 1 && !(r->jbeg <= st->j0_xz && st->j0_xz <= r->jend)) continue;
        if (sec == 2 && !(r->ibeg <= st->i0_yz && st->i0_yz <= r->iend)) continue;
        const int a0 = sec == 2 ? js0 : is0, a1 = sec == 2 ? js1 : is1;
        const int b0 = (sec == 1 || sec == 2) ? ks0 : js0, b1 = (sec == 1 || sec == 2) ? ks1 : js1;
        for (int b = b0; b <= b1; b++)
            for (int a = a0; a <= a1; a++) {
                int i, j, k;
                if (sec == 1) { i = a * st->idec - st->idec / 2; j = st->j0_xz; k = b * st->kdec - st->kdec / 2; }
                else if (sec == 2) { i = st->i0_yz; j = a * st->jdec - st->jdec / 2; k = b * st->kdec - st->kdec / 2; }
                else { i = a * st->idec - st->idec / 2; j = b * st->jdec - st->jdec / 2; k = st->k0_xy; }
                if (sec == 3) k = r->kfs[ora_idx2(r, i, j)] + 1;
                if (sec == 4) k = r->kob[ora_idx2(r, i, j)] + 1;
                const size_t n = ora_idx3(r, k, i, j);
                float v;
                if (which == 0) v = r->rho[n];
                else if (which == 1) v = r->lam[n];
                else if (which == 2) v = r->mu[n];
                else if (which == 3) v = -r->bddep[ora_idx2(r, i, j)] * 1000;
                else {
                    float lo, la;
                    ora_geomap_c2g(st->xsnp[a - 1], st->ysnp[b - 1], c->clon, c->clat, c->phi, &lo, &la);
                    v = which == 4 ? lo : la;
                }
                out[(size_t)(a - 1) + (size_t)p->n1 * (size_t)(b - 1)] = v;
            }
    }
    return 0;
}
