/*
 * oracle/ora_step.c -- restatement of one swpc_3d time step (TEST INFRASTRUCTURE, see ora.h).
 *
 * Step order is main.f90:119-139.  Loop bodies keep the declared kinds of the reference
 * (Appendix A of SURVEY.md): float for real(SP), ora_mp for real(MP); C's usual arithmetic
 * conversions coincide with Fortran's mixed-kind promotion for every expression below
 * (float*double -> double, int*float -> float), and the file is compiled with
 * -ffp-contract=off so that no product/sum is fused.
 */
#include "ora.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FLT_EPS 1.1920929e-07f /* epsilon(1.0) */

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/* m_kernel.f90:103-104 / :188-189 : isign = sign(1, max(..)) -> +1 inside either 2nd-order band */
static inline int fd_sign(int k, int kfs_top, int kfs_bot, int kob_top, int kob_bot) {
    int a = (k - kfs_top) * (kfs_bot - k);
    int b = (k - kob_top) * (kob_bot - k);
    return (imax(a, b) >= 0) ? 1 : -1;
}

/* ---------------------------------------------------------------------------------------- */
/* kernel__update_vel  m_kernel.f90:75-140                                                    */
static void kernel_update_vel(const ora_cfg *c, ora_rank *r) {
    const ptrdiff_t si = r->nzm, sj = (ptrdiff_t)r->nzm * r->nxm;
    const float dt = c->dt;
    const ora_mp *restrict Sxx = r->Sxx, *restrict Syy = r->Syy, *restrict Szz = r->Szz;
    const ora_mp *restrict Syz = r->Syz, *restrict Sxz = r->Sxz, *restrict Sxy = r->Sxy;
    const float *restrict rho = r->rho;
    ora_mp *restrict Vx = r->Vx, *restrict Vy = r->Vy, *restrict Vz = r->Vz;
#pragma omp parallel for schedule(static, 1)
    for (int j = r->jbeg_k; j <= r->jend_k; j++) {
        for (int i = r->ibeg_k; i <= r->iend_k; i++) {
            size_t n2 = ora_idx2(r, i, j);
            int kft = r->kfs_top[n2], kfb = r->kfs_bot[n2], kot = r->kob_top[n2], kobb = r->kob_bot[n2];
            for (int k = r->kbeg_k; k <= r->kend_k; k++) {
                ptrdiff_t n = (ptrdiff_t)ora_idx3(r, k, i, j);
                int isign = fd_sign(k, kft, kfb, kot, kobb);
                ora_mp re40x = c->rc40x + isign * c->rd40x, re41x = c->rc41x + isign * c->rd41x;
                ora_mp re40y = c->rc40y + isign * c->rd40y, re41y = c->rc41y + isign * c->rd41y;
                ora_mp re40z = c->rc40z + isign * c->rd40z, re41z = c->rc41z + isign * c->rd41z;

                ora_mp d3Sx3 = (Sxx[n + si] - Sxx[n]) * re40x - (Sxx[n + 2 * si] - Sxx[n - si]) * re41x +
                               (Sxy[n] - Sxy[n - sj]) * re40y - (Sxy[n + sj] - Sxy[n - 2 * sj]) * re41y +
                               (Sxz[n] - Sxz[n - 1]) * re40z - (Sxz[n + 1] - Sxz[n - 2]) * re41z;
                ora_mp d3Sy3 = (Sxy[n] - Sxy[n - si]) * re40x - (Sxy[n + si] - Sxy[n - 2 * si]) * re41x +
                               (Syy[n + sj] - Syy[n]) * re40y - (Syy[n + 2 * sj] - Syy[n - sj]) * re41y +
                               (Syz[n] - Syz[n - 1]) * re40z - (Syz[n + 1] - Syz[n - 2]) * re41z;
                ora_mp d3Sz3 = (Sxz[n] - Sxz[n - si]) * re40x - (Sxz[n + si] - Sxz[n - 2 * si]) * re41x +
                               (Syz[n] - Syz[n - sj]) * re40y - (Syz[n + sj] - Syz[n - 2 * sj]) * re41y +
                               (Szz[n + 1] - Szz[n]) * re40z - (Szz[n + 2] - Szz[n - 1]) * re41z;

                Vx[n] = Vx[n] + 2.0f / (rho[n] + rho[n + si]) * d3Sx3 * dt;
                Vy[n] = Vy[n] + 2.0f / (rho[n] + rho[n + sj]) * d3Sy3 * dt;
                Vz[n] = Vz[n] + 2.0f / (rho[n] + rho[n + 1]) * d3Sz3 * dt;
            }
        }
    }
}

/* harmonic 4-point mean with epsilon in the denominator, m_kernel.f90:293-309 */
static inline float mu_harm(float a, float b, float cc, float d) {
    return 4 * a * b * cc * d / (a * b * cc + a * b * d + a * cc * d + b * cc * d + FLT_EPS);
}

/* ---------------------------------------------------------------------------------------- */
/* kernel__update_stress  m_kernel.f90:142-348 (normal sweep :179-242, shear sweep :266-337)  */
static void kernel_update_stress(const ora_cfg *c, ora_rank *r) {
    const ptrdiff_t si = r->nzm, sj = (ptrdiff_t)r->nzm * r->nxm;
    const float dt = c->dt;
    const int nm = c->nm;
    const float d2 = c->d2;
    const ora_mp *restrict Vx = r->Vx, *restrict Vy = r->Vy, *restrict Vz = r->Vz;
    const float *restrict mu = r->mu, *restrict lam = r->lam, *restrict taup = r->taup, *restrict taus = r->taus;

    /* ---- normal components */
#pragma omp parallel for schedule(static, 1)
    for (int j = r->jbeg_k; j <= r->jend_k; j++) {
        for (int i = r->ibeg_k; i <= r->iend_k; i++) {
            size_t n2 = ora_idx2(r, i, j);
            int kft = r->kfs_top[n2], kfb = r->kfs_bot[n2], kot = r->kob_top[n2], kobb = r->kob_bot[n2];
            for (int k = r->kbeg_k; k <= r->kend_k; k++) {
                ptrdiff_t n = (ptrdiff_t)ora_idx3(r, k, i, j);
                int isign = fd_sign(k, kft, kfb, kot, kobb);
                ora_mp re40x = c->rc40x + isign * c->rd40x, re41x = c->rc41x + isign * c->rd41x;
                ora_mp re40y = c->rc40y + isign * c->rd40y, re41y = c->rc41y + isign * c->rd41y;
                ora_mp re40z = c->rc40z + isign * c->rd40z, re41z = c->rc41z + isign * c->rd41z;

                ora_mp dxVx = (Vx[n] - Vx[n - si]) * re40x - (Vx[n + si] - Vx[n - 2 * si]) * re41x;
                ora_mp dyVy = (Vy[n] - Vy[n - sj]) * re40y - (Vy[n + sj] - Vy[n - 2 * sj]) * re41y;
                ora_mp dzVz = (Vz[n] - Vz[n - 1]) * re40z - (Vz[n + 1] - Vz[n - 2]) * re41z;

                float mu2 = 2 * mu[n];
                float lam2mu = lam[n] + mu2;
                float taup1 = taup[n], taus1 = taus[n];

                float d3v3 = (float)(dxVx + dyVy + dzVz);
                float dyVy_dzVz = (float)(dyVy + dzVz);
                float dxVx_dzVz = (float)(dxVx + dzVz);
                float dxVx_dyVy = (float)(dxVx + dyVy);

                float Rxx_n = 0.0f, Ryy_n = 0.0f, Rzz_n = 0.0f;
                if (nm > 0) {
                    size_t nr = (size_t)nm * ((size_t)(k - r->kbeg_k) +
                                              (size_t)r->nzk * ((size_t)(i - r->ibeg_k) + (size_t)r->nxk * (size_t)(j - r->jbeg_k)));
                    float *Rxx = r->Rxx + nr, *Ryy = r->Ryy + nr, *Rzz = r->Rzz + nr;
                    for (int m = 0; m < nm; m++) {
                        Rxx[m] = c->c1[m] * Rxx[m] - c->c2[m] * (lam2mu * taup1 * d3v3 - mu2 * taus1 * dyVy_dzVz) * dt;
                        Ryy[m] = c->c1[m] * Ryy[m] - c->c2[m] * (lam2mu * taup1 * d3v3 - mu2 * taus1 * dxVx_dzVz) * dt;
                        Rzz[m] = c->c1[m] * Rzz[m] - c->c2[m] * (lam2mu * taup1 * d3v3 - mu2 * taus1 * dxVx_dyVy) * dt;
                        Rxx_n = Rxx_n + c->d1[m] * Rxx[m];
                        Ryy_n = Ryy_n + c->d1[m] * Ryy[m];
                        Rzz_n = Rzz_n + c->d1[m] * Rzz[m];
                    }
                }
                float taup_plus1 = 1 + taup1 * (1 + d2);
                float taus_plus1 = 1 + taus1 * (1 + d2);

                r->Sxx[n] = r->Sxx[n] + (lam2mu * taup_plus1 * d3v3 - mu2 * taus_plus1 * dyVy_dzVz + Rxx_n) * dt;
                r->Syy[n] = r->Syy[n] + (lam2mu * taup_plus1 * d3v3 - mu2 * taus_plus1 * dxVx_dzVz + Ryy_n) * dt;
                r->Szz[n] = r->Szz[n] + (lam2mu * taup_plus1 * d3v3 - mu2 * taus_plus1 * dxVx_dyVy + Rzz_n) * dt;
            }
        }
    }

    /* ---- shear components */
#pragma omp parallel for schedule(static, 1)
    for (int j = r->jbeg_k; j <= r->jend_k; j++) {
        for (int i = r->ibeg_k; i <= r->iend_k; i++) {
            size_t n2 = ora_idx2(r, i, j);
            int kft = r->kfs_top[n2], kfb = r->kfs_bot[n2], kot = r->kob_top[n2], kobb = r->kob_bot[n2];
            for (int k = r->kbeg_k; k <= r->kend_k; k++) {
                ptrdiff_t n = (ptrdiff_t)ora_idx3(r, k, i, j);
                int isign = fd_sign(k, kft, kfb, kot, kobb);
                ora_mp re40x = c->rc40x + isign * c->rd40x, re41x = c->rc41x + isign * c->rd41x;
                ora_mp re40y = c->rc40y + isign * c->rd40y, re41y = c->rc41y + isign * c->rd41y;
                ora_mp re40z = c->rc40z + isign * c->rd40z, re41z = c->rc41z + isign * c->rd41z;

                ora_mp dxVy_dyVx = (Vy[n + si] - Vy[n]) * re40x - (Vy[n + 2 * si] - Vy[n - si]) * re41x +
                                   (Vx[n + sj] - Vx[n]) * re40y - (Vx[n + 2 * sj] - Vx[n - sj]) * re41y;
                ora_mp dxVz_dzVx = (Vz[n + si] - Vz[n]) * re40x - (Vz[n + 2 * si] - Vz[n - si]) * re41x +
                                   (Vx[n + 1] - Vx[n]) * re40z - (Vx[n + 2] - Vx[n - 1]) * re41z;
                ora_mp dyVz_dzVy = (Vz[n + sj] - Vz[n]) * re40y - (Vz[n + 2 * sj] - Vz[n - sj]) * re41y +
                                   (Vy[n + 1] - Vy[n]) * re40z - (Vy[n + 2] - Vy[n - 1]) * re41z;

                float muxz = mu_harm(mu[n], mu[n + 1], mu[n + si], mu[n + 1 + si]);
                float muxy = mu_harm(mu[n], mu[n + si], mu[n + sj], mu[n + si + sj]);
                float muyz = mu_harm(mu[n], mu[n + 1], mu[n + sj], mu[n + 1 + sj]);
                float taus1 = taus[n];

                float Ryz_n = 0.0f, Rxz_n = 0.0f, Rxy_n = 0.0f;
                if (nm > 0) {
                    size_t nr = (size_t)nm * ((size_t)(k - r->kbeg_k) +
                                              (size_t)r->nzk * ((size_t)(i - r->ibeg_k) + (size_t)r->nxk * (size_t)(j - r->jbeg_k)));
                    float *Ryz = r->Ryz + nr, *Rxz = r->Rxz + nr, *Rxy = r->Rxy + nr;
                    for (int m = 0; m < nm; m++) {
                        Ryz[m] = (float)(c->c1[m] * Ryz[m] - c->c2[m] * muyz * taus1 * dyVz_dzVy * dt);
                        Rxz[m] = (float)(c->c1[m] * Rxz[m] - c->c2[m] * muxz * taus1 * dxVz_dzVx * dt);
                        Rxy[m] = (float)(c->c1[m] * Rxy[m] - c->c2[m] * muxy * taus1 * dxVy_dyVx * dt);
                        Ryz_n = Ryz_n + c->d1[m] * Ryz[m];
                        Rxz_n = Rxz_n + c->d1[m] * Rxz[m];
                        Rxy_n = Rxy_n + c->d1[m] * Rxy[m];
                    }
                }
                float taus_plus1 = 1 + taus1 * (1 + d2);

                r->Syz[n] = r->Syz[n] + (muyz * taus_plus1 * dyVz_dzVy + Ryz_n) * dt;
                r->Sxz[n] = r->Sxz[n] + (muxz * taus_plus1 * dxVz_dzVx + Rxz_n) * dt;
                r->Sxy[n] = r->Sxy[n] + (muxy * taus_plus1 * dxVy_dyVx + Rxy_n) * dt;
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------- */
/* Horizontal zero-derivative boundary of the plane-wave mode (m_absorb_p.f90:137-243 for the stresses, :332-424 for the
 * velocities): linear extrapolation into the first plane outside the model on the outer ranks, owned rows only */
static void pw_edges(const ora_cfg *c, ora_rank *r, ora_mp **f, int nf) {
    for (int q = 0; q < nf; q++) {
        ora_mp *a = f[q];
        if (r->idx == 0)
            for (int j = r->jbeg; j <= r->jend; j++)
                for (int k = 1; k <= c->nz; k++) a[ora_idx3(r, k, 0, j)] = 2 * a[ora_idx3(r, k, 1, j)] - a[ora_idx3(r, k, 2, j)];
        if (r->idx == c->nproc_x - 1)
            for (int j = r->jbeg; j <= r->jend; j++)
                for (int k = 1; k <= c->nz; k++) a[ora_idx3(r, k, c->nx + 1, j)] = 2 * a[ora_idx3(r, k, c->nx, j)] - a[ora_idx3(r, k, c->nx - 1, j)];
        if (r->idy == 0)
            for (int i = r->ibeg; i <= r->iend; i++)
                for (int k = 1; k <= c->nz; k++) a[ora_idx3(r, k, i, 0)] = 2 * a[ora_idx3(r, k, i, 1)] - a[ora_idx3(r, k, i, 2)];
        if (r->idy == c->nproc_y - 1)
            for (int i = r->ibeg; i <= r->iend; i++)
                for (int k = 1; k <= c->nz; k++) a[ora_idx3(r, k, i, c->ny + 1)] = 2 * a[ora_idx3(r, k, i, c->ny)] - a[ora_idx3(r, k, i, c->ny - 1)];
    }
}

/* ---------------------------------------------------------------------------------------- */
/* absorb_p__update_vel  m_absorb_p.f90:246-317 (time-marching part) */
static void absorb_p_update_vel(const ora_cfg *c, ora_rank *r) {
    const ptrdiff_t si = r->nzm, sj = (ptrdiff_t)r->nzm * r->nxm;
    const float dt = c->dt;
    const ora_mp r20x = c->r20x, r20y = c->r20y, r20z = c->r20z;
    const ora_mp *restrict Sxx = r->Sxx, *restrict Syy = r->Syy, *restrict Szz = r->Szz;
    const ora_mp *restrict Syz = r->Syz, *restrict Sxz = r->Sxz, *restrict Sxy = r->Sxy;
    const float *restrict rho = r->rho;
#pragma omp parallel for schedule(dynamic)
    for (int j = r->jbeg; j <= r->jend; j++) {
        const float *gyc = &r->gyc[4 * (j - r->jbeg)], *gye = &r->gye[4 * (j - r->jbeg)];
        for (int i = r->ibeg; i <= r->iend; i++) {
            const float *gxc = &r->gxc[4 * (i - r->ibeg)], *gxe = &r->gxe[4 * (i - r->ibeg)];
            int kb = r->kbeg_a[ora_idx2(r, i, j)];
            int64_t a0 = r->aoff[(size_t)(i - r->ibeg) + (size_t)r->nxp * (j - r->jbeg)] - kb;
            for (int k = kb; k <= r->kend; k++) {
                const float *gzc = &r->gzc[4 * (k - r->kbeg)], *gze = &r->gze[4 * (k - r->kbeg)];
                ptrdiff_t n = (ptrdiff_t)ora_idx3(r, k, i, j);
                int64_t a = a0 + k;

                ora_mp dxSxx = (Sxx[n + si] - Sxx[n]) * r20x;
                ora_mp dySyy = (Syy[n + sj] - Syy[n]) * r20y;
                ora_mp dzSzz = (Szz[n + 1] - Szz[n]) * r20z;
                ora_mp dySyz = (Syz[n] - Syz[n - sj]) * r20y;
                ora_mp dzSyz = (Syz[n] - Syz[n - 1]) * r20z;
                ora_mp dxSxz = (Sxz[n] - Sxz[n - si]) * r20x;
                ora_mp dzSxz = (Sxz[n] - Sxz[n - 1]) * r20z;
                ora_mp dxSxy = (Sxy[n] - Sxy[n - si]) * r20x;
                ora_mp dySxy = (Sxy[n] - Sxy[n - sj]) * r20y;

                float bx = 2.0f / (rho[n] + rho[n + si]);
                float by = 2.0f / (rho[n] + rho[n + sj]);
                float bz = 2.0f / (rho[n] + rho[n + 1]);

                r->Vx[n] = r->Vx[n] + bx * (float)(gxe[0] * dxSxx + gyc[0] * dySxy + gzc[0] * dzSxz +
                                                   gxe[1] * r->axSxx[a] + gyc[1] * r->aySxy[a] + gzc[1] * r->azSxz[a]) * dt;
                r->Vy[n] = r->Vy[n] + by * (float)(gxc[0] * dxSxy + gye[0] * dySyy + gzc[0] * dzSyz +
                                                   gxc[1] * r->axSxy[a] + gye[1] * r->aySyy[a] + gzc[1] * r->azSyz[a]) * dt;
                r->Vz[n] = r->Vz[n] + bz * (float)(gxc[0] * dxSxz + gyc[0] * dySyz + gze[0] * dzSzz +
                                                   gxc[1] * r->axSxz[a] + gyc[1] * r->aySyz[a] + gze[1] * r->azSzz[a]) * dt;

                r->axSxx[a] = gxe[2] * r->axSxx[a] + gxe[3] * (float)(dxSxx)*dt;
                r->aySxy[a] = gyc[2] * r->aySxy[a] + gyc[3] * (float)(dySxy)*dt;
                r->azSxz[a] = gzc[2] * r->azSxz[a] + gzc[3] * (float)(dzSxz)*dt;
                r->axSxy[a] = gxc[2] * r->axSxy[a] + gxc[3] * (float)(dxSxy)*dt;
                r->aySyy[a] = gye[2] * r->aySyy[a] + gye[3] * (float)(dySyy)*dt;
                r->azSyz[a] = gzc[2] * r->azSyz[a] + gzc[3] * (float)(dzSyz)*dt;
                r->axSxz[a] = gxc[2] * r->axSxz[a] + gxc[3] * (float)(dxSxz)*dt;
                r->aySyz[a] = gyc[2] * r->aySyz[a] + gyc[3] * (float)(dySyz)*dt;
                r->azSzz[a] = gze[2] * r->azSzz[a] + gze[3] * (float)(dzSzz)*dt;
            }
        }
    }
}

/* absorb_p__update_stress  m_absorb_p.f90:429-531 */
static void absorb_p_update_stress(const ora_cfg *c, ora_rank *r) {
    const ptrdiff_t si = r->nzm, sj = (ptrdiff_t)r->nzm * r->nxm;
    const float dt = c->dt;
    const ora_mp r20x = c->r20x, r20y = c->r20y, r20z = c->r20z;
    const ora_mp *restrict Vx = r->Vx, *restrict Vy = r->Vy, *restrict Vz = r->Vz;
    const float *restrict mu = r->mu, *restrict lam = r->lam;
#pragma omp parallel for schedule(dynamic)
    for (int j = r->jbeg; j <= r->jend; j++) {
        const float *gyc = &r->gyc[4 * (j - r->jbeg)], *gye = &r->gye[4 * (j - r->jbeg)];
        for (int i = r->ibeg; i <= r->iend; i++) {
            const float *gxc = &r->gxc[4 * (i - r->ibeg)], *gxe = &r->gxe[4 * (i - r->ibeg)];
            int kb = r->kbeg_a[ora_idx2(r, i, j)];
            int64_t a0 = r->aoff[(size_t)(i - r->ibeg) + (size_t)r->nxp * (j - r->jbeg)] - kb;
            /* loop 1 :453-474 */
            for (int k = kb; k <= r->kend; k++) {
                const float *gzc = &r->gzc[4 * (k - r->kbeg)];
                ptrdiff_t n = (ptrdiff_t)ora_idx3(r, k, i, j);
                int64_t a = a0 + k;
                ora_mp dxVx = (Vx[n] - Vx[n - si]) * r20x;
                ora_mp dyVy = (Vy[n] - Vy[n - sj]) * r20y;
                ora_mp dzVz = (Vz[n] - Vz[n - 1]) * r20z;
                float lam2mu_R = (lam[n] + 2 * mu[n]);
                float lam_R = lam2mu_R - 2 * mu[n];
                float dxVx_ade = gxc[0] * (float)(dxVx) + gxc[1] * r->axVx[a];
                float dyVy_ade = gyc[0] * (float)(dyVy) + gyc[1] * r->ayVy[a];
                float dzVz_ade = gzc[0] * (float)(dzVz) + gzc[1] * r->azVz[a];
                r->Sxx[n] = r->Sxx[n] + (lam2mu_R * dxVx_ade + lam_R * (dyVy_ade + dzVz_ade)) * dt;
                r->Syy[n] = r->Syy[n] + (lam2mu_R * dyVy_ade + lam_R * (dxVx_ade + dzVz_ade)) * dt;
                r->Szz[n] = r->Szz[n] + (lam2mu_R * dzVz_ade + lam_R * (dxVx_ade + dyVy_ade)) * dt;
                r->axVx[a] = gxc[2] * r->axVx[a] + gxc[3] * (float)(dxVx)*dt;
                r->ayVy[a] = gyc[2] * r->ayVy[a] + gyc[3] * (float)(dyVy)*dt;
                r->azVz[a] = gzc[2] * r->azVz[a] + gzc[3] * (float)(dzVz)*dt;
            }
            /* loop 2 :477-519 */
            for (int k = kb; k <= r->kend; k++) {
                const float *gze = &r->gze[4 * (k - r->kbeg)];
                ptrdiff_t n = (ptrdiff_t)ora_idx3(r, k, i, j);
                int64_t a = a0 + k;
                ora_mp dxVy = (Vy[n + si] - Vy[n]) * r20x;
                ora_mp dxVz = (Vz[n + si] - Vz[n]) * r20x;
                ora_mp dyVx = (Vx[n + sj] - Vx[n]) * r20y;
                ora_mp dyVz = (Vz[n + sj] - Vz[n]) * r20y;
                ora_mp dzVx = (Vx[n + 1] - Vx[n]) * r20z;
                ora_mp dzVy = (Vy[n + 1] - Vy[n]) * r20z;
                float muxz = mu_harm(mu[n], mu[n + 1], mu[n + si], mu[n + 1 + si]);
                float muxy = mu_harm(mu[n], mu[n + si], mu[n + sj], mu[n + si + sj]);
                float muyz = mu_harm(mu[n], mu[n + 1], mu[n + sj], mu[n + 1 + sj]);
                r->Syz[n] = r->Syz[n] + muyz * (gye[0] * dyVz + gze[0] * dzVy + gye[1] * r->ayVz[a] + gze[1] * r->azVy[a]) * dt;
                r->Sxz[n] = r->Sxz[n] + muxz * (gxe[0] * dxVz + gze[0] * dzVx + gxe[1] * r->axVz[a] + gze[1] * r->azVx[a]) * dt;
                r->Sxy[n] = r->Sxy[n] + muxy * (gxe[0] * dxVy + gye[0] * dyVx + gxe[1] * r->axVy[a] + gye[1] * r->ayVx[a]) * dt;
                r->ayVx[a] = gye[2] * r->ayVx[a] + gye[3] * (float)(dyVx)*dt;
                r->azVx[a] = gze[2] * r->azVx[a] + gze[3] * (float)(dzVx)*dt;
                r->axVy[a] = gxe[2] * r->axVy[a] + gxe[3] * (float)(dxVy)*dt;
                r->azVy[a] = gze[2] * r->azVy[a] + gze[3] * (float)(dzVy)*dt;
                r->axVz[a] = gxe[2] * r->axVz[a] + gxe[3] * (float)(dxVz)*dt;
                r->ayVz[a] = gye[2] * r->ayVz[a] + gye[3] * (float)(dyVz)*dt;
            }
        }
    }
}

/* ---------------------------------------------------------------------------------------- */
/* m_absorb_c.f90:113-168 / :170-198                                                          */
static void absorb_c_update_stress(ora_rank *r) {
#pragma omp parallel for schedule(dynamic)
    for (int j = r->jbeg; j <= r->jend; j++)
        for (int i = r->ibeg; i <= r->iend; i++)
            for (int k = r->kbeg; k <= r->kend_k; k++) {
                size_t n = ora_idx3(r, k, i, j);
                float gcc = r->gx_c[i - r->ibeg_m] * r->gy_c[j - r->jbeg_m] * r->gz_c[k - r->kbeg_m];
                r->Sxx[n] = r->Sxx[n] * gcc;
                r->Syy[n] = r->Syy[n] * gcc;
                r->Szz[n] = r->Szz[n] * gcc;
            }
#pragma omp parallel for schedule(dynamic)
    for (int j = r->jbeg; j <= r->jend; j++)
        for (int i = r->ibeg; i <= r->iend; i++)
            for (int k = r->kbeg; k <= r->kend_k; k++) {
                size_t n = ora_idx3(r, k, i, j);
                float gxc = r->gx_c[i - r->ibeg_m], gxb = r->gx_b[i - r->ibeg_m];
                float gyc = r->gy_c[j - r->jbeg_m], gyb = r->gy_b[j - r->jbeg_m];
                float gzc = r->gz_c[k - r->kbeg_m], gzb = r->gz_b[k - r->kbeg_m];
                r->Syz[n] = r->Syz[n] * gxc * gyb * gzb;
                r->Sxz[n] = r->Sxz[n] * gxb * gyc * gzb;
                r->Sxy[n] = r->Sxy[n] * gxb * gyb * gzc;
            }
}

static void absorb_c_update_vel(ora_rank *r) {
#pragma omp parallel for schedule(dynamic)
    for (int j = r->jbeg; j <= r->jend; j++)
        for (int i = r->ibeg; i <= r->iend; i++)
            for (int k = r->kbeg; k <= r->kend; k++) {
                size_t n = ora_idx3(r, k, i, j);
                float gxc = r->gx_c[i - r->ibeg_m], gxb = r->gx_b[i - r->ibeg_m];
                float gyc = r->gy_c[j - r->jbeg_m], gyb = r->gy_b[j - r->jbeg_m];
                float gzc = r->gz_c[k - r->kbeg_m], gzb = r->gz_b[k - r->kbeg_m];
                r->Vx[n] = r->Vx[n] * gxb * gyc * gzc;
                r->Vy[n] = r->Vy[n] * gxc * gyb * gzc;
                r->Vz[n] = r->Vz[n] * gxc * gyc * gzb;
            }
}

/* ---------------------------------------------------------------------------------------- */
void ora_update_stress(ora_sim *s) {
    int pml = !strcmp(s->cfg.abc_type, "pml");
    for (int q = 0; q < s->nranks; q++) {
        kernel_update_stress(&s->cfg, &s->r[q]);
        if (pml && s->cfg.pw_mode) { ora_mp *f[3] = {s->r[q].Vx, s->r[q].Vy, s->r[q].Vz}; pw_edges(&s->cfg, &s->r[q], f, 3); }
        if (pml) absorb_p_update_stress(&s->cfg, &s->r[q]);
        else absorb_c_update_stress(&s->r[q]);
    }
}

void ora_update_vel(ora_sim *s) {
    int pml = !strcmp(s->cfg.abc_type, "pml");
    for (int q = 0; q < s->nranks; q++) {
        kernel_update_vel(&s->cfg, &s->r[q]);
        if (pml && s->cfg.pw_mode) {
            ora_rank *r = &s->r[q];
            ora_mp *f[6] = {r->Sxx, r->Syy, r->Szz, r->Syz, r->Sxz, r->Sxy};
            pw_edges(&s->cfg, r, f, 6);
        }
        if (pml) absorb_p_update_vel(&s->cfg, &s->r[q]);
        else absorb_c_update_vel(&s->r[q]);
    }
}

/* source__stressglut  m_source.f90:776-848 */
void ora_stressglut(ora_sim *s, int it) {
    const ora_cfg *c = &s->cfg;
    if (c->bf_mode || c->green_mode) return;
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        const ptrdiff_t si = r->nzm, sj = (ptrdiff_t)r->nzm * r->nxm;
        for (int i = 0; i < r->nsrc; i++) {
            float t = c->tbeg + ((float)it - 0.5f) * c->dt;
            float stime = ora_momentrate(t, c->stftype, &r->srcprm[2 * i]);
            ora_mp sdrop = r->mo[i] * stime * c->dt_dxyz;
            ptrdiff_t n = (ptrdiff_t)ora_idx3(r, r->ksrc[i], r->isrc[i], r->jsrc[i]);
            r->Sxx[n] = r->Sxx[n] - r->mxx[i] * sdrop;
            r->Syy[n] = r->Syy[n] - r->myy[i] * sdrop;
            r->Szz[n] = r->Szz[n] - r->mzz[i] * sdrop;
            r->Sxy[n] = r->Sxy[n] - r->mxy[i] * sdrop / 4;
            r->Sxy[n - sj] = r->Sxy[n - sj] - r->mxy[i] * sdrop / 4;
            r->Sxy[n - si] = r->Sxy[n - si] - r->mxy[i] * sdrop / 4;
            r->Sxy[n - si - sj] = r->Sxy[n - si - sj] - r->mxy[i] * sdrop / 4;
            r->Sxz[n] = r->Sxz[n] - r->mxz[i] * sdrop / 4;
            r->Sxz[n - 1] = r->Sxz[n - 1] - r->mxz[i] * sdrop / 4;
            r->Sxz[n - si] = r->Sxz[n - si] - r->mxz[i] * sdrop / 4;
            r->Sxz[n - 1 - si] = r->Sxz[n - 1 - si] - r->mxz[i] * sdrop / 4;
            r->Syz[n] = r->Syz[n] - r->myz[i] * sdrop / 4;
            r->Syz[n - 1] = r->Syz[n - 1] - r->myz[i] * sdrop / 4;
            r->Syz[n - sj] = r->Syz[n - sj] - r->myz[i] * sdrop / 4;
            r->Syz[n - 1 - sj] = r->Syz[n - 1 - sj] - r->myz[i] * sdrop / 4;
        }
    }
}

/* source__bodyforce  m_source.f90:850-898 */
void ora_bodyforce(ora_sim *s, int it) {
    const ora_cfg *c = &s->cfg;
    if (!c->bf_mode) return;
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        const ptrdiff_t si = r->nzm, sj = (ptrdiff_t)r->nzm * r->nxm;
        const float *rho = r->rho;
        for (int i = 0; i < r->nsrc; i++) {
            float t = c->tbeg + it * c->dt;
            float stime = ora_momentrate(t, c->stftype, &r->srcprm[2 * i]);
            ptrdiff_t n = (ptrdiff_t)ora_idx3(r, r->ksrc[i], r->isrc[i], r->jsrc[i]);
            r->Vx[n] = r->Vx[n] + (2.0f / (rho[n] + rho[n + si])) * r->fx[i] * stime * c->dt_dxyz / 2;
            r->Vx[n - si] = r->Vx[n - si] + (2.0f / (rho[n] + rho[n - si])) * r->fx[i] * stime * c->dt_dxyz / 2;
            r->Vy[n] = r->Vy[n] + (2.0f / (rho[n] + rho[n + sj])) * r->fy[i] * stime * c->dt_dxyz / 2;
            r->Vy[n - sj] = r->Vy[n - sj] + (2.0f / (rho[n] + rho[n - sj])) * r->fy[i] * stime * c->dt_dxyz / 2;
            r->Vz[n] = r->Vz[n] + (2.0f / (rho[n] + rho[n + 1])) * r->fz[i] * stime * c->dt_dxyz / 2;
            r->Vz[n - 1] = r->Vz[n - 1] + (2.0f / (rho[n] + rho[n - 1])) * r->fz[i] * stime * c->dt_dxyz / 2;
        }
    }
}

/* ---------------------------------------------------------------------------------------- */
/* halo exchange  m_global.f90:391-607.  A plane is (k=1..nz, j=jbeg..jend) at fixed i for the
 * x faces, (k=1..nz, i=ibeg..iend) at fixed j for the y faces; no corners, no k halo.          */
static void pack_i(const ora_rank *r, int nz, ora_mp *buf, int slot, const ora_mp *F, int i) {
    size_t isize = (size_t)r->nyp * nz;
    for (int j = r->jbeg; j <= r->jend; j++)
        memcpy(buf + slot * isize + (size_t)(j - r->jbeg) * nz, F + ora_idx3(r, 1, i, j), sizeof(ora_mp) * (size_t)nz);
}
static void unpack_i(const ora_rank *r, int nz, const ora_mp *buf, int slot, ora_mp *F, int i) {
    size_t isize = (size_t)r->nyp * nz;
    for (int j = r->jbeg; j <= r->jend; j++)
        memcpy(F + ora_idx3(r, 1, i, j), buf + slot * isize + (size_t)(j - r->jbeg) * nz, sizeof(ora_mp) * (size_t)nz);
}
static void pack_j(const ora_rank *r, int nz, ora_mp *buf, int slot, const ora_mp *F, int j) {
    size_t jsize = (size_t)r->nxp * nz;
    for (int i = r->ibeg; i <= r->iend; i++)
        memcpy(buf + slot * jsize + (size_t)(i - r->ibeg) * nz, F + ora_idx3(r, 1, i, j), sizeof(ora_mp) * (size_t)nz);
}
static void unpack_j(const ora_rank *r, int nz, const ora_mp *buf, int slot, ora_mp *F, int j) {
    size_t jsize = (size_t)r->nxp * nz;
    for (int i = r->ibeg; i <= r->iend; i++)
        memcpy(F + ora_idx3(r, 1, i, j), buf + slot * jsize + (size_t)(i - r->ibeg) * nz, sizeof(ora_mp) * (size_t)nz);
}

static int nbr(const ora_sim *s, int idx, int idy) {
    int tw = s->cfg.nproc_x + 2;
    return s->itbl[(idx + 1) + tw * (idy + 1)];
}

/* deliver: neighbour's recv buffer <- my send buffer (what mpi_isend/irecv do).  A missing neighbour is
 * MPI_PROC_NULL (m_global.f90:630): nothing is delivered, the zero-initialised rbuf (:253-258) stays zero and the
 * unconditional unpack (:458-488) keeps writing those zeros into the outer halo. */
static void deliver(ora_sim *s, size_t isz_unit_planes_ip, size_t isz_unit_planes_im) {
    (void)isz_unit_planes_ip;
    (void)isz_unit_planes_im;
    const int nz = s->cfg.nz;
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        int p;
        /* to +x: my sbuf_ip -> their rbuf_im ; to -x: my sbuf_im -> their rbuf_ip */
        if ((p = nbr(s, r->idx + 1, r->idy)) >= 0) memcpy(s->r[p].rbuf_im, r->sbuf_ip, sizeof(ora_mp) * 5 * (size_t)r->nyp * nz);
        if ((p = nbr(s, r->idx - 1, r->idy)) >= 0) memcpy(s->r[p].rbuf_ip, r->sbuf_im, sizeof(ora_mp) * 5 * (size_t)r->nyp * nz);
        if ((p = nbr(s, r->idx, r->idy + 1)) >= 0) memcpy(s->r[p].rbuf_jm, r->sbuf_jp, sizeof(ora_mp) * 5 * (size_t)r->nxp * nz);
        if ((p = nbr(s, r->idx, r->idy - 1)) >= 0) memcpy(s->r[p].rbuf_jp, r->sbuf_jm, sizeof(ora_mp) * 5 * (size_t)r->nxp * nz);
    }
}

/* global__comm_vel  m_global.f90:391-494 */
void ora_comm_vel(ora_sim *s) {
    const int nz = s->cfg.nz;
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        int ib = r->ibeg, ie = r->iend, jb = r->jbeg, je = r->jend;
        pack_i(r, nz, r->sbuf_ip, 0, r->Vx, ie - 1); pack_i(r, nz, r->sbuf_ip, 1, r->Vx, ie);
        pack_i(r, nz, r->sbuf_ip, 2, r->Vy, ie);     pack_i(r, nz, r->sbuf_ip, 3, r->Vz, ie);
        pack_i(r, nz, r->sbuf_im, 0, r->Vx, ib);     pack_i(r, nz, r->sbuf_im, 1, r->Vy, ib);
        pack_i(r, nz, r->sbuf_im, 2, r->Vy, ib + 1); pack_i(r, nz, r->sbuf_im, 3, r->Vz, ib);
        pack_i(r, nz, r->sbuf_im, 4, r->Vz, ib + 1);
        pack_j(r, nz, r->sbuf_jp, 0, r->Vx, je);     pack_j(r, nz, r->sbuf_jp, 1, r->Vy, je - 1);
        pack_j(r, nz, r->sbuf_jp, 2, r->Vy, je);     pack_j(r, nz, r->sbuf_jp, 3, r->Vz, je);
        pack_j(r, nz, r->sbuf_jm, 0, r->Vx, jb);     pack_j(r, nz, r->sbuf_jm, 1, r->Vx, jb + 1);
        pack_j(r, nz, r->sbuf_jm, 2, r->Vy, jb);     pack_j(r, nz, r->sbuf_jm, 3, r->Vz, jb);
        pack_j(r, nz, r->sbuf_jm, 4, r->Vz, jb + 1);
    }
    deliver(s, 0, 0);
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        int ib = r->ibeg, ie = r->iend, jb = r->jbeg, je = r->jend;
        {
            unpack_i(r, nz, r->rbuf_im, 0, r->Vx, ib - 2); unpack_i(r, nz, r->rbuf_im, 1, r->Vx, ib - 1);
            unpack_i(r, nz, r->rbuf_im, 2, r->Vy, ib - 1); unpack_i(r, nz, r->rbuf_im, 3, r->Vz, ib - 1);
        }
        {
            unpack_i(r, nz, r->rbuf_ip, 0, r->Vx, ie + 1); unpack_i(r, nz, r->rbuf_ip, 1, r->Vy, ie + 1);
            unpack_i(r, nz, r->rbuf_ip, 2, r->Vy, ie + 2); unpack_i(r, nz, r->rbuf_ip, 3, r->Vz, ie + 1);
            unpack_i(r, nz, r->rbuf_ip, 4, r->Vz, ie + 2);
        }
        {
            unpack_j(r, nz, r->rbuf_jm, 0, r->Vx, jb - 1); unpack_j(r, nz, r->rbuf_jm, 1, r->Vy, jb - 2);
            unpack_j(r, nz, r->rbuf_jm, 2, r->Vy, jb - 1); unpack_j(r, nz, r->rbuf_jm, 3, r->Vz, jb - 1);
        }
        {
            unpack_j(r, nz, r->rbuf_jp, 0, r->Vx, je + 1); unpack_j(r, nz, r->rbuf_jp, 1, r->Vx, je + 2);
            unpack_j(r, nz, r->rbuf_jp, 2, r->Vy, je + 1); unpack_j(r, nz, r->rbuf_jp, 3, r->Vz, je + 1);
            unpack_j(r, nz, r->rbuf_jp, 4, r->Vz, je + 2);
        }
    }
}

/* global__comm_stress  m_global.f90:500-607 */
void ora_comm_stress(ora_sim *s) {
    const int nz = s->cfg.nz;
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        int ib = r->ibeg, ie = r->iend, jb = r->jbeg, je = r->jend;
        pack_i(r, nz, r->sbuf_ip, 0, r->Sxx, ie);     pack_i(r, nz, r->sbuf_ip, 1, r->Sxy, ie - 1);
        pack_i(r, nz, r->sbuf_ip, 2, r->Sxy, ie);     pack_i(r, nz, r->sbuf_ip, 3, r->Sxz, ie - 1);
        pack_i(r, nz, r->sbuf_ip, 4, r->Sxz, ie);
        pack_i(r, nz, r->sbuf_im, 0, r->Sxx, ib);     pack_i(r, nz, r->sbuf_im, 1, r->Sxx, ib + 1);
        pack_i(r, nz, r->sbuf_im, 2, r->Sxy, ib);     pack_i(r, nz, r->sbuf_im, 3, r->Sxz, ib);
        pack_j(r, nz, r->sbuf_jp, 0, r->Syy, je);     pack_j(r, nz, r->sbuf_jp, 1, r->Sxy, je - 1);
        pack_j(r, nz, r->sbuf_jp, 2, r->Sxy, je);     pack_j(r, nz, r->sbuf_jp, 3, r->Syz, je - 1);
        pack_j(r, nz, r->sbuf_jp, 4, r->Syz, je);
        pack_j(r, nz, r->sbuf_jm, 0, r->Syy, jb);     pack_j(r, nz, r->sbuf_jm, 1, r->Syy, jb + 1);
        pack_j(r, nz, r->sbuf_jm, 2, r->Sxy, jb);     pack_j(r, nz, r->sbuf_jm, 3, r->Syz, jb);
    }
    deliver(s, 0, 0);
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        int ib = r->ibeg, ie = r->iend, jb = r->jbeg, je = r->jend;
        {
            unpack_i(r, nz, r->rbuf_im, 0, r->Sxx, ib - 1); unpack_i(r, nz, r->rbuf_im, 1, r->Sxy, ib - 2);
            unpack_i(r, nz, r->rbuf_im, 2, r->Sxy, ib - 1); unpack_i(r, nz, r->rbuf_im, 3, r->Sxz, ib - 2);
            unpack_i(r, nz, r->rbuf_im, 4, r->Sxz, ib - 1);
        }
        {
            unpack_i(r, nz, r->rbuf_ip, 0, r->Sxx, ie + 1); unpack_i(r, nz, r->rbuf_ip, 1, r->Sxx, ie + 2);
            unpack_i(r, nz, r->rbuf_ip, 2, r->Sxy, ie + 1); unpack_i(r, nz, r->rbuf_ip, 3, r->Sxz, ie + 1);
        }
        {
            unpack_j(r, nz, r->rbuf_jm, 0, r->Syy, jb - 1); unpack_j(r, nz, r->rbuf_jm, 1, r->Sxy, jb - 2);
            unpack_j(r, nz, r->rbuf_jm, 2, r->Sxy, jb - 1); unpack_j(r, nz, r->rbuf_jm, 3, r->Syz, jb - 2);
            unpack_j(r, nz, r->rbuf_jm, 4, r->Syz, jb - 1);
        }
        {
            unpack_j(r, nz, r->rbuf_jp, 0, r->Syy, je + 1); unpack_j(r, nz, r->rbuf_jp, 1, r->Syy, je + 2);
            unpack_j(r, nz, r->rbuf_jp, 2, r->Sxy, je + 1); unpack_j(r, nz, r->rbuf_jp, 3, r->Syz, je + 1);
        }
    }
}

/* ---------------------------------------------------------------------------------------- */
/* wav__store  m_wav.f90:397-625: displacement / strain are accumulated EVERY step (:430-513), then, when
 * mod(it-1, ntdec_w) == 0, velocity, displacement, stress and strain are sampled (:515-617). */
void ora_wav_store(ora_sim *s, int it) {
    const ora_cfg *c = &s->cfg;
    if (!(c->sw_wav_v || c->sw_wav_u || c->sw_wav_stress || c->sw_wav_strain) || c->ntdec_w <= 0) return;
    const float dt = c->dt;
    /* m_wav.f90:127-132 */
    const ora_mp r40x = (ora_mp)9.0 / (ora_mp)8.0 / (ora_mp)c->dx, r40y = (ora_mp)9.0 / (ora_mp)8.0 / (ora_mp)c->dy, r40z = (ora_mp)9.0 / (ora_mp)8.0 / (ora_mp)c->dz;
    const ora_mp r41x = (ora_mp)1.0 / (ora_mp)24.0 / (ora_mp)c->dx, r41y = (ora_mp)1.0 / (ora_mp)24.0 / (ora_mp)c->dy, r41z = (ora_mp)1.0 / (ora_mp)24.0 / (ora_mp)c->dz;
    const int sample = ((it - 1) % c->ntdec_w == 0);
    const int itw = (it - 1) / c->ntdec_w + 1;
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        const ptrdiff_t si = r->nzm, sj = (ptrdiff_t)r->nzm * r->nxm;
        const ora_mp *Vx = r->Vx, *Vy = r->Vy, *Vz = r->Vz;
        for (int n = 0; n < r->nst; n++) {
            ptrdiff_t p = (ptrdiff_t)ora_idx3(r, r->kst[n], r->ist[n], r->jst[n]);
            if (c->sw_wav_u) { /* :439-444 */
                r->ux[n] = r->ux[n] + (float)(Vx[p] + Vx[p - si]) * 0.5f * dt;
                r->uy[n] = r->uy[n] + (float)(Vy[p] + Vy[p - sj]) * 0.5f * dt;
                r->uz[n] = r->uz[n] - (float)(Vz[p] + Vz[p - 1]) * 0.5f * dt;
            }
            if (c->sw_wav_strain) { /* :462-506 */
                ora_mp dxVx = (Vx[p] - Vx[p - si]) * r40x - (Vx[p + si] - Vx[p - 2 * si]) * r41x;
                ora_mp dyVy = (Vy[p] - Vy[p - sj]) * r40y - (Vy[p + sj] - Vy[p - 2 * sj]) * r41y;
                ora_mp dzVz = (Vz[p] - Vz[p - 1]) * r40z - (Vz[p + 1] - Vz[p - 2]) * r41z;
                ora_mp dxVy = ((Vy[p + si] - Vy[p]) * r40x - (Vy[p + 2 * si] - Vy[p - si]) * r41x +
                               (Vy[p + si - sj] - Vy[p - sj]) * r40x - (Vy[p + 2 * si - sj] - Vy[p - si - sj]) * r41x +
                               (Vy[p] - Vy[p - si]) * r40x - (Vy[p + si] - Vy[p - 2 * si]) * r41x +
                               (Vy[p - sj] - Vy[p - si - sj]) * r40x - (Vy[p + si - sj] - Vy[p - 2 * si - sj]) * r41x) / 4.0f;
                ora_mp dxVz = ((Vz[p + si] - Vz[p]) * r40x - (Vz[p + 2 * si] - Vz[p - si]) * r41x +
                               (Vz[p - 1 + si] - Vz[p - 1]) * r40x - (Vz[p - 1 + 2 * si] - Vz[p - 1 - si]) * r41x +
                               (Vz[p] - Vz[p - si]) * r40x - (Vz[p + si] - Vz[p - 2 * si]) * r41x +
                               (Vz[p - 1] - Vz[p - 1 - si]) * r40x - (Vz[p - 1 + si] - Vz[p - 1 - 2 * si]) * r41x) / 4.0f;
                ora_mp dyVx = ((Vx[p + sj] - Vx[p]) * r40y - (Vx[p + 2 * sj] - Vx[p - sj]) * r41y +
                               (Vx[p - si + sj] - Vx[p - si]) * r40y - (Vx[p - si + 2 * sj] - Vx[p - si - sj]) * r41y +
                               (Vx[p] - Vx[p - sj]) * r40y - (Vx[p + sj] - Vx[p - 2 * sj]) * r41y +
                               (Vx[p - si] - Vx[p - si - sj]) * r40y - (Vx[p - si + sj] - Vx[p - si - 2 * sj]) * r41y) / 4.0f;
                ora_mp dyVz = ((Vz[p + sj] - Vz[p]) * r40y - (Vz[p + 2 * sj] - Vz[p - sj]) * r41y +
                               (Vz[p - 1 + sj] - Vz[p - 1]) * r40y - (Vz[p - 1 + 2 * sj] - Vz[p - 1 - sj]) * r41y +
                               (Vz[p] - Vz[p - sj]) * r40y - (Vz[p + sj] - Vz[p - 2 * sj]) * r41y +
                               (Vz[p - 1] - Vz[p - 1 - sj]) * r40y - (Vz[p - 1 + sj] - Vz[p - 1 - 2 * sj]) * r41y) / 4.0f;
                ora_mp dzVx = ((Vx[p + 1] - Vx[p]) * r40z - (Vx[p + 2] - Vx[p - 1]) * r41z +
                               (Vx[p + 1 - si] - Vx[p - si]) * r40z - (Vx[p + 2 - si] - Vx[p - 1 - si]) * r41z +
                               (Vx[p] - Vx[p - 1]) * r40z - (Vx[p + 1] - Vx[p - 2]) * r41z +
                               (Vx[p - si] - Vx[p - 1 - si]) * r40z - (Vx[p + 1 - si] - Vx[p - 2 - si]) * r41z) / 4.0f;
                ora_mp dzVy = ((Vy[p + 1] - Vy[p]) * r40z - (Vy[p + 2] - Vy[p - 1]) * r41z +
                               (Vy[p + 1 - sj] - Vy[p - sj]) * r40z - (Vy[p + 2 - sj] - Vy[p - 1 - sj]) * r41z +
                               (Vy[p] - Vy[p - 1]) * r40z - (Vy[p + 1] - Vy[p - 2]) * r41z +
                               (Vy[p - sj] - Vy[p - 1 - sj]) * r40z - (Vy[p + 1 - sj] - Vy[p - 2 - sj]) * r41z) / 4.0f;
                r->exx[n] = r->exx[n] + (float)(dxVx)*dt;
                r->eyy[n] = r->eyy[n] + (float)(dyVy)*dt;
                r->ezz[n] = r->ezz[n] + (float)(dzVz)*dt;
                r->eyz[n] = r->eyz[n] + (float)(dyVz + dzVy) / 2.0f * dt;
                r->exz[n] = r->exz[n] + (float)(dxVz + dzVx) / 2.0f * dt;
                r->exy[n] = r->exy[n] + (float)(dxVy + dyVx) / 2.0f * dt;
            }
            if (!sample || itw > c->ntw) continue;
            const size_t ntw = (size_t)c->ntw;
            if (c->sw_wav_v) { /* :527-532 */
                float *w = r->wav_vel + ntw * 3 * n + (itw - 1);
                w[0 * ntw] = (float)(Vx[p] + Vx[p - si]) / 2.0f * c->M0 * c->UC * 1e9f;
                w[1 * ntw] = (float)(Vy[p] + Vy[p - sj]) / 2.0f * c->M0 * c->UC * 1e9f;
                w[2 * ntw] = -(float)(Vz[p] + Vz[p - 1]) / 2.0f * c->M0 * c->UC * 1e9f;
            }
            if (c->sw_wav_u) { /* :550-554 */
                float *w = r->wav_disp + ntw * 3 * n + (itw - 1);
                w[0 * ntw] = r->ux[n] * c->M0 * c->UC * 1e9f;
                w[1 * ntw] = r->uy[n] * c->M0 * c->UC * 1e9f;
                w[2 * ntw] = r->uz[n] * c->M0 * c->UC * 1e9f;
            }
            if (c->sw_wav_stress) { /* :572-583 */
                float *w = r->wav_stress + ntw * 6 * n + (itw - 1);
                w[0 * ntw] = (float)(r->Sxx[p]) * c->M0 * c->UC * 1e6f;
                w[1 * ntw] = (float)(r->Syy[p]) * c->M0 * c->UC * 1e6f;
                w[2 * ntw] = (float)(r->Szz[p]) * c->M0 * c->UC * 1e6f;
                w[3 * ntw] = (float)(r->Syz[p] + r->Syz[p - sj] + r->Syz[p - 1] + r->Syz[p - 1 - sj]) / 4.0f * c->M0 * c->UC * 1e6f;
                w[4 * ntw] = (float)(r->Sxz[p] + r->Sxz[p - si] + r->Sxz[p - 1] + r->Sxz[p - 1 - si]) / 4.0f * c->M0 * c->UC * 1e6f;
                w[5 * ntw] = (float)(r->Sxy[p] + r->Sxy[p - sj] + r->Sxy[p - si] + r->Sxy[p - si - sj]) / 4.0f * c->M0 * c->UC * 1e6f;
            }
            if (c->sw_wav_strain) { /* :601-608 */
                float *w = r->wav_strain + ntw * 6 * n + (itw - 1);
                w[0 * ntw] = r->exx[n] * c->M0 * c->UC * 1e-3f;
                w[1 * ntw] = r->eyy[n] * c->M0 * c->UC * 1e-3f;
                w[2 * ntw] = r->ezz[n] * c->M0 * c->UC * 1e-3f;
                w[3 * ntw] = r->eyz[n] * c->M0 * c->UC * 1e-3f;
                w[4 * ntw] = r->exz[n] * c->M0 * c->UC * 1e-3f;
                w[5 * ntw] = r->exy[n] * c->M0 * c->UC * 1e-3f;
            }
        }
    }
}

/* kernel__vmax m_kernel.f90:350-374 + mpi_reduce(MAX) and scaling m_report.f90:140,155 */
void ora_vmax(ora_sim *s, float out[3]) {
    const ora_cfg *c = &s->cfg;
    const int margin = 5;
    float va[3] = {0.0f, 0.0f, 0.0f};
    for (int q = 0; q < s->nranks; q++) {
        ora_rank *r = &s->r[q];
        float xm = 0.0f, ym = 0.0f, zm = 0.0f;
        for (int j = imax(c->na + margin + 1, r->jbeg_k); j <= imin(c->ny - c->na - margin, r->jend_k); j++)
            for (int i = imax(c->na + margin + 1, r->ibeg_k); i <= imin(c->nx - c->na - margin, r->iend_k); i++) {
                size_t n = ora_idx3(r, r->kob[ora_idx2(r, i, j)] + 1, i, j);
                xm = fmaxf(xm, (float)fabs((double)r->Vx[n]));
                ym = fmaxf(ym, (float)fabs((double)r->Vy[n]));
                zm = fmaxf(zm, (float)fabs((double)r->Vz[n]));
            }
        va[0] = fmaxf(va[0], xm);
        va[1] = fmaxf(va[1], ym);
        va[2] = fmaxf(va[2], zm);
    }
    for (int q = 0; q < 3; q++) out[q] = va[q] * c->UC * c->M0;
}

/* one iteration, main.f90:119-139 (report__progress is ora_vmax; snapshots/green are out of scope) */
void ora_step(ora_sim *s, int it) {
    ora_green_store(s, it);
    ora_wav_store(s, it);
    ora_snap_write(s, it);
    ora_update_stress(s);
    ora_stressglut(s, it);
    ora_comm_stress(s);
    ora_update_vel(s);
    ora_bodyforce(s, it);
    ora_green_source(s, it);
    ora_comm_vel(s);
}

int ora_run(ora_sim *s, int it0, int it1, float *vm, int nvm) {
    int nrec = 0;
    for (int it = it0; it <= it1; it++) {
        if (vm && s->cfg.ntdec_r > 0 && it % s->cfg.ntdec_r == 0 && nrec < nvm) {
            ora_vmax(s, vm + 3 * nrec);
            nrec++;
        }
        ora_step(s, it);
    }
    return nrec;
}
