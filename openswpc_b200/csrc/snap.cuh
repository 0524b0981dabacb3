// snap.cuh -- decimated 2-D snapshot slices of m_snap.f90 (15 products: sections xy / xz / yz / fs / ob times
// types ps / v / u), gathered on the device into the reference's buffers buf(n1, n2, nvar).
//
//   type v  : buf = V * UC * M0                                   (m_snap.f90:1476-1478, 1546-1548, ...)
//   type u  : buf = buf + V * UC * M0 * dt, EVERY step            (:1857-1859, 1923-1925, ...)
//   type ps : div / rot with the lam / shear-stress masks         (:1016-1036, ...)
//   fs / ob : k = kfs(i,j)+1 / kob(i,j)+1; v and u also keep running maxima max-V/H/A, every step (:1704-1706)
// Off the roofline path (2-D, every ntdec_s steps); kept bit-exact like everything else.
#pragma once

#include "kernels.cuh"

namespace swpc {

enum SnapSection { SEC_XY = 0, SEC_XZ = 1, SEC_YZ = 2, SEC_FS = 3, SEC_OB = 4 };
enum SnapType { TYP_PS = 0, TYP_V = 1, TYP_U = 2 };

struct SnapGeom {
    int idec, jdec, kdec;
    int nxs, nys, nzs;
    int is0, is1, js0, js1, ks0, ks1;   // slice index ranges covered by this rank (1-based, inclusive)
    int k0_xy, i0_yz, j0_xz;            // section positions (global indices)
    int ibeg, jbeg;                     // first owned global i / j
    float UC, M0;
    const int *kfs, *kob;               // (mi, mj) maps
};

template <typename F>
__global__ void snap_kernel(const __grid_constant__ KParams<F> p, const SnapGeom g, int section, int type, float *buf, float *maxbuf) {
    // local slice coordinates (a, b): xy/fs/ob -> (ii, jj); xz -> (ii, kk); yz -> (jj, kk)
    const int a0 = (section == SEC_YZ) ? g.js0 : g.is0, a1 = (section == SEC_YZ) ? g.js1 : g.is1;
    const int b0 = (section == SEC_XZ || section == SEC_YZ) ? g.ks0 : g.js0, b1 = (section == SEC_XZ || section == SEC_YZ) ? g.ks1 : g.js1;
    const int n1 = (section == SEC_YZ) ? g.nys : g.nxs;
    const int n2 = (section == SEC_XZ || section == SEC_YZ) ? g.nzs : g.nys;
    const int a = a0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int b = b0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (a > a1 || b > b1) return;
    int i, j, k;
    if (section == SEC_XZ) { i = a * g.idec - g.idec / 2; j = g.j0_xz; k = b * g.kdec - g.kdec / 2; }
    else if (section == SEC_YZ) { i = g.i0_yz; j = a * g.jdec - g.jdec / 2; k = b * g.kdec - g.kdec / 2; }
    else { i = a * g.idec - g.idec / 2; j = b * g.jdec - g.jdec / 2; k = g.k0_xy; }
    const int mi = i - g.ibeg + HALO, mj = j - g.jbeg + HALO;
    const long long col = (long long)mi + (long long)p.NXM * mj;
    if (section == SEC_FS) k = g.kfs[col] + 1;
    if (section == SEC_OB) k = g.kob[col] + 1;
    const long long n = (long long)(k + KOFF - 1) + (long long)p.NZP * col;
    const long long si = p.SI, sj = p.SJ;
    const long long plane = (long long)n1 * n2;
    float *o = buf + (long long)(a - 1) + (long long)n1 * (b - 1);
    const F *Vx = p.Vx, *Vy = p.Vy, *Vz = p.Vz;
    const float UC = g.UC, M0 = g.M0;

    if (type == TYP_V) {
        o[0] = (float)(Vx[n] * UC * M0);
        o[plane] = (float)(Vy[n] * UC * M0);
        o[2 * plane] = (float)(Vz[n] * UC * M0);
    } else if (type == TYP_U) {
        const float dt = p.dt;
        o[0] = (float)(o[0] + Vx[n] * UC * M0 * dt);
        o[plane] = (float)(o[plane] + Vy[n] * UC * M0 * dt);
        o[2 * plane] = (float)(o[2 * plane] + Vz[n] * UC * M0 * dt);
    } else {
        const F r20x = p.r20x, r20y = p.r20y, r20z = p.r20z;
        float div = (float)((Vx[n] - Vx[n - si]) * r20x + (Vy[n] - Vy[n - sj]) * r20y + (Vz[n] - Vz[n - 1]) * r20z);
        float rot_x = (float)((Vz[n + sj] - Vz[n]) * r20y - (Vy[n + 1] - Vy[n]) * r20z);
        float rot_y = (float)((Vx[n + 1] - Vx[n]) * r20z - (Vz[n + si] - Vz[n]) * r20x);
        float rot_z = (float)((Vy[n + si] - Vy[n]) * r20x - (Vx[n + sj] - Vx[n]) * r20y);
        const float lam = p.lam[n];
        const float eps = FLT_EPS_;
        div = div * lam / fabsf(lam + eps);
        rot_x = (float)(rot_x * fabs(p.Syz[n]) / fabs(p.Syz[n] + eps));
        rot_y = (float)(rot_y * fabs(p.Sxz[n]) / fabs(p.Sxz[n] + eps));
        rot_z = (float)(rot_z * fabs(p.Sxy[n]) / fabs(p.Sxy[n] + eps));
        o[0] = div * UC * M0 * 1e-3f;
        o[plane] = rot_x * UC * M0 * 1e-3f;
        o[2 * plane] = rot_y * UC * M0 * 1e-3f;
        o[3 * plane] = rot_z * UC * M0 * 1e-3f;
    }
    if (maxbuf && type != TYP_PS) {
        float *m = maxbuf + (long long)(a - 1) + (long long)n1 * (b - 1);
        const float b1v = o[0], b2v = o[plane], b3v = o[2 * plane];
        const float m1 = fmaxf(m[0], fabsf(b3v));
        const float m2 = fmaxf(m[plane], sqrtf(b1v * b1v + b2v * b2v));
        m[0] = m1;
        m[plane] = m2;
        m[2 * plane] = sqrtf(m1 * m1 + m2 * m2);
    }
}

}   // namespace swpc
