// psv_kernels.cuh -- sm_100a device code of the swpc_psv (2-D P-SV) time step.  Hand-written CUDA, no tensor cores:
// like the 3-D path this is an HBM-bound staggered-grid stencil (212 B per interior cell and step at NM = 3,
// float64 fields).
//
// Arithmetic contract: every temporary keeps the kind it is declared with in src/swpc_psv/*.f90 (F = real(MP), float =
// real(SP)) and every expression the reference's association; built with -fmad=false the results are bit-identical to a
// plain-IEEE evaluation (tests/test_gpu_psv.py).  Note the kinds differ from swpc_3d in places: the shear strain rate
// dxVz_dzVx is real(SP) here (m_kernel.f90:155), the PML 1/d factors are real(SP) (m_absorb_p.f90:54), the PML velocity
// bracket is NOT rounded to single (m_absorb_p.f90:189-195) and the ADE updates are rounded after a mixed-kind sum.
//
// Layout in HBM: idx(k,i) = (k + KOFF - 1) + NZP * (i - ibeg + 3), k fastest as m_kernel.f90:331; KOFF = 32 puts k = 1 of
// every column on a 128-byte boundary, NZP is a multiple of 32, the +-3 columns are the reference's halo/sleeve cells
// (m_global.f90:244-245), cells with k <= 0 or k > nz are real zero cells.  Memory variables are one array per
// (component, mechanism) instead of the reference's (m,k,i) so that each is a coalesced stream.  The 8 ADE arrays exist
// only for absorber cells: column i stores k = kbeg_a(i)..nz at aoff(i), padded so that element k keeps lane (k-1) mod 32.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"   // KOFF, HALO, MAXNM, ldro/lds_/sts_, fd_order_sel, mu_harm, pf_l2, momentrate_dev

namespace swpc {

template <typename F>
struct PsvParams {
    int nz, nxp, NZP, NXM;
    long long ncell;                  // NZP * NXM
    int li0_k, li1_k, k1_k;           // interior kernel box: local columns (inclusive) and kend_k
    int abc;                          // 1 PML, 2 Cerjan
    F *Vx, *Vz, *Sxx, *Szz, *Sxz;
    float *R;                         // 3*NM arrays of ncell floats: R[(c*NM + m)*ncell + idx], c: xx zz xz
    const float *rho, *lam, *mu, *taup, *taus;
    const int4 *band;                 // per memory column mi: kfs_top, kfs_bot, kob_top, kob_bot
    const int *kbeg_a, *kob;          // per memory column
    const long long *aoff;            // per owned column li
    float *aux;                       // 8 arrays of naux floats
    long long naux;
    const float4 *gxc, *gxe, *gzc, *gze;   // per owned column li / per k-1
    const float *cgx_c, *cgx_b, *cgz_c, *cgz_b;   // Cerjan: per memory column / per k + KOFF - 1
    F r40x[2], r41x[2], r40z[2], r41z[2];   // [0] 4th order (isign = -1), [1] 2nd order (isign = +1)
    float r20x, r20z;
    float c1[MAXNM], c2[MAXNM], d1[MAXNM], d2;
    float dt;
};

enum PsvAux { p_axVx = 0, p_azVx, p_axVz, p_azVz, p_axSxx, p_azSxz, p_axSxz, p_azSzz };
enum PsvPhase { PSV_FUSED = 0, PSV_INTERIOR = 1, PSV_ABSORBER = 2 };

// kernel__update_stress, m_kernel.f90:142-311: normal (:179-226) and shear (:262-300) loops in one pass
template <typename F, int NM>
__device__ __forceinline__ void psv_stress_interior(const PsvParams<F> &p, long long n, int k, int mi, bool cerjan) {
    const long long SI = p.NZP;
    const int o = fd_order_sel(k, p.band[mi]);
    const F re40x = p.r40x[o], re41x = p.r41x[o], re40z = p.r40z[o], re41z = p.r41z[o];
    const float dt = p.dt;
    const F *__restrict__ Vx = p.Vx, *__restrict__ Vz = p.Vz;
    const F vx0 = ldro(Vx + n), vz0 = ldro(Vz + n);
    const F dxVx = (vx0 - ldro(Vx + n - SI)) * re40x - (ldro(Vx + n + SI) - ldro(Vx + n - 2 * SI)) * re41x;
    const F dzVz = (vz0 - ldro(Vz + n - 1)) * re40z - (ldro(Vz + n + 1) - ldro(Vz + n - 2)) * re41z;
    const F dxVz = (ldro(Vz + n + SI) - vz0) * re40x - (ldro(Vz + n + 2 * SI) - ldro(Vz + n - SI)) * re41x;
    const F dzVx = (ldro(Vx + n + 1) - vx0) * re40z - (ldro(Vx + n + 2) - ldro(Vx + n - 1)) * re41z;

    const float mu0 = ldro(p.mu + n);
    const float mu2 = 2 * mu0;
    const float lam2mu = lds_(p.lam + n) + mu2;
    const float taup1 = (NM > 0) ? lds_(p.taup + n) : 0.0f, taus1 = (NM > 0) ? lds_(p.taus + n) : 0.0f;
    const float mu_xz = mu_harm(mu0, ldro(p.mu + n + 1), ldro(p.mu + n + SI), ldro(p.mu + n + 1 + SI));

    const float d2v2 = (float)(dxVx + dzVz);
    const float dxVz_dzVx = (float)(dxVz + dzVx);
    float Rxx_n = 0.0f, Rzz_n = 0.0f, Rxz_n = 0.0f;
    float rn[3][NM > 0 ? NM : 1];
    if (NM > 0) {
        const float f_Rxx = (float)(lam2mu * taup1 * d2v2 - mu2 * taus1 * dzVz);
        const float f_Rzz = (float)(lam2mu * taup1 * d2v2 - mu2 * taus1 * dxVx);
        const float f_Rxz = mu_xz * taus1 * dxVz_dzVx;
#pragma unroll
        for (int m = 0; m < NM; m++) {
            const float c1 = p.c1[m], c2 = p.c2[m], d1 = p.d1[m];
            const float nxx = c1 * lds_(p.R + n + (0 * NM + m) * p.ncell) - c2 * f_Rxx * dt;
            const float nzz = c1 * lds_(p.R + n + (1 * NM + m) * p.ncell) - c2 * f_Rzz * dt;
            const float nxz = c1 * lds_(p.R + n + (2 * NM + m) * p.ncell) - c2 * f_Rxz * dt;
            rn[0][m] = nxx; rn[1][m] = nzz; rn[2][m] = nxz;
            Rxx_n = Rxx_n + d1 * nxx; Rzz_n = Rzz_n + d1 * nzz; Rxz_n = Rxz_n + d1 * nxz;
        }
    }
    const float taup_plus1 = 1 + taup1 * (1 + p.d2), taus_plus1 = 1 + taus1 * (1 + p.d2);
    F sxx = lds_(p.Sxx + n) + (lam2mu * taup_plus1 * d2v2 - mu2 * taus_plus1 * dzVz + Rxx_n) * dt;
    F szz = lds_(p.Szz + n) + (lam2mu * taup_plus1 * d2v2 - mu2 * taus_plus1 * dxVx + Rzz_n) * dt;
    F sxz = lds_(p.Sxz + n) + (mu_xz * taus_plus1 * dxVz_dzVx + Rxz_n) * dt;
    if (cerjan) {   // m_absorb_c.f90:98-125
        const int kk = k + KOFF - 1;
        const float gc = p.cgx_c[mi] * p.cgz_c[kk], gb = p.cgx_b[mi] * p.cgz_b[kk];
        sxx = sxx * gc; szz = szz * gc; sxz = sxz * gb;
    }
    if (NM > 0) {
#pragma unroll
        for (int m = 0; m < NM; m++) {
            sts_(p.R + n + (0 * NM + m) * p.ncell, rn[0][m]);
            sts_(p.R + n + (1 * NM + m) * p.ncell, rn[1][m]);
            sts_(p.R + n + (2 * NM + m) * p.ncell, rn[2][m]);
        }
    }
    sts_(p.Sxx + n, sxx); sts_(p.Szz + n, szz); sts_(p.Sxz + n, sxz);
}

// absorb_p__update_stress, m_absorb_p.f90:266-400 (both k-loops of one column)
template <typename F>
__device__ __forceinline__ void psv_stress_pml(const PsvParams<F> &p, long long n, int k, int li, long long a) {
    const long long SI = p.NZP, na = p.naux;
    const float dt = p.dt, r20x = p.r20x, r20z = p.r20z;
    const F *__restrict__ Vx = p.Vx, *__restrict__ Vz = p.Vz;
    const float4 gxc = ldro(p.gxc + li), gxe = ldro(p.gxe + li), gzc = ldro(p.gzc + (k - 1)), gze = ldro(p.gze + (k - 1));
    float *__restrict__ A = p.aux + a;
    const F vx0 = ldro(Vx + n), vz0 = ldro(Vz + n);
    const F dxVx = (vx0 - ldro(Vx + n - SI)) * r20x;
    const F dzVz = (vz0 - ldro(Vz + n - 1)) * r20z;
    const float mu0 = ldro(p.mu + n);
    const float lam2mu_R = ldro(p.lam + n) + 2 * mu0;
    const float lam_R = lam2mu_R - 2 * mu0;
    const float a_xVx = A[p_axVx * na], a_zVz = A[p_azVz * na], a_zVx = A[p_azVx * na], a_xVz = A[p_axVz * na];
    const float dxVx_ade = (float)(gxc.x * dxVx + gxc.y * a_xVx);
    const float dzVz_ade = (float)(gzc.x * dzVz + gzc.y * a_zVz);
    const F sxx = p.Sxx[n] + (lam2mu_R * dxVx_ade + lam_R * dzVz_ade) * dt;
    const F szz = p.Szz[n] + (lam2mu_R * dzVz_ade + lam_R * dxVx_ade) * dt;
    const float n_xVx = (float)(gxc.z * a_xVx + gxc.w * dxVx * dt);
    const float n_zVz = (float)(gzc.z * a_zVz + gzc.w * dzVz * dt);

    const F dzVx = (ldro(Vx + n + 1) - vx0) * r20z;
    const F dxVz = (ldro(Vz + n + SI) - vz0) * r20x;
    const float muxz = mu_harm(mu0, ldro(p.mu + n + 1), ldro(p.mu + n + SI), ldro(p.mu + n + 1 + SI));
    const F sxz = p.Sxz[n] + muxz * (gxe.x * dxVz + gze.x * dzVx + gxe.y * a_xVz + gze.y * a_zVx) * dt;
    const float n_zVx = (float)(gze.z * a_zVx + gze.w * dzVx * dt);
    const float n_xVz = (float)(gxe.z * a_xVz + gxe.w * dxVz * dt);

    p.Sxx[n] = sxx; p.Szz[n] = szz; p.Sxz[n] = sxz;
    A[p_axVx * na] = n_xVx; A[p_azVz * na] = n_zVz; A[p_azVx * na] = n_zVx; A[p_axVz * na] = n_xVz;
}

// kernel__update_vel, m_kernel.f90:76-140 (+ absorb_c__update_vel m_absorb_c.f90:127-151 when fused): the arithmetic, on
// operands the caller fetched up front: issuing all 20 loads of a cell before the first dependent instruction (instead of
// interleaving them with the band lookup and the differences) took the velocity sweep from 1.80 to 1.41 ms at 16384 x 8192
template <typename F>
struct PsvVelOps {
    F sxx_im1, sxx0, sxx_ip1, sxx_ip2;      // Sxx(k, i-1 .. i+2)
    F szz_km1, szz0, szz_kp1, szz_kp2;      // Szz(k-1 .. k+2, i)
    F sxz_im2, sxz_im1, sxz0, sxz_ip1;      // Sxz(k, i-2 .. i+1)
    F sxz_km2, sxz_km1, sxz_kp1;            // Sxz(k-2, k-1, k+1; i)
    float rho0, rho_ip1, rho_kp1;
    F vx, vz;                               // in: old values, out: updated
};
template <typename F>
__device__ __forceinline__ void psv_vel_calc(const PsvParams<F> &p, PsvVelOps<F> &q, int k, int mi, bool cerjan) {
    const int o = fd_order_sel(k, p.band[mi]);
    const F re40x = p.r40x[o], re41x = p.r41x[o], re40z = p.r40z[o], re41z = p.r41z[o];
    const float dt = p.dt;
    const F dxSxx = (q.sxx_ip1 - q.sxx0) * re40x - (q.sxx_ip2 - q.sxx_im1) * re41x;
    const F dzSzz = (q.szz_kp1 - q.szz0) * re40z - (q.szz_kp2 - q.szz_km1) * re41z;
    const F dxSxz = (q.sxz0 - q.sxz_im1) * re40x - (q.sxz_ip1 - q.sxz_im2) * re41x;
    const F dzSxz = (q.sxz0 - q.sxz_km1) * re40z - (q.sxz_kp1 - q.sxz_km2) * re41z;
    const float bx = 2.0f / (q.rho0 + q.rho_ip1);
    const float bz = 2.0f / (q.rho0 + q.rho_kp1);
    F vx = q.vx + bx * (dxSxx + dzSxz) * dt;
    F vz = q.vz + bz * (dxSxz + dzSzz) * dt;
    if (cerjan) {
        const int kk = k + KOFF - 1;
        vx = vx * p.cgx_b[mi] * p.cgz_c[kk];
        vz = vz * p.cgx_c[mi] * p.cgz_b[kk];
    }
    q.vx = vx; q.vz = vz;
}

template <typename F>
__device__ __forceinline__ void psv_vel_interior(const PsvParams<F> &p, long long n, int k, int mi, bool cerjan) {
    const long long SI = p.NZP;
    const F *__restrict__ Sxx = p.Sxx, *__restrict__ Szz = p.Szz, *__restrict__ Sxz = p.Sxz;
    PsvVelOps<F> q;
    q.sxx0 = ldro(Sxx + n); q.szz0 = ldro(Szz + n); q.sxz0 = ldro(Sxz + n);
    q.sxx_ip1 = ldro(Sxx + n + SI); q.sxx_ip2 = ldro(Sxx + n + 2 * SI); q.sxx_im1 = ldro(Sxx + n - SI);
    q.szz_kp1 = ldro(Szz + n + 1); q.szz_kp2 = ldro(Szz + n + 2); q.szz_km1 = ldro(Szz + n - 1);
    q.sxz_im1 = ldro(Sxz + n - SI); q.sxz_ip1 = ldro(Sxz + n + SI); q.sxz_im2 = ldro(Sxz + n - 2 * SI);
    q.sxz_km1 = ldro(Sxz + n - 1); q.sxz_kp1 = ldro(Sxz + n + 1); q.sxz_km2 = ldro(Sxz + n - 2);
    q.rho0 = ldro(p.rho + n); q.rho_ip1 = ldro(p.rho + n + SI); q.rho_kp1 = ldro(p.rho + n + 1);
    q.vx = lds_(p.Vx + n); q.vz = lds_(p.Vz + n);
    psv_vel_calc<F>(p, q, k, mi, cerjan);
    sts_(p.Vx + n, q.vx); sts_(p.Vz + n, q.vz);
}

// absorb_p__update_vel, m_absorb_p.f90:157-201
template <typename F>
__device__ __forceinline__ void psv_vel_pml(const PsvParams<F> &p, long long n, int k, int li, long long a) {
    const long long SI = p.NZP, na = p.naux;
    const float dt = p.dt, r20x = p.r20x, r20z = p.r20z;
    const F *__restrict__ Sxx = p.Sxx, *__restrict__ Szz = p.Szz, *__restrict__ Sxz = p.Sxz;
    const float4 gxc = ldro(p.gxc + li), gxe = ldro(p.gxe + li), gzc = ldro(p.gzc + (k - 1)), gze = ldro(p.gze + (k - 1));
    float *__restrict__ A = p.aux + a;
    const F sxz0 = ldro(Sxz + n);
    const F dxSxx = (ldro(Sxx + n + SI) - ldro(Sxx + n)) * r20x;
    const F dzSzz = (ldro(Szz + n + 1) - ldro(Szz + n)) * r20z;
    const F dxSxz = (sxz0 - ldro(Sxz + n - SI)) * r20x;
    const F dzSxz = (sxz0 - ldro(Sxz + n - 1)) * r20z;
    const float rho0 = ldro(p.rho + n);
    const float bx = 2.0f / (rho0 + ldro(p.rho + n + SI));
    const float bz = 2.0f / (rho0 + ldro(p.rho + n + 1));
    const float a_xSxx = A[p_axSxx * na], a_zSxz = A[p_azSxz * na], a_xSxz = A[p_axSxz * na], a_zSzz = A[p_azSzz * na];
    p.Vx[n] = p.Vx[n] + bx * (gxe.x * dxSxx + gzc.x * dzSxz + gxe.y * a_xSxx + gzc.y * a_zSxz) * dt;
    p.Vz[n] = p.Vz[n] + bz * (gxc.x * dxSxz + gze.x * dzSzz + gxc.y * a_xSxz + gze.y * a_zSzz) * dt;
    A[p_axSxx * na] = (float)(gxe.z * a_xSxx + gxe.w * dxSxx * dt);
    A[p_azSxz * na] = (float)(gzc.z * a_zSxz + gzc.w * dzSxz * dt);
    A[p_axSxz * na] = (float)(gxc.z * a_xSxz + gxc.w * dxSxz * dt);
    A[p_azSzz * na] = (float)(gze.z * a_zSzz + gze.w * dzSzz * dt);
}

// L2 prefetch of what the cell `pf` columns ahead will stream from HBM
template <typename F, int NM, bool STRESS>
__device__ __forceinline__ void psv_prefetch(const PsvParams<F> &p, long long n, bool pml_target, long long a) {
    const long long SI = p.NZP;
    if (STRESS) {
        pf_l2(p.Sxx + n); pf_l2(p.Szz + n); pf_l2(p.Sxz + n);
        pf_l2(p.lam + n); pf_l2(p.mu + n + SI); pf_l2(p.Vx + n + SI); pf_l2(p.Vz + n + 2 * SI);
        if (pml_target) {
#pragma unroll
            for (int q = 0; q < 4; q++) pf_l2(p.aux + a + q * p.naux);
        } else {
            if (NM > 0) { pf_l2(p.taup + n); pf_l2(p.taus + n); }
#pragma unroll
            for (int q = 0; q < 3 * NM; q++) pf_l2(p.R + n + q * p.ncell);
        }
    } else {
        pf_l2(p.Vx + n); pf_l2(p.Vz + n); pf_l2(p.rho + n + SI);
        pf_l2(p.Szz + n); pf_l2(p.Sxx + n + 2 * SI); pf_l2(p.Sxz + n + SI);
        if (pml_target) {
#pragma unroll
            for (int q = 4; q < 8; q++) pf_l2(p.aux + a + q * p.naux);
        }
    }
}

// One sweep over owned columns li0..li1: thread = one k, block = TK consecutive k marching over `ilen` columns along i
// (x is the slow axis of the (k,i) layout, so every load of a warp is one or two full 128-byte lines and the +-2 column
// neighbours of the stencil were touched by the same block one or two iterations earlier: they are L1/L2 hits).
// phase: PSV_FUSED = interior + absorber (PML update or Cerjan multiply of the fresh value), PSV_INTERIOR = interior cells
// only, no Cerjan multiply, PSV_ABSORBER = PML cells only / Cerjan multiply only.
template <typename F, int NM, bool STRESS>
__global__ void __launch_bounds__(256, 3) psv_sweep(const __grid_constant__ PsvParams<F> p, int phase, int li0, int li1, int ilen, int pf, int k0) {
    const int k = k0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k > p.nz) return;
    const int lis = li0 + blockIdx.y * ilen;
    const int lie = min(lis + ilen, li1 + 1);
    const bool pml_mode = (p.abc == 1);
    for (int li = lis; li < lie; li++) {
        const int mi = li + HALO;
        const long long n = (long long)(k + KOFF - 1) + (long long)p.NZP * mi;
        if (pf > 0 && li + pf < lie) {
            bool pt = false;
            long long ap = 0;
            if (pml_mode) {
                const int kb = p.kbeg_a[mi + pf];
                pt = (k >= kb);
                if (pt) ap = p.aoff[li + pf] + (k - kb);
            }
            if (phase != PSV_ABSORBER || pt) psv_prefetch<F, NM, STRESS>(p, n + (long long)p.NZP * pf, pt, ap);
        }
        const bool is_pml = pml_mode && (k >= p.kbeg_a[mi]);
        if (is_pml) {
            if (phase != PSV_INTERIOR) {
                const long long a = p.aoff[li] + (k - p.kbeg_a[mi]);
                if (STRESS) psv_stress_pml<F>(p, n, k, li, a);
                else psv_vel_pml<F>(p, n, k, li, a);
            }
        } else if (phase != PSV_ABSORBER) {
            if (li >= p.li0_k && li <= p.li1_k && k <= p.k1_k) {
                if (STRESS) psv_stress_interior<F, NM>(p, n, k, mi, p.abc == 2 && phase == PSV_FUSED);
                else psv_vel_interior<F>(p, n, k, mi, p.abc == 2 && phase == PSV_FUSED);
            }
        } else if (p.abc == 2) {   // Cerjan multiply on its own (m_absorb_c.f90:98-151)
            const int kk = k + KOFF - 1;
            if (STRESS) {
                const float gc = p.cgx_c[mi] * p.cgz_c[kk], gb = p.cgx_b[mi] * p.cgz_b[kk];
                p.Sxx[n] = p.Sxx[n] * gc; p.Szz[n] = p.Szz[n] * gc; p.Sxz[n] = p.Sxz[n] * gb;
            } else {
                p.Vx[n] = p.Vx[n] * p.cgx_b[mi] * p.cgz_c[kk];
                p.Vz[n] = p.Vz[n] * p.cgx_c[mi] * p.cgz_b[kk];
            }
        }
    }
}

// snap__write, m_snap.f90:435-650: decimated xz slices.  One thread per snapshot node (ii, kk) of this rank's region;
// ps = masked divergence / rotation (:468-494), v = Vx, Vz (:550-551), u = running displacement, accumulated every step
// (:609-610).  Buffers are the reference's buf(nxs, nzs, 2) over the WHOLE snapshot grid (zero outside the rank's region,
// summed onto the I/O rank afterwards).
struct PsvSnap {
    int idec, kdec, nxs, nzs, is0, is1, ks0, ks1, ibeg;
    int do_ps, do_v, do_u;          // ps / v only on sampled steps, u every step
    float UC, M0;
    double r20x, r20z;              // 1/dx, 1/dz in the field kind
    float *buf_ps, *buf_v, *buf_u;
};

template <typename F>
__global__ void psv_snap_kernel(const __grid_constant__ PsvParams<F> p, const PsvSnap g) {
    const int kk = g.ks0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int ii = g.is0 + blockIdx.y;
    if (kk > g.ks1 || ii > g.is1) return;
    const int k = kk * g.kdec - g.kdec / 2, i = ii * g.idec - g.idec / 2;
    const long long SI = p.NZP;
    const long long n = (long long)(k + KOFF - 1) + SI * (i - g.ibeg + HALO);
    const long long o = (long long)(ii - 1) + (long long)g.nxs * (kk - 1), n2 = (long long)g.nxs * g.nzs;
    const F *Vx = p.Vx, *Vz = p.Vz;
    if (g.do_ps) {
        const F r20x = (F)g.r20x, r20z = (F)g.r20z;
        float div = (float)((Vx[n] - Vx[n - SI]) * r20x + (Vz[n] - Vz[n - 1]) * r20z);
        float rot = (float)((Vx[n + 1] - Vx[n]) * r20z - (Vz[n + SI] - Vz[n]) * r20x);
        const float mu_xz = mu_harm(p.mu[n], p.mu[n + 1], p.mu[n + SI], p.mu[n + 1 + SI]);
        const float lam0 = p.lam[n];
        div = div * lam0 / (fabsf(lam0) + FLT_EPS_);
        rot = rot * mu_xz / fabsf(mu_xz + FLT_EPS_);
        g.buf_ps[o] = div * g.UC * g.M0 * 1e-3f;
        g.buf_ps[n2 + o] = rot * g.UC * g.M0 * 1e-3f;
    }
    if (g.do_v) {
        g.buf_v[o] = (float)(Vx[n] * g.UC * g.M0);
        g.buf_v[n2 + o] = (float)(Vz[n] * g.UC * g.M0);
    }
    if (g.do_u) {
        g.buf_u[o] = (float)(g.buf_u[o] + Vx[n] * g.UC * g.M0 * p.dt);
        g.buf_u[n2 + o] = (float)(g.buf_u[n2 + o] + Vz[n] * g.UC * g.M0 * p.dt);
    }
}

// Plane-wave mode: horizontal zero-derivative boundary (m_absorb_p.f90:114-155 / :287-326): linear extrapolation into the
// first column outside the model on the outer ranks, k = 1..nz.  dst / s1 / s2 are memory columns; side < 0 = off.
template <typename F>
__global__ void psv_pw_edge_kernel(F *f0, F *f1, F *f2, int nz, int NZP, int dstL, int dstR) {
    const int k = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nz) return;
    const int side = blockIdx.y;
    const int dst = side == 0 ? dstL : dstR;
    if (dst < 0) return;
    const int s1 = side == 0 ? dst + 1 : dst - 1, s2 = side == 0 ? dst + 2 : dst - 2;
    const long long kk = k + KOFF - 1, d = kk + (long long)NZP * dst, a = kk + (long long)NZP * s1, b = kk + (long long)NZP * s2;
    f0[d] = 2 * f0[a] - f0[b];
    f1[d] = 2 * f1[a] - f1[b];
    if (f2) f2[d] = 2 * f2[a] - f2[b];
}

struct PsvSrc {
    int nsrc;
    const int *ik;          // 2*nsrc: memory column mi and k
    const double *mo;       // nsrc
    const double *m3;       // 3*nsrc: mxx mzz mxz (body force: fx fz -)
    const float *prm;       // 2*nsrc
    int stf;
    float t;
    double dt_dxz;
};

// source__stressglut m_source.f90:566-583
template <typename F>
__global__ void psv_stressglut_kernel(const __grid_constant__ PsvParams<F> p, const PsvSrc s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.nsrc) return;
    const F stime = (F)momentrate_dev(s.t, s.stf, s.prm[2 * i], s.prm[2 * i + 1]);
    const F sdrop = (F)s.mo[i] * stime * (F)s.dt_dxz;
    const long long SI = p.NZP;
    const long long n = (long long)(s.ik[2 * i + 1] + KOFF - 1) + SI * s.ik[2 * i];
    const F mxx = (F)s.m3[3 * i], mzz = (F)s.m3[3 * i + 1], mxz = (F)s.m3[3 * i + 2];
    atomicAdd(p.Sxx + n, -(mxx * sdrop));
    atomicAdd(p.Szz + n, -(mzz * sdrop));
    const F q = mxz * sdrop / 4;
    atomicAdd(p.Sxz + n, -q); atomicAdd(p.Sxz + n - 1, -q); atomicAdd(p.Sxz + n - SI, -q); atomicAdd(p.Sxz + n - 1 - SI, -q);
}

// source__bodyforce m_source.f90:607-623 (the same rho(kk,ii) in all four terms)
template <typename F>
__global__ void psv_bodyforce_kernel(const __grid_constant__ PsvParams<F> p, const PsvSrc s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.nsrc) return;
    const F stime = (F)momentrate_dev(s.t, s.stf, s.prm[2 * i], s.prm[2 * i + 1]);
    const long long SI = p.NZP;
    const long long n = (long long)(s.ik[2 * i + 1] + KOFF - 1) + SI * s.ik[2 * i];
    const F fx = (F)s.m3[3 * i], fz = (F)s.m3[3 * i + 1], dtd = (F)s.dt_dxz;
    const float rho = p.rho[n];
    const F ax = fx / rho * stime * dtd / 2, az = fz / rho * stime * dtd / 2;
    atomicAdd(p.Vx + n, ax); atomicAdd(p.Vx + n - SI, ax);
    atomicAdd(p.Vz + n, az); atomicAdd(p.Vz + n - 1, az);
}

struct PsvWav {
    int nst, ntw, itw, sample;
    int sw_v, sw_u, sw_stress, sw_strain;
    const int *ik;
    float *wav_v, *wav_u, *wav_s, *wav_e;   // (ntw,2,nst) (ntw,2,nst) (ntw,3,nst) (ntw,3,nst)
    float *acc;                             // 5 running sums per station: ux uz exx ezz exz
    float M0, UC;
    double r40x, r40z, r41x, r41z;
};

// wav__store m_wav.f90:143-306
template <typename F>
__global__ void psv_wav_kernel(const __grid_constant__ PsvParams<F> p, const PsvWav w) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= w.nst) return;
    const long long SI = p.NZP;
    const long long n = (long long)(w.ik[2 * s + 1] + KOFF - 1) + SI * w.ik[2 * s];
    const F *Vx = p.Vx, *Vz = p.Vz;
    const float dt = p.dt;
    float *a = w.acc + 5 * s;
    if (w.sw_u) {
        a[0] = a[0] + (float)(Vx[n] + Vx[n - SI]) * 0.5f * dt;
        a[1] = a[1] - (float)(Vz[n] + Vz[n - 1]) * 0.5f * dt;
    }
    if (w.sw_strain) {
        const F r40x = (F)w.r40x, r40z = (F)w.r40z, r41x = (F)w.r41x, r41z = (F)w.r41z;
        const F dxVx = (Vx[n] - Vx[n - SI]) * r40x - (Vx[n + SI] - Vx[n - 2 * SI]) * r41x;
        const F dzVz = (Vz[n] - Vz[n - 1]) * r40z - (Vz[n + 1] - Vz[n - 2]) * r41z;
        const F dxVz = ((Vz[n + SI] - Vz[n]) * r40x - (Vz[n + 2 * SI] - Vz[n - SI]) * r41x
                        + (Vz[n - 1 + SI] - Vz[n - 1]) * r40x - (Vz[n - 1 + 2 * SI] - Vz[n - 1 - SI]) * r41x
                        + (Vz[n] - Vz[n - SI]) * r40x - (Vz[n + SI] - Vz[n - 2 * SI]) * r41x
                        + (Vz[n - 1] - Vz[n - 1 - SI]) * r40x - (Vz[n - 1 + SI] - Vz[n - 1 - 2 * SI]) * r41x) / 4.0f;
        const F dzVx = ((Vx[n + 1] - Vx[n]) * r40z - (Vx[n + 2] - Vx[n - 1]) * r41z
                        + (Vx[n + 1 - SI] - Vx[n - SI]) * r40z - (Vx[n + 2 - SI] - Vx[n - 1 - SI]) * r41z
                        + (Vx[n] - Vx[n - 1]) * r40z - (Vx[n + 1] - Vx[n - 2]) * r41z
                        + (Vx[n - SI] - Vx[n - 1 - SI]) * r40z - (Vx[n + 1 - SI] - Vx[n - 2 - SI]) * r41z) / 4.0f;
        a[2] = a[2] + (float)(dxVx) * dt;
        a[3] = a[3] + (float)(dzVz) * dt;
        a[4] = a[4] + (float)(dxVz + dzVx) / 2.0f * dt;
    }
    if (!w.sample) return;
    const long long ntw = w.ntw;
    const float M0 = w.M0, UC = w.UC;
    if (w.sw_v) {
        float *o = w.wav_v + ntw * 2 * s + (w.itw - 1);
        o[0] = (float)(Vx[n] + Vx[n - SI]) / 2.0f * M0 * UC * 1e9f;
        o[ntw] = -(float)(Vz[n] + Vz[n - 1]) / 2.0f * M0 * UC * 1e9f;
    }
    if (w.sw_u) {
        float *o = w.wav_u + ntw * 2 * s + (w.itw - 1);
        o[0] = a[0] * M0 * UC * 1e9f;
        o[ntw] = a[1] * M0 * UC * 1e9f;
    }
    if (w.sw_stress) {
        float *o = w.wav_s + ntw * 3 * s + (w.itw - 1);
        o[0] = (float)(p.Sxx[n]) * M0 * UC * 1e6f;
        o[ntw] = (float)(p.Szz[n]) * M0 * UC * 1e6f;
        o[2 * ntw] = (float)(p.Sxz[n] + p.Sxz[n - SI] + p.Sxz[n - 1] + p.Sxz[n - 1 - SI]) / 4.0f * M0 * UC * 1e6f;
    }
    if (w.sw_strain) {
        float *o = w.wav_e + ntw * 3 * s + (w.itw - 1);
        o[0] = a[2] * M0 * UC * 1e-3f;
        o[ntw] = a[3] * M0 * UC * 1e-3f;
        o[2 * ntw] = a[4] * M0 * UC * 1e-3f;
    }
}

// kernel__vmax m_kernel.f90:313-327 over local columns li0..li1
template <typename F>
__global__ void psv_vmax_kernel(const __grid_constant__ PsvParams<F> p, int li0, int li1, unsigned int *out2) {
    float xm = 0.0f, zm = 0.0f;
    for (int li = li0 + blockIdx.x * blockDim.x + threadIdx.x; li <= li1; li += gridDim.x * blockDim.x) {
        const int mi = li + HALO;
        const long long n = (long long)(p.kob[mi] + 1 + KOFF - 1) + (long long)p.NZP * mi;
        xm = fmaxf(xm, fabsf((float)p.Vx[n]));
        zm = fmaxf(zm, fabsf((float)p.Vz[n]));
    }
    for (int o = 16; o > 0; o >>= 1) {
        xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, o));
        zm = fmaxf(zm, __shfl_xor_sync(0xffffffffu, zm, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(out2 + 0, __float_as_uint(xm));
        atomicMax(out2 + 1, __float_as_uint(zm));
    }
}

// halo columns, m_global.f90:312-418: three (field, column) pairs per direction, k = 1..nz
struct PsvCols { void *field[3]; int mi[3]; };

template <typename F, bool PACK>
__global__ void psv_halo_kernel(int nz, int NZP, const PsvCols c, F *buf) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;   // 0-based
    const int s = blockIdx.y;
    if (k >= nz) return;
    F *f = (F *)c.field[s];
    const long long n = (long long)(k + KOFF) + (long long)NZP * c.mi[s];
    if (PACK) buf[(long long)s * nz + k] = f[n];
    else f[n] = buf ? buf[(long long)s * nz + k] : F(0);   // no buffer: the never-written receive buffer of an MPI_PROC_NULL neighbour
}

}   // namespace swpc
