// kernels.cuh -- sm_100a device code of the swpc_3d time step (hand-written, no tensor cores: the
// path is an HBM-bound staggered-grid stencil).
//
// Arithmetic contract (parity with the reference): every temporary keeps the kind
// it is declared with in the Fortran source -- F (= real(MP): double by default, float for MP=SP)
// for field differences, float for medium / memory-variable / PML terms -- and every expression is
// written in the reference's evaluation order.  Built with -fmad=false the results are bit-identical
// to a plain-IEEE evaluation of those expressions (see tests/test_gpu_parity.py).
//
// Layout in HBM (all arrays): (k,i,j) with k fastest, as m_kernel.f90:380, but with device padding:
//   idx(k,i,j) = (k + KOFF - 1) + NZP * ((i - ibeg + 3) + NXM * (j - jbeg + 3))
// KOFF = 32 puts k = 1 of every column on a 128-byte boundary for float and double arrays; NZP is a
// multiple of 32; the +-3 i/j margins are the reference's halo/sleeve cells (m_global.f90:295-298).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace swpc {

constexpr int KOFF = 32;   // leading k padding (elements)
constexpr int HALO = 3;    // i/j margin cells each side (m_global.f90:295-298)
constexpr int MAXNM = 3;   // NM of the reference build (m_global.f90:31)

struct Band { int fs_top, fs_bot, ob_top, ob_bot; };   // kfs_top, kfs_bot, kob_top, kob_bot of one column

template <typename F>
struct KParams {
    // ---- geometry (local = index inside the owned box, 0-based)
    int nz, nxp, nyp;
    int NZP, NXM, NYM;
    long long SI, SJ;                 // element strides of i and j
    long long ncell;                  // NZP*NXM*NYM
    int li0_k, li1_k, lj0_k, lj1_k;   // interior kernel box, local, inclusive (empty if li0_k > li1_k)
    int k1_k;                         // kend_k (kbeg_k is always 1)
    int abc;                          // 1 = PML, 2 = Cerjan
    // ---- fields
    F *Vx, *Vy, *Vz, *Sxx, *Syy, *Szz, *Syz, *Sxz, *Sxy;
    float *R;                         // 6*NM arrays of ncell floats: R[(c*NM+m)*ncell + idx], c: xx yy zz yz xz xy
    const float *rho, *lam, *mu, *taup, *taus;
    const int4 *band;                 // per (mi,mj) over the memory box: NXM*NYM
    const int *kbeg_a;                // per (mi,mj)
    const int *kob;                   // per (mi,mj)
    // ---- PML
    const long long *aoff;            // per owned column (li + nxp*lj): start of the column in the aux arrays
    float *aux;                       // 18 arrays of naux floats
    long long naux;
    const float4 *gxc, *gxe, *gyc, *gye, *gzc, *gze;   // g(1:4) per owned i / j / k
    // ---- Cerjan (indexed by memory-box index mi / mj, and k + KOFF - 1)
    const float *cgx_c, *cgx_b, *cgy_c, *cgy_b, *cgz_c, *cgz_b;
    // ---- coefficients
    F r40x[2], r41x[2], r40y[2], r41y[2], r40z[2], r41z[2];   // [0]: 4th order (isign=-1), [1]: 2nd order (isign=+1)
    F r20x, r20y, r20z;
    float c1[MAXNM], c2[MAXNM], d1[MAXNM], d2;
    float dt;
};

// aux array order (m_absorb_p.f90:47-52)
enum Aux { axVx = 0, ayVx, azVx, axVy, ayVy, azVy, axVz, ayVz, azVz,
           axSxx, aySxy, azSxz, axSxy, aySyy, azSyz, axSxz, aySyz, azSzz };

constexpr float FLT_EPS_ = 1.1920929e-07f;   // epsilon(1.0)

template <typename T> __device__ __forceinline__ T ldro(const T *p) { return __ldg(p); }
// streamed (touched once per sweep) data: evict-first so that L1/L2 keep the stencil neighbours
#ifndef SWPC_STREAM_HINT
#define SWPC_STREAM_HINT 1
#endif
template <typename T> __device__ __forceinline__ T lds_(const T *p) {
#if SWPC_STREAM_HINT
    return __ldcs(p);
#else
    return *p;
#endif
}
template <typename T> __device__ __forceinline__ void sts_(T *p, T v) {
#if SWPC_STREAM_HINT
    __stcs(p, v);
#else
    *p = v;
#endif
}

// m_kernel.f90:103-104: isign = sign(1, max((k-kfs_top)(kfs_bot-k), (k-kob_top)(kob_bot-k)))
__device__ __forceinline__ int fd_order_sel(int k, const int4 b) {
    int a = (k - b.x) * (b.y - k);
    int c = (k - b.z) * (b.w - k);
    return (max(a, c) >= 0) ? 1 : 0;
}

// m_kernel.f90:293-309
__device__ __forceinline__ float mu_harm(float a, float b, float c, float d) {
    return 4 * a * b * c * d / (a * b * c + a * b * d + a * c * d + b * c * d + FLT_EPS_);
}

// ------------------------------------------------------------------------------------------------
// stress update of one interior cell: kernel__update_stress, m_kernel.f90:179-242 (normal) and
// :266-337 (shear) in ONE pass so that V and mu are read once; Cerjan multiply (m_absorb_c.f90:113-168)
// fused as a post-multiply of the freshly updated value.
//
// The arithmetic is written once, against an accessor `A` that says where the operands live:
//   AccDirect  -- global memory through the read-only / streaming paths (sweep_direct)
//   AccTma     -- shared-memory tiles staged by TMA (stress_tma), results stored straight to global
// V<f,dk,di,dj>: field f (0 Vx, 1 Vy, 2 Vz) at (k+dk, i+di, j+dj); S/R component order xx yy zz yz xz xy.
template <typename F, int NM>
struct AccDirect {
    const KParams<F> &p;
    long long n;
    __device__ __forceinline__ AccDirect(const KParams<F> &p_, long long n_) : p(p_), n(n_) {}
    template <int f, int dk, int di, int dj> __device__ __forceinline__ F V() const {
        const F *b = (f == 0) ? p.Vx : (f == 1) ? p.Vy : p.Vz;
        return ldro(b + n + dk + di * p.SI + dj * p.SJ);
    }
    template <int dk, int di, int dj> __device__ __forceinline__ float mu() const { return ldro(p.mu + n + dk + di * p.SI + dj * p.SJ); }
    __device__ __forceinline__ float lam() const { return ldro(p.lam + n); }
    __device__ __forceinline__ float taup() const { return ldro(p.taup + n); }
    __device__ __forceinline__ float taus() const { return ldro(p.taus + n); }
    __device__ __forceinline__ F *sptr(int c) const { return c == 0 ? p.Sxx : c == 1 ? p.Syy : c == 2 ? p.Szz : c == 3 ? p.Syz : c == 4 ? p.Sxz : p.Sxy; }
    __device__ __forceinline__ F S(int c) const { return lds_(sptr(c) + n); }
    __device__ __forceinline__ void setS(int c, F v) const { sts_(sptr(c) + n, v); }
    __device__ __forceinline__ float R(int q) const { return lds_(p.R + n + q * p.ncell); }
    __device__ __forceinline__ void setR(int q, float v) const { sts_(p.R + n + q * p.ncell, v); }
};

// DO_N: normal components (m_kernel.f90:179-242); DO_S: shear components (:266-337).  The two halves share no
// intermediate, so they may run in different warps (stress_tma) or back to back (sweep_direct) with identical results.
template <typename F, int NM, typename A, bool DO_N = true, bool DO_S = true>
__device__ __forceinline__ void stress_interior_t(const KParams<F> &p, const A &a, int k, int mi, int mj, const int4 bnd) {
    const int o = fd_order_sel(k, bnd);
    const F re40x = p.r40x[o], re41x = p.r41x[o], re40y = p.r40y[o], re41y = p.r41y[o], re40z = p.r40z[o], re41z = p.r41z[o];
    const float dt = p.dt;
    const float taus1 = (NM > 0) ? a.taus() : 0.0f;
    const float taus_plus1 = 1 + taus1 * (1 + p.d2);
    const float mu0 = a.template mu<0, 0, 0>();
    float gxc = 1.0f, gxb = 1.0f, gyc = 1.0f, gyb = 1.0f, gzc = 1.0f, gzb = 1.0f;
    const bool cerjan = (p.abc == 2);   // Cerjan sponge, m_absorb_c.f90:129-133, :155-157
    if (cerjan) {
        const int kk = k + KOFF - 1;
        gxc = p.cgx_c[mi]; gxb = p.cgx_b[mi]; gyc = p.cgy_c[mj]; gyb = p.cgy_b[mj]; gzc = p.cgz_c[kk]; gzb = p.cgz_b[kk];
    }
    F sn[6];
    float rn[6][NM > 0 ? NM : 1];

    if (DO_N) {
        const F vx0 = a.template V<0, 0, 0, 0>(), vy0 = a.template V<1, 0, 0, 0>(), vz0 = a.template V<2, 0, 0, 0>();
        const F vx_im1 = a.template V<0, 0, -1, 0>(), vx_ip1 = a.template V<0, 0, 1, 0>(), vx_im2 = a.template V<0, 0, -2, 0>();
        const F vy_jm1 = a.template V<1, 0, 0, -1>(), vy_jp1 = a.template V<1, 0, 0, 1>(), vy_jm2 = a.template V<1, 0, 0, -2>();
        const F vz_km1 = a.template V<2, -1, 0, 0>(), vz_kp1 = a.template V<2, 1, 0, 0>(), vz_km2 = a.template V<2, -2, 0, 0>();

        const F dxVx = (vx0 - vx_im1) * re40x - (vx_ip1 - vx_im2) * re41x;
        const F dyVy = (vy0 - vy_jm1) * re40y - (vy_jp1 - vy_jm2) * re41y;
        const F dzVz = (vz0 - vz_km1) * re40z - (vz_kp1 - vz_km2) * re41z;

        const float mu2 = 2 * mu0;
        const float lam2mu = a.lam() + mu2;
        const float taup1 = (NM > 0) ? a.taup() : 0.0f;

        const float d3v3 = (float)(dxVx + dyVy + dzVz);
        const float dyVy_dzVz = (float)(dyVy + dzVz);
        const float dxVx_dzVz = (float)(dxVx + dzVz);
        const float dxVx_dyVy = (float)(dxVx + dyVy);

        float Rxx_n = 0.0f, Ryy_n = 0.0f, Rzz_n = 0.0f;
        if (NM > 0) {
#pragma unroll
            for (int m = 0; m < NM; m++) {
                const float c1 = p.c1[m], c2 = p.c2[m], d1 = p.d1[m];
                const float oxx = a.R(0 * NM + m), oyy = a.R(1 * NM + m), ozz = a.R(2 * NM + m);
                const float nxx = c1 * (oxx) - c2 * (lam2mu * taup1 * d3v3 - mu2 * taus1 * dyVy_dzVz) * dt;
                const float nyy = c1 * (oyy) - c2 * (lam2mu * taup1 * d3v3 - mu2 * taus1 * dxVx_dzVz) * dt;
                const float nzz = c1 * (ozz) - c2 * (lam2mu * taup1 * d3v3 - mu2 * taus1 * dxVx_dyVy) * dt;
                rn[0][m] = nxx; rn[1][m] = nyy; rn[2][m] = nzz;
                Rxx_n = Rxx_n + d1 * nxx; Ryy_n = Ryy_n + d1 * nyy; Rzz_n = Rzz_n + d1 * nzz;
            }
        }
        const float taup_plus1 = 1 + taup1 * (1 + p.d2);
        sn[0] = a.S(0) + (lam2mu * taup_plus1 * d3v3 - mu2 * taus_plus1 * dyVy_dzVz + Rxx_n) * dt;
        sn[1] = a.S(1) + (lam2mu * taup_plus1 * d3v3 - mu2 * taus_plus1 * dxVx_dzVz + Ryy_n) * dt;
        sn[2] = a.S(2) + (lam2mu * taup_plus1 * d3v3 - mu2 * taus_plus1 * dxVx_dyVy + Rzz_n) * dt;
        if (cerjan) {
            const float gcc = gxc * gyc * gzc;
            sn[0] = sn[0] * gcc; sn[1] = sn[1] * gcc; sn[2] = sn[2] * gcc;
        }
    }

    if (DO_S) {
        const F vx0 = a.template V<0, 0, 0, 0>(), vy0 = a.template V<1, 0, 0, 0>(), vz0 = a.template V<2, 0, 0, 0>();
        const F dxVy_dyVx = (a.template V<1, 0, 1, 0>() - vy0) * re40x - (a.template V<1, 0, 2, 0>() - a.template V<1, 0, -1, 0>()) * re41x +
                            (a.template V<0, 0, 0, 1>() - vx0) * re40y - (a.template V<0, 0, 0, 2>() - a.template V<0, 0, 0, -1>()) * re41y;
        const F dxVz_dzVx = (a.template V<2, 0, 1, 0>() - vz0) * re40x - (a.template V<2, 0, 2, 0>() - a.template V<2, 0, -1, 0>()) * re41x +
                            (a.template V<0, 1, 0, 0>() - vx0) * re40z - (a.template V<0, 2, 0, 0>() - a.template V<0, -1, 0, 0>()) * re41z;
        const F dyVz_dzVy = (a.template V<2, 0, 0, 1>() - vz0) * re40y - (a.template V<2, 0, 0, 2>() - a.template V<2, 0, 0, -1>()) * re41y +
                            (a.template V<1, 1, 0, 0>() - vy0) * re40z - (a.template V<1, 2, 0, 0>() - a.template V<1, -1, 0, 0>()) * re41z;

        const float mu_k = a.template mu<1, 0, 0>(), mu_i = a.template mu<0, 1, 0>(), mu_j = a.template mu<0, 0, 1>();
        const float muxz = mu_harm(mu0, mu_k, mu_i, a.template mu<1, 1, 0>());
        const float muxy = mu_harm(mu0, mu_i, mu_j, a.template mu<0, 1, 1>());
        const float muyz = mu_harm(mu0, mu_k, mu_j, a.template mu<1, 0, 1>());

        float Ryz_n = 0.0f, Rxz_n = 0.0f, Rxy_n = 0.0f;
        if (NM > 0) {
#pragma unroll
            for (int m = 0; m < NM; m++) {
                const float c1 = p.c1[m], c2 = p.c2[m], d1 = p.d1[m];
                const float oyz = a.R(3 * NM + m), oxz = a.R(4 * NM + m), oxy = a.R(5 * NM + m);
                // shear R: the product with the F-kind strain rate is an F expression rounded on store (m_kernel.f90:320-322)
                const float nyz = (float)(c1 * (oyz) - c2 * muyz * taus1 * dyVz_dzVy * dt);
                const float nxz = (float)(c1 * (oxz) - c2 * muxz * taus1 * dxVz_dzVx * dt);
                const float nxy = (float)(c1 * (oxy) - c2 * muxy * taus1 * dxVy_dyVx * dt);
                rn[3][m] = nyz; rn[4][m] = nxz; rn[5][m] = nxy;
                Ryz_n = Ryz_n + d1 * nyz; Rxz_n = Rxz_n + d1 * nxz; Rxy_n = Rxy_n + d1 * nxy;
            }
        }
        sn[3] = a.S(3) + (muyz * taus_plus1 * dyVz_dzVy + Ryz_n) * dt;
        sn[4] = a.S(4) + (muxz * taus_plus1 * dxVz_dzVx + Rxz_n) * dt;
        sn[5] = a.S(5) + (muxy * taus_plus1 * dxVy_dyVx + Rxy_n) * dt;
        if (cerjan) {
            sn[3] = sn[3] * gxc * gyb * gzb;
            sn[4] = sn[4] * gxb * gyc * gzb;
            sn[5] = sn[5] * gxb * gyb * gzc;
        }
    }

    // all loads above, all stores below: without `restrict` knowledge the compiler may not hoist a load over a store
    if (DO_N) {
        if (NM > 0) {
#pragma unroll
            for (int m = 0; m < NM; m++) { a.setR(0 * NM + m, rn[0][m]); a.setR(1 * NM + m, rn[1][m]); a.setR(2 * NM + m, rn[2][m]); }
        }
        a.setS(0, sn[0]); a.setS(1, sn[1]); a.setS(2, sn[2]);
    }
    if (DO_S) {
        if (NM > 0) {
#pragma unroll
            for (int m = 0; m < NM; m++) { a.setR(3 * NM + m, rn[3][m]); a.setR(4 * NM + m, rn[4][m]); a.setR(5 * NM + m, rn[5][m]); }
        }
        a.setS(3, sn[3]); a.setS(4, sn[4]); a.setS(5, sn[5]);
    }
}

template <typename F, int NM>
__device__ __forceinline__ void stress_interior(const KParams<F> &p, long long n, int k, int mi, int mj, const int4 bnd) {
    stress_interior_t<F, NM>(p, AccDirect<F, NM>(p, n), k, mi, mj, bnd);
}

// stress update of one PML cell: absorb_p__update_stress, m_absorb_p.f90:453-519 (both k-loops fused).  As for the interior
// body the arithmetic is written once against an accessor: AccPmlDirect (global memory, sweep_direct) or AccPmlTma (shared-memory
// tiles staged by TMA, pml_tma.cuh).  V<f,dk,di,dj>, mu<dk,di,dj>, lam(), S(c) / setS(c, v) with c = xx yy zz yz xz xy,
// aux(q) / setAux(q, v) with q from enum Aux.
template <typename F>
struct AccPmlDirect {
    const KParams<F> &p;
    long long n;
    float *A;
    __device__ __forceinline__ AccPmlDirect(const KParams<F> &p_, long long n_, long long a_) : p(p_), n(n_), A(p_.aux + a_) {}
    template <int f, int dk, int di, int dj> __device__ __forceinline__ F V() const {
        const F *b = (f == 0) ? p.Vx : (f == 1) ? p.Vy : p.Vz;
        return ldro(b + n + dk + di * p.SI + dj * p.SJ);
    }
    template <int dk, int di, int dj> __device__ __forceinline__ float mu() const { return ldro(p.mu + n + dk + di * p.SI + dj * p.SJ); }
    __device__ __forceinline__ float lam() const { return ldro(p.lam + n); }
    __device__ __forceinline__ F *sptr(int c) const { return c == 0 ? p.Sxx : c == 1 ? p.Syy : c == 2 ? p.Szz : c == 3 ? p.Syz : c == 4 ? p.Sxz : p.Sxy; }
    __device__ __forceinline__ F S(int c) const { return sptr(c)[n]; }
    __device__ __forceinline__ void setS(int c, F v) const { sptr(c)[n] = v; }
    __device__ __forceinline__ float aux(int q) const { return A[q * p.naux]; }
    __device__ __forceinline__ void setAux(int q, float v) const { A[q * p.naux] = v; }
    // velocity side
    template <int c, int dk, int di, int dj> __device__ __forceinline__ F Sn() const {
        const F *b = c == 0 ? p.Sxx : c == 1 ? p.Syy : c == 2 ? p.Szz : c == 3 ? p.Syz : c == 4 ? p.Sxz : p.Sxy;
        return ldro(b + n + dk + di * p.SI + dj * p.SJ);
    }
    template <int dk, int di, int dj> __device__ __forceinline__ float rho() const { return ldro(p.rho + n + dk + di * p.SI + dj * p.SJ); }
    __device__ __forceinline__ F Vc(int f) const { return (f == 0 ? p.Vx : f == 1 ? p.Vy : p.Vz)[n]; }
    __device__ __forceinline__ void setV(int f, F v) const { (f == 0 ? p.Vx : f == 1 ? p.Vy : p.Vz)[n] = v; }
};

template <typename F, typename A>
__device__ __forceinline__ void stress_pml_t(const KParams<F> &p, const A &a, const float4 gxc, const float4 gxe, const float4 gyc,
                                             const float4 gye, const float4 gzc, const float4 gze) {
    const float dt = p.dt;
    const F r20x = p.r20x, r20y = p.r20y, r20z = p.r20z;

    const F vx0 = a.template V<0, 0, 0, 0>(), vy0 = a.template V<1, 0, 0, 0>(), vz0 = a.template V<2, 0, 0, 0>();
    const F dxVx = (vx0 - a.template V<0, 0, -1, 0>()) * r20x;
    const F dyVy = (vy0 - a.template V<1, 0, 0, -1>()) * r20y;
    const F dzVz = (vz0 - a.template V<2, -1, 0, 0>()) * r20z;
    const float mu0 = a.template mu<0, 0, 0>(), mu_k = a.template mu<1, 0, 0>(), mu_i = a.template mu<0, 1, 0>(), mu_j = a.template mu<0, 0, 1>();
    const float lam2mu_R = (a.lam() + 2 * mu0);
    const float lam_R = lam2mu_R - 2 * mu0;

    const float a_xVx = a.aux(axVx), a_yVy = a.aux(ayVy), a_zVz = a.aux(azVz);
    const float dxVx_ade = gxc.x * (float)(dxVx) + gxc.y * a_xVx;
    const float dyVy_ade = gyc.x * (float)(dyVy) + gyc.y * a_yVy;
    const float dzVz_ade = gzc.x * (float)(dzVz) + gzc.y * a_zVz;

    a.setS(0, a.S(0) + (lam2mu_R * dxVx_ade + lam_R * (dyVy_ade + dzVz_ade)) * dt);
    a.setS(1, a.S(1) + (lam2mu_R * dyVy_ade + lam_R * (dxVx_ade + dzVz_ade)) * dt);
    a.setS(2, a.S(2) + (lam2mu_R * dzVz_ade + lam_R * (dxVx_ade + dyVy_ade)) * dt);

    a.setAux(axVx, gxc.z * a_xVx + gxc.w * (float)(dxVx) * dt);
    a.setAux(ayVy, gyc.z * a_yVy + gyc.w * (float)(dyVy) * dt);
    a.setAux(azVz, gzc.z * a_zVz + gzc.w * (float)(dzVz) * dt);

    const F dxVy = (a.template V<1, 0, 1, 0>() - vy0) * r20x;
    const F dxVz = (a.template V<2, 0, 1, 0>() - vz0) * r20x;
    const F dyVx = (a.template V<0, 0, 0, 1>() - vx0) * r20y;
    const F dyVz = (a.template V<2, 0, 0, 1>() - vz0) * r20y;
    const F dzVx = (a.template V<0, 1, 0, 0>() - vx0) * r20z;
    const F dzVy = (a.template V<1, 1, 0, 0>() - vy0) * r20z;

    const float muxz = mu_harm(mu0, mu_k, mu_i, a.template mu<1, 1, 0>());
    const float muxy = mu_harm(mu0, mu_i, mu_j, a.template mu<0, 1, 1>());
    const float muyz = mu_harm(mu0, mu_k, mu_j, a.template mu<1, 0, 1>());

    const float a_yVx = a.aux(ayVx), a_zVx = a.aux(azVx), a_xVy = a.aux(axVy);
    const float a_zVy = a.aux(azVy), a_xVz = a.aux(axVz), a_yVz = a.aux(ayVz);

    a.setS(3, a.S(3) + muyz * (gye.x * dyVz + gze.x * dzVy + gye.y * a_yVz + gze.y * a_zVy) * dt);
    a.setS(4, a.S(4) + muxz * (gxe.x * dxVz + gze.x * dzVx + gxe.y * a_xVz + gze.y * a_zVx) * dt);
    a.setS(5, a.S(5) + muxy * (gxe.x * dxVy + gye.x * dyVx + gxe.y * a_xVy + gye.y * a_yVx) * dt);

    a.setAux(ayVx, gye.z * a_yVx + gye.w * (float)(dyVx) * dt);
    a.setAux(azVx, gze.z * a_zVx + gze.w * (float)(dzVx) * dt);
    a.setAux(axVy, gxe.z * a_xVy + gxe.w * (float)(dxVy) * dt);
    a.setAux(azVy, gze.z * a_zVy + gze.w * (float)(dzVy) * dt);
    a.setAux(axVz, gxe.z * a_xVz + gxe.w * (float)(dxVz) * dt);
    a.setAux(ayVz, gye.z * a_yVz + gye.w * (float)(dyVz) * dt);
}

template <typename F>
__device__ __forceinline__ void stress_pml(const KParams<F> &p, long long n, int k, int li, int lj, long long a) {
    const float4 gxc = ldro(p.gxc + li), gxe = ldro(p.gxe + li), gyc = ldro(p.gyc + lj), gye = ldro(p.gye + lj);
    const float4 gzc = ldro(p.gzc + (k - 1)), gze = ldro(p.gze + (k - 1));
    stress_pml_t<F>(p, AccPmlDirect<F>(p, n, a), gxc, gxe, gyc, gye, gzc, gze);
}

// ------------------------------------------------------------------------------------------------
// velocity update of one interior cell: kernel__update_vel m_kernel.f90:99-129 (+ Cerjan m_absorb_c.f90:185-187).
// Accessor form as for the stress body: S<c,dk,di,dj> with c = 0 xx, 1 yy, 2 zz, 3 yz, 4 xz, 5 xy.
template <typename F>
struct AccVelDirect {
    const KParams<F> &p;
    long long n;
    __device__ __forceinline__ AccVelDirect(const KParams<F> &p_, long long n_) : p(p_), n(n_) {}
    template <int c, int dk, int di, int dj> __device__ __forceinline__ F S() const {
        const F *b = c == 0 ? p.Sxx : c == 1 ? p.Syy : c == 2 ? p.Szz : c == 3 ? p.Syz : c == 4 ? p.Sxz : p.Sxy;
        return ldro(b + n + dk + di * p.SI + dj * p.SJ);
    }
    template <int dk, int di, int dj> __device__ __forceinline__ float rho() const { return ldro(p.rho + n + dk + di * p.SI + dj * p.SJ); }
    __device__ __forceinline__ F V(int f) const { return (f == 0 ? p.Vx : f == 1 ? p.Vy : p.Vz)[n]; }
    __device__ __forceinline__ void setV(int f, F v) const { (f == 0 ? p.Vx : f == 1 ? p.Vy : p.Vz)[n] = v; }
};

template <typename F, typename A>
__device__ __forceinline__ void vel_interior_calc(const KParams<F> &p, const A &a, int k, int mi, int mj, const int4 bnd, F &vx, F &vy, F &vz) {
    const int o = fd_order_sel(k, bnd);
    const F re40x = p.r40x[o], re41x = p.r41x[o], re40y = p.r40y[o], re41y = p.r41y[o], re40z = p.r40z[o], re41z = p.r41z[o];
    const float dt = p.dt;
    const F sxy0 = a.template S<5, 0, 0, 0>(), sxz0 = a.template S<4, 0, 0, 0>(), syz0 = a.template S<3, 0, 0, 0>();

    const F d3Sx3 = (a.template S<0, 0, 1, 0>() - a.template S<0, 0, 0, 0>()) * re40x - (a.template S<0, 0, 2, 0>() - a.template S<0, 0, -1, 0>()) * re41x +
                    (sxy0 - a.template S<5, 0, 0, -1>()) * re40y - (a.template S<5, 0, 0, 1>() - a.template S<5, 0, 0, -2>()) * re41y +
                    (sxz0 - a.template S<4, -1, 0, 0>()) * re40z - (a.template S<4, 1, 0, 0>() - a.template S<4, -2, 0, 0>()) * re41z;
    const F d3Sy3 = (sxy0 - a.template S<5, 0, -1, 0>()) * re40x - (a.template S<5, 0, 1, 0>() - a.template S<5, 0, -2, 0>()) * re41x +
                    (a.template S<1, 0, 0, 1>() - a.template S<1, 0, 0, 0>()) * re40y - (a.template S<1, 0, 0, 2>() - a.template S<1, 0, 0, -1>()) * re41y +
                    (syz0 - a.template S<3, -1, 0, 0>()) * re40z - (a.template S<3, 1, 0, 0>() - a.template S<3, -2, 0, 0>()) * re41z;
    const F d3Sz3 = (sxz0 - a.template S<4, 0, -1, 0>()) * re40x - (a.template S<4, 0, 1, 0>() - a.template S<4, 0, -2, 0>()) * re41x +
                    (syz0 - a.template S<3, 0, 0, -1>()) * re40y - (a.template S<3, 0, 0, 1>() - a.template S<3, 0, 0, -2>()) * re41y +
                    (a.template S<2, 1, 0, 0>() - a.template S<2, 0, 0, 0>()) * re40z - (a.template S<2, 2, 0, 0>() - a.template S<2, -1, 0, 0>()) * re41z;

    const float rho0 = a.template rho<0, 0, 0>();
    vx = a.V(0) + 2.0f / (rho0 + a.template rho<0, 1, 0>()) * d3Sx3 * dt;
    vy = a.V(1) + 2.0f / (rho0 + a.template rho<0, 0, 1>()) * d3Sy3 * dt;
    vz = a.V(2) + 2.0f / (rho0 + a.template rho<1, 0, 0>()) * d3Sz3 * dt;
    if (p.abc == 2) {
        const int kk = k + KOFF - 1;
        const float gxc = p.cgx_c[mi], gxb = p.cgx_b[mi], gyc = p.cgy_c[mj], gyb = p.cgy_b[mj], gzc = p.cgz_c[kk], gzb = p.cgz_b[kk];
        vx = vx * gxb * gyc * gzc;
        vy = vy * gxc * gyb * gzc;
        vz = vz * gxc * gyc * gzb;
    }
}

template <typename F, typename A>
__device__ __forceinline__ void vel_interior_t(const KParams<F> &p, const A &a, int k, int mi, int mj, const int4 bnd) {
    F vx, vy, vz;
    vel_interior_calc<F, A>(p, a, k, mi, mj, bnd, vx, vy, vz);
    a.setV(0, vx); a.setV(1, vy); a.setV(2, vz);
}

template <typename F>
__device__ __forceinline__ void vel_interior(const KParams<F> &p, long long n, int k, int mi, int mj, const int4 bnd) {
    vel_interior_t<F>(p, AccVelDirect<F>(p, n), k, mi, mj, bnd);
}

// velocity update of one PML cell: absorb_p__update_vel m_absorb_p.f90:261-308.  Accessor: Sn<c,dk,di,dj> (c = xx yy zz yz xz xy),
// rho<dk,di,dj>, Vc(f) / setV(f, v), aux(q) / setAux(q, v).
template <typename F, typename A>
__device__ __forceinline__ void vel_pml_t(const KParams<F> &p, const A &a, const float4 gxc, const float4 gxe, const float4 gyc,
                                          const float4 gye, const float4 gzc, const float4 gze) {
    const float dt = p.dt;
    const F r20x = p.r20x, r20y = p.r20y, r20z = p.r20z;
    const F sxy0 = a.template Sn<5, 0, 0, 0>(), sxz0 = a.template Sn<4, 0, 0, 0>(), syz0 = a.template Sn<3, 0, 0, 0>();

    const F dxSxx = (a.template Sn<0, 0, 1, 0>() - a.template Sn<0, 0, 0, 0>()) * r20x;
    const F dySyy = (a.template Sn<1, 0, 0, 1>() - a.template Sn<1, 0, 0, 0>()) * r20y;
    const F dzSzz = (a.template Sn<2, 1, 0, 0>() - a.template Sn<2, 0, 0, 0>()) * r20z;
    const F dySyz = (syz0 - a.template Sn<3, 0, 0, -1>()) * r20y;
    const F dzSyz = (syz0 - a.template Sn<3, -1, 0, 0>()) * r20z;
    const F dxSxz = (sxz0 - a.template Sn<4, 0, -1, 0>()) * r20x;
    const F dzSxz = (sxz0 - a.template Sn<4, -1, 0, 0>()) * r20z;
    const F dxSxy = (sxy0 - a.template Sn<5, 0, -1, 0>()) * r20x;
    const F dySxy = (sxy0 - a.template Sn<5, 0, 0, -1>()) * r20y;

    const float rho0 = a.template rho<0, 0, 0>();
    const float bx = 2.0f / (rho0 + a.template rho<0, 1, 0>());
    const float by = 2.0f / (rho0 + a.template rho<0, 0, 1>());
    const float bz = 2.0f / (rho0 + a.template rho<1, 0, 0>());

    const float a_xSxx = a.aux(axSxx), a_ySxy = a.aux(aySxy), a_zSxz = a.aux(azSxz);
    const float a_xSxy = a.aux(axSxy), a_ySyy = a.aux(aySyy), a_zSyz = a.aux(azSyz);
    const float a_xSxz = a.aux(axSxz), a_ySyz = a.aux(aySyz), a_zSzz = a.aux(azSzz);

    a.setV(0, a.Vc(0) + bx * (float)(gxe.x * dxSxx + gyc.x * dySxy + gzc.x * dzSxz + gxe.y * a_xSxx + gyc.y * a_ySxy + gzc.y * a_zSxz) * dt);
    a.setV(1, a.Vc(1) + by * (float)(gxc.x * dxSxy + gye.x * dySyy + gzc.x * dzSyz + gxc.y * a_xSxy + gye.y * a_ySyy + gzc.y * a_zSyz) * dt);
    a.setV(2, a.Vc(2) + bz * (float)(gxc.x * dxSxz + gyc.x * dySyz + gze.x * dzSzz + gxc.y * a_xSxz + gyc.y * a_ySyz + gze.y * a_zSzz) * dt);

    a.setAux(axSxx, gxe.z * a_xSxx + gxe.w * (float)(dxSxx) * dt);
    a.setAux(aySxy, gyc.z * a_ySxy + gyc.w * (float)(dySxy) * dt);
    a.setAux(azSxz, gzc.z * a_zSxz + gzc.w * (float)(dzSxz) * dt);
    a.setAux(axSxy, gxc.z * a_xSxy + gxc.w * (float)(dxSxy) * dt);
    a.setAux(aySyy, gye.z * a_ySyy + gye.w * (float)(dySyy) * dt);
    a.setAux(azSyz, gzc.z * a_zSyz + gzc.w * (float)(dzSyz) * dt);
    a.setAux(axSxz, gxc.z * a_xSxz + gxc.w * (float)(dxSxz) * dt);
    a.setAux(aySyz, gyc.z * a_ySyz + gyc.w * (float)(dySyz) * dt);
    a.setAux(azSzz, gze.z * a_zSzz + gze.w * (float)(dzSzz) * dt);
}

template <typename F>
__device__ __forceinline__ void vel_pml(const KParams<F> &p, long long n, int k, int li, int lj, long long a) {
    const float4 gxc = ldro(p.gxc + li), gxe = ldro(p.gxe + li), gyc = ldro(p.gyc + lj), gye = ldro(p.gye + lj);
    const float4 gzc = ldro(p.gzc + (k - 1)), gze = ldro(p.gze + (k - 1));
    vel_pml_t<F>(p, AccPmlDirect<F>(p, n, a), gxc, gxe, gyc, gye, gzc, gze);
}

// ------------------------------------------------------------------------------------------------
// Sweep kernels, version 1 ("direct"): thread = one (k,i) column position, block = TK x TI tile of the
// (k,i) plane, marching over jlen planes along j (the slowest axis).  Neighbour values come through
// L1/L2 (read-only path); every streamed array (S, R, medium, aux) is touched exactly once.
// Interior cells and absorber cells partition the owned box (m_global.f90:334-376) and read only the
// other field family, so one pass per family is order-independent.
#ifndef SWPC_MINB
#define SWPC_MINB 3
#endif

__device__ __forceinline__ void pf_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// L2 prefetch of everything the cell at linear index n (pf planes ahead along j) will stream from HBM: turns the
// DRAM latency of the next plane into an L2 hit without holding registers.
template <typename F, int NM, bool STRESS>
__device__ __forceinline__ void prefetch_cell(const KParams<F> &p, long long n, bool pml_target, long long a) {
    if (STRESS) {
        pf_l2(p.Sxx + n); pf_l2(p.Syy + n); pf_l2(p.Szz + n); pf_l2(p.Syz + n); pf_l2(p.Sxz + n); pf_l2(p.Sxy + n);
        pf_l2(p.lam + n); pf_l2(p.mu + n + p.SJ);
        pf_l2(p.Vx + n + 2 * p.SJ); pf_l2(p.Vy + n + p.SJ); pf_l2(p.Vz + n + 2 * p.SJ);
        if (pml_target) {
#pragma unroll
            for (int q = 0; q < 9; q++) pf_l2(p.aux + a + q * p.naux);
        } else {
            pf_l2(p.taup + n); pf_l2(p.taus + n);
#pragma unroll
            for (int q = 0; q < 6 * NM; q++) pf_l2(p.R + n + q * p.ncell);
        }
    } else {
        pf_l2(p.Vx + n); pf_l2(p.Vy + n); pf_l2(p.Vz + n); pf_l2(p.rho + n + p.SJ);
        pf_l2(p.Sxx + n); pf_l2(p.Szz + n); pf_l2(p.Sxz + n);
        pf_l2(p.Syy + n + 2 * p.SJ); pf_l2(p.Sxy + n + p.SJ); pf_l2(p.Syz + n + p.SJ);
        if (pml_target) {
#pragma unroll
            for (int q = 9; q < 18; q++) pf_l2(p.aux + a + q * p.naux);
        }
    }
}

// inclusive: k 1-based, li / lj local 0-based; skip_interior: absorber cells only; flat: threads are numbered over the
// (k, i) cells of the box, k fastest, instead of one warp per 32 consecutive k -- for boxes that are thin in k (the
// bottom absorber slab: 20 cells per column) every lane then has a cell and a warp spans one and a half columns
struct Box3 { int k0, k1, li0, li1, lj0, lj1; int skip_interior; int flat; };

template <typename F, int NM, bool STRESS>
__global__ void __launch_bounds__(256, SWPC_MINB) sweep_direct(const __grid_constant__ KParams<F> p, const Box3 b, int jlen, int pf) {
    int k, li;
    if (b.flat) {
        const int nk = b.k1 - b.k0 + 1;
        const long long t = (long long)blockIdx.x * (blockDim.x * blockDim.y) + threadIdx.y * blockDim.x + threadIdx.x;
        k = b.k0 + (int)(t % nk);
        li = b.li0 + (int)(t / nk);
    } else {
        k = b.k0 + blockIdx.x * blockDim.x + threadIdx.x;
        li = b.li0 + blockIdx.y * blockDim.y + threadIdx.y;
    }
    if (k > b.k1 || li > b.li1) return;
    const int mi = li + HALO;
    const int ljs = b.lj0 + blockIdx.z * jlen;
    const int lje = min(ljs + jlen, b.lj1 + 1);
    const bool pml_mode = (p.abc == 1);
    const bool i_interior = (li >= p.li0_k && li <= p.li1_k);
    for (int lj = ljs; lj < lje; lj++) {
        const int mj = lj + HALO;
        const long long col = (long long)mi + (long long)p.NXM * mj;
        const long long n = (long long)(k + KOFF - 1) + (long long)p.NZP * col;
        if (pf > 0 && lj + pf < lje) {
            const long long colp = col + (long long)p.NXM * pf;
            bool pt = false;
            long long ap = 0;
            if (pml_mode) {
                const int kb = p.kbeg_a[colp];
                pt = (k >= kb);
                if (pt) ap = p.aoff[li + (long long)p.nxp * (lj + pf)] + (k - kb);
            }
            // (an absorber-only box skips its interior lanes: prefetching their S / R / medium lines would be pure waste)
            if (!b.skip_interior || pt) prefetch_cell<F, NM, STRESS>(p, n + p.SJ * pf, pt, ap);
        }
        bool is_pml = false;
        if (pml_mode) is_pml = (k >= p.kbeg_a[col]);
        if (is_pml) {
            const long long a = p.aoff[li + (long long)p.nxp * lj] + (k - p.kbeg_a[col]);
            if (STRESS) stress_pml<F>(p, n, k, li, lj, a);
            else vel_pml<F>(p, n, k, li, lj, a);
        } else if (!b.skip_interior && i_interior && lj >= p.lj0_k && lj <= p.lj1_k && k <= p.k1_k) {
            const int4 bnd = p.band[col];
            if (STRESS) stress_interior<F, NM>(p, n, k, mi, mj, bnd);
            else vel_interior<F>(p, n, k, mi, mj, bnd);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Velocity sweep over INTERIOR cells, version 2 ("ring"): software-pipelined in registers.  The direct kernel issues its
// ~36 loads per cell in register-limited batches that each wait for an L2 round trip (long-scoreboard stalls at 36 %
// occupancy, profiles/r01_ncu_full_sweeps_*.txt).  Here a thread marches along j and
//   * keeps the j-direction stencil of its own column in a register ring (Syy j-1..j+2, Sxy / Syz j-2..j+1),
//   * loads everything plane j+1 needs from its own column (Sxx, Szz, Sxz, Vx, Vy, Vz, rho and the ring's new heads
//     Syy(j+3), Sxy(j+2), Syz(j+2)) while plane j is being computed, so those HBM-latency loads have a full iteration
//     to land, and
//   * reads only the in-plane (k, i) neighbours on demand -- lines its neighbour threads pulled into L1 one or two
//     iterations earlier.
// Same arithmetic body (vel_interior_calc) -> bit-identical results.  Absorber cells stay with sweep_direct.
template <typename F>
struct AccVelRing {
    const KParams<F> &p;
    long long n;
    F sxx0, szz0, sxz0;
    F syy[4];   // j-1 .. j+2
    F sxy[4];   // j-2 .. j+1
    F syz[4];   // j-2 .. j+1
    F v[3];
    float rho0, rho_j1;
    __device__ __forceinline__ AccVelRing(const KParams<F> &p_) : p(p_) {}
    template <int c, int dk, int di, int dj> __device__ __forceinline__ F S() const {
        if constexpr (dk == 0 && di == 0) {
            if constexpr (c == 0) return sxx0;
            else if constexpr (c == 1) return syy[dj + 1];
            else if constexpr (c == 2) return szz0;
            else if constexpr (c == 3) return syz[dj + 2];
            else if constexpr (c == 4) return sxz0;
            else return sxy[dj + 2];
        } else {
            static_assert(dj == 0, "in-plane neighbour expected");
            const F *b = c == 0 ? p.Sxx : c == 1 ? p.Syy : c == 2 ? p.Szz : c == 3 ? p.Syz : c == 4 ? p.Sxz : p.Sxy;
            return ldro(b + n + dk + di * p.SI);
        }
    }
    template <int dk, int di, int dj> __device__ __forceinline__ float rho() const {
        if constexpr (dk == 0 && di == 0) return dj == 0 ? rho0 : rho_j1;
        else return ldro(p.rho + n + dk + di * p.SI);
    }
    __device__ __forceinline__ F V(int f) const { return v[f]; }
};

#ifndef SWPC_RING_MINB
#define SWPC_RING_MINB 2
#endif
#ifndef SWPC_RING_THREADS
#define SWPC_RING_THREADS 256
#endif
template <typename F>
__global__ void __launch_bounds__(SWPC_RING_THREADS, SWPC_RING_MINB) vel_ring(const __grid_constant__ KParams<F> p, const Box3 b, int jlen, int pf) {
    const int k = b.k0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int li = b.li0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (k > b.k1 || li > b.li1) return;
    const int mi = li + HALO;
    const int ljs = b.lj0 + blockIdx.z * jlen;
    const int lje = min(ljs + jlen, b.lj1 + 1);
    const long long sj = p.SJ;
    AccVelRing<F> a(p);
    long long col = (long long)mi + (long long)p.NXM * (ljs + HALO);
    a.n = (long long)(k + KOFF - 1) + (long long)p.NZP * col;
    {   // prologue: the ring and the own-column values of the first plane
        const long long n = a.n;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            a.syy[q] = ldro(p.Syy + n + (q - 1) * sj);
            a.sxy[q] = ldro(p.Sxy + n + (q - 2) * sj);
            a.syz[q] = ldro(p.Syz + n + (q - 2) * sj);
        }
        a.sxx0 = ldro(p.Sxx + n); a.szz0 = ldro(p.Szz + n); a.sxz0 = ldro(p.Sxz + n);
        a.v[0] = lds_(p.Vx + n); a.v[1] = lds_(p.Vy + n); a.v[2] = lds_(p.Vz + n);
        a.rho0 = ldro(p.rho + n); a.rho_j1 = ldro(p.rho + n + sj);
    }
    int4 bnd = p.band[col];
    for (int lj = ljs; lj < lje; lj++) {
        const long long n = a.n;
        const bool more = (lj + 1 < lje);
        // loads for plane j+1 (consumed next iteration)
        F nsxx = 0, nszz = 0, nsxz = 0, nsyy = 0, nsxy = 0, nsyz = 0, nv0 = 0, nv1 = 0, nv2 = 0;
        float nrho = 0.0f;
        int4 nbnd = bnd;
        if (more) {
            nsxx = ldro(p.Sxx + n + sj); nszz = ldro(p.Szz + n + sj); nsxz = ldro(p.Sxz + n + sj);
            nsyy = ldro(p.Syy + n + 3 * sj); nsxy = ldro(p.Sxy + n + 2 * sj); nsyz = ldro(p.Syz + n + 2 * sj);
            nv0 = lds_(p.Vx + n + sj); nv1 = lds_(p.Vy + n + sj); nv2 = lds_(p.Vz + n + sj);
            nrho = ldro(p.rho + n + 2 * sj);
            nbnd = p.band[col + p.NXM];
            if (pf > 0 && lj + 1 + pf < lje) {   // optional L2 prefetch further ahead
                const long long q = n + sj * (1 + pf);
                pf_l2(p.Sxx + q); pf_l2(p.Szz + q); pf_l2(p.Sxz + q); pf_l2(p.Vx + q); pf_l2(p.Vy + q); pf_l2(p.Vz + q);
                pf_l2(p.Syy + q + 2 * sj); pf_l2(p.Sxy + q + sj); pf_l2(p.Syz + q + sj); pf_l2(p.rho + q + sj);
            }
        }
        F vx, vy, vz;
        vel_interior_calc<F>(p, a, k, mi, lj + HALO, bnd, vx, vy, vz);
        sts_(p.Vx + n, vx); sts_(p.Vy + n, vy); sts_(p.Vz + n, vz);
        // rotate
        a.syy[0] = a.syy[1]; a.syy[1] = a.syy[2]; a.syy[2] = a.syy[3]; a.syy[3] = nsyy;
        a.sxy[0] = a.sxy[1]; a.sxy[1] = a.sxy[2]; a.sxy[2] = a.sxy[3]; a.sxy[3] = nsxy;
        a.syz[0] = a.syz[1]; a.syz[1] = a.syz[2]; a.syz[2] = a.syz[3]; a.syz[3] = nsyz;
        a.sxx0 = nsxx; a.szz0 = nszz; a.sxz0 = nsxz;
        a.v[0] = nv0; a.v[1] = nv1; a.v[2] = nv2;
        a.rho0 = a.rho_j1; a.rho_j1 = nrho;
        bnd = nbnd;
        a.n += sj;
        col += p.NXM;
    }
}

// ------------------------------------------------------------------------------------------------
// Velocity sweep over INTERIOR cells, version 3 ("ring, two cells per thread"): for float32 fields (the reference's MP = SP
// build) vel_ring moves half the bytes of the float64 run with the same instruction count and ends up instruction / latency
// bound (50 % of the HBM peak).  Here a thread owns the cells (k, k+1) of a column: every own-column stream and every in-plane
// neighbour comes in as ONE 64-bit load for both cells -- 25 vector loads per cell pair where vel_ring issues 72 scalar ones --
// and the k-neighbours that are the pair's own other cell come from registers.  Same arithmetic body (vel_interior_calc).
template <typename F> struct Vec2;
template <> struct Vec2<float> { using T = float2; };
template <> struct Vec2<double> { using T = double2; };

template <typename F>
struct RingPair {
    using F2 = typename Vec2<F>::T;
    F2 sxx0, szz0, sxz0;               // own column, plane j
    F2 syy[4];                         // j-1 .. j+2
    F2 sxy[4], syz[4];                 // j-2 .. j+1
    F2 v[3];
    float2 rho0, rho_j1, rho_i1;
    float rho_k2;
    F2 sxx_i[3];                       // columns i-1, i+1, i+2
    F2 sxy_i[3], sxz_i[3];             // columns i-2, i-1, i+1
    F2 sxz_k[2], syz_k[2], szz_k[2];   // rows (k-2, k-1) and (k+2, k+3)
};

template <typename F, int E>
struct AccVelPair {
    using F2 = typename Vec2<F>::T;
    const RingPair<F> &r;
    __device__ __forceinline__ AccVelPair(const RingPair<F> &r_) : r(r_) {}
    template <typename T2> static __device__ __forceinline__ auto el(const T2 &v) { return E == 0 ? v.x : v.y; }
    // value at row k + E + dk of a column whose rows (k-2, k-1), (k, k+1), (k+2, k+3) are lo, mid, hi
    template <int dk> static __device__ __forceinline__ F row(const F2 &lo, const F2 &mid, const F2 &hi) {
        constexpr int q = E + dk;
        static_assert(q >= -2 && q <= 3, "row out of the preloaded range");
        return q == -2 ? lo.x : q == -1 ? lo.y : q == 0 ? mid.x : q == 1 ? mid.y : q == 2 ? hi.x : hi.y;
    }
    template <int c, int dk, int di, int dj> __device__ __forceinline__ F S() const {
        if constexpr (c == 0) {
            static_assert(dk == 0 && dj == 0, "Sxx: i neighbours only");
            if constexpr (di == 0) return el(r.sxx0);
            else return el(r.sxx_i[di == -1 ? 0 : di == 1 ? 1 : 2]);
        } else if constexpr (c == 1) {
            static_assert(dk == 0 && di == 0, "Syy: j neighbours only");
            return el(r.syy[dj + 1]);
        } else if constexpr (c == 2) {
            static_assert(di == 0 && dj == 0, "Szz: k neighbours only");
            return row<dk>(r.szz_k[0], r.szz0, r.szz_k[1]);
        } else if constexpr (c == 3) {
            static_assert(di == 0, "Syz: j and k neighbours only");
            if constexpr (dk == 0) return el(r.syz[dj + 2]);
            else return row<dk>(r.syz_k[0], r.syz[2], r.syz_k[1]);
        } else if constexpr (c == 4) {
            static_assert(dj == 0, "Sxz: i and k neighbours only");
            if constexpr (di != 0) return el(r.sxz_i[di == -2 ? 0 : di == -1 ? 1 : 2]);
            else return row<dk>(r.sxz_k[0], r.sxz0, r.sxz_k[1]);
        } else {
            static_assert(dk == 0, "Sxy: i and j neighbours only");
            if constexpr (di != 0) return el(r.sxy_i[di == -2 ? 0 : di == -1 ? 1 : 2]);
            else return el(r.sxy[dj + 2]);
        }
    }
    template <int dk, int di, int dj> __device__ __forceinline__ float rho() const {
        if constexpr (dk == 1) return E == 0 ? r.rho0.y : r.rho_k2;
        else if constexpr (di == 1) return el(r.rho_i1);
        else if constexpr (dj == 1) return el(r.rho_j1);
        else return el(r.rho0);
    }
    __device__ __forceinline__ F V(int f) const { return el(r.v[f]); }
};

template <typename T2, typename T> __device__ __forceinline__ T2 ld2ro(const T *ptr) { return __ldg(reinterpret_cast<const T2 *>(ptr)); }
template <typename T2, typename T> __device__ __forceinline__ T2 ld2cs(const T *ptr) { return __ldcs(reinterpret_cast<const T2 *>(ptr)); }

// (float64 fields: the pair's neighbour registers need more than 128 registers -- one block of 256 threads per SM, option ring_pair = 2)
template <typename F>
__global__ void __launch_bounds__(SWPC_RING_THREADS, (sizeof(F) == 8 ? 1 : SWPC_RING_MINB)) vel_ring2(const __grid_constant__ KParams<F> p, const Box3 b, int jlen, int pf) {
    using F2 = typename Vec2<F>::T;
    const int k = b.k0 + 2 * (blockIdx.x * blockDim.x + threadIdx.x);   // b.k0 is odd: index k + KOFF - 1 is even, 2-element loads are aligned
    const int li = b.li0 + blockIdx.y * blockDim.y + threadIdx.y;
    if (k > b.k1 || li > b.li1) return;
    const bool two = k + 1 <= b.k1;
    const int mi = li + HALO;
    const int ljs = b.lj0 + blockIdx.z * jlen;
    const int lje = min(ljs + jlen, b.lj1 + 1);
    const long long sj = p.SJ, si = p.SI;
    RingPair<F> r;
    long long col = (long long)mi + (long long)p.NXM * (ljs + HALO);
    long long n = (long long)(k + KOFF - 1) + (long long)p.NZP * col;
    {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            r.syy[q] = ld2ro<F2>(p.Syy + n + (q - 1) * sj);
            r.sxy[q] = ld2ro<F2>(p.Sxy + n + (q - 2) * sj);
            r.syz[q] = ld2ro<F2>(p.Syz + n + (q - 2) * sj);
        }
        r.sxx0 = ld2ro<F2>(p.Sxx + n); r.szz0 = ld2ro<F2>(p.Szz + n); r.sxz0 = ld2ro<F2>(p.Sxz + n);
        r.v[0] = ld2cs<F2>(p.Vx + n); r.v[1] = ld2cs<F2>(p.Vy + n); r.v[2] = ld2cs<F2>(p.Vz + n);
        r.rho0 = ld2ro<float2>(p.rho + n); r.rho_j1 = ld2ro<float2>(p.rho + n + sj);
    }
    int4 bnd = p.band[col];
    for (int lj = ljs; lj < lje; lj++) {
        const bool more = (lj + 1 < lje);
        F2 nsxx{}, nszz{}, nsxz{}, nsyy{}, nsxy{}, nsyz{}, nv0{}, nv1{}, nv2{};
        float2 nrho{};
        int4 nbnd = bnd;
        if (more) {   // plane j+1, consumed next iteration
            nsxx = ld2ro<F2>(p.Sxx + n + sj); nszz = ld2ro<F2>(p.Szz + n + sj); nsxz = ld2ro<F2>(p.Sxz + n + sj);
            nsyy = ld2ro<F2>(p.Syy + n + 3 * sj); nsxy = ld2ro<F2>(p.Sxy + n + 2 * sj); nsyz = ld2ro<F2>(p.Syz + n + 2 * sj);
            nv0 = ld2cs<F2>(p.Vx + n + sj); nv1 = ld2cs<F2>(p.Vy + n + sj); nv2 = ld2cs<F2>(p.Vz + n + sj);
            nrho = ld2ro<float2>(p.rho + n + 2 * sj);
            nbnd = p.band[col + p.NXM];
            if (pf > 0 && lj + 1 + pf < lje) {
                const long long q = n + sj * (1 + pf);
                pf_l2(p.Sxx + q); pf_l2(p.Szz + q); pf_l2(p.Sxz + q); pf_l2(p.Vx + q); pf_l2(p.Vy + q); pf_l2(p.Vz + q);
                pf_l2(p.Syy + q + 2 * sj); pf_l2(p.Sxy + q + sj); pf_l2(p.Syz + q + sj); pf_l2(p.rho + q + sj);
            }
        }
        // in-plane neighbours of plane j: lines the neighbour threads pulled into L1 one or two iterations earlier
        r.sxx_i[0] = ld2ro<F2>(p.Sxx + n - si); r.sxx_i[1] = ld2ro<F2>(p.Sxx + n + si); r.sxx_i[2] = ld2ro<F2>(p.Sxx + n + 2 * si);
        r.sxy_i[0] = ld2ro<F2>(p.Sxy + n - 2 * si); r.sxy_i[1] = ld2ro<F2>(p.Sxy + n - si); r.sxy_i[2] = ld2ro<F2>(p.Sxy + n + si);
        r.sxz_i[0] = ld2ro<F2>(p.Sxz + n - 2 * si); r.sxz_i[1] = ld2ro<F2>(p.Sxz + n - si); r.sxz_i[2] = ld2ro<F2>(p.Sxz + n + si);
        r.sxz_k[0] = ld2ro<F2>(p.Sxz + n - 2); r.sxz_k[1] = ld2ro<F2>(p.Sxz + n + 2);
        r.syz_k[0] = ld2ro<F2>(p.Syz + n - 2); r.syz_k[1] = ld2ro<F2>(p.Syz + n + 2);
        r.szz_k[0] = ld2ro<F2>(p.Szz + n - 2); r.szz_k[1] = ld2ro<F2>(p.Szz + n + 2);
        r.rho_i1 = ld2ro<float2>(p.rho + n + si);
        r.rho_k2 = ldro(p.rho + n + 2);
        F2 ox, oy, oz;
        vel_interior_calc<F>(p, AccVelPair<F, 0>(r), k, mi, lj + HALO, bnd, ox.x, oy.x, oz.x);
        vel_interior_calc<F>(p, AccVelPair<F, 1>(r), k + 1, mi, lj + HALO, bnd, ox.y, oy.y, oz.y);
        if (two) {
            __stcs(reinterpret_cast<F2 *>(p.Vx + n), ox); __stcs(reinterpret_cast<F2 *>(p.Vy + n), oy); __stcs(reinterpret_cast<F2 *>(p.Vz + n), oz);
        } else {
            sts_(p.Vx + n, ox.x); sts_(p.Vy + n, oy.x); sts_(p.Vz + n, oz.x);
        }
        r.syy[0] = r.syy[1]; r.syy[1] = r.syy[2]; r.syy[2] = r.syy[3]; r.syy[3] = nsyy;
        r.sxy[0] = r.sxy[1]; r.sxy[1] = r.sxy[2]; r.sxy[2] = r.sxy[3]; r.sxy[3] = nsxy;
        r.syz[0] = r.syz[1]; r.syz[1] = r.syz[2]; r.syz[2] = r.syz[3]; r.syz[3] = nsyz;
        r.sxx0 = nsxx; r.szz0 = nszz; r.sxz0 = nsxz;
        r.v[0] = nv0; r.v[1] = nv1; r.v[2] = nv2;
        r.rho0 = r.rho_j1; r.rho_j1 = nrho;
        bnd = nbnd;
        n += sj;
        col += p.NXM;
    }
}

// ------------------------------------------------------------------------------------------------
// source time functions, m_fdtool.f90:339-497 (PI is real(DP) there)
__device__ __forceinline__ float momentrate_dev(float t, int stf, float ts, float tr) {
    const double PI = 3.14159265358979323846;
    switch (stf) {
    case 0:   // boxcar
        return (ts <= t && t <= ts + tr) ? 1.0f / tr : 0.0f;
    case 1:   // triangle
        if (ts <= t && t <= ts + tr / 2) return 4 * (t - ts) / (tr * tr);
        if (ts + tr / 2 < t && t <= ts + tr) return -4 * (t - ts - tr) / (tr * tr);
        return 0.0f;
    case 2: {   // herrmann
        const float t1 = ts + tr / 4, t2 = ts + 3 * tr / 4, tr3 = tr * tr * tr;
        if (ts <= t && t < t1) return 16 * ((t - ts) * (t - ts)) / tr3;
        if (t1 <= t && t < t2) return -2 * (8 * (t * t + tr * ts + ts * ts - t * tr - 2 * t * ts) + tr * tr) / tr3;
        if (t2 <= t && t <= ts + tr) return 16 * ((ts + tr - t) * (ts + tr - t)) / tr3;
        return 0.0f;
    }
    case 4:   // cosine
        return (ts <= t && t <= ts + tr) ? (float)((1 - cos(2 * PI * (double)(t - ts) / (double)tr)) / (double)tr) : 0.0f;
    case 5:   // texp
        if (ts <= t) {
            const float tt = t - ts;
            return (float)((2 * PI) * (2 * PI) * (double)tt / (double)(tr * tr) * exp(-2 * PI * (double)tt / (double)tr));
        }
        return 0.0f;
    default: {   // kupper (also the reference's default branch)
        if (ts <= t && t <= ts + tr) {
            const double s = sin(PI * (double)(t - ts) / (double)tr);
            return (float)(3 * PI * (s * s * s) / (double)(4 * tr));
        }
        return 0.0f;
    }
    }
}

struct SrcParams {
    int nsrc;
    const int *ijk;         // 3*nsrc: LOCAL memory-box indices (mi, mj) and k
    const double *mo;       // nsrc
    const double *mij;      // 6*nsrc: mxx myy mzz myz mxz mxy  (body force: fx fy fz in the first three)
    const float *prm;       // 2*nsrc
    const float *stime;     // optional host-evaluated moment rate for this step (nsrc), else nullptr
    int nstv;               // > 0: the host-evaluated values travel in the kernel parameters themselves (up to 16 sources: no copy at all)
    float stv[16];
    int stf;
    float t;                // evaluation time
    double dt_dxyz;
    // boundary-first overlap: phase 0 = every target cell; 1 = only targets OUTSIDE the core box (boundary planes and
    // halo cells); 2 = only targets inside the core box [c_li0,c_li1] x [c_lj0,c_lj1] (local 0-based owned indices)
    int phase, c_li0, c_li1, c_lj0, c_lj1;
    __device__ __forceinline__ bool hit(int mi, int mj) const {
        if (phase == 0) return true;
        const int li = mi - HALO, lj = mj - HALO;
        const bool core = li >= c_li0 && li <= c_li1 && lj >= c_lj0 && lj <= c_lj1;
        return phase == 2 ? core : !core;
    }
};

// source__stressglut m_source.f90:798-841: 15 atomic adds per source (sources may share cells)
template <typename F>
__global__ void stressglut_kernel(const __grid_constant__ KParams<F> p, const SrcParams s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.nsrc) return;
    const float stime = s.nstv > 0 ? s.stv[i] : s.stime ? s.stime[i] : momentrate_dev(s.t, s.stf, s.prm[2 * i], s.prm[2 * i + 1]);
    const F sdrop = (F)((F)s.mo[i] * stime * (F)s.dt_dxyz);
    const long long si = p.SI, sj = p.SJ;
    const long long n = (long long)(s.ijk[3 * i + 2] + KOFF - 1) + (long long)p.NZP * ((long long)s.ijk[3 * i] + (long long)p.NXM * s.ijk[3 * i + 1]);
    const F mxx = (F)s.mij[6 * i], myy = (F)s.mij[6 * i + 1], mzz = (F)s.mij[6 * i + 2];
    const F myz = (F)s.mij[6 * i + 3], mxz = (F)s.mij[6 * i + 4], mxy = (F)s.mij[6 * i + 5];
    const int mi = s.ijk[3 * i], mj = s.ijk[3 * i + 1];
    const bool h00 = s.hit(mi, mj), h10 = s.hit(mi - 1, mj), h01 = s.hit(mi, mj - 1), h11 = s.hit(mi - 1, mj - 1);
    const F qxy = mxy * sdrop / 4, qxz = mxz * sdrop / 4, qyz = myz * sdrop / 4;
    if (h00) {
        atomicAdd(p.Sxx + n, -(mxx * sdrop));
        atomicAdd(p.Syy + n, -(myy * sdrop));
        atomicAdd(p.Szz + n, -(mzz * sdrop));
        atomicAdd(p.Sxy + n, -qxy); atomicAdd(p.Sxz + n, -qxz); atomicAdd(p.Sxz + n - 1, -qxz); atomicAdd(p.Syz + n, -qyz); atomicAdd(p.Syz + n - 1, -qyz);
    }
    if (h01) { atomicAdd(p.Sxy + n - sj, -qxy); atomicAdd(p.Syz + n - sj, -qyz); atomicAdd(p.Syz + n - 1 - sj, -qyz); }
    if (h10) { atomicAdd(p.Sxy + n - si, -qxy); atomicAdd(p.Sxz + n - si, -qxz); atomicAdd(p.Sxz + n - 1 - si, -qxz); }
    if (h11) atomicAdd(p.Sxy + n - si - sj, -qxy);
}

// source__bodyforce m_source.f90:870-892
template <typename F>
__global__ void bodyforce_kernel(const __grid_constant__ KParams<F> p, const SrcParams s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.nsrc) return;
    const float stime = s.nstv > 0 ? s.stv[i] : s.stime ? s.stime[i] : momentrate_dev(s.t, s.stf, s.prm[2 * i], s.prm[2 * i + 1]);
    const long long si = p.SI, sj = p.SJ;
    const long long n = (long long)(s.ijk[3 * i + 2] + KOFF - 1) + (long long)p.NZP * ((long long)s.ijk[3 * i] + (long long)p.NXM * s.ijk[3 * i + 1]);
    const F fx = (F)s.mij[6 * i], fy = (F)s.mij[6 * i + 1], fz = (F)s.mij[6 * i + 2];
    const F dtd = (F)s.dt_dxyz;
    const float *rho = p.rho;
    const int mi = s.ijk[3 * i], mj = s.ijk[3 * i + 1];
    if (s.hit(mi, mj)) {
        atomicAdd(p.Vx + n, (F)((2.0f / (rho[n] + rho[n + si])) * fx * stime * dtd / 2));
        atomicAdd(p.Vy + n, (F)((2.0f / (rho[n] + rho[n + sj])) * fy * stime * dtd / 2));
        atomicAdd(p.Vz + n, (F)((2.0f / (rho[n] + rho[n + 1])) * fz * stime * dtd / 2));
        atomicAdd(p.Vz + n - 1, (F)((2.0f / (rho[n] + rho[n - 1])) * fz * stime * dtd / 2));
    }
    if (s.hit(mi - 1, mj)) atomicAdd(p.Vx + n - si, (F)((2.0f / (rho[n] + rho[n - si])) * fx * stime * dtd / 2));
    if (s.hit(mi, mj - 1)) atomicAdd(p.Vy + n - sj, (F)((2.0f / (rho[n] + rho[n - sj])) * fy * stime * dtd / 2));
}

// wav__store m_wav.f90:397-625: per station, every step: displacement / strain accumulation (:430-513); when
// `sample`: velocity, displacement, stress, strain traces (:515-617).  Products are switched by bit flags.
struct WavParams {
    int nst, ntw, itw, sample;
    int sw_v, sw_u, sw_stress, sw_strain;
    const int *ijk;
    float *wav_v, *wav_u, *wav_s, *wav_e;   // (ntw,3,nst) (ntw,3,nst) (ntw,6,nst) (ntw,6,nst)
    float *acc;                             // 9 running sums per station: ux uy uz exx eyy ezz eyz exz exy
    float M0, UC;
    double r40[3], r41[3];                  // 9/8/d, 1/24/d in the field kind (m_wav.f90:127-132)
};

template <typename F>
__global__ void wav_store_kernel(const __grid_constant__ KParams<F> p, const WavParams w) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= w.nst) return;
    const long long si = p.SI, sj = p.SJ;
    const long long n = (long long)(w.ijk[3 * s + 2] + KOFF - 1) + (long long)p.NZP * ((long long)w.ijk[3 * s] + (long long)p.NXM * w.ijk[3 * s + 1]);
    const F *Vx = p.Vx, *Vy = p.Vy, *Vz = p.Vz;
    const float dt = p.dt;
    float *a = w.acc + 9 * s;
    if (w.sw_u) {
        a[0] = a[0] + (float)(Vx[n] + Vx[n - si]) * 0.5f * dt;
        a[1] = a[1] + (float)(Vy[n] + Vy[n - sj]) * 0.5f * dt;
        a[2] = a[2] - (float)(Vz[n] + Vz[n - 1]) * 0.5f * dt;
    }
    if (w.sw_strain) {
        const F r40x = (F)w.r40[0], r40y = (F)w.r40[1], r40z = (F)w.r40[2], r41x = (F)w.r41[0], r41y = (F)w.r41[1], r41z = (F)w.r41[2];
        const F dxVx = (Vx[n] - Vx[n - si]) * r40x - (Vx[n + si] - Vx[n - 2 * si]) * r41x;
        const F dyVy = (Vy[n] - Vy[n - sj]) * r40y - (Vy[n + sj] - Vy[n - 2 * sj]) * r41y;
        const F dzVz = (Vz[n] - Vz[n - 1]) * r40z - (Vz[n + 1] - Vz[n - 2]) * r41z;
        const F dxVy = ((Vy[n + si] - Vy[n]) * r40x - (Vy[n + 2 * si] - Vy[n - si]) * r41x +
                        (Vy[n + si - sj] - Vy[n - sj]) * r40x - (Vy[n + 2 * si - sj] - Vy[n - si - sj]) * r41x +
                        (Vy[n] - Vy[n - si]) * r40x - (Vy[n + si] - Vy[n - 2 * si]) * r41x +
                        (Vy[n - sj] - Vy[n - si - sj]) * r40x - (Vy[n + si - sj] - Vy[n - 2 * si - sj]) * r41x) / 4.0f;
        const F dxVz = ((Vz[n + si] - Vz[n]) * r40x - (Vz[n + 2 * si] - Vz[n - si]) * r41x +
                        (Vz[n - 1 + si] - Vz[n - 1]) * r40x - (Vz[n - 1 + 2 * si] - Vz[n - 1 - si]) * r41x +
                        (Vz[n] - Vz[n - si]) * r40x - (Vz[n + si] - Vz[n - 2 * si]) * r41x +
                        (Vz[n - 1] - Vz[n - 1 - si]) * r40x - (Vz[n - 1 + si] - Vz[n - 1 - 2 * si]) * r41x) / 4.0f;
        const F dyVx = ((Vx[n + sj] - Vx[n]) * r40y - (Vx[n + 2 * sj] - Vx[n - sj]) * r41y +
                        (Vx[n - si + sj] - Vx[n - si]) * r40y - (Vx[n - si + 2 * sj] - Vx[n - si - sj]) * r41y +
                        (Vx[n] - Vx[n - sj]) * r40y - (Vx[n + sj] - Vx[n - 2 * sj]) * r41y +
                        (Vx[n - si] - Vx[n - si - sj]) * r40y - (Vx[n - si + sj] - Vx[n - si - 2 * sj]) * r41y) / 4.0f;
        const F dyVz = ((Vz[n + sj] - Vz[n]) * r40y - (Vz[n + 2 * sj] - Vz[n - sj]) * r41y +
                        (Vz[n - 1 + sj] - Vz[n - 1]) * r40y - (Vz[n - 1 + 2 * sj] - Vz[n - 1 - sj]) * r41y +
                        (Vz[n] - Vz[n - sj]) * r40y - (Vz[n + sj] - Vz[n - 2 * sj]) * r41y +
                        (Vz[n - 1] - Vz[n - 1 - sj]) * r40y - (Vz[n - 1 + sj] - Vz[n - 1 - 2 * sj]) * r41y) / 4.0f;
        const F dzVx = ((Vx[n + 1] - Vx[n]) * r40z - (Vx[n + 2] - Vx[n - 1]) * r41z +
                        (Vx[n + 1 - si] - Vx[n - si]) * r40z - (Vx[n + 2 - si] - Vx[n - 1 - si]) * r41z +
                        (Vx[n] - Vx[n - 1]) * r40z - (Vx[n + 1] - Vx[n - 2]) * r41z +
                        (Vx[n - si] - Vx[n - 1 - si]) * r40z - (Vx[n + 1 - si] - Vx[n - 2 - si]) * r41z) / 4.0f;
        const F dzVy = ((Vy[n + 1] - Vy[n]) * r40z - (Vy[n + 2] - Vy[n - 1]) * r41z +
                        (Vy[n + 1 - sj] - Vy[n - sj]) * r40z - (Vy[n + 2 - sj] - Vy[n - 1 - sj]) * r41z +
                        (Vy[n] - Vy[n - 1]) * r40z - (Vy[n + 1] - Vy[n - 2]) * r41z +
                        (Vy[n - sj] - Vy[n - 1 - sj]) * r40z - (Vy[n + 1 - sj] - Vy[n - 2 - sj]) * r41z) / 4.0f;
        a[3] = a[3] + (float)(dxVx) * dt;
        a[4] = a[4] + (float)(dyVy) * dt;
        a[5] = a[5] + (float)(dzVz) * dt;
        a[6] = a[6] + (float)(dyVz + dzVy) / 2.0f * dt;
        a[7] = a[7] + (float)(dxVz + dzVx) / 2.0f * dt;
        a[8] = a[8] + (float)(dxVy + dyVx) / 2.0f * dt;
    }
    if (!w.sample) return;
    const long long ntw = w.ntw;
    const float M0 = w.M0, UC = w.UC;
    if (w.sw_v) {
        float *o = w.wav_v + ntw * 3 * s + (w.itw - 1);
        o[0] = (float)(Vx[n] + Vx[n - si]) / 2.0f * M0 * UC * 1e9f;
        o[ntw] = (float)(Vy[n] + Vy[n - sj]) / 2.0f * M0 * UC * 1e9f;
        o[2 * ntw] = -(float)(Vz[n] + Vz[n - 1]) / 2.0f * M0 * UC * 1e9f;
    }
    if (w.sw_u) {
        float *o = w.wav_u + ntw * 3 * s + (w.itw - 1);
        o[0] = a[0] * M0 * UC * 1e9f;
        o[ntw] = a[1] * M0 * UC * 1e9f;
        o[2 * ntw] = a[2] * M0 * UC * 1e9f;
    }
    if (w.sw_stress) {
        float *o = w.wav_s + ntw * 6 * s + (w.itw - 1);
        o[0] = (float)(p.Sxx[n]) * M0 * UC * 1e6f;
        o[ntw] = (float)(p.Syy[n]) * M0 * UC * 1e6f;
        o[2 * ntw] = (float)(p.Szz[n]) * M0 * UC * 1e6f;
        o[3 * ntw] = (float)(p.Syz[n] + p.Syz[n - sj] + p.Syz[n - 1] + p.Syz[n - 1 - sj]) / 4.0f * M0 * UC * 1e6f;
        o[4 * ntw] = (float)(p.Sxz[n] + p.Sxz[n - si] + p.Sxz[n - 1] + p.Sxz[n - 1 - si]) / 4.0f * M0 * UC * 1e6f;
        o[5 * ntw] = (float)(p.Sxy[n] + p.Sxy[n - sj] + p.Sxy[n - si] + p.Sxy[n - si - sj]) / 4.0f * M0 * UC * 1e6f;
    }
    if (w.sw_strain) {
        float *o = w.wav_e + ntw * 6 * s + (w.itw - 1);
#pragma unroll
        for (int c = 0; c < 6; c++) o[c * ntw] = a[3 + c] * M0 * UC * 1e-3f;
    }
}

// Green's-function mode, m_green.f90.  green__store (:404-521): nine displacement-gradient sums per grid point (4th-order
// differences, the six off-diagonal ones averaged over the four surrounding staggered nodes) and, with green_bforce,
// three displacement sums; every ntdec_w steps the sums are written to gf(itw, (i-1)*ncmp + c) (:523-547).
struct GreenParams {
    int ng, ncmp, ntw, itw, sample, bforce;
    const int *ijk;          // 3*ng: LOCAL memory-box indices (mi, mj) and k
    float *acc;              // 12 per point: dxUx dxUy dxUz dyUx dyUy dyUz dzUx dzUy dzUz Ux Uy Uz
    float *gf;               // (ntw, ncmp*ng), itw fastest
    double r40[3], r41[3];   // C40/d, C41/d in the field kind (:111-116)
};

template <typename F>
__global__ void green_store_kernel(const __grid_constant__ KParams<F> p, const GreenParams g) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= g.ng) return;
    const long long si = p.SI, sj = p.SJ;
    const long long n = (long long)(g.ijk[3 * s + 2] + KOFF - 1) + (long long)p.NZP * ((long long)g.ijk[3 * s] + (long long)p.NXM * g.ijk[3 * s + 1]);
    const F *Vx = p.Vx, *Vy = p.Vy, *Vz = p.Vz;
    const float dt = p.dt;
    const F r40x = (F)g.r40[0], r40y = (F)g.r40[1], r40z = (F)g.r40[2], r41x = (F)g.r41[0], r41y = (F)g.r41[1], r41z = (F)g.r41[2];
    // d<a>V<b> at offset o: 4th-order difference of component b along a, staggered forward (fwd) or backward
    auto dxb = [&](const F *V, long long o) { return (V[n + o] - V[n + o - si]) * r40x - (V[n + o + si] - V[n + o - 2 * si]) * r41x; };
    auto dxf = [&](const F *V, long long o) { return (V[n + o + si] - V[n + o]) * r40x - (V[n + o + 2 * si] - V[n + o - si]) * r41x; };
    auto dyb = [&](const F *V, long long o) { return (V[n + o] - V[n + o - sj]) * r40y - (V[n + o + sj] - V[n + o - 2 * sj]) * r41y; };
    auto dyf = [&](const F *V, long long o) { return (V[n + o + sj] - V[n + o]) * r40y - (V[n + o + 2 * sj] - V[n + o - sj]) * r41y; };
    auto dzb = [&](const F *V, long long o) { return (V[n + o] - V[n + o - 1]) * r40z - (V[n + o + 1] - V[n + o - 2]) * r41z; };
    auto dzf = [&](const F *V, long long o) { return (V[n + o + 1] - V[n + o]) * r40z - (V[n + o + 2] - V[n + o - 1]) * r41z; };
    const F dxVx = dxb(Vx, 0), dyVy = dyb(Vy, 0), dzVz = dzb(Vz, 0);
    const F dxVy = (dxf(Vy, 0) + dxf(Vy, -sj) + dxb(Vy, 0) + dxb(Vy, -sj)) * 0.25f;
    const F dxVz = (dxf(Vz, 0) + dxf(Vz, -1) + dxb(Vz, 0) + dxb(Vz, -1)) * 0.25f;
    const F dyVx = (dyf(Vx, 0) + dyf(Vx, -si) + dyb(Vx, 0) + dyb(Vx, -si)) * 0.25f;
    const F dyVz = (dyf(Vz, 0) + dyf(Vz, -1) + dyb(Vz, 0) + dyb(Vz, -1)) * 0.25f;
    const F dzVx = (dzf(Vx, 0) + dzf(Vx, -si) + dzb(Vx, 0) + dzb(Vx, -si)) * 0.25f;
    const F dzVy = (dzf(Vy, 0) + dzf(Vy, -sj) + dzb(Vy, 0) + dzb(Vy, -sj)) * 0.25f;
    float *a = g.acc + 12 * (long long)s;
    a[0] = a[0] + (float)(dxVx * dt); a[1] = a[1] + (float)(dxVy * dt); a[2] = a[2] + (float)(dxVz * dt);
    a[3] = a[3] + (float)(dyVx * dt); a[4] = a[4] + (float)(dyVy * dt); a[5] = a[5] + (float)(dyVz * dt);
    a[6] = a[6] + (float)(dzVx * dt); a[7] = a[7] + (float)(dzVy * dt); a[8] = a[8] + (float)(dzVz * dt);
    if (g.bforce) {
        a[9] = a[9] + 0.5f * (float)(Vx[n] + Vx[n - si]) * dt;
        a[10] = a[10] + 0.5f * (float)(Vy[n] + Vy[n - sj]) * dt;
        a[11] = a[11] + 0.5f * (float)(Vz[n] + Vz[n - 1]) * dt;
    }
    if (!g.sample) return;
    const float UC_BF = 1e-12f, UC_DERIV = 1e-15f;   // m_green.f90:30-31
    const long long nt = g.ntw;
    float *o = g.gf + nt * g.ncmp * s + (g.itw - 1);
    o[0] = a[0] * UC_DERIV * 1e9f;
    o[nt] = a[4] * UC_DERIV * 1e9f;
    o[2 * nt] = a[8] * UC_DERIV * 1e9f;
    o[3 * nt] = (a[5] + a[7]) * UC_DERIV * 1e9f;
    o[4 * nt] = (a[2] + a[6]) * UC_DERIV * 1e9f;
    o[5 * nt] = (a[1] + a[3]) * UC_DERIV * 1e9f;
    if (g.bforce) {
        o[6 * nt] = a[9] * UC_BF * 1e9f;
        o[7 * nt] = a[10] * UC_BF * 1e9f;
        o[8 * nt] = a[11] * UC_BF * 1e9f;
    }
}

// green__source m_green.f90:606-649: unit body force at the pseudo source (a station), six single-thread adds
template <typename F>
__global__ void green_source_kernel(const __grid_constant__ KParams<F> p, int mi, int mj, int k, float fx, float fy, float fz) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const long long si = p.SI, sj = p.SJ;
    const long long n = (long long)(k + KOFF - 1) + (long long)p.NZP * ((long long)mi + (long long)p.NXM * mj);
    const float *rho = p.rho;
    p.Vx[n] = p.Vx[n] + (2.0f / (rho[n] + rho[n + si])) * fx / 2;
    p.Vx[n - si] = p.Vx[n - si] + (2.0f / (rho[n] + rho[n - si])) * fx / 2;
    p.Vy[n] = p.Vy[n] + (2.0f / (rho[n] + rho[n + sj])) * fy / 2;
    p.Vy[n - sj] = p.Vy[n - sj] + (2.0f / (rho[n] + rho[n - sj])) * fy / 2;
    p.Vz[n] = p.Vz[n] + (2.0f / (rho[n] + rho[n + 1])) * fz / 2;
    p.Vz[n - 1] = p.Vz[n - 1] + (2.0f / (rho[n] + rho[n - 1])) * fz / 2;
}

// kernel__vmax m_kernel.f90:360-372: max |V| at k = kob(i,j)+1 over the given local (i,j) window
template <typename F>
__global__ void vmax_kernel(const __grid_constant__ KParams<F> p, int li0, int li1, int lj0, int lj1, unsigned int *out3) {
    const int ni = li1 - li0 + 1, nj = lj1 - lj0 + 1;
    float xm = 0.0f, ym = 0.0f, zm = 0.0f;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < (long long)ni * nj; t += (long long)gridDim.x * blockDim.x) {
        const int li = li0 + (int)(t % ni), lj = lj0 + (int)(t / ni);
        const long long col = (long long)(li + HALO) + (long long)p.NXM * (lj + HALO);
        const long long n = (long long)(p.kob[col] + 1 + KOFF - 1) + (long long)p.NZP * col;
        xm = fmaxf(xm, (float)fabs((double)p.Vx[n]));
        ym = fmaxf(ym, (float)fabs((double)p.Vy[n]));
        zm = fmaxf(zm, (float)fabs((double)p.Vz[n]));
    }
    for (int o = 16; o > 0; o >>= 1) {
        xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, o));
        ym = fmaxf(ym, __shfl_xor_sync(0xffffffffu, ym, o));
        zm = fmaxf(zm, __shfl_xor_sync(0xffffffffu, zm, o));
    }
    if ((threadIdx.x & 31) == 0) {   // non-negative floats order like their bit patterns
        atomicMax(out3 + 0, __float_as_uint(xm));
        atomicMax(out3 + 1, __float_as_uint(ym));
        atomicMax(out3 + 2, __float_as_uint(zm));
    }
}

// ------------------------------------------------------------------------------------------------
// halo planes, m_global.f90:391-607.  A "plane list" names up to 5 (field, plane index) pairs; the
// message layout is the reference's: plane-major, then (j-jbeg)*nz + (k-1) for x faces /
// (i-ibeg)*nz + (k-1) for y faces, k = 1..nz only, no corners.
struct PlaneList {
    int n;
    void *field[5];
    int m[5];        // memory-box index (mi for x faces, mj for y faces) of the plane
};

template <typename F, bool XFACE, bool PACK>
__global__ void halo_kernel(int nz, int nline, int NZP, int NXM, const PlaneList pl, F *buf) {
    // one thread per (k, line) ; line = owned j (x face) or owned i (y face)
    const int k = blockIdx.x * blockDim.x + threadIdx.x;   // 0-based
    const int line = blockIdx.y;
    const int s = blockIdx.z;
    if (k >= nz || s >= pl.n) return;
    F *f = (F *)pl.field[s];
    long long col;
    if (XFACE) col = (long long)pl.m[s] + (long long)NXM * (line + HALO);
    else col = (long long)(line + HALO) + (long long)NXM * pl.m[s];
    const long long n = (long long)(k + KOFF) + (long long)NZP * col;
    const long long b = (long long)s * nline * nz + (long long)line * nz + k;
    if (PACK) buf[b] = f[n];
    else f[n] = buf ? buf[b] : F(0);   // no buffer: the never-written, zero-initialised rbuf of an MPI_PROC_NULL neighbour
}

// ------------------------------------------------------------------------------------------------
// The same exchange without a library in between (one node, NVLink / NVSwitch peer memory): the "pack" kernel of a face stores
// its planes STRAIGHT into the neighbour's receive buffer -- one hop over NVLink, no send buffer, no copy kernel of a
// communication library -- and, when its last block is done, publishes the exchange's sequence number in the neighbour's flag
// word with a system-scope release store.  The neighbour's "unpack" kernel spins on that flag (acquire) before it reads.
// Messages keep the reference's layout (plane lists of m_global.f90:416-443 ...), so both paths fill the halo identically.
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All faces of a rank in ONE launch (blockIdx.z runs over the (face, plane) pairs), 16 bytes per thread.
struct P2pFaces {
    PlaneList pl[4];          // push: the planes sent through face f; pull: the halo planes filled through face f
    void *buf[4];             // push: the neighbour's receive buffer (peer memory); pull: this rank's receive buffer
    unsigned int *flag[4];    // push: the neighbour's flag word (peer memory); pull: unused
    unsigned int *count[4];   // push: block counter of face f (own memory)
    int nline[4];             // owned j (x faces: f = 0, 1) or owned i (y faces: f = 2, 3) along the face; 0 = no neighbour
    int zend[4];              // running sum of pl[f].n over the faces with a neighbour
};

template <typename F, bool PUSH>
__global__ void halo_p2p(int nz, int NZP, int NXM, const __grid_constant__ P2pFaces q, unsigned int seq) {
    constexpr int VEC = 16 / (int)sizeof(F);
    int f = 0;
    while (f < 3 && (int)blockIdx.z >= q.zend[f]) f++;
    const int s = (int)blockIdx.z - (f > 0 ? q.zend[f - 1] : 0);
    const int line = blockIdx.y;
    const int nline = q.nline[f];
    if (line >= nline) return;                                 // (uniform per block; such blocks are not counted)
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;   // 0-based
    const PlaneList &pl = q.pl[f];
    F *fld = (F *)pl.field[s];
    long long col;
    if (f < 2) col = (long long)pl.m[s] + (long long)NXM * (line + HALO);
    else col = (long long)(line + HALO) + (long long)NXM * pl.m[s];
    const long long n = (long long)(k + KOFF) + (long long)NZP * col;
    const long long b = (long long)s * nline * nz + (long long)line * nz + k;
    F *buf = (F *)q.buf[f];
    const bool vec = (nz % VEC) == 0;                          // rows of the message stay 16-byte aligned
    if (k < nz) {
        if (vec) {
            if (PUSH) *reinterpret_cast<int4 *>(buf + b) = *reinterpret_cast<const int4 *>(fld + n);
            else *reinterpret_cast<int4 *>(fld + n) = __ldcg(reinterpret_cast<const int4 *>(buf + b));
        } else {
            for (int e = 0; e < VEC && k + e < nz; e++) {
                if (PUSH) buf[b + e] = fld[n + e];
                else fld[n + e] = __ldcg(buf + b + e);
            }
        }
    }
    if (!PUSH) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();                                  // this block's stores before the ticket
        const unsigned int total = gridDim.x * (unsigned int)nline * (unsigned int)pl.n;
        if (atomicAdd(q.count[f], 1u) == total - 1) {             // the last block of this face
            *q.count[f] = 0;
            __threadfence_system();
            st_release_sys(q.flag[f], seq);
        }
    }
}

// one warp waits for the flags of up to four faces, so that the pull kernel behind it in the stream need not spin in every block.
// A neighbour that never pushes (it died, or left the time loop) must not leave this kernel spinning for ever: after timeout_ns
// (0 = none) the lane gives up and raises *err, which the host reads at its next wait (stream_wait) and turns into the abort.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
static __global__ void halo_wait(const unsigned int *f0, const unsigned int *f1, const unsigned int *f2, const unsigned int *f3, unsigned int seq,
                                 unsigned long long timeout_ns, unsigned int *err) {
    const unsigned int *f = threadIdx.x == 0 ? f0 : threadIdx.x == 1 ? f1 : threadIdx.x == 2 ? f2 : threadIdx.x == 3 ? f3 : nullptr;
    if (!f || *(volatile unsigned int *)err) return;   // (an earlier exchange already gave up: do not wait again)
    const unsigned long long t0 = globaltimer_ns();
    while ((int)(ld_acquire_sys(f) - seq) < 0) {
        __nanosleep(200);
        if (timeout_ns && globaltimer_ns() - t0 > timeout_ns) {
            atomicExch(err, 1u + threadIdx.x);
            return;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Plane-wave mode: horizontal zero-derivative boundary (m_absorb_p.f90:137-243 / :332-424).  Linear extrapolation into
// the first plane outside the model on the outer ranks, owned rows and k = 1..nz only.  blockIdx.z = edge
// (0: i=0, 1: i=nx+1, 2: j=0, 3: j=ny+1); `dst/s1/s2[e]` are memory-box indices along the edge's normal, < 0 = edge off.
struct PwEdges { int dst[4], s1[4], s2[4]; int nrow[4]; };

template <typename F>
__global__ void pw_edge_kernel(F *fields, long long ncell, int nf, int nz, int NZP, int NXM, PwEdges e) {
    const int edge = blockIdx.z;
    if (e.dst[edge] < 0) return;
    const int k = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (k > nz || r >= e.nrow[edge]) return;
    const long long SI = NZP, SJ = (long long)NZP * NXM;
    const long long kk = k + KOFF - 1;
    long long d, a, b;
    if (edge < 2) { const long long row = (long long)(r + HALO) * SJ + kk; d = row + e.dst[edge] * SI; a = row + e.s1[edge] * SI; b = row + e.s2[edge] * SI; }
    else { const long long row = (long long)(r + HALO) * SI + kk; d = row + e.dst[edge] * SJ; a = row + e.s1[edge] * SJ; b = row + e.s2[edge] * SJ; }
    for (int q = 0; q < nf; q++) {
        F *f = fields + (long long)q * ncell;
        f[d] = 2 * f[a] - f[b];
    }
}

}   // namespace swpc
