// bottom_tma.cuh -- the bottom 32 rows of the interior columns, k = nz-31 .. nz, as WHOLE cache lines.
//
// Why: the absorber's bottom slab (k = kend_k+1 .. nz: 20 rows with the reference's default na) starts in the middle of a
// 128-byte line (96 bytes in for float64, 48 for the float arrays).  HBM is fetched in 64-byte pieces, so a kernel that sweeps
// only those rows moves 1.5x the bytes it needs, and the lines it shares with the rows above are fetched once more by the interior
// kernels -- 2.3 GB per step at 1024 x 1024 x 512, the largest single item of over-fetch left.  Here ONE
// tile covers the whole lines: 32 rows x 8 columns marching along j, whose first RI = kend_k - (nz-32) rows are interior cells
// (kernel__update_stress / _vel with memory variables, 4th-order stencils) and whose last 32 - RI rows are PML cells
// (absorb_p__update_*).  Threads are numbered flat over the cells of each kind, so the two bodies run in different warps (no
// divergence); both are the accessor-templated bodies of kernels.cuh, so results are bit-identical to the other kernels'.
// The interior kernels (stress_tma, vel_ring) then stop at k = nz-32 and pml_tma keeps only the walls.
//
// Pipeline as stress_tma_p: persistent blocks, work items (a tile column x a run of planes) from a ticket counter, TMA loads under
// one full-barrier per plane, rings indexed by the block's running step count, a drain of the pipeline before an item's ring
// prologue.  Per plane (stress sweep):
//   S box (32, 8, 1, 6)      R box (RIB, 8, 1, 6 NM) rows of the interior cells only      lam [taup taus] box (32, 8, 1, 1|3)
//   aux box (NA, 8, 1, 9) rows of the PML cells only         V box with halo (40, 12, 1, 3) of plane j+2 -> 5-plane ring (the PML
//   cells use planes j-1 .. j+1 of it)                        mu box (36, 9) of plane j+1 -> 2-plane ring
// velocity sweep:
//   V box (32, 8, 1, 3)      aux box (NA, 8, 1, 9)           Sxx Szz Sxz halo box (40, 12, 1, 3) of plane j
//   Syy Syz Sxy halo box of plane j+2 -> 5-plane ring        rho box (36, 9) of plane j+1 -> 2-plane ring
// RIB = RI rounded up to 4 rows, the aux box starts at RIA = RI rounded down to 4 rows (16-byte alignment of TMA boxes).
#pragma once

#include "pml_tma.cuh"

namespace swpc {

struct BotItem {
    int li0, ncol;       // first local column, active columns (<= 8)
    int lj0, nsteps;     // first local plane, planes to march
    int ai, aj;          // aux tensor coordinates (column, plane) of (li0, lj0)
    long long aux0;      // linear aux index of (row k0 + RIA, li0, lj0)
};

struct BotMaps {
    CUtensorMap C;       // centre box over the field tensor: (32, 8, 1, 6) stress / (32, 8, 1, 3) velocity
    CUtensorMap H;       // halo box over the field tensor: (40, 12, 1, 3)
    CUtensorMap M;       // medium centre box (32, 8, 1, NMED): lam [, taup, taus]  (stress only)
    CUtensorMap Mh;      // medium halo box (36, 9, 1, 1): mu (stress) / rho (velocity)
    CUtensorMap R;       // memory variables (RIB, 8, 1, 6 NM)  (stress only, NM > 0)
    CUtensorMap aux;     // (NA, 8, 1, 9) over the bottom columns' part of the aux arrays
};

struct BotGeom {
    int nitems;
    int k0;              // first row of the tile = nz - 31
    int RI;              // interior rows of the tile: cells k0 .. k0 + RI - 1
    int RIB, RIA, NA;    // R box rows; first row / rows of the aux box
    int ak;              // aux tensor coordinate (row within the column) of tile row RIA
    int asi;             // aux elements between neighbouring columns
    long long asj;       // aux elements between neighbouring planes
    int m_first, mh_index;   // stress: lam = 2 (taup, taus follow), mu = 1;  velocity: -, rho = 0
    int c_first, h_first, sa_first, a_first;   // as PmlGeom
    int nint, nint32;    // interior cells of the tile (RI * 8) and that rounded up to whole warps
};

template <typename F, int NM, bool STRESS>
struct BotCfg {
    static constexpr int TK = 32, TI = 8, NS = 3;
    static constexpr int NCW = 16;                     // consumer warps (roles are dealt at run time, see bottom_tma)
    static constexpr int THREADS = (NCW + 1) * 32;
    static constexpr int HK = 4, VK = TK + 2 * HK, VI = TI + 4;
    static constexpr int MUK = TK + 4, MUI = TI + 1;
    static constexpr int NC = STRESS ? 6 : 3;
    static constexpr int NMED = (NM > 0) ? 3 : 1;
    static constexpr int C_BYTES = NC * TK * TI * (int)sizeof(F);
    static constexpr int RA_MAX = (STRESS ? 6 * NM * TK * TI * 4 : 0);         // R box, worst case (RI = 32)
    static constexpr int AX_MAX = 9 * TK * TI * 4;                             // aux box, worst case (RI = 0)
    static constexpr int M_BYTES = STRESS ? NMED * TK * TI * 4 : 0;
    static constexpr int H_BYTES = 3 * VK * VI * (int)sizeof(F), H_STRIDE = align128(H_BYTES);
    static constexpr int MU_BYTES = MUK * MUI * 4, MU_STRIDE = align128(MU_BYTES);
    // R box and aux box share one area: RIB + NA <= 36 rows, so the worst case is RI = 31 (a full R box, 4 aux rows); without
    // memory variables (velocity sweep, NM = 0) it is the full aux box
    static constexpr int RAX = (RA_MAX > 0) ? align128(RA_MAX) + align128(9 * 4 * TI * 4) + 128 : align128(AX_MAX);
    static constexpr int R_OFF = align128(C_BYTES);                            // R box, then (at R_OFF + align128(R bytes)) the aux box
    static constexpr int M_OFF = R_OFF + RAX;
    static constexpr int SA_OFF = M_OFF + align128(M_BYTES);                   // velocity: in-plane S halo box
    static constexpr int STAGE = SA_OFF + (STRESS ? 0 : H_STRIDE);
    static constexpr int NV = NS + 4, NMU = NS + 1;
    static constexpr int V_OFF = NS * STAGE, MU_OFF = V_OFF + NV * H_STRIDE, BAR_OFF = MU_OFF + NMU * MU_STRIDE;
    static constexpr int SMEM = BAR_OFF + 128;
};

// ---- accessors: operands in shared memory, results to global memory.  Interior cells, stress sweep:
template <typename F, int NM>
struct AccBotInt {
    using C = BotCfg<F, NM, true>;
    const KParams<F> &p;
    const F *v[5];
    const float *mu0, *mu1;
    const F *s;
    const float *r;
    const float *m;
    int rstride;            // elements between two R arrays: 8 * RIB
    long long n;
    __device__ __forceinline__ AccBotInt(const KParams<F> &p_) : p(p_) {}
    template <int f, int dk, int di, int dj> __device__ __forceinline__ F V() const { return v[dj + 2][f * (C::VI * C::VK) + di * C::VK + dk]; }
    template <int dk, int di, int dj> __device__ __forceinline__ float mu() const { return (dj == 0 ? mu0 : mu1)[di * C::MUK + dk]; }
    __device__ __forceinline__ float lam() const { return m[0]; }
    __device__ __forceinline__ float taup() const { return m[1 * C::TI * C::TK]; }
    __device__ __forceinline__ float taus() const { return m[2 * C::TI * C::TK]; }
    __device__ __forceinline__ F *sptr(int c) const { return c == 0 ? p.Sxx : c == 1 ? p.Syy : c == 2 ? p.Szz : c == 3 ? p.Syz : c == 4 ? p.Sxz : p.Sxy; }
    __device__ __forceinline__ F S(int c) const { return s[(c == 0 ? 0 : c == 1 ? 3 : c == 2 ? 1 : c == 3 ? 4 : c == 4 ? 2 : 5) * (C::TI * C::TK)]; }
    __device__ __forceinline__ void setS(int c, F val) const { sts_(sptr(c) + n, val); }
    __device__ __forceinline__ float R(int q) const { return r[q * rstride]; }
    __device__ __forceinline__ void setR(int q, float val) const { sts_(p.R + n + q * p.ncell, val); }
};

// PML cells, both sweeps (interface of stress_pml_t / vel_pml_t)
template <typename F, int NM, bool STRESS>
struct AccBotPml {
    using C = BotCfg<F, NM, STRESS>;
    const KParams<F> &p;
    const F *h[3];          // ring planes j-1, j, j+1 of the halo'd fields [3][VI][VK]
    const F *sa;            // velocity: Sxx Szz Sxz of plane j
    const float *m0, *m1;   // mu / rho planes j, j+1
    const F *c;             // centre box [NC][8][32]
    const float *ax;        // aux box [9][8][NA]
    const float *lm;        // lam [8][32]
    int astride;            // elements between two aux arrays: 8 * NA
    long long n, an;
    __device__ __forceinline__ AccBotPml(const KParams<F> &p_) : p(p_) {}
    static constexpr int PH = C::VI * C::VK;
    template <int f, int dk, int di, int dj> __device__ __forceinline__ F V() const { return h[dj + 1][f * PH + di * C::VK + dk]; }
    template <int dk, int di, int dj> __device__ __forceinline__ float mu() const { return (dj == 0 ? m0 : m1)[di * C::MUK + dk]; }
    __device__ __forceinline__ float lam() const { return lm[0]; }
    __device__ __forceinline__ F *sptr(int q) const { return q == 0 ? p.Sxx : q == 1 ? p.Syy : q == 2 ? p.Szz : q == 3 ? p.Syz : q == 4 ? p.Sxz : p.Sxy; }
    __device__ __forceinline__ F S(int q) const { return c[(q == 0 ? 0 : q == 1 ? 3 : q == 2 ? 1 : q == 3 ? 4 : q == 4 ? 2 : 5) * (C::TI * C::TK)]; }
    __device__ __forceinline__ void setS(int q, F v) const { sts_(sptr(q) + n, v); }
    __device__ __forceinline__ float aux(int q) const { return ax[(q - (STRESS ? 0 : 9)) * astride]; }
    __device__ __forceinline__ void setAux(int q, float v) const { sts_(p.aux + an + q * p.naux, v); }
    template <int q, int dk, int di, int dj> __device__ __forceinline__ F Sn() const {
        if (q == 0 || q == 2 || q == 4) return sa[(q == 0 ? 0 : q == 2 ? 1 : 2) * PH + di * C::VK + dk];
        return h[dj + 1][(q == 1 ? 0 : q == 3 ? 1 : 2) * PH + di * C::VK + dk];
    }
    template <int dk, int di, int dj> __device__ __forceinline__ float rho() const { return (dj == 0 ? m0 : m1)[di * C::MUK + dk]; }
    __device__ __forceinline__ F Vc(int f) const { return c[f * (C::TI * C::TK)]; }
    __device__ __forceinline__ void setV(int f, F v) const { sts_((f == 0 ? p.Vx : f == 1 ? p.Vy : p.Vz) + n, v); }
};

// interior cells, velocity sweep (interface of vel_interior_calc)
template <typename F, int NM>
struct AccBotVelInt {
    using C = BotCfg<F, NM, false>;
    const KParams<F> &p;
    const F *sa;            // [3][VI][VK] plane j: Sxx Szz Sxz
    const F *sb[5];         // planes j-2 .. j+2: Syy Syz Sxy
    const float *rho0, *rho1;
    const F *v;             // [3][8][32]
    long long n;
    __device__ __forceinline__ AccBotVelInt(const KParams<F> &p_) : p(p_) {}
    template <int c, int dk, int di, int dj> __device__ __forceinline__ F S() const {
        constexpr int PL = C::VI * C::VK;
        if (c == 0 || c == 2 || c == 4) return sa[(c == 0 ? 0 : c == 2 ? 1 : 2) * PL + di * C::VK + dk];
        return sb[dj + 2][(c == 1 ? 0 : c == 3 ? 1 : 2) * PL + di * C::VK + dk];
    }
    template <int dk, int di, int dj> __device__ __forceinline__ float rho() const { return (dj == 0 ? rho0 : rho1)[di * C::MUK + dk]; }
    __device__ __forceinline__ F V(int f) const { return v[f * (C::TI * C::TK)]; }
    __device__ __forceinline__ void setV(int f, F val) const { sts_((f == 0 ? p.Vx : f == 1 ? p.Vy : p.Vz) + n, val); }
};

template <typename F, int NM, bool STRESS>
__global__ void __launch_bounds__((BotCfg<F, NM, STRESS>::THREADS), 1)
bottom_tma(const __grid_constant__ KParams<F> p, const __grid_constant__ BotMaps tm, const BotItem *__restrict__ items, unsigned int *ticket,
           unsigned int ticket_base, const BotGeom g) {
    using C = BotCfg<F, NM, STRESS>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::BAR_OFF);
    uint64_t *empty = full + C::NS;
    volatile int2 *meta = reinterpret_cast<volatile int2 *>(smem + C::BAR_OFF + 64);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r_bytes = (STRESS && NM > 0) ? 6 * NM * g.RIB * C::TI * 4 : 0;
    const int ax_off = C::R_OFF + (r_bytes + 127) / 128 * 128;
    const int ax_bytes = 9 * g.NA * C::TI * 4;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NS; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], C::NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == C::NCW) {
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            const uint32_t step_tx = C::C_BYTES + r_bytes + ax_bytes + C::M_BYTES + (STRESS ? 0 : C::H_BYTES) + C::H_BYTES + C::MU_BYTES;
            int gs = 0;
            for (;;) {
                const unsigned int it = atomicAdd(ticket, 1u) - ticket_base;
                if (it >= (unsigned int)g.nitems) break;
                const BotItem I = items[it];
                const int ck = g.k0 + KOFF - 1, ci = I.li0 + HALO, cj = I.lj0 + HALO;
                for (int t = 0; t < I.nsteps; t++, gs++) {
                    const int s = gs % C::NS;
                    if (gs >= C::NS) mbar_wait(&empty[s], ((gs / C::NS) - 1) & 1);
                    unsigned char *st = smem + s * C::STAGE;
                    meta[s].x = (int)it; meta[s].y = t;
                    mbar_expect_tx(&full[s], step_tx + (t == 0 ? 4 * C::H_BYTES + C::MU_BYTES : 0));
                    tma_load_4d(st, &tm.C, &full[s], ck, ci, cj + t, g.c_first);
                    if (STRESS && NM > 0) tma_load_4d(st + C::R_OFF, &tm.R, &full[s], ck, ci, cj + t, 0);
                    if (g.NA > 0) tma_load_4d(st + ax_off, &tm.aux, &full[s], g.ak, I.ai, I.aj + t, g.a_first);
                    if (STRESS) tma_load_4d(st + C::M_OFF, &tm.M, &full[s], ck, ci, cj + t, g.m_first);
                    else tma_load_4d(st + C::SA_OFF, &tm.H, &full[s], ck - C::HK, ci - 2, cj + t, g.sa_first);
                    if (t == 0) {
                        for (int q = C::NS - 1; q >= 1; q--) {   // drain: every earlier step released before the prologue touches the rings
                            const int gp = gs - q;
                            if (gp >= 0) mbar_wait(&empty[gp % C::NS], (gp / C::NS) & 1);
                        }
                        for (int q = 0; q < 4; q++)
                            tma_load_4d(smem + C::V_OFF + ((gs + q) % C::NV) * C::H_STRIDE, &tm.H, &full[s], ck - C::HK, ci - 2, cj - 2 + q, g.h_first);
                        tma_load_4d(smem + C::MU_OFF + (gs % C::NMU) * C::MU_STRIDE, &tm.Mh, &full[s], ck, ci, cj, g.mh_index);
                    }
                    tma_load_4d(smem + C::V_OFF + ((gs + 4) % C::NV) * C::H_STRIDE, &tm.H, &full[s], ck - C::HK, ci - 2, cj + t + 2, g.h_first);
                    tma_load_4d(smem + C::MU_OFF + ((gs + 1) % C::NMU) * C::MU_STRIDE, &tm.Mh, &full[s], ck, ci, cj + t + 1, g.mh_index);
                }
            }
            const int s = gs % C::NS;
            if (gs >= C::NS) mbar_wait(&empty[s], ((gs / C::NS) - 1) & 1);
            meta[s].x = -1; meta[s].y = 0;
            mbar_arrive(&full[s]);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers.  Roles by thread index:
    //   stress:   [0, nint) interior cells, normal components; [nint32, nint32 + nint) interior cells, shear components;
    //             [2 nint32, 2 nint32 + npml) PML cells           velocity: [0, nint) interior cells; [nint32, nint32 + npml) PML cells
    const int tid = threadIdx.x;
    const int npml = (C::TK - g.RI) * C::TI;
    int role = -1, cell = 0;   // 0 interior (normal / all), 1 interior shear, 2 PML
    if (tid < g.nint) { role = 0; cell = tid; }
    else if (STRESS && tid >= g.nint32 && tid < g.nint32 + g.nint) { role = 1; cell = tid - g.nint32; }
    else {
        const int b = (STRESS ? 2 : 1) * g.nint32;
        if (tid >= b && tid < b + npml) { role = 2; cell = tid - b; }
    }
    int r = 0, c = 0;
    if (role == 0 || role == 1) { c = cell / g.RI; r = cell - c * g.RI; }
    else if (role == 2) { const int np = C::TK - g.RI; c = cell / np; r = g.RI + (cell - c * np); }
    const int k = g.k0 + r;
    const int offC = c * C::TK + r, offH = (c + 2) * C::VK + (r + C::HK), offM = c * C::MUK + r;
    const int offR = c * g.RIB + r, offA = c * g.NA + (r - g.RIA);
    float4 gzc = make_float4(0, 0, 0, 0), gze = gzc, gxc = gzc, gxe = gzc;
    if (role == 2) { gzc = ldro(p.gzc + (k - 1)); gze = ldro(p.gze + (k - 1)); }
    bool active = false;
    int mi = 0, mj = 0, lj = 0;
    long long col = 0, n = 0, an = 0;
    for (int gs = 0;; gs++) {
        const int s = gs % C::NS;
        // the per-plane values that live in global memory (band of the column, y profile of the plane) are fetched BEFORE the wait:
        // col / lj already point at this step unless a new item starts, so their L2 latency hides behind the barrier
        int4 bnd = make_int4(0, 0, 0, 0);
        float4 gyc = make_float4(0, 0, 0, 0), gye = gyc;
        if (active) {
            if (role == 2) { const int q = min(lj, p.nyp - 1); gyc = ldro(p.gyc + q); gye = ldro(p.gye + q); }   // (lj = one past the item's last plane after its last step)
            else bnd = p.band[col];
        }
        mbar_wait(&full[s], (gs / C::NS) & 1);
        int it = 0, t = 0;
        if (lane == 0) { it = meta[s].x; t = meta[s].y; }   // the lane that arrives on the empty-barrier is the one that reads the mailbox
        it = __shfl_sync(0xffffffffu, it, 0);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (it < 0) break;
        if (t == 0) {
            const BotItem I = items[it];
            active = role >= 0 && c < I.ncol;
            const int li = I.li0 + c;
            mi = li + HALO; mj = I.lj0 + HALO; lj = I.lj0;
            col = (long long)mi + (long long)p.NXM * mj;
            n = (long long)(k + KOFF - 1) + (long long)p.NZP * col;
            an = I.aux0 + (r - g.RIA) + (long long)c * g.asi;
            if (active) {
                if (role == 2) { gxc = ldro(p.gxc + li); gxe = ldro(p.gxe + li); gyc = ldro(p.gyc + lj); gye = ldro(p.gye + lj); }
                else bnd = p.band[col];
            }
        }
        const unsigned char *st = smem + s * C::STAGE;
        if (active) {
            if (role == 2) {
                AccBotPml<F, NM, STRESS> a(p);
#pragma unroll
                for (int q = 0; q < 3; q++) a.h[q] = reinterpret_cast<const F *>(smem + C::V_OFF + ((gs + 1 + q) % C::NV) * C::H_STRIDE) + offH;
                a.m0 = reinterpret_cast<const float *>(smem + C::MU_OFF + (gs % C::NMU) * C::MU_STRIDE) + offM;
                a.m1 = reinterpret_cast<const float *>(smem + C::MU_OFF + ((gs + 1) % C::NMU) * C::MU_STRIDE) + offM;
                a.c = reinterpret_cast<const F *>(st) + offC;
                a.ax = reinterpret_cast<const float *>(st + ax_off) + offA;
                a.lm = reinterpret_cast<const float *>(st + C::M_OFF) + offC;
                a.sa = reinterpret_cast<const F *>(st + C::SA_OFF) + offH;
                a.astride = C::TI * g.NA;
                a.n = n; a.an = an;
                if (STRESS) stress_pml_t<F>(p, a, gxc, gxe, gyc, gye, gzc, gze);
                else vel_pml_t<F>(p, a, gxc, gxe, gyc, gye, gzc, gze);
            } else if (STRESS) {
                AccBotInt<F, NM> a(p);
#pragma unroll
                for (int q = 0; q < 5; q++) a.v[q] = reinterpret_cast<const F *>(smem + C::V_OFF + ((gs + q) % C::NV) * C::H_STRIDE) + offH;
                a.mu0 = reinterpret_cast<const float *>(smem + C::MU_OFF + (gs % C::NMU) * C::MU_STRIDE) + offM;
                a.mu1 = reinterpret_cast<const float *>(smem + C::MU_OFF + ((gs + 1) % C::NMU) * C::MU_STRIDE) + offM;
                a.s = reinterpret_cast<const F *>(st) + offC;
                a.r = reinterpret_cast<const float *>(st + C::R_OFF) + offR;
                a.m = reinterpret_cast<const float *>(st + C::M_OFF) + offC;
                a.rstride = C::TI * g.RIB;
                a.n = n;
                if (role == 1) stress_interior_t<F, NM, AccBotInt<F, NM>, false, true>(p, a, k, mi, mj, bnd);
                else stress_interior_t<F, NM, AccBotInt<F, NM>, true, false>(p, a, k, mi, mj, bnd);
            } else {
                AccBotVelInt<F, NM> a(p);
#pragma unroll
                for (int q = 0; q < 5; q++) a.sb[q] = reinterpret_cast<const F *>(smem + C::V_OFF + ((gs + q) % C::NV) * C::H_STRIDE) + offH;
                a.rho0 = reinterpret_cast<const float *>(smem + C::MU_OFF + (gs % C::NMU) * C::MU_STRIDE) + offM;
                a.rho1 = reinterpret_cast<const float *>(smem + C::MU_OFF + ((gs + 1) % C::NMU) * C::MU_STRIDE) + offM;
                a.v = reinterpret_cast<const F *>(st) + offC;
                a.sa = reinterpret_cast<const F *>(st + C::SA_OFF) + offH;
                a.n = n;
                vel_interior_t<F>(p, a, k, mi, mj, bnd);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        n += p.SJ;
        an += g.asj;
        col += p.NXM;
        mj++;
        lj++;
    }
}

}   // namespace swpc
