// pml_tma.cuh -- the absorber shell (ADE-CFS PML cells, m_absorb_p.f90:246-317 / :429-531), TMA-staged.
//
// Why: sweep_direct runs a PML cell as ~40 dependent global loads at 36 % occupancy (long-scoreboard bound: 61-63 % of the HBM
// peak on the shell, 47 % on the bottom slab whose 20-cell columns start mid-line).  Here the operands of a tile are staged in
// shared memory by TMA, several planes ahead, exactly as stress_tma does for the interior -- bytes in flight are bounded by
// shared memory, not by registers -- and ONE persistent block per SM walks a list of work items without ever draining its
// pipeline between them.
//
// Work item = a tile of BK rows (k) x TI columns (i) marching `nsteps` planes along j, all of its active cells PML cells:
//   walls   BK = 32 (k = 1 + 32 q ...), the 4 slabs of absorber columns (kbeg_a = 1) around the interior kernel box
//   bottom  the rows kend_k + 1 .. nz under the interior columns, as ONE box of BK rows that starts on a 16-byte boundary
// Threads are numbered flat over the active (row, column) cells of the tile, so every lane has a cell whatever the row count.
// Per step (plane j) the producer lane loads under ONE full-barrier
//   stress sweep:  S box (BK, TI, 1, 6)   aux box (BK, TI, 1, 9: the V-gradient ADE variables)   lam box (BK, TI)   -> stage
//                  V box with halo (BK+8, TI+2, 1, 3) of plane j+1 -> 3-plane ring      mu box (BK+4, TI+1) of plane j+1 -> 2-plane ring
//   velocity sweep: V box (BK, TI, 1, 3)  aux box (.., 9: the S-gradient ADE variables)  Sxx Szz Sxz halo box of plane j -> stage
//                  Syy Syz Sxy halo box of plane j+1 -> 3-plane ring                     rho box of plane j+1 -> 2-plane ring
// (an item's first step also brings planes j-1 and j of the ring arrays).  Ring slots are numbered by a running counter that
// jumps at every item start, so the loads of a new item never land on slots the previous item's last steps still read:
//   V / S ring  slot(g, m, q) = (g + 2 m + q) mod NV, q = 0..2     medium ring  slot(g, m, q) = (g + m + q) mod NM_, q = 0..1
// with g = steps this block has done, m = items it has started.  Items are at least NS steps long (host), which bounds the
// spread of live slots by NS + 3 (NS + 1); NV = NS + 4, NM_ = NS + 2.
// The arithmetic is stress_pml_t / vel_pml_t of kernels.cuh -- the body sweep_direct runs -- so the results are bit-identical.
#pragma once

#include "stress_tma.cuh"

namespace swpc {

struct PmlItem {
    int k0;              // first row of the boxes (1-based k); index k0 + KOFF - 1 is a multiple of 4 (16-byte aligned for float)
    int r0, r1;          // active rows of the box, 0-based inclusive: cells k0 + r0 .. k0 + r1
    int li0, ncol;       // first local column, active columns (<= TI)
    int lj0, nsteps;     // first local plane, planes to march
    int amap;            // aux tensor map of the item's region
    int ak, ai, aj;      // aux tensor coordinates of (k0, li0, lj0)
    int asi;             // aux elements between neighbouring columns of the region
    long long asj;       // aux elements between neighbouring planes of the region
    long long aux0;      // linear aux index of (k0, li0, lj0)
};

constexpr int PML_NMAP = 4;
struct PmlMaps {
    CUtensorMap C;       // centre box over the field tensor: (BK, TI, 1, 6) stress / (BK, TI, 1, 3) velocity
    CUtensorMap H;       // halo box over the field tensor: (BK + 8, TI + 2, 1, 3)
    CUtensorMap M1;      // medium centre box (BK, TI, 1, 1): lam (stress only)
    CUtensorMap Mh;      // medium halo box (BK + 4, TI + 1, 1, 1): mu (stress) / rho (velocity)
    CUtensorMap aux[PML_NMAP];   // (BK, TI, 1, 9) over the region's part of the aux arrays
};

#ifndef SWPC_PML_NS
#define SWPC_PML_NS 4
#endif

template <typename F, bool STRESS, int TI, int BK, int NS_>
struct PmlCfgN {
    static constexpr int NS = NS_;
    static constexpr int NCELL = TI * BK;
    static constexpr int NCT = (NCELL + 31) / 32 * 32;   // consumer threads: one cell each
    static constexpr int NCW = NCT / 32;
    static constexpr int THREADS = NCT + 32;
    static constexpr int HK = 4;                          // k halo of the halo box (keeps its start 16-byte aligned for float fields)
    static constexpr int BKH = BK + 2 * HK, TIH = TI + 2;
    static constexpr int BKM = BK + 4, TIM = TI + 1;
    static constexpr int NC = STRESS ? 6 : 3;
    static constexpr int C_BYTES = NC * NCELL * (int)sizeof(F);
    static constexpr int A_BYTES = 9 * NCELL * 4;
    static constexpr int L_BYTES = STRESS ? NCELL * 4 : 0;
    static constexpr int H_BYTES = 3 * BKH * TIH * (int)sizeof(F), H_STRIDE = align128(H_BYTES);
    static constexpr int M_BYTES = BKM * TIM * 4, M_STRIDE = align128(M_BYTES);
    static constexpr int A_OFF = align128(C_BYTES);
    static constexpr int L_OFF = A_OFF + align128(A_BYTES);
    static constexpr int SA_OFF = L_OFF + align128(L_BYTES);            // velocity sweep: in-plane S halo box
    static constexpr int STAGE = SA_OFF + (STRESS ? 0 : H_STRIDE);
    static constexpr int NV = NS + 4, NM_ = NS + 2;
    static constexpr int H_OFF = NS * STAGE, M_OFF = H_OFF + NV * H_STRIDE, BAR_OFF = M_OFF + NM_ * M_STRIDE;
    static constexpr int SMEM = BAR_OFF + 128;
    static constexpr int STEP_TX = C_BYTES + A_BYTES + L_BYTES + (STRESS ? 0 : H_BYTES) + H_BYTES + M_BYTES;
    static constexpr int FIRST_TX = 2 * H_BYTES + M_BYTES;
};

// the deepest pipeline (<= SWPC_PML_NS stages) that fits in 227 KB of shared memory
template <typename F, bool STRESS, int TI, int BK>
using PmlCfg = PmlCfgN<F, STRESS, TI, BK, (PmlCfgN<F, STRESS, TI, BK, SWPC_PML_NS>::SMEM <= 227 * 1024 ? SWPC_PML_NS : SWPC_PML_NS - 1)>;

// operands in shared memory, results to global memory
template <typename F, bool STRESS, int TI, int BK>
struct AccPmlTma {
    using C = PmlCfg<F, STRESS, TI, BK>;
    const KParams<F> &p;
    const F *h[3];          // ring planes j-1, j, j+1 of the halo'd fields [3][TIH][BKH], offset to this thread's cell
    const F *sa;            // velocity sweep: Sxx Szz Sxz of plane j [3][TIH][BKH]
    const float *m0, *m1;   // medium ring planes j, j+1 [TIM][BKM]
    const F *c;             // centre box [NC][TI][BK]
    const float *ax;        // aux box [9][TI][BK]
    const float *lm;        // lam box [TI][BK]
    long long n, an;        // global linear indices of the cell in the field arrays / in the aux arrays
    __device__ __forceinline__ AccPmlTma(const KParams<F> &p_) : p(p_) {}
    static constexpr int PH = C::TIH * C::BKH;
    // ---- stress sweep
    template <int f, int dk, int di, int dj> __device__ __forceinline__ F V() const { return h[dj + 1][f * PH + di * C::BKH + dk]; }
    template <int dk, int di, int dj> __device__ __forceinline__ float mu() const { return (dj == 0 ? m0 : m1)[di * C::BKM + dk]; }
    __device__ __forceinline__ float lam() const { return lm[0]; }
    __device__ __forceinline__ F *sptr(int q) const { return q == 0 ? p.Sxx : q == 1 ? p.Syy : q == 2 ? p.Szz : q == 3 ? p.Syz : q == 4 ? p.Sxz : p.Sxy; }
    // S box = field slots 3..8 = Sxx Szz Sxz Syy Syz Sxy; q = xx yy zz yz xz xy
    __device__ __forceinline__ F S(int q) const { return c[(q == 0 ? 0 : q == 1 ? 3 : q == 2 ? 1 : q == 3 ? 4 : q == 4 ? 2 : 5) * C::NCELL]; }
    __device__ __forceinline__ void setS(int q, F v) const { sts_(sptr(q) + n, v); }
    // aux box holds 9 consecutive arrays: q - 0 (stress sweep: axVx .. azVz) or q - 9 (velocity sweep: axSxx .. azSzz)
    __device__ __forceinline__ float aux(int q) const { return ax[(q - (STRESS ? 0 : 9)) * C::NCELL]; }
    __device__ __forceinline__ void setAux(int q, float v) const { sts_(p.aux + an + q * p.naux, v); }
    // ---- velocity sweep: xx zz xz of plane j live in the per-step box, yy yz xy in the ring
    template <int q, int dk, int di, int dj> __device__ __forceinline__ F Sn() const {
        if (q == 0 || q == 2 || q == 4) return sa[(q == 0 ? 0 : q == 2 ? 1 : 2) * PH + di * C::BKH + dk];
        return h[dj + 1][(q == 1 ? 0 : q == 3 ? 1 : 2) * PH + di * C::BKH + dk];
    }
    template <int dk, int di, int dj> __device__ __forceinline__ float rho() const { return (dj == 0 ? m0 : m1)[di * C::BKM + dk]; }
    __device__ __forceinline__ F Vc(int f) const { return c[f * C::NCELL]; }
    __device__ __forceinline__ void setV(int f, F v) const { sts_((f == 0 ? p.Vx : f == 1 ? p.Vy : p.Vz) + n, v); }
};

// first array of the boxes in the 4th tensor dimension (device slot order Vx Vy Vz | Sxx Szz Sxz | Syy Syz Sxy; medium rho mu lam taup taus)
struct PmlGeom {
    int nitems;
    int c_first, h_first, sa_first;   // stress: 3, 0, -;  velocity: 0, 6, 3
    int m1_index, mh_index;           // stress: lam = 2, mu = 1;  velocity: -, rho = 0
    int a_first;                      // stress: 0, velocity: 9
};

template <typename F, bool STRESS, int TI, int BK>
__global__ void __launch_bounds__((PmlCfg<F, STRESS, TI, BK>::THREADS), 1)
pml_tma(const __grid_constant__ KParams<F> p, const __grid_constant__ PmlMaps tm, const PmlItem *__restrict__ items, unsigned int *ticket,
        unsigned int ticket_base, const PmlGeom g) {
    using C = PmlCfg<F, STRESS, TI, BK>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::BAR_OFF);
    uint64_t *empty = full + C::NS;
    volatile int2 *meta = reinterpret_cast<volatile int2 *>(smem + C::BAR_OFF + 64);   // (item, plane) of each stage; item -1 ends the block
    static_assert(2 * C::NS * 8 <= 64 && C::NS * 8 <= 64, "barriers and mailbox share 128 bytes");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NS; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], C::NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == C::NCW) {
        // ------------------------------------------------------------------ producer: items from the global ticket counter (the SMs
        // do not stream at the same rate: a static split would wait for the slowest)
        if (lane == 0) {
            int gs = 0, m = 0;
            for (;; m++) {
                const unsigned int it = atomicAdd(ticket, 1u) - ticket_base;
                if (it >= (unsigned int)g.nitems) break;
                const PmlItem I = items[it];
                const int ck = I.k0 + KOFF - 1, ci = I.li0 + HALO, cj = I.lj0 + HALO;
                for (int t = 0; t < I.nsteps; t++, gs++) {
                    const int s = gs % C::NS;
                    if (gs >= C::NS) mbar_wait(&empty[s], ((gs / C::NS) - 1) & 1);
                    unsigned char *st = smem + s * C::STAGE;
                    meta[s].x = (int)it; meta[s].y = t;
                    mbar_expect_tx(&full[s], C::STEP_TX + (t == 0 ? C::FIRST_TX : 0));
                    const int v0 = gs + 2 * m, q0 = gs + m;
                    if (t == 0) {
                        tma_load_4d(smem + C::H_OFF + (v0 % C::NV) * C::H_STRIDE, &tm.H, &full[s], ck - C::HK, ci - 1, cj - 1, g.h_first);
                        tma_load_4d(smem + C::H_OFF + ((v0 + 1) % C::NV) * C::H_STRIDE, &tm.H, &full[s], ck - C::HK, ci - 1, cj, g.h_first);
                        tma_load_4d(smem + C::M_OFF + (q0 % C::NM_) * C::M_STRIDE, &tm.Mh, &full[s], ck, ci, cj, g.mh_index);
                    }
                    tma_load_4d(st, &tm.C, &full[s], ck, ci, cj + t, g.c_first);
                    tma_load_4d(st + C::A_OFF, &tm.aux[I.amap], &full[s], I.ak, I.ai, I.aj + t, g.a_first);
                    if (STRESS) tma_load_4d(st + C::L_OFF, &tm.M1, &full[s], ck, ci, cj + t, g.m1_index);
                    else tma_load_4d(st + C::SA_OFF, &tm.H, &full[s], ck - C::HK, ci - 1, cj + t, g.sa_first);
                    tma_load_4d(smem + C::H_OFF + ((v0 + 2) % C::NV) * C::H_STRIDE, &tm.H, &full[s], ck - C::HK, ci - 1, cj + t + 1, g.h_first);
                    tma_load_4d(smem + C::M_OFF + ((q0 + 1) % C::NM_) * C::M_STRIDE, &tm.Mh, &full[s], ck, ci, cj + t + 1, g.mh_index);
                }
            }
            const int s = gs % C::NS;
            if (gs >= C::NS) mbar_wait(&empty[s], ((gs / C::NS) - 1) & 1);
            meta[s].x = -1; meta[s].y = 0;
            mbar_arrive(&full[s]);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers: thread = one active cell of the tile
    AccPmlTma<F, STRESS, TI, BK> a(p);
    const int tid = threadIdx.x;
    int m = -1, lj = 0, offC = 0, offH = 0, offM = 0;
    bool active = false;
    long long asj = 0;
    float4 gxc = make_float4(0, 0, 0, 0), gxe = gxc, gzc = gxc, gze = gxc;
    for (int gs = 0;; gs++) {
        const int s = gs % C::NS;
        // the y profile of this plane comes from global memory: fetched BEFORE the wait (lj already points at this step unless a new
        // item starts), so that its L2 latency hides behind the barrier
        float4 gyc = make_float4(0, 0, 0, 0), gye = gyc;
        if (active) { const int q = min(lj, p.nyp - 1); gyc = ldro(p.gyc + q); gye = ldro(p.gye + q); }   // (lj = one past the item's last plane after its last step)
        mbar_wait(&full[s], (gs / C::NS) & 1);
        int it = 0, t = 0;
        if (lane == 0) { it = meta[s].x; t = meta[s].y; }   // the lane that arrives on the empty-barrier is the one that reads the mailbox
        it = __shfl_sync(0xffffffffu, it, 0);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (it < 0) break;
        if (t == 0) {
            const PmlItem I = items[it];
            m++;
            const int nrow = I.r1 - I.r0 + 1;
            const int ai = tid / nrow, kr = I.r0 + (tid - ai * nrow);
            active = ai < I.ncol;
            const int k = I.k0 + kr, li = I.li0 + ai;
            offC = ai * BK + kr; offH = (ai + 1) * C::BKH + (kr + C::HK); offM = ai * C::BKM + kr;
            if (active) { gxc = ldro(p.gxc + li); gxe = ldro(p.gxe + li); gzc = ldro(p.gzc + (k - 1)); gze = ldro(p.gze + (k - 1)); }
            a.n = (long long)(k + KOFF - 1) + (long long)p.NZP * ((long long)(li + HALO) + (long long)p.NXM * (I.lj0 + HALO));
            a.an = I.aux0 + kr + (long long)ai * I.asi;
            asj = I.asj;
            lj = I.lj0;
            if (active) { gyc = ldro(p.gyc + lj); gye = ldro(p.gye + lj); }
        }
        const int v0 = gs + 2 * m, q0 = gs + m;
        const unsigned char *st = smem + s * C::STAGE;
#pragma unroll
        for (int q = 0; q < 3; q++) a.h[q] = reinterpret_cast<const F *>(smem + C::H_OFF + ((v0 + q) % C::NV) * C::H_STRIDE) + offH;
        a.m0 = reinterpret_cast<const float *>(smem + C::M_OFF + (q0 % C::NM_) * C::M_STRIDE) + offM;
        a.m1 = reinterpret_cast<const float *>(smem + C::M_OFF + ((q0 + 1) % C::NM_) * C::M_STRIDE) + offM;
        a.c = reinterpret_cast<const F *>(st) + offC;
        a.ax = reinterpret_cast<const float *>(st + C::A_OFF) + offC;
        a.lm = reinterpret_cast<const float *>(st + C::L_OFF) + offC;
        a.sa = reinterpret_cast<const F *>(st + C::SA_OFF) + offH;
        if (active) {
            if (STRESS) stress_pml_t<F>(p, a, gxc, gxe, gyc, gye, gzc, gze);
            else vel_pml_t<F>(p, a, gxc, gxe, gyc, gye, gzc, gze);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        a.n += p.SJ;
        a.an += asj;
        lj++;
    }
}

}   // namespace swpc
