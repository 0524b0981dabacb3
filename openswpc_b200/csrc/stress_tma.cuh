// stress_tma.cuh -- fused interior stress sweep, version 2: TMA-staged, warp-specialised, mbarrier-pipelined.
//
// Why: the direct kernel (kernels.cuh) is latency-bound -- ~840 instructions per cell, 40 % of them 64-bit address
// arithmetic, and its memory-level parallelism is capped by registers.  Here one producer lane issues five
// `cp.async.bulk.tensor` (TMA) loads per j-plane into a shared-memory ring; sixteen consumer warps read their operands
// with immediate-offset LDS, run the SAME arithmetic body (stress_interior_t) and store S / R straight to HBM.
// Bytes in flight are bounded by shared memory (2-3 stages x 45 KB per SM), not by registers.
//
// Tile: TK = 32 cells along k (one warp lane each) x TI = 8 columns along i (one consumer warp each), marching along
// j.  Per step t (plane j = j0 + t) the producer loads, under ONE full-barrier:
//   S   box (TK, TI, 1, 6)        fields 3..8 of the field tensor      -> stage t % NS
//   R   box (TK, TI, 1, 6*NM)     memory variables                     -> stage t % NS
//   M   box (TK, TI, 1, 3|1)      lam, taup, taus (lam only if NM = 0) -> stage t % NS
//   V   box (TK+8, TI+4, 1, 3)    Vx, Vy, Vz of plane j+2, with halo   -> ring slot (t+4) % (NS+4)
//   mu  box (TK+4, TI+1, 1, 1)    mu of plane j+1                      -> ring slot (t+1) % (NS+1)
// (step 0 also brings V planes j0-2..j0+1 and mu plane j0).  A consumer warp arrives on the stage's empty-barrier when
// it has finished step t; the producer refills stage t % NS (and the V / mu slots whose last reader was step t-NS... t)
// only after that.  Only tiles that lie completely inside the interior kernel box are handled here; the absorber
// shell and the ragged edges stay with sweep_direct.
#pragma once

#include <cuda.h>

#include "kernels.cuh"

namespace swpc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// the same with an L2 eviction policy: data that is streamed once (S, R, medium centre boxes: 90 % of the traffic) is loaded evict-first
// so that it does not push the halo'd boxes (V, mu), which neighbouring tiles read again, out of L2 before they do
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_4d_hint(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(pol)
        : "memory");
}

struct TmaMaps {
    CUtensorMap S, V, R, M, Mu;
};
struct TmaGeom {
    int li0, lj0, lj1;   // first interior column handled (local i), j range (local, inclusive)
    int li1;             // last interior column handled: the last tile column may be partial (its idle columns are loaded and masked)
    int jl;              // planes per block
    int m_first;         // index of lam in the medium tensor's 4th dimension (lam, taup, taus are consecutive)
    int mu_index;        // index of mu in the medium tensor's 4th dimension
    int shift_last;      // 1: shift a partial last k-tile up (see stress_tma)
    int l2hint;          // 1: centre boxes evict-first; 2: + halo boxes evict-last
};

constexpr int align128(int x) { return (x + 127) / 128 * 128; }

#ifndef SWPC_TMA_NM0_TWO
#define SWPC_TMA_NM0_TWO 1
#endif
#ifndef SWPC_TMA_MAXREG
#define SWPC_TMA_MAXREG 80
#endif
#ifndef SWPC_TMA_NS
#define SWPC_TMA_NS 3
#endif
template <typename F, int NM>
struct TmaCfg {
#ifndef SWPC_TMA_TK
#define SWPC_TMA_TK 32
#endif
#ifndef SWPC_TMA_TI
#define SWPC_TMA_TI 8
#endif
    // elastic runs (NM = 0) move 128 B per cell instead of 280: the consumer warps of ONE resident block cannot keep up with
    // HBM, so that instantiation takes a 2-stage ring (100 KB) and a 56-register cap (17 warps x 1792 registers): two blocks per SM
    // (stress sweep 4.31 -> 3.89 ms at 512^3; the same change for float32 fields with NM = 3 measured 1 % slower, not kept)
    // float32 fields: a stage is 27 KB instead of 45 KB -- a 4-stage ring fits (164 KB) and covers more latency
#ifndef SWPC_TMA_NS_F32
#define SWPC_TMA_NS_F32 4
#endif
    static constexpr int TK = SWPC_TMA_TK, TI = SWPC_TMA_TI,
                         NS = (NM == 0 && sizeof(F) == 8 && SWPC_TMA_NM0_TWO) ? 2 : (sizeof(F) == 4 && NM > 0) ? SWPC_TMA_NS_F32 : SWPC_TMA_NS;
    static constexpr int MAXREG = (NM == 0 && sizeof(F) == 8 && SWPC_TMA_NM0_TWO) ? 56 : SWPC_TMA_MAXREG;
    static constexpr int NHW = (TK / 32) * TI;   // warps per half: one warp = 32 consecutive k of one column
    static constexpr int NCW = 2 * NHW;          // consumer warps: NHW for the normal, NHW for the shear components
    static constexpr int VHK = 4;   // k halo of the V box: 4 (not 2) so that the box start stays 16-byte aligned for float fields
    static constexpr int VK = TK + 2 * VHK, VI = TI + 4;
    static constexpr int NMED = (NM > 0) ? 3 : 1;
    static constexpr int S_BYTES = 6 * TK * TI * (int)sizeof(F);
    static constexpr int R_BYTES = 6 * NM * TK * TI * 4;
    static constexpr int M_BYTES = NMED * TK * TI * 4;
    static constexpr int R_OFF = align128(S_BYTES), M_OFF = R_OFF + align128(R_BYTES);
    static constexpr int STAGE = M_OFF + align128(M_BYTES);
    static constexpr int V_BYTES = 3 * VK * VI * (int)sizeof(F), V_STRIDE = align128(V_BYTES);
    static constexpr int MUK = TK + 4, MUI = TI + 1;
    static constexpr int MU_BYTES = MUK * MUI * 4, MU_STRIDE = align128(MU_BYTES);
    static constexpr int NV = NS + 4, NMU = NS + 1;
    static constexpr int V_OFF = NS * STAGE, MU_OFF = V_OFF + NV * V_STRIDE, BAR_OFF = MU_OFF + NMU * MU_STRIDE;
    static constexpr int SMEM = BAR_OFF + 128;
    static constexpr int STEP_TX = S_BYTES + R_BYTES + M_BYTES + V_BYTES + MU_BYTES;
    static constexpr int THREADS = (NCW + 1) * 32;
};

// operands in shared memory (tiles written by TMA), results to global
template <typename F, int NM>
struct AccTma {
    using C = TmaCfg<F, NM>;
    const KParams<F> &p;
    const F *v[5];          // V planes j-2 .. j+2, each [3][VI][VK]; pointers already offset to this thread's cell
    const float *mu0, *mu1; // mu planes j, j+1, each [MUI][MUK]; offset to this thread's cell
    const F *s;             // [6][TI][TK]
    const float *r;         // [6*NM][TI][TK]
    const float *m;         // [NMED][TI][TK]
    long long n;            // global linear index of the cell
    __device__ __forceinline__ AccTma(const KParams<F> &p_) : p(p_) {}
    template <int f, int dk, int di, int dj> __device__ __forceinline__ F V() const { return v[dj + 2][f * (C::VI * C::VK) + di * C::VK + dk]; }
    template <int dk, int di, int dj> __device__ __forceinline__ float mu() const { return (dj == 0 ? mu0 : mu1)[di * C::MUK + dk]; }
    __device__ __forceinline__ float lam() const { return m[0]; }
    __device__ __forceinline__ float taup() const { return m[1 * C::TI * C::TK]; }
    __device__ __forceinline__ float taus() const { return m[2 * C::TI * C::TK]; }
    __device__ __forceinline__ F *sptr(int c) const { return c == 0 ? p.Sxx : c == 1 ? p.Syy : c == 2 ? p.Szz : c == 3 ? p.Syz : c == 4 ? p.Sxz : p.Sxy; }
    // S box = field slots 3..8 = Sxx Szz Sxz Syy Syz Sxy (device slot order, see abi.cu); c = xx yy zz yz xz xy
    __device__ __forceinline__ F S(int c) const { return s[(c == 0 ? 0 : c == 1 ? 3 : c == 2 ? 1 : c == 3 ? 4 : c == 4 ? 2 : 5) * (C::TI * C::TK)]; }
    __device__ __forceinline__ void setS(int c, F val) const { sts_(sptr(c) + n, val); }
    __device__ __forceinline__ float R(int q) const { return r[q * (C::TI * C::TK)]; }
    __device__ __forceinline__ void setR(int q, float val) const { sts_(p.R + n + q * p.ncell, val); }
};

template <typename F, int NM>
#ifndef SWPC_TMA_MAXREG
#define SWPC_TMA_MAXREG 80
#endif
// register cap (instead of __launch_bounds__): 17 warps x 80 registers leave room for one sweep_direct block of the
// absorber shell on the same SM
__global__ void __maxnreg__((TmaCfg<F, NM>::MAXREG))
stress_tma(const __grid_constant__ KParams<F> p, const __grid_constant__ TmaMaps tm, const TmaGeom g) {
    using C = TmaCfg<F, NM>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::BAR_OFF);
    uint64_t *empty = full + C::NS;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // first k of the tile (1-based).  A last tile that would reach below the interior box is shifted up instead (to the first
    // k = 1 mod 4 -- the alignment class of every tile start, which TMA needs -- that still covers kend_k): the rows it then
    // shares with the tile above are masked here and come out of L2 (the neighbour block pulls them at the same time),
    // where the absorber rows below kend_k would have been fetched from HBM only to be thrown away.
    const int k0t = 1 + blockIdx.x * C::TK;
    const int k0 = (k0t + C::TK - 1 > p.k1_k && p.k1_k >= C::TK && g.shift_last) ? ((p.k1_k - C::TK + 3) / 4) * 4 + 1 : k0t;
    const int li0 = g.li0 + blockIdx.y * C::TI;            // first local i of the tile
    const int lj0 = g.lj0 + blockIdx.z * g.jl;             // first local j of the march
    const int nsteps = min(g.jl, g.lj1 - lj0 + 1);
    // tensor coordinates (elements): k -> k + KOFF - 1, i -> li + HALO, j -> lj + HALO
    const int ck = k0 + KOFF - 1, ci = li0 + HALO, cj = lj0 + HALO;

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
            const uint64_t pf = l2_policy_evict_first(), pl = l2_policy_evict_last();
            const bool hc = g.l2hint >= 1, hh = g.l2hint >= 2;
            auto ldc = [&](void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {   // streamed once
                if (hc) tma_load_4d_hint(dst, m, bar, c0, c1, c2, c3, pf);
                else tma_load_4d(dst, m, bar, c0, c1, c2, c3);
            };
            auto ldh = [&](void *dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1, int c2, int c3) {   // halo'd: read again by neighbours
                if (hh) tma_load_4d_hint(dst, m, bar, c0, c1, c2, c3, pl);
                else tma_load_4d(dst, m, bar, c0, c1, c2, c3);
            };
            for (int t = 0; t < nsteps; t++) {
                const int s = t % C::NS;
                if (t >= C::NS) mbar_wait(&empty[s], ((t / C::NS) - 1) & 1);
                unsigned char *st = smem + s * C::STAGE;
                uint32_t tx = C::STEP_TX;
                if (t == 0) tx += 4 * C::V_BYTES + C::MU_BYTES;
                mbar_expect_tx(&full[s], tx);
                if (t == 0) {
                    for (int q = 0; q < 4; q++)
                        ldh(smem + C::V_OFF + q * C::V_STRIDE, &tm.V, &full[s], ck - C::VHK, ci - 2, cj - 2 + q, 0);
                    ldh(smem + C::MU_OFF, &tm.Mu, &full[s], ck, ci, cj, g.mu_index);
                }
                ldc(st, &tm.S, &full[s], ck, ci, cj + t, 3);
                if (NM > 0) ldc(st + C::R_OFF, &tm.R, &full[s], ck, ci, cj + t, 0);
                ldc(st + C::M_OFF, &tm.M, &full[s], ck, ci, cj + t, g.m_first);
                ldh(smem + C::V_OFF + ((t + 4) % C::NV) * C::V_STRIDE, &tm.V, &full[s], ck - C::VHK, ci - 2, cj + t + 2, 0);
                ldh(smem + C::MU_OFF + ((t + 1) % C::NMU) * C::MU_STRIDE, &tm.Mu, &full[s], ck, ci, cj + t + 1, g.mu_index);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers: warp = column (x2: warps 0..TI-1
    // update the normal components, warps TI..2TI-1 the shear components of the same cells), lane = k
    const bool shear = warp >= C::NHW;
    const int wq = shear ? warp - C::NHW : warp;
    const int tk = (wq % (C::TK / 32)) * 32 + lane, ti = wq / (C::TK / 32);
    const int k = k0 + tk, li = li0 + ti, mi = li + HALO;
    // last k-tile: absorber rows / rows of the tile above idle; last i-tile: columns past the box idle
    const bool active = (k0 + tk) <= p.k1_k && (k0 + tk) >= k0t && li <= g.li1;
    AccTma<F, NM> a(p);
    const int voff = (ti + 2) * C::VK + (tk + C::VHK);
    const int muoff = ti * C::MUK + tk;
    const int coff = ti * C::TK + tk;
    long long col = (long long)mi + (long long)p.NXM * (lj0 + HALO);
    a.n = (long long)(k + KOFF - 1) + (long long)p.NZP * col;
    for (int t = 0; t < nsteps; t++) {
        const int s = t % C::NS;
        const int4 bnd = p.band[col];
        mbar_wait(&full[s], (t / C::NS) & 1);
        const unsigned char *st = smem + s * C::STAGE;
#pragma unroll
        for (int q = 0; q < 5; q++) a.v[q] = reinterpret_cast<const F *>(smem + C::V_OFF + ((t + q) % C::NV) * C::V_STRIDE) + voff;
        a.mu0 = reinterpret_cast<const float *>(smem + C::MU_OFF + (t % C::NMU) * C::MU_STRIDE) + muoff;
        a.mu1 = reinterpret_cast<const float *>(smem + C::MU_OFF + ((t + 1) % C::NMU) * C::MU_STRIDE) + muoff;
        a.s = reinterpret_cast<const F *>(st) + coff;
        a.r = reinterpret_cast<const float *>(st + C::R_OFF) + coff;
        a.m = reinterpret_cast<const float *>(st + C::M_OFF) + coff;
        if (active) {
            if (shear) stress_interior_t<F, NM, AccTma<F, NM>, false, true>(p, a, k, mi, lj0 + t + HALO, bnd);
            else stress_interior_t<F, NM, AccTma<F, NM>, true, false>(p, a, k, mi, lj0 + t + HALO, bnd);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        a.n += p.SJ;
        col += p.NXM;
    }
}

// ================================================================================================
// stress_tma, persistent form: ONE block per SM takes work items -- a tile (k0, li0) marching `nsteps` planes from lj0 --
// from a global ticket counter, in the order k-tile, i-tile, j-chunk (the order the hardware dispatches stress_tma's grid in,
// so that the blocks that run side by side are neighbours and their halos meet in L2).  Dynamic tickets, not a static split:
// the SMs do not stream at the same rate (a static split of this sweep measured 20 % slower than the plain grid).  What the
// persistent form buys over one block per chunk: the next item's loads are issued while the last planes of the previous one
// are still being computed, chunks can be long (the 4-plane V prologue is paid per item), and there is no block launch / exit
// in between.  Before an item's first ring loads the producer drains the pipeline (the rings have no room for two items'
// planes), so the V / mu rings are indexed by the block's running step count.  The producer tells the consumers which
// (item, plane) a stage holds through a small mailbox written before the stage's barrier is armed; item -1 ends the block.
struct TmaItem {
    int k0;        // first k of the boxes (1-based; 1 mod 4)
    int kown;      // first k this tile owns: rows k0 .. kown-1 of a shifted last tile belong to the tile above
    int li0, lj0;  // first local column / plane
    int nsteps;
};

template <typename F, int NM>
__global__ void __maxnreg__((TmaCfg<F, NM>::MAXREG))
stress_tma_p(const __grid_constant__ KParams<F> p, const __grid_constant__ TmaMaps tm, const TmaItem *__restrict__ items, int nitems,
             unsigned int *ticket, unsigned int ticket_base, const TmaGeom g) {
    using C = TmaCfg<F, NM>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::BAR_OFF);
    uint64_t *empty = full + C::NS;
    volatile int2 *meta = reinterpret_cast<volatile int2 *>(smem + C::BAR_OFF + 64);   // (item, plane) of each stage
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
        // ------------------------------------------------------------------ producer
        if (lane == 0) {
            int gs = 0;
            for (;;) {
                const unsigned int it = atomicAdd(ticket, 1u) - ticket_base;
                if (it >= (unsigned int)nitems) break;
                const TmaItem I = items[it];
                const int ck = I.k0 + KOFF - 1, ci = I.li0 + HALO, cj = I.lj0 + HALO;
                for (int t = 0; t < I.nsteps; t++, gs++) {
                    const int s = gs % C::NS;
                    if (gs >= C::NS) mbar_wait(&empty[s], ((gs / C::NS) - 1) & 1);
                    unsigned char *st = smem + s * C::STAGE;
                    uint32_t tx = C::STEP_TX;
                    if (t == 0) tx += 4 * C::V_BYTES + C::MU_BYTES;
                    meta[s].x = (int)it; meta[s].y = t;
                    mbar_expect_tx(&full[s], tx);
                    tma_load_4d(st, &tm.S, &full[s], ck, ci, cj + t, 3);
                    if (NM > 0) tma_load_4d(st + C::R_OFF, &tm.R, &full[s], ck, ci, cj + t, 0);
                    tma_load_4d(st + C::M_OFF, &tm.M, &full[s], ck, ci, cj + t, g.m_first);
                    if (t == 0) {
                        // drain: every earlier step released, so the prologue may overwrite any ring slot
                        for (int q = C::NS - 1; q >= 1; q--) {
                            const int gp = gs - q;
                            if (gp >= 0) mbar_wait(&empty[gp % C::NS], (gp / C::NS) & 1);
                        }
                        for (int q = 0; q < 4; q++)
                            tma_load_4d(smem + C::V_OFF + ((gs + q) % C::NV) * C::V_STRIDE, &tm.V, &full[s], ck - C::VHK, ci - 2, cj - 2 + q, 0);
                        tma_load_4d(smem + C::MU_OFF + (gs % C::NMU) * C::MU_STRIDE, &tm.Mu, &full[s], ck, ci, cj, g.mu_index);
                    }
                    tma_load_4d(smem + C::V_OFF + ((gs + 4) % C::NV) * C::V_STRIDE, &tm.V, &full[s], ck - C::VHK, ci - 2, cj + t + 2, 0);
                    tma_load_4d(smem + C::MU_OFF + ((gs + 1) % C::NMU) * C::MU_STRIDE, &tm.Mu, &full[s], ck, ci, cj + t + 1, g.mu_index);
                }
            }
            const int s = gs % C::NS;   // end of work: an empty stage that carries item -1
            if (gs >= C::NS) mbar_wait(&empty[s], ((gs / C::NS) - 1) & 1);
            meta[s].x = -1; meta[s].y = 0;
            mbar_arrive(&full[s]);
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers (as stress_tma)
    const bool shear = warp >= C::NHW;
    const int wq = shear ? warp - C::NHW : warp;
    const int tk = (wq % (C::TK / 32)) * 32 + lane, ti = wq / (C::TK / 32);
    AccTma<F, NM> a(p);
    const int voff = (ti + 2) * C::VK + (tk + C::VHK);
    const int muoff = ti * C::MUK + tk;
    const int coff = ti * C::TK + tk;
    int k = 0, mi = 0, mj = 0;
    bool active = false;
    long long col = 0;
    for (int gs = 0;; gs++) {
        const int s = gs % C::NS;
        mbar_wait(&full[s], (gs / C::NS) & 1);
        int it = 0, t = 0;
        if (lane == 0) { it = meta[s].x; t = meta[s].y; }   // the lane that arrives on the empty-barrier is the one that reads the mailbox
        it = __shfl_sync(0xffffffffu, it, 0);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (it < 0) break;
        if (t == 0) {
            const TmaItem I = items[it];
            k = I.k0 + tk;
            mi = I.li0 + ti + HALO;
            mj = I.lj0 + HALO;
            active = k <= p.k1_k && k >= I.kown && I.li0 + ti <= g.li1;
            col = (long long)mi + (long long)p.NXM * mj;
            a.n = (long long)(k + KOFF - 1) + (long long)p.NZP * col;
        }
        const int4 bnd = p.band[col];
        const unsigned char *st = smem + s * C::STAGE;
#pragma unroll
        for (int q = 0; q < 5; q++) a.v[q] = reinterpret_cast<const F *>(smem + C::V_OFF + ((gs + q) % C::NV) * C::V_STRIDE) + voff;
        a.mu0 = reinterpret_cast<const float *>(smem + C::MU_OFF + (gs % C::NMU) * C::MU_STRIDE) + muoff;
        a.mu1 = reinterpret_cast<const float *>(smem + C::MU_OFF + ((gs + 1) % C::NMU) * C::MU_STRIDE) + muoff;
        a.s = reinterpret_cast<const F *>(st) + coff;
        a.r = reinterpret_cast<const float *>(st + C::R_OFF) + coff;
        a.m = reinterpret_cast<const float *>(st + C::M_OFF) + coff;
        if (active) {
            if (shear) stress_interior_t<F, NM, AccTma<F, NM>, false, true>(p, a, k, mi, mj, bnd);
            else stress_interior_t<F, NM, AccTma<F, NM>, true, false>(p, a, k, mi, mj, bnd);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        a.n += p.SJ;
        col += p.NXM;
        mj++;
    }
}

// ================================================================================================
// velocity sweep, same scheme.  Per step t (plane j = j0 + t), one full-barrier:
//   V    box (TK, TI, 1, 3)        Vx Vy Vz (field slots 0..2), read-modify-write   -> stage t % NS
//   SA   box (TK+8, TI+4, 1, 3)    Sxx Szz Sxz (slots 3..5): only plane j is needed -> stage t % NS
//   SB   box (TK+8, TI+4, 1, 3)    Syy Syz Sxy (slots 6..8) of plane j+2            -> ring slot (t+4) % (NS+4)
//   rho  box (TK+4, TI+1, 1, 1)    rho of plane j+1                                 -> ring slot (t+1) % (NS+1)
struct TmaMapsVel {
    CUtensorMap Vv, Sh, Rho;   // Sh serves SA and SB (same box, different 4th coordinate)
};

template <typename F>
struct TmaCfgVel {
    static constexpr int TK = SWPC_TMA_TK, TI = SWPC_TMA_TI, NS = 3;
    static constexpr int NCW = (TK / 32) * TI;
    static constexpr int HK = 4;
    static constexpr int SK = TK + 2 * HK, SI_ = TI + 4;
    static constexpr int V_BYTES = 3 * TK * TI * (int)sizeof(F);
    static constexpr int S_BYTES = 3 * SK * SI_ * (int)sizeof(F), S_STRIDE = align128(S_BYTES);
    static constexpr int SA_OFF = align128(V_BYTES);
    static constexpr int STAGE = SA_OFF + S_STRIDE;
    static constexpr int RK = TK + 4, RI = TI + 1;
    static constexpr int RHO_BYTES = RK * RI * 4, RHO_STRIDE = align128(RHO_BYTES);
    static constexpr int NB = NS + 4, NRHO = NS + 1;
    static constexpr int SB_OFF = NS * STAGE, RHO_OFF = SB_OFF + NB * S_STRIDE, BAR_OFF = RHO_OFF + NRHO * RHO_STRIDE;
    static constexpr int SMEM = BAR_OFF + 128;
    static constexpr int STEP_TX = V_BYTES + 2 * S_BYTES + RHO_BYTES;
    static constexpr int THREADS = (NCW + 1) * 32;
};

template <typename F>
struct AccVelTma {
    using C = TmaCfgVel<F>;
    const KParams<F> &p;
    const F *sa;            // [3][SI_][SK] plane j: Sxx Szz Sxz, offset to this thread's cell
    const F *sb[5];         // planes j-2..j+2: [3][SI_][SK] Syy Syz Sxy
    const float *rho0, *rho1;
    const F *v;             // [3][TI][TK]
    long long n;
    __device__ __forceinline__ AccVelTma(const KParams<F> &p_) : p(p_) {}
    template <int c, int dk, int di, int dj> __device__ __forceinline__ F S() const {
        constexpr int PL = C::SI_ * C::SK;
        if (c == 0 || c == 2 || c == 4) {   // xx zz xz live in the per-step tile (dj is always 0 for them)
            return sa[(c == 0 ? 0 : c == 2 ? 1 : 2) * PL + di * C::SK + dk];
        } else {                              // yy yz xy in the j ring
            return sb[dj + 2][(c == 1 ? 0 : c == 3 ? 1 : 2) * PL + di * C::SK + dk];
        }
    }
    template <int dk, int di, int dj> __device__ __forceinline__ float rho() const { return (dj == 0 ? rho0 : rho1)[di * C::RK + dk]; }
    __device__ __forceinline__ F V(int f) const { return v[f * (C::TI * C::TK)]; }
    __device__ __forceinline__ void setV(int f, F val) const { (f == 0 ? p.Vx : f == 1 ? p.Vy : p.Vz)[n] = val; }
};

template <typename F>
__global__ void __maxnreg__(SWPC_TMA_MAXREG)
vel_tma(const __grid_constant__ KParams<F> p, const __grid_constant__ TmaMapsVel tm, const TmaGeom g) {
    using C = TmaCfgVel<F>;
    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::BAR_OFF);
    uint64_t *empty = full + C::NS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k0 = 1 + blockIdx.x * C::TK;
    const int li0 = g.li0 + blockIdx.y * C::TI;
    const int lj0 = g.lj0 + blockIdx.z * g.jl;
    const int nsteps = min(g.jl, g.lj1 - lj0 + 1);
    const int ck = k0 + KOFF - 1, ci = li0 + HALO, cj = lj0 + HALO;

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NS; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], C::NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == C::NCW) {
        if (lane == 0) {
            for (int t = 0; t < nsteps; t++) {
                const int s = t % C::NS;
                if (t >= C::NS) mbar_wait(&empty[s], ((t / C::NS) - 1) & 1);
                unsigned char *st = smem + s * C::STAGE;
                uint32_t tx = C::STEP_TX;
                if (t == 0) tx += 4 * C::S_BYTES + C::RHO_BYTES;
                mbar_expect_tx(&full[s], tx);
                if (t == 0) {
                    for (int q = 0; q < 4; q++)
                        tma_load_4d(smem + C::SB_OFF + q * C::S_STRIDE, &tm.Sh, &full[s], ck - C::HK, ci - 2, cj - 2 + q, 6);
                    tma_load_4d(smem + C::RHO_OFF, &tm.Rho, &full[s], ck, ci, cj, 0);
                }
                tma_load_4d(st, &tm.Vv, &full[s], ck, ci, cj + t, 0);
                tma_load_4d(st + C::SA_OFF, &tm.Sh, &full[s], ck - C::HK, ci - 2, cj + t, 3);
                tma_load_4d(smem + C::SB_OFF + ((t + 4) % C::NB) * C::S_STRIDE, &tm.Sh, &full[s], ck - C::HK, ci - 2, cj + t + 2, 6);
                tma_load_4d(smem + C::RHO_OFF + ((t + 1) % C::NRHO) * C::RHO_STRIDE, &tm.Rho, &full[s], ck, ci, cj + t + 1, 0);
            }
        }
        return;
    }

    const int tk = (warp % (C::TK / 32)) * 32 + lane, ti = warp / (C::TK / 32);
    const int k = k0 + tk, li = li0 + ti, mi = li + HALO;
    const bool active = (k0 + tk) <= p.k1_k && li <= g.li1;
    AccVelTma<F> a(p);
    const int soff = (ti + 2) * C::SK + (tk + C::HK);
    const int roff = ti * C::RK + tk;
    const int coff = ti * C::TK + tk;
    long long col = (long long)mi + (long long)p.NXM * (lj0 + HALO);
    a.n = (long long)(k + KOFF - 1) + (long long)p.NZP * col;
    for (int t = 0; t < nsteps; t++) {
        const int s = t % C::NS;
        const int4 bnd = p.band[col];
        mbar_wait(&full[s], (t / C::NS) & 1);
        const unsigned char *st = smem + s * C::STAGE;
#pragma unroll
        for (int q = 0; q < 5; q++) a.sb[q] = reinterpret_cast<const F *>(smem + C::SB_OFF + ((t + q) % C::NB) * C::S_STRIDE) + soff;
        a.rho0 = reinterpret_cast<const float *>(smem + C::RHO_OFF + (t % C::NRHO) * C::RHO_STRIDE) + roff;
        a.rho1 = reinterpret_cast<const float *>(smem + C::RHO_OFF + ((t + 1) % C::NRHO) * C::RHO_STRIDE) + roff;
        a.v = reinterpret_cast<const F *>(st) + coff;
        a.sa = reinterpret_cast<const F *>(st + C::SA_OFF) + soff;
        if (active) vel_interior_t<F>(p, a, k, mi, lj0 + t + HALO, bnd);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        a.n += p.SJ;
        col += p.NXM;
    }
}

}   // namespace swpc
