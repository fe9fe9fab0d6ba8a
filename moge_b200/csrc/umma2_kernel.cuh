// 2-CTA variant of the encoder GEMM: a thread-block cluster of two CTAs (one SM pair) computes a 256 x 256 output tile
// with `tcgen05.mma.cta_group::2` (UMMA M = 256).  Each CTA stages its own 128 rows of A and its own HALF of the B tile
// (128 of the 256 output columns), so the B traffic L2 -> SM per flop is halved and the smem ring holds 6 stages instead
// of 4; the accumulator halves live in each CTA's own TMEM and each CTA runs the usual epilogue on its 128 rows.
//
// Protocol (barriers sit at identical smem offsets in both CTAs; "leader" = cluster rank 0):
//   full[s]   leader only.  Both CTAs' TMA loads (cp.async.bulk.tensor ... cta_group::2) complete their bytes on it.
//   empty[s]  per CTA.  The leader's tcgen05.commit multicasts the arrive to both CTAs when the MMAs that read stage s retire.
//   tfull[a]  per CTA (multicast commit): accumulator stage a is complete.
//   tempty[a] leader only, count = epilogue warps of BOTH CTAs (the follower arrives remotely).
#pragma once
#include "umma_kernel.cuh"

namespace mg {

constexpr uint32_t kPeerMask = 0xFEFFFFFFu;     // clears the CTA-rank bit of a shared::cluster address -> the pair's leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {        // arrive on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
                 : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// RS > 0: the EPI_RESID residual is staged through shared memory by TMA (RS 4 KB chunk buffers per epilogue warp, 32 KB per
// buffer set = one ring stage each): RS = 2 for the short reductions (proj), RS = 1 where the main loop needs the ring (fc2).
template <int RS> struct Umma2CfgT {
    static constexpr int BN = 256;                          // output columns of the pair tile; each CTA stages 128 of them
    static constexpr int kStageBytes = TILE_M * 128 + 128 * 128;   // A: 128 rows, B: 128 rows (this CTA's half), 64 k each
    static constexpr int kStages = 6 - RS;                  // RS = number of 4 KB residual chunk buffers per epilogue warp (0, 1 or 2)
    static constexpr int kEpiWarps = 8;
    static constexpr int kThreads = 64 + 32 * kEpiWarps;
    static constexpr int kTmemCols = 512;
    static constexpr int kScratchBytes = kEpiWarps * 4096;
    static constexpr int kResidBytes = kEpiWarps * RS * 4096;
    static constexpr int kBarBytes = 1024;                  // full/empty rings, accumulator barriers, residual barriers, TMEM slot (keeps what follows 1024-byte aligned)
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 + kBarBytes + kScratchBytes + kResidBytes;
    static_assert(kSmemBytes <= kMaxDynSmem, "umma2_kernel: stage ring + scratch exceed the shared memory of one CTA");
    static_assert((2 * kStages + 4 + 2 * kEpiWarps) * 8 + 4 <= kBarBytes, "umma2_kernel: barrier block overflows");
};
using Umma2Cfg = Umma2CfgT<0>;

template <int EPI, bool BF16, int RS = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Umma2Cfg::kThreads, 1)
umma2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const __grid_constant__ CUtensorMap mapR,
             const UmmaParams p) {
    static_assert(RS == 0 || EPI == EPI_RESID, "the staged residual belongs to EPI_RESID");
    pdl_launch_dependents();      // (the wait sits after the barrier / TMEM set-up below: that prologue overlaps the previous kernel's tail)
    using Cfg = Umma2CfgT<RS>;
    constexpr int S = Cfg::kStages;
    constexpr int BN = Cfg::BN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
    uint64_t* empty = full + S;
    uint64_t* tfull = empty + S;
    uint64_t* tempty = tfull + 2;
    uint64_t* rfull = tempty + 2;                              // RS: [epilogue warp][2 buffers]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull + 2 * Cfg::kEpiWarps);
    float* scratch_base = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + Cfg::kBarBytes);
    uint8_t* resid_base = smem + S * Cfg::kStageBytes + Cfg::kBarBytes + Cfg::kScratchBytes;     // 1024-byte aligned (all terms are)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int m_pairs = (p.num_m_tiles + 1) / 2;                 // 256-row tiles
    const int total = m_pairs * p.num_n_tiles;
    const int kb_total = p.kb_main;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * Cfg::kEpiWarps); }
        if (RS) { tma_prefetch_desc(&mapR); for (int i = 0; i < 2 * Cfg::kEpiWarps; ++i) mbar_init(&rfull[i], 1); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    cluster_sync_all();                                           // barriers of both CTAs initialised before any remote signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ================================================================== TMA producer (both CTAs; converged warp, elected lane issues)
        {
            int s = 0; uint32_t ph = 0;
            for (int t = pair; t < total; t += npairs) {
                const int mp = t / p.num_n_tiles, nt = t % p.num_n_tiles;
                const int m0 = (2 * mp + static_cast<int>(rank)) * TILE_M;
                const int n0 = nt * BN + static_cast<int>(rank) * 128;
                for (int i = 0; i < kb_total; ++i) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* sa = smem + s * Cfg::kStageBytes;
                    uint8_t* sb = sa + TILE_M * 128;
                    const uint32_t lbar = smem_u32(&full[s]) & kPeerMask;
                    if (elect_one()) {
                        if (leader) mbar_arrive_expect_tx(&full[s], 2 * Cfg::kStageBytes);      // bytes of BOTH CTAs
                        tma_load_2d_2sm(sa, &mapA, lbar, i * TILE_K, m0);
                        tma_load_2d_2sm(sb, &mapB, lbar, i * TILE_K, n0);
                    }
                    __syncwarp();
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer (leader CTA only)
        if (leader) {
            constexpr uint32_t idesc = make_idesc(256, BN, BF16 ? 1u : 0u);
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int t = pair; t < total; t += npairs, ++it) {
                const int acc = it & 1;
                mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int i = 0; i < kb_total; ++i) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * Cfg::kStageBytes);
                    const uint64_t adesc = make_sdesc_sw128(sa);
                    const uint64_t bdesc = make_sdesc_sw128(sa + TILE_M * 128);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < TILE_K / 16; ++k) umma_f16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
                        umma_commit_2sm(&empty[s]);
                        if (i == kb_total - 1) umma_commit_2sm(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        // ================================================================== epilogue (both CTAs, own 128 rows)
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int col_begin = (ew >> 2) * (BN / 2);
        float4* scr = reinterpret_cast<float4*>(scratch_base + ew * 1024);
        int it = 0;
        LnRows lnr{}, lnn{};
        const bool ln_on = (EPI == EPI_STORE16 || EPI == EPI_GELU16) && p.ln_rstd != nullptr;
        if (ln_on && pair < total) ln_load(p, 2 * (pair / p.num_n_tiles) + static_cast<int>(rank), quarter, lane, lnn);
        // RS: the fp32 residual of this warp's 32 rows x 128 columns arrives in 32 x 32 chunks by TMA, two chunks ahead of their use;
        // the first two chunks of a tile are requested while the previous tile is still being drained, i.e. they land under the
        // MMA main loop instead of stalling the epilogue (proj / fc2 were bound by the latency of these reads).
        ResidPipe rp{};
        if (RS) {
            rp.buf = resid_base + ew * (RS * 4096); rp.bars = rfull + 2 * ew; rp.map = &mapR; rp.rc = 0; rp.nbuf = RS;
            if (pair < total) {
                const int row0 = (2 * (pair / p.num_n_tiles) + static_cast<int>(rank)) * TILE_M + quarter * 32;
                const int col0 = (pair % p.num_n_tiles) * BN + col_begin;
                if (elect_one()) {
                    for (int b = 0; b < RS; ++b) {
                        mbar_arrive_expect_tx(&rp.bars[b], 4096);
                        tma_load_2d(rp.buf + b * 4096, &mapR, &rp.bars[b], col0 + 32 * b, row0);
                    }
                }
                __syncwarp();
            }
        }
        for (int t = pair; t < total; t += npairs, ++it) {
            const int mp = t / p.num_n_tiles, nt = t % p.num_n_tiles;
            const int mt = 2 * mp + static_cast<int>(rank);
            const int acc = it & 1;
            if (ln_on) {        // rstd of this tile's rows (loaded one iteration ago); then issue the next tile's loads
                lnr = lnn;
                if (t + npairs < total) ln_load(p, 2 * ((t + npairs) / p.num_n_tiles) + static_cast<int>(rank), quarter, lane, lnn);
            }
            if (RS) {
                const int tn = t + npairs;
                rp.next_row0 = (tn < total) ? (2 * (tn / p.num_n_tiles) + static_cast<int>(rank)) * TILE_M + quarter * 32 : -1;
                rp.next_col0 = (tn < total) ? (tn % p.num_n_tiles) * BN + col_begin : 0;
            }
            mbar_wait(&tfull[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + col_begin;
            epilogue_tile<BN, BN / 2, AMODE_ROWS, EPI, BF16>(p, mt, nt, t_addr, scr, quarter, lane, col_begin, ln_on, lnr, RS ? &rp : nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&tempty[acc]);
                else mbar_arrive_remote(smem_u32(&tempty[acc]) & kPeerMask);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();            // nobody exits (or frees TMEM) while the peer may still signal / read
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace mg
