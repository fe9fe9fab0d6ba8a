// 3x3 replicate-padded convolution for C_in >= 128 (decoder levels 1-2): the halo-box trick of conv64_kernel with STREAMED
// weights.  umma_kernel<TILES> fetches one 16x8-pixel box per filter tap and 64-channel block (9 boxes per block) and a
// weight block with every one of them; at C = 128 that is 576 KB of L2->SM traffic per 128-pixel tile against 4608 tensor
// cycles -- 125 B/clk per SM, the level is L2-bandwidth-bound (0.3 ms measured vs 0.115 ms of tensor time).  Here the pixels
// come as 3 boxes per channel block (one per horizontal tap offset, 16 px x 10 rows; the three vertical taps are 2 KB-aligned
// row windows of the same box), so the A traffic drops from 9x to 3.75x the tile; the weights keep streaming through their
// own ring (one [BN][64] block per tap and channel block).
//
// Two independent smem rings: A boxes (full_a/empty_a) and weight blocks (full_w/empty_w).  The producer issues them in the
// exact order the MMA warp consumes them: for each channel block c and horizontal tap i: box(c, i), W(c, i, dy = 0..2);
// then the optional aux blocks (fused 1x1 source): 8-row box, one weight block.
// Accumulators, epilogue warps and the epilogue itself are those of umma_kernel<BN, AMODE_TILES, EPI_DEC>.
#pragma once
#include "umma_kernel.cuh"

namespace mg {

template <int BN> struct ConvhCfg {
    static constexpr int kABytes = 160 * 128;                       // box {64 ch, 16 px, 10 rows}
    static constexpr int kWBytes = BN * 128;                        // [BN][64] 128B-swizzled
    // Weight blocks per ring stage.  With one block per stage the MMA warp goes through a barrier wait, an elect, the descriptor
    // arithmetic and a commit for every FOUR MMAs; at BN = 128 those last 256 tensor cycles while the loop iteration takes ~530
    // (ncu source page of convh_kernel<128>: the issuing warp never finds the tensor queue full, tensor pipe 50 % active).  Grouping
    // the three vertical taps of one (channel block, horizontal tap) into one stage gives 12 MMAs per iteration.  The ring holds
    // the same bytes (2 x 3 blocks instead of 6 x 1); BN = 256 stages would not fit three blocks and are less overhead-bound.
    // Same-box A/B, B = 32 ViT-L: level-2 res_a 0.531 -> 0.465 ms, res_b 0.561 -> 0.510 ms.
#ifndef MG_CONVH_WGROUP
#define MG_CONVH_WGROUP 3      // 1 = one weight block per stage (A/B builds)
#endif
    static constexpr int kWGroup = (MG_CONVH_WGROUP == 3 && BN < 256) ? 3 : 1;
    static constexpr int kWStageBytes = kWGroup * kWBytes;
    static constexpr int kAStages = (BN >= 256) ? 3 : 4;
    static constexpr int kWStages = (BN >= 256) ? 4 : (kWGroup == 3 ? 2 : 6);
    static constexpr int kEpiWarps = 8;
    static constexpr int kThreads = 64 + 32 * kEpiWarps;
    static constexpr int kTmemCols = (2 * BN <= 256) ? 256 : 512;
    static constexpr int kScratchBytes = kEpiWarps * 4096;
    static constexpr int kSmemBytes = kAStages * kABytes + kWStages * kWStageBytes + 1024 + 256 + kScratchBytes;
    static constexpr int kColsPerWarp = BN / 2;
    static_assert(kSmemBytes <= kMaxDynSmem, "convh_kernel: box ring + weight ring + scratch exceed the shared memory of one CTA");
    static_assert((2 * kAStages + 2 * kWStages + 4) * 8 + 4 <= 256, "convh_kernel: barrier block overflows its 256 bytes");
};

template <int BN, bool BF16, int DF>
__global__ void __launch_bounds__(ConvhCfg<BN>::kThreads, 1)
convh_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAux,
             const __grid_constant__ CUtensorMap mapW, const UmmaParams p) {
    pdl_launch_dependents();      // (the wait sits after the barrier / TMEM set-up below: that prologue overlaps the previous kernel's tail)
    using Cfg = ConvhCfg<BN>;
    constexpr int SA = Cfg::kAStages, SW = Cfg::kWStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;
    uint8_t* sW = smem + SA * Cfg::kABytes;
    uint64_t* full_a = reinterpret_cast<uint64_t*>(sW + SW * Cfg::kWStageBytes);
    uint64_t* empty_a = full_a + SA;
    uint64_t* full_w = empty_a + SA;
    uint64_t* empty_w = full_w + SW;
    uint64_t* tfull = empty_w + SW;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* scratch_base = reinterpret_cast<float*>(sW + SW * Cfg::kWStageBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = p.num_m_tiles * p.num_n_tiles;
    const int kbm = p.kb_main, kba = p.kb_aux;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
        if (kba) tma_prefetch_desc(&mapAux);
        for (int s = 0; s < SA; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1); }
        for (int s = 0; s < SW; ++s) { mbar_init(&full_w[s], 1); mbar_init(&empty_w[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], Cfg::kEpiWarps); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // ================================================================== TMA producer (converged warp, elected lane issues)
        int sa = 0, sw = 0; uint32_t pha = 0, phw = 0;
        const int per_img = p.tiles_x * p.tiles_y;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
            const int b = mt / per_img, r = mt % per_img;
            const int y0 = (r / p.tiles_x) * TILE_PH, x0 = (r % p.tiles_x) * TILE_PW;
            for (int c = 0; c < kbm + kba; ++c) {
                const bool aux = c >= kbm;
                const int ni = aux ? 1 : 3;
                for (int i = 0; i < ni; ++i) {
                    mbar_wait(&empty_a[sa], pha ^ 1);
                    if (elect_one()) {
                        uint8_t* dst = sA + sa * Cfg::kABytes;
                        if (!aux) {
                            mbar_arrive_expect_tx(&full_a[sa], Cfg::kABytes);
                            tma_load_4d(dst, &mapA, &full_a[sa], c * TILE_K, x0 + i, y0, b);     // padded rows y0..y0+9 = taps dy 0..2
                        } else {
                            mbar_arrive_expect_tx(&full_a[sa], TILE_M * 128);
                            tma_load_4d(dst, &mapAux, &full_a[sa], (c - kbm) * TILE_K, x0 + 1, y0 + 1, b);
                        }
                    }
                    __syncwarp();
                    if (++sa == SA) { sa = 0; pha ^= 1; }
                    const int nw = aux ? 1 : 3;
                    if (Cfg::kWGroup == 3) {         // one stage = the (up to) three vertical-tap blocks of this box
                        mbar_wait(&empty_w[sw], phw ^ 1);
                        if (elect_one()) {
                            mbar_arrive_expect_tx(&full_w[sw], nw * Cfg::kWBytes);
                            for (int dy = 0; dy < nw; ++dy) {
                                const int kblk = aux ? 9 * kbm + (c - kbm) : (dy * 3 + i) * kbm + c;
                                tma_load_2d(sW + sw * Cfg::kWStageBytes + dy * Cfg::kWBytes, &mapW, &full_w[sw], kblk * TILE_K, nt * BN);
                            }
                        }
                        __syncwarp();
                        if (++sw == SW) { sw = 0; phw ^= 1; }
                    } else
                    for (int dy = 0; dy < nw; ++dy) {
                        mbar_wait(&empty_w[sw], phw ^ 1);
                        if (elect_one()) {
                            const int kblk = aux ? 9 * kbm + (c - kbm) : (dy * 3 + i) * kbm + c;
                            mbar_arrive_expect_tx(&full_w[sw], Cfg::kWBytes);
                            tma_load_2d(sW + sw * Cfg::kWBytes, &mapW, &full_w[sw], kblk * TILE_K, nt * BN);
                        }
                        __syncwarp();
                        if (++sw == SW) { sw = 0; phw ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        constexpr uint32_t idesc = make_idesc(TILE_M, BN, BF16 ? 1u : 0u);
        int sa = 0, sw = 0; uint32_t pha = 0, phw = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            uint32_t first = 0;
            for (int c = 0; c < kbm + kba; ++c) {
                const bool aux = c >= kbm;
                const int ni = aux ? 1 : 3;
                for (int i = 0; i < ni; ++i) {
                    mbar_wait(&full_a[sa], pha);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(sA + sa * Cfg::kABytes);
                    const int nw = aux ? 1 : 3;
                    if (Cfg::kWGroup == 3) {
                        mbar_wait(&full_w[sw], phw);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t wbase = smem_u32(sW + sw * Cfg::kWStageBytes);
                            if (!aux) {
#pragma unroll
                                for (int dy = 0; dy < 3; ++dy) {
                                    const uint64_t adesc = make_sdesc_sw128(a0 + dy * TILE_PW * 128);
                                    const uint64_t bdesc = make_sdesc_sw128(wbase + dy * Cfg::kWBytes);
#pragma unroll
                                    for (int k = 0; k < TILE_K / 16; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, first | dy | k);
                                }
                            } else {
                                const uint64_t adesc = make_sdesc_sw128(a0);
                                const uint64_t bdesc = make_sdesc_sw128(wbase);
#pragma unroll
                                for (int k = 0; k < TILE_K / 16; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, first | k);
                            }
                            umma_commit(&empty_w[sw]);
                            umma_commit(&empty_a[sa]);
                            if (c == kbm + kba - 1 && i == ni - 1) umma_commit(&tfull[acc]);
                        }
                        __syncwarp();
                        first = 1;
                        if (++sw == SW) { sw = 0; phw ^= 1; }
                    } else
                    for (int dy = 0; dy < nw; ++dy) {
                        mbar_wait(&full_w[sw], phw);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint64_t adesc = make_sdesc_sw128(a0 + dy * TILE_PW * 128);
                            const uint64_t bdesc = make_sdesc_sw128(smem_u32(sW + sw * Cfg::kWBytes));
#pragma unroll
                            for (int k = 0; k < TILE_K / 16; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, first | k);
                            umma_commit(&empty_w[sw]);
                            if (dy == nw - 1) {
                                umma_commit(&empty_a[sa]);
                                if (c == kbm + kba - 1 && i == ni - 1) umma_commit(&tfull[acc]);
                            }
                        }
                        __syncwarp();
                        first = 1;
                        if (++sw == SW) { sw = 0; phw ^= 1; }
                    }
                    if (++sa == SA) { sa = 0; pha ^= 1; }
                }
            }
        }
    } else {
        // ================================================================== epilogue (as umma_kernel<BN, AMODE_TILES, EPI_DEC>)
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int col_begin = (ew >> 2) * (BN / 2);
        float4* scr = reinterpret_cast<float4*>(scratch_base + ew * 1024);
        int it = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
            const int acc = it & 1;
            mbar_wait(&tfull[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + col_begin;
            epilogue_tile<BN, Cfg::kColsPerWarp, AMODE_TILES, EPI_DEC, BF16, DF>(p, mt, nt, t_addr, scr, quarter, lane, col_begin);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace mg
