// 3x3 replicate-padded convolution for 64-channel inputs (decoder levels 3 and 4 -- the HBM-bound end of the pyramid).
// Same tcgen05/TMEM machinery and the same epilogues as umma_kernel<TILES>, but with the two changes that matter when
// K per tap is a single 64-channel block:
//   * the packed weights of this CTA's output-channel tile (9 taps [+ the fused 1x1 input block]) are loaded ONCE and
//     stay resident in shared memory for the whole persistent loop;
//   * the input halo is fetched as 3 boxes per tile (one per horizontal tap offset, 16 px x 10 rows) instead of 9: the
//     three vertical taps of a box are 2 KB-aligned row windows of the same box, so each becomes a UMMA A operand
//     without another copy (L2->SM traffic 3.75x the tile instead of 9x + weights).
#pragma once
#include "umma_kernel.cuh"

namespace mg {

#ifdef MG_C64_DEBUG
// wait-time attribution (debug builds only): [0] producer waits for a free stage, [1] MMA waits for data, [2] MMA waits for
// a free accumulator, [3] epilogue warp 2 waits for an accumulator, [4] epilogue warp 2 busy, [5] CTA lifetime, [6] tiles
__device__ unsigned long long mg_c64_dbg[8];
#define C64_T0() const long long _t0 = clock64()
#define C64_ACC(var) var += clock64() - _t0
#else
#define C64_T0()
#define C64_ACC(var)
#endif

template <int BN, int MW = 1> struct Conv64Cfg {
    static constexpr int kABytes = 160 * 128;                       // box {64 ch, 16 px, 10 rows}
    // resident weights: 9 taps + 1 aux block, each [BN][64] 128B-swizzled.  The two-issuer form never runs aux launches
    // (launch_conv64), so it keeps 9 blocks and spends the 8 KB on a sixth halo slot: two rings of three slots = one whole tile each
#ifndef MG_C64_MW2_STAGES
#define MG_C64_MW2_STAGES 6
#endif
    static constexpr int kWBlocks = (MW == 2) ? 9 : 10;
    static constexpr int kWBytes = BN * 128 * kWBlocks;
#ifndef MG_C64_GROUPS
#define MG_C64_GROUPS 2
#endif
    static constexpr int kEpiGroups = MG_C64_GROUPS;
    static constexpr int kStages = (BN >= 64) ? (kEpiGroups > 2 ? 4 : (MW == 2 ? MG_C64_MW2_STAGES : 5)) : 6;
    // 2 TMEM accumulator stages per epilogue group; the epilogue warps form kEpiGroups groups of 4 (one warp per TMEM lane quarter) that drain
    // tiles round-robin, concurrently, so the per-tile epilogue latency chain (TMEM load -> transpose -> global) overlaps.
    static constexpr int kAccStages = 2 * kEpiGroups;
    static constexpr int kEpiWarps = 4 * kEpiGroups;
    // MW = MMA-issuing warps.  With K = 64 per tap an N = 64 MMA lasts ~48 cycles and a stage is only 12 of them: the issuing
    // warp's own per-stage work (barrier wait, elect, ~60 uniform-datapath instructions of descriptor arithmetic, commits; ncu
    // source page: the warp is blocked on a full tensor queue for only 24 % of its samples) is longer than the MMAs it feeds.
    // With MW = 2 the two issuing warps take alternate tiles (separate accumulator stages), so one warp's bookkeeping runs under
    // the other's MMAs.  Same-box A/B, B = 32 ViT-L: level-3 res_a 0.637 -> 0.560 ms, res_b 0.691 -> 0.645, the N = 16 / 32 output
    // convs 1.18 -> 0.89 ms; launches with the extra 1x1 aux stage (4 stages per tile on rings of 3 + 2 slots) lose 4 % and
    // keep MW = 1 (launch_conv64 picks per launch).
    static constexpr int kMmaWarps = MW;
    // two issuing warps: the halo ring is split into one ring per warp (slots [0, kRing0) and [kRing0, kStages)), so that every
    // full / empty barrier is waited on by exactly one consumer and the phase-parity bookkeeping stays valid
    static constexpr int kRing0 = (kMmaWarps > 1) ? (kStages + 1) / 2 : kStages;
    static constexpr int kFirstEpiWarp = 1 + kMmaWarps;
    static constexpr int kThreads = 32 * (1 + kMmaWarps + kEpiWarps);
    static_assert(kMmaWarps == 1 || (kMmaWarps == 2 && kAccStages % 2 == 0), "conv64_kernel: 1 or 2 MMA warps (tiles alternate; accumulator stages must split evenly)");
    static constexpr int kTmemCols = (kAccStages * BN <= 32) ? 32 : (kAccStages * BN <= 64) ? 64 : (kAccStages * BN <= 128) ? 128 : (kAccStages * BN <= 256) ? 256 : 512;
    static constexpr int kScratchBytes = kEpiWarps * 4096;
    static constexpr int kSmemBytes = kStages * kABytes + kWBytes + 1024 + 256 + kScratchBytes;
    static constexpr int kColsPerWarp = BN;
    static_assert(kSmemBytes <= kMaxDynSmem, "conv64_kernel: halo ring + resident weights + scratch exceed the shared memory of one CTA");
    static_assert((2 * kStages + 2 * kAccStages + 1) * 8 + 4 <= 256, "conv64_kernel: barrier block overflows its 256 bytes");
    static_assert(kAccStages * BN <= 512, "conv64_kernel: accumulator stages exceed tensor memory");
};

template <int BN, int EPI, bool BF16, int DF, int MMAW>
__global__ void __launch_bounds__(Conv64Cfg<BN, MMAW>::kThreads, 1)
conv64_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAux,
              const __grid_constant__ CUtensorMap mapW, const UmmaParams p) {
    pdl_launch_dependents();      // (the wait sits after the barrier / TMEM set-up below: that prologue overlaps the previous kernel's tail)
    using Cfg = Conv64Cfg<BN, MMAW>;
    constexpr int S = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem + S * Cfg::kABytes;
    uint64_t* full = reinterpret_cast<uint64_t*>(sW + Cfg::kWBytes);
    uint64_t* empty = full + S;
    uint64_t* tfull = empty + S;
    uint64_t* tempty = tfull + Cfg::kAccStages;
    uint64_t* wfull = tempty + Cfg::kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
    float* scratch_base = reinterpret_cast<float*>(sW + Cfg::kWBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef MG_C64_DEBUG
    long long dbg_a = 0, dbg_b = 0;
    const long long dbg_start = clock64();
#endif
    const int nnt = p.num_n_tiles;
    const int nt = blockIdx.x % nnt;                  // this CTA's output-channel tile (weights stay resident)
    const int mt0 = blockIdx.x / nnt, mstep = gridDim.x / nnt;
    const int nstage = 3 + p.kb_aux;                  // stages per tile: dx = 0,1,2 (+ the 1x1 aux source)

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
        if (p.kb_aux) tma_prefetch_desc(&mapAux);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < Cfg::kAccStages; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
        mbar_init(wfull, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == 0) {
        // The whole warp walks the loop (converged) and ONE ELECTED lane issues: inside `if (lane == 0)` the compiler must
        // assume an arbitrary active mask and wraps every TMA / tcgen05 instruction in an elect-and-retry loop (~9
        // dependent instructions per MMA -- more than a 48-cycle N = 64 MMA takes to execute).
        {
            const int nblk = (Cfg::kWBlocks == 9) ? 9 : 9 + p.kb_aux;
            if (elect_one()) {
                mbar_arrive_expect_tx(wfull, nblk * BN * 128);
                for (int t = 0; t < nblk; ++t) tma_load_2d(sW + t * BN * 128, &mapW, wfull, t * TILE_K, nt * BN);
            }
            __syncwarp();
            int s0 = 0, s1 = 0; uint32_t ph0 = 0, ph1 = 0;    // ring positions (ring 1 only with two issuing warps)
            int it = 0;
            const int per_img = p.tiles_x * p.tiles_y;
            for (int mt = mt0; mt < p.num_m_tiles; mt += mstep, ++it) {
                const int b = mt / per_img, r = mt % per_img;
                const int y0 = (r / p.tiles_x) * TILE_PH, x0 = (r % p.tiles_x) * TILE_PW;
                const bool ring1 = Cfg::kMmaWarps > 1 && (it & 1);
                const int rbase = ring1 ? Cfg::kRing0 : 0, rdepth = ring1 ? S - Cfg::kRing0 : Cfg::kRing0;
                int s = ring1 ? s1 : s0; uint32_t ph = ring1 ? ph1 : ph0;
                for (int i = 0; i < nstage; ++i) {
                    const int slot = rbase + s;
                    { C64_T0(); mbar_wait(&empty[slot], ph ^ 1); C64_ACC(dbg_a); }
                    uint8_t* sa = smem + slot * Cfg::kABytes;
                    if (elect_one()) {
                        if (i < 3) {
                            mbar_arrive_expect_tx(&full[slot], Cfg::kABytes);
                            tma_load_4d(sa, &mapA, &full[slot], 0, x0 + i, y0, b);          // padded rows y0..y0+9 = taps dy 0..2
                        } else {
                            mbar_arrive_expect_tx(&full[slot], TILE_M * 128);
                            tma_load_4d(sa, &mapAux, &full[slot], 0, x0 + 1, y0 + 1, b);
                        }
                    }
                    __syncwarp();
                    if (++s == rdepth) { s = 0; ph ^= 1; }
                }
                if (ring1) { s1 = s; ph1 = ph; } else { s0 = s; ph0 = ph; }
            }
        }
    } else if (warp <= Cfg::kMmaWarps) {
        {
            constexpr uint32_t idesc = make_idesc(TILE_M, BN, BF16 ? 1u : 0u);
            constexpr int MW = Cfg::kMmaWarps;
            const int mw = warp - 1;                   // this warp issues the MMAs of tiles mw, mw + MW, ... of the CTA
            mbar_wait(wfull, 0);
            const uint32_t w0 = smem_u32(sW);
            const int rbase = mw ? Cfg::kRing0 : 0, rdepth = mw ? S - Cfg::kRing0 : Cfg::kRing0;      // this warp's halo ring
            int s = 0; uint32_t ph = 0;
            int it = mw;
            for (int mt = mt0 + mw * mstep; mt < p.num_m_tiles; mt += MW * mstep, it += MW) {
                const int acc = it % Cfg::kAccStages;
                { C64_T0(); mbar_wait(&tempty[acc], ((it / Cfg::kAccStages) & 1) ^ 1); C64_ACC(dbg_b); }
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int i = 0; i < nstage; ++i) {
                    const int slot = rbase + s;
                    { C64_T0(); mbar_wait(&full[slot], ph); C64_ACC(dbg_a); }
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + slot * Cfg::kABytes);
                    if (elect_one()) {
                        if (i < 3) {
#pragma unroll
                            for (int dy = 0; dy < 3; ++dy) {
                                const uint64_t adesc = make_sdesc_sw128(sa + dy * TILE_PW * 128);
                                const uint64_t bdesc = make_sdesc_sw128(w0 + (dy * 3 + i) * BN * 128);
#pragma unroll
                                for (int k = 0; k < TILE_K / 16; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (i | dy | k) != 0);
                            }
                        } else {
                            const uint64_t adesc = make_sdesc_sw128(sa);
                            const uint64_t bdesc = make_sdesc_sw128(w0 + 9 * BN * 128);
#pragma unroll
                            for (int k = 0; k < TILE_K / 16; ++k) umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
                        }
                        umma_commit(&empty[slot]);
                        if (i == nstage - 1) umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++s == rdepth) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        const int ew = warp - Cfg::kFirstEpiWarp;
        const int quarter = warp & 3;
        const int group = ew >> 2;                     // epilogue group g drains tiles g, g + G, ...
        float4* scr = reinterpret_cast<float4*>(scratch_base + ew * 1024);
        int it = group;
        for (int mt = mt0 + group * mstep; mt < p.num_m_tiles; mt += Cfg::kEpiGroups * mstep, it += Cfg::kEpiGroups) {
            const int acc = it % Cfg::kAccStages;
            { C64_T0(); mbar_wait(&tfull[acc], (it / Cfg::kAccStages) & 1); C64_ACC(dbg_a); }
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
            { C64_T0();
            epilogue_tile<BN, Cfg::kColsPerWarp, AMODE_TILES, EPI, BF16, DF>(p, mt, nt, t_addr, scr, quarter, lane, 0);
            C64_ACC(dbg_b); }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
#ifdef MG_C64_DEBUG
    if (lane == 0) {
        if (warp == 0) { atomicAdd(&mg_c64_dbg[0], (unsigned long long)dbg_a); atomicAdd(&mg_c64_dbg[5], (unsigned long long)(clock64() - dbg_start)); }
        if (warp == 1) { atomicAdd(&mg_c64_dbg[1], (unsigned long long)dbg_a); atomicAdd(&mg_c64_dbg[2], (unsigned long long)dbg_b); }
        if (warp == Cfg::kFirstEpiWarp) { atomicAdd(&mg_c64_dbg[3], (unsigned long long)dbg_a); atomicAdd(&mg_c64_dbg[4], (unsigned long long)dbg_b); }
    }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::kTmemCols);
    }
}

}  // namespace mg
