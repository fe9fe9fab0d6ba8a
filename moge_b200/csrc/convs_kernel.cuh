// "Swapped" implicit-GEMM 3x3 convolution for 64-channel maps (decoder levels 3 and 4).
//
// Idea: make the WEIGHTS the M-side operand (M = 64 output channels, resident in shared memory) and a 16x16-pixel window
// the N-side operand (N = 256), D^T[co, pixel] accumulating in TMEM, so that one instruction covers 256 pixels; the three
// vertical taps of a halo box are 2 KB-aligned row windows of the same box.
//
// STATUS (round 1): correct (tests/test_gpu_ops.py::test_conv3x3_swapped_operands) but measured ~15 % SLOWER end to end than
// conv64_kernel on the level-3 convs: tools/mma_cost.cu shows an M=64 tcgen05.mma costs exactly as much as an M=128 one
// (N/2 cycles, 48-cycle floor), so the swap halves the tensor throughput per instruction, and the scalar smem transpose plus
// half-idle store lanes add epilogue cost.  The engine only routes to it when MOGE_B200_CONVS=1.
//
// TMEM layout of an M=64 accumulator (cta_group::1): row r sits in lane (r%16) + 32*(r/16), i.e. each warp quarter owns
// 16 output channels in its lower 16 lanes.  The epilogue transposes 32-pixel chunks through shared memory so that
// every lane ends up with 4 channels of one pixel and the decoder epilogue logic (bias, UV, skip, ReLU copy, border
// replication, pixel shuffle) is the same as in umma_kernel.
#pragma once
#include "umma_kernel.cuh"

namespace mg {

constexpr int CS_TH = 16, CS_TW = 16;                       // output tile (pixels)
constexpr int CS_BOX_BYTES = (CS_TH + 2) * CS_TW * 128;     // box {64 ch, 16 px, 18 rows} = 36864 B

struct ConvSCfg {
    static constexpr int kStages = 3;
    static constexpr int kWBytes = 64 * 128 * 10;           // 9 taps + aux, each [64 co][64 k] swizzled
    static constexpr int kEpiWarps = 8;
    static constexpr int kThreads = 64 + 32 * kEpiWarps;
    static constexpr int kScratchBytes = kEpiWarps * 4096;
    static constexpr int kSmemBytes = kStages * CS_BOX_BYTES + kWBytes + 1024 + 256 + kScratchBytes;
};

// Epilogue of one 64-channel x 256-pixel accumulator for one warp: TMEM lane quarter `quarter` (channels 16*quarter ..
// +15 in lanes 0..15), pixel columns [col_begin, col_begin + 128).
template <bool BF16, int DF>
__device__ __forceinline__ void convs_epilogue_dec(const UmmaParams& p, int mt, int nt, uint32_t t_addr, float* scr, int quarter,
                                                   int lane, int col_begin) {
    using H = H16<BF16>;
    const bool has_raw = (DF < 0) ? (p.out0 != nullptr) : ((DF & DF_RAW) != 0);
    const bool has_relu = (DF < 0) ? (p.out1 != nullptr) : ((DF & DF_RELU) != 0);
    const bool has_skip = (DF < 0) ? (p.skip != nullptr) : ((DF & DF_SKIP) != 0);
    const bool has_uv = (DF < 0) ? (p.vec1 != nullptr) : ((DF & DF_UV) != 0);
    const bool shuffle = (DF < 0) ? (p.shuffle != 0) : ((DF & DF_SHUFFLE) != 0);
    const int sub = lane >> 3, q4 = lane & 7;
    const bool ch_ok = q4 < 4;                               // 16 channels per quarter -> 4 float4 groups
    const int per_img = p.tiles_x * p.tiles_y;
    const int b = mt / per_img, r = mt % per_img;
    const int y0 = (r / p.tiles_x) * CS_TH, x0 = (r % p.tiles_x) * CS_TW;
    // accumulator row (= output column of the GEMM) handled by this lane after the transpose
    const int n = nt * 64 + quarter * 16 + 4 * q4;
    int co = n, qd = 0;
    if (shuffle) { qd = n / p.ldo; co = n - qd * p.ldo; }    // N = 4*C_out: phase-major (ldo == C_out)
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = bias4, h4 = bias4;
    if (ch_ok) {
        bias4 = *reinterpret_cast<const float4*>(p.bias + co);
        if (has_uv) { g4 = *reinterpret_cast<const float4*>(p.vec1 + co); h4 = *reinterpret_cast<const float4*>(p.vec2 + co); }
    }
    const int sh = shuffle ? 2 : 1;
    const size_t tile_off = ((static_cast<size_t>(b) * p.Hop + sh * y0 + 1 + (qd >> 1)) * p.Wop + sh * x0 + 1 + (qd & 1)) * p.ldo * 2 +
                            static_cast<size_t>(co) * 2;
    const size_t row_pitch = static_cast<size_t>(sh) * p.Wop * p.ldo * 2, px_pitch = static_cast<size_t>(sh) * p.ldo * 2;
    const float inv_wo = 1.0f / static_cast<float>(p.Wo), inv_ho = 1.0f / static_cast<float>(p.Ho);
#pragma unroll 1
    for (int c = 0; c < 128; c += 32) {
        float v[32];
        tmem_ld32(t_addr + c, v);
        tc_wait_ld();
        // transpose [channel lane][32 pixels] -> pixel-major through a rotated (bank-conflict-free) smem tile
#pragma unroll
        for (int j = 0; j < 32; ++j) scr[lane * 32 + ((j + lane) & 31)] = v[j];
        __syncwarp();
        const int pix0 = col_begin + c;                      // window pixel index of column 0 of this chunk
        float4 pre[8];
        bool ok[8];
        size_t off[8];
        int eflag[8], PY[8], PX[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int pj = 4 * i + sub;
            const int wq = pix0 + pj, wy = wq >> 4, wx = wq & 15;
            const int y = y0 + wy, x = x0 + wx;
            ok[i] = ch_ok && (y < p.H) && (x < p.W);
            off[i] = tile_off + wy * row_pitch + wx * px_pitch;
            PY[i] = sh * y + (qd >> 1); PX[i] = sh * x + (qd & 1);
            eflag[i] = (PY[i] == 0 ? 1 : 0) | (PY[i] == p.Ho - 1 ? 2 : 0) | (PX[i] == 0 ? 4 : 0) | (PX[i] == p.Wo - 1 ? 8 : 0);
            pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (has_skip && ok[i]) {
                const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(p.skip) + off[i]);
                const float2 f0 = H::unpack(u.x), f1 = H::unpack(u.y);
                pre[i] = make_float4(f0.x, f0.y, f1.x, f1.y);
            }
        }
        uint2 pk_raw[8], pk_relu[8];
        unsigned edge_rows = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int pj = 4 * i + sub;
            const int c0 = 4 * q4;
            float4 a;
            a.x = scr[(c0 + 0) * 32 + ((pj + c0 + 0) & 31)];
            a.y = scr[(c0 + 1) * 32 + ((pj + c0 + 1) & 31)];
            a.z = scr[(c0 + 2) * 32 + ((pj + c0 + 2) & 31)];
            a.w = scr[(c0 + 3) * 32 + ((pj + c0 + 3) & 31)];
            if (!ok[i]) continue;
            a.x += bias4.x + pre[i].x; a.y += bias4.y + pre[i].y; a.z += bias4.z + pre[i].z; a.w += bias4.w + pre[i].w;
            if (has_uv) {
                const float uu = p.su * ((2 * PX[i] + 1) * inv_wo - 1.0f);
                const float vv = p.sv * ((2 * PY[i] + 1) * inv_ho - 1.0f);
                a.x += g4.x * uu + h4.x * vv; a.y += g4.y * uu + h4.y * vv;
                a.z += g4.z * uu + h4.z * vv; a.w += g4.w * uu + h4.w * vv;
            }
            if (eflag[i] != 0) edge_rows |= 1u << i;
            if (has_raw) {
                pk_raw[i].x = H::pack(a.x, a.y); pk_raw[i].y = H::pack(a.z, a.w);
                *reinterpret_cast<uint2*>(static_cast<uint8_t*>(p.out0) + off[i]) = pk_raw[i];
            }
            if (has_relu) {
                pk_relu[i].x = H::pack(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)); pk_relu[i].y = H::pack(fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
                *reinterpret_cast<uint2*>(static_cast<uint8_t*>(p.out1) + off[i]) = pk_relu[i];
            }
        }
        if (edge_rows != 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (edge_rows >> i & 1) {
                    if (has_raw) store_px_border8(static_cast<uint8_t*>(p.out0), b, PY[i], PX[i], p.Ho, p.Wo, p.Hop, p.Wop, p.ldo, co, pk_raw[i]);
                    if (has_relu) store_px_border8(static_cast<uint8_t*>(p.out1), b, PY[i], PX[i], p.Ho, p.Wo, p.Hop, p.Wop, p.ldo, co, pk_relu[i]);
                }
        }
        __syncwarp();
    }
}

template <int EPI, bool BF16, int DF>
__global__ void __launch_bounds__(ConvSCfg::kThreads, 1)
convs_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAux,
             const __grid_constant__ CUtensorMap mapW, const UmmaParams p) {
    using Cfg = ConvSCfg;
    constexpr int S = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sW = smem + S * CS_BOX_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(sW + Cfg::kWBytes);
    uint64_t* empty = full + S;
    uint64_t* tfull = empty + S;
    uint64_t* tempty = tfull + 2;
    uint64_t* wfull = tempty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull + 1);
    float* scratch_base = reinterpret_cast<float*>(sW + Cfg::kWBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nnt = p.num_n_tiles;                    // 64-channel output groups (N / 64)
    const int nt = blockIdx.x % nnt;
    const int mt0 = blockIdx.x / nnt, mstep = gridDim.x / nnt;
    const int nstage = 3 + p.kb_aux;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapW);
        if (p.kb_aux) tma_prefetch_desc(&mapAux);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], Cfg::kEpiWarps); }
        mbar_init(wfull, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const int nblk = 9 + p.kb_aux;
            mbar_arrive_expect_tx(wfull, nblk * 64 * 128);
            for (int t = 0; t < nblk; ++t) tma_load_2d(sW + t * 64 * 128, &mapW, wfull, t * TILE_K, nt * 64);
            int s = 0; uint32_t ph = 0;
            const int per_img = p.tiles_x * p.tiles_y;
            for (int mt = mt0; mt < p.num_m_tiles; mt += mstep) {
                const int b = mt / per_img, r = mt % per_img;
                const int y0 = (r / p.tiles_x) * CS_TH, x0 = (r % p.tiles_x) * CS_TW;
                for (int i = 0; i < nstage; ++i) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* sa = smem + s * CS_BOX_BYTES;
                    if (i < 3) {
                        mbar_arrive_expect_tx(&full[s], CS_BOX_BYTES);
                        tma_load_4d(sa, &mapA, &full[s], 0, x0 + i, y0, b);          // padded rows y0 .. y0+17
                    } else {
                        mbar_arrive_expect_tx(&full[s], CS_TH * CS_TW * 128);
                        tma_load_4d(sa, &mapAux, &full[s], 0, x0 + 1, y0 + 1, b);    // box {64, 16, 16}
                    }
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(64, 256, BF16 ? 1u : 0u);          // M = 64 output channels, N = 256 pixels
            mbar_wait(wfull, 0);
            int s = 0; uint32_t ph = 0;
            int it = 0;
            const uint32_t w0 = smem_u32(sW);
            for (int mt = mt0; mt < p.num_m_tiles; mt += mstep, ++it) {
                const int acc = it & 1;
                mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int i = 0; i < nstage; ++i) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * CS_BOX_BYTES);
                    if (i < 3) {
#pragma unroll
                        for (int dy = 0; dy < 3; ++dy) {
                            const uint64_t wdesc = make_sdesc_sw128(w0 + (dy * 3 + i) * 64 * 128);
                            const uint64_t xdesc = make_sdesc_sw128(sa + dy * CS_TW * 128);
#pragma unroll
                            for (int k = 0; k < TILE_K / 16; ++k) umma_f16(d_tmem, wdesc + 2 * k, xdesc + 2 * k, idesc, (i | dy | k) != 0);
                        }
                    } else {
                        const uint64_t wdesc = make_sdesc_sw128(w0 + 9 * 64 * 128);
                        const uint64_t xdesc = make_sdesc_sw128(sa);
#pragma unroll
                        for (int k = 0; k < TILE_K / 16; ++k) umma_f16(d_tmem, wdesc + 2 * k, xdesc + 2 * k, idesc, 1);
                    }
                    umma_commit(&empty[s]);
                    if (++s == S) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int col_begin = (ew >> 2) * 128;          // the two warps of a lane quarter split the 256 pixel columns
        float* scr = scratch_base + ew * 1024;
        int it = 0;
        for (int mt = mt0; mt < p.num_m_tiles; mt += mstep, ++it) {
            const int acc = it & 1;
            mbar_wait(&tfull[acc], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256 + col_begin;
            convs_epilogue_dec<BF16, DF>(p, mt, nt, t_addr, scr, quarter, lane, col_begin);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace mg
