// The one tensor-core kernel of the engine: a persistent, warp-specialised tcgen05 GEMM whose A operand is either
// a row-major matrix (encoder linears) or an 8x16-pixel window of a padded NHWC image shifted by a filter tap
// (decoder implicit-GEMM convolutions).  TMA -> 128B-swizzled smem ring -> tcgen05.mma (fp32 accumulators in
// TMEM, double-buffered) -> fused epilogue straight from TMEM.
//
// Roles (warp-uniform): warp 0 = TMA producer (1 lane), warp 1 = MMA issuer (1 lane) + TMEM owner,
// warps 2.. = epilogue (each warp owns the TMEM lane quarter (warp & 3); with 8 epilogue warps the two warps of a
// quarter split the accumulator columns).
//
// Reference ops covered: K2 patch-embed, K4 qkv, K6 proj+LayerScale+residual, K7 fc1+GELU,
// K8 fc2+LayerScale+residual, K9 tap projection sum, K10/K11 1x1 + UV, K12 ConvTranspose2d k2s2, K13/K14 3x3
// replicate-padded conv with ReLU/residual, K17 head output projection.
#pragma once
#include "common.cuh"
#include <type_traits>

namespace mg {

enum : int { AMODE_ROWS = 0, AMODE_TILES = 1 };
enum : int {
    EPI_STORE16 = 0,   // out16[row, col] = T(acc + bias[col])
    EPI_GELU16 = 1,    // out16[row, col] = T(gelu_erf(acc + bias[col]))
    EPI_RESID = 2,     // x32[row, col] += gamma[col] * (acc + bias[col])            (fp32 residual stream, in place)
    EPI_PATCH = 3,     // x32[b*(T+1)+1+t, col] = acc + table[t, col]                (patch embed + pos embed)
    EPI_DEC = 4,       // padded-NHWC decoder store: acc + bias (+ UV rank-2) (+ skip) -> raw and/or ReLU copies
    EPI_HEADOUT = 5,   // BN=16: (bilinear x2 + 3x3 conv + output 1x1) folded into one low-res 3x3 conv with 4 output phases,
                       //        + folded 1x1 of the high-res neck map -> fp32 maps at the high resolution
    EPI_NECKOUT = 6,   // BN=32: the neck's last level (bilinear x2 + 3x3 conv + UV 1x1) folded THROUGH the heads' last-level
                       //        input + output blocks: 4 phases x 8 components (points xyz, normal xyz, mask logit, pad) written
                       //        straight into the heads' fp32 output maps; the heads' EPI_HEADOUT launches then accumulate
};

#ifndef MG_STAGES256
#define MG_STAGES256 4
#endif
constexpr int TILE_M = 128;
constexpr int TILE_K = 64;     // 64 x 16-bit = one 128-byte swizzle row
constexpr int TILE_PW = 16;    // pixel tile = 8 rows x 16 columns
constexpr int TILE_PH = 8;

struct UmmaParams {
    // GEMM shape
    int M;                 // AMODE_ROWS: number of valid rows
    int N;                 // total output columns (multiple of BN)
    int ntaps;             // 1 (centre) or 9 (3x3)
    int kb_main;           // 64-wide K blocks per tap from the main source
    int kb_aux;            // 64-wide K blocks from the aux source (centre tap), appended after the taps
    int num_m_tiles, num_n_tiles;
    // pixel geometry (AMODE_TILES: of the A source; EPI_DEC/HEADOUT: output pixel grid derives from it)
    int B, H, W;           // A-source unpadded size (AMODE_ROWS + EPI_DEC: the h x w token grid)
    int tiles_x, tiles_y;
    // epilogue operands
    void* out0;            // EPI_STORE16/GELU16: T* ; EPI_RESID/PATCH: float* ; EPI_DEC: raw T* (or null) ; HEADOUT: float*
    void* out1;            // EPI_DEC: ReLU copy T* (or null); NECKOUT: normal map float4* (or null)
    void* out2;            // NECKOUT: mask-logit map float* (or null)
    int accum;             // HEADOUT: 1 = add to the values already in out0 (written by the NECKOUT launch)
    const float* bias;     // [N] (EPI_DEC shuffle: [C_out]); HEADOUT: [16]
    const float* vec1;     // EPI_RESID: gamma[N]; EPI_PATCH: table[T, N]; EPI_DEC: wu[C] (or null); HEADOUT: waux[ncomp,32] (or null)
    const float* vec2;     // EPI_DEC: wv[C]
    const void* skip;      // EPI_DEC: residual input, same geometry as out ; HEADOUT: neck level-4 map (T*, 32 ch, padded)
    int ldo;               // row pitch of out (elements): ROWS epilogues: N total; DEC: channels of the out buffer
    int T;                 // EPI_PATCH / ROWS+EPI_DEC: tokens per image
    int Ho, Wo, Hop, Wop;  // EPI_DEC/HEADOUT: output pixel grid and its padded allocation
    int shuffle;           // EPI_DEC: 1 = ConvTranspose2d k2s2 pixel shuffle (Ho = 2H, Wo = 2W), N = 4*C_out
    int ncomp;             // HEADOUT: 3 (float4 per pixel) or 1 (float per pixel)
    float su, sv;          // UV half extents
    // LayerNorm folded into the consuming GEMM (ROWS epilogues; see elementwise.cu "LayerNorm folded into the next GEMM")
    void* x16;                 // producers (EPI_RESID / EPI_PATCH): 16-bit copy of the rows written to out0 (pitch ldo), or null
    float2* stats_out;         // producers: [row, stats_ld] partial (sum, sum of squares) of the ROUNDED row, one per column group
    const float* ln_rstd;      // consumers (EPI_STORE16 / EPI_GELU16): 1/sqrt(var + eps) of the A rows; null = plain bias epilogue
    int stats_ld;              // partial sums per row in stats_out
};

template <int BN> struct UmmaCfg {
    static constexpr int kStageBytes = TILE_M * 128 + BN * 128;
    static constexpr int kStages = (BN >= 256) ? MG_STAGES256 : (BN >= 128) ? 6 : 8;
    static constexpr int kEpiWarps = (BN >= 64) ? 8 : 4;
    static constexpr int kThreads = 64 + 32 * kEpiWarps;
    static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int kScratchBytes = kEpiWarps * 4096;   // per-epilogue-warp 32x32 fp32 transpose tile
    static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kScratchBytes;
    static constexpr int kColsPerWarp = (kEpiWarps == 8) ? BN / 2 : BN;
    static_assert(kSmemBytes <= kMaxDynSmem, "umma_kernel: stage ring + scratch exceed the shared memory of one CTA");
    static_assert((2 * kStages + 4) * 8 + 4 <= 256, "umma_kernel: barrier block overflows its 256 bytes");
};

// Exact-erf GELU (nn.GELU() default, vision_transformer.py:61): gelu(x) = relu(x) - 0.5 |x| erfc(|x|/sqrt2), with
// erfc(|x|/sqrt2) = 2^q(|x|), q a degree-5 fit of log2(erfc) on [0,6] (weighted for the GELU error; max |gelu error|
// 7e-7 in fp32, far below the 16-bit output rounding).  5 FMA + 1 MUFU.EX2 + 4 other instructions instead of libdevice
// erff's ~25: the fc1 epilogue is ALU-issue-bound, this is what keeps the tensor pipe fed.  (An A&S 7.1.26 variant with
// MUFU.RCP + MUFU.EX2 was measured 25 % slower than erff: two quarter-rate MUFU ops per element are too many.)
// (No clamp of |x|: the polynomial keeps decreasing beyond the fitted range -- q(6) = -29.2, q(8) = -51.8, q(20) = -1019 -- so
// |x| * 2^q just underflows to 0 for large |x|; the clamp cost one instruction per element in an epilogue that is bound by
// instruction issue.)
__device__ __forceinline__ float gelu_erf(float x) {
    const float ax = fabsf(x);
    float q = fmaf(-4.804994387e-04f, ax, 7.133518346e-03f);
    q = fmaf(q, ax, -5.194617063e-02f);
    q = fmaf(q, ax, -4.598676562e-01f);
    q = fmaf(q, ax, -1.150842190e+00f);
    q = fmaf(q, ax, -3.041332639e-05f);
    const float t = ax * ex2_approx(q);
    return fmaf(-0.5f, t, fmaxf(x, 0.0f));
}

// Two lanes of gelu_erf with packed FFMA2/FMUL2 (bit-identical to the scalar version: fma.rn.f32x2 is fmaf per lane).
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
    const float2 ab = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 ax = ab;
    float2 q = ffma2(make_float2(-4.804994387e-04f, -4.804994387e-04f), ax, make_float2(7.133518346e-03f, 7.133518346e-03f));
    q = ffma2(q, ax, make_float2(-5.194617063e-02f, -5.194617063e-02f));
    q = ffma2(q, ax, make_float2(-4.598676562e-01f, -4.598676562e-01f));
    q = ffma2(q, ax, make_float2(-1.150842190e+00f, -1.150842190e+00f));
    q = ffma2(q, ax, make_float2(-3.041332639e-05f, -3.041332639e-05f));
    const float2 t = fmul2(ab, make_float2(ex2_approx(q.x), ex2_approx(q.y)));
    return ffma2(make_float2(-0.5f, -0.5f), t, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}

// 8-byte (4 x 16-bit) store of one pixel's channel group into a padded NHWC buffer, replicating into the 1-pixel border
// when the pixel lies on the image edge (so that 3x3 taps of the consumer never need clamping).  Cold path: edge pixels only.
static __device__ __noinline__ void store_px_border8(uint8_t* base, int b, int Y, int X, int Ho, int Wo, int Hop, int Wop,
                                                 int ld, int c0, uint2 q) {
    uint8_t* centre = base + (((static_cast<size_t>(b) * Hop + Y + 1) * Wop + X + 1) * ld + c0) * 2;
    *reinterpret_cast<uint2*>(centre) = q;
    if (Y != 0 && Y != Ho - 1 && X != 0 && X != Wo - 1) return;      // interior pixel: done
    const int dy0 = (Y == 0) ? -1 : 0, dy1 = (Y == Ho - 1) ? 1 : 0;
    const int dx0 = (X == 0) ? -1 : 0, dx1 = (X == Wo - 1) ? 1 : 0;
    const ptrdiff_t rowp = static_cast<ptrdiff_t>(Wop) * ld * 2, colp = static_cast<ptrdiff_t>(ld) * 2;
    for (int dy = dy0; dy <= dy1; ++dy)
        for (int dx = dx0; dx <= dx1; ++dx)
            if (dy != 0 || dx != 0) *reinterpret_cast<uint2*>(centre + dy * rowp + dx * colp) = q;
}

// LayerNorm fold, consumer side: rstd of the 8 rows a lane serves after the transpose (row 4*i + sub of the warp's 32).  The
// loads of tile i+1 are issued while tile i is drained and consumed one iteration later (no latency on the critical path).
struct LnRows { float rs[8]; };
__device__ __forceinline__ void ln_load(const UmmaParams& p, int mt, int quarter, int lane, LnRows& r) {
    const int sub = lane >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long grow = static_cast<long>(mt) * TILE_M + quarter * 32 + 4 * i + sub;
        r.rs[i] = (grow < p.M) ? p.ln_rstd[grow] : 1.f;
    }
}

// EPI_RESID with the residual staged by TMA (umma2_kernel): per epilogue warp two 32 x 32 fp32 chunk buffers (128-byte rows,
// SWIZZLE_128B) and their mbarriers; `rc` counts the chunks this warp has consumed (buffer = rc & 1, phase = (rc >> 1) & 1).
struct ResidPipe {
    uint8_t* buf;                 // 2 x 4096 bytes, 1024-byte aligned
    uint64_t* bars;               // 2 mbarriers
    const CUtensorMap* map;       // fp32 [M, N], box {32, 32}
    uint32_t rc;
    int nbuf;                     // 1 or 2 chunk buffers
    int next_row0, next_col0;     // first row / first column (of this warp's column group) of the NEXT tile, next_row0 < 0: none
};

// DF < 0: the EPI_DEC variant (raw / ReLU copies, skip, UV, pixel shuffle) is decided at run time from the params;
// DF >= 0: compile-time bit mask (DF_RAW | DF_RELU | DF_SKIP | DF_UV | DF_SHUFFLE) -- a much smaller hot loop.
enum : int { DF_RAW = 1, DF_RELU = 2, DF_SKIP = 4, DF_UV = 8, DF_SHUFFLE = 16 };

// 16-byte flavour of the border-replicating store (8 x 16-bit channels of one pixel).  Cold path: edge pixels only.
static __device__ __noinline__ void store_px_border16(uint8_t* base, int b, int Y, int X, int Ho, int Wo, int Hop, int Wop,
                                                      int ld, int c0, uint4 q) {
    uint8_t* centre = base + (((static_cast<size_t>(b) * Hop + Y + 1) * Wop + X + 1) * ld + c0) * 2;
    const int dy0 = (Y == 0) ? -1 : 0, dy1 = (Y == Ho - 1) ? 1 : 0;
    const int dx0 = (X == 0) ? -1 : 0, dx1 = (X == Wo - 1) ? 1 : 0;
    const ptrdiff_t rowp = static_cast<ptrdiff_t>(Wop) * ld * 2, colp = static_cast<ptrdiff_t>(ld) * 2;
    for (int dy = dy0; dy <= dy1; ++dy)
        for (int dx = dx0; dx <= dx1; ++dx)
            if (dy != 0 || dx != 0) *reinterpret_cast<uint4*>(centre + dy * rowp + dx * colp) = q;
}

// EPI_DEC epilogue (decoder maps, padded NHWC): after the 32x32 transpose every lane owns EIGHT channels of one pixel
// (4 lanes per pixel row, 8 pixels per warp instruction, 4 passes per chunk) and moves 16 bytes per load/store -- half the
// per-byte instruction overhead of the 8-byte variant used for the row-major epilogues.
template <int BN, int COLS, int AMODE, bool BF16, int DF>
__device__ __forceinline__ void epilogue_dec16(const UmmaParams& p, int mt, int nt, uint32_t t_addr, float4* scr, int quarter, int lane,
                                               int col_begin) {
    using H = H16<BF16>;
    const bool has_raw = (DF < 0) ? (p.out0 != nullptr) : ((DF & DF_RAW) != 0);
    const bool has_relu = (DF < 0) ? (p.out1 != nullptr) : ((DF & DF_RELU) != 0);
    const bool has_skip = (DF < 0) ? (p.skip != nullptr) : ((DF & DF_SKIP) != 0);
    const bool has_uv = (DF < 0) ? (p.vec1 != nullptr) : ((DF & DF_UV) != 0);
    const bool shuffle = (DF < 0) ? (p.shuffle != 0) : ((DF & DF_SHUFFLE) != 0);
    const int sub = lane >> 2;        // which of the 8 rows of a pass this lane serves
    const int q8 = lane & 3;          // which group of 8 channels of the 32-column chunk
    int tile_b = 0, tile_y0 = 0, tile_x0 = 0;
    if (AMODE == AMODE_TILES) {
        const int per_img = p.tiles_x * p.tiles_y;
        tile_b = mt / per_img;
        const int r = mt % per_img;
        tile_y0 = (r / p.tiles_x) * TILE_PH;
        tile_x0 = (r % p.tiles_x) * TILE_PW;
    }
    bool ok[4];
    int rb[4], ry[4], rx[4], eflags[4];
    size_t roff[4];                   // byte offset of the centre output pixel (channel 0)
    const int sh = shuffle ? 2 : 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int rl = quarter * 32 + 8 * i + sub;
        if (AMODE == AMODE_ROWS) {
            const long grow = static_cast<long>(mt) * TILE_M + rl;
            ok[i] = grow < p.M;
            rb[i] = static_cast<int>(grow / p.T);
            const int t = static_cast<int>(grow % p.T);
            ry[i] = t / p.W; rx[i] = t % p.W;
        } else {
            rb[i] = tile_b;
            ry[i] = tile_y0 + rl / TILE_PW;
            rx[i] = tile_x0 + rl % TILE_PW;
            ok[i] = (ry[i] < p.H) && (rx[i] < p.W);
        }
        roff[i] = ((static_cast<size_t>(rb[i]) * p.Hop + sh * ry[i] + 1) * p.Wop + sh * rx[i] + 1) * p.ldo * 2;
        eflags[i] = (ry[i] == 0 ? 1 : 0) | (ry[i] == p.H - 1 ? 2 : 0) | (rx[i] == 0 ? 4 : 0) | (rx[i] == p.W - 1 ? 8 : 0);
    }
    const float inv_wo = 1.0f / static_cast<float>(p.Wo), inv_ho = 1.0f / static_cast<float>(p.Ho);
    // Interior tiles (every pixel inside the image and none on its border: ~70 % of the tiles of a 148x148 map) take a
    // straight-line path: no per-row validity predicates, no border-replication bookkeeping.  ncu of the ConvTranspose launch
    // showed this epilogue spending 57 % of its instructions on address arithmetic, predicates and branches (IMAD / ISETP / BRA /
    // LOP3 / SEL) around 21 % of useful work (FADD, F2FP, LDS / STS, STG).  The phase of a pixel-shuffle chunk comes from a shift
    // when C_out is a power of two (every MoGe config) instead of an integer division per chunk.
    bool interior;
    if (AMODE == AMODE_TILES) interior = tile_y0 > 0 && tile_x0 > 0 && tile_y0 + TILE_PH < p.H && tile_x0 + TILE_PW < p.W;
    else interior = false;
    const int ldo_shift = ((p.ldo & (p.ldo - 1)) == 0) ? (31 - __clz(p.ldo)) : -1;
    auto chunk = [&](int c, auto fast_tag) {
        constexpr bool FAST = decltype(fast_tag)::value;
        float v[32];
        tmem_ld32(t_addr + c, v);
        tc_wait_ld();
#pragma unroll
        for (int q = 0; q < 8; ++q) scr[lane * 8 + (q ^ (lane & 7))] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        __syncwarp();
        const int col = nt * BN + col_begin + c;
        int co = col + 8 * q8, qd = 0, qmask = 15;
        size_t qoff = 0;
        if (shuffle) {
            qd = (ldo_shift >= 0) ? (col >> ldo_shift) : (col / p.ldo);      // ldo == C_out; a 32-column chunk never straddles a phase
            co -= qd * p.ldo;
            qoff = (static_cast<size_t>(qd >> 1) * p.Wop + (qd & 1)) * p.ldo * 2;
            qmask = ((qd >> 1) ? 2 : 1) | ((qd & 1) ? 8 : 4);
        }
        qoff += static_cast<size_t>(co) * 2;
        const float4 b0 = *reinterpret_cast<const float4*>(p.bias + co), b1 = *reinterpret_cast<const float4*>(p.bias + co + 4);
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, h0 = g0, h1 = g0;
        if (has_uv) {
            g0 = *reinterpret_cast<const float4*>(p.vec1 + co); g1 = *reinterpret_cast<const float4*>(p.vec1 + co + 4);
            h0 = *reinterpret_cast<const float4*>(p.vec2 + co); h1 = *reinterpret_cast<const float4*>(p.vec2 + co + 4);
        }
        // phase 1: all global reads of the chunk back to back
        uint4 sk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            sk[i] = make_uint4(0u, 0u, 0u, 0u);
            if (has_skip && (FAST || ok[i])) sk[i] = *reinterpret_cast<const uint4*>(static_cast<const uint8_t*>(p.skip) + roff[i] + qoff);
        }
        // phase 2: math + 16-byte stores
        uint4 pk_raw[4], pk_relu[4];
        unsigned edge_rows = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rl = 8 * i + sub;
            float4 a0 = scr[rl * 8 + ((2 * q8) ^ (rl & 7))];
            float4 a1 = scr[rl * 8 + ((2 * q8 + 1) ^ (rl & 7))];
            if (!FAST && !ok[i]) continue;
            a0.x += b0.x; a0.y += b0.y; a0.z += b0.z; a0.w += b0.w;
            a1.x += b1.x; a1.y += b1.y; a1.z += b1.z; a1.w += b1.w;
            if (has_skip) {
                const float2 s0 = H::unpack(sk[i].x), s1 = H::unpack(sk[i].y), s2 = H::unpack(sk[i].z), s3 = H::unpack(sk[i].w);
                a0.x += s0.x; a0.y += s0.y; a0.z += s1.x; a0.w += s1.y;
                a1.x += s2.x; a1.y += s2.y; a1.z += s3.x; a1.w += s3.y;
            }
            if (has_uv) {
                const int X = sh * rx[i] + (qd & 1), Y = sh * ry[i] + (qd >> 1);
                const float uu = p.su * ((2 * X + 1) * inv_wo - 1.0f);
                const float vv = p.sv * ((2 * Y + 1) * inv_ho - 1.0f);
                a0.x += g0.x * uu + h0.x * vv; a0.y += g0.y * uu + h0.y * vv; a0.z += g0.z * uu + h0.z * vv; a0.w += g0.w * uu + h0.w * vv;
                a1.x += g1.x * uu + h1.x * vv; a1.y += g1.y * uu + h1.y * vv; a1.z += g1.z * uu + h1.z * vv; a1.w += g1.w * uu + h1.w * vv;
            }
            if (!FAST && (eflags[i] & qmask) != 0) edge_rows |= 1u << i;
            if (has_raw) {
                pk_raw[i] = make_uint4(H::pack(a0.x, a0.y), H::pack(a0.z, a0.w), H::pack(a1.x, a1.y), H::pack(a1.z, a1.w));
                *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out0) + roff[i] + qoff) = pk_raw[i];
            }
            if (has_relu) {
                pk_relu[i] = make_uint4(H::pack(fmaxf(a0.x, 0.f), fmaxf(a0.y, 0.f)), H::pack(fmaxf(a0.z, 0.f), fmaxf(a0.w, 0.f)),
                                        H::pack(fmaxf(a1.x, 0.f), fmaxf(a1.y, 0.f)), H::pack(fmaxf(a1.z, 0.f), fmaxf(a1.w, 0.f)));
                *reinterpret_cast<uint4*>(static_cast<uint8_t*>(p.out1) + roff[i] + qoff) = pk_relu[i];
            }
        }
        if (!FAST && edge_rows != 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (edge_rows >> i & 1) {
                    const int X = sh * rx[i] + (qd & 1), Y = sh * ry[i] + (qd >> 1);
                    if (has_raw) store_px_border16(static_cast<uint8_t*>(p.out0), rb[i], Y, X, p.Ho, p.Wo, p.Hop, p.Wop, p.ldo, co, pk_raw[i]);
                    if (has_relu) store_px_border16(static_cast<uint8_t*>(p.out1), rb[i], Y, X, p.Ho, p.Wo, p.Hop, p.Wop, p.ldo, co, pk_relu[i]);
                }
        }
        __syncwarp();
    };
    if (interior) {
#pragma unroll 1
        for (int c = 0; c < COLS; c += 32) chunk(c, std::true_type{});
    } else {
#pragma unroll 1
        for (int c = 0; c < COLS; c += 32) chunk(c, std::false_type{});
    }
}


// Epilogue of ONE accumulator tile for one epilogue warp (TMEM lane quarter `quarter`, columns [col_begin, col_begin+COLS)).
// TMEM hands each thread one accumulator ROW; global memory wants warps on contiguous COLUMNS.  Every 32x32 chunk is
// therefore transposed through a per-warp swizzled smem tile: afterwards 8 lanes x float4 cover the 32 columns of one
// row and each warp instruction touches 4 rows (4 x 128 B), fully coalesced.

template <int BN, int COLS, int AMODE, int EPI, bool BF16, int DF = -1>
__device__ __forceinline__ void epilogue_tile(const UmmaParams& p, int mt, int nt, uint32_t t_addr, float4* scr, int quarter,
                                              int lane, int col_begin, bool ln_rows_on = false, LnRows lnr = LnRows{},
                                              ResidPipe* rp = nullptr) {
    using H = H16<BF16>;
    const bool has_raw = (DF < 0) ? (p.out0 != nullptr) : ((DF & DF_RAW) != 0);
    const bool has_relu = (DF < 0) ? (p.out1 != nullptr) : ((DF & DF_RELU) != 0);
    const bool has_skip = (DF < 0) ? (p.skip != nullptr) : ((DF & DF_SKIP) != 0);
    const bool has_uv = (DF < 0) ? (p.vec1 != nullptr) : ((DF & DF_UV) != 0);
    const bool shuffle = (DF < 0) ? (p.shuffle != 0) : ((DF & DF_SHUFFLE) != 0);
    if (EPI == EPI_DEC) {
        epilogue_dec16<BN, COLS, AMODE, BF16, DF>(p, mt, nt, t_addr, scr, quarter, lane, col_begin);
        return;
    }
    const int row = quarter * 32 + lane;
    const int sub = lane >> 3;        // which of the 4 rows of a pass this lane serves
    const int q4 = lane & 7;          // which float4 (4 columns) of the 32-column chunk
    if (EPI == EPI_HEADOUT) {
        // lane = pixel: 16 B (or 4 B) per lane, consecutive lanes = consecutive pixels of a tile row
        const int per_img = p.tiles_x * p.tiles_y;
        const int b = mt / per_img;
        const int r = mt % per_img;
        const int py = (r / p.tiles_x) * TILE_PH + row / TILE_PW;
        const int px = (r % p.tiles_x) * TILE_PW + row % TILE_PW;
        const bool valid = (py < p.H) && (px < p.W);
        float v[16];
        tmem_ld16(t_addr, v);
        tc_wait_ld();
        if (valid) {
            // accumulator columns: (phase, component); phase (qy,qx) -> output pixel (2*py+qy, 2*px+qx)
#pragma unroll
            for (int ph = 0; ph < 4; ++ph) {
                const int Y = 2 * py + (ph >> 1), X = 2 * px + (ph & 1);
                float o[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) o[k] = (k < p.ncomp) ? v[ph * p.ncomp + k] + p.bias[k] : 0.f;
                if (p.vec1 != nullptr) {
                    const uint4* src = reinterpret_cast<const uint4*>(
                        static_cast<const uint8_t*>(p.skip) + ((static_cast<size_t>(b) * p.Hop + Y + 1) * p.Wop + X + 1) * 64);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 u = src[q];
                        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = H::unpack(w[e]);
                            const int c = q * 8 + e * 2;
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                if (k < p.ncomp) o[k] += p.vec1[k * 32 + c] * f.x + p.vec1[k * 32 + c + 1] * f.y;
                        }
                    }
                }
                const size_t pix = (static_cast<size_t>(b) * p.Ho + Y) * p.Wo + X;
                if (p.accum) {       // the neck's folded contribution is already there (EPI_NECKOUT)
                    if (p.ncomp == 1) o[0] += static_cast<const float*>(p.out0)[pix];
                    else { const float4 t = static_cast<const float4*>(p.out0)[pix]; o[0] += t.x; o[1] += t.y; o[2] += t.z; }
                }
                if (p.ncomp == 1) static_cast<float*>(p.out0)[pix] = o[0];
                else static_cast<float4*>(p.out0)[pix] = make_float4(o[0], o[1], o[2], 0.f);
            }
        }
        __syncwarp();
    } else if (EPI == EPI_NECKOUT) {
        // lane = low-resolution pixel; 32 accumulator columns = (phase, component)
        const int per_img = p.tiles_x * p.tiles_y;
        const int b = mt / per_img;
        const int r = mt % per_img;
        const int py = (r / p.tiles_x) * TILE_PH + row / TILE_PW;
        const int px = (r % p.tiles_x) * TILE_PW + row % TILE_PW;
        const bool valid = (py < p.H) && (px < p.W);
        float v[32];
        tmem_ld32(t_addr, v);
        tc_wait_ld();
        if (valid) {
            float b8[8], gu[8], gv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { b8[k] = p.bias[k]; gu[k] = p.vec1[k]; gv[k] = p.vec2[k]; }
            const float inv_wo = 1.0f / static_cast<float>(p.Wo), inv_ho = 1.0f / static_cast<float>(p.Ho);
#pragma unroll
            for (int ph = 0; ph < 4; ++ph) {
                const int Y = 2 * py + (ph >> 1), X = 2 * px + (ph & 1);
                const float uu = p.su * ((2 * X + 1) * inv_wo - 1.0f);
                const float vv = p.sv * ((2 * Y + 1) * inv_ho - 1.0f);
                float o[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = v[ph * 8 + k] + b8[k] + gu[k] * uu + gv[k] * vv;
                const size_t pix = (static_cast<size_t>(b) * p.Ho + Y) * p.Wo + X;
                if (p.out0) static_cast<float4*>(p.out0)[pix] = make_float4(o[0], o[1], o[2], 0.f);
                if (p.out1) static_cast<float4*>(p.out1)[pix] = make_float4(o[3], o[4], o[5], 0.f);
                if (p.out2) static_cast<float*>(p.out2)[pix] = o[6];
            }
        }
        __syncwarp();
    } else {
        // ---- coordinates of the 8 rows this lane serves after the transpose (row = 4*i + sub of the warp's 32)
        int tile_b = 0, tile_y0 = 0, tile_x0 = 0;
        if (AMODE == AMODE_TILES) {
            const int per_img = p.tiles_x * p.tiles_y;
            tile_b = mt / per_img;
            const int r = mt % per_img;
            tile_y0 = (r / p.tiles_x) * TILE_PH;
            tile_x0 = (r % p.tiles_x) * TILE_PW;
        }
        bool ok[8];
        int rb[8], ry[8], rx[8];
        size_t roff[8];      // element offset of the row (ROWS epilogues) / BYTE offset of the centre output pixel (EPI_DEC)
        int eflags[8];       // EPI_DEC: bit0 y==0, bit1 y==H-1, bit2 x==0, bit3 x==W-1 (source grid)
        int xrow[8];         // ROWS epilogues: row index in the output / statistics arrays
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rl = quarter * 32 + 4 * i + sub;
            rb[i] = 0; ry[i] = 0; rx[i] = 0; roff[i] = 0; eflags[i] = 0; xrow[i] = 0;
            if (AMODE == AMODE_ROWS) {
                const long grow = static_cast<long>(mt) * TILE_M + rl;
                ok[i] = grow < p.M;
                xrow[i] = static_cast<int>(grow);
                roff[i] = static_cast<size_t>(grow) * p.ldo;
                if (EPI == EPI_DEC || EPI == EPI_PATCH) {
                    rb[i] = static_cast<int>(grow / p.T);
                    const int t = static_cast<int>(grow % p.T);
                    ry[i] = t / p.W; rx[i] = t % p.W;
                    if (EPI == EPI_PATCH) {
                        xrow[i] = rb[i] * (p.T + 1) + 1 + t;
                        roff[i] = static_cast<size_t>(xrow[i]) * p.ldo;
                    }
                }
            } else {
                rb[i] = tile_b;
                ry[i] = tile_y0 + rl / TILE_PW;
                rx[i] = tile_x0 + rl % TILE_PW;
                ok[i] = (ry[i] < p.H) && (rx[i] < p.W);
            }
            if (EPI == EPI_DEC) {
                const int sh = shuffle ? 2 : 1;
                roff[i] = ((static_cast<size_t>(rb[i]) * p.Hop + sh * ry[i] + 1) * p.Wop + sh * rx[i] + 1) * p.ldo * 2;
                eflags[i] = (ry[i] == 0 ? 1 : 0) | (ry[i] == p.H - 1 ? 2 : 0) | (rx[i] == 0 ? 4 : 0) | (rx[i] == p.W - 1 ? 8 : 0);
            }
        }
        const float inv_wo = 1.0f / static_cast<float>(p.Wo), inv_ho = 1.0f / static_cast<float>(p.Ho);
        // LayerNorm fold, consumer side: rstd and -rstd*mean of this lane's 8 rows from the producers' partial sums (the 8
        // lanes that share a row split the partials, fixed order -> deterministic)
        constexpr bool LN_CONS = (EPI == EPI_STORE16 || EPI == EPI_GELU16) && AMODE == AMODE_ROWS;
        constexpr bool LN_PROD = (EPI == EPI_RESID || EPI == EPI_PATCH) && AMODE == AMODE_ROWS;
        const bool ln_on = LN_CONS && ln_rows_on;
        const bool x16_on = LN_PROD && p.x16 != nullptr;
        float ln_rs[8], st1[8], st2[8];        // st1/st2 (producers): partial sums of x and x^2 of this lane's 8 rows
#pragma unroll
        for (int i = 0; i < 8; ++i) { ln_rs[i] = ln_on ? lnr.rs[i] : 1.f; st1[i] = 0.f; st2[i] = 0.f; }
        // (Prefetching the fp32 residual of chunk c + 1 before chunk c is drained -- two chunks of loads in flight per warp -- was
        // measured on one box, A/B: proj 3.68 -> 4.14 ms, fc2 7.24 -> 7.33 ms per step, i.e. SLOWER; the extra 32 registers and the
        // deeper load queue cost more than the latency they hide.  Kept load-then-use.)
        // Full row tiles (all but the last of a GEMM) take a path without the per-row validity predicates.
        const bool full_tile = (AMODE == AMODE_ROWS) && (static_cast<long>(mt) + 1) * TILE_M <= static_cast<long>(p.M);
        auto chunk = [&](int c, auto full_tag) {
            constexpr bool FULL = decltype(full_tag)::value;
            const int col = nt * BN + col_begin + c;       // first global output column of this chunk
            int co = col + 4 * q4;                         // this lane's 4 columns
            float4 pre[8];
            // (EPI_RESID -- proj / fc2 -- stalls on the latency of the fp32 residual reads: ncu long_scoreboard 7.6 warps per issue
            // cycle at 48 % of the DRAM peak.  Issuing those reads HERE, ahead of the TMEM load and the transpose, measured SLOWER in a
            // same-box A/B -- proj 3.41 -> 3.65 ms per step -- like the two-chunk prefetch before it: the longer live ranges cost
            // registers / spills in an epilogue that sits at the 168-register cap.  Kept load-then-use.)
            float v[32];
            tmem_ld32(t_addr + c, v);
            tc_wait_ld();
#pragma unroll
            for (int q = 0; q < 8; ++q)
                scr[lane * 8 + (q ^ (lane & 7))] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            __syncwarp();
            int qd = 0;
            size_t qoff = 0;                               // EPI_DEC: byte offset of this chunk relative to the centre pixel
            int qmask = 15;                                // which source-grid edges replicate for this chunk
            if (EPI == EPI_DEC && shuffle) {
                qd = col / p.ldo;                          // ldo == C_out; a 32-column chunk never straddles qd
                co -= qd * p.ldo;
                qoff = (static_cast<size_t>(qd >> 1) * p.Wop + (qd & 1)) * p.ldo * 2;
                qmask = ((qd >> 1) ? 2 : 1) | ((qd & 1) ? 8 : 4);
            }
            if (EPI == EPI_DEC) qoff += static_cast<size_t>(co) * 2;
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), g4 = bias4, h4 = bias4;
            if (EPI != EPI_PATCH) bias4 = *reinterpret_cast<const float4*>(p.bias + co);
            if (EPI == EPI_RESID) g4 = *reinterpret_cast<const float4*>(p.vec1 + co);
            if (EPI == EPI_DEC && has_uv) {
                g4 = *reinterpret_cast<const float4*>(p.vec1 + co);
                h4 = *reinterpret_cast<const float4*>(p.vec2 + co);
            }
            // ---- phase 1: issue every global READ of this chunk (residual / pos table / skip) back to back, so
            //      their L2 latencies overlap instead of serialising behind the stores of the previous row
            int sY[8], sX[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                pre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                sY[i] = ry[i]; sX[i] = rx[i];
                if (EPI == EPI_DEC && shuffle) { sY[i] = 2 * ry[i] + (qd >> 1); sX[i] = 2 * rx[i] + (qd & 1); }
                if (!FULL && !ok[i]) continue;
                if (EPI == EPI_RESID) {
                    if (rp == nullptr) pre[i] = *reinterpret_cast<const float4*>(static_cast<const float*>(p.out0) + roff[i] + co);
                } else if (EPI == EPI_PATCH) {
                    const int t = ry[i] * p.W + rx[i];
                    pre[i] = *reinterpret_cast<const float4*>(p.vec1 + static_cast<size_t>(t) * p.ldo + co);
                } else if (EPI == EPI_DEC && has_skip) {
                    const uint2 u = *reinterpret_cast<const uint2*>(static_cast<const uint8_t*>(p.skip) + roff[i] + qoff);
                    const float2 f0 = H::unpack(u.x), f1 = H::unpack(u.y);
                    pre[i] = make_float4(f0.x, f0.y, f1.x, f1.y);
                }
            }
            if (EPI == EPI_RESID && rp != nullptr) {
                // residual chunk from the TMA-staged buffer: row rl = 4 i + sub is one swizzled 128-byte line, this lane's 4 columns
                // are its 16-byte chunk q4 ^ (rl & 7)
                const uint32_t b = (rp->nbuf == 2) ? (rp->rc & 1u) : 0u;
                mbar_wait(&rp->bars[b], ((rp->nbuf == 2) ? (rp->rc >> 1) : rp->rc) & 1u);
                const uint8_t* rb_ = rp->buf + b * 4096;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rl = 4 * i + sub;
                    pre[i] = *reinterpret_cast<const float4*>(rb_ + rl * 128 + ((q4 ^ (rl & 7)) << 4));
                }
                fence_proxy_async_smem();           // these generic-proxy reads are ordered before the TMA write that refills the buffer
                __syncwarp();
                // refill the buffer `nbuf` chunks ahead: a later chunk of this tile, or the first chunk(s) of the next tile (they land
                // under its main loop)
                const int ahead = 32 * rp->nbuf;
                int ncol = -1, nrow = 0;
                if (c + ahead < COLS) { ncol = nt * BN + col_begin + c + ahead; nrow = mt * TILE_M + quarter * 32; }
                else if (rp->next_row0 >= 0) { ncol = rp->next_col0 + (c + ahead - COLS); nrow = rp->next_row0; }
                if (ncol >= 0 && elect_one()) {
                    mbar_arrive_expect_tx(&rp->bars[b], 4096);
                    tma_load_2d(rp->buf + b * 4096, rp->map, &rp->bars[b], ncol, nrow);
                }
                __syncwarp();
                rp->rc++;
            }
            // ---- phase 2: math + stores
            uint2 pk_raw[8], pk_relu[8];
            unsigned edge_rows = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rl = 4 * i + sub;
                float4 a = scr[rl * 8 + (q4 ^ (rl & 7))];
                if (!FULL && !ok[i]) continue;
                if (EPI == EPI_STORE16 || EPI == EPI_GELU16) {
                    // (packed FADD2 for the plain bias add measured ~5 % SLOWER on these GEMMs than four scalar FADDs -- three boxes
                    // each way -- although the packed GELU polynomial and the packed attention softmax are wins; kept scalar)
                    float2 a01, a23;
                    if (ln_on) {      // LN(x) W^T + b = rstd * (x16 W''^T) + b'   (mean removal lives in the centred weight W'')
                        a01 = make_float2(fmaf(ln_rs[i], a.x, bias4.x), fmaf(ln_rs[i], a.y, bias4.y));
                        a23 = make_float2(fmaf(ln_rs[i], a.z, bias4.z), fmaf(ln_rs[i], a.w, bias4.w));
                    } else {
                        a01 = make_float2(a.x + bias4.x, a.y + bias4.y);
                        a23 = make_float2(a.z + bias4.z, a.w + bias4.w);
                    }
                    if (EPI == EPI_GELU16) { a01 = gelu_erf2(a01); a23 = gelu_erf2(a23); }
                    uint2 pk;
                    pk.x = H::pack(a01.x, a01.y); pk.y = H::pack(a23.x, a23.y);
                    *reinterpret_cast<uint2*>(static_cast<typename H::T*>(p.out0) + roff[i] + co) = pk;
                } else if (EPI == EPI_RESID) {
                    float4 x = pre[i];
                    x.x += g4.x * (a.x + bias4.x); x.y += g4.y * (a.y + bias4.y);
                    x.z += g4.z * (a.z + bias4.z); x.w += g4.w * (a.w + bias4.w);
                    *reinterpret_cast<float4*>(static_cast<float*>(p.out0) + roff[i] + co) = x;
                    const float2 x01 = make_float2(x.x, x.y), x23 = make_float2(x.z, x.w);
                    if (x16_on) {
                        uint2 pk;
                        pk.x = H::pack(x01.x, x01.y); pk.y = H::pack(x23.x, x23.y);
                        *reinterpret_cast<uint2*>(static_cast<typename H::T*>(p.x16) + roff[i] + co) = pk;
                        const float2 f0 = H::unpack(pk.x), f1 = H::unpack(pk.y);
                        const float2 s1 = fadd2(f0, f1), s2 = ffma2(f0, f0, fmul2(f1, f1));
                        st1[i] += s1.x + s1.y;
                        st2[i] += s2.x + s2.y;
                    }
                } else if (EPI == EPI_PATCH) {
                    const float4 x = make_float4(a.x + pre[i].x, a.y + pre[i].y, a.z + pre[i].z, a.w + pre[i].w);
                    *reinterpret_cast<float4*>(static_cast<float*>(p.out0) + roff[i] + co) = x;
                    if (x16_on) {
                        uint2 pk;
                        pk.x = H::pack(x.x, x.y); pk.y = H::pack(x.z, x.w);
                        *reinterpret_cast<uint2*>(static_cast<typename H::T*>(p.x16) + roff[i] + co) = pk;
                        const float2 f0 = H::unpack(pk.x), f1 = H::unpack(pk.y);
                        const float2 s1 = fadd2(f0, f1), s2 = ffma2(f0, f0, fmul2(f1, f1));
                        st1[i] += s1.x + s1.y;
                        st2[i] += s2.x + s2.y;
                    }
                } else if (EPI == EPI_DEC) {
                    a.x += bias4.x + pre[i].x; a.y += bias4.y + pre[i].y; a.z += bias4.z + pre[i].z; a.w += bias4.w + pre[i].w;
                    if (has_uv) {
                        const float uu = p.su * ((2 * sX[i] + 1) * inv_wo - 1.0f);
                        const float vv = p.sv * ((2 * sY[i] + 1) * inv_ho - 1.0f);
                        a.x += g4.x * uu + h4.x * vv; a.y += g4.y * uu + h4.y * vv;
                        a.z += g4.z * uu + h4.z * vv; a.w += g4.w * uu + h4.w * vv;
                    }
                    // centre pixel: straight-line predicated stores; the (rare) replicated-border copies are done after the loop
                    if ((eflags[i] & qmask) != 0) edge_rows |= 1u << i;
                    if (has_raw) {
                        pk_raw[i].x = H::pack(a.x, a.y); pk_raw[i].y = H::pack(a.z, a.w);
                        *reinterpret_cast<uint2*>(static_cast<uint8_t*>(p.out0) + roff[i] + qoff) = pk_raw[i];
                    }
                    if (has_relu) {
                        pk_relu[i].x = H::pack(fmaxf(a.x, 0.f), fmaxf(a.y, 0.f)); pk_relu[i].y = H::pack(fmaxf(a.z, 0.f), fmaxf(a.w, 0.f));
                        *reinterpret_cast<uint2*>(static_cast<uint8_t*>(p.out1) + roff[i] + qoff) = pk_relu[i];
                    }
                }
            }
            if (EPI == EPI_DEC && edge_rows != 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (edge_rows >> i & 1) {
                        if (has_raw) store_px_border8(static_cast<uint8_t*>(p.out0), rb[i], sY[i], sX[i], p.Ho, p.Wo, p.Hop, p.Wop, p.ldo, co, pk_raw[i]);
                        if (has_relu) store_px_border8(static_cast<uint8_t*>(p.out1), rb[i], sY[i], sX[i], p.Ho, p.Wo, p.Hop, p.Wop, p.ldo, co, pk_relu[i]);
                    }
            }
            __syncwarp();
        };
        if (full_tile) {
#pragma unroll 1
            for (int c = 0; c < COLS; c += 32) chunk(c, std::true_type{});
        } else {
#pragma unroll 1
            for (int c = 0; c < COLS; c += 32) chunk(c, std::false_type{});
        }
        // LayerNorm fold, producer side: partial (sum, sum of squares) of this warp's column group for each of its rows
        if (x16_on) {
            const int part = (nt * BN + col_begin) / COLS;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float a = st1[i], b = st2[i];
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
                if (q4 == 0 && ok[i]) p.stats_out[static_cast<size_t>(xrow[i]) * p.stats_ld + part] = make_float2(a, b);
            }
        }
    }
}

template <int BN, int AMODE, int EPI, bool BF16, int DF = -1>
__global__ void __launch_bounds__(UmmaCfg<BN>::kThreads, 1)
umma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapAux,
            const __grid_constant__ CUtensorMap mapB, const UmmaParams p) {
    pdl_launch_dependents();      // (the wait sits after the barrier / TMEM set-up below: that prologue overlaps the previous kernel's tail)
    using Cfg = UmmaCfg<BN>;
    using H = H16<BF16>;
    constexpr int S = Cfg::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * Cfg::kStageBytes);
    uint64_t* empty = full + S;
    uint64_t* tfull = empty + S;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    float* scratch_base = reinterpret_cast<float*>(smem + S * Cfg::kStageBytes + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.num_m_tiles * p.num_n_tiles;
    const int kb_taps = p.ntaps * p.kb_main;
    const int kb_total = kb_taps + p.kb_aux;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        if (p.kb_aux) tma_prefetch_desc(&mapAux);
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
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
        // ================================================================== TMA producer
        // (this warp and the MMA warp stay CONVERGED and one elected lane issues: under `if (lane == 0)` the compiler has
        // to assume an arbitrary active mask and wraps every TMA / tcgen05 instruction in an elect-and-retry loop)
        {
            int s = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
                int b = 0, x0 = 0, y0 = 0;
                if (AMODE == AMODE_TILES) {
                    const int per_img = p.tiles_x * p.tiles_y;
                    b = mt / per_img;
                    const int r = mt % per_img;
                    y0 = (r / p.tiles_x) * TILE_PH;
                    x0 = (r % p.tiles_x) * TILE_PW;
                }
                for (int i = 0; i < kb_total; ++i) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* sa = smem + s * Cfg::kStageBytes;
                    uint8_t* sb = sa + TILE_M * 128;
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&full[s], Cfg::kStageBytes);
                        if (AMODE == AMODE_ROWS) {
                            tma_load_2d(sa, &mapA, &full[s], i * TILE_K, mt * TILE_M);
                        } else if (i < kb_taps) {
                            const int tap = i / p.kb_main, c = i % p.kb_main;
                            const int tx = (p.ntaps == 9) ? tap % 3 : 1, ty = (p.ntaps == 9) ? tap / 3 : 1;
                            tma_load_4d(sa, &mapA, &full[s], c * TILE_K, x0 + tx, y0 + ty, b);
                        } else {
                            tma_load_4d(sa, &mapAux, &full[s], (i - kb_taps) * TILE_K, x0 + 1, y0 + 1, b);
                        }
                        tma_load_2d(sb, &mapB, &full[s], i * TILE_K, nt * BN);
                    }
                    __syncwarp();
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================================== MMA issuer
        {
            constexpr uint32_t idesc = make_idesc(TILE_M, BN, BF16 ? 1u : 0u);
            int s = 0; uint32_t ph = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(&tempty[acc], aph ^ 1);
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
                        for (int k = 0; k < TILE_K / 16; ++k)
                            umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0);
                        umma_commit(&empty[s]);
                        if (i == kb_total - 1) umma_commit(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++s == S) { s = 0; ph ^= 1; }
                }
            }
        }
    } else {
        // ================================================================== epilogue
        // TMEM hands each thread one accumulator ROW; global memory wants warps on contiguous COLUMNS.  Every 32x32
        // chunk is therefore transposed through a per-warp swizzled smem tile: afterwards 8 lanes x float4 cover the
        // 32 columns of one row and each warp instruction touches 4 rows (4 x 128 B), fully coalesced.
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int col_begin = (Cfg::kEpiWarps == 8) ? (ew >> 2) * (BN / 2) : 0;
        float4* scr = reinterpret_cast<float4*>(scratch_base + ew * 1024);
        int it = 0;
        LnRows lnr{}, lnn{};
        const bool ln_on = AMODE == AMODE_ROWS && (EPI == EPI_STORE16 || EPI == EPI_GELU16) && p.ln_rstd != nullptr;
        if (ln_on && static_cast<int>(blockIdx.x) < total_tiles) ln_load(p, static_cast<int>(blockIdx.x) / p.num_n_tiles, quarter, lane, lnn);
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int mt = tile / p.num_n_tiles, nt = tile % p.num_n_tiles;
            const int acc = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            if (ln_on) {        // rstd of this tile's rows (loaded one iteration ago); then issue the next tile's loads
                lnr = lnn;
                const int nxt = tile + static_cast<int>(gridDim.x);
                if (nxt < total_tiles) ln_load(p, nxt / p.num_n_tiles, quarter, lane, lnn);
            }
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + col_begin;
            epilogue_tile<BN, Cfg::kColsPerWarp, AMODE, EPI, BF16, DF>(p, mt, nt, t_addr, scr, quarter, lane, col_begin, ln_on, lnr);
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
