// Multi-head self-attention of the DINOv2 blocks (reference: dinov2/layers/attention.py:70-81, SDPA with
// scale hd^-0.5, no mask) as one tcgen05 kernel: S = Q K^T and O~ = P V on the tensor cores with TMEM accumulators,
// single-pass online softmax in registers (the whole 128-wide score row lives in registers; FMNMX3 + MUFU.EX2),
// O accumulated in TMEM by the MMA itself and rescaled lazily (only when the running max grows by more than 2^8),
// two 128-row query tiles per CTA ping-ponging on the tensor pipe.  The probabilities P never touch shared memory: the
// softmax threads store them (16-bit, packed) into tensor memory and the P V MMA reads its A operand from there -- with
// P in smem the kernel was shared-memory-bandwidth-bound (P write + P read were half of all smem traffic and the MMAs
// ran at 2.3x their nominal duration waiting for operands).
//
// Layout: qkv is the QKV-GEMM output [B, N, 3*D] (16-bit); Q/K/V tiles of head h are the column windows
// [h*64, D+h*64, 2D+h*64) fetched by TMA straight from that buffer (no head-major repack): Q and K tiles are
// K-major UMMA operands, the V tile ([kv][hd], hd contiguous) is consumed as an MN-major B operand.
// Output: out[b*N + q, h*64 + d] (16-bit), the A operand of the projection GEMM.
#include "common.cuh"
#include "host_api.h"

namespace mg {

constexpr int ATT_HD = 64;
constexpr int ATT_BQ = 128;       // query rows per softmax warpgroup
constexpr int ATT_BKV = 128;      // keys per tile
constexpr int ATT_KV_STAGES = 5;
constexpr int ATT_TILE_BYTES = 128 * 128;   // [128 rows][64 x 16-bit]
constexpr int ATT_THREADS = 128 + 256;   // warpgroup 0: TMA warp, MMA warp, 2 idle; warpgroups 1,2: softmax
// smem: Q0,Q1 | K[stages] | V[stages] | barriers
constexpr int ATT_SMEM = (2 + 2 * ATT_KV_STAGES) * ATT_TILE_BYTES + 1024 + 256;
static_assert(ATT_SMEM <= kMaxDynSmem, "attention_kernel: Q + K/V ring exceed the shared memory of one CTA");
static_assert((1 + 4 * ATT_KV_STAGES + 8) * 8 + 4 <= 256, "attention_kernel: barrier block overflows its 256 bytes");

#ifdef MG_ATT_DEBUG
// wait-time attribution (debug builds only; tools/att_debug.py), summed over CTAs, warp 4 lane 0 / warp 1 lane 0:
// [0] softmax waits S  [1] softmax waits PV(j-1)  [2] softmax loop total  [3] MMA waits P  [4] MMA waits K/V
// [5] CTA lifetime  [6] prologue (start -> first S)  [7] epilogue (O normalise + store)
__device__ unsigned long long mg_att_dbg[8];
#define ATT_T0() const long long _t0 = clock64()
#define ATT_ACC(var) var += clock64() - _t0
#else
#define ATT_T0()
#define ATT_ACC(var)
#endif

struct AttnParams {
    void* out;        // [B*N, D] 16-bit
    int B, N, D, heads;
    float scale_log2; // hd^-0.5 * log2(e)
};

template <bool BF16>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_kernel(const __grid_constant__ CUtensorMap mapQKV, const AttnParams p) {
    using H = H16<BF16>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int ST = ATT_KV_STAGES;
    uint8_t* sQ = smem;                                        // 2 tiles
    uint8_t* sK = sQ + 2 * ATT_TILE_BYTES;
    uint8_t* sV = sK + ST * ATT_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * ATT_TILE_BYTES);
    uint64_t* q_full = bars;                 // 1
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = k_full + ST;
    uint64_t* v_full = k_empty + ST;
    uint64_t* v_empty = v_full + ST;
    uint64_t* s_full = v_empty + ST;         // 2
    uint64_t* p_full = s_full + 2;           // 2
    uint64_t* o_full = p_full + 2;           // 2 (one per group; completes once per kv tile)
    uint64_t* s_free = o_full + 2;           // 2: the softmax threads hold the whole score tile in registers -> S may be overwritten
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef MG_ATT_DEBUG
    long long dbg_a = 0, dbg_b = 0, dbg_c = 0, dbg_d = 0;
    const long long dbg_start = clock64();
#endif
    const int q0 = blockIdx.x * 2 * ATT_BQ;
    const int h = blockIdx.y, b = blockIdx.z;
    const int nkv = (p.N + ATT_BKV - 1) / ATT_BKV;
    const int ng = (q0 + ATT_BQ < p.N) ? 2 : 1;       // the second query tile of the last CTA may be entirely past N: skipped

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapQKV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_KV_STAGES; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
        }
        for (int g = 0; g < 2; ++g) {
            mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4);
            mbar_init(&o_full[g], 1); mbar_init(&s_free[g], 4);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM columns: S0 [0,128)  S1 [128,256)  O0 [256,320)  O1 [320,384)  P0 [384,448)  P1 [448,512)  (P: 2 x 16-bit per column)

    // register re-balancing between the control warpgroup and the two softmax warpgroups (row of 128 scores in registers)
    if (warp == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        // The warp stays converged and ONE ELECTED lane issues (here and in the MMA warp): under `if (lane == 0)` the
        // compiler wraps every TMA / tcgen05 instruction in an elect-and-retry loop, ~9 dependent instructions per MMA.
        if (elect_one()) {
            mbar_arrive_expect_tx(q_full, 2 * ATT_TILE_BYTES);
            tma_load_3d(sQ, &mapQKV, q_full, h * ATT_HD, q0, b);
            tma_load_3d(sQ + ATT_TILE_BYTES, &mapQKV, q_full, h * ATT_HD, q0 + ATT_BQ, b);
        }
        __syncwarp();
        int s = 0; uint32_t ph = 0;
        for (int j = 0; j < nkv; ++j) {
            mbar_wait(&k_empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&k_full[s], ATT_TILE_BYTES);
                tma_load_3d(sK + s * ATT_TILE_BYTES, &mapQKV, &k_full[s], p.D + h * ATT_HD, j * ATT_BKV, b);
            }
            __syncwarp();
            mbar_wait(&v_empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(&v_full[s], ATT_TILE_BYTES);
                tma_load_3d(sV + s * ATT_TILE_BYTES, &mapQKV, &v_full[s], 2 * p.D + h * ATT_HD, j * ATT_BKV, b);
            }
            __syncwarp();
            if (++s == ATT_KV_STAGES) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        {
            constexpr uint32_t idesc_s = make_idesc(128, ATT_BKV, BF16 ? 1u : 0u, 0, 0);   // Q (K-major) x K (K-major)
            constexpr uint32_t idesc_o = make_idesc(128, ATT_HD, BF16 ? 1u : 0u, 0, 1);    // P (TMEM) x V (MN-major)
            // S_g = Q_g K_stage^T; `release` != 0: the K stage is free once these MMAs retire
            auto issue_s = [&](int g, int stage, bool release) {
                if (elect_one()) {
                    const uint64_t a = make_sdesc_sw128(smem_u32(sQ + g * ATT_TILE_BYTES));
                    const uint64_t bd = make_sdesc_sw128(smem_u32(sK + stage * ATT_TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < ATT_HD / 16; ++k) umma_f16(tmem + g * 128, a + 2 * k, bd + 2 * k, idesc_s, k != 0);
                    umma_commit(&s_full[g]);
                    if (release) umma_commit(&k_empty[stage]);
                }
                __syncwarp();
            };
            // O_g (+)= P_g V_stage
            auto issue_pv = [&](int g, int stage, bool first, bool release) {
                if (elect_one()) {
                    const uint32_t pa = tmem + 384 + g * 64;
                    const uint32_t va = smem_u32(sV + stage * ATT_TILE_BYTES);
                    const uint32_t d = tmem + 256 + g * 64;
#pragma unroll
                    for (int k = 0; k < ATT_BKV / 16; ++k) {
                        // A: P in tensor memory, 16 keys (= 8 columns of packed 16-bit pairs) per K-step
                        // B: V tile [kv][hd]: MN-major, 16 kv rows (= 2 groups of 8 x 128 B) per K-step
                        const uint64_t bd = make_sdesc_sw128(va + k * 16 * 128, /*lbo=*/ATT_TILE_BYTES, /*sbo=*/1024);
                        umma_f16_ts(d, pa + 8 * k, bd, idesc_o, !(first && k == 0));   // O accumulates across kv tiles in TMEM
                    }
                    umma_commit(&o_full[g]);
                    if (release) umma_commit(&v_empty[stage]);
                }
                __syncwarp();
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tc_fence_after();
            issue_s(0, 0, ng == 1);
            if (ng == 2) issue_s(1, 0, true);
            int sv = 0; uint32_t phv = 0;         // V stage of tile j
            int sk = 1 % ATT_KV_STAGES; uint32_t phk = (ATT_KV_STAGES == 1) ? 1u : 0u;   // K stage of tile j + 1
            for (int j = 0; j < nkv; ++j) {
                // S(j+1) of both groups first: it only needs the score registers of tile j to be loaded (s_free), not the
                // softmax of tile j to be finished -- the next scores are ready before the softmax threads ask for them
                if (j + 1 < nkv) {
                    { ATT_T0(); mbar_wait(&k_full[sk], phk); ATT_ACC(dbg_b); }
                    for (int g = 0; g < ng; ++g) {
                        { ATT_T0(); mbar_wait(&s_free[g], j & 1); ATT_ACC(dbg_a); }
                        tc_fence_after();
                        issue_s(g, sk, g == ng - 1);
                    }
                }
                { ATT_T0(); mbar_wait(&v_full[sv], phv); ATT_ACC(dbg_b); }
                for (int g = 0; g < ng; ++g) {
                    { ATT_T0(); mbar_wait(&p_full[g], j & 1); ATT_ACC(dbg_a); }
                    tc_fence_after();
                    issue_pv(g, sv, j == 0, g == ng - 1);
                }
                if (++sv == ATT_KV_STAGES) { sv = 0; phv ^= 1; }
                if (++sk == ATT_KV_STAGES) { sk = 0; phk ^= 1; }
            }
        }
    } else if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ========================================================== softmax / accumulate / store (one row per thread)
        const int g = (warp - 4) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int qrow = q0 + g * ATT_BQ + row;
        const uint32_t lane_sel = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t tS = tmem + lane_sel + g * 128;
        const uint32_t tO = tmem + lane_sel + 256 + g * 64;
        const uint32_t tP = tmem + lane_sel + 384 + g * 64;
        float m = -INFINITY;            // reference max of the exponent (may lag the true running max by < 2^8)
        float l = 0.f;
        const float sc = p.scale_log2;
#ifdef MG_ATT_DEBUG
        const long long dbg_loop0 = clock64();
#endif
        for (int j = 0; j < (g < ng ? nkv : 0); ++j) {
            { ATT_T0(); mbar_wait(&s_full[g], j & 1); ATT_ACC(dbg_a); }
#ifdef MG_ATT_DEBUG
            if (j == 0) dbg_d = dbg_a;
#endif
            tc_fence_after();
            float v[ATT_BKV];
#pragma unroll
            for (int c = 0; c < ATT_BKV; c += 32) tmem_ld32(tS + c, v + c);
            tc_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_free[g]);
            const int kv_left = p.N - j * ATT_BKV;
            if (kv_left < ATT_BKV) {                    // only the last tile has padding columns
#pragma unroll
                for (int i = 0; i < ATT_BKV; ++i) v[i] = (i < kv_left) ? v[i] : -INFINITY;
            }
            // four independent max chains (a single serial chain of 64 dependent FMNMX3 costs ~300 cycles per tile with only
            // two softmax warps per scheduler to hide it)
            float mx0 = fmax3(v[0], v[1], v[2]), mx1 = fmax3(v[3], v[4], v[5]), mx2 = fmax3(v[6], v[7], v[8]), mx3 = fmax3(v[9], v[10], v[11]);
#pragma unroll
            for (int i = 12; i + 7 < ATT_BKV; i += 8) {
                mx0 = fmax3(mx0, v[i], v[i + 1]); mx1 = fmax3(mx1, v[i + 2], v[i + 3]);
                mx2 = fmax3(mx2, v[i + 4], v[i + 5]); mx3 = fmax3(mx3, v[i + 6], v[i + 7]);
            }
            mx0 = fmax3(mx0, v[ATT_BKV - 4], v[ATT_BKV - 3]); mx1 = fmax3(mx1, v[ATT_BKV - 2], v[ATT_BKV - 1]);
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            const float m_new = fmaxf(m, mx);
            const bool grow = (m_new - m) * sc > 8.0f;        // true on the first tile (m = -inf)
            if (j > 0) {
                // PV(j-1) must have finished before P is overwritten (and before O is touched)
                { ATT_T0(); mbar_wait(&o_full[g], (j - 1) & 1); ATT_ACC(dbg_b); }
                tc_fence_after();
                if (__any_sync(0xffffffffu, grow)) {
                    const float alpha = grow ? ex2_approx((m - m_new) * sc) : 1.0f;
#pragma unroll
                    for (int c = 0; c < ATT_HD; c += 32) {
                        float o[32];
                        tmem_ld32(tO + c, o);
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) o[i] *= alpha;
                        tmem_st32(tO + c, o);
                    }
                    tc_wait_st();
                    l *= alpha;
                }
            }
            if (grow) m = m_new;
            const float mb = m * sc;
            // exponent argument and row sum with packed 2 x fp32 instructions (FFMA2 / FADD2): the loop is bound by issue slots
            // and the fma pipe next to the MUFU.EX2 stream, not by the tensor pipe
            const float2 sc2 = make_float2(sc, sc), nmb2 = make_float2(-mb, -mb);
            float2 ls01 = make_float2(0.f, 0.f), ls23 = ls01;       // independent partial sums (no serial FADD chain)
#pragma unroll
            for (int c = 0; c < ATT_BKV; c += 32) {
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float2 t0 = ffma2(make_float2(v[c + i], v[c + i + 1]), sc2, nmb2);
                    const float2 t1 = ffma2(make_float2(v[c + i + 2], v[c + i + 3]), sc2, nmb2);
                    const float2 e0 = make_float2(ex2_approx(t0.x), ex2_approx(t0.y));
                    const float2 e1 = make_float2(ex2_approx(t1.x), ex2_approx(t1.y));
                    ls01 = fadd2(ls01, e0); ls23 = fadd2(ls23, e1);
                    w[(i >> 1)] = H::pack(e0.x, e0.y); w[(i >> 1) + 1] = H::pack(e1.x, e1.y);
                }
                tmem_st16(tP + (c >> 1), w);
            }
            l += (ls01.x + ls01.y) + (ls23.x + ls23.y);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[g]);
        }
#ifdef MG_ATT_DEBUG
        dbg_c = clock64() - dbg_loop0;
#endif
        if (g < ng) {
        mbar_wait(&o_full[g], (nkv - 1) & 1);
        tc_fence_after();
        const float inv = 1.0f / l;
#pragma unroll
        for (int c = 0; c < ATT_HD; c += 32) {
            float o[32];
            tmem_ld32(tO + c, o);
            tc_wait_ld();
            if (qrow < p.N) {
                uint4 q[4];
                uint32_t* qw = reinterpret_cast<uint32_t*>(q);
#pragma unroll
                for (int i = 0; i < 32; i += 2) qw[i >> 1] = H::pack(o[i] * inv, o[i + 1] * inv);
                uint4* dst = reinterpret_cast<uint4*>(static_cast<typename H::T*>(p.out) +
                                                      (static_cast<size_t>(b) * p.N + qrow) * p.D + h * ATT_HD + c);
                dst[0] = q[0]; dst[1] = q[1]; dst[2] = q[2]; dst[3] = q[3];
            }
            __syncwarp();
        }
        }
#ifdef MG_ATT_DEBUG
        if (warp == 4 && lane == 0) {
            const long long now = clock64();
            atomicAdd(&mg_att_dbg[0], (unsigned long long)dbg_a); atomicAdd(&mg_att_dbg[1], (unsigned long long)dbg_b);
            atomicAdd(&mg_att_dbg[2], (unsigned long long)dbg_c); atomicAdd(&mg_att_dbg[5], (unsigned long long)(now - dbg_start));
            atomicAdd(&mg_att_dbg[6], (unsigned long long)(dbg_loop0 - dbg_start + dbg_d));
            atomicAdd(&mg_att_dbg[7], (unsigned long long)(now - dbg_loop0 - dbg_c));
        }
#endif
    }
#ifdef MG_ATT_DEBUG
    if (warp == 1 && lane == 0) { atomicAdd(&mg_att_dbg[3], (unsigned long long)dbg_a); atomicAdd(&mg_att_dbg[4], (unsigned long long)dbg_b); }
#endif
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------ host
int launch_attention(const CUtensorMap& mapQKV, void* out, int B, int N, int D, int heads, bool bf16, cudaStream_t st) {
    if (D != heads * ATT_HD) return set_error("attention: head dim must be 64 (D=%d heads=%d)", D, heads);
    AttnParams p;
    p.out = out; p.B = B; p.N = N; p.D = D; p.heads = heads;
    p.scale_log2 = 0.125f * 1.4426950408889634f;
    dim3 grid((N + 2 * ATT_BQ - 1) / (2 * ATT_BQ), heads, B);
    auto kern = bf16 ? attention_kernel<true> : attention_kernel<false>;
    static bool attr_set[2] = {false, false};
    if (!attr_set[bf16]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
        attr_set[bf16] = true;
    }
    kern<<<grid, ATT_THREADS, ATT_SMEM, st>>>(mapQKV, p);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace mg

#ifdef MG_ATT_DEBUG
extern "C" int mg_debug_att(unsigned long long* out, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, mg::mg_att_dbg, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(mg::mg_att_dbg, z, sizeof(z)); }
    return 0;
}
#endif
