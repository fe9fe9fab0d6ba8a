// Micro-benchmark: duration of tcgen05.mma sequences as the attention kernel issues them (S = Q K^T: M128 N128, smem x smem;
// O += P V: M128 N64, A from tensor memory or smem, B MN-major), alone and with the other warps of the CTA generating
// TMEM-load / TMEM-store / shared-memory traffic.   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_mix tools/mma_mix.cu
#include "../moge_b200/csrc/common.cuh"
#include <cstdio>
using namespace mg;

// mode: 0 = SS K-major N=128 | 1 = SS, B MN-major, N=64 | 2 = TS (A in TMEM), B K-major, N=64 | 3 = TS, B MN-major, N=64
//       4 = attention mix: 4 x mode 0 then 8 x mode 3, repeated
// noise: 0 none | 1 tcgen05.ld x32 loops | 2 STS.128 loops | 3 tcgen05.st x16 loops | 4 MUFU loops
__global__ void __launch_bounds__(288, 1) k(int mode, int noise, int reps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t id_s = make_idesc(128, 128, 0, 0, 0), id_o_k = make_idesc(128, 64, 0, 0, 0), id_o_mn = make_idesc(128, 64, 0, 0, 1);
            const uint32_t sa = smem_u32(smem);
            const uint64_t a = make_sdesc_sw128(sa), b = make_sdesc_sw128(sa + 32768);
            const uint64_t bmn = make_sdesc_sw128(sa + 32768, 16384, 1024);
            auto one = [&](int m, int r) {
                if (m == 0) umma_f16(tmem, a + 2 * (r & 3), b + 2 * (r & 3), id_s, 1);
                else if (m == 1) umma_f16(tmem + 256, a + 2 * (r & 3), bmn + (uint64_t)((r & 7) * 128), id_o_mn, 1);
                else if (m == 2) umma_f16_ts(tmem + 256, tmem + 384 + 8 * (r & 7), b + 2 * (r & 3), id_o_k, 1);
                else umma_f16_ts(tmem + 256, tmem + 384 + 8 * (r & 7), bmn + (uint64_t)((r & 7) * 128), id_o_mn, 1);
            };
            one(mode == 4 ? 0 : mode, 0);
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t0 = clock64();
            if (mode < 4) for (int r = 0; r < reps; ++r) one(mode, r);
            else for (int r = 0; r < reps; r += 12) { for (int i = 0; i < 4; ++i) one(0, i); for (int i = 0; i < 8; ++i) one(3, i); }
            const long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, 1);
            const long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
            stop = 1;
        }
    } else if (noise != 0) {
        const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
        float acc = 0.f; long long n = 0;
        uint8_t* mine = smem + 65536 + (warp - 1) * 4096 + lane * 16;
        while (!stop) {
            if (noise == 1) {
                float v[32];
                tmem_ld32(tmem + lane_sel + ((n & 3) * 32), v);
                tc_wait_ld();
                acc += v[0] + v[31];
            } else if (noise == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(mine + i * 512)), "r"((uint32_t)n), "r"(i), "r"(0), "r"(0) : "memory");
            } else if (noise == 3) {
                uint32_t w[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = n + i;
                tmem_st16(tmem + lane_sel + 448 + ((n & 3) * 16), w);
                tc_wait_st();
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) acc += ex2_approx(acc * 0.001f + i);
            }
            ++n;
        }
        if (acc == 12345.f) out[3] = n;
        if (threadIdx.x == 32) out[2] = n;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 32);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const char* mn[] = {"SS  N=128 K-major      ", "SS  N=64  B MN-major   ", "TS  N=64  B K-major    ", "TS  N=64  B MN-major   ", "mix 4xS + 8xPV(TS,MN)  "};
    const char* nn[] = {"quiet", "tmem-ld", "sts", "tmem-st", "mufu"};
    for (int mode = 0; mode < 5; ++mode)
        for (int noise = 0; noise < 5; ++noise) {
            const int reps = 240;
            cudaMemset(d, 0, 32);
            k<<<1, 288, 100 * 1024>>>(mode, noise, reps, d);
            long long h[4];
            cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            printf("%s noise=%-8s: issue %.1f cyc/mma, complete %.1f cyc/mma, noise iters/warp %lld (%s)\n", mn[mode], nn[noise], double(h[0]) / reps,
                   double(h[1]) / reps, h[2], cudaGetErrorString(e));
        }
    return 0;
}
