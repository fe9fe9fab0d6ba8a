// Micro-benchmark: cycles per tcgen05.mma (kind::f16, cta_group::1) as a function of the instruction shape (M, N),
// issued back to back by one thread on garbage operands in shared memory.  Ground truth for the decoder/attention
// tile-shape decisions in DESIGN.md.   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/mma_cost tools/mma_cost.cu
#include "../moge_b200/csrc/common.cuh"
#include <cstdio>
using namespace mg;

__global__ void __launch_bounds__(128, 1) k(int M, int N, int reps, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 1.0
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc(M, N, 0);
        const uint64_t a = make_sdesc_sw128(smem_u32(smem)), b = make_sdesc_sw128(smem_u32(smem) + 32768);
        // warm-up
        umma_f16(tmem, a, b, idesc, 0);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) umma_f16(tmem, a + 2 * (r & 3), b + 2 * (r & 3), idesc, 1);
        const long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait(&bar, 1);
        const long long t2 = clock64();
        out[0] = t1 - t0;   // issue time
        out[1] = t2 - t0;   // completion time
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
    long long* d;
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    const int shapes[][2] = {{128, 256}, {128, 128}, {128, 64}, {128, 32}, {128, 16}, {64, 256}, {64, 128}, {64, 64}, {64, 16}};
    for (auto& s : shapes)
        for (int reps : {64, 256}) {
            k<<<1, 128, 66 * 1024>>>(s[0], s[1], reps, d);
            long long h[2];
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            printf("M=%3d N=%3d reps=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma  (%s)\n", s[0], s[1], reps, double(h[0]) / reps,
                   double(h[1]) / reps, cudaGetErrorString(e));
        }
    return 0;
}
