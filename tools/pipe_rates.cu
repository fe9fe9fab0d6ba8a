// Issue / pipe rates of the instructions the attention softmax is made of, measured on the GPU at hand.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pipe_rates tools/pipe_rates.cu && tools/bin/pipe_rates
// One CTA per SM, W warps per SMSP (W = 1, 2, 4), every warp runs ITER x 8 independent chains of one instruction kind.
// Output: cycles per warp-instruction per SMSP (reciprocal throughput) at W = 2 (the softmax's occupancy) and at W = 4.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>

constexpr int ITER = 2048;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra, rb, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

template <int KIND>
__global__ void rate_kernel(float* out, long long* cyc, float seed) {
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 1e-3f + i; b[i] = seed * 0.5f + i * 0.25f; }
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0) { asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a[i]) : "f"(a[i])); }                 // MUFU.EX2
            else if (KIND == 1) { asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(a[i]) : "f"(a[i]), "f"(b[i]), "f"(a[i])); }      // FFMA
            else if (KIND == 2) { float2 r = ffma2(make_float2(a[i], b[i]), make_float2(b[i], a[i]), make_float2(a[i], b[i])); a[i] = r.x; b[i] = r.y; }   // FFMA2
            else if (KIND == 3) { float2 r = fadd2(make_float2(a[i], b[i]), make_float2(b[i], a[i])); a[i] = r.x; b[i] = r.y; }      // FADD2
            else if (KIND == 4) { asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(a[i]) : "f"(a[i]), "f"(b[i]), "f"(a[(i + 1) & 7])); }   // FMNMX3
            else if (KIND == 5) { asm volatile("max.f32 %0, %1, %2;" : "=f"(a[i]) : "f"(a[i]), "f"(b[i])); }                           // FMNMX
            else if (KIND == 6) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(b[i])); a[i] = __uint_as_float(r); }  // F2FP
            else if (KIND == 7) { uint32_t r = __float_as_uint(a[i]) + (__float_as_uint(b[i]) << 23); asm volatile("" : "+r"(r)); a[i] = __uint_as_float(r); }  // shift-add (IMAD/LEA)
            else if (KIND == 8) { asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(a[i]) : "f"(a[i]), "f"(b[i])); }                        // FMUL
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
double run(int warps_per_smsp) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 32 * 4 * warps_per_smsp;
    float* out; long long* cyc;
    cudaMalloc(&out, sms * threads * 4);
    cudaMalloc(&cyc, sms * 8);
    rate_kernel<KIND><<<sms, threads>>>(out, cyc, 0.37f);
    rate_kernel<KIND><<<sms, threads>>>(out, cyc, 0.37f);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += h[i];
    avg /= sms;
    cudaFree(out); cudaFree(cyc);
    return avg / (static_cast<double>(ITER) * 8 * warps_per_smsp);     // cycles per warp-instruction per SMSP
}

int main() {
    const char* names[9] = {"MUFU.EX2", "FFMA", "FFMA2 (fma.rn.f32x2)", "FADD2 (add.rn.f32x2)", "FMNMX3 (max.f32 a,b,c)", "FMNMX", "F2FP (cvt.rn.f16x2.f32)",
                            "shift-add (exponent insert)", "FMUL"};
    printf("%-30s %10s %10s %10s   (cycles per warp-instruction per SMSP; 8 independent chains per warp)\n", "instruction", "1 warp", "2 warps", "4 warps");
    double r[9][3];
    const int w[3] = {1, 2, 4};
    for (int j = 0; j < 3; ++j) {
        r[0][j] = run<0>(w[j]); r[1][j] = run<1>(w[j]); r[2][j] = run<2>(w[j]); r[3][j] = run<3>(w[j]); r[4][j] = run<4>(w[j]);
        r[5][j] = run<5>(w[j]); r[6][j] = run<6>(w[j]); r[7][j] = run<7>(w[j]); r[8][j] = run<8>(w[j]);
    }
    for (int k = 0; k < 9; ++k) printf("%-30s %10.2f %10.2f %10.2f\n", names[k], r[k][0], r[k][1], r[k][2]);
    return 0;
}
