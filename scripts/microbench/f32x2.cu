// Microbenchmark: packed FP32 (FFMA2 / FADD2, new on sm_100) against scalar FFMA — issue slots vs FMA-pipe cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    unsigned long long r, x = *reinterpret_cast<unsigned long long*>(&a), y = *reinterpret_cast<unsigned long long*>(&b),
                          z = *reinterpret_cast<unsigned long long*>(&c);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(x), "l"(y), "l"(z));
    return *reinterpret_cast<float2*>(&r);
}
template <int MODE>
__global__ void __launch_bounds__(256) k(int iters, float a, float* sink, unsigned int* isink) {
    float v[16];
    float2 w[8];
    unsigned int q[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] = make_float2(v[2 * i], v[2 * i + 1]); q[i] = threadIdx.x + i; }
    const float2 a2 = make_float2(a, a), b2 = make_float2(0.5f, 0.25f);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {   // 16 scalar FFMA
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, 0.5f);
        } else if (MODE == 1) {   // 8 FFMA2 (= 16 fp32 fma)
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = fma2(w[i], a2, b2);
        } else if (MODE == 2) {   // 16 FFMA + 8 integer adds: 24 issue slots, 16 FMA-pipe cycles, 16 ALU-pipe cycles
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, 0.5f);
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] += 0x9e3779b9u + (unsigned)it;
        } else {   // 8 FFMA2 + 8 integer adds: 16 issue slots if FFMA2 is one slot, same pipe cycles
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = fma2(w[i], a2, b2);
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] += 0x9e3779b9u + (unsigned)it;
        }
    }
    float s = 0.f; unsigned int t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s += w[i].x + w[i].y; t ^= q[i]; }
    if (s == 123.456f) sink[threadIdx.x] = s;
    if (t == 0x12345u) isink[threadIdx.x] = t;
}
template <int MODE> double run(int iters) {
    float* sink; unsigned int* isink; cudaMalloc(&sink, 4096); cudaMalloc(&isink, 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0); k<MODE><<<sms * 8, 256>>>(iters, 1.0001f, sink, isink); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
    }
    return best;
}
int main() {
    const int iters = 1 << 14;
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double fma_per_launch = 16.0 * iters * sms * 8 * 256;
    const double t0 = run<0>(iters), t1 = run<1>(iters), t2 = run<2>(iters), t3 = run<3>(iters);
    printf("{\"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f, \"ffma_plus_int_tflops\": %.2f, \"ffma2_plus_int_tflops\": %.2f, "
           "\"ms\": [%.3f, %.3f, %.3f, %.3f]}\n", 2 * fma_per_launch / t0 / 1e9, 2 * fma_per_launch / t1 / 1e9,
           2 * fma_per_launch / t2 / 1e9, 2 * fma_per_launch / t3 / 1e9, t0, t1, t2, t3);
    return 0;
}
