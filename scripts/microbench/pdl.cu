// Launch-gap microbenchmark: chains of small dependent kernels, plain stream order vs programmatic dependent launch.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_plain(float* p, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] * 1.0001f + 1.f;
}
__global__ void k_pdl(float* p, int n) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] * 1.0001f + 1.f;
}
int main() {
    const int chain = 200;
    for (int n : {1 << 10, 1 << 20, 1 << 22}) {
        float* d;
        cudaMalloc(&d, sizeof(float) * n);
        cudaMemset(d, 0, sizeof(float) * n);
        cudaStream_t st;
        cudaStreamCreate(&st);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int g = (n + 255) / 256;
        for (int mode = 0; mode < 2; ++mode) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0, st);
                for (int k = 0; k < chain; ++k) {
                    if (mode == 0) {
                        k_plain<<<g, 256, 0, st>>>(d, n);
                    } else {
                        cudaLaunchConfig_t cfg = {};
                        cfg.gridDim = dim3(g); cfg.blockDim = dim3(256); cfg.stream = st;
                        cudaLaunchAttribute at[1];
                        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                        at[0].val.programmaticStreamSerializationAllowed = 1;
                        cfg.attrs = at; cfg.numAttrs = 1;
                        cudaLaunchKernelEx(&cfg, k_pdl, d, n);
                    }
                }
                cudaEventRecord(e1, st);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("n=%d mode=%s per-kernel %.2f us (%s)\n", n, mode ? "pdl" : "plain", best * 1e3f / chain, cudaGetErrorString(cudaGetLastError()));
        }
        cudaFree(d);
    }
    return 0;
}
