// mb_smsp.cu -- which warps of a CTA share a scheduler (SM sub-partition)?  Warps 0 and k run a dependent-free FFMA stream
// (one warp alone issues it at ~1 instruction per clock); the pair takes twice as long when both sit on the same scheduler.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/mb_smsp tools/mb/mb_smsp.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) k_pair(int other, long long* cycles, float* sink)
{
    const int warp = threadIdx.x >> 5;
    float a0 = threadIdx.x, a1 = 1.f, a2 = 2.f, a3 = 3.f, a4 = 4.f, a5 = 5.f, a6 = 6.f, a7 = 7.f;
    __syncthreads();
    const long long t0 = clock64();
    if (warp == 0 || warp == other) {
#pragma unroll 1
        for (int i = 0; i < 20000; ++i) {
            a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
            a4 = fmaf(a4, 1.0001f, 0.5f); a5 = fmaf(a5, 1.0001f, 0.5f); a6 = fmaf(a6, 1.0001f, 0.5f); a7 = fmaf(a7, 1.0001f, 0.5f);
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    sink[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main()
{
    long long* d; float* s; cudaMalloc(&d, 8); cudaMalloc(&s, 384 * 4);
    for (int other = 0; other < 12; ++other) {
        long long h = 0;
        k_pair<<<1, 384>>>(other, d, s);
        k_pair<<<1, 384>>>(other, d, s);
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("warps 0 and %2d: %lld cycles (%.2f cycles per FFMA of warp 0)\n", other, h, h / 160000.0);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
