// Micro-benchmark: the butterfly passes of stft_pkz2048_kernel (fft64_pk_after_stage1 + two twisted 32-point passes per
// row pair = 960 packed FP32x2 instructions per stereo frame) alone, on registers, no memory traffic: what fraction of the
// FP32 pipe (one FFMA2 per 2 cycles and scheduler) does this instruction stream reach with W warps per SM?  Tooling only.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I jadespectrogram_b200/csrc -o tools/bin/mb_fp_passes tools/mb/mb_fp_passes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "jade_pkz.cuh"
using namespace jade;
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) k(float* o, const uint32_t* tab, int iters, long long* cyc)
{
    f2 v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = pk(1e-30f * (threadIdx.x + i), 1e-30f * i);
    uint32_t ta[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) ta[i] = tab[(threadIdx.x & 31) * 32 + i];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        fft64_pk_after_stage1(v);
        fft32_twisted_lo(v, ta);
        fft32_twisted_hi(v, ta);
        fft32_twisted_lo(v + 32, ta);
        fft32_twisted_hi(v + 32, ta);
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = mul2(v[i], pk(1e-3f, 1e-3f)); // keep the values finite (64 more FMUL2)
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 64; ++i) s += lo(v[i]) + hi(v[i]);
    if (s == 12345.f) o[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int WARPS>
void run(float* o, uint32_t* tab, long long* cyc, int iters)
{
    k<WARPS><<<148, WARPS * 32>>>(o, tab, iters, cyc);
    cudaDeviceSynchronize();
    k<WARPS><<<148, WARPS * 32>>>(o, tab, iters, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0;
    for (int i = 0; i < 148; ++i) c += h[i];
    c /= 148;
    const double per_frame_smsp = c / iters / (WARPS / 4.0); // cycles per frame and scheduler
    printf("%2d warps: %.0f cycles per frame and scheduler (962 packed instructions -> floor 1924): FP32 pipe %.1f %%   [%s]\n", WARPS,
           per_frame_smsp, 100.0 * 1924 / per_frame_smsp, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    float* o;
    uint32_t* tab;
    long long* cyc;
    cudaMalloc(&o, 4);
    cudaMalloc(&tab, 32 * 32 * 4);
    cudaMalloc(&cyc, 148 * 8);
    float h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = 0.5f + 0.001f * i;
    cudaMemcpy(tab, h, sizeof h, cudaMemcpyHostToDevice);
    run<4>(o, tab, cyc, 2000);
    run<8>(o, tab, cyc, 2000);
    run<12>(o, tab, cyc, 2000);
    return 0;
}
