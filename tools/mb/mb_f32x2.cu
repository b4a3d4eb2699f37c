// Micro-benchmark: FP32 pipe throughput on sm_100a for scalar (FFMA/FADD) vs packed (FFMA2/FADD2) forms, and with
// shared-memory loads interleaved.  Prints warp-instructions/clk/SM and flops/clk/SM.  Tooling only (not product).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ float fma1(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(r):"f"(a),"f"(b),"f"(c)); return r;}
__device__ __forceinline__ float add1(float a, float b){ float r; asm volatile("add.rn.f32 %0, %1, %2;":"=f"(r):"f"(a),"f"(b)); return r;}

constexpr int ITERS = 4096, ACC = 16;
template <int MODE>
__global__ void __launch_bounds__(256) k(float* o, float x, float y, long long* cyc)
{
    __shared__ float2 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_float2(x * i, y);
    __syncthreads();
    long long t0 = clock64();
    float res = 0.f;
    if (MODE == 0) { // FFMA
        float a[ACC];
        for (int i = 0; i < ACC; i++) a[i] = x + i;
        for (int it = 0; it < ITERS; it++)
#pragma unroll
            for (int i = 0; i < ACC; i++) a[i] = fma1(a[i], x, y);
        for (int i = 0; i < ACC; i++) res += a[i];
    } else if (MODE == 1) { // FFMA2
        u64 a[ACC]; u64 bx = pk(x, x), by = pk(y, y);
        for (int i = 0; i < ACC; i++) a[i] = pk(x + i, y + i);
        for (int it = 0; it < ITERS; it++)
#pragma unroll
            for (int i = 0; i < ACC; i++) a[i] = fma2(a[i], bx, by);
        for (int i = 0; i < ACC; i++) { float u, v; upk(a[i], u, v); res += u + v; }
    } else if (MODE == 2) { // FADD
        float a[ACC];
        for (int i = 0; i < ACC; i++) a[i] = x + i;
        for (int it = 0; it < ITERS; it++)
#pragma unroll
            for (int i = 0; i < ACC; i++) a[i] = add1(a[i], y);
        for (int i = 0; i < ACC; i++) res += a[i];
    } else if (MODE == 3) { // FADD2
        u64 a[ACC]; u64 by = pk(y, y);
        for (int i = 0; i < ACC; i++) a[i] = pk(x + i, y + i);
        for (int it = 0; it < ITERS; it++)
#pragma unroll
            for (int i = 0; i < ACC; i++) a[i] = add2(a[i], by);
        for (int i = 0; i < ACC; i++) { float u, v; upk(a[i], u, v); res += u + v; }
    } else if (MODE == 4) { // FFMA + LDS.64 4:1
        float a[ACC];
        for (int i = 0; i < ACC; i++) a[i] = x + i;
        int idx = threadIdx.x;
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int i = 0; i < ACC; i++) a[i] = fma1(a[i], x, y);
#pragma unroll
            for (int i = 0; i < ACC / 4; i++) { float2 v = sm[(idx + 32 * i + it) & 1023]; a[i] += v.x; a[i + 4] += v.y; }
        }
        for (int i = 0; i < ACC; i++) res += a[i];
    } else if (MODE == 5) { // FFMA2 + LDS.64 (same flops as mode 4)
        u64 a[ACC / 2]; u64 bx = pk(x, x), by = pk(y, y);
        for (int i = 0; i < ACC / 2; i++) a[i] = pk(x + i, y + i);
        int idx = threadIdx.x;
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int i = 0; i < ACC / 2; i++) a[i] = fma2(a[i], bx, by);
#pragma unroll
            for (int i = 0; i < ACC / 4; i++) { float2 v = sm[(idx + 32 * i + it) & 1023]; a[i] = add2(a[i], pk(v.x, v.y)); }
        }
        for (int i = 0; i < ACC / 2; i++) { float u, v; upk(a[i], u, v); res += u + v; }
    } else if (MODE == 6) { // FFMA2 and FFMA interleaved 1:1 (do they share the pipe?)
        u64 a[ACC / 2]; u64 bx = pk(x, x), by = pk(y, y);
        float b[ACC / 2];
        for (int i = 0; i < ACC / 2; i++) { a[i] = pk(x + i, y + i); b[i] = x - i; }
        for (int it = 0; it < ITERS; it++)
#pragma unroll
            for (int i = 0; i < ACC / 2; i++) { a[i] = fma2(a[i], bx, by); b[i] = fma1(b[i], x, y); }
        for (int i = 0; i < ACC / 2; i++) { float u, v; upk(a[i], u, v); res += u + v + b[i]; }
    }
    long long t1 = clock64();
    o[blockIdx.x * blockDim.x + threadIdx.x] = res;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, double warp_instr_per_iter, double flops_per_thread_iter, int blocks_per_sm)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int grid = sms * blocks_per_sm;
    float* o; long long* c;
    cudaMalloc(&o, grid * 256 * 4); cudaMalloc(&c, grid * 8);
    k<MODE><<<grid, 256>>>(o, 1.0001f, 0.5f, c);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(o, 1.0001f, 0.5f, c);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long hc[4096]; cudaMemcpy(hc, c, grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += hc[i]; avg /= grid;
    double warps_per_sm = 8.0 * blocks_per_sm;
    double winstr = warp_instr_per_iter * ITERS * warps_per_sm; // per SM
    printf("%-28s blocks/SM %d  cycles %.0f  ms %.3f  warp-instr/clk/SM %.3f  flops/clk/SM %.1f  (err %s)\n", name, blocks_per_sm, avg, ms,
           winstr / avg, flops_per_thread_iter * ITERS * warps_per_sm * 32 / avg, cudaGetErrorString(cudaGetLastError()));
    cudaFree(o); cudaFree(c);
}
int main()
{
    for (int b = 1; b <= 4; b *= 2) {
        run<0>("FFMA", ACC, 2.0 * ACC, b);
        run<1>("FFMA2", ACC, 4.0 * ACC, b);
        run<2>("FADD", ACC, 1.0 * ACC, b);
        run<3>("FADD2", ACC, 2.0 * ACC, b);
        run<4>("FFMA+LDS64 4:1", ACC + ACC / 4 + ACC / 2, 2.0 * ACC + ACC / 2, b);
        run<5>("FFMA2+LDS64", ACC / 2 + ACC / 4 + ACC / 4, 2.0 * ACC + ACC / 2, b);
        run<6>("FFMA2:FFMA 1:1", ACC, 3.0 * ACC, b);
    }
    return 0;
}
