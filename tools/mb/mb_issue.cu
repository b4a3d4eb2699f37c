// Micro-benchmark: does the second (pipe) cycle of a packed FFMA2 block the warp scheduler's issue port, or can another
// warp's / the same warp's ALU / LSU instruction issue in it?  Tooling only (not product).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ float fma1(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(r):"f"(a),"f"(b),"f"(c)); return r;}
__device__ __forceinline__ unsigned xr(unsigned a, unsigned b){ unsigned r; asm volatile("shf.l.wrap.b32 %0, %1, %2, 3;":"=r"(r):"r"(a),"r"(b)); return r;}
constexpr int ITERS = 4096, ACC = 8;
// MODE: 0 FFMA2 only; 1 FFMA2 + XOR 1:1; 2 FFMA2 + 2 XOR; 3 FFMA only; 4 FFMA + XOR 1:1; 5 XOR only; 6 FFMA2 + LDS.64 1:1; 7 FFMA2 + XOR + LDS (2:2:1)
template <int MODE>
__global__ void __launch_bounds__(512) k(float* o, float x, float y, long long* cyc)
{
    __shared__ float2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(x * i, y);
    __syncthreads();
    u64 a[ACC]; u64 bx = pk(x, x), by = pk(y, y);
    float f[ACC]; unsigned n[2 * ACC];
    for (int i = 0; i < ACC; i++) { a[i] = pk(x + i, y + i); f[i] = x - i; n[i] = threadIdx.x + i; n[i + ACC] = i; }
    unsigned m = __float_as_uint(y);
    int idx = threadIdx.x;
    float2 acc2 = make_float2(0, 0);
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ACC; i++) {
            if (MODE == 0 || MODE == 1 || MODE == 2 || MODE == 6 || MODE == 7) a[i] = fma2(a[i], bx, by);
            if (MODE == 3 || MODE == 4) f[i] = fma1(f[i], x, y);
            if (MODE == 1 || MODE == 2 || MODE == 4 || MODE == 5 || MODE == 7) n[i] = xr(n[i], m + it);
            if (MODE == 2 || MODE == 5) n[i + ACC] = xr(n[i + ACC], m + it);
            if (MODE == 6 || (MODE == 7 && (i & 1))) { float2 v = sm[(idx + 32 * i + it) & 2047]; acc2.x += v.x; acc2.y += v.y; }
        }
    }
    long long t1 = clock64();
    float res = acc2.x + acc2.y;
    for (int i = 0; i < ACC; i++) { float u, v; upk(a[i], u, v); res += u + v + f[i] + n[i] + n[i + ACC]; }
    o[blockIdx.x * blockDim.x + threadIdx.x] = res;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, double instr_per_inner, int threads)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* o; long long* c;
    cudaMalloc(&o, sms * threads * 4); cudaMalloc(&c, sms * 8);
    k<MODE><<<sms, threads>>>(o, 1.0001f, 0.5f, c);
    k<MODE><<<sms, threads>>>(o, 1.0001f, 0.5f, c);
    cudaDeviceSynchronize();
    long long hc[1024]; cudaMemcpy(hc, c, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; i++) avg += hc[i]; avg /= sms;
    double winstr = instr_per_inner * ACC * ITERS * (threads / 32);
    printf("%-34s warps/SM %2d  cycles/iter %.2f  warp-instr/clk/scheduler %.3f  (%s)\n", name, threads / 32, avg / ITERS,
           winstr / avg / 4.0, cudaGetErrorString(cudaGetLastError()));
    cudaFree(o); cudaFree(c);
}
int main()
{
    for (int t = 128; t <= 512; t *= 2) {
        run<0>("FFMA2", 1, t);
        run<1>("FFMA2 + SHF 1:1", 2, t);
        run<2>("FFMA2 + 2 SHF", 3, t);
        run<3>("FFMA", 1, t);
        run<4>("FFMA + SHF 1:1", 2, t);
        run<5>("SHF x2", 2, t);
        run<6>("FFMA2 + LDS.64 (+2 FADD) 1:1", 4, t);
        run<7>("FFMA2 + SHF + LDS/2", 3.5, t);
    }
    return 0;
}
