// mb_tmem.cu -- tcgen05.ld (LDTM) throughput per SM for 1..12 warps, alone and next to a shared-memory load stream, and
// tcgen05.st -> tcgen05.ld round trips.  Answers: is the tensor-memory read port per SM or per sub-partition, and does it
// share bandwidth with the shared-memory pipe?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/mb_tmem tools/mb/mb_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD16(addr, r)                                                                                                                      \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                   \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),  \
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])                                             \
                 : "r"(addr))
#define WAITLD(r)                                                                                                                          \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                                          \
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),  \
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]))

// mode 0: tm_warps warps stream LDTM.x16 (4 independent loads in flight), the others idle
// mode 1: the same while the remaining warps stream LDS.128
// mode 2: only the LDS.128 stream (warps >= tm_warps)
__global__ void __launch_bounds__(384, 1) k_tm(int tm_warps, int mode, int iters, long long* cycles, unsigned* sink)
{
    __shared__ uint32_t slot;
    __shared__ float4 buf[384 * 4];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        const uint32_t a = (uint32_t)__cvta_generic_to_shared(&slot);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 384 * 4; i += blockDim.x) buf[i] = make_float4(i, 1, 2, 3);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tq = slot + (((warp & 3) * 32u) << 16);
    unsigned acc = 0;
    float facc = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < tm_warps && mode != 2) {
        uint32_t r0[16], r1[16], r2[16], r3[16];
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
            LD16(tq, r0);
            LD16(tq + 16, r1);
            LD16(tq + 32, r2);
            LD16(tq + 48, r3);
            WAITLD(r0);
            WAITLD(r1);
            WAITLD(r2);
            WAITLD(r3);
            acc += r0[0] ^ r1[3] ^ r2[7] ^ r3[15];
        }
    } else if (warp >= tm_warps && mode != 0) {
        const float4* p = buf + (threadIdx.x & 31);
#pragma unroll 1
        for (int i = 0; i < iters; ++i) {
            const float4 a = p[0], b = p[32], c = p[64], d = p[96];
            facc += a.x + b.y + c.z + d.w;
            p = buf + ((threadIdx.x + (int)facc) & 31);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * 384 + threadIdx.x] = acc + (unsigned)facc;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(128) : "memory");
}

int main()
{
    long long* d;
    unsigned* s;
    cudaMalloc(&d, 8 * 148);
    cudaMalloc(&s, 148 * 384 * 4);
    const int iters = 20000;
    for (int mode = 0; mode < 3; ++mode)
        for (int w : {1, 2, 3, 4, 8, 12}) {
            if (mode != 0 && w == 12) continue;
            long long h = 0;
            k_tm<<<148, 384>>>(w, mode, iters, d, s);
            k_tm<<<148, 384>>>(w, mode, iters, d, s);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
            const double tm_bytes = (mode == 2 ? 0.0 : (double)w) * iters * 4 * 16 * 4 * 32;
            const double lds_bytes = (mode == 0 ? 0.0 : (double)(12 - w)) * iters * 4 * 16 * 32;
            printf("mode %d  tmem warps %2d: %9lld cycles  LDTM %.1f B/clk/SM  LDS %.1f B/clk/SM  (%s)\n", mode, w, h, tm_bytes / h, lds_bytes / h,
                   cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
