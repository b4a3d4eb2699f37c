// jade_tmem.cuh -- per-lane constant tables in Blackwell tensor memory (TMEM).
//
// The packed N = 2048 kernels read two per-lane tables for every frame: 64 window values and 32 twisted twiddles per lane,
// 16 KB per warp and frame, a quarter of all shared-memory wavefronts of the kernel.  They are constants of the launch, and they
// are addressed [lane][index] -- exactly the shape of tensor memory (128 lanes x 512 32-bit columns per SM, a warp reaches
// the 32 lanes of its quadrant warp % 4).  Held there, they come back through tcgen05.ld (.32x32b: thread l receives
// consecutive columns of lane 32 (warp % 4) + l), on a datapath that does not touch the shared-memory pipe.  No tensor-core
// instruction is involved; TMEM is used as a 64 KB per-quadrant lookup store next to the register file.
//
// tm_ld<N> is asynchronous: its destination registers are valid after tm_wait_ld(), which takes the registers as
// read-write operands so that neither nvcc nor ptxas can move a use above the wait.
#pragma once
#include <stdint.h>
#include <string.h>

#include "jade_fft_regs.cuh"

namespace jade {

#if defined(JADE_EMU)
// host emulator: a [128][512] word array per block (cuda_emu.h)
inline void tm_alloc(uint32_t* slot, uint32_t)
{
    if ((threadIdx.x & 31) == 0) *slot = 0u;
}
inline void tm_dealloc(uint32_t, uint32_t) {}
inline uint32_t* tm_row_emu() { return jade_emu::tmem() + (size_t)(((threadIdx.x >> 5) & 3) * 32 + (threadIdx.x & 31)) * 512; }
template <int N>
inline void tm_st(uint32_t taddr, const uint32_t* r)
{
    uint32_t* row = tm_row_emu() + (taddr & 0xffffu);
    for (int i = 0; i < N; ++i) row[i] = r[i];
}
template <int N>
inline void tm_ld(uint32_t taddr, uint32_t* r)
{
    const uint32_t* row = tm_row_emu() + (taddr & 0xffffu);
    for (int i = 0; i < N; ++i) r[i] = row[i];
}
template <int N>
inline void tm_wait_ld(uint32_t*) {}
template <int N>
inline void tm_tie(uint32_t*) {}
inline void tm_wait_st() {}
inline void tm_fence_before_sync() {}
inline void tm_fence_after_sync() {}
inline uint32_t tm_quadrant_base(uint32_t taddr) { return taddr; }
#else
// allocate ncols (power of two >= 32) columns for this CTA; called by every lane of ONE warp; the address lands in *slot
__device__ __forceinline__ void tm_alloc(uint32_t* slot, uint32_t ncols)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(slot);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(a), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_dealloc(uint32_t taddr, uint32_t ncols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// address of this warp's quadrant (lane field = bits 31..16)
__device__ __forceinline__ uint32_t tm_quadrant_base(uint32_t taddr) { return taddr + ((((threadIdx.x >> 5) & 3u) * 32u) << 16); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int N>
__device__ __forceinline__ void tm_st(uint32_t taddr, const uint32_t* r);
template <>
__device__ __forceinline__ void tm_st<8>(uint32_t taddr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
template <>
__device__ __forceinline__ void tm_st<16>(uint32_t taddr, const uint32_t* r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                 "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
template <>
__device__ __forceinline__ void tm_st<32>(uint32_t taddr, const uint32_t* r)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

template <int N>
__device__ __forceinline__ void tm_ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tm_ld<4>(uint32_t taddr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
template <>
__device__ __forceinline__ void tm_ld<8>(uint32_t taddr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
template <>
__device__ __forceinline__ void tm_ld<16>(uint32_t taddr, uint32_t* r)
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
template <>
__device__ __forceinline__ void tm_ld<32>(uint32_t taddr, uint32_t* r)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}

// wait for every tcgen05.ld of this thread; the N registers are tied to the wait
#if defined(JADE_TM_WAIT_MEMORY)
#define JADE_TM_CLOB ::"memory"
#else
#define JADE_TM_CLOB
#endif
template <int N>
__device__ __forceinline__ void tm_wait_ld(uint32_t* r);
template <>
__device__ __forceinline__ void tm_wait_ld<4>(uint32_t* r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3])JADE_TM_CLOB);
}
template <>
__device__ __forceinline__ void tm_wait_ld<8>(uint32_t* r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])JADE_TM_CLOB);
}
template <>
__device__ __forceinline__ void tm_wait_ld<16>(uint32_t* r)
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])JADE_TM_CLOB);
}
// tie further registers of loads that the preceding tm_wait_ld has completed
template <int N>
__device__ __forceinline__ void tm_tie(uint32_t* r);
template <>
__device__ __forceinline__ void tm_tie<8>(uint32_t* r)
{
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
}
template <>
__device__ __forceinline__ void tm_tie<16>(uint32_t* r)
{
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]));
}
template <>
__device__ __forceinline__ void tm_wait_ld<32>(uint32_t* r)
{
    tm_wait_ld<16>(r);
    // the second half is tied by an empty statement: the wait above has already completed every load
    asm volatile("" : "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                   "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])JADE_TM_CLOB);
}
#endif

// raw words <-> floats (tensor-memory traffic is in 32-bit words)
JADE_HD float u2f(uint32_t u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}
JADE_HD uint32_t f2u(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

} // namespace jade
