// jade_pk.cuh -- the headline kernel: N = 2048 (BASELINE configs[1] / [3]) with packed FP32x2 arithmetic.
//
// Same fused path as jade_kernels.cuh (framing, window, real FFT, |X|^2, channel mix, dB, flip, palette, packed pixel
// store; reference lines cited there), restructured around what limits it on sm_100a (profiles/r01*, r01_s2_*):
//   * every complex value lives in ONE 64-bit register pair and all butterflies / twiddle products are FFMA2 / FADD2 /
//     FMUL2 (PTX fma/add/mul.rn.f32x2): a complex add is 1 instruction, a complex multiply 2, a general radix-2
//     butterfly 3 (a' = a + W b by two chained FFMA2, b' = 2a - a').  ptxas folds the half-swap, the per-half sign
//     and the scalar broadcast into operand modifiers (R.F32x2.LO_HI.NP, R.F32), so no repacking moves are needed.
//   * a warp never waits for global memory: while it runs pass 2, the split and the epilogue of one channel, the TMA
//     engine (cp.async.bulk, one instruction of lane 0, completion on a per-warp mbarrier) copies the 8 KB of the NEXT
//     channel / frame into the warp's own shared-memory buffer -- the same buffer the transpose went through a moment
//     before.  No registers are tied up by loads in flight.
//   * the inter-pass twiddles are folded into the pass-2 butterflies ("twisted" DIT, see fft32_twisted): 16 table values
//     per lane instead of 31 and no separate twiddle multiply.
//   * the real-FFT split is done per PAIR (k, M-k): A = Z[k]+conj Z[M-k], B = Z[k]-conj Z[M-k], T = (-i W_N^k) B,
//     X[k] = A+T, X[M-k] = conj(A-T): 8 FP32 instructions per bin instead of 13.  Z[M-k] sits in lane 32 - s, register
//     31 - q and arrives by SHFL.IDX; nothing goes through shared memory for it.
//   * the last channel's split loop goes straight on to dB, palette index (one FFMA from the log2, colour_of_lg) and the
//     coalesced pixel store, so accumulators die as they are consumed and the epilogue overlaps FP work.
//   * per-lane tables (window, twisted twiddles, split twiddles) are laid out [lane][index] with a 16-byte row pad, so
//     that they are read with conflict-free LDS.128 (two table entries per instruction); the transpose rows are 16-byte
//     aligned for the same reason.
//   * 12 warps per SM with up to 168 registers each (one CTA per SM): the 16-warp / 128-register shape spills the stereo
//     accumulators and is 12 % slower; 8 warps lose 8 % (profiles/r01_s2_pk2048_variants.txt).
//
// One warp transforms one frame: M = 1024 complex points z[m] = x[2m] + i x[2m+1], 32 per lane (m = s + 32 n1), radix-32
// in registers, one transpose through shared memory, twisted radix-32 in registers: lane s ends with Z[s + 32 k2].
#pragma once
#include <type_traits>
#include "jade_kernels.cuh"
#include "jade_tmem.cuh"

namespace jade {

// ---------------------------------------------------------------------------------------------------------
// packed pair of floats
// ---------------------------------------------------------------------------------------------------------
#if defined(JADE_EMU)
struct alignas(8) f2 {
    float x, y;
};
inline f2 pk(float a, float b)
{
    f2 r;
    r.x = a;
    r.y = b;
    return r;
}
inline float lo(f2 v) { return v.x; }
inline float hi(f2 v) { return v.y; }
inline f2 add2(f2 a, f2 b) { return pk(a.x + b.x, a.y + b.y); }
inline f2 sub2(f2 a, f2 b) { return pk(a.x - b.x, a.y - b.y); }
inline f2 mul2(f2 a, f2 b) { return pk(a.x * b.x, a.y * b.y); }
inline f2 fma2(f2 a, f2 b, f2 c) { return pk(__builtin_fmaf(a.x, b.x, c.x), __builtin_fmaf(a.y, b.y, c.y)); }
#else
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float a, float b)
{
    f2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lo(f2 v)
{
    float a, b;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return a;
}
__device__ __forceinline__ float hi(f2 v)
{
    float a, b;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return b;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b)
{
    f2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 sub2(f2 a, f2 b)
{
    f2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b)
{
    f2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c)
{
    f2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
#endif
// The three re-packings below never cost an instruction: ptxas folds them into the operand modifiers of the consuming
// FFMA2 / FADD2 (R.F32x2.LO_HI = halves swapped, .NP = per-half sign, leading '-' = both negated).  Constants and
// broadcast scalars must sit on the OTHER operand (pk(c, c) becomes an immediate or R.F32) for this to happen.
JADE_DEVICE f2 neg2(f2 v) { return pk(-lo(v), -hi(v)); }
JADE_DEVICE f2 mul_mi(f2 v) { return pk(hi(v), -lo(v)); }   // v * (-i)
JADE_DEVICE f2 mul_pi(f2 v) { return pk(-hi(v), lo(v)); }   // v * (+i)
JADE_DEVICE f2 conj2(f2 v) { return pk(lo(v), -hi(v)); }
struct alignas(16) f2x2 {
    f2 a, b;
};

// a * w, w = (wx, wy):  (ax wx - ay wy, ay wx + ax wy)  -- FMUL2 + FFMA2
JADE_DEVICE f2 cmul2(f2 a, f2 w)
{
    const float wx = lo(w), wy = hi(w);
    return fma2(mul_pi(a), pk(wy, wy), mul2(a, pk(wx, wx)));
}

// radix-2 DIT butterfly with W = exp(-2 pi i M32/32) = c - i s:  a' = a + W b, b' = a - W b
template <int M32>
JADE_DEVICE void bfly2(f2& a, f2& b)
{
    if (M32 == 0) {
        const f2 t = b;
        b = sub2(a, t);
        a = add2(a, t);
    } else if (M32 == 8) { // W b = -i b = (b.y, -b.x)
        const f2 t = mul_mi(b);
        b = sub2(a, t);
        a = add2(a, t);
    } else {
        constexpr float c = cos32(M32);
        constexpr float s = sin32(M32);
        // W b = c b + s (-i b) = (c bx + s by, c by - s bx)
        const f2 n = fma2(mul_mi(b), pk(s, s), fma2(b, pk(c, c), a));
        b = fma2(a, pk(2.0f, 2.0f), neg2(n));
        a = n;
    }
}
template <int LEN, int BASE, int J>
JADE_DEVICE void pk_inner(f2* a)
{
    if constexpr (J < LEN / 2) {
        bfly2<(J * 32) / LEN>(a[BASE + J], a[BASE + J + LEN / 2]);
        pk_inner<LEN, BASE, J + 1>(a);
    }
}
template <int R, int LEN, int BASE>
JADE_DEVICE void pk_blocks(f2* a)
{
    if constexpr (BASE < R) {
        pk_inner<LEN, BASE, 0>(a);
        pk_blocks<R, LEN, BASE + LEN>(a);
    }
}
template <int R, int LEN>
JADE_DEVICE void pk_stages(f2* a)
{
    if constexpr (LEN <= R) {
        pk_blocks<R, LEN, 0>(a);
        pk_stages<R, LEN * 2>(a);
    }
}
// in-place R-point DFT (R = 2..32), bit-reversed input, natural-order output (same convention as fft_dit<R>)
template <int R>
JADE_DEVICE void fft_pk(f2* a)
{
    pk_stages<R, 2>(a);
}
// the same after its first (twiddle-free) stage has been applied by the caller
template <int R>
JADE_DEVICE void fft_pk_after_stage1(f2* a)
{
    pk_stages<R, 4>(a);
}
JADE_DEVICE void fft32_pk(f2* a) { fft_pk<32>(a); }
JADE_DEVICE void fft32_pk_after_stage1(f2* a) { fft_pk_after_stage1<32>(a); }

// Window multiply fused with the first radix-2 stage of the pass-1 DFT.  Stage 1 of an R-point DFT pairs n1 = j and
// j + R/2 and leaves them in v[2 brev(j)], v[2 brev(j) + 1]:   v0 = x_j w_j + x_{j+R/2} w_{j+R/2},  v1 = x_j w_j - x_{j+R/2} w_{j+R/2}
// written as ONE product and two FFMA2 (3 instructions instead of 4).  The fusion is spelled out because ptxas
// contracts mul.f32x2 + add.f32x2 pairs on its own where it can, which would make differently-compiled instantiations
// of the kernel round differently; with the explicit form every instantiation (and the CPU emulator) agrees bit for bit.
template <int R = 32>
JADE_DEVICE void win_stage1(f2* v, int j, f2 xa, f2 wa, f2 xb, f2 wb)
{
    const int i = brev(j, ilog2c(R) - 1);
    const f2 va = mul2(xa, wa);
    v[2 * i] = fma2(xb, wb, va);
    v[2 * i + 1] = fma2(neg2(xb), wb, va);
}

// ---------------------------------------------------------------------------------------------------------
// asynchronous global -> shared copy (LDGSTS; only the -DJADE_PK_LDGSTS experiment build uses it) and the 64-bit lane shuffle
// ---------------------------------------------------------------------------------------------------------
#if defined(JADE_EMU)
inline void cp_async16(void* dst, const void* src) { std::memcpy(dst, src, 16); }
inline void cp_async_wait_all() {}
inline f2 shfl2(f2 v, int src_lane)
{
    unsigned long long b;
    std::memcpy(&b, &v, 8);
    b = __shfl_sync(0xffffffffu, b, src_lane);
    std::memcpy(&v, &b, 8);
    return v;
}
inline f2 sel2(bool c, f2 a, f2 b) { return c ? a : b; }
#else
__device__ __forceinline__ void cp_async16(void* dst, const void* src)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
    // .cg: L2 only -- the 4x frame overlap is served by L2; measured faster than .ca (gpurun_out/variants4.txt)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ f2 shfl2(f2 v, int src_lane) { return __shfl_sync(0xffffffffu, v, src_lane); }
__device__ __forceinline__ f2 sel2(bool c, f2 a, f2 b) { return c ? a : b; }
#endif

// Bulk asynchronous copy (TMA, cp.async.bulk / UBLKCP): one lane moves a whole 8 KB frame, completion is signalled on a
// per-warp mbarrier.  The bytes reach shared memory through the async proxy, not through LSU instructions.
#if defined(JADE_EMU)
inline void mbar_init(unsigned long long*, int) {}
inline void bulk_copy_g2s(void* dst, const void* src, int bytes, unsigned long long*) { std::memcpy(dst, src, (size_t)bytes); }
inline void bulk_copy2_g2s(void* dst0, const void* src0, void* dst1, const void* src1, int bytes, unsigned long long*)
{
    std::memcpy(dst0, src0, (size_t)bytes);
    std::memcpy(dst1, src1, (size_t)bytes);
}
inline void mbar_expect_tx(unsigned long long*, int) {}
inline void bulk_copy_issue(void* dst, const void* src, int bytes, unsigned long long*) { std::memcpy(dst, src, (size_t)bytes); }
inline void mbar_wait(unsigned long long*, unsigned) {}
#else
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // visible to the async proxy (the __syncthreads follows)
}
// executed by ONE lane after a __syncwarp(): earlier generic-proxy accesses of the warp to dst are ordered before the copy
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, int bytes, unsigned long long* bar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src),
                 "r"(bytes), "r"(b)
                 : "memory");
}
// the same in two steps, for several copies issued by different lanes on one barrier: ONE lane announces the total
// (mbar_expect_tx, after a __syncwarp()), then -- after another __syncwarp() -- any lane issues its copy
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes)
{
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_issue(void* dst, const void* src, int bytes, unsigned long long* bar)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst), b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the issuing lane's own view, as in bulk_copy_g2s
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src),
                 "r"(bytes), "r"(b)
                 : "memory");
}
// two copies of `bytes` each, one completion (2 * bytes) on the barrier
__device__ __forceinline__ void bulk_copy2_g2s(void* dst0, const void* src0, void* dst1, const void* src1, int bytes,
                                               unsigned long long* bar)
{
    const unsigned d0 = (unsigned)__cvta_generic_to_shared(dst0), d1 = (unsigned)__cvta_generic_to_shared(dst1);
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(2 * bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d0), "l"(src0),
                 "r"(bytes), "r"(b)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d1), "l"(src1),
                 "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(b), "r"(parity)
                 : "memory");
}
#endif

// ---------------------------------------------------------------------------------------------------------
// Pass 2 with the inter-pass twiddle folded into the butterflies ("twisted" DIT).  After the transpose lane k1 holds
// y[j] (j = n2 = 0..31) and needs  Z[k1 + 32 k2] = sum_j y[j] u^j W_32^{j k2},  u = W_1024^{k1}.  Writing the DIT recursion
// for the polynomial in (u W_32^{k2}) gives the ordinary radix-2 flow graph whose stage of length LEN uses
//     u^{32/LEN} W_LEN^J = W_1024^{(32/LEN)(k1 + 32 J)},   J = 0 .. LEN/2-1,
// instead of W_LEN^J.  J >= LEN/4 is -i times entry J - LEN/4, so a lane needs 1+1+2+4+8 = 16 table values per
// transform instead of 31 separate twiddle factors, the 31 complex multiplies disappear, and every butterfly is the
// 3-instruction FFMA2 form.  Each table value is one correctly rounded root of unity (from P.twP).
// ---------------------------------------------------------------------------------------------------------
// a' = a + W b, b' = a - W b with W = (wr, wi) held in registers
JADE_DEVICE void bfly_w(f2& a, f2& b, f2 w)
{
    const float wr = lo(w), wi = hi(w);
    const f2 n = fma2(mul_pi(b), pk(wi, wi), fma2(b, pk(wr, wr), a)); // a + wr b + wi (i b)
    b = fma2(a, pk(2.0f, 2.0f), neg2(n));
    a = n;
}
// the same with W = -i (wr, wi) = (wi, -wr)
JADE_DEVICE void bfly_wmi(f2& a, f2& b, f2 w)
{
    const float wr = lo(w), wi = hi(w);
    const f2 n = fma2(mul_mi(b), pk(wr, wr), fma2(b, pk(wi, wi), a)); // a + wi b + wr (-i b)
    b = fma2(a, pk(2.0f, 2.0f), neg2(n));
    a = n;
}
template <int LEN, int BASE, int J>
JADE_DEVICE void tw_inner(f2* a, const f2* tw)
{
    if constexpr (J < LEN / 2) {
        constexpr int Q = (LEN >= 4) ? LEN / 4 : 1;
        if constexpr (J < Q) bfly_w(a[BASE + J], a[BASE + J + LEN / 2], tw[J]);
        else bfly_wmi(a[BASE + J], a[BASE + J + LEN / 2], tw[J - Q]);
        tw_inner<LEN, BASE, J + 1>(a, tw);
    }
}
template <int LEN, int BASE>
JADE_DEVICE void tw_blocks(f2* a, const f2* tw)
{
    if constexpr (BASE < 32) {
        tw_inner<LEN, BASE, 0>(a, tw);
        tw_blocks<LEN, BASE + LEN>(a, tw);
    }
}
// (Computing eleven of the sixteen values as base x constant instead -- 5 table reads and 22 extra FMUL2/FFMA2 per
// transform, and likewise the split twiddles from one base -- removes 14 % of the shared-memory wavefronts but was 2 %
// slower: the FP32 pipe is the scarcer resource, gpurun_out/variants7.txt.)
template <int J0>
JADE_DEVICE void tw32_half(f2* u, const f2x2 ta, const f2x2 tb)
{
    const f2 w[4] = {ta.a, ta.b, tb.a, tb.b};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bfly_w(u[J0 + i], u[J0 + i + 16], w[i]);
        bfly_wmi(u[J0 + i + 8], u[J0 + i + 24], w[i]);
    }
}
// exponent e of table entry t (0..15) of lane k1: the entry is W_1024^e
JADE_HD int tw2_exponent(int k1, int t)
{
    if (t == 0) return 16 * k1;                  // LEN 2
    if (t == 1) return 8 * k1;                   // LEN 4
    if (t < 4) return 4 * (k1 + 32 * (t - 2));   // LEN 8,  J = 0,1
    if (t < 8) return 2 * (k1 + 32 * (t - 4));   // LEN 16, J = 0..3
    return k1 + 32 * (t - 8);                    // LEN 32, J = 0..7
}
// in-place twisted 32-point pass: bit-reversed input, natural-order output; trow = this lane's 16 table values
JADE_DEVICE void fft32_twisted(f2* u, const f2x2* trow)
{
    {
        const f2x2 t = trow[0];
        tw_blocks<2, 0>(u, &t.a);
        tw_blocks<4, 0>(u, &t.b);
    }
    {
        const f2x2 t = trow[1];
        const f2 w[2] = {t.a, t.b};
        tw_blocks<8, 0>(u, w);
    }
    {
        const f2x2 t0 = trow[2], t1 = trow[3];
        const f2 w[4] = {t0.a, t0.b, t1.a, t1.b};
        tw_blocks<16, 0>(u, w);
    }
    // LEN 32 in two halves (entries J = 0..3, then 4..7; J + 8 is -i times entry J) so that only four twiddles are live
    tw32_half<0>(u, trow[4], trow[5]);
    tw32_half<4>(u, trow[6], trow[7]);
}

// two transforms at once: every table value is read once and used for both
template <int J0>
JADE_DEVICE void tw32_half2(f2* ua, f2* ub, const f2x2 ta, const f2x2 tb)
{
    const f2 w[4] = {ta.a, ta.b, tb.a, tb.b};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bfly_w(ua[J0 + i], ua[J0 + i + 16], w[i]);
        bfly_w(ub[J0 + i], ub[J0 + i + 16], w[i]);
        bfly_wmi(ua[J0 + i + 8], ua[J0 + i + 24], w[i]);
        bfly_wmi(ub[J0 + i + 8], ub[J0 + i + 24], w[i]);
    }
}
#if defined(JADE_ABL_NOTW)
#define JADE_TROW(i) trow[0]
#else
#define JADE_TROW(i) trow[i]
#endif
JADE_DEVICE void fft32_twisted2(f2* ua, f2* ub, const f2x2* trow)
{
    {
        const f2x2 t = JADE_TROW(0);
        tw_blocks<2, 0>(ua, &t.a);
        tw_blocks<2, 0>(ub, &t.a);
        tw_blocks<4, 0>(ua, &t.b);
        tw_blocks<4, 0>(ub, &t.b);
    }
    {
        const f2x2 t = JADE_TROW(1);
        const f2 w[2] = {t.a, t.b};
        tw_blocks<8, 0>(ua, w);
        tw_blocks<8, 0>(ub, w);
    }
    {
        const f2x2 t0 = JADE_TROW(2), t1 = JADE_TROW(3);
        const f2 w[4] = {t0.a, t0.b, t1.a, t1.b};
        tw_blocks<16, 0>(ua, w);
        tw_blocks<16, 0>(ub, w);
    }
    tw32_half2<0>(ua, ub, JADE_TROW(4), JADE_TROW(5));
    tw32_half2<4>(ua, ub, JADE_TROW(6), JADE_TROW(7));
}

// the twisted pass in two halves with the table held in registers (raw words of f2 entries t = 0..7, then 8..15)
JADE_DEVICE void fft32_twisted_lo(f2* u, const uint32_t* r)
{
    f2 w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = pk(u2f(r[2 * i]), u2f(r[2 * i + 1]));
    tw_blocks<2, 0>(u, w);
    tw_blocks<4, 0>(u, w + 1);
    tw_blocks<8, 0>(u, w + 2);
    tw_blocks<16, 0>(u, w + 4);
}
JADE_DEVICE void fft32_twisted_hi(f2* u, const uint32_t* r)
{
    f2x2 t[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        t[i].a = pk(u2f(r[4 * i]), u2f(r[4 * i + 1]));
        t[i].b = pk(u2f(r[4 * i + 2]), u2f(r[4 * i + 3]));
    }
    tw32_half<0>(u, t[0], t[1]);
    tw32_half<4>(u, t[2], t[3]);
}


// ---------------------------------------------------------------------------------------------------------
// occupancy: JADE_PK_WARPS warps per CTA, JADE_PK_CTAS CTAs per SM (register cap = 65536 / (32 WARPS CTAS))
#ifndef JADE_PK_WARPS
#define JADE_PK_WARPS 12
#endif
#ifndef JADE_PK_CTAS
#define JADE_PK_CTAS 1
#endif
// one contributing channel (MIX_NONE): 32 complex values per lane fit 128 registers, so 16 warps share the SM (+2..3 % over 12,
// gpurun_out/ab2.txt); the mixing instantiations carry the accumulators as well and spill at 128 (12 warps x 168 registers)
#ifndef JADE_PK_WARPS_MONO
#define JADE_PK_WARPS_MONO 16
#endif
template <int W>
struct PkCfgW {
    static constexpr int M = 1024, N = 2048, B = 1025;
    static constexpr int WARPS = W;
    static constexpr int ROW = 34;   // f2 words per lane row of the window table (32 + 16 B pad)
    static constexpr int TROW = 18;  // f2 words per lane row of the twisted pass-2 table (16 + 16 B pad)
    static constexpr int PROW = 18;  // f2 words per lane row of the split-twiddle table (16 + 16 B pad)
    static constexpr int XROW = 34;  // f2 words per transpose row: 16-byte aligned rows, conflict-free LDS.128
    static constexpr int XCH = 32 * XROW; // f2 words per warp buffer: transpose, and landing area of the next frame (1024)
#ifndef JADE_PK_TMEM
#define JADE_PK_TMEM 1
#endif
    // per-lane window / twisted / split tables in tensor memory (jade_tmem.cuh) instead of shared memory; one CTA per SM only
    // (the engine's occupancy query answers 1 for kernels with tcgen05.alloc)
    static constexpr bool TM = JADE_PK_TMEM != 0 && JADE_PK_CTAS == 1;
    static constexpr int TM_COLS = 512; // window 0..63 (in the order pass 1 consumes it), twisted table 64..95, split table 96..127; 128 + 64 (warp / 4) ..:
                                        // that warp's sample ring (PK_LD_RING*; the warps w, w + 4, w + 8 share the lanes of quadrant w % 4)
    static constexpr int off_win = 0;
    static constexpr int off_tw2 = off_win + (TM ? 0 : 32 * ROW * 8);
    static constexpr int off_twP = off_tw2 + (TM ? 0 : 32 * TROW * 8);
    static constexpr int off_pal = off_twP + (TM ? 0 : 32 * PROW * 8);
    static JADE_HD int off_bar(int npal) { return off_pal + ((npal + 1) * 4 + 15) / 16 * 16; } // palette + its `>= m_Max` entry (colour_of_lg1); then one mbarrier per warp (+ the tensor-memory address)
    static JADE_HD int off_xch(int npal) { return off_bar(npal) + (WARPS * 8 + 8 + 15) / 16 * 16; }
    static JADE_HD int smem_bytes(int npal) { return off_xch(npal) + WARPS * XCH * 8; }
};
using PkCfg = PkCfgW<JADE_PK_WARPS>;
template <int MIXK>
using PkCfgFor = PkCfgW<(MIXK == MIX_NONE && JADE_PK_CTAS == 1) ? JADE_PK_WARPS_MONO : JADE_PK_WARPS>;

// How a frame reaches the registers:
//   PK_LD_ASYNC  : frames lie inside [0, nsamples) and start on a multiple of 4 samples with 16-byte aligned channel
//                  bases (P.aligned4).  Each warp copies the NEXT channel / frame it will transform into its own
//                  buffer with cp.async (LDGSTS.128, no registers, no waiting) right after the transpose of the current
//                  one has been read back, so the global-memory latency is covered by pass 2, the split and the
//                  epilogue (profiles/r01e: 17.5 % of all warp stalls were the first use of a global load).
//   PK_LD_DIRECT : interior frames on an even sample (P.aligned2): LDG.64 straight to registers.
//   PK_LD_GUARD  : bounds-checked scalar loads for the few columns that touch the signal boundary, and unaligned
//                  geometries; routed by launch_stft in jade_gpu.cu.
// The arithmetic after the load is the same code in all three, so streaming, batch and sharded renderings of a column
// agree bit for bit whichever instantiation produced it.
//   PK_LD_RING4 / PK_LD_RING8 : staged like PK_LD_ASYNC, for long runs of evenly spaced columns with hop = 256 / 512 and one
//                  contributing channel: every warp takes a contiguous run of columns and keeps the frame's 32 complex
//                  values per lane in a tensor-memory ring of 8 / 4 chunks, so that a frame whose start lies one hop after
//                  its predecessor's reads only its NEW chunk from shared memory and the TMA engine stages 1 / 2 KB per frame
//                  instead of 8 (cf. PKZ_RING in jade_pkz.cuh).
enum { PK_LD_ASYNC = 0, PK_LD_DIRECT = 1, PK_LD_GUARD = 2, PK_LD_RING4 = 3, PK_LD_RING8 = 4 };

struct PkUnit {
    int stream;
    long long j, st;
};
JADE_DEVICE PkUnit pk_unit(const KParams& P, unsigned g)
{
    PkUnit u;
    u.stream = (int)(g / (unsigned)P.ncols);
    u.j = P.first_col + (g - (unsigned)u.stream * (unsigned)P.ncols);
    u.st = frame_start(P, u.j);
    return u;
}
// 8 KB of one channel's frame -> the warp buffer.  Default: one bulk copy issued by lane 0 (call after a __syncwarp()).
// -DJADE_PK_LDGSTS: 16 x LDGSTS.128 per lane instead (512 contiguous bytes per instruction).
JADE_DEVICE void pk_prefetch(f2* xw, const float* src, int s, unsigned long long* bar)
{
#if defined(JADE_PK_LDGSTS)
    char* d = reinterpret_cast<char*>(xw) + 16 * s;
    const char* g = reinterpret_cast<const char*>(src) + 16 * s;
#pragma unroll
    for (int t = 0; t < 16; ++t) cp_async16(d + 512 * t, g + 512 * t);
#else
    if (s == 0) bulk_copy_g2s(xw, src, PkCfg::M * 8, bar);
#if defined(JADE_EMU)
    __syncwarp();
#endif
#endif
}
// wait for the copy started by the last pk_prefetch of this warp; parity = number of completed copies & 1
JADE_DEVICE void pk_prefetch_wait(unsigned long long* bar, unsigned parity)
{
#if defined(JADE_PK_LDGSTS)
    cp_async_wait_all();
    __syncwarp();
#else
    mbar_wait(bar, parity);
#endif
}

// MIXK: MIX_NONE (one contributing channel) or MIX_SUM (AbsMean over 2^n channels).  Identity rows in the reference
// orientation, hardware log2 for the dB.  WANT_DB additionally stores the float dB column (streaming ring / getMem).
template <int MIXK, bool WANT_DB, int LD>
#if defined(JADE_PK_MAXREG) && !defined(JADE_EMU)
__global__ void __maxnreg__(JADE_PK_MAXREG) stft_pk2048_kernel(const KParams P)
#else
JADE_KERNEL(PkCfgFor<MIXK>::WARPS * 32, JADE_PK_CTAS) stft_pk2048_kernel(const KParams P)
#endif
{
    using Cfg = PkCfgFor<MIXK>;
    constexpr int M = Cfg::M;
    constexpr bool RING = LD == PK_LD_RING4 || LD == PK_LD_RING8, STAGED = LD == PK_LD_ASYNC || RING;
    constexpr int CH = LD == PK_LD_RING8 ? 8 : 4, NCH = 32 / CH; // ring: NCH chunks of CH values n1 (hop = 64 CH samples)
    static_assert(!RING || (Cfg::TM && MIXK == MIX_NONE && Cfg::WARPS <= 24), "the sample ring: tensor memory, one contributing channel");
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    f2* s_win = reinterpret_cast<f2*>(sm + Cfg::off_win);
    f2* s_tw2 = reinterpret_cast<f2*>(sm + Cfg::off_tw2);
    f2* s_twP = reinterpret_cast<f2*>(sm + Cfg::off_twP);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(sm + Cfg::off_bar(P.npal));
    f2* s_xch = reinterpret_cast<f2*>(sm + Cfg::off_xch(P.npal));

    if (STAGED && threadIdx.x < Cfg::WARPS) mbar_init(s_bar + threadIdx.x, 1);
    uint32_t tq = 0; // tensor-memory address of this warp's quadrant of the tables
    if constexpr (Cfg::TM) {
        uint32_t* s_tm = reinterpret_cast<uint32_t*>(s_bar + Cfg::WARPS);
        if (threadIdx.x < 32) tm_alloc(s_tm, Cfg::TM_COLS);
        tm_fence_before_sync();
        __syncthreads();
        tm_fence_after_sync();
        tq = tm_quadrant_base(*s_tm);
        if (threadIdx.x < 128) { // warp q fills quadrant q: thread l writes the tables of lane l
            const int l = threadIdx.x & 31;
            uint32_t r[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) { // window chunk c: points n1 = 4c .. 4c+3, then 16 + 4c .. 16 + 4c + 3 (m = l + 32 n1)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ma = l + 32 * (4 * c + i), mb = ma + 512;
                    r[2 * i] = f2u(P.window[2 * ma]);
                    r[2 * i + 1] = f2u(P.window[2 * ma + 1]);
                    r[8 + 2 * i] = f2u(P.window[2 * mb]);
                    r[8 + 2 * i + 1] = f2u(P.window[2 * mb + 1]);
                }
                tm_st<16>(tq + 16 * c, r);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) { // twisted pass-2 table
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const cpx w = P.twP[2 * tw2_exponent(l, 8 * c + i)]; // W_1024^e = W_2048^{2e}
                    r[2 * i] = f2u(w.x);
                    r[2 * i + 1] = f2u(w.y);
                }
                tm_st<16>(tq + 64 + 16 * c, r);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) { // split table: k = l + 32 q, entry -i W_N^k = (w.y, -w.x)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const cpx w = P.twP[l + 32 * (8 * c + i)];
                    r[2 * i] = f2u(w.y);
                    r[2 * i + 1] = f2u(-w.x);
                }
                tm_st<16>(tq + 96 + 16 * c, r);
            }
            tm_wait_st();
        }
        tm_fence_before_sync();
    } else {
        // ---- per-lane tables: entry for (lane s, index i) at [s*ROW + i]
        for (int i = threadIdx.x; i < M; i += blockDim.x) {
            const int s = i & 31, n1 = i >> 5;                       // m = s + 32 n1
            s_win[s * Cfg::ROW + n1] = pk(P.window[2 * i], P.window[2 * i + 1]);
        }
        for (int i = threadIdx.x; i < 512; i += blockDim.x) {
            const int s = i & 31, q = i >> 5;
            const cpx t = P.twP[2 * tw2_exponent(s, q)];             // W_1024^e = W_2048^{2e}
            s_tw2[s * Cfg::TROW + q] = pk(t.x, t.y);
            const cpx w = P.twP[s + 32 * q];                         // k = s + 32 q: table holds -i W_N^k = (w.y, -w.x)
            s_twP[s * Cfg::PROW + q] = pk(w.y, -w.x);
        }
    }
    for (int i = threadIdx.x; i <= P.npal; i += blockDim.x) s_pal[i] = P.palette[i < P.npal ? i : P.ci_hi];
    __syncthreads();
    if constexpr (Cfg::TM) tm_fence_after_sync();
    grid_dep_wait();

    const int s = threadIdx.x & 31, warp = threadIdx.x >> 5;
    f2* xw = s_xch + warp * Cfg::XCH;
    unsigned long long* bar = s_bar + warp;
    unsigned copies = 0; // bulk copies of this warp waited for so far (mbarrier phase parity)
    const f2x2* wrow = reinterpret_cast<const f2x2*>(s_win + s * Cfg::ROW);
    const f2x2* trow = reinterpret_cast<const f2x2*>(s_tw2 + s * Cfg::TROW);
    const f2x2* prow = reinterpret_cast<const f2x2*>(s_twP + s * Cfg::PROW);
    f2* tr_wr = xw + s;                                                      // transpose: word k1*XROW + s
    const f2x2* tr_rd = reinterpret_cast<const f2x2*>(xw + s * Cfg::XROW);   //            row s, words j (two per load)
    const int partner = (32 - s) & 31; // lane holding Z[M - k] for this lane's k = s + 32 q  (lane 0: itself)

    const unsigned total = (unsigned)P.ncols * (unsigned)P.nstreams;
    const unsigned gstep = gridDim.x * Cfg::WARPS;
    int ch0, ch1;
    channel_range(P, ch0, ch1);
    if (MIXK == MIX_NONE) ch1 = ch0 + 1;
    const float scale = (MIXK == MIX_SUM) ? (1.0f / (float)P.channels) : 1.0f;

    // One bin: dB, palette, packed pixel (and the dB value for the streaming ring).  Reference orientation: bin k lands
    // in row M - k, so lane s writes rows M-s-32q (bins s+32q) and rows s+32q (bins M-s-32q): both coalesced.
    // (one contributing channel: the + 1e-11 is seeded into the power FMAs)
    auto emit = [&](auto u8, float p, uint32_t* pix, float* db) {
        emit_bin1<MIXK, WANT_DB, decltype(u8)::value, MIXK == MIX_NONE>(p, scale, pix, db, P, s_pal);
    };

    // frames of this warp: dealt round-robin (g, g + gstep, ...), or -- PK_LD_RING* -- one contiguous run
    unsigned g = blockIdx.x * Cfg::WARPS + warp, g_end = total, g_inc = gstep;
    if (RING) {
        const unsigned per = (total + gstep - 1) / gstep;
        g = min(total, g * per);
        g_end = min(total, g + per);
        g_inc = 1;
    }
    // PK_LD_RING*: chunk c of the current frame (n1 = CH c .. CH c + CH - 1) sits in ring slot (ring + c) % NCH, columns
    // tring + 2 CH slot + 2 i (+ 1) for n1 = CH c + i
    unsigned ring = 0;
    bool warm = false; // the current frame starts one hop after the previous frame of this warp: all chunks but the last are in the ring
    const uint32_t tring = tq + 128u + 64u * ((unsigned)warp >> 2);
    auto ingest = [&](int c) { // chunk c of the staged frame -> its ring slot
        uint32_t xs[2 * CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const f2 z = xw[s + 32 * (CH * c + i)];
            xs[2 * i] = f2u(lo(z));
            xs[2 * i + 1] = f2u(hi(z));
        }
        tm_st<2 * CH>(tring + 2 * CH * ((ring + c) % NCH), xs);
    };
    // the frame's last chunk only (its first sample at src): 256 CH bytes, in place at the end of the buffer
    auto prefetch_last = [&](const float* src) {
        if (s == 0) bulk_copy_g2s(xw + 32 * (32 - CH), src + 64 * (32 - CH), 256 * CH, bar);
#if defined(JADE_EMU)
        __syncwarp();
#endif
    };
    // PK_LD_RING* (evenly spaced columns one hop = 64 CH samples apart, by dispatch): the walk is (stream, column) plus one source
    // and one output pointer, advanced by one addition per frame -- no division by the column count and no 64-bit multiplies
    // per frame (~100 of 1200 instructions)
    unsigned r_stream = 0, r_col = 0;
    const float* r_x = nullptr;
    ColOut r_o{nullptr, nullptr};
    if (STAGED && g < g_end) {
        const PkUnit un = pk_unit(P, g);
        r_stream = (unsigned)un.stream;
        r_col = (unsigned)(un.j - P.first_col);
        r_x = P.samples + un.stream * P.stream_stride + ch0 * P.channel_stride + un.st;
        r_o = col_out(P, un.stream, un.j);
        if (!WANT_DB) r_o.db = nullptr;
        pk_prefetch(xw, r_x, s, bar);
    }
    for (; g < g_end; g += g_inc) {
        PkUnit un;
        ColOut o;
        const float* x; // frame, channel ch
        if constexpr (RING) {
            un.stream = (int)r_stream, un.j = 0, un.st = 0; // (un.j / un.st: guarded loads only)
            o = r_o;
            x = r_x;
        } else {
            un = pk_unit(P, g);
            o = col_out(P, un.stream, un.j);
            x = P.samples + un.stream * P.stream_stride + ch0 * P.channel_stride + un.st;
        }

        float alo[16], ahi[16], amid = 0.f; // power of bins s+32q / M-(s+32q) / 512 (lane 0) of the channels so far
#pragma unroll
        for (int q = 0; q < 16; ++q) alo[q] = ahi[q] = 0.f;

        for (int ch = ch0; ch < ch1; ++ch, x += P.channel_stride) {
            const bool last = (MIXK == MIX_NONE) || (ch + 1 == ch1);
            f2 v[32];
            if (STAGED) {
                pk_prefetch_wait(bar, copies & 1u);
                ++copies;
            }
            if constexpr (RING) {
                if (!warm) {
                    ring = 0;
#pragma unroll
                    for (int c = 0; c < NCH - 1; ++c) ingest(c);
                }
                ingest(NCH - 1);
                tm_wait_st();
            }
            uint32_t wq[2][16]; // window chunks from tensor memory, one ahead
            uint32_t xq[2][2][8]; // PK_LD_RING*: the samples of the same points from the ring
            auto ring_fetch = [&](int c) { // points n1 = 4c .. 4c+3 and 16 + 4c .. 16 + 4c + 3
                const int na = 4 * c, nb = 16 + 4 * c;
                tm_ld<8>(tring + 2 * CH * ((ring + na / CH) % NCH) + 2 * (na % CH), xq[c & 1][0]);
                tm_ld<8>(tring + 2 * CH * ((ring + nb / CH) % NCH) + 2 * (nb % CH), xq[c & 1][1]);
            };
            if constexpr (RING) ring_fetch(0);
            if constexpr (Cfg::TM) tm_ld<16>(tq, wq[0]);
#pragma unroll
            for (int jj = 0; jj < 16; jj += 2) { // n1 = jj, jj+1 paired with n1 + 16
                f2x2 wa, wb;
                if constexpr (Cfg::TM) {
                    const int c = jj / 4, o = 4 * ((jj / 2) & 1);
                    if ((jj & 2) == 0) {
                        tm_wait_ld<16>(wq[c & 1]);
                        if constexpr (RING) {
                            tm_tie<8>(xq[c & 1][0]);
                            tm_tie<8>(xq[c & 1][1]);
                            if (c < 3) ring_fetch(c + 1);
                        }
                        if (c < 3) tm_ld<16>(tq + 16 * (c + 1), wq[(c + 1) & 1]);
                    }
                    const uint32_t* w = wq[c & 1];
                    wa.a = pk(u2f(w[o]), u2f(w[o + 1]));
                    wa.b = pk(u2f(w[o + 2]), u2f(w[o + 3]));
                    wb.a = pk(u2f(w[8 + o]), u2f(w[8 + o + 1]));
                    wb.b = pk(u2f(w[8 + o + 2]), u2f(w[8 + o + 3]));
                } else {
                    wa = wrow[jj / 2];
                    wb = wrow[(jj + 16) / 2];
                }
                f2 xa0, xa1, xb0, xb1;
                if (RING) {
                    const int c = jj / 4, o = 4 * ((jj / 2) & 1);
                    const uint32_t* qa = xq[c & 1][0];
                    const uint32_t* qb = xq[c & 1][1];
                    xa0 = pk(u2f(qa[o]), u2f(qa[o + 1]));
                    xa1 = pk(u2f(qa[o + 2]), u2f(qa[o + 3]));
                    xb0 = pk(u2f(qb[o]), u2f(qb[o + 1]));
                    xb1 = pk(u2f(qb[o + 2]), u2f(qb[o + 3]));
                } else if (LD == PK_LD_ASYNC) {
                    const f2* xz = xw + s;
                    xa0 = xz[32 * jj];
                    xa1 = xz[32 * (jj + 1)];
                    xb0 = xz[32 * (jj + 16)];
                    xb1 = xz[32 * (jj + 17)];
                } else if (LD == PK_LD_DIRECT) {
                    const f2* xz = reinterpret_cast<const f2*>(x) + s;
                    xa0 = xz[32 * jj];
                    xa1 = xz[32 * (jj + 1)];
                    xb0 = xz[32 * (jj + 16)];
                    xb1 = xz[32 * (jj + 17)];
                } else {
                    const float* xc = x - un.st; // channel base; un.st may be negative / past the end here
                    const cpx a0 = load_pair_guarded(xc, un.st + 2 * (s + 32 * jj), P.nsamples);
                    const cpx a1 = load_pair_guarded(xc, un.st + 2 * (s + 32 * (jj + 1)), P.nsamples);
                    const cpx b0 = load_pair_guarded(xc, un.st + 2 * (s + 32 * (jj + 16)), P.nsamples);
                    const cpx b1 = load_pair_guarded(xc, un.st + 2 * (s + 32 * (jj + 17)), P.nsamples);
                    xa0 = pk(a0.x, a0.y);
                    xa1 = pk(a1.x, a1.y);
                    xb0 = pk(b0.x, b0.y);
                    xb1 = pk(b1.x, b1.y);
                }
                win_stage1(v, jj, xa0, wa.a, xb0, wb.a);
                win_stage1(v, jj + 1, xa1, wa.b, xb1, wb.b);
            }
            if (RING) ring = (ring + 1) % NCH;
            if (STAGED) __syncwarp(); // every lane has read its samples before the transpose overwrites them
            fft32_pk_after_stage1(v);
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) tr_wr[k1 * Cfg::XROW] = v[k1];
            __syncwarp();
            f2 u[32];
#pragma unroll
            for (int jx = 0; jx < 32; jx += 2) {
                const f2x2 t = tr_rd[jx / 2];
                u[brev(jx, 5)] = t.a;
                u[brev(jx + 1, 5)] = t.b;
            }
            __syncwarp(); // the buffer is free again
            if (STAGED) {
                // start the copy of what this warp transforms next: the next channel of this frame, or the first
                // channel of its next frame
                if (!last) {
                    pk_prefetch(xw, x + P.channel_stride, s, bar);
                } else if (g + g_inc < g_end) {
                    if constexpr (RING) {
                        // the next frame continues this one (same stream, one hop further): only its last chunk is new
                        warm = ++r_col != (unsigned)P.ncols;
                        if (warm) {
                            r_x += 64 * CH;
                            prefetch_last(r_x);
                        } else {
                            r_col = 0;
                            ++r_stream;
                            r_x = P.samples + (long long)r_stream * P.stream_stride + ch0 * P.channel_stride + frame_start(P, P.first_col);
                            pk_prefetch(xw, r_x, s, bar);
                        }
                    } else {
                        const PkUnit nx = pk_unit(P, g + g_inc);
                        pk_prefetch(xw, P.samples + nx.stream * P.stream_stride + ch0 * P.channel_stride + nx.st, s, bar);
                    }
                }
            }
            if constexpr (Cfg::TM) { // twisted table from tensor memory, eight entries at a time
                uint32_t ta[16], tb[16];
                tm_ld<16>(tq + 64, ta);
                tm_wait_ld<16>(ta);
                tm_ld<16>(tq + 80, tb);
                fft32_twisted_lo(u, ta);
                tm_wait_ld<16>(tb);
                fft32_twisted_hi(u, tb);
            } else {
                fft32_twisted(u, trow); // u[k2] = Z[s + 32 k2]
            }
            // split table: pairs 0..7, then 8..15 (tensor memory) / two entries per LDS.128 (shared memory)
            uint32_t pw[16];
            auto split_tw = [&](int q) {
                if constexpr (Cfg::TM) {
                    if ((q & 7) == 0) {
                        tm_ld<16>(tq + 96 + 2 * q, pw);
                        tm_wait_ld<16>(pw);
                    }
                    return pk(u2f(pw[2 * (q & 7)]), u2f(pw[2 * (q & 7) + 1]));
                } else {
                    const f2x2 wq2 = prow[q / 2];
                    return (q & 1) ? wq2.b : wq2.a;
                }
            };
            // Real-FFT split per pair (k, M-k), k = s + 32 q.  Z[M-k] is register 31 - q of lane 32 - s; lane 0 pairs with
            // its own register 32 - q.  The last channel goes straight on to dB / colour / store.
            if (!last) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const f2 zp = sel2(s == 0, u[(32 - q) & 31], shfl2(u[31 - q], partner));
                    const f2 A = add2(u[q], conj2(zp));  // Z[k] + conj Z[M-k]
                    const f2 Bv = sub2(u[q], conj2(zp)); // Z[k] - conj Z[M-k]
                    const f2 T = cmul2(Bv, split_tw(q));
                    const f2 xp = add2(A, T), xm = sub2(A, T);
                    alo[q] = fm(lo(xp), lo(xp), fm(hi(xp), hi(xp), alo[q]));
                    ahi[q] = fm(lo(xm), lo(xm), fm(hi(xm), hi(xm), ahi[q]));
                }
                const float a = lo(u[16]), b = hi(u[16]); // bin 512 (lane 0, self-paired): X = 2 conj Z
                amid = fm(JADE_FMUL(4.0f, a), a, fm(JADE_FMUL(4.0f, b), b, amid));
            } else {
                // (two copies behind a launch-uniform branch: with the u8 palette the float -> integer conversion is the clamp)
                auto finish = [&](auto u8) {
                    uint32_t* p_lo = o.pix ? o.pix + (M - s) : nullptr;
                    uint32_t* p_hi = o.pix ? o.pix + s : nullptr;
                    float* d_lo = (WANT_DB && o.db) ? o.db + s : nullptr;
                    float* d_hi = (WANT_DB && o.db) ? o.db + (M - s) : nullptr;
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const f2 zp = sel2(s == 0, u[(32 - q) & 31], shfl2(u[31 - q], partner));
                        const f2 A = add2(u[q], conj2(zp));
                        const f2 Bv = sub2(u[q], conj2(zp));
                        const f2 T = cmul2(Bv, split_tw(q));
                        const f2 xp = add2(A, T), xm = sub2(A, T);
                        const float pl = fm(lo(xp), lo(xp), fm(hi(xp), hi(xp), MIXK == MIX_NONE ? 1e-11f : alo[q]));
                        const float ph = fm(lo(xm), lo(xm), fm(hi(xm), hi(xm), MIXK == MIX_NONE ? 1e-11f : ahi[q]));
                        emit(u8, pl, p_lo ? p_lo - 32 * q : nullptr, d_lo ? d_lo + 32 * q : nullptr);
                        emit(u8, ph, p_hi ? p_hi + 32 * q : nullptr, d_hi ? d_hi - 32 * q : nullptr);
                    }
                    if (s == 0) {
                        const float a = lo(u[16]), b = hi(u[16]);
                        const float pm = fm(JADE_FMUL(4.0f, a), a, fm(JADE_FMUL(4.0f, b), b, MIXK == MIX_NONE ? 1e-11f : amid));
                        emit(u8, pm, o.pix ? o.pix + 512 : nullptr, (WANT_DB && o.db) ? o.db + 512 : nullptr);
                    }
                };
                if (P.pal_u8) finish(std::true_type{});
                else finish(std::false_type{});
            }
        }
        if constexpr (RING) {
            if (g + g_inc < g_end) { // output column of the next frame
                if (warm && P.ring_w == 0) {
                    if (r_o.pix) r_o.pix += P.R;
                    if (r_o.db) r_o.db += P.B;
                } else {
                    r_o = col_out(P, (int)r_stream, P.first_col + r_col);
                    if (!WANT_DB) r_o.db = nullptr;
                }
            }
        }
    }
    if constexpr (Cfg::TM) {
        tm_fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32) tm_dealloc(tq, Cfg::TM_COLS); // warp 0: quadrant 0 = the allocation's base address
    }
}

// ---------------------------------------------------------------------------------------------------------
// Stereo kernel (AbsMean over exactly two channels, the BASELINE configs[1] shape): one warp transforms BOTH channels of a
// frame at once, so that every window, twiddle and split-twiddle value is read from shared memory once for the two
// (the shared-memory pipe is the limiter, profiles/r01_s2_pk2048_stereo.txt: the tables are 256 of 813 wavefronts per stereo
// frame), the two powers add up without accumulators living through a transform, and the two independent dependency
// chains give the scheduler extra parallelism.  12 warps per SM, <= 168 registers.  Interior, 16-byte aligned frames only
// (TMA staging as in PK_LD_ASYNC, one mbarrier completion for the two 8 KB copies); everything else goes to
// stft_pk2048_kernel, whose per-channel arithmetic is the same code in the same order, so the results are bit-identical.
// (The same pairing applied to two consecutive mono columns is slower than one transform per warp: 433 M vs 485 M
// frames/s, gpurun_out/variants10.txt -- mono keeps stft_pk2048_kernel.  Replacing the TMA staging by LDG.64 of the next
// frame into the registers the split loop frees, which would save the 256 staging wavefronts, fits in 168 registers
// but runs at 245 M instead of 291 M frames/s.  A per-warp sample ring for mono renderings -- a warp walks consecutive
// columns and copies only the hop new samples of each, 1 KB instead of 8 KB at hop 256 -- changes nothing: 484 M vs
// 486 M frames/s, gpurun_out/variants13.txt; the staging writes are not on the critical path.)
// ---------------------------------------------------------------------------------------------------------
struct PkPairCfg {
    static constexpr int WARPS = 12;
    // (its own shared-memory tables: PkCfg's moved to tensor memory)
    static constexpr int off_win = 0;
    static constexpr int off_tw2 = off_win + 32 * PkCfg::ROW * 8;
    static constexpr int off_twP = off_tw2 + 32 * PkCfg::TROW * 8;
    static constexpr int off_pal = off_twP + 32 * PkCfg::PROW * 8;
    static JADE_HD int off_bar(int npal) { return off_pal + ((npal + 1) * 4 + 15) / 16 * 16; }
    static JADE_HD int off_xch(int npal) { return off_bar(npal) + (WARPS * 8 + 15) / 16 * 16; }
    static JADE_HD int smem_bytes(int npal) { return off_xch(npal) + WARPS * 2 * PkCfg::XCH * 8; }
};

template <bool WANT_DB>
JADE_KERNEL(PkPairCfg::WARPS * 32, 1) stft_pk2048x2_kernel(const KParams P)
{
    using Cfg = PkCfg;
    constexpr int M = Cfg::M, WARPS = PkPairCfg::WARPS, XCH = Cfg::XCH;
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    f2* s_win = reinterpret_cast<f2*>(sm + PkPairCfg::off_win);
    f2* s_tw2 = reinterpret_cast<f2*>(sm + PkPairCfg::off_tw2);
    f2* s_twP = reinterpret_cast<f2*>(sm + PkPairCfg::off_twP);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + PkPairCfg::off_pal);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(sm + PkPairCfg::off_bar(P.npal));
    f2* s_xch = reinterpret_cast<f2*>(sm + PkPairCfg::off_xch(P.npal));

    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const int s = i & 31, n1 = i >> 5;
        s_win[s * Cfg::ROW + n1] = pk(P.window[2 * i], P.window[2 * i + 1]);
    }
    if (threadIdx.x < WARPS) mbar_init(s_bar + threadIdx.x, 1);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        const int s = i & 31, q = i >> 5;
        const cpx t = P.twP[2 * tw2_exponent(s, q)];
        s_tw2[s * Cfg::TROW + q] = pk(t.x, t.y);
        const cpx w = P.twP[s + 32 * q];
        s_twP[s * Cfg::PROW + q] = pk(w.y, -w.x);
    }
    for (int i = threadIdx.x; i <= P.npal; i += blockDim.x) s_pal[i] = P.palette[i < P.npal ? i : P.ci_hi];
    __syncthreads();
    grid_dep_wait();

    const int s = threadIdx.x & 31, warp = threadIdx.x >> 5;
    f2* xa = s_xch + warp * 2 * XCH; // buffer of channel 0
    f2* xb = xa + XCH;               //           channel 1
    unsigned long long* bar = s_bar + warp;
    unsigned copies = 0;
    const f2x2* wrow = reinterpret_cast<const f2x2*>(s_win + s * Cfg::ROW);
    const f2x2* trow = reinterpret_cast<const f2x2*>(s_tw2 + s * Cfg::TROW);
    const f2x2* prow = reinterpret_cast<const f2x2*>(s_twP + s * Cfg::PROW);
    const int partner = (32 - s) & 31;

    const unsigned total = (unsigned)P.ncols * (unsigned)P.nstreams;
    const unsigned gstep = gridDim.x * WARPS;
    constexpr float scale = 0.5f;
    auto stage = [&](unsigned g) { // both channels of frame g -> the two buffers (lane 0, after a __syncwarp())
        const PkUnit un = pk_unit(P, g);
        const float* a = P.samples + un.stream * P.stream_stride + un.st;
        if (s == 0) bulk_copy2_g2s(xa, a, xb, a + P.channel_stride, M * 8, bar);
#if defined(JADE_EMU)
        __syncwarp();
#endif
    };
    auto emit = [&](float p, uint32_t* pix, float* db) { emit_bin1<MIX_SUM, WANT_DB, false, false>(p, scale, pix, db, P, s_pal); };

    unsigned g = blockIdx.x * WARPS + warp;
    if (g < total) stage(g);
    for (; g < total; g += gstep) {
        const PkUnit un = pk_unit(P, g);
        const ColOut o = col_out(P, un.stream, un.j);

        f2 va[32], vb[32];
#if defined(JADE_ABL_NOSTAGE)
        if (copies == 0)
#endif
        mbar_wait(bar, copies & 1u);
        ++copies;
#pragma unroll
        for (int jj = 0; jj < 16; jj += 2) {
#if defined(JADE_ABL_NOWIN)
            const f2x2 wa = wrow[0], wb = wrow[1];
#else
            const f2x2 wa = wrow[jj / 2], wb = wrow[(jj + 16) / 2];
#endif
            const f2* za = xa + s;
            const f2* zb = xb + s;
#if defined(JADE_ABL_NOLOAD)
#define ABL_LD(z, i) z[32 * ((i) & 1)]
#else
#define ABL_LD(z, i) z[32 * (i)]
#endif
            win_stage1(va, jj, ABL_LD(za, jj), wa.a, ABL_LD(za, jj + 16), wb.a);
            win_stage1(va, jj + 1, ABL_LD(za, jj + 1), wa.b, ABL_LD(za, jj + 17), wb.b);
            win_stage1(vb, jj, ABL_LD(zb, jj), wa.a, ABL_LD(zb, jj + 16), wb.a);
            win_stage1(vb, jj + 1, ABL_LD(zb, jj + 1), wa.b, ABL_LD(zb, jj + 17), wb.b);
        }
        __syncwarp(); // every lane has read its samples before the transposes overwrite them
#if !defined(JADE_ABL_NOFP1)
        fft32_pk_after_stage1(va);
        fft32_pk_after_stage1(vb);
#endif
        f2 ua[32], ub[32];
#if defined(JADE_ABL_NOXPOSE)
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) {
            ua[k1] = va[k1];
            ub[k1] = vb[k1];
        }
#else
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) {
            xa[k1 * Cfg::XROW + s] = va[k1];
            xb[k1 * Cfg::XROW + s] = vb[k1];
        }
        __syncwarp();
        {
            const f2x2* ra = reinterpret_cast<const f2x2*>(xa + s * Cfg::XROW);
            const f2x2* rb = reinterpret_cast<const f2x2*>(xb + s * Cfg::XROW);
#pragma unroll
            for (int jx = 0; jx < 32; jx += 2) {
                const f2x2 ta = ra[jx / 2], tb = rb[jx / 2];
                ua[brev(jx, 5)] = ta.a;
                ua[brev(jx + 1, 5)] = ta.b;
                ub[brev(jx, 5)] = tb.a;
                ub[brev(jx + 1, 5)] = tb.b;
            }
        }
#endif
        __syncwarp(); // the buffers are free again
#if !defined(JADE_ABL_NOSTAGE)
        if (g + gstep < total) stage(g + gstep); // stage the next frame of this warp
#endif
#if !defined(JADE_ABL_NOFP2)
        fft32_twisted2(ua, ub, trow);
#endif

        uint32_t* p_lo = o.pix ? o.pix + (M - s) : nullptr; // bin k -> row M - k
        uint32_t* p_hi = o.pix ? o.pix + s : nullptr;
        float* d_lo = (WANT_DB && o.db) ? o.db + s : nullptr;
        float* d_hi = (WANT_DB && o.db) ? o.db + (M - s) : nullptr;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
#if defined(JADE_ABL_NOSHFL)
            const f2 zpa = ua[31 - q], zpb = ub[31 - q];
#else
            const f2 zpa = sel2(s == 0, ua[(32 - q) & 31], shfl2(ua[31 - q], partner));
            const f2 zpb = sel2(s == 0, ub[(32 - q) & 31], shfl2(ub[31 - q], partner));
#endif
#if defined(JADE_ABL_NOTW)
            const f2x2 wq2 = prow[0];
#else
            const f2x2 wq2 = prow[q / 2];
#endif
            const f2 wq = (q & 1) ? wq2.b : wq2.a;
            const f2 Aa = add2(ua[q], conj2(zpa)), Ba = sub2(ua[q], conj2(zpa));
            const f2 Ab = add2(ub[q], conj2(zpb)), Bb = sub2(ub[q], conj2(zpb));
            const f2 Ta = cmul2(Ba, wq), Tb = cmul2(Bb, wq);
            const f2 xpa = add2(Aa, Ta), xma = sub2(Aa, Ta);
            const f2 xpb = add2(Ab, Tb), xmb = sub2(Ab, Tb);
            // channel 0 first, channel 1 on top: the accumulation order of stft_pk2048_kernel
            const float pl = fm(lo(xpb), lo(xpb), fm(hi(xpb), hi(xpb), fm(lo(xpa), lo(xpa), fm(hi(xpa), hi(xpa), 0.f))));
            const float ph = fm(lo(xmb), lo(xmb), fm(hi(xmb), hi(xmb), fm(lo(xma), lo(xma), fm(hi(xma), hi(xma), 0.f))));
            emit(pl, (!WANT_DB || p_lo) ? p_lo - 32 * q : nullptr, d_lo ? d_lo + 32 * q : nullptr);
            emit(ph, (!WANT_DB || p_hi) ? p_hi + 32 * q : nullptr, d_hi ? d_hi - 32 * q : nullptr);
        }
        if (s == 0) { // bin 512 (lane 0, self-paired): X = 2 conj Z
            const float a0 = lo(ua[16]), a1 = hi(ua[16]), b0 = lo(ub[16]), b1 = hi(ub[16]);
            const float pm = fm(JADE_FMUL(4.0f, b0), b0, fm(JADE_FMUL(4.0f, b1), b1, fm(JADE_FMUL(4.0f, a0), a0, fm(JADE_FMUL(4.0f, a1), a1, 0.f))));
            emit(pm, (!WANT_DB || o.pix) ? o.pix + 512 : nullptr, (WANT_DB && o.db) ? o.db + 512 : nullptr);
        }
    }
}

} // namespace jade
