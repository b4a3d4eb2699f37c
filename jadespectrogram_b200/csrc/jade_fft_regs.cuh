// jade_fft_regs.cuh -- register-resident complex FFT building blocks (sm_100a).
//
// Replaces the reference's external FFT primitive `spectrum::power` (call site Spectrogram.cpp:144; class
// declared in the absent TGM "FFT.h", Spectrogram.h:16,157).  Everything is fully unrolled so every array
// index is a compile-time constant and the arrays live in registers.
//
// The file is plain C++ apart from the JADE_DEVICE qualifiers so that tests/emu can execute the very same
// code on the CPU (kernel-logic debugging aid, test infrastructure only).
#pragma once

#if defined(__CUDACC__)
#define JADE_DEVICE __device__ __forceinline__
#define JADE_HD __host__ __device__ __forceinline__
#else
#define JADE_DEVICE inline
#define JADE_HD inline
#endif

namespace jade {

struct alignas(8) cpx { // 8-byte aligned so that loads/stores are single 64-bit accesses (LDG.64 / LDS.64)
    float x, y;
};

JADE_HD cpx mk(float x, float y)
{
    cpx r;
    r.x = x;
    r.y = y;
    return r;
}

// fused multiply-add that is an FFMA on the device and std::fma on the host emulator
JADE_HD float fm(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return __builtin_fmaf(a, b, c);
#endif
}

// a * w  (complex)
JADE_HD cpx cmul(cpx a, cpx w) { return mk(fm(a.x, w.x, -(a.y * w.y)), fm(a.x, w.y, a.y * w.x)); }

// cos(2*pi*m/32) for m in [0,8]
JADE_HD constexpr float cos32_q(int m)
{
    return m == 0 ? 1.0f
         : m == 1 ? 0.98078528040323044913f
         : m == 2 ? 0.92387953251128675613f
         : m == 3 ? 0.83146961230254523708f
         : m == 4 ? 0.70710678118654752440f
         : m == 5 ? 0.55557023301960222474f
         : m == 6 ? 0.38268343236508977173f
         : m == 7 ? 0.19509032201612826785f
                  : 0.0f;
}
// cos / sin of 2*pi*m/32 for m in [0,16]
JADE_HD constexpr float cos32(int m) { return m <= 8 ? cos32_q(m) : -cos32_q(16 - m); }
JADE_HD constexpr float sin32(int m) { return m <= 8 ? cos32_q(8 - m) : cos32_q(m - 8); }

// cos(2 pi m / 64), m = 0..16
JADE_HD constexpr float cos64_q(int m)
{
    return m == 0 ? 1.0f
         : m == 1 ? 0.99518472667219692873f
         : m == 2 ? 0.98078528040323043058f
         : m == 3 ? 0.95694033573220882438f
         : m == 4 ? 0.92387953251128673848f
         : m == 5 ? 0.88192126434835504956f
         : m == 6 ? 0.83146961230254523567f
         : m == 7 ? 0.77301045336273699338f
         : m == 8 ? 0.70710678118654757274f
         : m == 9 ? 0.63439328416364548779f
         : m == 10 ? 0.55557023301960228867f
         : m == 11 ? 0.47139673682599780857f
         : m == 12 ? 0.38268343236508983729f
         : m == 13 ? 0.29028467725446233105f
         : m == 14 ? 0.19509032201612833135f
         : m == 15 ? 0.09801714032956077016f
                   : 0.0f;
}
JADE_HD constexpr float cos64(int m) { return m <= 16 ? cos64_q(m) : -cos64_q(32 - m); }
JADE_HD constexpr float sin64(int m) { return m <= 16 ? cos64_q(16 - m) : cos64_q(m - 16); }

// log2 of a power of two <= 32 and bit reversal of a `bits`-bit index; written without recursion or loops so that
// they fold to constants after loop unrolling (register arrays must only ever be indexed by constants).
JADE_HD constexpr int ilog2c(int n) { return n >= 32 ? 5 : n >= 16 ? 4 : n >= 8 ? 3 : n >= 4 ? 2 : n >= 2 ? 1 : 0; }
JADE_HD constexpr int brev5(int v) { return ((v & 1) << 4) | ((v & 2) << 2) | (v & 4) | ((v & 8) >> 2) | ((v & 16) >> 4); }
JADE_HD constexpr int brev(int v, int bits) { return brev5(v) >> (5 - bits); }

// One radix-2 decimation-in-time butterfly with twiddle W = exp(-2*pi*i*M32/32):
//   a' = a + W b ; b' = a - W b
// General twiddles use the 6-FMA form (b' = 2a - a'); W = 1 and W = -i need 4 adds.
template <int M32>
JADE_DEVICE void bfly(cpx& a, cpx& b)
{
    if (M32 == 0) {
        const cpx t = b;
        b = mk(a.x - t.x, a.y - t.y);
        a = mk(a.x + t.x, a.y + t.y);
    } else if (M32 == 8) { // W = -i : W b = (b.y, -b.x)
        const cpx t = b;
        b = mk(a.x - t.y, a.y + t.x);
        a = mk(a.x + t.y, a.y - t.x);
    } else {
        constexpr float c = cos32(M32);
        constexpr float s = sin32(M32); // W = c - i s
        // W b = (c b.x + s b.y, c b.y - s b.x)
        const float nr = fm(s, b.y, fm(c, b.x, a.x));
        const float ni = fm(-s, b.x, fm(c, b.y, a.y));
        b = mk(fm(2.0f, a.x, -nr), fm(2.0f, a.y, -ni));
        a = mk(nr, ni);
    }
}

// In-place R-point DFT (R = 1,2,4,8,16,32) on a[0..R-1] (unit stride).
// Input must be supplied in BIT-REVERSED order (a[brev(n)] = x[n]); output is in natural order.
template <int LEN, int BASE, int J>
JADE_DEVICE void dit_inner(cpx* a)
{
    if constexpr (J < LEN / 2) {
        bfly<(J * 32) / LEN>(a[BASE + J], a[BASE + J + LEN / 2]);
        dit_inner<LEN, BASE, J + 1>(a);
    }
}
template <int R, int LEN, int BASE>
JADE_DEVICE void dit_blocks(cpx* a)
{
    if constexpr (BASE < R) {
        dit_inner<LEN, BASE, 0>(a);
        dit_blocks<R, LEN, BASE + LEN>(a);
    }
}
template <int R, int LEN>
JADE_DEVICE void dit_stages(cpx* a)
{
    if constexpr (LEN <= R) {
        dit_blocks<R, LEN, 0>(a);
        dit_stages<R, LEN * 2>(a);
    }
}

template <int R>
JADE_DEVICE void fft_dit(cpx* a)
{
    static_assert(R >= 1 && R <= 32 && (R & (R - 1)) == 0, "radix must be a power of two <= 32");
    dit_stages<R, 2>(a);
}

} // namespace jade
