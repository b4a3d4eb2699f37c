// jade_pk.cuh -- the headline kernel: N = 2048 (BASELINE configs[1] / [3]) with packed FP32x2 arithmetic.
//
// Same fused path as jade_kernels.cuh (framing, window, real FFT, |X|^2, channel mix, dB, flip, palette, packed pixel
// store; reference lines cited there), restructured around what limits it on sm_100a:
//   * every complex value lives in ONE 64-bit register pair and all butterflies / twiddle products are FFMA2 / FADD2 /
//     FMUL2 (PTX fma/add/mul.rn.f32x2): a complex add is 1 instruction, a complex multiply 2, a general radix-2
//     butterfly 3 (a' = a + W b by two chained FFMA2, b' = 2a - a').  ptxas folds the half-swap, the per-half sign
//     and the scalar broadcast into operand modifiers (R.F32x2.LO_HI.NP, R.F32), so no repacking moves are needed.
//     This halves the issue slots of the FP32 work, which was the measured limiter (profiles/r01b: issue-active 64 %).
//   * the real-FFT split is done per PAIR (k, M-k): A = Z[k]+conj Z[M-k], B = Z[k]-conj Z[M-k], T = (-i W_N^k) B,
//     X[k] = A+T, X[M-k] = conj(A-T): 8 FP32 instructions per bin instead of 13, and only the upper half of the
//     spectrum crosses lanes (17 shared-memory words per lane instead of 33+32).
//   * per-lane tables (window, inter-pass twiddles, split twiddles) are laid out [lane][index] with a 16-byte row pad, so
//     that they are read with conflict-free LDS.128 (two table entries per instruction).
//
// One warp transforms one frame: M = 1024 complex points z[m] = x[2m] + i x[2m+1], 32 per lane (m = s + 32 n1), radix-32
// in registers, one padded transpose through shared memory, radix-32 in registers: lane s ends with Z[s + 32 k2].
#pragma once
#include "jade_kernels.cuh"

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
struct PkCfg {
    static constexpr int M = 1024, N = 2048, B = 1025;
    static constexpr int WARPS = 8;
    static constexpr int ROW = 34;   // f2 words per lane row of the window / inter-pass twiddle tables (32 + 16 B pad)
    static constexpr int PROW = 18;  // f2 words per lane row of the split-twiddle table (16 + 16 B pad)
    static constexpr int XCH = 32 * 33; // f2 words per warp exchange buffer (transpose: 32 rows of 33)
    static constexpr int off_win = 0;
    static constexpr int off_twI = off_win + 32 * ROW * 8;
    static constexpr int off_twP = off_twI + 32 * ROW * 8;
    static constexpr int off_pal = off_twP + 32 * PROW * 8;
    static JADE_HD int off_xch(int npal) { return off_pal + (npal * 4 + 15) / 16 * 16; }
    static JADE_HD int smem_bytes(int npal) { return off_xch(npal) + WARPS * XCH * 8; }
};

// GUARD = false: frames must lie entirely inside [0, nsamples) and start on an even sample (8-byte aligned float2
// loads).  The host (launch_stft in jade_gpu.cu) sends the few columns that touch the signal boundary, and unaligned
// geometries, to the GUARD = true instantiation: bounds-checked scalar loads, bit-identical arithmetic afterwards (so
// streaming, batch and sharded renderings of the same column agree bit for bit whichever instantiation produced it).
// MIXK: MIX_NONE (one contributing channel) or MIX_SUM (AbsMean over 2^n channels).  Identity rows in the reference
// orientation, hardware log2 for the dB.  WANT_DB additionally stores the float dB column (streaming ring / getMem).
template <int MIXK, bool WANT_DB, bool GUARD>
JADE_KERNEL(PkCfg::WARPS * 32, 2) stft_pk2048_kernel(const KParams P)
{
    using Cfg = PkCfg;
    constexpr int M = Cfg::M;
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    f2* s_win = reinterpret_cast<f2*>(sm + Cfg::off_win);
    f2* s_twI = reinterpret_cast<f2*>(sm + Cfg::off_twI);
    f2* s_twP = reinterpret_cast<f2*>(sm + Cfg::off_twP);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);
    f2* s_xch = reinterpret_cast<f2*>(sm + Cfg::off_xch(P.npal));

    // ---- per-lane tables: entry for (lane s, index i) at [s*ROW + i]
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const int s = i & 31, n1 = i >> 5;                       // m = s + 32 n1
        s_win[s * Cfg::ROW + n1] = pk(P.window[2 * i], P.window[2 * i + 1]);
        const cpx t = P.twI[n1 * 32 + s];                        // W_1024^(k1 s), k1 = n1 here
        s_twI[s * Cfg::ROW + n1] = pk(t.x, t.y);
    }
    for (int i = threadIdx.x; i < 512; i += blockDim.x) {
        const int s = i & 31, q = i >> 5;                        // k = s + 32 q
        const cpx w = P.twP[s + 32 * q];                         // W_N^k ; table holds -i W_N^k = (w.y, -w.x)
        s_twP[s * Cfg::PROW + q] = pk(w.y, -w.x);
    }
    for (int i = threadIdx.x; i < P.npal; i += blockDim.x) s_pal[i] = P.palette[i];
    __syncthreads();

    const int s = threadIdx.x & 31, warp = threadIdx.x >> 5;
    f2* xw = s_xch + warp * Cfg::XCH;
    const f2x2* wrow = reinterpret_cast<const f2x2*>(s_win + s * Cfg::ROW);
    const f2x2* trow = reinterpret_cast<const f2x2*>(s_twI + s * Cfg::ROW);
    const f2x2* prow = reinterpret_cast<const f2x2*>(s_twP + s * Cfg::PROW);
    f2* tr_wr = xw + s;            // transpose: word k1*33 + s
    const f2* tr_rd = xw + s * 33; //            row s, words j
    f2* ex_wr = xw + s;            // exchange: word (i-16)*32 + s holds Z[s + 32 i], i = 16..31; row 16 holds Z[s]
    const f2* ex_rd = xw + (32 - s); // partner of k = s + 32 q is word (15-q)*32 + (32-s)   (lane 0: its own column)

    const unsigned total = (unsigned)P.ncols * (unsigned)P.nstreams;
    int ch0, ch1;
    channel_range(P, ch0, ch1);
    const float scale = (MIXK == MIX_SUM) ? (1.0f / (float)P.channels) : 1.0f;

    for (unsigned g = blockIdx.x * Cfg::WARPS + warp; g < total; g += gridDim.x * Cfg::WARPS) {
        const int stream = (int)(g / (unsigned)P.ncols);
        const long long j = P.first_col + (g - (unsigned)stream * (unsigned)P.ncols);
        const long long st = frame_start(P, j);

        float alo[16], ahi[16], amid = 0.f; // power of bins s+32q / M-(s+32q) / 512 (lane 0)
#pragma unroll
        for (int q = 0; q < 16; ++q) alo[q] = ahi[q] = 0.f;

        for (int ch = ch0; ch < (MIXK == MIX_NONE ? ch0 + 1 : ch1); ++ch) {
            const float* x = P.samples + stream * P.stream_stride + ch * P.channel_stride;
            f2 v[32];
#pragma unroll
            for (int j = 0; j < 16; j += 2) { // n1 = j, j+1 paired with n1 + 16
                const f2x2 wa = wrow[j / 2], wb = wrow[(j + 16) / 2];
                f2 xa0, xa1, xb0, xb1;
                if (!GUARD) {
                    const f2* xz = reinterpret_cast<const f2*>(x + st) + s;
                    xa0 = xz[32 * j];
                    xa1 = xz[32 * (j + 1)];
                    xb0 = xz[32 * (j + 16)];
                    xb1 = xz[32 * (j + 17)];
                } else {
                    const cpx a0 = load_pair_guarded(x, st + 2 * (s + 32 * j), P.nsamples);
                    const cpx a1 = load_pair_guarded(x, st + 2 * (s + 32 * (j + 1)), P.nsamples);
                    const cpx b0 = load_pair_guarded(x, st + 2 * (s + 32 * (j + 16)), P.nsamples);
                    const cpx b1 = load_pair_guarded(x, st + 2 * (s + 32 * (j + 17)), P.nsamples);
                    xa0 = pk(a0.x, a0.y);
                    xa1 = pk(a1.x, a1.y);
                    xb0 = pk(b0.x, b0.y);
                    xb1 = pk(b1.x, b1.y);
                }
                win_stage1(v, j, xa0, wa.a, xb0, wb.a);
                win_stage1(v, j + 1, xa1, wa.b, xb1, wb.b);
            }
            fft32_pk_after_stage1(v);
#pragma unroll
            for (int k1 = 0; k1 < 32; k1 += 2) {
                const f2x2 t = trow[k1 / 2];
                tr_wr[k1 * 33] = (k1 == 0) ? v[0] : cmul2(v[k1], t.a);
                tr_wr[(k1 + 1) * 33] = cmul2(v[k1 + 1], t.b);
            }
            __syncwarp();
            f2 u[32];
#pragma unroll
            for (int jx = 0; jx < 32; ++jx) u[brev(jx, 5)] = tr_rd[jx];
            fft32_pk(u); // u[k2] = Z[s + 32 k2]
            __syncwarp();
#pragma unroll
            for (int i = 16; i < 32; ++i) ex_wr[(i - 16) * 32] = u[i];
            ex_wr[16 * 32] = u[0];
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const f2 zp = ex_rd[(15 - q) * 32];
                const f2x2 wq = prow[q / 2];
                const f2 A = add2(u[q], conj2(zp));  // Z[k] + conj Z[M-k]
                const f2 Bv = sub2(u[q], conj2(zp)); // Z[k] - conj Z[M-k]
                const f2 T = cmul2(Bv, (q & 1) ? wq.b : wq.a);
                const f2 xp = add2(A, T), xm = sub2(A, T);
                alo[q] = fm(lo(xp), lo(xp), fm(hi(xp), hi(xp), alo[q]));
                ahi[q] = fm(lo(xm), lo(xm), fm(hi(xm), hi(xm), ahi[q]));
            }
            { // bin 512 (lane 0, self-paired): X = 2 conj Z
                const float a = lo(u[16]), b = hi(u[16]);
                amid = fm(JADE_FMUL(4.0f, a), a, fm(JADE_FMUL(4.0f, b), b, amid));
            }
            __syncwarp();
        }

        const ColOut o = col_out(P, stream, j);
        // reference orientation: bin k lands in row M - k.  lane s: bins s+32q -> rows M-s-32q ; bins M-s-32q -> rows s+32q
        uint32_t* p_lo = o.pix ? o.pix + (M - s) : nullptr;
        uint32_t* p_hi = o.pix ? o.pix + s : nullptr;
        float* d_lo = (WANT_DB && o.db) ? o.db + s : nullptr;
        float* d_hi = (WANT_DB && o.db) ? o.db + (M - s) : nullptr;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const float dl = to_db_fast(MIXK == MIX_SUM ? JADE_FMUL(alo[q], scale) : alo[q]);
            const float dh = to_db_fast(MIXK == MIX_SUM ? JADE_FMUL(ahi[q], scale) : ahi[q]);
            if (WANT_DB && d_lo) {
                d_lo[32 * q] = dl;
                d_hi[-32 * q] = dh;
            }
            if (p_lo) {
                p_lo[-32 * q] = colour_of(dl, P, s_pal);
                p_hi[32 * q] = colour_of(dh, P, s_pal);
            }
        }
        if (s == 0) {
            const float dm = to_db_fast(MIXK == MIX_SUM ? JADE_FMUL(amid, scale) : amid);
            if (WANT_DB && o.db) o.db[512] = dm;
            if (o.pix) o.pix[512] = colour_of(dm, P, s_pal);
        }
    }
}

} // namespace jade
