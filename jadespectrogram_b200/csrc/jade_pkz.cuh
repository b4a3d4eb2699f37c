// jade_pkz.cuh -- N = 2048, AbsMean over exactly two channels (BASELINE configs[1], the plugin's live default): both channels
// of a frame in ONE 2048-point complex transform.
//
// The reference computes each channel's power spectrum and averages them (Spectrogram.cpp:52-58,67-74).  With
//     z[n] = w[n] (x_L[n] + i x_R[n]),   Z = DFT_2048(z):     X_L[k] = (Z[k] + conj Z[N-k]) / 2,   X_R[k] = (Z[k] - conj Z[N-k]) / 2i
//     |X_L[k]|^2 + |X_R[k]|^2 = ( |Z[k]|^2 + |Z[N-k]|^2 ) / 2
// so the two real transforms, their pair splits (96 packed instructions and 32 64-bit shuffles per channel in
// stft_pk2048x2_kernel, jade_pk.cuh) and the split-twiddle table disappear for one extra radix-2 stage, and only |Z|^2
// is ever needed.  The device window table carries the factor 1/2 (upload_window), hence mean power = |Z'[k]|^2 + |Z'[N-k]|^2.
//
// One warp transforms one stereo frame: n = s + 32 n1 (lane s, n1 = 0..63), 64-point DFT over n1 in registers (window
// fused into its first stage), one transpose through shared memory, twisted 32-point DFTs over s (inter-pass twiddle
// W_2048^{s k1} folded into the butterflies, fft32_twisted).  Lane l takes the two rows k1 = l and k1 = 64 - l (lane 0:
// rows 0 and 32): bin k = k1 + 64 k2 and its mirror N - k = (64 - k1) + 64 (31 - k2) then sit in the SAME lane, so the
// mirror sum needs no shuffle at all -- nothing crosses lanes after the transpose.  Lane l emits the bins l + 64 k2 and
// (64 - l) + 64 k2, k2 = 0..15: per k2 the warp stores two runs of 32 consecutive rows.
//
// Epilogue per bin: log2 (MUFU), one FFMA to the palette position, F2I, ONE integer clamp to [0, npal] (entry npal of the
// shared-memory table is the colour of the `value >= m_Max` rule), palette load, store; the + 1e-11 rides on the power FMAs.
// (Tried and dropped, profiles/r02_pkz_variants.txt: interleaving this epilogue with the next frame's loads in one basic
// block, -4 %; passing an "FP token" round-robin between the three warps of a scheduler through named barriers so that
// only one of them runs a butterfly pass at a time, -6 %.)
//
// Interior, 16-byte aligned frames are staged by the TMA engine (cp.async.bulk, two 8 KB copies on one mbarrier) into
// the warp's buffer -- the buffer the transpose went through a moment before; PKZ_GUARD instantiations read global
// memory with bounds checks (boundary columns, unaligned geometries) and run the same arithmetic, so streaming, batch
// and sharded renderings agree bit for bit.
#pragma once
#include "jade_pk.cuh"
#include "jade_tmem.cuh"

// JADE_PKZ_TMEM = 1: the per-lane window and twisted-twiddle tables live in tensor memory (jade_tmem.cuh) instead of shared memory
#ifndef JADE_PKZ_TMEM
#define JADE_PKZ_TMEM 1
#endif

namespace jade {

JADE_HD constexpr int brev6(int v) { return (brev5(v & 31) << 1) | (v >> 5); }

// radix-2 DIT butterfly with W = exp(-2 pi i M64/64), M64 in [0, 32)
template <int M64>
JADE_DEVICE void bfly2_64(f2& a, f2& b)
{
    if (M64 == 0) {
        const f2 t = b;
        b = sub2(a, t);
        a = add2(a, t);
    } else if (M64 == 16) {
        const f2 t = mul_mi(b);
        b = sub2(a, t);
        a = add2(a, t);
    } else {
        constexpr float c = cos64(M64);
        constexpr float s = sin64(M64);
        const f2 n = fma2(mul_mi(b), pk(s, s), fma2(b, pk(c, c), a));
        b = fma2(a, pk(2.0f, 2.0f), neg2(n));
        a = n;
    }
}
template <int LEN, int BASE, int J>
JADE_DEVICE void pk64_inner(f2* a)
{
    if constexpr (J < LEN / 2) {
        bfly2_64<(J * 64) / LEN>(a[BASE + J], a[BASE + J + LEN / 2]);
        pk64_inner<LEN, BASE, J + 1>(a);
    }
}
template <int LEN, int BASE>
JADE_DEVICE void pk64_blocks(f2* a)
{
    if constexpr (BASE < 64) {
        pk64_inner<LEN, BASE, 0>(a);
        pk64_blocks<LEN, BASE + LEN>(a);
    }
}
// stages 2..6 of the in-place 64-point DFT (bit-reversed input, natural-order output); stage 1 is win_stage1_64
JADE_DEVICE void fft64_pk_after_stage1(f2* a)
{
    pk64_blocks<4, 0>(a);
    pk64_blocks<8, 0>(a);
    pk64_blocks<16, 0>(a);
    pk64_blocks<32, 0>(a);
    pk64_blocks<64, 0>(a);
}
// window multiply fused with stage 1: pairs n1 = j and j + 32 (see win_stage1); the window is real here, so it is a
// broadcast scalar operand
JADE_DEVICE void win_stage1_64(f2* v, int j, f2 xa, float wa, f2 xb, float wb)
{
    const int i = brev5(j);
    const f2 va = mul2(xa, pk(wa, wa));
    v[2 * i] = fma2(xb, pk(wb, wb), va);
    v[2 * i + 1] = fma2(neg2(xb), pk(wb, wb), va);
}

// exponent e of twisted-table entry t (0..15) of row k1 (0..63): the entry is W_2048^e  (cf. tw2_exponent)
JADE_HD int twz_exponent(int k1, int t)
{
    if (t == 0) return 16 * k1;
    if (t == 1) return 8 * k1;
    if (t < 4) return 4 * (k1 + 64 * (t - 2));
    if (t < 8) return 2 * (k1 + 64 * (t - 4));
    return k1 + 64 * (t - 8);
}
// second row of lane l
JADE_HD int pkz_row_b(int l) { return l == 0 ? 32 : 64 - l; }

struct PkzCfg {
    static constexpr int N = 2048, B = 1025;
#ifndef JADE_PKZ_WARPS
#define JADE_PKZ_WARPS 12
#endif
    static constexpr int WARPS = JADE_PKZ_WARPS;
    static constexpr int WROW = 68;   // floats per lane row of the window table (64 + 16 B pad): conflict-free LDS.128
    static constexpr int TROW = 18;   // f2 words per lane row of a twisted table (16 + 16 B pad)
    static constexpr int XROW = 34;   // f2 words per transpose row
    static constexpr int XCH = 64 * XROW; // f2 words per warp buffer (17 408 B): transpose, and landing area of the next frame
    static constexpr int CH1 = 32 * XROW; // f2 offset of channel 1's 8 KB inside the buffer
    static constexpr bool TM = JADE_PKZ_TMEM != 0;
    static constexpr int TM_COLS = 512; // tensor-memory columns: 0..31 twisted table of row a, 32..63 of row b, 64..127 window; 128 (1 + warp / 4) ..: that warp's
                                        // sample ring (the warps w, w + 4, w + 8 share the 32 lanes of quadrant w % 4)
    static_assert(WARPS <= 12, "one 128-column sample ring per warp of a quadrant");
    static constexpr int off_win = 0;
    static constexpr int off_twa = off_win + (TM ? 0 : 32 * WROW * 4);
    static constexpr int off_twb = off_twa + (TM ? 0 : 32 * TROW * 8);
    static constexpr int off_pal = off_twb + (TM ? 0 : 32 * TROW * 8);
    static JADE_HD int off_bar(int npal) { return off_pal + ((npal + 1) * 4 + 15) / 16 * 16; } // table + the `>= m_Max` entry
    static JADE_HD int off_xch(int npal) { return off_bar(npal) + (WARPS * 8 + 8 + 15) / 16 * 16; } // mbarriers + the TMEM address
    static JADE_HD int smem_bytes(int npal) { return off_xch(npal) + WARPS * XCH * 8; }
};

// PKZ_ASYNC: interior aligned frames staged by the TMA engine, frames dealt round-robin to the warps;  PKZ_GUARD: bounds-checked
// global loads;  PKZ_RING: TMA-staged like PKZ_ASYNC, but every warp takes a contiguous run of columns and keeps the frame's
// samples in a ring of four 512-sample chunks in tensor memory, so that a frame whose start lies one chunk after its
// predecessor's reads only its NEW chunk from shared memory (evenly spaced columns with hop = N/4, the batch renderer's case).
enum { PKZ_ASYNC = 0, PKZ_GUARD = 1, PKZ_RING = 2 };

// Epilogue of one bin (cf. emit_bin): p already carries the reference's + 1e-11 (Spectrogram.cpp:36,107).  lg = log2 p; dB value
// (optional) = 3.0103 lg; palette index = trunc(lg ck1 + ck0) clamped to [0, npal] by ONE integer min/max, where entry
// npal of the shared-memory table holds the colour of CColorPalette's `value >= m_Max` rule (index of 0.9999 m_Max,
// CColorpalette.h:34-35): lg ck1 + ck0 >= npal exactly when the dB value reaches m_Max.
template <bool WANT_DB, bool U8 = false>
JADE_DEVICE void pkz_emit(float p, uint32_t* pix, float* db, const KParams& P, const uint32_t* pal)
{
    emit_bin1<MIX_NONE, WANT_DB, U8, true>(p, 1.0f, pix, db, P, pal);
}

// (stream, column) walked incrementally: g -> g + gstep without a division per frame
struct PkzWalk {
    unsigned stream, col, dq, dr;
    template <bool UNIT> // UNIT: gstep == 1 (a contiguous run): no quotient / remainder to carry
    JADE_DEVICE void init(unsigned g, unsigned gstep, unsigned nc)
    {
        stream = g / nc;
        col = g - stream * nc;
        dq = UNIT ? 0u : gstep / nc;
        dr = UNIT ? 1u : gstep - dq * nc;
    }
    template <bool UNIT>
    JADE_DEVICE void next(unsigned nc)
    {
        if (UNIT) {
            if (++col == nc) {
                col = 0;
                ++stream;
            }
        } else {
            stream += dq;
            col += dr;
            if (col >= nc) {
                col -= nc;
                ++stream;
            }
        }
    }
};

// U8: instantiation for P.pal_u8 palettes (colour_of_lg1: the float -> u8 conversion is the clamp; identical colours), picked by
// launch_one for the pixel-only launches -- a launch-uniform branch around two epilogue copies costs registers this kernel
// does not have (-1.5 %, gpurun_out/ab8.txt)
template <bool WANT_DB, int LD, bool U8 = false>
JADE_KERNEL(PkzCfg::WARPS * 32, 1) stft_pkz2048_kernel(const KParams P)
{
    using Cfg = PkzCfg;
    constexpr int WARPS = Cfg::WARPS;
    constexpr bool STAGED = LD != PKZ_GUARD, RING = LD == PKZ_RING;
    static_assert(!RING || Cfg::TM, "the sample ring lives in tensor memory");
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    float* s_win = reinterpret_cast<float*>(sm + Cfg::off_win);
    f2* s_twa = reinterpret_cast<f2*>(sm + Cfg::off_twa);
    f2* s_twb = reinterpret_cast<f2*>(sm + Cfg::off_twb);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(sm + Cfg::off_bar(P.npal));
    f2* s_xch = reinterpret_cast<f2*>(sm + Cfg::off_xch(P.npal));

    // ---- per-lane tables
    if (STAGED && threadIdx.x < WARPS) mbar_init(s_bar + threadIdx.x, 1);
    for (int i = threadIdx.x; i <= P.npal; i += blockDim.x) s_pal[i] = P.palette[i < P.npal ? i : P.ci_hi];
    uint32_t tq = 0; // tensor-memory address of this warp's quadrant of the tables
    if constexpr (Cfg::TM) {
        uint32_t* s_tm = reinterpret_cast<uint32_t*>(s_bar + WARPS);
        const int l = threadIdx.x & 31;
        if (threadIdx.x < 32) tm_alloc(s_tm, Cfg::TM_COLS);
        tm_fence_before_sync();
        __syncthreads();
        tm_fence_after_sync();
        tq = tm_quadrant_base(*s_tm);
        if (threadIdx.x < 128) { // warp q fills quadrant q: thread l writes lane 32 q + l
            uint32_t r[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) { // columns 16 c ..: entries t = 8 (c & 1) .. + 7 of row a (c < 2) / row b
                const int k1 = c < 2 ? l : pkz_row_b(l);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const cpx a = P.twP[twz_exponent(k1, 8 * (c & 1) + i)]; // W_2048^e, e < 1024
                    r[2 * i] = f2u(a.x);
                    r[2 * i + 1] = f2u(a.y);
                }
                tm_st<16>(tq + 16 * c, r);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) { // columns 64 + 16 c ..: window of n1 = 8 c .. + 7, then of n1 = 32 + 8 c .. + 7 (n = l + 32 n1)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    r[i] = f2u(P.window[l + 32 * (8 * c + i)]);
                    r[8 + i] = f2u(P.window[l + 32 * (32 + 8 * c + i)]);
                }
                tm_st<16>(tq + 64 + 16 * c, r);
            }
            tm_wait_st();
        }
        tm_fence_before_sync();
    } else {
        for (int i = threadIdx.x; i < Cfg::N; i += blockDim.x) s_win[(i & 31) * Cfg::WROW + (i >> 5)] = P.window[i]; // n = s + 32 n1
        for (int i = threadIdx.x; i < 512; i += blockDim.x) {
            const int l = i & 31, t = i >> 5;
            const cpx a = P.twP[twz_exponent(l, t)]; // W_2048^e, e < 1024
            s_twa[l * Cfg::TROW + t] = pk(a.x, a.y);
            const cpx b = P.twP[twz_exponent(pkz_row_b(l), t)];
            s_twb[l * Cfg::TROW + t] = pk(b.x, b.y);
        }
    }
    __syncthreads();
    if constexpr (Cfg::TM) tm_fence_after_sync();
    grid_dep_wait();

    const int s = threadIdx.x & 31, warp = threadIdx.x >> 5;
    f2* xw = s_xch + warp * Cfg::XCH;
    const float* x0 = reinterpret_cast<const float*>(xw) + s;            // staged channel 0, sample s + 32 n1 at [32 n1]
    const float* x1 = reinterpret_cast<const float*>(xw + Cfg::CH1) + s; //        channel 1
    unsigned long long* bar = s_bar + warp;
    unsigned copies = 0;
    const float4* wrow = reinterpret_cast<const float4*>(s_win + s * Cfg::WROW);
    const f2x2* trowa = reinterpret_cast<const f2x2*>(s_twa + s * Cfg::TROW);
    const f2x2* trowb = reinterpret_cast<const f2x2*>(s_twb + s * Cfg::TROW);
    const int kb = pkz_row_b(s);
    // Transpose through the warp buffer in 16-byte units {Y[s', k1], Y[s', k1']} of the two rows (k1, k1' = 64 - k1; 0 and 32) ONE
    // lane transforms in pass 2: unit (pair u, lane s') at f2x2 index 33 u + s' (32 pairs of 32 units + 16 B pad = 16 896 B).  Lane s'
    // writes its 64 values as 32 STS.128 (consecutive lanes, consecutive units), lane u reads its two rows as 32 LDS.128 (a quarter-
    // warp's units 528 B apart: all 32 banks) -- 32 store instructions per frame fewer than row-wise 8-byte stores.
    f2x2* const tr_wr = reinterpret_cast<f2x2*>(xw) + s;
    const f2x2* const tr_rd = reinterpret_cast<const f2x2*>(xw) + 33 * s;

    const unsigned total = (unsigned)P.ncols * (unsigned)P.nstreams;
    const unsigned gstep = gridDim.x * WARPS;
    const unsigned g0 = blockIdx.x * WARPS; // the CTA's first frame
    if (!Cfg::TM && g0 >= total) return;    // (CTA-uniform; with tensor memory allocated the CTA runs on to the release at the end)
    unsigned g = g0 + warp, my_iters; // first frame and number of frames of this warp
    PkzWalk cur, nxt;
    if (RING) { // a contiguous run of columns per warp
        const unsigned per = (total + gstep - 1) / gstep;
        const unsigned b = min(total, g * per), e = min(total, b + per);
        my_iters = e - b;
        cur.template init<true>(b < total ? b : 0u, 1u, (unsigned)P.ncols);
    } else {
        my_iters = g < total ? (total - g + gstep - 1) / gstep : 0;
        cur.template init<false>(g < total ? g : 0u, gstep, (unsigned)P.ncols);
    }
    nxt = cur;

    auto frame_ptr = [&](const PkzWalk& u, long long& st) { // channel 0 of the frame; st = its first sample index (may be < 0)
        st = frame_start(P, P.first_col + u.col);
        return P.samples + (long long)u.stream * P.stream_stride;
    };
    // both channels of a frame -> the warp buffer (lane 0, after a __syncwarp()); last_chunk: only its last 512 samples, in place
    auto stage = [&](const PkzWalk& u, bool last_chunk = false) {
        long long st;
        const float* a = frame_ptr(u, st) + st;
        if (s == 0) {
            if (last_chunk) bulk_copy2_g2s(xw + 768, a + 1536, xw + Cfg::CH1 + 768, a + P.channel_stride + 1536, 512 * 4, bar);
            else bulk_copy2_g2s(xw, a, xw + Cfg::CH1, a + P.channel_stride, Cfg::N * 4, bar);
        }
#if defined(JADE_EMU)
        __syncwarp();
#endif
    };
    // samples x window and the first butterfly stage of the 64-point DFT over n1 -> v (bit-reversed order); pair j = n1 j, j + 32.
    // c0 / st: channel-0 base and first sample index of the frame (PKZ_GUARD only).
    auto load_pair = [&](f2* v, int j, const float* c0, long long st, float wa, float wb) {
        if (STAGED) {
            win_stage1_64(v, j, pk(x0[32 * j], x1[32 * j]), wa, pk(x0[32 * (j + 32)], x1[32 * (j + 32)]), wb);
        } else {
            const float* c1 = c0 + P.channel_stride;
            const long long ia = st + s + 32 * j, ib = ia + 1024;
            const bool ina = ia >= 0 && ia < P.nsamples, inb = ib >= 0 && ib < P.nsamples;
            win_stage1_64(v, j, pk(ina ? c0[ia] : 0.f, ina ? c1[ia] : 0.f), wa, pk(inb ? c0[ib] : 0.f, inb ? c1[ib] : 0.f), wb);
        }
    };

    if (STAGED && my_iters > 0) stage(cur);
    // PKZ_RING: chunk c of the current frame sits in ring slot (ring + c) & 3, columns tring + 32 slot + 2 i + channel for n1 = 16 c + i
    unsigned ring = 0;
    const uint32_t tring = tq + 128u * (1u + ((unsigned)warp >> 2));
    bool warm = false; // the current frame starts one chunk after the previous frame of this warp (three chunks in the ring already)
    auto ingest = [&](int c) { // chunk c of the staged frame -> its ring slot
        uint32_t xs[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            xs[2 * i] = f2u(x0[32 * (16 * c + i)]);
            xs[2 * i + 1] = f2u(x1[32 * (16 * c + i)]);
        }
        tm_st<32>(tring + 32 * ((ring + c) & 3u), xs);
    };

    // eps = 1e-11 (Spectrogram.cpp:36) rides on the power FMAs: every output is the sum of one "low" power (register 0..15 of
    // a row, seeded with eps) and one "high" power (16..31, unseeded); lane 0 pairs row 0 with itself, so its DC and Nyquist
    // terms (register 0 / 16, added to themselves) carry eps / 2 each.
    const float eps = 1e-11f, eps0 = s == 0 ? 0.5e-11f : 1e-11f, eps16 = s == 0 ? 0.5e-11f : 0.0f;

    for (unsigned it = 0; it < my_iters; ++it) {
        const ColOut o = col_out(P, (int)cur.stream, P.first_col + cur.col);
        // ---- samples x window, stage 1 (the frame was staged while the previous one was in pass 2)
        f2 v[64];
        if (STAGED && !RING) {
            mbar_wait(bar, copies & 1u);
            ++copies;
        }
        {
            long long st;
            const float* c0 = frame_ptr(cur, st);
            if constexpr (RING) {
                // Eight pairs at a time: samples n1 = 8 c .. + 7 (slot of chunk c >> 1) and 32 + 8 c .. + 7 (chunk (c >> 1) + 2), window
                // chunk c.  Order c = 2, 3, 0, 1: the NEW chunk (3 = the high samples of c = 2, 3) is consumed from the registers that
                // carry it from shared to tensor memory (no store -> wait -> load round trip in front of the arithmetic), and the
                // tensor-memory loads of c = 2 are in flight before the mbarrier wait.
                uint32_t xa[2][16], xb[2][16], wq[2][16], xs[32];
                auto fetch = [&](int c, bool hi) {
                    tm_ld<16>(tring + 32 * ((ring + (c >> 1)) & 3u) + 16 * (c & 1), xa[c & 1]);
                    if (hi) tm_ld<16>(tring + 32 * ((ring + (c >> 1) + 2) & 3u) + 16 * (c & 1), xb[c & 1]);
                    tm_ld<16>(tq + 64 + 16 * c, wq[c & 1]);
                };
                if (!warm) {
                    mbar_wait(bar, copies & 1u);
                    ring = 0;
                    ingest(0);
                    ingest(1);
                    ingest(2);
                    tm_wait_st();
                    fetch(2, false);
                } else {
                    fetch(2, false);
                    mbar_wait(bar, copies & 1u);
                }
                ++copies;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    xs[2 * i] = f2u(x0[32 * (48 + i)]);
                    xs[2 * i + 1] = f2u(x1[32 * (48 + i)]);
                }
                tm_st<32>(tring + 32 * ((ring + 3) & 3u), xs);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const int c = (cc + 2) & 3;
                    tm_wait_ld<16>(xa[c & 1]);
                    if (c < 2) tm_tie<16>(xb[c & 1]);
                    tm_tie<16>(wq[c & 1]);
                    if (cc < 3) fetch((c + 1) & 3, cc >= 1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const f2 hi_s = c < 2 ? pk(u2f(xb[c & 1][2 * i]), u2f(xb[c & 1][2 * i + 1]))
                                              : pk(u2f(xs[16 * (c & 1) + 2 * i]), u2f(xs[16 * (c & 1) + 2 * i + 1]));
                        win_stage1_64(v, 8 * c + i, pk(u2f(xa[c & 1][2 * i]), u2f(xa[c & 1][2 * i + 1])), u2f(wq[c & 1][i]), hi_s, u2f(wq[c & 1][8 + i]));
                    }
                }
                tm_wait_st(); // (the slot is read again by the next frame of this warp)
                ring = (ring + 1) & 3u;
            } else if constexpr (Cfg::TM) {
                // window chunk c (n1 = 8 c .. + 7 and 32 + 8 c .. + 7) comes back from tensor memory while chunk c - 1 is consumed
                uint32_t wq[2][16];
                tm_ld<16>(tq + 64, wq[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    tm_wait_ld<16>(wq[c & 1]);
                    if (c < 3) tm_ld<16>(tq + 64 + 16 * (c + 1), wq[(c + 1) & 1]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) load_pair(v, 8 * c + i, c0, st, u2f(wq[c & 1][i]), u2f(wq[c & 1][8 + i]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float4 wa4 = wrow[j / 4], wb4 = wrow[(j + 32) / 4];
                    const float wa = (j & 3) == 0 ? wa4.x : (j & 3) == 1 ? wa4.y : (j & 3) == 2 ? wa4.z : wa4.w;
                    const float wb = (j & 3) == 0 ? wb4.x : (j & 3) == 1 ? wb4.y : (j & 3) == 2 ? wb4.z : wb4.w;
                    load_pair(v, j, c0, st, wa, wb);
                }
            }
        }
        if (STAGED) __syncwarp(); // every lane has read its samples before the transpose overwrites them
        fft64_pk_after_stage1(v); // v[k1] = Y[s, k1]
#pragma unroll
        for (int u = 0; u < 32; ++u) {
            f2x2 t;
            t.a = v[u];
            t.b = v[pkz_row_b(u)];
            tr_wr[33 * u] = t;
        }
        __syncwarp();
        f2 ua[32], ub[32];
#pragma unroll
        for (int jx = 0; jx < 32; ++jx) {
            const f2x2 t = tr_rd[jx];
            ua[brev(jx, 5)] = t.a;
            ub[brev(jx, 5)] = t.b;
        }
        __syncwarp(); // the buffer is free again
        const bool more = it + 1 < my_iters;
        if (more) {
            nxt.template next<RING>((unsigned)P.ncols);
            if constexpr (RING) {
                // the next frame continues this one (same stream, one chunk further): three of its chunks are in the ring already
                // (evenly spaced columns N/4 apart by dispatch, launch_one: the next column of the same stream is one chunk on)
                warm = nxt.col != 0;
                stage(nxt, warm);
            } else if (STAGED) {
                stage(nxt);
            }
        }
        if constexpr (Cfg::TM) {
            // twisted tables from tensor memory, eight entries (16 columns) at a time, the next eight in flight behind them
            uint32_t ta[16], tb[16];
            tm_ld<16>(tq, ta);
            tm_wait_ld<16>(ta);
            tm_ld<16>(tq + 16, tb);
            fft32_twisted_lo(ua, ta); // ua[k2] = Z[s  + 64 k2]
            tm_wait_ld<16>(tb);
            tm_ld<16>(tq + 32, ta);
            fft32_twisted_hi(ua, tb);
            tm_wait_ld<16>(ta);
            tm_ld<16>(tq + 48, tb);
            fft32_twisted_lo(ub, ta); // ub[k2] = Z[kb + 64 k2]
            tm_wait_ld<16>(tb);
            fft32_twisted_hi(ub, tb);
        } else {
            fft32_twisted(ua, trowa); // ua[k2] = Z[s  + 64 k2]
            fft32_twisted(ub, trowb); // ub[k2] = Z[kb + 64 k2]
        }
        // mean power of bin k (+ eps): |Z[k]|^2 + |Z[N-k]|^2.  Row s: bins s + 64 k2 mirror into row 64 - s at 31 - k2 (lane 0: row 0
        // at 32 - k2); row kb likewise into row s (lane 0: row 32 into itself).
        // (the packed form -- (re_a^2 + re_b^2 + eps, im_a^2 + im_b^2) by two FFMA2, then one FADD -- issues one instruction
        // less per output but was 3 % slower: the lane-0 selects become 64-bit, profiles/r02_pkz_variants.txt)
        float pa[32], pb[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            const float ea = q == 0 ? eps0 : q < 16 ? eps : q == 16 ? eps16 : 0.0f, eb = q < 16 ? eps : 0.0f;
            pa[q] = fm(lo(ua[q]), lo(ua[q]), q <= 16 ? fm(hi(ua[q]), hi(ua[q]), ea) : JADE_FMUL(hi(ua[q]), hi(ua[q])));
            pb[q] = fm(lo(ub[q]), lo(ub[q]), q < 16 ? fm(hi(ub[q]), hi(ub[q]), eb) : JADE_FMUL(hi(ub[q]), hi(ub[q])));
        }
        float oa[16], ob[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            oa[q] = JADE_FADD(pa[q], s == 0 ? pa[(32 - q) & 31] : pb[31 - q]);
            ob[q] = JADE_FADD(pb[q], s == 0 ? pb[31 - q] : pa[31 - q]);
        }
        const float omid = JADE_FADD(pa[16], pa[16]); // bin 1024 (lane 0)

        // ---- epilogue: dB, palette, store
        {
            uint32_t* p_a = o.pix ? o.pix + (1024 - s) : nullptr;  // bin k -> row 1024 - k
            uint32_t* p_b = o.pix ? o.pix + (1024 - kb) : nullptr;
            float* d_a = (WANT_DB && o.db) ? o.db + s : nullptr;
            float* d_b = (WANT_DB && o.db) ? o.db + kb : nullptr;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                pkz_emit<WANT_DB, U8>(oa[q], (!WANT_DB || p_a) ? p_a - 64 * q : nullptr, d_a ? d_a + 64 * q : nullptr, P, s_pal);
                pkz_emit<WANT_DB, U8>(ob[q], (!WANT_DB || p_b) ? p_b - 64 * q : nullptr, d_b ? d_b + 64 * q : nullptr, P, s_pal);
            }
            if (s == 0) pkz_emit<WANT_DB, U8>(omid, (!WANT_DB || o.pix) ? o.pix : nullptr, (WANT_DB && o.db) ? o.db + 1024 : nullptr, P, s_pal);
        }
        cur = nxt;
    }
    if constexpr (Cfg::TM) {
        tm_fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32) tm_dealloc(tq, Cfg::TM_COLS); // warp 0: quadrant 0 = the allocation's base address
    }
}

} // namespace jade
