// jade_pk_cta.cuh -- large transforms (N = 4096 ... 65536) with the packed FP32x2 arithmetic of jade_pk.cuh.
//
// One CTA of R1 warps transforms one frame of M = 1024 R1 complex points z[m] = x[2m] + i x[2m+1] that live in shared
// memory (BASELINE configs[2]: N = 16384 -> 64 KB; "large-FFT smem staging"):
//   staging     : interior, 16-byte aligned frames are copied into the row matrix by the TMA engine (cp.async.bulk, R1 rows
//                 of 8 KB, one mbarrier) while the CTA is still in the epilogue of the previous frame;
//   column pass : thread owns columns n2 (32/R1 of them), reads z[n2 + 1024 n1] from the staged rows (or straight from
//                 global memory for boundary / unaligned frames) with the window multiply fused into the first radix-2 stage, radix-R1 DFT in registers, twiddle W_M^(n2 k1),
//                 row k1 of the shared-memory matrix;
//   row pass    : warp k1 transforms its 1024-point row with the register code of the N = 2048 kernel (radix-32,
//                 twiddle, in-place XOR-swizzled transpose, radix-32) and leaves it in natural order;
//   split       : thread handles the pairs (k, M-k), k = t + 32 R1 q: A = Z[k] + conj Z[M-k], B = Z[k] - conj Z[M-k],
//                 T = (-i W_N^k) B, X[k] = A + T, X[M-k] = conj(A - T); powers accumulate over channels in registers;
//   epilogue    : dB (MUFU.LG2), CColorPalette lookup, one packed pixel per row, rows M-k and k, coalesced.
// N = 65536 (configs[4]) does not fit (256 KB): it is built from two half-size real FFTs (even / odd samples) combined
// by one radix-2 step, with E and the mixed power spectrum in an L2-resident per-CTA scratch slot (general epilogue:
// log-frequency max-pool rows).
// Reference lines replaced: Spectrogram.cpp:50-56,137-145 (framing, window, spectrum::power), :64-107 (mix, dB),
// :634-647 + CColorpalette.h:32-47 (pixel loops).
#pragma once
#include <type_traits>
#include "jade_pk.cuh"

namespace jade {

template <int R1>
struct PkCtaCfg {
    static constexpr int M = 1024 * R1;
    static constexpr int N = 2 * M;
    static constexpr int B = M + 1;
    static constexpr int THREADS = 32 * R1;
    static constexpr int RS = 1024 + 16 / R1; // row stride (complex words): conflict-free split reads (see rowget)
    static constexpr int TROW = 34;           // per-lane row of the 1024-point inter-pass twiddle table (32 + 16 B pad)
    static constexpr int MINB = R1 >= 16 ? 1 : 16 / R1; // CTAs per SM aimed at (16 warps of 128 registers)
    static constexpr int off_row = 0;
    static constexpr int off_twI = off_row + R1 * RS * 8;
    static constexpr int off_pal = off_twI + 32 * TROW * 8;
    static JADE_HD int off_spec(int npal) { return off_pal + ((npal + 1) * 4 + 15) / 16 * 16; } // palette + its `>= m_Max` entry (colour_of_lg1)
    // stft_pkcta_kernel (general == false): one mbarrier (frame staging) where the general kernels keep their spectrum
    static JADE_HD int off_bar(int npal) { return off_spec(npal); }
    static JADE_HD int smem_bytes(int npal, bool general) { return off_spec(npal) + (general ? ((B + 3) / 4) * 16 : 16); }
    // two-half kernel (N = 4096 R1): + the pooled-row table
    static JADE_HD int smem_bytes2(int npal, int pooled_rows) { return off_spec(npal) + (pooled_rows * 8 + 15) / 16 * 16; }
};

// Z[k] of the M-point transform inside the row matrix: k1 = k % R1 is the row, k / R1 the row-FFT bin
template <int R1>
JADE_DEVICE f2 rowget(const f2* rowbuf, int k)
{
    return rowbuf[(k & (R1 - 1)) * PkCtaCfg<R1>::RS + (k / R1)];
}

// M = 1024 R1 point complex FFT by the whole CTA.  ld(m, x, w) yields the raw sample pair and the window pair of
// complex point m (the product is formed here, fused with stage 1).  Result in rowbuf (see rowget).
// wa = W_M^t of this thread (P.twA[1024 + t], loaded once per kernel): column n2 = t + 32 R1 c needs W_M^n2 = wa W_32^c.
template <int R1, typename Loader>
JADE_DEVICE void cta_fft_pk(f2* rowbuf, const f2* s_twI, f2 wa, Loader ld)
{
    using Cfg = PkCtaCfg<R1>;
    constexpr int RS = Cfg::RS, CPT = 32 / R1, H = R1 / 2;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int n2 = t + Cfg::THREADS * c;
        f2 a[R1];
#pragma unroll
        for (int j = 0; j < H; ++j) { // stage 1 pairs n1 = j and j + R1/2
            f2 xa, wa, xb, wb;
            ld(n2 + 1024 * j, xa, wa);
            ld(n2 + 1024 * (j + H), xb, wb);
            win_stage1<R1>(a, j, xa, wa, xb, wb);
        }
        fft_pk_after_stage1<R1>(a);
        // twiddles W_M^(n2 k1), k1 = 1 .. R1-1: k1 = 1 is the thread's base value times a compile-time 32nd root of unity
        // (no table read: the 8 bytes per column and frame used to come from L2, 64 KB per frame at N = 16384), the others
        // by squaring / one more product (depth <= log2 R1, a few ulp)
        f2 wk[R1];
        wk[1] = c == 0 ? wa : cmul2(wa, pk(cos32(c & 15), -sin32(c & 15)));
#pragma unroll
        for (int k1 = 2; k1 < R1; ++k1) wk[k1] = (k1 & 1) ? cmul2(wk[k1 - 1], wk[1]) : cmul2(wk[k1 / 2], wk[k1 / 2]);
        rowbuf[n2] = a[0];
#pragma unroll
        for (int k1 = 1; k1 < R1; ++k1) rowbuf[k1 * RS + n2] = cmul2(a[k1], wk[k1]);
    }
    __syncthreads();
    // row pass: warp `warp` transforms row k1 = warp (1024 points) in place
    f2* row = rowbuf + warp * RS;
    f2 v[32], u[32];
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) v[brev(n1, 5)] = row[lane + 32 * n1];
    __syncwarp();
    fft32_pk(v);
    const f2x2* trow = reinterpret_cast<const f2x2*>(s_twI + lane * Cfg::TROW);
#pragma unroll
    for (int k1 = 0; k1 < 32; k1 += 2) {
        const f2x2 tw = trow[k1 / 2];
        row[k1 * 32 + (lane ^ k1)] = (k1 == 0) ? v[0] : cmul2(v[k1], tw.a);
        row[(k1 + 1) * 32 + (lane ^ (k1 + 1))] = cmul2(v[k1 + 1], tw.b);
    }
    __syncwarp();
#pragma unroll
    for (int jx = 0; jx < 32; ++jx) u[brev(jx, 5)] = row[lane * 32 + (jx ^ lane)];
    fft32_pk(u);
    __syncwarp();
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) row[lane + 32 * k2] = u[k2];
    __syncthreads();
}

// per-lane [lane][k1] copy of the 1024-point inter-pass twiddles W_1024^(k1 lane) (P.twI is [k1][lane])
template <int R1>
JADE_DEVICE void stage_row_twiddles(f2* s_twI, const cpx* JADE_RESTRICT twI)
{
    for (int i = threadIdx.x; i < 1024; i += PkCtaCfg<R1>::THREADS) {
        const int s = i & 31, k1 = i >> 5;
        const cpx t = twI[k1 * 32 + s];
        s_twI[s * PkCtaCfg<R1>::TROW + k1] = pk(t.x, t.y);
    }
}

// N = 2048 R1 (R1 = 2, 4, 8, 16).  Identity rows in the reference orientation, hardware log2 dB.  Loads are always
// bounds-checked per frame (one warp-uniform test selects the unchecked loop), the arithmetic is identical either way.
template <int R1, int MIXK, bool WANT_DB>
JADE_KERNEL(32 * R1, PkCtaCfg<R1>::MINB) stft_pkcta_kernel(const KParams P)
{
    using Cfg = PkCtaCfg<R1>;
    constexpr int M = Cfg::M, N = Cfg::N, THREADS = Cfg::THREADS;
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    f2* rowbuf = reinterpret_cast<f2*>(sm + Cfg::off_row);
    f2* s_twI = reinterpret_cast<f2*>(sm + Cfg::off_twI);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);

    unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + Cfg::off_bar(P.npal));

    const int t = threadIdx.x;
    stage_row_twiddles<R1>(s_twI, P.twI);
    for (int i = t; i <= P.npal; i += THREADS) s_pal[i] = P.palette[i < P.npal ? i : P.ci_hi];
    if (t == 0) mbar_init(bar, 1);
    __syncthreads();
    grid_dep_wait();

    const unsigned total = (unsigned)P.ncols * (unsigned)P.nstreams;
    int ch0, ch1;
    channel_range(P, ch0, ch1);
    if (MIXK == MIX_NONE) ch1 = ch0 + 1;
    const f2* JADE_RESTRICT winp = reinterpret_cast<const f2*>(P.window);
    const float scale = (MIXK == MIX_SUM) ? (1.0f / (float)P.channels) : 1.0f;
    // per-thread twiddle bases, loaded once: W_M^t (column pass) and W_N^t (real-FFT split); every other twiddle of this thread
    // is one of them times a compile-time root of unity (they used to be two 8-byte L2 reads per point and frame)
    f2 wa, ws;
    {
        const cpx a = P.twA[1024 + t], b = P.twP[t];
        wa = pk(a.x, a.y);
        ws = pk(b.x, b.y);
    }

    // Interior frames on a multiple of 4 samples (P.aligned4) are STAGED: the TMA engine copies the frame's R1 rows of
    // 1024 complex samples into the row matrix (cp.async.bulk, one mbarrier completion) while the CTA is still in the
    // epilogue of the previous frame, and the column pass then works in place on shared memory.
    auto frame_of = [&](unsigned gg, int& stream_, long long& j_, long long& st_, bool& staged_) {
        stream_ = (int)(gg / (unsigned)P.ncols);
        j_ = P.first_col + (gg - (unsigned)stream_ * (unsigned)P.ncols);
        st_ = frame_start(P, j_);
        // (R1 = 16: rows of 1025 words are only 8-byte aligned, cp.async.bulk needs 16 -- N = 32768 keeps the global loads)
        staged_ = (Cfg::RS % 2 == 0) && P.aligned4 && st_ >= 0 && st_ + N <= P.nsamples;
    };
    auto stage = [&](const float* src) { // thread 0, after a __syncthreads(): rows n1 = 0 .. R1-1 of the frame at src
        if (t == 0) {
            mbar_expect_tx(bar, M * 8);
#pragma unroll
            for (int n1 = 0; n1 < R1; ++n1) bulk_copy_issue(rowbuf + n1 * Cfg::RS, src + 2048 * n1, 8192, bar);
        }
#if defined(JADE_EMU)
        __syncthreads();
#endif
    };
    unsigned copies = 0;
    {
        int stream_;
        long long j_, st_;
        bool staged_ = false;
        if (blockIdx.x < total) frame_of(blockIdx.x, stream_, j_, st_, staged_);
        if (staged_) stage(P.samples + stream_ * P.stream_stride + ch0 * P.channel_stride + st_);
    }

    for (unsigned g = blockIdx.x; g < total; g += gridDim.x) {
        int stream;
        long long j, st;
        bool staged;
        frame_of(g, stream, j, st, staged);
        const bool fast = P.aligned2 && st >= 0 && st + N <= P.nsamples;

        float alo[16], ahi[16], amid = 0.f; // bins t + THREADS q / M - (t + THREADS q) / M/2 (thread 0)
#pragma unroll
        for (int q = 0; q < 16; ++q) alo[q] = ahi[q] = 0.f;

        for (int ch = ch0; ch < ch1; ++ch) {
            const float* x = P.samples + stream * P.stream_stride + ch * P.channel_stride;
            const long long ns = P.nsamples;
            if (staged) {
                mbar_wait(bar, copies & 1u);
                ++copies;
                cta_fft_pk<R1>(rowbuf, s_twI, wa, [&](int m, f2& xv, f2& wv) {
                    xv = rowbuf[(m >> 10) * Cfg::RS + (m & 1023)]; // the thread's own column: replaced in place below
                    wv = winp[m];
                });
            } else if (fast) {
                const f2* xz = reinterpret_cast<const f2*>(x + st);
                cta_fft_pk<R1>(rowbuf, s_twI, wa, [&](int m, f2& xv, f2& wv) {
                    xv = xz[m];
                    wv = winp[m];
                });
            } else {
                cta_fft_pk<R1>(rowbuf, s_twI, wa, [&](int m, f2& xv, f2& wv) {
                    const cpx z = load_pair_guarded(x, st + 2 * m, ns);
                    xv = pk(z.x, z.y);
                    wv = winp[m];
                });
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int k = t + THREADS * q;
                const f2 zk = rowget<R1>(rowbuf, k);
                const f2 zp = rowget<R1>(rowbuf, (M - k) & (M - 1));
                const f2 w = cmul2(ws, pk(cos64(q), -sin64(q))); // W_N^k = W_N^t W_64^q ; -i W_N^k = (w.y, -w.x)
                const f2 A = add2(zk, conj2(zp));
                const f2 Bv = sub2(zk, conj2(zp));
                const f2 T = cmul2(Bv, pk(hi(w), -lo(w)));
                const f2 xp = add2(A, T), xm = sub2(A, T);
                alo[q] = fm(lo(xp), lo(xp), fm(hi(xp), hi(xp), alo[q]));
                ahi[q] = fm(lo(xm), lo(xm), fm(hi(xm), hi(xm), ahi[q]));
            }
            { // bin M/2 (self-paired): X = 2 conj Z
                const f2 zm = rowget<R1>(rowbuf, M / 2);
                const float a = lo(zm), b = hi(zm);
                amid = fm(JADE_FMUL(4.0f, a), a, fm(JADE_FMUL(4.0f, b), b, amid));
            }
            __syncthreads();
            // the row matrix is free: stage what this CTA transforms next (covered by the epilogue below)
            if (ch + 1 < ch1) {
                if (staged) stage(x + P.channel_stride + st);
            } else if (g + gridDim.x < total) {
                int stream_;
                long long j_, st_;
                bool staged_;
                frame_of(g + gridDim.x, stream_, j_, st_, staged_);
                if (staged_) stage(P.samples + stream_ * P.stream_stride + ch0 * P.channel_stride + st_);
            }
        }

        const ColOut o = col_out(P, stream, j);
        uint32_t* p_lo = o.pix ? o.pix + (M - t) : nullptr; // bin k -> row M - k
        uint32_t* p_hi = o.pix ? o.pix + t : nullptr;       // bin M - k -> row k
        float* d_lo = (WANT_DB && o.db) ? o.db + t : nullptr;
        float* d_hi = (WANT_DB && o.db) ? o.db + (M - t) : nullptr;
        // (two copies behind a launch-uniform branch: with the u8 palette the float -> integer conversion is the clamp, colour_of_lg1)
        auto finish = [&](auto u8) {
            constexpr bool U8 = decltype(u8)::value;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                // (kept as two guarded blocks per q: the straight-line emit_bin form schedules 7 % slower here, variants run)
                const float ll = JADE_LOG2F(MIXK == MIX_SUM ? fm(alo[q], scale, 1e-11f) : JADE_FADD(alo[q], 1e-11f));
                const float lh = JADE_LOG2F(MIXK == MIX_SUM ? fm(ahi[q], scale, 1e-11f) : JADE_FADD(ahi[q], 1e-11f));
                if (WANT_DB && d_lo) {
                    d_lo[THREADS * q] = JADE_FMUL(3.01029995663981195f, ll);
                    d_hi[-THREADS * q] = JADE_FMUL(3.01029995663981195f, lh);
                }
                if (p_lo) {
                    p_lo[-THREADS * q] = colour_of_lg1<U8>(ll, P, s_pal);
                    p_hi[THREADS * q] = colour_of_lg1<U8>(lh, P, s_pal);
                }
            }
            if (t == 0) emit_bin1<MIXK, WANT_DB, U8, false>(amid, scale, (!WANT_DB || o.pix) ? o.pix + M / 2 : nullptr, (WANT_DB && o.db) ? o.db + M / 2 : nullptr, P, s_pal);
        };
        if (P.pal_u8) finish(std::true_type{});
        else finish(std::false_type{});
    }
}

// N = 4096 R1 (R1 = 16 -> N = 65536) from two half-size real FFTs (decimation in time on the REAL data):
//   X[k] = E[k] + W_N^k O[k],  X[N/2-k] = conj(E[k] - W_N^k O[k]),  k = 0..N/4
// E / O = real FFT (size N/2, M2 = N/4 = 1024 R1 complex points) of the even / odd windowed samples.
template <int R1, int MIXK>
JADE_KERNEL(32 * R1, 1) stft_pkcta2_kernel(const KParams P)
{
    using Cfg = PkCtaCfg<R1>;
    constexpr int M2 = Cfg::M;      // complex points per half
    constexpr int NH = 2 * M2;      // real points per half  (= N/2)
    constexpr int N = 2 * NH;
    constexpr int THREADS = Cfg::THREADS;
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    f2* rowbuf = reinterpret_cast<f2*>(sm + Cfg::off_row);
    f2* s_twI = reinterpret_cast<f2*>(sm + Cfg::off_twI);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);

    i2* s_rows = reinterpret_cast<i2*>(sm + Cfg::off_spec(P.npal)); // pooled rows: [R] bin ranges (smem_bytes2)

    const int t = threadIdx.x;
    stage_row_twiddles<R1>(s_twI, P.twI);
    for (int i = t; i < P.npal; i += THREADS) s_pal[i] = P.palette[i];
    if (P.pooled)
        for (int i = t; i < P.R; i += THREADS) s_rows[i] = P.row_bins[i];
    __syncthreads();
    grid_dep_wait();

    f2 wa;
    {
        const cpx a = P.twA[1024 + t];
        wa = pk(a.x, a.y);
    }
    f2* se = reinterpret_cast<f2*>(P.scratch_e) + (long long)blockIdx.x * (M2 + 1);
    float* sp = P.scratch_p + (long long)blockIdx.x * (NH + 1);
    const unsigned total = (unsigned)P.ncols * (unsigned)P.nstreams;
    int ch0, ch1;
    channel_range(P, ch0, ch1);

    for (unsigned g = blockIdx.x; g < total; g += gridDim.x) {
        const int stream = (int)(g / (unsigned)P.ncols);
        const long long j = P.first_col + (g - (unsigned)stream * (unsigned)P.ncols);
        const long long st = frame_start(P, j);
        const bool fast = st >= 0 && st + N <= P.nsamples;

        for (int ch = ch0; ch < (MIXK == MIX_NONE ? ch0 + 1 : ch1); ++ch) {
            const float* x = P.samples + stream * P.stream_stride + ch * P.channel_stride;
            const long long ns = P.nsamples;
            const float* JADE_RESTRICT win = P.window;
            for (int half = 0; half < 2; ++half) {
                // z[m] = xw[4m + half] + i xw[4m + 2 + half].  Three loaders, identical arithmetic afterwards:
                // 16-byte vector loads (interior frame, 16-byte aligned), scalar loads (interior), branch-free guarded.
                const float* xs = x + st;
                if (fast && ((reinterpret_cast<uintptr_t>(xs) | reinterpret_cast<uintptr_t>(win)) & 15) == 0) {
                    const float4* x4 = reinterpret_cast<const float4*>(xs);
                    const float4* w4 = reinterpret_cast<const float4*>(win);
                    if (half == 0) {
                        cta_fft_pk<R1>(rowbuf, s_twI, wa, [&](int m, f2& xv, f2& wv) {
                            const float4 a = x4[m], w = w4[m];
                            xv = pk(a.x, a.z);
                            wv = pk(w.x, w.z);
                        });
                    } else {
                        cta_fft_pk<R1>(rowbuf, s_twI, wa, [&](int m, f2& xv, f2& wv) {
                            const float4 a = x4[m], w = w4[m];
                            xv = pk(a.y, a.w);
                            wv = pk(w.y, w.w);
                        });
                    }
                } else if (fast) {
                    cta_fft_pk<R1>(rowbuf, s_twI, wa, [&](int m, f2& xv, f2& wv) {
                        xv = pk(xs[4 * m + half], xs[4 * m + 2 + half]);
                        wv = pk(win[4 * m + half], win[4 * m + 2 + half]);
                    });
                } else {
                    cta_fft_pk<R1>(rowbuf, s_twI, wa, [&](int m, f2& xv, f2& wv) {
                        const long long i0 = st + 4LL * m + half, i1 = i0 + 2;
                        const long long c0 = i0 < 0 ? 0 : (i0 < ns ? i0 : ns - 1), c1 = i1 < 0 ? 0 : (i1 < ns ? i1 : ns - 1);
                        const float a = x[c0], b = x[c1]; // clamped address, value selected afterwards: no branches
                        xv = pk((i0 >= 0 && i0 < ns) ? a : 0.f, (i1 >= 0 && i1 < ns) ? b : 0.f);
                        wv = pk(win[4 * m + half], win[4 * m + 2 + half]);
                    });
                }
                // 32 independent iterations per thread (+ k = M2 on thread 0), unrolled so that the L2 round trips of
                // several iterations are in flight together (one CTA of 16 warps per SM cannot hide them otherwise)
                auto split_one = [&](int k) {
                    const f2 zk = rowget<R1>(rowbuf, k & (M2 - 1));
                    const f2 zp = rowget<R1>(rowbuf, (M2 - k) & (M2 - 1));
                    const cpx wh = P.twH[k];
                    // E[k] or O[k] = (Z[k] + conj Z[M2-k]) - i W (Z[k] - conj Z[M2-k])
                    const f2 hv = add2(add2(zk, conj2(zp)), cmul2(sub2(zk, conj2(zp)), pk(wh.y, -wh.x)));
                    if (half == 0) {
                        se[k] = hv;
                    } else {
                        const f2 e = se[k];
                        const cpx wp = P.twP[k];
                        const f2 qv = cmul2(hv, pk(wp.x, wp.y));
                        const f2 xa = add2(e, qv), xb = sub2(e, qv);
                        const float p1 = fm(lo(xa), lo(xa), JADE_FMUL(hi(xa), hi(xa)));
                        const float p2 = fm(lo(xb), lo(xb), JADE_FMUL(hi(xb), hi(xb)));
                        const int k2 = NH - k;
                        float a1 = mix_init<MIXK>(P.mix_mode), a2 = a1;
                        if (MIXK != MIX_NONE && ch != ch0) {
                            a1 = sp[k];
                            a2 = sp[k2];
                        }
                        mix_add<MIXK>(a1, p1, P.mix_mode);
                        mix_add<MIXK>(a2, p2, P.mix_mode);
                        sp[k] = a1;
                        if (k2 != k) sp[k2] = a2;
                    }
                };
#pragma unroll 4
                for (int q = 0; q < M2 / THREADS; ++q) split_one(t + THREADS * q);
                if (t == 0) split_one(M2);
                __syncthreads();
            }
        }
        const ColOut o = col_out(P, stream, j);
        if (P.pooled && !o.db) {
            // Log-frequency max-pool rows (configs[4]): the mixed power spectrum moves from the L2 scratch slot into the
            // (now idle) shared-memory row matrix and is pooled there -- a scan of the L2 copy serialises ~250 dependent
            // L2 round trips per row.
            float* spec_s = reinterpret_cast<float*>(rowbuf); // B floats <= R1 * RS * 8 bytes
            const bool mean = P.mix_mode == K_MIX_ABSMEAN && P.channels > 1;
            const float nchf = (float)P.channels;
#pragma unroll 8
            for (int q = 0; q < NH / THREADS; ++q) {
                const int k = t + THREADS * q;
                const float p = sp[k];
                spec_s[k] = mean ? JADE_FDIV(p, nchf) : p;
            }
            if (t == 0) spec_s[NH] = mean ? JADE_FDIV(sp[NH], nchf) : sp[NH];
            __syncthreads();
            // one thread per row, scanning its band in shared memory with four independent max chains; consecutive
            // threads own consecutive rows, so the pixel stores coalesce.  (A warp-per-row scan with a shuffle max
            // spends ~45 instructions per row on loop control for the many one-bin rows: measured 70 k warp
            // instructions per frame against ~8 k for this form.)
            for (int r = t; r < P.R; r += THREADS) {
                const i2 rb = s_rows[r];
                float m0 = spec_s[rb.lo], m1 = m0, m2 = m0, m3 = m0; // bands are never empty (jade_host::log_rows)
                int k = rb.lo + 1;
                for (; k + 3 < rb.hi; k += 4) {
                    m0 = fmaxf(m0, spec_s[k]);
                    m1 = fmaxf(m1, spec_s[k + 1]);
                    m2 = fmaxf(m2, spec_s[k + 2]);
                    m3 = fmaxf(m3, spec_s[k + 3]);
                }
                for (; k < rb.hi; ++k) m0 = fmaxf(m0, spec_s[k]);
                const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                if (o.pix) o.pix[P.flip ? (P.R - 1 - r) : r] = colour_of(to_db(mx, P.db_precise), P, s_pal);
            }
            __syncthreads();
        } else {
            emit_general_bins(P, s_pal, o, sp, t, THREADS);
            __syncthreads();
            emit_general_rows(P, s_pal, o, sp, t, THREADS);
            __syncthreads();
        }
    }
}

} // namespace jade
