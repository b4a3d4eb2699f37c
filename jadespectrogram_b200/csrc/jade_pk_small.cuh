// jade_pk_small.cuh -- N = 64 T for T = 2, 4, 8, 16 (N = 128 ... 1024; BASELINE configs[0] is N = 1024) with the packed
// FP32x2 arithmetic of jade_pk.cuh.  The plugin's GUI offers FFT sizes 512 ... 8192 (Spectrogram.cpp:413-417): 512 and
// 1024 land here, 2048 in jade_pk.cuh, 4096 and 8192 in jade_pk_cta.cuh.
//
// T lanes transform one frame (F = 32 / T frames per warp), M = 32 T complex points z[m] = x[2m] + i x[2m+1], 32 per lane
// (m = s + T n1, s = lane % T):
//   pass 1 : radix-32 over n1 in registers (window fused into stage 1);
//   transpose inside the frame's T lanes through a padded shared-memory tile (32 rows of T+1);
//   pass 2 : lane s owns rows k1 = s + T i (i < F) and runs F twisted radix-T DFTs whose butterflies carry the inter-pass
//            twiddle W_M^(k1 j) (fft_twisted: 16 table values per lane instead of 31 and no twiddle multiply):
//            u[i T + k2] = Z[(s + T i) + 32 k2];
//   split  : pairs (k, M-k) with k2 < T/2: pair q = i T/2 + k2 of lane s has its partner in register (F-1-i) T + (T-1-k2)
//            of lane T - s of the same frame and receives it by SHFL.IDX (nothing goes through shared memory; the smem
//            exchange tile it replaces cost 32 wavefronts per warp more and two more barriers); lane s = 0, whose
//            partners are its own values at irregular indices (lane0_partner), selects them instead;
//   staging: the F frames (8 KB) a warp transforms next are copied into its tiles by the TMA engine right after the
//            transpose has been read back (cp.async.bulk, one mbarrier per warp);
//   epilogue as in jade_pk.cuh.
// Same reference lines replaced as jade_kernels.cuh (Spectrogram.cpp:50-56,137-145,64-107,634-647; CColorpalette.h:32-47).
#pragma once
#include <type_traits>
#include "jade_pk.cuh"
#include "jade_tmem.cuh"

// JADE_PKS_TMEM = 1: the per-lane window, twisted-twiddle and split-twiddle tables (128 words per lane) live in tensor memory
// (jade_tmem.cuh) instead of shared memory: a quarter of the kernel's shared-memory wavefronts
#ifndef JADE_PKS_TMEM
#define JADE_PKS_TMEM 1
#endif

// occupancy: T >= 8 (N = 512, 1024): 12 warps per SM with up to 168 registers (one CTA); smaller T: 16 warps of 128
// registers (two CTAs of 8) -- measured both ways for every T (gpurun_out: 12 x 1 is +6 % at N = 1024, -7 % at N = 128).
// -DJADE_PKS_WARPS / -DJADE_PKS_CTAS override both for experiments.
namespace jade {

// Twisted R-point DIT pass (R = 2 .. 16), the small sibling of fft32_twisted (jade_pk.cuh): the inter-pass twiddle
// W_M^{k1 j} of row k1 rides in the butterflies, stage LEN uses W_M^{(R/LEN)(k1 + 32 J)} (M = 32 R), J < LEN/2, and
// J >= LEN/4 is -i times entry J - LEN/4: R/2 table values per row at tw[0] (LEN 2), tw[1] (LEN 4), tw[2..3] (LEN 8),
// tw[4..7] (LEN 16).  Bit-reversed input, natural-order output.
template <int R, int LEN, int BASE, int J>
JADE_DEVICE void twr_inner(f2* a, const f2* tw)
{
    if constexpr (J < LEN / 2) {
        constexpr int Q = (LEN >= 4) ? LEN / 4 : 1;
        if constexpr (J < Q) bfly_w(a[BASE + J], a[BASE + J + LEN / 2], tw[J]);
        else bfly_wmi(a[BASE + J], a[BASE + J + LEN / 2], tw[J - Q]);
        twr_inner<R, LEN, BASE, J + 1>(a, tw);
    }
}
template <int R, int LEN, int BASE>
JADE_DEVICE void twr_blocks(f2* a, const f2* tw)
{
    if constexpr (BASE < R) {
        twr_inner<R, LEN, BASE, 0>(a, tw);
        twr_blocks<R, LEN, BASE + LEN>(a, tw);
    }
}
template <int R, int LEN>
JADE_DEVICE void twr_stages(f2* a, const f2* tw)
{
    if constexpr (LEN <= R) {
        twr_blocks<R, LEN, 0>(a, tw + (LEN >= 4 ? LEN / 4 : 0));
        twr_stages<R, LEN * 2>(a, tw);
    }
}
template <int R>
JADE_DEVICE void fft_twisted(f2* a, const f2* tw)
{
    twr_stages<R, 2>(a, tw);
}
// exponent e (entry = W_M^e, M = 32 R) of table value t (0 .. R/2-1) of row k1
template <int R>
JADE_HD int twr_exponent(int k1, int t)
{
    const int LEN = t == 0 ? 2 : (t == 1 ? 4 : (t < 4 ? 8 : 16));
    const int J = t < 2 ? 0 : t - LEN / 4;
    return (R / LEN) * (k1 + 32 * J);
}

// MONO (one contributing channel, MIX_NONE): no channel accumulators, the kernel fits 128 registers and 16 warps share the SM
// for N = 512 / 1024 (cfg1 +3 %, gpurun_out/ab3.txt); the mixing instantiations spill at 128 and stay at 12 warps
template <int T, bool MONO = false>
struct PkSmallCfg {
    static constexpr int M = 32 * T, N = 2 * M, B = M + 1;
    static constexpr int F = 32 / T;
    static constexpr int H = T / 2;   // k2 values per pair half
#if defined(JADE_PKS_WARPS) && defined(JADE_PKS_CTAS)
    static constexpr int WARPS = JADE_PKS_WARPS, CTAS = JADE_PKS_CTAS;
#else
    static constexpr int WARPS = (T >= 8) ? (MONO ? 16 : 12) : 8, CTAS = (T >= 8) ? 1 : 2;
#endif
    static constexpr int ROW = 34;    // f2 words per s-row of the window / inter-pass twiddle tables (32 + 16 B pad)
    static constexpr int PROW = 18;   // f2 words per s-row of the split-twiddle table (16 + 16 B pad)
    // per-frame tile (f2 words): 32 rows of T+1 for the transpose (and the landing area of the next frame); the stride
    // between the F tiles of a warp is kept == T (mod 16) so that frames sharing a half-warp hit disjoint banks
    static constexpr int FS = 32 * (T + 1) + (T < 16 ? T : 0);
    static constexpr bool TM = JADE_PKS_TMEM != 0 && T >= 8; // (N = 128 / 256, two CTAs per SM: 3 % slower with it, gpurun_out/pks1.txt)
    static constexpr int TM_COLS = 128; // tensor-memory columns per lane: window 0..63 (in the order pass 1 consumes it), twisted table 64..95, split table 96..127
    static constexpr int off_win = 0;
    static constexpr int off_twI = off_win + (TM ? 0 : T * ROW * 8);
    static constexpr int off_twP = off_twI + (TM ? 0 : T * ROW * 8);
    static constexpr int off_pal = off_twP + (TM ? 0 : T * PROW * 8);
    static JADE_HD int off_bar(int npal) { return off_pal + ((npal + 1) * 4 + 15) / 16 * 16; } // palette + its `>= m_Max` entry (colour_of_lg1); then one mbarrier per warp (+ the tensor-memory address)
    static JADE_HD int off_xch(int npal) { return off_bar(npal) + (WARPS * 8 + 8 + 15) / 16 * 16; }
    static JADE_HD int smem_bytes(int npal) { return off_xch(npal) + WARPS * F * FS * 8; }
};

// u-index of the value lane 0 pairs with its own pair q = i H + k2 (k = T i + 32 k2):  M - k lives in lane 0 as well
template <int T>
JADE_HD constexpr int lane0_partner(int q)
{
    constexpr int F = 32 / T, H = T / 2;
    const int i = q / H, k2 = q % H;
    return i >= 1 ? (F - i) * T + (T - 1 - k2) : (k2 >= 1 ? T - k2 : 0);
}

template <int T, int MIXK, bool WANT_DB, bool GUARD>
JADE_KERNEL((PkSmallCfg<T, MIXK == MIX_NONE>::WARPS * 32), (PkSmallCfg<T, MIXK == MIX_NONE>::CTAS)) stft_pksmall_kernel(const KParams P)
{
    using Cfg = PkSmallCfg<T, MIXK == MIX_NONE>;
    constexpr int M = Cfg::M, F = Cfg::F, H = Cfg::H, FS = Cfg::FS, TS = T + 1;
    static_assert(T >= 2 && T <= 16, "T = 32 is jade_pk.cuh");
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    f2* s_win = reinterpret_cast<f2*>(sm + Cfg::off_win);
    f2* s_twI = reinterpret_cast<f2*>(sm + Cfg::off_twI);
    f2* s_twP = reinterpret_cast<f2*>(sm + Cfg::off_twP);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(sm + Cfg::off_bar(P.npal));
    f2* s_xch = reinterpret_cast<f2*>(sm + Cfg::off_xch(P.npal));

    if (!GUARD && threadIdx.x < Cfg::WARPS) mbar_init(s_bar + threadIdx.x, 1);
    uint32_t tq = 0; // tensor-memory address of this warp's quadrant of the tables
    if constexpr (Cfg::TM) {
        uint32_t* s_tm = reinterpret_cast<uint32_t*>(s_bar + Cfg::WARPS);
        if (threadIdx.x < 32) tm_alloc(s_tm, Cfg::TM_COLS);
        tm_fence_before_sync();
        __syncthreads();
        tm_fence_after_sync();
        tq = tm_quadrant_base(*s_tm);
        if (threadIdx.x < 128) { // warp q fills quadrant q: the tables of lane l depend on s = l % T only
            const int s = (threadIdx.x & 31) % T;
            uint32_t r[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) { // window chunk c: points n1 = 4c .. 4c+3, then 16 + 4c .. 16 + 4c + 3 (m = s + T n1)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ma = s + T * (4 * c + i), mb = ma + 16 * T;
                    r[2 * i] = f2u(P.window[2 * ma]);
                    r[2 * i + 1] = f2u(P.window[2 * ma + 1]);
                    r[8 + 2 * i] = f2u(P.window[2 * mb]);
                    r[8 + 2 * i + 1] = f2u(P.window[2 * mb + 1]);
                }
                tm_st<16>(tq + 16 * c, r);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) { // twisted pass-2 table: entry q = ip (T/2) + t for row k1 = s + T ip
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int q = 8 * c + i, ip = q / H, t = q % H;
                    const cpx w = P.twP[2 * twr_exponent<T>(s + T * ip, t)]; // W_M^e = W_N^{2e}
                    r[2 * i] = f2u(w.x);
                    r[2 * i + 1] = f2u(w.y);
                }
                tm_st<16>(tq + 64 + 16 * c, r);
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) { // split table: pair q = ip H + k2: k = s + T ip + 32 k2, entry -i W_N^k = (w.y, -w.x)
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int q = 8 * c + i;
                    const cpx w = P.twP[s + T * (q / H) + 32 * (q % H)];
                    r[2 * i] = f2u(w.y);
                    r[2 * i + 1] = f2u(-w.x);
                }
                tm_st<16>(tq + 96 + 16 * c, r);
            }
            tm_wait_st();
        }
        tm_fence_before_sync();
    } else {
    // ---- per-s tables: entry (s, index) at [s*ROW + index]
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        const int s = i % T, n1 = i / T;                         // m = s + T n1
        s_win[s * Cfg::ROW + n1] = pk(P.window[2 * i], P.window[2 * i + 1]);
    }
    for (int i = threadIdx.x; i < 16 * T; i += blockDim.x) { // twisted pass-2 table: [s][ip (T/2) + t] for row k1 = s + T ip
        const int s = i % T, q = i / T, ip = q / H, t = q % H;
        const cpx w = P.twP[2 * twr_exponent<T>(s + T * ip, t)]; // W_M^e = W_N^{2e}
        s_twI[s * Cfg::ROW + q] = pk(w.x, w.y);
    }
    for (int i = threadIdx.x; i < 16 * T; i += blockDim.x) {
        const int s = i % T, q = i / T;                          // pair q = ip H + k2: k = s + T ip + 32 k2
        const cpx w = P.twP[s + T * (q / H) + 32 * (q % H)];     // W_N^k ; table holds -i W_N^k = (w.y, -w.x)
        s_twP[s * Cfg::PROW + q] = pk(w.y, -w.x);
    }
    }
    for (int i = threadIdx.x; i <= P.npal; i += blockDim.x) s_pal[i] = P.palette[i < P.npal ? i : P.ci_hi];
    __syncthreads();
    if constexpr (Cfg::TM) tm_fence_after_sync();
    grid_dep_wait();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = lane / T, s = lane % T;
    f2* xw = s_xch + (warp * F + f) * FS;
    const f2x2* wrow = reinterpret_cast<const f2x2*>(s_win + s * Cfg::ROW);
    const f2x2* trow = reinterpret_cast<const f2x2*>(s_twI + s * Cfg::ROW);
    const f2x2* prow = reinterpret_cast<const f2x2*>(s_twP + s * Cfg::PROW);
    f2* tr_wr = xw + s;               // transpose: word k1*TS + s
    const f2* tr_rd = xw + s * TS;    //            row k1 = s + T i at tr_rd[T*i*TS + j]
    const int plane = (lane & ~(T - 1)) | ((T - s) & (T - 1)); // lane of the same frame holding the mirrored bins

    const unsigned groups = (unsigned)((P.ncols + F - 1) / F);
    const unsigned total = groups * (unsigned)P.nstreams;
    const unsigned gstep = gridDim.x * Cfg::WARPS;
    int ch0, ch1;
    channel_range(P, ch0, ch1);
    if (MIXK == MIX_NONE) ch1 = ch0 + 1;
    const float scale = (MIXK == MIX_SUM) ? (1.0f / (float)P.channels) : 1.0f;
    unsigned long long* bar = s_bar + warp;
    unsigned copies = 0; // staged groups of this warp waited for so far (mbarrier phase parity)

    // frame of this lane in group grp of a stream: column (clamped to the last one), first sample
    auto frame_of = [&](unsigned grp, long long& j_, long long& st_, bool& active_) {
        int jrel = (int)grp * F + f;
        active_ = jrel < P.ncols;
        if (!active_) jrel = P.ncols - 1; // transform the last column again, store nothing
        j_ = P.first_col + jrel;
        st_ = frame_start(P, j_);
    };
    // the warp's groups g, g + gstep, ... as (stream, group in the stream), advanced without a division per group
    const unsigned w_dq = gstep / groups, w_dr = gstep - w_dq * groups;
    // (T >= 8; the 128-register instantiations for N <= 256 lose 2..3 % to the two more live registers and keep the division)
    constexpr bool WALK = T >= 8;
    auto advance = [&](unsigned& stream_, unsigned& grp_) {
        stream_ += w_dq;
        grp_ += w_dr;
        if (grp_ >= groups) {
            grp_ -= groups;
            ++stream_;
        }
    };
    // GUARD = false: interior frames starting on a multiple of 4 samples with 16-byte aligned channel bases
    // (P.aligned4).  The TMA engine copies the F frames (8 KB together) a warp transforms next -- channel ch of group gg --
    // into the warp's F tiles while the warp is still busy with the split and the epilogue of the previous ones: lane 0
    // announces the bytes on the warp's mbarrier, then the first lane of every frame issues its cp.async.bulk.
    auto stage = [&](unsigned stream_, unsigned grp_, int ch) {
        long long j_, st_;
        bool act_;
        frame_of(grp_, j_, st_, act_);
        if (lane == 0) mbar_expect_tx(bar, F * M * 8);
        __syncwarp();
        if (s == 0) bulk_copy_issue(xw, P.samples + (long long)stream_ * P.stream_stride + ch * P.channel_stride + st_, M * 8, bar);
#if defined(JADE_EMU)
        __syncwarp();
#endif
    };

    unsigned g = blockIdx.x * Cfg::WARPS + warp;
    unsigned w_stream = g / groups, w_grp = g - w_stream * groups;
    unsigned n_stream = w_stream, n_grp = w_grp; // the warp's next group
    if (!GUARD && g < total) stage(w_stream, w_grp, ch0);
    for (; g < total; g += gstep, w_stream = n_stream, w_grp = n_grp) {
        if constexpr (WALK) {
            advance(n_stream, n_grp);
        } else {
            w_stream = g / groups;
            w_grp = g - w_stream * groups;
        }
        const int stream = (int)w_stream;
        long long j, st;
        bool active;
        frame_of(w_grp, j, st, active);

        constexpr float seed = MIXK == MIX_NONE ? 1e-11f : 0.f; // one contributing channel: the + 1e-11 (Spectrogram.cpp:36) rides on the power FMAs
        float alo[16], ahi[16], amid = seed; // bins k(q) = s + T i + 32 k2 / M - k(q) / M/2 (lane s = 0)
#pragma unroll
        for (int q = 0; q < 16; ++q) alo[q] = ahi[q] = seed;

        for (int ch = ch0; ch < ch1; ++ch) {
            const float* x = P.samples + stream * P.stream_stride + ch * P.channel_stride;
            f2 v[32];
            if (!GUARD) {
                mbar_wait(bar, copies & 1u);
                ++copies;
            }
            uint32_t wq[2][16]; // window chunks from tensor memory, one ahead
            if constexpr (Cfg::TM) tm_ld<16>(tq, wq[0]);
#pragma unroll
            for (int jj = 0; jj < 16; jj += 2) { // n1 = jj, jj+1 paired with n1 + 16
                f2x2 wa, wb;
                if constexpr (Cfg::TM) {
                    const int c = jj / 4, o = 4 * ((jj / 2) & 1);
                    if ((jj & 2) == 0) {
                        tm_wait_ld<16>(wq[c & 1]);
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
                if (!GUARD) {
                    const f2* xz = xw + s; // staged frame: z[m] at word m of the lane's tile
                    xa0 = xz[T * jj];
                    xa1 = xz[T * (jj + 1)];
                    xb0 = xz[T * (jj + 16)];
                    xb1 = xz[T * (jj + 17)];
                } else {
                    const cpx a0 = load_pair_guarded(x, st + 2 * (s + T * jj), P.nsamples);
                    const cpx a1 = load_pair_guarded(x, st + 2 * (s + T * (jj + 1)), P.nsamples);
                    const cpx b0 = load_pair_guarded(x, st + 2 * (s + T * (jj + 16)), P.nsamples);
                    const cpx b1 = load_pair_guarded(x, st + 2 * (s + T * (jj + 17)), P.nsamples);
                    xa0 = pk(a0.x, a0.y);
                    xa1 = pk(a1.x, a1.y);
                    xb0 = pk(b0.x, b0.y);
                    xb1 = pk(b1.x, b1.y);
                }
                win_stage1(v, jj, xa0, wa.a, xb0, wb.a);
                win_stage1(v, jj + 1, xa1, wa.b, xb1, wb.b);
            }
            if (!GUARD) __syncwarp(); // every lane has read its samples before the transpose overwrites them
            fft32_pk_after_stage1(v);
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) tr_wr[k1 * TS] = v[k1];
            __syncwarp();
            f2 u[32], twl[16]; // this lane's 16 twisted twiddles: T/2 per row
            if constexpr (Cfg::TM) {
                uint32_t tw[32];
                tm_ld<32>(tq + 64, tw);
                tm_wait_ld<32>(tw);
#pragma unroll
                for (int i = 0; i < 16; ++i) twl[i] = pk(u2f(tw[2 * i]), u2f(tw[2 * i + 1]));
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const f2x2 t = trow[i];
                    twl[2 * i] = t.a;
                    twl[2 * i + 1] = t.b;
                }
            }
#pragma unroll
            for (int i = 0; i < F; ++i) {
#pragma unroll
                for (int jx = 0; jx < T; ++jx) u[i * T + brev(jx, ilog2c(T))] = tr_rd[T * i * TS + jx];
                fft_twisted<T>(u + i * T, twl + i * H); // u[i T + k2] = Z[(s + T i) + 32 k2], twiddle W_M^(k1 j) included
            }
            __syncwarp(); // the tiles are free again
            if (!GUARD) { // stage what this warp transforms next: the next channel of this group, or its next group
                if (ch + 1 < ch1) stage(w_stream, w_grp, ch + 1);
                else if (g + gstep < total) {
                    if constexpr (WALK) {
                        stage(n_stream, n_grp, ch0);
                    } else {
                        const unsigned gn = g + gstep, sn = gn / groups;
                        stage(sn, gn - sn * groups, ch0);
                    }
                }
            }
            // Pair split.  Z[M - k] of pair q = i H + k2 (k = s + T i + 32 k2) is register (F-1-i) T + (T-1-k2) of lane
            // T - s of the same frame and arrives by SHFL.IDX; lane s = 0 pairs with its own lane0_partner<T>(q).
            uint32_t pw[16]; // split table from tensor memory: pairs 0..7, then 8..15
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                if constexpr (Cfg::TM) {
                    if ((q & 7) == 0) {
                        tm_ld<16>(tq + 96 + 2 * q, pw);
                        tm_wait_ld<16>(pw);
                    }
                }
                const int iq = q / H, k2 = q % H;
                const int uq = iq * T + k2;
                const f2 zp = sel2(s == 0, u[lane0_partner<T>(q)], shfl2(u[(F - 1 - iq) * T + (T - 1 - k2)], plane));
                f2 wsp;
                if constexpr (Cfg::TM) {
                    wsp = pk(u2f(pw[2 * (q & 7)]), u2f(pw[2 * (q & 7) + 1]));
                } else {
                    const f2x2 wq2 = prow[q / 2];
                    wsp = (q & 1) ? wq2.b : wq2.a;
                }
                const f2 A = add2(u[uq], conj2(zp));  // Z[k] + conj Z[M-k]
                const f2 Bv = sub2(u[uq], conj2(zp)); // Z[k] - conj Z[M-k]
                const f2 Tw = cmul2(Bv, wsp);
                const f2 xp = add2(A, Tw), xm = sub2(A, Tw);
                alo[q] = fm(lo(xp), lo(xp), fm(hi(xp), hi(xp), alo[q]));
                ahi[q] = fm(lo(xm), lo(xm), fm(hi(xm), hi(xm), ahi[q]));
            }
            { // bin M/2 = 16 T (lane s = 0, i = 0, k2 = T/2; self-paired): X = 2 conj Z
                const float a = lo(u[H]), b = hi(u[H]);
                amid = fm(JADE_FMUL(4.0f, a), a, fm(JADE_FMUL(4.0f, b), b, amid));
            }
        }

        const ColOut o = active ? col_out(P, stream, j) : ColOut{nullptr, nullptr};
        // reference orientation: bin k lands in row M - k
        if (!WANT_DB && !o.pix) continue; // padding lane group of the last column group: nothing to store
        uint32_t* p_lo = o.pix ? o.pix + (M - s) : nullptr; // bin k(q)   -> row M - k(q)
        uint32_t* p_hi = o.pix ? o.pix + s : nullptr;       // bin M-k(q) -> row k(q)
        float* d_lo = (WANT_DB && o.db) ? o.db + s : nullptr;
        float* d_hi = (WANT_DB && o.db) ? o.db + (M - s) : nullptr;
        // (two copies behind a launch-uniform branch: with the u8 palette the float -> integer conversion is the clamp)
        auto finish = [&](auto u8) {
            constexpr bool U8 = decltype(u8)::value;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int koff = T * (q / H) + 32 * (q % H); // k(q) - s
                emit_bin1<MIXK, WANT_DB, U8, MIXK == MIX_NONE>(alo[q], scale, (!WANT_DB || p_lo) ? p_lo - koff : nullptr, d_lo ? d_lo + koff : nullptr, P, s_pal);
                emit_bin1<MIXK, WANT_DB, U8, MIXK == MIX_NONE>(ahi[q], scale, (!WANT_DB || p_hi) ? p_hi + koff : nullptr, d_hi ? d_hi - koff : nullptr, P, s_pal);
            }
            if (s == 0) emit_bin1<MIXK, WANT_DB, U8, MIXK == MIX_NONE>(amid, scale, (!WANT_DB || o.pix) ? o.pix + M / 2 : nullptr, (WANT_DB && o.db) ? o.db + M / 2 : nullptr, P, s_pal);
        };
        if (P.pal_u8) finish(std::true_type{});
        else finish(std::false_type{});
    }
    if constexpr (Cfg::TM) {
        tm_fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32) tm_dealloc(tq, Cfg::TM_COLS); // warp 0: quadrant 0 = the allocation's base address
    }
}

} // namespace jade
