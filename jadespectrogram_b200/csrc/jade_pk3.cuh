// jade_pk3.cuh -- N = 16384 (BASELINE configs[2]: mono 96 kHz, hop 4096, Blackman-Harris), one contributing channel:
// the M = 8192 complex points z[m] = x[2m] + i x[2m+1] of a frame as THREE register passes 32 x 16 x 16 by one CTA of
// 256 threads (32 complex values per thread), the real-FFT split and the epilogue straight from registers.
//
//   m = t + 256 n1            (t = thread = c_lo + 16 c_hi,  n1 = 0..31)
//   k = k1 + 32 k2 + 512 k3   (k1 = 0..31, k2, k3 = 0..15)
//   W_M^{mk} = W_32^{n1 k1} . W_512^{c_hi k1} W_16^{c_hi k2} . W_M^{c_lo (k1 + 32 k2)} W_16^{c_lo k3}
//
//   pass 1 : thread t, 32-point DFT over n1 (window fused into its first stage), result row k1 of the shared matrix
//            X[k1][t] -- the thread's own column of the staged frame, so the pass is in place;
//   pass 2 : thread (k1, c_lo) twice, TWISTED 16-point DFT over c_hi with base W_512^{k1} (the inter-pass twiddle rides
//            on the butterflies, cf. fft32_twisted); a half-warp owns two rows k1 and rewrites them in place as
//            X[k1][k2][c_lo];
//   pass 3 : thread (k1, k2) twice, twisted 16-point DFT over c_lo with base W_M^{k1 + 32 k2}: Z[k1 + 32 k2 + 512 k3]
//            in register k3.  Lane i of a half-warp takes A = (k1 = i, k2 = kappa) and B = (32 - i, 15 - kappa): bin k
//            of A and its mirror M - k then sit in the SAME thread (registers q and 15 - q), so the real-FFT split
//                X[k] = A' + T,  X[M-k] = conj(A' - T),  A' = Z[k] + conj Z[M-k],  T = -i W_N^k (Z[k] - conj Z[M-k])
//            needs no exchange, and the 15 lanes i = 1..15 of a half-warp emit 15 consecutive bins per store.  Lane 0 of
//            the sixteen half-warps takes the self-mirrored rows: (16, kappa) + (16, 15 - kappa) for kappa < 8,
//            (0, j) + (0, 16 - j) for kappa = 8 + j; only (0, 0) + (0, 8) (kappa = 8) pairs inside its own transforms.
//
// Shared memory carries the frame five times per transform (read, write, read, write, read: 2560 wavefronts) instead of
// nine in stft_pkcta_kernel<8> (column pass, row pass with its transpose, split), the window (64 KB, formerly re-read
// through L1 for every frame) and both twisted tables are per-thread constants held in TENSOR MEMORY (jade_tmem.cuh:
// 128 columns per warp, two warps per quadrant, two CTAs per SM = all 512 columns), and the next frame is staged by
// the TMA engine as soon as pass 3 has picked up its inputs.
// Reference lines replaced: Spectrogram.cpp:50-56,137-145 (framing, window, spectrum::power), :107 (dB), :634-647 +
// CColorpalette.h:32-47 (pixel loop).
#pragma once
#include <type_traits>

#include "jade_pkz.cuh"
#include "jade_tmem.cuh"

namespace jade {

// ---- twisted 16-point pass: bit-reversed input, natural-order output; w[0..7] = table entries of this transform:
// stage LEN uses u^{16/LEN} W_LEN^J, J < LEN/4 from the table (J >= LEN/4 is -i times entry J - LEN/4)
template <int LEN, int BASE, int J>
JADE_DEVICE void tw16_inner(f2* a, const f2* tw)
{
    if constexpr (J < LEN / 2) {
        constexpr int Q = (LEN >= 4) ? LEN / 4 : 1;
        if constexpr (J < Q) bfly_w(a[BASE + J], a[BASE + J + LEN / 2], tw[J]);
        else bfly_wmi(a[BASE + J], a[BASE + J + LEN / 2], tw[J - Q]);
        tw16_inner<LEN, BASE, J + 1>(a, tw);
    }
}
template <int LEN, int BASE>
JADE_DEVICE void tw16_blocks(f2* a, const f2* tw)
{
    if constexpr (BASE < 16) {
        tw16_inner<LEN, BASE, 0>(a, tw);
        tw16_blocks<LEN, BASE + LEN>(a, tw);
    }
}
JADE_DEVICE void fft16_twisted(f2* u, const f2* w)
{
    tw16_blocks<2, 0>(u, w);
    tw16_blocks<4, 0>(u, w + 1);
    tw16_blocks<8, 0>(u, w + 2);
    tw16_blocks<16, 0>(u, w + 4);
}
// exponent (of W_16384) of table entry e = 0..7 of a twisted 16-point pass whose base is W_16384^b:
// stage LEN = 2, 4, 8, 8, 16 x 4 with J = 0, 0, 0, 1, 0..3:  (16 / LEN) (b + J 16384 / 16)
JADE_HD int tw16_exponent(int b, int e)
{
    const int len = e == 0 ? 2 : e == 1 ? 4 : e < 4 ? 8 : 16;
    const int j = e < 2 ? 0 : e < 4 ? e - 2 : e - 4;
    return (16 / len) * (b + j * 1024);
}
JADE_HD constexpr int brev4(int v) { return ((v & 1) << 3) | ((v & 2) << 1) | ((v & 4) >> 1) | ((v & 8) >> 3); }

struct Pk3Cfg {
    static constexpr int M = 8192, N = 16384, B = M + 1, THREADS = 256;
    static constexpr int TM_COLS = 256; // tensor memory per CTA: 128 columns per warp of a quadrant (window 64, pass-2 table 32, pass-3 table 32)
    static constexpr int off_row = 0;
    static constexpr int off_pal = off_row + M * 8;
    static JADE_HD int off_bar(int npal) { return off_pal + ((npal + 1) * 4 + 15) / 16 * 16; } // table + the `>= m_Max` entry (pkz_emit)
    static JADE_HD int smem_bytes(int npal) { return off_bar(npal) + 16; }
};

// the two transforms (row k1, column k2) of lane l = 16 h + i of warp w in pass 3, and whether the lane is the one that
// pairs inside its own transforms
struct Pk3Lane {
    int rowA, k2A, rowB, k2B;
    bool self;
};
JADE_HD Pk3Lane pk3_lane(int warp, int lane)
{
    const int kappa = 2 * warp + (lane >> 4), i = lane & 15;
    Pk3Lane r;
    r.self = false;
    if (i != 0) {
        r.rowA = i, r.k2A = kappa, r.rowB = 32 - i, r.k2B = 15 - kappa;
    } else if (kappa < 8) {
        r.rowA = 16, r.k2A = kappa, r.rowB = 16, r.k2B = 15 - kappa;
    } else if (kappa > 8) {
        r.rowA = 0, r.k2A = kappa - 8, r.rowB = 0, r.k2B = 24 - kappa;
    } else {
        r.rowA = 0, r.k2A = 0, r.rowB = 0, r.k2B = 8, r.self = true;
    }
    return r;
}

enum { PK3_STAGED = 0, PK3_GUARD = 1 };

// LD = PK3_STAGED: every frame of the launch is interior and starts on a multiple of 4 samples: the TMA engine stages it;
// PK3_GUARD: boundary frames / unaligned geometries, the thread fills its own column with bounds-checked loads first.  The
// arithmetic is the same, so streaming, batch and sharded renderings agree bit for bit.  (One instantiation that decides per
// frame is 7 % slower on interior frames -- registers -- than the pair, profiles/r02c_pk3_variants.txt.)
// U8: instantiation for P.pal_u8 palettes (colour_of_lg1), picked by launch_one for the pixel-only launches
template <bool WANT_DB, int LD, bool U8 = false>
JADE_KERNEL(Pk3Cfg::THREADS, 2) stft_pk3_kernel(const KParams P)
{
    using Cfg = Pk3Cfg;
    constexpr int M = Cfg::M, N = Cfg::N;
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    f2* buf = reinterpret_cast<f2*>(sm + Cfg::off_row);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm + Cfg::off_bar(P.npal));
    uint32_t* s_tm = reinterpret_cast<uint32_t*>(bar + 1);

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // pass 2: two rows per half-warp
    const int c_lo = lane & 15, row2 = 4 * warp + 2 * (lane >> 4);
    const Pk3Lane L3 = pk3_lane(warp, lane);

    for (int i = t; i <= P.npal; i += Cfg::THREADS) s_pal[i] = P.palette[i < P.npal ? i : P.ci_hi];
    if (t == 0) mbar_init(bar, 1);
    if (t < 32) tm_alloc(s_tm, Cfg::TM_COLS);
    tm_fence_before_sync();
    __syncthreads();
    tm_fence_after_sync();
    const uint32_t tq = tm_quadrant_base(*s_tm) + 128u * ((unsigned)warp >> 2);
    {
        // ---- this thread's constants -> tensor memory.  Columns 0..63: window pairs in the order pass 1 consumes them: chunk c
        // (16 columns) = points n1 = 4c .. 4c+3, then n1 = 16 + 4c .. 16 + 4c + 3;  64..95: pass-2 tables (row2, row2 + 1);
        // 96..127: pass-3 tables (A, B).  Every table value is one correctly rounded root of unity (P.twP[e] = W_N^e).
        uint32_t r[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ma = t + 256 * (4 * c + i), mb = ma + 256 * 16;
                r[2 * i] = f2u(P.window[2 * ma]);
                r[2 * i + 1] = f2u(P.window[2 * ma + 1]);
                r[8 + 2 * i] = f2u(P.window[2 * mb]);
                r[8 + 2 * i + 1] = f2u(P.window[2 * mb + 1]);
            }
            tm_st<16>(tq + 16 * c, r);
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            // base exponent (of W_N): pass 2 W_512^{k1} = W_N^{32 k1};  pass 3 W_M^{k1 + 32 k2} = W_N^{2 (k1 + 32 k2)}
            const int b = d == 0 ? 32 * row2 : d == 1 ? 32 * (row2 + 1) : d == 2 ? 2 * (L3.rowA + 32 * L3.k2A) : 2 * (L3.rowB + 32 * L3.k2B);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const cpx a = P.twP[tw16_exponent(b, e)];
                r[2 * e] = f2u(a.x);
                r[2 * e + 1] = f2u(a.y);
            }
            tm_st<16>(tq + 64 + 16 * d, r);
        }
        tm_wait_st();
    }
    // split twiddles W_N^k = W_N^{kb} W_32^q (kb = k - 512 q): one base per thread; the self-pairing lane's slots q >= 8 are the
    // bins 256 + 512 (q - 8) = (256 - 4096) + 512 q
    const int kb_lo = L3.rowA + 32 * L3.k2A, kb_hi = L3.self ? 256 - 4096 : kb_lo;
    f2 ws_lo, ws_hi;
    {
        const cpx a = P.twP[kb_lo], b = P.twP[L3.self ? 3840 : kb_lo];
        ws_lo = pk(a.x, a.y);
        ws_hi = L3.self ? pk(b.x, -b.y) : pk(b.x, b.y); // W_N^{-3840} = conj W_N^{3840}
    }
    tm_fence_before_sync();
    __syncthreads();
    tm_fence_after_sync();
    grid_dep_wait();

    const unsigned total = (unsigned)P.ncols * (unsigned)P.nstreams;
    int ch0, ch1;
    channel_range(P, ch0, ch1);
    // the CTA's frames g, g + gridDim.x, ... as (stream, column), advanced without a division per frame
    const unsigned w_dq = gridDim.x / (unsigned)P.ncols, w_dr = gridDim.x - w_dq * (unsigned)P.ncols;
    auto advance = [&](unsigned& stream_, unsigned& col_) {
        stream_ += w_dq;
        col_ += w_dr;
        if (col_ >= (unsigned)P.ncols) {
            col_ -= (unsigned)P.ncols;
            ++stream_;
        }
    };
    auto frame_of = [&](unsigned col_, long long& j_, long long& st_, bool& staged_) {
        j_ = P.first_col + col_;
        st_ = frame_start(P, j_);
        staged_ = LD == PK3_STAGED || (P.aligned4 && st_ >= 0 && st_ + N <= P.nsamples);
    };
    auto stage = [&](unsigned stream_, unsigned col_) { // after a __syncthreads(): the frame's 64 KB -> buf (thread 0; nothing for a boundary frame)
        long long j_, st_;
        bool staged_;
        frame_of(col_, j_, st_, staged_);
        const float* src = P.samples + (long long)stream_ * P.stream_stride + ch0 * P.channel_stride + st_;
        if (t == 0 && staged_) {
            mbar_expect_tx(bar, M * 8);
#pragma unroll
            for (int i = 0; i < 8; ++i) bulk_copy_issue(buf + 1024 * i, src + 2048 * i, 8192, bar);
        }
#if defined(JADE_EMU)
        __syncthreads();
#endif
    };
    unsigned copies = 0;
    unsigned w_stream = blockIdx.x / (unsigned)P.ncols, w_col = blockIdx.x - w_stream * (unsigned)P.ncols;
    unsigned n_stream = w_stream, n_col = w_col; // the CTA's next frame
    if (blockIdx.x < total) stage(w_stream, w_col);

    // exchange-2 addresses (f2 words inside a row): element (k2, c) of a row lives at 16 k2 + 2 ((c/2 + row) & 7) + (c & 1)
    const int rotA = L3.rowA & 7, rotB = L3.rowB & 7;

    for (unsigned g = blockIdx.x; g < total; g += gridDim.x, w_stream = n_stream, w_col = n_col) {
        const int stream = (int)w_stream;
        long long j, st;
        bool staged;
        frame_of(w_col, j, st, staged);
        advance(n_stream, n_col);
        f2 v[32];
        // ---- pass 1: samples x window, 32-point DFT over n1, row k1 <- Y[t][k1]
        {
            const float* x = P.samples + stream * P.stream_stride + ch0 * P.channel_stride;
            if (staged) {
                mbar_wait(bar, copies & 1u);
                ++copies;
            } else {
                // boundary frame / unaligned geometry: the thread fills its own column with bounds-checked loads (zeros outside the
                // signal) and then runs the same pass as for a staged frame -- one rolled loop instead of a second copy of the pass
#pragma unroll 1
                for (int n1 = 0; n1 < 32; ++n1) {
                    const cpx z = load_pair_guarded(x, st + 2 * (t + 256 * n1), P.nsamples);
                    buf[t + 256 * n1] = pk(z.x, z.y);
                }
            }
            uint32_t wq[2][16];
            tm_ld<16>(tq, wq[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tm_wait_ld<16>(wq[c & 1]);
                if (c < 3) tm_ld<16>(tq + 16 * (c + 1), wq[(c + 1) & 1]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int n1 = 4 * c + i, ma = t + 256 * n1, mb = ma + 256 * 16;
                    const uint32_t* w = wq[c & 1];
                    win_stage1<32>(v, n1, buf[ma], pk(u2f(w[2 * i]), u2f(w[2 * i + 1])), buf[mb], pk(u2f(w[8 + 2 * i]), u2f(w[8 + 2 * i + 1])));
                }
            }
            fft32_pk_after_stage1(v);
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) buf[256 * k1 + t] = v[k1];
        }
        __syncthreads();
        // ---- pass 2: rows row2, row2 + 1: twisted 16-point DFTs over c_hi, in place as [k2][c_lo] (rotated 16-byte chunks)
        {
            f2* ra = buf + 256 * row2;
            uint32_t tw[2][16];
            tm_ld<16>(tq + 64, tw[0]);
            tm_ld<16>(tq + 80, tw[1]);
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                v[brev4(c)] = ra[c_lo + 16 * c];
                v[16 + brev4(c)] = ra[256 + c_lo + 16 * c];
            }
            tm_wait_ld<16>(tw[0]);
            tm_tie<16>(tw[1]);
            __syncwarp(); // both rows have been read by the sixteen lanes that own them
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                f2 w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = pk(u2f(tw[d][2 * e]), u2f(tw[d][2 * e + 1]));
                fft16_twisted(v + 16 * d, w);
                const int rot = (row2 + d) & 7;
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) ra[256 * d + 16 * k2 + 2 * (((c_lo >> 1) + rot) & 7) + (c_lo & 1)] = v[16 * d + k2];
            }
        }
        __syncthreads();
        // ---- pass 3: transforms A and B: twisted 16-point DFTs over c_lo
        {
            uint32_t tw[2][16];
            tm_ld<16>(tq + 96, tw[0]);
            tm_ld<16>(tq + 112, tw[1]);
            const f2* ga = buf + 256 * L3.rowA + 16 * L3.k2A;
            const f2* gb = buf + 256 * L3.rowB + 16 * L3.k2B;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const f2x2 a = *reinterpret_cast<const f2x2*>(ga + 2 * ((c + rotA) & 7));
                const f2x2 b = *reinterpret_cast<const f2x2*>(gb + 2 * ((c + rotB) & 7));
                v[brev4(2 * c)] = a.a;
                v[brev4(2 * c + 1)] = a.b;
                v[16 + brev4(2 * c)] = b.a;
                v[16 + brev4(2 * c + 1)] = b.b;
            }
            tm_wait_ld<16>(tw[0]);
            tm_tie<16>(tw[1]);
            __syncthreads(); // the matrix is free: stage the frame this CTA transforms next (covered by the rest of this one)
            if (g + gridDim.x < total) stage(n_stream, n_col);
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                f2 w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = pk(u2f(tw[d][2 * e]), u2f(tw[d][2 * e + 1]));
                fft16_twisted(v + 16 * d, w);
            }
        }
        // ---- split + epilogue: slot q pairs zk = A[q] with zp = B[15 - q] (bins k = kb + 512 q and M - k)
        const ColOut o = col_out(P, stream, j);
        const f2* ua = v;
        const f2* ub = v + 16;
        // The self-pairing lane -- kappa = 8, lane 0 of warp 4: rows (0, 0) and (0, 8) -- pairs A[q] with A[16 - q] in its slots q < 8
        // and B[q - 8] with B[23 - q] in its slots q >= 8.  Two copies of the loop behind a warp-uniform branch: only warp 4 executes
        // the selects (48 per thread and frame).  (A re-sort of that warp's registers in front of ONE loop costs spills and 6 %.)
        auto split_emit = [&](auto self_warp) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                f2 zk = ua[q], zp = ub[15 - q];
                f2 wb = q < 8 ? ws_lo : ws_hi;
                int kb = q < 8 ? kb_lo : kb_hi;
                if constexpr (decltype(self_warp)::value) {
                    if (q >= 8) zk = sel2(L3.self, ub[q - 8], zk);
                    zp = sel2(L3.self, q < 8 ? ua[(16 - q) & 15] : ub[23 - q], zp);
                } else {
                    wb = ws_lo; // (regular lanes: ws_hi == ws_lo, kb_hi == kb_lo)
                    kb = kb_lo;
                }
                const f2 w = cmul2(wb, pk(cos32(q), -sin32(q))); // W_N^k ; -i W_N^k = (w.y, -w.x)
                const f2 A = add2(zk, conj2(zp));
                const f2 Bv = sub2(zk, conj2(zp));
                const f2 T = cmul2(Bv, pk(hi(w), -lo(w)));
                const f2 xp = add2(A, T), xm = sub2(A, T);
                // + 1e-11 (Spectrogram.cpp:36,107) rides on the power FMAs; dB, palette index by one FFMA and one integer clamp (pkz_emit)
                const float plo = fm(lo(xp), lo(xp), fm(hi(xp), hi(xp), 1e-11f));
                const float phi = fm(lo(xm), lo(xm), fm(hi(xm), hi(xm), 1e-11f));
                const int k = kb + 512 * q;
                pkz_emit<WANT_DB, U8>(plo, o.pix ? o.pix + (M - k) : nullptr, (WANT_DB && o.db) ? o.db + k : nullptr, P, s_pal); // bin k -> row M - k
                pkz_emit<WANT_DB, U8>(phi, o.pix ? o.pix + k : nullptr, (WANT_DB && o.db) ? o.db + (M - k) : nullptr, P, s_pal);
            }
        };
        if (warp == 4) split_emit(std::true_type{});
        else split_emit(std::false_type{});
        if (L3.self) { // bin M/2 (self-paired, A[8] of that lane): X = 2 conj Z
            const float a = lo(ua[8]), b = hi(ua[8]);
            const float p = fm(JADE_FMUL(4.0f, a), a, fm(JADE_FMUL(4.0f, b), b, 1e-11f));
            pkz_emit<WANT_DB, U8>(p, o.pix ? o.pix + M / 2 : nullptr, (WANT_DB && o.db) ? o.db + M / 2 : nullptr, P, s_pal);
        }
    }
    tm_fence_before_sync();
    __syncthreads();
    if (t < 32) tm_dealloc(*s_tm, Cfg::TM_COLS);
}

} // namespace jade
