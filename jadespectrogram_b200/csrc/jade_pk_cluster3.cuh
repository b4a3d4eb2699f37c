// jade_pk_cluster3.cuh -- N = 65536 (BASELINE configs[4]: mono 192 kHz, hop 1024, log-frequency rows), one contributing channel,
// on a cluster of two CTAs: the three-pass register scheme of jade_pk3.cuh applied to the two 16384-point halves that the
// decimation-in-frequency split of jade_pk_cluster.cuh produces.
//
// CTA c of the cluster computes the bins of parity c of the M = 32768 point packed transform:
//     Z[2 k' + c] = sum_{m < 16384} d_c[m] W_M^{m (2 k' + c)},      d_0 = w z[m] + w' z[m + M/2],  d_1 = w z[m] - w' z[m + M/2]
// with m = t + 512 n1 (t = thread = c_lo + 16 c_hi), k' = k1 + 32 k2 + 1024 k3:
//     W_M^{m (2k' + c)} = W_64^{n1 (2 k1 + c)} . W_2048^{c_hi (2 k1 + c)} W_32^{c_hi k2} . W_M^{c_lo (2 k1 + c + 64 k2)} W_16^{c_lo k3}
//   pass 1 : thread t, 32-point DFT over n1 -- plain on CTA 0, TWISTED with the compile-time base W_64 on CTA 1: the odd-bin
//            factor W_M^m costs no multiply at all, its other two parts ride on the bases of pass 2 and pass 3;
//            samples come straight from global memory (two 8-byte loads per point), the window from TENSOR MEMORY: the whole
//            256 KB table (second half negated on CTA 1), 128 columns per warp x 4 warps per quadrant = all 512 columns;
//   pass 2 : thread (k1, c_lo), twisted 32-point DFT over c_hi, base W_2048^{2 k1 + c} (4 KB table in shared memory, read as a
//            half-warp broadcast); a half-warp owns row k1 and rewrites it in place as X[k1][k2][c_lo];
//   pass 3 : thread (k1, k2) twice, twisted 16-point DFTs over c_lo, base W_M^{2 k1 + c + 64 k2} (64 KB table in shared memory);
//            transform A = (i, kappa) and B = its mirror row / column, so that Z[k] and Z[M - k] (same parity, same CTA) meet in
//            one thread: split and power from registers.  CTA 1: B = (31 - i, 31 - kappa) for every lane; CTA 0: B =
//            (32 - i, 31 - kappa), lane 0 of the 32 half-warps takes the self-mirrored rows 16 and 0 (cf. pk3_lane).
// The power spectrum (16385 floats) overwrites the transform buffer once pass 3 has read it; rows (per-bin or log max-pool with
// the partial maxima of the two CTAs meeting through DSMEM) are those of stft_pkcl65536_kernel.
// Shared memory per CTA and frame: write / read / write / read of 128 KB (4096 wavefronts) + 68 KB of tables + 64 KB of powers,
// against ~12 800 in stft_pkcl65536_kernel; no separate twiddle multiplies (6 900 FMUL2 per frame there).
// Reference lines replaced: Spectrogram.cpp:50-56,137-145 (framing, window, spectrum::power), :107 (dB), :634-647 +
// CColorpalette.h:32-47 (pixel loop); the log rows are an extension.
#pragma once
#include "jade_pk3.cuh"
#include "jade_pk_cluster.cuh"

namespace jade {

struct PkCl3Cfg {
    static constexpr int MC = 16384;  // complex points per CTA
    static constexpr int N = 65536;
    static constexpr int THREADS = 512;
    static constexpr int TM_COLS = 512;
    static constexpr int off_row = 0;                       // 128 KB transform buffer, later the power spectrum (MC + 1 floats)
    static constexpr int off_tw3 = off_row + MC * 8;        // [4][1024] f2x2: entries 2 j, 2 j + 1 of transform (k1, k2) at [j][k1 + 32 k2]
    static constexpr int off_tw2 = off_tw3 + 4 * 1024 * 16; // [32][16] f2: the sixteen entries of row k1
    static constexpr int off_pal = off_tw2 + 32 * 16 * 8;
    static JADE_HD int off_rows(int npal) { return off_pal + (npal * 4 + 15) / 16 * 16; }
    static JADE_HD int off_part(int npal, int pooled_rows) { return off_rows(npal) + (pooled_rows * 8 + 15) / 16 * 16; }
    static JADE_HD int off_tm(int npal, int pooled_rows) { return off_part(npal, pooled_rows) + 2 * ((pooled_rows * 4 + 15) / 16 * 16); }
    static JADE_HD int smem_bytes(int npal, int pooled_rows) { return off_tm(npal, pooled_rows) + 16; }
};

// exponent (of W_65536) of entry e = 0..15 of a twisted 32-point pass with base W_65536^b (order of fft32_twisted: stage LEN = 2, 4,
// 8 x 2, 16 x 4, 32 x 8 with J = 0 .. LEN/4 - 1):  (32 / LEN) (b + J 65536 / 32)
JADE_HD int tw32_exponent_n(int b, int e)
{
    const int len = e == 0 ? 2 : e == 1 ? 4 : e < 4 ? 8 : e < 8 ? 16 : 32;
    const int j = e < 2 ? 0 : e < 4 ? e - 2 : e < 8 ? e - 4 : e - 8;
    return (32 / len) * (b + j * 2048);
}
// the same for a twisted 16-point pass (cf. tw16_exponent, which is in units of W_16384)
JADE_HD int tw16_exponent_n(int b, int e)
{
    const int len = e == 0 ? 2 : e == 1 ? 4 : e < 4 ? 8 : 16;
    const int j = e < 2 ? 0 : e < 4 ? e - 2 : e - 4;
    return (16 / len) * (b + j * 4096);
}
// pass-3 transforms of lane l = 16 h + i of warp w on CTA c (kappa = 2 w + h = 0..31)
JADE_HD Pk3Lane pkcl3_lane(int c, int warp, int lane)
{
    const int kappa = 2 * warp + (lane >> 4), i = lane & 15;
    Pk3Lane r;
    r.self = false;
    if (c == 1) {
        r.rowA = i, r.k2A = kappa, r.rowB = 31 - i, r.k2B = 31 - kappa;
    } else if (i != 0) {
        r.rowA = i, r.k2A = kappa, r.rowB = 32 - i, r.k2B = 31 - kappa;
    } else if (kappa < 16) {
        r.rowA = 16, r.k2A = kappa, r.rowB = 16, r.k2B = 31 - kappa;
    } else if (kappa > 16) {
        r.rowA = 0, r.k2A = kappa - 16, r.rowB = 0, r.k2B = 48 - kappa;
    } else {
        r.rowA = 0, r.k2A = 0, r.rowB = 0, r.k2B = 16, r.self = true;
    }
    return r;
}
// compile-time table of the twisted 32-point pass with base W_64 (pass 1 of the odd-bin CTA): entry of stage LEN, index J is
// W_64^{(32 / LEN)(1 + 2 J)}
JADE_DEVICE void fft32_twisted_w64(f2* u)
{
#define JADE_W64(e) pk(cos64(e), -sin64(e))
    f2x2 tb[8];
    tb[0].a = JADE_W64(16), tb[0].b = JADE_W64(8);
    tb[1].a = JADE_W64(4), tb[1].b = JADE_W64(12);
    tb[2].a = JADE_W64(2), tb[2].b = JADE_W64(6);
    tb[3].a = JADE_W64(10), tb[3].b = JADE_W64(14);
    tb[4].a = JADE_W64(1), tb[4].b = JADE_W64(3);
    tb[5].a = JADE_W64(5), tb[5].b = JADE_W64(7);
    tb[6].a = JADE_W64(9), tb[6].b = JADE_W64(11);
    tb[7].a = JADE_W64(13), tb[7].b = JADE_W64(15);
#undef JADE_W64
    fft32_twisted(u, tb);
}

template <int MIXK>
JADE_CLUSTER_KERNEL(PkCl3Cfg::THREADS) stft_pkcl3_kernel(const KParams P)
{
    static_assert(MIXK == MIX_NONE, "one contributing channel");
    using Cfg = PkCl3Cfg;
    constexpr int MC = Cfg::MC, N = Cfg::N, THREADS = Cfg::THREADS;
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    f2* buf = reinterpret_cast<f2*>(sm + Cfg::off_row);
    float* s_spec = reinterpret_cast<float*>(sm + Cfg::off_row);
    f2x2* s_tw3 = reinterpret_cast<f2x2*>(sm + Cfg::off_tw3);
    f2* s_tw2 = reinterpret_cast<f2*>(sm + Cfg::off_tw2);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);
    i2* s_rows = reinterpret_cast<i2*>(sm + Cfg::off_rows(P.npal));
    float* s_part = reinterpret_cast<float*>(sm + Cfg::off_part(P.npal, P.pooled ? P.R : 0));
    uint32_t* s_tm = reinterpret_cast<uint32_t*>(sm + Cfg::off_tm(P.npal, P.pooled ? P.R : 0));
    const int part_stride = P.pooled ? (P.R * 4 + 15) / 16 * 4 : 0; // floats per partial-maxima buffer

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int c = (int)cl_rank(); // parity of the bins this CTA computes
    const int c_lo = lane & 15, row2 = 2 * warp + (lane >> 4);
    const Pk3Lane L3 = pkcl3_lane(c, warp, lane);

    for (int i = t; i < P.npal; i += THREADS) s_pal[i] = P.palette[i];
    if (P.pooled)
        for (int i = t; i < P.R; i += THREADS) s_rows[i] = P.row_bins[i];
    // twisted tables (every value one correctly rounded root of unity, P.twP[e] = W_N^e)
    for (int i = t; i < 32 * 16; i += THREADS) { // pass 2: base W_2048^{2 k1 + c} = W_N^{32 (2 k1 + c)}
        const cpx a = P.twP[tw32_exponent_n(32 * (2 * (i >> 4) + c), i & 15)];
        s_tw2[i] = pk(a.x, a.y);
    }
    for (int i = t; i < 4 * 1024; i += THREADS) { // pass 3: base W_M^{2 k1 + c + 64 k2} = W_N^{2 (2 k1 + c + 64 k2)}
        const int j = i >> 10, combo = i & 1023, b = 2 * (2 * (combo & 31) + c + 64 * (combo >> 5));
        const cpx a0 = P.twP[tw16_exponent_n(b, 2 * j)], a1 = P.twP[tw16_exponent_n(b, 2 * j + 1)];
        f2x2 v;
        v.a = pk(a0.x, a0.y);
        v.b = pk(a1.x, a1.y);
        s_tw3[i] = v;
    }
    if (t < 32) tm_alloc(s_tm, Cfg::TM_COLS);
    tm_fence_before_sync();
    __syncthreads();
    tm_fence_after_sync();
    const uint32_t tq = tm_quadrant_base(*s_tm) + 128u * ((unsigned)warp >> 2);
    {
        // window of this thread's points m = t + 512 n1 and m + 16384: chunk n1 / 4 (16 columns) = 4 x { w[2m], w[2m+1], w'[..], w'[..] }
        uint32_t r[16];
#pragma unroll 1
        for (int ch = 0; ch < 8; ++ch) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = t + 512 * (4 * ch + i);
                r[4 * i] = f2u(P.window[2 * m]);
                r[4 * i + 1] = f2u(P.window[2 * m + 1]);
                r[4 * i + 2] = f2u(c ? -P.window[2 * m + 32768] : P.window[2 * m + 32768]); // (the odd-bin CTA subtracts the second half)
                r[4 * i + 3] = f2u(c ? -P.window[2 * m + 32769] : P.window[2 * m + 32769]);
            }
            tm_st<16>(tq + 16 * ch, r);
        }
        tm_wait_st();
    }
    // split twiddle W_N^{2 k' + c} = W_N^{2 kb + c} W_32^q (k' = kb + 1024 q); the self-pairing lane's slots q >= 8 are the bins
    // k' = 512 + 1024 (q - 8) = (512 - 8192) + 1024 q
    const int kb_lo = L3.rowA + 32 * L3.k2A, kb_hi = L3.self ? 512 - 8192 : kb_lo;
    f2 ws_lo, ws_hi;
    {
        const cpx a = P.twP[2 * kb_lo + c], b = P.twP[L3.self ? 15360 : 2 * kb_lo + c];
        ws_lo = pk(a.x, a.y);
        ws_hi = L3.self ? pk(b.x, -b.y) : pk(b.x, b.y); // W_N^{-15360} = conj W_N^{15360}
    }
    const peer_addr peer_part = cl_map(s_part, (unsigned)(c ^ 1));
    tm_fence_before_sync();
    __syncthreads();
    tm_fence_after_sync();
    grid_dep_wait();
    cl_sync(); // both CTAs of the cluster are resident before the first remote access

    const unsigned total = (unsigned)P.ncols * (unsigned)P.nstreams;
    const unsigned nclusters = gridDim.x / 2, cid = blockIdx.x / 2;
    int ch0, ch1;
    channel_range(P, ch0, ch1);
    unsigned parity = 0; // which partial-maxima buffer this frame uses
    const int rotA = L3.rowA & 7, rotB = L3.rowB & 7;

    for (unsigned g = cid; g < total; g += nclusters, parity ^= 1u) {
        const int stream = (int)(g / (unsigned)P.ncols);
        const long long j = P.first_col + (g - (unsigned)stream * (unsigned)P.ncols);
        const long long st = frame_start(P, j);
        const long long ns = P.nsamples;
        const float* x = P.samples + stream * P.stream_stride + ch0 * P.channel_stride;
        f2 v[32];
        // ---- pass 1: samples x window, DIF stage, 32-point DFT over n1, row k1 <- Y[t][k1]
        {
            const bool fast = st >= 0 && st + N <= ns && (reinterpret_cast<uintptr_t>(x + st) & 7) == 0;
            auto point = [&](int n1, f2 xa, f2 xb, const uint32_t* w) {
                const f2 a = mul2(xa, pk(u2f(w[0]), u2f(w[1])));
                v[brev(n1, 5)] = fma2(xb, pk(u2f(w[2]), u2f(w[3])), a); // (w' carries the sign of the DIF stage)
            };
            if (fast) {
                const f2* xz = reinterpret_cast<const f2*>(x + st) + t;
                uint32_t wq[2][16];
                f2 xa[2][4], xb[2][4];
                auto fetch = [&](int ch) {
                    tm_ld<16>(tq + 16 * ch, wq[ch & 1]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        xa[ch & 1][i] = xz[512 * (4 * ch + i)];
                        xb[ch & 1][i] = xz[512 * (4 * ch + i) + MC];
                    }
                };
                fetch(0);
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    tm_wait_ld<16>(wq[ch & 1]);
                    if (ch < 7) fetch(ch + 1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) point(4 * ch + i, xa[ch & 1][i], xb[ch & 1][i], wq[ch & 1] + 4 * i);
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    uint32_t wq[16];
                    tm_ld<16>(tq + 16 * ch, wq);
                    tm_wait_ld<16>(wq);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int m = t + 512 * (4 * ch + i);
                        const cpx za = load_pair_guarded(x, st + 2 * m, ns), zb = load_pair_guarded(x, st + 2 * m + 32768, ns);
                        point(4 * ch + i, pk(za.x, za.y), pk(zb.x, zb.y), wq + 4 * i);
                    }
                }
            }
            if (c == 0) fft32_pk(v);
            else fft32_twisted_w64(v);
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) buf[512 * k1 + t] = v[k1];
        }
        __syncthreads();
        // ---- pass 2: row row2: twisted 32-point DFT over c_hi, in place as [k2][c_lo] (rotated 16-byte chunks)
        {
            f2* ra = buf + 512 * row2;
#pragma unroll
            for (int q = 0; q < 32; ++q) v[brev(q, 5)] = ra[c_lo + 16 * q];
            __syncwarp(); // the row has been read by the sixteen lanes that own it
            fft32_twisted(v, reinterpret_cast<const f2x2*>(s_tw2 + 16 * row2));
            const int off = 2 * (((c_lo >> 1) + (row2 & 7)) & 7) + (c_lo & 1);
#pragma unroll
            for (int k2 = 0; k2 < 32; ++k2) ra[16 * k2 + off] = v[k2];
        }
        __syncthreads();
        // ---- pass 3: transforms A and B: twisted 16-point DFTs over c_lo
        {
            const f2* ga = buf + 512 * L3.rowA + 16 * L3.k2A;
            const f2* gb = buf + 512 * L3.rowB + 16 * L3.k2B;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const f2x2 a = *reinterpret_cast<const f2x2*>(ga + 2 * ((q + rotA) & 7));
                const f2x2 b = *reinterpret_cast<const f2x2*>(gb + 2 * ((q + rotB) & 7));
                v[brev4(2 * q)] = a.a;
                v[brev4(2 * q + 1)] = a.b;
                v[16 + brev4(2 * q)] = b.a;
                v[16 + brev4(2 * q + 1)] = b.b;
            }
            __syncthreads(); // the buffer is free: the power spectrum goes into it
#pragma unroll
            for (int d = 0; d < 2; ++d) {
                const f2x2* tw = s_tw3 + (d == 0 ? L3.rowA + 32 * L3.k2A : L3.rowB + 32 * L3.k2B);
                f2 w[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const f2x2 tv = tw[1024 * e];
                    w[2 * e] = tv.a;
                    w[2 * e + 1] = tv.b;
                }
                fft16_twisted(v + 16 * d, w);
            }
        }
        // ---- split + power: slot q pairs zk = A[q] with zp = B[15 - q] (bins 2 k' + c, k' = kb + 1024 q, and M - bin); the power of
        // bin 2 i + c goes to s_spec[i] (index MC for bin N/2)
        {
            const f2* ua = v;
            const f2* ub = v + 16;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                f2 zk = ua[q], zp = ub[15 - q];
                if (q >= 8) zk = sel2(L3.self, ub[q - 8], zk);
                zp = sel2(L3.self, q < 8 ? ua[(16 - q) & 15] : ub[23 - q], zp);
                const f2 w = cmul2(q < 8 ? ws_lo : ws_hi, pk(cos32(q), -sin32(q))); // W_N^{2k'+c} ; -i W = (w.y, -w.x)
                const f2 A = add2(zk, conj2(zp));
                const f2 Bv = sub2(zk, conj2(zp));
                const f2 T = cmul2(Bv, pk(hi(w), -lo(w)));
                const f2 xp = add2(A, T), xm = sub2(A, T);
                const int k = (q < 8 ? kb_lo : kb_hi) + 1024 * q;
                s_spec[k] = fm(lo(xp), lo(xp), JADE_FMUL(hi(xp), hi(xp)));
                s_spec[MC - c - k] = fm(lo(xm), lo(xm), JADE_FMUL(hi(xm), hi(xm)));
            }
            if (L3.self) { // k' = 8192: bin N/4 pairs with itself, X = 2 conj Z
                const float a = lo(ua[8]), b = hi(ua[8]);
                s_spec[MC / 2] = fm(JADE_FMUL(4.0f, a), a, JADE_FMUL(JADE_FMUL(4.0f, b), b));
            }
        }
        __syncthreads();

        // ---- rows (as in stft_pkcl65536_kernel)
        const ColOut o = col_out(P, stream, j);
        const int nown = c == 0 ? MC + 1 : MC; // bins 2 i + c, i < nown
        if (o.db || !P.pooled) { // per-bin outputs: every CTA emits the bins of its parity
            for (int i = t; i < nown; i += THREADS) {
                const int k = 2 * i + c;
                const float d = to_db(s_spec[i], P.db_precise);
                if (o.db) o.db[k] = d;
                if (!P.pooled && o.pix && k >= P.k_lo && k < P.k_hi) o.pix[P.flip ? (P.k_hi - 1 - k) : (k - P.k_lo)] = colour_of(d, P, s_pal);
            }
        }
        if (P.pooled && o.pix) {
            // log max-pool: every CTA reduces the bins of its parity for EVERY row (-1 where it has none: powers are >= 0) ...
            float* part = s_part + parity * part_stride;
            for (int r = t; r < P.R; r += THREADS) {
                const i2 rb = s_rows[r];
                int i = (rb.lo + 1 - c) >> 1;         // first index with 2 i + c >= lo
                const int i1 = (rb.hi - c + 1) >> 1; // first index with 2 i + c >= hi
                float m0 = -1.0f, m1 = -1.0f, m2 = -1.0f, m3 = -1.0f;
                for (; i + 3 < i1; i += 4) {
                    m0 = fmaxf(m0, s_spec[i]);
                    m1 = fmaxf(m1, s_spec[i + 1]);
                    m2 = fmaxf(m2, s_spec[i + 2]);
                    m3 = fmaxf(m3, s_spec[i + 3]);
                }
                for (; i < i1; ++i) m0 = fmaxf(m0, s_spec[i]);
                part[r] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            }
            cl_sync(); // ... the partial maxima of both CTAs are visible cluster-wide ...
            // ... and each CTA finishes alternate groups of 32 rows: its own partial and the peer's through DSMEM
            for (int r = t; r < P.R; r += THREADS) {
                if (((r >> 5) & 1) != c) continue;
                const float mx = fmaxf(part[r], cl_ld_f32(cl_at(peer_part, (int)(parity * part_stride + r) * 4)));
                o.pix[P.flip ? (P.R - 1 - r) : r] = colour_of(to_db(mx, P.db_precise), P, s_pal);
            }
        } else {
            __syncthreads(); // the power spectrum has been read before the next frame overwrites the buffer
        }
    }
    cl_sync(); // nobody leaves while the peer may still read its shared memory
    tm_fence_before_sync();
    __syncthreads();
    if (t < 32) tm_dealloc(*s_tm, Cfg::TM_COLS);
}

} // namespace jade
