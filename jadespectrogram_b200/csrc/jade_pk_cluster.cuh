// jade_pk_cluster.cuh -- N = 65536 (BASELINE configs[4]: mono 192 kHz, hop 1024, log-frequency rows) on a thread-block CLUSTER of
// two CTAs with distributed shared memory.
//
// The M = 32768 complex points z[m] = x[2m] + i x[2m+1] of the packed real transform are 256 KB -- more than one SM's shared
// memory.  They are split by ONE decimation-in-frequency stage, fused with the window multiply into the load:
//     CTA 0:  y[m] =  w z[m] + w' z[m + M/2]                 -> FFT_16384(y) = Z[2 k']        (even bins of Z)
//     CTA 1:  y[m] = (w z[m] - w' z[m + M/2]) W_M^m           -> FFT_16384(y) = Z[2 k' + 1]    (odd bins)
// so the two SMs of a cluster each run a 16384-point transform (128 KB) in their own shared memory AT THE SAME TIME.  The
// real-FFT split pairs Z[k] with Z[M - k] -- same parity -- so it stays inside a CTA, and so does the power spectrum
// (CTA c holds the bins of parity c).  The only thing that crosses the cluster is the log max-pool: every CTA reduces its
// own bins of a row, and the two partial maxima (1080 floats) meet through DSMEM -- one cluster barrier per frame.
// (A first version split even / odd SAMPLES and exchanged 128 KB per frame and direction with st / ld.shared::cluster: 2.2 M
// frames/s against 3.4 M of the one-CTA kernel -- remote accesses cost ~55 cycles per warp instruction, profiles/r02_cfg5_*.)
// All twiddles come from two per-thread base values times compile-time roots of unity (the one-CTA kernel streams 640 KB of
// twiddle tables and ~770 KB of E / power scratch through L2 per frame).
// Reference lines replaced: Spectrogram.cpp:50-56,137-145 (framing, window, spectrum::power), :64-107 (mix, dB), :634-647 +
// CColorpalette.h:32-47 (pixel loops); the log rows are an extension.
#pragma once
#include "jade_pk_cta.cuh"
#include "jade_pkz.cuh"

namespace jade {

// ---------------------------------------------------------------------------------------------------------
// cluster primitives (PTX: %cluster_ctarank, barrier.cluster, mapa, ld / st.shared::cluster); tests/emu runs the two CTAs
// of a cluster concurrently
// ---------------------------------------------------------------------------------------------------------
#if defined(JADE_EMU)
typedef char* peer_addr;
inline unsigned cl_rank() { return jade_emu::cluster_rank(); }
inline void cl_sync() { jade_emu::cluster_barrier(); }
inline peer_addr cl_map(const void* local, unsigned rank) { return jade_emu::cluster_peer_ptr(local, rank); }
inline void cl_st_f2(peer_addr a, f2 v) { std::memcpy(a, &v, 8); }
inline f2 cl_ld_f2(peer_addr a)
{
    f2 v;
    std::memcpy(&v, a, 8);
    return v;
}
inline float cl_ld_f32(peer_addr a)
{
    float v;
    std::memcpy(&v, a, 4);
    return v;
}
#define JADE_CLUSTER_KERNEL(threads) inline void
#else
typedef unsigned peer_addr; // shared::cluster window address
__device__ __forceinline__ unsigned cl_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cl_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ peer_addr cl_map(const void* local, unsigned rank)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(local);
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void cl_st_f2(peer_addr a, f2 v) { asm volatile("st.shared::cluster.b64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ f2 cl_ld_f2(peer_addr a)
{
    f2 v;
    asm volatile("ld.shared::cluster.b64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float cl_ld_f32(peer_addr a)
{
    float v;
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
    return v;
}
#define JADE_CLUSTER_KERNEL(threads) __global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(threads, 1)
#endif

// byte offset inside the peer's window
JADE_DEVICE peer_addr cl_at(peer_addr a, int byte_off) { return a + byte_off; }

struct PkClCfg {
    static constexpr int R1 = 16;
    static constexpr int M2 = 1024 * R1; // complex points per CTA (16384)
    static constexpr int NH = 2 * M2;    // M = N/2 = highest bin (32768)
    static constexpr int N = 2 * NH;     // 65536
    static constexpr int THREADS = 32 * R1;
    static constexpr int RS = PkCtaCfg<R1>::RS;     // 1025
    static constexpr int TROW = PkCtaCfg<R1>::TROW; // 34
    static constexpr int SPEC = M2 + 1;             // power values a CTA holds: bins 2 i + c (CTA 0 also bin N/2 at i = M2)
    static constexpr int off_row = 0;
    static constexpr int off_twI = off_row + R1 * RS * 8;
    static constexpr int off_spec = off_twI + 32 * TROW * 8;
    static constexpr int off_pal = off_spec + (SPEC * 4 + 15) / 16 * 16;
    static JADE_HD int off_rows(int npal) { return off_pal + (npal * 4 + 15) / 16 * 16; }
    // pooled rows: the band table [R] and two buffers of R partial maxima (double-buffered across frames)
    static JADE_HD int off_part(int npal, int pooled_rows) { return off_rows(npal) + (pooled_rows * 8 + 15) / 16 * 16; }
    static JADE_HD int smem_bytes(int npal, int pooled_rows) { return off_part(npal, pooled_rows) + 2 * ((pooled_rows * 4 + 15) / 16 * 16); }
};

// 16384-point complex FFT of the (already windowed) row matrix, in place; cf. cta_fft_pk.  wa = W_16384^t of this thread.
JADE_DEVICE void cl_fft_pk(f2* rowbuf, const f2* s_twI, f2 wa)
{
    using Cfg = PkClCfg;
    constexpr int R1 = Cfg::R1, RS = Cfg::RS;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int c = 0; c < 2; ++c) { // columns n2 = t and t + 512
        const int n2 = t + Cfg::THREADS * c;
        f2 a[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) a[brev(n1, 4)] = rowbuf[n1 * RS + n2];
        fft_pk<R1>(a);
        // twiddles W_M2^(n2 k1): k1 = 1 from the thread's base (x W_32^1 for the second column), the others by squaring
        f2 wk[R1];
        wk[1] = c == 0 ? wa : cmul2(wa, pk(cos32(1), -sin32(1)));
#pragma unroll
        for (int k1 = 2; k1 < R1; ++k1) wk[k1] = (k1 & 1) ? cmul2(wk[k1 - 1], wk[1]) : cmul2(wk[k1 / 2], wk[k1 / 2]);
        rowbuf[n2] = a[0];
#pragma unroll
        for (int k1 = 1; k1 < R1; ++k1) rowbuf[k1 * RS + n2] = cmul2(a[k1], wk[k1]);
    }
    __syncthreads();
    // row pass: warp k1 transforms its 1024-point row in place (as in cta_fft_pk)
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

// position of Y[k'] (k' < M2) inside the row matrix, see rowget
JADE_DEVICE int cl_pos(int k) { return (k & (PkClCfg::R1 - 1)) * PkClCfg::RS + (k / PkClCfg::R1); }

template <int MIXK>
JADE_CLUSTER_KERNEL(PkClCfg::THREADS) stft_pkcl65536_kernel(const KParams P)
{
    using Cfg = PkClCfg;
    constexpr int M2 = Cfg::M2, N = Cfg::N, THREADS = Cfg::THREADS;
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    f2* rowbuf = reinterpret_cast<f2*>(sm + Cfg::off_row);
    f2* s_twI = reinterpret_cast<f2*>(sm + Cfg::off_twI);
    float* s_spec = reinterpret_cast<float*>(sm + Cfg::off_spec);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);
    i2* s_rows = reinterpret_cast<i2*>(sm + Cfg::off_rows(P.npal));
    float* s_part = reinterpret_cast<float*>(sm + Cfg::off_part(P.npal, P.pooled ? P.R : 0));
    const int part_stride = P.pooled ? (P.R * 4 + 15) / 16 * 4 : 0; // floats per partial-maxima buffer

    const int t = threadIdx.x;
    const int c = (int)cl_rank(); // parity of the bins this CTA computes
    stage_row_twiddles<Cfg::R1>(s_twI, P.twI);
    for (int i = t; i < P.npal; i += THREADS) s_pal[i] = P.palette[i];
    if (P.pooled)
        for (int i = t; i < P.R; i += THREADS) s_rows[i] = P.row_bins[i];
    // per-thread twiddle bases (registers, for the whole kernel)
    f2 wa, ws;
    {
        const cpx a = P.twA[1024 + t]; // W_16384^t = W_32768^(2t): column-pass twiddle of column t, and the DIF twiddle of m = 2 t
        const cpx h = P.twP[2 * t + c]; // W_65536^(2 t + c): real-FFT split twiddle of bin k = 2 k' + c at k' = t
        wa = pk(a.x, a.y);
        ws = pk(h.x, h.y);
    }
    const peer_addr peer_part = cl_map(s_part, (unsigned)(c ^ 1));
    __syncthreads();
    grid_dep_wait();
    cl_sync(); // both CTAs of the cluster are resident before the first remote access

    const unsigned total = (unsigned)P.ncols * (unsigned)P.nstreams;
    const unsigned nclusters = gridDim.x / 2, cid = blockIdx.x / 2;
    int ch0, ch1;
    channel_range(P, ch0, ch1);
    if (MIXK == MIX_NONE) ch1 = ch0 + 1;
    const bool mean = P.mix_mode == K_MIX_ABSMEAN && P.channels > 1;
    const float nchf = (float)P.channels;
    unsigned parity = 0; // which partial-maxima buffer this frame uses

    for (unsigned g = cid; g < total; g += nclusters, parity ^= 1u) {
        const int stream = (int)(g / (unsigned)P.ncols);
        const long long j = P.first_col + (g - (unsigned)stream * (unsigned)P.ncols);
        const long long st = frame_start(P, j);
        const long long ns = P.nsamples;
        for (int ch = ch0; ch < ch1; ++ch) {
            const float* x = P.samples + stream * P.stream_stride + ch * P.channel_stride;
            // ---- load + window + DIF stage: complex points m = 2 (t + 512 i) and m + 1 (one 16-byte load of samples and one of
            // window values each, and the same M/2 = 16384 points further on)
            const bool fast = st >= 0 && st + N <= ns && ((reinterpret_cast<uintptr_t>(x + st) | reinterpret_cast<uintptr_t>(P.window)) & 15) == 0;
            auto dif_store = [&](int i, float4 xa, float4 wva, float4 xb, float4 wvb) {
                const int m = 2 * (t + THREADS * i);
                const f2 a0 = mul2(pk(xa.x, xa.y), pk(wva.x, wva.y)), a1 = mul2(pk(xa.z, xa.w), pk(wva.z, wva.w));
                f2 y0, y1;
                if (c == 0) {
                    y0 = fma2(pk(xb.x, xb.y), pk(wvb.x, wvb.y), a0);
                    y1 = fma2(pk(xb.z, xb.w), pk(wvb.z, wvb.w), a1);
                } else {
                    // W_32768^m = W_32768^(2t) W_32^i ; W_32768^(m+1) = that times W_32768^1
                    const f2 w0 = cmul2(wa, pk(cos32(i & 15), -sin32(i & 15)));
                    const f2 w1 = cmul2(w0, pk(0.99999998161642929f, -0.00019174759731070331f));
                    y0 = cmul2(fma2(neg2(pk(xb.x, xb.y)), pk(wvb.x, wvb.y), a0), w0);
                    y1 = cmul2(fma2(neg2(pk(xb.z, xb.w)), pk(wvb.z, wvb.w), a1), w1);
                }
                const int pos = (m >> 10) * Cfg::RS + (m & 1023);
                rowbuf[pos] = y0;
                rowbuf[pos + 1] = y1;
            };
            if (fast) {
                const float4* x4 = reinterpret_cast<const float4*>(x + st);
                const float4* w4 = reinterpret_cast<const float4*>(P.window);
#pragma unroll
                for (int b = 0; b < 4; ++b) { // 16 loads in flight per thread
                    float4 xa[4], xb[4], va[4], vb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int q = t + THREADS * (4 * b + i); // float4 index: samples 4 q .. 4 q + 3 = points 2 q, 2 q + 1
                        xa[i] = x4[q];
                        xb[i] = x4[q + M2 / 2]; // + 32768 samples = M/2 complex points
                        va[i] = w4[q];
                        vb[i] = w4[q + M2 / 2];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) dif_store(4 * b + i, xa[i], va[i], xb[i], vb[i]);
                }
            } else {
                const float4* w4 = reinterpret_cast<const float4*>(P.window);
                for (int i = 0; i < 16; ++i) {
                    const int q = t + THREADS * i;
                    auto ld4 = [&](long long base) { // guarded: clamped address, value selected afterwards
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const long long idx = base + e;
                            const long long cc = idx < 0 ? 0 : (idx < ns ? idx : ns - 1);
                            const float sv = ns > 0 ? x[cc] : 0.f;
                            v[e] = (idx >= 0 && idx < ns) ? sv : 0.f;
                        }
                        float4 r;
                        r.x = v[0];
                        r.y = v[1];
                        r.z = v[2];
                        r.w = v[3];
                        return r;
                    };
                    // (the window table itself is 16-byte aligned device memory whenever this kernel runs)
                    dif_store(i, ld4(st + 4LL * q), w4[q], ld4(st + 4LL * q + 32768), w4[q + M2 / 2]);
                }
            }
            __syncthreads();

            cl_fft_pk(rowbuf, s_twI, wa);

            // ---- real-FFT split + power: bin k = 2 k' + c pairs with M - k, i.e. Y[k'] with Y[(M2 - c - k') mod M2]; both powers
            // go to this CTA's array (index = bin >> 1).  k' = t + 512 q < 8192 (+ the self-paired k' = 8192 on CTA 0).
            auto put = [&](int idx, float p) {
                float a = mix_init<MIXK>(P.mix_mode);
                if (MIXK != MIX_NONE && ch != ch0) a = s_spec[idx];
                mix_add<MIXK>(a, p, P.mix_mode);
                s_spec[idx] = a;
            };
#pragma unroll 4
            for (int q = 0; q < 16; ++q) {
                const int k = t + THREADS * q;
                const int km = (M2 - c - k) & (M2 - 1); // partner inside the transform (k' = 0 on CTA 0: itself)
                const f2 zk = rowbuf[cl_pos(k)], zp = rowbuf[cl_pos(km)];
                // W_65536^(2 k' + c) = W_65536^(2 t + c) W_64^q ; the split multiplies by -i W
                const f2 w = cmul2(ws, pk(cos64(q), -sin64(q)));
                const f2 A = add2(zk, conj2(zp)), Bv = sub2(zk, conj2(zp));
                const f2 T = cmul2(Bv, pk(hi(w), -lo(w)));
                const f2 xp = add2(A, T), xm = sub2(A, T);
                put(k, fm(lo(xp), lo(xp), JADE_FMUL(hi(xp), hi(xp))));                  // bin 2 k' + c
                put(M2 - c - k, fm(lo(xm), lo(xm), JADE_FMUL(hi(xm), hi(xm))));         // bin M - (2 k' + c)  (index M2 for bin N/2)
            }
            if (c == 0 && t == 0) { // k' = 8192: bin N/4 pairs with itself, X = 2 conj Z
                const f2 z = rowbuf[cl_pos(M2 / 2)];
                put(M2 / 2, fm(JADE_FMUL(4.0f, lo(z)), lo(z), JADE_FMUL(JADE_FMUL(4.0f, hi(z)), hi(z))));
            }
            __syncthreads();
        }

        // ---- rows
        const ColOut o = col_out(P, stream, j);
        const int nown = c == 0 ? M2 + 1 : M2; // bins 2 i + c, i < nown
        if (o.db || !P.pooled) { // per-bin outputs: every CTA emits the bins of its parity
            for (int i = t; i < nown; i += THREADS) {
                const int k = 2 * i + c;
                const float p = mean ? JADE_FDIV(s_spec[i], nchf) : s_spec[i];
                const float d = to_db(p, P.db_precise);
                if (o.db) o.db[k] = d;
                if (!P.pooled && o.pix && k >= P.k_lo && k < P.k_hi) o.pix[P.flip ? (P.k_hi - 1 - k) : (k - P.k_lo)] = colour_of(d, P, s_pal);
            }
        }
        if (P.pooled && o.pix) {
            // log max-pool: every CTA reduces the bins of its parity for EVERY row (-1 where it has none: powers are >= 0) ...
            float* part = s_part + parity * part_stride;
            for (int r = t; r < P.R; r += THREADS) {
                const i2 rb = s_rows[r];
                int i = (rb.lo + 1 - c) >> 1;          // first index with 2 i + c >= lo
                const int i1 = (rb.hi - c + 1) >> 1;  // first index with 2 i + c >= hi
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
                float mx = fmaxf(part[r], cl_ld_f32(cl_at(peer_part, (int)(parity * part_stride + r) * 4)));
                if (mean) mx = JADE_FDIV(mx, nchf);
                o.pix[P.flip ? (P.R - 1 - r) : r] = colour_of(to_db(mx, P.db_precise), P, s_pal);
            }
            // (the other partial buffer is used next frame; this one is overwritten the frame after, behind that frame's barrier)
        }
    }
    cl_sync(); // nobody leaves while the peer may still read its shared memory
}

} // namespace jade
