// jade_kernels.cuh -- fused STFT -> power -> channel mix -> dB -> row map -> palette kernels for sm_100a.
//
// One launch covers the whole hot path of the reference (file:line relative to /root/reference):
//   framing            Spectrogram.cpp:50-56      (frame start computed per column, no copy pass)
//   window multiply    Spectrogram.cpp:137-141    (fused into the load)
//   real FFT -> power  Spectrogram.cpp:144        (spectrum::power, external; written from scratch here)
//   channel mix        Spectrogram.cpp:64-106
//   10*log10(p+1e-11)  Spectrogram.cpp:36,107
//   bin -> row (flip)  Spectrogram.cpp:642        (+ linear crop :441-459, + log max-pool extension)
//   getRGBColor | 0xFF000000   CColorpalette.h:32-47, Spectrogram.cpp:636-637
//
// Kernel families
//   stft_warp_kernel<T>   N = 64*T <= 2048   : one FFT per T lanes, 32 complex values per thread in registers,
//                                              two radix passes (32 x T) with one shared-memory transpose per warp.
//   stft_cta_kernel<R1>   N = 2048*R1 <= 32768: CTA of R1 warps; radix-R1 column pass, then one 1024-point row FFT
//                                              per warp (same register code), rows staged in shared memory.
//   The fast paths of N >= 2048 live in jade_pk.cuh / jade_pk_cta.cuh (packed FP32x2 arithmetic); N = 65536 is built
//   from two half-size real FFTs there (stft_pkcta2_kernel).
//
// The real-input FFT packs z[m] = x[2m] + i x[2m+1] (M = N/2 complex points) and finishes with the split
//   X[k] = (Z[k] + conj Z[M-k]) - i W_N^k (Z[k] - conj Z[M-k])      (window table is pre-multiplied by 1/2)
//
// All code here also compiles for the host SIMT emulator in tests/emu (JADE_EMU) -- test infrastructure only.
#pragma once
#include <stdint.h>

#include "jade_fft_regs.cuh"

#if defined(JADE_EMU)
#include "cuda_emu.h"
#define JADE_KERNEL(...) inline void
#define JADE_DYN_SMEM(name) float4* name = reinterpret_cast<float4*>(jade_emu::dyn_smem())
#define JADE_RESTRICT
#define JADE_LOG2F(x) ::log2f(x)
#define JADE_FDIV(a, b) ((a) / (b))
#define JADE_FMUL(a, b) ((a) * (b))
#define JADE_FADD(a, b) ((a) + (b))
#else
#define JADE_KERNEL(...) __global__ void __launch_bounds__(__VA_ARGS__)
#define JADE_DYN_SMEM(name) extern __shared__ float4 name[]
#define JADE_RESTRICT __restrict__
// raw MUFU.LG2: the argument is p + 1e-11 >= 1e-11, never denormal, so __log2f's denormal pre-scaling is dead weight
__device__ __forceinline__ float jade_lg2(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
#define JADE_LOG2F(x) jade_lg2(x)
#define JADE_FDIV(a, b) __fdiv_rn((a), (b))
// explicitly rounded (never contracted into an FFMA): every kernel instantiation must produce the same bits
#define JADE_FMUL(a, b) __fmul_rn((a), (b))
#define JADE_FADD(a, b) __fadd_rn((a), (b))
#endif

namespace jade {

enum { K_MIX_ABSMEAN = 0, K_MIX_MAX = 1, K_MIX_MIN = 2, K_MIX_LEFT = 3, K_MIX_RIGHT = 4 };

struct i2 {
    int lo, hi;
};

struct KParams {
    // ---- input samples (device), planar: samples[stream*stream_stride + channel*channel_stride + i]
    const float* samples;
    long long stream_stride;
    long long channel_stride;
    long long nsamples;     // valid samples per channel; reads outside [0,nsamples) give 0
    long long sample_base;  // absolute sample index of samples[..][0]
    int aligned2;           // every frame start is even and the channel bases are 8-byte aligned
    int aligned4;           // every frame start is a multiple of 4 samples and the channel bases are 16-byte aligned
    // ---- geometry: column j starts at (j / fb) * bstride + (j % fb) * hop - preroll   (absolute sample index)
    int N, M, B;
    int hop, fb, bstride, preroll;
    long long first_col;
    int ncols;              // columns per stream in this launch
    int nstreams;
    int channels;
    int mix_mode;
    // ---- tables (device)
    const float* window;    // N floats, already multiplied by 0.5*sqrt(power_scale)
    const cpx* twI;         // [32][T]   W_{32T}^{k1*s}       (warp FFT inter-pass twiddles; T=32 for row FFTs)
    const cpx* twP;         // [M+1]     W_N^k                (real-FFT split)
    const cpx* twA;         // [R1][1024] W_M^{k1*n2}         (CTA kernels: column-pass twiddles)
    const uint32_t* palette; // npal entries, alpha / byte order already baked in
    int npal;
    float pmin, pmax, pmaxc, pmult; // CColorPalette m_Min, m_Max, m_Max*0.9999f, m_AccessMult
    // the same lookup folded onto lg = log2(p + 1e-11) (colour_of_lg, N = 2048 kernel): index = trunc(lg*ck1 + ck0),
    // lg >= clg_hi (dB >= m_Max) gives ci_hi = index of m_Max*0.9999
    float ck1, ck0, clg_hi;
    int ci_hi;
    int pal_u8;             // 256 colours whose `>= m_Max` entry equals entry 255: the saturating float -> u8 conversion is the whole clamp (colour_of_lg1)
    int db_precise;         // 1: float(10.0*log10(double(p+1e-11f))) exactly as the reference; 0: MUFU log2
    // ---- rows
    int pooled;             // 0: rows are bins [k_lo,k_hi); 1: row r = max over bins [row_bins[r].lo, row_bins[r].hi)
    int R;                  // rows per column
    int k_lo, k_hi;
    int flip;               // 1: row 0 is the highest frequency (reference orientation, Spectrogram.cpp:642)
    const i2* row_bins;     // [R] (pooled only)
    // ---- outputs (device)
    uint32_t* pix;          // [stream][col][R]   (may be null)
    float* db;              // [stream][col][B]   (may be null)
    long long pix_stream_stride; // in elements
    long long db_stream_stride;
    int ring_w;             // >0: column slot = (ring_col0 + j - first_col) % ring_w (streaming ring); 0: j - first_col
    long long ring_col0;    // ring column counter of the first column of this launch
    // ---- scratch for the two-half kernels (N = 65536)
    cpx* scratch_e;         // [grid][M/2+1] complex
    float* scratch_p;       // [grid][B] floats
    const cpx* twH;         // [M/2+1]  W_{N/2}^k  (split of the half-size real FFTs)
};

// Programmatic dependent launch (streaming path): the engine launches the STFT kernel of a push with
// cudaLaunchAttributeProgrammaticStreamSerialization right behind ingest_kernel, which releases its dependents at once, so
// that the STFT kernel's table prologue overlaps the ingest; grid_dep_wait() -- after the prologue, before the first read
// of the samples -- blocks until the ingest grid has completed and its writes are visible.  Both are no-ops for a
// kernel launched the ordinary way.
#if defined(JADE_EMU)
inline void grid_dep_wait() {}
inline void grid_dep_launch() {}
#else
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// 10*log10(p + 1e-11) (Spectrogram.cpp:36,107).  Fast: one MUFU.LG2 + one FMUL (|err| ~ 1e-5 dB).
JADE_DEVICE float to_db_fast(float p) { return JADE_FMUL(3.01029995663981195f, JADE_LOG2F(JADE_FADD(p, 1e-11f))); }
// Exactly the reference's arithmetic: float add, double log10, double multiply, float store.
#if defined(__CUDACC__)
static __device__ __noinline__
#else
inline
#endif
float to_db_precise(float p)
{
    const float sh = p + 1e-11f;
    return (float)(10.0 * log10((double)sh));
}
JADE_DEVICE float to_db(float p, int precise) { return precise ? to_db_precise(p) : to_db_fast(p); }

// CColorPalette::getRGBColor (CColorpalette.h:32-47) on the baked table.  The index is additionally clamped at 0
// (the reference would read out of bounds when m_Min > m_Max).
JADE_DEVICE uint32_t colour_of(float v, const KParams& P, const uint32_t* pal)
{
    v = (v >= P.pmax) ? P.pmaxc : v;
    v = fmaxf(v, P.pmin);            // == `if (v < m_Min) v = m_Min` for every non-NaN v
    const float d = JADE_FADD(v, -P.pmin);
    int idx = (int)JADE_FMUL(d, P.pmult);
    idx = max(min(idx, P.npal - 1), 0);
    return pal[idx];
}

// The same lookup from lg = log2(p + 1e-11) with the dB scale folded in: one FFMA instead of FMUL, FADD, FMUL, and the lower
// clamp left to the integer RELU.  The index can differ from colour_of(3.0103 lg) only where the dB value lies within
// ~1e-5 dB of an index edge (the parity tolerance is 1e-3 dB, tests/parity.py).
JADE_DEVICE uint32_t colour_of_lg(float lg, const KParams& P, const uint32_t* pal)
{
    int idx = (int)fm(lg, P.ck1, P.ck0);
    idx = max(min(idx, P.npal - 1), 0);
    idx = (lg >= P.clg_hi) ? P.ci_hi : idx;
#if defined(JADE_ABL_NOPAL)
    return (uint32_t)idx;
#else
    return pal[idx];
#endif
}
// The lookup of the packed N <= 2048 / N = 16384 kernels: their shared-memory table has npal + 1 entries, entry npal holding
// the colour of CColorPalette's `value >= m_Max` rule (index ci_hi of 0.9999 m_Max, CColorpalette.h:34-35) --
// lg ck1 + ck0 >= npal exactly when the dB value reaches m_Max -- so ONE integer clamp to [0, npal] does it (VIMNMX.RELU).
// U8 (P.pal_u8: npal == 256 and entry 256 == entry 255, true for every 256-colour scheme with the reference's ranges): the
// conversion itself saturates to [0, 255] (F2IP.U8.F32), no clamp instruction at all; same colours by construction.
template <bool U8>
JADE_DEVICE uint32_t colour_of_lg1(float lg, const KParams& P, const uint32_t* pal)
{
    const float x = fm(lg, P.ck1, P.ck0);
    if (U8) {
#if defined(JADE_EMU)
        const int i = (int)x;
        return pal[i < 0 ? 0 : (i > 255 ? 255 : i)];
#else
        unsigned idx;
        asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(idx) : "f"(x));
        return pal[idx];
#endif
    }
    return pal[max(min((int)x, P.npal), 0)];
}
// host side: fill ck1, ck0, clg_hi, ci_hi from pmin, pmax, pmaxc, pmult, npal
inline void colour_fold(KParams& P)
{
    P.ck1 = 3.01029995663981195f * P.pmult;
    P.ck0 = -P.pmin * P.pmult;
    P.clg_hi = P.pmax / 3.01029995663981195f;
    int ih = (int)((P.pmaxc > P.pmin ? P.pmaxc - P.pmin : 0.0f) * P.pmult);
    P.ci_hi = ih < 0 ? 0 : (ih > P.npal - 1 ? P.npal - 1 : ih);
}

// emit_bin with the one-clamp lookup (table of npal + 1 entries); EPS_IN: p already carries the + 1e-11 (seeded into the
// power FMAs of a one-channel kernel)
template <int MIXK, bool WANT_DB, bool U8, bool EPS_IN>
JADE_DEVICE void emit_bin1(float p, float scale, uint32_t* pix, float* db, const KParams& P, const uint32_t* pal)
{
    const float lg = JADE_LOG2F(EPS_IN ? p : (MIXK == 1 /* MIX_SUM */ ? fm(p, scale, 1e-11f) : JADE_FADD(p, 1e-11f)));
    if (WANT_DB) {
        if (db) *db = JADE_FMUL(3.01029995663981195f, lg);
        if (pix) *pix = colour_of_lg1<U8>(lg, P, pal);
    } else {
        *pix = colour_of_lg1<U8>(lg, P, pal);
    }
}

// Fast-path epilogue of one bin (packed kernels): mixed power -> lg = log2(p' + 1e-11) -> dB value (optional) and pixel.
// MIX_SUM is only routed here for 2^n channels, so the single FFMA rounds exactly like (p * scale) + 1e-11; dB = 3.0103 lg
// is to_db_fast.  pix / db may be null when WANT_DB (the plain instantiations are only launched with a pixel buffer).
template <int MIXK, bool WANT_DB>
JADE_DEVICE void emit_bin(float p, float scale, uint32_t* pix, float* db, const KParams& P, const uint32_t* pal)
{
#if defined(JADE_ABL_NOEPI)
    if (!WANT_DB) {
        *pix = (uint32_t)(int)fm(p, scale, 1e-11f);
        return;
    }
#endif
    const float lg = JADE_LOG2F(MIXK == 1 /* MIX_SUM */ ? fm(p, scale, 1e-11f) : JADE_FADD(p, 1e-11f));
    if (WANT_DB) {
        if (db) *db = JADE_FMUL(3.01029995663981195f, lg);
        if (pix) *pix = colour_of_lg(lg, P, pal);
    } else {
        *pix = colour_of_lg(lg, P, pal);
    }
}

JADE_DEVICE cpx load_pair_guarded(const float* JADE_RESTRICT x, long long idx, long long ns)
{
    cpx r;
    r.x = (idx >= 0 && idx < ns) ? x[idx] : 0.f;
    r.y = (idx + 1 >= 0 && idx + 1 < ns) ? x[idx + 1] : 0.f;
    return r;
}

// |X'[k]|^2 from Z[k], Z[M-k] and W_N^k (see header comment)
JADE_DEVICE float split_power(cpx zk, cpx zp, cpx w)
{
    const float ax = zk.x + zp.x, ay = zk.y - zp.y;
    const float bx = zk.x - zp.x, by = zk.y + zp.y;
    const float xr = ax + fm(w.x, by, w.y * bx);
    const float xi = ay - fm(w.x, bx, -(w.y * by));
    return fm(xr, xr, xi * xi);
}
JADE_DEVICE cpx split_value(cpx zk, cpx zp, cpx w)
{
    const float ax = zk.x + zp.x, ay = zk.y - zp.y;
    const float bx = zk.x - zp.x, by = zk.y + zp.y;
    return mk(ax + fm(w.x, by, w.y * bx), ay - fm(w.x, bx, -(w.y * by)));
}

// Channel mix (Spectrogram.cpp:64-106) as a compile-time kind so the per-bin code stays branch-free:
//   MIX_NONE : one contributing channel (mono, Left, Right)          acc = p
//   MIX_SUM  : AbsMean                                               acc += p, finally / channels
//   MIX_SEL  : Max (start 0) / Min (start 1e6)                       acc = max/min(acc, p)
enum { MIX_NONE = 0, MIX_SUM = 1, MIX_SEL = 2 };
template <int MIXK>
JADE_DEVICE float mix_init(int mode)
{
    return (MIXK == MIX_SEL && mode == K_MIX_MIN) ? 1000000.0f : 0.0f;
}
template <int MIXK>
JADE_DEVICE void mix_add(float& a, float p, int mode)
{
    if (MIXK == MIX_NONE) a = p;
    else if (MIXK == MIX_SUM) a += p;
    else a = (mode == K_MIX_MIN) ? ((p < a) ? p : a) : ((p > a) ? p : a);
}

struct ColOut {
    uint32_t* pix; // column base or null
    float* db;     // column base or null
};
JADE_DEVICE ColOut col_out(const KParams& P, int stream, long long j)
{
    const long long slot = P.ring_w > 0 ? ((P.ring_col0 + (j - P.first_col)) % P.ring_w) : (j - P.first_col);
    ColOut o;
    o.pix = P.pix ? P.pix + stream * P.pix_stream_stride + slot * P.R : nullptr;
    o.db = P.db ? P.db + stream * P.db_stream_stride + slot * P.B : nullptr;
    return o;
}
JADE_DEVICE long long frame_start(const KParams& P, long long j)
{
    if (P.fb == 1) return j * (long long)P.bstride - P.preroll - P.sample_base;
    return (j / P.fb) * (long long)P.bstride + (j % P.fb) * (long long)P.hop - P.preroll - P.sample_base;
}
// channels that contribute to the mix (Left / Right need a single one)
JADE_DEVICE void channel_range(const KParams& P, int& ch0, int& ch1)
{
    ch0 = 0;
    ch1 = P.channels;
    if (P.mix_mode == K_MIX_LEFT) ch1 = 1;
    if (P.mix_mode == K_MIX_RIGHT) {
        ch0 = P.channels > 1 ? 1 : 0;
        ch1 = ch0 + 1;
    }
}

// General ("slow") epilogue from a mixed power spectrum in shared or global memory: any row map, precise dB,
// non power-of-two channel means.  Rolled loops: small code, only used off the headline path.
// Phase 1 (per bin) must be followed by a barrier of the participating threads before phase 2 (pooled rows).
JADE_DEVICE void emit_general_bins(const KParams& P, const uint32_t* pal, const ColOut& o, float* spec, int tid, int step)
{
    const bool mean = P.mix_mode == K_MIX_ABSMEAN && P.channels > 1;
    const float nchf = (float)P.channels;
    for (int k = tid; k < P.B; k += step) {
        float p = spec[k];
        if (mean) p = JADE_FDIV(p, nchf);
        spec[k] = p;
        if (o.db || !P.pooled) {
            const float d = to_db(p, P.db_precise);
            if (o.db) o.db[k] = d;
            if (!P.pooled && o.pix && k >= P.k_lo && k < P.k_hi)
                o.pix[P.flip ? (P.k_hi - 1 - k) : (k - P.k_lo)] = colour_of(d, P, pal);
        }
    }
}
JADE_DEVICE void emit_general_rows(const KParams& P, const uint32_t* pal, const ColOut& o, const float* spec, int tid, int step)
{
    if (!P.pooled || !o.pix) return;
    for (int r = tid; r < P.R; r += step) {
        const i2 rb = P.row_bins[r];
        float mx = spec[rb.lo];
        for (int k = rb.lo + 1; k < rb.hi; ++k) mx = fmaxf(mx, spec[k]);
        o.pix[P.flip ? (P.R - 1 - r) : r] = colour_of(to_db(mx, P.db_precise), P, pal);
    }
}

// ---------------------------------------------------------------------------------------------------------
// Warp-level M = 32*T point complex FFT, T lanes per transform (F = 32/T transforms per warp).
// in : v[brev5(n1)] = z[s + T*n1]           (s = lane % T)
// out: u[i*T + k2]  = Z[(s + T*i) + 32*k2]  (i < 32/T, k2 < T)  i.e. bin k = s + T*q sits in u[(q % F)*T + q / F]
// xw : this transform's shared-memory scratch; synchronised with __syncwarp only.
// PADDED = true : transpose through rows of T+1 words (needs 32*(T+1) words).  Every access is base+immediate and
//                 conflict-free for 8-byte words.
// PADDED = false: in-place XOR swizzle inside exactly 32*T words (used on the CTA kernels' row buffers).
// ---------------------------------------------------------------------------------------------------------
template <int T, bool PADDED>
JADE_DEVICE void warp_fft(cpx* v, cpx* u, cpx* xw, const cpx* twI, int s)
{
    constexpr int F = 32 / T;
    fft_dit<32>(v);
    const cpx* tw = twI + s;
    if (PADDED) {
        cpx* wr = xw + s;
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) wr[k1 * (T + 1)] = (k1 == 0) ? v[0] : cmul(v[k1], tw[k1 * T]);
        __syncwarp();
        const cpx* rd = xw + s * (T + 1);
#pragma unroll
        for (int i = 0; i < F; ++i) {
#pragma unroll
            for (int jx = 0; jx < T; ++jx) u[i * T + brev(jx, ilog2c(T))] = rd[T * i * (T + 1) + jx];
            fft_dit<T>(u + i * T);
        }
    } else {
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1)
            xw[k1 * T + (s ^ (k1 & (T - 1)))] = (k1 == 0) ? v[0] : cmul(v[k1], tw[k1 * T]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < F; ++i) {
            const int k1 = s + T * i;
#pragma unroll
            for (int jx = 0; jx < T; ++jx) u[i * T + brev(jx, ilog2c(T))] = xw[k1 * T + (jx ^ s)];
            fft_dit<T>(u + i * T);
        }
    }
    __syncwarp();
}

// =========================================================================================================
// Class A: N = 64*T
// =========================================================================================================
constexpr int WARP_KERNEL_WARPS = 8;

template <int T>
struct WarpCfg {
    static constexpr int M = 32 * T;
    static constexpr int N = 2 * M;
    static constexpr int B = M + 1;
    static constexpr int F = 32 / T;
    // per-transform scratch (complex words): 32 rows of T+1 (transpose) / M+1 natural-order words; for T < 16 the
    // stride is made == T (mod 16) so that the transforms sharing a half-warp hit disjoint banks
    static constexpr int FS = 32 * T + 32 + (T < 16 ? T : 0);
    static constexpr int SPEC_STRIDE = ((B + 3) / 4) * 4; // floats
    // shared memory layout in bytes
    static constexpr int off_twI = 0;
    static constexpr int off_win = off_twI + 32 * T * 8;
    static constexpr int off_twP = off_win + M * 8;
    static constexpr int off_pal = off_twP + ((M + 1) * 8 + 15) / 16 * 16;
    static JADE_HD int off_xch(int npal) { return off_pal + (npal * 4 + 15) / 16 * 16; }
    static JADE_HD int smem_bytes(int npal, bool general)
    {
        return off_xch(npal) + WARP_KERNEL_WARPS * F * FS * 8 + (general ? WARP_KERNEL_WARPS * F * SPEC_STRIDE * 4 : 0);
    }
};

// GENERAL = false: headline path (identity rows in the reference orientation, hardware log2, exact reciprocal for the
//                  mean); all addressing is base + immediate.  GENERAL = true: every other option (emit_general_*).
template <int T, int MIXK, bool GENERAL>
JADE_KERNEL(WARP_KERNEL_WARPS * 32, (T >= 4 && !GENERAL) ? 2 : 1) stft_warp_kernel(const KParams P)
{
    using Cfg = WarpCfg<T>;
    constexpr int M = Cfg::M, N = Cfg::N, F = Cfg::F, FS = Cfg::FS;
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    cpx* s_twI = reinterpret_cast<cpx*>(sm + Cfg::off_twI);
    cpx* s_win = reinterpret_cast<cpx*>(sm + Cfg::off_win);
    cpx* s_twP = reinterpret_cast<cpx*>(sm + Cfg::off_twP);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);
    cpx* s_xch = reinterpret_cast<cpx*>(sm + Cfg::off_xch(P.npal));
    float* s_spec = reinterpret_cast<float*>(s_xch + WARP_KERNEL_WARPS * F * FS);

    for (int i = threadIdx.x; i < 32 * T; i += blockDim.x) s_twI[i] = P.twI[i];
    for (int i = threadIdx.x; i < M; i += blockDim.x) s_win[i] = mk(P.window[2 * i], P.window[2 * i + 1]);
    for (int i = threadIdx.x; i <= M; i += blockDim.x) s_twP[i] = P.twP[i];
    for (int i = threadIdx.x; i < P.npal; i += blockDim.x) s_pal[i] = P.palette[i];
    __syncthreads();
    grid_dep_wait();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int f = lane / T, s = lane % T;
    cpx* xw = s_xch + (warp * F + f) * FS;
    float* spec = s_spec + (warp * F + f) * Cfg::SPEC_STRIDE;
    const cpx* win_l = s_win + s;
    const cpx* twP_l = s_twP + s;
    cpx* nat_wr = xw + s;            // natural-order copy of Z: word k = s + T*q
    const cpx* nat_rd = xw + (T - s); // partner Z[M-k] = word (T - s) + T*(31 - q); word M duplicates Z[0]

    const unsigned groups = (unsigned)((P.ncols + F - 1) / F);
    const unsigned total = groups * (unsigned)P.nstreams;
    int ch0, ch1;
    channel_range(P, ch0, ch1);

    for (unsigned g = blockIdx.x * WARP_KERNEL_WARPS + warp; g < total; g += gridDim.x * WARP_KERNEL_WARPS) {
        const int stream = (int)(g / groups);
        const int jfirst = (int)(g - (unsigned)stream * groups) * F; // first column of this warp's group
        int jrel = jfirst + f;
        const bool active = jrel < P.ncols;
        if (!active) jrel = P.ncols - 1;
        const long long j = P.first_col + jrel;
        const long long st = frame_start(P, j);
        // warp-uniform: every transform of this warp lies inside the signal (frame starts are monotone in j)
        bool fast;
        if (F == 1) {
            fast = P.aligned2 && st >= 0 && st + N <= P.nsamples;
        } else {
            const int jlast = (jfirst + F - 1 < P.ncols) ? jfirst + F - 1 : P.ncols - 1;
            fast = P.aligned2 && frame_start(P, P.first_col + jfirst) >= 0 &&
                   frame_start(P, P.first_col + jlast) + N <= P.nsamples;
        }

        float acc[33];
#pragma unroll
        for (int q = 0; q < 33; ++q) acc[q] = mix_init<MIXK>(P.mix_mode);

        for (int ch = ch0; ch < (MIXK == MIX_NONE ? ch0 + 1 : ch1); ++ch) {
            const float* x = P.samples + stream * P.stream_stride + ch * P.channel_stride;
            cpx v[32], u[32];
            if (fast) {
                const cpx* xz = reinterpret_cast<const cpx*>(x + st) + s;
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) {
                    const cpx z = xz[T * n1];
                    const cpx w = win_l[T * n1];
                    v[brev(n1, 5)] = mk(z.x * w.x, z.y * w.y);
                }
            } else { // frame touches the signal boundary or is unaligned: guarded loads staged through shared memory
                for (int m = s; m < M; m += T) {
                    const cpx z = load_pair_guarded(x, st + 2 * m, P.nsamples);
                    const cpx w = s_win[m];
                    xw[m] = mk(z.x * w.x, z.y * w.y);
                }
#pragma unroll
                for (int n1 = 0; n1 < 32; ++n1) v[brev(n1, 5)] = nat_wr[T * n1]; // the thread's own words
                __syncwarp();
            }
            warp_fft<T, true>(v, u, xw, s_twI, s);
#pragma unroll
            for (int q = 0; q < 32; ++q) nat_wr[T * q] = u[(q % F) * T + q / F];
            if (s == 0) xw[M] = u[0];
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const float p = split_power(u[(q % F) * T + q / F], nat_rd[T * (31 - q)], twP_l[T * q]);
                mix_add<MIXK>(acc[q], p, P.mix_mode);
            }
            {   // Nyquist bin k = M (kept by sub-lane 0; computed by every lane to stay convergent)
                const cpx z0 = xw[0];
                mix_add<MIXK>(acc[32], split_power(z0, z0, s_twP[M]), P.mix_mode);
            }
            __syncwarp();
        }
        const ColOut o = active ? col_out(P, stream, j) : ColOut{nullptr, nullptr};
        if (!GENERAL) {
            // identity rows: bin k = s + T*q -> row (flip ? M - k : k); mean over 2^n channels is an exact multiply
            const float scale = (MIXK == MIX_SUM) ? (1.0f / (float)P.channels) : 1.0f;
            // reference orientation (flip): bin k = s + T*q lands in row M - k; every store is base + immediate
            uint32_t* prow = o.pix ? o.pix + (M - s) : nullptr;
            float* drow = o.db ? o.db + s : nullptr;
#pragma unroll
            for (int q = 0; q < 33; ++q) {
                if (q == 32 && s != 0) break;
                const float d = to_db_fast(MIXK == MIX_SUM ? JADE_FMUL(acc[q], scale) : acc[q]);
                const uint32_t c = colour_of(d, P, s_pal);
                if (drow) drow[T * q] = d;
                if (prow) prow[-T * q] = c;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) spec[s + T * q] = acc[q];
            if (s == 0) spec[M] = acc[32];
            __syncwarp();
            emit_general_bins(P, s_pal, o, spec, s, T);
            __syncwarp();
            emit_general_rows(P, s_pal, o, spec, s, T);
            __syncwarp();
        }
    }
}

// =========================================================================================================
// Class B: N = 2048*R1, one frame per CTA iteration, CTA = R1 warps
// =========================================================================================================
template <int R1>
struct CtaCfg {
    static constexpr int M = 1024 * R1;
    static constexpr int N = 2 * M;
    static constexpr int B = M + 1;
    static constexpr int THREADS = 32 * R1;
    static constexpr int RS = 1024 + 16 / R1; // row stride (complex words)
    static constexpr int off_row = 0;
    static constexpr int off_twI = off_row + R1 * RS * 8;
    static constexpr int off_pal = off_twI + 1024 * 8;
    static JADE_HD int off_spec(int npal) { return off_pal + (npal * 4 + 15) / 16 * 16; }
    static JADE_HD int smem_bytes(int npal, bool general) { return off_spec(npal) + (general ? ((B + 3) / 4) * 16 : 0); }
};

// Z[k] of the M-point transform inside the row buffer (k1 = k % R1 is the row, k / R1 the row-FFT bin)
template <int R1>
JADE_DEVICE cpx rowbuf_get(const cpx* rowbuf, int k)
{
    return rowbuf[(k & (R1 - 1)) * CtaCfg<R1>::RS + (k / R1)];
}

// M = 1024*R1 point complex FFT of z[m] = ld(m) by the whole CTA; result left in rowbuf (see rowbuf_get).
template <int R1, typename Loader>
JADE_DEVICE void cta_fft(cpx* rowbuf, const cpx* s_twI, const cpx* JADE_RESTRICT twA, Loader ld)
{
    using Cfg = CtaCfg<R1>;
    constexpr int RS = Cfg::RS, CPT = 32 / R1;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    // column pass: radix R1 over n1, thread owns columns n2 = t + THREADS*c
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
        const int n2 = t + Cfg::THREADS * c;
        cpx a[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) a[brev(n1, ilog2c(R1))] = ld(n2 + 1024 * n1);
        fft_dit<R1>(a);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) rowbuf[k1 * RS + n2] = (k1 == 0) ? a[0] : cmul(a[k1], twA[k1 * 1024 + n2]);
    }
    __syncthreads();
    // row pass: warp `warp` transforms row k1 = warp (1024 points) in place
    cpx* row = rowbuf + warp * RS;
    cpx v[32], u[32];
#pragma unroll
    for (int n1 = 0; n1 < 32; ++n1) v[brev(n1, 5)] = row[lane + 32 * n1];
    __syncwarp();
    warp_fft<32, false>(v, u, row, s_twI, lane);
#pragma unroll
    for (int k2 = 0; k2 < 32; ++k2) row[lane + 32 * k2] = u[k2];
    __syncthreads();
}

template <int R1, int MIXK, bool GENERAL>
JADE_KERNEL(32 * R1, 1) stft_cta_kernel(const KParams P)
{
    using Cfg = CtaCfg<R1>;
    constexpr int M = Cfg::M, N = Cfg::N, THREADS = Cfg::THREADS;
    JADE_DYN_SMEM(smem);
    char* sm = reinterpret_cast<char*>(smem);
    cpx* rowbuf = reinterpret_cast<cpx*>(sm + Cfg::off_row);
    cpx* s_twI = reinterpret_cast<cpx*>(sm + Cfg::off_twI);
    uint32_t* s_pal = reinterpret_cast<uint32_t*>(sm + Cfg::off_pal);
    float* spec = reinterpret_cast<float*>(sm + Cfg::off_spec(P.npal));

    const int t = threadIdx.x;
    for (int i = t; i < 1024; i += THREADS) s_twI[i] = P.twI[i];
    for (int i = t; i < P.npal; i += THREADS) s_pal[i] = P.palette[i];
    __syncthreads();
    grid_dep_wait();

    const unsigned total = (unsigned)P.ncols * (unsigned)P.nstreams;
    int ch0, ch1;
    channel_range(P, ch0, ch1);
    const cpx* JADE_RESTRICT winp = reinterpret_cast<const cpx*>(P.window);

    for (unsigned g = blockIdx.x; g < total; g += gridDim.x) {
        const int stream = (int)(g / (unsigned)P.ncols);
        const long long j = P.first_col + (g - (unsigned)stream * (unsigned)P.ncols);
        const long long st = frame_start(P, j);
        const bool fast = P.aligned2 && st >= 0 && st + N <= P.nsamples;

        float acc[33];
#pragma unroll
        for (int q = 0; q < 33; ++q) acc[q] = mix_init<MIXK>(P.mix_mode);

        for (int ch = ch0; ch < (MIXK == MIX_NONE ? ch0 + 1 : ch1); ++ch) {
            const float* x = P.samples + stream * P.stream_stride + ch * P.channel_stride;
            const long long ns = P.nsamples;
            if (fast) {
                const cpx* xz = reinterpret_cast<const cpx*>(x + st);
                cta_fft<R1>(rowbuf, s_twI, P.twA, [&](int m) {
                    const cpx z = xz[m];
                    const cpx w = winp[m];
                    return mk(z.x * w.x, z.y * w.y);
                });
            } else {
                cta_fft<R1>(rowbuf, s_twI, P.twA, [&](int m) {
                    const cpx z = load_pair_guarded(x, st + 2 * m, ns);
                    const cpx w = winp[m];
                    return mk(z.x * w.x, z.y * w.y);
                });
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                const int k = t + THREADS * q;
                const cpx zk = rowbuf_get<R1>(rowbuf, k);
                const cpx zp = rowbuf_get<R1>(rowbuf, (M - k) & (M - 1));
                mix_add<MIXK>(acc[q], split_power(zk, zp, P.twP[k]), P.mix_mode);
            }
            {
                const cpx z0 = rowbuf[0];
                mix_add<MIXK>(acc[32], split_power(z0, z0, P.twP[M]), P.mix_mode);
            }
            __syncthreads();
        }
        const ColOut o = col_out(P, stream, j);
        if (!GENERAL) {
            const float scale = (MIXK == MIX_SUM) ? (1.0f / (float)P.channels) : 1.0f;
#pragma unroll
            for (int q = 0; q < 33; ++q) {
                if (q == 32 && t != 0) break;
                const int k = t + THREADS * q;
                const float d = to_db_fast(MIXK == MIX_SUM ? JADE_FMUL(acc[q], scale) : acc[q]);
                const uint32_t c = colour_of(d, P, s_pal);
                if (o.db) o.db[k] = d;
                if (o.pix) o.pix[M - k] = c;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) spec[t + THREADS * q] = acc[q];
            if (t == 0) spec[M] = acc[32];
            __syncthreads();
            emit_general_bins(P, s_pal, o, spec, t, THREADS);
            __syncthreads();
            emit_general_rows(P, s_pal, o, spec, t, THREADS);
            __syncthreads();
        }
    }
}

// =========================================================================================================
// Small helper kernels (non-template: compiled only into the translation unit that defines JADE_HELPER_KERNELS)
// =========================================================================================================
#if defined(JADE_HELPER_KERNELS) || defined(JADE_EMU)
// Re-colour stored dB columns (ring or batch) -- SpectrogramComponent's m_recomputeAll path (Spectrogram.cpp:623-657)
JADE_KERNEL(256) recolor_kernel(const KParams P, const float* dbcols, long long ncolumns)
{
    for (long long c = blockIdx.x; c < ncolumns; c += gridDim.x) {
        const float* d = dbcols + c * P.B;
        uint32_t* o = P.pix + c * P.R;
        if (P.pooled) { // log max-pool rows: 10 log10 is monotone, so the band's largest dB value is the dB of its largest power
            for (int r = threadIdx.x; r < P.R; r += blockDim.x) {
                const i2 rb = P.row_bins[r];
                float mx = d[rb.lo];
                for (int k = rb.lo + 1; k < rb.hi; ++k) mx = fmaxf(mx, d[k]);
                o[P.flip ? (P.R - 1 - r) : r] = colour_of(mx, P, P.palette);
            }
            continue;
        }
        for (int k = P.k_lo + threadIdx.x; k < P.k_hi; k += blockDim.x) {
            const int row = P.flip ? (P.k_hi - 1 - k) : (k - P.k_lo);
            o[row] = colour_of(d[k], P, P.palette);
        }
    }
}

// Append `n` staged samples per channel to the device history (streaming path).  If `slide` > 0 the last `keep`
// samples are first moved to the front (source and destination ranges never overlap, see engine).
JADE_KERNEL(256) ingest_kernel(float* hist, long long channel_stride, int channels, const float* stage, int n,
                               long long write_pos, long long slide_from, int keep)
{
    grid_dep_launch(); // the STFT kernel behind may start its prologue now; it waits for this grid before reading hist
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, step = gridDim.x * blockDim.x;
    for (int ch = 0; ch < channels; ++ch) {
        float* h = hist + ch * channel_stride;
        if (keep > 0)
            for (int i = tid; i < keep; i += step) h[i] = h[slide_from + i];
        const float* s = stage + (long long)ch * n;
        for (int i = tid; i < n; i += step) h[write_pos + i] = s[i];
    }
}

// Deterministic synthetic signals: kind 0 = linear sine sweep 20 Hz -> 0.475 fs, 1 = white noise
// uniform(-0.5,0.5) from splitmix64(seed, stream, n), 2 = 0.1*noise + sweep.
JADE_DEVICE float synth_noise(unsigned long long seed, unsigned long long stream, unsigned long long n)
{
    unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (n + 1) + 0xD1B54A32D192ED03ULL * (stream + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f) - 0.5f;
}
JADE_KERNEL(256) synth_kernel(float* out, long long stream_stride, long long channel_stride, int nstreams, int channels,
                              long long nsamples, int kind, unsigned long long seed, float fs)
{
    const long long total = (long long)nstreams * channels * nsamples;
    const double dur = (double)nsamples / fs, f0 = 20.0, f1 = 0.475 * fs;
    const double kr = (f1 - f0) / dur;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long n = i % nsamples;
        const int ch = (int)((i / nsamples) % channels);
        const int st = (int)(i / (nsamples * channels));
        float v = 0.f;
        if (kind != 1) {
            const double tt = (double)n / fs;
            double ph = f0 * tt + 0.5 * kr * tt * tt + 0.25 * ch; // cycles
            ph -= floor(ph);
            v = 0.5f * sinpif((float)(2.0 * ph));
        }
        if (kind == 1) v = synth_noise(seed, (unsigned long long)st * channels + ch, (unsigned long long)n);
        if (kind == 2) v += 0.1f * synth_noise(seed, (unsigned long long)st * channels + ch, (unsigned long long)n);
        out[st * stream_stride + ch * channel_stride + n] = v;
    }
}
#endif // JADE_HELPER_KERNELS

} // namespace jade
