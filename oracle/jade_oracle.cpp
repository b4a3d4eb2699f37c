/*
 * jade_oracle.cpp -- CPU oracle for the JadeSpectrogram hot path.  TEST INFRASTRUCTURE ONLY (see jade_oracle.h).
 *
 * Every function cites the reference lines it restates (paths relative to /root/reference).
 * Arithmetic types are kept exactly as the reference evaluates them on gcc/x86-64:
 *   - cos/exp/sqrt/log10/log/pow/fabs resolve to the double overloads (the reference calls them unqualified),
 *   - sample, window, power and dB storage is float,
 *   - no FMA contraction (build with -ffp-contract=off).
 * FFT: PARITY UNPINNED (external TGM "FFT.h" is not in the reference tree) -- textbook unnormalised DFT power.
 */
#include "jade_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>
#include <thread>
#include <vector>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace {

/* ------------------------------------------------------------------------------------------------
 * FFT stand-in for `spectrum` (FFT.h, absent): call sites Spectrogram.cpp:19 (ctor), :215 (setFFTSize),
 * :144 (power(float*, std::vector<float>&)).  Real forward DFT via an N/2-point complex radix-2 FFT of
 * the even/odd packed signal plus the split post-pass; power[k] = Re^2 + Im^2, k = 0..N/2, no scaling.
 * ---------------------------------------------------------------------------------------------- */
template <typename T>
class RealFftPower {
public:
    explicit RealFftPower(size_t n = 0) { setFFTSize(n); }
    void setFFTSize(size_t n)
    {
        if (n == m_n) return;
        m_n = n;
        m_half = n / 2;
        m_tw.clear();
        m_split.clear();
        m_rev.clear();
        if (n < 2 || (n & (n - 1)) != 0) return;
        m_work.resize(m_half);
        m_tw.resize(m_half / 2 + 1);
        for (size_t k = 0; k < m_tw.size(); ++k) {
            double a = -2.0 * M_PI * double(k) / double(m_half ? m_half : 1);
            m_tw[k] = std::complex<T>(T(std::cos(a)), T(std::sin(a)));
        }
        m_split.resize(m_half / 2 + 1);
        for (size_t k = 0; k < m_split.size(); ++k) {
            double a = -2.0 * M_PI * double(k) / double(n);
            m_split[k] = std::complex<T>(T(std::cos(a)), T(std::sin(a)));
        }
        m_rev.resize(m_half);
        size_t bits = 0;
        while ((size_t(1) << bits) < m_half) ++bits;
        for (size_t i = 0; i < m_half; ++i) {
            size_t r = 0;
            for (size_t b = 0; b < bits; ++b)
                if (i & (size_t(1) << b)) r |= size_t(1) << (bits - 1 - b);
            m_rev[i] = r;
        }
    }
    bool valid() const { return !m_rev.empty() || m_n == 2; }

    /* in: n real samples (already windowed); out: n/2+1 powers */
    template <typename O>
    void power(const float* in, O* out)
    {
        const size_t h = m_half;
        if (m_n == 2) {
            T a = T(in[0]) + T(in[1]), b = T(in[0]) - T(in[1]);
            out[0] = O(a * a);
            out[1] = O(b * b);
            return;
        }
        /* Raw interleaved (re, im) arrays instead of std::complex element access: the same operations in the same order
         * (bit-identical results), but the loops compile to straight scalar code (about 10x faster at -O2), so that the
         * CPU baseline timed through this stand-in is not handicapped by the container class. */
        T* __restrict w = reinterpret_cast<T*>(m_work.data());
        const T* __restrict tw = reinterpret_cast<const T*>(m_tw.data());
        for (size_t i = 0; i < h; ++i) {
            w[2 * m_rev[i]] = T(in[2 * i]);
            w[2 * m_rev[i] + 1] = T(in[2 * i + 1]);
        }
        for (size_t len = 2; len <= h; len <<= 1) {
            const size_t step = h / len, half = len / 2;
            for (size_t base = 0; base < h; base += len) {
                T* __restrict pa = w + 2 * base;
                T* __restrict pb = w + 2 * (base + half);
                for (size_t j = 0; j < half; ++j) {
                    const T wr = tw[2 * j * step], wi = tw[2 * j * step + 1];
                    const T ar = pa[2 * j], ai = pa[2 * j + 1], br = pb[2 * j], bi = pb[2 * j + 1];
                    const T tr = wr * br - wi * bi;
                    const T ti = wr * bi + wi * br;
                    pa[2 * j] = ar + tr;
                    pa[2 * j + 1] = ai + ti;
                    pb[2 * j] = ar - tr;
                    pb[2 * j + 1] = ai - ti;
                }
            }
        }
        /* split: X[k] = (Z[k]+conj Z[h-k])/2 - i/2 * W_N^k * (Z[k]-conj Z[h-k]) */
        for (size_t k = 0; k <= h / 2; ++k) {
            const std::complex<T> zk = m_work[k];
            const std::complex<T> zp = std::conj(m_work[(h - k) % h]);
            const std::complex<T> e = (zk + zp) * T(0.5);
            const std::complex<T> d = (zk - zp) * T(0.5);
            const std::complex<T> w = m_split[k];
            /* -i * w * d */
            const T pr = w.real() * d.real() - w.imag() * d.imag();
            const T pi = w.real() * d.imag() + w.imag() * d.real();
            const std::complex<T> o(pi, -pr);
            const std::complex<T> xk = e + o;
            /* X[h-k] = conj(e) - ... : by symmetry X[h-k] = conj(e - o) with e,o taken at k */
            const std::complex<T> xm = std::conj(e - o);
            out[k] = O(xk.real() * xk.real() + xk.imag() * xk.imag());
            out[h - k] = O(xm.real() * xm.real() + xm.imag() * xm.imag());
        }
        /* k = 0 and k = h (DC / Nyquist) are real: Z[0].re +- Z[0].im */
        {
            const T a = m_work[0].real() + m_work[0].imag();
            const T b = m_work[0].real() - m_work[0].imag();
            out[0] = O(a * a);
            out[h] = O(b * b);
        }
    }

private:
    size_t m_n = 0, m_half = 0;
    std::vector<std::complex<T>> m_tw, m_split, m_work;
    std::vector<size_t> m_rev;
};

/* Spectrogram.cpp:36 */
const float kMinValForLog = 0.00000000001f;

inline float to_db(float p)
{
    /* Spectrogram.cpp:107 : float add, double log10, double multiply, float store */
    float shifted = p + kMinValForLog;
    return float(10.0 * std::log10(double(shifted)));
}

/* Spectrogram.cpp:239-293 */
void make_window(int choice, size_t n, std::vector<float>& w)
{
    w.resize(n);
    float norm = 0.f;
    for (size_t kk = 0; kk < n; ++kk) {
        const double ph = 2.0 * M_PI * kk / n;
        float v = 0.f;
        switch (choice) {
        case JO_WIN_RECT: v = 1.f; break;
        case JO_WIN_HANN: v = float(0.5f * (1.f - std::cos(ph))); break;
        case JO_WIN_HAMMING: v = float(25.0 / 46.0 - (1.0 - 25.0 / 46.0) * std::cos(ph)); break;
        case JO_WIN_BLACKMANHARRIS: {
            const float a0 = 0.35875, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
            v = float(a0 - a1 * std::cos(2.0 * M_PI * kk / n) + a2 * std::cos(4.0 * M_PI * kk / n)
                      - a3 * std::cos(6.0 * M_PI * kk / n));
            break;
        }
        case JO_WIN_FLATTOP: {
            const float a0 = 0.21557895, a1 = 0.41663158, a2 = 0.277263158, a3 = 0.083578947, a4 = 0.006947368;
            v = float(a0 - a1 * std::cos(2.0 * M_PI * kk / n) + a2 * std::cos(4.0 * M_PI * kk / n)
                      - a3 * std::cos(6.0 * M_PI * kk / n) + a4 * std::cos(8.0 * M_PI * kk / n));
            break;
        }
        case JO_WIN_HANNPOISSON: {
            const float alpha = 2.0;
            /* :280 -- (m_fftsize - 2*kk) is size_t arithmetic: wraps for kk > n/2, exp(-huge) == 0 */
            const size_t d = n - 2 * kk;
            v = float(0.5f * (1.f - std::cos(ph)) * std::exp(-alpha * std::fabs(double(d)) / n));
            break;
        }
        default: v = 1.f; break;
        }
        w[kk] = v;
        norm += w[kk] * w[kk]; /* sequential float accumulation, :283 */
    }
    norm /= n;                     /* float / size_t -> float, :285 */
    norm = float(std::sqrt(double(norm)));
    for (size_t kk = 0; kk < n; ++kk) w[kk] /= norm;
}

} // namespace

/* ================================================================================================
 * Restated Spectrogram
 * ============================================================================================== */
struct jo_spec {
    /* Spectrogram.cpp:16-24 constructor defaults */
    float m_fs = 48000.0;
    size_t m_channels = 2;
    float m_feed_percent = 100.0;
    int m_feed_samples = 1024;
    int m_feedblocks = 1;
    float m_memsize_s = 1.0;
    int m_memsize_blocks = 0;
    size_t m_freqsize = 0;
    size_t m_fftsize = 1024;
    int m_mode = JO_MIX_ABSMEAN;
    std::vector<std::vector<float>> m_mem;
    int m_newEntryCounter = int(100000000000LL);
    int m_memCounter = 0;
    std::vector<float> m_intime;
    std::vector<std::vector<float>> m_indatamem;
    int m_inCounter = 0;
    std::vector<std::vector<float>> m_power;
    std::vector<float> m_powerfinal;
    RealFftPower<float> m_fft{1024};
    RealFftPower<double> m_fft64{1024};
    bool m_useDouble = false;
    int m_windowChoice = JO_WIN_HANN;
    std::vector<float> m_window;
    bool m_PauseMode = false;

    jo_spec() { buildmem(); }

    /* Spectrogram.cpp:213-238 */
    void buildmem()
    {
        m_fft.setFFTSize(m_fftsize);
        m_fft64.setFFTSize(m_fftsize);
        m_feed_samples = int(m_feed_percent * 0.01 * m_fftsize + 0.5);
        m_memsize_blocks = int(m_memsize_s * m_fs / m_feed_samples + 0.5);
        m_freqsize = m_fftsize / 2 + 1;
        m_mem.resize(m_memsize_blocks);
        for (int kk = 0; kk < m_memsize_blocks; kk++) {
            m_mem.at(kk).resize(m_freqsize);
            std::fill(m_mem.at(kk).begin(), m_mem.at(kk).end(), -120.0);
        }
        m_intime.resize(m_fftsize);
        m_powerfinal.resize(m_freqsize);
        m_power.resize(m_channels);
        m_indatamem.resize(m_channels);
        for (size_t cc = 0; cc < m_channels; ++cc) {
            m_power.at(cc).resize(m_freqsize);
            m_indatamem.at(cc).resize(2 * m_fftsize);
            std::fill(m_indatamem.at(cc).begin(), m_indatamem.at(cc).end(), 0.0);
        }
        m_inCounter = int(m_fftsize);
        m_newEntryCounter = int(100000000000LL);
        m_memCounter = 0;
    }

    /* Spectrogram.cpp:137-145 */
    void computePowerSpectrum(std::vector<float>& in, std::vector<float>& power)
    {
        for (size_t kk = 0; kk < m_fftsize; ++kk) in[kk] *= m_window[kk];
        if (m_useDouble) {
            std::vector<double> p(m_freqsize);
            m_fft64.power(in.data(), p.data());
            for (size_t k = 0; k < m_freqsize; ++k) power[k] = float(p[k]);
        } else {
            m_fft.power(in.data(), power.data());
        }
    }

    /* Spectrogram.cpp:37-135; data is planar [channels][fftsize] */
    int process(const float* data)
    {
        for (size_t kk = 0; kk < m_fftsize; ++kk) {
            for (size_t cc = 0; cc < m_channels; ++cc) m_indatamem[cc][m_inCounter] = data[cc * m_fftsize + kk];
            m_inCounter++;
        }
        for (int bb = 0; bb < m_feedblocks; bb++) {
            for (size_t cc = 0; cc < m_channels; ++cc) {
                for (size_t kk = 0; kk < m_fftsize; ++kk) m_intime[kk] = m_indatamem[cc][kk + m_feed_samples * bb];
                computePowerSpectrum(m_intime, m_power[cc]);
            }
            for (size_t kk = 0; kk < m_freqsize; ++kk) {
                switch (m_mode) {
                case JO_MIX_ABSMEAN:
                    m_powerfinal[kk] = 0.0;
                    for (size_t cc = 0; cc < m_channels; ++cc) m_powerfinal[kk] += m_power[cc][kk];
                    m_powerfinal[kk] /= m_channels;
                    break;
                case JO_MIX_MAX:
                    m_powerfinal[kk] = 0.0;
                    for (size_t cc = 0; cc < m_channels; ++cc)
                        if (m_power[cc][kk] > m_powerfinal[kk]) m_powerfinal[kk] = m_power[cc][kk];
                    break;
                case JO_MIX_MIN:
                    m_powerfinal[kk] = 1000000.0;
                    for (size_t cc = 0; cc < m_channels; ++cc)
                        if (m_power[cc][kk] < m_powerfinal[kk]) m_powerfinal[kk] = m_power[cc][kk];
                    break;
                case JO_MIX_LEFT: m_powerfinal[kk] = m_power[0][kk]; break;
                case JO_MIX_RIGHT:
                    /* :98 guards with m_channels>0 (a latent bug for mono); the oracle guards >1 and
                     * documents the difference. */
                    if (m_channels > 1) m_powerfinal[kk] = m_power[1][kk];
                    else m_powerfinal[kk] = m_power[0][kk];
                    break;
                }
                m_powerfinal[kk] = to_db(m_powerfinal[kk]);
            }
            if (!m_PauseMode) {
                m_newEntryCounter++;
                m_mem.at(m_memCounter) = m_powerfinal;
                m_memCounter++;
                if (m_memCounter == m_memsize_blocks) m_memCounter = 0;
            }
        }
        if (m_inCounter == int(2 * m_fftsize)) {
            m_inCounter = int(m_fftsize);
            for (size_t kk = 0; kk < m_fftsize; ++kk)
                for (size_t cc = 0; cc < m_channels; ++cc) m_indatamem[cc][kk] = m_indatamem[cc][kk + m_fftsize];
        }
        return 0;
    }

    /* Spectrogram.cpp:295-331; mem is [w][freqsize] */
    int getMem(float* mem, int w, int& pos)
    {
        if (size_t(w) != m_mem.size()) return -1;
        auto copycol = [&](size_t kk) { std::copy(m_mem.at(kk).begin(), m_mem.at(kk).end(), mem + kk * m_freqsize); };
        if (size_t(m_newEntryCounter) >= size_t(w)) {
            for (size_t kk = 0; kk < size_t(w); kk++) copycol(kk);
        } else {
            int startpos = m_memCounter - m_newEntryCounter;
            if (startpos >= 0) {
                for (size_t kk = startpos; kk < size_t(m_memCounter); kk++) copycol(kk);
            } else {
                for (size_t kk = 0; kk < size_t(m_memCounter); kk++) copycol(kk);
                for (size_t kk = size_t(w) + startpos; kk < size_t(w); kk++) copycol(kk);
            }
        }
        int newVals = m_newEntryCounter;
        m_newEntryCounter = 0;
        pos = m_memCounter;
        return newVals;
    }

    size_t nextpow2(float ms) const
    {
        /* Spectrogram.cpp:171-176 */
        float firstguess = float(ms * 0.001 * m_fs);
        int np2 = int(std::log(double(firstguess)) / std::log(double(2.f))) + 1;
        return size_t(std::pow(double(2.f), np2));
    }
};

/* ================================================================================================
 * Restated CColorPalette.  State semantics (persistent table, write order) follow CColorpalette.cpp so that
 * the kMono+invert quirk (entries above n/2 keep their previous content) is reproduced.
 * ============================================================================================== */
#include "jade_oracle_cmaps.inc"

struct jo_pal {
    std::vector<int> col;
    int n = 2;
    float mx = 1.f, mn = 0.f, mult = 2.f;
    int scheme = JO_PAL_MONO;
    int invert = 0;

    jo_pal(int ncolors, int sch) : n(ncolors), scheme(sch)
    {
        /* CColorpalette.cpp:3-32 */
        mn = 0.f;
        mx = 1.f;
        mult = float(n) / (mx - mn);
        allocate();
    }
    void allocate() /* :95-99 */
    {
        col.resize(n);
        compute();
    }
    void put(int kk, int c) /* the invert write pattern shared by every scheme */
    {
        if (invert) col[n - kk - 1] = c;
        else col[kk] = c;
    }
    static int pack(int r, int g, int b) { return (r << 16) | (g << 8) | b; }

    void compute() /* CColorpalette.cpp:100-339 */
    {
        const int half = n / 2;
        switch (scheme) {
        case JO_PAL_MONO: /* :105-124 */
            for (int kk = 0; kk < n; kk++) {
                if (kk <= half) col[kk] = 0;
                else put(kk, pack(255, 255, 255));
            }
            break;
        case JO_PAL_BW: /* :126-139 */
            for (int kk = 0; kk < n; kk++) {
                const int g = int(255.f * float(kk) / n);
                put(kk, pack(g, g, g));
            }
            break;
        case JO_PAL_RAINBOW: /* :140-209 */
            for (int kk = 0; kk < n; kk++) {
                int r, g, b;
                const float slope = 4.f / float(n);
                if (kk < n / 8) {
                    b = int(255.f * (float(kk) * slope + 0.5));
                    g = 0;
                    r = 0;
                } else if (kk < 3 * n / 8) { /* two identical branches in the reference (:152-165) */
                    b = 255;
                    g = int(255.f * float(kk - n / 8) * slope);
                    r = 0;
                } else if (kk < 5 * n / 8) { /* :166-177, two identical branches */
                    b = int(255.f * float(1.f - float(kk - 3 * n / 8) * slope));
                    g = 255;
                    r = int(255.f * float(kk - 3 * n / 8) * slope);
                } else if (kk < 7 * n / 8) { /* :178-191 */
                    b = 0;
                    g = int(255.f * float(1.f - float(kk - 5 * n / 8) * slope));
                    r = 255;
                } else { /* :192-198 */
                    b = 0;
                    g = 0;
                    r = int(255.f * float(1.f - float(kk - 7 * n / 8) * slope));
                }
                put(kk, (r << 16) | (g << 8) | b);
            }
            break;
        case JO_PAL_HOT: /* :210-250 */
            for (int kk = 0; kk < n; kk++) {
                int r, g, b;
                const float s3 = 8.f / float(3 * n);
                const float s2 = 8.f / float(2 * n);
                if (kk < 3 * n / 8) {
                    b = 0;
                    g = 0;
                    r = int(255.f * (float(kk) * s3));
                } else if (kk < 6 * n / 8) {
                    b = 0;
                    g = int(255.f * float(kk - 3 * n / 8) * s3);
                    r = 255;
                } else {
                    b = int(255.f * float(kk - 6 * n / 8) * s2);
                    g = 255;
                    r = 255;
                }
                put(kk, (r << 16) | (g << 8) | b);
            }
            break;
        case JO_PAL_VIRIDIS: /* :252-269 */
        case JO_PAL_PLASMA:  /* :270-287 */
            for (int kk = 0; kk < n; kk++) {
                const int src = 256;
                const int index = int(float(kk) / n * src);
                /* int(cm[index][c]*255) per channel, precomputed per source entry (tools/gen_cmap_rgb8.py) */
                put(kk, (scheme == JO_PAL_VIRIDIS) ? jo_cm_viridis_rgb8[index] : jo_cm_plasma_rgb8[index]);
            }
            break;
        case JO_PAL_JADE: { /* :288-335 */
            const float rs = 0.3529, rm = 0.89019, re = 0.95;
            const float gs = 0.372549, gm = 0.023529, ge = 0.95;
            const float bs = 0.33725, bm = 0.074509, be = 0.95;
            const int mix = 2 * n / 4;
            for (int kk = 0; kk < n; kk++) {
                int r, g, b;
                if (kk < mix) {
                    b = int(255 * (float(kk) / mix * (bm - bs) + bs));
                    g = int(255 * (float(kk) / mix * (gm - gs) + gs));
                    r = int(255 * (float(kk) / mix * (rm - rs) + rs));
                } else {
                    b = int(255 * (float(kk - mix) / mix * (be - bm) + bm));
                    g = int(255 * (float(kk - mix) / mix * (ge - gm) + gm));
                    r = int(255 * (float(kk - mix) / mix * (re - rm) + rm));
                }
                put(kk, (r << 16) | (g << 8) | b);
            }
            break;
        }
        default: break;
        }
    }
    void setRange(float a, float b) /* :39-54 */
    {
        if (b >= a) {
            mn = a;
            mx = b;
        } else {
            mn = b;
            mx = a;
        }
        if (mx == mn) mn = float(0.99 * mx);
        mult = float(n) / (mx - mn);
    }
    int rgb(float v) const /* CColorpalette.h:32-47 */
    {
        if (v >= mx) v = mx * 0.9999f;
        if (v < mn) v = mn;
        int index = int((v - mn) * mult);
        if (index < n) return col[index];
        return col[n - 1];
    }
    float value(int c) const /* :82-94 */
    {
        for (int kk = 0; kk < n; kk++)
            if (col[kk] == c) return float(kk) / mult + mn;
        return 100000000000000000000000000000.f;
    }
};

/* ================================================================================================
 * Restated SpectrogramComponent::timerCallback pixel loops (Spectrogram.cpp:590-724).
 * JUCE Image::RGB -> row-major uint32 [H][W]; setPixelColour(x,y,Colour(c)) -> pix[y*W+x] = c;
 * moveImageSection(0,0,newVals,0,W-newVals,H) -> shift every row left by newVals;
 * juce::Colours::red -> 0xFFFF0000.
 * ============================================================================================== */
struct jo_view {
    jo_spec* spec;
    jo_pal* pal;
    std::vector<float> disp; /* m_displaymem [W][H] */
    std::vector<uint32_t> img;
    int W = 1, H = 1;
    bool recomputeAll = true;
    bool running = true;
    float minC = -50.f, maxC = 50.f; /* PlugInGUISettings.h:37-38 */

    int tick()
    {
        int wData = spec->m_memsize_blocks;
        int hData = int(spec->m_freqsize);
        if (wData != W || hData != H) { /* :595-605 (rescaled() content is irrelevant: fully overwritten below) */
            W = wData;
            H = hData;
            img.assign(size_t(W) * H, 0xFF000000u);
            disp.assign(size_t(W) * H, 0.f);
        }
        int pos = 0;
        int newVals = spec->getMem(disp.data(), W, pos);
        if (newVals > W) recomputeAll = true;
        pal->setRange(minC, maxC);
        auto px = [&](int x, int y, uint32_t c) { img[size_t(y) * W + x] = c; };
        auto colour = [&](int col, int hh) { return uint32_t(pal->rgb(disp[size_t(col) * H + hh])) | 0xFF000000u; };
        const uint32_t red = 0xFFFF0000u;
        if (recomputeAll) { /* :623-657 */
            recomputeAll = false;
            int newwstart = W - pos;
            for (int ww = 0; ww < W; ++ww) {
                int neww = ww + newwstart;
                if (neww >= W) neww -= W;
                for (int hh = 0; hh < H; ++hh) {
                    if (running) px(neww, H - 1 - hh, colour(ww, hh));
                    else px(ww, H - 1 - hh, colour(ww, hh));
                }
            }
            if (!running)
                for (int hh = 0; hh < H; ++hh) px(pos % W, H - 1 - hh, red);
        } else { /* :658-724 */
            int startread = pos - newVals;
            if (running) {
                if (newVals > 0 && newVals < W)
                    for (int y = 0; y < H; ++y)
                        std::memmove(&img[size_t(y) * W], &img[size_t(y) * W + newVals], size_t(W - newVals) * 4);
                for (int ww = W - newVals; ww < W; ++ww) {
                    int readpos = startread < 0 ? wData + startread : startread;
                    for (int hh = 0; hh < H; ++hh) px(ww, H - 1 - hh, colour(readpos, hh));
                    startread++;
                }
            } else {
                for (int ww = 0; ww < newVals; ++ww) {
                    int readpos = startread < 0 ? wData + startread : startread;
                    for (int hh = 0; hh < H; ++hh) px(readpos, H - 1 - hh, colour(readpos, hh));
                    startread++;
                }
                int drawwidth = 1;
                if (H < 2048) drawwidth++;
                if (H < 1024) drawwidth += 2;
                for (int hh = 0; hh < H; ++hh)
                    for (int dd = 0; dd < drawwidth; ++dd) {
                        int drawpos = pos + dd;
                        /* :715 only wraps the == W case; the oracle wraps with % to stay in bounds */
                        if (drawpos >= W) drawpos -= W;
                        px(drawpos, H - 1 - hh, red);
                    }
            }
        }
        return newVals;
    }
};

/* ================================================================================================
 * C interface
 * ============================================================================================== */
extern "C" {

int jo_window(int window, int n, float* out)
{
    if (n <= 0 || !out) return -1;
    std::vector<float> w;
    make_window(window, size_t(n), w);
    std::copy(w.begin(), w.end(), out);
    return 0;
}
/* The plan (twiddles, split table, bit reversal) is kept per thread and rebuilt only when n changes, as any plan-caching
 * FFT class would between setFFTSize calls (Spectrogram.cpp:215); results are bit-identical to a fresh plan. */
int jo_power_f32(const float* x, int n, float* power)
{
    static thread_local RealFftPower<float> f;
    f.setFFTSize(size_t(n));
    if (!f.valid()) return -1;
    f.power(x, power);
    return 0;
}
int jo_power_f64(const float* x, int n, double* power)
{
    static thread_local RealFftPower<double> f;
    f.setFFTSize(size_t(n));
    if (!f.valid()) return -1;
    f.power(x, power);
    return 0;
}
float jo_db(float p) { return to_db(p); }

jo_spec* jo_spec_create(void)
{
    jo_spec* s = new jo_spec();
    make_window(s->m_windowChoice, s->m_fftsize, s->m_window);
    return s;
}
void jo_spec_destroy(jo_spec* s) { delete s; }
void jo_spec_set_samplerate(jo_spec* s, float fs) /* :148-152 */
{
    s->m_fs = fs;
    s->buildmem();
}
void jo_spec_set_channels(jo_spec* s, size_t n) /* :153-157 */
{
    s->m_channels = n;
    s->buildmem();
}
void jo_spec_set_fftsize(jo_spec* s, size_t n) /* :160-170 */
{
    s->m_fftsize = n;
    s->buildmem();
    make_window(s->m_windowChoice, s->m_fftsize, s->m_window);
    s->m_newEntryCounter = int(100000000000LL);
}
void jo_spec_set_closest_fftsize_ms(jo_spec* s, float ms) /* :177-183 */
{
    s->m_fftsize = s->nextpow2(ms);
    s->buildmem();
    make_window(s->m_windowChoice, s->m_fftsize, s->m_window);
}
void jo_spec_set_memory_time_s(jo_spec* s, float t) /* :184-188 */
{
    s->m_memsize_s = t;
    s->buildmem();
}
void jo_spec_set_feed_percent(jo_spec* s, int feed) /* :189-211 */
{
    switch (feed) {
    case JO_FEED_100: s->m_feed_percent = 100.0; s->m_feedblocks = 1; break;
    case JO_FEED_50: s->m_feed_percent = 50.0; s->m_feedblocks = 2; break;
    case JO_FEED_25: s->m_feed_percent = 25.0; s->m_feedblocks = 4; break;
    case JO_FEED_10: s->m_feed_percent = 10.0; s->m_feedblocks = 10; break;
    default: break;
    }
    s->buildmem();
}
void jo_spec_set_pause(jo_spec* s, int on) { s->m_PauseMode = on != 0; }
void jo_spec_set_window(jo_spec* s, int w) /* Spectrogram.h:123 */
{
    s->m_windowChoice = w;
    make_window(s->m_windowChoice, s->m_fftsize, s->m_window);
}
void jo_spec_set_mix_mode(jo_spec* s, int m) { s->m_mode = m; }
void jo_spec_set_fft_double(jo_spec* s, int on) { s->m_useDouble = on != 0; }
size_t jo_spec_next_pow2(jo_spec* s, float ms) { return s->nextpow2(ms); }
int jo_spec_spectrum_size(jo_spec* s) { return int(s->m_freqsize); }
int jo_spec_memory_size(jo_spec* s) { return s->m_memsize_blocks; }
int jo_spec_feed_samples(jo_spec* s) { return s->m_feed_samples; }
int jo_spec_feed_blocks(jo_spec* s) { return s->m_feedblocks; }
float jo_spec_samplerate(jo_spec* s) { return s->m_fs; }
int jo_spec_process_block(jo_spec* s, const float* planar) { return s->process(planar); }
int jo_spec_get_mem(jo_spec* s, float* mem, int w, int* pos)
{
    int p = 0;
    int r = s->getMem(mem, w, p);
    if (pos) *pos = p;
    return r;
}

jo_pal* jo_pal_create(int n, int scheme) { return new jo_pal(n, scheme); }
jo_pal* jo_pal_create_default(void) { return new jo_pal(2, JO_PAL_MONO); }
void jo_pal_destroy(jo_pal* p) { delete p; }
void jo_pal_set_value_range(jo_pal* p, float a, float b) { p->setRange(a, b); }
void jo_pal_set_nr_of_colors(jo_pal* p, int n) /* :55-61 */
{
    p->n = n;
    p->mult = float(p->n) / (p->mx - p->mn);
    p->allocate();
}
void jo_pal_set_color_scheme(jo_pal* p, int s) /* :62-66 */
{
    p->scheme = s;
    p->compute();
}
void jo_pal_set_invert(jo_pal* p, int on) { p->invert = on; } /* CColorpalette.h:29: no recompute */
int jo_pal_get_rgb(jo_pal* p, float v) { return p->rgb(v); }
float jo_pal_get_value(jo_pal* p, int c) { return p->value(c); }
int jo_pal_table(jo_pal* p, int* out, int cap)
{
    for (int i = 0; i < p->n && i < cap; ++i) out[i] = p->col[i];
    return p->n;
}
void jo_pal_get_range(jo_pal* p, float* a, float* b, float* m)
{
    if (a) *a = p->mn;
    if (b) *b = p->mx;
    if (m) *m = p->mult;
}
void jo_pal_lookup_many(jo_pal* p, const float* v, int n, int* out)
{
    for (int i = 0; i < n; ++i) out[i] = p->rgb(v[i]);
}

jo_view* jo_view_create(jo_spec* s, jo_pal* p)
{
    jo_view* v = new jo_view();
    v->spec = s;
    v->pal = p;
    return v;
}
void jo_view_destroy(jo_view* v) { delete v; }
void jo_view_set_running(jo_view* v, int r) { v->running = r != 0; }
void jo_view_set_color_range(jo_view* v, float a, float b)
{
    v->minC = a;
    v->maxC = b;
}
void jo_view_force_recompute(jo_view* v) { v->recomputeAll = true; }
int jo_view_tick(jo_view* v) { return v->tick(); }
int jo_view_width(jo_view* v) { return v->W; }
int jo_view_height(jo_view* v) { return v->H; }
const uint32_t* jo_view_pixels(jo_view* v) { return v->img.data(); }

/* ---- batch ---- */
namespace {
struct BatchWorker {
    jo_batch_cfg c;
    std::vector<float> win, frame;
    std::vector<std::vector<float>> pw;
    std::vector<float> pf;
    RealFftPower<float> f32;
    RealFftPower<double> f64;
    jo_pal pal;
    explicit BatchWorker(const jo_batch_cfg& cfg)
        : c(cfg), f32(cfg.fft_size), f64(cfg.fft_size), pal(cfg.palette_size, JO_PAL_MONO)
    {
        make_window(c.window, size_t(c.fft_size), win);
        frame.resize(c.fft_size);
        pw.assign(c.channels, std::vector<float>(c.fft_size / 2 + 1));
        pf.resize(c.fft_size / 2 + 1);
        pal.invert = c.palette_invert;
        pal.scheme = c.palette_scheme;
        pal.compute();
        pal.setRange(c.min_db, c.max_db);
    }
    /* one column: same arithmetic as jo_spec::process for one sub-frame */
    void column(const float* samples, long nsamples, long j, float* db_out, uint32_t* pix_out)
    {
        const int N = c.fft_size, B = N / 2 + 1;
        const long start = j * long(c.hop) - N;
        for (int cc = 0; cc < c.channels; ++cc) {
            const float* x = samples + long(cc) * nsamples;
            for (int k = 0; k < N; ++k) {
                long idx = start + k;
                frame[k] = (idx >= 0 && idx < nsamples) ? x[idx] : 0.f;
            }
            for (int k = 0; k < N; ++k) frame[k] *= win[k];
            if (c.use_double_fft) {
                std::vector<double> p(B);
                f64.power(frame.data(), p.data());
                for (int k = 0; k < B; ++k) pw[cc][k] = float(p[k]);
            } else {
                f32.power(frame.data(), pw[cc].data());
            }
        }
        for (int kk = 0; kk < B; ++kk) {
            float v = 0.f;
            switch (c.mix_mode) {
            case JO_MIX_ABSMEAN:
                v = 0.0;
                for (int cc = 0; cc < c.channels; ++cc) v += pw[cc][kk];
                v /= size_t(c.channels);
                break;
            case JO_MIX_MAX:
                v = 0.0;
                for (int cc = 0; cc < c.channels; ++cc)
                    if (pw[cc][kk] > v) v = pw[cc][kk];
                break;
            case JO_MIX_MIN:
                v = 1000000.0;
                for (int cc = 0; cc < c.channels; ++cc)
                    if (pw[cc][kk] < v) v = pw[cc][kk];
                break;
            case JO_MIX_LEFT: v = pw[0][kk]; break;
            case JO_MIX_RIGHT: v = c.channels > 1 ? pw[1][kk] : pw[0][kk]; break;
            }
            pf[kk] = to_db(v);
        }
        if (db_out) std::copy(pf.begin(), pf.end(), db_out);
        if (pix_out)
            for (int kk = 0; kk < B; ++kk) pix_out[B - 1 - kk] = uint32_t(pal.rgb(pf[kk])) | 0xFF000000u;
    }
};
} // namespace

long jo_render_batch(const jo_batch_cfg* cfg, const float* samples, long nsamples, long first_col, long ncols,
                     float* db_out, uint32_t* pix_out)
{
    if (!cfg || cfg->fft_size < 2 || (cfg->fft_size & (cfg->fft_size - 1)) || cfg->hop <= 0 || cfg->channels <= 0)
        return -1;
    BatchWorker w(*cfg);
    const long B = cfg->fft_size / 2 + 1;
    for (long j = 0; j < ncols; ++j)
        w.column(samples, nsamples, first_col + j, db_out ? db_out + j * B : nullptr, pix_out ? pix_out + j * B : nullptr);
    return ncols;
}

double jo_bench_batch(const jo_batch_cfg* cfg, const float* samples, long nsamples, int nstreams, int nthreads,
                      long* frames_out)
{
    const long B = cfg->fft_size / 2 + 1;
    const long cols = nsamples / cfg->hop;
    if (nthreads < 1) nthreads = 1;
    std::vector<std::thread> th;
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t]() {
            BatchWorker w(*cfg);
            std::vector<float> db(B);
            std::vector<uint32_t> px(B);
            volatile uint32_t sink = 0;
            for (int s = t; s < nstreams; s += nthreads)
                for (long j = 0; j < cols; ++j) {
                    w.column(samples, nsamples, j, db.data(), px.data());
                    sink = sink + px[0];
                }
        });
    }
    for (auto& x : th) x.join();
    auto t1 = std::chrono::steady_clock::now();
    if (frames_out) *frames_out = cols * long(nstreams);
    double sec = std::chrono::duration<double>(t1 - t0).count();
    return double(cols) * nstreams / sec;
}

double jo_bench_stream(jo_spec* s, const float* planar_block, int nblocks)
{
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < nblocks; ++i) s->process(planar_block);
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
