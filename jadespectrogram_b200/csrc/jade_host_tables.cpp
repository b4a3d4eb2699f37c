// jade_host_tables.cpp -- see jade_host_tables.h.  Arithmetic types follow what gcc/x86-64 evaluates for the
// reference expressions (unqualified cos/exp/sqrt/fabs take the double overloads; tables are stored as float/int).
#include "jade_host_tables.h"

#include <cmath>
#include <cstddef>

#include "cmap_rgb8.inc"

namespace jade_host {

static const double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------------------------------------------------
// Windows.  Generalised cosine sums are accumulated left to right in double with float-rounded coefficients,
// which is bit-identical to `a0 - a1*cos(..) + a2*cos(..) - ...` (Spectrogram.cpp:261-275).
// ---------------------------------------------------------------------------------------------------------
namespace {
struct CosSum {
    int terms;
    float coef[5];
};
const CosSum kBlackmanHarris = {4, {0.35875f, 0.48829f, 0.14128f, 0.01168f, 0.f}};
const CosSum kFlatTop = {5, {0.21557895f, 0.41663158f, 0.277263158f, 0.083578947f, 0.006947368f}};

double cos_sum(const CosSum& cs, std::size_t k, std::size_t n)
{
    double acc = cs.coef[0];
    for (int i = 1; i < cs.terms; ++i) {
        const double ang = (2.0 * i) * kPi * k / n; // 2.0*M_PI, 4.0*M_PI, 6.0*M_PI, 8.0*M_PI times kk over N
        const double term = cs.coef[i] * std::cos(ang);
        acc = (i & 1) ? acc - term : acc + term;
    }
    return acc;
}
} // namespace

void make_window(int kind, int n_in, std::vector<float>& w)
{
    const std::size_t n = std::size_t(n_in);
    w.assign(n, 0.f);
    float energy = 0.f;
    for (std::size_t k = 0; k < n; ++k) {
        const double ph = 2.0 * kPi * k / n;
        double v;
        switch (kind) {
        case 1: v = 0.5 * (1.0 - std::cos(ph)); break;                              // Hann       (:256)
        case 2: v = 25.0 / 46.0 - (1.0 - 25.0 / 46.0) * std::cos(ph); break;        // Hamming    (:259)
        case 3: v = cos_sum(kBlackmanHarris, k, n); break;                          // (:262-267)
        case 4: v = cos_sum(kFlatTop, k, n); break;                                 // (:269-276)
        case 5: {                                                                   // HannPoisson (:278-281)
            const std::size_t wrapped = n - 2 * k; // unsigned on purpose: the reference wraps for k > n/2
            v = 0.5 * (1.0 - std::cos(ph)) * std::exp(-2.0 * std::fabs(double(wrapped)) / n);
            break;
        }
        default: v = 1.0; break;                                                    // Rect       (:253)
        }
        w[k] = float(v);
        energy += w[k] * w[k];
    }
    energy /= float(n);
    const float rms = float(std::sqrt(double(energy)));
    for (std::size_t k = 0; k < n; ++k) w[k] /= rms;
}

// ---------------------------------------------------------------------------------------------------------
// Palette.  Every scheme is a function kk -> (r,g,b) evaluated with the reference's float expressions; the
// common writer reproduces its invert handling and its kMono special case.
// ---------------------------------------------------------------------------------------------------------
namespace {
struct Rgb {
    int r, g, b;
};
inline int word(const Rgb& c) { return (c.r << 16) | (c.g << 8) | c.b; }

Rgb grey_ramp(int kk, int n)
{
    const int g = int(255.f * float(kk) / n); // :129-131
    return {g, g, g};
}
Rgb rainbow(int kk, int n)
{
    const float slope = 4.f / float(n);
    const int e = n / 8;
    if (kk < e) return {0, 0, int(255.f * (float(kk) * slope + 0.5))};                                  // :147-152
    if (kk < 3 * n / 8) return {0, int(255.f * float(kk - e) * slope), 255};                            // :153-166
    if (kk < 5 * n / 8) {                                                                               // :167-178
        const float x = float(kk - 3 * n / 8) * slope;
        return {int(255.f * float(kk - 3 * n / 8) * slope), 255, int(255.f * float(1.f - x))};
    }
    if (kk < 7 * n / 8) return {255, int(255.f * float(1.f - float(kk - 5 * n / 8) * slope)), 0};      // :179-192
    return {int(255.f * float(1.f - float(kk - 7 * n / 8) * slope)), 0, 0};                             // :193-199
}
Rgb hot(int kk, int n)
{
    const float s3 = 8.f / float(3 * n), s2 = 8.f / float(2 * n);
    if (kk < 3 * n / 8) return {int(255.f * (float(kk) * s3)), 0, 0};                                   // :224-229
    if (kk < 6 * n / 8) return {255, int(255.f * float(kk - 3 * n / 8) * s3), 0};                       // :230-236
    return {255, 255, int(255.f * float(kk - 6 * n / 8) * s2)};                                         // :237-243
}
Rgb listed(const int* rgb8, int kk, int n)
{
    const int idx = int(float(kk) / n * 256); // :257,275
    const int c = rgb8[idx];
    return {(c >> 16) & 255, (c >> 8) & 255, c & 255};
}
Rgb jade(int kk, int n)
{
    // anchors (:290-300) as float variables, like the reference
    const float r0 = 0.3529, r1 = 0.89019, r2 = 0.95;
    const float g0 = 0.372549, g1 = 0.023529, g2 = 0.95;
    const float b0 = 0.33725, b1 = 0.074509, b2 = 0.95;
    const int mix = 2 * n / 4;
    auto seg = [&](int pos, float from, float to) { return int(255 * (float(pos) / mix * (to - from) + from)); };
    if (kk < mix) return {seg(kk, r0, r1), seg(kk, g0, g1), seg(kk, b0, b1)};                           // :309-314
    return {seg(kk - mix, r1, r2), seg(kk - mix, g1, g2), seg(kk - mix, b1, b2)};                       // :316-320
}
} // namespace

void palette_build(int scheme, int n, int invert, int32_t* t)
{
    auto store = [&](int kk, int c) { t[invert ? n - kk - 1 : kk] = c; };
    for (int kk = 0; kk < n; ++kk) {
        switch (scheme) {
        case 0: // kMono :105-124 -- the lower half is written un-inverted
            if (kk <= n / 2) t[kk] = 0;
            else store(kk, 0xFFFFFF);
            break;
        case 1: store(kk, word(grey_ramp(kk, n))); break;
        case 2: store(kk, word(hot(kk, n))); break;
        case 3: store(kk, word(rainbow(kk, n))); break;
        case 4: store(kk, word(listed(jade_cm_viridis_rgb8, kk, n))); break;
        case 5: store(kk, word(listed(jade_cm_plasma_rgb8, kk, n))); break;
        case 6: store(kk, word(jade(kk, n))); break;
        default: break;
        }
    }
}

void ValueRange::set(float a, float b, int ncolors)
{
    mn = b >= a ? a : b;
    mx = b >= a ? b : a;
    if (mx == mn) mn = float(0.99 * mx);
    mult = float(ncolors) / (mx - mn);
}

int palette_index(float v, const ValueRange& r, int n)
{
    if (v >= r.mx) v = r.mx * 0.9999f;
    if (v < r.mn) v = r.mn;
    int idx = int((v - r.mn) * r.mult);
    if (idx >= n) idx = n - 1;
    if (idx < 0) idx = 0;
    return idx;
}

void twiddles(int size, int count, long long step, std::vector<cpxf>& out)
{
    out.resize(count);
    for (int k = 0; k < count; ++k) {
        const long long idx = (k * step) % size;
        const double a = 2.0 * kPi * double(idx) / double(size);
        out[k] = {float(std::cos(a)), float(-std::sin(a))};
    }
}
void twiddle_matrix(int size, int rows, int cols, std::vector<cpxf>& out)
{
    out.resize(std::size_t(rows) * cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) {
            const long long idx = (1LL * r * c) % size;
            const double a = 2.0 * kPi * double(idx) / double(size);
            out[std::size_t(r) * cols + c] = {float(std::cos(a)), float(-std::sin(a))};
        }
}

void linear_crop(float fs, int H, float fmin, float fmax, int& k_lo, int& k_hi)
{
    // Spectrogram.cpp:444-453 clamps
    if (fmin >= fs * 0.5) fmin = float(0.9 * fs * 0.5);
    if (fmax >= fs * 0.5) fmax = float(fs * 0.5);
    if (fmin >= fmax) fmin = float(0.9 * fmax);
    // :455-459
    const int endPix = int(2.0 * fmax / fs * H + 0.5);
    const int interval = int(2.0 * fmax / fs * H - 2.0 * fmin / fs * H + 0.5);
    // image rows [H-endPix, H-endPix+interval) of the flipped image <-> bins [endPix-interval, endPix)
    k_hi = endPix;
    k_lo = endPix - interval;
    if (k_hi > H) k_hi = H;
    if (k_lo < 0) k_lo = 0;
    if (k_lo >= k_hi) k_lo = k_hi > 0 ? k_hi - 1 : 0;
}

void log_rows(float fs, int n, int rows, float fmin, float fmax, std::vector<int32_t>& lo, std::vector<int32_t>& hi)
{
    const int B = n / 2 + 1;
    lo.resize(rows);
    hi.resize(rows);
    const double f0 = fmin > 0.f ? double(fmin) : 1.0;
    const double f1 = fmax > f0 ? double(fmax) : f0 * 2.0;
    const double ratio = f1 / f0;
    const double binhz = double(fs) / double(n);
    for (int r = 0; r < rows; ++r) {
        const double e0 = f0 * std::pow(ratio, double(r) / rows);
        const double e1 = f0 * std::pow(ratio, double(r + 1) / rows);
        long a = long(std::ceil(e0 / binhz));
        long b = long(std::ceil(e1 / binhz));
        if (b <= a) { // no bin centre inside the band: take the nearest bin to the band centre
            a = long(std::floor(std::sqrt(e0 * e1) / binhz + 0.5));
            b = a + 1;
        }
        if (a < 0) a = 0;
        if (a > B - 1) a = B - 1;
        if (b > B) b = B;
        if (b <= a) b = a + 1;
        lo[r] = int32_t(a);
        hi[r] = int32_t(b);
    }
}

} // namespace jade_host
