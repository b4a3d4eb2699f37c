// jade_host_tables.h -- host-side, one-off table generation of the engine: window, palette, twiddles, row maps.
// (The reference also computes these on the host once per reconfiguration: Spectrogram.cpp:239-293,
//  CColorpalette.cpp:100-339.)  Pure C++17, no CUDA; shared by jade_gpu.cu, the drop-in classes and tests/emu.
#pragma once
#include <cstdint>
#include <vector>

namespace jade_host {

struct alignas(8) cpxf {
    float x, y;
};

// Unit-RMS periodic window table, bit-exact restatement of Spectrogram::setWindowFkt (Spectrogram.cpp:239-293).
void make_window(int kind, int n, std::vector<float>& w);

// Reference colour table builder (CColorpalette.cpp:100-339): writes into `table` (size n) in the reference's order.
void palette_build(int scheme, int n, int invert, int32_t* table);

// CColorPalette::setValueRange (CColorpalette.cpp:39-54)
struct ValueRange {
    float mn = 0.f, mx = 1.f, mult = 2.f;
    void set(float a, float b, int ncolors);
    float maxclamp() const { return mx * 0.9999f; }
};
// CColorPalette::getRGBColor (CColorpalette.h:32-47); index additionally clamped at 0
int palette_index(float v, const ValueRange& r, int n);

// exp(-2*pi*i*k*step/size) for k in [0,count)
void twiddles(int size, int count, long long step, std::vector<cpxf>& out);
// table[k1*cols + n] = exp(-2*pi*i*k1*n/size)
void twiddle_matrix(int size, int rows, int cols, std::vector<cpxf>& out);

// SpectrogramComponent::paint crop maths (Spectrogram.cpp:441-459): bins [k_lo,k_hi) shown between fmin and fmax
void linear_crop(float fs, int bins, float fmin, float fmax, int& k_lo, int& k_hi);
// Log-spaced max-pool bands (extension, see DESIGN.md "row maps")
void log_rows(float fs, int fft_size, int rows, float fmin, float fmax, std::vector<int32_t>& lo, std::vector<int32_t>& hi);

} // namespace jade_host
