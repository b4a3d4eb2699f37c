// jade_axis.cpp -- host-side display arithmetic of SpectrogramComponent::paint (Spectrogram.cpp:432-545): the clamp rules of the
// two frequency sliders, the tick values / label boxes of the frequency axis and of the colourbar axis.  Plain integer / float
// arithmetic, no GPU, no JUCE; every expression keeps the reference's operand types (float members, int pixel sizes, double
// literals) so that the truncations land on the same integers.  The colourbar pixels themselves come from the engine's
// palette: jade_colorbar in jade_gpu.cu.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/jade_gpu.h"

namespace {
// label text of a frequency tick (Spectrogram.cpp:476-493).  The reference formats with juce::String(float), whose digits
// are JUCE's business; here: kHz with as many decimals as the value has (the value is a multiple of 100 Hz), else whole Hz.
void freq_label(float hz, char* out, size_t n)
{
    if (hz >= 1000.f) {
        const int hundreds = int(hz * 0.01 + 0.5);
        if (hundreds % 10 == 0) snprintf(out, n, "%dk", hundreds / 10);
        else snprintf(out, n, "%d.%dk", hundreds / 10, hundreds % 10);
    } else {
        snprintf(out, n, "%d", int(hz));
    }
}
// y of a label box (Spectrogram.cpp:498-507, :536-543): all but the last tick are centred on their value
int label_y(int comp_height, float scale, int menu_height, int text_height, int ydelta, bool last)
{
    if (!last) return int(comp_height - scale * menu_height - 0.5 * text_height * scale - ydelta);
    return int(comp_height - scale * menu_height - ydelta);
}
} // namespace

extern "C" {

int jade_display_freq_clamp(float fs, float* min_hz, float* max_hz)
{
    if (!min_hz || !max_hz || !(fs > 0.f)) return JADE_ERR_ARG;
    float mn = *min_hz, mx = *max_hz; // m_minDisplayFreq / m_maxDisplayFreq are floats (Spectrogram.h:195-196)
    if (mn >= fs * 0.5) mn = float(0.9 * fs * 0.5); // :444-445
    if (mx >= fs * 0.5) mx = float(fs * 0.5);       // :446-447
    if (mn >= mx) mn = float(0.9 * mx);             // :449-453
    *min_hz = mn;
    *max_hz = mx;
    return JADE_OK;
}

int jade_freq_axis_ticks(float min_hz, float max_hz, int comp_height, float scale, int menu_height, int text_height, int nticks,
                         jade_axis_tick* out)
{
    if (!out || nticks < 2 || !(max_hz > min_hz)) return JADE_ERR_ARG;
    const float range_per_tick = float(max_hz - min_hz) / (nticks - 1); // :467
    for (int kk = 0; kk < nticks; ++kk) {
        float v = float(int(min_hz + range_per_tick * kk + 0.5)); // :472
        if (v >= 1000.f) v = float(int(v * 0.01 + 0.5) * 100);     // :476  nearest 100 Hz
        else if (v >= 150.f) v = float(int(v * 0.1 + 0.5) * 10);   // :484  nearest 10 Hz
        const int ydelta = int((comp_height - scale * menu_height) * (v - min_hz) / (max_hz - min_hz)); // :498
        out[kk].value = v;
        out[kk].y = label_y(comp_height, scale, menu_height, text_height, ydelta, kk == nticks - 1);
        freq_label(v, out[kk].label, sizeof out[kk].label);
    }
    return JADE_OK;
}

int jade_color_axis_ticks(float min_val, float max_val, int comp_height, float scale, int menu_height, int text_height, int nticks,
                          jade_axis_tick* out)
{
    if (!out || nticks < 2 || !(max_val > min_val)) return JADE_ERR_ARG;
    const float range_per_tick = float(max_val - min_val) / (nticks - 1); // :526
    for (int kk = 0; kk < nticks; ++kk) {
        const float v = float(int((min_val + range_per_tick * kk) * 0.1) * 10); // :529  truncated to a multiple of 10
        const int ydelta = int((comp_height - scale * menu_height) * (v - min_val) / (max_val - min_val)); // :533
        out[kk].value = v;
        out[kk].y = label_y(comp_height, scale, menu_height, text_height, ydelta, kk == nticks - 1);
        snprintf(out[kk].label, sizeof out[kk].label, "%d", int(v));
    }
    return JADE_OK;
}

int jade_colorbar_height(int comp_height, float scale, int menu_height) { return int(comp_height - scale * menu_height); } // :510

} // extern "C"
