// dropin_shim.cpp -- C entry points onto the drop-in C++ classes (include/Spectrogram.h, include/CColorpalette.h) so the
// Python tests can drive them with the same calls they use on the oracle (jo_*) and on the compiled reference (jr_*).
// Test glue only; the classes themselves are header-only wrappers over the C ABI of libjade_gpu.so.
#include "Spectrogram.h"

#include <cstring>

extern "C" {
// ---- CColorPalette ----
void* jd_pal_create(int n, int scheme) { return new CColorPalette(n, scheme); }
void* jd_pal_create_default(void) { return new CColorPalette(); }
void jd_pal_destroy(void* p) { delete static_cast<CColorPalette*>(p); }
void jd_pal_set_value_range(void* p, float a, float b) { static_cast<CColorPalette*>(p)->setValueRange(a, b); }
void jd_pal_set_nr_of_colors(void* p, int n) { static_cast<CColorPalette*>(p)->setNrOfColors(n); }
void jd_pal_set_color_scheme(void* p, int s) { static_cast<CColorPalette*>(p)->setColorSceme(s); }
void jd_pal_set_invert(void* p, int on) { static_cast<CColorPalette*>(p)->setInvertStatus(on != 0); }
int jd_pal_get_rgb(void* p, float v) { return static_cast<CColorPalette*>(p)->getRGBColor(v); }
float jd_pal_get_value(void* p, int c) { return static_cast<CColorPalette*>(p)->getValue(c); }
void jd_pal_lookup_many(void* p, const float* v, int n, int* out)
{
    CColorPalette* c = static_cast<CColorPalette*>(p);
    for (int i = 0; i < n; ++i) out[i] = c->getRGBColor(v[i]);
}

// ---- Spectrogram ----
void* jd_spec_create(void) { return new Spectrogram(); }
void jd_spec_destroy(void* s) { delete static_cast<Spectrogram*>(s); }
int jd_spec_ok(void* s) { return static_cast<Spectrogram*>(s)->ok() ? 1 : 0; }
const char* jd_spec_error(void* s) { return static_cast<Spectrogram*>(s)->lastError().c_str(); }
void jd_spec_set_samplerate(void* s, float fs) { static_cast<Spectrogram*>(s)->setSamplerate(fs); }
void jd_spec_set_channels(void* s, size_t n) { static_cast<Spectrogram*>(s)->setchannels(n); }
void jd_spec_set_fftsize(void* s, size_t n) { static_cast<Spectrogram*>(s)->setFFTSize(n); }
void jd_spec_set_closest_fftsize_ms(void* s, float ms) { static_cast<Spectrogram*>(s)->setclosestFFTSize_ms(ms); }
void jd_spec_set_memory_time_s(void* s, float t) { static_cast<Spectrogram*>(s)->setmemoryTime_s(t); }
void jd_spec_set_feed_percent(void* s, int f) { static_cast<Spectrogram*>(s)->setfeed_percent(static_cast<Spectrogram::FeedPercentage>(f)); }
void jd_spec_set_pause(void* s, int on) { static_cast<Spectrogram*>(s)->setPauseMode(on != 0); }
void jd_spec_set_window(void* s, int w) { static_cast<Spectrogram*>(s)->setWindow(static_cast<Spectrogram::Windows>(w)); }
void jd_spec_set_mix_mode(void* s, int m) { static_cast<Spectrogram*>(s)->setMixMode(static_cast<Spectrogram::ChannelMixMode>(m)); }
size_t jd_spec_next_pow2(void* s, float ms) { return static_cast<Spectrogram*>(s)->getnextpowerof2(ms); }
int jd_spec_spectrum_size(void* s) { return static_cast<Spectrogram*>(s)->getSpectrumSize(); }
int jd_spec_memory_size(void* s) { return static_cast<Spectrogram*>(s)->getMemorySize(); }
float jd_spec_samplerate(void* s) { return static_cast<Spectrogram*>(s)->getSamplerate(); }
// planar [channels][fftsize] -> processSynchronBlock
int jd_spec_process_block(void* s, const float* planar, int channels, int fftsize)
{
    std::vector<std::vector<float>> data(static_cast<size_t>(channels));
    for (int c = 0; c < channels; ++c) data[size_t(c)].assign(planar + size_t(c) * fftsize, planar + size_t(c + 1) * fftsize);
    juce::MidiBuffer midi;
    return static_cast<Spectrogram*>(s)->processSynchronBlock(data, midi);
}
// arbitrary host block through the re-blocker (the plugin's processBlock path, PluginProcessor.cpp:148)
int jd_spec_prepare(void* s, int channels, int max_block)
{
    static_cast<Spectrogram*>(s)->preparetoProcess(channels, max_block);
    return 0;
}
int jd_spec_process_audio(void* s, const float* planar, int channels, int nsamples)
{
    std::vector<const float*> ptr(static_cast<size_t>(channels));
    for (int c = 0; c < channels; ++c) ptr[size_t(c)] = planar + size_t(c) * nsamples;
    juce::MidiBuffer midi;
    return static_cast<Spectrogram*>(s)->processBlock(ptr.data(), channels, nsamples, midi);
}
// mem [w][bins] row-major
int jd_spec_get_mem(void* s, float* mem, int w, int* pos)
{
    Spectrogram* sp = static_cast<Spectrogram*>(s);
    const int B = sp->getSpectrumSize();
    std::vector<std::vector<float>> m(static_cast<size_t>(w), std::vector<float>(size_t(B)));
    for (int i = 0; i < w; ++i) std::memcpy(m[size_t(i)].data(), mem + size_t(i) * B, size_t(B) * 4);
    int p = 0;
    const int r = sp->getMem(m, p);
    if (r >= 0)
        for (int i = 0; i < w; ++i) std::memcpy(mem + size_t(i) * B, m[size_t(i)].data(), size_t(B) * 4);
    if (pos) *pos = p;
    return r;
}
}
