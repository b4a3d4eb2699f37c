/*
 * ref_glue.cpp -- C entry points onto the REFERENCE's own classes, compiled together with the reference
 * sources where they lie under /root/reference (never copied).  TEST INFRASTRUCTURE ONLY.
 * Output: oracle/_ref/libjade_ref.so (git-ignored; travels to the GPU box as a built file).
 *
 *   jr_pal_*   -> CColorPalette         (CColorpalette.h:20-48, CColorpalette.cpp)      the real code, no stand-ins
 *   jr_spec_*  -> Spectrogram           (Spectrogram.h:81-169, Spectrogram.cpp:16-331)   real code; JUCE/TGM types are
 *                                        the inert stubs of oracle/shim/, and the absent TGM FFT (`spectrum`) is the
 *                                        oracle's float32 stand-in (oracle/shim/FFT.h) -- FFT parity stays unpinned
 *   jr_view_*  -> SpectrogramComponent  (Spectrogram.cpp:333-431,590-731)                real timerCallback pixel loops
 *                                        writing into the stub Image (a plain ARGB32 array)
 *   jr_bench_* -> the reference classes timed on host threads (bench.py cpu_baseline / --impl reference)
 */
#define private public
#define protected public
#include "JadeLookAndFeel.h"
#include "Spectrogram.h"
#include "PluginEditor.h"
#undef private
#undef protected

#include <chrono>
#include <cstring>
#include <thread>

/* SpectrogramComponent is defined in the reference's Spectrogram.h only up to its declaration; its member functions
 * live in Spectrogram.cpp which is compiled next to this file. */

namespace {
/* A never-constructed editor/processor pair: timerCallback only calls m_editor.getRunningStatus() (Spectrogram.cpp:728),
 * which reads JadeSpectrogramAudioProcessor::isRunning through the editor's m_processorRef.  Both objects are zeroed
 * storage; every pointer-sized slot of the editor storage is aimed at the processor storage so that the reference
 * member, wherever the compiler placed it, refers to valid (all-zero => not running) memory. */
struct FakeHost {
    alignas(16) unsigned char proc[sizeof(JadeSpectrogramAudioProcessor)];
    alignas(16) unsigned char ed[sizeof(JadeSpectrogramAudioProcessorEditor)];
    FakeHost()
    {
        std::memset(proc, 0, sizeof proc);
        void* p = proc;
        for (size_t off = 0; off + sizeof(void*) <= sizeof ed; off += sizeof(void*)) std::memcpy(ed + off, &p, sizeof p);
    }
    JadeSpectrogramAudioProcessorEditor& editor() { return *reinterpret_cast<JadeSpectrogramAudioProcessorEditor*>(ed); }
};

struct RefView {
    FakeHost host;
    AudioProcessorValueTreeState vts;
    SpectrogramComponent comp;
    RefView(Spectrogram& s) : comp(vts, s, host.editor())
    {
        comp.m_DisplayMinColorSlider.setValue(g_minColorVal);
        comp.m_DisplayMaxColorSlider.setValue(g_maxColorVal);
    }
};

Spectrogram::FeedPercentage feed_of(int f)
{
    switch (f) {
    case 0: return Spectrogram::FeedPercentage::perc100;
    case 1: return Spectrogram::FeedPercentage::perc50;
    case 2: return Spectrogram::FeedPercentage::perc25;
    default: return Spectrogram::FeedPercentage::perc10;
    }
}
} // namespace

extern "C" {
/* ---------------- CColorPalette ---------------- */
void* jr_pal_create(int n, int scheme) { return new CColorPalette(n, scheme); }
void* jr_pal_create_default(void) { return new CColorPalette(); }
void jr_pal_destroy(void* p) { delete static_cast<CColorPalette*>(p); }
void jr_pal_set_value_range(void* p, float a, float b) { static_cast<CColorPalette*>(p)->setValueRange(a, b); }
void jr_pal_set_nr_of_colors(void* p, int n) { static_cast<CColorPalette*>(p)->setNrOfColors(n); }
void jr_pal_set_color_scheme(void* p, int s) { static_cast<CColorPalette*>(p)->setColorSceme(s); }
void jr_pal_set_invert(void* p, int on) { static_cast<CColorPalette*>(p)->setInvertStatus(on != 0); }
int jr_pal_get_rgb(void* p, float v) { return static_cast<CColorPalette*>(p)->getRGBColor(v); }
float jr_pal_get_value(void* p, int c) { return static_cast<CColorPalette*>(p)->getValue(c); }
void jr_pal_lookup_many(void* p, const float* v, int n, int* out)
{
    CColorPalette* c = static_cast<CColorPalette*>(p);
    for (int i = 0; i < n; ++i) out[i] = c->getRGBColor(v[i]);
}

/* ---------------- Spectrogram (the real class) ---------------- */
void* jr_spec_create(void) { return new Spectrogram(); }
void jr_spec_destroy(void* s) { delete static_cast<Spectrogram*>(s); }
void jr_spec_set_samplerate(void* s, float fs) { static_cast<Spectrogram*>(s)->setSamplerate(fs); }
void jr_spec_set_channels(void* s, size_t n) { static_cast<Spectrogram*>(s)->setchannels(n); }
void jr_spec_set_fftsize(void* s, size_t n) { static_cast<Spectrogram*>(s)->setFFTSize(n); }
void jr_spec_set_closest_fftsize_ms(void* s, float ms) { static_cast<Spectrogram*>(s)->setclosestFFTSize_ms(ms); }
void jr_spec_set_memory_time_s(void* s, float t) { static_cast<Spectrogram*>(s)->setmemoryTime_s(t); }
void jr_spec_set_feed_percent(void* s, int f) { static_cast<Spectrogram*>(s)->setfeed_percent(feed_of(f)); }
void jr_spec_set_pause(void* s, int on) { static_cast<Spectrogram*>(s)->setPauseMode(on != 0); }
void jr_spec_set_window(void* s, int w) { static_cast<Spectrogram*>(s)->setWindow(static_cast<Spectrogram::Windows>(w)); }
/* the reference has no setter: m_mode is fixed to AbsMean in the constructor (Spectrogram.cpp:21); poked directly so
 * that the other four branches of the mix switch (Spectrogram.cpp:64-106) can be executed */
void jr_spec_set_mix_mode(void* s, int m) { static_cast<Spectrogram*>(s)->m_mode = static_cast<Spectrogram::ChannelMixMode>(m); }
size_t jr_spec_next_pow2(void* s, float ms) { return static_cast<Spectrogram*>(s)->getnextpowerof2(ms); }
int jr_spec_spectrum_size(void* s) { return static_cast<Spectrogram*>(s)->getSpectrumSize(); }
int jr_spec_memory_size(void* s) { return static_cast<Spectrogram*>(s)->getMemorySize(); }
int jr_spec_feed_samples(void* s) { return static_cast<Spectrogram*>(s)->m_feed_samples; }
int jr_spec_feed_blocks(void* s) { return static_cast<Spectrogram*>(s)->m_feedblocks; }
float jr_spec_samplerate(void* s) { return static_cast<Spectrogram*>(s)->getSamplerate(); }
int jr_spec_window(void* s, float* out, int n)
{
    Spectrogram* sp = static_cast<Spectrogram*>(s);
    if (int(sp->m_window.size()) != n) return -1;
    std::memcpy(out, sp->m_window.data(), size_t(n) * sizeof(float));
    return 0;
}
/* planar [channels][fftsize] */
int jr_spec_process_block(void* s, const float* planar)
{
    Spectrogram* sp = static_cast<Spectrogram*>(s);
    const size_t C = sp->m_channels, N = sp->m_fftsize;
    std::vector<std::vector<float>> data(C, std::vector<float>(N));
    for (size_t c = 0; c < C; ++c) std::memcpy(data[c].data(), planar + c * N, N * sizeof(float));
    juce::MidiBuffer midi;
    return sp->processSynchronBlock(data, midi);
}
/* mem is [w][B] row-major */
int jr_spec_get_mem(void* s, float* mem, int w, int* pos)
{
    Spectrogram* sp = static_cast<Spectrogram*>(s);
    const int B = sp->getSpectrumSize();
    std::vector<std::vector<float>> m;
    m.assign(size_t(w), std::vector<float>(size_t(B)));
    for (int i = 0; i < w; ++i) std::memcpy(m[size_t(i)].data(), mem + size_t(i) * B, size_t(B) * sizeof(float));
    int p = 0;
    const int r = sp->getMem(m, p);
    for (int i = 0; i < w; ++i) std::memcpy(mem + size_t(i) * B, m[size_t(i)].data(), size_t(B) * sizeof(float));
    if (pos) *pos = p;
    return r;
}

/* ---------------- SpectrogramComponent (the real timerCallback) ---------------- */
void* jr_view_create(void* spec) { return new RefView(*static_cast<Spectrogram*>(spec)); }
void jr_view_destroy(void* v) { delete static_cast<RefView*>(v); }
void jr_view_set_running(void* v, int running) { static_cast<RefView*>(v)->comp.m_isRunningDisplay = running != 0; }
void jr_view_set_color_range(void* v, float mn, float mx)
{
    RefView* r = static_cast<RefView*>(v);
    r->comp.m_DisplayMinColorSlider.setValue(mn);
    r->comp.m_DisplayMaxColorSlider.setValue(mx);
    r->comp.m_recomputeAll = true; /* the sliders' onValueChange lambdas do this (Spectrogram.cpp:370,379) */
}
void jr_view_set_scheme(void* v, int idx)
{
    RefView* r = static_cast<RefView*>(v);
    r->comp.m_colorScheme.setSelectedItemIndex(idx, dontSendNotification);
    if (r->comp.m_colorScheme.onChange) r->comp.m_colorScheme.onChange(); /* Spectrogram.cpp:400 */
}
void jr_view_force_recompute(void* v) { static_cast<RefView*>(v)->comp.m_recomputeAll = true; }
void jr_view_tick(void* v) { static_cast<RefView*>(v)->comp.timerCallback(); }
int jr_view_width(void* v) { return static_cast<RefView*>(v)->comp.m_internalImg.w; }
int jr_view_height(void* v) { return static_cast<RefView*>(v)->comp.m_internalImg.h; }
const uint32_t* jr_view_pixels(void* v) { return static_cast<RefView*>(v)->comp.m_internalImg.px.data(); }

/* The real SpectrogramComponent::paint (Spectrogram.cpp:432-545) against the recording Graphics stub.  Inputs: component size,
 * scale factor and the two frequency sliders (log Hz).  Outputs: the source rectangle of the spectrogram blit (hStart,
 * heightInterval: the crop maths :441-464), the 11 + 11 text boxes of the two axes (value parsed from the label the
 * reference produced with the stub String, and the box's y) and the colourbar pixels.  Returns the colourbar height. */
int jr_view_paint(void* v, int width, int height, float scale, float log_min_hz, float log_max_hz, int* crop /*[2]*/,
                  float* tick_val /*[22]*/, int* tick_y /*[22]*/, uint32_t* colorbar, int colorbar_cap)
{
    RefView* r = static_cast<RefView*>(v);
    r->comp.stubWidth = width;
    r->comp.stubHeight = height;
    r->comp.setScaleFactor(scale);
    r->comp.m_DisplayMinFreqSlider.setValue(log_min_hz, dontSendNotification);
    r->comp.m_DisplayMaxFreqSlider.setValue(log_max_hz, dontSendNotification);
    Graphics g;
    r->comp.paint(g);
    if (g.images.size() < 2 || g.texts.size() < 22) return -1;
    crop[0] = g.images[0].sy;
    crop[1] = g.images[0].sh;
    for (int i = 0; i < 22; ++i) {
        std::string t = g.texts[size_t(i)].text;
        float val = float(std::atof(t.c_str()));
        if (!t.empty() && t.back() == 'k') val *= 1000.f;
        tick_val[i] = val;
        tick_y[i] = g.texts[size_t(i)].y;
    }
    const Graphics::ImageCall& cb = g.images[1];
    const int n = cb.ih;
    for (int i = 0; i < n && i < colorbar_cap; ++i) colorbar[i] = cb.px[size_t(i)];
    return n;
}
float jr_view_min_display_freq(void* v) { return static_cast<RefView*>(v)->comp.m_minDisplayFreq; }
float jr_view_max_display_freq(void* v) { return static_cast<RefView*>(v)->comp.m_maxDisplayFreq; }

/* ---------------- timing: the reference's own classes on host threads ---------------- */
/* Each thread owns one Spectrogram + one CColorPalette, configured in the plugin's order (PluginProcessor.cpp:108-112),
 * and runs its share of `nstreams` streams: processSynchronBlock per fft_size block, getMem, and per new column the
 * pixel loop body of timerCallback (getRGBColor | 0xFF000000, Spectrogram.cpp:634-637).  samples planar
 * [channels][nsamples].  Returns frames (columns) per second; *frames_out = columns produced. */
double jr_bench_batch(float fs, int fft_size, int feed, int window, int channels, int scheme, int ncolors, float mn, float mx,
                      const float* samples, long nsamples, int nstreams, int nthreads, long* frames_out)
{
    if (nthreads < 1) nthreads = 1;
    std::vector<long> frames;
    frames.assign(size_t(nthreads), 0);
    std::vector<unsigned> sink;
    sink.assign(size_t(nthreads), 0u);
    auto work = [&](int t) {
        Spectrogram sp;
        sp.setSamplerate(fs);
        sp.setmemoryTime_s(10.0f);
        sp.setchannels(size_t(channels));
        sp.setFFTSize(size_t(fft_size));
        sp.setfeed_percent(feed_of(feed));
        sp.setWindow(static_cast<Spectrogram::Windows>(window));
        CColorPalette pal(ncolors, scheme);
        pal.setValueRange(mn, mx);
        const int W = sp.getMemorySize(), B = sp.getSpectrumSize();
        std::vector<std::vector<float>> mem;
        mem.assign(size_t(W), std::vector<float>(size_t(B)));
        std::vector<std::vector<float>> block;
        block.assign(size_t(channels), std::vector<float>(size_t(fft_size)));
        std::vector<uint32_t> column;
        column.assign(size_t(B), 0u);
        juce::MidiBuffer midi;
        int pos = 0;
        sp.getMem(mem, pos); /* clears the "everything is new" state */
        unsigned acc = 0;
        long nf = 0;
        for (int s = t; s < nstreams; s += nthreads) {
            for (long b = 0; (b + 1) * fft_size <= nsamples; ++b) {
                for (int c = 0; c < channels; ++c)
                    std::memcpy(block[size_t(c)].data(), samples + size_t(c) * nsamples + b * fft_size, size_t(fft_size) * sizeof(float));
                sp.processSynchronBlock(block, midi);
                int nv = sp.getMem(mem, pos);
                if (nv > W) nv = W;
                int rd = pos - nv;
                for (int i = 0; i < nv; ++i, ++rd) {
                    const std::vector<float>& col = mem[size_t(rd < 0 ? rd + W : rd)];
                    for (int hh = 0; hh < B; ++hh) column[size_t(B - 1 - hh)] = uint32_t(pal.getRGBColor(col[size_t(hh)])) | 0xFF000000u;
                    acc += column[size_t(B / 2)];
                }
                nf += nv;
            }
        }
        frames[size_t(t)] = nf;
        sink[size_t(t)] = acc;
    };
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < nthreads; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    long total = 0;
    for (long f : frames) total += f;
    if (frames_out) *frames_out = total;
    return dt > 0 ? double(total) / dt : 0.0;
}
}
