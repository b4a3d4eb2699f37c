// tsan_dropin.cpp -- two-thread exercise of the drop-in Spectrogram class, built with -fsanitize=thread against the CPU fake of
// the C ABI (fake_jade_gpu.cpp).  TEST INFRASTRUCTURE ONLY.  The audio thread feeds odd-sized host blocks through the
// re-blocker (PluginProcessor.cpp:148) while the GUI thread calls getMem and the non-structural setters (phase 1: every
// column must arrive exactly once, in order) and then also the structural ones (phase 2: races only).
#include <atomic>
#include <cstdio>
#include <thread>

#include "Spectrogram.h"

int main()
{
    Spectrogram spec;
    spec.preparetoProcess(2, 512);
    spec.setSamplerate(48000.f);
    spec.setmemoryTime_s(10.f);
    spec.setFFTSize(512);
    spec.setfeed_percent(Spectrogram::FeedPercentage::perc50);
    const int W = spec.getMemorySize(), B = spec.getSpectrumSize();
    std::vector<std::vector<float>> mem(size_t(W), std::vector<float>(size_t(B), 0.f));
    int pos = -1;
    if (spec.getMem(mem, pos) < W || pos != 0) { // "everything is new" before the first block
        std::printf("FAIL: first getMem\n");
        return 1;
    }
    std::atomic<bool> stop{false};
    std::atomic<long long> blocks{0};
    auto audio = [&](int nblocks) {
        std::vector<float> l(480, 0.25f), r(480, -0.25f);
        juce::MidiBuffer midi;
        for (int i = 0; i < nblocks; ++i) {
            const int n = 64 + (i * 37) % 417; // 64..480 samples
            const float* ch[2] = {l.data(), r.data()};
            spec.processBlock(ch, 2, n, midi);
            blocks.fetch_add(1);
        }
        stop.store(true);
    };
    // ---- phase 1: strict accounting
    long long delivered = 0;
    int errors = 0;
    {
        std::thread t(audio, 40000);
        int tick = 0;
        while (!stop.load() || delivered == 0) {
            const int n = spec.getMem(mem, pos);
            if (n < 0 || n > W) { ++errors; break; }
            for (int i = 0; i < n; ++i) {
                const long long j = delivered + i;
                if (mem[size_t(j % W)][0] != float(j) || mem[size_t(j % W)][size_t(B - 1)] != float(j)) { if (errors < 5) std::printf("col %lld slot holds %g (n %d pos %d)\n", j, mem[size_t(j % W)][0], n, pos); ++errors; }
            }
            delivered += n;
            if (pos != int(delivered % W)) { if (errors < 5) std::printf("pos %d delivered %lld n %d\n", pos, delivered, n); ++errors; }
            if ((++tick & 15) == 0) spec.setWindow((tick & 16) ? Spectrogram::Windows::Hann : Spectrogram::Windows::Hamming);
            if ((tick & 1023) == 0) { spec.setPauseMode(true); spec.setPauseMode(false); }
        }
        t.join();
        const int n = spec.getMem(mem, pos);
        for (int i = 0; i < n; ++i)
            if (mem[size_t((delivered + i) % W)][0] != float(delivered + i)) ++errors;
        delivered += n;
        int64_t total = 0;
        jade_ring_info(spec.engine(), nullptr, nullptr, nullptr, &total);
        if (delivered != total || spec.getMem(mem, pos) != 0) ++errors;
        std::printf("phase 1: %lld host blocks, %lld columns emitted, %lld delivered, %d errors\n", blocks.load(), (long long)total, delivered, errors);
    }
    // ---- phase 2: structural setters against a running audio thread (ThreadSanitizer is the judge)
    {
        stop.store(false);
        std::thread t(audio, 20000);
        int k = 0;
        while (!stop.load()) {
            switch (k++ % 5) {
            case 0: spec.setFFTSize((k & 8) ? 512 : 1024); break;
            case 1: spec.setfeed_percent((k & 16) ? Spectrogram::FeedPercentage::perc25 : Spectrogram::FeedPercentage::perc50); break;
            case 2: spec.setmemoryTime_s(5.f + float(k % 3)); break;
            default: break;
            }
            const int w = spec.getMemorySize(), b = spec.getSpectrumSize();
            if ((int)mem.size() != w || (int)mem[0].size() != b) mem.assign(size_t(w), std::vector<float>(size_t(b), 0.f));
            spec.getMem(mem, pos);
        }
        t.join();
        std::printf("phase 2: %d reconfigurations\n", k);
    }
    std::printf(errors ? "FAIL\n" : "OK\n");
    return errors ? 1 : 0;
}
