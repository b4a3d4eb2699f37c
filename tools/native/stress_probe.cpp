// stress_probe.cpp -- the real-time path under contention, through the C ABI only: the "audio thread" pushes 512-sample
// stereo blocks at a fixed cadence while the "GUI thread" fetches the new pixel columns and keeps changing the value
// range, the palette and the window-independent display state (jade_set_value_range every 8th tick, jade_recolor_ring
// every 32nd, jade_set_palette_scheme every 64th) -- what a slider drag on the plugin's message thread does
// (Spectrogram.cpp:376-400,608-617).  Checks
//   * the latency of jade_push_samples alone (p50 / p99 / max): the audio thread must not wait for the GUI thread's GPU work;
//   * every fetched column is complete (no zero pixel) and the columns arrive in order, none lost;
//   * the dB ring at the end is bit-identical to the one an undisturbed, single-threaded engine computes from the same
//     blocks (dB values do not depend on palette or range).
// Prints one JSON object.  Built by `make -C tools/native`; run by tests/test_gpu_threads.py.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "jade_gpu.h"

static jade_engine* make_engine(int device)
{
    jade_engine* e = nullptr;
    if (jade_create(device, &e) != 0) return nullptr;
    jade_config c;
    jade_config_default(&c);
    c.sample_rate = 48000.f;
    c.fft_size = 2048;
    c.hop = 512;
    c.frames_per_block = 1;
    c.block_stride = 512;
    c.emit_mode = JADE_EMIT_HOP;
    c.channels = 2;
    c.max_push = 512;
    c.ring_columns = 4096; // longer than the run: the final ring holds every column
    if (jade_configure(e, &c) != 0 || jade_set_palette_scheme(e, JADE_PAL_JADE, 256, 0) != 0 || jade_set_value_range(e, -50.f, 50.f) != 0) {
        fprintf(stderr, "configure failed: %s\n", jade_last_error(e));
        jade_destroy(e);
        return nullptr;
    }
    return e;
}

int main(int argc, char** argv)
{
    const int blocks = std::min(argc > 1 ? atoi(argv[1]) : 3000, 4000);
    const int cadence_us = argc > 2 ? atoi(argv[2]) : 150;
    const int device = argc > 3 ? atoi(argv[3]) : 0;
    jade_engine* e = make_engine(device);
    jade_engine* ref = make_engine(device);
    if (!e || !ref) {
        fprintf(stderr, "jade_create failed: %s\n", jade_last_error(nullptr));
        return 2;
    }
    int W = 0, R = 0, B = 0;
    jade_ring_info(e, &W, &R, &B, nullptr);
    std::vector<float> l(512 * 64), r(512 * 64);
    for (size_t i = 0; i < l.size(); ++i) {
        l[i] = 0.3f * std::sin(0.01f * i + 1e-6f * i * i);
        r[i] = 0.1f * (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f - 0.05f;
    }
    auto block_ptrs = [&](int b, const float** p) {
        p[0] = l.data() + (b % 64) * 512;
        p[1] = r.data() + (b % 64) * 512;
    };

    std::atomic<bool> done{false};
    std::atomic<int> gui_errors{0};
    std::vector<double> push_us;
    push_us.reserve(blocks);
    long long incomplete = 0, fetched_cols = 0, out_of_order = 0, gui_ticks = 0;
    std::thread gui([&] {
        std::vector<uint32_t> pix((size_t)64 * R), ring((size_t)W * R);
        int64_t expect = 0;
        for (int tick = 1;; ++tick) {
            const bool last = done.load();
            int n = 0;
            int64_t first = 0;
            if (jade_fetch_columns(e, pix.data(), nullptr, 64, &n, &first) != 0) gui_errors++;
            if (n > 0 && first != expect) ++out_of_order;
            expect = first + n;
            fetched_cols += n;
            for (int c = 0; c < n; ++c)
                for (int k = 0; k < R; ++k)
                    if ((pix[(size_t)c * R + k] >> 24) != 0xFFu) {
                        ++incomplete;
                        break;
                    }
            if (tick % 8 == 0 && jade_set_value_range(e, -50.f - float(tick % 5), 50.f - float(tick % 7)) != 0) gui_errors++;
            if (tick % 32 == 0 && jade_recolor_ring(e, ring.data()) != 0) gui_errors++;
            if (tick % 64 == 0 && jade_set_palette_scheme(e, (tick / 64) % 7, 256, 0) != 0) gui_errors++;
            ++gui_ticks;
            if (last) break;
        }
    });
    const auto t_start = std::chrono::steady_clock::now();
    for (int b = 0; b < blocks; ++b) {
        const float* p[2];
        block_ptrs(b, p);
        const auto t0 = std::chrono::steady_clock::now();
        if (jade_push_samples(e, p, 2, 512) != 0) return 3;
        const auto t1 = std::chrono::steady_clock::now();
        push_us.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
        while (std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() < cadence_us) {
        }
    }
    const double wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    done.store(true);
    gui.join();

    // the same blocks through an undisturbed engine
    for (int b = 0; b < blocks; ++b) {
        const float* p[2];
        block_ptrs(b, p);
        if (jade_push_samples(ref, p, 2, 512) != 0) return 3;
    }
    std::vector<float> db_a((size_t)W * B), db_b((size_t)W * B);
    if (jade_read_ring_db(e, db_a.data()) != 0 || jade_read_ring_db(ref, db_b.data()) != 0) return 3;
    int64_t total_a = 0, total_b = 0;
    jade_ring_info(e, nullptr, nullptr, nullptr, &total_a);
    jade_ring_info(ref, nullptr, nullptr, nullptr, &total_b);
    const bool identical = total_a == total_b && std::memcmp(db_a.data(), db_b.data(), db_a.size() * 4) == 0;

    std::vector<double> v(push_us.begin() + std::min<size_t>(200, push_us.size() / 4), push_us.end());
    std::sort(v.begin(), v.end());
    auto pct = [&](double q) { return v[(size_t)(q * (v.size() - 1))]; };
    printf("{\"blocks\": %d, \"cadence_us\": %d, \"wall_s\": %.3f, \"push_p50_us\": %.3f, \"push_p99_us\": %.3f, \"push_max_us\": %.3f, "
           "\"gui_ticks\": %lld, \"columns\": %lld, \"fetched_columns\": %lld, \"incomplete_columns\": %lld, \"out_of_order\": %lld, "
           "\"gui_errors\": %d, \"db_ring_identical\": %s}\n",
           blocks, cadence_us, wall_s, pct(0.5), pct(0.99), v.back(), gui_ticks, (long long)total_a, fetched_cols, incomplete,
           out_of_order, gui_errors.load(), identical ? "true" : "false");
    jade_destroy(e);
    jade_destroy(ref);
    return 0;
}
