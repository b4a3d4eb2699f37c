// latency_probe.cpp -- the real-time path as the C++ host of the plugin would drive it: jade_push_samples of one 512-sample
// stereo block followed by jade_fetch_columns, timed with std::chrono around both calls, through the C ABI only.
// Prints one JSON object (p50 / p99 of the pair, and of each call).  Built by `make -C tools/native`; bench.py runs it for
// its `latency` leg when the binary exists (the Python wrappers add ~10 us of ctypes / numpy overhead per block).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "jade_gpu.h"

int main(int argc, char** argv)
{
    const int blocks = argc > 1 ? atoi(argv[1]) : 2000;
    const int device = argc > 2 ? atoi(argv[2]) : 0;
    jade_engine* e = nullptr;
    if (jade_create(device, &e) != 0) {
        fprintf(stderr, "jade_create failed: %s\n", jade_last_error(nullptr));
        return 2;
    }
    jade_config c;
    jade_config_default(&c);
    c.sample_rate = 48000.f;
    c.fft_size = 2048;
    c.hop = 512;
    c.frames_per_block = 1;
    c.block_stride = 512;
    c.emit_mode = JADE_EMIT_HOP; // one column per 512-sample block, as soon as its samples exist
    c.channels = 2;
    c.max_push = 512;
    if (jade_configure(e, &c) != 0 || jade_set_palette_scheme(e, 0, 256, 0) != 0 || jade_set_value_range(e, -50.f, 50.f) != 0) {
        fprintf(stderr, "configure failed: %s\n", jade_last_error(e));
        return 2;
    }
    int W = 0, R = 0, B = 0;
    int64_t total = 0;
    jade_ring_info(e, &W, &R, &B, &total);
    std::vector<float> l(512 * 64), r(512 * 64);
    for (size_t i = 0; i < l.size(); ++i) {
        l[i] = 0.3f * std::sin(0.01f * i + 1e-6f * i * i);
        r[i] = 0.1f * (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f - 0.05f;
    }
    std::vector<uint32_t> pix((size_t)4 * R);
    std::vector<double> both, tp, tf;
    long long cols = 0;
    for (int b = 0; b < blocks + 200; ++b) {
        const float* planar[2] = {l.data() + (b % 64) * 512, r.data() + (b % 64) * 512};
        int n = 0;
        int64_t first = 0;
        const auto t0 = std::chrono::steady_clock::now();
        if (jade_push_samples(e, planar, 2, 512) != 0) return 3;
        const auto t1 = std::chrono::steady_clock::now();
        if (jade_fetch_columns(e, pix.data(), nullptr, 4, &n, &first) != 0) return 3;
        const auto t2 = std::chrono::steady_clock::now();
        if (b >= 200) {
            both.push_back(std::chrono::duration<double, std::micro>(t2 - t0).count());
            tp.push_back(std::chrono::duration<double, std::micro>(t1 - t0).count());
            tf.push_back(std::chrono::duration<double, std::micro>(t2 - t1).count());
            cols += n;
        }
    }
    auto pct = [](std::vector<double>& v, double p) {
        std::sort(v.begin(), v.end());
        return v[(size_t)(p * (v.size() - 1))];
    };
    printf("{\"p50_us\": %.3f, \"p99_us\": %.3f, \"push_p50_us\": %.3f, \"fetch_p50_us\": %.3f, \"blocks\": %d, \"block_samples\": 512, "
           "\"columns\": %lld, \"checksum\": %u}\n",
           pct(both, 0.5), pct(both, 0.99), pct(tp, 0.5), pct(tf, 0.5), blocks, cols, pix[R / 2]);
    jade_destroy(e);
    return 0;
}
