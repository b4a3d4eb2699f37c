// fake_jade_gpu.cpp -- a CPU FAKE of the few C-ABI entry points include/Spectrogram.h calls.  TEST INFRASTRUCTURE ONLY:
// it exists so that the drop-in class (re-blocker, atomic column counter, m_protect, getMem) can run under ThreadSanitizer
// in the GPU-less build container (tests/test_dropin_threads.py).  It computes NO spectrogram: column j of the fake
// ring is the constant float(j), which lets the test check that no column is lost or delivered twice.  The product never
// links this file.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "jade_gpu.h"

struct jade_engine {
    std::mutex mu;
    jade_config cfg{};
    bool configured = false, paused = false;
    int W = 1, B = 1;
    long long pushed = 0, emitted = 0, fetched = 0;
    std::vector<float> ring; // [W][B]
    std::string err;
};

extern "C" {
int jade_abi_version(void) { return JADE_ABI_VERSION; }
int jade_create(int, jade_engine** out)
{
    *out = new jade_engine();
    return JADE_OK;
}
int jade_destroy(jade_engine* e)
{
    delete e;
    return JADE_OK;
}
const char* jade_last_error(jade_engine* e) { return e ? e->err.c_str() : "fake"; }
int jade_config_default(jade_config* c)
{
    std::memset(c, 0, sizeof *c);
    c->sample_rate = 48000.f;
    c->fft_size = 2048;
    c->hop = 1024;
    c->frames_per_block = 2;
    c->block_stride = 2048;
    c->channels = 2;
    c->memory_time_s = 10.f;
    return JADE_OK;
}
int jade_configure(jade_engine* e, const jade_config* c)
{
    std::lock_guard<std::mutex> lk(e->mu);
    e->cfg = *c;
    e->W = std::max(1, c->ring_columns);
    e->B = c->fft_size / 2 + 1;
    e->ring.assign(size_t(e->W) * e->B, -120.f);
    e->pushed = e->emitted = e->fetched = 0;
    e->configured = true;
    return JADE_OK;
}
int jade_set_pause(jade_engine* e, int on)
{
    std::lock_guard<std::mutex> lk(e->mu);
    e->paused = on != 0;
    return JADE_OK;
}
int jade_set_window(jade_engine* e, int w)
{
    std::lock_guard<std::mutex> lk(e->mu);
    e->cfg.window = w;
    return JADE_OK;
}
// whole blocks only (the drop-in class pushes fft_size samples per call): frames_per_block columns per block
int jade_push_samples(jade_engine* e, const float* const* planar, int nch, int nsamples)
{
    std::lock_guard<std::mutex> lk(e->mu);
    if (!e->configured || nch != e->cfg.channels || !planar) return JADE_ERR_ARG;
    e->pushed += nsamples;
    if (e->paused) return JADE_OK;
    for (int f = 0; f < e->cfg.frames_per_block; ++f) {
        float* col = e->ring.data() + size_t(e->emitted % e->W) * e->B;
        std::fill(col, col + e->B, float(e->emitted));
        ++e->emitted;
    }
    return JADE_OK;
}
int jade_fetch_columns(jade_engine* e, uint32_t*, float* db, int max_cols, int* ncols, int64_t* first_col)
{
    std::lock_guard<std::mutex> lk(e->mu);
    long long from = e->fetched, to = e->emitted;
    if (to - from > e->W) from = to - e->W;
    if (max_cols >= 0 && to - from > max_cols) from = to - max_cols;
    e->fetched = to;
    if (first_col) *first_col = from;
    *ncols = int(to - from);
    if (db)
        for (long long j = from; j < to; ++j)
            std::memcpy(db + size_t(j - from) * e->B, e->ring.data() + size_t(j % e->W) * e->B, size_t(e->B) * 4);
    return JADE_OK;
}
int jade_ring_info(jade_engine* e, int* w, int* r, int* b, int64_t* total)
{
    std::lock_guard<std::mutex> lk(e->mu);
    if (w) *w = e->W;
    if (r) *r = e->B;
    if (b) *b = e->B;
    if (total) *total = e->emitted;
    return JADE_OK;
}
int jade_read_ring_db(jade_engine* e, float* db)
{
    std::lock_guard<std::mutex> lk(e->mu);
    std::memcpy(db, e->ring.data(), e->ring.size() * 4);
    return JADE_OK;
}
// CColorPalette's table builder is not exercised by the thread test
int jade_palette_build(int, int n, int, int32_t* t)
{
    for (int i = 0; i < n; ++i) t[i] = 0;
    return JADE_OK;
}
}
