// jade_view.cpp -- display-image assembly on top of the engine's public C ABI (include/jade_gpu.h).
//
// Replaces the two pixel loops of SpectrogramComponent::timerCallback (Spectrogram.cpp:590-724): every colour comes from
// the GPU (the pixel ring the STFT kernels fill, or jade_recolor_ring for a full redraw); what is left for the host is
// where a column goes in the image -- scroll mode (moveImageSection + new columns on the right, :660-681), fixed mode
// (columns at their ring position + the red cursor, :683-721) and the full redraw (m_recomputeAll, :623-657) -- a few
// memmoves and column blits per timer tick, with the reference's exact index rules.
// The image is row-major ARGB32 [H][W], x = column, y = H-1-bin (Spectrogram.cpp:642); W = ring columns, H = rows.
#include <cstdint>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/jade_gpu.h"

struct jade_view {
    jade_engine* e = nullptr;
    int W = 0, H = 0;
    std::vector<uint32_t> img;   // [H][W]
    std::vector<uint32_t> cols;  // [<=W][H] fetched / recoloured columns
    bool recompute_all = true;   // m_recomputeAll (true after construction, Spectrogram.cpp:337)
    bool running = true;         // scroll mode (m_isRunning: the "Fix" button toggles it, Spectrogram.cpp:777-790)
    int64_t seen = 0;            // columns accounted for by earlier ticks
    bool first = true;           // the reference's first getMem reports "everything is new" (Spectrogram.cpp:18,236)
    bool flip = true;            // engine rows: row 0 = highest bin (jade_config.flip_y); otherwise columns are turned over here
};

namespace {
constexpr uint32_t kRed = 0xFFFF0000u; // juce::Colours::red
// col[r] in the engine's row order; the image always has the lowest frequency at the bottom (y = H-1-hh, Spectrogram.cpp:642)
inline void put_column(jade_view* v, int x, const uint32_t* col)
{
    uint32_t* p = v->img.data() + x;
    if (v->flip)
        for (int r = 0; r < v->H; ++r) p[(size_t)r * v->W] = col[r];
    else
        for (int r = 0; r < v->H; ++r) p[(size_t)(v->H - 1 - r) * v->W] = col[r];
}
inline void red_column(jade_view* v, int x)
{
    uint32_t* p = v->img.data() + x;
    for (int r = 0; r < v->H; ++r) p[(size_t)r * v->W] = kRed;
}
} // namespace

extern "C" {

int jade_view_create(jade_engine* e, jade_view** out)
{
    if (!e || !out) return -1;
    jade_view* v = new (std::nothrow) jade_view();
    if (!v) return -2;
    v->e = e;
    *out = v;
    return 0;
}

int jade_view_destroy(jade_view* v)
{
    delete v;
    return 0;
}

int jade_view_set_running(jade_view* v, int running)
{
    if (!v) return -1;
    v->running = running != 0;
    return 0;
}

int jade_view_invalidate(jade_view* v)
{
    if (!v) return -1;
    v->recompute_all = true;
    return 0;
}

int jade_view_set_value_range(jade_view* v, float min_db, float max_db)
{
    if (!v) return -1;
    if (int r = jade_set_value_range(v->e, min_db, max_db)) return r;
    v->recompute_all = true; // the sliders' listeners (Spectrogram.cpp:376-399)
    return 0;
}

int jade_view_tick(jade_view* v, int* new_columns)
{
    if (!v) return -1;
    int W = 0, H = 0, B = 0;
    if (int r = jade_ring_info(v->e, &W, &H, &B, nullptr)) return r;
    if (W != v->W || H != v->H) { // Spectrogram.cpp:595-605: the image follows the data size
        v->W = W;
        v->H = H;
        v->img.assign((size_t)W * H, 0xFF000000u);
        v->cols.assign((size_t)W * H, 0u);
        v->recompute_all = true;
        jade_config c;
        if (int r = jade_get_config(v->e, &c)) return r;
        v->flip = c.flip_y != 0;
    }
    // ONE call both moves the fetch cursor and tells how many columns exist (first_col + n): the count and the columns
    // cannot be separated by a concurrent jade_push_samples.  Columns older than a ring are dropped by the engine.
    int n = 0;
    int64_t first_col = 0;
    if (int r = jade_fetch_columns(v->e, v->cols.data(), nullptr, W, &n, &first_col)) return r;
    const int64_t total = first_col + n;
    if (total < v->seen) { // engine was reset / reconfigured
        v->seen = 0;
        v->first = true;
    }
    // m_newEntryCounter: columns since the previous tick; "everything" on the first one
    // (the reference's counter starts at int(100000000000) = 1215752192 and keeps counting, Spectrogram.cpp:18,112)
    const int64_t new_vals = v->first ? (int64_t)1215752192 + total : total - v->seen;
    v->first = false;
    v->seen = total;
    const int pos = (int)(total % W); // ring write index (m_memCounter)
    if (new_columns) *new_columns = (int)(new_vals > 2147483647 ? 2147483647 : new_vals);
    if (new_vals > W) v->recompute_all = true; // :611

    if (v->recompute_all) { // :623-657
        v->recompute_all = false;
        if (int r = jade_recolor_ring(v->e, v->cols.data())) return r; // [slot][row], includes the columns fetched above
        const int newwstart = W - pos;
        for (int ww = 0; ww < W; ++ww) {
            int neww = ww + newwstart;
            if (neww >= W) neww -= W;
            put_column(v, v->running ? neww : ww, v->cols.data() + (size_t)ww * H);
        }
        if (!v->running) red_column(v, pos % W);
        return 0;
    }
    // :658-724 -- only the new columns (n == new_vals here: new_vals <= W and the engine keeps a full ring)
    if (v->running) {
        if (n > 0 && n < W)
            for (int y = 0; y < H; ++y) std::memmove(&v->img[(size_t)y * W], &v->img[(size_t)y * W + n], (size_t)(W - n) * 4);
        for (int i = 0; i < n; ++i) put_column(v, W - n + i, v->cols.data() + (size_t)i * H);
    } else {
        for (int i = 0; i < n; ++i) put_column(v, (int)((first_col + i) % W), v->cols.data() + (size_t)i * H);
        int drawwidth = 1;
        if (H < 2048) drawwidth++;
        if (H < 1024) drawwidth += 2;
        for (int dd = 0; dd < drawwidth; ++dd) {
            int drawpos = pos + dd;
            if (drawpos >= W) drawpos -= W; // (:715 only wraps the == W case; stay in bounds)
            red_column(v, drawpos);
        }
    }
    return 0;
}

int jade_view_image(jade_view* v, const uint32_t** pixels, int* width, int* height)
{
    if (!v) return -1;
    if (pixels) *pixels = v->img.data();
    if (width) *width = v->W;
    if (height) *height = v->H;
    return 0;
}

} // extern "C"
