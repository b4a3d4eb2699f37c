/*
 * jade_gpu.h -- C ABI of the B200 spectrogram engine (libjade_gpu.so).
 *
 * This is the drop-in boundary for the hot path of JoergBitzer/JadeSpectrogram: the entry points are what the
 * reference's C++ classes bind when their CPU loops are replaced (the JUCE-free drop-in classes in
 * include/Spectrogram.h and include/CColorpalette.h are thin wrappers over exactly these calls).
 * extern "C", plain pointers and sizes, opaque handle, caller-owned buffers, int status (0 ok, <0 error),
 * no exceptions cross the boundary.  File:line citations are relative to /root/reference.
 *
 *   reference interface                                         replaced by
 *   ----------------------------------------------------------  -------------------------------------------
 *   Spectrogram::Spectrogram / buildmem  (Spectrogram.cpp:16,213) jade_create + jade_configure
 *   setSamplerate/setchannels/setFFTSize/setmemoryTime_s/
 *     setfeed_percent               (Spectrogram.cpp:148-211)   jade_configure (jade_config fields)
 *   setWindow / setWindowFkt        (Spectrogram.h:123, .cpp:239) jade_configure.window, jade_get_window
 *   setPauseMode                    (Spectrogram.h:122)          jade_set_pause
 *   processSynchronBlock            (Spectrogram.cpp:37-135)     jade_push_samples
 *   getMem / getMemorySize / getSpectrumSize (.cpp:295-331)      jade_fetch_columns, jade_ring_info
 *   CColorPalette tables            (CColorpalette.cpp:100-339)  jade_palette_build, jade_set_palette
 *   CColorPalette::setValueRange    (CColorpalette.cpp:39-54)    jade_set_value_range
 *   getRGBColor | 0xFF000000 pixel loops (Spectrogram.cpp:623-724) fused into the kernels; jade_recolor_ring
 *   timerCallback image assembly    (Spectrogram.cpp:590-724)    jade_view_*
 *   paint() crop maths              (Spectrogram.cpp:441-459)    jade_linear_crop, jade_config.row_map
 *   paint() axis ticks / colourbar  (Spectrogram.cpp:466-545)    jade_freq_axis_ticks, jade_color_axis_ticks, jade_colorbar
 *   (offline batch rendering, north star)                        jade_render_batch / jade_render_device
 */
#ifndef JADE_GPU_H
#define JADE_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JADE_ABI_VERSION 1

/* status codes */
enum {
    JADE_OK = 0,
    JADE_ERR_ARG = -1,     /* bad argument / shape mismatch (the reference returns -1, Spectrogram.cpp:297-298) */
    JADE_ERR_CUDA = -2,    /* a CUDA call failed; see jade_last_error */
    JADE_ERR_STATE = -3,   /* engine not configured */
    JADE_ERR_NOGPU = -4    /* no usable CUDA device -- there is NO CPU fallback */
};

/* Spectrogram::ChannelMixMode (Spectrogram.h:84-91) */
enum { JADE_MIX_ABSMEAN = 0, JADE_MIX_MAX, JADE_MIX_MIN, JADE_MIX_LEFT, JADE_MIX_RIGHT };
/* Spectrogram::Windows (Spectrogram.h:92-100) */
enum { JADE_WIN_RECT = 0, JADE_WIN_HANN, JADE_WIN_HAMMING, JADE_WIN_BLACKMANHARRIS, JADE_WIN_FLATTOP, JADE_WIN_HANNPOISSON };
/* CColorPalette schemes (CColorpalette.h:9-18) */
enum { JADE_PAL_MONO = 0, JADE_PAL_BW, JADE_PAL_HOT, JADE_PAL_RAINBOW, JADE_PAL_VIRIDIS, JADE_PAL_PLASMA, JADE_PAL_JADE };
/* bin -> row maps */
enum {
    JADE_ROWS_IDENTITY = 0,    /* one row per bin, the reference (Spectrogram.cpp:592-599,642) */
    JADE_ROWS_LINEAR_CROP = 1, /* bins [k_lo,k_hi) from the reference's paint() crop maths (Spectrogram.cpp:441-459) */
    JADE_ROWS_LOG_MAXPOOL = 2  /* extension (no reference): `rows` log-spaced bands fmin..fmax, max power per band */
};
enum { JADE_PIX_ARGB32 = 0 /* 0xAARRGGBB word, the reference's juce::Colour(uint32) */, JADE_PIX_RGBA8 = 1 /* bytes R,G,B,A */ };
enum { JADE_EMIT_HOP = 0 /* a column as soon as its samples exist */, JADE_EMIT_BLOCK = 1 /* per full block (reference) */ };
enum { JADE_SYNTH_SWEEP = 0, JADE_SYNTH_NOISE = 1, JADE_SYNTH_MIX = 2 };

typedef struct jade_engine jade_engine;

typedef struct jade_config {
    float sample_rate;        /* m_fs */
    int32_t fft_size;         /* m_fftsize, power of two in [64, 65536] */
    int32_t hop;              /* m_feed_samples: samples between sub-frames inside a block */
    int32_t frames_per_block; /* m_feedblocks (>=1) */
    int32_t block_stride;     /* samples between blocks; 0 -> hop*frames_per_block (reference: fft_size) */
    int32_t preroll;          /* zeros assumed in front of the first sample; <0 -> fft_size (reference, .cpp:233) */
    int32_t emit_mode;        /* JADE_EMIT_* */
    int32_t window;           /* JADE_WIN_* */
    int32_t channels;         /* m_channels, 1..16 */
    int32_t mix_mode;         /* JADE_MIX_* */
    int32_t row_map;          /* JADE_ROWS_* */
    int32_t rows;             /* LOG_MAXPOOL: number of rows; otherwise derived */
    float fmin, fmax;         /* Hz, for LINEAR_CROP / LOG_MAXPOOL */
    int32_t flip_y;           /* 1: row 0 = highest frequency (reference image orientation) */
    int32_t pixel_format;     /* JADE_PIX_* */
    float power_scale;        /* multiplies |X|^2 (FFT normalisation knob; 1 = unnormalised DFT) */
    float memory_time_s;      /* m_memsize_s: ring length in seconds */
    int32_t ring_columns;     /* >0 overrides memory_time_s */
    int32_t db_precise;       /* 1: dB through double log10 exactly like the reference; 0: hardware log2 */
    int32_t max_push;         /* largest nsamples of one jade_push_samples call (0 -> fft_size) */
} jade_config;

/* ---- life cycle ---- */
int jade_abi_version(void);
int jade_device_count(void);
int jade_create(int device, jade_engine** out);
int jade_destroy(jade_engine* e);
const char* jade_last_error(jade_engine* e); /* e may be NULL: last error of the calling thread */

/* ---- configuration ---- */
/* The plugin's live defaults (PluginProcessor.cpp:102-114): 48 kHz, N=2048, 50 % feed, Hann, 2 ch, AbsMean, 10 s. */
int jade_config_default(jade_config* c);
/* Reference feed percentages (Spectrogram.cpp:189-216): fills hop, frames_per_block, block_stride for 100/50/25/10. */
int jade_config_set_feed_percent(jade_config* c, int percent);
int jade_configure(jade_engine* e, const jade_config* c); /* (re)allocates: the reference's buildmem() */
int jade_get_config(jade_engine* e, jade_config* out);    /* resolved values */
int jade_set_pause(jade_engine* e, int on);
int jade_set_window(jade_engine* e, int window);          /* like setWindow: table only, no buffer reset */
int jade_get_window(jade_engine* e, float* out, int n);   /* the unit-RMS window table (Spectrogram.cpp:239-293) */
/* the same table without an engine (host-only, works without a GPU): Spectrogram::setWindowFkt for `window`, length n */
int jade_window_build(int window, int n, float* out);
int jade_reset(jade_engine* e);                           /* buildmem() without reconfiguration */

/* ---- palette ---- */
/* Builds the reference colour table (0x00RRGGBB). `table` must hold n ints and is updated IN PLACE with the
 * reference's write order, so stale entries survive exactly like m_Color does (kMono + invert quirk). */
int jade_palette_build(int scheme, int n, int invert, int32_t* table);
/* n in [1, 65536].  Transactional: every kernel keeps the table in shared memory, so a long table makes the engine fall back
 * to kernels with a smaller footprint (default stereo N = 2048 configuration: the two-channel kernel up to 1384 colours, the
 * per-channel ones up to ~27 000); if nothing fits the call fails with JADE_ERR_ARG and the PREVIOUS table, range and
 * kernel choice stay in force.  The new table is uploaded in stream order: columns already pushed keep the old colours. */
int jade_set_palette(jade_engine* e, const int32_t* rgb, int n);
int jade_set_palette_scheme(jade_engine* e, int scheme, int n, int invert);
int jade_set_value_range(jade_engine* e, float min_db, float max_db); /* swap / equal rules of the reference */
int jade_get_value_range(jade_engine* e, float* mn, float* mx, float* mult);
/* host-side scalar lookup with the engine's current table/range (CColorPalette::getRGBColor) */
int jade_lookup_color(jade_engine* e, float value_db, int32_t* rgb);
/* crop maths of SpectrogramComponent::paint (Spectrogram.cpp:441-459): bins [k_lo,k_hi) shown for fmin..fmax */
int jade_linear_crop(float fs, int bins, float fmin, float fmax, int* k_lo, int* k_hi);
/* ---- display arithmetic of SpectrogramComponent::paint (Spectrogram.cpp:432-545); host-only, no GPU needed ---- */
/* slider clamp rules (:441-453): min >= fs/2 -> 0.9 fs/2; max >= fs/2 -> fs/2; min >= max -> 0.9 max */
int jade_display_freq_clamp(float fs, float* min_hz, float* max_hz);
typedef struct jade_axis_tick {
    float value;    /* the rounded tick value the reference prints (Hz / dB) */
    int32_t y;      /* top of its label box in component pixels */
    char label[16]; /* text; the reference formats with juce::String(float), so only `value` is pinned */
} jade_axis_tick;
/* frequency axis (:466-503): nticks values from min_hz to max_hz, rounded to 1 / 10 / 100 Hz; comp_height = component
 * height, scale = m_scaleFactor, menu_height = g_menuHeight (20), text_height = 20 in the reference */
int jade_freq_axis_ticks(float min_hz, float max_hz, int comp_height, float scale, int menu_height, int text_height, int nticks,
                         jade_axis_tick* out);
/* colourbar axis (:524-545): tick values truncated to multiples of 10 between min_val and max_val (g_minColorVal / g_maxColorVal) */
int jade_color_axis_ticks(float min_val, float max_val, int comp_height, float scale, int menu_height, int text_height, int nticks,
                          jade_axis_tick* out);
/* colourbar height in pixels (:510): int(comp_height - scale * menu_height) */
int jade_colorbar_height(int comp_height, float scale, int menu_height);
/* The colourbar itself (:511-521): out[y], y = 0 at the top, is the engine's current palette / value range applied to
 * val = float(kk) / height * (ramp_max - ramp_min) + ramp_min at kk = height - 1 - y, in the engine's pixel format. */
int jade_colorbar(jade_engine* e, int height, float ramp_min, float ramp_max, uint32_t* out);

/* log max-pool band table (extension): lo/hi hold `rows` entries */
int jade_log_rows(float fs, int fft_size, int rows, float fmin, float fmax, int32_t* lo, int32_t* hi);

/* ---- streaming (real-time) path ---- */
/* planar[ch] points to nsamples host floats; nch must equal config.channels.  The audio-thread call: it copies the block
 * into a pinned staging slot and launches two kernels; it never waits for GPU work (the only event it checks guards the
 * reuse of a staging slot eight pushes later) and the engine lock it takes is never held by another call across a GPU
 * synchronisation or a ring-sized copy -- jade_set_value_range / jade_set_palette / jade_set_window / jade_recolor_ring
 * from the GUI thread delay it by microseconds.  Only jade_configure / jade_reset (the reference's buildmem()) stop it. */
int jade_push_samples(jade_engine* e, const float* const* planar, int nch, int nsamples);
/* Copies the columns produced since the previous fetch (oldest first, at most max_cols -- older ones are dropped
 * like the reference's ring overwrite) into pixels[ncols][rows] and, if not NULL, db[ncols][bins].  *first_col is
 * the absolute index of the first returned column.  Returns when those columns are complete: pixel-only fetches of
 * freshly pushed columns poll the pinned, device-mapped ring (no wait for kernel retirement), everything else waits for
 * the GPU work of earlier pushes.  (Invariant behind the poll: both pixel formats bake the alpha byte 0xFF into every
 * pixel, so a zeroed slot is "not written yet"; a pixel format that can produce 0 must take the waiting path.) */
int jade_fetch_columns(jade_engine* e, uint32_t* pixels, float* db, int max_cols, int* ncols, int64_t* first_col);
int jade_ring_info(jade_engine* e, int* ring_columns, int* rows, int* bins, int64_t* total_columns);
/* Re-colour the whole ring from the stored dB values with the current palette/range (m_recomputeAll path,
 * Spectrogram.cpp:623-657).  pixels[ring_columns][rows], ring order (slot = column % ring_columns); every row map
 * (log max-pool rows take the largest dB of their band).  Columns pushed WHILE the call copies the ring may be caught
 * half-written in `pixels`; they are returned complete by the next jade_fetch_columns. */
int jade_recolor_ring(jade_engine* e, uint32_t* pixels);
/* dB ring exactly as stored (slot order), db[ring_columns][bins]; unwritten slots hold -120 (Spectrogram.cpp:223) */
int jade_read_ring_db(jade_engine* e, float* db);

/* ---- display image (SpectrogramComponent::timerCallback, Spectrogram.cpp:590-724) ----
 * Assembles the W x H ARGB32 image the component paints, with the reference's index rules: scroll mode (image shifted
 * left, new columns on the right), fixed mode (columns at their ring position, red cursor at the write position) and the
 * full redraw after a range / scheme change.  Colours come from the engine (pixel ring, jade_recolor_ring); the view only
 * places columns.  It owns the engine's fetch cursor: do not mix with jade_fetch_columns on the same engine. */
typedef struct jade_view jade_view;
int jade_view_create(jade_engine* e, jade_view** out);
int jade_view_destroy(jade_view* v);
int jade_view_set_running(jade_view* v, int running);                     /* 1: scroll (default), 0: fixed + cursor */
int jade_view_set_value_range(jade_view* v, float min_db, float max_db);  /* jade_set_value_range + full redraw */
int jade_view_invalidate(jade_view* v);                                   /* m_recomputeAll = true */
int jade_view_tick(jade_view* v, int* new_columns);                       /* one timerCallback; new_columns may be NULL */
int jade_view_image(jade_view* v, const uint32_t** pixels, int* width, int* height); /* row-major [H][W], row 0 = top */

/* ---- batch path ---- */
/* number of columns the configured geometry yields for nsamples samples per channel */
int64_t jade_columns_for(jade_engine* e, int64_t nsamples);
/* Host buffers.  samples[nstreams][channels][nsamples] -> pixels[nstreams][ncols][rows] (and db[nstreams][ncols][bins]
 * if not NULL), columns first_col .. first_col+ncols-1 of every stream.  Copies are pipelined with the kernels. */
int jade_render_batch(jade_engine* e, const float* samples, int nstreams, int64_t nsamples, int64_t first_col,
                      int64_t ncols, uint32_t* pixels, float* db);
/* Same over several engines (one per GPU): streams are range-partitioned, no inter-GPU exchange. */
int jade_render_batch_multi(jade_engine* const* engines, int nengines, const float* samples, int nstreams,
                            int64_t nsamples, int64_t first_col, int64_t ncols, uint32_t* pixels, float* db);
/* Device buffers (already resident in HBM).  stream strides in elements; cuda_stream is a cudaStream_t or NULL
 * for the engine's own stream.  Asynchronous: call jade_sync. */
int jade_render_device(jade_engine* e, const float* d_samples, int nstreams, int64_t nsamples,
                       int64_t stream_stride, int64_t channel_stride, int64_t first_col, int64_t ncols,
                       uint32_t* d_pixels, float* d_db, void* cuda_stream);
int jade_sync(jade_engine* e);
/* deterministic synthetic input written on the device (bench / tests) */
int jade_synth_device(jade_engine* e, float* d_out, int nstreams, int channels, int64_t nsamples,
                      int64_t stream_stride, int64_t channel_stride, int kind, uint64_t seed, void* cuda_stream);

/* Page-locked host memory for caller-side sample / pixel buffers: buffers allocated here are DMA'd directly by
 * jade_render_batch; any other host pointer is staged through internal pinned buffers. */
void* jade_host_alloc(size_t bytes);
int jade_host_free(void* p);

/* ---- introspection ---- */
int64_t jade_kernel_launches(jade_engine* e); /* kernels launched by this engine so far */
/* name of the kernel the current configuration dispatches interior, aligned frames to ("pkz2048", "pk2048",
 * "pksmall<T>", "pkcta<R1>", "pkcta2<16>", "warp<T>", "cta<R1>") */
const char* jade_kernel_name(jade_engine* e);
/* seconds of device time of the kernels of the most recent jade_render_device call (CUDA events on its stream); -1 before
 * the first one.  jade_render_batch overlaps copies and kernels on several streams and is not covered. */
double jade_last_kernel_seconds(jade_engine* e);

#ifdef __cplusplus
}
#endif
#endif
