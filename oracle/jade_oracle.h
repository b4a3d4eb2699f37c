/*
 * jade_oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and there
 * only as the checker / the timed CPU baseline.  The product library (libjade_gpu.so) never links,
 * loads or calls anything declared here.
 *
 * The oracle is a CPU restatement of the hot path of JoergBitzer/JadeSpectrogram:
 *   Spectrogram.cpp:16-24,36-135,137-145,148-238,239-293,295-331  (STFT core, windows, ring, getMem)
 *   Spectrogram.cpp:590-724                                       (column -> image pixel loops)
 *   CColorpalette.h:32-47, CColorpalette.cpp:1-339                (palette tables and lookup)
 *
 * PARITY STATUS
 *   palette tables / lookup : pinned  -- checked bit-exact against the reference's own CColorpalette.cpp
 *                             compiled in place (oracle/_ref/libjade_ref.so, see oracle/Makefile).
 *   windows, framing, mix, dB, ring, getMem : restated from the reference text; pinned against the real
 *                             Spectrogram.cpp compiled with a stub JUCE/TGM shim when oracle/_ref was built.
 *   FFT (spectrum::power)   : PARITY UNPINNED.  The reference's FFT lives in the author's external TGM
 *                             library ("FFT.h", class spectrum; CMakeLists.txt:65 ${TGMLIBCPPS}) which is
 *                             not in the tree and has no pinned version.  The oracle restates the published
 *                             textbook algorithm: power[k] = |X[k]|^2, k = 0..N/2, X = unnormalised forward
 *                             DFT of the windowed frame (float32 radix-2 real FFT; a float64 version bounds
 *                             the rounding error).
 */
#ifndef JADE_ORACLE_H
#define JADE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums mirror the reference (Spectrogram.h:84-107, CColorpalette.h:9-18) ---- */
enum { JO_MIX_ABSMEAN = 0, JO_MIX_MAX, JO_MIX_MIN, JO_MIX_LEFT, JO_MIX_RIGHT };
enum { JO_WIN_RECT = 0, JO_WIN_HANN, JO_WIN_HAMMING, JO_WIN_BLACKMANHARRIS, JO_WIN_FLATTOP, JO_WIN_HANNPOISSON };
enum { JO_FEED_100 = 0, JO_FEED_50, JO_FEED_25, JO_FEED_10 };
enum { JO_PAL_MONO = 0, JO_PAL_BW, JO_PAL_HOT, JO_PAL_RAINBOW, JO_PAL_VIRIDIS, JO_PAL_PLASMA, JO_PAL_JADE };

/* ---- stage functions ---- */
/* Spectrogram.cpp:239-293 setWindowFkt */
int jo_window(int window, int n, float* out);
/* stand-in for spectrum::power (Spectrogram.cpp:144): n real floats -> n/2+1 power values. */
int jo_power_f32(const float* x, int n, float* power);
int jo_power_f64(const float* x, int n, double* power);
/* Spectrogram.cpp:36,107 */
float jo_db(float power);

/* ---- restated Spectrogram (Spectrogram.h:81-169) ---- */
typedef struct jo_spec jo_spec;
jo_spec* jo_spec_create(void);
void jo_spec_destroy(jo_spec*);
void jo_spec_set_samplerate(jo_spec*, float fs);
void jo_spec_set_channels(jo_spec*, size_t nch);
void jo_spec_set_fftsize(jo_spec*, size_t n);
void jo_spec_set_closest_fftsize_ms(jo_spec*, float ms);
void jo_spec_set_memory_time_s(jo_spec*, float s);
void jo_spec_set_feed_percent(jo_spec*, int feed);
void jo_spec_set_pause(jo_spec*, int on);
void jo_spec_set_window(jo_spec*, int window);
/* extension: the reference has no setter for m_mode (fixed AbsMean, Spectrogram.cpp:21) */
void jo_spec_set_mix_mode(jo_spec*, int mode);
/* extension: use the float64 FFT (error-bound reference) instead of the float32 stand-in */
void jo_spec_set_fft_double(jo_spec*, int on);
size_t jo_spec_next_pow2(jo_spec*, float ms);
int jo_spec_spectrum_size(jo_spec*);
int jo_spec_memory_size(jo_spec*);
int jo_spec_feed_samples(jo_spec*);
int jo_spec_feed_blocks(jo_spec*);
float jo_spec_samplerate(jo_spec*);
/* planar input [channels][fftsize]; Spectrogram.cpp:37-135 */
int jo_spec_process_block(jo_spec*, const float* planar);
/* mem is [W][B] row-major, W must equal memory_size else -1; Spectrogram.cpp:295-331 */
int jo_spec_get_mem(jo_spec*, float* mem, int w, int* pos);

/* ---- restated CColorPalette (CColorpalette.h:20-48) ---- */
typedef struct jo_pal jo_pal;
jo_pal* jo_pal_create(int ncolors, int scheme);
jo_pal* jo_pal_create_default(void); /* CColorPalette() : 2 colours, kMono */
void jo_pal_destroy(jo_pal*);
void jo_pal_set_value_range(jo_pal*, float mn, float mx);
void jo_pal_set_nr_of_colors(jo_pal*, int n);
void jo_pal_set_color_scheme(jo_pal*, int scheme);
void jo_pal_set_invert(jo_pal*, int on);
int jo_pal_get_rgb(jo_pal*, float value);
float jo_pal_get_value(jo_pal*, int color);
int jo_pal_table(jo_pal*, int* out, int cap); /* returns n */
void jo_pal_get_range(jo_pal*, float* mn, float* mx, float* mult);
void jo_pal_lookup_many(jo_pal*, const float* v, int n, int* out);

/* ---- restated image assembly (Spectrogram.cpp:590-724), JUCE Image replaced by a W x H ARGB32 array ---- */
typedef struct jo_view jo_view;
jo_view* jo_view_create(jo_spec* spec, jo_pal* pal);
void jo_view_destroy(jo_view*);
void jo_view_set_running(jo_view*, int running_display); /* m_isRunningDisplay */
void jo_view_set_color_range(jo_view*, float mn, float mx);
void jo_view_force_recompute(jo_view*);
int jo_view_tick(jo_view*);                /* one timerCallback; returns newVals */
int jo_view_width(jo_view*);
int jo_view_height(jo_view*);
const uint32_t* jo_view_pixels(jo_view*);  /* row-major [H][W], 0xAARRGGBB */

/* ---- batch convenience used by parity tests and the CPU baseline ---- */
typedef struct jo_batch_cfg {
    float fs;
    int fft_size;
    int hop;             /* generalised hop (reference: only 100/50/25/10 %) */
    int window;
    int channels;
    int mix_mode;
    int palette_scheme;
    int palette_size;
    int palette_invert;
    float min_db, max_db;
    int use_double_fft;
} jo_batch_cfg;
/* Uniform-hop batch: column j analyses x[j*hop - N, j*hop) (zeros before the start), exactly the frames the
 * reference emits when N % hop == 0.  samples planar [channels][nsamples].  Writes ncols columns
 * of dB [ncols][B] (may be NULL) and ARGB32 pixels [ncols][B] with row r <-> bin B-1-r (may be NULL).
 * Returns number of columns written. */
long jo_render_batch(const jo_batch_cfg* cfg, const float* samples, long nsamples,
                     long first_col, long ncols, float* db_out, uint32_t* pix_out);
/* Times jo_render_batch-equivalent work on nthreads host threads over nstreams independent streams of
 * nsamples each (same data per stream); returns frames per second; *frames_out = frames processed. */
double jo_bench_batch(const jo_batch_cfg* cfg, const float* samples, long nsamples, int nstreams,
                      int nthreads, long* frames_out);
/* Streaming through the restated class exactly like the plugin: push fft_size blocks, then getMem.
 * Returns p50 seconds per 512-sample-equivalent block is computed by the caller; this returns total seconds
 * for nblocks processSynchronBlock calls. */
double jo_bench_stream(jo_spec* s, const float* planar_block, int nblocks);

#ifdef __cplusplus
}
#endif
#endif
