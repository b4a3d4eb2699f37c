// emu_harness.cpp -- runs the product kernels (jade_kernels.cuh compiled with -DJADE_EMU) on the CPU SIMT emulator.
// TEST INFRASTRUCTURE ONLY: used by tests/test_emu_kernels.py to debug kernel logic in the GPU-less build container.
#define JADE_EMU 1
#include "cuda_emu.h"

#include "../../jadespectrogram_b200/csrc/jade_kernels.cuh"
#include "../../jadespectrogram_b200/csrc/jade_pk.cuh"
#include "../../jadespectrogram_b200/csrc/jade_pk_cta.cuh"
#include "../../jadespectrogram_b200/csrc/jade_pk_small.cuh"
#include "../../jadespectrogram_b200/csrc/jade_pkz.cuh"
#include "../../jadespectrogram_b200/csrc/jade_pk3.cuh"
#include "../../jadespectrogram_b200/csrc/jade_pk_cluster3.cuh"
#include "../../jadespectrogram_b200/csrc/jade_pk_cluster.cuh"
#include "../../jadespectrogram_b200/csrc/jade_host_tables.h"
#include "../../include/jade_gpu.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

using jade::KParams;

namespace {
template <int T>
void run_warp(const KParams& P, int mixk, bool general, int grid)
{
    const int smem = jade::WarpCfg<T>::smem_bytes(P.npal, general);
    const int block = jade::WARP_KERNEL_WARPS * 32;
    if (mixk == jade::MIX_SEL) jade_emu::launch(jade::stft_warp_kernel<T, jade::MIX_SEL, true>, grid, block, smem, P);
    else if (mixk == jade::MIX_SUM && general) jade_emu::launch(jade::stft_warp_kernel<T, jade::MIX_SUM, true>, grid, block, smem, P);
    else if (mixk == jade::MIX_SUM) jade_emu::launch(jade::stft_warp_kernel<T, jade::MIX_SUM, false>, grid, block, smem, P);
    else if (general) jade_emu::launch(jade::stft_warp_kernel<T, jade::MIX_NONE, true>, grid, block, smem, P);
    else jade_emu::launch(jade::stft_warp_kernel<T, jade::MIX_NONE, false>, grid, block, smem, P);
}
template <int R1>
void run_cta(const KParams& P, int mixk, bool general, int grid)
{
    const int smem = jade::CtaCfg<R1>::smem_bytes(P.npal, general);
    const int block = 32 * R1;
    if (mixk == jade::MIX_SEL) jade_emu::launch(jade::stft_cta_kernel<R1, jade::MIX_SEL, true>, grid, block, smem, P);
    else if (mixk == jade::MIX_SUM && general) jade_emu::launch(jade::stft_cta_kernel<R1, jade::MIX_SUM, true>, grid, block, smem, P);
    else if (mixk == jade::MIX_SUM) jade_emu::launch(jade::stft_cta_kernel<R1, jade::MIX_SUM, false>, grid, block, smem, P);
    else if (general) jade_emu::launch(jade::stft_cta_kernel<R1, jade::MIX_NONE, true>, grid, block, smem, P);
    else jade_emu::launch(jade::stft_cta_kernel<R1, jade::MIX_NONE, false>, grid, block, smem, P);
}
// packed N <= 2048 kernels: T = 32 is jade_pk.cuh, smaller T jade_pk_small.cuh
template <int T, int MIXK, bool WDB, bool GUARD>
void run_pk_one(const KParams& P, int npal, int grid)
{
    const int block = (T == 32 ? jade::PkCfgFor<MIXK>::WARPS : jade::PkSmallCfg<(T == 32 ? 16 : T), MIXK == jade::MIX_NONE>::WARPS) * 32;
    if constexpr (T == 32) {
        // same routing as launch_stft: AbsMean over two channels -> the one-complex-transform kernel (jade_pkz.cuh) for
        // every column (TMA-staged when 16-byte aligned and interior, guarded otherwise) unless JADE_EMU_NOPAIR is set
        // (JADE_EMU_PAIR2: the older two-real-transforms stereo kernel); else guard / cp.async staging / LDG to registers
        const bool pair_ok = MIXK == jade::MIX_SUM && P.channels == 2 && !getenv("JADE_EMU_NOPAIR");
        const bool pair2 = getenv("JADE_EMU_PAIR2") != nullptr;
        if (pair_ok && !pair2) {
            const int zs = jade::PkzCfg::smem_bytes(npal), zb = jade::PkzCfg::WARPS * 32;
            if (GUARD || !P.aligned4) jade_emu::launch(jade::stft_pkz2048_kernel<true, jade::PKZ_GUARD>, grid, zb, zs, P);
            // (launch_one: pixel-only launches with a u8 palette take the instantiations without a clamp instruction)
            else if (getenv("JADE_EMU_RING")) { // launch_one: long evenly spaced runs
                if (!WDB && P.pal_u8) jade_emu::launch(jade::stft_pkz2048_kernel<false, jade::PKZ_RING, true>, grid, zb, zs, P);
                else jade_emu::launch(jade::stft_pkz2048_kernel<WDB, jade::PKZ_RING>, grid, zb, zs, P);
            }
            else if (!WDB && P.pal_u8) jade_emu::launch(jade::stft_pkz2048_kernel<false, jade::PKZ_ASYNC, true>, grid, zb, zs, P);
            else jade_emu::launch(jade::stft_pkz2048_kernel<WDB, jade::PKZ_ASYNC>, grid, zb, zs, P);
        }
        else if (GUARD) jade_emu::launch(jade::stft_pk2048_kernel<MIXK, WDB, jade::PK_LD_GUARD>, grid, block, jade::PkCfgFor<MIXK>::smem_bytes(npal), P);
        else if (P.aligned4 && pair_ok)
            jade_emu::launch(jade::stft_pk2048x2_kernel<WDB>, grid, jade::PkPairCfg::WARPS * 32, jade::PkPairCfg::smem_bytes(npal), P);
        else if (P.aligned4 && MIXK == jade::MIX_NONE && getenv("JADE_EMU_RING") && (P.hop == 256 || P.hop == 512)) { // launch_one: long evenly spaced runs
            if (P.hop == 256) jade_emu::launch(jade::stft_pk2048_kernel<jade::MIX_NONE, WDB, jade::PK_LD_RING4>, grid, block, jade::PkCfgFor<MIXK>::smem_bytes(npal), P);
            else jade_emu::launch(jade::stft_pk2048_kernel<jade::MIX_NONE, WDB, jade::PK_LD_RING8>, grid, block, jade::PkCfgFor<MIXK>::smem_bytes(npal), P);
        }
        else if (P.aligned4) jade_emu::launch(jade::stft_pk2048_kernel<MIXK, WDB, jade::PK_LD_ASYNC>, grid, block, jade::PkCfgFor<MIXK>::smem_bytes(npal), P);
        else jade_emu::launch(jade::stft_pk2048_kernel<MIXK, WDB, jade::PK_LD_DIRECT>, grid, block, jade::PkCfgFor<MIXK>::smem_bytes(npal), P);
    }
    else jade_emu::launch(jade::stft_pksmall_kernel<T, MIXK, WDB, GUARD>, grid, block, jade::PkSmallCfg<T, MIXK == jade::MIX_NONE>::smem_bytes(npal), P);
}
template <int T>
void run_pk(const KParams& P, int mixk, bool wdb, bool guard, int npal, int grid)
{
    if (mixk == jade::MIX_SUM) {
        if (guard) run_pk_one<T, jade::MIX_SUM, true, true>(P, npal, grid);
        else if (wdb) run_pk_one<T, jade::MIX_SUM, true, false>(P, npal, grid);
        else run_pk_one<T, jade::MIX_SUM, false, false>(P, npal, grid);
    } else {
        if (guard) run_pk_one<T, jade::MIX_NONE, true, true>(P, npal, grid);
        else if (wdb) run_pk_one<T, jade::MIX_NONE, true, false>(P, npal, grid);
        else run_pk_one<T, jade::MIX_NONE, false, false>(P, npal, grid);
    }
}
template <int R1>
void run_pkcta(const KParams& P, int mixk, bool want_db, int npal, int grid)
{
    const int smem = jade::PkCtaCfg<R1>::smem_bytes(npal, false);
    const int block = 32 * R1;
    if (mixk == jade::MIX_SUM && want_db) jade_emu::launch(jade::stft_pkcta_kernel<R1, jade::MIX_SUM, true>, grid, block, smem, P);
    else if (mixk == jade::MIX_SUM) jade_emu::launch(jade::stft_pkcta_kernel<R1, jade::MIX_SUM, false>, grid, block, smem, P);
    else if (want_db) jade_emu::launch(jade::stft_pkcta_kernel<R1, jade::MIX_NONE, true>, grid, block, smem, P);
    else jade_emu::launch(jade::stft_pkcta_kernel<R1, jade::MIX_NONE, false>, grid, block, smem, P);
}
} // namespace

extern "C" int emu_render(const jade_config* cin, const int32_t* palette, int npal, float min_db, float max_db,
                          const float* samples, int nstreams, long long nsamples, long long first_col, int ncols,
                          int grid, uint32_t* pix, float* db)
{
    jade_config c = *cin;
    if (c.frames_per_block < 1) c.frames_per_block = 1;
    if (c.block_stride <= 0) c.block_stride = c.hop * c.frames_per_block;
    if (c.preroll < 0) c.preroll = c.fft_size;
    if (c.power_scale <= 0.f) c.power_scale = 1.f;
    const int N = c.fft_size, M = N / 2, B = M + 1;
    std::vector<float> win;
    jade_host::make_window(c.window, N, win);
    const float g = (c.power_scale == 1.0f) ? 0.5f : 0.5f * std::sqrt(c.power_scale);
    for (auto& v : win) v *= g;
    std::vector<uint32_t> baked(npal);
    for (int i = 0; i < npal; ++i) {
        const uint32_t r = (palette[i] >> 16) & 255, gg = (palette[i] >> 8) & 255, b = palette[i] & 255;
        baked[i] = c.pixel_format == JADE_PIX_RGBA8 ? (0xFF000000u | (b << 16) | (gg << 8) | r) : (0xFF000000u | (r << 16) | (gg << 8) | b);
    }
    jade_host::ValueRange range;
    range.set(min_db, max_db, npal);
    int k_lo = 0, k_hi = B, R = B;
    bool pooled = false;
    std::vector<jade::i2> rb;
    if (c.row_map == JADE_ROWS_LINEAR_CROP) {
        jade_host::linear_crop(c.sample_rate, B, c.fmin, c.fmax, k_lo, k_hi);
        R = k_hi - k_lo;
    } else if (c.row_map == JADE_ROWS_LOG_MAXPOOL) {
        std::vector<int32_t> lo, hi;
        jade_host::log_rows(c.sample_rate, N, c.rows, c.fmin, c.fmax, lo, hi);
        rb.resize(c.rows);
        for (int r = 0; r < c.rows; ++r) rb[r] = {lo[r], hi[r]};
        R = c.rows;
        pooled = true;
    }
    const int contributing = (c.mix_mode == JADE_MIX_LEFT || c.mix_mode == JADE_MIX_RIGHT) ? 1 : c.channels;
    int multi = jade::MIX_NONE;
    if (c.mix_mode == JADE_MIX_MIN || (c.mix_mode == JADE_MIX_MAX && contributing > 1)) multi = jade::MIX_SEL;
    else if (c.mix_mode == JADE_MIX_ABSMEAN && contributing > 1) multi = jade::MIX_SUM;
    const bool pow2ch = (c.channels & (c.channels - 1)) == 0;
    const bool general = pooled || c.row_map != JADE_ROWS_IDENTITY || c.db_precise != 0 || c.flip_y == 0 || multi == jade::MIX_SEL ||
                         (multi == jade::MIX_SUM && !pow2ch);
    std::vector<jade_host::cpxf> twP, twI, twA, twH;
    jade_host::twiddles(N, M + 1, 1, twP);
    KParams P;
    memset(&P, 0, sizeof P);
    P.samples = samples;
    P.stream_stride = (long long)c.channels * nsamples;
    P.channel_stride = nsamples;
    P.nsamples = nsamples;
    P.sample_base = 0;
    P.aligned2 = ((nsamples % 2) == 0 && (c.hop % 2) == 0 && (c.block_stride % 2) == 0 && (c.preroll % 2) == 0) ? 1 : 0;
    P.aligned4 = ((nsamples % 4) == 0 && (c.hop % 4) == 0 && (c.block_stride % 4) == 0 && (c.preroll % 4) == 0 &&
                  ((uintptr_t)samples % 16) == 0) ? 1 : 0;
    P.N = N; P.M = M; P.B = B;
    P.hop = c.hop; P.fb = c.frames_per_block; P.bstride = c.block_stride; P.preroll = c.preroll;
    P.first_col = first_col; P.ncols = ncols; P.nstreams = nstreams;
    P.channels = c.channels; P.mix_mode = c.mix_mode;
    P.window = win.data();
    P.twP = reinterpret_cast<const jade::cpx*>(twP.data());
    P.palette = baked.data(); P.npal = npal;
    P.pmin = range.mn; P.pmax = range.mx; P.pmaxc = range.maxclamp(); P.pmult = range.mult;
    jade::colour_fold(P);
    P.pal_u8 = (npal == 256 && baked[P.ci_hi] == baked[255] && !getenv("JADE_EMU_NO_U8")) ? 1 : 0; // as fill_params in jade_gpu.cu
    P.db_precise = c.db_precise;
    P.pooled = pooled; P.R = R; P.k_lo = k_lo; P.k_hi = k_hi; P.flip = c.flip_y;
    P.row_bins = rb.data();
    P.pix = pix; P.db = db;
    P.pix_stream_stride = (long long)ncols * R;
    P.db_stream_stride = (long long)ncols * B;
    std::vector<jade::cpx> se;
    std::vector<float> sp;
    if (N >= 128 && N <= 2048 && !general && multi != jade::MIX_SEL) {
        // same routing as launch_stft in jade_gpu.cu: interior columns -> packed kernel, boundary columns -> its guarded form
        const int T = N / 64;
        jade_host::twiddle_matrix(32 * T, 32, T, twI);
        P.twI = reinterpret_cast<const jade::cpx*>(twI.data());
        auto start = [&](long long j) {
            return (j / c.frames_per_block) * (long long)c.block_stride + (j % c.frames_per_block) * (long long)c.hop - c.preroll;
        };
        const long long j0 = first_col, j1 = first_col + ncols;
        long long lo = j0, hi = j1;
        while (lo < j1 && start(lo) < 0) ++lo;
        while (hi > lo && start(hi - 1) + N > nsamples) --hi;
        if (!P.aligned2 || (T < 32 && !P.aligned4) || getenv("JADE_EMU_FORCE_GUARD")) lo = hi = j0; // everything through the guarded instantiation
        auto sub = [&](long long a, long long b) {
            KParams Q = P;
            Q.first_col = a;
            Q.ncols = (int)(b - a);
            if (Q.pix) Q.pix += (a - j0) * R;
            if (Q.db) Q.db += (a - j0) * B;
            return Q;
        };
        for (int part = 0; part < 3; ++part) {
            const long long a = part == 0 ? lo : (part == 1 ? j0 : hi), b = part == 0 ? hi : (part == 1 ? lo : j1);
            if (b <= a) continue;
            const KParams Q = sub(a, b);
            const bool guard = part != 0, wdb = guard || db != nullptr;
            switch (T) {
            case 2: run_pk<2>(Q, multi, wdb, guard, npal, grid); break;
            case 4: run_pk<4>(Q, multi, wdb, guard, npal, grid); break;
            case 8: run_pk<8>(Q, multi, wdb, guard, npal, grid); break;
            case 16: run_pk<16>(Q, multi, wdb, guard, npal, grid); break;
            case 32: run_pk<32>(Q, multi, wdb, guard, npal, grid); break;
            default: return -1;
            }
        }
    } else if (N <= 2048) {
        const int T = N / 64;
        jade_host::twiddle_matrix(32 * T, 32, T, twI);
        P.twI = reinterpret_cast<const jade::cpx*>(twI.data());
        switch (T) {
        case 1: run_warp<1>(P, multi, general, grid); break;
        case 2: run_warp<2>(P, multi, general, grid); break;
        case 4: run_warp<4>(P, multi, general, grid); break;
        case 8: run_warp<8>(P, multi, general, grid); break;
        case 16: run_warp<16>(P, multi, general, grid); break;
        case 32: run_warp<32>(P, multi, general, grid); break;
        default: return -1;
        }
    } else {
        jade_host::twiddle_matrix(1024, 32, 32, twI);
        P.twI = reinterpret_cast<const jade::cpx*>(twI.data());
        const int Mfft = (N == 65536) ? N / 4 : M;
        const int R1 = Mfft / 1024;
        jade_host::twiddle_matrix(Mfft, R1, 1024, twA);
        P.twA = reinterpret_cast<const jade::cpx*>(twA.data());
        if (N == 65536) {
            jade_host::twiddles(N / 2, N / 4 + 1, 1, twH);
            P.twH = reinterpret_cast<const jade::cpx*>(twH.data());
            se.resize((size_t)grid * (N / 4 + 1));
            sp.resize((size_t)grid * (N / 2 + 1));
            P.scratch_e = se.data();
            P.scratch_p = sp.data();
            if (!getenv("JADE_EMU_CTA2")) { // the product route: a cluster of two CTAs per frame (jade_pk_cluster.cuh)
                const int smem = jade::PkClCfg::smem_bytes(npal, pooled ? R : 0);
                const unsigned cg = 2u * (unsigned)((grid + 1) / 2);
                if (multi == jade::MIX_NONE && !getenv("JADE_EMU_PKCL")) { // one contributing channel: jade_pk_cluster3.cuh
                    jade_emu::launch_cluster2(jade::stft_pkcl3_kernel<jade::MIX_NONE>, cg, 512, jade::PkCl3Cfg::smem_bytes(npal, pooled ? R : 0), P);
                    return R;
                }
                if (multi == jade::MIX_SEL) jade_emu::launch_cluster2(jade::stft_pkcl65536_kernel<jade::MIX_SEL>, cg, 512, smem, P);
                else if (multi == jade::MIX_SUM) jade_emu::launch_cluster2(jade::stft_pkcl65536_kernel<jade::MIX_SUM>, cg, 512, smem, P);
                else jade_emu::launch_cluster2(jade::stft_pkcl65536_kernel<jade::MIX_NONE>, cg, 512, smem, P);
                return R;
            }
            const int smem = jade::PkCtaCfg<16>::smem_bytes2(npal, pooled ? R : 0);
            if (multi == jade::MIX_SEL) jade_emu::launch(jade::stft_pkcta2_kernel<16, jade::MIX_SEL>, grid, 512, smem, P);
            else if (multi == jade::MIX_SUM) jade_emu::launch(jade::stft_pkcta2_kernel<16, jade::MIX_SUM>, grid, 512, smem, P);
            else jade_emu::launch(jade::stft_pkcta2_kernel<16, jade::MIX_NONE>, grid, 512, smem, P);
        } else if (N == 16384 && !general && multi == jade::MIX_NONE && !getenv("JADE_EMU_PKCTA")) {
            // the product route (choose_kernel / launch_stft): three-pass kernel, TMA-staged for interior 16-byte aligned frames,
            // guarded for the rest
            auto start = [&](long long j) {
                return (j / c.frames_per_block) * (long long)c.block_stride + (j % c.frames_per_block) * (long long)c.hop - c.preroll;
            };
            const long long j0 = first_col, j1 = first_col + ncols;
            long long lo = j0, hi = j1;
            while (lo < j1 && start(lo) < 0) ++lo;
            while (hi > lo && start(hi - 1) + N > nsamples) --hi;
            if (!P.aligned4 || getenv("JADE_EMU_FORCE_GUARD")) lo = hi = j0;
            const int smem = jade::Pk3Cfg::smem_bytes(npal);
            for (int part = 0; part < 3; ++part) {
                const long long a = part == 0 ? lo : (part == 1 ? j0 : hi), b = part == 0 ? hi : (part == 1 ? lo : j1);
                if (b <= a) continue;
                KParams Q = P;
                Q.first_col = a;
                Q.ncols = (int)(b - a);
                if (Q.pix) Q.pix += (a - j0) * R;
                if (Q.db) Q.db += (a - j0) * B;
                if (part != 0) jade_emu::launch(jade::stft_pk3_kernel<true, jade::PK3_GUARD>, grid, 256, smem, Q);
                else if (db) jade_emu::launch(jade::stft_pk3_kernel<true, jade::PK3_STAGED>, grid, 256, smem, Q);
                else if (Q.pal_u8) jade_emu::launch(jade::stft_pk3_kernel<false, jade::PK3_STAGED, true>, grid, 256, smem, Q);
                else jade_emu::launch(jade::stft_pk3_kernel<false, jade::PK3_STAGED>, grid, 256, smem, Q);
            }
        } else if (!general && multi != jade::MIX_SEL) {
            switch (R1) {
            case 2: run_pkcta<2>(P, multi, db != nullptr, npal, grid); break;
            case 4: run_pkcta<4>(P, multi, db != nullptr, npal, grid); break;
            case 8: run_pkcta<8>(P, multi, db != nullptr, npal, grid); break;
            case 16: run_pkcta<16>(P, multi, db != nullptr, npal, grid); break;
            default: return -1;
            }
        } else {
            switch (R1) {
            case 2: run_cta<2>(P, multi, general, grid); break;
            case 4: run_cta<4>(P, multi, general, grid); break;
            case 8: run_cta<8>(P, multi, general, grid); break;
            case 16: run_cta<16>(P, multi, general, grid); break;
            default: return -1;
            }
        }
    }
    return R;
}
