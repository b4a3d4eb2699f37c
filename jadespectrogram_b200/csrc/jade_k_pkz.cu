// jade_k_pkz.cu -- instantiations of the N = 2048 stereo kernel that transforms both channels as one complex signal
// (jade_pkz.cuh); dispatch in jade_gpu.cu.
#include "jade_pkz.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
// guard: bounds-checked global loads (boundary columns, unaligned geometries); otherwise TMA-staged interior frames
kernel_fn pkz2048_kernel(bool want_db, bool guard)
{
    using namespace jade;
    if (guard) return want_db ? (kernel_fn)stft_pkz2048_kernel<true, PKZ_GUARD> : (kernel_fn)stft_pkz2048_kernel<false, PKZ_GUARD>;
    return want_db ? (kernel_fn)stft_pkz2048_kernel<true, PKZ_ASYNC> : (kernel_fn)stft_pkz2048_kernel<false, PKZ_ASYNC>;
}
// long runs of evenly spaced columns (hop = N/4): contiguous columns per warp, samples kept in a tensor-memory ring
kernel_fn pkz2048_run_kernel(bool want_db)
{
    using namespace jade;
    return want_db ? (kernel_fn)stft_pkz2048_kernel<true, PKZ_RING> : (kernel_fn)stft_pkz2048_kernel<false, PKZ_RING>;
}
// pixel-only instantiations for palettes with P.pal_u8 (run: the long-run instantiation)
kernel_fn pkz2048_u8_kernel(bool run)
{
    using namespace jade;
    return run ? (kernel_fn)stft_pkz2048_kernel<false, PKZ_RING, true> : (kernel_fn)stft_pkz2048_kernel<false, PKZ_ASYNC, true>;
}
} // namespace jade_k
