// jade_k_pk3.cu -- instantiations of the N = 16384 three-pass kernel (jade_pk3.cuh); dispatch in jade_gpu.cu.
#include "jade_pk3.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
// guard: bounds-checked global loads (boundary columns, unaligned geometries); otherwise TMA-staged interior frames
kernel_fn pk3_kernel(bool want_db, bool guard)
{
    using namespace jade;
    if (guard) return want_db ? (kernel_fn)stft_pk3_kernel<true, PK3_GUARD> : (kernel_fn)stft_pk3_kernel<false, PK3_GUARD>;
    return want_db ? (kernel_fn)stft_pk3_kernel<true, PK3_STAGED> : (kernel_fn)stft_pk3_kernel<false, PK3_STAGED>;
}
// pixel-only, TMA-staged instantiation for palettes with P.pal_u8
kernel_fn pk3_u8_kernel() { return (kernel_fn)jade::stft_pk3_kernel<false, jade::PK3_STAGED, true>; }
} // namespace jade_k
