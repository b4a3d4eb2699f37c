// jade_k_warp_b.cu -- instantiations of stft_warp_kernel<T> for T in {16,32} (jade_kernels.cuh); see jade_gpu.cu for the dispatch.
#include "jade_kernels.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
namespace {
template <int T>
kernel_fn pick(int mixk, bool general)
{
    using namespace jade;
    if (mixk == MIX_SEL) return (kernel_fn)stft_warp_kernel<T, MIX_SEL, true>;
    if constexpr (T >= 2) {
        // N >= 128: the fast path is the packed kernel (jade_k_pk*.cu); only the general epilogue lives here
        if (!general) return nullptr;
        return mixk == MIX_SUM ? (kernel_fn)stft_warp_kernel<T, MIX_SUM, true> : (kernel_fn)stft_warp_kernel<T, MIX_NONE, true>;
    } else {
        if (mixk == MIX_SUM) return general ? (kernel_fn)stft_warp_kernel<T, MIX_SUM, true> : (kernel_fn)stft_warp_kernel<T, MIX_SUM, false>;
        return general ? (kernel_fn)stft_warp_kernel<T, MIX_NONE, true> : (kernel_fn)stft_warp_kernel<T, MIX_NONE, false>;
    }
}
} // namespace
kernel_fn warp_kernel_small(int T, int mixk, bool general); // jade_k_warp_a.cu
kernel_fn warp_kernel(int T, int mixk, bool general)
{
    switch (T) {
    case 16: return pick<16>(mixk, general);
    case 32: return pick<32>(mixk, general);
    default: return warp_kernel_small(T, mixk, general);
    }
}
} // namespace jade_k
