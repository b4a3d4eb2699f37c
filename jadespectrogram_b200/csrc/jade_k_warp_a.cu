// jade_k_warp_a.cu -- instantiations of stft_warp_kernel<T> for T in {1,2,4,8} (jade_kernels.cuh); see jade_gpu.cu for the dispatch.
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
kernel_fn warp_kernel_small(int T, int mixk, bool general)
{
    switch (T) {
    case 1: return pick<1>(mixk, general);
    case 2: return pick<2>(mixk, general);
    case 4: return pick<4>(mixk, general);
    case 8: return pick<8>(mixk, general);
    default: return nullptr;
    }
}
} // namespace jade_k
