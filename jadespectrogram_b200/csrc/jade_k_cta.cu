// jade_k_cta.cu -- instantiations of stft_cta_kernel<R1> (jade_kernels.cuh; general-epilogue path of N >= 4096).
#include "jade_kernels.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
namespace {
template <int R1>
kernel_fn pick(int mixk, bool general)
{
    using namespace jade;
    // the fast path of these sizes is stft_pkcta_kernel (jade_k_pkcta.cu); only the general epilogue lives here
    if (!general && mixk != MIX_SEL) return nullptr;
    if (mixk == MIX_SEL) return (kernel_fn)stft_cta_kernel<R1, MIX_SEL, true>;
    return mixk == MIX_SUM ? (kernel_fn)stft_cta_kernel<R1, MIX_SUM, true> : (kernel_fn)stft_cta_kernel<R1, MIX_NONE, true>;
}
} // namespace
kernel_fn cta_kernel(int R1, int mixk, bool general)
{
    switch (R1) {
    case 2: return pick<2>(mixk, general);
    case 4: return pick<4>(mixk, general);
    case 8: return pick<8>(mixk, general);
    case 16: return pick<16>(mixk, general);
    default: return nullptr;
    }
}
} // namespace jade_k
