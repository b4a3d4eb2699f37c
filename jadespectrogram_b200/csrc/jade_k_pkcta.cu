// jade_k_pkcta.cu -- instantiations of the packed large-N kernels (jade_pk_cta.cuh); see jade_gpu.cu for the dispatch.
#include "jade_pk_cta.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
namespace {
template <int R1>
kernel_fn pick(int mixk, bool want_db)
{
    using namespace jade;
    if (mixk == MIX_SUM) return want_db ? (kernel_fn)stft_pkcta_kernel<R1, MIX_SUM, true> : (kernel_fn)stft_pkcta_kernel<R1, MIX_SUM, false>;
    return want_db ? (kernel_fn)stft_pkcta_kernel<R1, MIX_NONE, true> : (kernel_fn)stft_pkcta_kernel<R1, MIX_NONE, false>;
}
} // namespace
kernel_fn pkcta_kernel(int R1, int mixk, bool want_db)
{
    switch (R1) {
    case 2: return pick<2>(mixk, want_db);
    case 4: return pick<4>(mixk, want_db);
    case 8: return pick<8>(mixk, want_db);
    case 16: return pick<16>(mixk, want_db);
    default: return nullptr;
    }
}
kernel_fn pkcta2_kernel(int mixk)
{
    using namespace jade;
    return mixk == MIX_SEL ? (kernel_fn)stft_pkcta2_kernel<16, MIX_SEL>
         : mixk == MIX_SUM ? (kernel_fn)stft_pkcta2_kernel<16, MIX_SUM>
                           : (kernel_fn)stft_pkcta2_kernel<16, MIX_NONE>;
}
} // namespace jade_k
