// jade_k_pk.cu -- instantiations of the packed N = 2048 kernel (jade_pk.cuh); see jade_gpu.cu for the dispatch.
#include "jade_pk.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
kernel_fn pk2048_kernel(int mixk, bool want_db, bool guard)
{
    using namespace jade;
    if (mixk == MIX_SUM) {
        if (guard) return (kernel_fn)stft_pk2048_kernel<MIX_SUM, true, true>;
        return want_db ? (kernel_fn)stft_pk2048_kernel<MIX_SUM, true, false> : (kernel_fn)stft_pk2048_kernel<MIX_SUM, false, false>;
    }
    if (guard) return (kernel_fn)stft_pk2048_kernel<MIX_NONE, true, true>;
    return want_db ? (kernel_fn)stft_pk2048_kernel<MIX_NONE, true, false> : (kernel_fn)stft_pk2048_kernel<MIX_NONE, false, false>;
}
} // namespace jade_k
