// jade_k_pksmall_a.cu -- instantiations of stft_pksmall_kernel<T> for T in {2,4} (jade_pk_small.cuh); see jade_gpu.cu for the dispatch.
#include "jade_pk_small.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
namespace {
template <int T>
kernel_fn pick(int mixk, bool want_db, bool guard)
{
    using namespace jade;
    if (mixk == MIX_SUM) {
        if (guard) return (kernel_fn)stft_pksmall_kernel<T, MIX_SUM, true, true>;
        return want_db ? (kernel_fn)stft_pksmall_kernel<T, MIX_SUM, true, false> : (kernel_fn)stft_pksmall_kernel<T, MIX_SUM, false, false>;
    }
    if (guard) return (kernel_fn)stft_pksmall_kernel<T, MIX_NONE, true, true>;
    return want_db ? (kernel_fn)stft_pksmall_kernel<T, MIX_NONE, true, false> : (kernel_fn)stft_pksmall_kernel<T, MIX_NONE, false, false>;
}
} // namespace
kernel_fn pksmall_kernel_a(int T, int mixk, bool want_db, bool guard)
{
    switch (T) {
    case 2: return pick<2>(mixk, want_db, guard);
    case 4: return pick<4>(mixk, want_db, guard);
    default: return nullptr;
    }
}
} // namespace jade_k
