// jade_k_pksmall_b.cu -- instantiations of stft_pksmall_kernel<T> for T in {8,16} (jade_pk_small.cuh); see jade_gpu.cu for the dispatch.
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
kernel_fn pksmall_kernel_a(int T, int mixk, bool want_db, bool guard); // jade_k_pksmall_a.cu
kernel_fn pksmall_kernel(int T, int mixk, bool want_db, bool guard)
{
    switch (T) {
    case 8: return pick<8>(mixk, want_db, guard);
    case 16: return pick<16>(mixk, want_db, guard);
    default: return pksmall_kernel_a(T, mixk, want_db, guard);
    }
}
} // namespace jade_k
