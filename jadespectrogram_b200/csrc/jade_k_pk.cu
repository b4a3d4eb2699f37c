// jade_k_pk.cu -- instantiations of the packed N = 2048 kernel (jade_pk.cuh); see jade_gpu.cu for the dispatch.
#include "jade_pk.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
// load: jade::PK_LD_ASYNC (cp.async staging, 16-byte aligned interior frames), PK_LD_DIRECT (8-byte aligned interior
// frames), PK_LD_GUARD (bounds-checked; always the dB-storing form)
kernel_fn pk2048_kernel(int mixk, bool want_db, int load)
{
    using namespace jade;
    if (mixk == MIX_SUM) {
        if (load == PK_LD_GUARD) return (kernel_fn)stft_pk2048_kernel<MIX_SUM, true, PK_LD_GUARD>;
        if (load == PK_LD_DIRECT)
            return want_db ? (kernel_fn)stft_pk2048_kernel<MIX_SUM, true, PK_LD_DIRECT> : (kernel_fn)stft_pk2048_kernel<MIX_SUM, false, PK_LD_DIRECT>;
        return want_db ? (kernel_fn)stft_pk2048_kernel<MIX_SUM, true, PK_LD_ASYNC> : (kernel_fn)stft_pk2048_kernel<MIX_SUM, false, PK_LD_ASYNC>;
    }
    if (load == PK_LD_GUARD) return (kernel_fn)stft_pk2048_kernel<MIX_NONE, true, PK_LD_GUARD>;
    if (load == PK_LD_DIRECT)
        return want_db ? (kernel_fn)stft_pk2048_kernel<MIX_NONE, true, PK_LD_DIRECT> : (kernel_fn)stft_pk2048_kernel<MIX_NONE, false, PK_LD_DIRECT>;
    return want_db ? (kernel_fn)stft_pk2048_kernel<MIX_NONE, true, PK_LD_ASYNC> : (kernel_fn)stft_pk2048_kernel<MIX_NONE, false, PK_LD_ASYNC>;
}
} // namespace jade_k
