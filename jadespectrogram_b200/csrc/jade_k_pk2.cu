// jade_k_pk2.cu -- instantiations of the N = 2048 stereo kernel (both channels of a frame per warp, jade_pk.cuh); dispatch in
// jade_gpu.cu.
#include "jade_pk.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
kernel_fn pk2048x2_kernel(bool want_db)
{
    using namespace jade;
    return want_db ? (kernel_fn)stft_pk2048x2_kernel<true> : (kernel_fn)stft_pk2048x2_kernel<false>;
}
// long runs of evenly spaced columns, one contributing channel, hop 256 / 512: contiguous columns per warp, samples in a
// tensor-memory ring (PK_LD_RING4 / PK_LD_RING8)
kernel_fn pk2048_run_kernel(bool want_db, int hop)
{
    using namespace jade;
    if (hop == 256) return want_db ? (kernel_fn)stft_pk2048_kernel<MIX_NONE, true, PK_LD_RING4> : (kernel_fn)stft_pk2048_kernel<MIX_NONE, false, PK_LD_RING4>;
    if (hop == 512) return want_db ? (kernel_fn)stft_pk2048_kernel<MIX_NONE, true, PK_LD_RING8> : (kernel_fn)stft_pk2048_kernel<MIX_NONE, false, PK_LD_RING8>;
    return nullptr;
}
} // namespace jade_k
