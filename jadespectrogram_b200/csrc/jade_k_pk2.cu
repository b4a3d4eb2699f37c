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
} // namespace jade_k
