// jade_k_pkcl.cu -- instantiations of the N = 65536 cluster kernel (two CTAs + distributed shared memory, jade_pk_cluster.cuh);
// dispatch in jade_gpu.cu.
#include "jade_pk_cluster.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
kernel_fn pkcl65536_kernel(int mixk)
{
    using namespace jade;
    return mixk == MIX_SEL ? (kernel_fn)stft_pkcl65536_kernel<MIX_SEL>
         : mixk == MIX_SUM ? (kernel_fn)stft_pkcl65536_kernel<MIX_SUM>
                           : (kernel_fn)stft_pkcl65536_kernel<MIX_NONE>;
}
} // namespace jade_k
