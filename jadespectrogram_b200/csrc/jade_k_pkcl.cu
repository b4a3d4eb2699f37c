// jade_k_pkcl.cu -- instantiations of the N = 65536 cluster kernel (two CTAs + distributed shared memory, jade_pk_cluster.cuh);
// dispatch in jade_gpu.cu.
#include "jade_pk_cluster.cuh"
#include "jade_pk_cluster3.cuh"
namespace jade_k {
typedef void (*kernel_fn)(const jade::KParams);
kernel_fn pkcl65536_kernel(int mixk)
{
    using namespace jade;
    return mixk == MIX_SEL ? (kernel_fn)stft_pkcl65536_kernel<MIX_SEL>
         : mixk == MIX_SUM ? (kernel_fn)stft_pkcl65536_kernel<MIX_SUM>
                           : (kernel_fn)stft_pkcl65536_kernel<MIX_NONE>;
}
// one contributing channel: three register passes per CTA, window in tensor memory (jade_pk_cluster3.cuh)
kernel_fn pkcl3_kernel() { return (kernel_fn)jade::stft_pkcl3_kernel<jade::MIX_NONE>; }
} // namespace jade_k
