// k3_inst_cg2.cu -- the CTA-pair (cta_group::2) instantiations of K3, in their own translation unit so that they
// compile beside the single-CTA ones.
#include "k3_kernel.cuh"

namespace cv {
namespace k3 {
MmaKernel pick_mma_kernel_cg2(bool dump, int k, int ew) { return pick_mma_kernel<2>(dump, k, ew); }
}  // namespace k3
}  // namespace cv
