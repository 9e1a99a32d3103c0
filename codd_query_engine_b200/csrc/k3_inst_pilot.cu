// k3_inst_pilot.cu -- the PILOT (two-pass first tile) instantiations of K3, in their own translation unit so that they
// compile beside the others.
#include "k3_kernel.cuh"

namespace cv {
namespace k3 {
MmaKernel pick_mma_kernel_pilot(int k) {
    if (k <= 10) return k_mma_topk<false, 10, 1, 4, true>;
    if (k <= 16) return k_mma_topk<false, 16, 1, 4, true>;
    if (k <= 32) return k_mma_topk<false, 32, 1, 4, true>;
    return k_mma_topk<false, 64, 1, 4, true>;
}
}  // namespace k3
}  // namespace cv
