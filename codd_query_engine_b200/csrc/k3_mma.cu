// k3_mma.cu -- K3 `gemm_topk`.  Placeholder: the tcgen05 path is not built yet.
#include "k3_mma.cuh"
namespace cv {
bool mma_supported(int) { return false; }
std::string& mma_error() { static thread_local std::string e; return e; }
int mma_enqueue(const MmaArgs&, MmaScratch&, std::atomic<int64_t>*) {
    mma_error() = "tensor-core path not built";
    return -5;
}
}  // namespace cv
