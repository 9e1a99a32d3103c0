// K2 instantiations, 4 queries per pass (row lengths up to 384 elements).
#define CV_K2_INSTANTIATE_QB 4
#include "k2_scan.cuh"
namespace cv {
ScanFn pick_scan_qb4(int nch, bool pred) {
    switch (nch) {
        case 1: return scan_fn<1, 4>(pred);
        case 2: return scan_fn<2, 4>(pred);
        case 3: return scan_fn<3, 4>(pred);
        default: return nullptr;
    }
}
}  // namespace cv
