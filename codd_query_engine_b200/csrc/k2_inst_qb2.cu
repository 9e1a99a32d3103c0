// K2 instantiations, 2 queries per pass (row lengths up to 768 elements).
#define CV_K2_INSTANTIATE_QB 2
#include "k2_scan.cuh"
namespace cv {
ScanFn pick_scan_qb2(int nch, bool pred) {
    switch (nch) {
        case 1: return scan_fn<1, 2>(pred);
        case 2: return scan_fn<2, 2>(pred);
        case 3: return scan_fn<3, 2>(pred);
        case 4: return scan_fn<4, 2>(pred);
        case 5: return scan_fn<5, 2>(pred);
        case 6: return scan_fn<6, 2>(pred);
        default: return nullptr;
    }
}
}  // namespace cv
