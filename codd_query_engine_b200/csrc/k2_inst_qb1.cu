// K2 instantiations, 1 query per pass (register-resident query up to 1024 elements, wide kernel above).
#define CV_K2_INSTANTIATE_QB 1
#include "k2_scan.cuh"
namespace cv {
ScanFn pick_scan_qb1(int nch, bool pred) {
    switch (nch) {
        case 1: return scan_fn<1, 1>(pred);
        case 2: return scan_fn<2, 1>(pred);
        case 3: return scan_fn<3, 1>(pred);
        case 4: return scan_fn<4, 1>(pred);
        case 5: return scan_fn<5, 1>(pred);
        case 6: return scan_fn<6, 1>(pred);
        case 7: return scan_fn<7, 1>(pred);
        case 8: return scan_fn<8, 1>(pred);
        default:  // rows longer than 1024 elements: query in shared memory
            return pred ? static_cast<ScanFn>(k_scan_topk_wide<true>) : static_cast<ScanFn>(k_scan_topk_wide<false>);
    }
}
}  // namespace cv
