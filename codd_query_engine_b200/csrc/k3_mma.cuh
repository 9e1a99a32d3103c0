// k3_mma.cuh -- interface of K3 `gemm_topk` (implementation in k3_mma.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>

namespace cv {

struct MmaScratch {  // device scratch owned by the query context, grown on demand
    unsigned char* p = nullptr;      // work area: bf16 queries (long rows) + candidate keys; contents do not matter between launches
    size_t cap = 0;
    unsigned char* state = nullptr;  // cross-CTA thresholds / counters for state_q queries: zero between launches (self-resetting)
    int state_q = 0;
    uint32_t epoch = 0;              // launch counter: tags the super-tile progress words (stale tags are ignored, no reset needed)
};

struct MmaArgs {
    const void* E;          // [n][Dp] bf16
    const float* scale;
    const uint32_t* ns;
    const uint32_t* label;
    const uint32_t* mask;
    uint32_t ns_code, label_code;
    int64_t n;
    int D, Dp;
    bool cosine;
    uint32_t row_offset;
    uint32_t row_stride;    // global row = row_offset + local row * row_stride (0 / 1 = dense)
    int precision;          // CV_PRECISION_*
    bool scale_bounded;     // every row scale is 0, NaN or within 2^-7 of 1
    const float* q;         // [B][D] fp32 raw (device)
    int B, k;
    uint64_t* out;          // [B][k]
    float* dump;            // debug: write every score to dump[b * dump_ld + row] (device), or nullptr
    int64_t dump_ld;
    unsigned int* counter;
    int sm_count;
    int variant;            // 0 = pick by batch size, 1 = single-CTA tiles, 2 = CTA pairs (cta_group::2)
    cudaStream_t stream;
};

bool mma_supported(int Dp);
// returns 0 or a CV_ERR_* code; text in mma_error().  *launches is incremented per kernel launch.
int mma_enqueue(const MmaArgs& a, MmaScratch& scratch, std::atomic<int64_t>* launches);
std::string& mma_error();

// K4 launcher (kernel lives in coddvec.cu): merge G lists [G][b_stride][k_in] -> out[nq][k_out]
cudaError_t launch_merge_keys(const uint64_t* in, int G, int b_stride, int k_in, int k_out, uint64_t* out, int nq,
                              cudaStream_t stream, const uint32_t* bound32 = nullptr);

}  // namespace cv
