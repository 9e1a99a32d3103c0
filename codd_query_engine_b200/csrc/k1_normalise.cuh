// k1_normalise.cuh -- K1 `normalise_append`: fp32 rows -> L2-normalise -> bf16 row + scale + codes.
//
// Replaces what chromadb/hnswlib does inside Collection.upsert (call site
// codd_dal/metrics/metrics_semantic_metadata_store.py:236-238): cosine space normalises at insert.
// One warp per row.  Norms are accumulated in fp64 and rounded once to fp32 so that the stored
// bf16 bits are identical to oracle/brute_force.py::normalise_store.
// HBM-bound: reads 4*D, writes 2*Dp + 12 bytes per row (cfg 5).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace cv {

struct NormArgs {
    const float* x;            // [n][D] staged input (device)
    const int64_t* dst_rows;   // [n] local destination rows, or nullptr => base_row + i
    int64_t base_row;
    int64_t n;
    int D;
    int Dp;                    // padded row stride (elements)
    int cosine;
    const uint32_t* ns_in;     // [n] or nullptr
    const uint32_t* label_in;  // [n] or nullptr
    __nv_bfloat16* E;
    float* scale;
    uint32_t* ns;
    uint32_t* label;
    unsigned int* flags;       // flags[1] is raised when a stored scale is not 0 and not within 2^-7 of 1 (K3's prefilter bound)
};

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

__global__ void __launch_bounds__(256) k_normalise_append(NormArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
    for (int64_t i = warp; i < a.n; i += nwarps) {
        const float* r = a.x + i * a.D;
        const int64_t dst = a.dst_rows ? a.dst_rows[i] : a.base_row + i;
        __nv_bfloat16* o = a.E + dst * a.Dp;
        float safe = 1.0f;
        if (a.cosine) {
            double ss = 0.0;
            for (int d = lane; d < a.D; d += 32) {
                double v = static_cast<double>(r[d]);
                ss += v * v;
            }
            ss = warp_sum_f64(ss);
            float norm = static_cast<float>(sqrt(ss));
            safe = norm > 0.0f ? norm : 1.0f;
        }
        double ss2 = 0.0;
        for (int d = lane; d < a.Dp; d += 32) {
            float v = (d < a.D) ? __fdiv_rn(r[d], safe) : 0.0f;
            __nv_bfloat16 b = __float2bfloat16_rn(v);
            o[d] = b;
            double w = static_cast<double>(__bfloat162float(b));
            ss2 += w * w;
        }
        if (a.cosine) ss2 = warp_sum_f64(ss2);
        if (lane == 0) {
            const float sc = a.cosine ? (ss2 > 0.0 ? static_cast<float>(1.0 / sqrt(ss2)) : 0.0f) : 1.0f;
            if (sc != 0.0f && !(fabsf(sc - 1.0f) <= 0.0078125f)) atomicOr(a.flags + 1, 1u);
            a.scale[dst] = sc;
            a.ns[dst] = a.ns_in ? a.ns_in[i] : 0u;
            a.label[dst] = a.label_in ? a.label_in[i] : 0u;
        }
    }
}

// Counts non-finite inputs (NaN / Inf) so that an upsert can be rejected before anything is written.
__global__ void __launch_bounds__(256) k_count_nonfinite(const float* x, int64_t count, unsigned int* out) {
    unsigned int bad = 0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        uint32_t u = __float_as_uint(x[i]);
        bad += ((u & 0x7F800000u) == 0x7F800000u) ? 1u : 0u;
    }
    bad = __reduce_add_sync(kFullMask, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(out, bad);
}

}  // namespace cv
