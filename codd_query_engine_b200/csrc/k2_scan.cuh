// k2_scan.cuh -- K2 `scan_topk`: small-batch (B <= 4) HBM streaming scan + fused top-k + merge.
//
// Replaces hnswlib's searchKnn behind Collection.query (call site
// codd_dal/metrics/metrics_semantic_metadata_store.py:314-316) with an EXACT linear scan:
//   score_i = (sum_d E[i][d] * qhat[d]) * scale[i]        (bf16 rows, fp32 query, fp32 accumulate)
// HBM-bound: algorithmic bytes = rows_read * (2*Dp + 4) (+ 4 or 8 per row of codes, + n/8 mask bytes).
//
// Mapping: a half-warp (16 lanes) owns one row at a time and reads it as 16-byte chunks, lane l
// taking chunks l, l+16, ... (each load instruction covers 256 contiguous bytes per half-warp, fully
// sectored).  Every warp iteration handles S = 2*RPH rows ("slots") so that 12-ish independent
// 16-byte loads per lane are in flight before any arithmetic.  The query lives in registers
// (normalised in the CTA prologue, bit-identical to the oracle).  Partial sums of the V = RPH*QB
// (row, query) pairs of a half-warp are reduced with a transposing butterfly (V/2 + V/4 + ... shuffles
// instead of 4*V).  Each warp keeps a sorted k-list per query in shared memory behind a register
// threshold; CTAs merge their warps' lists, publish one k-list per query, and the last CTA to finish
// merges all CTA lists (threadfence + counter), so a query batch is ONE kernel launch.
#pragma once
#include "common.cuh"

namespace cv {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;

struct ScanArgs {
    const uint4* E;         // [n][C] 16-byte chunks (bf16 x 8)
    const float* scale;     // [n]
    const uint32_t* ns;     // [n]
    const uint32_t* label;  // [n]
    const uint32_t* mask;   // nullptr or ceil(n/32) words
    uint32_t ns_code;
    uint32_t label_code;
    int64_t n;              // committed rows (snapshot)
    int C;                  // chunks per row = Dp / 8
    int D;                  // true dimension
    int cosine;
    const float* q;         // [QB][D] raw fp32 queries (device)
    int k;
    uint32_t row_offset;    // global row of local row 0
    uint64_t* partial;      // [QB][gridDim.x][k]
    unsigned int* counter;  // self-resetting "CTAs done" counter
    uint64_t* out;          // [QB][k] final keys (desc, 0 padded)
    PeerExchange px;        // world > 1: the last CTA also exchanges + merges across GPUs (fused)
    uint64_t seq;           // exchange sequence number of this batch
    int q_base;             // global query index of this launch's query 0
};

// Transposing butterfly over the 16 lanes of a half-warp.  On return the lane holds the full sum of
// element idx where bit l of idx = (hl >> (3 - l)) & 1 for l < log2(V).
template <int V>
__device__ __forceinline__ float reduce16(float (&v)[V], int hl) {
    static_assert(V == 1 || V == 2 || V == 4 || V == 8 || V == 16, "V must be a power of two <= 16");
    int cur = V;
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        if (cur > 1) {
            const bool upper = (hl & off) != 0;
#pragma unroll
            for (int j = 0; j < V / 2; ++j) {
                if (j < cur / 2) {
                    float a = v[2 * j], b = v[2 * j + 1];
                    float keep = upper ? b : a;
                    float send = upper ? a : b;
                    v[j] = keep + __shfl_xor_sync(kFullMask, send, off);
                }
            }
            cur >>= 1;
        } else {
            v[0] += __shfl_xor_sync(kFullMask, v[0], off);
        }
    }
    return v[0];
}

template <int V>
__device__ __forceinline__ int reduce16_index(int hl) {
    int idx = 0, level = 0, cur = V;
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) {
        if (cur > 1) {
            idx |= ((hl & off) ? 1 : 0) << level;
            ++level;
            cur >>= 1;
        }
    }
    return idx;
}

__device__ __forceinline__ uint64_t block_max_u64(uint64_t v, uint64_t* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        uint64_t t = shfl_u64(v, (threadIdx.x & 31) ^ o);
        v = t > v ? t : v;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    uint64_t r = 0;
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) r = s_red[w] > r ? s_red[w] : r;
    return r;
}

// Shared by K2 and K3: the CTA has k-lists of `nlists` producers per query in shared memory
// (lists[(p * QBn + b) * k + j]); merge -> publish -> last CTA merges all CTAs.
// selbuf: kSelCap keys of shared scratch.
template <int QBn>
__device__ __forceinline__ void publish_and_final_merge(const uint64_t* lists, int nlists, int k,
                                                        uint64_t* partial, unsigned int* counter,
                                                        uint64_t* out, uint64_t* selbuf, int* s_cnt,
                                                        uint64_t* s_red, const PeerExchange& px, uint64_t seq,
                                                        int q_base) {
    __shared__ int s_last;
    __syncthreads();
#pragma unroll 1
    for (int b = 0; b < QBn; ++b) {
        auto src = [&](int i) { int p = i / k, j = i - p * k; return lists[(p * QBn + b) * k + j]; };
        block_select_topk(src, nlists * k, k, 0ull, selbuf, s_cnt);
        for (int j = threadIdx.x; j < k; j += blockDim.x)
            partial[(static_cast<size_t>(b) * gridDim.x + blockIdx.x) * k + j] = selbuf[j];
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int G = gridDim.x;
#pragma unroll 1
    for (int b = 0; b < QBn; ++b) {
        const uint64_t* pb = partial + static_cast<size_t>(b) * G * k;
        uint64_t tau = 0;
        for (int c = threadIdx.x; c < G; c += blockDim.x) {
            uint64_t kth = __ldcg(pb + static_cast<size_t>(c) * k + (k - 1));
            tau = kth > tau ? kth : tau;
        }
        tau = block_max_u64(tau, s_red);
        auto src = [&](int i) { return __ldcg(pb + i); };
        block_select_topk<8>(src, G * k, k, tau, selbuf, s_cnt);
        if (px.world > 1) exchange_and_merge(px, seq, q_base + b, selbuf, k, selbuf, s_cnt);  // fused NVLink exchange
        for (int j = threadIdx.x; j < k; j += blockDim.x) out[static_cast<size_t>(b) * k + j] = selbuf[j];
        __syncthreads();
    }
    if (threadIdx.x == 0) *counter = 0u;
}

// rows-per-half per iteration: keep ~12 (<= 16) independent 16-byte loads per lane in flight.
__host__ __device__ constexpr int scan_rph(int nch, int qb) {
    int r = nch <= 1 ? 8 : nch <= 3 ? 4 : nch <= 6 ? 2 : 1;
    while (r * qb > 16) r >>= 1;
    return r < 1 ? 1 : r;
}

template <int NCH, int QB, bool PRED>
__global__ void __launch_bounds__(kScanThreads, (NCH * QB <= 6) ? 2 : 1) k_scan_topk(ScanArgs a) {
    constexpr int RPH = scan_rph(NCH, QB);
    constexpr int S = 2 * RPH;   // row slots per warp iteration
    constexpr int V = RPH * QB;  // partial sums per half-warp
    static_assert(V <= 16, "too many partial sums per half-warp");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* selbuf = reinterpret_cast<uint64_t*>(smem_raw);                 // kSelCap keys (32 KB)
    uint64_t* lists = selbuf + kSelCap;                                       // [warps][QB][k]
    float* sq = reinterpret_cast<float*>(smem_raw);                           // prologue alias: [QB][D]
    __shared__ float s_norm[QB];
    __shared__ int s_cnt;
    __shared__ uint64_t s_red[kScanWarps];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int half = lane >> 4;
    const int hl = lane & 15;
    const int k = a.k;
    const int C = a.C;

    // ---- prologue: normalise the queries exactly like the oracle (fp64 norm, fp32 divide) ----
    for (int i = tid; i < QB * a.D; i += kScanThreads) sq[i] = a.q[i];
    for (int i = tid; i < kScanWarps * QB * k; i += kScanThreads) lists[i] = 0ull;
    __syncthreads();
    if (warp < QB) {
        float nrm = 1.0f;
        if (a.cosine) {
            double ss = 0.0;
            for (int d = lane; d < a.D; d += 32) {
                double v = static_cast<double>(sq[warp * a.D + d]);
                ss += v * v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFullMask, ss, o);
            float n32 = static_cast<float>(sqrt(ss));
            nrm = n32 > 0.0f ? n32 : 1.0f;
        }
        if (lane == 0) s_norm[warp] = nrm;
    }
    __syncthreads();
    float qr[QB][NCH][8];
#pragma unroll
    for (int b = 0; b < QB; ++b)
#pragma unroll
        for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                int d = (hl + 16 * c) * 8 + e;
                qr[b][c][e] = (d < a.D) ? __fdiv_rn(sq[b * a.D + d], s_norm[b]) : 0.0f;
            }
    __syncthreads();  // sq aliases selbuf; everyone is done reading it

    // which (row slot, query) this lane ends up holding after the butterfly
    const int my_idx = reduce16_index<V>(hl);
    const int my_b = my_idx % QB;
    const int my_slot = 2 * (my_idx / QB) + half;
    const bool owner = (hl & ((16 / V) - 1)) == 0;
    uint64_t* my_lists = lists + static_cast<size_t>(warp) * QB * k;
    uint64_t thr = 0ull;  // k-th best key of my query's list (0 while not full)

    const int64_t gw = static_cast<int64_t>(blockIdx.x) * kScanWarps + warp;
    const int64_t nw = static_cast<int64_t>(gridDim.x) * kScanWarps;

    auto process = [&](const int64_t (&rows)[S]) {
        // rows[s] < 0 => empty slot.  Issue every load first.
        uint4 ld[RPH][NCH];
#pragma unroll
        for (int i = 0; i < RPH; ++i) {
            const int64_t r = half ? rows[2 * i + 1] : rows[2 * i];  // (select, not a dynamic index)
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int ci = hl + 16 * c;
                if (r >= 0 && ci < C) ld[i][c] = ldg_stream(a.E + r * C + ci);
                else ld[i][c] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        int64_t my_row = -1;
#pragma unroll
        for (int s = 0; s < S; ++s) my_row = (s == my_slot) ? rows[s] : my_row;
        const float sc = (my_row >= 0) ? __ldg(a.scale + my_row) : 0.0f;

        float acc[V];
#pragma unroll
        for (int i = 0; i < RPH; ++i) {
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const uint4 w = ld[i][c];
                    s0 = fmaf(bf16lo_f32(w.x), qr[b][c][0], s0);
                    s1 = fmaf(bf16hi_f32(w.x), qr[b][c][1], s1);
                    s0 = fmaf(bf16lo_f32(w.y), qr[b][c][2], s0);
                    s1 = fmaf(bf16hi_f32(w.y), qr[b][c][3], s1);
                    s0 = fmaf(bf16lo_f32(w.z), qr[b][c][4], s0);
                    s1 = fmaf(bf16hi_f32(w.z), qr[b][c][5], s1);
                    s0 = fmaf(bf16lo_f32(w.w), qr[b][c][6], s0);
                    s1 = fmaf(bf16hi_f32(w.w), qr[b][c][7], s1);
                }
                acc[i * QB + b] = s0 + s1;
            }
        }
        const float dot = reduce16<V>(acc, hl);
        const float score = dot * sc;
        const uint64_t key = pack_key(score, a.row_offset + static_cast<uint32_t>(my_row));
        const bool cand = owner && my_row >= 0 && key > thr;
        unsigned int pending = __ballot_sync(kFullMask, cand);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint64_t ck = shfl_u64(key, src);
            const int cb = __shfl_sync(kFullMask, my_b, src);
            const uint64_t kth = warp_list_insert(my_lists + cb * k, k, ck, lane);
            if (cb == my_b) thr = kth;
        }
    };

    if (!PRED) {
        const int64_t groups = (a.n + S - 1) / S;
        for (int64_t g = gw; g < groups; g += nw) {
            int64_t rows[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                int64_t r = g * S + s;
                rows[s] = r < a.n ? r : -1;
            }
            process(rows);
        }
    } else {
        // 128 rows per warp step: build the eligibility words, then consume set bits S at a time.  The
        // codes / mask words of the NEXT step are loaded before the current step's rows are processed, so
        // their latency overlaps the row loads instead of preceding them.
        const int64_t blocks = (a.n + 127) / 128;
        const int64_t nwords = (a.n + 31) / 32;
        const bool use_ns = a.ns_code != kAny, use_label = a.label_code != kAny;
        uint32_t nsv[4], lbv[4], mkv[4];
        auto fetch = [&](int64_t blk) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int64_t r = blk * 128 + 32 * w + lane;
                const bool in = blk < blocks && r < a.n;
                nsv[w] = (in && use_ns) ? __ldg(a.ns + r) : a.ns_code;
                lbv[w] = (in && use_label) ? __ldg(a.label + r) : a.label_code;
                const int64_t wi = blk * 4 + w;
                mkv[w] = (a.mask && blk < blocks && wi < nwords) ? __ldg(a.mask + wi) : (a.mask ? 0u : 0xFFFFFFFFu);
            }
        };
        fetch(gw);
        for (int64_t blk = gw; blk < blocks; blk += nw) {
            const int64_t base = blk * 128;
            uint32_t words[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int64_t r = base + 32 * w + lane;
                const bool ok = r < a.n && nsv[w] == a.ns_code && lbv[w] == a.label_code;
                words[w] = __ballot_sync(kFullMask, ok) & mkv[w];
            }
            fetch(blk + nw);
            int remaining = __popc(words[0]) + __popc(words[1]) + __popc(words[2]) + __popc(words[3]);
            while (remaining > 0) {
                int64_t rows[S];
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    int64_t r = -1;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        if (r < 0 && words[w]) {
                            const int bit = __ffs(words[w]) - 1;
                            words[w] &= words[w] - 1;
                            r = base + 32 * w + bit;
                        }
                    }
                    rows[s] = r;
                }
                remaining -= S;
                process(rows);
            }
        }
    }

    publish_and_final_merge<QB>(lists, kScanWarps, k, a.partial, a.counter, a.out, selbuf, &s_cnt, s_red, a.px, a.seq,
                                a.q_base);
}

// K2 for long rows (Dp > 1024, up to CV_MAX_DIM): the query no longer fits the register file, so it
// lives normalised in shared memory; a full warp owns kWideRows rows at a time and its lanes stride over
// the rows' 16-byte chunks (each load instruction covers 512 contiguous bytes of one row).  One query per
// launch.  Same per-warp lists, CTA merge and last-CTA global merge as the register-resident kernel.
constexpr int kWideRows = 4;

template <bool PRED>
__global__ void __launch_bounds__(kScanThreads, 2) k_scan_topk_wide(ScanArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* selbuf = reinterpret_cast<uint64_t*>(smem_raw);          // kSelCap keys
    uint64_t* lists = selbuf + kSelCap;                                // [warps][k]
    float* sq = reinterpret_cast<float*>(lists + kScanWarps * a.k);    // [C * 8] normalised query, zero padded
    __shared__ float s_norm;
    __shared__ int s_cnt;
    __shared__ uint64_t s_red[kScanWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k, C = a.C;
    for (int i = tid; i < C * 8; i += kScanThreads) sq[i] = i < a.D ? a.q[i] : 0.0f;
    for (int i = tid; i < kScanWarps * k; i += kScanThreads) lists[i] = 0ull;
    __syncthreads();
    if (warp == 0) {
        float nrm = 1.0f;
        if (a.cosine) {
            double ss = 0.0;
            for (int d = lane; d < a.D; d += 32) {
                double v = static_cast<double>(sq[d]);
                ss += v * v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(kFullMask, ss, o);
            float n32 = static_cast<float>(sqrt(ss));
            nrm = n32 > 0.0f ? n32 : 1.0f;
        }
        if (lane == 0) s_norm = nrm;
    }
    __syncthreads();
    for (int i = tid; i < a.D; i += kScanThreads) sq[i] = __fdiv_rn(sq[i], s_norm);
    __syncthreads();

    uint64_t* my_list = lists + static_cast<size_t>(warp) * k;
    uint64_t thr = 0ull;
    const int64_t gw = static_cast<int64_t>(blockIdx.x) * kScanWarps + warp;
    const int64_t nw = static_cast<int64_t>(gridDim.x) * kScanWarps;
    const float4* sq4 = reinterpret_cast<const float4*>(sq);

    auto process = [&](const int64_t (&rows)[kWideRows]) {
        float acc[kWideRows];
#pragma unroll
        for (int r = 0; r < kWideRows; ++r) acc[r] = 0.0f;
#pragma unroll 2
        for (int c = lane; c < C; c += 32) {
            uint4 w[kWideRows];
#pragma unroll
            for (int r = 0; r < kWideRows; ++r)
                w[r] = rows[r] >= 0 ? ldg_stream(a.E + rows[r] * C + c) : make_uint4(0u, 0u, 0u, 0u);
            const float4 q0 = sq4[2 * c], q1 = sq4[2 * c + 1];
#pragma unroll
            for (int r = 0; r < kWideRows; ++r) {
                float s0 = fmaf(bf16lo_f32(w[r].x), q0.x, 0.0f), s1 = bf16hi_f32(w[r].x) * q0.y;
                s0 = fmaf(bf16lo_f32(w[r].y), q0.z, s0);
                s1 = fmaf(bf16hi_f32(w[r].y), q0.w, s1);
                s0 = fmaf(bf16lo_f32(w[r].z), q1.x, s0);
                s1 = fmaf(bf16hi_f32(w[r].z), q1.y, s1);
                s0 = fmaf(bf16lo_f32(w[r].w), q1.z, s0);
                s1 = fmaf(bf16hi_f32(w[r].w), q1.w, s1);
                acc[r] += s0 + s1;
            }
        }
#pragma unroll
        for (int r = 0; r < kWideRows; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(kFullMask, acc[r], o);
            if (rows[r] >= 0) {  // warp-uniform
                const float score = acc[r] * __ldg(a.scale + rows[r]);
                const uint64_t key = pack_key(score, a.row_offset + static_cast<uint32_t>(rows[r]));
                if (key > thr) thr = warp_list_insert(my_list, k, key, lane);
            }
        }
    };

    if (!PRED) {
        const int64_t groups = (a.n + kWideRows - 1) / kWideRows;
        for (int64_t g = gw; g < groups; g += nw) {
            int64_t rows[kWideRows];
#pragma unroll
            for (int r = 0; r < kWideRows; ++r) {
                const int64_t row = g * kWideRows + r;
                rows[r] = row < a.n ? row : -1;
            }
            process(rows);
        }
    } else {
        const int64_t blocks = (a.n + 127) / 128;
        const int64_t nwords = (a.n + 31) / 32;
        for (int64_t blk = gw; blk < blocks; blk += nw) {
            const int64_t base = blk * 128;
            uint32_t words[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int64_t r = base + 32 * w + lane;
                bool ok = r < a.n;
                if (ok && a.ns_code != kAny) ok = __ldg(a.ns + r) == a.ns_code;
                if (ok && a.label_code != kAny) ok = __ldg(a.label + r) == a.label_code;
                uint32_t word = __ballot_sync(kFullMask, ok);
                if (a.mask) {
                    const int64_t wi = (base >> 5) + w;
                    word &= (wi < nwords) ? __ldg(a.mask + wi) : 0u;
                }
                words[w] = word;
            }
            int remaining = __popc(words[0]) + __popc(words[1]) + __popc(words[2]) + __popc(words[3]);
            while (remaining > 0) {
                int64_t rows[kWideRows];
#pragma unroll
                for (int r = 0; r < kWideRows; ++r) {
                    int64_t row = -1;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {
                        if (row < 0 && words[w]) {
                            const int bit = __ffs(words[w]) - 1;
                            words[w] &= words[w] - 1;
                            row = base + 32 * w + bit;
                        }
                    }
                    rows[r] = row;
                }
                remaining -= kWideRows;
                process(rows);
            }
        }
    }
    publish_and_final_merge<1>(lists, kScanWarps, k, a.partial, a.counter, a.out, selbuf, &s_cnt, s_red, a.px, a.seq,
                               a.q_base);
}

using ScanFn = void (*)(ScanArgs);
// largest query-block size the register-resident K2 supports for a row of nch*16 16-byte chunks
inline int scan_max_qb(int nch) { return nch <= 3 ? 4 : nch <= 6 ? 2 : 1; }  // nch > 8: wide kernel, 1 query
inline bool scan_is_wide(int nch) { return nch > 8; }
ScanFn pick_scan_qb1(int nch, bool pred);
ScanFn pick_scan_qb2(int nch, bool pred);
ScanFn pick_scan_qb4(int nch, bool pred);

#ifdef CV_K2_INSTANTIATE_QB
template <int NCH, int QB>
inline ScanFn scan_fn(bool pred) {
    return pred ? static_cast<ScanFn>(k_scan_topk<NCH, QB, true>) : static_cast<ScanFn>(k_scan_topk<NCH, QB, false>);
}
#endif

// K4 `merge_topk`: one CTA per query merges G lists of k_in keys ([G][B][k_in]) into k_out keys.
// Used (a) after the NCCL all-gather of per-GPU top-k and (b) to combine query-tile passes.
#ifdef CV_MAIN_TU
__global__ void __launch_bounds__(256) k_merge_keys(const uint64_t* in, int G, int B, int k_in, int k_out,
                                                    uint64_t* out) {
    __shared__ __align__(16) uint64_t selbuf[kSelCap];
    __shared__ int s_cnt;
    __shared__ uint64_t s_red[8];
    const int b = blockIdx.x;
    uint64_t tau = 0;
    if (k_in >= k_out) {
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            uint64_t kth = in[(static_cast<size_t>(g) * B + b) * k_in + (k_out - 1)];
            tau = kth > tau ? kth : tau;
        }
        tau = block_max_u64(tau, s_red);
    }
    auto src = [&](int i) { int g = i / k_in, j = i - g * k_in; return in[(static_cast<size_t>(g) * B + b) * k_in + j]; };
    block_select_topk<8>(src, G * k_in, k_out, tau, selbuf, &s_cnt);
    for (int j = threadIdx.x; j < k_out; j += blockDim.x) out[static_cast<size_t>(b) * k_out + j] = selbuf[j];
}

// Batched path: K3 leaves this GPU's top-k per query in `local` [B][k]; one CTA per query does the
// fused NVLink exchange + cross-GPU merge.
__global__ void __launch_bounds__(256) k_exchange_merge(PeerExchange px, uint64_t seq, int q_base, const uint64_t* local,
                                                        int k, uint64_t* out) {
    __shared__ __align__(16) uint64_t selbuf[kSelCap];
    __shared__ int s_cnt;
    const int b = blockIdx.x;
    for (int j = threadIdx.x; j < k; j += blockDim.x) selbuf[j] = local[static_cast<size_t>(b) * k + j];
    __syncthreads();
    exchange_and_merge(px, seq, q_base + b, selbuf, k, selbuf, &s_cnt);
    for (int j = threadIdx.x; j < k; j += blockDim.x) out[static_cast<size_t>(b) * k + j] = selbuf[j];
}
#endif  // CV_MAIN_TU

}  // namespace cv
