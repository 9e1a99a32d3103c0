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
// instead of 4*V).
//
// Top-k: a candidate that beats the lane's register threshold is appended to an unsorted per-(warp, query)
// buffer in shared memory (common.cuh: warp_buf_append / warp_prune).  The threshold is the larger of the
// buffer's k-th best at its last prune and the cross-CTA HEADS BOUND (heads_bound_key below), which tracks the
// rows every warp of the grid has seen.  Tail (publish_and_final_merge): each CTA appends its survivors
// compactly to one global list, the last CTA to finish (threadfence + counter) selects the k best, runs the
// fused cross-GPU exchange if there is one and writes the result -- a query batch is ONE kernel launch.
//
// Three flavours of the inner step (process): FAST -- the unfiltered scan's full groups of consecutive rows (one
// base pointer, constant offsets, no validity tests); DENSE -- the predicate scan's full groups gathered from the
// per-warp ring of eligible rows; general -- ragged ends and rows whose last 16-lane chunk group is partial.
#pragma once
#include "common.cuh"

namespace cv {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;

constexpr int kInlineQ = 1024;

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
    uint32_t row_stride;    // global row = row_offset + local row * row_stride
    uint64_t* partial;      // [QB][gridDim.x][k]
    unsigned int* counter;  // self-resetting "CTAs done" counter
    uint64_t* out;          // [QB][k] final keys (desc, 0 padded)
    int q_inline;           // 1: the queries of this pass travel in qinline[] with the launch instead of behind q
    uint64_t* done_flag;    // nullptr, or host-mapped word the last CTA sets to done_value once out[] is written
    uint64_t done_value;    // (cv_query's low-latency completion: the host polls it instead of a stream sync)
    uint64_t* heads;        // [QB][kHeadSlots] group maxima of the keys warps have kept (self-resetting, see heads_bound_key)
    PeerExchange px;        // world > 1: the last CTA also exchanges + merges across GPUs (fused)
    uint64_t seq;           // exchange sequence number of this batch
    int q_base;             // global query index of this launch's query 0
    // Low-latency path of cv_query (one interactive query, D <= kInlineQ): the query rides in the kernel
    // parameters, so no copy-engine operation precedes the launch and no CTA reads it over PCIe.
    float qinline[kInlineQ];
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

// ---- cross-CTA "heads" bound ------------------------------------------------------------------------
// Every warp of the grid folds the best key it has kept for query b into heads[b][its slot] (atomicMax; a
// slot pools the warps with the same global warp index mod kHeadSlots).  The slot maxima are keys of
// DISTINCT rows, so the k-th largest of them is a lower bound on the k-th best key of the whole scan -- and a
// tight one: it tracks the rows ALL warps have seen, where a warp's own k-th best only reflects its few
// hundred rows.  One warp per CTA recomputes it at iteration counts 1, 2, 4, 8, ... (the warps take turns) and
// publishes it in shared memory; every warp folds it into its register threshold.  With it a warp stops
// buffering candidates after its first handful of rows, which is what makes large k cheap when the rows per
// warp are few (predicate scans, small shards).  Only the score half of the keys is ranked (a bit-wise binary
// search over 32 bits: 4 ballots per bit), the returned key has row bits 0, i.e. it is <= the exact bound.
constexpr int kHeadSlots = 128;  // >= kMaxK

__device__ __forceinline__ uint64_t heads_bound_key(const uint64_t* heads, int k, int lane) {
    uint32_t h[kHeadSlots / 32];
#pragma unroll
    for (int j = 0; j < kHeadSlots / 32; ++j)
        h[j] = static_cast<uint32_t>(__ldcg(reinterpret_cast<const unsigned long long*>(heads) + j * 32 + lane) >> 32);
    uint32_t t = 0u;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t c = t | (1u << bit);
        int n = 0;
#pragma unroll
        for (int j = 0; j < kHeadSlots / 32; ++j) n += __popc(__ballot_sync(kFullMask, h[j] >= c));
        if (n >= k) t = c;
    }
    return static_cast<uint64_t>(t) << 32;
}

// it = process() calls this warp has finished; true on the calls where this warp refreshes the CTA's bound
__device__ __forceinline__ bool heads_refresh_turn(uint32_t it, int warp) {
    return (it & (it - 1u)) == 0u && ((31 - __clz(it)) & (kScanWarps - 1)) == warp;
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

// The CTA has candidate buffers of `nlists` producers per query in shared memory
// (lists[(p * QBn + b) * len + j], any order, 0 = empty).  Tail of the launch:
//   1. every CTA selects its k best above the heads bound it already holds in shared memory (bound[b], from its
//      last refresh; on a long scan within ~2x of the exact k-th best, so only a handful of keys in the whole
//      grid survive) and appends the survivors COMPACTLY to partial[b][...] (one atomicAdd on counter[1 + b]
//      per CTA reserves the range);
//   2. the last CTA to finish (threadfence + counter[0]) selects the k best of the survivors -- a few hundred
//      keys instead of gridDim.x * k; if the bounds were still loose (short scans) it first recomputes the
//      final heads bound so that one small sort suffices -- runs the fused cross-GPU exchange if there is
//      one, writes out[b][0..k) and resets the counters and the heads for the next launch.
// partial: [QBn][gridDim.x * k] keys.  selbuf: kSelCap keys of shared scratch.
template <int QBn>
__device__ __forceinline__ void publish_and_final_merge(const uint64_t* lists, int nlists, int len, int k,
                                                        uint64_t* partial, unsigned int* counter,
                                                        uint64_t* out, uint64_t* selbuf, int* s_cnt,
                                                        uint64_t* s_red, const PeerExchange& px, uint64_t seq,
                                                        int q_base, uint64_t* heads, const unsigned long long* bound,
                                                        uint64_t* done_flag = nullptr, uint64_t done_value = 0) {
    __shared__ int s_last;
    __shared__ unsigned int s_pos;
    __shared__ unsigned long long s_hb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t stride = static_cast<size_t>(gridDim.x) * k;
    (void)s_red;
    __syncthreads();
#pragma unroll 1
    for (int b = 0; b < QBn; ++b) {
        auto src = [&](int i) { int p = i / len, j = i - p * len; return lists[(p * QBn + b) * len + j]; };
        block_select_topk(src, nlists * len, k, bound[b], selbuf, s_cnt);
        const int nsurv = __syncthreads_count(static_cast<int>(threadIdx.x) < k && selbuf[threadIdx.x] != 0ull);
        if (threadIdx.x == 0) s_pos = nsurv ? atomicAdd(counter + 1 + b, static_cast<unsigned int>(nsurv)) : 0u;
        __syncthreads();
        CV_ASSERT(nsurv <= k && s_pos + nsurv <= stride);
        if (static_cast<int>(threadIdx.x) < nsurv) partial[b * stride + s_pos + threadIdx.x] = selbuf[threadIdx.x];
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
#pragma unroll 1
    for (int b = 0; b < QBn; ++b) {
        const uint64_t* pb = partial + b * stride;
        const int m = static_cast<int>(__ldcg(counter + 1 + b));
        CV_ASSERT(m >= 0 && static_cast<size_t>(m) <= stride);
        if (warp == 0) {
            const uint64_t hb = m > 512 ? heads_bound_key(heads + b * kHeadSlots, k, lane) : 0ull;
            if (lane == 0) s_hb = hb;
        }
        __syncthreads();
        auto src = [&](int i) { return __ldcg(pb + i); };
        block_select_topk<8>(src, m, k, s_hb, selbuf, s_cnt);
        if (px.world > 1) exchange_and_merge(px, seq, q_base + b, selbuf, k, selbuf, s_cnt);  // fused NVLink exchange
        for (int j = threadIdx.x; j < k; j += blockDim.x) out[static_cast<size_t>(b) * k + j] = selbuf[j];
        __syncthreads();
    }
    if (threadIdx.x <= QBn) counter[threadIdx.x] = 0u;
    for (int i = threadIdx.x; i < QBn * kHeadSlots; i += blockDim.x) heads[i] = 0ull;  // every other CTA is done
    if (done_flag) {  // out[] (host-mapped in this mode) is complete: tell the polling host thread
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) st_release_sys(done_flag, done_value);
    }
}

// rows-per-half per iteration: keep ~12 (<= 16) independent 16-byte loads per lane in flight.
__host__ __device__ constexpr int scan_rph(int nch, int qb) {
    int r = nch <= 1 ? 8 : nch <= 3 ? 4 : nch <= 6 ? 2 : 1;
    while (r * qb > 16) r >>= 1;
    return r < 1 ? 1 : r;
}

// Row slots handed to process(): the unfiltered scan counts rows in int64 (-1 = empty slot); the predicate
// scan carries LOCAL rows as u32 (keys hold 32-bit global rows, so a shard never exceeds 2^32 - 1 rows) with
// 0xFFFFFFFF = empty, which halves the index arithmetic of its row-gather.
template <bool PRED> struct ScanRow { using type = int64_t; };
template <> struct ScanRow<true> { using type = uint32_t; };
__device__ __forceinline__ bool row_valid(int64_t r) { return r >= 0; }
__device__ __forceinline__ bool row_valid(uint32_t r) { return r != 0xFFFFFFFFu; }

constexpr int kRingCap = 256;  // per-warp ring of eligible rows (predicate scan): >= 128 + max slots - 1
constexpr int kCodeStages = 4; // 128-row steps of code columns in flight per warp (cp.async staging)

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// QS ("queries in shared memory"): variants whose queries would need more than 48 registers per lane
// (NCH * QB > 6) keep the normalised queries in shared memory instead -- two 16-byte arrays per 8-element chunk
// (elements 0-3 / 4-7) so that the lanes' LDS.128 reads are conflict-free -- and fetch them per (chunk, query)
// inside the FMA loop (1 LDS.128 per 16 FMAs).  That keeps every variant at two CTAs per SM; the register
// variants are unchanged.  The summation order is the register variant's, so results are bit-identical.
__host__ __device__ constexpr bool scan_queries_in_smem(int nch, int qb) { return nch * qb > 6; }

template <int NCH, int QB, bool PRED>
__global__ void __launch_bounds__(kScanThreads, 2) k_scan_topk(ScanArgs a) {
    constexpr bool QS = scan_queries_in_smem(NCH, QB);
    using row_t = typename ScanRow<PRED>::type;
    constexpr row_t kEmptyRow = static_cast<row_t>(-1);
    constexpr int RPH = scan_rph(NCH, QB);
    constexpr int S = 2 * RPH;   // row slots per warp iteration
    constexpr int V = RPH * QB;  // partial sums per half-warp
    static_assert(V <= 16, "too many partial sums per half-warp");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* selbuf = reinterpret_cast<uint64_t*>(smem_raw);                 // kSelCap keys (32 KB)
    uint64_t* lists = selbuf + kSelCap;                                       // [warps][QB][cap] candidate buffers
    float* sq = reinterpret_cast<float*>(smem_raw);                           // prologue alias: [QB][D]
    // QS only: normalised queries behind the lists, [QB][C] x (elements 0-3) then [QB][C] x (elements 4-7)
    float4* sqlo = reinterpret_cast<float4*>(lists + static_cast<size_t>(kScanWarps) * QB * scan_list_cap(a.k));
    float4* sqhi = sqlo + QB * a.C;
    __shared__ float s_norm[QB];
    __shared__ int s_cnt;
    __shared__ uint64_t s_red[kScanWarps];
    __shared__ unsigned long long s_bound[QB];  // heads bound of each query, refreshed by the warps in turn

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int half = lane >> 4;
    const int hl = lane & 15;
    const int k = a.k;
    const int cap = scan_list_cap(k);
    const int C = a.C;

    // ---- prologue: normalise the queries exactly like the oracle (fp64 norm, fp32 divide) ----
    if (a.q_inline) {  // QB * D <= kInlineQ
        for (int i = tid; i < QB * a.D; i += kScanThreads) sq[i] = a.qinline[i];
    } else {
        for (int i = tid; i < QB * a.D; i += kScanThreads) sq[i] = a.q[i];
    }
    for (int i = tid; i < kScanWarps * QB * cap; i += kScanThreads) lists[i] = 0ull;
    if (tid < QB) s_bound[tid] = 0ull;
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
    float qr[QS ? 1 : QB][QS ? 1 : NCH][8];
    if constexpr (QS) {
        for (int i = tid; i < QB * C * 8; i += kScanThreads) {
            const int b = i / (C * 8), d = i - b * (C * 8);
            const float v = (d < a.D) ? __fdiv_rn(sq[b * a.D + d], s_norm[b]) : 0.0f;
            float* dst = reinterpret_cast<float*>(((d & 4) ? sqhi : sqlo) + b * C + (d >> 3));
            dst[d & 3] = v;
        }
    } else {
#pragma unroll
        for (int b = 0; b < QB; ++b)
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    int d = (hl + 16 * c) * 8 + e;
                    qr[b][c][e] = (d < a.D) ? __fdiv_rn(sq[b * a.D + d], s_norm[b]) : 0.0f;
                }
    }
    __syncthreads();  // sq aliases selbuf; everyone is done reading it

    // which (row slot, query) this lane ends up holding after the butterfly
    const int my_idx = reduce16_index<V>(hl);
    const int my_b = my_idx % QB;
    const int my_slot = 2 * (my_idx / QB) + half;
    const bool owner = (hl & ((16 / V) - 1)) == 0;
    uint64_t* my_lists = lists + static_cast<size_t>(warp) * QB * cap;
    uint64_t thr = 0ull;  // k-th best key of my query's buffer at its last prune (0 before the first)
    int cnt[QB];          // keys in each query's buffer (warp-uniform)
    uint64_t best[QB];    // best key this warp has kept per query (what it has folded into the heads)
#pragma unroll
    for (int b = 0; b < QB; ++b) { cnt[b] = 0; best[b] = 0ull; }
    uint32_t it = 0;      // process() calls finished

    const int64_t gw = static_cast<int64_t>(blockIdx.x) * kScanWarps + warp;
    const int64_t nw = static_cast<int64_t>(gridDim.x) * kScanWarps;
    const int head_slot = static_cast<int>(gw % kHeadSlots);

    // fast = std::true_type: the S slots are the consecutive rows rows[0] .. rows[0] + S - 1, all of them exist and
    // every lane's chunks exist (C == 16 * NCH); gp = address of this lane's first chunk of the row in slot `half`.
    // No validity tests, no row selects: the loads are base + constant offsets.  That is every group of the
    // unfiltered scan except a ragged last one.
    // (FAST is a literal at every call site and the lambda is inlined, so the unused side folds away; a generic
    // lambda with `if constexpr` here crashes nvcc 12.9's front end.)
    // DENSE (predicate scan, full groups out of the ring): arbitrary rows, but every slot is filled and every
    // lane's chunks exist, so the gather needs no validity tests either.
    auto process = [&](const bool FAST, const bool DENSE, const row_t (&rows)[S], const uint4* gp) {
        // rows[s] == kEmptyRow => empty slot.  Issue every load first.
        uint4 ld[RPH][NCH];
        row_t my_row = kEmptyRow;
        if (FAST) {
#pragma unroll
            for (int i = 0; i < RPH; ++i) {
                const uint4* rp = gp + static_cast<size_t>(2 * i) * C;
#pragma unroll
                for (int c = 0; c < NCH; ++c) ld[i][c] = ldg_stream(rp + 16 * c);
            }
            my_row = rows[0] + my_slot;
        } else {
#pragma unroll
            for (int i = 0; i < RPH; ++i) {
                const row_t r = half ? rows[2 * i + 1] : rows[2 * i];  // (select, not a dynamic index)
                const uint4* rp = a.E + static_cast<size_t>(r) * C + hl;
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (DENSE || (row_valid(r) && hl + 16 * c < C)) ld[i][c] = ldg_stream(rp + 16 * c);
                    else ld[i][c] = make_uint4(0u, 0u, 0u, 0u);
                }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) my_row = (s == my_slot) ? rows[s] : my_row;
        }
        CV_ASSERT(!(FAST || DENSE || row_valid(my_row)) || static_cast<int64_t>(my_row) < a.n);
        const float sc = (FAST || DENSE || row_valid(my_row)) ? __ldg(a.scale + my_row) : 0.0f;

        float acc[V];
        if constexpr (QS) {
            float a0[V], a1[V];
#pragma unroll
            for (int v = 0; v < V; ++v) { a0[v] = 0.0f; a1[v] = 0.0f; }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const int ci = min(hl + 16 * c, C - 1);  // (a chunk past the row's end was loaded as zeros)
#pragma unroll
                for (int b = 0; b < QB; ++b) {
                    const float4 ql = sqlo[b * C + ci], qh = sqhi[b * C + ci];
#pragma unroll
                    for (int i = 0; i < RPH; ++i) {
                        const uint4 w = ld[i][c];
                        float s0 = a0[i * QB + b], s1 = a1[i * QB + b];
                        s0 = fmaf(bf16lo_f32(w.x), ql.x, s0);
                        s1 = fmaf(bf16hi_f32(w.x), ql.y, s1);
                        s0 = fmaf(bf16lo_f32(w.y), ql.z, s0);
                        s1 = fmaf(bf16hi_f32(w.y), ql.w, s1);
                        s0 = fmaf(bf16lo_f32(w.z), qh.x, s0);
                        s1 = fmaf(bf16hi_f32(w.z), qh.y, s1);
                        s0 = fmaf(bf16lo_f32(w.w), qh.z, s0);
                        s1 = fmaf(bf16hi_f32(w.w), qh.w, s1);
                        a0[i * QB + b] = s0;
                        a1[i * QB + b] = s1;
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = a0[v] + a1[v];
        } else {
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
        }
        const float dot = reduce16<V>(acc, hl);
        const float score = dot * sc;
        const uint64_t key = pack_key(score, a.row_offset + static_cast<uint32_t>(my_row) * a.row_stride);
        {
            const uint64_t sb = s_bound[my_b];
            thr = sb > thr ? sb : thr;
        }
        const bool cand = owner && (FAST || DENSE || row_valid(my_row)) && sc == sc && key > thr;  // NaN scale = deleted row
        unsigned int pending = __ballot_sync(kFullMask, cand);
        while (pending) {
            const int src = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint64_t ck = shfl_u64(key, src);
            const int cb = __shfl_sync(kFullMask, my_b, src);
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                if (b == cb) {  // warp-uniform
                    uint64_t kth;
                    if (warp_buf_append(my_lists + b * cap, cap, k, ck, lane, cnt[b], kth) && b == my_b && kth > thr) thr = kth;
                    if (ck > best[b]) {
                        best[b] = ck;
                        if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(a.heads) + b * kHeadSlots + head_slot, ck);
                    }
                }
            }
        }
        ++it;
        if (heads_refresh_turn(it, warp)) {  // warp-uniform, rare
#pragma unroll 1
            for (int b = 0; b < QB; ++b) {
                const uint64_t hb = heads_bound_key(a.heads + b * kHeadSlots, k, lane);
                if (lane == 0 && hb != 0ull) atomicMax(&s_bound[b], hb);
            }
        }
    };

    if constexpr (!PRED) {
        const int64_t groups = (a.n + S - 1) / S;
        const int64_t full_groups = a.n / S;
        auto ragged = [&](int64_t g) {
            int64_t rows[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                int64_t r = g * S + s;
                rows[s] = r < a.n ? r : -1;
            }
            process(false, false, rows, nullptr);
        };
        if (C == 16 * NCH) {
            const uint4* gp = a.E + static_cast<size_t>(gw * S + half) * C + hl;
            const size_t gstride = static_cast<size_t>(nw) * S * C;
            for (int64_t g = gw; g < full_groups; g += nw, gp += gstride) {
                int64_t rows[S];
                rows[0] = g * S;
                process(true, false, rows, gp);
            }
            if (full_groups < groups && full_groups % nw == gw) ragged(full_groups);
        } else {
            for (int64_t g = gw; g < groups; g += nw) ragged(g);
        }
    } else {
        // 128 rows per warp step.  Lane l tests rows base + 4l .. base + 4l + 3 (one 16-byte load per code
        // column, one mask word), four ballots turn that into eligibility words, and the eligible rows are
        // COMPACTED into a per-warp ring in shared memory (position = running tail + popc of the lower lanes).
        // The ring is consumed S rows at a time, so every process() call runs with all its slots filled no
        // matter how the selected rows fall into 128-row steps; only the final flush is partial.  The codes /
        // mask word of the NEXT step are loaded before this step's rows are processed, so their latency
        // overlaps the row loads.
        __shared__ __align__(16) uint32_t s_ring[kScanWarps][kRingCap];
        // Code columns are staged through shared memory with cp.async, kCodeStages 128-row steps ahead: the copies
        // hold no registers while in flight, so the prefetch can be deep enough to hide the L2/HBM latency even
        // when almost no row is eligible and a step is only a few dozen instructions long (1 % selectivity).
        __shared__ __align__(16) uint4 s_codes[kScanWarps][kCodeStages][2][32];  // [stage][ns | label][lane]
        uint32_t* ring = s_ring[warp];
        const uint32_t lt_mask = (1u << lane) - 1u;
        const int64_t blocks = (a.n + 127) / 128;
        const int64_t nwords = (a.n + 31) / 32;
        const bool use_ns = a.ns_code != kAny, use_label = a.label_code != kAny;
        uint32_t mkv;
        // issue the copies of step `blk` into stage `st` (nothing for columns not in use / rows past the end), then
        // close the group -- one group per step, empty or not, so that wait_group counts steps
        auto stage_codes = [&](int64_t blk, int st) {
            const int64_t r0 = blk * 128 + 4 * lane;  // 16-byte aligned; the arrays are mapped in 16-byte multiples
            if (blk < blocks && r0 < a.n) {
                if (use_ns) cp_async16(&s_codes[warp][st][0][lane], a.ns + r0);
                if (use_label) cp_async16(&s_codes[warp][st][1][lane], a.label + r0);
            }
            cp_async_commit();
        };
        auto fetch_mask = [&](int64_t blk) {
            const int64_t wi = blk * 4 + (lane >> 3);
            mkv = a.mask ? ((blk < blocks && wi < nwords) ? __ldg(a.mask + wi) : 0u) : 0xFFFFFFFFu;
        };
        const bool fullc = C == 16 * NCH;
        auto consume = [&](uint32_t at, uint32_t avail) {  // avail >= S: full group
            row_t rows[S];
#pragma unroll
            for (int s = 0; s < S; ++s) rows[s] = (static_cast<uint32_t>(s) < avail) ? ring[(at + s) & (kRingCap - 1)] : kEmptyRow;
            if (fullc && avail >= static_cast<uint32_t>(S)) process(false, true, rows, nullptr);
            else process(false, false, rows, nullptr);
        };
        uint32_t head = 0, tail = 0;  // warp-uniform; head stays a multiple of S
#pragma unroll
        for (int st = 0; st < kCodeStages; ++st) stage_codes(gw + st * nw, st);
        fetch_mask(gw);
        int st = 0;
        for (int64_t blk = gw; blk < blocks; blk += nw) {
            cp_async_wait<kCodeStages - 1>();  // this step's group has landed (the younger ones may be in flight)
            __syncwarp();
            const int64_t rem = a.n - (blk * 128 + 4 * lane);  // rows of this lane's quad that exist
            const int left = rem >= 4 ? 4 : (rem > 0 ? static_cast<int>(rem) : 0);
            const uint4 none_ns = make_uint4(a.ns_code, a.ns_code, a.ns_code, a.ns_code);
            const uint4 none_lb = make_uint4(a.label_code, a.label_code, a.label_code, a.label_code);
            const uint4 nsv = (use_ns && left > 0) ? s_codes[warp][st][0][lane] : none_ns;
            const uint4 lbv = (use_label && left > 0) ? s_codes[warp][st][1][lane] : none_lb;
            __syncwarp();  // every lane has read the stage before it is refilled
            stage_codes(blk + static_cast<int64_t>(kCodeStages) * nw, st);
            st = st + 1 == kCodeStages ? 0 : st + 1;
            const uint32_t base = static_cast<uint32_t>(blk * 128) + 4u * lane;
            const uint32_t mbits = mkv >> ((4 * lane) & 31);
            const bool ok0 = left > 0 && nsv.x == a.ns_code && lbv.x == a.label_code && (mbits & 1u);
            const bool ok1 = left > 1 && nsv.y == a.ns_code && lbv.y == a.label_code && (mbits & 2u);
            const bool ok2 = left > 2 && nsv.z == a.ns_code && lbv.z == a.label_code && (mbits & 4u);
            const bool ok3 = left > 3 && nsv.w == a.ns_code && lbv.w == a.label_code && (mbits & 8u);
            const uint32_t w0 = __ballot_sync(kFullMask, ok0), w1 = __ballot_sync(kFullMask, ok1);
            const uint32_t w2 = __ballot_sync(kFullMask, ok2), w3 = __ballot_sync(kFullMask, ok3);
            fetch_mask(blk + nw);
            uint32_t t = tail;
            if (ok0) ring[(t + __popc(w0 & lt_mask)) & (kRingCap - 1)] = base;
            t += __popc(w0);
            if (ok1) ring[(t + __popc(w1 & lt_mask)) & (kRingCap - 1)] = base + 1;
            t += __popc(w1);
            if (ok2) ring[(t + __popc(w2 & lt_mask)) & (kRingCap - 1)] = base + 2;
            t += __popc(w2);
            if (ok3) ring[(t + __popc(w3 & lt_mask)) & (kRingCap - 1)] = base + 3;
            tail = t + __popc(w3);
            __syncwarp();
            CV_ASSERT(tail - head <= static_cast<uint32_t>(kRingCap));
            while (tail - head >= static_cast<uint32_t>(S)) {
                consume(head, S);
                head += S;
            }
            __syncwarp();  // ring slots read above may be overwritten by the next step's appends
        }
        cp_async_wait<0>();  // nothing may still be landing in shared memory when the tail reuses it
        if (tail != head) consume(head, tail - head);
    }

    // slots past each buffer's count hold keys a prune dropped: blank them before the CTA merge reads all `cap`
    __syncwarp();
#pragma unroll
    for (int b = 0; b < QB; ++b)
        for (int i = cnt[b] + lane; i < cap; i += 32) my_lists[b * cap + i] = 0ull;
    publish_and_final_merge<QB>(lists, kScanWarps, cap, k, a.partial, a.counter, a.out, selbuf, &s_cnt, s_red, a.px, a.seq,
                                a.q_base, a.heads, s_bound, a.done_flag, a.done_value);
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
    uint64_t* lists = selbuf + kSelCap;                                // [warps][cap] candidate buffers
    const int cap = scan_list_cap(a.k);
    float* sq = reinterpret_cast<float*>(lists + kScanWarps * cap);    // [C * 8] normalised query, zero padded
    __shared__ float s_norm;
    __shared__ int s_cnt;
    __shared__ uint64_t s_red[kScanWarps];
    __shared__ unsigned long long s_bound;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = a.k, C = a.C;
    if (tid == 0) s_bound = 0ull;
    for (int i = tid; i < C * 8; i += kScanThreads) sq[i] = i < a.D ? a.q[i] : 0.0f;
    for (int i = tid; i < kScanWarps * cap; i += kScanThreads) lists[i] = 0ull;
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

    uint64_t* my_list = lists + static_cast<size_t>(warp) * cap;
    uint64_t thr = 0ull, best = 0ull;
    int cnt = 0;
    uint32_t it = 0;
    const int64_t gw = static_cast<int64_t>(blockIdx.x) * kScanWarps + warp;
    const int64_t nw = static_cast<int64_t>(gridDim.x) * kScanWarps;
    const int head_slot = static_cast<int>(gw % kHeadSlots);
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
        {
            const uint64_t sb = s_bound;
            thr = sb > thr ? sb : thr;
        }
#pragma unroll
        for (int r = 0; r < kWideRows; ++r) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[r] += __shfl_xor_sync(kFullMask, acc[r], o);
            if (rows[r] >= 0) {  // warp-uniform
                const float sc = __ldg(a.scale + rows[r]);
                const float score = acc[r] * sc;
                const uint64_t key = pack_key(score, a.row_offset + static_cast<uint32_t>(rows[r]) * a.row_stride);
                if (sc == sc && key > thr) {  // NaN scale = deleted row
                    uint64_t kth;
                    if (warp_buf_append(my_list, cap, k, key, lane, cnt, kth) && kth > thr) thr = kth;
                    if (key > best) {
                        best = key;
                        if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(a.heads) + head_slot, key);
                    }
                }
            }
        }
        ++it;
        if (heads_refresh_turn(it, warp)) {
            const uint64_t hb = heads_bound_key(a.heads, k, lane);
            if (lane == 0 && hb != 0ull) atomicMax(&s_bound, hb);
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
    __syncwarp();
    for (int i = cnt + lane; i < cap; i += 32) my_list[i] = 0ull;
    publish_and_final_merge<1>(lists, kScanWarps, cap, k, a.partial, a.counter, a.out, selbuf, &s_cnt, s_red, a.px, a.seq,
                               a.q_base, a.heads, &s_bound, a.done_flag, a.done_value);
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
// bound32 (nullable): per query, the high word of a key known to be <= the k_out-th best (K3's shared threshold:
// the largest k-th best any CTA reached) -- drops almost every candidate before it touches shared memory.
__global__ void __launch_bounds__(256) k_merge_keys(const uint64_t* in, int G, int B, int k_in, int k_out,
                                                    uint64_t* out, const uint32_t* bound32) {
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
    if (bound32) {
        const uint64_t bk = static_cast<uint64_t>(bound32[b]) << 32;
        tau = bk > tau ? bk : tau;
    }
    auto src = [&](int i) { int g = i / k_in, j = i - g * k_in; return in[(static_cast<size_t>(g) * B + b) * k_in + j]; };
    block_select_topk<8>(src, G * k_in, k_out, tau, selbuf, &s_cnt);
    for (int j = threadIdx.x; j < k_out; j += blockDim.x) out[static_cast<size_t>(b) * k_out + j] = selbuf[j];
}

// Batched path: K3 leaves this GPU's top-k per query in `local` [nq][k]; a FEW CTAs (kExchangeCtas at most) do the
// fused NVLink exchange + cross-GPU merge, each for the queries b = blockIdx.x, blockIdx.x + gridDim.x, ...: all of a
// CTA's queries are published before it waits for any, so the waits cannot cycle, and only a handful of SMs sit in
// the spin-wait while a slower peer is still scanning.
constexpr int kExchangeCtas = 32;
__global__ void __launch_bounds__(256) k_exchange_merge(PeerExchange px, uint64_t seq, int q_base, const uint64_t* local,
                                                        int k, uint64_t* out, int nq) {
    __shared__ __align__(16) uint64_t selbuf[kSelCap];
    __shared__ int s_cnt;
    for (int b = blockIdx.x; b < nq; b += gridDim.x) {
        exchange_publish(px, seq, q_base + b, local + static_cast<size_t>(b) * k, k);
        __syncthreads();
    }
    for (int b = blockIdx.x; b < nq; b += gridDim.x) {
        exchange_wait_merge(px, seq, q_base + b, k, selbuf, &s_cnt);
        for (int j = threadIdx.x; j < k; j += blockDim.x) out[static_cast<size_t>(b) * k + j] = selbuf[j];
        __syncthreads();
    }
}
#endif  // CV_MAIN_TU

}  // namespace cv
