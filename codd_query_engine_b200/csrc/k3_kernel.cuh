// k3_kernel.cuh -- K3 `gemm_topk` (kernel; host side in k3_mma.cu, CTA-pair instantiations in k3_inst_cg2.cu): batched scoring on the 5th-gen tensor cores with a fused top-k epilogue.
//
// Replaces the batched form of Collection.query (call site
// codd_dal/metrics/metrics_semantic_metadata_store.py:314-316; chromadb accepts query batches) with an
// exact scan:  S[q][row] = (Qhat_bf16[q] . E_bf16[row]) * scale[row], fp32 accumulate in TMEM.
//
// Shape of the work.  The query tile is the UMMA *M* operand (128 queries = 128 TMEM lanes, resident
// in shared memory for the whole kernel up to 512-d), a tile of 128 corpus rows is the *N* operand (streamed once
// from HBM by TMA, K-major, SWIZZLE_128B, 64-element K blocks through an up-to-8-deep mbarrier ring), D is
// the short K dimension (Dp/16 tcgen05.mma per tile).  With queries on the lanes every epilogue
// thread owns ONE query: it keeps that query's running threshold in a register and its k-list
// privately, so the fused top-k needs no atomics, no shuffles and no shared memory -- a max-reduction of the raw
// dots per 32-column chunk against a slightly relaxed threshold, exact scaling only on the chunks that pass --
// and the B x N score matrix never leaves TMEM/registers.  The per-row scale
// and the namespace / label / bitmask predicate are folded into one shared-memory vector per tile
// (NaN = not eligible, so an ineligible score fails every comparison).
//
// Warp roles (224 threads with four epilogue warps, one CTA per SM, persistent): warp 0 = TMA producer (one elected
// lane), warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2..5 = epilogue (TMEM lane quarter
// = warp_id % 4), last warp = bound warp (keeps the cross-CTA lower bound of each duty query's k-th best fresh, off
// the epilogue's critical path).  All 512 TMEM columns are accumulator stages -- four of 128 columns for single
// CTAs, two of 256 for CTA pairs -- so the MMAs of the next tiles overlap the epilogue of tile t and a burst of list
// insertions does not stop the HBM stream.
//
// What keeps the fixed costs down on short scans (profiles/README.md, round 2): query prep as a programmatic
// dependent launch, rank-based list insertion, heads published once per tile, the bound warp, a two-pass first tile
// (PILOT instantiations), a per-query last-owner merge instead of a merge launch, self-resetting cross-CTA state;
// on long scans with several query super-tiles: their lockstep (HBM reads every tile once for all of them).
//
// Roofline: algorithmic bytes = rows * (2*Dp + 4) per query tile pass, flops =
// 2 * B * rows * Dp.  HBM-bound for B <~ 200, tensor-bound above.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/coddvec.h"
#include "common.cuh"
#include "k3_mma.cuh"
#include "vmm.hpp"

namespace cv {

namespace k3 {

constexpr int kQ = 128;          // queries per tile  (UMMA M, TMEM lanes)
constexpr int kTN = 128;         // corpus rows per tile (UMMA N, TMEM columns per accumulator stage)
constexpr int kKB = 64;          // K elements per shared-memory block = 128 bytes = one swizzle row
constexpr int kBlockBytes = kTN * kKB * 2;  // 16 KiB: one [128 rows][64] bf16 block (Q or corpus)
constexpr int kMaxStages = 8;    // corpus ring depth: as many 16 KiB stages as fit next to the query tile (<= 8)
constexpr int kMaxAccStages = 4;
// TMEM accumulator stages: all 512 columns either way -- four 128-column stages for single CTAs, two 256-column
// stages for CTA pairs.  The stages are the slack between the MMA stream and the epilogue: a burst of list insertions
// (every CTA's first tiles) holds the epilogue for several microseconds, and the HBM stream only keeps running through
// it while the MMA issuer still has a free accumulator to write.
__host__ __device__ constexpr int acc_stages(int cg) { return cg == 1 ? 4 : 2; }
// threads per CTA: producer warp + MMA warp + EW epilogue warps (4, or 8: two warps per TMEM lane quarter on alternate
// column chunks) + the bound warp
__host__ __device__ constexpr int mma_threads(int ew) { return 96 + 32 * ew; }
constexpr int kMaxKBlocks = 64;  // Dp <= 4096 (CV_MAX_DIM)

constexpr int kHeadsLd = 320;    // >= max list owners per query (148 CTAs x 2 column halves), multiple of 4

constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

struct KArgs {
    const float* scale;
    const uint32_t* ns;
    const uint32_t* label;
    const uint32_t* mask;
    uint32_t ns_code, label_code;
    int64_t n;
    int kblocks;
    int stages;     // ring depth actually used (<= kMaxStages)
    int q_resident; // 1: the query tile stays in shared memory for the whole kernel (short rows);
                    // 0: its K blocks are streamed through the ring next to the corpus blocks (long rows)
    int k;
    uint32_t row_offset;
    int n_tiles;
    int q_tiles;
    int P;          // CTAs sharing one query tile
    int Bpad;
    int B;            // real queries in this slice (lanes >= B are padding)
    int smem_lists;   // k > 32: 1 = the per-query lists live in shared memory during the scan
    uint64_t* lists;  // [P][Bpad][k]
    uint32_t* gthr;   // [Bpad] shared per-query lower bounds on the k-th best score (ordered-float encoding)
    uint32_t* heads;  // [Bpad][kHeadsLd] best score each CTA (slot p) has seen for the query, ordered floats, 0 = none
    float* dump;      // debug: [Bpad][dump_ld] scores, or nullptr
    int64_t dump_ld;
    uint64_t corpus_hint;
    uint32_t row_stride;   // global row = row_offset + local row * row_stride (1 = dense block, G = cyclic shard)
    int scale_bounded;     // 1: every row scale is 0, NaN or within 2^-7 of 1 (enables the raw-dot prefilter)
    // fused query preparation (q_resident only): the CTA normalises its 128 raw queries itself and writes the bf16
    // tile (and, with `split`, the bf16 residual tile) straight into the swizzled shared-memory layout
    const float* q_raw;    // [B][D] fp32 raw queries of this slice, or nullptr (the tile comes from tmap_q)
    int D;
    int cosine;
    int split;             // 1: Q = hi + lo (two bf16 tiles, two MMAs per K step) -> fp32-accurate scores
    // fused tail (k <= 64): survivors above the cross-CTA bound are appended to cand[q][...]; the last CTA of a
    // query-tile group merges them, writes out[] and resets gthr / heads / cnt / done for the next launch
    int fused_tail;
    uint64_t* cand;        // [Bpad][cand_ld]
    int cand_ld;           // P * k
    uint32_t* cnt;         // [Bpad] survivors appended per query
    uint32_t* done;        // [Bpad] list owners of the query that have published (the last one merges)
    uint64_t* out;         // [B][k] final keys of this slice
    // lockstep of the query super-tiles that read the same corpus tiles (q_tiles > 1): one progress word per unit
    unsigned long long* prog;  // [q_tiles * P]: (epoch << 32) | tiles issued; nullptr = off
    uint32_t epoch;            // this launch's tag
    int lock_window;           // a unit issues tile i only once every sibling has issued tile i - lock_window
    unsigned long long* dbg;  // nullptr, or 4 words (atomicMax): prep ns, publish ns, merge ns, survivors per query
    unsigned long long* tl;   // nullptr, or [gridDim.x][16] globaltimer stamps / wait totals of this CTA's phases (CODDVEC_K3_TIMELINE)
};

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "CV_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra CV_DONE;\n\t"
        "bra CV_WAIT;\n\t"
        "CV_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// CG = 2: issued by both CTAs of the pair; the peer bit of the barrier address is cleared so that the
// transaction bytes are counted on the LEADER's barrier (the MMA issuer waits there for both halves).
template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, uint64_t hint) {
    if (CG == 1) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
            " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
            : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
            " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1), "l"(hint)
            : "memory");
    }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 remote;\n\t"
        "mapa.shared::cluster.u32 remote, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [remote];\n\t"
        "}\n" ::"r"(bar), "r"(cta)
        : "memory");
}

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
    if (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> fp32, issued by one thread for the CTA
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (CG == 1) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}\n" ::"r"(tmem_d),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// arrive on an mbarrier when every tcgen05.mma issued so far by this thread has completed
// (CG = 2: on the barrier at this offset in BOTH CTAs of the pair)
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CG == 1) {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    } else {
        const uint16_t mask = 3;
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                     "h"(mask)
                     : "memory");
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B, rows at a 128-byte pitch:
//   bits [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1),
//   [32,46) stride byte offset >> 4 (8 rows * 128 B = 1024 B between 8-row core-matrix groups),
//   [46,48) descriptor version = 1 (sm_100), [61,64) layout type = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}
// Instruction descriptor for kind::f16: fp32 accumulate (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15, 16 = 0), N >> 3 at bits 17-22, M >> 4 at bits 24-28.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// Fire-and-forget maximum (RED, not ATOM: the issuing warp must not wait an L2 round trip per published threshold --
// nvcc turned the value-less atomicMax of this kernel into a returning ATOMG, which cost 35-100 us per batch).
__device__ __forceinline__ void red_max_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// order-preserving float <-> uint32 (same mapping as the high word of a result key); 0 = "nothing yet"
__device__ __forceinline__ uint32_t ord_of(float f) {
    uint32_t u = __float_as_uint(f + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_to_float(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o);
}
// lower bound published by other CTAs -> strict threshold: s > result  <=>  s >= published score.
// (The predecessor of +0.0 in the ordered encoding is -0.0, which compares EQUAL to +0.0 as a float and
// would wrongly reject exact-zero ties, so it is skipped.)
__device__ __forceinline__ float shared_threshold(uint32_t g) {
    if (g <= 1u) return -INFINITY;
    uint32_t o = g - 1u;
    if (o == 0x7FFFFFFFu) o = 0x7FFFFFFEu;
    return ord_to_float(o);
}

// ---- cross-CTA "heads" bound (register lists, k <= 64) ----------------------------------------------------
// Every list owner scanning rows for query q publishes the BEST score it has seen in heads[q][owner] (a plain store
// to its own slot, once per tile).  The k-th largest of those heads is a lower bound on the global k-th best score
// (k different owners each hold a distinct row at least that good) and -- because the global top-k rows are almost
// surely spread over different CTAs -- a tight one.  It is computed by the bound warp of the query's duty CTA
// (k_mma_topk) and travels to every owner through gthr[q].
// Whole warp: the k-th LARGEST of the 32 * J values the lanes hold in v[] (destroyed), 0 if fewer than k are non-zero.
// k rounds of "warp maximum, one holder drops one copy" -- ~20 instructions per round, no memory traffic.
template <int J>
__device__ __forceinline__ uint32_t warp_kth_largest(uint32_t (&v)[J], int k, int lane) {
    uint32_t M = 0u;
#pragma unroll 1
    for (int i = 0; i < k; ++i) {
        uint32_t m = v[0];
#pragma unroll
        for (int j = 1; j < J; ++j) m = max(m, v[j]);
        M = __reduce_max_sync(kFullMask, m);
        if (M == 0u) return 0u;
        const uint32_t holders = __ballot_sync(kFullMask, m == M);
        if (lane == __ffs(holders) - 1) {
            bool dropped = false;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const bool hit = !dropped && v[j] == M;
                v[j] = hit ? 0u : v[j];
                dropped |= hit;
            }
        }
    }
    return M;
}

// Register-resident descending list, no memory traffic.  Short lists insert by RANK: every slot compares the key with
// its own and its left neighbour's old value (independent compares and selects, dependency depth ~4), where a bubble
// pass chains KREG compare-exchanges (depth ~4 * KREG) -- and a lone epilogue warp has nothing to hide that latency
// behind: every CTA's first tile is a burst of ~k * (1 + ln(128 / k)) insertions per query, 18 us with the bubble.
// Long lists keep the bubble (the rank form needs old and new values live at once).
template <int KREG>
__device__ __forceinline__ void list_insert_regs(uint64_t (&L)[KREG], uint64_t key) {
    if (KREG <= 16) {
        bool gt[KREG];
#pragma unroll
        for (int j = 0; j < KREG; ++j) gt[j] = key > L[j];  // monotone: false ... false true ... true
#pragma unroll
        for (int j = KREG - 1; j > 0; --j) L[j] = gt[j] ? (gt[j - 1] ? L[j - 1] : key) : L[j];
        L[0] = gt[0] ? key : L[0];
    } else {
        uint64_t cur = key;
#pragma unroll
        for (int j = 0; j < KREG; ++j) {
            const bool g = cur > L[j];
            const uint64_t hi = g ? cur : L[j];
            cur = g ? L[j] : cur;
            L[j] = hi;
        }
    }
}

// v[j] for a runtime j without spilling v to local memory: a 31-select binary tree
__device__ __forceinline__ uint32_t pick32(const uint32_t (&v)[32], int j) {
    uint32_t a[16], b[8], c[4], d[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = (j & 2) ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = (j & 4) ? b[2 * i + 1] : b[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) d[i] = (j & 8) ? c[2 * i + 1] : c[2 * i];
    return (j & 16) ? d[1] : d[0];
}

// ------------------------------------------------------------------------------------------ query prep
// Normalise fp32 queries exactly like the oracle (fp64 norm, fp32 divide), round to bf16, zero-pad to
// [Bpad][Dp].  One warp per query row.  Only the long-row mode (query tile streamed through the ring by TMA)
// needs this launch; short rows build their tile inside k_mma_topk (prep_query_tile).
static __global__ void __launch_bounds__(128) k_prep_queries(const float* q, int B, int Bpad, int D, int Dp, int cosine,
                                                      __nv_bfloat16* out) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
    // Programmatic dependent launch: k_mma_topk may start now (its prologue, tensor-map prefetch and first corpus
    // loads do not depend on this kernel); it executes griddepcontrol.wait before it fetches the query tile.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (row >= Bpad) return;
    __nv_bfloat16* o = out + static_cast<size_t>(row) * Dp;
    if (row >= B) {
        for (int d = lane; d < Dp; d += 32) o[d] = __float2bfloat16_rn(0.0f);
        return;
    }
    const float* r = q + static_cast<size_t>(row) * D;
    float safe = 1.0f;
    if (cosine) {
        double ss = 0.0;
        for (int d = lane; d < D; d += 32) {
            double v = static_cast<double>(r[d]);
            ss += v * v;
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) ss += __shfl_xor_sync(kFullMask, ss, o2);
        float nrm = static_cast<float>(sqrt(ss));
        safe = nrm > 0.0f ? nrm : 1.0f;
    }
    for (int d = lane; d < Dp; d += 32) o[d] = __float2bfloat16_rn(d < D ? __fdiv_rn(r[d], safe) : 0.0f);
}

// Fused query preparation: the whole CTA turns its kQ raw fp32 queries into the bf16 UMMA operand tile in
// shared memory -- normalised exactly like the oracle (fp64 norm, fp32 divide), rounded to bf16, laid out as the
// TMA would have written it: K block kb = 16 KiB of [128 rows][64 elements], 128-byte rows, 16-byte chunk cc of
// row r stored at chunk position cc ^ (r & 7) (SWIZZLE_128B; block bases are 1024-byte aligned).  With `split`
// a second tile holds bf16(qhat - hi): hi + lo carries 16 mantissa bits, so the two MMAs per K step reproduce
// the fp32 query to ~2^-17 relative.  Rows past B and elements past D are zero.  One warp per row; lane l owns
// the 16-byte chunks l and l + 32 (rows of up to 512 elements stay resident, see mma_enqueue).
template <int THREADS>
__device__ __forceinline__ void prep_query_tile(unsigned char* sm_q, const float* q_raw, int q_row0, int B, int D,
                                                int kblocks, int cosine, int split) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarps = THREADS / 32;
    const int nchunks = kblocks * 8;
    const bool vec = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(q_raw) & 15) == 0;
    const int live_rows = max(0, min(kQ, B - q_row0));
    // padding rows: plain zero stores, no arithmetic
    for (int i = live_rows * nchunks + threadIdx.x; i < kQ * nchunks; i += THREADS) {
        const int r = i / nchunks, c = i - r * nchunks;
        unsigned char* dst = sm_q + static_cast<size_t>(c >> 3) * kBlockBytes + r * 128 + (((c & 7) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
        if (split) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(kblocks) * kBlockBytes) = make_uint4(0u, 0u, 0u, 0u);
    }
    // live rows: kRows rows per warp iteration, all their loads issued before any arithmetic
    constexpr int kRows = 4;
    for (int r0 = warp * kRows; r0 < live_rows; r0 += kWarps * kRows) {
        float x[kRows][2][8];
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
            const bool live = r0 + i < live_rows;  // warp-uniform
            const float* src = q_raw + static_cast<size_t>(q_row0 + (live ? r0 + i : 0)) * D;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int d0 = (lane + 32 * j) * 8;
                if (live && vec && d0 + 8 <= D) {
                    const float4 u = __ldg(reinterpret_cast<const float4*>(src + d0));
                    const float4 w = __ldg(reinterpret_cast<const float4*>(src + d0 + 4));
                    x[i][j][0] = u.x; x[i][j][1] = u.y; x[i][j][2] = u.z; x[i][j][3] = u.w;
                    x[i][j][4] = w.x; x[i][j][5] = w.y; x[i][j][6] = w.z; x[i][j][7] = w.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[i][j][e] = (live && d0 + e < D) ? __ldg(src + d0 + e) : 0.0f;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kRows; ++i) {
            const int r = r0 + i;
            if (r >= live_rows) break;  // warp-uniform
            float safe = 1.0f;
            if (cosine) {
                // ||q|| as the oracle computes it -- fp32(sqrt(sum of squares in fp64)) -- without the FP64 pipe (a
                // trickle on this part: 128 rows took 34 us): the sum is carried as an unevaluated fp32 pair
                // (error-free products and sums: ~2^-45 relative), the root as y + c with y = sqrtf(hi) and one
                // Newton correction from the exact residual.  Equal to the fp64 result except when it lies within
                // 2^-21 ulp of a rounding boundary.
                float hi = 0.0f, lo = 0.0f;
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float v = x[i][j][e];
                        const float pr = v * v, pe = fmaf(v, v, -pr);      // v*v = pr + pe exactly
                        const float sum = hi + pr, bb = sum - hi;
                        lo += ((hi - (sum - bb)) + (pr - bb)) + pe;         // TwoSum error + product error
                        hi = sum;
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ohi = __shfl_xor_sync(kFullMask, hi, o), olo = __shfl_xor_sync(kFullMask, lo, o);
                    const float sum = hi + ohi, bb = sum - hi;
                    lo = (lo + olo) + ((hi - (sum - bb)) + (ohi - bb));
                    hi = sum;
                }
                const float sh = hi + lo, sl = lo - (sh - hi);              // renormalised pair
                float nrm = 0.0f;
                if (sh > 0.0f) {
                    const float y = sqrtf(sh);
                    const float resid = fmaf(-y, y, sh) + sl;               // sh + sl - y*y (the fma is exact)
                    nrm = y + __fdiv_rn(resid, 2.0f * y);
                }
                safe = nrm > 0.0f ? nrm : 1.0f;
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = lane + 32 * j;
                if (c < nchunks) {
                    CV_ASSERT(r >= 0 && r < kQ && (c >> 3) < kblocks);
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // (x == 0 -- padding past D, sparse vectors -- would send the IEEE division down its slow path)
                        const float a0 = x[i][j][2 * e], a1 = x[i][j][2 * e + 1];
                        const float v0 = a0 == 0.0f ? 0.0f : __fdiv_rn(a0, safe), v1 = a1 == 0.0f ? 0.0f : __fdiv_rn(a1, safe);
                        const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                        hi[e] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
                        const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
                        const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
                        lo[e] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
                    }
                    const int kb = c >> 3, cc = c & 7;
                    unsigned char* dst = sm_q + static_cast<size_t>(kb) * kBlockBytes + r * 128 + ((cc ^ (r & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    if (split)
                        *reinterpret_cast<uint4*>(dst + static_cast<size_t>(kblocks) * kBlockBytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
    }
    fence_proxy_async();  // generic-proxy stores above -> visible to the tensor core's async-proxy reads
}

// ------------------------------------------------------------------------------------------ main kernel
// CG = 1: one CTA per SM works alone (M = 128 queries, N = 128 corpus rows per tile).
// CG = 2: a CTA pair (cluster of 2, `cta_group::2`) shares every MMA: M = 256 queries (128 per CTA),
//         N = 256 corpus rows per tile of which each CTA stages 128 -- so each SM reads half the B
//         operand from its shared memory and L2 serves each corpus tile once per 256 queries.
//         Only the leader CTA (cluster rank 0) issues MMAs; both CTAs' TMA loads complete on the
//         leader's `full` barrier; `tcgen05.commit ... multicast` releases ring slots / publishes
//         accumulators in both CTAs; the peer's epilogue warps arrive remotely on the leader's
//         `tmem_empty` barrier.
// EW = epilogue warps: 4 (one per TMEM lane quarter; thread = query) or 8 (two per quarter: the threads of warps w and
//      w + 4 share a query and take alternate 32-column chunks of every tile, each with its own top-k list -- twice
//      the issue slots and latency hiding for the score scan, which is what bounds large batches).
// PILOT = two-pass first tile (see the epilogue): its own instantiations, so that the kernels that do not use it --
//      CTA pairs, long scans -- keep exactly the code they had (the mere presence of the branch in the tile loop cost
//      them 2-5 % in same-box A/B runs).
template <bool DUMP, int KREG, int CG, int EW, bool PILOT = false>
__global__ void __launch_bounds__(mma_threads(EW), 1)
k_mma_topk(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_e, const KArgs a) {
    constexpr int TN = kTN * CG;                       // corpus rows (= accumulator columns) per tile
    constexpr int kEpi = 32 * EW;                      // epilogue threads
    constexpr int kHalves = EW / 4;                    // list owners per (CTA, query)
    constexpr int kAccStages = acc_stages(CG);
    constexpr uint32_t kCols = kAccStages * TN;        // TMEM columns: 512
    extern __shared__ __align__(1024) unsigned char smem[];
    // carve-up (all block bases 1024-byte aligned, as SWIZZLE_128B requires)
    // q_resident: [query tile: kblocks (x2 when split) x 16 KiB][ring: stages x 16 KiB corpus blocks]
    // streamed:   [ring: stages x (16 KiB query block + 16 KiB corpus block)]
    const bool qres = a.q_resident != 0;
    const int kStages = a.stages;
    const uint32_t stage_bytes = qres ? kBlockBytes : 2 * kBlockBytes;
    unsigned char* sm_q = smem;
    const int qblocks = a.kblocks * (a.split ? 2 : 1);  // resident query tile: hi blocks, then the residual (lo) blocks
    unsigned char* sm_ring = smem + (qres ? static_cast<size_t>(qblocks) * kBlockBytes : 0);
    float* sm_scale = reinterpret_cast<float*>(sm_ring + static_cast<size_t>(kStages) * stage_bytes);  // [2][TN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_scale + kAccStages * TN);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 1 + 2 * kMaxAccStages);
    uint64_t* sm_lists = reinterpret_cast<uint64_t*>(tmem_slot + 4);  // [128][k] when a.smem_lists

    const uint32_t bar_full = smem_u32(bars);                  // [kMaxStages] (leader's are the live ones)
    const uint32_t bar_empty = bar_full + 8 * kMaxStages;      // [kMaxStages]
    const uint32_t bar_q = bar_empty + 8 * kMaxStages;         // [1]
    const uint32_t bar_tfull = bar_q + 8;                      // [kAccStages]
    const uint32_t bar_tempty = bar_tfull + 8 * kMaxAccStages; // [kAccStages] (leader's are the live ones)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // (phase stamps exist only in the diagnostic build, -DCV_K3_TIMELINE: `tl` is a compile-time nullptr otherwise)
#ifdef CV_K3_TIMELINE
    unsigned long long* const tl = a.tl ? a.tl + static_cast<size_t>(blockIdx.x) * 16 : nullptr;
#else
    constexpr unsigned long long* tl = nullptr;
#endif
    if (tl && threadIdx.x == 0) tl[0] = global_timer_ns();     // 0: CTA entry
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;   // 0 = leader
    const int unit = blockIdx.x / CG;                         // CTA (CG=1) or CTA pair (CG=2)
    const int qt = unit % a.q_tiles;                          // query tile of 128*CG queries
    const int p = unit / a.q_tiles;

    if (threadIdx.x == 0) {
        tmem_slot[1] = 0u;  // scan_done: the epilogue has scanned its last tile (-> bound warp)
        tmem_slot[2] = 0u;  // bound_ack: the bound warp has published its last bound (-> epilogue, before the tail)
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, CG);      // own arrive.expect_tx (+ the peer's remote arrive)
            mbar_init(bar_empty + 8 * s, 1);      // one tcgen05.commit
        }
        mbar_init(bar_q, CG);
        for (int s = 0; s < kAccStages; ++s) {
            mbar_init(bar_tfull + 8 * s, 1);
            mbar_init(bar_tempty + 8 * s, EW * CG);  // one arrival per epilogue warp of every CTA in the group
        }
        fence_barrier_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<CG>(smem_u32(tmem_slot), kCols);
    const bool fused_prep = a.q_raw != nullptr;  // (q_resident only)
    if (fused_prep) {
        const uint64_t t0 = a.dbg ? global_timer_ns() : 0;
        prep_query_tile<mma_threads(EW)>(sm_q, a.q_raw, (qt * CG + static_cast<int>(rank)) * kQ, a.B, a.D, a.kblocks, a.cosine, a.split);
        if (a.dbg && (threadIdx.x & 31) == 0) atomicMax(a.dbg, global_timer_ns() - t0);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync(); else __syncthreads();  // barriers, TMEM address and both CTAs' query tiles are ready
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tl && threadIdx.x == 0) tl[1] = global_timer_ns();     // 1: prologue (barriers, TMEM, fused query prep) done

    if (warp == 0) {
        // ===================================================================== TMA producer (every CTA)
        if (lane == 0) {
            if (!fused_prep) tma_prefetch_desc(&tmap_q);
            tma_prefetch_desc(&tmap_e);
            const int q_row = (qt * CG + static_cast<int>(rank)) * kQ;
            // The resident query tile comes from k_prep_queries, the previous launch in the stream.  This kernel is
            // launched as its programmatic dependent, so everything up to here -- and the first ring-full of corpus
            // blocks -- overlaps it; the tile itself is fetched after griddepcontrol.wait (prep complete and visible).
            bool q_pending = qres && !fused_prep;
            auto fetch_query_tile = [&]() {
                asm volatile("griddepcontrol.wait;" ::: "memory");
                const uint32_t q_bytes = static_cast<uint32_t>(a.kblocks) * kBlockBytes;
                if (rank == 0) mbar_arrive_expect_tx(bar_q, q_bytes * CG);
                else mbar_arrive_cluster(bar_q, 0);
                for (int kb = 0; kb < a.kblocks; ++kb)
                    tma_load_2d<CG>(smem_u32(sm_q + static_cast<size_t>(kb) * kBlockBytes), &tmap_q, bar_q, kb * kKB, q_row,
                                    kEvictLast);
                q_pending = false;
            };
            if (!qres && !fused_prep) asm volatile("griddepcontrol.wait;" ::: "memory");  // streamed query blocks ride in the ring
            // Lockstep of sibling super-tiles: the q_tiles units with the same p walk the SAME corpus tiles, and L2 serves
            // a tile to all of them only while they stay close together.  Left alone they drift apart (different
            // candidate loads) and on a long scan HBM ends up reading the corpus 1.4-2x (ncu, 12.5M rows x batch 1024;
            // 1.00x at 1M rows).  So every leader publishes how many tiles it has issued and does not run more than
            // lock_window tiles ahead of its slowest sibling.  Stale words carry an older launch tag and are ignored;
            // the wait is bounded (a sibling that is not resident yet must not stall this unit for good).
            const bool lockstep = a.prog != nullptr && a.q_tiles > 1 && rank == 0;
            unsigned long long* const my_prog = a.prog + unit;
            const unsigned long long* const sib_prog = a.prog + static_cast<size_t>(p) * a.q_tiles;
            const unsigned long long tag = static_cast<unsigned long long>(a.epoch) << 32;
            uint32_t tiles_issued = 0;
            bool lock_off = false;
            // (the siblings' words are read one tile AHEAD of their use: the loads are in flight while this tile's TMA
            // requests go out, so the producer never sits on an L2 round trip unless it really has to wait; an older
            // value only understates a sibling's progress.  Up to 4 super-tiles -- batches of up to 1024 queries -- keep
            // their 3 words in registers; wider batches check synchronously every 4th tile.)
            constexpr int kSib = 3;
            unsigned long long seen[kSib] = {0ull, 0ull, 0ull};
            const bool few = a.q_tiles <= kSib + 1;
            auto sibling = [&](int j) { return j < qt ? j : j + 1; };  // j-th sibling's super-tile index
            auto progress_of = [&](unsigned long long w) { return (w >> 32) == a.epoch ? static_cast<uint32_t>(w) : 0u; };
            int stage = 0, issued = 0;
            uint32_t phase = 0;
            for (int t = p; t < a.n_tiles; t += a.P) {
                if (lockstep) {
                    st_relaxed_gpu_u64(my_prog, tag | tiles_issued);
                    if (!lock_off && tiles_issued > static_cast<uint32_t>(a.lock_window) && (few || (tiles_issued & 3u) == 0u)) {
                        uint64_t t_wait = 0;
                        bool fresh = !few;
                        while (true) {
                            uint32_t slowest = 0xFFFFFFFFu;
                            if (few) {
#pragma unroll
                                for (int j = 0; j < kSib; ++j) {
                                    if (j < a.q_tiles - 1) {
                                        if (fresh) seen[j] = ld_relaxed_gpu_u64(sib_prog + sibling(j));
                                        slowest = min(slowest, progress_of(seen[j]));
                                    }
                                }
                            } else {
                                for (int s = 0; s < a.q_tiles; ++s)
                                    if (s != qt) slowest = min(slowest, progress_of(ld_relaxed_gpu_u64(sib_prog + s)));
                            }
                            if (tiles_issued <= slowest + static_cast<uint32_t>(a.lock_window)) break;
                            fresh = true;  // (the prefetched words may just be old: look again before waiting)
                            const uint64_t now = global_timer_ns();
                            if (t_wait == 0) t_wait = now;
                            else if (now - t_wait > 30000) { lock_off = true; break; }  // 30 us: give up for this launch
                        }
                    }
                    if (few) {
#pragma unroll
                        for (int j = 0; j < kSib; ++j)
                            if (j < a.q_tiles - 1) seen[j] = ld_relaxed_gpu_u64(sib_prog + sibling(j));  // for the next tile
                    }
                    ++tiles_issued;
                }
                for (int kb = 0; kb < a.kblocks; ++kb) {
                    if (q_pending && issued == kStages) fetch_query_tile();  // the ring is full of requests: now the query tile
                    ++issued;
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, stage_bytes * CG);
                    else mbar_arrive_cluster(bar_full + 8 * stage, 0);
                    unsigned char* slot = sm_ring + static_cast<size_t>(stage) * stage_bytes;
                    if (!qres) {  // this K block of the query tile rides along (L2 resident, evict-last)
                        tma_load_2d<CG>(smem_u32(slot), &tmap_q, bar_full + 8 * stage, kb * kKB, q_row, kEvictLast);
                        slot += kBlockBytes;
                    }
                    tma_load_2d<CG>(smem_u32(slot), &tmap_e, bar_full + 8 * stage, kb * kKB,
                                    t * TN + static_cast<int>(rank) * kTN, a.corpus_hint);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
            if (q_pending) fetch_query_tile();
            if (lockstep) st_relaxed_gpu_u64(my_prog, tag | 0xFFFFFFFFull);  // done: nobody waits for this unit any more
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA only)
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(kQ * CG, TN);
            if (qres && !fused_prep) {
                mbar_wait(bar_q, 0);
                tc_fence_after();
            }
            int stage = 0, as = 0;
            uint32_t phase = 0, aphase = 0;
            uint64_t w_empty = 0, w_full = 0;  // (timeline only) ns spent waiting for the epilogue / for TMA data
            for (int t = p; t < a.n_tiles; t += a.P) {
                const uint64_t tw0 = tl ? global_timer_ns() : 0;
                mbar_wait(bar_tempty + 8 * as, aphase ^ 1);  // every epilogue warp has drained this accumulator stage
                tc_fence_after();
                const uint64_t tw1 = tl ? global_timer_ns() : 0;
                w_empty += tw1 - tw0;
                const int tidx = (t - p) / a.P;
                unsigned long long* const tr = (tl && blockIdx.x == 0 && tidx < 64) ? a.tl + 16 * 296 + tidx * 6 : nullptr;
                if (tr) { tr[0] = tw0; tr[1] = tw1; }
                const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * TN);
                for (int kb = 0; kb < a.kblocks; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);  // TMA bytes of every CTA in the group have landed
                    tc_fence_after();
                    if (tl && t == p && kb == 0) tl[2] = global_timer_ns();  // 2: first corpus block landed
                    const uint32_t slot = smem_u32(sm_ring + static_cast<size_t>(stage) * stage_bytes);
                    const uint32_t a_addr = qres ? smem_u32(sm_q + static_cast<size_t>(kb) * kBlockBytes) : slot;
                    const uint32_t b_addr = qres ? slot : slot + kBlockBytes;
#pragma unroll
                    for (int ks = 0; ks < kKB / 16; ++ks)  // UMMA_K = 16 bf16 = 32 bytes inside the swizzle row
                        umma_bf16<CG>(tmem_d, umma_desc_sw128(a_addr + ks * 32), umma_desc_sw128(b_addr + ks * 32), idesc,
                                      (kb | ks) != 0 ? 1u : 0u);
                    if (a.split) {  // + (residual query tile) x (the same corpus block)
                        const uint32_t lo_addr = a_addr + static_cast<uint32_t>(a.kblocks) * kBlockBytes;
#pragma unroll
                        for (int ks = 0; ks < kKB / 16; ++ks)
                            umma_bf16<CG>(tmem_d, umma_desc_sw128(lo_addr + ks * 32), umma_desc_sw128(b_addr + ks * 32), idesc, 1u);
                    }
                    umma_commit<CG>(bar_empty + 8 * stage);  // frees the ring slot (in both CTAs) once read
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit<CG>(bar_tfull + 8 * as);  // accumulator complete -> epilogue warps (of both CTAs)
                if (++as == kAccStages) { as = 0; aphase ^= 1; }
                if (tl) w_full += global_timer_ns() - tw1;  // (issue time is negligible: this is the wait for TMA bytes)
                if (tr) tr[2] = global_timer_ns();
            }
            if (tl) {
                tl[3] = global_timer_ns();                     // 3: last MMA issued
                tl[8] = w_empty;                               // 8: MMA issuer waited for a free accumulator (epilogue-bound)
                tl[9] = w_full;                                // 9: MMA issuer waited for corpus blocks (HBM-bound)
            }
        }
    } else if (warp < 2 + EW) {
        // ===================================================================== epilogue (32 * EW threads per CTA)
        const int et = threadIdx.x - 64;                 // 0 .. kEpi - 1
        const int half = (warp - 2) >> 2;                // which alternate column chunks this thread scans (EW = 8)
        const int qlane = 32 * (warp & 3) + lane;        // TMEM lane = query within this CTA's 128
        const int qglobal = (qt * CG + static_cast<int>(rank)) * kQ + qlane;
        const int owner = half * a.P + p;                // this thread's slot among the lists / heads of its query
        const int owners = a.P * kHalves;
        CV_ASSERT(qglobal < a.Bpad && owner < owners && kHalves * ((a.P + 3) / 4 * 4) <= kHeadsLd && a.k <= (KREG > 0 ? KREG : kMaxK));
        const bool fused_tail = KREG > 0 && a.fused_tail != 0;
        uint64_t* const glist = fused_tail ? nullptr : a.lists + (static_cast<size_t>(owner) * a.Bpad + qglobal) * a.k;
        // large k (KREG == 0): the working list is in shared memory when it fits (fast warp-cooperative
        // insertion), else directly in global memory; lanes of a warp own consecutive lists either way
        uint64_t* list = (KREG == 0 && a.smem_lists) ? sm_lists + static_cast<size_t>(qlane) * a.k : glist;
        constexpr int KR = KREG > 0 ? KREG : 1;
        // k <= KREG: the list lives in registers, right-aligned -- slots [KREG-k, KREG) hold the top-k
        // descending, the slots before them are never-beaten sentinels -- so the k-th best is always
        // L[KREG-1] (a compile-time register, no dynamic indexing).
        uint64_t L[KR];
#pragma unroll
        for (int j = 0; j < KR; ++j) L[j] = (j < KR - a.k) ? ~0ull : 0ull;
        if (KREG == 0) {
            for (int j = 0; j < a.k; ++j) list[j] = 0ull;
            __syncwarp();
        }
        const bool live = qglobal < a.B;           // padding lanes never produce candidates
        const bool share = owners > 1 && live;     // other list owners scan other rows for the same query: trade thresholds
        float thr_local = live ? -INFINITY : INFINITY;
        float thr = thr_local;
        uint32_t g_seen = 0u;                      // largest shared lower bound (ordered float) folded into thr so far
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(32 * (warp & 3)) << 16);
        const bool pred = a.mask != nullptr || a.ns_code != kAny || a.label_code != kAny;
        // Unfiltered scans compare the RAW dot product against a slightly relaxed threshold first: row scales are 0,
        // NaN (deleted) or within 2^-7 of 1 (K1 normalises; the host verifies loaded snapshots), so
        // dot * scale > thr implies dot > thr - |thr| / 64, and a 32-column chunk costs a max-reduction instead of 32
        // multiplies -- the scales are only applied to the few chunks that pass.
        const bool prefilter = !pred && a.scale_bounded != 0 && !DUMP;
        int as = 0;
        uint32_t aphase = 0;
        // scale * eligibility of one tile row (NaN = skip); loaded one tile ahead so that the L2/HBM latency
        // of these small loads is never on the epilogue's critical path
        auto row_scale = [&](int t, int col) -> float {
            const int64_t row = static_cast<int64_t>(t) * TN + col;
            float sc = __int_as_float(0x7FC00000);
            if (t < a.n_tiles && row < a.n) {
                bool ok = true;
                if (pred) {
                    if (a.ns_code != kAny) ok = __ldg(a.ns + row) == a.ns_code;
                    if (ok && a.label_code != kAny) ok = __ldg(a.label + row) == a.label_code;
                    if (ok && a.mask) ok = (__ldg(a.mask + (row >> 5)) >> (row & 31)) & 1u;
                }
                if (ok) sc = __ldg(a.scale + row);
            }
            return sc;
        };
        constexpr int kScPer = (TN + kEpi - 1) / kEpi;   // tile rows whose scale each epilogue thread stages
        uint32_t* const head_row = a.heads + static_cast<size_t>(qglobal) * kHeadsLd;
        float best_local = -INFINITY;  // best score this thread has inserted (its head for this query)
        uint32_t g_pub = 0u;           // the last k-th best this thread published to gthr (ordered float)
        int tcount = 0;
        float sc_next[kScPer];
#pragma unroll
        for (int h = 0; h < kScPer; ++h) sc_next[h] = (h * kEpi + et < TN) ? row_scale(p, h * kEpi + et) : 0.0f;
        uint32_t g_next = share ? __ldcg(a.gthr + qglobal) : 0u;
        auto fold_shared = [&](uint32_t g) {  // a lower bound on the global k-th best published by other list owners
            if (g > g_seen) { g_seen = g; thr = fmaxf(thr, shared_threshold(g)); }
        };
        uint64_t e_top = 0, e_wait = 0, e_scan = 0;  // (timeline only) ns per phase of the epilogue loop, thread 0
        for (int t = p; t < a.n_tiles; t += a.P) {
            const uint64_t te0 = tl ? global_timer_ns() : 0;
#pragma unroll
            for (int h = 0; h < kScPer; ++h)
                if (h * kEpi + et < TN) sm_scale[as * TN + h * kEpi + et] = sc_next[h];
            if (share) fold_shared(g_next);
            ++tcount;
            // prefetch for the next tile of this CTA (in flight while this tile is processed)
#pragma unroll
            for (int h = 0; h < kScPer; ++h) sc_next[h] = (h * kEpi + et < TN) ? row_scale(t + a.P, h * kEpi + et) : 0.0f;
            if (share) g_next = __ldcg(a.gthr + qglobal);
            asm volatile("bar.sync 1, %0;" ::"n"(kEpi) : "memory");  // this CTA's epilogue warps only
            const uint64_t te1 = tl ? global_timer_ns() : 0;
            mbar_wait(bar_tfull + 8 * as, aphase);
            tc_fence_after();
            const uint64_t te2 = tl ? global_timer_ns() : 0;
            if (tl && et == 0 && t == p) tl[4] = te2;          // 4: first accumulator ready
            const float* sc_tile = sm_scale + as * TN;
            const uint32_t lrow0 = static_cast<uint32_t>(t) * TN;  // local row of the tile's first column
            // ---- pilot (first tile only): every CTA starts with empty lists, and converging a list from -inf costs
            // ~k * (1 + ln(columns / k)) insertions per query -- one warp per 32 queries, ~0.25 us per warp-wide
            // insertion round: 18 us on a 128-column tile, 26 us on a 256-column one (per-tile trace,
            // profiles/README.md), far beyond what the accumulator stages can hide.  So the first tile is read twice.
            // Pass A only takes each query's maximum over the tile and publishes it as this owner's head; the duty
            // CTA's bound warp answers with the k-th largest of ALL owners' first heads (rank ~k among P x 128 rows
            // instead of among 128); pass B is the normal scan under that bound and inserts almost nothing.  The wait
            // (~3 us: two L2 round trips on either side) is bounded, and the MMA issuer keeps the stream running into
            // the free accumulator stages meanwhile.
            if (PILOT && KREG > 0 && !DUMP && tcount == 1 && owners > 1) {  // (warp-uniform: tcgen05.ld is collective)
                float mx = -INFINITY;
#pragma unroll 1
                for (int c = half; c < TN / 32; c += kHalves) {
                    uint32_t v[32];
                    tmem_ld32(lane_base + static_cast<uint32_t>(as * TN + c * 32), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]) * sc_tile[c * 32 + j]);  // (NaN = skip)
                }
                if (share) {
                    best_local = mx;
                    __stcg(head_row + half * ((a.P + 3) / 4 * 4) + p, ord_of(mx));
                    uint32_t g = 0u;
#pragma unroll 1
                    for (int tries = 0; tries < 12 && g == 0u; ++tries) {
                        if (tries) __nanosleep(250);
                        g = __ldcg(a.gthr + qglobal);
                    }
                    fold_shared(g);
                }
                __syncwarp();
            }
            // Cross-CTA publications happen ONCE per tile, after its scan (this owner's head if it rose, its k-th best
            // if that can still raise the shared bound): a store and at most one RED per query and tile.
            bool tile_changed = false, head_changed = false;
#pragma unroll 1
            for (int c = half; c < TN / 32; c += kHalves) {
                uint32_t v[32];
                tmem_ld32(lane_base + static_cast<uint32_t>(as * TN + c * 32), v);
                tmem_ld_wait();
                bool any = false;
                if (prefilter) {
                    float m = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
                    for (int j = 2; j < 32; j += 2) m = fmaxf(m, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
                    const float relaxed = fabsf(thr) <= 3.0e38f ? thr - fabsf(thr) * 0.015625f : thr;
                    any = m > relaxed;
                    if (any) {  // apply the scales and test exactly
                        any = false;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float s = __uint_as_float(v[j]) * sc_tile[c * 32 + j];
                            v[j] = __float_as_uint(s);
                            any |= s > thr;
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float s = __uint_as_float(v[j]) * sc_tile[c * 32 + j];
                        v[j] = __float_as_uint(s);
                        any |= s > thr;
                    }
                }
                if (DUMP) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int64_t row = static_cast<int64_t>(t) * TN + c * 32 + j;
                        if (row < a.n) a.dump[static_cast<size_t>(qglobal) * a.dump_ld + row] = __uint_as_float(v[j]);
                    }
                }
                if (KREG > 0) {
                    if (any) {
                        // rare: build the candidate bitmap, then handle candidates one at a time through a
                        // single inlined insertion site (keeps the list in registers, keeps the code small)
                        uint32_t cand = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) cand |= (__uint_as_float(v[j]) > thr) ? (1u << j) : 0u;
                        while (cand) {
                            const int j = __ffs(cand) - 1;
                            cand &= cand - 1;
                            const float s = __uint_as_float(pick32(v, j));
                            if (s > thr) {  // thr may have risen since the bitmap was built
                                list_insert_regs<KR>(L, pack_key(s, a.row_offset + (lrow0 + c * 32 + j) * a.row_stride));
                                if (L[KR - 1] != 0ull) thr_local = key_score(L[KR - 1]);
                                thr = fmaxf(thr, thr_local);
                                if (s > best_local) {
                                    best_local = s;
                                    head_changed = true;
                                }
                                tile_changed = true;
                            }
                        }
                    }
                } else if (__any_sync(kFullMask, any)) {
                    // large k: the lists live in global memory (L1/L2 resident) and the whole warp inserts
                    // one candidate at a time (warp_list_insert: ballot-located position, parallel shift)
                    uint32_t cand = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) cand |= (__uint_as_float(v[j]) > thr) ? (1u << j) : 0u;
                    uint32_t owners_mask = __ballot_sync(kFullMask, cand != 0u);
                    while (owners_mask) {
                        const int src = __ffs(owners_mask) - 1;
                        owners_mask &= owners_mask - 1;
                        uint32_t cm = __shfl_sync(kFullMask, cand, src);
                        float thr_src = __shfl_sync(kFullMask, thr, src);
                        float local_src = __shfl_sync(kFullMask, thr_local, src);
                        uint64_t* lst = list + static_cast<ptrdiff_t>(src - lane) * a.k;  // src lane's list
                        bool changed = false;
                        while (cm) {
                            const int j = __ffs(cm) - 1;
                            cm &= cm - 1;
                            const float s = __shfl_sync(kFullMask, __uint_as_float(pick32(v, j)), src);
                            if (s > thr_src) {
                                const uint64_t kth =
                                    warp_list_insert(lst, a.k, pack_key(s, a.row_offset + (lrow0 + c * 32 + j) * a.row_stride), lane);
                                if (kth != 0ull) local_src = key_score(kth);
                                thr_src = fmaxf(thr_src, local_src);
                                changed = true;
                            }
                        }
                        if (lane == src && changed) {
                            thr_local = local_src;
                            thr = thr_src;
                            if (share && local_src > -INFINITY && ord_of(local_src) > max(g_seen, g_pub)) {
                                g_pub = ord_of(local_src);
                                red_max_u32(a.gthr + qglobal, g_pub);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(bar_tempty + 8 * as);
                else mbar_arrive_cluster(bar_tempty + 8 * as, 0);
            }
            if (KREG > 0 && share) {
                // this owner's new head for the query, and its k-th best if that can still raise the shared bound (once
                // the heads bound is folded in, a single owner's k-th best almost never can)
                if (head_changed) __stcg(head_row + half * ((a.P + 3) / 4 * 4) + p, ord_of(best_local));
                if (tile_changed && L[KR - 1] != 0ull) {
                    const uint32_t o = ord_of(thr_local);
                    if (o > max(g_seen, g_pub)) {
                        g_pub = o;
                        red_max_u32(a.gthr + qglobal, o);
                    }
                }
            }
            if (++as == kAccStages) { as = 0; aphase ^= 1; }
            if (tl) {
                const uint64_t te3 = global_timer_ns();
                e_top += te1 - te0;
                e_wait += te2 - te1;
                e_scan += te3 - te2;
                if (blockIdx.x == 0 && et == 0 && tcount <= 64) {
                    unsigned long long* const tr = a.tl + 16 * 296 + (tcount - 1) * 6;
                    tr[3] = te1; tr[4] = te2; tr[5] = te3;
                }
            }
        }
        // the bound warp stops publishing before any owner reports "done" (the last owner of a query resets gthr)
        if (et == 0) *reinterpret_cast<volatile uint32_t*>(tmem_slot + 1) = 1u;
        if (fused_tail && owners > 1)
            while (*reinterpret_cast<volatile uint32_t*>(tmem_slot + 2) == 0u) {}
        if (tl && et == 0) {
            tl[5] = global_timer_ns();                          // 5: last tile scanned
            tl[10] = e_top;                                     // 10: epilogue loop top (scales, shared bound, bar.sync)
            tl[11] = e_wait;                                    // 11: epilogue waited for an accumulator (it was ahead)
            tl[12] = e_scan;                                    // 12: TMEM loads + score scan + insertions
        }
        if (fused_tail) {
            // ---- fused tail: no merge launch.  A single list owner: its list is final.  Otherwise every owner
            // publishes the keys that can still be in the global top-k -- score >= the largest shared lower bound it
            // knows (gthr: every owner's k-th best and the bound warps' k-th largest head): both bound the global k-th
            // best from below, so every global top-k key survives in its owner.
            if (owners == 1) {
                if (live) {
#pragma unroll
                    for (int j = 0; j < KR; ++j)
                        if (j >= KR - a.k) a.out[static_cast<size_t>(qglobal) * a.k + (j - (KR - a.k))] = L[j];
                }
            } else if (live) {
                // Publish, then merge wherever the LAST owner of a query happens to run: every owner appends its
                // surviving keys to cand[q][...] (atomicAdd on cnt[q] reserves the range), fences, and bumps the
                // query's "owners published" counter; the thread that completes the count merges that query on
                // the spot.  The owners of a query finish within microseconds of each other all over the grid, so
                // the B merges run in parallel on many SMs instead of queueing in one CTA.
                const uint64_t t_tail = a.dbg ? global_timer_ns() : 0;
                const uint32_t g = max(g_seen, __ldcg(a.gthr + qglobal));
                int ns = 0;
#pragma unroll
                for (int j = 0; j < KR; ++j)
                    ns += (j >= KR - a.k && L[j] != 0ull && static_cast<uint32_t>(L[j] >> 32) >= g) ? 1 : 0;
                if (ns) {  // the list is sorted: the survivors are its first ns entries
                    const uint32_t pos = atomicAdd(a.cnt + qglobal, static_cast<uint32_t>(ns));
                    CV_ASSERT(ns <= a.k && pos + ns <= static_cast<uint32_t>(a.cand_ld));
                    uint64_t* dst = a.cand + static_cast<size_t>(qglobal) * a.cand_ld + pos;
#pragma unroll
                    for (int j = 0; j < KR; ++j) {
                        const int i = j - (KR - a.k);
                        if (i >= 0 && i < ns) __stcg(dst + i, L[j]);
                    }
                }
                __threadfence();
                const bool last_owner = atomicAdd(a.done + qglobal, 1u) == static_cast<unsigned>(owners - 1);
                if (a.dbg && lane == 0) atomicMax(a.dbg + 1, global_timer_ns() - t_tail);
                if (last_owner) {
                    __threadfence();
                    const uint64_t t_merge = a.dbg ? global_timer_ns() : 0;
                    const uint32_t m = __ldcg(a.cnt + qglobal);
                    CV_ASSERT(m <= static_cast<uint32_t>(a.cand_ld));
                    if (a.dbg) atomicMax(a.dbg + 3, static_cast<unsigned long long>(m));
                    const uint64_t* src = a.cand + static_cast<size_t>(qglobal) * a.cand_ld;
#pragma unroll
                    for (int j = 0; j < KR; ++j) L[j] = (j < KR - a.k) ? ~0ull : 0ull;
                    constexpr int U = KR <= 16 ? 16 : 8;  // keys in flight per L2 round trip
#pragma unroll 1
                    for (uint32_t i0 = 0; i0 < m; i0 += U) {
                        uint64_t kk[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) kk[u] = (i0 + u < m) ? __ldcg(src + i0 + u) : 0ull;
#pragma unroll
                        for (int u = 0; u < U; ++u)
                            if (kk[u] > L[KR - 1]) list_insert_regs<KR>(L, kk[u]);
                    }
#pragma unroll
                    for (int j = 0; j < KR; ++j)
                        if (j >= KR - a.k) a.out[static_cast<size_t>(qglobal) * a.k + (j - (KR - a.k))] = L[j];
                    // every owner of this query is done: leave its shared state zero for the next launch
                    a.cnt[qglobal] = 0u;
                    a.done[qglobal] = 0u;
                    a.gthr[qglobal] = 0u;
                    for (int v = 0; v < kHalves * ((a.P + 3) / 4); ++v) reinterpret_cast<uint4*>(head_row)[v] = make_uint4(0u, 0u, 0u, 0u);
                    if (a.dbg) atomicMax(a.dbg + 2, global_timer_ns() - t_merge);
                }
            }
        } else if (KREG > 0) {
#pragma unroll
            for (int j = 0; j < KR; ++j)
                if (j >= KR - a.k) glist[j - (KR - a.k)] = L[j];
        } else if (a.smem_lists) {
            __syncwarp();
            for (int j = 0; j < a.k; ++j) glist[j] = list[j];
        }
    } else {
        // ===================================================================== bound warp
        // Keeps the shared lower bound of the global k-th best fresh WITHOUT the epilogue's help.  Every list owner
        // publishes its best score per query (its "head"); the k-th largest head bounds the global k-th best from
        // below (k distinct owners each hold a row at least that good) and, unlike one owner's own k-th best, reflects
        // every row the grid has scanned.  Each query has ONE duty CTA (p == query % P): its bound warp re-reads the
        // query's heads row over and over (coalesced, one L2 round trip), ranks it with warp reductions and raises
        // gthr[q], which every epilogue thread reads once per tile anyway.  (When the epilogue threads ranked the heads
        // themselves -- every live thread its own query's 148 heads, at tiles 1, 4, 16, ... and before the last tile --
        // each refresh held its CTA's epilogue for 7-12 us: per-tile trace, profiles/README.md.)
        volatile uint32_t* const scan_done = tmem_slot + 1;
        volatile uint32_t* const bound_ack = tmem_slot + 2;
        constexpr int J = 5;                               // 32 * J >= P (at most 148 list owners per column half)
        const int Pr = (a.P + 3) / 4 * 4;
        const int q0 = (qt * CG + static_cast<int>(rank)) * kQ;
        if (KREG > 0 && !DUMP && a.P >= a.k && a.P <= 32 * J && p < kQ && q0 + p < a.B) {
            bool dirty = false;  // a RED has been issued since the last fence
            int passes = 0;      // pilot: the first bound of a query is taken over EVERY owner's first head (or after ~8 us)
            while (*scan_done == 0u) {
                if (dirty) {  // (REDs are rare once the heads have settled: no fence in the steady state)
                    __threadfence();
                    dirty = false;
                }
                for (int ql = p; ql < kQ && q0 + ql < a.B && *scan_done == 0u; ql += a.P) {
                    const int qg = q0 + ql;
#pragma unroll 1
                    for (int h = 0; h < kHalves; ++h) {
                        const uint32_t* row = a.heads + static_cast<size_t>(qg) * kHeadsLd + h * Pr;
                        uint32_t v[J];
#pragma unroll
                        for (int j = 0; j < J; ++j) v[j] = (lane + 32 * j < a.P) ? __ldcg(row + lane + 32 * j) : 0u;
                        const uint32_t cur = __ldcg(a.gthr + qg);
                        bool all_in = true;
#pragma unroll
                        for (int j = 0; j < J; ++j) all_in &= (lane + 32 * j >= a.P) || v[j] != 0u;
                        if (PILOT && passes < 8 && !__all_sync(kFullMask, all_in)) continue;
                        const uint32_t kth = warp_kth_largest<J>(v, a.k, lane);
                        if (kth > cur && *scan_done == 0u) {  // (warp-uniform)
                            if (lane == 0) red_max_u32(a.gthr + qg, kth);
                            dirty = true;
                        }
                    }
                }
                ++passes;
                __nanosleep(200);
            }
            if (dirty) __threadfence();  // every RED above is visible before the ack below can be
        }
        __syncwarp();
        if (lane == 0) *bound_ack = 1u;
    }
    if (tl && threadIdx.x == 64) tl[6] = global_timer_ns();    // 6: tail (publish / merge / list write-out) done
    tc_fence_before();
    if (CG == 2) cluster_sync(); else __syncthreads();  // nobody leaves while its partner may still touch it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, kCols);
    }
    if (tl && threadIdx.x == 32) tl[7] = global_timer_ns();    // 7: CTA exit
}

using MmaKernel = void (*)(const CUtensorMap, const CUtensorMap, const KArgs);

// k <= 16: four or eight epilogue warps (the host picks; four by default); everything else four (register budget of
// eight: 65536 / 320 threads = 204 per thread).
template <int CG>
inline MmaKernel pick_mma_kernel(bool dump, int k, int ew) {
    if (dump) return k_mma_topk<true, 0, CG, 4>;
    if (k <= 10) return ew == 4 ? static_cast<MmaKernel>(k_mma_topk<false, 10, CG, 4>) : static_cast<MmaKernel>(k_mma_topk<false, 10, CG, 8>);
    if (k <= 16) return ew == 4 ? static_cast<MmaKernel>(k_mma_topk<false, 16, CG, 4>) : static_cast<MmaKernel>(k_mma_topk<false, 16, CG, 8>);
    if (k <= 32) return k_mma_topk<false, 32, CG, 4>;
    if (k <= 64) return k_mma_topk<false, 64, CG, 4>;
    return k_mma_topk<false, 0, CG, 4>;
}
inline int mma_kernel_epilogue_warps(bool dump, int k, int ew) { return (!dump && k <= 16 && ew == 8) ? 8 : 4; }

MmaKernel pick_mma_kernel_cg2(bool dump, int k, int ew);  // k3_inst_cg2.cu
// PILOT instantiations (two-pass first tile; single CTAs, four epilogue warps, register lists): k3_inst_pilot.cu
MmaKernel pick_mma_kernel_pilot(int k);

}  // namespace k3
}  // namespace cv
