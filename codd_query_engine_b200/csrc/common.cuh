// common.cuh -- shared device utilities for the coddvec kernels (sm_100a).
//
// Result keys: u64 = ordered(score) << 32 | (0xFFFFFFFF - global_row).  Unsigned order of keys
// == (score descending, row ascending) -- the tie rule of SURVEY.md section 8(c) -- so every
// top-k / merge in this library is "k largest u64".  Key 0 is the empty slot.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// CV_CHECKED builds (python -m codd_query_engine_b200.build --checked -> libcoddvec_checked.so) turn the index
// arithmetic the kernels rely on into device-side assertions that trap: the bounds-check pass this project runs in
// place of compute-sanitizer, which is closed on the GPU pool (profiles/README.md).
#ifdef CV_CHECKED
#include <cstdio>
#define CV_ASSERT(cond)                                                                                          \
    do {                                                                                                         \
        if (!(cond)) {                                                                                           \
            printf("CV_ASSERT failed: %s (%s:%d) block %d thread %d\n", #cond, __FILE__, __LINE__, (int)blockIdx.x, \
                   (int)threadIdx.x);                                                                            \
            __trap();                                                                                            \
        }                                                                                                        \
    } while (0)
#else
#define CV_ASSERT(cond) ((void)0)
#endif

namespace cv {

constexpr int kMaxK = 128;
constexpr uint32_t kAny = 0xFFFFFFFFu;
constexpr uint32_t kFullMask = 0xFFFFFFFFu;

__host__ __device__ __forceinline__ uint32_t f32_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
__host__ __device__ __forceinline__ float bits_f32(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

__host__ __device__ __forceinline__ uint64_t pack_key(float s, uint32_t row) {
    if (!(s == s)) s = bits_f32(0xFF800000u);  // NaN -> -inf
    s = s + 0.0f;                              // -0 -> +0
    uint32_t u = f32_bits(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (static_cast<uint64_t>(u) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - row);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) {
    uint32_t u = static_cast<uint32_t>(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return bits_f32(u);
}
__host__ __device__ __forceinline__ uint32_t key_row(uint64_t key) {
    return 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull);
}

#ifdef __CUDACC__
__device__ __forceinline__ float bf16lo_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi_f32(uint32_t packed) { return __uint_as_float(packed & 0xFFFF0000u); }

// streaming 16-byte load: read-only path, do not allocate in L1 (corpus rows are read once).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
    uint32_t lo = __shfl_sync(kFullMask, static_cast<uint32_t>(v), src);
    uint32_t hi = __shfl_sync(kFullMask, static_cast<uint32_t>(v >> 32), src);
    return (static_cast<uint64_t>(hi) << 32) | lo;
}

// ---------------------------------------------------------------------------------------------
// Warp-parallel insertion into a descending k-list held in shared memory (empty slots = 0).
// All 32 lanes call with the same (list, k, key).  Returns the new k-th entry (the threshold).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t warp_list_insert(uint64_t* list, int k, uint64_t key, int lane) {
    uint64_t kth = list[k - 1];
    if (key <= kth) return kth;
    uint64_t v[kMaxK / 32];
    int pos = 0;
#pragma unroll
    for (int j = 0; j < kMaxK / 32; ++j) {
        int i = j * 32 + lane;
        v[j] = (i < k) ? list[i] : 0ull;
        pos += __popc(__ballot_sync(kFullMask, v[j] > key));
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kMaxK / 32; ++j) {
        int i = j * 32 + lane;
        if (i >= pos && i + 1 < k) list[i + 1] = v[j];
    }
    if (lane == 0) list[pos] = key;
    __syncwarp();
    return list[k - 1];
}

// ---------------------------------------------------------------------------------------------
// K2's per-(warp, query) candidate store: an UNSORTED buffer of cap = scan_list_cap(k) keys.  A candidate
// that beats the register threshold is appended by one lane (a single shared-memory store); only when the
// buffer is full does the warp prune it to its k best -- every lane ranks its cap/32 keys against all cap
// keys (keys are unique: they carry the row id), keys of rank < k are scattered to buf[rank], so the survivors
// come out sorted and buf[k-1] is the new threshold.  Compared with keeping the list sorted on every insert
// (warp_list_insert) the fill phase costs one store per row instead of a ballot/shift pass, which is what
// dominates when a warp sees only a few hundred rows (predicate scans, small shards) and k is large.
// The threshold is stale between prunes, so somewhat more candidates are appended; each costs ~one store.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int scan_list_cap(int k) { return k <= 32 ? 64 : (k <= 64 ? 128 : 256); }

template <int CAP>
__device__ __forceinline__ uint64_t warp_prune_impl(uint64_t* buf, int k, int lane) {
    constexpr int PER = CAP / 32;
    uint64_t mine[PER];
    int rank[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        mine[j] = buf[j * 32 + lane];
        rank[j] = 0;
    }
    const ulonglong2* pairs = reinterpret_cast<const ulonglong2*>(buf);  // warp-uniform address: broadcast reads
#pragma unroll 4
    for (int i = 0; i < CAP / 2; ++i) {
        const ulonglong2 o = pairs[i];
#pragma unroll
        for (int j = 0; j < PER; ++j) rank[j] += (o.x > mine[j] ? 1 : 0) + (o.y > mine[j] ? 1 : 0);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        CV_ASSERT(rank[j] >= 0 && rank[j] < CAP);
        if (rank[j] < k) buf[rank[j]] = mine[j];
    }
    __syncwarp();
    return buf[k - 1];
}

// All 32 lanes call; buf holds exactly `cap` distinct non-zero keys.  Returns the k-th best.
__device__ __forceinline__ uint64_t warp_prune(uint64_t* buf, int cap, int k, int lane) {
    if (cap == 64) return warp_prune_impl<64>(buf, k, lane);
    if (cap == 128) return warp_prune_impl<128>(buf, k, lane);
    return warp_prune_impl<256>(buf, k, lane);
}

// Append `key` (warp-uniform) to the buffer; prune when it fills.  cnt is warp-uniform.  Returns true when a
// prune happened, in which case `kth` holds the new k-th best key of this buffer.
__device__ __forceinline__ bool warp_buf_append(uint64_t* buf, int cap, int k, uint64_t key, int lane, int& cnt,
                                                uint64_t& kth) {
    CV_ASSERT(cnt >= 0 && cnt < cap);
    if (lane == 0) buf[cnt] = key;
    if (++cnt < cap) return false;
    __syncwarp();
    kth = warp_prune(buf, cap, k, lane);
    cnt = k;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Block-wide bitonic sort (descending) of n = power of two u64 keys in shared memory.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_bitonic_desc(uint64_t* buf, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = buf[lo], b = buf[hi];
                if ((a < b) == desc) { buf[lo] = b; buf[hi] = a; }
            }
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Block-wide streaming top-k selection over m candidate keys (any order, 0 = empty) read through
// `src(i)`.  Exact.  The whole block must call it.  buf: shared scratch of kSelCap keys, cnt:
// shared int.  On return buf[0..k) holds the k best keys descending (0 padded).
//   tau: lower bound on the k-th best key (0 = none); candidates below it are dropped early.
// Strategy: append survivors to buf; whenever fewer than blockDim.x free slots remain, sort buf,
// keep the best k, and raise tau to the k-th.
// ---------------------------------------------------------------------------------------------
constexpr int kSelCap = 4096;

// U = candidates fetched per thread per round: the final cross-CTA merge reads its candidates from
// L2, so all U loads of a round are issued before any is consumed (one L2 round trip per round
// instead of one per candidate).  Requires blockDim.x * U <= kSelCap / 2.
template <int U = 1, typename Src>
__device__ __forceinline__ void block_select_topk(Src src, int m, int k, uint64_t tau, uint64_t* buf,
                                                  int* cnt) {
    __shared__ uint64_t s_tau;
    __syncthreads();  // buf / cnt may still be in use by the caller's previous phase
    if (threadIdx.x == 0) { *cnt = 0; s_tau = tau; }
    __syncthreads();
    const int round = static_cast<int>(blockDim.x) * U;
    int c = 0;  // block-uniform running count (tracked with __syncthreads_count, no re-read races)
    for (int base = 0; base < m; base += round) {
        if (c > kSelCap - round) {
            for (int i = c + threadIdx.x; i < kSelCap; i += blockDim.x) buf[i] = 0ull;
            block_bitonic_desc(buf, kSelCap);
            if (threadIdx.x == 0) {
                int keep = 0;
                while (keep < k && buf[keep] != 0ull) ++keep;
                *cnt = keep;
                if (keep == k && buf[k - 1] > s_tau) s_tau = buf[k - 1];
            }
            __syncthreads();
            c = *cnt;
            __syncthreads();
        }
        uint64_t key[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * static_cast<int>(blockDim.x) + threadIdx.x;
            key[u] = (i < m) ? src(i) : 0ull;
        }
        int mine = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (key[u] != 0ull && key[u] >= s_tau) {
                int p = atomicAdd(cnt, 1);
                CV_ASSERT(p >= 0 && p < kSelCap);
                buf[p] = key[u];
                ++mine;
            }
        }
        if (U == 1) {
            c += __syncthreads_count(mine);
        } else {
            __syncthreads();
            c = *cnt;          // everyone reads the same value ...
            __syncthreads();   // ... before anyone appends again
        }
    }
    int n = 32;
    while (n < c || n < k) n <<= 1;  // c <= kSelCap, k <= kMaxK
    __syncthreads();
    for (int i = c + threadIdx.x; i < n; i += blockDim.x) buf[i] = 0ull;
    block_bitonic_desc(buf, n);
}
// ---------------------------------------------------------------------------------------------
// Fused multi-GPU exchange (row-sharded search): instead of "scan kernel -> NCCL all-gather -> merge
// kernel", the block that holds a query's final per-GPU top-k stores it straight into EVERY peer's
// exchange buffer over NVLink (peer-mapped pointers), raises a per-(source rank, query) sequence flag
// with a system-scope release, waits for the other ranks' flags, and merges the G lists that landed in
// its own buffer.  One kernel per GPU per query batch; no collective launch on the critical path.
//
// Exchange buffer of one rank (device memory, zero-initialised, mapped into every peer):
//     flags [G][max_q]              u64   flags[src][q]  = seq of the last batch src published for query q
//     keys  [2][G][max_q][max_k]    u64   keys[seq & 1][src][q][0..k)
// Two key slots suffice: a rank can only start batch s+1 after it has seen every peer's batch-s flag,
// and a peer publishes batch s only after it finished merging batch s-1 (stream order).
// ---------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 8;

struct PeerExchange {
    int world;       // 0 / 1 = no exchange
    int rank;
    int max_q;
    int max_k;
    uint64_t* base[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

constexpr uint64_t kExchangeTimeoutNs = 30ull * 1000 * 1000 * 1000;
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Whole block.  Publish: local[0..k) (this GPU's final top-k for global query index q; shared or global memory) goes
// to every rank's buffer, then this rank's flag for q is raised everywhere.
__device__ __forceinline__ void exchange_publish(const PeerExchange& px, uint64_t seq, int q, const uint64_t* local, int k) {
    CV_ASSERT(q >= 0 && q < px.max_q && k <= px.max_k && px.rank < px.world && px.world <= kMaxPeers);
    const int G = px.world;
    const size_t flag_words = static_cast<size_t>(G) * px.max_q;
    const size_t slot_words = static_cast<size_t>(G) * px.max_q * px.max_k;
    const size_t key_off = flag_words + (seq & 1ull) * slot_words + (static_cast<size_t>(px.rank) * px.max_q + q) * px.max_k;
    // peer stores of my k keys (coalesced 8-byte stores over NVLink; my own copy goes the same way)
    for (int i = threadIdx.x; i < G * k; i += blockDim.x) {
        const int g = i / k, j = i - g * k;
        st_relaxed_sys(px.base[g] + key_off + j, local[j]);
    }
    __threadfence_system();
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < G)
        st_release_sys(px.base[threadIdx.x] + static_cast<size_t>(px.rank) * px.max_q + q, seq);
}

// Whole block.  Wait for every rank's flag for q in my own buffer, then merge the G lists that landed there: on
// return buf[0..k) holds the merged top-k of all ranks.
__device__ __forceinline__ void exchange_wait_merge(const PeerExchange& px, uint64_t seq, int q, int k, uint64_t* buf, int* cnt) {
    const int G = px.world;
    const size_t flag_words = static_cast<size_t>(G) * px.max_q;
    const size_t slot_words = static_cast<size_t>(G) * px.max_q * px.max_k;
    if (static_cast<int>(threadIdx.x) < G) {
        const uint64_t* f = px.base[px.rank] + static_cast<size_t>(threadIdx.x) * px.max_q + q;
        // bounded: a peer that never publishes (its launch failed, its process died) must not hang this GPU --
        // after kExchangeTimeoutNs the kernel traps and the failure surfaces on the host as a CUDA error
        uint64_t t0 = 0;
        unsigned int spins = 0;
        while (ld_acquire_sys(f) < seq) {
            if ((++spins & 0xFFFu) == 0u) {
                const uint64_t now = global_timer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > kExchangeTimeoutNs) __trap();
            }
        }
    }
    __syncthreads();
    const uint64_t* mine = px.base[px.rank] + flag_words + (seq & 1ull) * slot_words;
    auto src = [&](int i) {
        const int g = i / k, j = i - g * k;
        return ld_relaxed_sys(mine + (static_cast<size_t>(g) * px.max_q + q) * px.max_k + j);
    };
    block_select_topk<1>(src, G * k, k, 0ull, buf, cnt);
}

// Publish + wait + merge for one query (K2's last CTA).  local may alias buf.
__device__ __forceinline__ void exchange_and_merge(const PeerExchange& px, uint64_t seq, int q, const uint64_t* local,
                                                   int k, uint64_t* buf, int* cnt) {
    exchange_publish(px, seq, q, local, k);
    exchange_wait_merge(px, seq, q, k, buf, cnt);
}
#endif  // __CUDACC__

}  // namespace cv
