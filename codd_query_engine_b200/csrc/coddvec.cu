// coddvec.cu -- host runtime + C-ABI of libcoddvec.so (see include/coddvec.h).
//
// Everything the reference delegates to the third-party chromadb client for the
// `search_relevant_metrics` path (codd_dal/metrics/metrics_semantic_metadata_store.py:60-69,
// :236-238, :260, :314-316) that is data-parallel lives here: the device-resident corpus, K1
// normalise/append, K2 streaming scan, K3 tensor-core scoring, K4 merge.  No CPU fallback.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <map>
#include <unordered_map>
#include <vector>

#define CV_MAIN_TU 1
#include "../../include/coddvec.h"
#include "common.cuh"
#include "k1_normalise.cuh"
#include "k2_scan.cuh"
#include "k3_mma.cuh"
#include "vmm.hpp"

namespace {

using namespace cv;

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CV_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            (void)cudaGetLastError();                                                              \
            return fail(e__ == cudaErrorMemoryAllocation ? CV_ERR_OOM : CV_ERR_CUDA, "%s failed: %s (%s:%d)", \
                        #call, cudaGetErrorString(e__), __FILE__, __LINE__);                       \
        }                                                                                          \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { (void)cudaGetLastError(); prev = -1; }
        ok = cudaSetDevice(dev) == cudaSuccess;
        if (!ok) (void)cudaGetLastError();
    }
    ~DeviceGuard() {
        if (prev >= 0) (void)cudaSetDevice(prev);
    }
};

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    int reserve(size_t n) {
        if (n <= cap) return CV_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 2 + 64;
        CV_CUDA(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
        return CV_OK;
    }
    ~DevBuf() { if (p) cudaFree(p); }
};

template <typename T>
struct PinBuf {
    T* p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return CV_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = n + n / 2 + 64;
        CV_CUDA(cudaMallocHost(&p, want * sizeof(T)));
        cap = want;
        return CV_OK;
    }
    ~PinBuf() { if (p) cudaFreeHost(p); }
};

// AUTO path: batches up to this size take the fp32-exact K2 scan, larger ones the tensor-core K3.
// (Rows up to 384 elements: two queries still fit one K2 pass at HBM speed -- 0.140 ms against K3's 0.151 ms on
// 1M x 384 -- four do not, the FMA work doubles again and the pass drops to 3.4 TB/s.)
constexpr int kAutoScanMaxBatch = 1;
constexpr int kAutoScanMaxBatchShortRows = 2;  // rows up to 768 elements: two queries per HBM-speed K2 pass

// Per-query scratch.  A context is used by one query batch at a time; `busy` orders reuse on a
// different stream without blocking the host.
struct QueryCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t busy = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    bool timed = false;
    DevBuf<float> d_q;
    DevBuf<uint64_t> d_partial;
    DevBuf<uint64_t> d_out;
    DevBuf<uint64_t> d_local;  // this GPU's top-k before the fused cross-GPU exchange
    DevBuf<uint32_t> d_mask;
    cv::MmaScratch mma;  // K3 scratch (bf16 queries, per-CTA lists)
    unsigned int* d_counter = nullptr;
    uint64_t* d_heads = nullptr;  // K2 heads bound: [4 queries][kHeadSlots], zero between launches (self-resetting)
    PinBuf<float> h_q;
    PinBuf<uint64_t> h_out;
    PinBuf<uint64_t> h_flag;   // [1] completion word of the low-latency cv_query path (written by the kernel)
    uint64_t ticket = 0;       // last value the flag was asked to take
    struct Plan { cv::ScanFn fn; int grid; size_t smem; };
    std::unordered_map<int, Plan> plans;  // (qb, pred, k) -> launch plan
    ~QueryCtx() {
        if (d_counter) cudaFree(d_counter);
        if (d_heads) cudaFree(d_heads);
        if (mma.p) cudaFree(mma.p);
        if (mma.state) cudaFree(mma.state);
        if (busy) cudaEventDestroy(busy);
        if (t0) cudaEventDestroy(t0);
        if (t1) cudaEventDestroy(t1);
        if (stream) cudaStreamDestroy(stream);
    }
};

}  // namespace

struct cv_index {
    cv_config cfg{};
    int D = 0, Dp = 0, C = 0;
    int64_t row_stride = 1;
    int sm_count = 148;
    int64_t max_rows = 0;
    VmmArray E, scale, ns, label;
    std::atomic<int64_t> n_committed{0};
    int64_t capacity_rows = 0;

    std::mutex upsert_mu;          // serialises writers
    std::shared_mutex overwrite_mu;  // writers overwriting live rows exclude query enqueue
    cudaStream_t upsert_stream = nullptr;
    DevBuf<float> stage_x;
    DevBuf<uint32_t> stage_ns, stage_label;
    DevBuf<int64_t> stage_rows;
    unsigned int* d_flag = nullptr;

    std::mutex ctx_mu;
    std::vector<std::unique_ptr<QueryCtx>> ctx_all;  // owner
    std::vector<QueryCtx*> ctx_free;                  // idle contexts

    std::atomic<int64_t> queries{0}, launches{0}, upserted{0};
    std::mutex stats_mu;
    QueryCtx* last_timed = nullptr;
    int64_t last_bytes = 0;
    int32_t last_path = 0;
    std::atomic<int32_t> profiling{0};
    std::atomic<bool> scale_bounded{true};  // every stored scale is 0, NaN or within 2^-7 of 1 (K3's raw-dot prefilter)

    ~cv_index() {
        ctx_free.clear();
        ctx_all.clear();
        if (d_flag) cudaFree(d_flag);
        if (upsert_stream) cudaStreamDestroy(upsert_stream);
    }
};

namespace {

int ensure_capacity(cv_index* idx, int64_t rows) {
    if (rows <= idx->capacity_rows) return CV_OK;
    if (rows > idx->max_rows)
        return fail(CV_ERR_INVALID, "index is full: %lld rows requested, max_rows = %lld", (long long)rows,
                    (long long)idx->max_rows);
    auto grow = [&](VmmArray& a, size_t bytes) -> int {
        std::string e = a.ensure(bytes);
        if (e.empty()) return CV_OK;
        return fail(e.rfind("OOM", 0) == 0 ? CV_ERR_OOM : CV_ERR_CUDA, "%s", e.c_str());
    };
    int rc;
    if ((rc = grow(idx->E, static_cast<size_t>(rows) * idx->Dp * 2))) return rc;
    if ((rc = grow(idx->scale, static_cast<size_t>(rows) * 4))) return rc;
    if ((rc = grow(idx->ns, static_cast<size_t>(rows) * 4))) return rc;
    if ((rc = grow(idx->label, static_cast<size_t>(rows) * 4))) return rc;
    int64_t cap = static_cast<int64_t>(idx->E.mapped_bytes() / (static_cast<size_t>(idx->Dp) * 2));
    cap = std::min<int64_t>(cap, idx->scale.mapped_bytes() / 4);
    cap = std::min<int64_t>(cap, idx->ns.mapped_bytes() / 4);
    cap = std::min<int64_t>(cap, idx->label.mapped_bytes() / 4);
    idx->capacity_rows = std::min(cap, idx->max_rows);
    return CV_OK;
}

int acquire_ctx(cv_index* idx, QueryCtx** out) {
    {
        std::lock_guard<std::mutex> g(idx->ctx_mu);
        if (!idx->ctx_free.empty()) {
            *out = idx->ctx_free.back();
            idx->ctx_free.pop_back();
            return CV_OK;
        }
    }
    std::unique_ptr<QueryCtx> c(new QueryCtx());
    CV_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CV_CUDA(cudaEventCreateWithFlags(&c->busy, cudaEventDisableTiming));
    CV_CUDA(cudaEventCreate(&c->t0));
    CV_CUDA(cudaEventCreate(&c->t1));
    CV_CUDA(cudaMalloc(&c->d_counter, 64 * sizeof(unsigned int)));
    // stream-ordered: the context's streams are non-blocking, so a NULL-stream cudaMemset would not be
    // ordered before the first kernel that uses the counter; `busy` carries the dependency to any stream
    CV_CUDA(cudaMemsetAsync(c->d_counter, 0, 64 * sizeof(unsigned int), c->stream));
    CV_CUDA(cudaMalloc(&c->d_heads, 4 * cv::kHeadSlots * sizeof(uint64_t)));
    CV_CUDA(cudaMemsetAsync(c->d_heads, 0, 4 * cv::kHeadSlots * sizeof(uint64_t), c->stream));
    CV_CUDA(cudaEventRecord(c->busy, c->stream));
    {
        std::lock_guard<std::mutex> g(idx->ctx_mu);
        *out = c.get();
        idx->ctx_all.push_back(std::move(c));
    }
    return CV_OK;
}

void release_ctx(cv_index* idx, QueryCtx* c) {
    std::lock_guard<std::mutex> g(idx->ctx_mu);
    idx->ctx_free.push_back(c);
}

// ------------------------------------------------------------------------------------------ K2 dispatch
// (instantiations live in k2_inst_qb{1,2,4}.cu so that they compile in parallel)
ScanFn pick_scan(int nch, int qb, bool pred) {
    switch (qb) {
        case 1: return pick_scan_qb1(nch, pred);
        case 2: return pick_scan_qb2(nch, pred);
        case 4: return pick_scan_qb4(nch, pred);
        default: return nullptr;
    }
}

struct ScanLaunch {
    ScanFn fn;
    int grid;
    size_t smem;
};

int plan_scan(cv_index* idx, QueryCtx* c, int qb, int k, bool pred, ScanLaunch* out) {
    const int plan_key = qb | (pred ? 16 : 0) | (k << 8);
    auto it = c->plans.find(plan_key);
    if (it != c->plans.end()) {
        out->fn = it->second.fn;
        out->grid = it->second.grid;
        out->smem = it->second.smem;
        return CV_OK;
    }
    const int nch = (idx->C + 15) / 16;
    ScanFn fn = pick_scan(nch, qb, pred);
    if (!fn) return fail(CV_ERR_UNSUPPORTED, "scan path: no kernel for dim=%d with %d queries per pass", idx->D, qb);
    auto smem_for = [&](int kk) {
        size_t b = (static_cast<size_t>(kSelCap) + static_cast<size_t>(kScanWarps) * qb * scan_list_cap(kk)) * sizeof(uint64_t);
        if (scan_is_wide(nch)) b += static_cast<size_t>(idx->Dp) * sizeof(float);  // the normalised query
        else if (scan_queries_in_smem(nch, qb)) b += static_cast<size_t>(qb) * idx->C * 32;  // queries as 2 x float4 per chunk
        return b;
    };
    const size_t smem = smem_for(k);
    // The dynamic shared-memory limit is an attribute per function AND device, shared by the whole process, and one
    // function serves every k (and every index of this dimension class): only ever raise it, to what the largest k
    // needs -- on every device the function is launched on (a striped collection runs it on several).
    {
        static std::mutex attr_mu;
        static std::map<std::pair<int, const void*>, size_t> attr_set;
        std::lock_guard<std::mutex> g(attr_mu);
        int dev = 0;
        CV_CUDA(cudaGetDevice(&dev));
        size_t& have = attr_set[{dev, reinterpret_cast<const void*>(fn)}];
        const size_t want = smem_for(CV_MAX_K);
        if (have < want) {
            CV_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(want)));
            have = want;
        }
    }
    int occ = 0;
    CV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kScanThreads, smem));
    if (occ < 1) return fail(CV_ERR_CUDA, "scan kernel cannot be resident (smem=%zu)", smem);
    if (occ > 8) occ = 8;
    out->fn = fn;
    out->grid = idx->sm_count * occ;
    out->smem = smem;
    c->plans[plan_key] = QueryCtx::Plan{fn, out->grid, smem};
    return CV_OK;
}

// Enqueue one query batch on `stream`.  d_q: [B][D] fp32 device, d_out: [B][k] device keys.
// done (optional): a host-mapped word + value; when the batch takes the K2 path its last launch stores the
// value there after writing d_out (which the caller then points at host-mapped memory) and *done->armed is
// set; on any other path it is left untouched and the caller synchronises the stream as usual.
struct DoneFlag {
    uint64_t* flag;
    uint64_t value;
    bool armed;
    const float* h_q;  // the queries in host memory: a single short query is passed inside the kernel parameters
};

int enqueue_query(cv_index* idx, QueryCtx* c, cudaStream_t stream, const float* d_q, const cv_query_params* p,
                  const uint32_t* d_mask, uint64_t* d_out, float* dump = nullptr, int64_t dump_ld = 0,
                  const PeerExchange* px = nullptr, uint64_t seq = 0, DoneFlag* done_flag = nullptr) {
    const int B = p->n_queries, k = p->k;
    const int64_t n = idx->n_committed.load(std::memory_order_acquire);
    const bool pred = d_mask != nullptr || p->ns_code != CV_ANY || p->label_code != CV_ANY;
    const bool exchange = px != nullptr && px->world > 1;
    if (n == 0) {
        if (!exchange) {
            CV_CUDA(cudaMemsetAsync(d_out, 0, sizeof(uint64_t) * B * k, stream));
            return CV_OK;
        }
        // an empty shard still has to take part in the exchange (it publishes empty lists)
        int rc0 = c->d_local.reserve(static_cast<size_t>(B) * k);
        if (rc0) return rc0;
        CV_CUDA(cudaMemsetAsync(c->d_local.p, 0, sizeof(uint64_t) * B * k, stream));
        k_exchange_merge<<<std::min(B, kExchangeCtas), 256, 0, stream>>>(*px, seq, 0, c->d_local.p, k, d_out, B);
        CV_CUDA(cudaGetLastError());
        idx->launches.fetch_add(1);
        return CV_OK;
    }
    const int nch = (idx->C + 15) / 16;
    int path = p->path;
    const int max_qb = scan_max_qb(nch);
    if (path == CV_PATH_AUTO) {
        // B = 1: fp32-exact K2 scan.  Batches: K3 on the tensor cores, except where K3 has no fast top-k
        // (k > 64 keeps lists in shared / global memory) and the batch is small enough for a few K2 passes.
        const int auto_scan = nch <= 6 ? kAutoScanMaxBatchShortRows : kAutoScanMaxBatch;
        const bool scan = B <= std::min(max_qb, auto_scan) || !mma_supported(idx->Dp) ||
                          (k > 64 && max_qb > 0 && B <= 4 * max_qb);
        path = scan ? CV_PATH_SCAN : CV_PATH_MMA;
    }
    const bool timed = idx->profiling.load() != 0;
    if (timed) CV_CUDA(cudaEventRecord(c->t0, stream));
    int64_t passes_bytes = 0;

    int variant = 0;
    if (path == CV_PATH_MMA_1CTA) { variant = 1; path = CV_PATH_MMA; }
    if (path == CV_PATH_MMA_2CTA) { variant = 2; path = CV_PATH_MMA; }
    if (path == CV_PATH_MMA) {
        MmaArgs m;
        m.variant = variant;
        m.E = idx->E.ptr<void>();
        m.scale = idx->scale.ptr<float>();
        m.ns = idx->ns.ptr<uint32_t>();
        m.label = idx->label.ptr<uint32_t>();
        m.mask = d_mask;
        m.ns_code = p->ns_code;
        m.label_code = p->label_code;
        m.n = n;
        m.D = idx->D;
        m.Dp = idx->Dp;
        m.cosine = idx->cfg.metric == CV_METRIC_COSINE;
        m.row_offset = static_cast<uint32_t>(idx->cfg.row_offset);
        m.row_stride = static_cast<uint32_t>(idx->row_stride);
        m.precision = p->precision;
        m.scale_bounded = idx->scale_bounded.load();
        m.q = d_q;
        m.B = B;
        m.k = k;
        if (exchange) {
            int rcl = c->d_local.reserve(static_cast<size_t>(B) * k);
            if (rcl) return rcl;
        }
        m.out = exchange ? c->d_local.p : d_out;
        m.dump = dump;
        m.dump_ld = dump_ld;
        m.counter = c->d_counter;
        m.sm_count = idx->sm_count;
        m.stream = stream;
        int rc = mma_enqueue(m, c->mma, &idx->launches);
        if (rc != 0) return fail(rc, "%s", mma_error().c_str());
        if (exchange) {
            k_exchange_merge<<<std::min(B, kExchangeCtas), 256, 0, stream>>>(*px, seq, 0, c->d_local.p, k, d_out, B);
            CV_CUDA(cudaGetLastError());
            idx->launches.fetch_add(1);
        }
        passes_bytes = n * (2ll * idx->Dp + 4) * ((B + 255) / 256);
    } else {
        int done = 0;
        while (done < B) {
            int qb = 1;
            while (qb * 2 <= max_qb && qb * 2 <= B - done) qb *= 2;
            ScanLaunch L;
            int rc = plan_scan(idx, c, qb, k, pred, &L);
            if (rc) return rc;
            rc = c->d_partial.reserve(static_cast<size_t>(qb) * L.grid * k);
            if (rc) return rc;
            ScanArgs a;
            a.E = idx->E.ptr<uint4>();
            a.scale = idx->scale.ptr<float>();
            a.ns = idx->ns.ptr<uint32_t>();
            a.label = idx->label.ptr<uint32_t>();
            a.mask = d_mask;
            a.ns_code = p->ns_code;
            a.label_code = p->label_code;
            a.n = n;
            a.C = idx->C;
            a.D = idx->D;
            a.cosine = idx->cfg.metric == CV_METRIC_COSINE;
            a.q = d_q + static_cast<size_t>(done) * idx->D;
            a.k = k;
            a.row_offset = static_cast<uint32_t>(idx->cfg.row_offset);
            a.row_stride = static_cast<uint32_t>(idx->row_stride);
            a.partial = c->d_partial.p;
            a.counter = c->d_counter;
            a.heads = c->d_heads;
            a.out = d_out + static_cast<size_t>(done) * k;
            if (exchange) a.px = *px; else a.px.world = 0;
            a.seq = seq;
            a.q_base = done;
            const bool last_launch = done + qb >= B;
            a.done_flag = (done_flag && last_launch) ? done_flag->flag : nullptr;
            a.done_value = done_flag ? done_flag->value : 0;
            a.q_inline = 0;
            if (done_flag && done_flag->h_q && qb * idx->D <= kInlineQ && !scan_is_wide(nch)) {
                memcpy(a.qinline, done_flag->h_q + static_cast<size_t>(done) * idx->D, sizeof(float) * qb * idx->D);
                a.q_inline = 1;
            }
            if (a.done_flag) done_flag->armed = true;
            // small collections: no more CTAs than it takes to give every warp one iteration of rows
            const int64_t want_ctas = (n + 63) / 64;
            const int grid = static_cast<int>(std::min<int64_t>(L.grid, std::max<int64_t>(want_ctas, 1)));
            L.fn<<<grid, kScanThreads, L.smem, stream>>>(a);
            CV_CUDA(cudaGetLastError());
            idx->launches.fetch_add(1);
            passes_bytes += n * (2ll * idx->Dp + 4);
            done += qb;
        }
    }
    if (timed) {
        CV_CUDA(cudaEventRecord(c->t1, stream));
        c->timed = true;
        std::lock_guard<std::mutex> g(idx->stats_mu);
        idx->last_timed = c;
        idx->last_bytes = passes_bytes;
    }
    {
        std::lock_guard<std::mutex> g(idx->stats_mu);
        idx->last_path = path;
    }
    idx->queries.fetch_add(1);
    return CV_OK;
}

// Poll the host-mapped completion word for a short while (the kernel of an interactive query on a small
// collection finishes in microseconds), then fall back to a blocking stream synchronise so that a long scan
// does not burn a core.  A launch failure surfaces through cudaStreamQuery / cudaStreamSynchronize.
int wait_done_flag(const uint64_t* flag, uint64_t want, cudaStream_t s) {
    for (int spins = 0; spins < (1 << 16); ++spins) {
        if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) == want) return CV_OK;
        if ((spins & 0x3FFF) == 0x3FFF) {
            cudaError_t e = cudaStreamQuery(s);
            if (e != cudaSuccess && e != cudaErrorNotReady) {
                (void)cudaGetLastError();
                return fail(CV_ERR_CUDA, "query kernel failed: %s", cudaGetErrorString(e));
            }
        }
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    CV_CUDA(cudaStreamSynchronize(s));
    if (__atomic_load_n(flag, __ATOMIC_ACQUIRE) != want) return fail(CV_ERR_CUDA, "query completion flag missing after stream sync");
    return CV_OK;
}

int check_query_params(const cv_index* idx, const void* q, const cv_query_params* p) {
    if (!idx || !q || !p) return fail(CV_ERR_INVALID, "null argument");
    if (p->n_queries < 1) return fail(CV_ERR_INVALID, "n_queries must be >= 1 (got %d)", p->n_queries);
    if (p->k < 1 || p->k > CV_MAX_K) return fail(CV_ERR_INVALID, "k must be in [1, %d] (got %d)", CV_MAX_K, p->k);
    if (p->path < CV_PATH_AUTO || p->path > CV_PATH_MMA_2CTA)
        return fail(CV_ERR_INVALID, "unknown path %d", p->path);
    if (p->precision < CV_PRECISION_AUTO || p->precision > CV_PRECISION_SPLIT)
        return fail(CV_ERR_INVALID, "unknown precision %d", p->precision);
    return CV_OK;
}

}  // namespace

// =============================================================================================== C-ABI
extern "C" {

const char* cv_last_error(void) { return g_err.c_str(); }
const char* cv_version(void) { return "coddvec 0.1.0 (sm_100a)"; }

uint64_t cv_encode_key(float score, uint32_t global_row) { return cv::pack_key(score, global_row); }

void cv_decode_keys(const uint64_t* keys, int64_t n, float* out_scores, int64_t* out_rows) {
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t key = keys[i];
        if (key == 0) {
            if (out_scores) out_scores[i] = 0.0f;
            if (out_rows) out_rows[i] = -1;
        } else {
            if (out_scores) out_scores[i] = cv::key_score(key);
            if (out_rows) out_rows[i] = static_cast<int64_t>(cv::key_row(key));
        }
    }
}

int cv_create(const cv_config* cfg, cv_index** out) {
    if (!cfg || !out) return fail(CV_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->dim < 1 || cfg->dim > CV_MAX_DIM) return fail(CV_ERR_INVALID, "dim must be in [1, %d] (got %d)", CV_MAX_DIM, cfg->dim);
    if (cfg->metric != CV_METRIC_COSINE && cfg->metric != CV_METRIC_IP)
        return fail(CV_ERR_UNSUPPORTED, "unsupported metric %d (cosine and ip only)", cfg->metric);
    if (cfg->capacity_rows < 0 || cfg->max_rows < 0 || cfg->row_offset < 0 || cfg->row_stride < 0)
        return fail(CV_ERR_INVALID, "negative capacity / max_rows / row_offset / row_stride");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        return fail(CV_ERR_CUDA, "no CUDA device available (%s); coddvec has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(CV_ERR_INVALID, "device %d out of range [0, %d)", cfg->device, ndev);
    DeviceGuard dg(cfg->device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", cfg->device);
    CV_CUDA(cudaFree(nullptr));  // create the primary context
    cudaDeviceProp prop;
    CV_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(CV_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                    cfg->device, prop.major, prop.minor);

    std::unique_ptr<cv_index> idx(new cv_index());
    idx->cfg = *cfg;
    idx->D = cfg->dim;
    idx->Dp = (cfg->dim + 63) / 64 * 64;
    idx->C = idx->Dp / 8;
    idx->sm_count = prop.multiProcessorCount;
    idx->max_rows = cfg->max_rows > 0 ? cfg->max_rows : (int64_t(1) << 28);
    if (cfg->capacity_rows > idx->max_rows) idx->max_rows = cfg->capacity_rows;
    idx->row_stride = cfg->row_stride > 0 ? cfg->row_stride : 1;
    if (idx->row_stride > 1 && idx->max_rows > ((int64_t(1) << 32) - 2 - cfg->row_offset) / idx->row_stride)
        idx->max_rows = ((int64_t(1) << 32) - 2 - cfg->row_offset) / idx->row_stride;  // default reservation of a striped shard
    if (cfg->row_offset + (idx->max_rows - 1) * idx->row_stride > (int64_t(1) << 32) - 2)
        return fail(CV_ERR_INVALID, "row_offset + max_rows * row_stride exceeds the 32-bit global row space");

    std::string err;
    if (!(err = idx->E.reserve(cfg->device, static_cast<size_t>(idx->max_rows) * idx->Dp * 2)).empty() ||
        !(err = idx->scale.reserve(cfg->device, static_cast<size_t>(idx->max_rows) * 4)).empty() ||
        !(err = idx->ns.reserve(cfg->device, static_cast<size_t>(idx->max_rows) * 4)).empty() ||
        !(err = idx->label.reserve(cfg->device, static_cast<size_t>(idx->max_rows) * 4)).empty())
        return fail(CV_ERR_CUDA, "%s", err.c_str());
    CV_CUDA(cudaStreamCreateWithFlags(&idx->upsert_stream, cudaStreamNonBlocking));
    CV_CUDA(cudaMalloc(&idx->d_flag, 2 * sizeof(unsigned int)));  // [0] non-finite inputs, [1] a scale outside K3's bound
    if (cfg->capacity_rows > 0) {
        int rc = ensure_capacity(idx.get(), cfg->capacity_rows);
        if (rc) return rc;
    }
    *out = idx.release();
    return CV_OK;
}

void cv_destroy(cv_index* idx) {
    if (!idx) return;
    DeviceGuard dg(idx->cfg.device);
    (void)cudaDeviceSynchronize();
    delete idx;
}

int32_t cv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

int64_t cv_count(const cv_index* idx) { return idx ? idx->n_committed.load(std::memory_order_acquire) : 0; }
int32_t cv_dim(const cv_index* idx) { return idx ? idx->D : 0; }

int cv_set_profiling(cv_index* idx, int32_t enabled) {
    if (!idx) return fail(CV_ERR_INVALID, "null index");
    idx->profiling.store(enabled ? 1 : 0);
    return CV_OK;
}

int cv_synchronize(cv_index* idx) {
    if (!idx) return fail(CV_ERR_INVALID, "null index");
    DeviceGuard dg(idx->cfg.device);
    CV_CUDA(cudaDeviceSynchronize());
    return CV_OK;
}

static int upsert_impl(cv_index* idx, const int64_t* rows, const float* emb, const uint32_t* ns_codes,
                       const uint32_t* label_codes, int64_t n, int64_t* out_rows, bool on_device) {
    if (!idx) return fail(CV_ERR_INVALID, "null index");
    if (n < 0) return fail(CV_ERR_INVALID, "negative row count");
    if (n == 0) return CV_OK;
    if (!emb) return fail(CV_ERR_INVALID, "null embeddings");
    DeviceGuard dg(idx->cfg.device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", idx->cfg.device);
    std::lock_guard<std::mutex> wg(idx->upsert_mu);
    if (on_device) CV_CUDA(cudaDeviceSynchronize());  // whatever produced d_emb (any stream) must be done

    const int64_t committed = idx->n_committed.load(std::memory_order_acquire);
    // resolve destination rows: appends get committed, committed+1, ...
    std::vector<int64_t> dst(static_cast<size_t>(n));
    int64_t appended = 0;
    bool overwrite = false;
    for (int64_t i = 0; i < n; ++i) {
        int64_t r = rows ? rows[i] : -1;
        if (r < 0) {
            dst[i] = committed + appended++;
        } else {
            if (r >= committed) return fail(CV_ERR_INVALID, "row %lld is not committed (count = %lld)", (long long)r, (long long)committed);
            dst[i] = r;
            overwrite = true;
        }
    }
    if (overwrite && !on_device) {
        // overwrites destroy live rows: reject non-finite input before touching anything
        for (int64_t i = 0; i < n; ++i) {  // every row of the batch, appended ones too: a failure after K1 has rewritten
                                           // live rows could not be undone
            const float* r = emb + i * idx->D;
            for (int d = 0; d < idx->D; ++d) {
                uint32_t u;
                memcpy(&u, r + d, 4);
                if ((u & 0x7F800000u) == 0x7F800000u)
                    return fail(CV_ERR_NONFINITE, "embedding %lld contains NaN/Inf", (long long)i);
            }
        }
    }
    int rc = ensure_capacity(idx, committed + appended);
    if (rc) return rc;

    std::unique_lock<std::shared_mutex> ow(idx->overwrite_mu, std::defer_lock);
    cudaStream_t s = idx->upsert_stream;
    if (overwrite) {
        ow.lock();  // no query may be enqueued while live rows are being rewritten
        std::lock_guard<std::mutex> g(idx->ctx_mu);
        for (auto& c : idx->ctx_all) CV_CUDA(cudaStreamWaitEvent(s, c->busy, 0));
    }

    // host input is staged through a 64 MiB device buffer; device input is consumed in place
    const int64_t chunk_rows = on_device ? n : std::max<int64_t>(1, std::min<int64_t>(n, (int64_t(64) << 20) / (4ll * idx->D)));
    if (!on_device) {
        if ((rc = idx->stage_x.reserve(static_cast<size_t>(chunk_rows) * idx->D))) return rc;
        if (ns_codes && (rc = idx->stage_ns.reserve(chunk_rows))) return rc;
        if (label_codes && (rc = idx->stage_label.reserve(chunk_rows))) return rc;
    }
    if ((rc = idx->stage_rows.reserve(chunk_rows))) return rc;
    CV_CUDA(cudaMemsetAsync(idx->d_flag, 0, 2 * sizeof(unsigned int), s));
    if (overwrite && on_device) {
        // device input cannot be pre-checked on the host: count non-finite values over the WHOLE input before the
        // first K1 launch touches a live row
        const int64_t total = n * idx->D;
        int cgrid = static_cast<int>(std::min<int64_t>((total + 255) / 256, idx->sm_count * 8));
        k_count_nonfinite<<<cgrid, 256, 0, s>>>(emb, total, idx->d_flag);
        CV_CUDA(cudaGetLastError());
        unsigned int bad = 0;
        CV_CUDA(cudaMemcpyAsync(&bad, idx->d_flag, sizeof(bad), cudaMemcpyDeviceToHost, s));
        CV_CUDA(cudaStreamSynchronize(s));
        if (bad) return fail(CV_ERR_NONFINITE, "embeddings contain %u NaN/Inf values; nothing was written", bad);
    }

    for (int64_t lo = 0; lo < n; lo += chunk_rows) {
        const int64_t m = std::min(chunk_rows, n - lo);
        NormArgs a;
        if (on_device) {
            a.x = emb + lo * idx->D;
            a.ns_in = ns_codes ? ns_codes + lo : nullptr;
            a.label_in = label_codes ? label_codes + lo : nullptr;
        } else {
            CV_CUDA(cudaMemcpyAsync(idx->stage_x.p, emb + lo * idx->D, sizeof(float) * m * idx->D, cudaMemcpyHostToDevice, s));
            if (ns_codes) CV_CUDA(cudaMemcpyAsync(idx->stage_ns.p, ns_codes + lo, sizeof(uint32_t) * m, cudaMemcpyHostToDevice, s));
            if (label_codes) CV_CUDA(cudaMemcpyAsync(idx->stage_label.p, label_codes + lo, sizeof(uint32_t) * m, cudaMemcpyHostToDevice, s));
            a.x = idx->stage_x.p;
            a.ns_in = ns_codes ? idx->stage_ns.p : nullptr;
            a.label_in = label_codes ? idx->stage_label.p : nullptr;
        }
        CV_CUDA(cudaMemcpyAsync(idx->stage_rows.p, dst.data() + lo, sizeof(int64_t) * m, cudaMemcpyHostToDevice, s));
        a.dst_rows = idx->stage_rows.p;
        a.base_row = 0;
        a.n = m;
        a.D = idx->D;
        a.Dp = idx->Dp;
        a.cosine = idx->cfg.metric == CV_METRIC_COSINE;
        a.E = idx->E.ptr<__nv_bfloat16>();
        a.scale = idx->scale.ptr<float>();
        a.ns = idx->ns.ptr<uint32_t>();
        a.label = idx->label.ptr<uint32_t>();
        a.flags = idx->d_flag;
        const int64_t total = m * idx->D;
        int cgrid = static_cast<int>(std::min<int64_t>((total + 255) / 256, idx->sm_count * 8));
        k_count_nonfinite<<<cgrid, 256, 0, s>>>(a.x, total, idx->d_flag);
        CV_CUDA(cudaGetLastError());
        int grid = static_cast<int>(std::min<int64_t>((m + 7) / 8, idx->sm_count * 8));
        k_normalise_append<<<grid, 256, 0, s>>>(a);
        CV_CUDA(cudaGetLastError());
        idx->launches.fetch_add(2);
        if (!on_device || lo + chunk_rows < n) CV_CUDA(cudaStreamSynchronize(s));  // staging buffers are reused
    }
    unsigned int flags[2] = {0, 0};
    CV_CUDA(cudaMemcpyAsync(flags, idx->d_flag, sizeof(flags), cudaMemcpyDeviceToHost, s));
    CV_CUDA(cudaStreamSynchronize(s));
    const unsigned int bad = flags[0];
    if (flags[1]) idx->scale_bounded.store(false);  // (rows whose fp32 norm underflowed: K3 applies every scale from now on)
    if (bad) return fail(CV_ERR_NONFINITE, "embeddings contain %u NaN/Inf values; nothing was committed", bad);
    idx->n_committed.store(committed + appended, std::memory_order_release);
    idx->upserted.fetch_add(n);
    if (out_rows) memcpy(out_rows, dst.data(), sizeof(int64_t) * n);
    return CV_OK;
}

int cv_upsert(cv_index* idx, const int64_t* rows, const float* emb, const uint32_t* ns_codes,
              const uint32_t* label_codes, int64_t n, int64_t* out_rows) {
    return upsert_impl(idx, rows, emb, ns_codes, label_codes, n, out_rows, false);
}

int cv_upsert_device(cv_index* idx, const int64_t* rows, const float* d_emb, const uint32_t* d_ns_codes,
                     const uint32_t* d_label_codes, int64_t n, int64_t* out_rows) {
    return upsert_impl(idx, rows, d_emb, d_ns_codes, d_label_codes, n, out_rows, true);
}

int cv_query_device(cv_index* idx, const float* d_q, const cv_query_params* p, uint64_t* d_out_keys, void* stream) {
    int rc = check_query_params(idx, d_q, p);
    if (rc) return rc;
    if (!d_out_keys) return fail(CV_ERR_INVALID, "null output");
    DeviceGuard dg(idx->cfg.device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", idx->cfg.device);
    std::shared_lock<std::shared_mutex> rd(idx->overwrite_mu);
    QueryCtx* c = nullptr;
    if ((rc = acquire_ctx(idx, &c))) return rc;
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    cudaError_t e = cudaStreamWaitEvent(s, c->busy, 0);
    if (e == cudaSuccess) {
        rc = enqueue_query(idx, c, s, d_q, p, p->mask, d_out_keys);
        e = cudaEventRecord(c->busy, s);
    }
    release_ctx(idx, c);
    if (e != cudaSuccess) return fail(CV_ERR_CUDA, "stream ordering failed: %s", cudaGetErrorString(e));
    return rc;
}

int64_t cv_peer_exchange_bytes(int32_t world, int32_t max_queries, int32_t max_k) {
    if (world < 1 || world > CV_MAX_PEERS || max_queries < 1 || max_k < 1 || max_k > CV_MAX_K) return -1;
    const int64_t flags = static_cast<int64_t>(world) * max_queries;
    const int64_t keys = 2ll * world * max_queries * max_k;
    return (flags + keys) * static_cast<int64_t>(sizeof(uint64_t));
}

// validate a cv_peer_exchange against the batch and turn it into the kernels' PeerExchange
static int resolve_exchange(const cv_query_params* p, const cv_peer_exchange* px, uint64_t seq, PeerExchange* dx) {
    if (px->world < 2 || px->world > CV_MAX_PEERS || px->rank < 0 || px->rank >= px->world)
        return fail(CV_ERR_INVALID, "bad peer exchange: world=%d rank=%d", px->world, px->rank);
    if (p->n_queries > px->max_queries || p->k > px->max_k)
        return fail(CV_ERR_INVALID, "batch %d x k %d exceeds the exchange buffer capacity %d x %d", p->n_queries, p->k,
                    px->max_queries, px->max_k);
    if (seq == 0) return fail(CV_ERR_INVALID, "seq must start at 1");
    dx->world = px->world;
    dx->rank = px->rank;
    dx->max_q = px->max_queries;
    dx->max_k = px->max_k;
    for (int g = 0; g < kMaxPeers; ++g) dx->base[g] = g < px->world ? reinterpret_cast<uint64_t*>(px->buffers[g]) : nullptr;
    for (int g = 0; g < px->world; ++g)
        if (!dx->base[g]) return fail(CV_ERR_INVALID, "peer buffer %d is null", g);
    return CV_OK;
}

int cv_query_device_exchange(cv_index* idx, const float* d_q, const cv_query_params* p, const cv_peer_exchange* px,
                             uint64_t seq, uint64_t* d_out_keys, void* stream) {
    int rc = check_query_params(idx, d_q, p);
    if (rc) return rc;
    if (!d_out_keys || !px) return fail(CV_ERR_INVALID, "null argument");
    PeerExchange dx;
    if ((rc = resolve_exchange(p, px, seq, &dx))) return rc;
    DeviceGuard dg(idx->cfg.device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", idx->cfg.device);
    std::shared_lock<std::shared_mutex> rd(idx->overwrite_mu);
    QueryCtx* c = nullptr;
    if ((rc = acquire_ctx(idx, &c))) return rc;
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : c->stream;
    cudaError_t e = cudaStreamWaitEvent(s, c->busy, 0);
    if (e == cudaSuccess) {
        rc = enqueue_query(idx, c, s, d_q, p, p->mask, d_out_keys, nullptr, 0, &dx, seq);
        e = cudaEventRecord(c->busy, s);
    }
    release_ctx(idx, c);
    if (e != cudaSuccess) return fail(CV_ERR_CUDA, "stream ordering failed: %s", cudaGetErrorString(e));
    return rc;
}

// A host-buffer query in two phases, so that a striped collection can put one batch in flight on every device
// before waiting for any of them.
struct PendingQuery {
    cv_index* idx = nullptr;
    QueryCtx* c = nullptr;
    DoneFlag df{nullptr, 0, false, nullptr};
    bool direct = false;
    int B = 0, k = 0;
    ~PendingQuery() {
        if (idx && c) release_ctx(idx, c);
    }
};

static int check_finite_queries(const cv_index* idx, const float* q, int B) {
    const size_t qn = static_cast<size_t>(B) * idx->D;
    for (size_t i = 0; i < qn; ++i) {
        uint32_t u;
        memcpy(&u, q + i, 4);
        if ((u & 0x7F800000u) == 0x7F800000u) return fail(CV_ERR_NONFINITE, "query %zu contains NaN/Inf", i / idx->D);
    }
    return CV_OK;
}

// Phase 1: stage the queries / mask and enqueue the kernels on one of the index's query contexts.  Nothing blocks.
static int query_host_begin(cv_index* idx, const float* q, const cv_query_params* p, const uint32_t* mask,
                            const PeerExchange* px, uint64_t seq, PendingQuery* pend) {
    int rc = CV_OK;
    const int B = p->n_queries, k = p->k;
    const size_t qn = static_cast<size_t>(B) * idx->D;
    DeviceGuard dg(idx->cfg.device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", idx->cfg.device);
    std::shared_lock<std::shared_mutex> rd(idx->overwrite_mu);  // (an overwrite waits on every context's `busy` event)
    QueryCtx* c = nullptr;
    if ((rc = acquire_ctx(idx, &c))) return rc;
    pend->idx = idx;
    pend->c = c;
    pend->B = B;
    pend->k = k;
    cudaStream_t s = c->stream;
    CV_CUDA(cudaStreamWaitEvent(s, c->busy, 0));
    if ((rc = c->d_q.reserve(qn)) || (rc = c->h_q.reserve(qn)) || (rc = c->d_out.reserve(static_cast<size_t>(B) * k)) ||
        (rc = c->h_out.reserve(static_cast<size_t>(B) * k)) || (rc = c->h_flag.reserve(1)))
        return rc;
    memcpy(c->h_q.p, q, qn * sizeof(float));
    // Low-latency mode (interactive batches, B <= 4 -- the reference issues one query per call,
    // metrics_semantic_metadata_store.py:314-316): no copy-engine operations and no stream synchronise.  The
    // kernels read the queries from and write the keys to page-locked host memory directly (unified
    // addressing), the K2 tail raises a host-mapped flag, and this thread polls it.  Saves the two DMA
    // hand-offs and the interrupt-driven wake-up: ~50 us -> ~20 us end to end on a 10k-row collection.
    const bool direct = B <= 4 && idx->n_committed.load(std::memory_order_acquire) > 0;
    pend->direct = direct;
    const float* q_arg = c->h_q.p;
    uint64_t* out_arg = c->h_out.p;
    if (!direct) {
        CV_CUDA(cudaMemcpyAsync(c->d_q.p, c->h_q.p, qn * sizeof(float), cudaMemcpyHostToDevice, s));
        q_arg = c->d_q.p;
        out_arg = c->d_out.p;
    }
    if (c->ticket == 0) c->h_flag.p[0] = 0;  // fresh allocation
    pend->df = DoneFlag{c->h_flag.p, ++c->ticket, false, c->h_q.p};
    const uint32_t* d_mask = nullptr;
    if (mask) {
        const int64_t n = idx->n_committed.load(std::memory_order_acquire);
        const size_t words = static_cast<size_t>((n + 31) / 32);
        if (words) {
            if ((rc = c->d_mask.reserve(words))) return rc;
            CV_CUDA(cudaMemcpyAsync(c->d_mask.p, mask, words * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
            d_mask = c->d_mask.p;
        }
    }
    rc = enqueue_query(idx, c, s, q_arg, p, d_mask, out_arg, nullptr, 0, px, seq, direct ? &pend->df : nullptr);
    if (rc) { (void)cudaStreamSynchronize(s); return rc; }
    if (!direct) CV_CUDA(cudaMemcpyAsync(c->h_out.p, c->d_out.p, sizeof(uint64_t) * B * k, cudaMemcpyDeviceToHost, s));
    CV_CUDA(cudaEventRecord(c->busy, s));
    return CV_OK;
}

// Phase 2: wait for the batch and decode its keys (all three outputs may be NULL: wait only).
static int query_host_finish(PendingQuery* pend, float* out_scores, int64_t* out_rows, int32_t* out_counts) {
    cv_index* idx = pend->idx;
    QueryCtx* c = pend->c;
    const int B = pend->B, k = pend->k;
    DeviceGuard dg(idx->cfg.device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", idx->cfg.device);
    int rc;
    if (pend->df.armed) {
        if ((rc = wait_done_flag(c->h_flag.p, pend->df.value, c->stream))) return rc;
    } else {
        CV_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (out_scores || out_rows) cv_decode_keys(c->h_out.p, static_cast<int64_t>(B) * k, out_scores, out_rows);
    if (out_counts) {
        for (int b = 0; b < B; ++b) {
            int cnt = 0;
            while (cnt < k && c->h_out.p[static_cast<size_t>(b) * k + cnt] != 0) ++cnt;
            out_counts[b] = cnt;
        }
    }
    return CV_OK;
}

static int query_host(cv_index* idx, const float* q, const cv_query_params* p, const PeerExchange* px, uint64_t seq,
                      float* out_scores, int64_t* out_rows, int32_t* out_counts) {
    int rc = check_finite_queries(idx, q, p->n_queries);
    if (rc) return rc;
    PendingQuery pend;
    if ((rc = query_host_begin(idx, q, p, p->mask, px, seq, &pend))) return rc;
    return query_host_finish(&pend, out_scores, out_rows, out_counts);
}

int cv_query(cv_index* idx, const float* q, const cv_query_params* p, float* out_scores, int64_t* out_rows,
             int32_t* out_counts) {
    int rc = check_query_params(idx, q, p);
    if (rc) return rc;
    return query_host(idx, q, p, nullptr, 0, out_scores, out_rows, out_counts);
}

}  // extern "C"

// G shards of one collection striped over the devices of THIS process (global row r lives on shard r % G at local
// row r / G) and the peer-mapped exchange buffers their kernels trade per-GPU top-k through.
struct cv_group {
    std::vector<cv_index*> shards;
    std::vector<uint64_t*> buffers;  // buffers[g]: exchange buffer in device g's memory, peer-mapped into every other device
    int max_queries = 0, max_k = 0;
    std::mutex mu;                   // one batch in flight per group: all shards see the same batches in the same order
    uint64_t seq = 0;
    bool broken = false;             // a shard failed mid-batch: its peers' flags can no longer be trusted
    // Shapes every shard has already served: buffers (query staging, K2 partials, K3 scratch, masks) grow on first
    // use, and cudaMalloc / cudaFree wait for the device -- which must not happen while a peer's kernel is already
    // spinning on this shard's keys.  A new shape is first run on every shard WITHOUT the exchange ("priming").
    int primed_B = 0, primed_k = 0;
    unsigned primed_paths = 0;       // bit per (path, precision, filtered)
};

extern "C" {

int cv_query_exchange(cv_index* idx, const float* q, const cv_query_params* p, const cv_peer_exchange* px, uint64_t seq,
                      float* out_scores, int64_t* out_rows, int32_t* out_counts) {
    int rc = check_query_params(idx, q, p);
    if (rc) return rc;
    if (!px) return fail(CV_ERR_INVALID, "null argument");
    PeerExchange dx;
    if ((rc = resolve_exchange(p, px, seq, &dx))) return rc;
    return query_host(idx, q, p, &dx, seq, out_scores, out_rows, out_counts);
}

}  // extern "C"

namespace cv {
cudaError_t launch_merge_keys(const uint64_t* in, int G, int b_stride, int k_in, int k_out, uint64_t* out, int nq,
                              cudaStream_t stream, const uint32_t* bound32) {
    k_merge_keys<<<nq, 256, 0, stream>>>(in, G, b_stride, k_in, k_out, out, bound32);
    return cudaGetLastError();
}
}  // namespace cv

extern "C" {

int cv_debug_mma_scores(cv_index* idx, const float* q, int32_t B, int32_t path, int32_t precision, float* out_scores) {
    if (!idx || !q || !out_scores || B < 1) return fail(CV_ERR_INVALID, "bad argument");
    if (path != CV_PATH_MMA && path != CV_PATH_MMA_1CTA && path != CV_PATH_MMA_2CTA)
        return fail(CV_ERR_INVALID, "path must be one of the tensor-core paths");
    DeviceGuard dg(idx->cfg.device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", idx->cfg.device);
    const int64_t n = idx->n_committed.load(std::memory_order_acquire);
    if (n == 0) return CV_OK;
    std::shared_lock<std::shared_mutex> rd(idx->overwrite_mu);
    QueryCtx* c = nullptr;
    int rc;
    if ((rc = acquire_ctx(idx, &c))) return rc;
    struct Releaser {
        cv_index* i; QueryCtx* c;
        ~Releaser() { release_ctx(i, c); }
    } rel{idx, c};
    cudaStream_t s = c->stream;
    CV_CUDA(cudaStreamWaitEvent(s, c->busy, 0));
    const int Bpad = (B + 255) / 256 * 256;
    DevBuf<float> d_dump, d_q;
    DevBuf<uint64_t> d_out;
    if ((rc = d_dump.reserve(static_cast<size_t>(Bpad) * n)) || (rc = d_q.reserve(static_cast<size_t>(B) * idx->D)) ||
        (rc = d_out.reserve(static_cast<size_t>(B) * 1)))
        return rc;
    CV_CUDA(cudaMemcpyAsync(d_q.p, q, sizeof(float) * B * idx->D, cudaMemcpyHostToDevice, s));
    CV_CUDA(cudaMemsetAsync(d_dump.p, 0xFF, sizeof(float) * Bpad * n, s));
    cv_query_params p{};
    p.n_queries = B;
    p.k = 1;
    p.ns_code = CV_ANY;
    p.label_code = CV_ANY;
    p.mask = nullptr;
    p.path = path;
    p.precision = precision;
    rc = enqueue_query(idx, c, s, d_q.p, &p, nullptr, d_out.p, d_dump.p, n);
    if (rc) { (void)cudaStreamSynchronize(s); return rc; }
    CV_CUDA(cudaMemcpyAsync(out_scores, d_dump.p, sizeof(float) * B * n, cudaMemcpyDeviceToHost, s));
    CV_CUDA(cudaEventRecord(c->busy, s));
    CV_CUDA(cudaStreamSynchronize(s));
    return CV_OK;
}

int cv_merge_keys_device(int32_t device, const uint64_t* d_keys_in, int32_t G, int32_t B, int32_t k_in, int32_t k_out,
                         uint64_t* d_keys_out, void* stream) {
    if (!d_keys_in || !d_keys_out) return fail(CV_ERR_INVALID, "null argument");
    if (G < 1 || B < 1 || k_in < 1 || k_out < 1 || k_out > CV_MAX_K)
        return fail(CV_ERR_INVALID, "bad merge shape G=%d B=%d k_in=%d k_out=%d", G, B, k_in, k_out);
    DeviceGuard dg(device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    k_merge_keys<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_keys_in, G, B, k_in, k_out, d_keys_out, nullptr);
    CV_CUDA(cudaGetLastError());
    return CV_OK;
}

int cv_load_rows(cv_index* idx, const uint16_t* bf16, const float* scale, const uint32_t* ns_codes,
                 const uint32_t* label_codes, int64_t n, int64_t* out_rows) {
    if (!idx || n < 0 || (n > 0 && (!bf16 || !scale))) return fail(CV_ERR_INVALID, "bad argument");
    if (n == 0) return CV_OK;
    DeviceGuard dg(idx->cfg.device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", idx->cfg.device);
    std::lock_guard<std::mutex> wg(idx->upsert_mu);
    for (int64_t i = 0; i < n; ++i) {  // snapshots normally carry what cv_get_rows returned; anything else switches K3's prefilter off
        const float sc = scale[i];
        if (sc == sc && sc != 0.0f && !(sc - 1.0f <= 0.0078125f && 1.0f - sc <= 0.0078125f)) { idx->scale_bounded.store(false); break; }
    }
    const int64_t base = idx->n_committed.load(std::memory_order_acquire);
    int rc = ensure_capacity(idx, base + n);
    if (rc) return rc;
    cudaStream_t s = idx->upsert_stream;
    const size_t pitch = static_cast<size_t>(idx->Dp) * 2;
    if (idx->Dp != idx->D) CV_CUDA(cudaMemsetAsync(idx->E.ptr<unsigned char>() + base * pitch, 0, pitch * n, s));
    CV_CUDA(cudaMemcpy2DAsync(idx->E.ptr<unsigned char>() + base * pitch, pitch, bf16, static_cast<size_t>(idx->D) * 2,
                              static_cast<size_t>(idx->D) * 2, n, cudaMemcpyHostToDevice, s));
    CV_CUDA(cudaMemcpyAsync(idx->scale.ptr<float>() + base, scale, sizeof(float) * n, cudaMemcpyHostToDevice, s));
    if (ns_codes) CV_CUDA(cudaMemcpyAsync(idx->ns.ptr<uint32_t>() + base, ns_codes, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));
    else CV_CUDA(cudaMemsetAsync(idx->ns.ptr<uint32_t>() + base, 0, sizeof(uint32_t) * n, s));
    if (label_codes) CV_CUDA(cudaMemcpyAsync(idx->label.ptr<uint32_t>() + base, label_codes, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, s));
    else CV_CUDA(cudaMemsetAsync(idx->label.ptr<uint32_t>() + base, 0, sizeof(uint32_t) * n, s));
    CV_CUDA(cudaStreamSynchronize(s));
    idx->n_committed.store(base + n, std::memory_order_release);
    if (out_rows)
        for (int64_t i = 0; i < n; ++i) out_rows[i] = base + i;
    return CV_OK;
}

int cv_get_rows(cv_index* idx, const int64_t* rows, int64_t n, uint16_t* out_bf16, float* out_scale, uint32_t* out_ns,
                uint32_t* out_label) {
    if (!idx || (!rows && n > 0)) return fail(CV_ERR_INVALID, "null argument");
    DeviceGuard dg(idx->cfg.device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", idx->cfg.device);
    const int64_t committed = idx->n_committed.load(std::memory_order_acquire);
    for (int64_t i = 0; i < n; ++i)
        if (rows[i] < 0 || rows[i] >= committed) return fail(CV_ERR_INVALID, "row %lld out of range [0, %lld)", (long long)rows[i], (long long)committed);
    std::lock_guard<std::mutex> wg(idx->upsert_mu);
    int64_t i = 0;
    while (i < n) {  // copy maximal contiguous runs
        int64_t j = i + 1;
        while (j < n && rows[j] == rows[j - 1] + 1) ++j;
        const int64_t r0 = rows[i], m = j - i;
        if (out_bf16)
            CV_CUDA(cudaMemcpy2D(out_bf16 + i * idx->D, sizeof(uint16_t) * idx->D,
                                 idx->E.ptr<uint16_t>() + r0 * idx->Dp, sizeof(uint16_t) * idx->Dp,
                                 sizeof(uint16_t) * idx->D, m, cudaMemcpyDeviceToHost));
        if (out_scale) CV_CUDA(cudaMemcpy(out_scale + i, idx->scale.ptr<float>() + r0, sizeof(float) * m, cudaMemcpyDeviceToHost));
        if (out_ns) CV_CUDA(cudaMemcpy(out_ns + i, idx->ns.ptr<uint32_t>() + r0, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
        if (out_label) CV_CUDA(cudaMemcpy(out_label + i, idx->label.ptr<uint32_t>() + r0, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost));
        i = j;
    }
    return CV_OK;
}

// Row-level edits that must be ordered against running queries exactly like an in-place overwrite: take the writer
// locks, make the upsert stream wait for every query context, then write the few words.
static int edit_rows_begin(cv_index* idx, const int64_t* rows, int64_t n, std::unique_lock<std::shared_mutex>& ow) {
    const int64_t committed = idx->n_committed.load(std::memory_order_acquire);
    for (int64_t i = 0; i < n; ++i)
        if (rows[i] < 0 || rows[i] >= committed)
            return fail(CV_ERR_INVALID, "row %lld out of range [0, %lld)", (long long)rows[i], (long long)committed);
    ow.lock();
    std::lock_guard<std::mutex> g(idx->ctx_mu);
    for (auto& c : idx->ctx_all) CV_CUDA(cudaStreamWaitEvent(idx->upsert_stream, c->busy, 0));
    return CV_OK;
}

int cv_delete_rows(cv_index* idx, const int64_t* rows, int64_t n) {
    if (!idx || (!rows && n > 0) || n < 0) return fail(CV_ERR_INVALID, "bad argument");
    if (n == 0) return CV_OK;
    DeviceGuard dg(idx->cfg.device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", idx->cfg.device);
    std::lock_guard<std::mutex> wg(idx->upsert_mu);
    std::unique_lock<std::shared_mutex> ow(idx->overwrite_mu, std::defer_lock);
    int rc = edit_rows_begin(idx, rows, n, ow);
    if (rc) return rc;
    // a dead row keeps its slot (rows are dense) but carries scale = NaN: every kernel treats that as "not
    // eligible", so deletes cost no per-query mask; an upsert into the slot revives it
    static const uint32_t kNaN = 0x7FC00000u;
    cudaStream_t s = idx->upsert_stream;
    int64_t i = 0;
    std::vector<uint32_t> run;
    while (i < n) {
        int64_t j = i + 1;
        while (j < n && rows[j] == rows[j - 1] + 1) ++j;
        run.assign(static_cast<size_t>(j - i), kNaN);
        CV_CUDA(cudaMemcpyAsync(idx->scale.ptr<float>() + rows[i], run.data(), sizeof(uint32_t) * (j - i), cudaMemcpyHostToDevice, s));
        CV_CUDA(cudaStreamSynchronize(s));  // `run` is reused
        i = j;
    }
    return CV_OK;
}

int cv_set_codes(cv_index* idx, const int64_t* rows, const uint32_t* ns_codes, const uint32_t* label_codes, int64_t n) {
    if (!idx || (!rows && n > 0) || n < 0) return fail(CV_ERR_INVALID, "bad argument");
    if (n == 0 || (!ns_codes && !label_codes)) return CV_OK;
    DeviceGuard dg(idx->cfg.device);
    if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", idx->cfg.device);
    std::lock_guard<std::mutex> wg(idx->upsert_mu);
    std::unique_lock<std::shared_mutex> ow(idx->overwrite_mu, std::defer_lock);
    int rc = edit_rows_begin(idx, rows, n, ow);
    if (rc) return rc;
    cudaStream_t s = idx->upsert_stream;
    for (int64_t i = 0; i < n; ++i) {
        if (ns_codes) CV_CUDA(cudaMemcpyAsync(idx->ns.ptr<uint32_t>() + rows[i], ns_codes + i, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        if (label_codes) CV_CUDA(cudaMemcpyAsync(idx->label.ptr<uint32_t>() + rows[i], label_codes + i, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    }
    CV_CUDA(cudaStreamSynchronize(s));
    return CV_OK;
}

int cv_group_create(cv_index* const* shards, int32_t n_shards, int32_t max_queries, int32_t max_k, cv_group** out) {
    if (!shards || !out) return fail(CV_ERR_INVALID, "null argument");
    *out = nullptr;
    if (n_shards < 2 || n_shards > CV_MAX_PEERS) return fail(CV_ERR_INVALID, "a group has 2..%d shards (got %d)", CV_MAX_PEERS, n_shards);
    const int64_t bytes = cv_peer_exchange_bytes(n_shards, max_queries, max_k);
    if (bytes < 0) return fail(CV_ERR_INVALID, "bad exchange capacity %d x %d", max_queries, max_k);
    for (int g = 0; g < n_shards; ++g) {
        if (!shards[g]) return fail(CV_ERR_INVALID, "shard %d is null", g);
        if (shards[g]->D != shards[0]->D || shards[g]->cfg.metric != shards[0]->cfg.metric)
            return fail(CV_ERR_INVALID, "shards disagree on dimension / metric");
        if (shards[g]->row_stride != n_shards || shards[g]->cfg.row_offset != g)
            return fail(CV_ERR_INVALID, "shard %d must be created with row_stride = %d and row_offset = %d", g, n_shards, g);
    }
    std::unique_ptr<cv_group> grp(new cv_group());
    grp->shards.assign(shards, shards + n_shards);
    grp->max_queries = max_queries;
    grp->max_k = max_k;
    // every device must be able to store into every other device's buffer
    for (int a = 0; a < n_shards; ++a) {
        const int da = shards[a]->cfg.device;
        DeviceGuard dg(da);
        if (!dg.ok) return fail(CV_ERR_CUDA, "cudaSetDevice(%d) failed", da);
        for (int b = 0; b < n_shards; ++b) {
            const int db = shards[b]->cfg.device;
            if (da == db) continue;
            int can = 0;
            CV_CUDA(cudaDeviceCanAccessPeer(&can, da, db));
            if (!can) return fail(CV_ERR_UNSUPPORTED, "device %d cannot map device %d's memory (no P2P): a striped collection needs NVLink / PCIe peer access", da, db);
            cudaError_t e = cudaDeviceEnablePeerAccess(db, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                (void)cudaGetLastError();
                return fail(CV_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", da, db, cudaGetErrorString(e));
            }
            (void)cudaGetLastError();
        }
    }
    for (int g = 0; g < n_shards; ++g) {
        DeviceGuard dg(shards[g]->cfg.device);
        void* p = nullptr;
        CV_CUDA(cudaMalloc(&p, static_cast<size_t>(bytes)));
        grp->buffers.push_back(static_cast<uint64_t*>(p));
        CV_CUDA(cudaMemset(p, 0, static_cast<size_t>(bytes)));
        CV_CUDA(cudaDeviceSynchronize());
    }
    *out = grp.release();
    return CV_OK;
}

void cv_group_destroy(cv_group* grp) {
    if (!grp) return;
    for (size_t g = 0; g < grp->buffers.size(); ++g) {
        DeviceGuard dg(grp->shards[g]->cfg.device);
        (void)cudaDeviceSynchronize();
        if (grp->buffers[g]) cudaFree(grp->buffers[g]);
    }
    delete grp;
}

int cv_group_query(cv_group* grp, const float* q, const cv_query_params* p, const uint32_t* const* masks, float* out_scores,
                   int64_t* out_rows, int32_t* out_counts) {
    if (!grp) return fail(CV_ERR_INVALID, "null group");
    const int G = static_cast<int>(grp->shards.size());
    int rc = check_query_params(grp->shards[0], q, p);
    if (rc) return rc;
    if (p->n_queries > grp->max_queries || p->k > grp->max_k)
        return fail(CV_ERR_INVALID, "batch %d x k %d exceeds the group's exchange capacity %d x %d", p->n_queries, p->k,
                    grp->max_queries, grp->max_k);
    if ((rc = check_finite_queries(grp->shards[0], q, p->n_queries))) return rc;
    std::lock_guard<std::mutex> lock(grp->mu);
    if (grp->broken) return fail(CV_ERR_CUDA, "a shard of this striped collection failed earlier; the collection must be reopened");
    const unsigned path_bit = 1u << ((p->path * 3 + p->precision) * 2 + (masks != nullptr ? 1 : 0));
    if (p->n_queries > grp->primed_B || p->k > grp->primed_k || !(grp->primed_paths & path_bit)) {
        for (int g = 0; g < G; ++g) {
            PendingQuery warm;
            if ((rc = query_host_begin(grp->shards[g], q, p, masks ? masks[g] : nullptr, nullptr, 0, &warm))) return rc;
            if ((rc = query_host_finish(&warm, nullptr, nullptr, nullptr))) return rc;
        }
        grp->primed_B = std::max(grp->primed_B, p->n_queries);
        grp->primed_k = std::max(grp->primed_k, p->k);
        grp->primed_paths |= path_bit;
    }
    const uint64_t seq = ++grp->seq;
    std::vector<PendingQuery> pend(G);
    // shard 0 last: it is the one whose merged result is read, and its kernel waits for everybody else's keys anyway
    for (int i = 0; i < G; ++i) {
        const int g = (i + 1) % G;
        PeerExchange dx;
        dx.world = G;
        dx.rank = g;
        dx.max_q = grp->max_queries;
        dx.max_k = grp->max_k;
        for (int j = 0; j < kMaxPeers; ++j) dx.base[j] = j < G ? grp->buffers[j] : nullptr;
        rc = query_host_begin(grp->shards[g], q, p, masks ? masks[g] : nullptr, &dx, seq, &pend[g]);
        if (rc) {
            // the shards already launched are waiting for this one's keys: they give up after their bounded wait.
            grp->broken = true;
            return rc;
        }
    }
    rc = query_host_finish(&pend[0], out_scores, out_rows, out_counts);
    // every shard must be done before the next batch may reuse the exchange slots (they finish within microseconds
    // of shard 0: all of them wait for the same flags)
    for (int g = 1; g < G; ++g) {
        const int r2 = query_host_finish(&pend[g], nullptr, nullptr, nullptr);
        if (!rc) rc = r2;
    }
    if (rc) grp->broken = true;
    return rc;
}

int cv_stats(cv_index* idx, cv_stats_t* out) {
    if (!idx || !out) return fail(CV_ERR_INVALID, "null argument");
    memset(out, 0, sizeof(*out));
    out->rows = idx->n_committed.load(std::memory_order_acquire);
    out->capacity_rows = idx->capacity_rows;
    out->bytes_per_row = 2ll * idx->Dp + 12;
    out->queries = idx->queries.load();
    out->kernel_launches = idx->launches.load();
    out->upserted_rows = idx->upserted.load();
    out->dim_padded = idx->Dp;
    std::lock_guard<std::mutex> g(idx->stats_mu);
    out->last_path = idx->last_path;
    out->last_query_bytes = idx->last_bytes;
    if (idx->last_timed && idx->last_timed->timed) {
        DeviceGuard dg(idx->cfg.device);
        float ms = 0.0f;
        if (cudaEventSynchronize(idx->last_timed->t1) == cudaSuccess &&
            cudaEventElapsedTime(&ms, idx->last_timed->t0, idx->last_timed->t1) == cudaSuccess)
            out->last_query_ms = ms;
        else
            (void)cudaGetLastError();
    }
    return CV_OK;
}

}  // extern "C"
