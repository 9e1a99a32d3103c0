/* coddvec.h -- C-ABI of libcoddvec.so, the B200-native (sm_100a) semantic top-k store.
 *
 * This is the drop-in boundary for the ONE data-parallel hot path of
 * sathish316/codd_query_engine: the cosine / inner-product top-k retrieval behind
 * `search_relevant_metrics`.  In the reference that arithmetic runs inside the
 * third-party `chromadb` client object that
 *   codd_dal/metrics/metrics_semantic_metadata_store.py:43-75   (constructor, get_or_create_collection)
 * is handed; the entry points below are what a binding for that object's
 *   Collection.upsert  (call site metrics_semantic_metadata_store.py:236-238)
 *   Collection.get     (call site metrics_semantic_metadata_store.py:260)
 *   Collection.query   (call site metrics_semantic_metadata_store.py:314-316)
 * needs from native code.  id-string <-> row maps, documents and metadata stay in
 * the host language (they are touched for k rows per query).
 *
 * Conventions
 *   - plain C, no exceptions / callbacks cross the ABI; every int-returning entry
 *     point returns CV_OK (0) or a negative CV_ERR_*; the message for the calling
 *     thread is available from cv_last_error().
 *   - "host" pointers are ordinary process memory owned by the caller; "device"
 *     pointers are CUDA device memory on the index's device, owned by the caller
 *     (e.g. a torch tensor's data_ptr()).  The library owns the index handle and
 *     all corpus memory.
 *   - rows are dense, 0-based *local* row numbers; results carry *global* row
 *     numbers = cv_config.row_offset + local row (row-sharded multi-GPU use).
 *   - result keys: u64 = ordered(score fp32) << 32 | (0xFFFFFFFF - global_row);
 *     unsigned order == (score descending, row ascending); 0 == empty slot.
 *   - thread safety: all entry points may be called concurrently on one index;
 *     appends never disturb running queries, in-place overwrites are ordered
 *     against them.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     fails with CV_ERR_CUDA.
 */
#ifndef CODDVEC_H_
#define CODDVEC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CV_OK 0
#define CV_ERR_INVALID (-1)     /* bad argument */
#define CV_ERR_CUDA (-2)        /* CUDA runtime / driver failure (incl. no device) */
#define CV_ERR_OOM (-3)         /* device or host memory exhausted */
#define CV_ERR_NONFINITE (-4)   /* NaN / Inf in an embedding or query */
#define CV_ERR_UNSUPPORTED (-5) /* valid request this build cannot serve */

#define CV_METRIC_COSINE 0 /* `hnsw:space = cosine` (metrics_semantic_metadata_store.py:64) */
#define CV_METRIC_IP 1     /* `hnsw:space = ip` */

#define CV_ANY 0xFFFFFFFFu /* namespace / label code meaning "no predicate" */
#define CV_MAX_K 128       /* >= MAX_N_RESULTS = 100 (metrics_semantic_metadata_store.py:24) */
#define CV_MAX_DIM 4096

#define CV_PATH_AUTO 0 /* pick by batch size */
#define CV_PATH_SCAN 1 /* K2: CUDA-core HBM streaming scan (small batches) */
#define CV_PATH_MMA 2  /* K3: TMA + tcgen05 tensor-core scoring (batched); CTA pairs when B > 128 */
#define CV_PATH_MMA_1CTA 3 /* K3 forced to single-CTA tiles (cta_group::1, 128 queries x 128 rows) */
#define CV_PATH_MMA_2CTA 4 /* K3 forced to CTA pairs (cta_group::2, 256 queries x 256 rows) */

/* Query precision of the tensor-core path (K2 always scores with the fp32 query).  AUTO: bf16 queries, except on
 * rows of up to 128 padded elements, where the hi/lo split (two bf16 tiles, two MMAs per K step, ~2^-17 relative)
 * keeps every score within the 1e-3 the north star allows. */
#define CV_PRECISION_AUTO 0
#define CV_PRECISION_BF16 1  /* one bf16 query tile at every dimension */
#define CV_PRECISION_SPLIT 2 /* hi/lo split at every dimension whose double tile fits shared memory (dim <= 384) */

typedef struct cv_index cv_index;

typedef struct cv_config {
    int32_t dim;           /* embedding dimension D (1..CV_MAX_DIM) */
    int32_t metric;        /* CV_METRIC_* */
    int32_t device;        /* CUDA device ordinal */
    int32_t flags;         /* reserved, 0 */
    int64_t capacity_rows; /* rows to back with physical memory up front (0 = grow on demand) */
    int64_t max_rows;      /* virtual reservation: hard upper bound on rows (0 = 2^28) */
    int64_t row_offset;    /* global row number of local row 0 */
    int64_t row_stride;    /* global row = row_offset + local row * row_stride; 0 or 1 = a dense block (row-sharded
                              SPMD use), G = this index holds every G-th row of a collection striped over G devices */
} cv_config;

typedef struct cv_query_params {
    int32_t n_queries;    /* B >= 1 */
    int32_t k;            /* 1..CV_MAX_K */
    uint32_t ns_code;     /* CV_ANY or the namespace code rows must carry */
    uint32_t label_code;  /* CV_ANY or the label code rows must carry */
    const uint32_t* mask; /* NULL or ceil(count/32) words, bit (r & 31) of word r>>5 set => row r eligible;
                             shared by the whole batch; host memory for cv_query, device memory for cv_query_device */
    int32_t path;         /* CV_PATH_* */
    int32_t precision;    /* CV_PRECISION_* (0 = AUTO) */
} cv_query_params;

typedef struct cv_stats_t {
    int64_t rows;             /* committed rows */
    int64_t capacity_rows;    /* rows currently backed by physical memory */
    int64_t bytes_per_row;    /* device bytes per row incl. scale + codes */
    int64_t queries;          /* query batches served */
    int64_t kernel_launches;  /* kernels launched by this index since creation */
    int64_t upserted_rows;    /* rows written by K1 since creation */
    double last_query_ms;     /* device time of the last profiled query (cv_set_profiling) */
    int64_t last_query_bytes; /* algorithmic bytes that query scanned */
    int32_t last_path;        /* CV_PATH_SCAN / CV_PATH_MMA */
    int32_t dim_padded;       /* row stride in elements */
} cv_stats_t;

/* Lifecycle -- replaces client.get_or_create_collection (metrics_semantic_metadata_store.py:60-69). */
int cv_create(const cv_config* cfg, cv_index** out);
void cv_destroy(cv_index* idx);

/* K1: normalise (cosine) + bf16 cast + write.  rows == NULL or rows[i] < 0 => append; otherwise
 * overwrite that local row in place (upsert of an existing id, metrics_semantic_metadata_store.py:236-238).
 * emb: [n][dim] fp32.  ns_codes / label_codes: [n] or NULL (code 0).  out_rows: [n] or NULL. */
int cv_upsert(cv_index* idx, const int64_t* rows, const float* emb, const uint32_t* ns_codes,
              const uint32_t* label_codes, int64_t n, int64_t* out_rows);
/* Same with emb / ns_codes / label_codes in device memory (bulk loads generated on the GPU).
 * rows / out_rows stay host pointers.  Non-finite checking is done on the device. */
int cv_upsert_device(cv_index* idx, const int64_t* rows, const float* d_emb, const uint32_t* d_ns_codes,
                     const uint32_t* d_label_codes, int64_t n, int64_t* out_rows);

/* K2 / K3 (+ fused merge): replaces Collection.query (metrics_semantic_metadata_store.py:314-316).
 * q: [B][dim] fp32 (raw, normalised on the device).  out_scores [B][k] cosine (or dot) descending,
 * out_rows [B][k] global rows, -1 padded; out_counts [B]. */
int cv_query(cv_index* idx, const float* q, const cv_query_params* p, float* out_scores,
             int64_t* out_rows, int32_t* out_counts);
/* Asynchronous, device-resident variant: d_q [B][dim] fp32 and d_out_keys [B][k] u64 are device
 * memory; work is enqueued on `stream` (a cudaStream_t; NULL = the index's own stream; pass
 * cudaStreamLegacy / cudaStreamPerThread to name a default stream) and NOT synchronised. */
int cv_query_device(cv_index* idx, const float* d_q, const cv_query_params* p, uint64_t* d_out_keys,
                    void* stream);

/* Fused multi-GPU search (row-sharded corpus, one process per GPU).  Every rank owns an exchange buffer of
 * cv_peer_exchange_bytes() zero-initialised device bytes that is mapped into all peers (CUDA IPC / torch
 * symmetric memory); `buffers[g]` is rank g's buffer as addressable from THIS process.  The kernel that
 * finishes a query's per-GPU top-k stores it into every peer's buffer over NVLink, raises a sequence flag,
 * waits for the peers' flags and merges -- no collective launch.  All ranks must issue the same batches in
 * the same order with the same `seq` (1, 2, 3, ...).  d_out_keys receives the merged global top-k. */
#define CV_MAX_PEERS 8
typedef struct cv_peer_exchange {
    int32_t world;       /* number of ranks, 2..CV_MAX_PEERS */
    int32_t rank;        /* this rank */
    int32_t max_queries; /* capacity: queries per batch */
    int32_t max_k;       /* capacity: k */
    uint64_t buffers[CV_MAX_PEERS]; /* device addresses */
} cv_peer_exchange;
int64_t cv_peer_exchange_bytes(int32_t world, int32_t max_queries, int32_t max_k);
int cv_query_device_exchange(cv_index* idx, const float* d_q, const cv_query_params* p, const cv_peer_exchange* px,
                             uint64_t seq, uint64_t* d_out_keys, void* stream);
/* The same search with HOST buffers, synchronous like cv_query (every rank passes the same queries): replaces
 * Collection.query (metrics_semantic_metadata_store.py:314-316) when the collection is row-sharded over GPUs.
 * Interactive batches take cv_query's low-latency path (queries in the kernel parameters, keys written to
 * page-locked host memory, completion by a polled flag). */
int cv_query_exchange(cv_index* idx, const float* q, const cv_query_params* p, const cv_peer_exchange* px,
                      uint64_t seq, float* out_scores, int64_t* out_rows, int32_t* out_counts);

/* One PROCESS serving G GPUs (the reference builds one client in one process: codd_lib/codd_lib/client/provider/
 * promql_module.py:59-67): a collection striped over G devices -- shard g is a cv_index created with row_stride = G
 * and row_offset = g, i.e. global row r lives on shard r % G at local row r / G, so appends stay balanced and ties
 * still break by global row id.  cv_group_create allocates one exchange buffer per device and maps it into the
 * others (cudaDeviceEnablePeerAccess); cv_group_query is cv_query for the whole collection: it puts the batch in
 * flight on every device (same kernels, fused NVLink exchange as in cv_query_exchange), waits for shard 0's merged
 * result and returns it.  masks: NULL, or one bitmask per shard over that shard's LOCAL rows (entries may be NULL).
 * p->mask is ignored.  Calls on one group are serialised.  The group does not own its shards. */
typedef struct cv_group cv_group;
int cv_group_create(cv_index* const* shards, int32_t n_shards, int32_t max_queries, int32_t max_k, cv_group** out);
void cv_group_destroy(cv_group* grp);
int cv_group_query(cv_group* grp, const float* q, const cv_query_params* p, const uint32_t* const* masks,
                   float* out_scores, int64_t* out_rows, int32_t* out_counts);

/* K4: merge G per-shard key lists [G][B][k_in] (device) into [B][k_out] (device) on `stream`
 * of `device`.  Used after the NCCL all-gather of per-GPU top-k. */
int cv_merge_keys_device(int32_t device, const uint64_t* d_keys_in, int32_t G, int32_t B, int32_t k_in,
                         int32_t k_out, uint64_t* d_keys_out, void* stream);

/* Host helpers (no device needed). */
void cv_decode_keys(const uint64_t* keys, int64_t n, float* out_scores, int64_t* out_rows);
uint64_t cv_encode_key(float score, uint32_t global_row);

/* Read back what K1 stored for local rows [n]: bf16 bit patterns [n][dim] and scale [n]
 * (either may be NULL) -- replaces Collection.get(include=["embeddings"]) and lets tests run the
 * oracle on the same stored embeddings. */
int cv_get_rows(cv_index* idx, const int64_t* rows, int64_t n, uint16_t* out_bf16, float* out_scale,
                uint32_t* out_ns, uint32_t* out_label);

/* Collection.delete (chromadb API; the reference's store never deletes, the indexer's re-runs upsert): mark local
 * rows dead.  A dead row keeps its slot -- rows stay dense -- but no query returns it (its scale becomes NaN, which
 * every kernel treats as "not eligible"), so deletes cost nothing per query; cv_upsert into the slot revives it. */
int cv_delete_rows(cv_index* idx, const int64_t* rows, int64_t n);
/* Collection.update(metadatas=...) that changes a row's namespace / label: rewrite the predicate codes of local rows
 * in place (either array may be NULL = leave that column). */
int cv_set_codes(cv_index* idx, const int64_t* rows, const uint32_t* ns_codes, const uint32_t* label_codes, int64_t n);

/* Append n rows verbatim -- bf16 bit patterns [n][dim], scale [n], codes [n] (NULL = 0), all host memory,
 * exactly what cv_get_rows returned -- without re-normalising: restores a snapshot (the role Chroma's
 * on-disk volume plays for the reference, docker-compose.yml:8-9).  out_rows as for cv_upsert. */
int cv_load_rows(cv_index* idx, const uint16_t* bf16, const float* scale, const uint32_t* ns_codes,
                 const uint32_t* label_codes, int64_t n, int64_t* out_rows);

/* Testing aid: run the tensor-core path (K3) for B queries and return EVERY score it computed,
 * out_scores[B][count] (host), so that tests can compare the whole score matrix with the oracle.
 * path: CV_PATH_MMA, CV_PATH_MMA_1CTA or CV_PATH_MMA_2CTA; precision: CV_PRECISION_*. */
int cv_debug_mma_scores(cv_index* idx, const float* q, int32_t B, int32_t path, int32_t precision, float* out_scores);

int32_t cv_device_count(void); /* CUDA devices visible to this process (0 = none: every compute call will fail) */
int64_t cv_count(const cv_index* idx);
int32_t cv_dim(const cv_index* idx);
int cv_set_profiling(cv_index* idx, int32_t enabled);
int cv_stats(cv_index* idx, cv_stats_t* out);
int cv_synchronize(cv_index* idx);

const char* cv_last_error(void);
const char* cv_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CODDVEC_H_ */
