/* TEST INFRASTRUCTURE ONLY -- plain-C twin of oracle/brute_force.py.
 *
 * CPU restatement of the cosine / inner-product top-k that runs behind the
 * reference's vector-DB boundary (call sites:
 *   codd_dal/metrics/metrics_semantic_metadata_store.py:236-238  collection.upsert
 *   codd_dal/metrics/metrics_semantic_metadata_store.py:314-316  collection.query
 *   codd_dal/metrics/metrics_semantic_metadata_store.py:336      similarity = 1 - distance).
 * The arithmetic itself is in the un-vendored third-party `chromadb` package
 * (>=0.4.0 / >=1.4.0; hnswlib cosine space = normalise + 1 - dot, fp32).
 * PARITY UNPINNED: the reference holds no golden vectors for this path; see
 * oracle/brute_force.py for what is and is not pinned.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  It is never on the product path.
 *
 * Build: make -C oracle   (gcc -O3 -march=native -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define CVO_ANY 0xFFFFFFFFu
#define CVO_COSINE 0
#define CVO_IP 1

static inline uint16_t f32_to_bf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return 0x7FFF; /* NaN */
    uint32_t lsb = (u >> 16) & 1u;
    return (uint16_t)((u + 0x7FFFu + lsb) >> 16);
}

static inline float bf16_to_f32(uint16_t b) {
    uint32_t u = ((uint32_t)b) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* (score, row) -> u64 whose unsigned order is (score desc, row asc); 0 = empty. */
static inline uint64_t pack_key(float s, uint32_t row) {
    if (s != s) s = -INFINITY;
    s = s + 0.0f;
    uint32_t u;
    memcpy(&u, &s, 4);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((uint64_t)u << 32) | (uint64_t)(0xFFFFFFFFu - row);
}

/* K1 twin: normalise + bf16 store + scale. */
void cvo_normalise_store(const float* x, int64_t n, int32_t D, int32_t metric,
                         uint16_t* stored, float* scale) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        const float* r = x + i * (int64_t)D;
        uint16_t* o = stored + i * (int64_t)D;
        if (metric == CVO_IP) {
            for (int d = 0; d < D; ++d) o[d] = f32_to_bf16(r[d]);
            scale[i] = 1.0f;
            continue;
        }
        double ss = 0.0;
        for (int d = 0; d < D; ++d) ss += (double)r[d] * (double)r[d];
        float norm = (float)sqrt(ss);
        float safe = norm > 0.0f ? norm : 1.0f;
        double ss2 = 0.0;
        for (int d = 0; d < D; ++d) {
            uint16_t b = f32_to_bf16(r[d] / safe);
            o[d] = b;
            double v = (double)bf16_to_f32(b);
            ss2 += v * v;
        }
        scale[i] = ss2 > 0.0 ? (float)(1.0 / sqrt(ss2)) : 0.0f;
    }
}

void cvo_normalise_query(const float* q, int32_t B, int32_t D, int32_t metric, float* qhat) {
    for (int b = 0; b < B; ++b) {
        const float* r = q + (int64_t)b * D;
        float* o = qhat + (int64_t)b * D;
        if (metric == CVO_IP) {
            memcpy(o, r, sizeof(float) * (size_t)D);
            continue;
        }
        double ss = 0.0;
        for (int d = 0; d < D; ++d) ss += (double)r[d] * (double)r[d];
        float norm = (float)sqrt(ss);
        float safe = norm > 0.0f ? norm : 1.0f;
        for (int d = 0; d < D; ++d) o[d] = r[d] / safe;
    }
}

/* sorted-descending insertion into a k-list of keys (0 = empty). */
static inline void list_insert(uint64_t* list, int k, uint64_t key) {
    if (key <= list[k - 1]) return;
    int p = k - 1;
    while (p > 0 && list[p - 1] < key) {
        list[p] = list[p - 1];
        --p;
    }
    list[p] = key;
}

static inline float dot_bf16(const uint16_t* e, const float* q, int D) {
    /* 16 fixed partial sums (SIMD-friendly, deterministic order) then a tree. */
    float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    int d = 0;
    for (; d + 16 <= D; d += 16)
        for (int j = 0; j < 16; ++j) acc[j] += bf16_to_f32(e[d + j]) * q[d + j];
    for (int j = 0; d < D; ++d, ++j) acc[j] += bf16_to_f32(e[d]) * q[d];
    for (int w = 8; w >= 1; w >>= 1)
        for (int j = 0; j < w; ++j) acc[j] += acc[j + w];
    return acc[0];
}

/* K2/K3 twin.  qhat: already-normalised queries [B][D].  out_keys: [B][k] desc, 0-padded.
 * threads <= 0 -> all OpenMP threads.  Returns the number of threads used. */
int32_t cvo_query(const uint16_t* stored, const float* scale, int64_t n, int32_t D,
                  const float* qhat, int32_t B, int32_t k, const uint32_t* mask,
                  const uint32_t* ns, uint32_t ns_code, const uint32_t* label,
                  uint32_t label_code, int64_t row_offset, uint64_t* out_keys,
                  int32_t threads) {
    int nt = 1;
#ifdef _OPENMP
    nt = threads > 0 ? threads : omp_get_max_threads();
#endif
    size_t per = (size_t)B * (size_t)k;
    uint64_t* lists = (uint64_t*)calloc((size_t)nt * per, sizeof(uint64_t));
#pragma omp parallel num_threads(nt)
    {
        int t = 0;
#ifdef _OPENMP
        t = omp_get_thread_num();
#endif
        uint64_t* mine = lists + (size_t)t * per;
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            if (mask && !((mask[i >> 5] >> (i & 31)) & 1u)) continue;
            if (ns_code != CVO_ANY && ns[i] != ns_code) continue;
            if (label_code != CVO_ANY && label[i] != label_code) continue;
            const uint16_t* e = stored + i * (int64_t)D;
            float sc = scale[i];
            if (sc != sc) continue; /* scale = NaN marks a deleted row (cv_delete_rows) */
            uint32_t row = (uint32_t)(row_offset + i);
            for (int b = 0; b < B; ++b) {
                float s = dot_bf16(e, qhat + (int64_t)b * D, D) * sc;
                list_insert(mine + (size_t)b * k, k, pack_key(s, row));
            }
        }
    }
    for (int b = 0; b < B; ++b) {
        uint64_t* o = out_keys + (size_t)b * k;
        memset(o, 0, sizeof(uint64_t) * (size_t)k);
        for (int t = 0; t < nt; ++t) {
            const uint64_t* l = lists + (size_t)t * per + (size_t)b * k;
            for (int j = 0; j < k && l[j]; ++j) list_insert(o, k, l[j]);
        }
    }
    free(lists);
    return nt;
}

/* K4 twin: merge G per-shard [B][kin] key lists into [B][k]. */
void cvo_merge_keys(const uint64_t* in, int32_t G, int32_t B, int32_t kin, int32_t k,
                    uint64_t* out) {
    for (int b = 0; b < B; ++b) {
        uint64_t* o = out + (size_t)b * k;
        memset(o, 0, sizeof(uint64_t) * (size_t)k);
        for (int g = 0; g < G; ++g) {
            const uint64_t* l = in + ((size_t)g * B + b) * kin;
            for (int j = 0; j < kin; ++j)
                if (l[j]) list_insert(o, k, l[j]);
        }
    }
}
