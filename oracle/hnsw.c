/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the approximate index the reference queries.
 *
 * What it restates.  The reference asks its vector DB for a cosine HNSW index
 *   codd_dal/metrics/metrics_semantic_metadata_store.py:60-69
 *       get_or_create_collection(metadata={"hnsw:space": "cosine", "hnsw:construction_ef": 200,
 *                                          "hnsw:search_ef": 100, "hnsw:M": 16})
 *   codd_dal/metrics/metrics_semantic_metadata_store.py:236-238   collection.upsert  (-> graph insertion)
 *   codd_dal/metrics/metrics_semantic_metadata_store.py:314-316   collection.query   (-> graph search)
 * and the arithmetic lives in the un-vendored third-party dependency `chromadb` (>=0.4.0 in
 * codd_dal/pyproject.toml:11, >=1.4.0 in codd_lib/pyproject.toml:11), which bundles `hnswlib`.  Neither is
 * under /root/reference nor installable here, so this file follows the PUBLISHED algorithm:
 *   Malkov & Yashunin, "Efficient and robust approximate nearest neighbor search using Hierarchical
 *   Navigable Small World graphs", IEEE TPAMI 2018 -- Alg. 1 (INSERT), Alg. 2 (SEARCH-LAYER),
 *   Alg. 4 (SELECT-NEIGHBORS-HEURISTIC, without extendCandidates / keepPrunedConnections, which is the
 *   variant hnswlib ships) and Alg. 5 (K-NN-SEARCH), with hnswlib's documented parameter conventions:
 *   maxM = M above level 0, maxM0 = 2*M at level 0, level multiplier mL = 1/ln(M), the new element's own
 *   links are chosen with M at every level, search ef = max(ef_search, k), and the "cosine" space =
 *   L2-normalise at insert and at query, distance = 1 - <a, b> in fp32.
 *
 * PARITY UNPINNED.  Level draws come from this file's own splitmix64 stream (hnswlib uses
 * std::default_random_engine(100)), so the graph is A valid HNSW graph for these parameters, not
 * byte-for-byte hnswlib's.  It is used for two things only: (1) the "recall@k against the approximate
 * HNSW answer" report the north-star asks for (tests/bench_configs.py cfg1, bench.py `configs`), and
 * (2) a CPU timing of the reference's query algorithm on the reference's own CPU-runnable case
 * (BASELINE.json configs[0]).  It is never on the product path.
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { float d; int id; } pair_t;

/* ---- binary heap of (distance, id); `sign` = +1 keeps the LARGEST distance on top, -1 the smallest ---- */
typedef struct { pair_t* a; int n, cap; int sign; } heap_t;

static void heap_init(heap_t* h, int cap, int sign) {
    h->a = (pair_t*)malloc(sizeof(pair_t) * (size_t)(cap > 16 ? cap : 16));
    h->n = 0;
    h->cap = cap > 16 ? cap : 16;
    h->sign = sign;
}
static void heap_free(heap_t* h) { free(h->a); h->a = NULL; }
static inline int heap_before(const heap_t* h, pair_t x, pair_t y) { /* x belongs above y */
    return h->sign > 0 ? (x.d > y.d) : (x.d < y.d);
}
static void heap_push(heap_t* h, pair_t v) {
    if (h->n == h->cap) {
        h->cap *= 2;
        h->a = (pair_t*)realloc(h->a, sizeof(pair_t) * (size_t)h->cap);
    }
    int i = h->n++;
    while (i > 0) {
        int p = (i - 1) >> 1;
        if (!heap_before(h, v, h->a[p])) break;
        h->a[i] = h->a[p];
        i = p;
    }
    h->a[i] = v;
}
static pair_t heap_pop(heap_t* h) {
    pair_t top = h->a[0];
    pair_t v = h->a[--h->n];
    int i = 0;
    for (;;) {
        int c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && heap_before(h, h->a[c + 1], h->a[c])) ++c;
        if (!heap_before(h, h->a[c], v)) break;
        h->a[i] = h->a[c];
        i = c;
    }
    if (h->n > 0) h->a[i] = v;
    return top;
}

/* ---- the index ---- */
typedef struct {
    int dim, M, maxM0, efc;
    int64_t cap, n;
    float* data;     /* [cap][dim], L2-normalised when cosine */
    int* level;      /* [cap] */
    int* link0;      /* [cap][1 + maxM0]: count, ids */
    int** linkup;    /* [cap] -> [level][1 + M] */
    int enter, maxlevel;
    double mult;
    uint64_t rng;
    int cosine;
    int64_t dist_evals; /* statistics: distance computations since the last reset */
} hnsw_t;

static inline uint64_t splitmix64(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

static inline float dist_ip(const float* a, const float* b, int dim) {
    float s = 0.0f;
#pragma omp simd reduction(+ : s)
    for (int i = 0; i < dim; ++i) s += a[i] * b[i];
    return 1.0f - s;
}

static inline int* links_of(const hnsw_t* h, int id, int lev) {
    return lev == 0 ? h->link0 + (size_t)id * (size_t)(1 + h->maxM0) : h->linkup[id] + (size_t)(lev - 1) * (size_t)(1 + h->M);
}

hnsw_t* hnsw_create(int dim, int M, int ef_construction, int64_t capacity, int cosine, uint64_t seed) {
    hnsw_t* h = (hnsw_t*)calloc(1, sizeof(hnsw_t));
    if (!h) return NULL;
    h->dim = dim;
    h->M = M;
    h->maxM0 = 2 * M;
    h->efc = ef_construction > M ? ef_construction : M;
    h->cap = capacity;
    h->data = (float*)malloc(sizeof(float) * (size_t)capacity * (size_t)dim);
    h->level = (int*)calloc((size_t)capacity, sizeof(int));
    h->link0 = (int*)calloc((size_t)capacity * (size_t)(1 + h->maxM0), sizeof(int));
    h->linkup = (int**)calloc((size_t)capacity, sizeof(int*));
    h->enter = -1;
    h->maxlevel = -1;
    h->mult = 1.0 / log((double)M);
    h->rng = seed;
    h->cosine = cosine;
    if (!h->data || !h->level || !h->link0 || !h->linkup) return NULL;
    return h;
}

void hnsw_destroy(hnsw_t* h) {
    if (!h) return;
    for (int64_t i = 0; i < h->n; ++i) free(h->linkup[i]);
    free(h->linkup);
    free(h->link0);
    free(h->level);
    free(h->data);
    free(h);
}

int64_t hnsw_count(const hnsw_t* h) { return h->n; }
int hnsw_max_level(const hnsw_t* h) { return h->maxlevel; }
int64_t hnsw_dist_evals(hnsw_t* h, int reset) {
    int64_t v = h->dist_evals;
    if (reset) h->dist_evals = 0;
    return v;
}

/* Alg. 2 SEARCH-LAYER: ef closest elements to q reachable from `ep` on layer `lev`.
 * `visited` is a per-caller tag array ([cap] uint32), `tag` the current generation.  Result in `top`
 * (largest distance on top, at most ef entries). */
static int64_t search_layer(const hnsw_t* h, const float* q, int ep, float ep_dist, int ef, int lev, uint32_t* visited,
                            uint32_t tag, heap_t* top, heap_t* cand) {
    int64_t evals = 0;
    top->n = 0;
    cand->n = 0;
    pair_t e = {ep_dist, ep};
    heap_push(top, e);
    heap_push(cand, e);
    visited[ep] = tag;
    float bound = ep_dist;
    while (cand->n > 0) {
        pair_t c = cand->a[0];
        if (c.d > bound && top->n >= ef) break; /* every remaining candidate is farther than the worst kept */
        heap_pop(cand);
        const int* l = links_of(h, c.id, lev);
        const int cnt = l[0];
        for (int j = 1; j <= cnt; ++j) {
            const int nb = l[j];
            if (visited[nb] == tag) continue;
            visited[nb] = tag;
            const float d = dist_ip(q, h->data + (size_t)nb * (size_t)h->dim, h->dim);
            ++evals;
            if (top->n < ef || d < bound) {
                pair_t p = {d, nb};
                heap_push(cand, p);
                heap_push(top, p);
                if (top->n > ef) heap_pop(top);
                bound = top->a[0].d;
            }
        }
    }
    return evals;
}

static int cmp_pair_asc(const void* x, const void* y) {
    const pair_t* a = (const pair_t*)x;
    const pair_t* b = (const pair_t*)y;
    if (a->d < b->d) return -1;
    if (a->d > b->d) return 1;
    return (a->id > b->id) - (a->id < b->id);
}

/* Alg. 4 SELECT-NEIGHBORS-HEURISTIC (no candidate extension, pruned connections dropped): walk the
 * candidates from the closest outwards and keep one only if it is closer to the base point than to
 * every neighbour kept so far.  cands[] holds distances TO THE BASE POINT; returns the kept count. */
static int select_heuristic(hnsw_t* h, pair_t* cands, int nc, int M, int* out) {
    if (nc <= M) {
        for (int i = 0; i < nc; ++i) out[i] = cands[i].id;
        return nc;
    }
    qsort(cands, (size_t)nc, sizeof(pair_t), cmp_pair_asc);
    int kept = 0;
    for (int i = 0; i < nc && kept < M; ++i) {
        const float* ci = h->data + (size_t)cands[i].id * (size_t)h->dim;
        int good = 1;
        for (int j = 0; j < kept; ++j) {
            const float d = dist_ip(ci, h->data + (size_t)out[j] * (size_t)h->dim, h->dim);
            ++h->dist_evals;
            if (d < cands[i].d) { good = 0; break; }
        }
        if (good) out[kept++] = cands[i].id;
    }
    return kept;
}

/* Alg. 1 INSERT, one element. */
static void insert_one(hnsw_t* h, const float* x, uint32_t* visited, uint32_t* tag, heap_t* top, heap_t* cand,
                       pair_t* scratch, int* sel) {
    const int id = (int)h->n;
    float* v = h->data + (size_t)id * (size_t)h->dim;
    if (h->cosine) {
        double ss = 0.0;
        for (int i = 0; i < h->dim; ++i) ss += (double)x[i] * (double)x[i];
        const float nrm = (float)sqrt(ss);
        const float inv = nrm > 0.0f ? 1.0f / nrm : 1.0f;
        for (int i = 0; i < h->dim; ++i) v[i] = x[i] * inv;
    } else {
        memcpy(v, x, sizeof(float) * (size_t)h->dim);
    }
    /* level ~ floor(-ln(U) * mL) */
    double u = ((double)(splitmix64(&h->rng) >> 11) + 1.0) * (1.0 / 9007199254740993.0);
    int lev = (int)(-log(u) * h->mult);
    h->level[id] = lev;
    h->linkup[id] = lev > 0 ? (int*)calloc((size_t)lev * (size_t)(1 + h->M), sizeof(int)) : NULL;
    h->link0[(size_t)id * (size_t)(1 + h->maxM0)] = 0;
    h->n = id + 1;
    if (h->enter < 0) {
        h->enter = id;
        h->maxlevel = lev;
        return;
    }
    int ep = h->enter;
    float epd = dist_ip(v, h->data + (size_t)ep * (size_t)h->dim, h->dim);
    ++h->dist_evals;
    /* greedy descent (ef = 1) through the layers above the new element's top layer */
    for (int l = h->maxlevel; l > lev; --l) {
        int changed = 1;
        while (changed) {
            changed = 0;
            const int* ls = links_of(h, ep, l);
            for (int j = 1; j <= ls[0]; ++j) {
                const float d = dist_ip(v, h->data + (size_t)ls[j] * (size_t)h->dim, h->dim);
                ++h->dist_evals;
                if (d < epd) { epd = d; ep = ls[j]; changed = 1; }
            }
        }
    }
    for (int l = lev < h->maxlevel ? lev : h->maxlevel; l >= 0; --l) {
        ++*tag;
        h->dist_evals += search_layer(h, v, ep, epd, h->efc, l, visited, *tag, top, cand);
        /* candidates -> scratch (distance to the new element) */
        int nc = top->n;
        for (int i = 0; i < nc; ++i) scratch[i] = top->a[i];
        /* the closest found element seeds the next layer */
        int best = 0;
        for (int i = 1; i < nc; ++i)
            if (scratch[i].d < scratch[best].d) best = i;
        ep = scratch[best].id;
        epd = scratch[best].d;
        const int ns = select_heuristic(h, scratch, nc, h->M, sel);
        int* mine = links_of(h, id, l);
        mine[0] = ns;
        for (int i = 0; i < ns; ++i) mine[1 + i] = sel[i];
        /* back-links, shrinking a full neighbour list with the same heuristic */
        const int maxm = l == 0 ? h->maxM0 : h->M;
        for (int i = 0; i < ns; ++i) {
            const int nb = sel[i];
            int* ls = links_of(h, nb, l);
            if (ls[0] < maxm) {
                ls[1 + ls[0]] = id;
                ++ls[0];
            } else {
                const float* nbv = h->data + (size_t)nb * (size_t)h->dim;
                pair_t* c2 = scratch; /* scratch is free again: sel[] holds what we still need */
                int n2 = 0;
                c2[n2].d = dist_ip(nbv, v, h->dim);
                c2[n2++].id = id;
                for (int j = 1; j <= ls[0]; ++j) {
                    c2[n2].d = dist_ip(nbv, h->data + (size_t)ls[j] * (size_t)h->dim, h->dim);
                    c2[n2++].id = ls[j];
                }
                h->dist_evals += n2;
                int keep[128];
                const int nk = select_heuristic(h, c2, n2, maxm, keep);
                ls[0] = nk;
                for (int j = 0; j < nk; ++j) ls[1 + j] = keep[j];
            }
        }
    }
    if (lev > h->maxlevel) {
        h->maxlevel = lev;
        h->enter = id;
    }
}

/* Insert n fp32 rows (single-threaded, like one Collection.upsert after another).  0 = ok. */
int hnsw_add(hnsw_t* h, const float* x, int64_t n) {
    if (h->n + n > h->cap || h->M > 63) return -1;
    uint32_t* visited = (uint32_t*)calloc((size_t)h->cap, sizeof(uint32_t));
    heap_t top, cand;
    heap_init(&top, h->efc + 2, +1);
    heap_init(&cand, 4 * h->efc, -1);
    const int sc = (h->efc > h->maxM0 ? h->efc : h->maxM0) + 2;
    pair_t* scratch = (pair_t*)malloc(sizeof(pair_t) * (size_t)sc);
    int* sel = (int*)malloc(sizeof(int) * (size_t)(h->maxM0 + 1));
    uint32_t tag = 0;
    for (int64_t i = 0; i < n; ++i) insert_one(h, x + (size_t)i * (size_t)h->dim, visited, &tag, &top, &cand, scratch, sel);
    free(sel);
    free(scratch);
    heap_free(&cand);
    heap_free(&top);
    free(visited);
    return 0;
}

/* ---- concurrent insertion (what `add_items(..., num_threads)` of the library the reference uses does for a
 * batch).  Test-infrastructure convenience: it makes the recall / timing report practical at 1M rows.  One
 * lock per element guards its link lists; a thread never holds two locks (read a list: lock, copy, unlock;
 * add a back-link: lock the neighbour, modify, unlock), so there is no lock-order to get wrong.  Vectors and
 * levels of the whole batch are stored before the threads start; an element becomes reachable when its first
 * back-link lands.  The resulting graph depends on the interleaving (as the library's does). ---- */
#ifdef _OPENMP
typedef struct {
    omp_lock_t* node;
    omp_lock_t global; /* entry point + top level */
} par_locks_t;

static int read_links(const hnsw_t* h, par_locks_t* L, int id, int lev, int* buf) {
    omp_set_lock(&L->node[id]);
    const int* l = links_of(h, id, lev);
    const int cnt = l[0];
    memcpy(buf, l + 1, sizeof(int) * (size_t)cnt);
    omp_unset_lock(&L->node[id]);
    return cnt;
}

static int64_t search_layer_par(const hnsw_t* h, par_locks_t* L, const float* q, int ep, float ep_dist, int ef, int lev,
                                uint32_t* visited, uint32_t tag, heap_t* top, heap_t* cand, int* nbuf) {
    int64_t evals = 0;
    top->n = 0;
    cand->n = 0;
    pair_t e = {ep_dist, ep};
    heap_push(top, e);
    heap_push(cand, e);
    visited[ep] = tag;
    float bound = ep_dist;
    while (cand->n > 0) {
        pair_t c = cand->a[0];
        if (c.d > bound && top->n >= ef) break;
        heap_pop(cand);
        const int cnt = read_links(h, L, c.id, lev, nbuf);
        for (int j = 0; j < cnt; ++j) {
            const int nb = nbuf[j];
            if (visited[nb] == tag) continue;
            visited[nb] = tag;
            const float d = dist_ip(q, h->data + (size_t)nb * (size_t)h->dim, h->dim);
            ++evals;
            if (top->n < ef || d < bound) {
                pair_t p = {d, nb};
                heap_push(cand, p);
                heap_push(top, p);
                if (top->n > ef) heap_pop(top);
                bound = top->a[0].d;
            }
        }
    }
    return evals;
}

/* select_heuristic without the shared counter (threads keep their own) */
static int select_heuristic_par(const hnsw_t* h, pair_t* cands, int nc, int M, int* out, int64_t* evals) {
    if (nc <= M) {
        for (int i = 0; i < nc; ++i) out[i] = cands[i].id;
        return nc;
    }
    qsort(cands, (size_t)nc, sizeof(pair_t), cmp_pair_asc);
    int kept = 0;
    for (int i = 0; i < nc && kept < M; ++i) {
        const float* ci = h->data + (size_t)cands[i].id * (size_t)h->dim;
        int good = 1;
        for (int j = 0; j < kept; ++j) {
            const float d = dist_ip(ci, h->data + (size_t)out[j] * (size_t)h->dim, h->dim);
            ++*evals;
            if (d < cands[i].d) { good = 0; break; }
        }
        if (good) out[kept++] = cands[i].id;
    }
    return kept;
}

static void insert_par(hnsw_t* h, par_locks_t* L, int id, uint32_t* visited, uint32_t* tag, heap_t* top, heap_t* cand,
                       pair_t* scratch, int* sel, int* nbuf, int64_t* evals) {
    const float* v = h->data + (size_t)id * (size_t)h->dim;
    const int lev = h->level[id];
    omp_set_lock(&L->global);
    int ep = h->enter;
    const int maxl = h->maxlevel;
    omp_unset_lock(&L->global);
    float epd = dist_ip(v, h->data + (size_t)ep * (size_t)h->dim, h->dim);
    ++*evals;
    for (int l = maxl; l > lev; --l) {
        int changed = 1;
        while (changed) {
            changed = 0;
            const int cnt = read_links(h, L, ep, l, nbuf);
            for (int j = 0; j < cnt; ++j) {
                const float d = dist_ip(v, h->data + (size_t)nbuf[j] * (size_t)h->dim, h->dim);
                ++*evals;
                if (d < epd) { epd = d; ep = nbuf[j]; changed = 1; }
            }
        }
    }
    for (int l = lev < maxl ? lev : maxl; l >= 0; --l) {
        ++*tag;
        *evals += search_layer_par(h, L, v, ep, epd, h->efc, l, visited, *tag, top, cand, nbuf);
        int nc = 0;
        for (int i = 0; i < top->n; ++i)
            if (top->a[i].id != id) scratch[nc++] = top->a[i]; /* (a concurrent back-link can make me reachable) */
        if (nc == 0) continue;
        int best = 0;
        for (int i = 1; i < nc; ++i)
            if (scratch[i].d < scratch[best].d) best = i;
        ep = scratch[best].id;
        epd = scratch[best].d;
        const int ns = select_heuristic_par(h, scratch, nc, h->M, sel, evals);
        const int maxm = l == 0 ? h->maxM0 : h->M;
        omp_set_lock(&L->node[id]);
        {   /* merge with back-links other threads may already have given me */
            int* mine = links_of(h, id, l);
            int have = mine[0];
            for (int i = 0; i < ns && have < maxm; ++i) {
                int dup = 0;
                for (int j = 1; j <= have; ++j)
                    if (mine[j] == sel[i]) { dup = 1; break; }
                if (!dup) mine[++have] = sel[i];
            }
            mine[0] = have;
        }
        omp_unset_lock(&L->node[id]);
        for (int i = 0; i < ns; ++i) {
            const int nb = sel[i];
            omp_set_lock(&L->node[nb]);
            int* ls = links_of(h, nb, l);
            int dup = 0;
            for (int j = 1; j <= ls[0]; ++j)
                if (ls[j] == id) { dup = 1; break; }
            if (!dup) {
                if (ls[0] < maxm) {
                    ls[1 + ls[0]] = id;
                    ++ls[0];
                } else {
                    const float* nbv = h->data + (size_t)nb * (size_t)h->dim;
                    pair_t* c2 = scratch;
                    int n2 = 0;
                    c2[n2].d = dist_ip(nbv, v, h->dim);
                    c2[n2++].id = id;
                    for (int j = 1; j <= ls[0]; ++j) {
                        c2[n2].d = dist_ip(nbv, h->data + (size_t)ls[j] * (size_t)h->dim, h->dim);
                        c2[n2++].id = ls[j];
                    }
                    *evals += n2;
                    int keep[128];
                    const int nk = select_heuristic_par(h, c2, n2, maxm, keep, evals);
                    ls[0] = nk;
                    for (int j = 0; j < nk; ++j) ls[1 + j] = keep[j];
                }
            }
            omp_unset_lock(&L->node[nb]);
        }
    }
    if (lev > maxl) {
        omp_set_lock(&L->global);
        if (lev > h->maxlevel) {
            h->maxlevel = lev;
            h->enter = id;
        }
        omp_unset_lock(&L->global);
    }
}
#endif

/* Insert n fp32 rows with `threads` OpenMP threads (<= 1 or no OpenMP: the sequential hnsw_add).  0 = ok. */
int hnsw_add_parallel(hnsw_t* h, const float* x, int64_t n, int threads) {
#ifndef _OPENMP
    (void)threads;
    return hnsw_add(h, x, n);
#else
    if (threads <= 1 || n < 2) return hnsw_add(h, x, n);
    if (h->n + n > h->cap || h->M > 63) return -1;
    int64_t first = 0;
    if (h->n == 0) { /* the graph needs an entry point before anyone can search it */
        if (hnsw_add(h, x, 1) != 0) return -1;
        first = 1;
    }
    const int64_t base = h->n;
    for (int64_t i = first; i < n; ++i) { /* store vectors, draw levels, allocate link lists -- sequentially */
        const int id = (int)(base + (i - first));
        const float* xi = x + (size_t)i * (size_t)h->dim;
        float* v = h->data + (size_t)id * (size_t)h->dim;
        if (h->cosine) {
            double ss = 0.0;
            for (int d = 0; d < h->dim; ++d) ss += (double)xi[d] * (double)xi[d];
            const float nrm = (float)sqrt(ss);
            const float inv = nrm > 0.0f ? 1.0f / nrm : 1.0f;
            for (int d = 0; d < h->dim; ++d) v[d] = xi[d] * inv;
        } else {
            memcpy(v, xi, sizeof(float) * (size_t)h->dim);
        }
        double u = ((double)(splitmix64(&h->rng) >> 11) + 1.0) * (1.0 / 9007199254740993.0);
        const int lev = (int)(-log(u) * h->mult);
        h->level[id] = lev;
        h->linkup[id] = lev > 0 ? (int*)calloc((size_t)lev * (size_t)(1 + h->M), sizeof(int)) : NULL;
        h->link0[(size_t)id * (size_t)(1 + h->maxM0)] = 0;
    }
    const int64_t total = base + (n - first);
    par_locks_t L;
    L.node = (omp_lock_t*)malloc(sizeof(omp_lock_t) * (size_t)total);
    for (int64_t i = 0; i < total; ++i) omp_init_lock(&L.node[i]);
    omp_init_lock(&L.global);
    h->n = total; /* every id below `total` has a vector, a level and (empty) link lists */
    int64_t evals_total = 0;
#pragma omp parallel num_threads(threads) reduction(+ : evals_total)
    {
        uint32_t* visited = (uint32_t*)calloc((size_t)total, sizeof(uint32_t));
        heap_t top, cand;
        heap_init(&top, h->efc + 2, +1);
        heap_init(&cand, 4 * h->efc, -1);
        const int sc = (h->efc > h->maxM0 ? h->efc : h->maxM0) + 2;
        pair_t* scratch = (pair_t*)malloc(sizeof(pair_t) * (size_t)sc);
        int* sel = (int*)malloc(sizeof(int) * (size_t)(h->maxM0 + 1));
        int* nbuf = (int*)malloc(sizeof(int) * (size_t)(h->maxM0 + 1));
        uint32_t tag = 0;
        int64_t evals = 0;
#pragma omp for schedule(dynamic, 8)
        for (int64_t id = base; id < total; ++id)
            insert_par(h, &L, (int)id, visited, &tag, &top, &cand, scratch, sel, nbuf, &evals);
        evals_total += evals;
        free(nbuf);
        free(sel);
        free(scratch);
        heap_free(&cand);
        heap_free(&top);
        free(visited);
    }
    h->dist_evals += evals_total;
    for (int64_t i = 0; i < total; ++i) omp_destroy_lock(&L.node[i]);
    omp_destroy_lock(&L.global);
    free(L.node);
    return 0;
#endif
}

/* Alg. 5 K-NN-SEARCH for nq queries.  out_ids [nq][k] (-1 padded), out_dist [nq][k] ascending
 * (1 - cosine).  threads <= 1: one query after another (how the reference issues them,
 * store.py:314-316 is one query per call); > 1: OpenMP over queries. */
int hnsw_search(hnsw_t* h, const float* q, int64_t nq, int k, int ef_search, int threads, int64_t* out_ids,
                float* out_dist) {
    const int ef = ef_search > k ? ef_search : k;
    int64_t total_evals = 0;
#ifdef _OPENMP
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads) reduction(+ : total_evals)
#endif
    {
        uint32_t* visited = (uint32_t*)calloc((size_t)(h->n > 0 ? h->n : 1), sizeof(uint32_t));
        float* qn = (float*)malloc(sizeof(float) * (size_t)h->dim);
        heap_t top, cand;
        heap_init(&top, ef + 2, +1);
        heap_init(&cand, 4 * ef, -1);
        uint32_t tag = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int64_t b = 0; b < nq; ++b) {
            int64_t* oi = out_ids + (size_t)b * (size_t)k;
            float* od = out_dist + (size_t)b * (size_t)k;
            for (int j = 0; j < k; ++j) { oi[j] = -1; od[j] = INFINITY; }
            if (h->n == 0) continue;
            const float* qb = q + (size_t)b * (size_t)h->dim;
            if (h->cosine) {
                double ss = 0.0;
                for (int i = 0; i < h->dim; ++i) ss += (double)qb[i] * (double)qb[i];
                const float nrm = (float)sqrt(ss);
                const float inv = nrm > 0.0f ? 1.0f / nrm : 1.0f;
                for (int i = 0; i < h->dim; ++i) qn[i] = qb[i] * inv;
            } else {
                memcpy(qn, qb, sizeof(float) * (size_t)h->dim);
            }
            int ep = h->enter;
            float epd = dist_ip(qn, h->data + (size_t)ep * (size_t)h->dim, h->dim);
            ++total_evals;
            for (int l = h->maxlevel; l > 0; --l) {
                int changed = 1;
                while (changed) {
                    changed = 0;
                    const int* ls = links_of(h, ep, l);
                    for (int j = 1; j <= ls[0]; ++j) {
                        const float d = dist_ip(qn, h->data + (size_t)ls[j] * (size_t)h->dim, h->dim);
                        ++total_evals;
                        if (d < epd) { epd = d; ep = ls[j]; changed = 1; }
                    }
                }
            }
            ++tag;
            total_evals += search_layer(h, qn, ep, epd, ef, 0, visited, tag, &top, &cand);
            while (top.n > k) heap_pop(&top);
            int m = top.n;
            while (top.n > 0) {
                pair_t p = heap_pop(&top);
                oi[top.n] = p.id;
                od[top.n] = p.d;
            }
            (void)m;
        }
        heap_free(&cand);
        heap_free(&top);
        free(qn);
        free(visited);
    }
    h->dist_evals += total_evals;
    return 0;
}
