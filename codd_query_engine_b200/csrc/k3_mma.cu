// k3_mma.cu -- host side of K3 `gemm_topk` (kernel: k3_kernel.cuh): shapes, scratch, tensor maps, launches.
#include "k3_kernel.cuh"
#include <algorithm>
#include <vector>

namespace cv {

using namespace k3;

namespace {

constexpr size_t kSmemLimit = 232448;  // 227 KiB: the per-CTA maximum on sm_100

// qblocks: 16 KiB blocks of the resident query tile (0 when the tile is streamed through the ring)
size_t mma_smem_bytes(int qblocks, int stages, int cg, size_t list_bytes, bool q_resident) {
    const size_t blocks = q_resident ? static_cast<size_t>(qblocks + stages) : 2 * static_cast<size_t>(stages);
    return blocks * kBlockBytes + acc_stages(cg) * kTN * cg * sizeof(float) +
           (2 * kMaxStages + 1 + 2 * kMaxAccStages) * sizeof(uint64_t) + 16 + list_bytes;
}

int mma_stages(int qblocks, int cg, size_t list_bytes, bool q_resident) {
    int s = kMaxStages;
    while (s > 1 && mma_smem_bytes(qblocks, s, cg, list_bytes, q_resident) > kSmemLimit) --s;
    return s;
}

thread_local std::string g_mma_err;

int mma_fail(int code, const std::string& msg) {
    g_mma_err = msg;
    return code;
}

int encode_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint32_t box_outer) {
    const DriverApi& d = driver_api();
    if (!d.ok) return mma_fail(CV_ERR_CUDA, d.why);
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {inner * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kKB), box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = d.tensorMapEncodeTiled(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return mma_fail(CV_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(r));
    return 0;
}

}  // namespace

bool mma_supported(int Dp) { return Dp % kKB == 0 && Dp / kKB <= kMaxKBlocks; }
std::string& mma_error() { return g_mma_err; }

#define MMA_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            (void)cudaGetLastError();                                                                   \
            return mma_fail(e__ == cudaErrorMemoryAllocation ? CV_ERR_OOM : CV_ERR_CUDA,                \
                            std::string(#call) + " failed: " + cudaGetErrorString(e__));                \
        }                                                                                               \
    } while (0)

// The dynamic shared-memory limit is an attribute per function and device, shared by the whole process: raise it to
// the architectural maximum once per (kernel, device) instead of setting it to each launch's need (a later, smaller
// setting would make a concurrent or cached larger launch fail).  Per DEVICE: a striped collection launches the same
// kernels on several GPUs from one process.
int raise_smem_limit(MmaKernel kernel) {
    static std::mutex mu;
    static std::vector<std::pair<int, MmaKernel>> done;
    std::lock_guard<std::mutex> g(mu);
    int dev = 0;
    MMA_CUDA(cudaGetDevice(&dev));
    for (const auto& d : done)
        if (d.first == dev && d.second == kernel) return 0;
    MMA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemLimit)));
    done.emplace_back(dev, kernel);
    return 0;
}

// Self-resetting cross-CTA state for `queries` queries: heads [q][kHeadsLd], gthr [q], cnt [q], done [q].
// All of it is zero between launches (the fused tail's last owner of a query clears what the query dirtied), so it
// is zeroed once, when (re)allocated.
constexpr size_t kDebugBytes = 64 + (296 * 16 + 64 * 6) * sizeof(unsigned long long);  // CODDVEC_K3_DEBUG words + CODDVEC_K3_TIMELINE stamps
constexpr size_t kProgBytes = 296 * sizeof(unsigned long long);                          // super-tile progress words (one per unit)

size_t mma_state_bytes(int queries) {
    return ((static_cast<size_t>(queries) * (kHeadsLd + 3) + 4) * sizeof(uint32_t) + 63) / 64 * 64;
}

int mma_enqueue(const MmaArgs& m, MmaScratch& scratch, std::atomic<int64_t>* launches) {
    if (!mma_supported(m.Dp))
        return mma_fail(CV_ERR_UNSUPPORTED, "tensor-core path supports padded dim <= 4096 (got " + std::to_string(m.Dp) + ")");
    const int kblocks = m.Dp / kKB;
    // CTA pairs (cta_group::2, 256 queries x 256 rows per tile) once a batch fills more than one
    // 128-query tile; single CTAs below that (the scan is HBM-bound there and pairs would idle lanes).
    const int cg = m.variant == 1 ? 1 : (m.variant == 2 ? 2 : (m.B > kQ ? 2 : 1));
    const int tile_rows = kTN * cg, tile_q = kQ * cg;
    const int n_tiles = static_cast<int>((m.n + tile_rows - 1) / tile_rows);
    const int units = m.sm_count / cg;  // CTAs or CTA pairs that fit the chip
    // Query precision.  bf16 queries carry an error of up to 2^-9 * sum|q_i c_i| per score: ~1e-4 on dense rows of
    // >= 128 dimensions, but up to 2^-8 on very short or sparse ones -- outside the 1e-3 the north star allows.  Rows
    // of up to 128 padded elements therefore take the hi/lo split by default (two bf16 tiles, two MMAs per K step:
    // ~2^-17 relative, and the scan is HBM-bound there anyway); CV_PRECISION_SPLIT asks for it at any dimension whose
    // double tile still fits shared memory.
    bool split = m.precision == CV_PRECISION_SPLIT || (m.precision == CV_PRECISION_AUTO && m.Dp <= 128);
    if (split && mma_smem_bytes(2 * kblocks, 2, cg, 0, true) > kSmemLimit) {
        if (m.precision == CV_PRECISION_SPLIT)
            return mma_fail(CV_ERR_UNSUPPORTED, "split-precision queries need the double query tile resident in shared memory: "
                                                "dim <= 384 (got " + std::to_string(m.Dp) + ")");
        split = false;
    }
    const int qblocks = kblocks * (split ? 2 : 1);
    // k > 64: per-query lists go to shared memory if a >= 3-stage ring still fits beside them
    size_t smem_list_bytes = m.k > 64 && !m.dump ? static_cast<size_t>(kQ) * m.k * sizeof(uint64_t) : 0;
    // Short rows keep the query tile resident (a deep corpus-only ring); long rows stream the query's K
    // blocks through the ring with the corpus blocks (7 x 32 KiB stages) -- the tile would crowd the ring out.
    const bool q_resident = split || mma_smem_bytes(kblocks, 6, cg, 0, true) <= kSmemLimit;
    if (smem_list_bytes && mma_smem_bytes(qblocks, 3, cg, smem_list_bytes, q_resident) > kSmemLimit) smem_list_bytes = 0;
    const int stages = mma_stages(qblocks, cg, smem_list_bytes, q_resident);
    const size_t smem = mma_smem_bytes(qblocks, stages, cg, smem_list_bytes, q_resident);
    if (smem > kSmemLimit)
        return mma_fail(CV_ERR_UNSUPPORTED, "tensor-core path: query tile does not fit shared memory at dim " + std::to_string(m.Dp));
    // (CODDVEC_K3_SPLIT_LAUNCHES: bit 0 = separate query-prep launch, bit 1 = separate merge launch -- A/B switches)
    static const int split_launches = getenv("CODDVEC_K3_SPLIT_LAUNCHES") ? atoi(getenv("CODDVEC_K3_SPLIT_LAUNCHES")) : 0;
    // The query tile is built once, by k_prep_queries, and fetched by TMA; k_mma_topk is launched as its programmatic
    // dependent, so the prep hides behind the main kernel's prologue and first corpus loads (same-box A/B: that beat
    // building the tile inside every CTA even for four queries, 0.143 against 0.146 ms at 1M rows).  Only the hi/lo
    // split builds its tiles inside the kernel (its second tile exists in shared memory only).
    // (CODDVEC_K3_FUSED_PREP_MAX_B=<B> builds the tile in-kernel for batches up to B: A/B switch.)
    static const int fused_prep_max_b = getenv("CODDVEC_K3_FUSED_PREP_MAX_B") ? atoi(getenv("CODDVEC_K3_FUSED_PREP_MAX_B")) : 0;
    const bool fused_prep = q_resident && kblocks <= 8 && (split || (m.B <= fused_prep_max_b && !(split_launches & 1)));
    const bool fused_tail = m.k <= 64 && m.dump == nullptr && !(split_launches & 2);          // register-resident lists
    // Epilogue warps: four (thread = query) by default; eight (two threads per query on alternate column chunks) are
    // built for k <= 16 and selectable with CODDVEC_K3_EW=8.
    static const int ew_env = getenv("CODDVEC_K3_EW") ? atoi(getenv("CODDVEC_K3_EW")) : 0;
    const int ew_pref = ew_env ? ew_env : 4;  // (same-box A/B, profiles/README.md: eight never beat four by more than noise once
                                              //  the thresholds were tight; they stay selectable for k <= 16)
    const bool dump = m.dump != nullptr;
    MmaKernel kernel = cg == 2 ? pick_mma_kernel_cg2(dump, m.k, ew_pref) : pick_mma_kernel<1>(dump, m.k, ew_pref);
    const int ew = mma_kernel_epilogue_warps(dump, m.k, ew_pref);
    const int halves = ew / 4;  // list owners per (CTA, query)
    int rc;
    if ((rc = raise_smem_limit(kernel))) return rc;
    // Two-pass first tile (PILOT instantiations): pays where the first tile's insertion burst is a visible share of the
    // scan -- single CTAs with up to 80 tiles each, i.e. corpora of up to ~1.5M rows per GPU (1M x 384: -3 % at
    // B = 16...128, 100k rows: -25 %); at 158 tiles per CTA (3M rows) the pilot kernel's slightly slower tile loop
    // already costs more than the burst it saves (+2.5 %), and CTA pairs measured neutral to negative.
    // (CODDVEC_K3_PILOT=0|1 forces it: A/B switch.)
    static const int pilot_env = getenv("CODDVEC_K3_PILOT") ? atoi(getenv("CODDVEC_K3_PILOT")) : -1;
    MmaKernel pilot_kernel = nullptr;
    if (cg == 1 && ew == 4 && !dump && fused_tail && pilot_env != 0) {
        pilot_kernel = pick_mma_kernel_pilot(m.k);
        if ((rc = raise_smem_limit(pilot_kernel))) return rc;
    }
    // batches larger than one wave of query tiles are processed in slices (stream order keeps the scratch safe)
    int done = 0;
    while (done < m.B) {
        const int Bs = std::min(m.B - done, units * tile_q);
        const int QT = (Bs + tile_q - 1) / tile_q;
        const int Bpad = QT * tile_q;
        int P = units / QT;
        if (P > n_tiles) P = n_tiles;
        if (P < 1) P = 1;
        // work scratch (contents do not matter between launches): [bf16 queries Bpad x Dp (long rows only)][keys Bpad x P x k]
        const size_t q_bytes = fused_prep ? 0 : (static_cast<size_t>(Bpad) * m.Dp * 2 + 1023) / 1024 * 1024;
        const size_t list_bytes = static_cast<size_t>(P) * halves * Bpad * m.k * sizeof(uint64_t);
        const size_t need = q_bytes + list_bytes;
        if (need > scratch.cap) {
            MMA_CUDA(cudaStreamSynchronize(m.stream));
            if (scratch.p) cudaFree(scratch.p);
            scratch.p = nullptr;
            scratch.cap = 0;
            MMA_CUDA(cudaMalloc(&scratch.p, need + need / 2));
            scratch.cap = need + need / 2;
        }
        if (Bpad > scratch.state_q) {
            MMA_CUDA(cudaStreamSynchronize(m.stream));
            if (scratch.state) cudaFree(scratch.state);
            scratch.state = nullptr;
            scratch.state_q = 0;
            const int cap_q = std::max(Bpad, 1024);
            MMA_CUDA(cudaMalloc(&scratch.state, mma_state_bytes(cap_q) + kDebugBytes + kProgBytes));  // (+ debug words, timeline, progress)
            MMA_CUDA(cudaMemsetAsync(scratch.state, 0, mma_state_bytes(cap_q) + kDebugBytes + kProgBytes, m.stream));
            scratch.state_q = cap_q;
        }
        __nv_bfloat16* qb = reinterpret_cast<__nv_bfloat16*>(scratch.p);
        uint64_t* lists = reinterpret_cast<uint64_t*>(scratch.p + q_bytes);
        uint32_t* heads = reinterpret_cast<uint32_t*>(scratch.state);                    // [state_q][kHeadsLd], 16-byte aligned rows
        uint32_t* gthr = heads + static_cast<size_t>(scratch.state_q) * kHeadsLd;         // [state_q]
        uint32_t* cnt = gthr + scratch.state_q;                                           // [state_q]
        uint32_t* owners_done = cnt + scratch.state_q;                                    // [state_q]

        CUtensorMap tq, te;
        memset(&tq, 0, sizeof(tq));
        if (!fused_prep) {
            k_prep_queries<<<(Bpad + 3) / 4, 128, 0, m.stream>>>(m.q + static_cast<size_t>(done) * m.D, Bs, Bpad, m.D, m.Dp,
                                                                 m.cosine ? 1 : 0, qb);
            MMA_CUDA(cudaGetLastError());
            if (launches) launches->fetch_add(1);
            if ((rc = encode_map(&tq, qb, m.Dp, Bpad, kQ))) return rc;
        }
        if ((rc = encode_map(&te, m.E, m.Dp, static_cast<uint64_t>(m.n), kTN))) return rc;

        KArgs a;
        a.scale = m.scale;
        a.ns = m.ns;
        a.label = m.label;
        a.mask = m.mask;
        a.ns_code = m.ns_code;
        a.label_code = m.label_code;
        a.n = m.n;
        a.kblocks = kblocks;
        a.stages = stages;
        a.q_resident = q_resident ? 1 : 0;
        a.k = m.k;
        a.row_offset = m.row_offset;
        a.row_stride = m.row_stride ? m.row_stride : 1u;
        static const bool no_prefilter = getenv("CODDVEC_K3_NO_PREFILTER") != nullptr;  // A/B switch
        a.scale_bounded = m.scale_bounded && !no_prefilter ? 1 : 0;
        a.n_tiles = n_tiles;
        a.q_tiles = QT;
        a.P = P;
        a.Bpad = Bpad;
        a.B = Bs;
        a.smem_lists = smem_list_bytes ? 1 : 0;
        a.lists = lists;
        a.gthr = gthr;
        a.heads = heads;
        a.dump = m.dump ? m.dump + static_cast<size_t>(done) * m.dump_ld : nullptr;
        a.dump_ld = m.dump_ld;
        a.corpus_hint = QT > 1 ? kEvictNormal : kEvictFirst;
        a.q_raw = fused_prep ? m.q + static_cast<size_t>(done) * m.D : nullptr;
        a.D = m.D;
        a.cosine = m.cosine ? 1 : 0;
        a.split = split ? 1 : 0;
        a.fused_tail = fused_tail ? 1 : 0;
        // lockstep of sibling super-tiles (CODDVEC_K3_LOCKSTEP=<window in tiles>, 0 = off: A/B switch)
        static const int lock_window = getenv("CODDVEC_K3_LOCKSTEP") ? atoi(getenv("CODDVEC_K3_LOCKSTEP")) : 4;
        // (two super-tiles stay together on their own -- 1.1x at 12.5M rows -- and short scans end before anybody has
        // drifted: 1.00x at 1M rows, 1.14x at 4M; the lockstep is for three or more super-tiles on long scans)
        static const int lock_min_tiles = getenv("CODDVEC_K3_LOCKSTEP_MIN_TILES") ? atoi(getenv("CODDVEC_K3_LOCKSTEP_MIN_TILES")) : 1000;
        a.prog = (QT > 2 && lock_window > 0 && n_tiles >= static_cast<int64_t>(lock_min_tiles) * P)
                     ? reinterpret_cast<unsigned long long*>(scratch.state + mma_state_bytes(scratch.state_q) + kDebugBytes)
                     : nullptr;
        a.epoch = ++scratch.epoch;
        a.lock_window = lock_window;
        a.cand = lists;
        a.cand_ld = P * halves * m.k;
        a.cnt = cnt;
        a.done = owners_done;
        a.out = m.out + static_cast<size_t>(done) * m.k;
        static const bool debug = getenv("CODDVEC_K3_DEBUG") != nullptr;
        a.dbg = debug ? reinterpret_cast<unsigned long long*>(scratch.state + mma_state_bytes(scratch.state_q)) : nullptr;
#ifdef CV_K3_TIMELINE
        static const bool timeline = getenv("CODDVEC_K3_TIMELINE") != nullptr;
#else
        static const bool timeline = false;  // (python -m codd_query_engine_b200.build --timeline builds the diagnostic library)
#endif
        a.tl = timeline ? reinterpret_cast<unsigned long long*>(scratch.state + mma_state_bytes(scratch.state_q) + 64) : nullptr;

        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(static_cast<unsigned>(QT * P * cg));
        cfg.blockDim = dim3(static_cast<unsigned>(mma_threads(ew)));
        cfg.dynamicSmemBytes = smem;
        cfg.stream = m.stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = static_cast<unsigned>(cg);
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        // after k_prep_queries: programmatic dependent launch -- this kernel's prologue and first corpus loads overlap
        // the query preparation (the kernel waits with griddepcontrol.wait before it touches the prepared tile)
        static const bool no_pdl = getenv("CODDVEC_K3_PDL") && atoi(getenv("CODDVEC_K3_PDL")) == 0;  // A/B switch
        if (!fused_prep && !no_pdl) {
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.numAttrs = 2;
        }
        const bool pilot = pilot_kernel != nullptr && P >= m.k && P > 1 && P <= 160 && (pilot_env == 1 || n_tiles <= 80 * P);
        MMA_CUDA(cudaLaunchKernelEx(&cfg, pilot ? pilot_kernel : kernel, tq, te, a));
        if (launches) launches->fetch_add(1);
        if (!fused_tail) {
            // k > 64 / score dump: per-CTA lists [P][Bpad][k] -> K4, then clear the thresholds the scan left behind
            MMA_CUDA(launch_merge_keys(lists, P, Bpad, m.k, m.k, m.out + static_cast<size_t>(done) * m.k, Bs, m.stream,
                                       P > 1 && !m.dump ? gthr : nullptr));
            MMA_CUDA(cudaMemsetAsync(heads, 0, static_cast<size_t>(Bpad) * kHeadsLd * sizeof(uint32_t), m.stream));
            MMA_CUDA(cudaMemsetAsync(gthr, 0, static_cast<size_t>(Bpad) * sizeof(uint32_t), m.stream));
            if (launches) launches->fetch_add(1);
        }
        done += Bs;
        if (a.tl) {  // CODDVEC_K3_TIMELINE: per-CTA phase stamps (synchronises; never set in production)
            const int ctas = QT * P * cg;
            std::vector<unsigned long long> w(static_cast<size_t>(ctas) * 16);
            MMA_CUDA(cudaStreamSynchronize(m.stream));
            MMA_CUDA(cudaMemcpy(w.data(), a.tl, w.size() * 8, cudaMemcpyDeviceToHost));
            unsigned long long t0 = ~0ull;
            for (int c = 0; c < ctas; ++c) t0 = std::min(t0, w[c * 16]);
            static const char* names[13] = {"entry", "prologue", "first block", "last mma", "first acc", "last scan", "tail", "exit",
                                            "mma: wait acc", "mma: wait tma", "epi: loop top", "epi: wait acc", "epi: scan"};
            fprintf(stderr, "[k3 timeline] B=%d cg=%d P=%d tiles/cta=%.1f (us after the first CTA's entry: min / median / max)\n", Bs, cg, P,
                    static_cast<double>(n_tiles) / P);
            for (int e = 0; e < 13; ++e) {
                std::vector<double> v;
                for (int c = 0; c < ctas; ++c)
                    if (w[c * 16 + e]) v.push_back((w[c * 16 + e] - (e < 8 ? t0 : 0ull)) * 1e-3);
                if (v.empty()) continue;
                std::sort(v.begin(), v.end());
                fprintf(stderr, "[k3 timeline]   %-14s %8.1f %8.1f %8.1f\n", names[e], v.front(), v[v.size() / 2], v.back());
            }
            static const bool trace = getenv("CODDVEC_K3_TRACE") != nullptr;
            if (trace) {  // CTA 0, tile by tile: MMA issuer (asks for acc, has acc, tile committed), epilogue (asks, has acc, done)
                std::vector<unsigned long long> tr(64 * 6);
                MMA_CUDA(cudaMemcpy(tr.data(), a.tl + 16 * 296, tr.size() * 8, cudaMemcpyDeviceToHost));
                for (int i = 0; i < 64 && tr[i * 6]; ++i)
                    fprintf(stderr, "[k3 trace] tile %2d  mma ask %7.2f got %7.2f committed %7.2f | epi ask %7.2f got %7.2f done %7.2f\n", i,
                            (tr[i * 6] - t0) * 1e-3, (tr[i * 6 + 1] - t0) * 1e-3, (tr[i * 6 + 2] - t0) * 1e-3, (tr[i * 6 + 3] - t0) * 1e-3,
                            (tr[i * 6 + 4] - t0) * 1e-3, (tr[i * 6 + 5] - t0) * 1e-3);
            }
            MMA_CUDA(cudaMemset(a.tl, 0, (296 * 16 + 64 * 6) * 8));
        }
        if (a.dbg) {  // CODDVEC_K3_DEBUG: in-kernel timers of the fused phases (synchronises; never set in production)
            unsigned long long w[4] = {0, 0, 0, 0};
            MMA_CUDA(cudaStreamSynchronize(m.stream));
            MMA_CUDA(cudaMemcpy(w, a.dbg, sizeof(w), cudaMemcpyDeviceToHost));
            MMA_CUDA(cudaMemset(a.dbg, 0, sizeof(w)));
            fprintf(stderr, "[k3 debug] B=%d cg=%d P=%d: prep %.1f us, publish %.1f us, merge %.1f us, max survivors/query %llu\n", Bs, cg, P,
                    w[0] * 1e-3, w[1] * 1e-3, w[2] * 1e-3, w[3]);
        }
    }
    return 0;
}

}  // namespace cv
