#!/usr/bin/env python
"""Parity-checked measurement of the secondary BASELINE.json configs on one B200 (lives under tests/ because
it uses the oracle as its checker) (bench.py covers configs[1] and the 100M corpus):

  cfg1  the reference's CPU-runnable case: 10k x 384 in the semantic store, 1 query, top-10 -- end to end through
        Collection.query / search_metadata, beside the exact CPU oracle and the HNSW restatement (timing + recall@k)
  cfg2r (on request) 1M x 384 clustered rows: this package's exact search beside the HNSW restatement (recall + timing)
  cfg3  10M x 768 bf16, fused namespace predicate at 10 % selectivity, top-50, batch 1 (K2 PRED)
  cfg5  bulk upsert of 1M x 384 fp32 host embeddings in chunks of 1000 (MAX_BULK_OPERATIONS,
        codd_dal/metrics/metrics_semantic_metadata_store.py:23) interleaved with live top-10 queries

Each result is parity-checked against the oracle (checker only) and printed as one JSON line.
    python tests/bench_configs.py [cfg3] [cfg5] [--rows3 N] [--rows5 N]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402  (peaks, C oracle loader)
from codd_query_engine_b200 import native  # noqa: E402
from oracle import brute_force as bf  # noqa: E402  (checker only)


def arg(name, default):
    return int(sys.argv[sys.argv.index(name) + 1]) if name in sys.argv else default


def time_device(fn, iters, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cfg3(rows):
    D, k, sel_ns = 768, arg("--k3", 50), 7
    pk = bench.peaks()
    idx = native.NativeIndex(D, capacity_rows=rows, max_rows=rows)
    gen = torch.Generator(device="cuda").manual_seed(2)
    ns_all = []
    for lo in range(0, rows, 500_000):
        m = min(500_000, rows - lo)
        x = torch.randn((m, D), device="cuda", generator=gen)
        # 10 namespaces, the target one holds exactly 10 % of every chunk
        ns = (torch.arange(m, device="cuda") % 10 + 1).to(torch.int32)
        ns = ns[torch.randperm(m, device="cuda", generator=gen)].contiguous()
        idx.upsert_device(x.data_ptr(), m, ns_ptr=ns.data_ptr())
        ns_all.append(ns.cpu().numpy().astype(np.uint32))
    del x
    ns_all = np.concatenate(ns_all)
    q = torch.randn((1, D), device="cuda", generator=gen)
    out = torch.zeros((1, k), dtype=torch.int64, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ms_pred = time_device(lambda: idx.query_device(q.data_ptr(), 1, k, out.data_ptr(), ns_code=sel_ns, stream=s), 50)
    keys_pred = out.cpu().numpy().view(np.uint64).copy()
    ms_full = time_device(lambda: idx.query_device(q.data_ptr(), 1, k, out.data_ptr(), stream=s), 20)
    n_sel = int((ns_all == sel_ns).sum())
    st = idx.stats()
    sel_bytes = n_sel * (2 * st.dim_padded + 4) + rows * 4  # selected rows + the namespace column
    # parity: oracle over the eligible rows only (same stored embeddings)
    el = np.nonzero(ns_all == sel_ns)[0]
    stored, scale, _, _ = idx.get_rows(el)
    es, er, ec = bf.query(stored, scale, q.cpu().numpy(), k)
    gs, gr = native.decode_keys(keys_pred)
    parity = bool(np.array_equal(gr[0], el[er[0]]) and np.abs(gs - es).max() <= 2e-6)
    idx.close()
    return {"config": "cfg3", "workload": f"{rows} x 768 bf16, namespace predicate {n_sel / rows:.3f} selectivity, top-{k}, batch 1",
            "ms_per_query": ms_pred, "qps": 1e3 / ms_pred, "selected_rows": n_sel,
            "selected_bytes_gbs": sel_bytes / ms_pred / 1e6, "selected_bytes_frac_of_hbm": sel_bytes / ms_pred / 1e6 / pk["hbm_gbs"],
            "logical_scan_gbs": rows * (2 * st.dim_padded + 4) / ms_pred / 1e6,
            "unfiltered_ms_per_query": ms_full, "unfiltered_gbs": rows * (2 * st.dim_padded + 4) / ms_full / 1e6,
            "parity_vs_oracle": parity}


def cfg5(rows):
    D, k, chunk = 384, 10, 1000
    rng = np.random.default_rng(5)
    x = rng.standard_normal((rows, D), dtype=np.float32)
    q = rng.standard_normal((1, D), dtype=np.float32)
    idx = native.NativeIndex(D, capacity_rows=rows, max_rows=rows)
    idx.upsert(x[:chunk])  # warm the path
    lat, t_up = [], 0.0
    t0 = time.perf_counter()
    for lo in range(chunk, rows, chunk):
        a = time.perf_counter()
        idx.upsert(x[lo:lo + chunk])
        b = time.perf_counter()
        t_up += b - a
        if (lo // chunk) % 10 == 0:  # a live top-10 query every 10 chunks, over the growing corpus
            idx.query(q, k)
            lat.append(time.perf_counter() - b)
    wall = time.perf_counter() - t0
    # one bulk call for comparison (what a batched indexer would issue)
    idx2 = native.NativeIndex(D, capacity_rows=rows, max_rows=rows)
    a = time.perf_counter()
    idx2.upsert(x)
    bulk = time.perf_counter() - a
    # the same from page-locked host memory (cudaMemcpyAsync then runs at PCIe speed instead of through
    # the driver's pageable bounce buffers)
    xp = torch.empty((rows, D), dtype=torch.float32, pin_memory=True)
    xp.numpy()[...] = x
    idx3 = native.NativeIndex(D, capacity_rows=rows, max_rows=rows)
    a = time.perf_counter()
    idx3.upsert(xp.numpy())
    bulk_pinned = time.perf_counter() - a
    idx4 = native.NativeIndex(D, capacity_rows=rows, max_rows=rows)
    a = time.perf_counter()
    for lo in range(0, rows, chunk):
        idx4.upsert(xp.numpy()[lo:lo + chunk])
    chunked_pinned = time.perf_counter() - a
    idx3.close()
    idx4.close()
    # parity of the final state: stored bits == oracle K1, final query == oracle
    s, r, c = idx.query(q, k)
    sample = np.arange(0, rows, max(1, rows // 100_000))
    stored, scale, _, _ = idx.get_rows(sample)
    est, esc = bf.normalise_store(x[sample])
    k1_exact = bool(np.array_equal(stored, est) and np.array_equal(scale, esc))
    lib = bench.load_c_oracle()
    allst, allsc, _, _ = idx.get_rows(np.arange(rows))
    _, _, okeys = bench.cpu_time_queries(lib, allst, allsc, q, k, 1)
    os_, or_ = bf.unpack_keys(okeys)
    parity = bool(np.array_equal(or_, r) and np.abs(os_ - s).max() <= 2e-6)
    idx.close()
    idx2.close()
    lat = np.array(lat) * 1e3
    return {"config": "cfg5", "workload": f"{rows} x 384 fp32 host embeddings upserted in chunks of {chunk}, top-10 query every 10 chunks",
            "rows_per_s_chunked": (rows - chunk) / t_up, "wall_s_interleaved": wall, "queries": int(lat.size),
            "query_ms_p50": float(np.percentile(lat, 50)), "query_ms_p99": float(np.percentile(lat, 99)),
            "rows_per_s_single_bulk_call": rows / bulk, "h2d_gbs_single_bulk_call": rows * D * 4 / bulk / 1e9,
            "rows_per_s_single_bulk_call_pinned": rows / bulk_pinned, "h2d_gbs_single_bulk_call_pinned": rows * D * 4 / bulk_pinned / 1e9,
            "rows_per_s_chunked_pinned": rows / chunked_pinned,
            "k1_bit_exact": k1_exact, "final_query_parity_vs_oracle": parity}


def cfg1(rows=10_000):
    """BASELINE configs[0]: the repo's semantic store over 10k synthetic 384-d metric embeddings, one
    query, top-10 -- through the chromadb-compatible Collection and the store facade (end to end, host
    in / host out), beside the CPU oracle on the same vectors."""
    from codd_query_engine_b200.client import Client
    from codd_query_engine_b200.semantic_store import MetricsSemanticMetadataStore

    D, k = 384, 10
    rng = np.random.default_rng(0)
    x = rng.standard_normal((rows, D), dtype=np.float32)
    client = Client(endpoint="cfg1")
    warm = client.get_or_create_collection("warm")  # CUDA context + library initialisation are not indexing time
    warm.upsert(ids=["w"], embeddings=x[:1])
    store = MetricsSemanticMetadataStore(client, collection_name="cfg1")
    coll = store.collection
    t0 = time.perf_counter()
    for lo in range(0, rows, 1000):
        coll.upsert(ids=[f"ns#metric.{i}" for i in range(lo, lo + 1000)], embeddings=x[lo:lo + 1000],
                    metadatas=[{"namespace": "ns", "description": f"synthetic metric {i}"} for i in range(lo, lo + 1000)],
                    documents=[f"synthetic metric {i}" for i in range(lo, lo + 1000)])
    t_index = time.perf_counter() - t0
    q = rng.standard_normal((1, D), dtype=np.float32)
    for _ in range(20):
        coll.query(query_embeddings=q, n_results=k)
    lat = []
    for _ in range(500):
        a = time.perf_counter()
        res = coll.query(query_embeddings=q, n_results=k)
        lat.append(time.perf_counter() - a)
    lat_text = []
    for i in range(200):
        a = time.perf_counter()
        hits = store.search_metadata(f"synthetic metric {i * 37}", n_results=k)  # text -> hashed embedder -> K2
        lat_text.append(time.perf_counter() - a)
    # CPU oracle on the same stored vectors (checker + baseline)
    idx = coll._index
    stored, scale, _, _ = idx.get_rows(np.arange(rows))
    lib = bench.load_c_oracle()
    bench.cpu_time_queries(lib, stored, scale, q, k, 5)
    t_cpu, threads, okeys = bench.cpu_time_queries(lib, stored, scale, q, k, 200)
    os_, or_ = bf.unpack_keys(okeys)
    parity = [f"ns#metric.{r}" for r in or_[0]] == res["ids"][0] and bool(np.allclose(1 - os_[0], res["distances"][0], atol=2e-6))
    shape_ok = len(hits) == k and all("similarity_score" in h and h["namespace"] == "ns" for h in hits)
    # The reference's own index type on the same vectors: HNSW (cosine, M=16, ef 200/100 -- store.py:63-68),
    # restated in oracle/hnsw.c.  Reported: its single-thread query time (one query per call, as the reference
    # issues them) and recall@k of its answer / of this package's answer against the exact brute force.
    from oracle import hnsw  # checker / baseline only
    nq = 200
    qs = rng.standard_normal((nq, D), dtype=np.float32)
    t0 = time.perf_counter()
    graph = hnsw.HnswIndex(D, rows)
    graph.add(x)
    t_graph = time.perf_counter() - t0
    graph.search(qs[:20], k)
    t0 = time.perf_counter()
    h_ids, _ = graph.search(qs, k, threads=1)
    t_hnsw = (time.perf_counter() - t0) / nq
    ex_keys = bench.cpu_time_queries(lib, stored, scale, qs, k, 1)[2]
    _, ex_rows = bf.unpack_keys(ex_keys)
    # one query per call, as the reference issues them (-> the fp32-exact K2 scan); the same batch in ONE call takes
    # the tensor path, whose bf16-rounded queries may swap near-ties inside the 1e-3 tolerance
    our_rows = np.array([[int(i.rsplit(".", 1)[1]) for i in coll.query(query_embeddings=qs[i:i + 1], n_results=k, include=[])["ids"][0]]
                         for i in range(nq)])
    ours_b = coll.query(query_embeddings=qs, n_results=k, include=[])
    our_rows_batched = np.array([[int(i.rsplit(".", 1)[1]) for i in row] for row in ours_b["ids"]])
    recall_hnsw = hnsw.recall_at_k(h_ids, ex_rows)
    recall_ours = hnsw.recall_at_k(our_rows, ex_rows)
    recall_ours_batched = hnsw.recall_at_k(our_rows_batched, ex_rows)
    graph.close()
    # clustered variant (256 Gaussian centroids, sigma 0.3 -- SURVEY.md section 8d): embeddings with structure, where
    # the graph walk does well; iid Gaussian rows above are its worst case
    cent = rng.standard_normal((256, D), dtype=np.float32)
    xc = cent[rng.integers(0, 256, rows)] + 0.3 * rng.standard_normal((rows, D), dtype=np.float32)
    qc = cent[rng.integers(0, 256, nq)] + 0.3 * rng.standard_normal((nq, D), dtype=np.float32)
    graph = hnsw.HnswIndex(D, rows)
    graph.add(xc)
    hc_ids, _ = graph.search(qc, k, threads=1)
    graph.close()
    cst, csc = bf.normalise_store(xc)
    _, exc_rows = bf.unpack_keys(bench.cpu_time_queries(lib, cst, csc, qc, k, 1)[2])
    cl = client.get_or_create_collection("cfg1_clustered")
    cl.upsert(ids=[f"c.{i}" for i in range(rows)], embeddings=xc)
    oc_rows = np.array([[int(i.rsplit(".", 1)[1]) for i in cl.query(query_embeddings=qc[i:i + 1], n_results=k, include=[])["ids"][0]]
                        for i in range(nq)])
    recall_hnsw_c, recall_ours_c = hnsw.recall_at_k(hc_ids, exc_rows), hnsw.recall_at_k(oc_rows, exc_rows)
    client.reset()
    lat, lat_text = np.array(lat) * 1e3, np.array(lat_text) * 1e3
    return {"config": "cfg1", "workload": f"{rows} x 384 metric embeddings in the semantic store, 1 query, top-{k}",
            "collection_query_ms_p50": float(np.percentile(lat, 50)), "collection_query_ms_p99": float(np.percentile(lat, 99)),
            "collection_qps": float(1e3 / np.mean(lat)),
            "store_search_metadata_ms_p50": float(np.percentile(lat_text, 50)),
            "index_rows_per_s": rows / t_index,
            "cpu_oracle_ms_per_query": float(np.median(t_cpu) * 1e3), "cpu_oracle_qps": float(1 / np.median(t_cpu)), "cpu_threads": threads,
            "hnsw_port_ms_per_query": t_hnsw * 1e3, "hnsw_port_qps": 1 / t_hnsw, "hnsw_port_threads": 1,
            "hnsw_port_build_rows_per_s": rows / t_graph, "hnsw_params": "cosine M=16 ef_construction=200 ef_search=100",
            "recall_at_k_hnsw_vs_exact": recall_hnsw, "recall_at_k_ours_vs_exact": recall_ours, "recall_at_k_ours_batched_tensor_path_vs_exact": recall_ours_batched,
            "recall_queries": nq,
            "recall_at_k_hnsw_vs_exact_clustered": recall_hnsw_c, "recall_at_k_ours_vs_exact_clustered": recall_ours_c,
            "parity_vs_oracle": bool(parity), "store_result_shape_ok": bool(shape_ok)}


def cfg2r(rows):
    """BASELINE configs[1]'s corpus size with structure (256 Gaussian centroids, sigma 0.3): this package's exact
    one-query-per-call search beside the reference's index type -- the HNSW restatement with the reference's
    parameters, graph built with all host threads (add_items(num_threads) analogue), queried on one thread --
    and the recall@10 of both against the exact CPU oracle.  (Only on request: the graph build takes ~a minute.)"""
    from oracle import hnsw  # checker / baseline only

    D, k, nq = 384, 10, 200
    rng = np.random.default_rng(0)
    cent = rng.standard_normal((256, D), dtype=np.float32)
    x = cent[rng.integers(0, 256, rows)] + 0.3 * rng.standard_normal((rows, D), dtype=np.float32)
    qs = cent[rng.integers(0, 256, nq)] + 0.3 * rng.standard_normal((nq, D), dtype=np.float32)
    idx = native.NativeIndex(D, capacity_rows=rows, max_rows=rows)
    idx.upsert(x)
    for i in range(20):
        idx.query(qs[i:i + 1], k)
    t0 = time.perf_counter()
    ours = np.concatenate([idx.query(qs[i:i + 1], k)[1] for i in range(nq)])
    t_ours = (time.perf_counter() - t0) / nq
    stored, scale, _, _ = idx.get_rows(np.arange(rows))
    idx.close()
    lib = bench.load_c_oracle()
    t_bf, threads, _ = bench.cpu_time_queries(lib, stored, scale, qs[:1], k, 20)
    _, exact = bf.unpack_keys(bench.cpu_time_queries(lib, stored, scale, qs, k, 1)[2])
    t0 = time.perf_counter()
    graph = hnsw.HnswIndex(D, rows)
    graph.add(x, threads=os.cpu_count() or 1)
    t_graph = time.perf_counter() - t0
    graph.search(qs[:20], k)
    t0 = time.perf_counter()
    approx = np.concatenate([graph.search(qs[i:i + 1], k)[0] for i in range(nq)])
    t_hnsw = (time.perf_counter() - t0) / nq
    graph.close()
    return {"config": "cfg2r", "workload": f"{rows} x 384 clustered rows, 1 query per call, top-{k}",
            "ours_e2e_ms_per_query": t_ours * 1e3, "ours_e2e_qps": 1 / t_ours,
            "recall_at_k_ours_vs_exact": hnsw.recall_at_k(ours, exact),
            "hnsw_port_ms_per_query_1_thread": t_hnsw * 1e3, "hnsw_port_qps_1_thread": 1 / t_hnsw,
            "hnsw_port_build_s": t_graph, "hnsw_port_build_threads": os.cpu_count() or 1,
            "recall_at_k_hnsw_vs_exact": hnsw.recall_at_k(approx, exact),
            "cpu_exact_oracle_ms_per_query": float(np.median(t_bf)) * 1e3, "cpu_exact_oracle_threads": threads,
            "hnsw_params": "cosine M=16 ef_construction=200 ef_search=100"}


def main():
    want = [a for a in sys.argv[1:] if a.startswith("cfg")] or ["cfg1", "cfg3", "cfg5"]
    if "cfg2r" in want:
        print(json.dumps(cfg2r(arg("--rows2", 1_000_000))), flush=True)
    if "cfg1" in want:
        print(json.dumps(cfg1()), flush=True)
    if "cfg3" in want:
        print(json.dumps(cfg3(arg("--rows3", 10_000_000))), flush=True)
    if "cfg5" in want:
        print(json.dumps(cfg5(arg("--rows5", 1_000_000))), flush=True)


if __name__ == "__main__":
    main()
