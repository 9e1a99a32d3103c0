"""Server-side micro-batching of concurrent single-query searches.

The reference issues one query per call (`collection.query(query_texts=[q], ...)`,
codd_dal/metrics/metrics_semantic_metadata_store.py:314-316) and the service handles requests one
at a time, so on its own traffic the tensor-core path (K3, batches >= 2) would never run.
``MicroBatcher`` sits between concurrent callers (uvicorn worker threads, a thread pool, ...) and one
``Collection``: requests that arrive within ``max_wait_ms`` of each other and agree on
(n_results, where, include) are answered by ONE ``Collection.query`` with B embeddings -- one K3
launch that reads the corpus once for the whole batch instead of once per request.

    batcher = MicroBatcher(collection, max_batch=256, max_wait_ms=0.5)
    hit = batcher.query(query_texts=["API latency"], n_results=5)      # same result shape as Collection.query
    batcher.close()
"""

from __future__ import annotations

import json
import threading
from concurrent.futures import Future
from typing import Any

import numpy as np


class MicroBatcher:
    def __init__(self, collection, max_batch: int = 256, max_wait_ms: float = 0.5):
        if max_batch < 1:
            raise ValueError("max_batch must be >= 1")
        self.collection = collection
        # the collection's own entry point below any CODDVEC_MICROBATCH routing (a batch of one must not come back here)
        self._search = getattr(collection, "_query_direct", collection.query)
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) / 1e3
        self._cv = threading.Condition()
        self._pending: list[tuple[Any, np.ndarray, dict, Future]] = []
        self._closed = False
        self.batches = 0          # Collection.query calls issued
        self.requests = 0         # single-query requests served
        self._worker = threading.Thread(target=self._run, name="coddvec-microbatch", daemon=True)
        self._worker.start()

    # ------------------------------------------------------------------ caller side
    def query(self, query_embeddings=None, query_texts=None, n_results: int = 10, where=None, include=None) -> dict:
        """One query in, one chromadb-shaped result out (blocks until its batch has been answered)."""
        if (query_embeddings is None) == (query_texts is None):
            raise ValueError("You must provide exactly one of query_embeddings or query_texts")
        if query_texts is not None:
            texts = [query_texts] if isinstance(query_texts, str) else list(query_texts)
            emb = self.collection._embed_texts(texts)  # embedding stays on the caller's thread
        else:
            emb = np.asarray(query_embeddings, dtype=np.float32)
            if emb.ndim == 1:
                emb = emb[None, :]
        if emb.shape[0] != 1:  # already a batch: nothing to coalesce
            return self._search(query_embeddings=emb, n_results=n_results, where=where, include=include)
        group = (int(n_results), json.dumps(where, sort_keys=True, default=str), None if include is None else tuple(include))
        fut: Future = Future()
        with self._cv:
            if self._closed:
                raise RuntimeError("MicroBatcher is closed")
            self._pending.append((group, emb[0], {"n_results": n_results, "where": where, "include": include}, fut))
            self._cv.notify()
        return fut.result()

    def close(self) -> None:
        with self._cv:
            self._closed = True
            self._cv.notify_all()
        self._worker.join()

    # ------------------------------------------------------------------ worker side
    def _take_batch(self):
        with self._cv:
            while not self._pending and not self._closed:
                self._cv.wait()
            if not self._pending:
                return None
            # give concurrent callers a moment to pile on, unless the batch is already full
            if len(self._pending) < self.max_batch and self.max_wait > 0:
                self._cv.wait(self.max_wait)
            group = self._pending[0][0]
            batch, rest = [], []
            for item in self._pending:
                (batch if item[0] == group and len(batch) < self.max_batch else rest).append(item)
            self._pending = rest
            return batch

    def _run(self) -> None:
        while True:
            batch = self._take_batch()
            if batch is None:
                return
            kw = batch[0][2]
            try:
                res = self._search(query_embeddings=np.stack([b[1] for b in batch]), **kw)
                self.batches += 1
                self.requests += len(batch)
                for i, (_, _, _, fut) in enumerate(batch):
                    one = {}
                    for key, val in res.items():
                        one[key] = [val[i]] if isinstance(val, list) and key != "included" else val
                    fut.set_result(one)
            except Exception as exc:  # every waiter sees the failure
                for _, _, _, fut in batch:
                    if not fut.done():
                        fut.set_exception(exc)
