"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper of oracle/hnsw.c, the restatement of the approximate
HNSW index the reference configures (codd_dal/metrics/metrics_semantic_metadata_store.py:60-69:
cosine, M=16, construction_ef=200, search_ef=100) and queries (:314-316) through the un-vendored
``chromadb`` / ``hnswlib`` dependency.

PARITY UNPINNED: follows the published algorithm (Malkov & Yashunin 2018) with hnswlib's documented
conventions; see the header of oracle/hnsw.c.  Used only for the recall@k report and the CPU timing of
the reference's query algorithm (tests/, bench.py); never on the product path.
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

# the reference's collection metadata (store.py:63-68)
REF_M = 16
REF_EF_CONSTRUCTION = 200
REF_EF_SEARCH = 100


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle_hnsw.so")
        if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(os.path.join(_HERE, "hnsw.c")):
            env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
            subprocess.run(["make", "-C", _HERE, "liboracle_hnsw.so"], check=True, capture_output=True, env=env)
        lib = ctypes.CDLL(path)
        lib.hnsw_create.restype = ctypes.c_void_p
        lib.hnsw_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_uint64]
        lib.hnsw_destroy.argtypes = [ctypes.c_void_p]
        lib.hnsw_destroy.restype = None
        lib.hnsw_count.argtypes = [ctypes.c_void_p]
        lib.hnsw_count.restype = ctypes.c_int64
        lib.hnsw_max_level.argtypes = [ctypes.c_void_p]
        lib.hnsw_max_level.restype = ctypes.c_int
        lib.hnsw_dist_evals.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.hnsw_dist_evals.restype = ctypes.c_int64
        lib.hnsw_add.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        lib.hnsw_add.restype = ctypes.c_int
        lib.hnsw_add_parallel.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int]
        lib.hnsw_add_parallel.restype = ctypes.c_int
        lib.hnsw_search.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.hnsw_search.restype = ctypes.c_int
        _LIB = lib
    return _LIB


class HnswIndex:
    """Cosine (or inner-product) HNSW graph with the reference's parameters by default."""

    def __init__(self, dim: int, capacity: int, M: int = REF_M, ef_construction: int = REF_EF_CONSTRUCTION,
                 cosine: bool = True, seed: int = 100):
        self.dim = int(dim)
        self._h = _lib().hnsw_create(self.dim, int(M), int(ef_construction), int(capacity), int(bool(cosine)), int(seed))
        if not self._h:
            raise MemoryError("hnsw_create failed")

    def add(self, x: np.ndarray, threads: int = 1) -> None:
        """threads <= 1: one insertion after another (deterministic); > 1: concurrent insertion of the batch, as
        the library's add_items(num_threads) does (graph depends on the interleaving)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2 or x.shape[1] != self.dim:
            raise ValueError("embeddings must be [n][dim]")
        rc = (_lib().hnsw_add(self._h, x.ctypes.data, x.shape[0]) if threads <= 1 else
              _lib().hnsw_add_parallel(self._h, x.ctypes.data, x.shape[0], int(threads)))
        if rc != 0:
            raise RuntimeError("hnsw_add: capacity exceeded")

    def search(self, q: np.ndarray, k: int, ef_search: int = REF_EF_SEARCH, threads: int = 1):
        """-> (ids int64 [B][k] (-1 padded), distances fp32 [B][k] ascending = 1 - cosine)."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        ids = np.empty((q.shape[0], k), dtype=np.int64)
        dist = np.empty((q.shape[0], k), dtype=np.float32)
        _lib().hnsw_search(self._h, q.ctypes.data, q.shape[0], int(k), int(ef_search), int(threads), ids.ctypes.data,
                           dist.ctypes.data)
        return ids, dist

    def __len__(self) -> int:
        return int(_lib().hnsw_count(self._h))

    @property
    def max_level(self) -> int:
        return int(_lib().hnsw_max_level(self._h))

    def dist_evals(self, reset: bool = False) -> int:
        return int(_lib().hnsw_dist_evals(self._h, int(reset)))

    def close(self) -> None:
        if self._h:
            _lib().hnsw_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def recall_at_k(approx_ids: np.ndarray, exact_ids: np.ndarray) -> float:
    """Mean |approx ∩ exact| / |exact| over queries (exact rows < 0 are padding)."""
    hits = total = 0
    for a, e in zip(approx_ids, exact_ids):
        e = set(int(v) for v in e if v >= 0)
        hits += len(e & set(int(v) for v in a if v >= 0))
        total += len(e)
    return hits / total if total else 1.0
