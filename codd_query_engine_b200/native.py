"""ctypes binding of libcoddvec.so (include/coddvec.h).

This is the only way the package computes anything: there is no CPU fallback.  If the shared
library is missing or no B200 is visible, every compute call raises ``NativeError``.

The boundary mirrors what the reference's chromadb collection object does for
codd_dal/metrics/metrics_semantic_metadata_store.py:236-238 (upsert), :260 (get) and :314-316
(query); id strings, documents and metadata stay in Python (collection.py).
"""

from __future__ import annotations

import ctypes
import os
from pathlib import Path

import numpy as np

CV_OK = 0
CV_ERR_INVALID = -1
CV_ERR_CUDA = -2
CV_ERR_OOM = -3
CV_ERR_NONFINITE = -4
CV_ERR_UNSUPPORTED = -5

METRIC_COSINE = 0
METRIC_IP = 1
ANY = 0xFFFFFFFF
MAX_K = 128
MAX_DIM = 4096

PATH_AUTO = 0
PATH_SCAN = 1
PATH_MMA = 2
PATH_MMA_1CTA = 3
PATH_MMA_2CTA = 4

PRECISION_AUTO = 0   # bf16 queries; hi/lo split on rows of up to 128 padded elements
PRECISION_BF16 = 1
PRECISION_SPLIT = 2  # hi/lo split (fp32-accurate scores) wherever the double query tile fits (dim <= 384)

LIB_NAME = "libcoddvec.so"


class NativeError(RuntimeError):
    """libcoddvec.so reported a failure (or is missing)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"coddvec error {code}: {message}")
        self.code = code


class _Config(ctypes.Structure):
    _fields_ = [
        ("dim", ctypes.c_int32),
        ("metric", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("flags", ctypes.c_int32),
        ("capacity_rows", ctypes.c_int64),
        ("max_rows", ctypes.c_int64),
        ("row_offset", ctypes.c_int64),
        ("row_stride", ctypes.c_int64),
    ]


class _QueryParams(ctypes.Structure):
    _fields_ = [
        ("n_queries", ctypes.c_int32),
        ("k", ctypes.c_int32),
        ("ns_code", ctypes.c_uint32),
        ("label_code", ctypes.c_uint32),
        ("mask", ctypes.c_void_p),
        ("path", ctypes.c_int32),
        ("precision", ctypes.c_int32),
    ]


class _PeerExchange(ctypes.Structure):
    _fields_ = [
        ("world", ctypes.c_int32),
        ("rank", ctypes.c_int32),
        ("max_queries", ctypes.c_int32),
        ("max_k", ctypes.c_int32),
        ("buffers", ctypes.c_uint64 * 8),
    ]


class Stats(ctypes.Structure):
    _fields_ = [
        ("rows", ctypes.c_int64),
        ("capacity_rows", ctypes.c_int64),
        ("bytes_per_row", ctypes.c_int64),
        ("queries", ctypes.c_int64),
        ("kernel_launches", ctypes.c_int64),
        ("upserted_rows", ctypes.c_int64),
        ("last_query_ms", ctypes.c_double),
        ("last_query_bytes", ctypes.c_int64),
        ("last_path", ctypes.c_int32),
        ("dim_padded", ctypes.c_int32),
    ]


_lib = None


def library_path() -> Path:
    override = os.environ.get("CODDVEC_LIB")
    return Path(override) if override else Path(__file__).resolve().parent / LIB_NAME


def load_library() -> ctypes.CDLL:
    """dlopen libcoddvec.so and declare every prototype of include/coddvec.h."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise NativeError(
            CV_ERR_CUDA,
            f"{path} not found: build it with `python -m codd_query_engine_b200.build` "
            "(there is no CPU fallback)",
        )
    lib = ctypes.CDLL(str(path))
    vp, i32, i64, u32, u64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint32, ctypes.c_uint64
    protos = {
        "cv_create": (ctypes.c_int, [ctypes.POINTER(_Config), ctypes.POINTER(vp)]),
        "cv_destroy": (None, [vp]),
        "cv_upsert": (ctypes.c_int, [vp, vp, vp, vp, vp, i64, vp]),
        "cv_upsert_device": (ctypes.c_int, [vp, vp, vp, vp, vp, i64, vp]),
        "cv_query": (ctypes.c_int, [vp, vp, ctypes.POINTER(_QueryParams), vp, vp, vp]),
        "cv_query_device": (ctypes.c_int, [vp, vp, ctypes.POINTER(_QueryParams), vp, vp]),
        "cv_peer_exchange_bytes": (i64, [i32, i32, i32]),
        "cv_query_device_exchange": (ctypes.c_int, [vp, vp, ctypes.POINTER(_QueryParams), ctypes.POINTER(_PeerExchange),
                                                    u64, vp, vp]),
        "cv_query_exchange": (ctypes.c_int, [vp, vp, ctypes.POINTER(_QueryParams), ctypes.POINTER(_PeerExchange),
                                             u64, vp, vp, vp]),
        "cv_group_create": (ctypes.c_int, [ctypes.POINTER(vp), i32, i32, i32, ctypes.POINTER(vp)]),
        "cv_group_destroy": (None, [vp]),
        "cv_group_query": (ctypes.c_int, [vp, vp, ctypes.POINTER(_QueryParams), ctypes.POINTER(vp), vp, vp, vp]),
        "cv_merge_keys_device": (ctypes.c_int, [i32, vp, i32, i32, i32, i32, vp, vp]),
        "cv_decode_keys": (None, [vp, i64, vp, vp]),
        "cv_encode_key": (u64, [ctypes.c_float, u32]),
        "cv_get_rows": (ctypes.c_int, [vp, vp, i64, vp, vp, vp, vp]),
        "cv_load_rows": (ctypes.c_int, [vp, vp, vp, vp, vp, i64, vp]),
        "cv_delete_rows": (ctypes.c_int, [vp, vp, i64]),
        "cv_set_codes": (ctypes.c_int, [vp, vp, vp, vp, i64]),
        "cv_debug_mma_scores": (ctypes.c_int, [vp, vp, i32, i32, i32, vp]),
        "cv_device_count": (i32, []),
        "cv_count": (i64, [vp]),
        "cv_dim": (i32, [vp]),
        "cv_set_profiling": (ctypes.c_int, [vp, i32]),
        "cv_stats": (ctypes.c_int, [vp, ctypes.POINTER(Stats)]),
        "cv_synchronize": (ctypes.c_int, [vp]),
        "cv_last_error": (ctypes.c_char_p, []),
        "cv_version": (ctypes.c_char_p, []),
    }
    for name, (res, args) in protos.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            # only for same-box A/B runs against a library built from an OLDER commit (tests/ab_build.sh + CODDVEC_LIB)
            if os.environ.get("CODDVEC_LIB") and os.environ.get("CODDVEC_LIB_ALLOW_MISSING") == "1":
                continue
            raise
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = (
    "cv_create", "cv_destroy", "cv_upsert", "cv_upsert_device", "cv_query", "cv_query_device",
    "cv_peer_exchange_bytes", "cv_query_device_exchange", "cv_query_exchange", "cv_group_create", "cv_group_destroy",
    "cv_group_query", "cv_merge_keys_device", "cv_decode_keys",
    "cv_encode_key", "cv_get_rows", "cv_load_rows", "cv_delete_rows", "cv_set_codes", "cv_debug_mma_scores", "cv_device_count", "cv_count", "cv_dim",
    "cv_set_profiling", "cv_stats", "cv_synchronize", "cv_last_error", "cv_version",
)


def _check(rc: int) -> None:
    if rc != CV_OK:
        msg = load_library().cv_last_error().decode("utf-8", "replace")
        if rc == CV_ERR_INVALID:
            raise ValueError(f"coddvec: {msg}")
        if rc == CV_ERR_NONFINITE:
            raise ValueError(f"coddvec: {msg}")
        if rc == CV_ERR_OOM:
            raise MemoryError(f"coddvec: {msg}")
        raise NativeError(rc, msg)


def _ptr(a) -> ctypes.c_void_p:
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def decode_keys(keys: np.ndarray):
    """u64 result keys -> (scores fp32, global rows int64, -1 = empty)."""
    keys = np.ascontiguousarray(keys, dtype=np.uint64)
    scores = np.empty(keys.shape, dtype=np.float32)
    rows = np.empty(keys.shape, dtype=np.int64)
    load_library().cv_decode_keys(_ptr(keys), keys.size, _ptr(scores), _ptr(rows))
    return scores, rows


def encode_key(score: float, row: int) -> int:
    return int(load_library().cv_encode_key(float(score), int(row)))


def device_count() -> int:
    """CUDA devices this process can see (0 without a driver / GPU)."""
    return int(load_library().cv_device_count())


def peer_exchange_bytes(world: int, max_queries: int, max_k: int) -> int:
    n = int(load_library().cv_peer_exchange_bytes(world, max_queries, max_k))
    if n < 0:
        raise ValueError("bad peer exchange shape")
    return n


def make_peer_exchange(world: int, rank: int, max_queries: int, max_k: int, buffer_ptrs) -> _PeerExchange:
    px = _PeerExchange(int(world), int(rank), int(max_queries), int(max_k))
    for g, ptr in enumerate(buffer_ptrs):
        px.buffers[g] = int(ptr)
    return px


CUDA_STREAM_LEGACY = 1  # cudaStreamLegacy: the handle that names the default (NULL) stream explicitly


def _stream_arg(stream):
    """None -> the library's own stream; 0 (torch's default-stream handle) -> cudaStreamLegacy."""
    if stream is None:
        return None
    return ctypes.c_void_p(CUDA_STREAM_LEGACY if int(stream) == 0 else int(stream))


def merge_keys_device(device: int, keys_in_ptr: int, G: int, B: int, k_in: int, k_out: int, out_ptr: int,
                      stream=0) -> None:
    """K4 on device pointers (e.g. torch tensors' data_ptr()) -- used after the NCCL all-gather.

    stream: a cudaStream_t handle (torch.cuda.current_stream().cuda_stream); 0/None = default stream."""
    _check(load_library().cv_merge_keys_device(device, keys_in_ptr, G, B, k_in, k_out, out_ptr,
                                               _stream_arg(0 if stream is None else stream)))


class NativeIndex:
    """One device-resident shard: bf16 rows + scale + namespace/label codes on one B200."""

    def __init__(self, dim: int, metric: int = METRIC_COSINE, device: int = 0, capacity_rows: int = 0,
                 max_rows: int = 0, row_offset: int = 0, row_stride: int = 1):
        self._lib = load_library()
        self._h = ctypes.c_void_p()
        cfg = _Config(int(dim), int(metric), int(device), 0, int(capacity_rows), int(max_rows), int(row_offset),
                      int(row_stride))
        _check(self._lib.cv_create(ctypes.byref(cfg), ctypes.byref(self._h)))
        self.dim = int(dim)
        self.metric = int(metric)
        self.device = int(device)
        self.row_offset = int(row_offset)
        self.row_stride = int(row_stride) or 1

    # -- lifecycle
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cv_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return int(self._lib.cv_count(self._h))

    count = __len__

    # -- K1
    def upsert(self, embeddings: np.ndarray, rows=None, ns_codes=None, label_codes=None) -> np.ndarray:
        emb = np.ascontiguousarray(embeddings, dtype=np.float32)
        if emb.ndim != 2 or emb.shape[1] != self.dim:
            raise ValueError(f"embeddings must be [n][{self.dim}], got {emb.shape}")
        n = emb.shape[0]
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        ns = None if ns_codes is None else np.ascontiguousarray(ns_codes, dtype=np.uint32)
        lb = None if label_codes is None else np.ascontiguousarray(label_codes, dtype=np.uint32)
        for name, arr in (("rows", r), ("ns_codes", ns), ("label_codes", lb)):
            if arr is not None and arr.shape != (n,):
                raise ValueError(f"{name} must have shape ({n},)")
        out = np.empty(n, dtype=np.int64)
        _check(self._lib.cv_upsert(self._h, _ptr(r), _ptr(emb), _ptr(ns), _ptr(lb), n, _ptr(out)))
        return out

    def upsert_device(self, emb_ptr: int, n: int, rows=None, ns_ptr: int = 0, label_ptr: int = 0) -> np.ndarray:
        """Embeddings already on this device ([n][dim] fp32, e.g. a torch tensor's data_ptr())."""
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        out = np.empty(n, dtype=np.int64)
        _check(self._lib.cv_upsert_device(self._h, _ptr(r), emb_ptr, ns_ptr or None, label_ptr or None, n, _ptr(out)))
        return out

    def load_rows(self, stored_bf16: np.ndarray, scale: np.ndarray, ns_codes=None, label_codes=None) -> np.ndarray:
        """Append rows exactly as get_rows returned them (snapshot restore; no re-normalisation)."""
        vec = np.ascontiguousarray(stored_bf16, dtype=np.uint16)
        if vec.ndim != 2 or vec.shape[1] != self.dim:
            raise ValueError(f"stored rows must be [n][{self.dim}] uint16, got {vec.shape}")
        n = vec.shape[0]
        sc = np.ascontiguousarray(scale, dtype=np.float32)
        ns = None if ns_codes is None else np.ascontiguousarray(ns_codes, dtype=np.uint32)
        lb = None if label_codes is None else np.ascontiguousarray(label_codes, dtype=np.uint32)
        for name, arr in (("scale", sc), ("ns_codes", ns), ("label_codes", lb)):
            if arr is not None and arr.shape != (n,):
                raise ValueError(f"{name} must have shape ({n},)")
        out = np.empty(n, dtype=np.int64)
        _check(self._lib.cv_load_rows(self._h, _ptr(vec), _ptr(sc), _ptr(ns), _ptr(lb), n, _ptr(out)))
        return out

    def delete_rows(self, rows) -> None:
        """Mark local rows dead: no query returns them until an upsert reuses the slot (no per-query mask needed)."""
        r = np.ascontiguousarray(rows, dtype=np.int64)
        _check(self._lib.cv_delete_rows(self._h, _ptr(r), r.shape[0]))

    def set_codes(self, rows, ns_codes=None, label_codes=None) -> None:
        """Rewrite the namespace / label predicate codes of local rows in place (metadata-only updates)."""
        r = np.ascontiguousarray(rows, dtype=np.int64)
        ns = None if ns_codes is None else np.ascontiguousarray(ns_codes, dtype=np.uint32)
        lb = None if label_codes is None else np.ascontiguousarray(label_codes, dtype=np.uint32)
        _check(self._lib.cv_set_codes(self._h, _ptr(r), _ptr(ns), _ptr(lb), r.shape[0]))

    # -- K2 / K3 / K4
    def query(self, queries: np.ndarray, k: int, ns_code: int = ANY, label_code: int = ANY, mask=None,
              path: int = PATH_AUTO, precision: int = PRECISION_AUTO, mask_bits=None):
        """Host API.  Returns (scores [B][k] fp32 desc, rows [B][k] int64 global (-1 pad), counts [B])."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [B][{self.dim}], got {q.shape}")
        B = q.shape[0]
        m = None
        if mask is None and mask_bits is not None:  # the bool array a mask would be packed from (see where.pack_mask)
            bits = np.asarray(mask_bits, dtype=bool)
            pad = (-bits.shape[0]) % 32
            mask = np.packbits(np.concatenate([bits, np.zeros(pad, dtype=bool)]) if pad else bits, bitorder="little").view(np.uint32)
        if mask is not None:
            m = np.ascontiguousarray(mask, dtype=np.uint32)
            need = (len(self) + 31) // 32
            if m.ndim != 1 or m.shape[0] < need:
                raise ValueError(f"mask needs {need} uint32 words, got shape {m.shape}")
        p = _QueryParams(B, int(k), int(ns_code), int(label_code), None if m is None else m.ctypes.data, int(path),
                         int(precision))
        scores = np.empty((B, int(k) if k > 0 else 0), dtype=np.float32)
        rows = np.empty((B, int(k) if k > 0 else 0), dtype=np.int64)
        counts = np.empty(B, dtype=np.int32)
        _check(self._lib.cv_query(self._h, _ptr(q), ctypes.byref(p), _ptr(scores), _ptr(rows), _ptr(counts)))
        return scores, rows, counts

    def query_device(self, q_ptr: int, B: int, k: int, out_keys_ptr: int, ns_code: int = ANY, label_code: int = ANY,
                     mask_ptr: int = 0, path: int = PATH_AUTO, stream=None, precision: int = PRECISION_AUTO) -> None:
        """Asynchronous device-resident API: enqueue on `stream`, no sync.

        stream: None = the index's own stream; otherwise a cudaStream_t handle such as
        torch.cuda.current_stream().cuda_stream (0 = the default stream)."""
        p = _QueryParams(int(B), int(k), int(ns_code), int(label_code), mask_ptr or None, int(path), int(precision))
        _check(self._lib.cv_query_device(self._h, q_ptr, ctypes.byref(p), out_keys_ptr, _stream_arg(stream)))

    def query_device_exchange(self, q_ptr: int, B: int, k: int, out_keys_ptr: int, px: _PeerExchange, seq: int,
                              ns_code: int = ANY, label_code: int = ANY, path: int = PATH_AUTO, stream=None) -> None:
        """Row-sharded search with the cross-GPU exchange fused into the kernels (no NCCL call): scan this
        shard, store the per-GPU top-k into every peer's exchange buffer over NVLink, wait for the peers'
        flags, merge.  `out_keys_ptr` receives the global top-k.  Every rank must call with the same seq."""
        p = _QueryParams(int(B), int(k), int(ns_code), int(label_code), None, int(path), 0)
        _check(self._lib.cv_query_device_exchange(self._h, q_ptr, ctypes.byref(p), ctypes.byref(px), int(seq),
                                                  out_keys_ptr, _stream_arg(stream)))

    def query_exchange(self, queries: np.ndarray, k: int, px: _PeerExchange, seq: int, ns_code: int = ANY,
                       label_code: int = ANY, path: int = PATH_AUTO):
        """Host API of the row-sharded search (every rank calls it with the same queries and seq).  Returns the
        GLOBAL (scores [B][k], rows [B][k] (-1 pad), counts [B]) on every rank."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [B][{self.dim}], got {q.shape}")
        B = q.shape[0]
        p = _QueryParams(B, int(k), int(ns_code), int(label_code), None, int(path), 0)
        scores = np.empty((B, int(k)), dtype=np.float32)
        rows = np.empty((B, int(k)), dtype=np.int64)
        counts = np.empty(B, dtype=np.int32)
        _check(self._lib.cv_query_exchange(self._h, _ptr(q), ctypes.byref(p), ctypes.byref(px), int(seq), _ptr(scores),
                                           _ptr(rows), _ptr(counts)))
        return scores, rows, counts

    # -- readback / bookkeeping
    def get_rows(self, rows, want_vectors: bool = True):
        r = np.ascontiguousarray(rows, dtype=np.int64)
        n = r.shape[0]
        vec = np.empty((n, self.dim), dtype=np.uint16) if want_vectors else None
        scale = np.empty(n, dtype=np.float32)
        ns = np.empty(n, dtype=np.uint32)
        lb = np.empty(n, dtype=np.uint32)
        _check(self._lib.cv_get_rows(self._h, _ptr(r), n, _ptr(vec), _ptr(scale), _ptr(ns), _ptr(lb)))
        return vec, scale, ns, lb

    def debug_mma_scores(self, queries: np.ndarray, path: int = PATH_MMA, precision: int = PRECISION_AUTO) -> np.ndarray:
        """Every score the tensor-core path computes: [B][count] fp32 (testing aid)."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        out = np.empty((q.shape[0], len(self)), dtype=np.float32)
        _check(self._lib.cv_debug_mma_scores(self._h, _ptr(q), q.shape[0], int(path), int(precision), _ptr(out)))
        return out

    def set_profiling(self, enabled: bool) -> None:
        _check(self._lib.cv_set_profiling(self._h, 1 if enabled else 0))

    def stats(self) -> Stats:
        s = Stats()
        _check(self._lib.cv_stats(self._h, ctypes.byref(s)))
        return s

    def synchronize(self) -> None:
        _check(self._lib.cv_synchronize(self._h))


class StripedIndex:
    """One collection striped over G devices of THIS process (include/coddvec.h: cv_group).

    Global row r lives on shard r % G at local row r // G: appends (global rows are handed out densely, in order)
    stay balanced to within one row, and result keys carry the global row, so ties break exactly as on one GPU.
    The interface is NativeIndex's -- `Collection` uses either without knowing which -- and a query is ONE C call
    (cv_group_query) that puts the batch in flight on every device, lets the kernels exchange their per-GPU top-k
    over NVLink peer stores, and returns shard 0's merged answer.  This is how the reference's single-process
    service (codd_lib/codd_lib/client/provider/promql_module.py:59-67 builds one client for the whole process)
    gets more than one GPU.
    """

    def __init__(self, dim: int, metric: int = METRIC_COSINE, devices=(0, 1), capacity_rows: int = 0, max_rows: int = 0,
                 max_queries: int = 1024, max_k: int = MAX_K):
        devices = [int(d) for d in devices]
        if len(devices) < 2:
            raise ValueError("a striped index needs at least two shards")
        self.dim, self.metric, self.devices = int(dim), int(metric), devices
        G = self.G = len(devices)
        self.max_queries, self.max_k = int(max_queries), int(max_k)
        per = lambda n: -(-int(n) // G) if n else 0  # noqa: E731
        self.shards = [self._make_shard(dim, metric=metric, device=d, capacity_rows=per(capacity_rows),
                                        max_rows=per(max_rows), row_offset=g, row_stride=G) for g, d in enumerate(devices)]
        try:
            self._make_group()
        except Exception:
            for sh in self.shards:
                sh.close()
            raise
        self.device = devices[0]
        self.row_offset, self.row_stride = 0, 1  # (as seen from outside: dense global rows)

    # the three places that touch the C library (tests substitute CPU doubles for them; the product has no CPU path)
    def _make_shard(self, dim, **kw):
        return NativeIndex(dim, **kw)

    def _make_group(self) -> None:
        self._lib = load_library()
        handles = (ctypes.c_void_p * self.G)(*[sh._h for sh in self.shards])
        self._g = ctypes.c_void_p()
        _check(self._lib.cv_group_create(handles, self.G, self.max_queries, self.max_k, ctypes.byref(self._g)))

    def _group_query(self, q, k, ns_code, label_code, path, precision, shard_masks):
        B = q.shape[0]
        masks = None
        if shard_masks is not None:
            masks = (ctypes.c_void_p * self.G)()
            for g, w in enumerate(shard_masks):
                masks[g] = w.ctypes.data
        scores = np.empty((B, int(k)), dtype=np.float32)
        rows = np.empty((B, int(k)), dtype=np.int64)
        counts = np.empty(B, dtype=np.int32)
        p = _QueryParams(B, int(k), int(ns_code), int(label_code), None, int(path), int(precision))
        _check(self._lib.cv_group_query(self._g, _ptr(q), ctypes.byref(p), masks, _ptr(scores), _ptr(rows), _ptr(counts)))
        return scores, rows, counts

    def _destroy_group(self) -> None:
        if getattr(self, "_g", None) is not None and self._g:
            self._lib.cv_group_destroy(self._g)
            self._g = ctypes.c_void_p()

    # -- lifecycle
    def close(self) -> None:
        if getattr(self, "shards", None):
            self._destroy_group()
            for sh in self.shards:
                sh.close()
            self.shards = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self) -> int:
        return sum(len(sh) for sh in self.shards)

    count = __len__

    # -- K1
    def upsert(self, embeddings: np.ndarray, rows=None, ns_codes=None, label_codes=None) -> np.ndarray:
        emb = np.ascontiguousarray(embeddings, dtype=np.float32)
        if emb.ndim != 2 or emb.shape[1] != self.dim:
            raise ValueError(f"embeddings must be [n][{self.dim}], got {emb.shape}")
        n = emb.shape[0]
        if not np.isfinite(emb).all():  # checked for the WHOLE batch before any shard is touched
            raise ValueError("coddvec: embeddings contain NaN/Inf; nothing was committed")
        total = len(self)
        want = np.full(n, -1, dtype=np.int64) if rows is None else np.ascontiguousarray(rows, dtype=np.int64).copy()
        if want.shape != (n,):
            raise ValueError(f"rows must have shape ({n},)")
        if (want >= total).any():
            raise ValueError(f"coddvec: row {int(want.max())} is not committed (count = {total})")
        fresh = want < 0
        glob = want.copy()
        glob[fresh] = total + np.arange(int(fresh.sum()), dtype=np.int64)  # dense global rows, in batch order
        ns = None if ns_codes is None else np.ascontiguousarray(ns_codes, dtype=np.uint32)
        lb = None if label_codes is None else np.ascontiguousarray(label_codes, dtype=np.uint32)
        for g, sh in enumerate(self.shards):
            pos = np.nonzero(glob % self.G == g)[0]
            if pos.size == 0:
                continue
            local = np.where(fresh[pos], -1, glob[pos] // self.G)
            got = sh.upsert(emb[pos], rows=local, ns_codes=None if ns is None else ns[pos],
                            label_codes=None if lb is None else lb[pos])
            if not np.array_equal(got * self.G + g, glob[pos]):
                raise NativeError(CV_ERR_CUDA, "striped index lost its row numbering (a shard was written behind its back)")
        return glob

    def load_rows(self, stored_bf16: np.ndarray, scale: np.ndarray, ns_codes=None, label_codes=None) -> np.ndarray:
        vec = np.ascontiguousarray(stored_bf16, dtype=np.uint16)
        n = vec.shape[0]
        glob = len(self) + np.arange(n, dtype=np.int64)
        sc = np.ascontiguousarray(scale, dtype=np.float32)
        for g, sh in enumerate(self.shards):
            pos = np.nonzero(glob % self.G == g)[0]
            if pos.size:
                sh.load_rows(vec[pos], sc[pos], None if ns_codes is None else np.asarray(ns_codes)[pos],
                             None if label_codes is None else np.asarray(label_codes)[pos])
        return glob

    # -- K2 / K3 + fused exchange
    def query(self, queries: np.ndarray, k: int, ns_code: int = ANY, label_code: int = ANY, mask=None,
              path: int = PATH_AUTO, precision: int = PRECISION_AUTO, mask_bits=None):
        """Same contract as NativeIndex.query; `mask` is over GLOBAL rows (u32 words), or pass the bool array it was
        packed from as `mask_bits` to skip the unpack."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.ndim != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [B][{self.dim}], got {q.shape}")
        B = q.shape[0]
        total = len(self)
        shard_masks = None
        if mask is not None or mask_bits is not None:
            if mask_bits is None:
                words = np.ascontiguousarray(mask, dtype=np.uint32)
                if words.shape[0] < (total + 31) // 32:
                    raise ValueError(f"mask needs {(total + 31) // 32} uint32 words, got shape {words.shape}")
                mask_bits = np.unpackbits(words.view(np.uint8), bitorder="little")[:total].astype(bool)
            bits = np.asarray(mask_bits, dtype=bool)[:total]
            shard_masks = []
            for g in range(self.G):
                part = bits[g::self.G]
                pad = (-part.shape[0]) % 32
                if pad:
                    part = np.concatenate([part, np.zeros(pad, dtype=bool)])
                shard_masks.append(np.packbits(part, bitorder="little").view(np.uint32) if part.size else np.zeros(1, np.uint32))
        if B > self.max_queries:  # larger than the exchange buffers: slices
            scores = np.empty((B, int(k)), dtype=np.float32)
            rows = np.empty((B, int(k)), dtype=np.int64)
            counts = np.empty(B, dtype=np.int32)
            for lo in range(0, B, self.max_queries):
                hi = min(B, lo + self.max_queries)
                scores[lo:hi], rows[lo:hi], counts[lo:hi] = self.query(q[lo:hi], k, ns_code, label_code, path=path,
                                                                       precision=precision, mask_bits=mask_bits)
            return scores, rows, counts
        if not 1 <= int(k) <= MAX_K:
            raise ValueError(f"coddvec: k must be in [1, {MAX_K}] (got {k})")
        return self._group_query(q, int(k), ns_code, label_code, path, precision, shard_masks)

    def delete_rows(self, rows) -> None:
        r = np.ascontiguousarray(rows, dtype=np.int64)
        for g, sh in enumerate(self.shards):
            pos = np.nonzero(r % self.G == g)[0]
            if pos.size:
                sh.delete_rows(r[pos] // self.G)

    def set_codes(self, rows, ns_codes=None, label_codes=None) -> None:
        r = np.ascontiguousarray(rows, dtype=np.int64)
        for g, sh in enumerate(self.shards):
            pos = np.nonzero(r % self.G == g)[0]
            if pos.size:
                sh.set_codes(r[pos] // self.G, None if ns_codes is None else np.asarray(ns_codes)[pos],
                             None if label_codes is None else np.asarray(label_codes)[pos])

    # -- readback / bookkeeping
    def get_rows(self, rows, want_vectors: bool = True):
        r = np.ascontiguousarray(rows, dtype=np.int64)
        n = r.shape[0]
        vec = np.empty((n, self.dim), dtype=np.uint16) if want_vectors else None
        scale = np.empty(n, dtype=np.float32)
        ns = np.empty(n, dtype=np.uint32)
        lb = np.empty(n, dtype=np.uint32)
        if n and (r.min() < 0 or r.max() >= len(self)):
            raise ValueError(f"coddvec: row out of range [0, {len(self)})")
        for g, sh in enumerate(self.shards):
            pos = np.nonzero(r % self.G == g)[0]
            if pos.size == 0:
                continue
            v, s, a, b = sh.get_rows(r[pos] // self.G, want_vectors=want_vectors)
            if want_vectors:
                vec[pos] = v
            scale[pos], ns[pos], lb[pos] = s, a, b
        return vec, scale, ns, lb

    def set_profiling(self, enabled: bool) -> None:
        for sh in self.shards:
            sh.set_profiling(enabled)

    def stats(self) -> Stats:
        """Shard 0's view with the additive counters summed over the shards."""
        parts = [sh.stats() for sh in self.shards]
        s = parts[0]
        for f in ("rows", "capacity_rows", "kernel_launches", "upserted_rows", "last_query_bytes"):
            setattr(s, f, sum(getattr(p, f) for p in parts))
        return s

    def synchronize(self) -> None:
        for sh in self.shards:
            sh.synchronize()
