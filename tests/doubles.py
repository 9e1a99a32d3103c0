"""Test double for native.NativeIndex, backed by the CPU oracle.

TEST INFRASTRUCTURE ONLY: lets the host-side logic of client.py / sharded.py run in the CPU-only
container (`-m "not gpu"`).  The product never falls back to this: client.py always constructs
native.NativeIndex, tests substitute the class with monkeypatch.
"""
import numpy as np

from oracle import brute_force as bf


class OracleIndex:
    def __init__(self, dim, metric=bf.METRIC_COSINE, device=0, capacity_rows=0, max_rows=0, row_offset=0, row_stride=1):
        self.dim, self.metric, self.device, self.row_offset = int(dim), int(metric), int(device), int(row_offset)
        self.row_stride = int(row_stride) or 1
        self._stored = np.zeros((0, dim), np.uint16)
        self._scale = np.zeros(0, np.float32)
        self._ns = np.zeros(0, np.uint32)
        self._label = np.zeros(0, np.uint32)
        self.queries = 0

    def __len__(self):
        return self._stored.shape[0]

    count = __len__

    def close(self):
        pass

    def upsert(self, embeddings, rows=None, ns_codes=None, label_codes=None):
        emb = np.ascontiguousarray(embeddings, dtype=np.float32)
        if emb.ndim != 2 or emb.shape[1] != self.dim:
            raise ValueError("bad embedding shape")
        if not np.isfinite(emb).all():
            raise ValueError("coddvec: embeddings contain NaN/Inf")
        n = emb.shape[0]
        rows = np.full(n, -1, np.int64) if rows is None else np.asarray(rows, np.int64)
        stored, scale = bf.normalise_store(emb, self.metric)
        ns = np.zeros(n, np.uint32) if ns_codes is None else np.asarray(ns_codes, np.uint32)
        lb = np.zeros(n, np.uint32) if label_codes is None else np.asarray(label_codes, np.uint32)
        out = np.empty(n, np.int64)
        committed = len(self)
        app = int((rows < 0).sum())
        if (rows >= committed).any():
            raise ValueError("coddvec: row is not committed")
        self._stored = np.concatenate([self._stored, np.zeros((app, self.dim), np.uint16)])
        self._scale = np.concatenate([self._scale, np.zeros(app, np.float32)])
        self._ns = np.concatenate([self._ns, np.zeros(app, np.uint32)])
        self._label = np.concatenate([self._label, np.zeros(app, np.uint32)])
        nxt = committed
        for i in range(n):
            r = int(rows[i])
            if r < 0:
                r = nxt
                nxt += 1
            self._stored[r], self._scale[r], self._ns[r], self._label[r] = stored[i], scale[i], ns[i], lb[i]
            out[i] = r
        return out

    def delete_rows(self, rows):
        self._scale[np.asarray(rows, np.int64)] = np.nan

    def set_codes(self, rows, ns_codes=None, label_codes=None):
        r = np.asarray(rows, np.int64)
        if ns_codes is not None:
            self._ns[r] = np.asarray(ns_codes, np.uint32)
        if label_codes is not None:
            self._label[r] = np.asarray(label_codes, np.uint32)

    def query(self, queries, k, ns_code=bf.ANY_CODE, label_code=bf.ANY_CODE, mask=None, path=0, precision=0, mask_bits=None):
        if mask is None and mask_bits is not None:
            bits = np.asarray(mask_bits, dtype=bool)
            pad = (-bits.shape[0]) % 32
            mask = np.packbits(np.concatenate([bits, np.zeros(pad, dtype=bool)]) if pad else bits, bitorder="little").view(np.uint32)
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if not 1 <= k <= 128:
            raise ValueError("coddvec: k must be in [1, 128]")
        self.queries += 1
        self.last_requested_path = path  # (tests of the client's path choice read it)
        return bf.query(self._stored, self._scale, q, k, metric=self.metric, mask_words=mask, ns=self._ns,
                        ns_code=ns_code, label=self._label, label_code=label_code, row_offset=self.row_offset)

    def load_rows(self, stored_bf16, scale, ns_codes=None, label_codes=None):
        n, base = stored_bf16.shape[0], len(self)
        self._stored = np.concatenate([self._stored, np.asarray(stored_bf16, np.uint16)])
        self._scale = np.concatenate([self._scale, np.asarray(scale, np.float32)])
        self._ns = np.concatenate([self._ns, np.zeros(n, np.uint32) if ns_codes is None else np.asarray(ns_codes, np.uint32)])
        self._label = np.concatenate([self._label, np.zeros(n, np.uint32) if label_codes is None else np.asarray(label_codes, np.uint32)])
        return np.arange(base, base + n, dtype=np.int64)

    def get_rows(self, rows, want_vectors=True):
        r = np.asarray(rows, np.int64)
        return self._stored[r], self._scale[r], self._ns[r], self._label[r]

    def stats(self):
        return None


# ---- pointer-based entry points (used by sharded.ShardedIndex with CPU tensors under gloo) ----
def _view(ptr, shape, dtype):
    import ctypes

    n = int(np.prod(shape))
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(int(ptr))
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def _query_device(self, q_ptr, B, k, out_keys_ptr, ns_code=bf.ANY_CODE, label_code=bf.ANY_CODE, mask_ptr=0, path=0,
                  stream=None):
    q = _view(q_ptr, (B, self.dim), np.float32)
    s, r, c = self.query(q, k, ns_code=ns_code, label_code=label_code)
    keys = np.where(r >= 0, bf.pack_keys(s.ravel(), np.maximum(r, 0).ravel()).reshape(r.shape), np.uint64(0))
    _view(out_keys_ptr, (B, k), np.uint64)[...] = keys


def _upsert_device(self, emb_ptr, n, rows=None, ns_ptr=0, label_ptr=0):
    return self.upsert(_view(emb_ptr, (n, self.dim), np.float32).copy(), rows=rows)


OracleIndex.query_device = _query_device
OracleIndex.upsert_device = _upsert_device


def oracle_merge_keys_device(device, keys_in_ptr, G, B, k_in, k_out, out_ptr, stream=0):
    keys = _view(keys_in_ptr, (G, B, k_in), np.uint64)
    _view(out_ptr, (B, k_out), np.uint64)[...] = bf.merge_keys(keys, k_out)


def make_oracle_striped_index():
    """CPU double of native.StripedIndex: the real routing / numbering logic over OracleIndex shards, with the
    device group query replaced by per-shard oracle queries + the oracle's key merge."""
    from codd_query_engine_b200.native import StripedIndex

    class OracleStripedIndex(StripedIndex):
        def _make_shard(self, dim, **kw):
            return OracleIndex(dim, **kw)

        def _make_group(self):
            self._g = None

        def _destroy_group(self):
            pass

        def _group_query(self, q, k, ns_code, label_code, path, precision, shard_masks):
            parts = []
            for g, sh in enumerate(self.shards):
                if len(sh) == 0:
                    parts.append(np.zeros((q.shape[0], k), np.uint64))
                    continue
                s, r, c = sh.query(q, k, ns_code=ns_code, label_code=label_code,
                                   mask=None if shard_masks is None else shard_masks[g], path=path)
                glob = np.where(r >= 0, sh.row_offset + (r - sh.row_offset) * sh.row_stride, 0)  # local -> global row
                keys = np.where(r >= 0, bf.pack_keys(s.ravel(), glob.ravel()).reshape(r.shape), np.uint64(0))
                parts.append(keys)
            best = bf.merge_keys(np.stack(parts), k)
            scores, rows = bf.unpack_keys(best)
            return scores, rows, (best != bf.EMPTY_KEY).sum(axis=1).astype(np.int32)

    return OracleStripedIndex
