"""Row-sharded multi-GPU search: one process per GPU, per-GPU top-k, NCCL all-gather, K4 merge.

SURVEY.md section 8(e): rows are independent and top-k is decomposable (top-k of a union == top-k
of the per-shard top-k's), so the corpus is split into contiguous row blocks -- rank g holds global
rows [g*ceil(N/G), (g+1)*ceil(N/G)) -- every rank scans its shard for the whole query batch with the
same kernels as the single-GPU path (keys carry *global* row ids, so tie-breaking by row id needs no
fix-up), and the only exchange step is one all-gather of B*k u64 keys per rank over NVLink/NVSwitch
followed by the K4 merge kernel on every rank.  The reference has no distributed path at all
(everything is one chromadb server); this is the scale-out of
codd_dal/metrics/metrics_semantic_metadata_store.py:314-316.

torch is used for what it is good at here: device buffers, streams and torch.distributed.
"""

from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import native
from .native import NativeIndex, merge_keys_device


def shard_bounds(total_rows: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block partition: [lo, hi) of global rows owned by `rank`."""
    per = -(-int(total_rows) // int(world)) if total_rows > 0 else 0
    lo = min(rank * per, total_rows)
    return lo, min(lo + per, total_rows)


class ShardedIndex:
    """SPMD handle: every rank constructs it and calls the same methods in the same order."""

    def __init__(self, dim: int, total_rows: int, metric: int = native.METRIC_COSINE,
                 group: Optional[dist.ProcessGroup] = None, device: Optional[torch.device] = None,
                 exchange: str = "auto", max_queries: int = 1024, max_k: int = 32):
        """exchange: "fused" = the kernels exchange per-GPU top-k themselves with NVLink peer stores into
        symmetric buffers (no collective call per query); "nccl" = all-gather + K4 merge; "auto" = fused when
        torch symmetric memory can map the peers, else nccl."""
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dim = int(dim)
        self.total_rows = int(total_rows)
        self.lo, self.hi = shard_bounds(total_rows, self.world, self.rank)
        dev_ordinal = self.device.index if self.device.type == "cuda" else 0
        self.local = NativeIndex(dim, metric=metric, device=dev_ordinal or 0, capacity_rows=max(self.hi - self.lo, 0),
                                 max_rows=max(self.hi - self.lo, 1), row_offset=self.lo)
        self._bufs: dict[tuple[int, int], tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}
        self._host: dict[tuple[int, int], tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}
        self._px = None
        self._seq = 0
        self.exchange = "nccl"
        if self.world > 1 and exchange in ("auto", "fused") and self.device.type == "cuda":
            try:
                self._setup_fused_exchange(int(max_queries), int(max_k))
                self.exchange = "fused"
            except Exception as exc:  # symmetric memory unavailable (no P2P / IPC): keep the NCCL path
                if exchange == "fused":
                    raise
                self._px = None
                self._exchange_error = repr(exc)
            # every rank must take the same path
            ok = torch.tensor([1 if self._px is not None else 0], device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
            if int(ok.item()) == 0:
                self._px, self.exchange = None, "nccl"

    def _setup_fused_exchange(self, max_queries: int, max_k: int) -> None:
        import torch.distributed._symmetric_memory as symm_mem

        nbytes = native.peer_exchange_bytes(self.world, max_queries, max_k)
        group = self.group if self.group is not None else dist.group.WORLD
        buf = symm_mem.empty(nbytes // 8, dtype=torch.int64, device=self.device)
        buf.zero_()
        handle = symm_mem.rendezvous(buf, group=group)
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)  # every rank's buffer is zeroed before anyone publishes into it
        self._xbuf, self._xhandle = buf, handle
        self._px = native.make_peer_exchange(self.world, self.rank, max_queries, max_k, list(handle.buffer_ptrs))
        self._px_caps = (max_queries, max_k)

    # ------------------------------------------------------------------ loading
    def owns(self, global_row: int) -> bool:
        return self.lo <= global_row < self.hi

    def append_local(self, embeddings: np.ndarray, ns_codes=None, label_codes=None) -> np.ndarray:
        """Append rows to this rank's shard (host embeddings); returns their *global* rows."""
        return self.local.upsert(embeddings, ns_codes=ns_codes, label_codes=label_codes) + self.lo

    def append_local_device(self, emb: torch.Tensor) -> None:
        """Append rows already resident on this rank's GPU ([n][dim] fp32)."""
        assert emb.dtype == torch.float32 and emb.dim() == 2 and emb.shape[1] == self.dim and emb.is_contiguous()
        self.local.upsert_device(emb.data_ptr(), emb.shape[0])

    def count(self) -> int:
        n = torch.tensor([len(self.local)], dtype=torch.int64, device=self.device)
        if self.world > 1:
            dist.all_reduce(n, group=self.group)
        return int(n.item())

    # ------------------------------------------------------------------ search
    def _buffers(self, B: int, k: int):
        key = (B, k)
        if key not in self._bufs:
            local = torch.zeros((B, k), dtype=torch.int64, device=self.device)
            gathered = torch.zeros((self.world, B, k), dtype=torch.int64, device=self.device)
            merged = torch.zeros((B, k), dtype=torch.int64, device=self.device)
            self._bufs[key] = (local, gathered, merged)
        return self._bufs[key]

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream if self.device.type == "cuda" else 0

    def query_device(self, q: torch.Tensor, k: int, ns_code: int = native.ANY, label_code: int = native.ANY,
                     path: int = native.PATH_AUTO, exchange: Optional[str] = None) -> torch.Tensor:
        """q: [B][dim] fp32 on this rank's device (same values on every rank).  Returns the merged
        [B][k] int64 key tensor (u64 bit patterns) on every rank.  Asynchronous on the current stream.
        exchange="nccl" forces the all-gather + K4 exchange for this call (every rank must pass the same value)."""
        B = q.shape[0]
        local, gathered, merged = self._buffers(B, k)
        stream = self._stream()
        fused = exchange != "nccl"
        if fused and self.world > 1 and self._px is not None and B <= self._px_caps[0] and k <= self._px_caps[1]:
            self._seq += 1  # fused: scan + NVLink peer stores + flag wait + merge, no collective launch
            self.local.query_device_exchange(q.data_ptr(), B, k, merged.data_ptr(), self._px, self._seq, ns_code=ns_code,
                                             label_code=label_code, path=path, stream=stream)
            return merged
        self.local.query_device(q.data_ptr(), B, k, local.data_ptr(), ns_code=ns_code, label_code=label_code,
                                path=path, stream=stream)
        if self.world == 1:
            return local
        dist.all_gather_into_tensor(gathered.view(self.world * B, k), local, group=self.group)  # NCCL over NVLink
        merge_keys_device(self.device.index or 0, gathered.data_ptr(), self.world, B, k, k, merged.data_ptr(), stream)
        return merged

    def query(self, queries: np.ndarray, k: int, **kw):
        """Host API (same on every rank): pinned H2D of the batch, sharded search, D2H of the merged keys.
        Returns (scores [B][k], global rows [B][k] (-1 pad), counts [B])."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if self.world == 1 and self.device.type == "cuda":
            # one shard: the C-ABI's own host entry point (cv_query) -- for interactive batches it has no
            # copy-engine operations or stream synchronise on the path (see native.NativeIndex.query)
            return self.local.query(q, k, **kw)
        B = q.shape[0]
        if (self.world > 1 and self._px is not None and B <= self._px_caps[0] and k <= self._px_caps[1]
                and os.environ.get("CODDVEC_HOST_EXCHANGE", "1") != "0"):
            # fused exchange through the C-ABI's host entry point: no torch copies or stream synchronise around it
            self._seq += 1
            return self.local.query_exchange(q, k, self._px, self._seq, **kw)
        key = (B, k)
        if key not in self._host:
            pin = self.device.type == "cuda"
            h_q = torch.empty((B, self.dim), dtype=torch.float32, pin_memory=pin)
            d_q = torch.empty((B, self.dim), dtype=torch.float32, device=self.device)
            h_out = torch.empty((B, k), dtype=torch.int64, pin_memory=pin)
            self._host[key] = (h_q, d_q, h_out)
        h_q, d_q, h_out = self._host[key]
        h_q.numpy()[...] = q
        d_q.copy_(h_q, non_blocking=True)
        keys = self.query_device(d_q, k, **kw)
        h_out.copy_(keys, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream(self.device).synchronize()
        arr = h_out.numpy().view(np.uint64)
        scores, rows = native.decode_keys(arr)
        counts = (arr != 0).sum(axis=1).astype(np.int32)
        return scores, rows, counts

    def close(self) -> None:
        self.local.close()
