"""Deterministic text embedder used when no embedding model is available.

The reference never passes an embedding function (codd_dal/metrics/metrics_semantic_metadata_store.py:60-69),
so chromadb falls back to its default ONNX all-MiniLM-L6-v2 model (384-d, L2-normalised) which is
downloaded on first use.  There are no model weights and no network here, so the default of this
package is a hashed bag-of-words / character-trigram embedder with the same shape contract:
``callable(list[str]) -> [n][384] float32``, unit length, non-negative (so cosine lies in [0, 1] as
every reference test assumes).  Any real model can be injected through ``embedding_function=``
exactly as with chromadb.
"""

from __future__ import annotations

import re
import zlib
from functools import lru_cache

import numpy as np

DEFAULT_DIM = 384  # all-MiniLM-L6-v2 output size, the dimension BASELINE.json quotes

_TOKEN = re.compile(r"[^\W_]+", re.UNICODE)
_STOP = frozenset(
    "a an and are as at be by for from in is it of on or per the to with all".split()
)


@lru_cache(maxsize=1 << 16)
def _token_features(tok: str, dim: int, trigram_weight: float) -> tuple:
    """(slot, weight) pairs of one lower-cased token, in the order they are accumulated.  Metric catalogues repeat
    the same few thousand words, so the hashing is done once per distinct token."""
    feats = [(zlib.crc32(("w:" + tok).encode("utf-8")) % dim, 1.0)]
    if trigram_weight > 0 and len(tok) > 3:
        marked = f"^{tok}$"
        feats += [(zlib.crc32(("g:" + marked[i:i + 3]).encode("utf-8")) % dim, trigram_weight) for i in range(len(marked) - 2)]
    return tuple(feats)


class HashedBagOfWordsEmbedding:
    """Feature-hashing embedder: words (weight 1) + boundary-marked character trigrams (weight 0.3)."""

    def __init__(self, dim: int = DEFAULT_DIM, trigram_weight: float = 0.3):
        if dim < 8:
            raise ValueError("dim must be >= 8")
        self.dim = int(dim)
        self.trigram_weight = float(trigram_weight)

    @staticmethod
    def _slot(feature: str, dim: int) -> int:
        return zlib.crc32(feature.encode("utf-8")) % dim

    def embed_one(self, text: str) -> np.ndarray:
        acc: dict[int, float] = {}
        for tok in _TOKEN.findall(text.lower()):
            if tok in _STOP:
                continue
            for slot, weight in _token_features(tok, self.dim, self.trigram_weight):
                acc[slot] = acc.get(slot, 0.0) + weight
        v = np.zeros(self.dim, dtype=np.float64)
        if acc:
            v[list(acc.keys())] = list(acc.values())
        np.sqrt(v, out=v)  # sub-linear term frequency
        n = np.linalg.norm(v)
        if n > 0:
            v /= n
        return v.astype(np.float32)

    def __call__(self, input):  # noqa: A002 - chromadb's EmbeddingFunction signature
        if isinstance(input, str):
            input = [input]
        if len(input) == 0:
            return np.zeros((0, self.dim), dtype=np.float32)
        return np.stack([self.embed_one(str(t)) for t in input])

    def name(self) -> str:
        return f"hashed-bow-{self.dim}"


def default_embedding_function() -> HashedBagOfWordsEmbedding:
    return HashedBagOfWordsEmbedding()
