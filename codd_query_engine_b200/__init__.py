"""codd_query_engine_b200 -- B200-native semantic top-k store for the Codd query engine.

One hot path, rebuilt from scratch for sm_100a: the cosine / inner-product top-k retrieval behind
``search_relevant_metrics`` (codd_dal/metrics/metrics_semantic_metadata_store.py in the reference)
and the matching upsert path.  Layout:

* ``csrc/``            hand-written CUDA kernels (K1 normalise/append, K2 streaming scan, K3 tcgen05
                       batched scoring, K4 merge) and the C-ABI (include/coddvec.h -> libcoddvec.so)
* ``native``           ctypes binding of that C-ABI (no CPU fallback)
* ``client``           chromadb-compatible Client / Collection (the drop-in boundary)
* ``semantic_store``   mirror of the reference's MetricsSemanticMetadataStore facade
* ``sharded``          row-sharded multi-GPU search (one process per GPU, NCCL all-gather + K4 merge)
* ``embedding``        deterministic text embedder used when no model is available
"""

from .client import Client, ClientAPI, Collection, EphemeralClient, HttpClient, PersistentClient  # noqa: F401
from .semantic_store import MetricsSemanticMetadataStore, ValidationError  # noqa: F401

__version__ = "0.1.0"
