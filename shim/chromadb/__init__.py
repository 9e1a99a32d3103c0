"""Import shim: put this directory on PYTHONPATH to make ``import chromadb`` resolve to the
B200-native store, so codd_lib / codd_jobs / codd_service / codd_mcp_server run with zero source
edits (their only chromadb uses are ``chromadb.HttpClient(host=..., port=...)``,
``chromadb.EphemeralClient()`` and the ``chromadb.Client`` / ``chromadb.ClientAPI`` annotations:
codd_lib/codd_lib/client/provider/promql_module.py:59-62,
codd_jobs/metrics_semantic_indexer_main.py:209-212, :250, :326-329,
codd_dal/metrics/metrics_semantic_metadata_store.py:45)."""

from codd_query_engine_b200.client import (  # noqa: F401
    Client,
    ClientAPI,
    Collection,
    DuplicateIDError,
    EphemeralClient,
    HttpClient,
    NotFoundError,
    PersistentClient,
)

__version__ = "coddvec-shim"
