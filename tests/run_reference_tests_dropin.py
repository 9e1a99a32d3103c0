"""Run the REFERENCE's own store tests (unmodified files under /root/reference/tests) against this
package used as a drop-in `chromadb` (shim/ on sys.path).  Build-container only; executed in a
subprocess by tests/test_dropin_reference.py.

    python tests/run_reference_tests_dropin.py [--native]

Without --native the vector index is the oracle-backed test double (no GPU here); with --native (GPU box
with the reference tree available) it is libcoddvec.so.
"""
import importlib.abc
import importlib.machinery
import importlib.util
import inspect
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("CODD_REFERENCE", "/root/reference")
MISSING_OK = ("opus_agent_base", "pydantic_ai", "logfire", "rapidfuzz", "redis", "fakeredis")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Empty stand-ins for third-party packages the reference imports but this image lacks
    (none of them is on the semantic-store path)."""

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in MISSING_OK:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = types.ModuleType(spec.name)
        mod.__path__ = []
        mod.__getattr__ = lambda attr: type(attr, (), {})  # any attribute -> a dummy class
        return mod

    def exec_module(self, module):
        pass


def main():
    # (codd_lib is a nested project: REF/codd_lib/codd_lib is the importable package)
    sys.path[:0] = [os.path.join(ROOT, "shim"), ROOT, os.path.join(REF, "codd_lib"), REF]
    sys.meta_path.append(_StubFinder())
    if "--native" not in sys.argv:
        from codd_query_engine_b200 import client as client_mod
        from tests.doubles import OracleIndex

        client_mod.NativeIndex = OracleIndex
    import chromadb  # noqa: F401  (the shim)

    assert chromadb.__version__ == "coddvec-shim"
    ran = 0
    for rel, cls_name in (
        ("tests/unit/codd_dal/metrics/test_metrics_semantic_metadata_store.py", "TestMetricsSemanticMetadataStore"),
        ("tests/integration/codd_dal/metrics/test_metrics_semantic_metadata_store_integration.py",
         "TestMetricsSemanticMetadataStoreIntegration"),
    ):
        spec = importlib.util.spec_from_file_location("ref_" + cls_name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)  # imports the reference's MetricsSemanticMetadataStore + `chromadb`
        store_cls = mod.MetricsSemanticMetadataStore
        assert store_cls.__module__ == "codd_dal.metrics.metrics_semantic_metadata_store"
        suite = getattr(mod, cls_name)()
        for name, fn in inspect.getmembers(suite, predicate=inspect.ismethod):
            if not name.startswith("test_"):
                continue
            client = chromadb.EphemeralClient()              # what the reference's fixture builds
            store = store_cls(client, collection_name=f"test_{name}")
            fn(store)
            client.delete_collection(f"test_{name}")
            ran += 1
            print("PASS", rel.split("/")[-1], name)
    print(f"reference tests passed against the drop-in: {ran}")
    assert ran == 9, ran
    check_client_projection(chromadb)
    check_indexer_job(chromadb)


def check_indexer_job(chromadb):
    """the indexer job's per-metric loop (`_process_batch`: metric_exists ->
    enrich -> index_metadata, codd_jobs/metrics_semantic_indexer_job.py:359-461) over the drop-in, with the
    LLM enrichment agent replaced by a canned one (the LLM is out of scope, the store calls are not)."""
    try:
        from codd_jobs.metrics_semantic_indexer_job import IndexingStats, MetricsSemanticIndexerJob
    except Exception as exc:
        print("SKIP indexer job:", type(exc).__name__, exc)
        return
    from codd_dal.metrics.metrics_semantic_metadata_store import MetricsSemanticMetadataStore

    class CannedAgent:
        def enrich_metric_to_dict(self, metric_name, metric_type, description=None):
            return {"metric_name": metric_name, "type": metric_type, "description": description or metric_name,
                    "category": "application", "golden_signal_type": "latency" if "latency" in metric_name else "traffic",
                    "meter_type": metric_type}

    job = object.__new__(MetricsSemanticIndexerJob)
    job.semantic_store = MetricsSemanticMetadataStore(chromadb.HttpClient(host="localhost", port=8001))
    job.enrichment_agent = CannedAgent()
    job.stats = IndexingStats()
    batch = [{"metric": "http_request_latency_seconds", "type": "histogram", "help": "HTTP request latency"},
             {"metric": "http_requests_total", "type": "counter", "help": "Total HTTP requests"},
             {"metric": "bad metric name", "type": "gauge", "help": "rejected by the store's validation"}]
    job._process_batch("prod", batch, 1, 1, skip_if_present=True)
    assert (job.stats.indexed_metrics, job.stats.failed_metrics, job.stats.skipped_metrics) == (2, 1, 0), job.stats
    job._process_batch("prod", batch[:2], 1, 1, skip_if_present=True)   # second run: already present
    assert job.stats.skipped_metrics == 2 and job.stats.indexed_metrics == 2
    hits = job.semantic_store.search_metadata("request latency", n_results=2)
    assert hits[0]["metric_name"] == "http_request_latency_seconds" and hits[0]["namespace"] == "prod"
    print("PASS codd_jobs MetricsSemanticIndexerJob._process_batch over the drop-in")


def check_client_projection(chromadb):
    """codd_lib's MetricsPromQLClient.search_relevant_metrics (the dict ->
    SearchResult projection the service and CLI call) on top of the reference store on top of the drop-in."""
    try:
        from codd_lib.client.metrics_promql_client import MetricsPromQLClient
    except Exception as exc:  # codd_lib pulls in the LLM stack; skip if the stubs are not enough
        print("SKIP codd_lib projection:", type(exc).__name__, exc)
        return
    from codd_dal.metrics.metrics_semantic_metadata_store import MetricsSemanticMetadataStore

    store = MetricsSemanticMetadataStore(chromadb.HttpClient(host="localhost", port=8000), collection_name="a7")
    store.index_metadata("prod", {"metric_name": "http.latency", "type": "histogram",
                                  "description": "HTTP request latency in milliseconds", "unit": "ms",
                                  "golden_signal_type": "latency"})
    store.index_metadata("prod", {"metric_name": "db.connections", "description": "Active database connections"})
    fake_self = types.SimpleNamespace(semantic_metadata_store=store)
    hits = MetricsPromQLClient.search_relevant_metrics(fake_self, "request latency", limit=2)
    assert [h["metric_name"] for h in hits] == ["http.latency", "db.connections"]
    assert set(hits[0]) == {"metric_name", "similarity_score", "description", "unit", "category", "subcategory",
                            "category_description", "golden_signal_type", "golden_signal_description", "meter_type",
                            "meter_type_description"}
    assert hits[0]["unit"] == "ms" and 0 <= hits[0]["similarity_score"] <= 1
    print("PASS codd_lib MetricsPromQLClient.search_relevant_metrics over the drop-in")


if __name__ == "__main__":
    main()
