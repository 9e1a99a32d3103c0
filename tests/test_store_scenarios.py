"""The reference's own executable spec for the path, re-run against this package.

Scenarios follow tests/unit/codd_dal/metrics/test_metrics_semantic_metadata_store.py:39-216 and
tests/integration/codd_dal/metrics/test_metrics_semantic_metadata_store_integration.py:26-175 of the
reference (8 live unit tests + 1 integration workflow).  They run twice: on the CPU with the
oracle-backed test double substituted for native.NativeIndex (host logic), and -- marked gpu --
through the real CUDA library.
"""
import pytest

from codd_query_engine_b200 import client as client_mod
from codd_query_engine_b200.semantic_store import MetricsSemanticMetadataStore
from tests.doubles import OracleIndex


@pytest.fixture(params=["double", pytest.param("native", marks=pytest.mark.gpu)])
def chromadb_client(request, monkeypatch):
    if request.param == "double":
        monkeypatch.setattr(client_mod, "NativeIndex", OracleIndex)
    c = client_mod.Client(endpoint=("test", request.node.name))
    yield c
    c.reset()


@pytest.fixture
def store(chromadb_client, request):
    return MetricsSemanticMetadataStore(chromadb_client, collection_name=f"test_{request.node.name}")


def test_index_metadata_basic(store):
    md = {"metric_name": "cpu.usage", "type": "gauge", "description": "CPU utilization percentage",
          "unit": "percent", "category": "system", "subcategory": "cpu"}
    assert store.index_metadata("test", md) == "test#cpu.usage"


def test_index_metadata_all_fields(store):
    md = {"metric_name": "http.request.duration", "type": "histogram",
          "description": "HTTP request duration in milliseconds", "unit": "ms", "category": "application",
          "subcategory": "http", "category_description": "Application-level metrics",
          "golden_signal_type": "latency", "golden_signal_description": "Measures request latency",
          "meter_type": "histogram", "meter_type_description": "Distribution of values over time"}
    assert store.index_metadata("test", md) == "test#http.request.duration"


def test_index_metadata_missing_required_field(store):
    with pytest.raises(KeyError) as exc:
        store.index_metadata("test", {"type": "gauge", "description": "Some metric"})
    assert "metric_name" in str(exc.value)


def test_index_metadata_upsert(store):
    ns, name = "test_namespace", "memory.usage"
    store.index_metadata(ns, {"metric_name": name, "description": "Memory usage version 1", "category": "system"})
    out = store.index_metadata(ns, {"metric_name": name, "description": "Memory usage version 2 updated",
                                    "category": "infrastructure"})
    assert out == f"{ns}#{name}"
    hits = store.search_metadata("memory usage version 2", n_results=1)
    assert len(hits) > 0 and hits[0]["metric_name"] == name and "version 2" in hits[0]["description"]
    assert store.collection.count() == 1


def test_search_metadata_basic(store):
    store.index_metadata("test", {"metric_name": "cpu.usage", "description": "CPU utilization percentage", "category": "system"})
    store.index_metadata("test", {"metric_name": "memory.usage", "description": "Memory utilization in bytes", "category": "system"})
    hits = store.search_metadata("CPU utilization")
    assert len(hits) > 0 and hits[0]["metric_name"] == "cpu.usage"
    assert "similarity_score" in hits[0] and 0 <= hits[0]["similarity_score"] <= 1


def test_search_metadata_no_results(store):
    assert store.search_metadata("some query") == []


def test_search_metadata_ranking(store):
    store.index_metadata("test", {"metric_name": "http.latency", "description": "HTTP request latency in milliseconds",
                                  "golden_signal_type": "latency"})
    store.index_metadata("test", {"metric_name": "db.query.time", "description": "Database query execution time",
                                  "category": "database"})
    store.index_metadata("test", {"metric_name": "network.bandwidth", "description": "Network bandwidth usage",
                                  "category": "network"})
    hits = store.search_metadata("request latency")
    assert len(hits) == 3 and hits[0]["metric_name"] == "http.latency"
    for a, b in zip(hits, hits[1:]):
        assert a["similarity_score"] >= b["similarity_score"]


def test_search_metadata_returns_all_fields(store):
    md = {"metric_name": "test.metric", "type": "gauge", "description": "Test description", "unit": "bytes",
          "category": "test_category", "subcategory": "test_subcategory", "category_description": "Category desc",
          "golden_signal_type": "throughput", "golden_signal_description": "Measures throughput",
          "meter_type": "gauge", "meter_type_description": "Gauge meter"}
    store.index_metadata("test", md)
    hits = store.search_metadata("test metric")
    assert len(hits) > 0
    for key, val in md.items():
        assert hits[0][key] == val
    assert hits[0]["namespace"] == "test"


def test_metric_exists(store):
    assert store.metric_exists("ns", "m") is False
    store.index_metadata("ns", {"metric_name": "m", "description": "d"})
    assert store.metric_exists("ns", "m") is True
    assert store.metric_exists("other", "m") is False


def test_complete_metric_semantic_metadata_workflow(store):
    ns = "test:order_service"
    common = {"category": "application", "subcategory": "http", "category_description": "Application-level performance metrics"}
    a = store.index_metadata(ns, {"metric_name": "http.request.duration.p99", "type": "histogram",
                                  "description": "99th percentile HTTP request duration for all endpoints",
                                  "unit": "milliseconds", **common, "golden_signal_type": "latency",
                                  "golden_signal_description": "Time taken to service requests",
                                  "meter_type": "histogram",
                                  "meter_type_description": "Distribution of request durations over time"})
    assert a == f"{ns}#http.request.duration.p99"
    b = store.index_metadata(ns, {"metric_name": "http.errors.5xx.rate", "type": "counter",
                                  "description": "Rate of 5xx server errors per second", "unit": "errors_per_second",
                                  **common, "golden_signal_type": "errors",
                                  "golden_signal_description": "Rate of failed requests", "meter_type": "counter",
                                  "meter_type_description": "Monotonically increasing count of errors"})
    assert b == f"{ns}#http.errors.5xx.rate"
    c = store.index_metadata(ns, {"metric_name": "http.requests.total", "type": "counter",
                                  "description": "Total number of HTTP requests received", "unit": "requests",
                                  **common, "golden_signal_type": "traffic",
                                  "golden_signal_description": "Volume of requests served", "meter_type": "counter",
                                  "meter_type_description": "Cumulative count of requests"})
    assert c == f"{ns}#http.requests.total"

    hits = store.search_metadata("request latency and response time", n_results=5)
    assert len(hits) >= 2
    assert hits[0]["metric_name"] == "http.request.duration.p99" and hits[0]["golden_signal_type"] == "latency"
    assert 0 <= hits[0]["similarity_score"] <= 1

    ns2 = "test:payment_service"
    store.index_metadata(ns2, {"metric_name": "db.query.execution.time", "type": "histogram",
                               "description": "Database query execution time in milliseconds", "unit": "milliseconds",
                               "category": "infrastructure", "subcategory": "database", "golden_signal_type": "latency",
                               "meter_type": "histogram"})
    store.index_metadata(ns2, {"metric_name": "db.connections.active", "type": "gauge",
                               "description": "Number of active database connections", "unit": "connections",
                               "category": "infrastructure", "subcategory": "database",
                               "golden_signal_type": "saturation", "meter_type": "gauge"})
    # search is NOT namespace-scoped (reference integration test :143-148)
    db_hits = store.search_metadata("latency", n_results=5)
    assert len(db_hits) >= 1 and "db.query.execution.time" in [h["metric_name"] for h in db_hits[:3]]

    again = store.index_metadata(ns, {"metric_name": "http.request.duration.p99", "type": "histogram",
                                      "description": "UPDATED: 99th percentile latency for HTTP requests with improved accuracy",
                                      "unit": "milliseconds", "category": "application", "subcategory": "http",
                                      "golden_signal_type": "latency", "meter_type": "histogram"})
    assert again == a
    upd = store.search_metadata("improved accuracy latency", n_results=1)
    assert len(upd) > 0 and upd[0]["metric_name"] == "http.request.duration.p99"
    assert "UPDATED" in upd[0]["description"] and "improved accuracy" in upd[0]["description"]
    assert store.collection.count() == 5


# ---- extensions: the bulk calls must be indistinguishable from the per-item ones
def _catalogue(n):
    kinds = [("cpu", "CPU utilization percentage of the host"), ("memory", "Memory usage in bytes"),
             ("http.latency", "HTTP request latency in milliseconds"), ("disk.io", "Disk read and write operations"),
             ("queue.depth", "Messages waiting in the queue")]
    return [{"metric_name": f"{kinds[i % 5][0]}.{i}", "description": f"{kinds[i % 5][1]} (series {i})",
             "category": "system" if i % 2 else "application", "type": "gauge"} for i in range(n)]


def test_bulk_index_equals_per_item_index(chromadb_client, request):
    one = MetricsSemanticMetadataStore(chromadb_client, collection_name=f"one_{request.node.name}")
    bulk = MetricsSemanticMetadataStore(chromadb_client, collection_name=f"bulk_{request.node.name}")
    items = _catalogue(2300)                                       # three upsert chunks of <= 1000
    ids_one = [one.index_metadata("prod", md) for md in items[:300]]
    ids_bulk = bulk.index_metadata_bulk("prod", items)
    assert ids_bulk[:300] == ids_one and len(ids_bulk) == 2300 and bulk.collection.count() == 2300
    for md in items[300:]:
        one.index_metadata("prod", md)
    for q in ("CPU utilization", "request latency milliseconds", "queue messages"):
        a, b = one.search_metadata(q, n_results=7), bulk.search_metadata(q, n_results=7)
        assert [h["metric_name"] for h in a] == [h["metric_name"] for h in b]
        assert all(abs(x["similarity_score"] - y["similarity_score"]) < 1e-6 for x, y in zip(a, b))
    assert bulk.metric_exists("prod", "cpu.0") and not bulk.metric_exists("prod", "nope")


def test_bulk_index_validates_everything_before_writing(store):
    bad = _catalogue(5) + [{"metric_name": "bad name with spaces", "description": "x"}]
    with pytest.raises(Exception) as exc:
        store.index_metadata_bulk("prod", bad)
    assert "ValidationError" in type(exc.value).__name__
    assert store.collection.count() == 0
    with pytest.raises(KeyError):
        store.index_metadata_bulk("prod", [{"description": "no name"}])
    # a repeated id keeps the LAST entry, as successive index_metadata calls would
    ids = store.index_metadata_bulk("prod", [{"metric_name": "m", "description": "first text about disks"},
                                             {"metric_name": "m", "description": "CPU utilization percentage"}])
    assert ids == ["prod#m", "prod#m"] and store.collection.count() == 1
    assert "CPU" in store.search_metadata("CPU utilization")[0]["description"]


def test_batched_search_equals_single_searches(store):
    store.index_metadata_bulk("prod", _catalogue(400))
    queries = ["CPU utilization", "", "request latency milliseconds", "   ", "disk read write", "queue messages waiting"]
    batched = store.search_metadata_batch(queries, n_results=5)
    assert len(batched) == len(queries) and batched[1] == [] and batched[3] == []
    for q, got in zip(queries, batched):
        want = store.search_metadata(q, n_results=5)
        assert len(got) == len(want)
        # the batch may run on the tensor path, which rounds the queries to bf16: |dscore| <= 2^-9 * sum|q_i c_i|.
        # Dense 384-d embeddings average that down to ~1e-4 (north-star tolerance 1e-3); the hashed bag-of-words
        # vectors here have a handful of non-zeros, so only the analytic bound 2^-8 (with the rows' own bf16
        # storage error) holds.  The catalogue's near-duplicate documents tie within it, so the hit sets may
        # differ -- but only among hits that tie with the single search's k-th
        tol = 2.0 ** -8
        if not want:
            continue
        kth = want[-1]["similarity_score"]
        names = {h["metric_name"] for h in want}
        for g in got:
            assert 0.0 <= g["similarity_score"] <= 1.0 + 1e-6 and g["namespace"] == "prod"
            assert g["metric_name"] in names or g["similarity_score"] >= kth - tol
        for g, w in zip(got, want):
            assert abs(g["similarity_score"] - w["similarity_score"]) < tol
    with pytest.raises(Exception):
        store.search_metadata_batch(["x" * 1001])
