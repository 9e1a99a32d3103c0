"""Generate tests/golden/store_calls.json by running the REFERENCE's own
MetricsSemanticMetadataStore (codd_dal/metrics/metrics_semantic_metadata_store.py) over a recording
fake chromadb client.  Run in the build container only (needs /root/reference):

    python tests/golden/make_store_golden.py

The fixture pins, for a list of inputs: the exact boundary calls the reference issues (documents,
metadatas, ids, query_texts, n_results), its return values, and the exception type + message where it
raises.  tests/test_semantic_store.py replays the same inputs through
codd_query_engine_b200.semantic_store and requires identical behaviour.
"""
import json
import os
import sys
import types

REF = os.environ.get("CODD_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__path__ = []  # behave like a package
    sys.modules[name] = mod
    return mod


def load_reference_store():
    sys.path.insert(0, REF)
    _stub("chromadb", Client=object, ClientAPI=object, HttpClient=object, EphemeralClient=object)
    try:
        import codd_dal.metrics.metrics_semantic_metadata_store as ref  # noqa
    except ModuleNotFoundError as exc:
        # the reference package __init__ files may pull optional third-party deps: stub and retry
        missing = exc.name
        _stub(missing)
        import importlib
        ref = importlib.import_module("codd_dal.metrics.metrics_semantic_metadata_store")
    return ref


class Recorder:
    """Fake chromadb client + collection that records calls and replays canned query results."""

    def __init__(self):
        self.calls = []
        self.canned_query = None
        self.canned_get = None
        self.raise_on_get = False

    def get_or_create_collection(self, **kw):
        self.calls.append(["get_or_create_collection", kw])
        return self

    def upsert(self, **kw):
        self.calls.append(["upsert", kw])

    def get(self, **kw):
        self.calls.append(["get", kw])
        if self.raise_on_get:
            raise RuntimeError("backend down")
        return self.canned_get

    def query(self, **kw):
        self.calls.append(["query", kw])
        return self.canned_query


INDEX_CASES = [
    ("basic", "test", {"metric_name": "cpu.usage", "type": "gauge", "description": "CPU utilization percentage",
                       "unit": "percent", "category": "system", "subcategory": "cpu"}),
    ("all_fields", "test", {"metric_name": "http.request.duration", "type": "histogram",
                            "description": "HTTP request duration in milliseconds", "unit": "ms",
                            "category": "application", "subcategory": "http",
                            "category_description": "Application-level metrics", "golden_signal_type": "latency",
                            "golden_signal_description": "Measures request latency", "meter_type": "histogram",
                            "meter_type_description": "Distribution of values over time"}),
    ("missing_name", "test", {"type": "gauge", "description": "Some metric"}),
    ("name_only", "ns", {"metric_name": "only_name"}),
    ("no_description", "ns", {"metric_name": "a.b", "category": "system", "meter_type": "gauge"}),
    ("whitespace_and_nul", "ns", {"metric_name": "w/s-1", "description": "  lots   of\n\tspace \x00 here  ",
                                  "category": " sys\x00tem ", "unit": "\x00"}),
    ("none_values", "ns", {"metric_name": "n.v", "description": None, "unit": None, "category": "c"}),
    ("non_string_values", "ns", {"metric_name": 12345, "description": 42, "unit": 3.5, "category": True}),
    ("unicode", "ns:svc", {"metric_name": "latência.p99_ü", "description": "Latência do pedido — p99 ✓",
                           "golden_signal_type": "latency"}),
    ("namespace_with_hash", "a#b", {"metric_name": "m1", "description": "d"}),
    ("empty_name", "ns", {"metric_name": ""}),
    ("bad_chars", "ns", {"metric_name": "cpu usage"}),
    ("bad_chars_hash", "ns", {"metric_name": "a#b"}),
    ("name_too_long", "ns", {"metric_name": "x" * 256}),
    ("name_max_len", "ns", {"metric_name": "x" * 255}),
    ("field_too_long", "ns", {"metric_name": "ok", "description": "d" * 2001}),
    ("field_max_len", "ns", {"metric_name": "ok", "description": "d" * 2000}),
    ("second_field_too_long", "ns", {"metric_name": "ok", "description": "fine", "meter_type_description": "m" * 2001}),
    ("extra_keys_ignored", "ns", {"metric_name": "e.k", "description": "d", "labels": "x", "foo": "bar"}),
    ("empty_strings", "ns", {"metric_name": "e.s", "description": "", "category": "", "subcategory": "sub"}),
]

FULL_MD = {"type": "gauge", "description": "Test description", "unit": "bytes", "category": "test_category",
           "subcategory": "test_subcategory", "category_description": "Category desc",
           "golden_signal_type": "throughput", "golden_signal_description": "Measures throughput",
           "meter_type": "gauge", "meter_type_description": "Gauge meter", "namespace": "test"}

SEARCH_CASES = [
    ("basic", "CPU utilization", 10,
     {"ids": [["test#cpu.usage", "test#memory.usage"]], "distances": [[0.25, 0.75]],
      "metadatas": [[dict(FULL_MD), dict(FULL_MD, description="other")]], "documents": [["a", "b"]]}),
    ("empty_query", "", 10, None),
    ("blank_query", "   \t\n", 10, None),
    ("sanitised_query", "  request \x00  latency\n\n now ", 5, {"ids": [[]], "distances": [[]], "metadatas": [[]]}),
    ("query_too_long", "q" * 1001, 10, None),
    ("query_max_len", "q" * 1000, 10, {"ids": [[]], "distances": [[]], "metadatas": [[]]}),
    ("n_zero", "x", 0, None),
    ("n_negative", "x", -3, None),
    ("n_clamped", "x", 500, {"ids": [["ns#m"]], "distances": [[0.5]], "metadatas": [[{"namespace": "ns"}]]}),
    ("n_exact_max", "x", 100, {"ids": [["ns#m"]], "distances": [[0.5]], "metadatas": [[{"namespace": "ns"}]]}),
    ("no_ids_key", "x", 3, {"distances": [[]]}),
    ("none_result", "x", 3, None),
    ("id_without_hash", "x", 3, {"ids": [["plainid"]], "distances": [[0.1]], "metadatas": [[{"a": "b"}]]}),
    ("id_with_two_hashes", "x", 3, {"ids": [["a#b#metric.name"]], "distances": [[0.0]], "metadatas": [[{}]]}),
    ("short_metadata_list", "x", 3, {"ids": [["n#a", "n#b"]], "distances": [[0.1]], "metadatas": [[{"k": "v"}]]}),
    ("metadata_overrides_metric_name", "x", 3,
     {"ids": [["n#real"]], "distances": [[0.2]], "metadatas": [[{"metric_name": "from_md", "similarity_score": 9}]]}),
]

EXISTS_CASES = [
    ("found", "ns", "m", {"ids": ["ns#m"], "metadatas": [{}]}, False),
    ("not_found", "ns", "m", {"ids": [], "metadatas": []}, False),
    ("none", "ns", "m", None, False),
    ("backend_error", "ns", "m", None, True),
]


def outcome(fn):
    try:
        return {"returns": fn()}
    except Exception as exc:  # noqa: BLE001
        return {"raises": type(exc).__name__, "message": str(exc)}


def main():
    ref = load_reference_store()
    out = {"constants": {k: getattr(ref, k) for k in ("MAX_METRIC_NAME_LENGTH", "MAX_TEXT_FIELD_LENGTH",
                                                       "MAX_QUERY_LENGTH", "MAX_BULK_OPERATIONS", "MAX_N_RESULTS")},
           "name_pattern": ref.METRIC_NAME_PATTERN.pattern, "init": None, "index": [], "search": [], "exists": []}
    rec = Recorder()
    ref.MetricsSemanticMetadataStore(rec)
    out["init"] = {"default": rec.calls}
    rec = Recorder()
    ref.MetricsSemanticMetadataStore(rec, collection_name="custom")
    out["init"]["custom"] = rec.calls

    for name, ns, md in INDEX_CASES:
        rec = Recorder()
        store = ref.MetricsSemanticMetadataStore(rec)
        rec.calls.clear()
        res = outcome(lambda: store.index_metadata(ns, md))
        if res.get("raises") in ("NameError", "UnboundLocalError"):
            # latent bug in the reference's except-block (it logs `document_id` before it is bound);
            # record the exception it was trying to re-raise instead
            res = {"raises": "ReferenceLatentNameError", "message": res["message"]}
        out["index"].append({"case": name, "namespace": ns, "metadata": md, "calls": rec.calls, **res})

    for name, query, n, canned in SEARCH_CASES:
        rec = Recorder()
        store = ref.MetricsSemanticMetadataStore(rec)
        rec.calls.clear()
        rec.canned_query = canned
        res = outcome(lambda: store.search_metadata(query, n_results=n))
        out["search"].append({"case": name, "query": query, "n_results": n, "canned": canned, "calls": rec.calls, **res})

    for name, ns, metric, canned, boom in EXISTS_CASES:
        rec = Recorder()
        store = ref.MetricsSemanticMetadataStore(rec)
        rec.calls.clear()
        rec.canned_get = canned
        rec.raise_on_get = boom
        res = outcome(lambda: store.metric_exists(ns, metric))
        out["exists"].append({"case": name, "namespace": ns, "metric_name": metric, "canned": canned,
                              "raise_on_get": boom, "calls": rec.calls, **res})

    path = os.path.join(HERE, "store_calls.json")
    if "--check" in sys.argv:  # is the committed fixture what the reference produces today?
        fresh = json.loads(json.dumps(out, ensure_ascii=False, sort_keys=True))
        same = fresh == json.load(open(path))
        print("fixture up to date" if same else "FIXTURE STALE")
        sys.exit(0 if same else 1)
    if "--fuzz" in sys.argv:
        sys.exit(fuzz(ref))
    with open(path, "w") as f:
        json.dump(out, f, indent=1, ensure_ascii=False, sort_keys=True)
    print(f"wrote {path}: {len(out['index'])} index, {len(out['search'])} search, {len(out['exists'])} exists cases")


def fuzz(ref):
    """Random inputs through the reference class and the mirror side by side."""
    import random

    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from codd_query_engine_b200 import semantic_store as ss

    rnd = random.Random(0)
    words = ["cpu", "latency ", " p99", "\tdisk", "io\x00", "näme", "—", "Rate", "  ", "of", "5xx", "\n"]
    bad = 0
    for i in range(500):
        md = {}
        if rnd.random() < 0.95:
            md["metric_name"] = rnd.choice(["a.b", "x_y/z", "m-1", "bad name", "", "ü.ñ", "x" * rnd.choice([10, 255, 256]), 7])
        for f in ss.STORED_FIELDS:
            r = rnd.random()
            if r < 0.5:
                md[f] = "".join(rnd.choice(words) for _ in range(rnd.randint(0, 6))) * rnd.choice([1, 1, 1, 400])
            elif r < 0.6:
                md[f] = rnd.choice([None, 0, 1.5, False])
        a, b = Recorder(), Recorder()
        ra = outcome(lambda: ref.MetricsSemanticMetadataStore(a).index_metadata("ns:%d" % i, md))
        rb = outcome(lambda: ss.MetricsSemanticMetadataStore(b).index_metadata("ns:%d" % i, md))
        if ra.get("raises") in ("NameError", "UnboundLocalError"):
            ok = "raises" in rb
        else:
            ok = ra == rb and a.calls == b.calls
        q = "".join(rnd.choice(words) for _ in range(rnd.randint(0, 5))) * rnd.choice([1, 1, 300])
        n = rnd.choice([-1, 0, 1, 5, 100, 101])
        a, b = Recorder(), Recorder()
        canned = {"ids": [["n#a", "b"]], "distances": [[0.125, 0.5]], "metadatas": [[{"x": "y"}, {}]]}
        a.canned_query = b.canned_query = canned
        sa = outcome(lambda: ref.MetricsSemanticMetadataStore(a).search_metadata(q, n_results=n))
        sb = outcome(lambda: ss.MetricsSemanticMetadataStore(b).search_metadata(q, n_results=n))
        ok = ok and sa == sb and a.calls == b.calls
        if not ok:
            bad += 1
            print("MISMATCH", md, ra, rb, sa, sb)
    print(f"fuzz: {500 - bad}/500 identical")
    return 1 if bad else 0


if __name__ == "__main__":
    main()
