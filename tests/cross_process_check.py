"""Two real processes, one snapshot directory: the unmodified-caller flow of INTEGRATION.md "Two processes, one server".

    python tests/cross_process_check.py            # needs a B200 (both processes use the native library)

The parent plays the service: it builds `chromadb.HttpClient(host, port)` through the shim, searches, then starts a
child that plays the indexer job -- the same `HttpClient(...)` call, `upsert`, and a plain exit (no persist() call) --
and searches again with the SAME store / collection objects.
"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shim"))
sys.path.insert(0, ROOT)

JOB = r'''
import sys, numpy as np
sys.path.insert(0, sys.argv[1] + "/shim"); sys.path.insert(0, sys.argv[1])
import chromadb
from codd_query_engine_b200.semantic_store import MetricsSemanticMetadataStore
store = MetricsSemanticMetadataStore(chromadb.HttpClient(host="chroma", port=8000), collection_name="metrics")
for i in range(int(sys.argv[2]), int(sys.argv[3])):
    store.index_metadata("prod", {"metric_name": f"svc.metric_{i}", "description": f"request latency of service number {i}",
                                  "category": "application"})
print("job indexed", sys.argv[2], sys.argv[3])
'''


def main():
    with tempfile.TemporaryDirectory() as vol:
        env = dict(os.environ, CODDVEC_SNAPSHOT_DIR=vol, CODDVEC_SNAPSHOT_RELOAD_S="0.05")
        os.environ.update(CODDVEC_SNAPSHOT_DIR=vol, CODDVEC_SNAPSHOT_RELOAD_S="0.05")
        subprocess.run([sys.executable, "-c", JOB, ROOT, "0", "50"], check=True, env=env)
        import chromadb
        from codd_query_engine_b200.semantic_store import MetricsSemanticMetadataStore

        store = MetricsSemanticMetadataStore(chromadb.HttpClient(host="chroma", port=8000), collection_name="metrics")
        coll = store.collection
        assert coll.count() == 50, coll.count()
        assert not store.metric_exists("prod", "svc.metric_77")
        subprocess.run([sys.executable, "-c", JOB, ROOT, "50", "100"], check=True, env=env)  # the job runs again, exits
        time.sleep(0.1)
        hits = store.search_metadata("request latency of service number 77", n_results=3)
        assert store.collection is coll and coll.count() == 100, coll.count()
        assert store.metric_exists("prod", "svc.metric_77")
        assert any(h["metric_name"] == "svc.metric_77" for h in hits), hits
        print("CROSS_PROCESS_CHECK OK: the service's collection object went from 50 to 100 rows without a restart;",
              "top hit", hits[0]["metric_name"])


if __name__ == "__main__":
    main()
