"""Drop-in proof: the reference's own MetricsSemanticMetadataStore class and its own test files run
unmodified on top of this package (`shim/chromadb`).  Needs /root/reference, so build-container only."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_exposes_the_chromadb_names_the_reference_uses():
    code = ("import sys; sys.path[:0]=[%r, %r]; import chromadb; "
            "assert all(hasattr(chromadb, n) for n in ('HttpClient','EphemeralClient','PersistentClient','Client','ClientAPI')); "
            "c = chromadb.HttpClient(host='localhost', port=8000); assert isinstance(c.heartbeat(), int); print('ok')"
            ) % (os.path.join(ROOT, "shim"), ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir("/root/reference/codd_dal"), reason="reference tree not present")
def test_reference_store_and_its_tests_run_on_the_dropin():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "run_reference_tests_dropin.py")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "reference tests passed against the drop-in: 9" in out.stdout
    assert "PASS codd_lib MetricsPromQLClient.search_relevant_metrics" in out.stdout
    assert "PASS codd_jobs MetricsSemanticIndexerJob._process_batch" in out.stdout
