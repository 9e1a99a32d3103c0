"""INTEGRATION.md's raw-ctypes binding example must actually run on a B200."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_integration_md_ctypes_example_runs(native_lib):
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    example = next(b for b in blocks if "cv_create" in b)
    example += "\nassert rc == 0 and out_rows[0, 0] == 0 and abs(distances[0, 0]) < 1e-3 and counts[0] == 10\nprint('doc example ok')\n"
    out = subprocess.run([sys.executable, "-c", example], capture_output=True, text=True, cwd=ROOT)
    assert out.returncode == 0 and "doc example ok" in out.stdout, out.stderr[-2000:]
