"""The C-ABI boundary without a GPU: the library loads, exports every symbol the header declares,
the host-only helpers agree with the oracle, and compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from oracle import brute_force as bf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "coddvec.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cv_[a-z_]+)\s*\(", text)))


def test_header_declares_the_documented_surface():
    fns = declared_functions()
    for must in ("cv_create", "cv_destroy", "cv_upsert", "cv_query", "cv_query_device", "cv_merge_keys_device",
                 "cv_count", "cv_stats", "cv_last_error", "cv_get_rows"):
        assert must in fns


def test_library_exports_every_declared_symbol(native_lib):
    from codd_query_engine_b200 import native

    out = subprocess.run(["nm", "-D", "--defined-only", str(native.library_path())], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (cv_[a-z_]+)", out))
    declared = set(declared_functions())
    assert declared <= exported, declared - exported
    assert set(native.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert hasattr(native_lib, name)


def test_header_is_plain_c():
    """No torch / C++ types in the signatures: the header must compile as C."""
    src = '#include "coddvec.h"\nint main(void){cv_config c; cv_query_params p; cv_stats_t s; (void)c;(void)p;(void)s; return CV_MAX_K==128?0:1;}\n'
    exe = "/tmp/coddvec_header_check"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe],
                   input=src, text=True, check=True, env={**os.environ, "CC": "gcc"})
    assert subprocess.run([exe]).returncode == 0


def test_struct_layouts_match_ctypes(native_lib):
    from codd_query_engine_b200 import native

    src = ('#include <stdio.h>\n#include "coddvec.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(cv_config), '
           'sizeof(cv_query_params), sizeof(cv_stats_t)); return 0;}\n')
    exe = "/tmp/coddvec_sizeof"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe], input=src, text=True, check=True)
    sizes = list(map(int, subprocess.run([exe], capture_output=True, text=True).stdout.split()))
    assert sizes == [ctypes.sizeof(native._Config), ctypes.sizeof(native._QueryParams), ctypes.sizeof(native.Stats)]


def test_key_helpers_match_oracle(native_lib):
    from codd_query_engine_b200 import native

    rng = np.random.default_rng(0)
    sc = np.concatenate([rng.standard_normal(1000).astype(np.float32),
                         np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1.0, -1.0], np.float32)])
    rows = rng.integers(0, 2**32 - 1, len(sc), dtype=np.uint64)
    want = bf.pack_keys(sc, rows)
    got = np.array([native.encode_key(float(s), int(r)) for s, r in zip(sc, rows)], np.uint64)
    assert np.array_equal(got, want)
    ds, dr = native.decode_keys(np.concatenate([want, np.zeros(2, np.uint64)]))
    os_, or_ = bf.unpack_keys(np.concatenate([want, np.zeros(2, np.uint64)]))
    assert np.array_equal(dr, or_) and np.array_equal(ds.view(np.uint32), os_.view(np.uint32))


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback(native_lib):
    from codd_query_engine_b200 import native
    from codd_query_engine_b200.client import Client

    with pytest.raises(native.NativeError) as e:
        native.NativeIndex(384)
    assert "no CPU fallback" in str(e.value)
    coll = Client("abi-test").get_or_create_collection("c")
    with pytest.raises(native.NativeError):
        coll.upsert(ids=["a"], embeddings=[[1.0] * 384])
    assert coll.count() == 0


def test_missing_library_fails_loudly(monkeypatch):
    from codd_query_engine_b200 import native

    monkeypatch.setattr(native, "_lib", None)
    monkeypatch.setenv("CODDVEC_LIB", "/nonexistent/libcoddvec.so")
    with pytest.raises(native.NativeError) as e:
        native.load_library()
    assert "no CPU fallback" in str(e.value)
