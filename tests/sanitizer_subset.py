"""Reduced pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck).

Small shapes (the tools slow kernels down 10-100x) but every code path: K1 append + overwrite, K2 with
1 / 2 / 4 queries per pass, K2 predicate (namespace, label, bitmask), K2 wide rows, K3 single CTAs and CTA
pairs (k in registers and k > 64), the K3 hi/lo query split, K4 merge.  Results are checked against the
oracle so that a tool-induced slowdown cannot hide a wrong answer.

    compute-sanitizer --tool memcheck python tests/sanitizer_subset.py [family ...]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from codd_query_engine_b200 import native  # noqa: E402
from oracle import brute_force as bf  # noqa: E402  (checker only)
from tests.parity import check_topk, oracle_full_scores  # noqa: E402


def build(n, D, seed=0, ns=False):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, D)).astype(np.float32)
    codes = rng.integers(1, 5, n).astype(np.uint32) if ns else None
    labels = rng.integers(1, 3, n).astype(np.uint32) if ns else None
    idx = native.NativeIndex(D)
    idx.upsert(x, ns_codes=codes, label_codes=labels)
    stored, scale, _, _ = idx.get_rows(np.arange(n))
    return idx, stored, scale, codes, labels, rng


def k1():
    idx, stored, scale, _, _, rng = build(3000, 100, 1)
    est, esc = bf.normalise_store(np.random.default_rng(1).standard_normal((3000, 100)).astype(np.float32))
    assert np.array_equal(stored, est) and np.array_equal(scale, esc)
    y = rng.standard_normal((7, 100)).astype(np.float32)
    idx.upsert(y, rows=np.array([0, 5, 17, 2999, 100, 200, 300]))
    got, _, _, _ = idx.get_rows(np.array([5, 2999]))
    assert np.array_equal(got, bf.normalise_store(y[[1, 3]])[0])
    idx.close()


def k2():
    for D, B, k in ((384, 1, 10), (384, 2, 10), (384, 4, 10), (768, 2, 50), (64, 3, 5), (1000, 1, 100)):
        idx, stored, scale, _, _, rng = build(6001, D, D + B)
        q = rng.standard_normal((B, D)).astype(np.float32)
        s, r, c = idx.query(q, k, path=native.PATH_SCAN)
        check_topk(s, r, c, oracle_full_scores(stored, scale, q), None, k, tol=2e-6)
        idx.close()


def k2_pred():
    idx, stored, scale, ns, lb, rng = build(9001, 384, 7, ns=True)
    q = rng.standard_normal((2, 384)).astype(np.float32)
    full = oracle_full_scores(stored, scale, q)
    s, r, c = idx.query(q, 10, ns_code=2, path=native.PATH_SCAN)
    check_topk(s, r, c, full, ns == 2, 10, tol=2e-6)
    s, r, c = idx.query(q[:1], 20, ns_code=3, label_code=1, path=native.PATH_SCAN)
    check_topk(s, r, c, full[:1], (ns == 3) & (lb == 1), 20, tol=2e-6)
    elig = rng.random(9001) < 0.05
    words = np.packbits(np.pad(elig, (0, (-len(elig)) % 32)), bitorder="little").view(np.uint32)
    s, r, c = idx.query(q, 10, mask=words, path=native.PATH_SCAN)
    check_topk(s, r, c, full, elig, 10, tol=2e-6)
    idx.close()


def k2_wide():
    idx, stored, scale, _, _, rng = build(1501, 1536, 9)
    q = rng.standard_normal((1, 1536)).astype(np.float32)
    s, r, c = idx.query(q, 10, path=native.PATH_SCAN)
    check_topk(s, r, c, oracle_full_scores(stored, scale, q), None, 10, tol=2e-6)
    idx.close()


def k3():
    for D, B, k, path in ((384, 16, 10, native.PATH_MMA_1CTA), (384, 130, 10, native.PATH_MMA_2CTA),
                          (768, 5, 50, native.PATH_MMA_1CTA), (64, 9, 10, native.PATH_MMA), (384, 3, 100, native.PATH_MMA_1CTA)):
        idx, stored, scale, ns, _, rng = build(5003, D, D + B, ns=True)
        q = rng.standard_normal((B, D)).astype(np.float32)
        full = oracle_full_scores(stored, scale, q)
        s, r, c = idx.query(q, k, path=path)
        check_topk(s, r, c, full, None, k, tol=1e-3)
        s, r, c = idx.query(q, k, ns_code=1, path=path)
        check_topk(s, r, c, full, ns == 1, k, tol=1e-3)
        idx.close()


def k4():
    import torch

    G, B, k = 5, 7, 10
    rng = np.random.default_rng(3)
    keys = np.sort(rng.integers(1, 2 ** 62, (G, B, k), dtype=np.uint64), axis=2)[:, :, ::-1].copy()
    d_in = torch.from_numpy(keys.view(np.int64)).cuda()
    d_out = torch.zeros((B, k), dtype=torch.int64, device="cuda")
    native.merge_keys_device(0, d_in.data_ptr(), G, B, k, k, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = np.sort(keys.transpose(1, 0, 2).reshape(B, G * k), axis=1)[:, ::-1][:, :k]
    assert np.array_equal(d_out.cpu().numpy().view(np.uint64), want)


FAMILIES = {"k1": k1, "k2": k2, "k2_pred": k2_pred, "k2_wide": k2_wide, "k3": k3, "k4": k4}

if __name__ == "__main__":
    names = sys.argv[1:] or list(FAMILIES)
    for name in names:
        FAMILIES[name]()
        print("ok", name, flush=True)
    print("sanitizer subset ok:", " ".join(names))
