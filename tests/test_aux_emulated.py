"""The kernels around the traversal, executed on the host by the SIMT emulator (tests/emu):
  * dist_pairs_kernel -- what hnsw_dist_func / the SQL distance operators run: bit-exact against
    BOTH checkers, i.e. also against the compiled reference's distfunc.c, for every dimension class of the vectorised loops;
  * norms_kernel -- the cached squared norms in cosine lane order;
  * scan_dist_kernel + scan_select_kernel -- the exact scan (8 f3) with its chunked running top-k and (dist,label) order;
  * merge_topk_kernel -- the shard merge (8 e)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from test_search_emulated import sqnorm_lane_order

pytestmark = pytest.mark.timeout(900, method="thread")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MID = {"l2": 0, "cosine": 1, "manhattan": 2}


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libaux_emu.so")
    src = [os.path.join(ROOT, "tests", "emu", f) for f in ("aux_emu.cpp", "emu_runtime.cpp")]
    res = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "tests", "emu", "fake_cuda"),
                          "-o", out] + src, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return C.CDLL(out)


@pytest.mark.parametrize("metric", ["l2", "cosine", "manhattan"])
def test_distance_kernel_bits_vs_both_checkers(emu, oracle_mod, metric):
    rng = np.random.default_rng(2)
    checkers = ["port"] + (["ref"] if oracle_mod.available("ref") else [])
    for dim in list(range(1, 36)) + [63, 64, 65, 100, 127, 128, 129, 300, 768, 769]:
        n = 9
        a = (rng.standard_normal((n, dim)) * rng.choice([1e-3, 1.0, 50.0], size=(n, 1))).astype(np.float32)
        b = rng.standard_normal((n, dim)).astype(np.float32)
        out = np.zeros(n, np.float32)
        emu.emu_dist_pairs(MID[metric], _p(a, C.c_float), _p(b, C.c_float), C.c_uint32(dim), C.c_uint32(n), 0, _p(out, C.c_float))
        for which in checkers:
            want = oracle_mod.dist_many(which, metric, a, b)
            assert out.tobytes() == want.tobytes(), (metric, dim, which)
        out1 = np.zeros(n, np.float32)   # one query against many rows (broadcast)
        emu.emu_dist_pairs(MID[metric], _p(a, C.c_float), _p(b, C.c_float), C.c_uint32(dim), C.c_uint32(n), 1, _p(out1, C.c_float))
        assert out1.tobytes() == oracle_mod.dist_many("port", metric, a[0], b).tobytes(), (metric, dim)


def test_norms_kernel_lane_order(emu):
    rng = np.random.default_rng(4)
    for dim in (1, 3, 4, 7, 33, 128, 203):
        n, row_f = 13, (dim + 3) & ~3
        x = np.zeros((n, row_f), np.float32)
        x[:, :dim] = rng.standard_normal((n, dim)).astype(np.float32) * 3
        norms = np.full(n, -1, np.float32)
        emu.emu_norms(_p(x, C.c_float), C.c_uint32(row_f), C.c_uint32(dim), C.c_uint32(2), C.c_uint32(n - 2), _p(norms, C.c_float))
        assert norms[0] == -1 and norms[1] == -1
        want = np.array([sqnorm_lane_order(x[i, :dim]) for i in range(2, n)], np.float32)
        assert norms[2:].tobytes() == want.tobytes(), dim


@pytest.mark.parametrize("metric", ["l2", "cosine", "manhattan"])
def test_exact_scan_kernels(emu, oracle_mod, metric):
    rng = np.random.default_rng(6)
    for dim, n, k, chunk, levels in ((3, 90, 7, 32, 3), (17, 300, 40, 128, 0), (8, 50, 64, 16, 2)):
        x = rng.integers(0, levels, (n, dim)).astype(np.float32) if levels else rng.standard_normal((n, dim)).astype(np.float32)
        q = rng.integers(0, levels, (5, dim)).astype(np.float32) if levels else rng.standard_normal((5, dim)).astype(np.float32)
        if metric == "cosine":
            x, q = x + 1.0, q + 1.0
        labels = rng.permutation(n).astype(np.uint64) + np.uint64(10)
        dead = np.arange(0, n, 6)
        labels[dead] |= np.uint64(1) << np.uint64(48)
        row_f = (dim + 3) & ~3
        xv = np.zeros((n, row_f), np.float32); xv[:, :dim] = x
        norms = np.array([sqnorm_lane_order(x[i]) for i in range(n)], np.float32)
        nq = q.shape[0]
        td = np.zeros((nq, k), np.uint32); tl = np.zeros((nq, k), np.uint64); tn = np.zeros(nq, np.uint32)
        emu.emu_scan_topk(MID[metric], _p(xv, C.c_float), _p(norms, C.c_float), _p(labels, C.c_uint64), C.c_uint32(row_f), C.c_uint32(dim), C.c_uint32(n),
                          _p(np.ascontiguousarray(q), C.c_float), C.c_uint32(nq), C.c_uint32(k), C.c_uint32(chunk), _p(td, C.c_uint32), _p(tl, C.c_uint64), _p(tn, C.c_uint32))
        alive = np.ones(n, bool); alive[dead] = False
        for i in range(nq):
            d = oracle_mod.dist_many("port", metric, q[i], x)
            order = sorted((float(d[j]), int(labels[j])) for j in range(n) if alive[j])[:k]
            assert int(tn[i]) == len(order)
            assert tl[i, :len(order)].tolist() == [o[1] for o in order], (metric, dim, i)


def test_merge_kernel(emu):
    rng = np.random.default_rng(8)
    nq, S, k = 9, 3, 6
    din = np.zeros((S, nq, k), np.float32); lin = np.zeros((S, nq, k), np.uint64); nin = np.zeros((S, nq), np.int32)
    for s in range(S):
        for q in range(nq):
            c = int(rng.integers(0, k + 1))
            pairs = sorted((float(rng.integers(0, 4)), int(rng.integers(0, 50))) for _ in range(c))   # ties on purpose
            nin[s, q] = c
            for i, (d, l) in enumerate(pairs):
                din[s, q, i], lin[s, q, i] = d, l
    dout = np.zeros((nq, k), np.float32); lout = np.zeros((nq, k), np.uint64); nout = np.zeros(nq, np.int32)
    emu.emu_merge_topk(C.c_uint32(nq), C.c_uint32(S), C.c_uint32(k), _p(din, C.c_float), _p(lin, C.c_uint64), _p(nin, C.c_int32), _p(dout, C.c_float),
                       _p(lout, C.c_uint64), _p(nout, C.c_int32))
    for q in range(nq):
        allp = sorted((float(din[s, q, i]), int(lin[s, q, i])) for s in range(S) for i in range(nin[s, q]))[:k]
        assert int(nout[q]) == len(allp)
        assert [(float(dout[q, i]), int(lout[q, i])) for i in range(len(allp))] == allp
