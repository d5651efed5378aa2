"""GPU parity tests of the OPT-IN PROTOTYPES: code that has been checked on the host emulator
against the oracle but has not run on a B200 yet.  It is compiled only into libpgemb_b200_proto.so (-DPGEMB_PROTO,
pg_embedding_b200/build.py); the product library does not contain it.  Run with

    PGEMB_LIB_VARIANT=proto python -m pytest tests/test_gpu_prototypes.py -m gpu

(tools/gpu_r2_first.sh does, followed by an A/B of each flag).  Under the product library these tests skip: the default
`pytest -m gpu` run exercises exactly the kernels that were measured.  tests/test_capi_emulated.py reuses the bodies
below on the emulated prototype build in the CPU suite."""
import numpy as np
import pytest

import test_gpu_parity as G
from test_gpu_parity import SEARCH_CFGS, BIND_CFGS, METRICS, _data, kernel_mode

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pg():
    import pg_embedding_b200 as pg
    from pg_embedding_b200 import _lib, build
    build.build()
    if pg.device_count() < 1:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    if b"+proto" not in _lib.load().pgemb_version():
        pytest.skip("prototype kernels live in libpgemb_b200_proto.so: set PGEMB_LIB_VARIANT=proto")
    return pg


@pytest.mark.parametrize("flags", [{"PGEMB_VISITED_PAIRS": "1"}, {"PGEMB_SMEM_VISITED": "4096"}, {"PGEMB_VISITED_PAIRS": "1", "PGEMB_SMEM_VISITED": "1024"}],
                         ids=["pairs", "smem-visited", "both"])
@pytest.mark.parametrize("cfg", [SEARCH_CFGS[0], SEARCH_CFGS[5], SEARCH_CFGS[6], SEARCH_CFGS[7]], ids=lambda c: f"d{c[0]}m{c[1]}")
def test_visited_pairs_mode(pg, oracle_mod, cfg, flags, monkeypatch):
    """PGEMB_VISITED_PAIRS=1: both 32-id halves of a link list are test-and-set concurrently.  Same results, same
    traversal counters; a graph whose lists repeat an id must be detected and served by the ordered path."""
    for k, v in flags.items():
        monkeypatch.setenv(k, v)
    dims, m, efc, n, kw, efs = cfg
    rng = np.random.default_rng(4242 + dims)
    x = _data(rng, n, dims, **kw)
    q = _data(rng, 200, dims, **{k: v for k, v in kw.items() if k == "levels"})
    orc = oracle_mod.FlatIndex("port", dims, m, efc, 64, "l2", capacity=n)
    orc.build(x)
    links = orc.links()
    for dup in (False, True):
        if dup:
            # repeat an id inside the lists of a few well-connected nodes (positions in different 32-id halves when possible)
            for node in np.argsort(-links[:, 0].astype(np.int64))[:5]:
                c = int(links[node, 0])
                if c >= 2:
                    links[node, c] = links[node, 1]
            orc.set_links(links)
        idx = pg.HnswIndex(dims, m, efc, 64, "l2", capacity=n)
        idx.append(x, None, links)
        for coop, nq in (("1", 64), ("0", 200)):
            with kernel_mode(coop):
                out = idx.search_batch(q[:nq], efs[-1], want_stats=True)
            want = orc.search_many(q[:nq], efs[-1], want_counters=True)
            assert out["labels"].tobytes() == want["labels"].tobytes(), (dup, coop)
            assert out["stats"][:, :3].tolist() == want["counters"][:, :3].tolist(), (dup, coop)
        idx.close()


def test_l2_eight_lanes_per_row(pg, oracle_mod, monkeypatch):
    """PGEMB_L2_TPR8=1 (prototype): long L2 rows scored by 8 lanes per row, rings of 4 rows."""
    monkeypatch.setenv("PGEMB_L2_TPR8", "1")
    monkeypatch.setenv("PGEMB_L2_TPR8_MIN_BYTES", "0")
    for cfg in (SEARCH_CFGS[3], SEARCH_CFGS[4], SEARCH_CFGS[7]):
        G.test_search_identical_to_oracle(pg, oracle_mod, "l2", cfg)
    G.test_bind_links_identical_to_oracle(pg, oracle_mod, "l2", BIND_CFGS[4])


@pytest.mark.parametrize("metric", METRICS)
def test_scan_topk_tiled(pg, oracle_mod, metric, monkeypatch):
    """PGEMB_SCAN_TILED=1: the exact scan's distance step through scan_tile_kernel (rows staged once per query tile)."""
    monkeypatch.setenv("PGEMB_SCAN_TILED", "1")
    G.test_scan_topk_matches_exact_order(pg, oracle_mod, metric)
    G.test_scan_topk_regress_seqscan(pg)


@pytest.mark.parametrize("metric", ["l2", "cosine"])
def test_scan_topk_tensor_core_filter(pg, oracle_mod, metric, monkeypatch):
    """PGEMB_SCAN_TC=1: one TF32 GEMM per chunk (cuBLAS, tensor cores) discards rows, survivors are re-scored with the
    reference-exact arithmetic -> the exact scan's labels, order and bits; the error-bound tripwire must stay silent."""
    import ctypes as C
    from pg_embedding_b200 import _lib

    def counters():
        out = (C.c_uint64 * 4)()
        _lib.load().pgemb_proto_counters(out)
        return dict(scans=out[0], fallbacks=out[1], rescored=out[2], pairs=out[3])

    rng = np.random.default_rng(23)
    for dims, n, k, nq in ((33, 3000, 64, 40), (128, 30000, 10, 70), (768, 40000, 10, 130)):
        x = _data(rng, n, dims)
        q = _data(rng, nq, dims)
        x[n // 2] = x[n // 3]
        labels = rng.permutation(n).astype(np.uint64) + np.uint64(3)
        labels[::11] |= np.uint64(1 << 48)
        idx = pg.HnswIndex(dims, 4, 8, 16, metric, capacity=n)
        idx.append(x, labels)
        monkeypatch.setenv("PGEMB_SCAN_TC", "0")
        want = idx.scan_topk(q, k)
        c0 = counters()
        monkeypatch.setenv("PGEMB_SCAN_TC", "1")
        got = idx.scan_topk(q, k)
        c1 = counters()
        assert got["labels"].tobytes() == want["labels"].tobytes(), (metric, dims)
        assert got["dists"].tobytes() == want["dists"].tobytes() and got["n"].tolist() == want["n"].tolist()
        assert c1["scans"] == c0["scans"] + 1 and c1["fallbacks"] == c0["fallbacks"], "the TF32 error bound was exceeded"
        frac = (c1["rescored"] - c0["rescored"]) / (c1["pairs"] - c0["pairs"])
        print(f"tc filter {metric} dims={dims} n={n} k={k}: {frac:.4f} of the pairs re-scored exactly")
        if n >= 30000:
            assert frac < 0.2
        idx.close()


def test_rows_gathered_with_cp_async_pieces(pg, oracle_mod, monkeypatch):
    """PGEMB_GATHER_LDGSTS=1: LDGSTS row gather (a warp instruction per 512 B) instead of one bulk copy per row."""
    monkeypatch.setenv("PGEMB_GATHER_LDGSTS", "1")
    for metric, cfg in (("l2", SEARCH_CFGS[4]), ("cosine", SEARCH_CFGS[7]), ("manhattan", SEARCH_CFGS[3]), ("l2", SEARCH_CFGS[0])):
        G.test_search_identical_to_oracle(pg, oracle_mod, metric, cfg)
    G.test_bind_links_identical_to_oracle(pg, oracle_mod, "cosine", BIND_CFGS[1])


def test_fast_small_batches_and_exact_build_clamp(pg, oracle_mod, monkeypatch):
    """PGEMB_FAST_SMALL=1 (single-stream host path for <= 64 queries, cached launch configuration) and
    PGEMB_EXACT_CLAMP_SMS=1 (exact parallel build keeps its batches at one search per SM): same results."""
    monkeypatch.setenv("PGEMB_FAST_SMALL", "1")
    for case in G.GOLD:
        G.test_kat_regress(pg, case)
    rng = np.random.default_rng(77)
    n, dims, m, efc = 4000, 48, 8, 40
    x = _data(rng, n, dims)
    q = _data(rng, 100, dims)
    orc = oracle_mod.FlatIndex("port", dims, m, efc, 64, "l2", capacity=n)
    orc.build(x)
    idx = pg.HnswIndex(dims, m, efc, 64, "l2", capacity=n)
    idx.append(x, orc.labels(), orc.links())
    for nq, ef in ((1, 10), (1, 10), (64, 16), (100, 16), (3, 40), (1, 64)):
        out = idx.search_batch(q[:nq], ef, want_stats=True)
        want = orc.search_many(q[:nq], ef, want_counters=True)
        assert out["labels"].tobytes() == want["labels"].tobytes() and out["stats"][:, :3].tolist() == want["counters"][:, :3].tolist(), (nq, ef)
    assert idx.search(q[0], 12).tolist() == orc.search(q[0], 12).tolist()
    idx.close()
    monkeypatch.setenv("PGEMB_EXACT_CLAMP_SMS", "1")
    idx = pg.HnswIndex(dims, m, efc, 64, "l2", capacity=n)
    idx.append(x)
    idx.build_exact(0, n, 1024)
    assert idx.links().tobytes() == orc.links().tobytes()
    idx.close()

