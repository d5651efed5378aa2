"""Golden fixtures produced by the UNMODIFIED compiled reference (tests/golden/gen_ref_fixtures.py, committed with its
output): the C restatement must reproduce them bit for bit WITHOUT the reference tree; on a GPU so must the CUDA path."""
import os

import numpy as np
import pytest

FIX = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures.npz"))
NAMES = sorted({k.split(".")[0] for k in FIX.files})
METRIC_NAME = {0: "l2", 1: "cosine", 2: "manhattan"}


def _case(name):
    dims, m, efc, n, metric = (int(v) for v in FIX[f"{name}.params"])
    return dims, m, efc, n, METRIC_NAME[metric]


@pytest.mark.parametrize("name", NAMES)
def test_port_reproduces_reference_fixture(oracle_mod, name):
    dims, m, efc, n, metric = _case(name)
    x, q, labels = FIX[f"{name}.x"], FIX[f"{name}.q"], FIX[f"{name}.labels"]
    clean = labels & ~(np.uint64(1) << np.uint64(48))
    idx = oracle_mod.FlatIndex("port", dims, m, efc, 16, metric, capacity=n)
    idx.build(x, clean)
    assert idx.links().tobytes() == FIX[f"{name}.links"].tobytes()
    for i in np.flatnonzero(labels != clean):
        idx.mark_deleted(int(i))
    for ef in (16, 5):
        r = idx.search_many(q, ef)
        assert r["labels"].tobytes() == FIX[f"{name}.search{ef}"].tobytes()
        assert r["n"].tolist() == FIX[f"{name}.n{ef}"].tolist()
    assert oracle_mod.dist_many("port", metric, q[0], x).view(np.uint32).tolist() == FIX[f"{name}.dist_bits"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_reference_fixture(name):
    import pg_embedding_b200 as pg
    dims, m, efc, n, metric = _case(name)
    x, q, labels = FIX[f"{name}.x"], FIX[f"{name}.q"], FIX[f"{name}.labels"]
    clean = labels & ~(np.uint64(1) << np.uint64(48))
    idx = pg.HnswIndex(dims, m, efc, 16, metric, capacity=n)
    idx.insert_many(x, clean)                                           # exact sequential binds on the device
    assert idx.links().tobytes() == FIX[f"{name}.links"].tobytes()
    idx.mark_deleted(np.flatnonzero(labels != clean))
    for ef in (16, 5):
        out = idx.search_batch(q, ef)
        assert out["labels"].tobytes() == FIX[f"{name}.search{ef}"].tobytes()
        assert out["n"].tolist() == FIX[f"{name}.n{ef}"].tolist()
    assert pg.dist_batch(metric, q[0], x).view(np.uint32).tolist() == FIX[f"{name}.dist_bits"].tolist()
    idx2 = pg.HnswIndex(dims, m, efc, 16, metric, capacity=n)
    idx2.append(x, clean)
    idx2.build_exact(0, n, 64)                                          # exact parallel build too
    assert idx2.links().tobytes() == FIX[f"{name}.links"].tobytes()
    idx.close(); idx2.close()
