"""bench.py's CPU-side legs (cpu_baseline / `--impl reference`) run here without a GPU: the device index is
replaced by a stub that hands out reference-format records of a graph built by the CPU checker, which is all
those legs ever ask of it (`export_records`).  Guards the JSON contract of the reference arm and the parity
bookkeeping of the cpu_baseline leg."""
import io
import json
import types

import numpy as np
import pytest


class _StubIndex:
    """Stands in for pg_embedding_b200.HnswIndex: only export_records() is used by the CPU legs."""

    def __init__(self, orc):
        self.rec = orc.records()

    def export_records(self, first=0, n=None):
        n = self.rec.shape[0] - first if n is None else n
        return self.rec[first:first + n]


class _HostTensor:
    def __init__(self, a):
        self.a = a

    def cpu(self):
        return self

    def numpy(self):
        return self.a


@pytest.fixture()
def small_bench(oracle_mod, monkeypatch):
    import bench
    monkeypatch.setattr(bench, "DIMS", 16)
    monkeypatch.setattr(bench, "M", 4)
    monkeypatch.setattr(bench, "EFC", 20)
    monkeypatch.setattr(bench, "EFS", 8)
    monkeypatch.setattr(bench, "METRIC", "cosine")
    rng = np.random.default_rng(5)
    n = 600
    x = rng.standard_normal((n, 16)).astype(np.float32) + 1.0
    q = rng.standard_normal((300, 16)).astype(np.float32) + 1.0
    which, _ = bench.pick_checker()
    orc = oracle_mod.FlatIndex(which, 16, 4, 20, 8, "cosine", capacity=n)
    orc.build(x)
    want = orc.search_many(q, 8, nthreads=2)
    return bench, _StubIndex(orc), n, q, want


def test_cpu_leg_parity_and_sample(small_bench):
    bench, idx, n, q, want = small_bench
    args = types.SimpleNamespace(cpu_seconds=0.2)
    labels = want["labels"].view(np.int64)          # what bench.py hands over: the GPU's int64 label tensor
    base, par = bench.cpu_leg(args, idx, _HostTensor(q), labels, want["n"], n)
    assert par == {"queries": q.shape[0], "labels_identical_to_cpu_reference": True}
    assert base["unit"] == "queries/s" and base["value"] > 0 and base["cores"] >= 1
    assert base["kind"] in ("reference", "port") and "median of 3" in base["sample"]
    # a single wrong label must be noticed
    bad = labels.copy()
    bad[7, 0] ^= 1
    _, par2 = bench.cpu_leg(args, idx, _HostTensor(q), bad, want["n"], n)
    assert par2["labels_identical_to_cpu_reference"] is False


def test_reference_arm_json_contract(small_bench, monkeypatch):
    bench, idx, n, q, want = small_bench
    buf = io.StringIO()
    monkeypatch.setattr(bench, "JSON_OUT", buf, raising=False)
    args = types.SimpleNamespace(gpus=1)
    assert bench.reference_arm(args, None, None, idx, None, _HostTensor(q), n, 2, 1) == 0
    lines = [l for l in buf.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert "workload" in d["config"] and "model" not in d["config"]
