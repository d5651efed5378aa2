"""Churn test of pgemb_sidecar + libpgemb_client.so on the host-emulated library (CPU only):

    python tools/sidecar_stress.py [--seconds 60] [--searchers 6]

One writer process inserts points one by one (record + hnsw_bind_point, as hnsw_add_point does) while searcher processes
call hnsw_search in a loop; every few seconds a random searcher is SIGKILLed mid-flight and replaced.  At the end:
the mirror's link lists must equal the oracle's sequential build of the same points (searches and dying clients must not
disturb the writer), every search that returned must have returned only labels that had been inserted, no call may hang, and
after the sidecar's reclaim pass no request slot or bulk lock may be left behind.
"""
import argparse
import json
import os
import signal
import struct
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

DIMS, M, EFC, EFS, METRIC, N = 12, 4, 12, 8, "l2", 260


def points():
    return np.random.default_rng(99).standard_normal((N, DIMS)).astype(np.float32)


def searcher(shm, out, seconds):
    from pg_embedding_b200 import sidecar
    sidecar.connect(shm)
    idx = sidecar.RemoteIndex(1, DIMS, M, EFC, EFS, METRIC, capacity=N)
    rng = np.random.default_rng(os.getpid())
    t_end, n, bad = time.time() + seconds, 0, 0
    while time.time() < t_end:
        res = idx.search(rng.standard_normal(DIMS).astype(np.float32), int(rng.choice([1, 4, 8])))
        n += 1
        if any(not (1000 <= int(l) < 1000 + N) for l in res):
            bad += 1
    json.dump({"calls": n, "bad": bad}, open(out, "w"))


def writer(shm, out):
    from pg_embedding_b200 import sidecar
    sidecar.connect(shm)
    idx = sidecar.RemoteIndex(1, DIMS, M, EFC, EFS, METRIC, capacity=N)
    x = points()
    rs = idx.record_bytes
    for i in range(N):
        rec = np.zeros((1, rs), np.uint8)
        rec[0, (2 * M + 1) * 4:(2 * M + 1) * 4 + DIMS * 4] = np.frombuffer(x[i].tobytes(), np.uint8)
        rec[0, rs - 8:] = np.frombuffer(np.uint64(1000 + i).tobytes(), np.uint8)
        idx.append_records(rec)
        idx.bind_point(i)
        time.sleep(0.01)
    json.dump({"inserted": N}, open(out, "w"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--searchers", type=int, default=6)
    ap.add_argument("--role", default="")
    ap.add_argument("--shm"), ap.add_argument("--out")
    a = ap.parse_args()
    if a.role == "searcher":
        return searcher(a.shm, a.out, a.seconds)
    if a.role == "writer":
        return writer(a.shm, a.out)

    from emu_build import build_emulated
    from oracle import oracle
    from pg_embedding_b200 import build, sidecar
    build.build_sidecar()
    oracle.build("port")
    tmp = tempfile.mkdtemp(prefix="sidecar_stress_")
    lib = build_emulated(tmp)
    shm = f"/pgemb_stress_{os.getpid()}"
    srv = sidecar.SidecarProcess(shm, lib=lib, slots=8, max_dim=16, max_ef=16, bulk_mb=1, env={"PGEMB_EMU_SMS": "2", "PGEMB_EMU_TMA": "late"})
    srv.wait_ready()
    me = [sys.executable, os.path.abspath(__file__)]
    try:
        sidecar.RemoteIndex(1, DIMS, M, EFC, EFS, METRIC, capacity=N)     # create the mirror before anybody searches it
        wout = os.path.join(tmp, "writer.json")
        w = subprocess.Popen(me + ["--role", "writer", "--shm", shm, "--out", wout])
        procs, outs, killed, k = {}, [], 0, 0
        def spawn():
            nonlocal k
            o = os.path.join(tmp, f"s{k}.json")
            k += 1
            p = subprocess.Popen(me + ["--role", "searcher", "--shm", shm, "--out", o, "--seconds", str(max(1.0, t_end - time.time()))])
            procs[p.pid] = (p, o)
        t_end = time.time() + a.seconds
        for _ in range(a.searchers):
            spawn()
        rng = np.random.default_rng(5)
        while time.time() < t_end - 1.0:
            time.sleep(float(rng.uniform(1.0, 3.0)))
            live = [pid for pid, (p, _) in procs.items() if p.poll() is None]
            if live:
                victim = int(rng.choice(live))
                procs[victim][0].send_signal(signal.SIGKILL)       # dies wherever it is: spinning, asleep on its slot, mid-copy
                procs[victim][0].wait()
                del procs[victim]
                killed += 1
                spawn()
        calls = bad = 0
        for pid, (p, o) in procs.items():
            rc = p.wait(timeout=120)
            assert rc == 0, f"searcher {pid} exited with {rc}"
            r = json.load(open(o))
            calls += r["calls"]
            bad += r["bad"]
        assert w.wait(timeout=600) == 0, "the writer failed"
        assert bad == 0, f"{bad} searches returned labels that were never inserted"
        idx = sidecar.RemoteIndex(1, DIMS, M, EFC, EFS, METRIC, capacity=N)
        orc = oracle.FlatIndex("port", DIMS, M, EFC, EFS, METRIC, capacity=N)
        orc.build(points(), np.arange(1000, 1000 + N, dtype=np.uint64))
        assert len(idx) == N and idx.links().tobytes() == orc.links().tobytes(), "link lists differ from the sequential build"
        time.sleep(2.5)                                             # two reclaim passes of the sidecar
        raw = open("/dev/shm" + shm, "rb").read()
        slots_off, stride, nslots = struct.unpack_from("<Q", raw, 24)[0], struct.unpack_from("<I", raw, 20)[0], struct.unpack_from("<I", raw, 8)[0]
        states = [struct.unpack_from("<I", raw, slots_off + i * stride)[0] for i in range(nslots)]
        assert all(s == 0 for s in states), f"request slots left behind: {states}"
        assert struct.unpack_from("<I", raw, 64)[0] == 0, "bulk lock left behind"
        print(json.dumps({"seconds": a.seconds, "searchers": a.searchers, "killed_and_replaced": killed, "searches": calls, "inserts": N,
                          "sidecar": sidecar.stats(), "result": "link lists == sequential build; no slot or lock leaked"}))
    finally:
        sidecar.client().pgemb_client_disconnect()
        assert srv.stop() == 0


if __name__ == "__main__":
    main()
