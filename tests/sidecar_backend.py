"""One 'backend' of tests/test_sidecar.py: a separate process with its own connection to the sidecar that issues one query
per hnsw_search call (embedding.c:317), like a Postgres backend scanning the index.  Usage:
    python sidecar_backend.py SHM REL_KEY DIMS M EFC EFS METRIC EF QUERIES.npy OUT.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    shm, rel_key, dims, m, efc, efs, metric, ef, qpath, out = sys.argv[1:11]
    from pg_embedding_b200 import sidecar
    sidecar.connect(shm)
    idx = sidecar.RemoteIndex(int(rel_key), int(dims), int(m), int(efc), int(efs), metric, capacity=1)  # attach = look the mirror up
    q = np.load(qpath)
    # start line: all backends of a test begin their scans together, so that their calls really are concurrent
    open(out + ".ready", "w").close()
    go = os.path.join(os.path.dirname(out), "go")
    deadline = time.time() + 120
    while not os.path.exists(go) and time.time() < deadline:
        time.sleep(0.002)
    res = [idx.search(v, int(ef)).tolist() for v in q]
    json.dump(res, open(out, "w"))


if __name__ == "__main__":
    main()
