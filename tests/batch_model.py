"""Array-level model of the CUDA traversal's per-hop queue update (search_kernel.cuh, "K2").

The kernel does not replay the reference's sequential push/pop loop (hnswalg.cpp:89-108); it scores all
unvisited neighbours of a hop at once and then applies an update that is claimed to be EQUIVALENT,
exact-distance ties included.  This model states that update with plain Python lists, mirroring the
kernel's data flow one-to-one (accept rule, rank-based merge, tie "overflow" list, pop order), so the
claim can be model-checked against the oracle on CPU (tests/test_batch_semantics.py) without a GPU.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import numpy as np


def search_base_layer_batched(dist_fn, links, n_items, q, ef, entry=0):
    """links: [n, maxM+1] uint32 ([count, ids...]); dist_fn(q, id) -> python float (np.float32 value).
    Returns (list of (dist, id) ascending by (dist,id), stats dict)."""
    stats = {"dist": 0, "hops": 0, "ovf_hw": 0}
    if n_items == 0 or entry >= n_items:
        return [], stats
    visited = set([entry])
    res = []          # list of [dist, id, expanded] ascending by (dist, id)
    ovf = []          # list of (dist, id): evicted, unexpanded, dist == current worst
    hop = [entry]
    while True:
        if hop:
            n = len(hop)
            stats["dist"] += n
            d = [dist_fn(q, i) for i in hop]
            r = len(res)
            W = res[-1][0] if r == ef else None
            acc = []
            for k in range(n):
                if r + n <= ef:
                    ok = True
                elif r == ef and d[k] >= W:
                    ok = False
                else:
                    cR = sum(1 for e in res if e[0] <= d[k])
                    if cR >= ef:
                        ok = False
                    else:
                        cP = sum(1 for j in range(k) if d[j] <= d[k])
                        ok = cR + cP < ef
                if ok:
                    acc.append((d[k], hop[k]))
            if acc:
                merged = sorted([(e[0], e[1], e[2]) for e in res] + [(x[0], x[1], False) for x in acc],
                                key=lambda e: (e[0], e[1]))
                total = len(merged)
                keep, evicted = merged[:ef], merged[ef:]
                res = [list(e) for e in keep]
                if total > ef:
                    Wn = res[-1][0]
                    if ovf and Wn != W:
                        ovf = []
                    for e in evicted:
                        if not e[2] and e[0] == Wn:
                            ovf.append((e[0], e[1]))
                    stats["ovf_hw"] = max(stats["ovf_hw"], len(ovf))
                    assert len(ovf) <= ef
        # pop: min dist, ties -> larger id first, over unexpanded results + overflow
        best = None
        for i, e in enumerate(res):
            if not e[2]:
                best = i
                break
        if best is not None:
            D = res[best][0]
            i = best + 1
            while i < len(res) and res[i][0] == D:
                if not res[i][2]:
                    best = i
                i += 1
        c = None
        if ovf:
            bo = min(range(len(ovf)), key=lambda t: (ovf[t][0], -ovf[t][1]))
            od, oid = ovf[bo]
            if best is None or (od, -oid) < (res[best][0], -res[best][1]):
                c = oid
                ovf[bo] = ovf[-1]
                ovf.pop()
        if c is None:
            if best is None:
                break
            c = res[best][1]
            res[best][2] = True
        stats["hops"] += 1
        cnt = int(links[c][0])
        hop = []
        for t in links[c][1:1 + cnt]:
            t = int(t)
            if t not in visited:
                visited.add(t)
                hop.append(t)
    return [(e[0], e[1]) for e in res], stats
