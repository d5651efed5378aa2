"""scan_tile_kernel (the tiled exact-scan distance step) executed on the HOST: the .cuh is compiled by g++ against a
stand-in <cuda_runtime.h> (tests/emu/fake_cuda: a small SIMT emulator, warps = OS threads, lanes = fibers) and run one
CTA at a time.  Checks the kernel's indexing, chunking, tails and summation order bit for bit against the oracle
without a GPU.  (It says nothing about speed or about nvcc's code generation.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F32P = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libscan_emu.so")
    cmd = ["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", os.path.join(ROOT, "tests", "emu", "fake_cuda"),
           "-o", out, os.path.join(ROOT, "tests", "emu", "scan_tile_emu.cpp"), os.path.join(ROOT, "tests", "emu", "emu_runtime.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    lib = C.CDLL(out)
    lib.emu_scan_tile.argtypes = [C.c_int, F32P, F32P, C.c_uint32, C.c_uint32, F32P, C.c_uint32, F32P, C.c_uint32, C.c_uint32, C.c_uint32, F32P]
    lib.emu_scan_tile.restype = None
    return lib


def sqnorm_lane_order(v):
    """|v|^2 in the cosine lane order (4 lane-strided partial sums, (s0+s2)+(s1+s3), scalar tail): distfunc.c:141-142 as built."""
    v = v.astype(np.float32)
    main = v.size & ~3
    s = np.zeros(4, np.float32)
    for i in range(0, main, 4):
        s = s + v[i:i + 4] * v[i:i + 4]
    res = np.float32(np.float32(s[0] + s[2]) + np.float32(s[1] + s[3]))
    for e in range(main, v.size):
        res = np.float32(res + np.float32(v[e] * v[e]))
    return res


def p(a):
    return a.ctypes.data_as(F32P)


@pytest.mark.parametrize("metric", ["l2", "cosine", "manhattan"])
@pytest.mark.parametrize("dim,nq,n,r0,nr", [(3, 1, 10, 0, 10), (16, 5, 70, 3, 65), (33, 33, 130, 1, 129), (64, 17, 64, 0, 64),
                                             (100, 40, 200, 7, 150), (203, 9, 90, 0, 90), (768, 6, 70, 2, 66)])
def test_tiled_scan_distances_bit_exact(emu, oracle_mod, metric, dim, nq, n, r0, nr):
    rng = np.random.default_rng(dim * 7 + nq)
    x = rng.standard_normal((n, dim)).astype(np.float32)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    if metric == "cosine":
        x, q = x + 0.5, q + 0.5
    row_f = (dim + 3) & ~3
    xv = np.zeros((n, row_f), np.float32)
    xv[:, :dim] = x
    norms = np.array([sqnorm_lane_order(x[i]) for i in range(n)], np.float32)
    qn = np.array([sqnorm_lane_order(q[i]) for i in range(nq)], np.float32)
    out = np.full((nq, nr), np.nan, np.float32)
    emu.emu_scan_tile({"l2": 0, "cosine": 1, "manhattan": 2}[metric], p(xv), p(norms), row_f, dim, p(q), dim, p(qn), nq, r0, nr, p(out))
    for i in range(nq):
        want = oracle_mod.dist_many("port", metric, q[i], x[r0:r0 + nr])
        assert out[i].tobytes() == want.tobytes(), (metric, dim, i, np.flatnonzero(out[i] != want)[:5])
