"""Build the host-emulated form of the C-ABI library (capi.cu compiled by g++ against tests/emu's stand-in CUDA runtime,
kernels run by the SIMT emulator).  Test infrastructure only: the outputs go to a temporary directory chosen by the
calling fixture and nothing in pg_embedding_b200/ knows about them."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_emulated(out_dir: str) -> str:
    out = os.path.join(str(out_dir), "libpgemb_emulated.so")
    cmd = ["g++", "-x", "c++", "-std=c++17", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread"] + [
        "-I", os.path.join(ROOT, "tests", "emu", "fake_cuda"), "-I", os.path.join(ROOT, "include"), "-o", out,
        os.path.join(ROOT, "pg_embedding_b200", "csrc", "capi.cu"), os.path.join(ROOT, "tests", "emu", "emu_runtime.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    return out
