"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"`: oracle vs golden vectors / vs the compiled reference, host logic, C-ABI symbol
checks.  `-m gpu`: parity tests proper, through the C-ABI on a real B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle

    oracle.build("all")
    return oracle


@pytest.fixture(scope="session")
def have_ref(oracle_mod):
    return oracle_mod.available("ref")


def checker_names(oracle_mod):
    names = ["port"]
    if oracle_mod.available("ref"):
        names.append("ref")
    return names
