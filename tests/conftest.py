import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `-m gpu` on the B200 box)")


@pytest.fixture(scope="session")
def cref():
    """the C oracle (test infrastructure), built on demand."""
    from oracle import cref as c
    c.lib()
    return c


@pytest.fixture(scope="session")
def ctx():
    """one libbzk context on cuda:0 — fails loudly if the extension or the GPU is missing."""
    import bazuka_b200 as B
    c = B.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def hostshim():
    """ff.cuh/ec.cuh compiled for the host (explicit-carry build of the device algorithm)."""
    import ctypes as ct
    src = os.path.join(ROOT, "tests", "hostshim", "shim.cpp")
    out = os.path.join(ROOT, "tests", "hostshim", "_shim.so")
    deps = [src] + [os.path.join(ROOT, "bazuka_b200", "csrc", h) for h in ("ff.cuh", "ec.cuh", "witness_core.cuh", "pairing.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++",
                               "-I", os.path.join(ROOT, "bazuka_b200", "csrc"), src, "-o", out])
    return ct.CDLL(out)


def fr_ints(a):
    from oracle.py import field as Fd
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    return [Fd.fr_from_mont_bytes(x.tobytes()) for x in a]


def fr_arr(xs):
    from oracle.py import field as Fd
    return np.frombuffer(Fd.fr_vec_to_mont(xs), dtype=np.uint64).reshape(-1, 4).copy()
