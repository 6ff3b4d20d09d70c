import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_sessionstart(session):
    """The C-ABI library and the pybind11 shim are build artefacts (git-ignored): build them if a fresh checkout has
    none (nvcc cross-compiles sm_100a without a GPU; ~2 minutes).  An up-to-date build is left alone."""
    from pailliercryptolib_python_b200 import build as B
    if not os.path.exists(B.LIB):
        B.build_all()
