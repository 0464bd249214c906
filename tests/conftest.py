import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "kinematicdriver.jl_b200"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build the CUDA library (nvcc cross-compiles
    without a GPU) and the oracle once, exactly as __graft_entry__.build() does."""
    lib = ROOT / "kinematicdriver.jl_b200" / "csrc" / "libkid_cuda.so"
    ora = ROOT / "oracle" / "libkid_oracle.so"
    if not lib.exists() or not ora.exists():
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle
    oracle.build()
    return oracle.load()
