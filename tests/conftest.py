"""Test configuration.

* ``-m "not gpu"``: oracle vs golden vectors, host front end, C-ABI loads/exports (no GPU).
* ``-m gpu``: the parity tests proper -- every one calls the CUDA path through the C-ABI
  (tiny-path-tracer_b200/lib/libtpt.so) and compares with the oracle / golden fixtures.
Only tests (and __graft_entry__.smoke, bench.py's cpu_baseline leg) may touch oracle/.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def T():
    import tpt_b200
    return tpt_b200


@pytest.fixture(scope="session")
def O():
    import oracle_ref
    if not oracle_ref.available(True):
        pytest.skip("oracle/_ref is not built (run __graft_entry__.build() where /root/reference exists)")
    return oracle_ref


@pytest.fixture(scope="session")
def gpu(T):
    """GPU tests must run the CUDA extension: a missing library is a failure, not a skip."""
    assert os.path.exists(T.LIBTPT), "libtpt.so missing on a GPU run: build it, there is no fallback"
    n = T.device_count()
    if n == 0:
        pytest.skip("no CUDA device visible")
    return n
