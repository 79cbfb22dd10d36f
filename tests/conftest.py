"""Test configuration.

* ``-m "not gpu"`` (CPU box): the oracle against the reference's known-answer
  tests and golden vectors, the host logic (recording, compiler, batching, Model
  API) driven through the oracle's program interpreter, and the C-ABI library's
  exported symbols.
* ``-m gpu`` (B200): the same parity cases and KATs through the CUDA library.
"""

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


@pytest.fixture(autouse=True)
def _executor(request):
    """GPU tests use the real CUDA executor; everything else the oracle-backed
    program interpreter (test infrastructure, never shipped)."""
    from qml_essentials_b200 import config, script

    config.set_precision("complex128")
    if request.node.get_closest_marker("gpu"):
        script._set_executor_for_testing(None)  # -> lazily creates CudaExecutor
    else:
        from _interp_executor import InterpExecutor

        script._set_executor_for_testing(InterpExecutor())
    yield
    script._set_executor_for_testing(None)
