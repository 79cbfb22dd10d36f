"""The C-ABI shared library: loads on a CPU box, exports every symbol the header
declares, and the ctypes structs match the header's layout.  No compute calls."""

import ctypes
import os
import re

import pytest

from qml_essentials_b200 import backend, compiler

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qmlb200.h")


@pytest.fixture(scope="module")
def lib():
    from qml_essentials_b200 import build

    build.build(force=False, verbose=False)
    return backend.load_library()


def test_library_exports_every_declared_symbol(lib):
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(qmlb_[a-z_]+)\s*\(", text))
    assert declared == set(backend.EXPORTED_SYMBOLS)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert lib.qmlb_version() == 100
    assert lib.qmlb_launch_count() == 0


def test_struct_layouts_match_header():
    assert compiler.OP_DTYPE.itemsize == 48
    assert compiler.SRC_DTYPE.itemsize == 40
    assert compiler.ANGLE_DTYPE.itemsize == 16
    assert compiler.TERM_DTYPE.itemsize == 16
    assert compiler.OBS_DTYPE.itemsize == 56
    assert ctypes.sizeof(backend._Arg) == 32
    text = open(HEADER).read()
    for name, val in (("QMLB_OP_MAT", compiler.OP_MAT), ("QMLB_OP_CTRL1", compiler.OP_CTRL1),
                      ("QMLB_OP_PERM", compiler.OP_PERM), ("QMLB_OP_DIAG", compiler.OP_DIAG),
                      ("QMLB_SRC_SUPER", compiler.SRC_SUPER),
                      ("QMLB_OUT_DENSITY", compiler.OUT_DENSITY),
                      ("QMLB_OBS_DENSE", compiler.OBS_DENSE),
                      ("QMLB_MAX_OP_BITS", compiler.MAX_OP_BITS)):
        assert re.search(rf"#define {name} {val}\b", text), name


def test_header_cites_reference_interfaces():
    text = open(HEADER).read()
    for cite in ("script.py:137-147", "simulation.py:131-201", "simulation.py:320-377",
                 "operations.py:485-512", "memory.py:54-139"):
        assert cite in text


def test_no_cpu_fallback_without_gpu():
    """On a box without CUDA the product must fail loudly, not compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from qml_essentials_b200 import operations as op, script

    script._set_executor_for_testing(None)
    with pytest.raises(backend.BackendUnavailable):
        script.Script(lambda: op.H(wires=0), 1).execute("probs")
