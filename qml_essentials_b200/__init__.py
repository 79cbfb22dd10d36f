"""qml-essentials B200 backend.

Drop-in for the circuit-execution hot path of cirKITers/qml-essentials
(``Model.__call__`` -> ``Script.execute`` -> statevector / density-matrix
evolution + measurement), executed by hand-written sm_100a CUDA kernels behind
the C ABI declared in ``include/qmlb200.h``.  Host code is NumPy + ctypes;
PyTorch only owns device memory, streams and ``torch.distributed``.
"""

from .config import get_precision, set_precision  # noqa: F401

__version__ = "0.1.0"
