"""Thread-local operation tapes (host side).

Same contract as the reference's ``qml_essentials/tape.py`` (tape.py:10-138):
operations append themselves to the innermost active tape of the current
thread; ``recording()`` pushes a fresh tape; ``copy_to_tape`` replays a
sub-circuit with shifted wires (used by the 2n/3n-qubit Bell / swap-test
scripts, entanglement.py:145-166).
"""

from __future__ import annotations

import copy
import threading
from contextlib import contextmanager
from typing import Callable, Iterator, List, Optional

_tls = threading.local()


def _stack(kind: str) -> list:
    st = getattr(_tls, kind, None)
    if st is None:
        st = []
        setattr(_tls, kind, st)
    return st


def active_tape() -> Optional[list]:
    """Innermost tape being recorded on this thread, or ``None`` (tape.py:24-34)."""
    st = _stack("ops")
    return st[-1] if st else None


@contextmanager
def recording() -> Iterator[list]:
    """Record operations instantiated inside the block (tape.py:37-55)."""
    st = _stack("ops")
    tape: list = []
    st.append(tape)
    try:
        yield tape
    finally:
        st.pop()


def active_pulse_tape() -> Optional[list]:
    """Pulse-event tape (tape.py:65-72); only kept for API compatibility."""
    st = _stack("pulse")
    return st[-1] if st else None


@contextmanager
def pulse_recording() -> Iterator[list]:
    st = _stack("pulse")
    tape: list = []
    st.append(tape)
    try:
        yield tape
    finally:
        st.pop()


def shift_and_append(tape_ops: List, offset: int) -> None:
    """Append wire-shifted shallow copies of ``tape_ops`` to the active tape
    (tape.py:92-112)."""
    dst = active_tape()
    if dst is None:
        return
    for o in tape_ops:
        clone = copy.copy(o)
        clone._wires = [w + offset for w in o.wires]
        dst.append(clone)


def copy_to_tape(fn: Callable[[], None], offset: int) -> None:
    """Record ``fn`` on a side tape and replay it shifted by ``offset``
    (tape.py:115-138)."""
    with recording() as side:
        fn()
    shift_and_append(side, offset)
