"""Generate tests/golden/topologies.json from the REFERENCE's own topologies.py.

``qml_essentials/topologies.py`` is the one hot-path module of the reference that
imports without JAX, so its outputs can be captured in this container and
committed as golden vectors (the GPU box has no /root/reference).

Run here:  python tools/gen_golden_topologies.py
"""

import json
import os
import sys

sys.path.insert(0, "/root/reference")
from qml_essentials.topologies import Topology  # noqa: E402

# The (topology, kwargs) combinations every ansatz in ansaetze.py:410-756 uses.
CASES = {
    "stairs_default": ("stairs", {}),
    "stairs_wrap_nomirror": ("stairs", dict(wrap=True, mirror=False)),
    "stairs_wrap_rev_nomirror": ("stairs", dict(wrap=True, reverse=True, mirror=False)),
    "stairs_wrap_norev_nomirror": ("stairs", dict(wrap=True, reverse=False, mirror=False)),
    "stairs_c10": ("stairs", dict(offset=-1, wrap=True)),
    "stairs_c13b": ("stairs", dict(reverse=False, mirror=False, offset="n-1", span=3, wrap=True)),
    "stairs_c20b": ("stairs", dict(reverse=False, offset="n-2", span=1, wrap=True)),
    "stairs_se_b": ("stairs", dict(reverse=False, span="n//2", wrap=True, mirror=False)),
    "stairs_ghz": ("stairs", dict(reverse=True)),
    "bricks_default": ("bricks", {}),
    "bricks_offset1": ("bricks", dict(offset=1)),
    "bricks_nomirror": ("bricks", dict(mirror=False)),
    "bricks_he_b": ("bricks", dict(offset=-1, modulo=True, wrap=True, mirror=False)),
    "all_to_all": ("all_to_all", {}),
}
LAMBDAS = {"n-1": lambda n: n - 1, "n-2": lambda n: n - 2, "n//2": lambda n: n // 2}


def main():
    out = {}
    for name, (topo, kw) in CASES.items():
        kw_real = {k: LAMBDAS.get(v, v) if isinstance(v, str) else v for k, v in kw.items()}
        per_n = {}
        for n in range(2, 13):
            pairs = getattr(Topology, topo)(n_qubits=n, **kw_real)
            per_n[str(n)] = [[int(a), int(b)] for a, b in pairs]
        out[name] = {"topology": topo, "kwargs": kw, "pairs": per_n}
    path = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "topologies.json")
    with open(path, "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("wrote", os.path.normpath(path), len(out), "cases")


if __name__ == "__main__":
    main()
