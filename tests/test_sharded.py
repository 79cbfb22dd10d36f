"""Qubit-sharded statevector (qml_essentials_b200/sharded.py): epoch planning on the CPU
and the full exchange logic over gloo with world sizes 2 and 4.  The shard-local work runs
through the oracle interpreter here; on GPUs it is qmlb_evolve / qmlb_zsums
(tools/sharded_gpu_check.py, run with `gpurun --gpus N`)."""

import json
import os
import socket
import subprocess
import sys
import warnings

import numpy as np
import pytest

from qml_essentials_b200 import compiler, script
from qml_essentials_b200.model import Model
from qml_essentials_b200.sharded import plan_epochs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _program(n, L, ct):
    captured = {}

    class Capture:
        def execute(self, plan, host_args, batch, chunk=None, to_host=True):
            captured["plan"] = plan
            return np.zeros((1, n))

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = Model(n, L, ct)
        m.script.executor = Capture()
        m(params=np.random.default_rng(0).uniform(0, 6, (1, *m._params_shape)),
          inputs=np.array([[0.3]]))
    return captured["plan"].program


@pytest.mark.parametrize("g", [0, 1, 2, 3])
def test_epochs_keep_every_op_local_and_track_the_permutation(g):
    prog = _program(8, 2, "Hardware_Efficient")
    n, nl = prog.n_bits, prog.n_bits - g
    counts = []
    for rank in range(1 << g):
        steps, pos, consts = plan_epochs(prog, g, rank)
        assert sorted(pos) == list(range(n))
        n_ops = 0
        for st in steps:
            if st[0] == "ops":
                for o in st[1]:
                    assert all(0 <= b < nl for b in o["bits"][: o["k"]])
                    if o["kind"] == compiler.OP_PERM and o["aux"] == len(prog.consts):
                        continue  # local SWAP inserted by the planner
                    n_ops += 1
        # gates controlled by a global bit are dropped on the ranks where the control reads 0
        assert n_ops <= len(prog.ops) and (n_ops == len(prog.ops) or g > 0)
        exchanges = sum(1 for st in steps if st[0] == "exchange")
        assert (exchanges == 0) == (g == 0)
        assert list(consts[len(prog.consts):]) == [0, 2, 1, 3, 1, 0]
        counts.append(tuple(st[0] for st in steps).count("exchange"))
    assert len(set(counts)) == 1  # the exchange schedule does not depend on the rank


def test_global_controls_save_exchanges():
    """gates whose global bits are controls need no communication.  The
    32-qubit Hardware_Efficient circuit on 8 ranks: fewer exchanges than one per op that
    touches a global bit would need (19 before this rule)."""
    prog = _program(16, 8, "Hardware_Efficient")
    steps, _, _ = plan_epochs(prog, 3, 5)
    assert sum(1 for st in steps if st[0] == "exchange") <= 17


def test_too_many_ranks_is_an_error():
    prog = _program(4, 1, "Circuit_19")
    with pytest.raises(ValueError):
        plan_epochs(prog, 3)


_RANK_CODE = r"""
import os, sys, json, warnings
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import numpy as np, torch.distributed as dist
from qml_essentials_b200 import script
from _interp_executor import InterpExecutor, InterpShardEngine
script._set_executor_for_testing(InterpExecutor())
dist.init_process_group("gloo")
from qml_essentials_b200.model import Model
from qml_essentials_b200.sharded import ShardedExecutor
warnings.simplefilter("ignore")
res = {{}}
for (n, L, ct) in ((6, 2, "Hardware_Efficient"), (7, 2, "Circuit_19"),
                   (6, 1, "Strongly_Entangling"), (6, 1, "Circuit_6")):
    m = Model(n, L, ct)
    p = np.random.default_rng(1).uniform(0, 6, (1, *m._params_shape))
    x = np.array([[0.37]])
    ref = np.asarray(m(params=p, inputs=x))
    ref_state = np.asarray(m(params=p, inputs=x, execution_type="state"))
    m2 = Model(n, L, ct)
    se = ShardedExecutor(engine=InterpShardEngine())
    m2.script.executor = se
    got = np.asarray(m2(params=p, inputs=x))
    got_state = np.asarray(m2(params=p, inputs=x, execution_type="state"))
    res[ct] = [float(np.abs(got - ref).max()), float(np.abs(got_state - ref_state).max()),
               se.stats["exchanges"]]
if dist.get_rank() == 0:
    print("RESULT " + json.dumps(res))
dist.destroy_process_group()
"""


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_circuit_equals_unsharded(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(port), "--no-python", sys.executable, "-c", _RANK_CODE.format(root=ROOT)]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")][-1]
    res = json.loads(line[len("RESULT "):])
    for ct, (e_ev, e_state, exchanges) in res.items():
        assert e_ev < 1e-12 and e_state < 1e-12, (ct, e_ev, e_state)
        assert exchanges >= 1
