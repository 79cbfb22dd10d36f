"""Qubit-sharded statevector on real GPUs (launch with torchrun, one rank per GPU):

    gpurun --gpus 2 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
        --master-addr 127.0.0.1 --master-port 29500 tools/sharded_gpu_check.py 20 31'

argv: the sizes n to run.  n <= 26 is compared against the unsharded single-GPU result of
rank 0; larger n reports time, exchanges and the product-state / norm invariants."""
import json
import os
import sys
import time
import warnings

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from qml_essentials_b200 import config  # noqa: E402
from qml_essentials_b200.model import Model  # noqa: E402
from qml_essentials_b200.sharded import ShardedExecutor  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
config.set_precision("complex64")
warnings.simplefilter("ignore")

from qml_essentials_b200.sharded import CudaShardEngine  # noqa: E402

results = {}
for n in [int(a) for a in sys.argv[1:]] or [20]:
    for mode in ("nccl", "fused"):
        m = Model(n_qubits=n, n_layers=8, circuit_type="Hardware_Efficient")
        p = np.random.default_rng(1000).uniform(0, 2 * np.pi, (1, *m._params_shape))
        x = np.array([[0.5]])
        se = ShardedExecutor(engine=CudaShardEngine(fused=(mode == "fused")))
        m.script.executor = se
        try:
            ev = np.asarray(m(params=p, inputs=x)).reshape(-1)  # warm-up + plan
        except Exception as exc:  # noqa: BLE001
            if rank == 0:
                print(json.dumps({"n": n, "mode": mode,
                                  "error": f"{type(exc).__name__}: {exc}"[:300]}), flush=True)
            continue
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        ev = np.asarray(m(params=p, inputs=x)).reshape(-1)
        torch.cuda.synchronize()
        dist.barrier()
        dt = time.perf_counter() - t0
        line = {"n": n, "mode": mode, "ranks": world, "seconds": dt, **se.stats,
                "abs_le_1": bool(np.all(np.abs(ev) <= 1 + 1e-4))}
        results[(n, mode)] = ev
        if (n, "nccl") in results and mode == "fused":
            line["max_abs_diff_vs_nccl"] = float(np.abs(ev - results[(n, "nccl")]).max())
        if n <= 26 and rank == 0:
            ref_m = Model(n_qubits=n, n_layers=8, circuit_type="Hardware_Efficient")
            ref = np.asarray(ref_m(params=p, inputs=x)).reshape(-1)
            line["max_abs_err_vs_single_gpu"] = float(np.abs(ev - ref).max())
        if rank == 0:
            print(json.dumps(line), flush=True)
        del m, se
        torch.cuda.empty_cache()
torch.cuda.synchronize()
os._exit(0)
