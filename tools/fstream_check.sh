#!/bin/bash
# k_fstream iteration check on one B200: the strategy-4 parity tests, the gate-pass leg, and
# one --set full capture of the kernel at n = 28.  Writes into gpurun_out/.
set -u
O=gpurun_out
R=${1:-r2_fs}
N=${2:-30}
timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_gpu_parity.py -m gpu -q -x -k "streamed or large or invariant or stream" > $O/${R}_pytest.log 2>&1; tail -2 $O/${R}_pytest.log
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config-legs --precision complex64 --gate-pass-qubits $N"
timeout 600 $CMD > $O/${R}_bench.json 2> $O/${R}_bench.err; echo "bench rc=$?"
python - <<P
import json
d=json.loads(open("$O/${R}_bench.json").read().strip().splitlines()[-1])
g=d["gate_pass"]; print({k:g[k] for k in ("n_qubits","passes","ms_per_circuit","ms_per_pass","achieved_gbs","frac") if k in g})
P
if [ "${3:-ncu}" = "ncu" ]; then
CMD3="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-config-legs --precision complex64 --gate-pass-qubits 28"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_fstream -s 3 -c 1 -o $O/${R}_kfstream -f $CMD3 > $O/${R}_ncu.log 2>&1; echo "ncu rc=$?"
fi
