"""Run the shared parity cases on a GPU through the C ABI and print max errors.
Usage (on a GPU box): python tools/gpu_parity.py [strategy]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
if len(sys.argv) > 1:
    os.environ["QMLB_FORCE_STRATEGY"] = sys.argv[1]  # 0 registers, 1 shared memory, 2 streamed
import parity_cases as pc  # noqa: E402

out = {}
for prec in ("complex128", "complex64"):
    t0 = time.time()
    res = {}
    for name in ("case_every_gate", "case_every_channel", "case_baseline_configs",
                 "case_all_ansaetze", "case_noise_keys", "case_shots"):
        try:
            res[name] = getattr(pc, name)(prec)
        except Exception as e:  # noqa: BLE001
            res[name] = f"ERROR {type(e).__name__}: {e}"
    res["seconds"] = round(time.time() - t0, 2)
    out[prec] = res
print(json.dumps(out, indent=1, default=str))
