"""bench.py contract that can be checked without a GPU: the reference arm prints exactly ONE
JSON line on stdout with the agreed keys (the CPU restatement of the reference's path on a
bounded sample), and bench.py only reaches into oracle/ from its CPU legs."""

import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1",
                        "--warmup", "3"], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["metric"] == "circuit evals/sec (batched)" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_product_never_imports_the_oracle():
    """Only tests/, bench.py's CPU legs and __graft_entry__.smoke() may touch oracle/."""
    pkg = os.path.join(ROOT, "qml_essentials_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
    bench = open(os.path.join(ROOT, "bench.py")).read()
    uses = [m.start() for m in re.finditer(r"from oracle import", bench)]
    assert uses, "bench.py times the CPU restatement"
    # every use sits inside oracle_cpu_evals_per_s (the cpu_baseline / reference legs)
    body_start = bench.index("def oracle_cpu_evals_per_s")
    body_end = bench.index("\ndef ", body_start + 10)
    assert all(body_start < u < body_end for u in uses)
