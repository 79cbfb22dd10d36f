"""Stall samples and shared-memory wavefronts of one ncu report (--set full, --import-source
on, kernels built with -lineinfo) aggregated by source line.
python tools/ncu_lines.py REPORT.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, hdr, lines, total = None, None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or not r[0]:
        continue  # SASS rows have an empty line number
    d = dict(zip(hdr, r))
    try:
        smp = int(d["# Samples"])
    except (KeyError, ValueError):
        continue
    total += smp
    lines.append((smp, cur, int(r[0]), r[1].strip(), int(d["Instructions Executed"]),
                  int(d.get("L1 Wavefronts Shared", 0) or 0),
                  int(d.get("L1 Wavefronts Shared Ideal", 0) or 0)))
lines.sort(reverse=True)
print(f"samples {total}")
for smp, f, ln, src, ins, wf, wfi in lines[:top]:
    print(f"{100.0 * smp / max(total, 1):5.1f}%  inst {ins:>11}  smem wf {wf:>10} ideal {wfi:>10}  {f}:{ln}  {src[:90]}")
