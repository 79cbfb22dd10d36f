"""Digest of one ncu report (--set full, --import-source on): headline counters and stall
samples aggregated by SASS opcode.   python tools/ncu_digest.py REPORT.ncu-rep [kernel-row]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2 + row]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__cluster_max_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
for h, u, v in zip(hdr, rows[1], vals):
    if h in want:
        print(f"{h} [{u}] = {v}")
stalls = [(h, float(v)) for h, v in zip(hdr, vals)
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
print("stalls per issue:", ", ".join(f"{h.split('stalled_')[1].split('_per')[0]}={v:.2f}"
                                      for h, v in sorted(stalls, key=lambda x: -x[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the source page lists kernels one after another; take the first block
body = []
for r in rows[2:]:
    if len(r) < 6 or not r[0].startswith("0x"):
        if body:
            break
        continue
    body.append(r)
tot = sum(int(r[2]) for r in body) or 1
toti = sum(int(r[5]) for r in body) or 1
by, byi = collections.Counter(), collections.Counter()
for r in body:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[1])
    op = m.group(2).split(".")[0] if m else "?"
    if "IMAD.MOV" in r[1]:
        op = "IMAD.MOV"
    by[op] += int(r[2])
    byi[op] += int(r[5])
print(f"samples {tot}, warp instructions {toti}")
for op, c in by.most_common(14):
    print(f"  {op:14s} samples {100 * c / tot:5.1f}%   instructions {100 * byi[op] / toti:5.1f}%")
print("hottest instructions:")
for r in sorted(body, key=lambda r: -int(r[2]))[:10]:
    print(f"  {int(r[2]):7d} {int(r[5]):10d}  {r[1].strip()[:80]}")
