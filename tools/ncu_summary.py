"""Print selected metrics per kernel from an .ncu-rep (runs `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_fp64.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
extra = [h for h in hdr if 'issue_stalled' in h and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:60])
    for k in KEYS:
        if k in hdr:
            print(f'  {k} = {r[hdr.index(k)]} {units[hdr.index(k)]}')
    st = sorted(((float(r[hdr.index(k)] or 0), k) for k in extra), reverse=True)[:6]
    for v, k in st:
        print(f'  stall {k.split("issue_stalled_")[1].split("_per_issue")[0]} = {v:.2f}')
