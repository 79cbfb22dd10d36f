"""Tabulate an `ncu --metrics ... --csv` launch list: per launch time, dram bytes, GB/s."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
H = rows[hdr]
data = {}
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    d = dict(zip(H, r))
    k = (int(d['ID']), d['Kernel Name'][:34])
    data.setdefault(k, {})[d['Metric Name']] = float(d['Metric Value'].replace(',', ''))
tot = 0
for (i, k), m in sorted(data.items()):
    t = m.get('gpu__time_duration.sum', 0) / 1e6
    rd = m.get('dram__bytes_read.sum', 0)
    wr = m.get('dram__bytes_write.sum', 0)
    ins = m.get('smsp__inst_executed.sum', 0)
    print(f"{i:4d} {k:36s} {t:8.3f} ms  rd {rd/1e9:7.3f} GB wr {wr/1e9:7.3f} GB "
          f"{(rd+wr)/1e9/max(t,1e-9)*1e3:8.1f} GB/s  inst {ins/1e6:8.1f} M")
    tot += t
print('total ms', tot)
