#!/bin/bash
# Round-end evidence on one B200: tests, smoke, bench (both precisions), launch list and
# --set full captures of the two hot kernels.  Writes into gpurun_out/.
set -u
O=gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > $O/final_pytest.log 2>&1; tail -1 $O/final_pytest.log
timeout 100 python __graft_entry__.py smoke > $O/final_smoke.log 2>&1; tail -1 $O/final_smoke.log
timeout 300 python bench.py > $O/final_bench_c128.json 2> $O/final_bench_c128.err; echo "bench c128 rc=$?"
timeout 200 python bench.py --precision complex64 --no-cpu-baseline > $O/final_bench_c64.json 2> $O/final_bench_c64.err; echo "bench c64 rc=$?"
timeout 100 python tools/probe_scale.py cfg5:30 cfg4:1024 cfg4f:1024 > $O/final_probe.jsonl 2>&1
QMLB_PAIR_RULE=0 timeout 100 python tools/probe_scale.py cfg5:30 cfg4f:1024 > $O/final_probe_nopair.jsonl 2>&1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --gate-pass-qubits 28"
timeout 100 $CMD > $O/final_plain.log 2>&1 && \
  timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --clock-control none -c 400 --csv --log-file $O/final_launches.csv $CMD > $O/final_ncu1.log 2>&1
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --gate-pass-qubits 0"
timeout 100 $CMD2 > $O/final_plain2.log 2>&1 && \
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_reg -s 6 -c 2 \
    -o $O/final_prof_kreg $CMD2 > $O/final_ncu2.log 2>&1
CMD3="python tools/probe_scale.py cfg5:28"
timeout 100 $CMD3 > $O/final_plain3.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_stream -s 30 -c 3 \
    -o $O/final_prof_kstream $CMD3 > $O/final_ncu3.log 2>&1
ls -la $O/final_* | awk '{print $5, $9}'
