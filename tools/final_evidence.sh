#!/bin/bash
# Round-end evidence on one B200: tests, smoke, bench (both precisions), launch list and
# --set full captures of the hot kernels.  Writes into gpurun_out/.
set -u
O=gpurun_out
R=${1:-r2}
timeout 900 python -m pytest tests -m gpu -q > $O/${R}_final_pytest.log 2>&1; tail -3 $O/${R}_final_pytest.log
timeout 200 python __graft_entry__.py smoke > $O/${R}_final_smoke.log 2>&1; tail -1 $O/${R}_final_smoke.log
timeout 900 python bench.py > $O/${R}_final_bench_c128.json 2> $O/${R}_final_bench_c128.err; echo "bench c128 rc=$?"
timeout 400 python bench.py --precision complex64 --no-cpu-baseline --no-config-legs --gate-pass-qubits 0 > $O/${R}_final_bench_c64.json 2> $O/${R}_final_bench_c64.err; echo "bench c64 rc=$?"
if [ "${2:-ref}" = "ref" ]; then
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_final_bench_reference.json 2> $O/${R}_final_bench_reference.err; echo "bench reference rc=$?"
fi
timeout 300 python tools/probe_dev.py cfg1 cfg3 cfg4:4096 cfg4f:4096 > $O/${R}_final_probe_dev.jsonl 2>&1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-config-legs --gate-pass-qubits 28"
timeout 200 $CMD > $O/${R}_final_plain.log 2>&1 && \
  timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
    --clock-control none -c 600 --csv --log-file $O/${R}_final_launches.csv $CMD > $O/${R}_final_ncu1.log 2>&1
if [ "${3:-all}" = "all" ]; then
CMD2="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-config-legs --gate-pass-qubits 0"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_reg -c 1 -o $O/${R}_final_kreg -f $CMD2 > $O/${R}_final_ncu2.log 2>&1
timeout 200 python tools/probe_dev.py cfg4:512 > /dev/null 2>&1 && \
  timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_frame_ptm -c 1 -o $O/${R}_final_kptm -f python tools/probe_dev.py cfg4:512 > $O/${R}_final_ncu3.log 2>&1
fi
CMD3="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-graph --no-config-legs --precision complex64 --gate-pass-qubits 28"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_fstream -s 3 -c 1 -o $O/${R}_final_kfstream -f $CMD3 > $O/${R}_final_ncu4.log 2>&1
ls -la $O/${R}_final_* | awk '{print $5, $9}'
