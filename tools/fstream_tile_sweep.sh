for cfg in "13 6" "12 6" "12 5" "13 5" "13 7" "14 6"; do
set -- $cfg
QMLB_FSTREAM_TILE_BITS=$1 QMLB_FSTREAM_LOW_BITS=$2 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-config-legs --precision complex64 --gate-pass-qubits 30 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); g=d['gate_pass']
print('T=$1 L=$2', g['passes'], round(g['ms_per_circuit'],2), round(g['achieved_gbs'],1), g['checks'])
"
done
