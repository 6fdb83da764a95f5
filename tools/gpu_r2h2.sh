#!/bin/bash
mkdir -p gpurun_out
T0=$(date +%s.%N)
timeout 900 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err; echo "bench(default) rc=$?"
T1=$(date +%s.%N); echo "wall seconds: $(echo "$T1 - $T0" | bc)"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2h_bench_default.json'))
print({k:(round(j[k]['value']) if isinstance(j[k],dict) and 'value' in j[k] else j[k]) for k in ('value','steps','ms_per_step','sustained','e2e','e2e_tensor','e2e_keep','cpu_baseline')})
print(j['roofline']['frac'], j['roofline']['traffic'], j['roofline']['traffic_info'], j['roofline']['issue']['frac'], j['clocks'])
print(j['streams'])
PY
tail -3 gpurun_out/r2h_bench_default.err
