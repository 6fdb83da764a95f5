#!/bin/bash
# round 2, call H: final suite + the default bench.py invocation (no flags) + reference arm + smoke
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2h_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/r2h_bench_default.json 2> gpurun_out/r2h_bench_default.err; echo "bench(default) rc=$?"; grep -E "Elapsed|Maximum resident" gpurun_out/r2h_bench_default.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2h_bench_default.json'))
print({k:(round(j[k]['value']) if isinstance(j[k],dict) and 'value' in j[k] else j[k]) for k in ('value','steps','ms_per_step','sustained','e2e','e2e_tensor','e2e_keep','cpu_baseline')})
print(j['roofline']['frac'], j['roofline']['traffic'], j['roofline']['traffic_info'], j['roofline']['issue']['frac'], j['clocks'])
print(j['streams'])
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2h_ref.json 2> gpurun_out/r2h_ref.err; echo "ref rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2h_ref.json')); print(j['value'], j['cpu_baseline']['kind'], j['cpu_baseline']['cores'])"
