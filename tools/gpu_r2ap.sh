#!/bin/bash
# round 2, call AP: final code of the round (LAB shift-and-add fusions, faster fog, k7 / k9 two-row medians, tapered host pipeline) -- GPU suite (shipped and debug-bounds builds), smoke, all config shapes,
# default bench.py invocation, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2ap_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2ap_pytest.log
RV_B200_LIB=librv_b200_dbg.so timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2ap_pytest_dbg.log 2>&1; echo "pytest(dbg) rc=$?"; tail -3 gpurun_out/r2ap_pytest_dbg.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python tests/perf/bench_configs.py --no-cpu > gpurun_out/r2ap_configs.jsonl 2> gpurun_out/r2ap_configs.err; echo "configs rc=$?"; cut -c1-200 gpurun_out/r2ap_configs.jsonl
timeout 900 python bench.py > gpurun_out/r2ap_bench.json 2> gpurun_out/r2ap_bench.err; echo "bench(default) rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2ap_bench.json'))
print({k:(round(j[k]['value']) if isinstance(j[k],dict) and 'value' in j[k] else j[k]) for k in ('value','steps','ms_per_step','sustained','e2e','e2e_tensor','e2e_keep','cpu_baseline')})
r=j['roofline']; print(r['frac'], r['avg_launch_ms'], r['traffic'], r['traffic_info'], r['issue']['frac'], r['kernel_share_of_step'], j['clocks'])
print(j['streams'])
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2ap_ref.json 2> gpurun_out/r2ap_ref.err; echo "ref rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2ap_ref.json')); print(j['value'], j['cpu_baseline']['kind'], j['cpu_baseline']['cores'])"
