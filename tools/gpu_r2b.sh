#!/bin/bash
# round 2, call B: parity suite on the single-frame graph path, per-frame latency, D2H experiment (1 GPU)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2b_pytest.log
timeout 600 python tests/perf/bench_latency.py > gpurun_out/r2b_latency.jsonl 2> gpurun_out/r2b_latency.err; echo "latency rc=$?"; cat gpurun_out/r2b_latency.jsonl; tail -5 gpurun_out/r2b_latency.err
timeout 600 python tools/exp_d2h.py > gpurun_out/r2b_d2h_n1.json 2> gpurun_out/r2b_d2h.err; echo "d2h rc=$?"; cat gpurun_out/r2b_d2h_n1.json; tail -5 gpurun_out/r2b_d2h.err
timeout 600 python bench.py --steps 20 --warmup 5 --stream-seconds 0 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; python - <<'PY'
import json
j=json.load(open('gpurun_out/r2b_bench.json'))
print({k:(j[k]['value'] if isinstance(j[k],dict) else j[k]) for k in ('value','e2e','e2e_tensor','e2e_keep')})
PY
tail -3 gpurun_out/r2b_bench.err
