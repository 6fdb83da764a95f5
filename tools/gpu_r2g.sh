#!/bin/bash
# round 2, call G: staged pageable upload -- parity + latency
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
timeout 600 python tests/perf/bench_latency.py > gpurun_out/r2g_latency.jsonl 2> gpurun_out/r2g_latency.err; echo "latency rc=$?"; cat gpurun_out/r2g_latency.jsonl; tail -3 gpurun_out/r2g_latency.err
timeout 600 python tests/perf/bench_fog.py > gpurun_out/r2g_fog.json 2> gpurun_out/r2g_fog.err; echo "fog rc=$?"; cat gpurun_out/r2g_fog.json
