#!/bin/bash
# per-config throughput + ncu capture of the LAB chain kernel
mkdir -p gpurun_out
timeout 900 python tests/perf/bench_configs.py --no-cpu > gpurun_out/configs_v4.jsonl 2>gpurun_out/configs_v4.err; echo "configs rc=$?"
cut -c1-200 gpurun_out/configs_v4.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 13 -c 1 -f -o gpurun_out/prof_chain_lab_v4 \
    python tests/perf/bench_configs.py --no-cpu > gpurun_out/ncu_lab.log 2>&1
echo "lab capture rc=$?"
