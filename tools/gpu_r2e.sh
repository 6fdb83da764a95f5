#!/bin/bash
# round 2, call E: full parity suite, latency / fog / config tables, then the ncu captures of the shipped kernels (one GPU)
mkdir -p gpurun_out
python -c "import rvb200; print(rvb200.kernel_source_hash())" > gpurun_out/r2e_hash.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2e_pytest.log
timeout 600 python tests/perf/bench_latency.py > gpurun_out/r2e_latency.jsonl 2> gpurun_out/r2e_latency.err; echo "latency rc=$?"; cat gpurun_out/r2e_latency.jsonl
timeout 600 python tests/perf/bench_fog.py > gpurun_out/r2e_fog.json 2> gpurun_out/r2e_fog.err; echo "fog rc=$?"; cat gpurun_out/r2e_fog.json; tail -3 gpurun_out/r2e_fog.err
timeout 900 python tests/perf/bench_configs.py > gpurun_out/r2e_configs.jsonl 2> gpurun_out/r2e_configs.err; echo "configs rc=$?"; cut -c1-230 gpurun_out/r2e_configs.jsonl
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --stream-seconds 0"
timeout 300 $B > gpurun_out/r2e_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv $B > gpurun_out/r2e_ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 8 -c 1 -f -o gpurun_out/r2e_prof_chain $B > gpurun_out/r2e_ncu_chain.log 2>&1; echo "chain capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_luma_hist -s 8 -c 1 -f -o gpurun_out/r2e_prof_hist $B > gpurun_out/r2e_ncu_hist.log 2>&1; echo "hist capture rc=$?"
for cfg in "1080p LAB k3" "1080p LAB k5"; do
  tag=$(echo "$cfg" | tr ' ' '_')
  timeout 300 python tests/perf/bench_configs.py --no-cpu --only "$cfg" > gpurun_out/r2e_plain_$tag.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 3 -c 1 -f -o gpurun_out/r2e_prof_chain_$tag python tests/perf/bench_configs.py --no-cpu --only "$cfg" > gpurun_out/r2e_ncu_$tag.log 2>&1
  echo "$cfg capture rc=$?"
done
ls -la gpurun_out | grep r2e
