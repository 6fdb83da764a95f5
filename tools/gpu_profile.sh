#!/bin/bash
# bench + ncu launch list + ncu full capture of the top kernel (one GPU). usage: tools/gpu_profile.sh <tag>
TAG=${1:-cur}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.log
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 8 -c 2 -f -o gpurun_out/prof_chain_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_luma_hist -s 8 -c 1 -f -o gpurun_out/prof_hist_$TAG \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full_hist_$TAG.log 2>&1
echo "hist capture rc=$?"
ls -la gpurun_out | tail -12
