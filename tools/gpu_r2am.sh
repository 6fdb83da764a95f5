#!/bin/bash
# round 2, call AM: ncu captures of the LAB kernels after calls AJ / AK (k_chain<LAB,3>, k_chain<LAB,5>, k_luma_hist<LAB>)
mkdir -p gpurun_out
python -c "import rvb200, json; json.dump(rvb200.kernel_sass_hashes(), open('gpurun_out/r2am_sass.json','w'), indent=1); print(rvb200.kernel_source_hash())" > gpurun_out/r2am_hash.txt
for cfg in "1080p LAB k3" "1080p LAB k5"; do
  tag=$(echo "$cfg" | tr ' ' '_')
  timeout 300 python tests/perf/bench_configs.py --no-cpu --only "$cfg" > gpurun_out/r2am_plain_$tag.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 3 -c 1 -f -o gpurun_out/r2am_prof_chain_$tag python tests/perf/bench_configs.py --no-cpu --only "$cfg" > gpurun_out/r2am_ncu_$tag.log 2>&1
  echo "$cfg capture rc=$?"; grep gpu_fps gpurun_out/r2am_plain_$tag.log | cut -c1-170
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_luma_hist -s 3 -c 1 -f -o gpurun_out/r2am_prof_hist_lab python tests/perf/bench_configs.py --no-cpu --only "1080p LAB k3" > gpurun_out/r2am_ncu_hist_lab.log 2>&1; echo "hist capture rc=$?"
ls -la gpurun_out | grep r2am
