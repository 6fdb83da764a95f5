#!/bin/bash
# round 2, call AB: new 0-1 frame test, deep property run (RV_PROP_SCALE=4: bigger frames, 480 examples), ncu capture of the two-row k9 / k7 kernels
mkdir -p gpurun_out
python -c "import rvb200, json; json.dump(rvb200.kernel_sass_hashes(), open('gpurun_out/r2ab_sass.json','w'), indent=1); print(rvb200.kernel_source_hash())" > gpurun_out/r2ab_hash.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "zero_one or median_alone" > gpurun_out/r2ab_pytest01.log 2>&1; echo "0-1 frames rc=$?"; tail -1 gpurun_out/r2ab_pytest01.log
RV_PROP_SCALE=4 timeout 1200 python -m pytest tests/test_properties.py -m gpu -x -q > gpurun_out/r2ab_props.log 2>&1; echo "props x4 rc=$?"; tail -1 gpurun_out/r2ab_props.log
for cfg in "1080p YCrCb k9" "1080p YCrCb k7"; do
  tag=$(echo "$cfg" | tr ' ' '_')
  timeout 300 python tests/perf/bench_configs.py --no-cpu --only "$cfg" > gpurun_out/r2ab_plain_$tag.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 3 -c 1 -f -o gpurun_out/r2ab_prof_chain_$tag python tests/perf/bench_configs.py --no-cpu --only "$cfg" > gpurun_out/r2ab_ncu_$tag.log 2>&1
  echo "$cfg capture rc=$?"; grep gpu_fps gpurun_out/r2ab_plain_$tag.log | cut -c1-160
done
ls -la gpurun_out | grep r2ab
