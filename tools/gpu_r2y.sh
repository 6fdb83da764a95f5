#!/bin/bash
# round 2, call Y: k = 7 / 9 medians -- wider groups (single-row networks, M 6) and two output rows per task (generic hierarchical
# two-row networks, M 4 / 6), with and without a 128-register cap (2 CTAs per SM)
mkdir -p gpurun_out
VAR="librv_b200_2r_m4c2.so librv_b200_2r_m6c2.so librv_b200_2r_m6c1.so librv_b200_2r_m4c1.so librv_b200_k7m6k9m6c2.so librv_b200_k9m6c2.so"
for lib in $VAR; do
  RV_B200_LIB=$lib timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "median_alone or chain_vs_oracle or size_independent" > gpurun_out/r2y_pytest_$lib.log 2>&1; echo "$lib pytest rc=$?"; tail -1 gpurun_out/r2y_pytest_$lib.log
done
for lib in librv_b200.so $VAR librv_b200.so librv_b200_2r_m6c2.so librv_b200_2r_m4c2.so; do
  echo "== $lib" | tee -a gpurun_out/r2y_variants.txt
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "1080p YCrCb k" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2y_variants.txt
done
