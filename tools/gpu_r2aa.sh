#!/bin/bash
# round 2, call AA: share of the compare-exchanges of the 7x7 / 9x9 two-row networks that takes the FMA-pipe form (shipped 1/2)
mkdir -p gpurun_out
for lib in librv_b200.so librv_b200_f79_1_3.so librv_b200_f79_2_5.so librv_b200_f79_3_7.so librv_b200_f79_3_5.so librv_b200.so librv_b200_f79_2_5.so librv_b200_f79_3_7.so; do
  echo "== $lib" | tee -a gpurun_out/r2aa_variants.txt
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "1080p YCrCb k" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        if d['ksize'] in (7, 9): print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2aa_variants.txt
done
