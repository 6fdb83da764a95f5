#!/bin/bash
# round 2, call AK: LAB inverse arguments with the constants folded (ft table + 10484) and the a / b differences as multiply-adds
mkdir -p gpurun_out
for lib in librv_b200_ftb.so librv_b200_ftb_ab.so; do
RV_B200_LIB=$lib timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "colour or chain or clahe or golden or sha or config2 or properties" > gpurun_out/r2ak_pytest_$lib.log 2>&1; echo "$lib pytest rc=$?"; tail -1 gpurun_out/r2ak_pytest_$lib.log
done
for lib in librv_b200.so librv_b200_ftb.so librv_b200_ftb_ab.so librv_b200.so librv_b200_ftb.so librv_b200_ftb_ab.so; do
  echo "== $lib" | tee -a gpurun_out/r2ak_variants.txt
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "LAB" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2ak_variants.txt
done
