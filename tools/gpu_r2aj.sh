#!/bin/bash
# round 2, call AJ: LAB inverse -- clamp before the shift so that shift + table base fuse into LEA.HI (RV_LAB_LEA_IG)
mkdir -p gpurun_out
RV_B200_LIB=librv_b200_leaig.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2aj_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2aj_pytest.log
for lib in librv_b200.so librv_b200_leaig.so librv_b200.so librv_b200_leaig.so; do
  echo "== $lib" | tee -a gpurun_out/r2aj_variants.txt
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "LAB" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2aj_variants.txt
done
