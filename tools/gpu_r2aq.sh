#!/bin/bash
# round 2, call AQ: LAB histogram pass with the gamma table replicated per lane (conflict-free gamma reads)
mkdir -p gpurun_out
RV_B200_LIB=librv_b200_hrep.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "luma or lut or chain or clahe or golden or sha or config2 or gate or tiny or large_tile or colour" > gpurun_out/r2aq_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2aq_pytest.log
for lib in librv_b200.so librv_b200_hrep.so librv_b200.so librv_b200_hrep.so; do
  RV_B200_LIB=$lib timeout 300 python tools/exp_kernel_shares.py 2>&1 | grep "^{" | grep LAB | tee -a gpurun_out/r2aq_shares.jsonl
done
for lib in librv_b200.so librv_b200_hrep.so librv_b200.so librv_b200_hrep.so; do
  echo "== $lib" | tee -a gpurun_out/r2aq_variants.txt
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "LAB" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2aq_variants.txt
done
