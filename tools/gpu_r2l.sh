#!/bin/bash
# round 2, call L: plane values without the 0x6400 bias (subnormal halves on the FMA pipe) -- parity + A/B timing; stream records
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2n_pytest_nobias.log 2>&1; echo "pytest(default build) rc=$?"; tail -3 gpurun_out/r2n_pytest_nobias.log
for lib in librv_b200.so librv_b200_nodp2a.so librv_b200.so librv_b200_nodp2a.so; do
  echo "== $lib"
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "k" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2n_variants.txt
done
