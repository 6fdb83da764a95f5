#!/bin/bash
# round 2, call P: LAB table indexes via IDP.2A -- parity (incl. all 2^24 colours) + A/B timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2p_pytest.log
for lib in librv_b200.so librv_b200_nolabdp.so librv_b200.so librv_b200_nolabdp.so; do
  echo "== $lib"
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "LAB" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2p_variants.txt
done
