#!/bin/bash
# round 2, call AN: fused letterbox store phase without divisions (2-D thread mapping, v * (1/255) instead of v / 255)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_properties.py -m gpu -x -q -k "letterbox or tensor or results_that_stay or chunk_schedules or gpu_equals" > gpurun_out/r2an_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2an_pytest.log
for lib in librv_b200_prev.so librv_b200.so librv_b200_prev.so librv_b200.so; do
  echo "== $lib" | tee -a gpurun_out/r2an_variants.txt
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "C" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        print('  %-72s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2', d.get('bit_exact_vs_oracle'))))
" | tee -a gpurun_out/r2an_variants.txt
done
