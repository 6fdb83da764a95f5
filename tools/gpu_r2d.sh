#!/bin/bash
# round 2, call D: debug-bounds build under the parity suite; LAB tuning variants
mkdir -p gpurun_out
RV_B200_LIB=librv_b200_dbg.so timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_dbg.log 2>&1; echo "pytest(dbg) rc=$?"; tail -6 gpurun_out/r2d_pytest_dbg.log
for lib in librv_b200.so librv_b200_mulhi.so librv_b200_fma3_11.so librv_b200_fma3_34.so librv_b200_fma5_23.so; do
  echo "== $lib"
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only LAB 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
    elif line: print(line)
" | tee -a gpurun_out/r2d_variants.txt
done
RV_B200_LIB=librv_b200_fma5_23.so timeout 300 python tests/perf/bench_configs.py --no-cpu --only headline | tee -a gpurun_out/r2d_variants.txt
