#!/bin/bash
# round 2, call V: 3x3 network compare-exchange mix re-swept after the ALU relief
mkdir -p gpurun_out
for lib in librv_b200.so librv_b200_k3_12.so librv_b200_k3_35.so librv_b200_k3_47.so librv_b200.so librv_b200_k3_12.so; do
  echo "== $lib"
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "k3" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2v_variants.txt
done
