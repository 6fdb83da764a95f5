#!/bin/bash
# round 2, call O: I2F blend and compare-exchange mixes after the ALU-pipe relief -- timing only (parity asserted per config)
mkdir -p gpurun_out
for lib in librv_b200.so librv_b200_i2f.so librv_b200_f49.so librv_b200_f37.so librv_b200_f25.so librv_b200_f59.so librv_b200.so librv_b200_i2f.so; do
  echo "== $lib"
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "k" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        if d['ksize'] in (3,5) and d['shape'][0] != 2160: print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2o_variants.txt
done
