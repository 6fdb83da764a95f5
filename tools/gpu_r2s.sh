#!/bin/bash
# round 2, call S: end-to-end (pinned in, pinned out) against the number of pipeline streams and the chunk size
mkdir -p gpurun_out
for lib in librv_b200.so librv_b200_np2.so librv_b200_np4.so librv_b200_np6.so; do
  for chunk in 0 1 2 6 12; do
    RV_B200_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --stream-seconds 0 --chunk $chunk 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); print('$lib chunk=$chunk  e2e %.0f  e2e_tensor %.0f  e2e_keep %.0f' % (d['e2e']['value'], d['e2e_tensor']['value'], d['e2e_keep']['value']))
" | tee -a gpurun_out/r2s_e2e_sweep.txt
  done
done
