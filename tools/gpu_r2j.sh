#!/bin/bash
# round 2, call J: plane skew (bank conflicts of the median loads) -- parity, A/B timing, conflict counters
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log
for lib in librv_b200.so librv_b200_skew0.so librv_b200.so librv_b200_skew0.so; do
  echo "== $lib"
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "k" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2j_variants.txt
done
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --stream-seconds 0"
timeout 300 $B > gpurun_out/r2j_plain.log 2>&1 && \
timeout 900 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,gpu__time_duration.sum --clock-control none -k regex:k_chain -s 8 -c 1 --csv --log-file gpurun_out/r2j_conflicts.csv $B > gpurun_out/r2j_ncu.log 2>&1; echo "ncu rc=$?"; tail -4 gpurun_out/r2j_conflicts.csv
