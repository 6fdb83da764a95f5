#!/bin/bash
# round 2, call T: channel-fastest median tasks + 2-word plane skew -- parity, A/B timing, conflict counters
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_pytest.log
for lib in librv_b200.so librv_b200_nocfast.so librv_b200.so librv_b200_nocfast.so; do
  echo "== $lib"
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "k" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        if d['ksize'] in (3,5): print('  %-36s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2')))
" | tee -a gpurun_out/r2t_variants.txt
done
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --stream-seconds 0"
timeout 300 $B > gpurun_out/r2t_plain.log 2>&1 && \
timeout 900 ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,gpu__time_duration.sum --clock-control none -k regex:k_chain -s 8 -c 1 --csv --log-file gpurun_out/r2t_conflicts.csv $B > gpurun_out/r2t_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r2t_conflicts.csv | cut -d, -f5,13,15
