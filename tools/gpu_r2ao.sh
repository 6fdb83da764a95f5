#!/bin/bash
# round 2, call AO: k_chain instantiations without the fused letterbox phase for plain jobs (RV_CHAIN_SPLIT_LB)
mkdir -p gpurun_out
RV_B200_LIB=librv_b200_split.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chain or letterbox or tensor or golden or sha" > gpurun_out/r2ao_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r2ao_pytest.log
for lib in librv_b200.so librv_b200_split.so librv_b200.so librv_b200_split.so; do
  echo "== $lib" | tee -a gpurun_out/r2ao_variants.txt
  RV_B200_LIB=$lib timeout 300 python tests/perf/bench_configs.py --no-cpu --only "C" 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line)
        print('  %-44s %9.1f fps  exact=%s' % (d['config'], d['gpu_fps'], d.get('bit_exact_vs_cv2', d.get('bit_exact_vs_oracle'))))
" | tee -a gpurun_out/r2ao_variants.txt
done
