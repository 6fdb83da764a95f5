#!/bin/bash
# k_chain L2 prefetch distance sweep (CTAs per SM ahead; 0 = off), device-resident bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for d in "$@"; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --prefetch $d 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('prefetch $d: value %.0f fps  ms/step %.3f  chain_ms %.3f hist_ms %.3f' % (d['value'], d['ms_per_step'], r['kernel_ms_per_step']['k_chain'], r['kernel_ms_per_step']['k_luma_hist']))"
done
