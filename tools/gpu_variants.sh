#!/bin/bash
# parity tests on the default build, then a device-resident bench for each tuning build named on the command line
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for lib in "$@"; do
  echo "== $lib"
  RV_B200_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('value %.0f fps  ms/step %.3f  chain_ms %.3f hist_ms %.3f lut_ms %.3f  frac %.4f clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms_per_step']['k_chain'], r['kernel_ms_per_step']['k_luma_hist'], r['kernel_ms_per_step']['k_build_lut'], r['frac'], d['clocks']['sm_mhz']))
    else: print(line)
"
done
