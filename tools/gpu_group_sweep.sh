#!/bin/bash
for g in 0 8 16 32; do
  echo "group=$g"; timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --group $g 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('  value %.0f fps  kernels %s launches %d' % (d['value'], {k: round(v,4) for k,v in r['kernel_ms_per_step'].items()}, d['gpu_launches']))"
done
