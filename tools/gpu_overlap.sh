#!/bin/bash
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for o in 0 2 4 8; do
  echo "overlap=$o"; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --overlap $o 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('  value %.0f fps ms/step %.4f  kernels %s launches %d' % (d['value'], d['ms_per_step'], {k: round(v,4) for k,v in r['kernel_ms_per_step'].items()}, d['gpu_launches']))
    elif 'rror' in l: print(l.strip())"
done
