#!/bin/bash
for c in 1 2 4 8 16; do
  echo "chunk=$c"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --chunk $c 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  e2e %.0f fps  value %.0f' % (d['e2e']['value'], d['value']))"
done
