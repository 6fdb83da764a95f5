#!/bin/bash
# round 2, call K: 8 (and 16) paced 1080p@30 camera streams on ONE GPU, one context + thread per stream; saturating replay
mkdir -p gpurun_out
for S in 8 16; do
  timeout 300 python tools/bench_streams.py --streams $S --fps 30 --seconds 12 --batch 1 > gpurun_out/r2k_streams_$S.json 2> gpurun_out/r2k_streams_$S.err; echo "streams=$S rc=$?"
  python - <<PY
import json
j=json.load(open('gpurun_out/r2k_streams_$S.json'))
ps=j['per_stream']
print('streams', j['streams'], 'total_fps', j['total_fps'], 'fps min', min(p['fps'] for p in ps), 'p50 max', max(p['lat_ms_p50'] for p in ps), 'p99 max', max(p['lat_ms_p99'] for p in ps), 'max', max(p['lat_ms_max'] for p in ps))
PY
done
timeout 300 python tools/bench_streams.py --streams 4 --fps 0 --seconds 2 --batch 16 > gpurun_out/r2k_streams_sat.json 2> gpurun_out/r2k_streams_sat.err; echo "saturating rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2k_streams_sat.json')); print('saturating: streams', j['streams'], 'batch', j['batch'], 'total fps', j['total_fps'])"
