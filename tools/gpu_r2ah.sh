#!/bin/bash
# round 2, call AH: fog synthesis -- faster float bilateral (table + ex2, no I2F) and lighter host side, against the first version
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fog.py -m gpu -x -q > gpurun_out/r2ah_pytest_fog.log 2>&1; echo "fog tests rc=$?"; tail -3 gpurun_out/r2ah_pytest_fog.log
RV_B200_LIB=librv_b200_fogv1.so timeout 300 python tools/exp_fog_ab.py gpurun_out/r2ah_fog_v1.npz
timeout 300 python tools/exp_fog_ab.py gpurun_out/r2ah_fog_v2.npz
python - <<'PY'
import numpy as np
a=np.load('gpurun_out/r2ah_fog_v1.npz'); b=np.load('gpurun_out/r2ah_fog_v2.npz')
worst=0
for k in a.files:
    x,y=a[k].astype(np.float64),b[k].astype(np.float64)
    d=np.abs(x-y)
    if k.endswith('_t') or k.endswith('_A'):
        print(f"{k:28s} float map: max abs diff {d.max():.3e}  rms {np.sqrt((d*d).mean()):.3e}")
    else:
        print(f"{k:28s} u8 frame : differing pixels {int((d>0).sum())} of {d.size}, max {int(d.max())}, mean abs {d.mean():.2e}")
PY
timeout 600 python tests/perf/bench_fog.py > gpurun_out/r2ah_fog.json 2> gpurun_out/r2ah_fog.err; cat gpurun_out/r2ah_fog.json
rm -f gpurun_out/r2ah_fog_v1.npz gpurun_out/r2ah_fog_v2.npz
