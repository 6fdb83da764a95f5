#!/bin/bash
# ncu capture of k_chain<0,3> (the reference's default chain: YCrCb, k3) on 64 x 720p
mkdir -p gpurun_out
cat > /tmp/k3.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, rvb200
from rvb200 import synth
ctx = rvb200.Context(0)
pool = synth.frame_pool(720, 1280, 4, base_seed=5)
host = np.stack([pool[i % 4] for i in range(64)])
d_in = torch.from_numpy(host).cuda(); d_out = torch.empty_like(d_in)
p = rvb200.Params.make("YCrCb", 2.0, 8, 3)
for _ in range(4):
    ctx.chain_device(d_in.data_ptr(), d_out.data_ptr(), 64, 720, 1280, p)
print("ok")
PY
timeout 120 python /tmp/k3.py && timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_chain -s 2 -c 1 -f -o gpurun_out/prof_chain_k3 python /tmp/k3.py > gpurun_out/ncu_k3.log 2>&1
echo "rc=$?"
