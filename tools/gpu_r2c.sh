#!/bin/bash
# round 2, call C (N GPUs): where full-frame results land at N ranks + bench legs at N
N=${1:-4}
mkdir -p gpurun_out
STEPS=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/exp_d2h.py > gpurun_out/r2c_d2h_n$N.json 2> gpurun_out/r2c_d2h_n$N.err; echo "d2h rc=$?"; cat gpurun_out/r2c_d2h_n$N.json; tail -3 gpurun_out/r2c_d2h_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --stream-seconds 12 > gpurun_out/r2c_bench_n$N.json 2> gpurun_out/r2c_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
j=json.load(open('gpurun_out/r2c_bench_n$N.json'))
print({k:(round(j[k]['value']) if isinstance(j[k],dict) else j[k]) for k in ('value','sustained','e2e','e2e_tensor','e2e_keep')}, j['streams'])
PY
tail -3 gpurun_out/r2c_bench_n$N.err
