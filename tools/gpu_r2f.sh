#!/bin/bash
# round 2, call F (8 GPUs): bench legs at N = 8 (device-resident, e2e, e2e_tensor, e2e_keep, 8 paced streams) + result placement experiment
N=${1:-8}
mkdir -p gpurun_out
python -c "import rvb200; print(rvb200.kernel_source_hash())" > gpurun_out/r2f_hash.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
for line in open('gpurun_out/r2f_bench_n$N.json'):
    if line.startswith('{'):
        j=json.loads(line)
        print({k:(round(j[k]['value']) if isinstance(j[k],dict) else j[k]) for k in ('value','sustained','e2e','e2e_tensor','e2e_keep')}, j['streams'], j['clocks'])
PY
tail -3 gpurun_out/r2f_bench_n$N.err
STEPS=8 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/exp_d2h.py > gpurun_out/r2f_d2h_n$N.json 2> gpurun_out/r2f_d2h_n$N.err; echo "d2h rc=$?"; grep '^{' gpurun_out/r2f_d2h_n$N.json
