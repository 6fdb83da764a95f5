#!/bin/bash
# round 2, call I (N GPUs): rows-only tensor download -- parity test on one GPU, bench legs at N
N=${1:-8}
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_rows_only or results_that_stay or contexts_are_independent" 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --stream-seconds 0 > gpurun_out/r2i_bench_n$N.json 2> gpurun_out/r2i_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
for line in open('gpurun_out/r2i_bench_n$N.json'):
    if line.startswith('{'):
        j=json.loads(line)
        print({k:(round(j[k]['value']) if isinstance(j[k],dict) else j[k]) for k in ('value','sustained','e2e','e2e_tensor','e2e_keep')})
PY
tail -3 gpurun_out/r2i_bench_n$N.err
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 20 --warmup 5 --stream-seconds 0 --no-cpu > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; python - <<PY
import json
j=json.load(open('gpurun_out/r2i_bench_n1.json'))
print('N=1', {k:(round(j[k]['value']) if isinstance(j[k],dict) else j[k]) for k in ('value','e2e','e2e_tensor','e2e_keep')})
PY
