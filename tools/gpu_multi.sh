#!/bin/bash
# N-GPU bench (torchrun) for N in the arguments + reference arm
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in "$@"; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/scale_n1.log 2>&1
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/scale_n$n.log 2>&1
  fi
  echo "N=$n rc=$?"; grep '^{' gpurun_out/scale_n$n.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('  value %.0f fps, e2e %.0f fps, ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']))"
done
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/ref_arm.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/ref_arm.log | cut -c1-600
