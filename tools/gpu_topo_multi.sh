#!/bin/bash
# host topology next to the GPUs + N-GPU bench lines (device-resident value and end-to-end from pinned host memory)
mkdir -p gpurun_out
{ nvidia-smi topo -m; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread"; numactl -H 2>/dev/null | head -20;
  for d in /sys/bus/pci/devices/*; do [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ -f $d/local_cpulist ] && echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist) class=$(cat $d/class)"; done;
  echo "affinity: $(python -c 'import os;print(len(os.sched_getaffinity(0)))') cores"; free -g | head -2; } > gpurun_out/topo.txt 2>&1
tail -25 gpurun_out/topo.txt
for n in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/multi_n$n.log 2>&1
  echo "N=$n rc=$?"; grep '^{' gpurun_out/multi_n$n.log | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print('  value %.0f fps, e2e %.0f fps, ms/step %.3f affinity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['host_affinity']))"
done
