#!/bin/bash
# tools/e2e_multi_ab.sh — host-buffer call at N GPUs with and without binding each rank to its GPU's CPUs
G=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m 2>/dev/null | head -14
nproc
for off in 0 1; do
  if [ $off = 1 ]; then export TSG_NO_NUMA_BIND=1; else unset TSG_NO_NUMA_BIND; fi
  python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port $((29600 + off)) --nproc-per-node $G bench.py --gpus $G --no-others 2>gpurun_out/e2e_ab_$off.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('bind off' if $off else 'bind on ', d['config'].get('host_affinity'), 'ms_per_step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],3), 'other path', round(d['e2e']['other_host_path']['ms_per_step'],3))"
done
