#!/bin/bash
# tools/scale_round.sh — strong scaling of c4 (and c3/c5a/c5b at the largest N) on one multi-GPU box:
#   gpurun --gpus 8 -- 'bash tools/scale_round.sh r2'   -> gpurun_out/<round>_bench_{N}gpu.json, ..._others_{N}gpu.jsonl
set -u
R=${1:-rX}
O=gpurun_out
mkdir -p $O
G=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
port=29500
for n in 8 4 2; do
  [ $n -le $G ] || continue
  port=$((port + 1))
  timeout 600 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --no-others > $O/${R}_bench_${n}gpu.json 2> $O/${R}_bench_${n}gpu.err
  echo "c4 N=$n rc=$?"; cut -c1-400 $O/${R}_bench_${n}gpu.json
done
n=$G
: > $O/${R}_bench_others_${n}gpu.jsonl
for w in c3 c5a c5b; do
  port=$((port + 1))
  timeout 600 $TR --nproc-per-node $n --master-port $port bench.py --gpus $n --workload $w --no-others >> $O/${R}_bench_others_${n}gpu.jsonl 2> $O/${R}_bench_others_${w}.err
  echo "$w N=$n rc=$?"
done
cut -c1-300 $O/${R}_bench_others_${n}gpu.jsonl
