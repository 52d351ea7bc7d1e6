#!/bin/bash
# tools/profile_tail.sh — launch list + full captures of c4 after the tail launch was added (subset of profile_round.sh)
set -u
R=${1:-rX}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
run() { echo "== $*"; "$@" > $O/${R}_last.log 2>&1 || { echo "FAILED: $*"; tail -5 $O/${R}_last.log; }; }
run python bench.py --steps 3 --warmup 1 --no-others --no-cpu-baseline --no-builder
run $NCU --metrics gpu__time_duration.sum -k regex:"dense_tc|split_tiles|gather|code_gemv" -c 400 --csv --log-file $O/${R}_c4_launches.csv \
    python bench.py --steps 3 --warmup 1 --no-others --no-cpu-baseline --no-builder
run python tools/sweep.py --workloads c4 --algos dense_tc --x real --steps 2
run $NCU --set full --import-source on -k regex:dense_tc -s 8 -c 2 -f -o $O/${R}_c4_dense_tc_real_tail python tools/sweep.py --workloads c4 --algos dense_tc --x real --steps 2
ls -la $O/${R}_*
