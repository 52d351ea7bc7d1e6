#!/bin/bash
# tools/profile_round.sh — the ncu captures kept under profiles/ for one round (run on the GPU box:
#   gpurun -- 'bash tools/profile_round.sh r2').  Every capture is taken only after the same
# command has exited 0 without ncu; numbers printed by a run under ncu are never bench values.
set -u
R=${1:-rX}
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
run() { echo "== $*"; "$@" > $O/${R}_last.log 2>&1 || { echo "FAILED: $*"; tail -5 $O/${R}_last.log; }; }
# 1. launch list of the contract line's workload (c4, both regimes)
run python bench.py --steps 3 --warmup 1 --no-others --no-cpu-baseline --no-builder
run $NCU --metrics gpu__time_duration.sum -k regex:"dense_tc|split_tiles|gather|code_gemv" -c 400 --csv --log-file $O/${R}_c4_launches.csv \
    python bench.py --steps 3 --warmup 1 --no-others --no-cpu-baseline --no-builder
# 2. full captures, one launch each (after warm-up launches)
full() { name=$1; kern=$2; skip=$3; shift 3; run "$@"; run $NCU --set full --import-source on -k regex:"$kern" -s $skip -c 1 -f -o $O/${R}_${name} "$@"; }
full c4_dense_tc_real   dense_tc    4 python tools/sweep.py --workloads c4 --algos dense_tc --x real --steps 2
full c4_split_real      split_tiles 4 python tools/sweep.py --workloads c4 --algos dense_tc --x real --steps 2
full c4_dense_tc_int    dense_tc    4 python tools/sweep.py --workloads c4 --algos dense_tc --x int --steps 2
full c4_split_int       split_tiles 4 python tools/sweep.py --workloads c4 --algos dense_tc --x int --steps 2
full c3_dense_tc_real   dense_tc    6 python tools/sweep.py --workloads c3 --algos dense_tc --x real --steps 2
full c5a_dense_tc_int   dense_tc    4 python tools/sweep.py --workloads c5a --algos dense_tc --x int --steps 2
full c2_code_gemv       code_gemv  20 python tools/sweep.py --workloads c2 --algos code_gemv --x int --steps 4
full c2_gather16        gather     20 python tools/sweep.py --workloads c2 --algos gather --x int --steps 4
full gather16_8192x28672_s8 gather  6 python tools/sweep.py --shape 1,8192,28672,8 --algos gather --steps 2
full build_encode       encode_planes_ballot 1 python tools/build_probe.py c4 1
full build_emit         emit_indices         1 python tools/build_probe.py c4 1
full build_codes        tile_codes           1 python tools/build_probe.py c4 1
run $NCU --metrics gpu__time_duration.sum -k regex:"encode|scan|emit|tile_codes" --csv --log-file $O/${R}_build_c4_launches.csv python tools/build_probe.py c4 1
ls -la $O/${R}_*
