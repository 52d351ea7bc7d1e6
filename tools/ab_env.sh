#!/bin/bash
# tools/ab_env.sh VAR "shape-or-workload args for sweep.py" — time sweep.py with and without an environment knob
V=$1; shift
for on in 0 1; do
  if [ $on = 1 ]; then export $V=1; else unset $V; fi
  echo "== $V=${!V:-unset}"
  python tools/sweep.py "$@" 2>&1 | grep -o "\"workload\": \"[A-Za-z0-9]*\".*\"x\": \"[a-z]*\", \"us\": [0-9.]*" | sed 's/"M".*"x"/"x"/'
done
