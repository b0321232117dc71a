#!/bin/bash
# Tuning sweep for the flat-resolution tile sweeps (run on the GPU box: rebuilds the library per variant).
#   bash scripts/sweep_flats.sh "<nvcc defines>" ...      e.g. "-DOFL_FL_SWEEP_CTAS=12" "-DOFL_FL_SWEEP_THREADS=256 -DOFL_FL_SWEEP_CTAS=8"
run() { for s in 8192 16384; do timeout 120 python scripts/bench_flats.py --size $s --kind 1 --relief $((s/41)) --steps 3 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['size'], round(d['ms'],2), {k:round(v,2) for k,v in d['phases_ms'].items()})"; done; }
echo "== default"; run
for v in "$@"; do
  echo "== $v"; OFL_NVCC_EXTRA="$v" python -m overflow_b200.build --force >/dev/null 2>&1 || { echo "build failed"; continue; }; run
done
python -m overflow_b200.build --force >/dev/null 2>&1
