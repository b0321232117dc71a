#!/bin/bash
# pass A tuning sweep (runs on the GPU box: rebuilds the library per configuration)
for cfg in "$@"; do
  OFL_NVCC_EXTRA="$cfg" python -m overflow_b200.build --force > /dev/null 2>&1 || { echo "$cfg: build failed"; continue; }
  for kind in ${KINDS:-0 2}; do
  python bench.py --size 32768 --kind $kind --steps 3 --warmup 2 --no-e2e --no-cpu --no-other --no-flats 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$cfg kind $kind', r['phases_ms_per_step'], d['parity']['accumulation_recurrence_violations'])"
  done
done
python -m overflow_b200.build --force > /dev/null 2>&1
