#!/bin/bash
# direction kernel tuning sweep (runs on the GPU box: rebuilds the library per configuration)
for cfg in "$@"; do
  OFL_NVCC_EXTRA="$cfg" python -m overflow_b200.build --force --verbose 2>&1 | grep -A3 "direction_kernel" | grep -E "Used|spill" | tr '\n' ' '
  python bench.py --size 65536 --steps 3 --warmup 2 --no-e2e --no-cpu --no-check --no-other --no-flats 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$cfg', r['phases_ms_per_step']['direction'])"
done
python -m overflow_b200.build --force > /dev/null 2>&1
