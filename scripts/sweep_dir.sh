#!/bin/bash
# direction kernel tuning sweep (runs on the GPU box: rebuilds the library per configuration)
for cfg in "-DOFL_DIR_RB=4 -DOFL_DIR_STAGES=4" "-DOFL_DIR_RB=8 -DOFL_DIR_STAGES=4" "-DOFL_DIR_RB=8 -DOFL_DIR_STAGES=3" "-DOFL_DIR_RB=6 -DOFL_DIR_STAGES=4" "-DOFL_DIR_RB=4 -DOFL_DIR_STAGES=6" "-DOFL_DIR_RB=2 -DOFL_DIR_STAGES=8"; do
  OFL_NVCC_EXTRA="$cfg" python -m overflow_b200.build --force > /dev/null 2>&1 || { echo "$cfg: build failed"; continue; }
  python bench.py --size 32768 --steps 3 --warmup 2 --no-e2e --no-cpu --no-check 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$cfg', r['phases_ms_per_step']['direction'])"
done
python -m overflow_b200.build --force > /dev/null 2>&1
