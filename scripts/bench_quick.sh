#!/bin/bash
# quick GPU check: parity tests, then 64k bench lines (fractal + tilted plane), phases printed
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for kind in 0 2; do
python bench.py --size 65536 --kind $kind --steps 3 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/bq_$kind.err | tee gpurun_out/bq_$kind.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print(round(d['value'],2), round(d['ms_per_step'],2), r.get('phases_ms_per_launch') or r.get('phases_ms_per_step'), d['parity'])"
done
