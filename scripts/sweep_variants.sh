#!/bin/bash
# A/B of prebuilt library variants (variants/*.so, built in the dev container with OFL_NVCC_EXTRA) on the GPU box:
# 64k^2 bench lines with phases and parity per variant, then the accumulation-facing parity tests on the fastest one.
#   scripts/sweep_variants.sh "base epi epi_blk" "epi_blk"      (second list: variants that also run the tilted plane)
cp overflow_b200/liboverflow_b200.so /tmp/lib_keep.so
best=""; best_ms=999999
for v in $1; do
  cp variants/$v.so overflow_b200/liboverflow_b200.so
  kinds="0"; for w in $2; do [ "$w" = "$v" ] && kinds="0 2"; done
  for kind in $kinds; do
    python bench.py --size 65536 --kind $kind --steps 5 --warmup 3 --no-e2e --no-cpu --no-other --no-flats 2>gpurun_out/sv_${v}_$kind.err > gpurun_out/sv_${v}_$kind.json
    python - "$v" "$kind" <<'P'
import json,sys
v,kind=sys.argv[1:3]
try:
    d=json.load(open(f"gpurun_out/sv_{v}_{kind}.json")); p=d["roofline"]["phases_ms_per_step"]; q=d["parity"]
    print(v,"kind",kind,"step",round(d["ms_per_step"],3),"A",p["acc_tile_a"],"solve",p["acc_solve"],"B",p["acc_tile_b"],"dir",p["direction"],
          "viol",q["accumulation_recurrence_violations"],"win",q["direction_windows_vs_oracle"])
except Exception as e:
    print(v,"kind",kind,"FAILED",e)
P
  done
done
# parity tests of the accumulation-facing entry points on the fastest variant (fractal, pass A, zero violations)
best=$(python - <<'P'
import json,glob,os
best=None
for f in glob.glob("gpurun_out/sv_*_0.json"):
    try:
        d=json.load(open(f))
        if d["parity"]["accumulation_recurrence_violations"]!=0: continue
        a=d["roofline"]["phases_ms_per_step"]["acc_tile_a"]; v=os.path.basename(f)[3:-7]
        if os.path.exists(f"variants/{v}.so") and (best is None or a<best[0]): best=(a,v)
    except Exception: pass
print(best[1] if best else "")
P
)
if [ -n "$best" ]; then
  echo "fastest: $best"
  cp variants/$best.so overflow_b200/liboverflow_b200.so
  timeout 90 python -m pytest tests/test_gpu_accumulation.py tests/test_gpu_tiles.py tests/test_gpu_strips.py tests/test_gpu_routing.py -x -q 2>&1 | tail -3
fi
cp /tmp/lib_keep.so overflow_b200/liboverflow_b200.so
