#!/bin/bash
# Round profile at the benchmark size: plain run (must exit 0), launch list, one full capture of the
# timed step's three main kernels, DRAM bytes of the solve's kernels.  Outputs under gpurun_out/;
# summarise with scripts/ncu_summary.py.
set -e
TAG=${1:-r02}
CMD="python bench.py --size 65536 --steps 1 --warmup 1 --no-e2e --no-cpu --no-check --no-flats --no-other"
$CMD > gpurun_out/plain_$TAG.json 2> gpurun_out/plain_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --import-source on --clock-control none -k 'regex:direction_kernel|acc_tile_kernel|acc_final_kernel' --launch-skip 3 -c 3 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:pj_' -c 200 --csv --log-file gpurun_out/solve_$TAG.csv $CMD > gpurun_out/ncu_s_$TAG.log 2>&1
tail -2 gpurun_out/ncu_f_$TAG.log
