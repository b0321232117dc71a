run() { timeout 120 python scripts/bench_flats.py --size 8192 --kind 1 --relief 200 --steps 2 2>&1 | tail -1| cut -c 1-120,560-900; timeout 120 python scripts/bench_flats.py --size 16384 --kind 1 --relief 400 --steps 2 2>&1 | tail -1 | cut -c 1-120,560-900; }
timeout 600 python -m pytest tests/test_gpu_flats.py -x -q 2>&1 | tail -2
echo "== default"; run
for v in "-DOFL_FL_SWEEP_CTAS=6" "-DOFL_FL_SWEEP_THREADS=128 -DOFL_FL_QCAP=1280 -DOFL_FL_SWEEP_CTAS=10" "-DOFL_FL_SWEEP_THREADS=128 -DOFL_FL_QCAP=1280"; do OFL_NVCC_EXTRA="$v" python -m overflow_b200.build --force > /dev/null 2>&1; echo "== $v"; run; done
python -m overflow_b200.build --force > /dev/null 2>&1
