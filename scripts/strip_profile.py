#!/usr/bin/env python
"""All N strips of the 64k x 64k benchmark raster on ONE GPU (loop-back exchange), for a per-kernel launch list of what
each rank of an N-GPU run executes:

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/strips8.csv \
        python scripts/strip_profile.py --strips 8
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from overflow_b200 import strips

ap = argparse.ArgumentParser()
ap.add_argument("--strips", type=int, default=8)
ap.add_argument("--size", type=int, default=65536)
ap.add_argument("--kind", type=int, default=0)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--only", type=int, default=-1, help="run the compute of this strip only (the others just exist)")
a = ap.parse_args()
pipes = [strips.StripPipeline(a.size, a.size, r, a.strips, device="cuda:0") for r in range(a.strips)]
for p in pipes:
    p.load_synthetic(seed=0, kind=a.kind, holes_permille=5)
for _ in range(a.steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    strips.step_in_process(pipes)
    torch.cuda.synchronize()
    print("step", (time.perf_counter() - t0) * 1e3, "ms for", a.strips, "strips in sequence")
