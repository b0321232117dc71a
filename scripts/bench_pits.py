"""Time single-cell pit breaching (ofl_breach_single_cell_pits_f32, device buffers) on a synthetic DEM.

    python scripts/bench_pits.py [--size 16384] [--kind 0] [--steps 3]

Prints one JSON line: ms per call (CUDA events), Gcells/s, algorithmic GB/s at 5 B/cell (DEM read 4, unsolved raster 1
written; the breached cells, about 1 %, are rewritten in place) against the measured HBM peak, pits found / unsolved, rounds.
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from overflow_b200 import _native, device as dev  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--kind", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    n = a.size
    peak = 6551.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    _native.init(0)
    dem0 = dev.synth_dem(n, n, seed=3, kind=a.kind, holes_permille=5)
    dem = dem0.clone()
    times, info = [], None
    for it in range(a.steps + 1):
        dem.copy_(dem0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, info = dev.breach_single_cell_pits(dem, -9999.0)
        e1.record()
        torch.cuda.synchronize()
        if it:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    cells = n * n
    print(json.dumps({
        "what": "ofl_breach_single_cell_pits_f32, device buffers", "size": n, "kind": a.kind, "ms": ms,
        "gcells_s": cells / ms / 1e6, "algorithmic_gbs_5B": cells * 5 / ms / 1e6, "frac_of_hbm_peak": cells * 5 / ms / 1e6 / peak,
        "pits": info[0], "unsolved": info[1], "rounds": info[2], "cells_changed": int((dem != dem0).sum().item()),
    }))


if __name__ == "__main__":
    main()
