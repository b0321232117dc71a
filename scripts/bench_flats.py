"""Time flat resolution (ofl_fix_flats_f32, device buffers) on synthetic DEMs.

    python scripts/bench_flats.py [--size 8192] [--kind 1] [--relief 200] [--steps 3]

kind 0 fractal (few, small flats), kind 1 terraces (flat-heavy), kind 2 tilted plane (none).
Prints one JSON line per run: ms per call (CUDA events on the launching stream), Gcells/s, the algorithmic
GB/s at 14 B/cell (DEM 4 + codes 1 read; codes 1 + flat_mask 4 + labels 4 written) against the measured HBM
peak, the sweep depths and the per-phase times.  (Parity against the oracle and the CPU baseline of this row are
in tests/test_gpu_flats.py and bench.py's next_rows.fix_flats.)
"""
import argparse
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
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--kind", type=int, default=1)
    ap.add_argument("--relief", type=float, default=200.0)
    ap.add_argument("--holes", type=int, default=5)
    ap.add_argument("--lake", action="store_true", help="one flat of (size - 18)^2 cells drained through one channel: size/2 and size levels")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--oracle-window", type=int, default=0, help="ignored (parity and the CPU baseline live in tests/ and bench.py)")
    a = ap.parse_args()
    n = a.size
    peak = 6551.4
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    _native.init(0)
    if a.lake:
        dem = torch.full((n, n), 10.0, dtype=torch.float32, device="cuda")
        dem[9:-9, 9:-9] = 5.0
        dem[n // 2, :10] = torch.linspace(1.0, 4.5, 10, device="cuda")
    else:
        dem = dev.synth_dem(n, n, seed=3, kind=a.kind, relief=a.relief, holes_permille=a.holes)
    fdr0 = dev.flow_direction(dem, -9999.0).contiguous()
    work = dev.flats_workspace(n, n)
    flat_mask = torch.empty((n, n), dtype=torch.int32, device="cuda")
    labels = torch.empty((n, n), dtype=torch.int32, device="cuda")
    fdr = fdr0.clone()
    times = []
    info = None
    _native.phase_timing_enable(True)
    for it in range(a.steps + 1):
        fdr.copy_(fdr0)
        torch.cuda.synchronize()
        _native.launch_count_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _, info = dev.fix_flats(dem, fdr, workspace=work, flat_mask=flat_mask, labels=labels)
        e1.record()
        torch.cuda.synchronize()
        if it:
            times.append(e0.elapsed_time(e1))
    launches = _native.launch_count()
    phases = {k: v[0] / max(1, a.steps + 1) for k, v in _native.phase_timing_read().items() if k.startswith("flats")}
    ms = float(np.median(times))
    cells = n * n
    out = {
        "what": "ofl_fix_flats_f32 (resolve_flats + d8_masked_flow_dirs), device buffers",
        "size": n, "kind": "lake" if a.lake else a.kind, "relief": a.relief, "ms": ms, "gcells_s": cells / ms / 1e6,
        "algorithmic_gbs_14B": cells * 14 / ms / 1e6, "frac_of_hbm_peak": cells * 14 / ms / 1e6 / peak, "peak_gbs": peak,
        "undefined_before": int((fdr0 == 8).sum()), "undefined_after": int((fdr == 8).sum()),
        "low_edges": info[0], "high_edges": info[1], "labels": info[2], "away_levels": info[3],
        "towards_levels": info[4], "launches_per_call": launches, "phases_ms": phases,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
