"""Scratch: direction kernel time for rasters of equal cell count and different shapes / pitches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from overflow_b200 import _native, device as dev
_native.init(0)
for rows, cols, pitch in [(16384, 16384, 16384), (4096, 65536, 65536), (65536, 4096, 4096), (16384, 16384, 16384 + 64),
                          (16384, 16000, 16000), (32768, 32768, 32768), (32768, 32768, 32768 + 64)]:
    buf = torch.empty((rows, pitch), dtype=torch.float32, device="cuda")
    dem = buf[:, :cols]
    tmp = dev.synth_dem(rows, cols, seed=0, kind=0, holes_permille=5)
    dem.copy_(tmp); del tmp
    out = torch.empty((rows, (cols + 15) // 16 * 16), dtype=torch.uint8, device="cuda")[:, :cols]
    for _ in range(3): dev.flow_direction(dem, -9999.0, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dev.flow_direction(dem, -9999.0, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(rows, cols, pitch, round(ms, 4), "ms", round(rows * cols / ms / 1e6, 1), "Gcells/s")
    del buf, dem, out
