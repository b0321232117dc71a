"""Generate tests/golden/direction_dtypes.npz by running the REFERENCE's flow_direction_for_tile on DEMs
of every element type a GDAL band can have (util/raster.py:22-34).

Run in the dev container only (needs /root/reference and numba):

    python oracle/gen_golden_dtypes.py

Under numba the elevation difference (flow_direction.py:94) is taken in the array's arithmetic: float64
for float64, int64 for signed integers, uint64 -- wrapping for uphill neighbours -- for unsigned ones.
The fixtures pin that behaviour: small rasters with plateaus, nodata cells and, for the wide integer
types, magnitudes beyond 2**53.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle.gen_golden import GOLD, import_reference  # noqa: E402


def cases():
    rng = np.random.default_rng(7)
    out = []
    for name, dt, lo, hi, nodata in [
        ("uint8", np.uint8, 0, 60, 255), ("uint16", np.uint16, 100, 3000, 65535), ("uint32", np.uint32, 0, 70000, 4294967295),
        ("uint64", np.uint64, 0, 5000, 9999), ("int8", np.int8, -50, 50, -128), ("int16", np.int16, -300, 3000, -32768),
        ("int32", np.int32, -100000, 100000, -9999), ("int64", np.int64, -5000, 5000, -9999),
    ]:
        for k, shape in enumerate([(9, 11), (24, 17)]):
            dem = rng.integers(lo, hi, size=shape).astype(dt)
            dem[2:4, 3:6] = dem[2, 3]            # a plateau
            dem[shape[0] // 2, shape[1] // 2] = nodata
            dem[1, 1] = nodata
            out.append((f"{name}_{k}", dem, float(nodata)))
    # magnitudes where int64 -> float64 rounds
    big = (rng.integers(-4, 4, size=(8, 8)) + (1 << 60)).astype(np.int64)
    big[3, 3] = -9999
    out.append(("int64_big", big, -9999.0))
    ubig = (rng.integers(0, 8, size=(8, 8)).astype(np.uint64) + np.uint64(1 << 63))
    ubig[2, 5] = 7
    out.append(("uint64_big", ubig, 7.0))
    for k, shape in enumerate([(10, 13), (31, 19)]):
        dem = rng.normal(size=shape) * 50.0
        dem[1:3, 2:5] = dem[1, 2]
        dem[shape[0] // 2, 4] = -9999.0
        out.append((f"float64_{k}", dem.astype(np.float64), -9999.0))
    f = rng.normal(size=(12, 12)) + 1e15  # differences far below float32 resolution
    out.append(("float64_fine", f.astype(np.float64), -9999.0))
    return out


def main():
    fd, _ = import_reference()
    store = {}
    for name, dem, nodata in cases():
        res = fd.flow_direction_for_tile(dem, nodata)
        store[name + "__dem"] = dem
        store[name + "__nodata"] = np.float64(nodata)
        store[name + "__fdr"] = res[1:-1, 1:-1].copy()
    np.savez_compressed(os.path.join(GOLD, "direction_dtypes.npz"), **store)
    print("wrote", len(store) // 3, "cases")


if __name__ == "__main__":
    main()
