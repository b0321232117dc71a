#!/usr/bin/env python
"""File-to-file timing of the streaming drivers and of out-of-core accumulation (rank 3).

    python scripts/bench_stream.py [--size 32768] [--dir /tmp] [--band-rows N] [--out-of-core]

Generates the synthetic benchmark DEM on the device, writes it as a (Big)TIFF, then times
DEM file -> codes file + counts file through the band pipeline next to the plain driver (read everything, one
library call, write everything), and prints one JSON line with both breakdowns.  The files live under --dir;
/dev/shm measures the pipeline without a disk, a real directory measures it with one.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32768)
    ap.add_argument("--dir", default="/tmp")
    ap.add_argument("--band-rows", type=int, default=None)
    ap.add_argument("--no-plain", action="store_true")
    ap.add_argument("--out-of-core", action="store_true", help="also time flow_accumulation_file_out_of_core")
    ap.add_argument("--strip-rows", type=int, default=4096)
    args = ap.parse_args()
    import numpy as np
    import torch

    from overflow_b200 import device as dev, streaming, strips
    from overflow_b200.flow_routing import flow_routing
    from overflow_b200.util.raster import create_raster, open_raster

    S = args.size
    d = os.path.join(args.dir, f"ofl_stream_{os.getpid()}")
    os.makedirs(d, exist_ok=True)
    p = lambda n: os.path.join(d, n)  # noqa: E731
    out = {"size": S, "dir": args.dir}
    try:
        dem = dev.synth_dem(S, S, seed=0, kind=0, holes_permille=5)
        t0 = time.perf_counter()
        ds = create_raster(p("dem.tif"), S, S, "Float32", geotransform=(0.0, 10.0, 0.0, 0.0, 0.0, -10.0))
        b = ds.GetRasterBand(1)
        for r in range(0, S, 2048):
            b.WriteArray(dem[r : r + 2048].cpu().numpy(), xoff=0, yoff=r)
        b.SetNoDataValue(-9999.0)
        ds.FlushCache()
        ds = b = None
        out["write_dem_s"] = time.perf_counter() - t0
        del dem
        torch.cuda.empty_cache()
        rep = streaming.stream_routing(p("dem.tif"), p("fdr_s.tif"), p("fac_s.tif"), band_rows=args.band_rows)
        rep = streaming.stream_routing(p("dem.tif"), p("fdr_s.tif"), p("fac_s.tif"), band_rows=args.band_rows)  # warm files
        rep["Gcells_per_s"] = S * S / rep["wall_s"] / 1e9
        out["streamed"] = rep
        if not args.no_plain:
            t0 = time.perf_counter()
            flow_routing(p("dem.tif"), p("fdr_p.tif"), p("fac_p.tif"), streamed=False)
            out["plain_wall_s"] = time.perf_counter() - t0
            a = open_raster(p("fac_s.tif")).GetRasterBand(1)
            c = open_raster(p("fac_p.tif")).GetRasterBand(1)
            same = all(np.array_equal(a.ReadAsArray(xoff=0, yoff=r, win_xsize=S, win_ysize=min(2048, S - r)),
                                      c.ReadAsArray(xoff=0, yoff=r, win_xsize=S, win_ysize=min(2048, S - r)))
                       for r in range(0, S, 2048))
            out["streamed_equals_plain"] = bool(same)
            os.remove(p("fac_p.tif"))
        if args.out_of_core:
            t0 = time.perf_counter()
            n = strips.flow_accumulation_file_out_of_core(p("fdr_s.tif"), p("fac_o.tif"), args.strip_rows)
            wall = time.perf_counter() - t0
            a = open_raster(p("fac_s.tif")).GetRasterBand(1)
            c = open_raster(p("fac_o.tif")).GetRasterBand(1)
            same = all(np.array_equal(a.ReadAsArray(xoff=0, yoff=r, win_xsize=S, win_ysize=min(2048, S - r)),
                                      c.ReadAsArray(xoff=0, yoff=r, win_xsize=S, win_ysize=min(2048, S - r)))
                       for r in range(0, S, 2048))
            out["out_of_core"] = {"strips": n, "strip_rows": args.strip_rows, "wall_s": wall,
                                  "Gcells_per_s": S * S / wall / 1e9, "equals_whole_raster": bool(same)}
    finally:
        for f in os.listdir(d):
            os.remove(os.path.join(d, f))
        os.rmdir(d)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
