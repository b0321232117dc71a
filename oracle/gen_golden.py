"""Generate tests/golden/*.npz by running the REFERENCE's own numba kernels.

Run in the dev container only (needs /root/reference and numba):

    python oracle/gen_golden.py [--big]

The reference (pure Python + numba) is imported from /root/reference/src with a stub
``osgeo`` module (its flow_direction.py:4 and util/raster.py:2 import GDAL at module
top but the tile kernels never touch it).  Outputs are small fixtures that travel to
the GPU box, where /root/reference does not exist.  ``--big`` adds the 1024x1024
config-1 anchor, which takes ~4 minutes because the reference accumulation is O(N^2)
(queue.pop(0), flow_accumulation.py:136).

Known-answer tests restated as data (values are the reference's fixtures):
  direction     tests/test_flow_direction.py:58-69 (7x7 DEM), :79-118 (expected 5x5 codes)
  accumulation  tests/test_flow_accumulation.py:20-87 (7x7 fdr), :93-104 (fac), :107-130 (links)
"""
import argparse
import hashlib
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402


def import_reference():
    osgeo = types.ModuleType("osgeo")
    gdal = types.ModuleType("osgeo.gdal")
    gdal.UseExceptions = lambda: None
    gdal.Band = object
    osgeo.gdal = gdal
    sys.modules.setdefault("osgeo", osgeo)
    sys.modules.setdefault("osgeo.gdal", gdal)
    sys.path.insert(0, "/root/reference/src")
    from overflow import flow_direction as fd  # noqa
    from overflow import flow_accumulation as fa  # noqa

    # codes 8/9 must resolve to "outside" (value 9) in the reference build.
    probe = np.array([[8, 9], [0, 4]], dtype=np.uint8)
    assert fa.get_next_cell(probe, 0, 0)[2] == 9 and fa.get_next_cell(probe, 0, 1)[2] == 9
    return fd, fa


E, NE, N, NW, W, SW, S, SE, UNDEF, ND = range(10)

KAT_DEM7 = np.array(
    [
        [-9999, -9999, -9999, -9999, -9999, -9999, -9999],
        [-9999, 5, 4, 3, 2, 1, -9999],
        [-9999, 5, 5, 5, 5, 5, -9999],
        [-9999, 5, 5, 5, 5, 5, -9999],
        [-9999, 5, 5, 4, 5, 5, -9999],
        [-9999, 5, 5, 5, 5, 5, -9999],
        [-9999, -9999, -9999, -9999, -9999, -9999, -9999],
    ],
    dtype=np.float32,
)
KAT_FDR5 = np.array(
    [
        [NE, NE, NE, NE, E],
        [NW, NE, NE, N, E],
        [NW, SE, S, SW, E],
        [NW, E, UNDEF, W, E],
        [NW, SW, SW, SW, E],
    ],
    dtype=np.uint8,
)
KAT_ACC_FDR = np.array(
    [
        [N, N, N, N, N, N, N],
        [NE, N, NW, NE, NE, N, NW],
        [N, N, N, NE, N, N, NW],
        [N, N, NW, W, N, NW, NW],
        [W, N, N, NW, NW, W, W],
        [NW, NW, SE, E, N, NW, W],
        [NW, SE, E, NE, N, N, W],
    ],
    dtype=np.int64,
)
KAT_ACC_FAC = np.array(
    [
        [1, 27, 1, 1, 2, 11, 1],
        [3, 21, 2, 1, 5, 4, 1],
        [2, 20, 1, 1, 3, 2, 1],
        [1, 2, 17, 14, 1, 1, 1],
        [2, 1, 1, 1, 13, 2, 1],
        [1, 1, 1, 1, 6, 4, 1],
        [1, 1, 1, 3, 1, 2, 1],
    ],
    dtype=np.int64,
)
EXT = (-2, -2)
KAT_ACC_LINKS = {
    (0, 0): EXT, (0, 1): EXT, (0, 2): EXT, (0, 3): EXT, (0, 4): EXT, (0, 5): EXT, (0, 6): EXT,
    (1, 0): (0, 1), (2, 0): (0, 1), (3, 0): (0, 1), (4, 0): EXT, (5, 0): EXT, (6, 0): EXT,
    (1, 6): (0, 5), (2, 6): (0, 5), (3, 6): (0, 5), (4, 6): (0, 1), (5, 6): (0, 1), (6, 6): (0, 1),
    (6, 1): EXT, (6, 2): (0, 1), (6, 3): (0, 1), (6, 4): (0, 1), (6, 5): (0, 1),
}


def perimeter_links(fa, fdr):
    """Run the reference and keep only what it defines: perimeter link entries."""
    fac, links = fa.single_tile_flow_accumulation(fdr)
    idx = np.array(fa.perimeter_indices(fdr.shape), dtype=np.int64).reshape(-1, 2)
    return fac, idx, links[idx[:, 0], idx[:, 1]].copy()


def discriminating_vectors():
    """3x3 tiles from SURVEY 8(a); expected centre codes are re-derived by the reference."""
    nd = np.float32(-9999.0)
    f = np.float32
    t = []

    def tile(c, **kw):
        a = np.full((3, 3), c, dtype=np.float32)
        pos = dict(E=(1, 2), NE=(0, 2), N=(0, 1), NW=(0, 0), W=(1, 0), SW=(2, 0), S=(2, 1), SE=(2, 2))
        for k, v in kw.items():
            a[pos[k]] = v
        return a

    t.append(tile(f(2.0**24), E=0.5, N=0.25))  # f32 subtraction ties -> 0
    t.append(tile(10, E=9, N=9, W=9, S=9))  # equal cardinals -> 0
    t.append(tile(10, NE=9, NW=9, SW=9, SE=9))  # equal diagonals -> 1
    t.append(tile(10, E=9, SE=8.6))  # 1 vs 1.4/sqrt2 -> 0
    t.append(tile(10, E=9, SE=8.5))  # 1 vs 1.5/sqrt2 -> 7
    t.append(tile(10))  # flat -> 8
    t.append(tile(10, E=11, NE=11, N=11, NW=11, W=11, SW=11, S=11, SE=11))  # pit -> 8
    t.append(tile(10, E=-1000, S=nd))  # nodata beats any drop -> 6
    t.append(tile(10, SE=nd, NW=nd))  # first nodata in scan order -> 3
    t.append(tile(nd))  # centre nodata -> 9
    t.append(tile(np.nan))  # centre NaN -> 8
    t.append(tile(10, E=np.nan, W=5))  # NaN neighbour never wins -> 4
    t.append(tile(np.inf))  # centre +inf, neighbours +inf -> NaN slopes -> 8
    t.append(tile(np.inf, E=3, N=2))  # centre +inf -> first finite neighbour -> 0
    t.append(tile(10, W=-np.inf, S=nd))  # -inf neighbour gives +inf slope before S -> 4
    t.append(tile(-np.inf, S=nd))  # centre -inf with nodata neighbour -> 6
    t.append(tile(np.nan, S=nd, NE=nd))  # centre NaN with nodata neighbours -> 1
    t.append(tile(f(-0.0), E=f(0.0)))  # signed zero flat -> 8
    return np.stack(t)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="also generate the 1024^2 config-1 anchor (~4 min)")
    args = ap.parse_args()
    fd, fa = import_reference()
    os.makedirs(GOLD, exist_ok=True)

    # --- known-answer tests, cross-checked against the live reference ---
    got = fd.flow_direction_for_tile(KAT_DEM7, -9999)[1:-1, 1:-1]
    assert np.array_equal(got, KAT_FDR5), "reference disagrees with its own direction KAT"
    fac, idx, pl = perimeter_links(fa, KAT_ACC_FDR)
    assert np.array_equal(fac, KAT_ACC_FAC)
    for (r, c), want in KAT_ACC_LINKS.items():
        k = np.flatnonzero((idx[:, 0] == r) & (idx[:, 1] == c))[0]
        assert tuple(pl[k]) == want
    np.savez_compressed(
        os.path.join(GOLD, "kat.npz"),
        dir_dem=KAT_DEM7, dir_nodata=np.float64(-9999), dir_expected=KAT_FDR5,
        acc_fdr=KAT_ACC_FDR, acc_fac=KAT_ACC_FAC, acc_perim_rc=idx, acc_perim_links=pl,
    )

    # --- discriminating 3x3 vectors ---
    tiles = discriminating_vectors()
    centre = np.array([fd.flow_direction_for_tile(t, -9999.0)[1, 1] for t in tiles], dtype=np.uint8)
    # nodata not representable in float32: compare happens in float64 -> never matches
    t_unrep = np.full((3, 3), np.float32(-1.1), dtype=np.float32)
    c_unrep = fd.flow_direction_for_tile(t_unrep, -1.1)[1, 1]
    np.savez_compressed(
        os.path.join(GOLD, "discriminating.npz"),
        tiles=tiles, nodata=np.float64(-9999.0), centre=centre,
        unrep_tile=t_unrep, unrep_nodata=np.float64(-1.1), unrep_centre=np.uint8(c_unrep),
    )
    print("discriminating centres:", centre.tolist(), "unrep:", int(c_unrep))

    # --- direction fuzz: every distribution, odd shapes, float32 ---
    out = {}
    for i, kind in enumerate(synth.FUZZ_KINDS):
        dem = synth.pad_nodata(synth.fuzz_dem(kind, 45, 83, seed=100 + i))
        fdr = fd.flow_direction_for_tile(dem, synth.NODATA)
        out[f"{kind}_dem"] = dem
        out[f"{kind}_fdr"] = fdr[1:-1, 1:-1].copy()
    # NaN as the nodata value: nothing is ever nodata
    dem = synth.fuzz_dem("special", 21, 33, seed=7)
    out["nan_nodata_dem"] = dem
    out["nan_nodata_fdr"] = fd.flow_direction_for_tile(dem, float("nan"))[1:-1, 1:-1].copy()
    # float64 DEM (subtraction happens in float64)
    dem64 = synth.pad_nodata(synth.fuzz_dem("uniform", 20, 31, seed=9).astype(np.float64) + 1e-9)
    out["f64_dem"] = dem64
    out["f64_fdr"] = fd.flow_direction_for_tile(dem64, synth.NODATA)[1:-1, 1:-1].copy()
    np.savez_compressed(os.path.join(GOLD, "direction_fuzz.npz"), **out)

    # --- accumulation + links on reference-made direction rasters (verbatim reference, small) ---
    cases = {
        "fractal_b2": synth.punch_holes(synth.fractal(150, 170, beta=2.0, seed=0), frac=0.01, seed=3),
        "fractal_b3": synth.fractal(130, 130, beta=3.0, seed=1),
        "fractal_b4": synth.fractal(128, 192, beta=4.0, seed=2),
        "terraced": synth.terraced(140, 150, seed=4),
        "tilted": synth.tilted_plane(96, 80),
        "tilted_diag": synth.tilted_plane(70, 90, a=1.0, b=-1.0),
        "serpentine": synth.serpentine(65, 67),
        "ints": synth.fuzz_dem("ints", 90, 75, seed=5),
        "row": synth.fractal(1, 40, beta=2.0, seed=6),
        "col": synth.fractal(40, 1, beta=2.0, seed=7),
        "one": np.array([[3.0]], dtype=np.float32),
    }
    out = {}
    for name, dem in cases.items():
        fdr = fd.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1].copy()
        fac, idx, pl = perimeter_links(fa, fdr)
        out[f"{name}_dem"] = dem
        out[f"{name}_fdr"] = fdr
        out[f"{name}_fac"] = fac
        out[f"{name}_perim_rc"] = idx
        out[f"{name}_perim_links"] = pl
        print(name, fdr.shape, "max fac", int(fac.max()), "nodata cells", int((fdr == 9).sum()))
    # an all-nodata and an all-undefined raster
    for name, code in (("all_nodata", 9), ("all_undefined", 8)):
        fdr = np.full((9, 11), code, dtype=np.uint8)
        fac, idx, pl = perimeter_links(fa, fdr)
        out[f"{name}_fdr"] = fdr
        out[f"{name}_fac"] = fac
        out[f"{name}_perim_rc"] = idx
        out[f"{name}_perim_links"] = pl
    np.savez_compressed(os.path.join(GOLD, "accumulation.npz"), **out)

    if args.big:
        # config 1: 1024^2 fractal, beta=2, seed 0, nodata ring; verbatim reference for both kernels
        dem = synth.fractal(1024, 1024, beta=2.0, seed=0)
        fdr = fd.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1].copy()
        fac, idx, pl = perimeter_links(fa, fdr)
        np.savez_compressed(
            os.path.join(GOLD, "config1_1024.npz"),
            dem_sha256=hashlib.sha256(dem.tobytes()).hexdigest(),
            fdr=fdr, fac=fac.astype(np.int32), perim_rc=idx.astype(np.int32), perim_links=pl.astype(np.int32),
        )
        print("config1 1024: max fac", int(fac.max()))


if __name__ == "__main__":
    main()
