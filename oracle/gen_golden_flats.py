"""Generate tests/golden/fix_flats.npz by running the REFERENCE's flat-resolution functions
(/root/reference/src/overflow/fix_flats.py: flat_edges, resolve_flats, d8_masked_flow_dirs).

Run in the dev container only (needs /root/reference and numba):

    python oracle/gen_golden_flats.py

Cases: the reference's known-answer fixture (tests/test_fix_flats.py:24-119 DEM and codes, :175-192
final mask, :291-372 rewritten codes), terraced fractals with nodata blobs whose codes come from the
reference's own flow_direction_for_tile, small-integer DEMs (many flats, pits, undrainable flats,
flats joined through same-height cells that have a direction), flats touching the raster edge,
NaN / inf cells, an all-flat DEM (no low edges), and "inconsistent" inputs -- random codes 0..9 over a
random small-integer DEM -- that pin the rules where real inputs never go (a NODATA cell can be a low
edge, label 0 matches label 0 in d8_masked_flow_dirs).
Per case the file holds dem, fdr, edges (bit 0 low, bit 1 high), flat_mask, labels, fdr_fixed.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402
from oracle.gen_golden import GOLD, import_reference  # noqa: E402

E, NE, N, NW, W, SW, S, SE, U, ND = range(10)

KAT_DEM = np.array(
    [
        [1, 1, 1, 1, 1, 1, 1],
        [1, 0, 0, 0, 0, 0, 1],
        [1, 0, 0, 0, 0, 0, 1],
        [1, 0, 0, 0, 0, 0, 1],
        [1, 0, 0, 0, 0, 0, 1],
        [1, 0, 0, 0, 0, 0, 1],
        [1, 1, -1, 1, 1, 1, 1],
    ],
    dtype=np.float32,
)
KAT_FDR = np.array(
    [
        [SE, S, S, S, S, S, SW],
        [E, U, U, U, U, U, W],
        [E, U, U, U, U, U, W],
        [E, U, U, U, U, U, W],
        [E, U, U, U, U, U, W],
        [E, SE, S, SW, U, U, W],
        [NE, N, S, N, N, N, NW],
    ],
    dtype=np.uint8,
)
KAT_MASK = np.array(
    [
        [0, 0, 0, 0, 0, 0, 0],
        [0, 12, 12, 12, 12, 12, 0],
        [0, 10, 9, 9, 9, 10, 0],
        [0, 8, 7, 6, 7, 8, 0],
        [0, 6, 5, 5, 5, 8, 0],
        [0, 2, 2, 2, 6, 8, 0],
        [0, 0, 0, 0, 0, 0, 0],
    ],
    dtype=np.int32,
)
KAT_FIXED = np.array(
    [
        [SE, S, S, S, S, S, SW],
        [E, SE, S, S, S, SW, W],
        [E, SE, SE, S, SW, SW, W],
        [E, SE, S, S, S, SW, W],
        [E, S, S, S, SW, W, W],
        [E, SE, S, SW, W, NW, W],
        [NE, N, S, N, N, N, NW],
    ],
    dtype=np.uint8,
)


def codes_of(fd, dem, nodata=synth.NODATA):
    """The reference's own flow direction of the whole raster (nodata ring outside)."""
    return np.ascontiguousarray(fd.flow_direction_for_tile(synth.pad_nodata(dem, nodata), nodata)[1:-1, 1:-1])


def cases(fd):
    rng = np.random.default_rng(11)
    out = [("kat", KAT_DEM, KAT_FDR)]
    for k, (n, m, step, relief) in enumerate([(96, 96, 1.0, 12.0), (80, 131, 2.0, 30.0), (128, 128, 1.0, 6.0)]):
        dem = synth.terraced(n, m, seed=20 + k, step=step, relief=relief, nodata_frac=0.04)
        out.append((f"terraced_{k}", dem, codes_of(fd, dem)))
    for k, (shape, hi) in enumerate([((40, 40), 3), ((33, 57), 2), ((64, 64), 5), ((70, 20), 4)]):
        dem = rng.integers(0, hi, size=shape).astype(np.float32)
        out.append((f"ints_{k}", dem, codes_of(fd, dem)))
    # big blocks of equal height: wide flats with long gradients, some closed (undrainable)
    blocks = np.kron(rng.integers(0, 4, size=(9, 9)), np.ones((8, 8))).astype(np.float32)
    blocks[30:34, 10:60] = 7.0  # a wall
    out.append(("blocks", blocks, codes_of(fd, blocks)))
    bowl = np.full((48, 48), 5.0, dtype=np.float32)
    bowl[8:40, 8:40] = 2.0  # closed flat: no low edge
    bowl[20:24, 20:24] = 3.0
    out.append(("closed_bowl", bowl, codes_of(fd, bowl)))
    outlet = bowl.copy()
    outlet[40:, 24] = 1.0  # ... and the same with an outlet channel
    out.append(("bowl_outlet", outlet, codes_of(fd, outlet)))
    allflat = np.zeros((17, 23), dtype=np.float32)
    out.append(("all_flat", allflat, np.full(allflat.shape, U, dtype=np.uint8)))  # no ring: nothing drains
    edge = np.zeros((30, 30), dtype=np.float32)
    edge[:, 15:] = 1.0
    edge[10:20, 10:20] = 0.0
    out.append(("edge_flat", edge, codes_of(fd, edge)))
    holes = synth.punch_holes(np.kron(rng.integers(0, 3, size=(8, 8)), np.ones((6, 6))).astype(np.float32), frac=0.08, seed=3)
    out.append(("blocks_nodata", holes, codes_of(fd, holes)))
    special = synth.fuzz_dem("special", 40, 40, seed=5)
    special[10:25, 5:30] = np.where(rng.random((15, 25)) < 0.1, np.nan, 1.0).astype(np.float32)
    out.append(("special", special, codes_of(fd, special)))
    zeros = rng.choice(np.array([0.0, -0.0, 1.0], dtype=np.float32), size=(36, 36))  # -0.0 == 0.0
    out.append(("signed_zero", zeros, codes_of(fd, zeros)))
    for k, shape in enumerate([(24, 24), (31, 45), (50, 50)]):
        dem = rng.integers(0, 3, size=shape).astype(np.float32)
        fdr = rng.choice(np.arange(10, dtype=np.uint8), size=shape, p=[0.06] * 8 + [0.42, 0.10])
        out.append((f"inconsistent_{k}", dem, fdr))
    for shape in [(1, 1), (1, 9), (9, 1), (2, 2)]:
        dem = np.zeros(shape, dtype=np.float32)
        fdr = np.full(shape, U, dtype=np.uint8)
        fdr.flat[0] = E
        out.append((f"tiny_{shape[0]}x{shape[1]}", dem, fdr))
    return out


def main():
    fd, _ = import_reference()
    from overflow import fix_flats as ff

    store = {}
    names = []
    for name, dem, fdr in cases(fd):
        dem = np.ascontiguousarray(dem, dtype=np.float32)
        fdr = np.ascontiguousarray(fdr, dtype=np.uint8)
        high, low = ff.flat_edges(dem, fdr)
        edges = np.zeros(dem.shape, dtype=np.uint8)
        for r, c in low:
            edges[r, c] |= 1
        for r, c in high:
            edges[r, c] |= 2
        flat_mask, labels = ff.resolve_flats(dem, fdr)
        fixed = fdr.copy()
        ff.d8_masked_flow_dirs(flat_mask, fixed, labels)
        if name == "kat":  # the reference's expected values, tests/test_fix_flats.py
            assert np.array_equal(flat_mask, KAT_MASK) and np.array_equal(fixed, KAT_FIXED)
        for key, arr in (("dem", dem), ("fdr", fdr), ("edges", edges), ("flat_mask", flat_mask.astype(np.int32)),
                         ("labels", labels.astype(np.int32)), ("fdr_fixed", fixed)):
            store[f"{name}__{key}"] = arr
        names.append(name)
        print(f"{name:18s} {dem.shape!s:12s} low {len(low):5d} high {len(high):5d} labels {int(labels.max()):4d} "
              f"max mask {int(flat_mask.max()):4d} undefined {int((fdr == U).sum()):5d} -> {int((fixed == U).sum()):5d}")
    np.savez_compressed(os.path.join(GOLD, "fix_flats.npz"), **store)
    print("wrote", len(names), "cases")


if __name__ == "__main__":
    main()
