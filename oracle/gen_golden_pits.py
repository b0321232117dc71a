"""Generate tests/golden/breach_pits.npz by running the REFERENCE's breach_single_cell_pits_in_chunk
(/root/reference/src/overflow/breach_single_cell_pits.py:9-63).

Run in the dev container only (needs /root/reference and numba):

    python oracle/gen_golden_pits.py

Cases: the reference's fixture (tests/test_breach_single_cell_pits.py:38-83), fractals with nodata holes, uniform
noise (a pit every few cells: the sequential second pass makes neighbouring pits depend on each other), small
integers (ties, plateaus), pits on lattices of spacing 2 and 3 (long dependency chains, cells written twice),
NaN / inf / nodata fuzz, nodata = -inf (the reference's file fixture), chunks too small to have an interior.
Per case: chunk_in, nodata, chunk_out, unsolved.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402
from oracle.gen_golden import GOLD, import_reference  # noqa: E402

KAT = np.full((9, 9), -999, dtype=np.float32)
KAT[2:7, 2:7] = 2
KAT[3, 2] = -1
KAT[4, 4] = 0


def cases():
    rng = np.random.default_rng(21)
    out = [("kat", KAT, -999.0)]
    out.append(("fractal_holes", synth.pad_nodata(synth.pad_nodata(synth.punch_holes(synth.fractal(120, 150, beta=1.5, seed=4), frac=0.02, seed=2))), synth.NODATA))
    out.append(("uniform", rng.uniform(0, 100, (90, 77)).astype(np.float32), synth.NODATA))
    out.append(("ints", rng.integers(0, 6, (64, 64)).astype(np.float32), synth.NODATA))
    for name, step in (("lattice2", 2), ("lattice3", 3)):
        z = rng.uniform(50, 60, (41, 47)).astype(np.float32)
        z[2:-2:step, 2:-2:step] = rng.uniform(0, 40, z[2:-2:step, 2:-2:step].shape).astype(np.float32)
        out.append((name, z, synth.NODATA))
    diag = rng.uniform(50, 60, (60, 60)).astype(np.float32)
    for i in range(2, 58, 2):
        diag[i, i] = 40 - i * 0.5  # a descending diagonal of pits: each reads what the previous one wrote
    out.append(("diagonal_chain", diag, synth.NODATA))
    sp = synth.fuzz_dem("special", 70, 70, seed=9)
    out.append(("special", sp, synth.NODATA))
    ninf = rng.uniform(0, 10, (50, 50)).astype(np.float32)
    ninf[rng.random((50, 50)) < 0.05] = -np.inf
    out.append(("nodata_neg_inf", ninf, float("-inf")))
    quant = (np.round(rng.uniform(0, 20, (80, 80)) * 2) / 2).astype(np.float32)
    quant[rng.random((80, 80)) < 0.03] = synth.NODATA
    out.append(("quantised_nodata", quant, synth.NODATA))
    for shape in [(4, 4), (5, 5), (5, 9), (1, 1), (3, 40)]:
        z = rng.uniform(1, 2, shape).astype(np.float32)
        if shape[0] >= 5 and shape[1] >= 5:
            z[2, 2] = 0.0
        out.append((f"tiny_{shape[0]}x{shape[1]}", z, synth.NODATA))
    return out


def main():
    import_reference()
    from overflow.breach_single_cell_pits import breach_single_cell_pits_in_chunk as ref

    store = {}
    for name, chunk, nodata in cases():
        work = np.ascontiguousarray(chunk, dtype=np.float32).copy()
        unsolved = ref(work, nodata)
        if name == "kat":  # tests/test_breach_single_cell_pits.py:62-83
            assert work[4, 3] == -0.5 and (np.delete(work.ravel(), 4 * 9 + 3) == np.delete(KAT.ravel(), 4 * 9 + 3)).all()
        store[f"{name}__chunk_in"] = np.ascontiguousarray(chunk, dtype=np.float32)
        store[f"{name}__nodata"] = np.float64(nodata)
        store[f"{name}__chunk_out"] = work
        store[f"{name}__unsolved"] = unsolved.astype(np.int8)
        changed = int((work.view(np.uint32) != np.ascontiguousarray(chunk, dtype=np.float32).view(np.uint32)).sum())
        print(f"{name:18s} {chunk.shape!s:11s} changed cells {changed:5d} unsolved {int(unsolved.sum()):4d}")
    np.savez_compressed(os.path.join(GOLD, "breach_pits.npz"), **store)
    print("wrote", len(store) // 4, "cases")


if __name__ == "__main__":
    main()
