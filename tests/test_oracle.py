"""Pins the CPU oracle (oracle/d8_oracle.c) to the reference.

Every fixture under tests/golden/ was produced by the reference's own numba kernels
(oracle/gen_golden.py); the two known-answer tests are the reference's
tests/test_flow_direction.py:58-118 and tests/test_flow_accumulation.py:20-130.
"""
import hashlib
import os

import numpy as np
import pytest

import oracle
from oracle import synth
from conftest import GOLDEN, load_golden


def test_kat_direction():
    g = load_golden("kat.npz")
    fdr = oracle.flow_direction_for_tile(g["dir_dem"], float(g["dir_nodata"]))
    assert np.array_equal(fdr[1:-1, 1:-1], g["dir_expected"])


def test_kat_accumulation_and_links():
    g = load_golden("kat.npz")
    fac, links = oracle.single_tile_flow_accumulation(g["acc_fdr"])  # int64 codes, like the reference test
    assert fac.dtype == np.int64 and links.dtype == np.int64
    assert np.array_equal(fac, g["acc_fac"])
    rc = g["acc_perim_rc"]
    assert np.array_equal(links[rc[:, 0], rc[:, 1]], g["acc_perim_links"])
    assert tuple(links[0, 0]) == oracle.oracle.FLOW_EXTERNAL
    assert tuple(links[1, 0]) == (0, 1)


def test_discriminating_vectors():
    g = load_golden("discriminating.npz")
    for tile, want in zip(g["tiles"], g["centre"]):
        assert oracle.flow_direction_for_tile(tile, float(g["nodata"]))[1, 1] == want
    got = oracle.flow_direction_for_tile(g["unrep_tile"], float(g["unrep_nodata"]))[1, 1]
    assert got == g["unrep_centre"]


@pytest.mark.parametrize("kind", synth.FUZZ_KINDS + ("nan_nodata", "f64"))
def test_direction_fuzz(kind):
    g = load_golden("direction_fuzz.npz")
    nodata = float("nan") if kind == "nan_nodata" else synth.NODATA
    fdr = oracle.flow_direction_for_tile(g[f"{kind}_dem"], nodata)
    assert np.array_equal(fdr[1:-1, 1:-1], g[f"{kind}_fdr"])


ACC_CASES = ["fractal_b2", "fractal_b3", "fractal_b4", "terraced", "tilted", "tilted_diag",
             "serpentine", "ints", "row", "col", "one", "all_nodata", "all_undefined"]


@pytest.mark.parametrize("name", ACC_CASES)
def test_accumulation_golden(name):
    g = load_golden("accumulation.npz")
    fdr = g[f"{name}_fdr"]
    if f"{name}_dem" in g:
        got = oracle.flow_direction_for_tile(synth.pad_nodata(g[f"{name}_dem"]), synth.NODATA)[1:-1, 1:-1]
        assert np.array_equal(got, fdr)
    fac = oracle.flow_accumulation(fdr)
    assert np.array_equal(fac, g[f"{name}_fac"])
    rc, pl = oracle.links_perimeter(fdr)
    assert np.array_equal(rc, g[f"{name}_perim_rc"])
    assert np.array_equal(pl, g[f"{name}_perim_links"])
    assert oracle.check_accumulation(fdr, fac) == 0
    if fac.size > 4 and (fdr != 9).any():
        bad = fac.copy()
        r, c = np.argwhere(fdr != 9)[0]
        bad[r, c] += 1
        assert oracle.check_accumulation(fdr, bad) >= 1


def test_nodata_cells_are_minus_9998():
    # the reference ends NODATA cells at -9998, not -9999
    g = load_golden("accumulation.npz")
    fdr, fac = g["terraced_fdr"], g["terraced_fac"]
    assert (fdr == 9).any()
    assert np.all(fac[fdr == 9] == -9998)
    assert np.all(oracle.flow_accumulation(fdr)[fdr == 9] == -9998)


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLDEN, "config1_1024.npz")), reason="big anchor not generated")
def test_config1_1024_anchor():
    g = load_golden("config1_1024.npz")
    dem = synth.fractal(1024, 1024, beta=2.0, seed=0)
    if hashlib.sha256(dem.tobytes()).hexdigest() != str(g["dem_sha256"]):
        pytest.skip("numpy FFT produced a different DEM than the one the fixture was made from")
    fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
    assert np.array_equal(fdr, g["fdr"])
    fac = oracle.flow_accumulation(fdr)
    assert np.array_equal(fac, g["fac"].astype(np.int64))
    rc, pl = oracle.links_perimeter(fdr)
    assert np.array_equal(rc, g["perim_rc"].astype(np.int64))
    assert np.array_equal(pl, g["perim_links"].astype(np.int64))


def _dtype_cases():
    g = load_golden("direction_dtypes.npz")
    return sorted({k.split("__")[0] for k in g.files})


@pytest.mark.parametrize("name", _dtype_cases())
def test_direction_dtypes_match_reference(name):
    """float64 and integer DEMs: the reference's arithmetic follows the array dtype (unsigned differences
    wrap); fixtures produced by the reference (oracle/gen_golden_dtypes.py)."""
    g = load_golden("direction_dtypes.npz")
    dem, nodata = g[name + "__dem"], float(g[name + "__nodata"])
    got = oracle.flow_direction_for_tile(dem, nodata)[1:-1, 1:-1]
    assert np.array_equal(got, g[name + "__fdr"])
