"""GPU parity: flow accumulation + perimeter links through the C ABI vs the oracle / golden vectors."""
import numpy as np
import pytest

import oracle
from oracle import synth
from conftest import load_golden
from test_oracle import ACC_CASES

pytestmark = pytest.mark.gpu


def stfa(fdr):
    from overflow_b200.flow_accumulation import single_tile_flow_accumulation

    return single_tile_flow_accumulation(fdr)


def test_kat():
    # reference tests/test_flow_accumulation.py:90-130
    g = load_golden("kat.npz")
    fac, links = stfa(g["acc_fdr"])
    assert fac.dtype == np.int64 and links.shape == (7, 7, 2)
    assert np.array_equal(fac, g["acc_fac"])
    rc = g["acc_perim_rc"]
    assert np.array_equal(links[rc[:, 0], rc[:, 1]], g["acc_perim_links"])


@pytest.mark.parametrize("name", ACC_CASES)
def test_golden(name):
    g = load_golden("accumulation.npz")
    fdr = g[f"{name}_fdr"]
    fac, links = stfa(fdr)
    assert np.array_equal(fac, g[f"{name}_fac"])
    rc = g[f"{name}_perim_rc"]
    assert np.array_equal(links[rc[:, 0], rc[:, 1]], g[f"{name}_perim_links"])


def _cases():
    yield "fractal_b2_holes", synth.punch_holes(synth.fractal(700, 900, beta=2.0, seed=0), frac=0.01, seed=1)
    yield "fractal_b3", synth.fractal(1000, 640, beta=3.0, seed=2)
    yield "fractal_b4", synth.fractal(513, 1027, beta=4.0, seed=3)
    yield "terraced", synth.terraced(600, 800, seed=4)
    yield "tilted", synth.tilted_plane(900, 300)
    yield "tilted_nw", synth.tilted_plane(300, 700, a=-1.0, b=-0.5)
    yield "serpentine", synth.serpentine(259, 263)
    yield "tile_edge_64", synth.fractal(64, 64, beta=2.0, seed=5)
    yield "tile_edge_65", synth.fractal(65, 129, beta=2.0, seed=6)
    yield "thin", synth.fractal(3, 2000, beta=2.0, seed=7)


@pytest.mark.parametrize("name,dem", list(_cases()), ids=[n for n, _ in _cases()])
def test_vs_oracle(name, dem):
    fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1].copy()
    want = oracle.flow_accumulation(fdr)
    rc, want_links = oracle.links_perimeter(fdr)
    fac, links = stfa(fdr)
    assert np.array_equal(fac, want)
    assert np.array_equal(links[rc[:, 0], rc[:, 1]], want_links)


def test_config1_1024_end_to_end():
    """configs[0]: direction then accumulation of the 1024^2 fractal, both on the GPU, vs the oracle."""
    from overflow_b200.flow_direction import flow_direction_for_raster

    dem = synth.fractal(1024, 1024, beta=2.0, seed=0)
    fdr = flow_direction_for_raster(dem, synth.NODATA)
    fac, links = stfa(fdr)
    want_fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
    assert np.array_equal(fdr, want_fdr)
    assert np.array_equal(fac, oracle.flow_accumulation(fdr))
    rc, want_links = oracle.links_perimeter(fdr)
    assert np.array_equal(links[rc[:, 0], rc[:, 1]], want_links)


def test_codes_8_and_9_and_invalid():
    rng = np.random.default_rng(3)
    dem = synth.terraced(200, 210, seed=9)
    fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1].copy()
    assert (fdr == 8).any() and (fdr == 9).any()
    fac, _ = stfa(fdr)
    assert np.all(fac[fdr == 9] == -9998)
    assert np.array_equal(fac, oracle.flow_accumulation(fdr))
    with pytest.raises(ValueError):
        stfa(np.full((4, 4), -1, dtype=np.int64))
    del rng


def test_cycle_is_reported():
    from overflow_b200 import OverflowB200Error

    fdr = np.full((80, 80), 8, dtype=np.uint8)
    fdr[10, 10], fdr[10, 11] = 0, 4  # two cells pointing at each other
    with pytest.raises(OverflowB200Error) as e:
        stfa(fdr)
    assert e.value.status == -5


def test_checker_agrees_with_oracle():
    from overflow_b200.flow_accumulation import check_flow_accumulation

    dem = synth.fractal(300, 333, beta=2.5, seed=12)
    fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1].copy()
    fac = oracle.flow_accumulation(fdr)
    assert check_flow_accumulation(fdr, fac) == 0
    bad = fac.copy()
    bad[17, 19] += 1
    assert check_flow_accumulation(fdr, bad) == oracle.check_accumulation(fdr, bad) >= 1


def test_large_device_pipeline_recurrence():
    """8k x 8k on the device: direction -> accumulation, exactness via the recurrence checker,
    and a cross-check of the whole result against the CPU oracle."""
    import torch

    from overflow_b200 import device as dev

    rows = cols = 8192
    dem = dev.synth_dem(rows, cols, seed=3, kind=0, holes_permille=5)
    fdr = dev.flow_direction(dem, synth.NODATA)
    fac, links = dev.flow_accumulation(fdr, with_links=True)
    assert dev.check_accumulation(fdr, fac) == 0
    h_fdr = fdr.cpu().numpy()
    want = oracle.flow_accumulation(h_fdr)
    assert np.array_equal(fac.cpu().numpy(), want)
    _, want_links = oracle.links_perimeter(h_fdr)
    assert np.array_equal(links.cpu().numpy(), want_links)
    del torch


@pytest.mark.parametrize("kind", [1, 2])
def test_large_device_adversarial(kind):
    from overflow_b200 import device as dev

    rows, cols = 4096, 8192
    dem = dev.synth_dem(rows, cols, seed=1, kind=kind, relief=60.0, holes_permille=10 if kind == 1 else 0)
    fdr = dev.flow_direction(dem, synth.NODATA)
    fac = dev.flow_accumulation(fdr)
    assert dev.check_accumulation(fdr, fac) == 0
    assert np.array_equal(fac.cpu().numpy(), oracle.flow_accumulation(fdr.cpu().numpy()))


def test_config1_1024_reference_golden():
    """configs[0] against the REFERENCE's own output (fixture made by oracle/gen_golden.py --big)."""
    import os

    from conftest import GOLDEN

    if not os.path.exists(os.path.join(GOLDEN, "config1_1024.npz")):
        pytest.skip("big anchor not generated")
    g = load_golden("config1_1024.npz")
    fac, links = stfa(g["fdr"])
    assert np.array_equal(fac, g["fac"].astype(np.int64))
    rc = g["perim_rc"].astype(np.int64)
    assert np.array_equal(links[rc[:, 0], rc[:, 1]], g["perim_links"].astype(np.int64))


def test_perimeter_links_device_large():
    """links of a 2048 x 3072 device raster vs the oracle (paths crossing many tiles)."""
    from overflow_b200 import device as dev

    dem = dev.synth_dem(2048, 3072, seed=21, kind=0, holes_permille=3)
    fdr = dev.flow_direction(dem, synth.NODATA)
    fac, links = dev.flow_accumulation(fdr, with_links=True)
    h = fdr.cpu().numpy()
    _, want = oracle.links_perimeter(h)
    assert np.array_equal(links.cpu().numpy(), want)
    assert np.array_equal(fac.cpu().numpy(), oracle.flow_accumulation(h))


@pytest.mark.parametrize("shape", [(1, 1), (1, 300), (300, 1), (63, 65), (64, 64), (129, 64), (2, 5000)])
def test_degenerate_and_tile_aligned_shapes(shape):
    dem = synth.fractal(shape[0], shape[1], beta=2.0, seed=shape[0] * 7 + shape[1])
    fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1].copy()
    fac, links = stfa(fdr)
    assert np.array_equal(fac, oracle.flow_accumulation(fdr))
    rc, want_links = oracle.links_perimeter(fdr)
    assert np.array_equal(links[rc[:, 0], rc[:, 1]], want_links)


def test_all_cells_one_chain_counts_exceed_32_bits_path():
    """A long single channel: the 64-bit path of the final pass (low word + carry) stays exact."""
    from overflow_b200 import device as dev
    import torch

    # serpentine on the device path; counts reach rows*cols-ish in the channel
    dem = synth.serpentine(513, 517)
    fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1].copy()
    want = oracle.flow_accumulation(fdr)
    pitched = torch.zeros((513, 528), dtype=torch.uint8, device="cuda")[:, :517]
    pitched.copy_(torch.from_numpy(fdr))
    fac = dev.flow_accumulation(pitched)
    assert np.array_equal(fac.cpu().numpy(), want)
    assert want.max() > 100000


def test_wide_final_pass_matches(monkeypatch):
    """The 64-bit variant of the final pass (taken when a count does not fit 32 bits) gives the same
    counts: force every tile with an inflow through it."""
    fdr = oracle.flow_direction_for_tile(
        synth.pad_nodata(synth.punch_holes(synth.fractal(300, 517, beta=2.5, seed=21), frac=0.02, seed=22)),
        synth.NODATA)[1:-1, 1:-1]
    from overflow_b200.flow_accumulation import single_tile_flow_accumulation

    fdr = np.ascontiguousarray(fdr)
    want = oracle.flow_accumulation(fdr)
    monkeypatch.setenv("OFL_FORCE_WIDE_FINAL", "1")
    fac, _ = single_tile_flow_accumulation(fdr)
    assert np.array_equal(fac, want)
    monkeypatch.delenv("OFL_FORCE_WIDE_FINAL")
    fac, _ = single_tile_flow_accumulation(fdr)
    assert np.array_equal(fac, want)


def test_wide_final_pass_many_tiles_per_cta(monkeypatch):
    """More listed tiles than the 64-bit kernel has CTAs (2 per SM): every CTA loops over several tiles,
    re-arming its TMA barrier each time."""
    from overflow_b200.flow_accumulation import flow_accumulation_for_raster

    dem = synth.punch_holes(synth.fractal(1500, 1700, beta=2.0, seed=31), frac=0.01, seed=32)
    fdr = np.ascontiguousarray(oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1])
    want = oracle.flow_accumulation(fdr)
    monkeypatch.setenv("OFL_FORCE_WIDE_FINAL", "1")
    assert np.array_equal(flow_accumulation_for_raster(fdr), want)


def test_random_shapes_counts_and_links():
    """Seeded sweep over ragged shapes around the 64-cell tile size (partial tiles, one-cell-wide
    tiles, nodata on tile corners): counts and perimeter links against the oracle."""
    rng = np.random.default_rng(2024)
    sides = [1, 2, 3, 63, 64, 65, 66, 127, 128, 129, 130, 191, 193, 257]
    for trial in range(40):
        rows, cols = int(rng.choice(sides)), int(rng.choice(sides))
        beta = float(rng.choice([2.0, 3.0, 4.0]))
        dem = synth.fractal(max(rows, 2), max(cols, 2), beta=beta, seed=100 + trial)[:rows, :cols]
        if trial % 3 == 0:
            dem = np.floor(dem / 25.0).astype(np.float32)  # terraces: plateaus of undefined cells
        dem = np.ascontiguousarray(dem)
        if rows * cols > 16 and trial % 2 == 0:
            dem = synth.punch_holes(dem, frac=0.05, seed=trial)
        for y in range(63, rows, 64):  # nodata straddling tile corners
            for x in range(63, cols, 64):
                if (y + x + trial) % 3 == 0:
                    dem[y : y + 2, x : x + 2] = synth.NODATA
        fdr = np.ascontiguousarray(
            oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1])
        fac, links = stfa(fdr)
        want_fac, want_links = oracle.single_tile_flow_accumulation(fdr)
        assert np.array_equal(fac, want_fac), (trial, rows, cols)
        assert np.array_equal(links, want_links), (trial, rows, cols)
