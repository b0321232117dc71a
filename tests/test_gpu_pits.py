"""GPU parity: single-cell pit breaching (csrc/pits.cu) through the C ABI vs fixtures produced by the reference's
breach_single_cell_pits_in_chunk and vs the C oracle, bit for bit; file driver and CLI as in the reference's
tests/test_breach_single_cell_pits.py (built-in GeoTIFF I/O standing in for GDAL's /vsimem)."""
import click.testing
import numpy as np
import pytest

import oracle
from oracle import synth
from conftest import load_golden

pytestmark = pytest.mark.gpu

Z = load_golden("breach_pits.npz")
NAMES = sorted({k.split("__")[0] for k in Z.files})


def same_bits(a, b):
    """Bit-identical elevations; a NaN matches a NaN whatever its payload (x86 hands an operand's payload on, the
    GPU writes the canonical NaN -- the reference's own test compares with allclose)."""
    if a.shape != b.shape:
        return False
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(nan_a, nan_b) and np.array_equal(a.view(np.uint32)[~nan_a], b.view(np.uint32)[~nan_b]))


def breach(chunk, nodata, **kw):
    from overflow.breach_single_cell_pits import breach_single_cell_pits_in_chunk

    return breach_single_cell_pits_in_chunk(chunk, nodata, **kw)


def test_ref_breach_single_cell_pits_in_chunk():
    """tests/test_breach_single_cell_pits.py:62-83 of the reference, same call, same expectation."""
    dem_chunk = Z["kat__chunk_in"].copy()
    expected = dem_chunk.copy()
    expected[4, 3] = -0.5
    breach(dem_chunk, -999)
    assert np.allclose(dem_chunk, expected)


@pytest.mark.parametrize("name", NAMES)
def test_golden(name):
    chunk = Z[f"{name}__chunk_in"].copy()
    unsolved = breach(chunk, float(Z[f"{name}__nodata"]))
    assert unsolved.dtype == np.int8 and np.array_equal(unsolved, Z[f"{name}__unsolved"])
    assert same_bits(chunk, Z[f"{name}__chunk_out"])


def _cases():
    rng = np.random.default_rng(5)
    yield "fractal", synth.pad_nodata(synth.pad_nodata(synth.punch_holes(synth.fractal(700, 900, beta=2.0, seed=0), frac=0.01, seed=1)))
    yield "rough_fractal", synth.fractal(513, 1027, beta=1.2, seed=3)
    yield "uniform", rng.uniform(0, 1000, (600, 800)).astype(np.float32)
    yield "ints", rng.integers(0, 5, (400, 333)).astype(np.float32)
    yield "special", synth.fuzz_dem("special", 300, 300, seed=2)
    lat = rng.uniform(50, 60, (301, 402)).astype(np.float32)
    lat[2:-2:2, 2:-2:2] = rng.uniform(0, 40, lat[2:-2:2, 2:-2:2].shape).astype(np.float32)
    yield "lattice", lat
    yield "thin", rng.uniform(0, 10, (5, 4000)).astype(np.float32)


@pytest.mark.parametrize("name,dem", list(_cases()), ids=[n for n, _ in _cases()])
def test_vs_oracle(name, dem):
    want, want_unsolved = oracle.breach_single_cell_pits_in_chunk(dem, synth.NODATA)
    chunk = dem.copy()
    unsolved, info = breach(chunk, synth.NODATA, return_info=True)
    assert np.array_equal(unsolved, want_unsolved)
    assert same_bits(chunk, want)
    assert info["unsolved"] == int(want_unsolved.sum()) and info["pits"] >= info["unsolved"]


def test_strided_chunk_and_dtype_rule():
    big = np.random.default_rng(1).uniform(0, 10, (64, 128)).astype(np.float32)
    view = big[:, ::2]  # not contiguous: handled through a copy, written back in place
    want, _ = oracle.breach_single_cell_pits_in_chunk(np.ascontiguousarray(view), synth.NODATA)
    breach(view, synth.NODATA)
    assert same_bits(np.ascontiguousarray(view), want)
    with pytest.raises(TypeError):
        breach(np.zeros((8, 8), dtype=np.float64), -9999.0)


@pytest.fixture
def raster_file_path(tmp_path):
    """The reference's file fixture (tests/test_breach_single_cell_pits.py:12-36): 5x5, nodata -inf."""
    from overflow_b200.util.raster import create_raster

    path = str(tmp_path / "test_raster_breach.tif")
    ds = create_raster(path, 5, 5, "Float32")
    band = ds.GetRasterBand(1)
    band.WriteArray(np.array([[2, 2, 2, 2, 2], [-1, 2, 2, 2, 2], [2, 2, 0, 2, 2], [2, 2, 2, 2, 2], [2, 2, 2, 2, 2]],
                             dtype=np.float32))
    band.SetNoDataValue(-np.inf)
    ds.FlushCache()
    return path


EXPECTED_FILE = np.array([[2, 2, 2, 2, 2], [-1, 2, 2, 2, 2], [2, -0.5, 0, 2, 2], [2, 2, 2, 2, 2], [2, 2, 2, 2, 2]],
                         dtype=np.float32)


def test_breach_single_cell_pits_file(raster_file_path, tmp_path):
    from overflow.breach_single_cell_pits import breach_single_cell_pits
    from overflow_b200.util.raster import open_raster

    out = str(tmp_path / "breached.tif")
    breach_single_cell_pits(raster_file_path, out, chunk_size=5)
    band = open_raster(out).GetRasterBand(1)
    assert np.allclose(band.ReadAsArray(), EXPECTED_FILE)
    assert band.GetNoDataValue() == -np.inf


def test_breach_single_cell_pits_cli(raster_file_path, tmp_path):
    from overflow_b200.util.raster import open_raster
    from overflow_cli import breach_single_cell_pits_cli

    out = str(tmp_path / "breached_cli.tif")
    result = click.testing.CliRunner().invoke(
        breach_single_cell_pits_cli, ["--input_file", raster_file_path, "--output_file", out, "--chunk_size", "5"])
    assert result.exit_code == 0
    assert np.allclose(open_raster(out).GetRasterBand(1).ReadAsArray(), EXPECTED_FILE)


def test_file_driver_chunking_matches_the_reference_loop(tmp_path):
    """Chunks are read from the INPUT file with a two-cell buffer, so a chunk never sees its neighbours' breaches:
    the driver must reproduce that chunk loop (stated here with the oracle), not the whole-raster result."""
    from overflow.breach_single_cell_pits import breach_single_cell_pits
    from overflow_b200.util.raster import create_raster, open_raster

    dem = np.random.default_rng(8).uniform(0, 100, (70, 90)).astype(np.float32)
    src = str(tmp_path / "dem.tif")
    ds = create_raster(src, 90, 70, "Float32")
    ds.GetRasterBand(1).WriteArray(dem)
    ds.GetRasterBand(1).SetNoDataValue(-9999.0)
    ds.FlushCache()
    out = str(tmp_path / "out.tif")
    size = 32
    breach_single_cell_pits(src, out, chunk_size=size)
    want = np.empty_like(dem)
    pad = np.full((70 + 2 * size, 90 + 2 * size), -9999.0, dtype=np.float32)
    pad[2 : 72, 2 : 92] = dem
    for r0 in range(0, 70, size):
        for c0 in range(0, 90, size):
            chunk = pad[r0 : r0 + size + 4, c0 : c0 + size + 4]
            res, _ = oracle.breach_single_cell_pits_in_chunk(chunk, -9999.0)
            h, w = min(size, 70 - r0), min(size, 90 - c0)
            want[r0 : r0 + h, c0 : c0 + w] = res[2 : 2 + h, 2 : 2 + w]
    assert same_bits(open_raster(out).GetRasterBand(1).ReadAsArray(), want)


def test_chunk_of_more_than_2_31_cells():
    """Cell indices are unsigned 32-bit: a 40000 x 65536 chunk (2.6e9 cells).  Terrain in a band at the top and in a
    band below cell 2^31, a plateau (no pits) in between; each band equals the oracle run on the band alone."""
    import torch

    from overflow_b200 import device as dev

    rows, cols, band = 40000, 65536, 192
    free, _ = torch.cuda.mem_get_info()
    if free < 40 * 2**30:
        pytest.skip("needs 40 GB of device memory")
    chunk = torch.full((rows, cols), 5000.0, dtype=torch.float32, device="cuda")
    lo = rows - band - 64  # far below row 32768, where the cell index passes 2^31
    width = 4000  # terrain in columns 8 .. 8 + width only: a plateau (no pits) everywhere else
    chunk[8 : 8 + band, 8 : 8 + width] = dev.synth_dem(band, width, seed=3, kind=0, holes_permille=5)
    chunk[lo : lo + band, 8 : 8 + width] = dev.synth_dem(band, width, seed=4, kind=0, holes_permille=5)
    want = []
    for r0 in (8, lo):  # the terrain with its plateau margin: exactly what the whole chunk shows the pits there
        w = chunk[r0 - 6 : r0 + band + 6, : width + 16].cpu().numpy().copy()
        want.append(oracle.breach_single_cell_pits_in_chunk(w, synth.NODATA))
    unsolved, info = dev.breach_single_cell_pits(chunk, synth.NODATA)
    assert info[0] > 5000 and info[2] >= 2
    n_pits = 0
    for r0, (w_chunk, w_uns) in zip((8, lo), want):
        got = chunk[r0 - 6 : r0 + band + 6, : width + 16].cpu().numpy()
        assert np.array_equal(got.view(np.uint32), w_chunk.view(np.uint32))
        # the oracle's unsolved raster counts the window's pits that stayed unsolved; its pits are the chunk's
        assert np.array_equal(unsolved[r0 - 6 : r0 + band + 6, : width + 16].cpu().numpy(), w_uns)
        n_pits += 1
    assert int(unsolved.sum().item()) == info[1]
