"""GPU parity: flow direction through the C ABI vs the CPU oracle and the reference's golden vectors."""
import zlib

import numpy as np
import pytest

import oracle
from oracle import synth
from conftest import load_golden

pytestmark = pytest.mark.gpu


def fd_tile(dem, nodata):
    from overflow_b200.flow_direction import flow_direction_for_tile

    return flow_direction_for_tile(dem, nodata)


def fd_raster(dem, nodata):
    from overflow_b200.flow_direction import flow_direction_for_raster

    return flow_direction_for_raster(dem, nodata)


def test_kat_from_dem():
    # reference tests/test_flow_direction.py:131-136
    g = load_golden("kat.npz")
    fdr = fd_tile(g["dir_dem"], -9999)
    assert fdr.dtype == np.uint8 and fdr.shape == g["dir_dem"].shape
    assert np.array_equal(fdr[1:-1, 1:-1], g["dir_expected"])
    assert np.all(fdr[0] == 9) and np.all(fdr[:, 0] == 9)


def test_kat_raster_mode_equals_padded_tile():
    g = load_golden("kat.npz")
    assert np.array_equal(fd_raster(g["dir_dem"][1:-1, 1:-1].copy(), -9999), g["dir_expected"])


def test_discriminating_vectors():
    g = load_golden("discriminating.npz")
    for k, (tile, want) in enumerate(zip(g["tiles"], g["centre"])):
        assert fd_tile(tile, float(g["nodata"]))[1, 1] == want, f"vector {k}"
    assert fd_tile(g["unrep_tile"], float(g["unrep_nodata"]))[1, 1] == g["unrep_centre"]


@pytest.mark.parametrize("kind", synth.FUZZ_KINDS + ("nan_nodata",))
def test_fuzz_golden(kind):
    g = load_golden("direction_fuzz.npz")
    nodata = float("nan") if kind == "nan_nodata" else synth.NODATA
    fdr = fd_tile(g[f"{kind}_dem"], nodata)
    assert np.array_equal(fdr[1:-1, 1:-1], g[f"{kind}_fdr"])


@pytest.mark.parametrize("kind", synth.FUZZ_KINDS)
@pytest.mark.parametrize("shape", [(3, 3), (4, 131), (67, 5), (130, 259), (300, 1030)])
def test_fuzz_vs_oracle(kind, shape):
    dem = synth.fuzz_dem(kind, shape[0], shape[1], seed=zlib.crc32(repr((kind, shape)).encode()) % 10007)
    want = oracle.flow_direction_for_tile(dem, synth.NODATA)
    got = fd_tile(dem, synth.NODATA)
    assert np.array_equal(got[1:-1, 1:-1], want[1:-1, 1:-1])


@pytest.mark.parametrize("shape", [(1, 1), (1, 7), (9, 1), (2, 2), (64, 128), (65, 129), (513, 257)])
def test_raster_mode_vs_padded_oracle(shape):
    dem = synth.punch_holes(synth.fractal(shape[0], shape[1], beta=2.0, seed=3), frac=0.02, seed=5)
    want = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
    assert np.array_equal(fd_raster(dem, synth.NODATA), want)


@pytest.mark.parametrize("nodata", [0.0, -0.0, np.inf, -np.inf, 3.5, -1.1, 1e300])
def test_nodata_values(nodata):
    rng = np.random.default_rng(11)
    dem = rng.integers(-3, 4, (60, 70)).astype(np.float32)
    dem[rng.random(dem.shape) < 0.1] = np.float32(3.5)
    dem[rng.random(dem.shape) < 0.05] = np.inf
    dem[rng.random(dem.shape) < 0.05] = -np.inf
    dem[rng.random(dem.shape) < 0.05] = -0.0
    want = oracle.flow_direction_for_tile(dem, nodata)
    got = fd_tile(dem, nodata)
    assert np.array_equal(got[1:-1, 1:-1], want[1:-1, 1:-1])
    # raster mode pads with float32(nodata), like util/raster.py:67 does
    pad = np.full((62, 72), np.float32(nodata), dtype=np.float32)
    pad[1:-1, 1:-1] = dem
    assert np.array_equal(fd_raster(dem, nodata), oracle.flow_direction_for_tile(pad, nodata)[1:-1, 1:-1])


def test_integer_dem_dtypes():
    rng = np.random.default_rng(2)
    dem = rng.integers(-300, 300, (40, 50)).astype(np.int16)
    want = oracle.flow_direction_for_tile(dem.astype(np.float64), -300)
    assert np.array_equal(fd_tile(dem, -300)[1:-1, 1:-1], want[1:-1, 1:-1])
    assert np.array_equal(fd_tile(dem.astype(np.float64), -300)[1:-1, 1:-1], want[1:-1, 1:-1])
    assert np.array_equal(fd_tile(dem.astype(np.int64), -300)[1:-1, 1:-1], want[1:-1, 1:-1])


def test_config1_1024_fractal():
    # BASELINE configs[0]: 1024x1024 float32 fractal, single tile, nodata ring
    for beta, seed in ((2.0, 0), (3.0, 1)):
        dem = synth.pad_nodata(synth.fractal(1024, 1024, beta=beta, seed=seed))
        want = oracle.flow_direction_for_tile(dem, synth.NODATA)
        got = fd_tile(dem, synth.NODATA)
        assert np.array_equal(got[1:-1, 1:-1], want[1:-1, 1:-1])


def test_flat_heavy_and_adversarial():
    for dem in (synth.terraced(700, 900, seed=4), synth.tilted_plane(300, 500), synth.serpentine(257, 263)):
        want = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
        assert np.array_equal(fd_raster(dem, synth.NODATA), want)


def test_tiling_invariance_device_strip_mode():
    """Row strips with halo rows give the same codes as the whole raster."""
    import torch

    from overflow_b200 import device as dev

    dem = synth.punch_holes(synth.fractal(512, 384, beta=2.0, seed=8), frac=0.01, seed=9)
    whole = fd_raster(dem, synth.NODATA)
    pad = synth.pad_nodata(dem)[:, 1:-1]  # nodata rows above and below only
    parts = []
    for r0, r1 in ((0, 100), (100, 101), (101, 384), (384, 512)):
        strip = torch.from_numpy(np.ascontiguousarray(pad[r0 : r1 + 2])).cuda()
        parts.append(dev.flow_direction(strip, synth.NODATA, mode="strip").cpu().numpy())
    assert np.array_equal(np.concatenate(parts), whole)


def test_large_device_raster_windows_vs_oracle():
    """16k x 4k synthetic raster generated on the device; random windows re-checked by the oracle."""
    import torch

    from overflow_b200 import device as dev

    rows, cols = 16384, 4096
    dem = dev.synth_dem(rows, cols, seed=5, kind=0, holes_permille=5)
    fdr = dev.flow_direction(dem, synth.NODATA)
    torch.cuda.synchronize()
    rng = np.random.default_rng(0)
    wins = [(0, 0), (rows - 400, cols - 400), (0, cols - 400), (rows - 400, 0)]
    wins += [(int(rng.integers(1, rows - 401)), int(rng.integers(1, cols - 401))) for _ in range(6)]
    for r, c in wins:
        r0, c0 = max(r - 1, 0), max(c - 1, 0)
        win = dem[r0 : r + 401, c0 : c + 401].cpu().numpy()
        want = oracle.flow_direction_for_tile(win, synth.NODATA)[1:-1, 1:-1]
        got = fdr[r0 + 1 : r0 + win.shape[0] - 1, c0 + 1 : c0 + win.shape[1] - 1].cpu().numpy()
        assert np.array_equal(got, want)
    # edges of the raster against a nodata-padded oracle run
    top = dem[:3].cpu().numpy()
    want = oracle.flow_direction_for_tile(synth.pad_nodata(top), synth.NODATA)[1:2, 1:-1]
    assert np.array_equal(fdr[:1].cpu().numpy(), want)


def _dtype_cases():
    g = load_golden("direction_dtypes.npz")
    return sorted({k.split("__")[0] for k in g.files})


@pytest.mark.parametrize("name", _dtype_cases())
def test_dtypes_match_reference(name):
    """float64 / integer DEMs through ofl_flow_direction_x64 (and int8 / int16 through the float32 kernel)
    against the reference's own outputs."""
    from overflow_b200.flow_direction import flow_direction_for_tile

    g = load_golden("direction_dtypes.npz")
    dem, nodata = g[name + "__dem"], float(g[name + "__nodata"])
    got = flow_direction_for_tile(dem, nodata)
    assert got.dtype == np.uint8 and got.shape == dem.shape
    assert np.array_equal(got[1:-1, 1:-1], g[name + "__fdr"])
    assert (got[0] == 9).all() and (got[:, 0] == 9).all()


@pytest.mark.parametrize("dtype", [np.float64, np.int32, np.uint16])
def test_dtypes_raster_mode_matches_oracle(dtype):
    """Whole-raster mode (out-of-raster neighbours read as nodata cast to the dtype) for the generic kernel."""
    from overflow_b200.flow_direction import flow_direction_for_raster

    rng = np.random.default_rng(5)
    dem = rng.integers(0, 500, size=(150, 203)).astype(dtype)
    nodata = 65535 if dtype == np.uint16 else -9999
    dem[40:50, 60:80] = nodata
    pad = np.full((152, 205), nodata, dtype=dtype)
    pad[1:-1, 1:-1] = dem
    want = oracle.flow_direction_for_tile(pad, float(nodata))[1:-1, 1:-1]
    assert np.array_equal(flow_direction_for_raster(dem, float(nodata)), want)


def test_unsupported_dtype_raises():
    from overflow_b200.flow_direction import flow_direction_for_tile

    with pytest.raises(TypeError):
        flow_direction_for_tile(np.zeros((5, 5), dtype=np.float16), -9999.0)


def test_near_ties_between_cardinal_and_diagonal():
    """Stress the guard band of the float32 class decision: windows whose best diagonal drop is within a
    few ulps of sqrt(2) times the best cardinal drop, at many magnitudes, signs and with plateaus around
    them; the exact path must take exactly the cases the fast path cannot decide."""
    rng = np.random.default_rng(77)
    rows, cols = 300, 512
    dem = np.full((rows, cols), 1000.0, dtype=np.float32)
    for r in range(1, rows - 1, 3):
        for c in range(1, cols - 1, 3):
            z = np.float32(rng.choice([1.0, 37.5, 1e3, 1e6, 1e-3, 3e-30]) * rng.uniform(0.5, 2.0))
            drop_c = np.float32(abs(z) * rng.choice([1e-6, 1e-3, 0.25, 0.9]))
            ulps = int(rng.integers(-6, 7))
            drop_d = np.float32(np.float64(drop_c) * np.sqrt(2.0))
            for _ in range(abs(ulps)):
                drop_d = np.nextafter(drop_d, np.float32(np.inf if ulps > 0 else -np.inf), dtype=np.float32)
            win = np.full((3, 3), z, dtype=np.float32)
            win[1, 1] = z
            ci = [(1, 2), (0, 1), (1, 0), (2, 1)][int(rng.integers(0, 4))]
            di = [(0, 2), (0, 0), (2, 0), (2, 2)][int(rng.integers(0, 4))]
            win[ci] = z - drop_c
            win[di] = z - drop_d
            dem[r - 1 : r + 2, c - 1 : c + 2] = win
    want = oracle.flow_direction_for_tile(dem, synth.NODATA)
    got = fd_tile(dem, synth.NODATA)
    assert np.array_equal(got[1:-1, 1:-1], want[1:-1, 1:-1])


def test_heavy_nodata_and_non_finite_values():
    """A quarter of the cells NODATA in blobs and singles, plus NaN / +-inf / denormal data cells: every
    combination of +inf slopes in the two neighbour classes, through the fast path and the fix-up."""
    rng = np.random.default_rng(78)
    dem = synth.fractal(700, 900, beta=2.0, seed=12)
    mask = rng.random(dem.shape)
    dem[mask < 0.15] = synth.NODATA
    for _ in range(300):
        r, c = int(rng.integers(0, 690)), int(rng.integers(0, 890))
        dem[r : r + int(rng.integers(1, 9)), c : c + int(rng.integers(1, 9))] = synth.NODATA
    special = rng.random(dem.shape)
    dem[special < 0.002] = np.nan
    dem[(special >= 0.002) & (special < 0.004)] = np.inf
    dem[(special >= 0.004) & (special < 0.006)] = -np.inf
    dem[(special >= 0.006) & (special < 0.008)] = np.float32(1e-42)
    want = oracle.flow_direction_for_tile(dem, synth.NODATA)
    got = fd_tile(dem, synth.NODATA)
    assert np.array_equal(got[1:-1, 1:-1], want[1:-1, 1:-1])
    pad = synth.pad_nodata(dem)
    assert np.array_equal(fd_raster(dem, synth.NODATA), oracle.flow_direction_for_tile(pad, synth.NODATA)[1:-1, 1:-1])
