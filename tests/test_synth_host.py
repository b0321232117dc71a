"""The host restatements of the device DEM generator (oracle/synth_host.c, oracle/synth.py::device_dem) agree bit
for bit, and the serpentine DEM (SURVEY 8d config 5 ii) is what it claims to be: one channel, one interior outlet."""
import numpy as np
import pytest

import oracle
from oracle import synth


@pytest.mark.parametrize("kind,holes", [(0, 5), (0, 0), (1, 50), (2, 0), (3, 0), (4, 0)])
@pytest.mark.parametrize("window", [(0, 0, 97, 130), (65400, 3000, 136, 70), (-1, 0, 40, 64)])
def test_c_and_numpy_generators_agree(kind, holes, window):
    row0, col0, rows, cols = window
    kw = dict(row0=row0, col0=col0, total_rows=65536, total_cols=65536, seed=11, kind=kind, holes_permille=holes)
    a = oracle.synth_dem(rows, cols, **kw)
    b = synth.device_dem(rows, cols, **kw)
    assert a.dtype == np.float32 and a.shape == (rows, cols)
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    if row0 < 0:
        assert np.all(a[0] == np.float32(synth.NODATA))  # a strip's halo row above the raster


def test_holes_fraction_and_seed():
    a = oracle.synth_dem(1024, 1024, seed=1, kind=0, holes_permille=50)
    frac = float((a == np.float32(synth.NODATA)).mean())
    assert 0.02 < frac < 0.09
    assert not np.array_equal(a, oracle.synth_dem(1024, 1024, seed=2, kind=0, holes_permille=50))
    t = oracle.synth_dem(256, 256, seed=1, kind=1, relief=50.0)
    assert np.array_equal(t, np.floor(t))


@pytest.mark.parametrize("shape", [(9, 8), (10, 7), (64, 64), (131, 70), (257, 129)])
def test_serpentine_is_one_chain_with_one_outlet(shape):
    rows, cols = shape
    z = oracle.synth_dem(rows, cols, kind=3)
    chan = z != synth.SERP_WALL
    # strictly decreasing along the channel: all channel values distinct, none equal to nodata
    vals = z[chan]
    assert len(np.unique(vals)) == vals.size and not np.any(vals == np.float32(synth.NODATA))
    fdr = np.ascontiguousarray(oracle.flow_direction_for_tile(synth.pad_nodata(z), synth.NODATA)[1:-1, 1:-1])
    assert int((fdr == 8).sum()) == 1  # the single interior outlet (a pit at the end of the channel)
    fac = oracle.flow_accumulation(fdr)
    py, px = np.argwhere(fdr == 8)[0]
    assert fac[py, px] == fac.max()
    # every cell that does not drain off the raster ends in the pit
    ring = np.zeros_like(chan)
    ring[0, :] = ring[-1, :] = ring[:, 0] = ring[:, -1] = True
    assert fac.max() >= chan.sum() and fac.max() <= rows * cols - ring.sum() + 2 * (rows + cols)
    assert oracle.check_accumulation(fdr, fac) == 0


def test_serpentine_window_equals_full_raster():
    full = oracle.synth_dem(200, 90, kind=3)
    win = oracle.synth_dem(64, 30, row0=100, col0=50, total_rows=200, total_cols=90, kind=3)
    assert np.array_equal(full[100:164, 50:80].view(np.uint32), win.view(np.uint32))
