"""GPU: rectangular tiles through the device -- the seeded accumulation entry point (inflow on all
four sides of a tile) and the two-pass tiled driver against the whole-raster result."""
import numpy as np
import pytest

import oracle
from oracle import synth
from overflow_b200 import tiles

pytestmark = pytest.mark.gpu


def codes_of(dem):
    return np.ascontiguousarray(oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1])


def cases():
    yield "fractal", synth.punch_holes(synth.fractal(700, 900, beta=2.0, seed=1), frac=0.02, seed=2)
    yield "terraced", synth.terraced(512, 640, seed=3)
    yield "tilted", synth.tilted_plane(600, 500, a=1.0, b=0.25)
    yield "serpentine", oracle.synth_dem(513, 389, kind=3)
    yield "diagonal", (np.add.outer(np.arange(400), np.arange(520)) * -1.0).astype(np.float32)


@pytest.mark.parametrize("name,dem", list(cases()), ids=[n for n, _ in cases()])
@pytest.mark.parametrize("tile", [(256, 256), (200, 333), (64, 1000), (1000, 100), (4096, 4096)])
def test_tiled_accumulation_on_the_device(name, dem, tile):
    fdr = codes_of(dem)
    want = oracle.flow_accumulation(fdr)
    rows, cols = fdr.shape
    got = np.full((rows, cols), -7, dtype=np.int64)

    def write_window(r0, c0, fac):
        got[r0 : r0 + fac.shape[0], c0 : c0 + fac.shape[1]] = fac

    tiles.flow_accumulation_tiled(lambda r0, r1, c0, c1: fdr[r0:r1, c0:c1], write_window, rows, cols, *tile)
    assert np.array_equal(got, want)


def test_seeded_accumulation_inflow_on_all_four_sides():
    """One tile cut out of a larger raster: seeded with what the rest of the raster sends into its perimeter cells, the
    device returns the cells' whole-raster counts."""
    dem = synth.punch_holes(synth.fractal(300, 340, beta=2.5, seed=5), frac=0.02, seed=6)
    fdr = codes_of(dem)
    whole = oracle.flow_accumulation(fdr)
    r0, r1, c0, c1 = 70, 231, 90, 277
    tile = np.ascontiguousarray(fdr[r0:r1, c0:c1])
    h, w = tile.shape
    pr, pc = tiles.perimeter_cells(h, w)
    first = tiles.perimeter_rank(pr, pc, h, w) == np.arange(len(pr))
    inflow = np.zeros(len(pr), dtype=np.int64)
    dy, dx = [0, -1, -1, -1, 0, 1, 1, 1], [1, 1, 0, -1, -1, -1, 0, 1]
    for k in np.nonzero(first)[0]:
        if tile[pr[k], pc[k]] == 9:
            continue
        for d in range(8):  # neighbours outside the tile that flow into this perimeter cell
            y, x = r0 + pr[k] + dy[d], c0 + pc[k] + dx[d]
            inside_tile = r0 <= y < r1 and c0 <= x < c1
            if not inside_tile and 0 <= y < 300 and 0 <= x < 340 and fdr[y, x] == (d + 4) % 8:
                inflow[k] += whole[y, x]
    assert (inflow > 0).sum() > 20 and {int(pr[k]) for k in np.nonzero(inflow)[0]} >= {0, h - 1}
    fac, links = tiles.CudaTileEngine().accumulate(tile, inflow, True)
    assert np.array_equal(fac, whole[r0:r1, c0:c1])
    _, want_links = oracle.links_perimeter(tile)
    assert np.array_equal(links, want_links)  # the links do not depend on the inflow
    # without the inflow: the tile-local counts
    fac0, _ = tiles.CudaTileEngine().accumulate(tile, None, False)
    assert np.array_equal(fac0, oracle.flow_accumulation(tile))


def test_tiled_device_raster_4k():
    """4096 x 3072 device-generated codes, 1500 x 1100 tiles: equals the single-raster device result."""
    from overflow_b200 import device as dev

    dem = dev.synth_dem(4096, 3072, seed=4, kind=0, holes_permille=8)
    fdr_d, fac_d = dev.flow_routing(dem, synth.NODATA)
    fdr = fdr_d.cpu().numpy()
    got = np.zeros(fdr.shape, dtype=np.int64)

    def write_window(r0, c0, fac):
        got[r0 : r0 + fac.shape[0], c0 : c0 + fac.shape[1]] = fac

    n = tiles.flow_accumulation_tiled(lambda r0, r1, c0, c1: fdr[r0:r1, c0:c1], write_window, 4096, 3072, 1500, 1100)
    assert n == 9 and np.array_equal(got, fac_d.cpu().numpy())
