"""The device DEM generator (csrc/synth.cu) against its host restatement (oracle/synth_host.c), bit for bit, and the
serpentine DEM (SURVEY 8d config 5 ii) through the whole device path."""
import numpy as np
import pytest
import torch

import oracle
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,holes,relief", [(0, 5, 1000.0), (0, 0, 37.5), (1, 50, 200.0), (2, 0, 1000.0), (3, 0, 1000.0), (4, 0, 1000.0)])
@pytest.mark.parametrize("rows,cols,row0,total", [(300, 517, 0, 300), (130, 1024, 65000, 65536), (66, 200, -1, 64)])
def test_device_generator_equals_host_restatement(kind, holes, relief, rows, cols, row0, total):
    from overflow_b200 import device as dev

    d = dev.synth_dem(rows, cols, row0=row0, total_rows=total, seed=9, kind=kind, relief=relief, holes_permille=holes)
    h = oracle.synth_dem(rows, cols, row0=row0, total_rows=total, total_cols=cols, seed=9, kind=kind, relief=relief,
                         holes_permille=holes)
    assert np.array_equal(d.cpu().numpy().view(np.uint32), h.view(np.uint32))


@pytest.mark.parametrize("shape", [(1024, 768), (513, 517), (2048, 4096)])
def test_serpentine_device_path_vs_oracle(shape):
    """One channel of rows*cols/2 cells through every tile: direction and accumulation equal the oracle's."""
    from overflow_b200 import device as dev

    rows, cols = shape
    dem = dev.synth_dem(rows, cols, kind=3)
    fdr, fac = dev.flow_routing(dem, synth.NODATA)
    h_dem = dem.cpu().numpy()
    want_fdr = oracle.flow_direction_for_tile(synth.pad_nodata(h_dem), synth.NODATA)[1:-1, 1:-1]
    assert np.array_equal(fdr.cpu().numpy(), want_fdr)
    want = oracle.flow_accumulation(np.ascontiguousarray(want_fdr))
    assert np.array_equal(fac.cpu().numpy(), want)
    assert int((want_fdr == 8).sum()) == 1 and want.max() > rows * cols // 2
    assert dev.check_accumulation(fdr, fac) == 0


@pytest.mark.parametrize("kind", [3, 4])
def test_serpentine_across_strips(kind):
    """One channel through every strip (kind 3: east-west runs, it crosses each strip boundary once; kind 4: north-south
    runs, it crosses every strip boundary cols / 2 times): 1, 3 and 8 strips equal the single raster."""
    from overflow_b200 import device as dev, strips

    rows, cols = 1024, 320
    dem = dev.synth_dem(rows, cols, kind=kind)
    one_fdr, one_fac = dev.flow_routing(dem, synth.NODATA)
    assert int(one_fac.max().item()) > rows * cols // 2
    for world in (1, 3, 8):
        pipes = [strips.StripPipeline(rows, cols, r, world, nodata=synth.NODATA, device="cuda:0") for r in range(world)]
        for p in pipes:
            p.load_synthetic(kind=kind)
        strips.step_in_process(pipes)
        assert torch.equal(torch.cat([p.fdr for p in pipes]), one_fdr)
        assert torch.equal(torch.cat([p.fac for p in pipes]), one_fac)
        assert strips.check_in_process(pipes) == 0


@pytest.mark.parametrize("cols", [130, 517, 1001])
def test_dense_odd_width_tensors_through_the_device_api(cols):
    """dev.flow_direction -> dev.flow_accumulation on dense tensors whose width is not a multiple of 16."""
    from overflow_b200 import device as dev

    dem = synth.punch_holes(synth.fractal(300, cols, beta=2.0, seed=cols), frac=0.01, seed=2)
    d_dem = torch.from_numpy(dem).cuda()  # dense: row pitch = cols
    fdr = dev.flow_direction(d_dem, synth.NODATA)
    want_fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
    assert np.array_equal(fdr.cpu().numpy(), want_fdr)
    fac = dev.flow_accumulation(fdr.contiguous())  # dense again
    assert np.array_equal(fac.cpu().numpy(), oracle.flow_accumulation(np.ascontiguousarray(want_fdr)))
    fdr2, fac2 = dev.flow_routing(d_dem, synth.NODATA)
    assert torch.equal(fdr2, fdr) and torch.equal(fac2, fac)
    with pytest.raises(ValueError):
        dev.flow_accumulation(fdr, out=torch.empty((300, cols), dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        dev.flow_accumulation(fdr, workspace=torch.empty((16,), dtype=torch.uint8, device="cuda"))
    with pytest.raises(ValueError):
        dev.check_accumulation(fdr, fac[:, :-1])
