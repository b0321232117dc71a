"""GPU: the fused host call (ofl_flow_routing_f32) and the banded host pipeline behind the host-memory
paths equal the two separate calls and the oracle, bit for bit."""
import numpy as np
import pytest

import oracle
from oracle import synth
from overflow_b200.flow_accumulation import flow_accumulation_for_raster
from overflow_b200.flow_direction import flow_direction_for_raster
from overflow_b200.flow_routing import flow_routing_for_raster

pytestmark = pytest.mark.gpu


def _oracle(dem):
    fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
    return fdr, oracle.flow_accumulation(np.ascontiguousarray(fdr))


@pytest.mark.parametrize("shape", [(1, 1), (3, 200), (257, 130), (700, 513)])
def test_fused_call_matches_oracle(shape):
    dem = synth.punch_holes(synth.fractal(*shape, beta=2.5, seed=3), frac=0.02, seed=4)
    want_fdr, want_fac = _oracle(dem)
    fdr, fac, links = flow_routing_for_raster(dem, synth.NODATA, with_links=True)
    assert np.array_equal(fdr, want_fdr)
    assert np.array_equal(fac, want_fac)
    rc, want_links = oracle.links_perimeter(np.ascontiguousarray(want_fdr))
    full = np.zeros(shape + (2,), dtype=np.int64)
    from overflow_b200.flow_accumulation import perimeter_indices

    idx = np.asarray(perimeter_indices(shape), dtype=np.int64).reshape(-1, 2)
    if len(idx):
        full[idx[:, 0], idx[:, 1]] = links
        assert np.array_equal(full[rc[:, 0], rc[:, 1]], want_links)


def test_fused_call_without_codes():
    dem = synth.fractal(300, 300, beta=2.0, seed=9)
    fdr, fac = flow_routing_for_raster(dem, synth.NODATA, want_fdr=False)
    assert fdr is None
    assert np.array_equal(fac, _oracle(dem)[1])


def test_many_bands_equal_separate_calls(monkeypatch):
    """64-row bands (71 of them, ragged last band), nodata on band boundaries: same bits as one band."""
    rows, cols = 4490, 2100
    dem = synth.fractal(rows, cols, beta=2.0, seed=11)
    dem[63:66, ::7] = synth.NODATA
    dem[64 * 30, :] = synth.NODATA
    fdr_one = flow_direction_for_raster(dem, synth.NODATA)
    fac_one = flow_accumulation_for_raster(fdr_one)
    monkeypatch.setenv("OFL_PIPE_BAND_BYTES", str(64 * cols * 4))
    fdr_sep = flow_direction_for_raster(dem, synth.NODATA)
    fdr, fac = flow_routing_for_raster(dem, synth.NODATA)
    assert np.array_equal(fdr_sep, fdr_one)
    assert np.array_equal(fdr, fdr_one)
    assert np.array_equal(fac, fac_one)
    # oracle on a window that spans several band boundaries
    want = oracle.flow_direction_for_tile(np.ascontiguousarray(synth.pad_nodata(dem)[0:402, 0:502]), synth.NODATA)
    assert np.array_equal(fdr[1:399, 1:499], want[2:-2, 2:-2])


def test_fused_call_on_device_buffers():
    """ofl_flow_routing_f32 with OFL_MEM_DEVICE: the codes and counts of a device-resident DEM."""
    import torch

    from overflow_b200 import _native, device as dev

    rows, cols = 777, 1040
    dem = synth.punch_holes(synth.fractal(rows, cols, beta=2.0, seed=41), frac=0.01, seed=42)
    want_fdr, want_fac = _oracle(dem)
    d_dem = torch.from_numpy(dem).cuda()
    d_fdr = torch.empty((rows, cols), dtype=torch.uint8, device="cuda")
    d_fac = torch.empty((rows, cols), dtype=torch.int64, device="cuda")
    assert d_dem.stride(0) % 4 == 0 and d_fdr.stride(0) % 16 == 0
    _native.init(0)
    _native.check(_native.lib().ofl_flow_routing_f32(
        d_dem.data_ptr(), rows, cols, d_dem.stride(0), float(synth.NODATA), d_fdr.data_ptr(), d_fdr.stride(0),
        d_fac.data_ptr(), d_fac.stride(0), None, _native.OFL_MEM_DEVICE,
        torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert np.array_equal(d_fdr.cpu().numpy(), want_fdr)
    assert np.array_equal(d_fac.cpu().numpy(), want_fac)
    assert dev.check_accumulation(d_fdr, d_fac) == 0
