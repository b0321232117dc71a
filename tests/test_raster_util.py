"""Raster tiling + built-in GeoTIFF I/O (CPU).  Follows the reference's tests/test_raster_util.py:
bands of 100x100 / 100x200 / 200x100, bounds handling incl. negative offsets and empty windows,
chunker with even and odd chunk sizes and a 2-cell buffer, write-back round trip."""
import numpy as np
import pytest

from overflow_b200.util import geotiff
from overflow_b200.util.raster import (
    RasterChunk,
    create_raster,
    open_raster,
    raster_chunker,
    read_raster_with_bounds_handling,
)

SHAPES = {"square": (100, 100), "tall": (200, 100), "wide": (100, 200)}


@pytest.fixture(params=list(SHAPES), scope="module")
def raster_band(request, tmp_path_factory):
    rows, cols = SHAPES[request.param]
    path = str(tmp_path_factory.mktemp("rast") / f"{request.param}.tif")
    ds = create_raster(path, cols, rows, "Float32", geotransform=(500000.0, 10.0, 0.0, 4400000.0, 0.0, -10.0))
    band = ds.GetRasterBand(1)
    band.WriteArray(np.random.default_rng(1).random((rows, cols)).astype(np.float32))
    band.SetNoDataValue(-9999)
    ds.FlushCache()
    ds2 = open_raster(path)
    return ds2.GetRasterBand(1)


def test_read_in_bounds(raster_band):
    a = read_raster_with_bounds_handling(10, 10, 20, 20, raster_band)
    assert a.shape == (20, 20) and a.dtype == np.float32
    assert np.all(a == raster_band.ReadAsArray(10, 10, 20, 20))


def test_read_out_of_bounds(raster_band):
    nd = raster_band.GetNoDataValue()
    a = read_raster_with_bounds_handling(-10, -10, 20, 20, raster_band)
    assert a.shape == (20, 20)
    assert np.all(a[:10, :] == nd) and np.all(a[:, :10] == nd)
    assert np.all(a[10:, 10:] == raster_band.ReadAsArray(0, 0, 10, 10))
    xs, ys = raster_band.XSize, raster_band.YSize
    b = read_raster_with_bounds_handling(xs - 10, ys - 10, 20, 20, raster_band)
    assert np.all(b[10:, :] == nd) and np.all(b[:, 10:] == nd)
    assert np.all(b[:10, :10] == raster_band.ReadAsArray(xs - 10, ys - 10, 10, 10))
    c = read_raster_with_bounds_handling(xs + 5, ys + 5, 7, 9, raster_band)
    assert c.shape == (9, 7) and np.all(c == nd)


def test_read_zero_and_negative_size(raster_band):
    assert read_raster_with_bounds_handling(0, 0, 0, 0, raster_band).shape == (0, 0)
    with pytest.raises(AssertionError):
        read_raster_with_bounds_handling(0, 0, -1, 5, raster_band)
    with pytest.raises(AssertionError):
        read_raster_with_bounds_handling(0, 0, 5, -1, raster_band)


@pytest.mark.parametrize("chunk_size", [10, 11])
def test_chunker_halo_and_roundtrip(raster_band, chunk_size, tmp_path):
    nd, buf = raster_band.GetNoDataValue(), 2
    whole = raster_band.ReadAsArray()
    out_ds = create_raster(str(tmp_path / "out.tif"), raster_band.XSize, raster_band.YSize, "Float32")
    out_band = out_ds.GetRasterBand(1)
    n = 0
    for chunk in raster_chunker(raster_band, chunk_size, buf):
        n += 1
        assert chunk.data.shape == (chunk_size + 2 * buf, chunk_size + 2 * buf)
        y0, x0 = chunk.row * chunk_size - buf, chunk.col * chunk_size - buf
        for (yy, xx), v in np.ndenumerate(chunk.data[:3, :3]):
            gy, gx = y0 + yy, x0 + xx
            inside = 0 <= gy < whole.shape[0] and 0 <= gx < whole.shape[1]
            assert v == (whole[gy, gx] if inside else nd)
        chunk.write(out_band)
    assert n == -(-whole.shape[0] // chunk_size) * -(-whole.shape[1] // chunk_size)
    out_ds.FlushCache()
    assert np.array_equal(out_band.ReadAsArray(), whole)


def test_unread_chunk_write_raises(raster_band):
    with pytest.raises(ValueError):
        RasterChunk(0, 0, 10, 1).write(raster_band)


@pytest.mark.parametrize("dtype_name,np_dtype", [("Byte", np.uint8), ("Int64", np.int64), ("Float32", np.float32)])
def test_geotiff_roundtrip_and_georeferencing(tmp_path, dtype_name, np_dtype):
    path = str(tmp_path / f"{dtype_name}.tif")
    gt = (1000.0, 2.5, 0.0, 9000.0, 0.0, -2.5)
    ds = create_raster(path, 37, 23, dtype_name, geotransform=gt)
    data = (np.arange(23 * 37).reshape(23, 37) % 200).astype(np_dtype)
    ds.GetRasterBand(1).WriteArray(data[:10], yoff=0)
    ds.GetRasterBand(1).WriteArray(data[10:], yoff=10)
    ds.GetRasterBand(1).SetNoDataValue(9)
    ds.FlushCache()
    rd = open_raster(path)
    assert (rd.RasterXSize, rd.RasterYSize) == (37, 23)
    band = rd.GetRasterBand(1)
    assert band.DataType == geotiff.GDT[dtype_name] and band.GetNoDataValue() == 9
    assert np.array_equal(band.ReadAsArray(), data)
    assert np.array_equal(band.ReadAsArray(xoff=5, yoff=7, win_xsize=11, win_ysize=4), data[7:11, 5:16])
    assert rd.GetGeoTransform() == pytest.approx(gt)
    # the file is a valid classic little-endian TIFF
    assert open(path, "rb").read(4) == b"II*\x00"


def test_missing_nodata_asserts(tmp_path):
    ds = create_raster(str(tmp_path / "n.tif"), 8, 8, "Float32")
    ds.FlushCache()
    band = open_raster(str(tmp_path / "n.tif")).GetRasterBand(1)
    with pytest.raises(AssertionError):
        read_raster_with_bounds_handling(0, 0, 4, 4, band)


def _write_striped_tiff(path, arr, rows_per_strip, gap=16, bits=None):
    """A classic little-endian TIFF with one strip per `rows_per_strip` rows, the strips NOT adjacent in the file
    (a gap after each) and the last strip short when rows % rows_per_strip != 0 -- what libtiff writers produce."""
    import struct

    rows, cols = arr.shape
    item = arr.dtype.itemsize
    fmt = {"u": 1, "i": 2, "f": 3}[arr.dtype.kind]
    strips = [arr[r : r + rows_per_strip].tobytes() for r in range(0, rows, rows_per_strip)]
    offsets, pos = [], 8
    body = b""
    for s in strips:
        offsets.append(pos)
        body += s + b"\xAA" * gap
        pos += len(s) + gap
    n = len(strips)
    off_offsets, off_counts = pos, pos + 4 * n
    extra = struct.pack(f"<{n}I", *offsets) + struct.pack(f"<{n}I", *[len(s) for s in strips])
    ifd_pos = off_counts + 4 * n
    tags = [(256, 4, 1, cols), (257, 4, 1, rows), (258, 3, 1, bits or item * 8), (259, 3, 1, 1), (262, 3, 1, 1),
            (273, 4, n, off_offsets if n > 1 else offsets[0]), (277, 3, 1, 1), (278, 4, 1, rows_per_strip),
            (279, 4, n, off_counts if n > 1 else len(strips[0])), (339, 3, 1, fmt)]
    ifd = struct.pack("<H", len(tags))
    for tag, ft, cnt, val in tags:
        ifd += struct.pack("<HHI", tag, ft, cnt) + (struct.pack("<HH", val, 0) if ft == 3 else struct.pack("<I", val))
    ifd += struct.pack("<I", 0)
    with open(path, "wb") as f:
        f.write(b"II" + struct.pack("<HI", 42, ifd_pos) + body + extra + ifd)


@pytest.mark.parametrize("rows,rps", [(37, 8), (40, 8), (5, 16), (33, 1)])
def test_striped_tiff_with_short_last_strip(tmp_path, rows, rps):
    """strips that are not adjacent in the file, the last one shorter than RowsPerStrip."""
    from overflow_b200.util.geotiff import open_geotiff

    arr = (np.arange(rows * 23, dtype=np.float32).reshape(rows, 23) * 0.5).astype(np.float32)
    path = str(tmp_path / "striped.tif")
    _write_striped_tiff(path, arr, rps)
    band = open_geotiff(path).GetRasterBand(1)
    assert (band.YSize, band.XSize) == (rows, 23)
    assert np.array_equal(band.ReadAsArray(), arr)
    assert np.array_equal(band.ReadAsArray(xoff=3, yoff=rows - 4, win_xsize=7, win_ysize=4), arr[rows - 4 :, 3:10])


def test_unsupported_bit_depth_is_refused(tmp_path):
    from overflow_b200.util.geotiff import open_geotiff

    path = str(tmp_path / "onebit.tif")
    _write_striped_tiff(path, np.zeros((8, 8), dtype=np.uint8), 8, bits=1)
    with pytest.raises(ValueError, match="bits per sample"):
        open_geotiff(path)
