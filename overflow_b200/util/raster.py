"""Raster tiling and I/O on the host -- mirror of the reference's src/overflow/util/raster.py:1-210.

Same public names and behaviour: `raster_chunker`, `RasterChunk`, `read_raster_with_bounds_handling`,
`gdal_data_type_to_numpy_data_type`.  They work on anything with GDAL's band surface (XSize, YSize,
DataType, GetNoDataValue, ReadAsArray, WriteArray).  `open_raster` / `create_raster` use GDAL when it
is installed and the built-in GeoTIFF reader-writer (util/geotiff.py) when it is not.
"""
from typing import Iterator

import numpy as np

from . import geotiff as _geotiff

try:  # GDAL is the primary I/O path, exactly as in the reference (util/raster.py:2,9)
    from osgeo import gdal as _gdal

    _gdal.UseExceptions()
except ImportError:  # this image has no GDAL: built-in GeoTIFF I/O
    _gdal = None

try:
    from tqdm import tqdm as _tqdm
except ImportError:  # pragma: no cover
    def _tqdm(it, **_):
        return it

_NUMPY_OF_GDAL_NAME = {
    "Byte": np.uint8, "Int8": np.int8, "UInt16": np.uint16, "Int16": np.int16, "UInt32": np.uint32,
    "Int32": np.int32, "UInt64": np.uint64, "Int64": np.int64, "Float32": np.float32, "Float64": np.float64,
    "CInt16": np.complex64, "CInt32": np.complex128, "CFloat32": np.complex64, "CFloat64": np.complex128,
}


def gdal_data_type_to_numpy_data_type(gdal_dtype: int):
    """numpy dtype of a GDAL data type code (reference util/raster.py:12-35)."""
    name = _gdal.GetDataTypeName(gdal_dtype) if _gdal is not None else _geotiff.GDT_NAME[gdal_dtype]
    return _NUMPY_OF_GDAL_NAME[name]


def read_raster_with_bounds_handling(x_offset: int, y_offset: int, x_size: int, y_size: int, raster_band) -> np.ndarray:
    """Read a window that may hang over the raster; the overhang is the band's nodata value.

    Reference util/raster.py:38-103: shape (y_size, x_size), dtype of the band, nodata required.
    """
    assert x_size >= 0, "x_size must be positive"
    assert y_size >= 0, "y_size must be positive"
    no_data_value = raster_band.GetNoDataValue()
    assert no_data_value is not None, "The raster band has no no data value"
    window = np.full((y_size, x_size), no_data_value, dtype=gdal_data_type_to_numpy_data_type(raster_band.DataType))
    x_start, y_start = max(x_offset, 0), max(y_offset, 0)
    # part of the window that overlaps the raster
    win_xsize = min(x_size - (x_start - x_offset), max(raster_band.XSize - x_start, 0))
    win_ysize = min(y_size - (y_start - y_offset), max(raster_band.YSize - y_start, 0))
    if win_xsize > 0 and win_ysize > 0:
        window[y_start - y_offset : y_start - y_offset + win_ysize, x_start - x_offset : x_start - x_offset + win_xsize] = (
            raster_band.ReadAsArray(xoff=x_start, yoff=y_start, win_xsize=win_xsize, win_ysize=win_ysize)
        )
    return window


class RasterChunk:
    """A square chunk of a raster band plus an overlapping buffer ring (reference util/raster.py:106-171)."""

    def __init__(self, row: int, col: int, size: int, buffer_size: int):
        self.data = None
        self.row = row
        self.col = col
        self.size = size
        self.buffer_size = buffer_size

    def from_numpy(self, data: np.ndarray):
        self.data = data

    def read(self, band):
        """Read the chunk with its buffer ring; parts outside the raster are nodata."""
        side = self.size + 2 * self.buffer_size
        self.data = read_raster_with_bounds_handling(
            self.col * self.size - self.buffer_size, self.row * self.size - self.buffer_size, side, side, band
        )

    def _get_unbuffered_data(self, band) -> np.ndarray:
        rows = min(self.size, max(band.YSize - self.row * self.size, 0))
        cols = min(self.size, max(band.XSize - self.col * self.size, 0))
        b = self.buffer_size
        return self.data[b : b + rows, b : b + cols]

    def write(self, band):
        """Write the chunk without its buffer ring and without anything outside the raster."""
        if self.data is None:
            raise ValueError("The chunk has not been read yet.")
        band.WriteArray(self._get_unbuffered_data(band), xoff=self.col * self.size, yoff=self.row * self.size)


def raster_chunker(band, chunk_size: int, chunk_buffer_size: int) -> Iterator[RasterChunk]:
    """Yield the raster's chunks row-major, each already read (reference util/raster.py:174-210)."""
    n_chunks_row = (band.YSize + chunk_size - 1) // chunk_size
    n_chunks_col = (band.XSize + chunk_size - 1) // chunk_size
    for chunk_row in _tqdm(range(n_chunks_row), desc=f"Processing {n_chunks_row} raster chunks"):
        for chunk_col in range(n_chunks_col):
            chunk = RasterChunk(chunk_row, chunk_col, chunk_size, chunk_buffer_size)
            chunk.read(band)
            yield chunk


# ---------------------------------------------------------------- dataset helpers (GDAL or built-in GeoTIFF)
def open_raster(path):
    if _gdal is not None:
        return _gdal.Open(path)
    return _geotiff.open_geotiff(path)


def create_raster(path, xsize, ysize, dtype_name, projection=None, geotransform=None):
    """New 1-band GeoTIFF of GDAL type `dtype_name` ("Byte", "Int64", ...), like flow_direction.py:109-118."""
    if _gdal is not None:
        ds = _gdal.GetDriverByName("GTiff").Create(path, xsize, ysize, 1, _gdal.GetDataTypeByName(dtype_name))
    else:
        ds = _geotiff.create_geotiff(path, xsize, ysize, dtype_name)
    if projection:
        ds.SetProjection(projection)
    if geotransform is not None:
        ds.SetGeoTransform(geotransform)
    return ds


def read_band(band, chunk_size):
    """Whole band as one array, read in row bands of chunk_size rows."""
    out = np.empty((band.YSize, band.XSize), dtype=gdal_data_type_to_numpy_data_type(band.DataType))
    for y in range(0, band.YSize, max(1, int(chunk_size))):
        n = min(int(chunk_size), band.YSize - y)
        out[y : y + n] = band.ReadAsArray(xoff=0, yoff=y, win_xsize=band.XSize, win_ysize=n)
    return out


def write_band(band, array, chunk_size):
    for y in range(0, array.shape[0], max(1, int(chunk_size))):
        band.WriteArray(array[y : y + int(chunk_size)], xoff=0, yoff=y)
