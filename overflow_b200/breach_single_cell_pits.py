"""Single-cell pit breaching -- host-side mirror of the reference's src/overflow/breach_single_cell_pits.py.

`breach_single_cell_pits_in_chunk` (reference :9-63) keeps its contract: the float32 chunk is breached IN PLACE
and the int8 raster of pits that could not be breached is returned; `breach_single_cell_pits` (:66-94) is the
file driver with the reference's chunking (buffer of two cells, every chunk read from the input file, so a
chunk never sees its neighbours' breaches -- results depend on chunk_size exactly as the reference's do).
The compute is csrc/pits.cu behind ofl_breach_single_cell_pits_f32; there is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _native
from .constants import DEFAULT_CHUNK_SIZE


def breach_single_cell_pits_in_chunk(chunk: np.ndarray, nodata_value: float, return_info: bool = False) -> np.ndarray:
    """Breach the single-cell pits of a DEM chunk in place; returns the int8 raster of unsolved pits.

    The reference's arithmetic follows the chunk's dtype; the device path implements the float32 case (what
    `breach_single_cell_pits` writes, GDT_Float32) and raises TypeError for anything else.
    """
    if not isinstance(chunk, np.ndarray) or chunk.ndim != 2:
        raise ValueError("chunk must be a 2-D numpy array (it is modified in place)")
    if chunk.dtype != np.float32:
        raise TypeError(f"breach_single_cell_pits_in_chunk works on float32 chunks, got {chunk.dtype}")
    rows, cols = chunk.shape
    if rows * cols > 2**32:
        raise ValueError("pit breaching works on one chunk of at most 2**32 cells")
    unsolved = np.zeros((rows, cols), dtype=np.int8)
    info = (ctypes.c_int64 * 3)()
    if rows and cols:
        work = chunk if chunk.flags.c_contiguous else np.ascontiguousarray(chunk)
        _native.check(
            _native.lib().ofl_breach_single_cell_pits_f32(
                work.ctypes.data, rows, cols, cols, float(nodata_value), unsolved.ctypes.data, info, None, 0,
                _native.OFL_MEM_HOST, None,
            )
        )
        if work is not chunk:
            chunk[...] = work
    if return_info:
        return unsolved, {"pits": int(info[0]), "unsolved": int(info[1]), "rounds": int(info[2])}
    return unsolved


def breach_single_cell_pits(input_path: str, output_path: str, chunk_size: int = DEFAULT_CHUNK_SIZE):
    """DEM file in, DEM file with breached single-cell pits out (reference :66-94): band 1, 1-band Float32
    GeoTIFF with the input's projection, geotransform and nodata value."""
    from .util import raster as _raster

    src = _raster.open_raster(input_path)
    band = src.GetRasterBand(1)
    nodata_value = band.GetNoDataValue()
    dst = _raster.create_raster(
        output_path, src.RasterXSize, src.RasterYSize, "Float32",
        projection=src.GetProjection(), geotransform=src.GetGeoTransform(),
    )
    out_band = dst.GetRasterBand(1)
    out_band.SetNoDataValue(nodata_value)
    for chunk in _raster.raster_chunker(band, chunk_size=chunk_size, chunk_buffer_size=2):
        if chunk.data.dtype != np.float32:
            raise TypeError(f"breach_single_cell_pits works on Float32 DEMs, band 1 is {chunk.data.dtype}")
        data = np.ascontiguousarray(chunk.data)
        breach_single_cell_pits_in_chunk(data, nodata_value)
        chunk.from_numpy(data)
        chunk.write(out_band)
    dst.FlushCache()
    dst = None
