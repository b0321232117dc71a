from overflow_b200.util.raster import (  # noqa: F401
    RasterChunk,
    gdal_data_type_to_numpy_data_type,
    raster_chunker,
    read_raster_with_bounds_handling,
)
