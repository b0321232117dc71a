from overflow_b200.flow_direction import flow_direction, flow_direction_for_tile, flow_direction_for_raster  # noqa: F401
