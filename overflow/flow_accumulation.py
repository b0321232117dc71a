from overflow_b200.flow_accumulation import (  # noqa: F401
    flow_accumulation,
    follow_path,
    get_next_cell,
    perimeter_indices,
    single_tile_flow_accumulation,
)
