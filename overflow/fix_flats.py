from overflow_b200.fix_flats import (  # noqa: F401
    away_from_higher,
    d8_masked_flow_dirs,
    fix_flats_for_tile,
    flat_edges,
    label_flats,
    resolve_flats,
    towards_lower,
)
