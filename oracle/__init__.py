"""CPU parity oracle for the D8 hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``overflow_b200/``
imports it; the product path fails loudly when the CUDA library is missing.

Parity status: pinned (see ``oracle/d8_oracle.c`` header and ``oracle/gen_golden.py``; flat resolution:
``oracle/flats_oracle.c`` and ``oracle/gen_golden_flats.py``).
"""
from .oracle import (  # noqa: F401
    build,
    flow_direction_for_tile,
    flow_accumulation,
    links_perimeter,
    single_tile_flow_accumulation,
    check_accumulation,
    num_threads,
    set_num_threads,
    flat_edges,
    resolve_flats,
    d8_masked_flow_dirs,
    breach_single_cell_pits_in_chunk,
    synth_dem,
)
