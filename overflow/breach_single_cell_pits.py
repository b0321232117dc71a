from overflow_b200.breach_single_cell_pits import (  # noqa: F401
    breach_single_cell_pits,
    breach_single_cell_pits_in_chunk,
)
