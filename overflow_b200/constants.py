"""Direction codes, scan order and sentinels of the D8 path.

Same names and values as the reference's src/overflow/constants.py:4-59 (they are API):
    3 | 2 | 1
    4 | 8 | 0        8 = undefined (pit / flat), 9 = nodata
    5 | 6 | 7
"""
import numpy as np

DEFAULT_SEARCH_RADIUS = 200
DEFAULT_MAX_PITS = 24
UNVISITED_INDEX = -1
EPSILON_GRADIENT = 1e-5
DEFAULT_CHUNK_SIZE = 2000

(
    FLOW_DIRECTION_EAST,
    FLOW_DIRECTION_NORTH_EAST,
    FLOW_DIRECTION_NORTH,
    FLOW_DIRECTION_NORTH_WEST,
    FLOW_DIRECTION_WEST,
    FLOW_DIRECTION_SOUTH_WEST,
    FLOW_DIRECTION_SOUTH,
    FLOW_DIRECTION_SOUTH_EAST,
    FLOW_DIRECTION_UNDEFINED,
    FLOW_DIRECTION_NODATA,
) = range(10)

# (d_row, d_col) of the downstream neighbour, indexed by direction code 0..7 (the scan order
# of the steepest-descent search); int64 like the reference's default-dtype table
NEIGHBOR_OFFSETS = np.array(
    [(0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1), (1, 0), (1, 1)], dtype=np.int64
)
# identity map code -> code, kept because callers index it
FLOW_DIRECTIONS = np.arange(10, dtype=np.uint8)

FLOW_ACCUMULATION_NODATA = -9999
# what NODATA cells actually hold after single_tile_flow_accumulation: the reference
# initialises them to -9999 and then dequeues them once (+1); see SURVEY.md fact 2
FLOW_ACCUMULATION_NODATA_EMITTED = -9998
FLOW_TERMINATES = (-1, -1)
FLOW_EXTERNAL = (-2, -2)
