"""Drop-in `overflow` package: the reference's module names, backed by overflow_b200 for the D8 path.

`from overflow.flow_direction import flow_direction, flow_direction_for_tile`,
`from overflow.flow_accumulation import single_tile_flow_accumulation` and
`from overflow.constants import ...` work exactly as with the reference (its tests import these names,
tests/test_flow_direction.py:5-17, tests/test_flow_accumulation.py:3-12).
"""
