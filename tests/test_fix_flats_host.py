"""Host-side behaviour of the fix_flats mirror that needs no device: argument checks happen before the library
is touched, so a wrong call fails the same way with or without a GPU."""
import numpy as np
import pytest

from overflow import fix_flats as ff
from overflow_b200 import fix_flats as impl


def test_reference_names_are_exported():
    # tests/test_fix_flats.py:3-10 of the reference imports exactly these
    for name in ("flat_edges", "label_flats", "away_from_higher", "towards_lower", "resolve_flats", "d8_masked_flow_dirs"):
        assert callable(getattr(ff, name))


def test_dem_dtype_rule():
    assert impl._dem_f32(np.array([[1, 2], [3, 4]], dtype=np.int16)).dtype == np.float32
    assert impl._dem_f32(np.array([[0.5, np.nan]], dtype=np.float64)).dtype == np.float32  # exact, NaN stays NaN
    with pytest.raises(TypeError):
        impl._dem_f32(np.array([[2**24 + 1]], dtype=np.int32))  # float32 cannot hold it: == and < would change
    with pytest.raises(TypeError):
        impl._dem_f32(np.array([[0.1]], dtype=np.float64))
    with pytest.raises(ValueError):
        impl._dem_f32(np.zeros(4, dtype=np.float32))


def test_code_and_mask_checks():
    dem = np.zeros((3, 4), dtype=np.float32)
    with pytest.raises(TypeError):
        ff.resolve_flats(dem, np.zeros((3, 4), dtype=np.float32))
    with pytest.raises(ValueError):
        ff.resolve_flats(dem, np.zeros((3, 5), dtype=np.uint8))
    with pytest.raises(ValueError):
        ff.flat_edges(dem, np.full((3, 4), 300, dtype=np.int32))
    with pytest.raises(TypeError):
        ff.d8_masked_flow_dirs(np.zeros((3, 4), np.int32), np.zeros((3, 4), np.int64), np.zeros((3, 4), np.int32))
    with pytest.raises(ValueError):
        ff.d8_masked_flow_dirs(np.zeros((3, 4), np.int64) + 2**40, np.zeros((3, 4), np.uint8), np.zeros((3, 4), np.int32))
    with pytest.raises(ValueError):
        ff.d8_masked_flow_dirs(np.zeros((2, 4), np.int32), np.zeros((3, 4), np.uint8), np.zeros((3, 4), np.int32))


def test_sweep_seed_checks():
    labels = np.ones((4, 4), dtype=np.int32)
    fdr = np.full((4, 4), 8, dtype=np.uint8)
    with pytest.raises(IndexError):
        ff.away_from_higher(labels, np.zeros((4, 4), np.int32), fdr, [(9, 0)], np.zeros(1, np.int32))
    with pytest.raises(IndexError):  # label 1 needs flat_height[0]
        ff.towards_lower(labels * 3, np.zeros((4, 4), np.int32), fdr, [(0, 0)], np.zeros(1, np.int32))
    with pytest.raises(TypeError):
        ff.away_from_higher(labels, np.zeros((4, 4), np.float32), fdr, [(0, 0)], np.zeros(1, np.int32))


def test_label_flats_degenerate_starts():
    dem = np.array([[1.0, np.nan], [2.0, 3.0]], dtype=np.float32)
    labels = np.zeros((2, 2), dtype=np.uint32)
    ff.label_flats(dem, labels, 7, 5, 5)  # the reference pops the out-of-bounds start and stops
    ff.label_flats(dem, labels, 7, 0, 1)  # NaN equals nothing, itself included
    assert not labels.any()
    ff.label_flats(dem, labels, 7, 1, 1)  # a flat of one cell: no equal neighbour, no device work needed
    assert labels.tolist() == [[0, 0], [0, 7]]
    ff.label_flats(dem, labels, 9, 1, 1)  # already labelled: the reference pops it and stops
    assert labels[1, 1] == 7


def test_pit_breaching_argument_checks():
    """breach_single_cell_pits_in_chunk rejects what the device path does not implement before touching the library."""
    from overflow.breach_single_cell_pits import breach_single_cell_pits, breach_single_cell_pits_in_chunk

    assert callable(breach_single_cell_pits)
    with pytest.raises(TypeError):
        breach_single_cell_pits_in_chunk(np.zeros((8, 8), dtype=np.float64), -9999.0)
    with pytest.raises(ValueError):
        breach_single_cell_pits_in_chunk(np.zeros(8, dtype=np.float32), -9999.0)
    with pytest.raises(ValueError):
        breach_single_cell_pits_in_chunk([[1.0, 2.0]], -9999.0)
    empty = np.zeros((0, 7), dtype=np.float32)  # nothing to do: no library call, an empty result
    assert breach_single_cell_pits_in_chunk(empty, -9999.0).shape == (0, 7)


def test_cli_exposes_the_reference_commands():
    import overflow_cli

    names = set(overflow_cli.main.commands)
    assert {"breach-single-cell-pits", "flow-direction"} <= names  # reference overflow_cli.py:18, :54
    opts = {p.name for p in overflow_cli.main.commands["breach-single-cell-pits"].params}
    assert opts == {"input_file", "output_file", "chunk_size"}
