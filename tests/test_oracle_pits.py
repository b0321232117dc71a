"""The C restatement of the reference's single-cell pit breaching (oracle/pits_oracle.c) against fixtures produced by
the reference itself (oracle/gen_golden_pits.py -> tests/golden/breach_pits.npz), bit for bit."""
import os

import numpy as np
import pytest

import oracle

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "breach_pits.npz"))
NAMES = sorted({k.split("__")[0] for k in Z.files})


def same_bits(a, b):
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("name", NAMES)
def test_breach_single_cell_pits_in_chunk(name):
    out, unsolved = oracle.breach_single_cell_pits_in_chunk(Z[f"{name}__chunk_in"], float(Z[f"{name}__nodata"]))
    assert np.array_equal(unsolved, Z[f"{name}__unsolved"])
    assert same_bits(out, Z[f"{name}__chunk_out"])


def test_reference_known_answer():
    """tests/test_breach_single_cell_pits.py:62-83: the pit at (4, 4) is breached towards (3, 2) through (4, 3)."""
    out, unsolved = oracle.breach_single_cell_pits_in_chunk(Z["kat__chunk_in"], -999)
    assert out[4, 3] == -0.5 and not unsolved.any()
    assert len(NAMES) >= 15
