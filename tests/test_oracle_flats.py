"""The C restatement of the reference's flat resolution (oracle/flats_oracle.c) against fixtures produced
by the reference itself (oracle/gen_golden_flats.py -> tests/golden/fix_flats.npz), including the
known-answer vectors of the reference's tests/test_fix_flats.py."""
import os

import numpy as np
import pytest

import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fix_flats.npz")


def load_cases():
    z = np.load(GOLD)
    names = sorted({k.split("__")[0] for k in z.files})
    return z, names


Z, NAMES = load_cases()


def edges_of(shape, high, low):
    e = np.zeros(shape, dtype=np.uint8)
    for r, c in low:
        e[r, c] |= 1
    for r, c in high:
        e[r, c] |= 2
    return e


def test_fixture_inventory():
    assert "kat" in NAMES and len(NAMES) >= 20


@pytest.mark.parametrize("name", NAMES)
def test_flat_edges(name):
    dem, fdr = Z[f"{name}__dem"], Z[f"{name}__fdr"]
    high, low = oracle.flat_edges(dem, fdr)
    assert np.array_equal(edges_of(dem.shape, high, low), Z[f"{name}__edges"])
    assert low == sorted(low) and high == sorted(high)  # row-major, the order the reference appends in


@pytest.mark.parametrize("name", NAMES)
def test_resolve_flats(name):
    flat_mask, labels = oracle.resolve_flats(Z[f"{name}__dem"], Z[f"{name}__fdr"])
    assert np.array_equal(labels, Z[f"{name}__labels"])
    assert np.array_equal(flat_mask, Z[f"{name}__flat_mask"])


@pytest.mark.parametrize("name", NAMES)
def test_d8_masked_flow_dirs(name):
    got = oracle.d8_masked_flow_dirs(Z[f"{name}__flat_mask"], Z[f"{name}__fdr"], Z[f"{name}__labels"])
    assert np.array_equal(got, Z[f"{name}__fdr_fixed"])


def test_reference_known_answers():
    """tests/test_fix_flats.py:137-143 (edges), :175-192 (final mask) of the reference."""
    dem, fdr = Z["kat__dem"], Z["kat__fdr"]
    high, low = oracle.flat_edges(dem, fdr)
    assert low == [(5, 1), (5, 2), (5, 3)]
    assert sorted(high) == sorted([(1, 1), (1, 2), (1, 3), (1, 4), (1, 5), (2, 1), (3, 1), (4, 1), (2, 5), (3, 5),
                                   (4, 5), (5, 5), (5, 4)])
    flat_mask, labels = oracle.resolve_flats(dem, fdr)
    assert flat_mask[1, 1] == 12 and flat_mask[3, 3] == 6 and flat_mask[5, 1] == 2
    assert labels[1:6, 1:6].min() == 1 and labels.sum() == 25
