"""CPU-side checks of the C-ABI library: it loads, exports every declared symbol and refuses to
compute without a CUDA device (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from overflow_b200 import _native, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _native.lib()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "overflow_b200.h")).read()
    declared = set(re.findall(r"\b(ofl_[a-z0-9_]+)\s*\(", header))
    declared -= {"ofl_status", "ofl_mem_kind", "ofl_dir_mode"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_native.SIGNATURES), "ctypes signature table out of sync with the header"


def test_abi_version(lib):
    assert lib.ofl_abi_version() == 2


def test_perimeter_count(lib):
    assert lib.ofl_perimeter_count(7, 7) == 24
    assert lib.ofl_perimeter_count(1, 40) == 2 + 2 * 38
    assert lib.ofl_perimeter_count(40, 1) == 80
    assert lib.ofl_perimeter_count(0, 5) == 0


def test_workspace_size_monotone(lib):
    a = lib.ofl_accumulation_workspace_bytes(1024, 1024)
    b = lib.ofl_accumulation_workspace_bytes(4096, 4096)
    assert 0 < a < b


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from overflow_b200.flow_direction import flow_direction_for_tile

    with pytest.raises(_native.OverflowB200Error):
        flow_direction_for_tile(np.zeros((5, 5), dtype=np.float32), -9999.0)


def test_dem_dtype_dispatch():
    """Host logic only: which kernel family a DEM dtype takes and how it is widened (no compute call)."""
    from overflow_b200 import _native
    from overflow_b200.flow_direction import _classify

    for dt in (np.float32, np.int8, np.int16):
        kind, arr = _classify(np.zeros((3, 3), dtype=dt))
        assert kind is None and arr.dtype == np.float32
    kind, arr = _classify(np.zeros((3, 3), dtype=np.float64))
    assert kind == _native.OFL_ELEM_F64 and arr.dtype == np.float64
    for dt in (np.int32, np.int64):
        kind, arr = _classify(np.full((3, 3), -7, dtype=dt))
        assert kind == _native.OFL_ELEM_I64 and arr.dtype == np.int64 and arr[0, 0] == -7
    for dt in (np.uint8, np.uint16, np.uint32, np.uint64):
        kind, arr = _classify(np.full((3, 3), 200, dtype=dt))
        assert kind == _native.OFL_ELEM_U64 and arr.dtype == np.uint64 and arr[0, 0] == 200
    for dt in (np.float16, np.complex64, np.bool_):
        with pytest.raises(TypeError):
            _classify(np.zeros((3, 3), dtype=dt))
