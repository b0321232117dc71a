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


def test_library_is_sm100a_code_with_tma(lib):
    """The shipped library holds sm_100a machine code whose two tile-staging kernels load through TMA
    (UTMALDG) behind mbarriers (SYNCS), and nothing in it is a tensor-core contraction."""
    import shutil
    import subprocess

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([tool, "-lelf", _native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in elf and not re.search(r"sm_(?!100a)\d+", elf), elf
    sass = subprocess.run([tool, "-sass", _native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    per_kernel, cur = {}, None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per_kernel[cur] = [0, 0]
        elif cur:
            per_kernel[cur][0] += "UTMALDG" in line
            per_kernel[cur][1] += "SYNCS" in line
    for kernel in ("direction_kernel", "acc_tile_kernel"):
        hits = [v for k, v in per_kernel.items() if re.search(r"\d+%sE" % kernel, k)]
        assert hits and hits[0][0] >= 1 and hits[0][1] >= 2, (kernel, hits)
    assert not re.search(r"\b(UTC\w*MMA|HMMA|IMMA|QGMMA)\b", sass)
