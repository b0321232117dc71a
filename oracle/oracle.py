"""ctypes front-end of ``oracle/liboracle_d8.so`` (the C restatement in d8_oracle.c).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  Function names follow the
reference (``/root/reference/src/overflow/flow_direction.py``,
``flow_accumulation.py``) so that parity tests read like the reference's own tests.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_d8.so")
_lib = None

FLOW_EXTERNAL = (-2, -2)
FLOW_TERMINATES = (-1, -1)


def build(force=False):
    """Compile d8_oracle.c with the committed Makefile (gcc, OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ("d8_oracle.c", "flats_oracle.c", "pits_oracle.c", "synth_host.c", "Makefile")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def _load():
    global _lib
    if _lib is not None:
        return _lib
    build()
    lib = ctypes.CDLL(_SO)
    i64, f64, vp = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    for name in ("orc_flow_direction_f32", "orc_flow_direction_f64", "orc_flow_direction_i64", "orc_flow_direction_u64"):
        fn = getattr(lib, name)
        fn.argtypes = [vp, i64, i64, i64, f64, vp, i64, ctypes.c_int]
        fn.restype = None
    lib.orc_flow_accumulation.argtypes = [vp, i64, i64, i64, vp, i64]
    lib.orc_flow_accumulation.restype = ctypes.c_int
    lib.orc_perimeter_count.argtypes = [i64, i64]
    lib.orc_perimeter_count.restype = i64
    lib.orc_links_perimeter.argtypes = [vp, i64, i64, i64, vp, vp]
    lib.orc_links_perimeter.restype = i64
    lib.orc_check_accumulation.argtypes = [vp, i64, i64, i64, vp, i64]
    lib.orc_check_accumulation.restype = i64
    lib.orc_num_threads.restype = ctypes.c_int
    lib.orc_breach_single_cell_pits_f32.argtypes = [vp, i64, i64, f64, vp]
    lib.orc_breach_single_cell_pits_f32.restype = i64
    lib.orc_flat_edges_f32.argtypes = [vp, vp, i64, i64, vp, ctypes.POINTER(i64)]
    lib.orc_flat_edges_f32.restype = i64
    lib.orc_resolve_flats_f32.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.orc_resolve_flats_f32.restype = i64
    lib.orc_d8_masked_flow_dirs.argtypes = [vp, vp, vp, i64, i64]
    lib.orc_d8_masked_flow_dirs.restype = None
    lib.orc_set_num_threads.argtypes = [ctypes.c_int]
    lib.orc_synth_dem_f32.argtypes = [vp, i64, i64, i64, i64, i64, i64, i64, ctypes.c_uint64, ctypes.c_int,
                                      ctypes.c_float, ctypes.c_int, ctypes.c_float]
    lib.orc_synth_dem_f32.restype = None
    _lib = lib
    return lib


def num_threads():
    return int(_load().orc_num_threads())


def set_num_threads(n):
    _load().orc_set_num_threads(int(n))


def flow_direction_for_tile(dem, nodata_value, border=9):
    """Restates flow_direction_for_tile (reference flow_direction.py:14-69).

    Interior cells get the D8 code; the border ring (uninitialised in the reference)
    is set to ``border``.
    """
    dem = np.asarray(dem)
    if dem.ndim != 2:
        raise ValueError("dem must be 2-D")
    if dem.dtype == np.float32:
        fn = _load().orc_flow_direction_f32
    elif dem.dtype == np.float64:
        fn = _load().orc_flow_direction_f64
    elif np.issubdtype(dem.dtype, np.signedinteger):
        fn = _load().orc_flow_direction_i64
        dem = dem.astype(np.int64)
    elif np.issubdtype(dem.dtype, np.unsignedinteger):
        fn = _load().orc_flow_direction_u64
        dem = dem.astype(np.uint64)
    else:
        raise TypeError(f"oracle supports float32/float64/integer DEMs, got {dem.dtype}")
    dem = np.ascontiguousarray(dem)
    rows, cols = dem.shape
    out = np.empty((rows, cols), dtype=np.uint8)
    fn(dem.ctypes.data, rows, cols, cols, float(nodata_value), out.ctypes.data, cols, int(border))
    return out


def _as_codes(fdr):
    fdr = np.asarray(fdr)
    if fdr.ndim != 2:
        raise ValueError("flow_direction must be 2-D")
    if fdr.dtype != np.uint8:
        if fdr.size and (fdr.min() < 0 or fdr.max() > 255):
            raise ValueError("flow direction codes must be in 0..255")
        fdr = fdr.astype(np.uint8)
    return np.ascontiguousarray(fdr)


def flow_accumulation(fdr):
    """Accumulation part of single_tile_flow_accumulation (flow_accumulation.py:95-144)."""
    fdr = _as_codes(fdr)
    rows, cols = fdr.shape
    fac = np.empty((rows, cols), dtype=np.int64)
    rc = _load().orc_flow_accumulation(fdr.ctypes.data, rows, cols, cols, fac.ctypes.data, cols)
    if rc != 0:
        raise MemoryError("oracle queue allocation failed")
    return fac


def links_perimeter(fdr):
    """(perim_rc[n,2], perim_links[n,2]) in perimeter_indices order (flow_accumulation.py:40-92)."""
    fdr = _as_codes(fdr)
    rows, cols = fdr.shape
    n = int(_load().orc_perimeter_count(rows, cols))
    links = np.empty((n, 2), dtype=np.int64)
    rc = np.empty((n, 2), dtype=np.int64)
    k = _load().orc_links_perimeter(fdr.ctypes.data, rows, cols, cols, links.ctypes.data, rc.ctypes.data)
    if k != n:
        raise RuntimeError("runaway flow path (cyclic flow direction raster)")
    return rc, links


def single_tile_flow_accumulation(fdr):
    """Restates single_tile_flow_accumulation (flow_accumulation.py:95-158).

    Returns (fac int64[R,C], links int64[R,C,2]); interior link entries (uninitialised
    in the reference) are zero here.
    """
    fdr = _as_codes(fdr)
    fac = flow_accumulation(fdr)
    rc, pl = links_perimeter(fdr)
    links = np.zeros(fdr.shape + (2,), dtype=np.int64)
    if len(rc):
        links[rc[:, 0], rc[:, 1]] = pl
    return fac, links


def check_accumulation(fdr, fac):
    """Number of cells violating fac[c] = 1 + sum(upstream fac) / -9998."""
    fdr = _as_codes(fdr)
    fac = np.ascontiguousarray(fac, dtype=np.int64)
    rows, cols = fdr.shape
    return int(_load().orc_check_accumulation(fdr.ctypes.data, rows, cols, cols, fac.ctypes.data, cols))


# ---------------------------------------------------------------- flat resolution (flats_oracle.c)
def _dem_f32(dem):
    dem = np.asarray(dem)
    if dem.ndim != 2:
        raise ValueError("dem must be 2-D")
    return np.ascontiguousarray(dem, dtype=np.float32)


def flat_edges(dem, fdr):
    """Restates flat_edges (reference fix_flats.py:13-62): (high_edges, low_edges) as lists of
    (row, col) in the reference's row-major order."""
    dem, fdr = _dem_f32(dem), _as_codes(fdr)
    rows, cols = fdr.shape
    edges = np.zeros((rows, cols), dtype=np.uint8)
    n_high = ctypes.c_int64(0)
    _load().orc_flat_edges_f32(dem.ctypes.data, fdr.ctypes.data, rows, cols, edges.ctypes.data, ctypes.byref(n_high))
    high = [(int(r), int(c)) for r, c in zip(*np.nonzero(edges & 2))]
    low = [(int(r), int(c)) for r, c in zip(*np.nonzero(edges & 1))]
    return high, low


def resolve_flats(dem, fdr):
    """Restates resolve_flats (fix_flats.py:227-288): (flat_mask int32, labels int32)."""
    dem, fdr = _dem_f32(dem), _as_codes(fdr)
    rows, cols = fdr.shape
    flat_mask = np.zeros((rows, cols), dtype=np.int32)
    labels = np.zeros((rows, cols), dtype=np.int32)
    rc = _load().orc_resolve_flats_f32(dem.ctypes.data, fdr.ctypes.data, rows, cols, flat_mask.ctypes.data,
                                       labels.ctypes.data)
    if rc < 0:
        raise MemoryError("oracle queue allocation failed")
    return flat_mask, labels


def d8_masked_flow_dirs(flat_mask, fdr, labels):
    """Restates d8_masked_flow_dirs (fix_flats.py:291-339); returns the rewritten codes (copy)."""
    out = _as_codes(fdr).copy()
    fm = np.ascontiguousarray(flat_mask, dtype=np.int32)
    lb = np.ascontiguousarray(labels, dtype=np.int32)
    rows, cols = out.shape
    _load().orc_d8_masked_flow_dirs(fm.ctypes.data, out.ctypes.data, lb.ctypes.data, rows, cols)
    return out


# ---------------------------------------------------------------- single-cell pit breaching (pits_oracle.c)
def breach_single_cell_pits_in_chunk(chunk, nodata_value):
    """Restates breach_single_cell_pits_in_chunk (reference breach_single_cell_pits.py:9-63) on a COPY:
    returns (breached chunk float32, unsolved int8)."""
    out = np.array(chunk, dtype=np.float32, order="C", copy=True)
    if out.ndim != 2:
        raise ValueError("chunk must be 2-D")
    rows, cols = out.shape
    unsolved = np.zeros((rows, cols), dtype=np.int8)
    _load().orc_breach_single_cell_pits_f32(out.ctypes.data, rows, cols, float(nodata_value), unsolved.ctypes.data)
    return out, unsolved


# ---------------------------------------------------------------- benchmark DEM on the host (synth_host.c)
def synth_dem(rows, cols, row0=0, col0=0, total_rows=None, total_cols=None, seed=0, kind=0, relief=1000.0,
              holes_permille=0, nodata=-9999.0):
    """The DEM ofl_synth_dem_f32 generates on the device (overflow_b200/csrc/synth.cu), computed on the host
    cores, bit for bit: rows row0 .. row0+rows and columns col0 .. col0+cols of the total raster."""
    total_rows = rows if total_rows is None else total_rows
    total_cols = cols if total_cols is None else total_cols
    out = np.empty((rows, cols), dtype=np.float32)
    _load().orc_synth_dem_f32(out.ctypes.data, rows, cols, cols, row0, col0, total_rows, total_cols, int(seed), int(kind),
                              float(relief), int(holes_permille), float(nodata))
    return out
