"""Seeded synthetic DEMs for parity tests, golden vectors and the CPU baseline.

TEST INFRASTRUCTURE (numpy, host).  The shapes follow SURVEY.md section 8(d):
spectral fractal terrain (config 1/2), flat-heavy terraces with nodata blobs
(config 4), tilted planes and a walled serpentine channel (config 5), plus the
float32 edge-case distributions used to pin the direction arithmetic.
"""
import numpy as np

NODATA = -9999.0


def fractal(rows, cols=None, beta=2.0, seed=0, lo=0.0, hi=1000.0):
    """Spectral-synthesis fractal surface, float32, rescaled to [lo, hi]."""
    cols = rows if cols is None else cols
    rng = np.random.default_rng(seed)
    ky = np.fft.fftfreq(rows)[:, None]
    kx = np.fft.rfftfreq(cols)[None, :]
    k = np.sqrt(ky * ky + kx * kx)
    k[0, 0] = 1.0
    amp = k ** (-(beta / 2.0 + 0.5))
    amp[0, 0] = 0.0
    spec = (rng.standard_normal(amp.shape) + 1j * rng.standard_normal(amp.shape)) * amp
    z = np.fft.irfft2(spec, s=(rows, cols))
    z -= z.min()
    m = z.max()
    if m > 0:
        z /= m
    return (lo + z * (hi - lo)).astype(np.float32)


def pad_nodata(dem, nodata=NODATA):
    """One-cell nodata ring (what raster_chunker's halo is at the raster edge)."""
    out = np.full((dem.shape[0] + 2, dem.shape[1] + 2), nodata, dtype=dem.dtype)
    out[1:-1, 1:-1] = dem
    return out


def punch_holes(dem, frac=0.005, seed=1, nodata=NODATA, max_side=None):
    """Rectangular nodata holes covering roughly `frac` of the raster (config 2)."""
    rng = np.random.default_rng(seed)
    rows, cols = dem.shape
    out = dem.copy()
    max_side = max_side or max(2, min(rows, cols) // 16)
    target = int(frac * rows * cols)
    covered = 0
    guard = 0
    while covered < target and guard < 10000:
        h = int(rng.integers(1, max_side + 1))
        w = int(rng.integers(1, max_side + 1))
        r = int(rng.integers(0, max(1, rows - h + 1)))
        c = int(rng.integers(0, max(1, cols - w + 1)))
        out[r : r + h, c : c + w] = nodata
        covered += h * w
        guard += 1
    return out


def terraced(rows, cols=None, seed=0, step=1.0, beta=2.5, relief=40.0, nodata_frac=0.05, nodata=NODATA):
    """Config 4: fractal quantised to `step` (large plateaus), nodata blobs + one nodata edge."""
    cols = rows if cols is None else cols
    z = fractal(rows, cols, beta=beta, seed=seed, lo=0.0, hi=relief)
    z = (np.floor(z / step) * step).astype(np.float32)
    blobs = fractal(rows, cols, beta=3.0, seed=seed + 101, lo=0.0, hi=1.0)
    thr = np.quantile(blobs, 1.0 - nodata_frac)
    z[blobs > thr] = nodata
    z[:, : max(1, cols // 64)] = nodata  # nodata along one raster edge
    return z


def tilted_plane(rows, cols=None, a=1.0, b=0.25):
    """Config 5(i): z = a*(rows-1-row) + b*col -> every cell drains the same way (long parallel chains)."""
    cols = rows if cols is None else cols
    r = np.arange(rows, dtype=np.float64)[:, None]
    c = np.arange(cols, dtype=np.float64)[None, :]
    return (a * (rows - 1 - r) + b * c).astype(np.float32)


def serpentine(rows, cols=None, wall=1.0e6):
    """Config 5(ii): one 1-cell channel snaking through the raster between 1-cell walls.

    Channel rows are the even rows; the channel runs east on rows 0,4,8.. and west on rows
    2,6,.. and steps down through a gap in the wall at the end of each run.  Elevations are
    integer-valued (exact in float32 while rows*cols/2 < 2**24) and strictly decrease along
    the channel, so the chain has about rows*cols/2 cells and a single outlet at its end.
    Walls are high and drain into the channel.
    """
    cols = rows if cols is None else cols
    z = np.full((rows, cols), wall, dtype=np.float64)
    n_chan = 0
    order = []
    # the outermost ring stays wall: raster-edge cells always drain off-raster, so the
    # channel lives in rows/cols 1..n-2 and ends in an interior pit
    for r in range(1, rows - 1, 2):
        east = (r // 2) % 2 == 0
        cs = range(1, cols - 1) if east else range(cols - 2, 0, -1)
        for c in cs:
            order.append((r, c))
        if r + 2 < rows - 1:
            order.append((r + 1, cols - 2 if east else 1))
    n_chan = len(order)
    for k, (r, c) in enumerate(order):
        z[r, c] = float(n_chan - k)
    return z.astype(np.float32)


# --- float32 edge-case distributions for the direction arithmetic (hard parts) ---

def fuzz_dem(kind, rows, cols, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.uniform(0, 1000, (rows, cols)).astype(np.float32)
    if kind == "ints":  # plateaus / exact ties
        return rng.integers(0, 7, (rows, cols)).astype(np.float32)
    if kind == "quant_nodata":
        z = (np.round(rng.uniform(0, 50, (rows, cols)) * 4) / 4).astype(np.float32)
        z[rng.random((rows, cols)) < 0.05] = NODATA
        return z
    if kind == "sqrt2_ties":
        # neighbours placed at z - k*sqrt(2) +- 1 ulp next to cardinal drops of k
        z = np.full((rows, cols), 100.0, dtype=np.float32)
        k = rng.integers(1, 9, (rows, cols)).astype(np.float32)
        diag = (np.float32(100.0) - (k * np.float32(np.sqrt(2.0))).astype(np.float32)).astype(np.float32)
        diag = np.nextafter(diag, np.float32(1e9) * rng.choice([-1, 1], (rows, cols)).astype(np.float32)).astype(np.float32)
        card = (np.float32(100.0) - k).astype(np.float32)
        pick = rng.integers(0, 3, (rows, cols))
        z = np.where(pick == 0, z, np.where(pick == 1, diag, card)).astype(np.float32)
        return z
    if kind == "special":  # NaN / +-inf / nodata fuzz
        z = rng.uniform(-10, 10, (rows, cols)).astype(np.float32)
        u = rng.random((rows, cols))
        z[u < 0.04] = np.nan
        z[(u >= 0.04) & (u < 0.08)] = np.inf
        z[(u >= 0.08) & (u < 0.12)] = -np.inf
        z[(u >= 0.12) & (u < 0.20)] = NODATA
        z[(u >= 0.20) & (u < 0.24)] = -0.0
        z[(u >= 0.24) & (u < 0.28)] = 0.0
        return z
    if kind == "denormal":
        return (rng.integers(0, 4000, (rows, cols)).astype(np.float64) * 1.4e-45).astype(np.float32)
    if kind == "huge":  # differences overflow to inf
        return (rng.choice([-3.0e38, -1.0e38, 1.0e38, 3.0e38, 0.0], (rows, cols))).astype(np.float32)
    if kind == "big_offset":  # float32 subtraction ties (centre 2**24-like offsets)
        return (np.float32(2.0**24) + rng.integers(-8, 8, (rows, cols)).astype(np.float32) * np.float32(0.25)).astype(np.float32)
    if kind == "near_zero":
        return rng.uniform(-1, 1, (rows, cols)).astype(np.float32) * np.float32(1e-3)
    raise ValueError(kind)


FUZZ_KINDS = (
    "uniform",
    "ints",
    "quant_nodata",
    "sqrt2_ties",
    "special",
    "denormal",
    "huge",
    "big_offset",
    "near_zero",
)


# --- host restatement of the device generator (overflow_b200/csrc/synth.cu) ---------------------------------
# bench.py's reference arm and the parity windows need the benchmark DEM without mapping the CUDA library into
# the process.  Every operation below is the float32 / uint32 operation the kernel performs, in its order (the
# library is built with --fmad=false, so nothing is contracted): the arrays are bit-identical to the device's
# (tests/test_gpu_synth.py).

def _hash3(x, y, s):
    with np.errstate(over="ignore"):
        x = x.astype(np.uint32)
        y = y.astype(np.uint32)
        s = np.uint32(s & 0xFFFFFFFF)
        h = (x * np.uint32(0x9E3779B1)) ^ (y * np.uint32(0x85EBCA77) + np.uint32(0x165667B1)) ^ (s * np.uint32(0xC2B2AE3D))
        h ^= h >> np.uint32(15)
        h *= np.uint32(0x2C1B3C6D)
        h ^= h >> np.uint32(12)
        h *= np.uint32(0x297A2D39)
        h ^= h >> np.uint32(15)
    return h


def _lattice(ix, iy, s):
    return (_hash3(ix, iy, s) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def _vnoise(fx, fy, s):
    f32 = np.float32
    flx, fly = np.floor(fx), np.floor(fy)
    ix, iy = flx.astype(np.int64), fly.astype(np.int64)
    tx, ty = fx - flx, fy - fly
    tx = (tx * tx) * (f32(3.0) - f32(2.0) * tx)
    ty = (ty * ty) * (f32(3.0) - f32(2.0) * ty)
    a, b = _lattice(ix, iy, s), _lattice(ix + 1, iy, s)
    c, d = _lattice(ix, iy + 1, s), _lattice(ix + 1, iy + 1, s)
    top = a + (b - a) * tx
    bot = c + (d - c) * tx
    return top + (bot - top) * ty


SERP_WALL = np.float32(8.0e37)
SERP_ORD0 = 0x7E000000


def serpentine_ulp(gr, gc, total_rows, total_cols):
    """Elevations of the device's kind-3 DEM at global rows `gr` (column vector) and columns `gc` (row vector)."""
    gr = np.asarray(gr, dtype=np.int64)
    gc = np.asarray(gc, dtype=np.int64)
    gr, gc = np.broadcast_arrays(gr, gc)
    wc = total_cols - 2
    n_runs = (total_rows - 1) // 2
    odd = (gr & 1) == 1
    m = np.where(odd, (gr - 1) >> 1, (gr >> 1) - 1)
    west = (m & 1) == 1
    pos = np.where(west, total_cols - 2 - gc, gc - 1)
    k_run = m * (wc + 1) + pos
    gap_col = np.where(west, 1, total_cols - 2)
    is_gap = (~odd) & (m + 1 < n_runs) & (gc == gap_col)
    k = np.where(odd, k_run, m * (wc + 1) + wc)
    inside = (gr > 0) & (gr < total_rows - 1) & (gc > 0) & (gc < total_cols - 1)
    chan = inside & (odd | is_gap)
    ordv = SERP_ORD0 - k
    bits = np.where(ordv >= 0, ordv, 0x80000000 | (-ordv)).astype(np.uint32)
    z = bits.view(np.float32)
    return np.where(chan, z, SERP_WALL).astype(np.float32)


def device_dem(rows, cols, row0=0, col0=0, total_rows=None, total_cols=None, seed=0, kind=0, relief=1000.0,
               holes_permille=0, nodata=NODATA):
    """The DEM ofl_synth_dem_f32 writes (kind 0 fractal value noise, 1 terraces, 2 tilted plane, 3 serpentine, 4 the
    serpentine transposed),
    rows row0 .. row0+rows and columns col0 .. col0+cols of a total_rows x total_cols raster, computed on the host."""
    f32 = np.float32
    total_rows = rows if total_rows is None else total_rows
    total_cols = cols if total_cols is None else total_cols
    gr = (row0 + np.arange(rows, dtype=np.int64))[:, None]
    gc = (col0 + np.arange(cols, dtype=np.int64))[None, :]
    seed32 = (seed ^ (seed >> 32)) & 0xFFFFFFFF
    if kind == 2:
        z = (total_rows - 1 - gr).astype(f32) + f32(0.25) * (total_cols - 1 - gc).astype(f32)
        z = np.broadcast_to(z, (rows, cols)).astype(f32)
    elif kind == 3:
        z = serpentine_ulp(gr, gc, total_rows, total_cols)
    elif kind == 4:  # the same channel transposed: it runs north-south
        z = serpentine_ulp(gc, gr, total_cols, total_rows)
    else:
        amp, norm, freq = f32(1.0), f32(0.0), f32(1.0 / 4096.0)
        total = np.zeros((rows, cols), dtype=f32)
        for o in range(12):
            fx = np.broadcast_to(gc.astype(f32) * freq, (rows, cols))
            fy = np.broadcast_to(gr.astype(f32) * freq, (rows, cols))
            total = total + amp * _vnoise(fx, fy, (seed32 + 31 * o) & 0xFFFFFFFF)
            norm = f32(norm + amp)
            amp = f32(amp * f32(0.55))
            freq = f32(freq * f32(2.0))
        z = (f32(relief) * total) / norm
        if kind == 1:
            z = np.floor(z)
        if holes_permille > 0:
            hb = _hash3(np.broadcast_to(gc >> 8, (rows, cols)), np.broadcast_to(gr >> 8, (rows, cols)),
                        seed32 ^ 0xA5A5A5A5)
            is_hole = (hb % np.uint32(1000)).astype(np.int64) < holes_permille * 16
            hx = ((hb >> np.uint32(10)) & np.uint32(127)).astype(np.int64)
            hy = ((hb >> np.uint32(17)) & np.uint32(127)).astype(np.int64)
            cx, cy = gc & 255, gr & 255
            z = np.where(is_hole & (cx >= hx) & (cx < hx + 64) & (cy >= hy) & (cy < hy + 64), f32(nodata), z)
    z = np.where((gr < 0) | (gr >= total_rows), f32(nodata), z)
    return np.ascontiguousarray(z, dtype=f32)
