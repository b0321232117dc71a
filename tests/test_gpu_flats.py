"""GPU parity: flat resolution (csrc/flats.cu) through the C ABI vs fixtures produced by the reference's
fix_flats.py and vs the C oracle.  The first block reads like the reference's tests/test_fix_flats.py."""
import numpy as np
import pytest

import oracle
from oracle import synth
from conftest import load_golden

pytestmark = pytest.mark.gpu

Z = load_golden("fix_flats.npz")
NAMES = sorted({k.split("__")[0] for k in Z.files})
U = 8


def ff():
    from overflow import fix_flats

    return fix_flats


def edges_of(shape, high, low):
    e = np.zeros(shape, dtype=np.uint8)
    for r, c in low:
        e[r, c] |= 1
    for r, c in high:
        e[r, c] |= 2
    return e


# ---- the reference's own tests (tests/test_fix_flats.py:195-372), same calls, same expectations
EXPECTED_HIGH = [(1, 1), (1, 2), (1, 3), (1, 4), (1, 5), (2, 1), (3, 1), (4, 1), (2, 5), (3, 5), (4, 5), (5, 5), (5, 4)]
EXPECTED_LOW = [(5, 1), (5, 2), (5, 3)]
AWAY_MASK = np.array(
    [
        [0, 0, 0, 0, 0, 0, 0],
        [0, 1, 1, 1, 1, 1, 0],
        [0, 1, 2, 2, 2, 1, 0],
        [0, 1, 2, 3, 2, 1, 0],
        [0, 1, 2, 2, 2, 1, 0],
        [0, 0, 0, 0, 1, 1, 0],
        [0, 0, 0, 0, 0, 0, 0],
    ],
    np.int32,
)


def test_ref_flat_edges():
    high, low = ff().flat_edges(Z["kat__dem"], Z["kat__fdr"])
    assert sorted(high) == sorted(EXPECTED_HIGH)
    assert sorted(low) == sorted(EXPECTED_LOW)


def test_ref_label_flats():
    labels = np.zeros((7, 7), dtype=np.uint32)
    ff().label_flats(Z["kat__dem"], labels, 1, 2, 2)
    assert np.array_equal(labels, Z["kat__labels"].astype(np.uint32))


def test_label_flats_stops_at_labelled_cells():
    """The reference's flood (fix_flats.py:102-103) does not pass through cells that already carry a label."""
    dem = np.zeros((6, 7), dtype=np.float32)
    labels = np.zeros((6, 7), dtype=np.int32)
    labels[:, 3] = 5  # a labelled column splits the one flat in two
    ff().label_flats(dem, labels, 2, 2, 1)
    assert (labels[:, :3] == 2).all() and (labels[:, 3] == 5).all() and (labels[:, 4:] == 0).all()
    ff().label_flats(dem, labels, 9, 0, 3)  # start on a labelled cell: nothing happens
    assert (labels[:, 3] == 5).all() and (labels[:, 4:] == 0).all()
    ff().label_flats(dem, labels, 3, 5, 6)
    assert (labels[:, 4:] == 3).all() and (labels[:, :3] == 2).all()


def test_ref_away_from_higher():
    flat_mask = np.zeros((7, 7), dtype=np.int32)
    flat_height = np.zeros((1), dtype=np.int32)
    ff().away_from_higher(Z["kat__labels"].astype(np.uint32), flat_mask, Z["kat__fdr"], list(EXPECTED_HIGH), flat_height)
    assert flat_height[0] == 3
    assert np.array_equal(flat_mask, AWAY_MASK)


def test_ref_towards_lower():
    flat_mask = AWAY_MASK.copy()
    flat_height = np.array([3], dtype=np.int32)
    ff().towards_lower(Z["kat__labels"].astype(np.uint32), flat_mask, Z["kat__fdr"], list(EXPECTED_LOW), flat_height)
    assert np.array_equal(flat_mask, Z["kat__flat_mask"])


def test_ref_resolve_flats():
    flat_mask, labels = ff().resolve_flats(Z["kat__dem"], Z["kat__fdr"])
    assert flat_mask.dtype == np.int32 and labels.dtype == np.int32
    assert np.array_equal(flat_mask, Z["kat__flat_mask"])
    assert np.array_equal(labels, Z["kat__labels"])


def test_ref_d8_masked_flow_dirs():
    test_fdr = Z["kat__fdr"].copy()
    ff().d8_masked_flow_dirs(Z["kat__flat_mask"].astype(np.uint32), test_fdr, Z["kat__labels"].astype(np.uint32))
    assert np.array_equal(test_fdr, Z["kat__fdr_fixed"])


# ---- every fixture the reference produced
@pytest.mark.parametrize("name", NAMES)
def test_golden_flat_edges(name):
    high, low = ff().flat_edges(Z[f"{name}__dem"], Z[f"{name}__fdr"])
    assert np.array_equal(edges_of(Z[f"{name}__dem"].shape, high, low), Z[f"{name}__edges"])


@pytest.mark.parametrize("name", NAMES)
def test_golden_resolve_flats(name):
    flat_mask, labels = ff().resolve_flats(Z[f"{name}__dem"], Z[f"{name}__fdr"])
    assert np.array_equal(labels, Z[f"{name}__labels"])
    assert np.array_equal(flat_mask, Z[f"{name}__flat_mask"])


@pytest.mark.parametrize("name", NAMES)
def test_golden_masked_dirs(name):
    fdr = Z[f"{name}__fdr"].copy()
    ff().d8_masked_flow_dirs(Z[f"{name}__flat_mask"], fdr, Z[f"{name}__labels"])
    assert np.array_equal(fdr, Z[f"{name}__fdr_fixed"])


@pytest.mark.parametrize("name", NAMES)
def test_golden_fix_flats_one_call(name):
    fixed, flat_mask, labels = ff().fix_flats_for_tile(Z[f"{name}__dem"], Z[f"{name}__fdr"], return_mask=True)
    assert np.array_equal(fixed, Z[f"{name}__fdr_fixed"])
    assert np.array_equal(flat_mask, Z[f"{name}__flat_mask"]) and np.array_equal(labels, Z[f"{name}__labels"])


@pytest.mark.parametrize("name", ["terraced_0", "blocks", "inconsistent_2", "special"])
def test_golden_sweeps_on_their_own(name):
    """away_from_higher / towards_lower from the reference's edge lists reproduce the reference's mask."""
    dem, fdr, labels = Z[f"{name}__dem"], Z[f"{name}__fdr"], Z[f"{name}__labels"]
    e = Z[f"{name}__edges"]
    high = [(int(r), int(c)) for r, c in zip(*np.nonzero(e & 2)) if labels[r, c] != 0]
    low = [(int(r), int(c)) for r, c in zip(*np.nonzero(e & 1))]
    flat_mask = np.zeros(dem.shape, dtype=np.int32)
    flat_height = np.zeros(int(labels.max()) + 1, dtype=np.int32)
    ff().away_from_higher(labels, flat_mask, fdr, high, flat_height)
    ff().towards_lower(labels, flat_mask, fdr, low, flat_height)
    assert np.array_equal(flat_mask, Z[f"{name}__flat_mask"])


# ---- larger rasters against the C oracle
def _cases():
    yield "terraced_600x800", synth.terraced(600, 800, seed=4, relief=25.0)
    yield "terraced_coarse", synth.terraced(1100, 700, seed=9, step=4.0, relief=40.0, nodata_frac=0.02)
    rng = np.random.default_rng(3)
    yield "ints", rng.integers(0, 4, size=(513, 771)).astype(np.float32)
    yield "blocks", np.kron(rng.integers(0, 5, size=(40, 50)), np.ones((16, 13))).astype(np.float32)
    yield "fractal", synth.punch_holes(synth.fractal(900, 900, beta=2.0, seed=0), frac=0.01, seed=1)
    wide = np.zeros((300, 2500), dtype=np.float32)  # one flat 2500 cells long: deep sweeps, long union-find chains
    wide[:, 0] = -1.0
    wide[0, :] = 5.0
    yield "one_long_flat", wide
    yield "thin", np.round(synth.fractal(2, 3000, beta=2.0, seed=7) / 100.0).astype(np.float32)
    # the sweeps relax 32 x 32 tiles: a lake whose levels cross many tiles, drained through one channel ...
    lake = np.full((700, 650), 10.0, dtype=np.float32)
    lake[9:-9, 9:-9] = 5.0
    lake[350, :10] = np.linspace(1.0, 4.5, 10, dtype=np.float32)
    yield "lake_with_outlet", lake
    # ... and one flat that winds between walls, entering and leaving every tile it meets several times
    maze = np.zeros((203, 330), dtype=np.float32)
    maze[0, :] = maze[-1, :] = maze[:, 0] = maze[:, -1] = 7.0
    for k, r in enumerate(range(3, 200, 3)):
        maze[r, (1 if k % 2 else 3):(-3 if k % 2 else -1)] = 7.0
    maze[1, 1] = -1.0
    yield "winding_flat", maze


def _cases_vector_path():
    """Widths that are multiples of 32 take the four-cells-per-thread edge stencil (flat_edges4_kernel)."""
    yield "terraced_257x160", synth.terraced(257, 160, seed=14, relief=12.0, nodata_frac=0.03)
    rng = np.random.default_rng(5)
    yield "ints_130x96", rng.integers(0, 3, size=(130, 96)).astype(np.float32)
    yield "one_span_9x32", rng.integers(0, 2, size=(9, 32)).astype(np.float32)
    runs = np.zeros((40, 384), dtype=np.float32)  # runs that cross the 128-cell seams of a warp's span, and rows of one run
    runs[::3, 100:300] = 1.0
    runs[1::3, :] = 2.0
    runs[5, 127] = 7.0
    runs[7, 128] = 7.0
    yield "runs_across_seams", runs
    odd = synth.terraced(64, 224, seed=15, relief=6.0)
    odd[10:14, 30:50] = np.nan
    odd[20, 60:70] = np.inf
    odd[22, 60:70] = -np.inf
    yield "nan_inf_64x224", odd


@pytest.mark.parametrize("name,dem", list(_cases_vector_path()), ids=[n for n, _ in _cases_vector_path()])
def test_vs_oracle_vector_edges(name, dem, monkeypatch):
    fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1].copy()
    want_mask, want_labels = oracle.resolve_flats(dem, fdr)
    assert ff().flat_edges(dem, fdr) == oracle.flat_edges(dem, fdr)
    want_fixed = oracle.d8_masked_flow_dirs(want_mask, fdr, want_labels)
    flat_mask, labels = ff().resolve_flats(dem, fdr)
    assert np.array_equal(labels, want_labels) and np.array_equal(flat_mask, want_mask)
    assert np.array_equal(ff().fix_flats_for_tile(dem, fdr), want_fixed)
    monkeypatch.setenv("OFL_FLATS_SCALAR", "1")  # the one-cell-per-thread kernels on the same raster
    assert ff().flat_edges(dem, fdr) == oracle.flat_edges(dem, fdr)
    flat_mask, labels = ff().resolve_flats(dem, fdr)
    assert np.array_equal(labels, want_labels) and np.array_equal(flat_mask, want_mask)
    assert np.array_equal(ff().fix_flats_for_tile(dem, fdr), want_fixed)


@pytest.mark.parametrize("name,dem", list(_cases()), ids=[n for n, _ in _cases()])
def test_vs_oracle(name, dem):
    fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1].copy()
    want_mask, want_labels = oracle.resolve_flats(dem, fdr)
    want_fixed = oracle.d8_masked_flow_dirs(want_mask, fdr, want_labels)
    flat_mask, labels, info = ff().resolve_flats(dem, fdr, return_info=True)
    assert np.array_equal(labels, want_labels)
    assert np.array_equal(flat_mask, want_mask)
    assert info["labels"] == int(want_labels.max())
    fixed = ff().fix_flats_for_tile(dem, fdr)
    assert np.array_equal(fixed, want_fixed)
    high, low = ff().flat_edges(dem, fdr)
    assert (high, low) == oracle.flat_edges(dem, fdr)


def test_device_tensors():
    """Device-resident chain: direction -> fix_flats, checked against the oracle chain."""
    import torch
    from overflow_b200 import device as dev

    dem = synth.terraced(1024, 1536, seed=12, relief=30.0)
    d_dem = torch.from_numpy(dem).cuda()
    d_fdr = dev.flow_direction(d_dem, synth.NODATA).contiguous()
    before = d_fdr.cpu().numpy()
    want_fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
    assert np.array_equal(before, want_fdr)
    d_fdr, info = dev.fix_flats(d_dem, d_fdr)
    want_mask, want_labels = oracle.resolve_flats(dem, before)
    want_fixed = oracle.d8_masked_flow_dirs(want_mask, before, want_labels)
    fixed = d_fdr.cpu().numpy()
    assert np.array_equal(fixed, want_fixed)
    assert info[2] == int(want_labels.max()) and (fixed == U).sum() < (before == U).sum()


def test_routing_through_resolved_flats():
    """direction -> fix_flats -> accumulation on the device.  A stepped plane has only drainable flats, so the
    rewritten codes are acyclic and the counts must satisfy the accumulation recurrence exactly.  (On DEMs with
    pits the reference's d8_masked_flow_dirs points a pit at its first label-0 neighbour, which can close a
    2-cycle; the reference breaches pits before this step.)"""
    import torch
    from overflow_b200 import device as dev

    rows, cols = 768, 1280
    r = np.arange(rows, dtype=np.float64)[:, None]
    c = np.arange(cols, dtype=np.float64)[None, :]
    dem = np.floor((rows - 1 - r) * 0.07 + c * 0.03).astype(np.float32)
    d_dem = torch.from_numpy(dem).cuda()
    d_fdr = dev.flow_direction(d_dem, synth.NODATA).contiguous()
    before = d_fdr.cpu().numpy()
    assert (before == U).sum() > rows * cols // 2
    d_fdr, info = dev.fix_flats(d_dem, d_fdr)
    fixed = d_fdr.cpu().numpy()
    want_mask, want_labels = oracle.resolve_flats(dem, before)
    assert np.array_equal(fixed, oracle.d8_masked_flow_dirs(want_mask, before, want_labels))
    assert (fixed == U).sum() == 0
    fac = dev.flow_accumulation(d_fdr)
    assert dev.check_accumulation(d_fdr, fac) == 0
    assert np.array_equal(fac.cpu().numpy(), oracle.flow_accumulation(fixed))
    # the same chain as one device call
    fdr2, fac2 = dev.flow_routing(d_dem, synth.NODATA, resolve_flats=True)
    assert torch.equal(fdr2, d_fdr) and torch.equal(fac2, fac)


@pytest.mark.parametrize("cols", [131, 132])  # 132: the four-cells-per-thread kernel (width a multiple of 4)
@pytest.mark.parametrize("span", [3, 1000, 2**20 - 1, 2**20 + 5, 2**30])
def test_masked_dirs_value_ranges(span, cols):
    """d8_masked_flow_dirs on arbitrary masks, from tiny to int32-wide differences, with near-ties between cardinal
    and diagonal steps planted: the float64 slopes must give the reference's (= the oracle's) choice everywhere."""
    rng = np.random.default_rng(span % 1000)
    shape = (97, cols)
    flat_mask = rng.integers(-span, span + 1, size=shape).astype(np.int32)
    near = np.round(flat_mask[:, :-1].astype(np.float64) * np.sqrt(2.0))  # near-ties between cardinal and diagonal steps
    flat_mask[:, 1:] = np.where(rng.random((97, cols - 1)) < 0.3, np.clip(near, -2**31 + 1, 2**31 - 1).astype(np.int64),
                                flat_mask[:, 1:]).astype(np.int32)
    labels = rng.integers(-1, 3, size=shape).astype(np.int32)  # -1: nothing special about it, it only has to match
    fdr = rng.choice(np.array([0, 3, 8, 8, 8, 9], dtype=np.uint8), size=shape)
    want = oracle.d8_masked_flow_dirs(flat_mask, fdr, labels)
    got = fdr.copy()
    ff().d8_masked_flow_dirs(flat_mask, got, labels)
    assert np.array_equal(got, want)


def test_rejects_inexact_dtypes_and_sizes():
    with pytest.raises(TypeError):
        ff().resolve_flats(np.array([[1e-50, 2.0]], dtype=np.float64), np.array([[8, 8]], dtype=np.uint8))
    fm, lb = ff().resolve_flats(np.array([[1, 1], [1, 0]], dtype=np.int16), np.array([[8, 8], [8, 7]], dtype=np.uint8))
    assert fm.shape == (2, 2) and lb.dtype == np.int32
    fm, lb = ff().resolve_flats(np.zeros((0, 5), dtype=np.float32), np.zeros((0, 5), dtype=np.uint8))
    assert fm.shape == (0, 5)


def test_tile_of_more_than_2_31_cells():
    """Cell indices are unsigned 32-bit on the device: a 532 480 x 4096 raster (2.18e9 cells) made of 1040 copies
    of one 512-row terraced block whose first and last rows are NODATA (inert for every step of the algorithm), so
    each copy must come out like the block on its own, with its labels shifted by the labels of the copies above."""
    import torch

    from overflow_b200 import device as dev

    h, cols, reps = 512, 4096, 1040
    free, _ = torch.cuda.mem_get_info()
    if free < 90 * 2**30:
        pytest.skip("needs 90 GB of device memory")
    block = np.full((h, cols), synth.NODATA, dtype=np.float32)
    block[1:-1] = synth.terraced(h - 2, cols, seed=21, relief=30.0, nodata_frac=0.01)
    fdr_b = oracle.flow_direction_for_tile(synth.pad_nodata(block), synth.NODATA)[1:-1, 1:-1].copy()
    want_mask, want_labels = oracle.resolve_flats(block, fdr_b)
    want_fixed = oracle.d8_masked_flow_dirs(want_mask, fdr_b, want_labels)
    n_lab = int(want_labels.max())
    assert n_lab > 100 and int(want_mask.max()) > 10

    dem = torch.from_numpy(block).cuda().repeat(reps, 1)
    assert dem.numel() > 2**31
    fdr = torch.from_numpy(fdr_b).cuda().repeat(reps, 1)
    flat_mask = torch.empty((h * reps, cols), dtype=torch.int32, device="cuda")
    labels = torch.empty((h * reps, cols), dtype=torch.int32, device="cuda")
    fdr, info = dev.fix_flats(dem, fdr, flat_mask=flat_mask, labels=labels)
    del dem
    assert info[2] == n_lab * reps
    assert bool((fdr.view(reps, h, cols) == torch.from_numpy(want_fixed).cuda()[None]).all())
    assert bool((flat_mask.view(reps, h, cols) == torch.from_numpy(want_mask).cuda()[None]).all())
    del flat_mask, fdr
    wl = torch.from_numpy(want_labels.astype(np.int32)).cuda()
    for k0 in range(0, reps, 130):  # in slices: the shifted comparison needs a temporary per slice
        got = labels.view(reps, h, cols)[k0 : k0 + 130]
        shift = (torch.arange(k0, k0 + got.shape[0], device="cuda", dtype=torch.int32) * n_lab)[:, None, None]
        assert bool((got == torch.where(wl[None] > 0, wl[None] + shift, 0)).all())
