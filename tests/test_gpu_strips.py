"""GPU parity of the row-strip path: N logical strips on ONE GPU (loop-back exchange) must equal the
single-raster result and the CPU oracle, cell for cell."""
import numpy as np
import pytest
import torch

import oracle
from oracle import synth

pytestmark = pytest.mark.gpu


def run_strips(dem, world):
    from overflow_b200 import strips

    rows, cols = dem.shape
    pipes = [strips.StripPipeline(rows, cols, r, world, nodata=synth.NODATA, device="cuda:0") for r in range(world)]
    for p in pipes:
        p.load_dem(torch.from_numpy(dem[p.r0 : p.r1]).cuda())
    strips.step_in_process(pipes)
    torch.cuda.synchronize()
    return (np.concatenate([p.fdr.cpu().numpy() for p in pipes]), np.concatenate([p.fac.cpu().numpy() for p in pipes]))


def cases():
    yield "fractal_holes", synth.punch_holes(synth.fractal(640, 300, beta=2.0, seed=1), frac=0.02, seed=2)
    yield "fractal_smooth", synth.fractal(512, 257, beta=4.0, seed=3)
    yield "terraced", synth.terraced(448, 200, seed=4)
    yield "tilted_south", synth.tilted_plane(512, 130)
    yield "tilted_north_west", synth.tilted_plane(384, 130, a=-1.0, b=-0.5)
    yield "serpentine", synth.serpentine(321, 67)


@pytest.mark.parametrize("name,dem", list(cases()), ids=[n for n, _ in cases()])
@pytest.mark.parametrize("world", [1, 2, 4, 5])
def test_strips_match_oracle(name, dem, world):
    want_fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
    want_fac = oracle.flow_accumulation(want_fdr)
    fdr, fac = run_strips(dem, world)
    assert np.array_equal(fdr, want_fdr)
    assert np.array_equal(fac, want_fac)


def test_eight_strips_large_device_raster():
    """8 strips of a 4096 x 2048 device-generated raster == single-raster device result == recurrence."""
    from overflow_b200 import device as dev, strips

    rows, cols, world = 4096, 2048, 8
    pipes = [strips.StripPipeline(rows, cols, r, world, nodata=synth.NODATA, device="cuda:0") for r in range(world)]
    for p in pipes:
        p.load_synthetic(seed=11, kind=0, holes_permille=10)
    strips.step_in_process(pipes)
    fdr = torch.cat([p.fdr for p in pipes]).contiguous()
    fac = torch.cat([p.fac for p in pipes]).contiguous()
    dem = dev.synth_dem(rows, cols, seed=11, kind=0, holes_permille=10)
    one_fdr = dev.flow_direction(dem, synth.NODATA)
    one_fac = dev.flow_accumulation(one_fdr)
    assert torch.equal(fdr, one_fdr)
    assert torch.equal(fac, one_fac)
    assert dev.check_accumulation(fdr, fac) == 0
    # the strip form of the recurrence check: clean, and it sees a count that is off on a strip boundary
    assert strips.check_in_process(pipes) == 0
    pipes[3].fac[0, 1000] += 1
    assert strips.check_in_process(pipes) >= 1


def test_cycle_is_reported_through_the_flags():
    """A cyclic code raster: the strip calls stay asynchronous and the step's one status read raises OFL_ERR_CYCLE."""
    from overflow_b200 import _native, strips

    rows, cols, world = 256, 128, 2
    pipes = [strips.StripPipeline(rows, cols, r, world, nodata=synth.NODATA, device="cuda:0") for r in range(world)]
    for p in pipes:
        p.fdr_halo.fill_(0)  # everything flows east ...
        p.fdr[:, cols - 1] = 8
    pipes[1].fdr[10, 5] = 4  # ... except one cell that flows back west: a two-cell cycle
    for p in pipes:
        p.accum_local()
    for p in pipes:
        for i, q in enumerate(pipes):
            if i != p.rank:
                p.rec_all[i].copy_(q.rec)
    for p in pipes:
        p.boundary_solve()
        p.accum_final()
        p.collect_flags()
    strips.raise_for_flags(pipes[0].flags.tolist())
    with pytest.raises(_native.OverflowB200Error) as ei:
        strips.raise_for_flags(pipes[1].flags.tolist())
    assert ei.value.status == _native.OFL_ERR_CYCLE


NCCL_SCRIPT = r"""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, os.environ["OFL_ROOT"])
from overflow_b200 import _native, device as dev, strips
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
_native.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for rows, cols, kind, holes in ((4096, 2048, 0, 10), (2048, 1024, 3, 0), (2048, 1024, 4, 0), (4096, 1024, 2, 0), (1024, 4096, 1, 50)):
    p = strips.StripPipeline(rows, cols, rank, world, nodata=-9999.0, device=torch.device("cuda", local))
    p.load_synthetic(seed=5, kind=kind, holes_permille=holes)
    for _ in range(3):  # repeated steps: stream ordering between NCCL and the library's kernels
        p.step()
    bad = torch.tensor([p.check()], device="cuda")
    dist.all_reduce(bad)
    dem = dev.synth_dem(rows, cols, seed=5, kind=kind, holes_permille=holes)
    fdr, fac = dev.flow_routing(dem, -9999.0)
    same = torch.tensor([int(torch.equal(p.fdr, fdr[p.r0:p.r1]) and torch.equal(p.fac, fac[p.r0:p.r1]))], device="cuda")
    dist.all_reduce(same)
    ok &= int(bad.item()) == 0 and int(same.item()) == world
    if rank == 0:
        print("case", rows, cols, kind, "violations", int(bad.item()), "strips equal", int(same.item()), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_strips_equal_single_gpu(world, tmp_path):
    """The path that actually travels over NCCL (torchrun, one rank per GPU): every rank's strip equals the same rows of
    a single-GPU run of the whole raster, cell for cell, and the strip recurrence check is clean."""
    import os
    import subprocess
    import sys

    from conftest import ROOT

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    script = tmp_path / "nccl_strips.py"
    script.write_text(NCCL_SCRIPT)
    env = dict(os.environ, OFL_ROOT=ROOT)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), str(script)],
                         capture_output=True, text=True, timeout=900, env=env)
    assert res.returncode == 0, (res.stdout + res.stderr)[-3000:]
    assert res.stdout.count("strips equal") == 5


@pytest.mark.parametrize("general", [False, True])
def test_both_routes_of_the_cross_strip_inflow(general, monkeypatch):
    """The inflow from other strips walks the forest from its entry cells when every path is short (terrain) and
    takes subtree sums over the whole forest otherwise (long channels); OFL_SEED_GENERAL forces the second route."""
    if general:
        monkeypatch.setenv("OFL_SEED_GENERAL", "1")
    dem = synth.punch_holes(synth.fractal(1024, 700, beta=2.5, seed=21), frac=0.02, seed=22)
    want_fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
    fdr, fac = run_strips(dem, 4)
    assert np.array_equal(fdr, want_fdr) and np.array_equal(fac, oracle.flow_accumulation(want_fdr))
    # one channel of 150 000 cells through four strips: far beyond the walk's limit, whatever the switch says
    chan = oracle.synth_dem(768, 400, kind=3)
    want_fdr = oracle.flow_direction_for_tile(synth.pad_nodata(chan), synth.NODATA)[1:-1, 1:-1]
    fdr, fac = run_strips(chan, 4)
    assert np.array_equal(fdr, want_fdr) and np.array_equal(fac, oracle.flow_accumulation(want_fdr))


def test_long_chain_across_all_strips():
    """config 5: a tilted plane drains every column through all 8 strips; counts reach the row count."""
    from overflow_b200 import strips

    rows, cols, world = 1024, 192, 8
    dem = synth.tilted_plane(rows, cols, a=1.0, b=0.0)
    fdr, fac = run_strips(dem, world)
    want_fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
    assert np.array_equal(fdr, want_fdr)
    assert np.array_equal(fac, oracle.flow_accumulation(want_fdr))
    assert fac.max() >= rows - 1
    del strips


@pytest.mark.parametrize("shape,world", [((257, 65), 3), ((200, 1000), 2), ((1000, 37), 5), ((449, 129), 7)])
def test_ragged_shapes_with_nodata_on_strip_boundaries(shape, world):
    """nodata rows / blobs exactly on strip boundaries, widths that are not multiples of the tile or band."""
    from overflow_b200 import strips

    rows, cols = shape
    dem = synth.punch_holes(synth.fractal(rows, cols, beta=2.5, seed=rows + cols), frac=0.03, seed=7)
    for r0, _ in strips.partition_rows(rows, world)[1:]:
        dem[r0 - 1 : r0 + 1, cols // 4 : cols // 2] = synth.NODATA  # nodata straddling the boundary
    want_fdr = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
    want_fac = oracle.flow_accumulation(want_fdr)
    fdr, fac = run_strips(dem, world)
    assert np.array_equal(fdr, want_fdr)
    assert np.array_equal(fac, want_fac)


@pytest.mark.parametrize("name,dem", [c for c in cases() if c[0] in ("fractal_holes", "tilted_south", "serpentine")],
                         ids=["fractal_holes", "tilted_south", "serpentine"])
@pytest.mark.parametrize("strip_rows", [64, 192])
def test_out_of_core_accumulation(name, dem, strip_rows):
    """SURVEY 8f rank 3, full-width tiles: the raster goes through the device strip by strip, twice, and the counts
    equal the whole-raster ones cell for cell."""
    from overflow_b200 import strips

    fdr = np.ascontiguousarray(oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1])
    want = oracle.flow_accumulation(fdr)
    got = np.full(fdr.shape, -7, dtype=np.int64)

    def write_rows(r0, fac):
        got[r0 : r0 + len(fac)] = fac

    n = strips.flow_accumulation_out_of_core(lambda r0, r1: fdr[r0:r1], write_rows, fdr.shape[0], fdr.shape[1],
                                             strip_rows, device="cuda:0")
    assert n == len(strips.strip_bounds(fdr.shape[0], strip_rows))
    assert np.array_equal(got, want)
