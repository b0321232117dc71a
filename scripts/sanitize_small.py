"""Small end-to-end run over odd shapes (ragged warps, ragged four-byte words): KATs, one odd-shaped raster, 3 strips
in-process, flat resolution, pit breaching, out-of-core strips, all against the oracle.  Written for
compute-sanitizer; where that tool is closed on the GPU pool it runs as a plain small check (OFL_DEBUG_SYNC=1 names
the kernel behind a fault)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from oracle import synth
from overflow_b200.flow_direction import flow_direction_for_tile, flow_direction_for_raster
from overflow_b200.flow_accumulation import single_tile_flow_accumulation
from overflow_b200 import strips

g = np.load("tests/golden/kat.npz")
assert np.array_equal(flow_direction_for_tile(g["dir_dem"], -9999)[1:-1, 1:-1], g["dir_expected"])
fac, links = single_tile_flow_accumulation(g["acc_fdr"])
assert np.array_equal(fac, g["acc_fac"])
dem = synth.punch_holes(synth.fractal(203, 331, beta=2.0, seed=1), frac=0.02, seed=2)
fdr = flow_direction_for_raster(dem, synth.NODATA)
want = oracle.flow_direction_for_tile(synth.pad_nodata(dem), synth.NODATA)[1:-1, 1:-1]
assert np.array_equal(fdr, want)
fac, links = single_tile_flow_accumulation(fdr)
assert np.array_equal(fac, oracle.flow_accumulation(fdr))
pipes = [strips.StripPipeline(203, 331, r, 3, nodata=synth.NODATA, device="cuda:0") for r in range(3)]
for p in pipes:
    p.load_dem(torch.from_numpy(dem[p.r0:p.r1]).cuda())
strips.step_in_process(pipes)
torch.cuda.synchronize()
assert np.array_equal(np.concatenate([p.fac.cpu().numpy() for p in pipes]), oracle.flow_accumulation(want))
# flat resolution and pit breaching (odd shapes: ragged warps, ragged four-byte words)
from overflow_b200.fix_flats import fix_flats_for_tile
from overflow_b200.breach_single_cell_pits import breach_single_cell_pits_in_chunk

tdem = synth.terraced(131, 203, seed=5, relief=9.0)
tfdr = np.ascontiguousarray(oracle.flow_direction_for_tile(synth.pad_nodata(tdem), synth.NODATA)[1:-1, 1:-1])
m, l = oracle.resolve_flats(tdem, tfdr)
assert np.array_equal(fix_flats_for_tile(tdem, tfdr), oracle.d8_masked_flow_dirs(m, tfdr, l))
chunk = np.random.default_rng(3).uniform(0, 50, (77, 131)).astype(np.float32)
want_chunk, want_unsolved = oracle.breach_single_cell_pits_in_chunk(chunk, synth.NODATA)
assert np.array_equal(breach_single_cell_pits_in_chunk(chunk, synth.NODATA), want_unsolved)
assert np.array_equal(chunk.view(np.uint32), want_chunk.view(np.uint32))
got = np.zeros(want.shape, dtype=np.int64)
strips.flow_accumulation_out_of_core(lambda a, b: want[a:b], lambda a, f: got.__setitem__(slice(a, a + len(f)), f),
                                     want.shape[0], want.shape[1], 64, device="cuda:0")
assert np.array_equal(got, oracle.flow_accumulation(want))
print("sanitize_small ok")
