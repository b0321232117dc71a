"""Small end-to-end run for compute-sanitizer (memcheck): KATs, one odd-shaped raster, 3 strips in-process."""
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
print("sanitize_small ok")
