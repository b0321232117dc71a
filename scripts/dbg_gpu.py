"""Scratch GPU debugging helper (not part of the product)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from oracle import synth
from overflow_b200 import _native
from overflow_b200.flow_direction import flow_direction_for_tile, flow_direction_for_raster
from overflow_b200.flow_accumulation import single_tile_flow_accumulation

g = np.load("tests/golden/kat.npz")
try:
    fdr = flow_direction_for_tile(g["dir_dem"], -9999)
    print("direction KAT:", np.array_equal(fdr[1:-1,1:-1], g["dir_expected"]))
    print(fdr)
except Exception as e:
    print("direction failed:", e)
try:
    fac, links = single_tile_flow_accumulation(g["acc_fdr"])
    print("acc KAT:", np.array_equal(fac, g["acc_fac"]))
    print(fac)
except Exception as e:
    print("acc failed:", e)
