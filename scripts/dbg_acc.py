"""Scratch GPU debugging helper (not part of the product): small accumulation cases vs the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from overflow_b200 import _native, device as dev

_native.init(0)
g = np.load("tests/golden/kat.npz")
cases = {"kat": g["acc_fdr"].astype(np.uint8)}
cases["one8"] = np.full((1, 1), 8, np.uint8)
cases["row_e"] = np.zeros((1, 5), np.uint8)
cases["col_s"] = np.full((5, 1), 6, np.uint8)
cases["sq_e"] = np.zeros((4, 4), np.uint8)
cases["sq_e64"] = np.zeros((64, 64), np.uint8)
cases["sq_s130"] = np.full((130, 70), 6, np.uint8)
for name, fdr in cases.items():
    want = oracle.flow_accumulation(fdr)
    rows, cols = fdr.shape
    pitch = (cols + 15) // 16 * 16
    buf = torch.zeros((rows, pitch), dtype=torch.uint8, device="cuda")
    d = buf[:, :cols]
    d.copy_(torch.from_numpy(fdr))
    out = torch.zeros((rows, cols), dtype=torch.int64, device="cuda")
    err = None
    try:
        dev.flow_accumulation(d, out=out)
    except Exception as e:
        err = e
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    ok = np.array_equal(got, want)
    print(name, "ok" if ok else "MISMATCH", "ERR %s" % err if err else "")
    if not ok and fdr.size <= 64:
        print(fdr); print(got); print(want)
