"""Scratch: time the direction kernel alone (device-resident), check parity on windows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from overflow_b200 import device as dev
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dem = dev.synth_dem(S, S, seed=0, kind=0, holes_permille=5)
out = torch.empty((S, S), dtype=torch.uint8, device="cuda")
for _ in range(3): dev.flow_direction(dem, -9999.0, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
for _ in range(N): dev.flow_direction(dem, -9999.0, out=out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
ok = True
for r, c in ((1, 1), (S // 2, S // 3), (S - 402, S - 402)):
    win = dem[r - 1:r + 401, c - 1:c + 401].cpu().numpy()
    ok &= bool(np.array_equal(out[r:r + 400, c:c + 400].cpu().numpy(), oracle.flow_direction_for_tile(win, -9999.0)[1:-1, 1:-1]))
print(f"S={S} direction {ms:.3f} ms  {S*S/ms/1e6:.1f} Gcells/s  {S*S*5/ms/1e6:.0f} GB/s  parity={ok}")
