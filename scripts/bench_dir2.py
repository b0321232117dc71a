import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from overflow_b200 import device as dev, _native
R, C = int(sys.argv[1]), int(sys.argv[2])
dem = dev.synth_dem(R, C, seed=0, kind=0, holes_permille=5)
out = torch.empty((R, C), dtype=torch.uint8, device="cuda")
for _ in range(3): dev.flow_direction(dem, -9999.0, out=out)
torch.cuda.synchronize()
_native.phase_timing_read(); _native.phase_timing_enable(True)
for _ in range(10): dev.flow_direction(dem, -9999.0, out=out)
torch.cuda.synchronize()
ms, n = _native.phase_timing_read()["direction"]
print(f"{R}x{C} direction kernel {ms/n:.3f} ms  {R*C/(ms/n)/1e6:.1f} Gcells/s")
