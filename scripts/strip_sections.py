#!/usr/bin/env python
"""Where a distributed step's time goes on one rank: CUDA events around every section of StripPipeline.step
(torchrun, one rank per GPU).  Prints rank 0's and the last rank's mean section times over the steps.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/strip_sections.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from overflow_b200 import _native, strips

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
_native.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S = int(os.environ.get("OFL_SIZE", "65536"))
p = strips.StripPipeline(S, S, rank, world, device=torch.device("cuda", local))
p.load_synthetic(seed=0, kind=0, holes_permille=5)
names = ["fill+dem halo", "direction", "fdr halo", "accum_local", "all_gather", "boundary_solve", "accum_final", "flags"]
acc = [0.0] * len(names)
steps = 10
for it in range(steps + 3):
    dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    ev[0].record()
    p.fill_edge_halos(); p._exchange_halo(p.dem_halo); ev[1].record()
    p.direction(); ev[2].record()
    p._exchange_halo(p.fdr_halo); ev[3].record()
    p.accum_local(); ev[4].record()
    dist.all_gather_into_tensor(p.rec_all, p.rec); ev[5].record()
    p.boundary_solve(); ev[6].record()
    p.accum_final(); ev[7].record()
    p.collect_flags(); dist.all_reduce(p.flags, op=dist.ReduceOp.MAX); strips.raise_for_flags(p.flags.tolist()); ev[8].record()
    torch.cuda.synchronize()
    if it >= 3:
        for k in range(len(names)):
            acc[k] += ev[k].elapsed_time(ev[k + 1])
if rank in (0, world - 1, world // 2):
    print(f"rank {rank}: " + ", ".join(f"{n} {a / steps:.3f}" for n, a in zip(names, acc)) + f"  | total {sum(acc) / steps:.3f} ms", flush=True)
# the step as shipped (DEM halo rows travel under the interior stencil)
for it in range(steps + 3):
    if it == 3:
        dist.barrier(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    p.step()
e1.record(); torch.cuda.synchronize()
if rank == 0:
    print(f"StripPipeline.step(): {e0.elapsed_time(e1) / steps:.3f} ms per step", flush=True)
dist.barrier()
dist.destroy_process_group()
