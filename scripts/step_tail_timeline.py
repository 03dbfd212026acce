"""What follows the fused sparse step inside one cfg-4 outer iteration (torch profiler, rank 0): every device activity after
the last sparse_fused_tc launch ends — overlap sum, halo exchange, Gram, eigensolver, recomposition, X / lambda update,
halo refresh — with start (us after the fused kernel), duration and name.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/step_tail_timeline.py   (or plain python)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

import bench
import lrs_pnp_dip_b200 as lrs
from lrs_pnp_dip_b200 import solver

rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
Y, pm, D = bench.make_inputs("cfg4")
R, C = Y.shape
st = solver.make_stripe(R, bench.BB, rank, world)
Yl = np.ascontiguousarray(Y[st.row_slice])
Ml = np.ascontiguousarray(np.repeat(pm[st.row_slice].astype(np.float32)[:, None], C, axis=1))
prm = lrs.Params(Nit=bench.NIT, bb=bench.BB, slidingDis=1, step="spectral")
sol = lrs.LRSPnP(torch.from_numpy(Yl), torch.from_numpy(Ml), torch.from_numpy(D), prm, stripe=st if world > 1 else None, device=dev)
for _ in range(2):
    sol.step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    sol.step()
    sol.step()
    torch.cuda.synchronize()
if rank == 0:
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    fused = [e for e in ev if "sparse_fused_tc" in e.name]
    # second profiled step: from the end of its last fused launch to the end of the step
    nf = len(fused) // 2
    t_first_end = fused[nf - 1].time_range.end
    t_second_begin = fused[nf].time_range.start
    print(f"world {world}: {len(fused)} fused launches in 2 steps; tail of step 1 = {t_second_begin - t_first_end:.1f} us "
          f"(last fused kernel end -> first fused kernel of the next step)")
    for e in ev:
        if t_first_end - 50 <= e.time_range.start <= t_second_begin:
            print(f"{e.time_range.start - t_first_end:10.1f} {e.time_range.end - e.time_range.start:9.1f}  {e.name[:90]}")
if world > 1:
    dist.destroy_process_group()
