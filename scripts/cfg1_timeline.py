"""Kernel timeline of bundled configuration 1 (bench.py --workload cfg1) through torch.profiler (CUPTI): one line per
kernel / memcpy of two outer iterations — start (us, relative), duration, stream, name — for reading where an outer
iteration spends its time outside the sparse step.   python scripts/cfg1_timeline.py [--no-overlap] > timeline.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

import bench
import lrs_pnp_dip_b200 as lrs

Y, M, D = bench.cfg1_inputs()
dev = torch.device("cuda", 0)
sol = lrs.LRSPnP(torch.from_numpy(Y), torch.from_numpy(M), torch.from_numpy(D), lrs.Params(), device=dev)
if "--no-overlap" in sys.argv:
    sol.overlap_low_rank = False
for _ in range(4):
    sol.reset(); sol.step(); sol.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    sol.reset(); sol.step(); sol.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
print(f"{len(ev)} device activities, span {ev[-1].time_range.end - t0:.1f} us")
last_name, run = None, 0
for e in ev:
    nm = e.name[:70]
    # compress the ISTA chain: print only the first and last of a run of GEMM/reduce launches
    chain = ("tc_gemm_splitk" in nm) or ("reduce_residual" in nm) or ("reduce_gradient" in nm)
    if chain and run > 3:
        run += 1
        last = (e, nm)
        continue
    if not chain and run > 3:
        le, ln = last
        print(f"   ... {run - 4} more chain launches ...")
        print(f"{le.time_range.start - t0:10.1f} {le.time_range.end - le.time_range.start:8.1f}  s{getattr(le, 'device_index', 0)} {ln}")
    run = run + 1 if chain else 0
    print(f"{e.time_range.start - t0:10.1f} {e.time_range.end - e.time_range.start:8.1f}  {nm}")
cpu = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CPU and e.name.startswith(("aten::linalg_eigh", "aten::_linalg_eigh", "cudaStreamSynchronize", "cudaDeviceSynchronize", "cudaMemcpyAsync", "aten::item", "aten::_local_scalar_dense"))]
print("--- host-side blocking calls (start us, duration us) ---")
for e in sorted(cpu, key=lambda e: e.time_range.start):
    print(f"{e.time_range.start - t0:10.1f} {e.time_range.end - e.time_range.start:8.1f}  {e.name}")
