"""Time lrs_col2im_accum_f32 at the cfg-4 geometry (run on B200); LRS_COL2IM selects an experimental tile shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrs_pnp_dip_b200 import ops

R, C, bb, s = 262144, 191, 8, 1
P = ops.patch_count(R, C, bb, s)
phi = torch.randn(64, P, device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
ops.col2im(phi, R, C, bb, s); torch.cuda.synchronize()
ts = []
for _ in range(5):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.col2im(phi, R, C, bb, s); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
nb = 4 * 64 * P + 4 * R * C
print(f"variant {os.environ.get('LRS_COL2IM', '0')}: {min(ts):.3f} ms  {nb / min(ts) / 1e6:.0f} GB/s")
