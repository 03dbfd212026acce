"""Time lrs_im2col_f32 at a cfg-4 row slice (run on B200)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrs_pnp_dip_b200 import ops

C, bb, s = 191, 8, 1
for R, with_l in ((32768, False), (32768, True), (131072, False)):
    P = ops.patch_count(R, C, bb, s)
    X = torch.randn(R, C, device="cuda"); L = torch.randn(R, C, device="cuda")
    flush = torch.empty(256 * 1024 * 1024 // 4, device="cuda")
    args = (X, bb, s, L, 0.15) if with_l else (X, bb, s)
    out = ops.im2col(*args); torch.cuda.synchronize(); del out
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = ops.im2col(*args); e1.record(); torch.cuda.synchronize(); del out
        ts.append(e0.elapsed_time(e1))
    nb = 4 * 64 * P + 4 * R * C * (2 if with_l else 1)
    print(f"im2col R={R} L={with_l}: {min(ts):.3f} ms  {nb / min(ts) / 1e6:.0f} GB/s")
    del X, L, flush
