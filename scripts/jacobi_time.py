"""Time lrs_sym_eig_jacobi_f64 + lrs_svt_weights_f64 (one cluster, no host sync) against torch.linalg.eigh (cuSOLVER syevd) on
Gram matrices of the band counts the hot path meets.  python scripts/jacobi_time.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from lrs_pnp_dip_b200 import ops

torch.manual_seed(0)
dev = torch.device("cuda")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for C in (31, 64, 128, 191, 224, 256):
    Z = torch.rand(8192, C, device=dev, dtype=torch.float64) @ torch.diag(torch.logspace(0, -3, C, device=dev, dtype=torch.float64)) \
        + 0.05 * torch.randn(8192, C, device=dev, dtype=torch.float64)
    G = Z.T @ Z
    sink = lambda s: None
    tj = timed(lambda: ops.svt_weights(G, 1.0 / 0.9, solver="jacobi", status_sink=sink))
    tl = timed(lambda: ops.svt_weights(G, 1.0 / 0.9, solver="library"))
    _, _, st = ops.sym_eig_jacobi(G)
    print(f"C={C:4d}: jacobi {tj:.3f} ms ({int(st[0])} sweeps)   library {tl:.3f} ms", flush=True)
