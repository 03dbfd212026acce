"""Time the explicit ISTA engine at the bundled shape (n=1296, K=2592, P=144, Nit=80); run on B200.
    python scripts/ista_tc_time.py [K] [P]
LRS_ISTA_ENGINE=simt selects the FFMA engine."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lrs_pnp_dip_b200 import ops, synth

n, K, Nit = 1296, int(sys.argv[1]) if len(sys.argv) > 1 else 2592, 80
P = int(sys.argv[2]) if len(sys.argv) > 2 else 144      # 2304 = the 144 x 144 crop of main_LRS_PnP.m
rng = np.random.default_rng(0)
D = torch.tensor(synth.synthetic_dictionary(n, K, seed=0)).cuda()
Y = torch.tensor(rng.standard_normal((n, P)).astype(np.float32)).cuda()
mask = torch.tensor((rng.random((n, 1)) < 0.7).repeat(P, axis=1))
bc = torch.where(mask.cuda(), Y + 3.0, torch.zeros_like(Y))
a = ops.step_constants(bc, D, "frob4")
for _ in range(2):
    ops.ista_batched(Y, bc, D, a, 0.1, Nit)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if os.environ.get("PROFILE_RANGE"):
    torch.cuda.profiler.start()
e0.record()
for _ in range(3):
    ops.ista_batched(Y, bc, D, a, 0.1, Nit)
e1.record(); torch.cuda.synchronize()
if os.environ.get("PROFILE_RANGE"):
    torch.cuda.profiler.stop()
ms = e0.elapsed_time(e1) / 3
print(f"engine {os.environ.get('LRS_ISTA_ENGINE', 'auto')}: {ms:.3f} ms per ISTA call ({Nit} iterations, K={K}, P={P}) = {ms / Nit * 1e3:.1f} us per iteration, "
      f"{4.0 * n * K * P * Nit / (ms * 1e-3) / 1e12:.1f} TFLOP/s useful")
