"""Achieved HBM GB/s of the gather / scatter / elementwise kernels at the cfg-4 geometry (run on B200).
Algorithmic bytes per SURVEY §8(d); peak = MEASURED_PEAKS.json hbm_gbs."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import lrs_pnp_dip_b200 as lrs
from lrs_pnp_dip_b200 import ops
from lrs_pnp_dip_b200._lib import lib, check, ptr, stream_ptr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
R, C, bb, s = 262144, 191, 8, 1
P = ops.patch_count(R, C, bb, s)
dev = "cuda"
X = torch.randn(R, C, device=dev); L1 = torch.randn(R, C, device=dev) * 0.01; L2 = torch.randn(R, C, device=dev) * 0.01
Y = X.clone(); M = torch.ones(R, C, device=dev); U = torch.randn(R, C, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)          # > L2

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

rows = []
def report(name, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    rows.append((name, ms, nbytes / 1e9, gbs, gbs / peak))
    print(f"{name:34s} {ms:9.3f} ms  {nbytes/1e9:8.3f} GB algorithmic  {gbs:8.1f} GB/s  {100*gbs/peak:5.1f} % of {peak:.0f}")

phi = torch.randn(64, P, device=dev)
report("col2im (cfg4, 12.3 GB Phi_z)", timed(lambda: ops.col2im(phi, R, C, bb, s)), 4 * 64 * P + 4 * R * C)
del phi
prm = lrs.Params(bb=bb, slidingDis=s)
IM = torch.randn(R, C, device=dev)
report("admm_update (X, lam1, lam2)", timed(lambda: lrs.admm_update(Y, M, IM, U, L1, L2, prm)), 36 * R * C)
report("soft threshold", timed(lambda: ops.soft_thresh(X, 0.1)), 8 * R * C)
G = torch.zeros(C, C, dtype=torch.float64, device=dev)
report("gram fp64 (Z = X + c L2)", timed(lambda: check(lib().lrs_gram_f64(ptr(X), ptr(L2), 1.1, R, C, ptr(G), stream_ptr()))), 8 * R * C)
W = torch.randn(C, C, device=dev); Uo = torch.empty(R, C, device=dev)
report("svt_apply (Z W)", timed(lambda: check(lib().lrs_svt_apply_f32(ptr(X), ptr(L2), 1.1, ptr(W), R, C, ptr(Uo), stream_ptr()))), 12 * R * C)
Rs = 32768
Xs = X[:Rs].contiguous(); Ps = ops.patch_count(Rs, C, bb, s)
report("im2col (32768 rows, 1.5 GB out)", timed(lambda: ops.im2col(Xs, bb, s)), 4 * Rs * C + 4 * 64 * Ps)
report("weight", timed(lambda: ops.coverage_weight(R, C, bb, s)), 4 * R * C)
