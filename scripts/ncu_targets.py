"""One launch of every helper kernel of the hot path at its benchmark geometry, for `ncu --set full` (run on the B200 box):
col2im (whole Phi_z slice and column-start range mode), im2col, admm_update, Gram, svt_apply, soft, and the explicit
tensor-core ISTA engine (split-K tcgen05 GEMM + reduce kernels) at the bundled shape.

    python scripts/ncu_targets.py && ncu --set full --clock-control none --import-source on \
        -k regex:'col2im|im2col|admm_update|gram_kernel|sgemm_kernel|tc_gemm_splitk|reduce_|soft_kernel' \
        -o gpurun_out/r02_helpers python scripts/ncu_targets.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import lrs_pnp_dip_b200 as lrs
from lrs_pnp_dip_b200 import ops, synth
from lrs_pnp_dip_b200._lib import check, lib, ptr, stream_ptr

dev = "cuda"
C, bb, s = 191, 8, 1
# ---- full cfg-4 row count for the elementwise / Gram kernels
R = 262144
X = torch.randn(R, C, device=dev)
L1, L2 = torch.randn(R, C, device=dev) * 0.01, torch.randn(R, C, device=dev) * 0.01
Y, M, U, IM = X.clone(), torch.ones(R, C, device=dev), torch.randn(R, C, device=dev), torch.randn(R, C, device=dev)
prm = lrs.Params(bb=bb, slidingDis=s)
lrs.admm_update(Y, M, IM, U, L1, L2, prm)
ops.soft_thresh(X, 0.1)
G = torch.zeros(C, C, dtype=torch.float64, device=dev)
check(lib().lrs_gram_f64(ptr(X), ptr(L2), 1.1, R, C, ptr(G), stream_ptr()))
W = torch.randn(C, C, device=dev)
Uo = torch.empty(R, C, device=dev)
check(lib().lrs_svt_apply_f32(ptr(X), ptr(L2), 1.1, ptr(W), R, C, ptr(Uo), stream_ptr()))
# ---- col2im: one column-start range of the cfg-4 pipeline (12 column starts x 262137 row starts, 0.8 GB) as
#      SparseCoder.imout launches it, and the whole-matrix kernel on a 32768-row slice
nR = R - 7
chunk = torch.randn(64, 12 * nR, device=dev)
out = torch.zeros(R, C, device=dev)
check(lib().lrs_col2im_accum_range_f32(ptr(chunk), R, C, bb, s, 24, 36, ptr(out), stream_ptr()))
del chunk
Rs = 32768
Ps = ops.patch_count(Rs, C, bb, s)
phi = torch.randn(64, Ps, device=dev)
ops.col2im(phi, Rs, C, bb, s)
del phi
ops.im2col(X[:Rs].contiguous(), bb, s, L1[:Rs].contiguous(), 0.15)
torch.cuda.synchronize()
# ---- explicit tensor-core engine at the bundled shape (n = 1296, K = 2592, P = 144), 2 iterations
n, K, P = 1296, 2592, 144
D = torch.from_numpy(synth.synthetic_dictionary(n, K, seed=0)).to(dev)
blocks = torch.randn(n, P, device=dev)
bc = (torch.rand(n, P, device=dev) < 0.7).float()
a = torch.full((P,), 9.0, device=dev)
ops.ista_batched(blocks, bc, D, a, 0.1, 2)
torch.cuda.synchronize()
print("ncu_targets: done")
