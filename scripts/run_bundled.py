"""BASELINE.json configs 1-3 on the bundled cubes (tests/golden/bundled_inputs.npz) with the seeded synthetic
dictionary (trained_dictionary.mat is not in the reference checkout): ms per outer iteration, ISTA
patch-iterations/s of the sparse step, and the reference-formula MPSNR (main_LRS_PnP.py:379-384).
    python scripts/run_bundled.py [K]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lrs_pnp_dip_b200 as lrs
from lrs_pnp_dip_b200 import matio, synth

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2592
g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bundled_inputs.npz")))
D = synth.synthetic_dictionary(1296, K, seed=0)

def mpsnr(clean, X):
    c = matio.fold_cube(clean.astype(np.float32), 36, 36); x = matio.fold_cube(X, 36, 36)
    return float(np.mean([10 * np.log10(255 / np.sqrt(np.mean((c[0, k] - x[0, k]) ** 2))) for k in range(c.shape[1])]))

def run(tag, prm, iters, low_rank=None, name=""):
    Y, pm, clean = g[f"{tag}_Y"], g[f"{tag}_pixmask"], g[f"{tag}_clean"]
    MtM = np.repeat(pm.astype(np.float32)[:, None], 128, axis=1)
    sol = lrs.LRSPnP(Y, MtM, D, prm, low_rank=low_rank)
    sol.step(); torch.cuda.synchronize()                       # warm-up (also builds cusolver handles)
    sol = lrs.LRSPnP(Y, MtM, D, prm, low_rank=low_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        sol.step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    # sparse step alone
    coder = sol.be.coder
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(); coder.imout(sol.X, sol.lambda_1); s1.record(); torch.cuda.synchronize()
    sp = s0.elapsed_time(s1)
    print(f"{name:44s} K={K}: {ms:8.2f} ms / outer iteration ({iters} its); sparse step {sp:7.2f} ms = "
          f"{144 * prm.Nit / (sp * 1e-3):.3e} patch-iters/s; MPSNR in {mpsnr(clean, Y):.3f} -> out {mpsnr(clean, sol.X.cpu().numpy()):.3f}")

run("img5", lrs.Params(), 2, name="cfg1 main_LRS_PnP.py (img5 + fourth_mask)")
run("base", lrs.Params(), 2, name="cfg1 main_LRS_PnP.py (base cube + mask)")
dip = lrs.Params(mu_1=0.1, mu_2=0.1, Nit=100, step="frob4")
run("img2", dip, 5, low_rank=lambda Z: Z.clone(), name="cfg2 1-LiP ISTA path (img2, U=Z stand-in)")
run("img5", dip, 5, low_rank=lambda Z: Z.clone(), name="cfg2 1-LiP ISTA path (img5, U=Z stand-in)")
run("base", dip, 5, low_rank=None, name="cfg3 DIP_pro sparse path + SVT stand-in")
