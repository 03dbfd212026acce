"""Barrier-wait breakdown of the fused tcgen05 kernel (block 0).  Run on the B200 box:
    LRS_TC_TIMING=1 python scripts/tc_timing.py [rows] [bands]"""
import ctypes, os, sys
os.environ["LRS_TC_TIMING"] = "1"
os.environ["LRS_PNP_DIAGNOSTICS"] = "1"      # the barrier-wait counters exist in liblrs_pnp_diag.so only
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import lrs_pnp_dip_b200 as lrs
from lrs_pnp_dip_b200 import _lib, synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
C = int(sys.argv[2]) if len(sys.argv) > 2 else 191
rng = np.random.default_rng(0)
X = (rng.standard_normal((R, C)) * 0.3 + 0.5).astype(np.float32)
pm = rng.random(R) < 0.5
Y = np.where(pm[:, None], X, 0).astype(np.float32)
D = synth.synthetic_dictionary(64, 256, 0)
prm = lrs.Params(Nit=80, bb=8, slidingDis=1, step="spectral")
sc = lrs.SparseCoder(torch.tensor(Y).cuda(), torch.tensor(D).cuda(), prm, engine="tc")
Xd = torch.tensor(X).cuda()
for _ in range(2):
    sc.phi_z(Xd, None)
torch.cuda.synchronize()
buf = (ctypes.c_uint64 * 32)()
_lib.check(_lib.diag_lib().lrs_tc_timing_read(buf))
t = np.array(buf[:], dtype=np.float64)
its = t[10]
print(f"iterations (block 0): {its:.0f}; MMA-warp cycles/iter {t[0]/its:.0f}")
print("  MMA warp waits/iter: bar_R[kk=0..3] " + " ".join(f"{v/its:6.0f}" for v in t[1:5]) + "   bar_S[0..3] " + " ".join(f"{v/its:6.0f}" for v in t[5:9]))
print(f"epilogue warp 4 cycles/iter {t[16]/its:.0f}: wait bar_A[3] {t[17]/its:.0f}, residual phase {t[21]/its:.0f} (incl. that wait), "
      f"wait bar_B {t[18]/its:.0f}, wait bar_A[0] {t[19]/its:.0f}, bar_A[1] {t[20]/its:.0f}, soft phase {t[22]/its:.0f} (incl. waits), "
      f"tile prologue {t[23]/its:.1f}, final epilogue {t[24]/its:.1f}")
