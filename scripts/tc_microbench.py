"""Print cycle counts of the tcgen05 micro-benchmarks (run on the B200 box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrs_pnp_dip_b200 import _lib

L = _lib.lib()
names = {0: "MMA TS N=64", 1: "MMA TS N=256", 2: "MMA SS N=64", 3: "MMA SS N=256", 4: "tmem ld x32 (8 warps)",
         5: "tmem st x32 (8 warps)", 6: "MMA TS N=64 + concurrent ld+st"}
for blocks in (1, 148):
    for mode in range(7):
        reps = 2048
        out = torch.zeros(blocks * 16, dtype=torch.int64, device="cuda")
        for _ in range(2):
            _lib.check(L.lrs_tc_microbench(mode, reps, blocks, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        o = out.view(blocks, 16).cpu()
        mma = o[:, 0].float().mean().item() / reps
        ep = o[:, 4:12].float().max(dim=1).values.mean().item() / reps
        print(f"blocks={blocks:3d} mode {mode} {names[mode]:34s} mma cyc/op={mma:8.1f}  epi-warp cyc/iter={ep:8.1f}")
