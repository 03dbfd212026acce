"""Print cycle counts of the tcgen05 micro-benchmarks (run on the B200 box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lrs_pnp_dip_b200 import _lib

L = _lib.diag_lib()
reps, blocks = 2048, 148
def run(do_mma, f16, ts, N, nacc, ldst, depth):
    out = torch.zeros(blocks * 16, dtype=torch.int64, device="cuda")
    for _ in range(2):
        _lib.check(L.lrs_tc_microbench(do_mma, f16, ts, N, nacc, ldst, depth, reps, blocks, out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    o = out.view(blocks, 16).cpu()
    return o[:, 0].float().mean().item() / reps, o[:, 4:12].float().max(dim=1).values.mean().item() / reps
print("MMA chains (cycles per MMA, M=128; K=8 tf32 / K=16 f16):")
for f16 in (0, 1):
    for ts in (1, 0):
        for N in (32, 64, 128, 192, 256):
            m, _ = run(1, f16, ts, N, 1, 0, 1)
            K = 16 if f16 else 8
            print(f"  {'f16 ' if f16 else 'tf32'} {'TS' if ts else 'SS'} N={N:3d}: {m:7.1f} cyc/MMA  ({128*N*K/m:7.0f} MAC/cyc)")
print("TMEM streams by 8 warps (cycles per loop iteration; each warp moves depth*4 KB per direction):")
for ldst, nm in ((1, "ld"), (2, "st"), (3, "ld+st")):
    for depth in (1, 2):
        _, e = run(0, 0, 1, 64, 1, ldst, depth)
        print(f"  {nm:5s} depth={depth}: {e:7.1f} cyc/iter")
print("MMA with concurrent epilogue ld+st traffic:")
for f16, ts, N in ((1, 1, 64), (1, 1, 128), (1, 1, 256), (0, 1, 256)):
    m, e = run(1, f16, ts, N, 1, 3, 2)
    print(f"  f16={f16} TS N={N}: mma {m:7.1f} cyc/MMA, epi {e:7.1f} cyc/iter")
