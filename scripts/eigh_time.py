"""Time torch.linalg.eigh (cuSOLVER syevd) on fp64 Gram matrices of the band counts the hot path meets, plain and
zero-padded to a larger order (the SVT only needs V f(L) V^T on the leading block; see ops.svt_weights)."""
import sys
import torch

torch.manual_seed(0)
dev = torch.device("cuda")


def t_eigh(G, reps=10):
    for _ in range(3):
        torch.linalg.eigh(G)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        torch.linalg.eigh(G)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for C in (31, 64, 128, 191, 224):
    Z = torch.randn(4096, C, device=dev, dtype=torch.float64) @ torch.diag(torch.logspace(0, -4, C, device=dev, dtype=torch.float64))
    G = Z.T @ Z
    row = [f"C={C}: plain {t_eigh(G):.3f} ms"]
    for n in sorted({C + 1, 129, 136, 144, 160, 192, 256}):
        if n <= C:
            continue
        Gp = torch.zeros((n, n), device=dev, dtype=torch.float64)
        Gp[:C, :C] = G
        row.append(f"pad{n} {t_eigh(Gp):.3f}")
    print("  ".join(row), flush=True)
