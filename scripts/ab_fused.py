"""A/B timing of fused-kernel variants on one box: python scripts/ab_fused.py <name|path>... [--rounds 3] [--cols 24]
Each variant (build_variants/liblrs_pnp_<name>.so, or 'product') is timed in its own process, in ABAB order, on the same
cfg-4 sub-problem: all 262137 row starts x `cols` column starts, K = 256, Nit = 80 (one launch)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, json
sys.path.insert(0, %r)
import numpy as np, torch
import lrs_pnp_dip_b200 as lrs
from lrs_pnp_dip_b200 import synth
cols, reps = int(sys.argv[1]), int(sys.argv[2])
R, C = 262144, 191
rng = np.random.default_rng(0)
X = (rng.standard_normal((R, C)) * 0.3 + 0.5).astype(np.float32)
pm = rng.random(R) < 0.5
Y = np.where(pm[:, None], X, 0).astype(np.float32)
D = synth.synthetic_dictionary(64, 256, 0)
prm = lrs.Params(Nit=80, bb=8, slidingDis=1, step="spectral")
sc = lrs.SparseCoder(torch.tensor(Y).cuda(), torch.tensor(D).cuda(), prm, engine="tc")
Xd = torch.tensor(X).cuda()
nR = R - 7
out = torch.empty(64, cols * nR, device="cuda")
ts = []
for i in range(reps + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sc._fused_range(Xd, None, 0, cols * nR, out=out); e1.record(); torch.cuda.synchronize()
    if i: ts.append(e0.elapsed_time(e1))
print(json.dumps({"ms": float(np.mean(ts)), "min": float(np.min(ts)), "chk": float(out[:, ::1000].double().abs().sum())}))
''' % ROOT


def main():
    args = sys.argv[1:]
    rounds, cols = 3, 24
    names = []
    while args:
        a = args.pop(0)
        if a == "--rounds":
            rounds = int(args.pop(0))
        elif a == "--cols":
            cols = int(args.pop(0))
        else:
            names.append(a)
    res = {n: [] for n in names}
    for r in range(rounds):
        for n in names:
            env = dict(os.environ)
            if n != "product":
                env["LRS_PNP_LIB"] = n if os.path.isfile(n) else os.path.join(ROOT, "build_variants", f"liblrs_pnp_{n}.so")
            out = subprocess.run([sys.executable, "-c", CHILD, str(cols), "3"], env=env, capture_output=True, text=True)
            if out.returncode != 0:
                print(n, "FAILED", out.stderr[-1500:])
                continue
            d = json.loads(out.stdout.strip().splitlines()[-1])
            res[n].append(d)
            print(f"round {r} {n:28s} {d['ms']:9.3f} ms (min {d['min']:.3f})  chk {d['chk']:.6e}", flush=True)
    base = None
    for n in names:
        if res[n]:
            m = sum(d["ms"] for d in res[n]) / len(res[n])
            base = base or m
            print(f"{n:28s} mean {m:9.3f} ms  {100 * (m / base - 1):+6.2f} % vs {names[0]}")


if __name__ == "__main__":
    main()
