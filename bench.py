#!/usr/bin/env python
"""Benchmark of the LRS-PnP hot path on B200:  python bench.py --gpus N --steps K --warmup W

A *step* is one ADMM outer iteration of main_LRS_PnP.py:250-362 on synthetic data: the sparse-coding
step (implicit 8x8 stride-1 patches → Nit soft-ISTA iterations → Phi_z), the overlap sum, the SVT
low-rank step and the closed-form X/λ update.  Metric (BASELINE.json): ISTA patch-iterations/s =
P*Nit / step time, whole job over all N GPUs; ms_per_step = ms per outer iteration.

Workload: BASELINE.json configs[3] — synthetic 512x512x191 cube (R = 262144 unfolded rows x 191
bands), 8x8 patches stride 1 (P = 48 233 208), 50 % per-pixel Bernoulli mask, K = 256 atoms,
Nit = 80.  N > 1: the same cube sharded as row stripes, one rank per GPU (strong scaling).

--impl reference: the CPU port of the reference path (oracle/, NumPy+BLAS on all host threads) on a
bounded patch sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (H, W, bands, mask kind, keep)
    "cfg4": (512, 512, 191, "bernoulli", 0.5),
    "cfg5": (1024, 1024, 224, "stripe+bernoulli", 0.75),
    "mini": (64, 64, 32, "bernoulli", 0.5),
}
K_ATOMS, NIT, BB, STRIDE = 256, 80, 8, 1
METRIC, UNIT = "ista_patch_iters_per_s", "patch-iters/s"


# The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner on the first
# communicator, "NCCL version 2.28.9+cuda12.9"), so the process keeps a private handle to the original stdout for that
# line and points file descriptor 1 at stderr for everything else.
_RESULT_OUT = None


def claim_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _RESULT_OUT if _RESULT_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sust=1400.0, src="fallback")


def ncu_traffic(workload, world):
    """DRAM bytes of the dominant kernel per step from the COMMITTED ncu capture of the same workload (or None): a stored
    figure (profiles/r02_traffic.json says how it was taken), not something this run measures."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        return json.load(open(p)).get(workload, {}).get(str(world))
    except (OSError, ValueError):
        return None


def make_inputs(workload):
    from lrs_pnp_dip_b200 import synth

    H, W, B, kind, keep = WORKLOADS[workload]
    clean, noisy = synth.synthetic_cube(H, W, B, rank=8, seed=0)
    pm = synth.pixel_mask(H, W, kind, keep=keep, seed=3)
    Y = synth.observe(noisy, pm)
    del clean, noisy
    D = synth.synthetic_dictionary(BB * BB, K_ATOMS, seed=0)
    return Y, pm, D


# ---------------------------------------------------------------------------------------------- CPU arm
CPU_SAMPLE_ROWS = 192      # FIXED sample: the first 192 unfolded rows = 185 patch row starts x ALL column starts


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to its ranks; the CPU legs are meant to use every host core, so the BLAS / OpenMP
    pools are resized at run time (threadpoolctl) and torch's intra-op pool too.  Returns the thread count in effect."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits

        threadpool_limits(limits=n)
        got = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") in ("blas", "openmp")]
        n_eff = max(got) if got else n
    except Exception:
        n_eff = n
    try:
        import torch

        torch.set_num_threads(n)
    except Exception:
        pass
    return int(n_eff)


def cpu_sample(Y):
    """The bounded CPU sample of a workload — independent of any time budget, so the figure is reproducible."""
    return np.ascontiguousarray(Y[:min(CPU_SAMPLE_ROWS, Y.shape[0])])


def cpu_pass(sub, D):
    """ONE pass of the oracle port (NumPy restatement of main_LRS_PnP.py:259-303, BLAS on all host threads) over the
    sample: im2col of the observed and current matrices, step constants, Nit masked soft-ISTA iterations, Phi_z.
    Returns (patches, seconds)."""
    from oracle import lrs_oracle as orc

    oprm = orc.Params(Nit=NIT, bb=BB, slidingDis=STRIDE, step="spectral")
    t0 = time.perf_counter()
    phi, _ = orc.sparse_step(sub, np.zeros_like(sub), sub, D, oprm)
    return phi.shape[1], time.perf_counter() - t0


def sample_text(P_s, rows, C):
    return (f"{P_s} patches = first {rows} unfolded rows ({rows - BB + 1} row starts) x all {C - BB + 1} column starts, "
            f"x {NIT} ISTA iterations per pass; fixed sample, independent of the time budget")


def cpu_patch_iters_per_s(Y, D, budget_s=15.0):
    """cpu_baseline leg: one untimed pass (thread pools, page faults), then whole passes over the FIXED sample until
    ``budget_s`` is used (at least one); value = patch-iterations / mean pass time."""
    sub = cpu_sample(Y)
    threads = use_all_host_threads()
    cpu_pass(sub, D)
    times, t_end = [], time.perf_counter() + budget_s
    while True:
        P_s, dt = cpu_pass(sub, D)
        times.append(dt)
        if time.perf_counter() + dt > t_end:
            break
    return P_s * NIT / float(np.mean(times)), threads, sample_text(P_s, sub.shape[0], sub.shape[1]), times


def run_reference(args, rank):
    """--impl reference: every step is one pass of the reference path's CPU port over the fixed sample; ms_per_step is
    the MEASURED wall time of such a pass (not an extrapolation)."""
    if rank != 0:
        return
    Y, pm, D = make_inputs(args.workload)
    sub = cpu_sample(Y)
    threads = use_all_host_threads()
    times, P_s = [], 0
    for i in range(args.warmup + args.steps):
        P_s, dt = cpu_pass(sub, D)
        if i >= args.warmup:
            times.append(dt)
    mean = float(np.mean(times))
    val = P_s * NIT / mean
    R, C = Y.shape
    P = (R - BB + 1) * (C - BB + 1)
    cores = threads
    sample = sample_text(P_s, sub.shape[0], C)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * mean, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "patches_per_step": P_s, "patches_full_workload": P,
                       "step": "one pass over the fixed CPU sample (sparse-coding step only)",
                       "extrapolated": False,
                       "full_workload_ms_per_step_extrapolated": 1e3 * P * NIT / val},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "pass_seconds_min_mean_max": [float(np.min(times)), mean, float(np.max(times))]},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def workload_name(w):
    H, W, B, kind, keep = WORKLOADS[w]
    return f"{w}: synthetic {H}x{W}x{B} cube, 8x8 patches stride 1, {kind} mask keep={keep}, K={K_ATOMS}, Nit={NIT}"


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- GPU arm
def measure_tf32_peak(torch, seconds=1.5):
    """Dense TF32 matmul throughput on this GPU (torch.matmul 8192^3, TF32 allowed), sustained over
    `seconds` — the denominator of the 3xTF32 roofline."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    n = 8192
    a = torch.randn(n, n, device="cuda")
    b = torch.randn(n, n, device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters, t0 = 0, time.perf_counter()
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(10):
            a @ b
        iters += 10
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    torch.backends.cuda.matmul.allow_tf32 = old
    return 2.0 * n ** 3 * iters / (e0.elapsed_time(e1) * 1e-3) / 1e12


def sharded_parity(torch, dist, lrs, solver, rank, world, dev, engine):
    """Outside the timed region (N > 1): two outer iterations of a small cube ('mini' workload) row-striped over the N
    ranks against the same run unsharded on rank 0's GPU; returns rel-L2 of X on rank 0 (the driver's pytest box has
    one GPU, so tests/test_gpu_multi.py cannot run there)."""
    Y, pm, D = make_inputs("mini")
    R, C = Y.shape
    M = np.repeat(pm.astype(np.float32)[:, None], C, axis=1)
    prm = lrs.Params(Nit=NIT, bb=BB, slidingDis=STRIDE, step="spectral")
    st = solver.make_stripe(R, BB, rank, world)
    sol = lrs.LRSPnP(torch.from_numpy(np.ascontiguousarray(Y[st.row_slice])), torch.from_numpy(np.ascontiguousarray(M[st.row_slice])),
                     torch.from_numpy(D), prm, engine=engine, stripe=st, device=dev)
    sol.run(2)
    own_max = max(solver.make_stripe(R, BB, r, world).rows_owned for r in range(world))
    pad = torch.zeros((own_max, C), dtype=torch.float32, device=dev)
    pad[:st.rows_owned] = sol.X[:st.rows_owned]
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    if rank != 0:
        return None
    Xs = torch.cat([parts[r][:solver.make_stripe(R, BB, r, world).rows_owned] for r in range(world)])
    one = lrs.LRSPnP(torch.from_numpy(Y), torch.from_numpy(M), torch.from_numpy(D), prm, engine=engine, device=dev)
    one.run(2)
    return float((Xs - one.X).norm() / one.X.norm())


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import lrs_pnp_dip_b200 as lrs
    from lrs_pnp_dip_b200 import _lib, solver

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    if args.range_mb:
        solver.SparseCoder.CHUNK_BYTES = args.range_mb << 20
    parity = sharded_parity(torch, dist, lrs, solver, rank, world, dev, args.engine) if world > 1 else None

    Y, pm, D = make_inputs(args.workload)
    R, C = Y.shape
    P_total = (R - BB + 1) * (C - BB + 1)
    prm = lrs.Params(Nit=NIT, bb=BB, slidingDis=STRIDE, step="spectral")
    st = solver.make_stripe(R, BB, rank, world)
    Yl = np.ascontiguousarray(Y[st.row_slice])
    Ml = np.ascontiguousarray(np.repeat(pm[st.row_slice].astype(np.float32)[:, None], C, axis=1))
    P_local = (st.b - st.a) * (C - BB + 1)
    if world > 1:
        del Y          # every rank generated the same seeded cube; keep only the stripe
        Y = None

    sol = lrs.LRSPnP(torch.from_numpy(Yl), torch.from_numpy(Ml), torch.from_numpy(D), prm, engine=args.engine,
                     stripe=st if world > 1 else None, device=dev)
    if args.no_hide:                                 # A/B: eigensolver after the sparse step instead of beside it
        sol.hide_eigensolver = False
    coder = sol.be.coder
    # Time the dominant kernel with CUDA events on the launching stream.  The sparse step launches the fused kernel once
    # per range of column starts (two alternating streams, each range followed by its overlap-sum kernel, so that Phi_z is
    # never materialised); consecutive launches overlap at their tails, so the figure is the SPAN of the whole sequence on
    # the caller's stream — all fused launches of the step plus the last range's overlap sum (an upper bound of the fused
    # kernel's own time; the ncu launch list under profiles/ has the per-launch durations).
    kern_ev = []
    orig_imout = coder.imout

    def timed_imout(X, l1, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig_imout(X, l1, **kw)
        e1.record()
        kern_ev.append((e0, e1))
        return out

    coder.imout = timed_imout

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The reference runs iteration_num = 2 outer iterations from X = Y_observed, λ = 0 (main_LRS_PnP.py:228-229,250).
    # Its update λ1 += μ1·(X − IMout) uses the overlap SUM (:346,361), which at stride 1 (Weight = 64) grows the state
    # ≈63x per outer iteration and overflows fp32 after ≈21 of them, so the bench never iterates further than the
    # reference does: the state is re-initialised every RESET_EVERY steps (three device copies/memsets on the timed
    # stream, inside the timed region), and every timed step is outer iteration 1 or 2 of the reference's own run.
    RESET_EVERY = 2
    n_done = 0

    def one_step():
        nonlocal n_done
        if n_done % RESET_EVERY == 0:
            sol.reset()
        sol.step()
        n_done += 1

    def assert_finite(where):
        if not bool(torch.isfinite(sol.X).all()):
            raise SystemExit(f"bench.py: non-finite ADMM state {where}")

    for _ in range(args.warmup):
        one_step()
    barrier()
    assert_finite("after warm-up")
    n_done = 0
    kern_ev.clear()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = L.lrs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_range:          # ncu --profile-from-start off: capture the timed steps only
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    if args.profile_range:
        torch.cuda.profiler.stop()
    assert_finite("after the timed steps")
    sol.validate()
    launches = int(L.lrs_launch_count() - launches0)
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    kms = torch.tensor([float(np.mean([a.elapsed_time(b) for a, b in kern_ev]))], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kms, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms.item()) / args.steps
    value = P_total * NIT / (ms_per_step * 1e-3)

    # ---- end to end through the host-facing call: host buffers in, host buffers out, every step ----
    h2d = d2h = 0
    e2e_value = None
    if not args.no_e2e:
        pin = lambda a: torch.from_numpy(a).pin_memory()
        hY, hM = pin(Yl), pin(Ml)
        hX, hL1, hL2 = pin(Yl.copy()), pin(np.zeros_like(Yl)), pin(np.zeros_like(Yl))
        oX = torch.empty_like(hX).pin_memory()
        oL1, oL2 = torch.empty_like(hX).pin_memory(), torch.empty_like(hX).pin_memory()
        Dd = torch.from_numpy(D).to(dev)

        def e2e_step():
            nonlocal h2d, d2h
            # host-facing call: pinned host buffers in (observation, mask, full ADMM state), one outer iteration, state out
            s2 = lrs.LRSPnP.from_host(hY, hM, Dd, prm, state=(hX, hL1, hL2), engine=args.engine,
                                      stripe=st if world > 1 else None, device=dev)
            s2.step()
            e2e_solvers.append(s2.be.coder)
            s2.to_host(oX, oL1, oL2)
            h2d = 5 * hY.numel() * 4
            d2h = 3 * hY.numel() * 4

        n_e2e = max(1, min(args.steps, args.e2e_steps))
        e2e_solvers = []
        e2e_step()
        barrier()
        e2e_solvers.clear()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n_e2e):
            e2e_step()
        t1.record()
        barrier()
        for s2 in e2e_solvers:       # deferred device-side input checks of the solvers built inside the timed region
            s2.validate()
        if not np.isfinite(oX.numpy()).all():
            raise SystemExit("bench.py: non-finite end-to-end result")
        e2e_solvers.clear()
        ems = torch.tensor([t0.elapsed_time(t1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        e2e_value = P_total * NIT / (float(ems.item()) / n_e2e * 1e-3)

    if rank == 0:
        pk = peaks()
        tf32 = measure_tf32_peak(torch)
        flops = 4.0 * 64 * K_ATOMS * P_local * NIT + 2.0 * 64 * K_ATOMS * P_local   # ISTA + Phi_z = D alpha
        achieved = flops / (float(kms.item()) * 1e-3) / 1e12
        # The tcgen05 engine emulates fp32 with 3 fp16-piece MMAs per product, so its ceiling is the dense
        # 16-bit tensor rate / 3 (driver-measured cuBLAS bf16, sustained: the kernel runs for seconds under
        # the power cap).  The 3xTF32 ceiling the spec names (measured TF32 / 3) is reported beside it.
        simt = (args.engine == "simt")
        peak = pk["bf16_sust"] / 3.0
        cpu_v, cores, sample = (None, None, None)
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu_v, cores, sample, ptimes = cpu_patch_iters_per_s(Y, D, budget_s=args.cpu_seconds)
            cpu = {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                   "passes": len(ptimes), "pass_seconds_mean": float(np.mean(ptimes))}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32 (fp32 FFMA)" if simt else "f32 via 3-pass fp16 split on tcgen05 (22-bit operands, fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "patches": P_total, "engine": args.engine,
                       "step_constant": "spectral", "parallelism": f"row-stripes x{world}",
                       "state": f"ADMM state re-initialised (X=Y, lambda=0) every {RESET_EVERY} steps inside the timed region: "
                                "every step is outer iteration 1 or 2 of the reference's 2-iteration run "
                                "(main_LRS_PnP.py:228-229); the literal update diverges at stride 1 beyond ~20 iterations",
                       "l2": "inputs exceed L2 (every step streams %.1f GB of Phi_z through two %.2f GB range buffers)"
                             % (64 * P_local * 4 / 1e9, 64 * 4 * coder._chunk_cols() * (coder.R - BB + 1) / 1e9),
                       "fused_launches_per_step": len(coder._ranges(sol.hide_eigensolver)),
                       "schedule": ("Gram first, Jacobi eigensolver on a high-priority stream beside the sparse step "
                                    "(work items claimed dynamically), recomposition after it") if sol.hide_eigensolver
                                   else "sparse step, then the low-rank step"},
            "clocks": clocks,
            "e2e": None if e2e_value is None else {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                                                   "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": ncu_traffic(args.workload, world),
                         "traffic_unit": "bytes per step = all fused launches of one step (ncu dram read+write)",
                         "traffic_source": "stored ncu capture profiles/r02_traffic.json (not measured in this run)",
                         "algorithmic_bytes": 3.0 * 4 * R * C / world + 4.0 * 64 * P_local,
                         "kernel": "sparse_fused_tc_kernel, all column-start-range launches of one step (span on the caller's stream, "
                                   "including the overlapped overlap-sum kernels)", "kernel_ms": float(kms.item()),
                         "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained {pk['bf16_sust']:.0f} ({pk['src']}) / 3 "
                                        "(3 fp16 MMAs per fp32 product; split products not counted as useful flops)",
                         "tf32x3_peak": tf32 / 3.0, "frac_of_tf32x3": achieved / (tf32 / 3.0),
                         "tf32_measured": tf32},
            "cpu_baseline": cpu,
        }
        if parity is not None:
            line["sharded_parity_rel_l2"] = parity
            line["config"]["sharded_parity"] = ("2 outer iterations of the 'mini' cube, row stripes over all ranks vs one GPU, "
                                                "outside the timed region")
            if not parity < 1e-4:
                raise SystemExit(f"bench.py: sharded run differs from the single-GPU run: rel-L2 {parity:.3e}")
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------- cfg 1 (bundled cube)
# BASELINE.json configs[0]: main_LRS_PnP.py as shipped — noisy_img5 + fourth_mask (main_LRS_PnP.py:170,183), 36x36 blocks
# at stride 36 of the 1296 x 128 unfolded cube (144 patches, n = 1296), Nit = 80, spectral step, SVT.  The trained
# dictionary is not part of the checkout: K = 2592 synthetic atoms.  A parity case first (tests/), offered as
# a bench workload so that the explicit tensor-core engine and the literal per-patch CPU loop (B0) have a line too.
CFG1_K, CFG1_NIT, CFG1_BB, CFG1_SAMPLE = 2592, 80, 36, (0, 29, 58, 87, 115, 143)


def cfg1_inputs():
    from lrs_pnp_dip_b200 import synth

    g = np.load(os.path.join(ROOT, "tests", "golden", "bundled_inputs.npz"))
    Y, pm = g["img5_Y"], g["img5_pixmask"]
    return Y, np.repeat(pm.astype(np.float32)[:, None], Y.shape[1], axis=1), synth.synthetic_dictionary(CFG1_BB ** 2, CFG1_K, seed=0)


CFG1_NAME = ("cfg1: bundled noisy_img5 + fourth_mask (1296x128 unfolded), 36x36 blocks stride 36 (144 patches, n=1296), "
             f"K={CFG1_K} synthetic atoms, Nit={CFG1_NIT}, spectral step, SVT")


def cfg1_cpu_pass(Y, D):
    """The literal per-patch loop (oracle/literal_loop.py = main_LRS_PnP.py:265-303: row deletion, one SVD per patch,
    Nit pairs of torch.mm) over a FIXED sample of 6 of the 144 patches."""
    from oracle import literal_loop as ll, lrs_oracle as orc

    blocks, _, _, _ = orc.get_image_block(Y, CFG1_BB, CFG1_BB)
    _, dt = ll.sparse_step_literal(blocks, blocks, D, 0.1, CFG1_NIT, "spectral", patches=np.array(CFG1_SAMPLE))
    return len(CFG1_SAMPLE), dt


def run_cfg1_reference(args, rank):
    if rank != 0:
        return
    Y, _, D = cfg1_inputs()
    threads = use_all_host_threads()
    times = []
    for i in range(args.warmup + args.steps):
        P_s, dt = cfg1_cpu_pass(Y, D)
        if i >= args.warmup:
            times.append(dt)
    mean = float(np.mean(times))
    val = P_s * CFG1_NIT / mean
    sample = f"{P_s} of the 144 patches (indices {list(CFG1_SAMPLE)}) x {CFG1_NIT} iterations per pass, literal per-patch loop with one SVD per patch"
    emit({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": 1e3 * mean, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "f32", "data": "bundled cube, synthetic dictionary",
                      "config": {"workload": CFG1_NAME, "patches_per_step": P_s, "patches_full_workload": 144, "extrapolated": False,
                                 "full_workload_ms_per_step_extrapolated": 1e3 * 144 * CFG1_NIT / val},
                      "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
                      "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


def run_cfg1_ours(args):
    import torch

    import lrs_pnp_dip_b200 as lrs
    from lrs_pnp_dip_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    L = _lib.lib()
    Y, M, D = cfg1_inputs()
    prm = lrs.Params()                                            # main_LRS_PnP.py:218-238
    sol = lrs.LRSPnP(torch.from_numpy(Y), torch.from_numpy(M), torch.from_numpy(D), prm, engine=args.engine, device=dev)
    if args.no_overlap:                                           # A/B: low-rank step after the sparse step, one stream
        sol.overlap_low_rank = False
    coder, P = sol.be.coder, sol.be.coder.P
    kern_ev, orig = [], coder.phi_z

    def timed_phi(X, l1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(X, l1)
        e1.record()
        kern_ev.append((e0, e1))
        return out

    coder.phi_z = timed_phi
    n_done = 0

    def one_step():
        nonlocal n_done
        if n_done % 2 == 0:
            sol.reset()                                           # the reference runs 2 outer iterations (:228)
        sol.step()
        n_done += 1

    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    n_done = 0
    kern_ev.clear()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = L.lrs_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(args.steps, 1)
    e0.record()
    for _ in range(reps):
        one_step()
    e1.record()
    torch.cuda.synchronize()
    launches = int(L.lrs_launch_count() - launches0)
    clocks = sampler.stop()
    if not bool(torch.isfinite(sol.X).all()):
        raise SystemExit("bench.py: non-finite ADMM state")
    ms_per_step = e0.elapsed_time(e1) / reps
    kms = float(np.mean([a.elapsed_time(b) for a, b in kern_ev]))
    # end to end: host buffers in, a fresh solver, one outer iteration, X back on the host
    pin = lambda a: torch.from_numpy(a).pin_memory()
    hY, hM, Dd = pin(Y), pin(M), torch.from_numpy(D).to(dev)
    oX = torch.empty_like(hY).pin_memory()

    def e2e_step():
        s2 = lrs.LRSPnP.from_host(hY, hM, Dd, prm, engine=args.engine, device=dev)
        s2.overlap_low_rank = sol.overlap_low_rank
        s2.step()
        s2.to_host(oX)

    e2e_step()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps):
        e2e_step()
    t1.record()
    torch.cuda.synchronize()
    e2e_value = P * CFG1_NIT / (t0.elapsed_time(t1) / reps * 1e-3)
    pk = peaks()
    n = CFG1_BB ** 2
    flops = (4.0 * n * CFG1_K * CFG1_NIT + 2.0 * n * CFG1_K) * P
    achieved = flops / (kms * 1e-3) / 1e12
    cpu = None
    if not args.no_cpu:
        threads = use_all_host_threads()
        cfg1_cpu_pass(Y, D)
        P_s, dt = cfg1_cpu_pass(Y, D)
        cpu = {"value": P_s * CFG1_NIT / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{P_s} of the 144 patches x {CFG1_NIT} iterations, literal per-patch loop with one SVD per patch (B0)"}
    emit({
        "metric": METRIC, "value": P * CFG1_NIT / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": reps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32 via 3-pass fp16 split on tcgen05 (22-bit operands, fp32 accumulate)", "data": "bundled cube, synthetic dictionary",
        "config": {"workload": CFG1_NAME, "patches": P, "engine": args.engine,
                   "state": "re-initialised every 2 steps (the reference's own 2-iteration run)",
                   "schedule": "SVT on a second stream beside the sparse step" if sol.overlap_low_rank else "sequential",
                   "e2e_note": "every e2e step builds a fresh solver from host buffers; the spectral step constants of its mask "
                               "patterns are memoised per (dictionary tensor, pattern) by ops.step_constants, so only the first "
                               "construction (the untimed warm-up) pays the 1296 x 1296 eigensolves",
                   "l2": "working set (27 MB of dictionary pieces) is L2-resident by design: the configuration is latency-bound"},
        "clocks": clocks, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * hY.numel() * 4, "d2h_bytes_per_step": hY.numel() * 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"] / 3.0, "unit": "TFLOP/s", "frac": achieved / (pk["bf16"] / 3.0),
                     "traffic": None, "kernel": "sparse step = im2col + (2 Nit + 1) split-K tcgen05 GEMM/reduce launches", "kernel_ms": kms,
                     "peak_source": f"MEASURED_PEAKS.json bf16_tflops {pk['bf16']:.0f} ({pk['src']}) / 3; informative only — "
                                    "144 patches cannot fill the tensor pipe, the path is launch/latency bound"},
        "cpu_baseline": cpu})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS) + ["cfg1"])
    ap.add_argument("--engine", default="auto", choices=["auto", "simt", "tc"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-range", action="store_true", help="cudaProfilerStart/Stop around the timed steps")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="cfg1: low-rank step after the sparse step instead of beside it")
    ap.add_argument("--no-hide", action="store_true", help="eigensolver of the SVT after the sparse step instead of beside it")
    ap.add_argument("--range-mb", type=int, default=None,
                    help="experiment: size (MiB) of each of the two Phi_z range buffers (default: SparseCoder.CHUNK_BYTES)")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    args = ap.parse_args()
    claim_stdout()                                  # stdout carries the one JSON line and nothing else
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "cfg1":
        if args.impl == "reference":
            run_cfg1_reference(args, rank)
        elif rank == 0:
            run_cfg1_ours(args)                     # 144 patches: replicas only, rank 0 reports
        return
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
