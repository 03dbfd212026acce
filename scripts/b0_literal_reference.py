"""B0 of BASELINE.md: the reference's OWN functions (AST-extracted from /root/reference, oracle/ref_extract.py) in the
literal per-patch loop of main_LRS_PnP.py:259-303 on configuration 1 AS SHIPPED (noisy_img5 + fourth_mask, bb = 36,
stride 36, Nit = 80, 144 patches) in full, timed on this container's host cores.  Runs only where the checkout exists
(the build container); the port that travels to the GPU box is oracle/literal_loop.py, timed by `bench.py --workload cfg1`.

    python scripts/b0_literal_reference.py [K=2592] > profiles/r02_b0_literal_cfg1.json
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lrs_pnp_dip_b200 import synth  # noqa: E402
from oracle import literal_loop as ll, ref_extract as rx  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 2592
g = np.load(os.path.join(ROOT, "tests", "golden", "bundled_inputs.npz"))
Y = torch.tensor(g["img5_Y"])
D = torch.tensor(synth.synthetic_dictionary(1296, K, seed=0))
ns = rx.extract("main_LRS_PnP.py")
ns["denoise_nl_means"] = rx.soft_shim(10.0)
bb = sd = 36
t0 = time.perf_counter()
blocks_copy, rows, cols, _ = ns["get_image_block"](Y, bb, sd)                 # :244
blocks, rows, cols, _ = ns["get_image_block"](Y, bb, sd)                      # :259 (X = Y_observed, lambda_1 = 0)
t_im2col = time.perf_counter() - t0
Phi_z = torch.zeros(blocks.size())
t0 = time.perf_counter()
for jj in range(Phi_z.size()[1]):                                              # :270-303
    pruned, valid = D, blocks[:, jj].view((bb ** 2, 1))
    missing = np.where(blocks_copy[:, jj].view((bb ** 2, 1)).flatten() == 0)[0]
    if len(missing) > 0:
        valid = ns["delete_element"](valid, torch.Tensor(missing).tolist())
        pruned = ns["delete_element"](pruned, torch.Tensor(missing).tolist())
    Phi_z[:, jj] = torch.mm(D, ns["ista"](valid, pruned, 0.1, 0, 80)).flatten()
t_loop = time.perf_counter() - t0
port, t_port = ll.sparse_step_literal(blocks.numpy(), blocks_copy.numpy(), D.numpy(), 0.1, 80, "spectral")
err = float(np.linalg.norm(port - Phi_z.numpy()) / np.linalg.norm(Phi_z.numpy()))
P = Phi_z.size()[1]
print(json.dumps({"what": "literal reference loop main_LRS_PnP.py:259-303 (AST-extracted functions), cfg 1 as shipped, full",
                  "patches": P, "n": bb * bb, "K": K, "Nit": 80, "cores": os.cpu_count(), "torch_threads": torch.get_num_threads(),
                  "get_image_block_s_two_calls": t_im2col, "patch_loop_s": t_loop, "patch_iters_per_s": P * 80 / t_loop,
                  "port_oracle_literal_loop_s": t_port, "port_patch_iters_per_s": P * 80 / t_port, "port_vs_literal_rel_l2": err}))
