"""TEST / BASELINE INFRASTRUCTURE — the reference's per-patch sparse-coding loop restated one patch at a time, the way
main_LRS_PnP.py:265-303 runs it (B0 of BASELINE.md): a Python ``for`` over the patches, row deletion of the missing
entries, ONE SVD PER PATCH for the step constant (``np.linalg.norm(H, 2)**2``, main_LRS_PnP.py:134), ``Nit`` pairs of
``torch.mm`` mat-vecs on CPU tensors (:138) and the full-dictionary reconstruction (:294/:302).  The denoiser is the
MATLAB twin's soft threshold (ista.m:23), as everywhere in this repository.

Only ``bench.py``'s CPU legs and ``tests/`` import this module.  It is pinned against the literal (AST-extracted)
reference functions in ``tests/test_oracle.py::test_literal_loop_port_matches_reference``.
"""
from __future__ import annotations

import time
from typing import Optional, Tuple

import numpy as np
import torch


def ista_one_patch(y: torch.Tensor, H: torch.Tensor, lambda_ista: float, Nit: int, step: str = "spectral") -> torch.Tensor:
    """main_LRS_PnP.py:131-149 (step='spectral') / main_LRS_PnP_DIP_pro.py:188-201 ('frob4'), soft denoiser."""
    x = torch.zeros((H.shape[1], 1))
    if step == "spectral":
        alpha = float(np.linalg.norm(H.numpy(), 2) ** 2)                       # :134 — an SVD of the pruned dictionary
    else:
        alpha = float(2 * (np.trace(torch.mm(H.T, H).numpy()) + np.trace(torch.mm(H.T, H).numpy())))   # DIP_pro.py:190
    T = lambda_ista / (2 * alpha)
    for _ in range(Nit):
        g = x + torch.mm(H.T, (y - torch.mm(H, x))) / alpha                    # :138
        x = torch.sign(g) * torch.clamp(g.abs() - T, min=0)                    # soft(g, T), ista.m:23
    return x


def sparse_step_literal(blocks: np.ndarray, blocks_copy: np.ndarray, D: np.ndarray, lambda_ista: float, Nit: int,
                        step: str = "spectral", patches: Optional[np.ndarray] = None) -> Tuple[np.ndarray, float]:
    """The jj loop of main_LRS_PnP.py:270-303 over ``patches`` (default: all).  Returns (Phi_z columns of those patches,
    seconds)."""
    Dt = torch.from_numpy(np.ascontiguousarray(D, dtype=np.float32))
    idx = np.arange(blocks.shape[1]) if patches is None else np.asarray(patches)
    out = np.zeros((D.shape[0], len(idx)), dtype=np.float32)
    t0 = time.perf_counter()
    for o, jj in enumerate(idx):
        keep = torch.from_numpy(blocks_copy[:, jj] != 0)                        # :276-280
        y = torch.from_numpy(np.ascontiguousarray(blocks[:, jj], dtype=np.float32)).view(-1, 1)
        if bool(keep.all()):
            coefs = ista_one_patch(y, Dt, lambda_ista, Nit, step)               # :300
        else:
            coefs = ista_one_patch(y[keep], Dt[keep], lambda_ista, Nit, step)   # :288-292 (delete_element on y and D)
        out[:, o] = torch.mm(Dt, coefs).flatten().numpy()                       # :294 / :302
    return out, time.perf_counter() - t0
