"""Substitute for the missing ``trained_dictionary.mat`` (main_LRS_PnP.py:159-165 and main_LRS_PnP.m:7 load a
``Dictionary`` the repository does not ship).

Method of optimal directions with the library's own sparse coder: alternate

    A  = argmin_a  ||y - D a||² + lambda |a|_1      (Nit soft-ISTA iterations on the device, ista.m:13-24)
    D  = Y Aᵀ (A Aᵀ + eps I)⁻¹ ,  columns l2-normalised (columnNormalise.m:1-4)

on the bb x bb patches of a training matrix (the unfolded clean cube).  Atoms that no patch uses are re-seeded
with the currently worst-represented patches.  The learner is a caller of the hot path, not part of it: the
patch gather and the ISTA loop are ``lrs_im2col_f32`` / ``lrs_ista_pnp_f32``; the K x K solve is torch.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib, ops


def column_normalise(A: torch.Tensor) -> torch.Tensor:
    """columnNormalise.m:1-4 — every column divided by its l2 norm (zero columns are left as they are)."""
    nrm = torch.linalg.vector_norm(A, dim=0, keepdim=True)
    return A / torch.where(nrm > 0, nrm, torch.ones_like(nrm))


def training_patches(Y: torch.Tensor, bb: int, s: int, max_patches: Optional[int] = None, seed: int = 0) -> torch.Tensor:
    """bb² x P patch matrix of ``Y`` in the reference's patch order (get_image_block, main_LRS_PnP.py:73-107),
    optionally a seeded random subset of the columns."""
    blocks = ops.im2col(Y, bb, s)
    if max_patches is not None and blocks.shape[1] > max_patches:
        g = torch.Generator(device="cpu").manual_seed(seed)
        keep = torch.randperm(blocks.shape[1], generator=g)[:max_patches].sort().values.to(blocks.device)
        blocks = blocks[:, keep].contiguous()
    return blocks


def learn_dictionary(patches: torch.Tensor, K: int, lambda_ista: float = 0.1, Nit: int = 40, rounds: int = 10,
                     eps: float = 1e-6, seed: int = 0, D0: Optional[torch.Tensor] = None,
                     step: str = "spectral") -> Tuple[torch.Tensor, List[float]]:
    """``patches`` [n, P] on a CUDA device -> (D [n, K] with unit columns, relative representation error
    ||Y - D A||_F / ||Y||_F after every round)."""
    _lib.require_cuda()
    if patches.device.type != "cuda":
        raise _lib.LrsError("learn_dictionary needs a CUDA tensor")
    Y = patches.contiguous().float()
    n, P = Y.shape
    if P < K:
        raise ValueError(f"need at least K = {K} training patches, got {P}")
    with torch.cuda.device(Y.device):
        g = torch.Generator(device="cpu").manual_seed(seed)
        if D0 is None:
            pick = torch.randperm(P, generator=g)[:K].to(Y.device)
            D = Y[:, pick] + 1e-3 * torch.randn((n, K), generator=g).to(Y.device)
        else:
            D = D0.to(Y.device).float()
        D = column_normalise(D).contiguous()
        observed = torch.ones_like(Y)                    # training patches are complete: blocks_copy != 0 everywhere
        ynorm = float(torch.linalg.matrix_norm(Y))
        history: List[float] = []
        eye = torch.eye(K, device=Y.device)
        for _ in range(rounds):
            a = ops.step_constants(observed, D, step)
            A, _ = ops.ista_batched(Y, observed, D, a, lambda_ista, Nit, want_coefs=True, want_phi=False)
            G = A @ A.T
            D = torch.linalg.solve(G + eps * max(float(G.diagonal().mean()), 1.0) * eye, A @ Y.T).T
            used = torch.linalg.vector_norm(A, dim=1) > 0
            if not bool(used.all()):                     # re-seed dead atoms with the worst-represented patches
                err = torch.linalg.vector_norm(Y - D @ A, dim=0)
                worst = torch.topk(err, int((~used).sum())).indices
                D[:, ~used] = Y[:, worst]
            D = column_normalise(D).contiguous()
            history.append(float(torch.linalg.matrix_norm(Y - D @ A)) / max(ynorm, 1e-30))
        return D, history


def save_dictionary(path: str, D) -> None:
    """Write ``Dictionary`` the way the reference loads it (scipy v5 .mat, main_LRS_PnP.py:163-165)."""
    from scipy.io import savemat

    savemat(path, {"Dictionary": np.asarray(D.detach().cpu() if isinstance(D, torch.Tensor) else D, dtype=np.float32)})
