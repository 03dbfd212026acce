"""Quality metrics of the reference scripts, as torch ops on whatever device the cubes live on
(no CPU round trip in the outer loop).  Image tensors are ``[1, bands, d2, d3]``.

* ``psnr_ref`` / ``mpsnr``: the reference's NON-standard ``10*log10(255/sqrt(mse))`` on [0,1] data
  (main_LRS_PnP.py:40-58, per band :379-384) — reproduced verbatim so that "PSNR within 0.01 dB" compares
  like with like.
* ``ssim``: Gaussian-window SSIM, window 11, sigma 1.5, C1 = 0.01², C2 = 0.03², zero-padded depthwise
  convolution, mean over the whole map (pytorch_ssim/__init__.py:7-37,65-73).
* ``state_convergence``: log of the 2-norm of the state change (main_LRS_PnP.py:23-25).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def psnr_ref(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = float(torch.mean((a.float() - b.float()) ** 2))
    if mse < 1.0e-10:
        return 100.0
    return 10 * math.log10(255 / math.sqrt(mse))


def mpsnr(clean: torch.Tensor, pred: torch.Tensor) -> float:
    """bach_mpsnr (main_LRS_PnP.py:48-58): mean over bands (and batch) of ``psnr_ref``."""
    mse = ((clean.float() - pred.float()) ** 2).mean(dim=(2, 3))             # [batch, bands]
    p = torch.where(mse < 1.0e-10, torch.full_like(mse, 100.0), 10 * torch.log10(255 / torch.sqrt(mse.clamp_min(1e-30))))
    return float(p.mean(dim=1).mean())


def _window(size: int, sigma: float, channels: int, device, dtype) -> torch.Tensor:
    x = torch.arange(size, dtype=torch.float64)
    g = torch.exp(-((x - size // 2) ** 2) / (2 * sigma ** 2))
    g = (g / g.sum()).to(torch.float32)
    w2 = (g[:, None] @ g[None, :]).to(device=device, dtype=dtype)
    return w2.expand(channels, 1, size, size).contiguous()


def ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11) -> float:
    img1, img2 = img1.float(), img2.float()
    ch = img1.shape[1]
    w = _window(window_size, 1.5, ch, img1.device, img1.dtype)
    pad = window_size // 2
    conv = lambda t: F.conv2d(t, w, padding=pad, groups=ch)  # noqa: E731
    mu1, mu2 = conv(img1), conv(img2)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = conv(img1 * img1) - mu1_sq
    s2 = conv(img2 * img2) - mu2_sq
    s12 = conv(img1 * img2) - mu12
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    m = ((2 * mu12 + c1) * (2 * s12 + c2)) / ((mu1_sq + mu2_sq + c1) * (s1 + s2 + c2))
    return float(m.mean())


def state_convergence(current: torch.Tensor, previous: torch.Tensor) -> float:
    return float(torch.log(torch.norm(current - previous, p=2)))


def fold(Y: torch.Tensor, d2: int, d3: int) -> torch.Tensor:
    """Unfolded ``[d3*d2, B]`` → image ``[1, B, d2, d3]`` on the device (inverse of main_LRS_PnP.py:209; the
    layout shuffles of main_LRS_PnP_DIP_pro.py:412,419 are this and :func:`unfold`)."""
    B = Y.shape[1]
    return Y.reshape(d3, d2, B).permute(2, 1, 0).reshape(1, B, d2, d3).contiguous()


def unfold(cube: torch.Tensor) -> torch.Tensor:
    _, B, d2, d3 = cube.shape
    return cube.reshape(B, d2, d3).permute(2, 1, 0).reshape(d3 * d2, B).contiguous()
