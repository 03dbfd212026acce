"""TEST INFRASTRUCTURE — not part of the product path.

Loads the reference's OWN function definitions out of its scripts, without
running the scripts' top-level code (which does file I/O on absolute paths and
imports h5py / skimage / matplotlib, none of which exist here).

The scripts under ``/root/reference`` are parsed with ``ast``; only the
``FunctionDef`` nodes named below are compiled and executed in a fresh
namespace that provides ``np``, ``torch`` and ``math``.  Nothing is copied
into this repository: the source stays where it lies and is only available in
the build container (``/root/reference`` does not exist on the GPU box), so
this module is used exclusively by ``tests/golden/make_golden.py`` (fixture
generation) and by CPU tests that skip when the reference is absent.

Extracted symbols (reference file:line):
  main_LRS_PnP.py         get_image_block :73, Shrinkage_Operator :112, SVT :118,
                          soft_thresh :128, ista :131 (spectral step),
                          delete_element :152, psnr :40, bach_mpsnr :48,
                          state_convergence :23
  main_LRS_PnP_DIP_pro.py ista :188 (4*||H||_F^2 step)
  admm_utils.py           l1_prox :72

``ista`` calls the global ``denoise_nl_means`` (skimage).  The caller injects
a replacement through ``namespace['denoise_nl_means']`` — identity, or the
soft-threshold of the MATLAB twin (ista.m:23) — which is how the gradient step
and the step constants of the literal function are pinned.
"""
from __future__ import annotations

import ast
import math
import os
from typing import Dict, Iterable

REFERENCE_ROOT = os.environ.get("LRS_REFERENCE_ROOT", "/root/reference")

_WANTED = {
    "main_LRS_PnP.py": (
        "state_convergence", "psnr", "bach_mpsnr", "get_image_block", "Shrinkage_Operator",
        "SVT", "soft_thresh", "ista", "delete_element",
    ),
    "main_LRS_PnP_DIP_pro.py": ("ista", "get_image_block", "delete_element"),
    "main_LRS_PnP_DIP_1-LiP.py": ("ista",),
    "admm_utils.py": ("l1_prox",),
}


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "main_LRS_PnP.py"))


def extract(script: str, names: Iterable[str] | None = None) -> Dict[str, object]:
    """Return ``{name: function}`` for the requested top-level defs of
    ``/root/reference/<script>``; the dict is also the functions' globals, so
    ``ns['denoise_nl_means'] = f`` rebinds the denoiser ``ista`` sees."""
    import numpy as np
    import torch

    path = os.path.join(REFERENCE_ROOT, script)
    import warnings

    with open(path, "r", encoding="utf-8", errors="replace") as fh, warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)
        tree = ast.parse(fh.read(), filename=path)
    want = set(names if names is not None else _WANTED[script])
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    missing = want - {n.name for n in body}
    if missing:
        raise KeyError(f"{script}: no top-level def for {sorted(missing)}")
    mod = ast.Module(body=body, type_ignores=[])
    ns: Dict[str, object] = {"np": np, "torch": torch, "math": math, "__name__": f"_ref_{script}"}
    exec(compile(mod, path, "exec"), ns)
    return ns


def soft_shim(scale: float):
    """Stand-in for ``denoise_nl_means(g, h=..., fast_mode=True, patch_size=3,
    patch_distance=3)`` that applies the MATLAB twin's ``soft(g, T)``
    (ista.m:23, soft.m:4).  The literal ``ista`` passes ``h = 0.1*T``
    (main_LRS_PnP.py:146) or ``h = T`` (main_LRS_PnP_DIP_pro.py:199);
    ``scale`` undoes that factor so the threshold is T."""
    import numpy as np

    def shim(g, h, fast_mode=True, **kw):
        g = np.asarray(g, dtype=np.float32)
        thr = np.float32(float(h) * scale)
        return (np.sign(g) * np.maximum(np.abs(g) - thr, np.float32(0))).astype(np.float32)

    return shim


def identity_shim(g, h=None, fast_mode=True, **kw):
    import numpy as np

    return np.asarray(g, dtype=np.float32)
