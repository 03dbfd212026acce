"""TEST INFRASTRUCTURE — CPU restatement (NumPy) of the LRS-PnP sparse-coding
hot path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the
product package never does.

Parity status: PINNED.  Every function below is checked in
``tests/test_oracle.py`` and ``tests/test_configs_parity.py`` against the
reference's own functions executed literally (``oracle/ref_extract.py``) when
``/root/reference`` is present, and against the committed fixtures those
functions produced (``tests/golden/*.npz``, generators
``tests/golden/make_golden.py`` and ``make_golden_r2.py``) everywhere else.  The one unpinned piece of the reference is the skimage NLM denoiser
(third-party, absent, version unpinned — SURVEY §8c); the denoiser restated
here is the MATLAB twin's soft threshold (ista.m:23, soft.m:4), which is the
update BASELINE.json's north_star names.

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------
# patch geometry  (main_LRS_PnP.py:73-107)
# --------------------------------------------------------------------------
def axis_starts(length: int, bb: int, s: int, append_last: Optional[bool] = None) -> np.ndarray:
    """Start offsets selected along one axis of the unfolded matrix.

    main_LRS_PnP.py:76-91: ``idx_Mat[0:row+1:s]`` with ``row = length-bb+1``
    marks 0, s, 2s, ... <= length-bb; the last start ``length-bb`` is added iff
    ``length % bb != 0`` (the test is on ``bb``, not on the stride — :83,:86).
    """
    n_pos = length - bb + 1
    if n_pos <= 0:
        raise ValueError("block larger than input")
    st = list(range(0, n_pos, s))
    if append_last is None:
        append_last = (length % bb) != 0
    if append_last and st[-1] != n_pos - 1:
        st.append(n_pos - 1)
    return np.asarray(st, dtype=np.int64)


def patch_index(R: int, C: int, bb: int, s: int) -> Tuple[np.ndarray, np.ndarray]:
    """(x_index, y_index) in the reference's order: ``argwhere`` over the
    column-major flattening of idx_Mat (main_LRS_PnP.py:94-99) ⇒ the row
    start varies fastest, the column start slowest."""
    rs = axis_starts(R, bb, s)
    cs = axis_starts(C, bb, s)
    x = np.tile(rs, len(cs))
    y = np.repeat(cs, len(rs))
    return x.astype(np.int64), y.astype(np.int64)


def idx_mat(R: int, C: int, bb: int, s: int) -> np.ndarray:
    """idx_Mat of main_LRS_PnP.py:76-91 (f32 zeros/ones)."""
    m = np.zeros((R - bb + 1, C - bb + 1), dtype=F32)
    rs = axis_starts(R, bb, s)
    cs = axis_starts(C, bb, s)
    m[np.ix_(rs, cs)] = 1
    return m


def get_image_block(X: np.ndarray, bb: int, s: int):
    """im2col of the unfolded matrix (main_LRS_PnP.py:73-107).
    ``blocks[k, p] = X[x[p] + k % bb, y[p] + k // bb]`` (column-major flatten
    of each bb×bb window, :103-105)."""
    X = np.asarray(X)
    R, C = X.shape
    x, y = patch_index(R, C, bb, s)
    k = np.arange(bb * bb)
    di = (k % bb)[:, None]
    dj = (k // bb)[:, None]
    blocks = X[x[None, :] + di, y[None, :] + dj].astype(F32)
    return blocks, x, y, idx_mat(R, C, bb, s)


def coverage_weight(R: int, C: int, bb: int, s: int) -> np.ndarray:
    """``Weight`` of main_LRS_PnP.py:332-341: how many selected patches cover
    each element (``+= torch.ones(bb)`` broadcasts to +1 over the window)."""
    rs = axis_starts(R, bb, s)
    cs = axis_starts(C, bb, s)
    wr = np.zeros(R, dtype=np.int64)
    wc = np.zeros(C, dtype=np.int64)
    for r in rs:
        wr[r:r + bb] += 1
    for c in cs:
        wc[c:c + bb] += 1
    return (wr[:, None] * wc[None, :]).astype(F32)


def col2im_accumulate(blocks: np.ndarray, R: int, C: int, bb: int, s: int) -> np.ndarray:
    """Overlap *sum* (not average) of patches, main_LRS_PnP.py:332-339:
    ``IMout[r:r+bb, c:c+bb] += reshape(blocks[:, p], (bb,bb)).T`` for
    p = 0..P-1 in order, float32 adds in that order."""
    x, y = patch_index(R, C, bb, s)
    out = np.zeros((R, C), dtype=F32)
    blocks = np.asarray(blocks, dtype=F32)
    P = len(x)
    # chunked np.add.at keeps the per-element summation in ascending patch order
    k = np.arange(bb * bb)
    di = k % bb
    dj = k // bb
    step = max(1, (1 << 22) // (bb * bb))
    for p0 in range(0, P, step):
        p1 = min(P, p0 + step)
        rr = (x[p0:p1, None] + di[None, :]).reshape(-1)
        cc = (y[p0:p1, None] + dj[None, :]).reshape(-1)
        np.add.at(out, (rr, cc), blocks[:, p0:p1].T.reshape(-1))
    return out


# --------------------------------------------------------------------------
# elementwise proximal operators
# --------------------------------------------------------------------------
def soft(x: np.ndarray, tau) -> np.ndarray:
    """soft.m:4 / soft_thresh main_LRS_PnP.py:128-129 / Shrinkage_Operator
    :112-116 / l1_prox admm_utils.py:72-75: sign(x)*max(|x|-tau, 0)."""
    x = np.asarray(x)
    tau = np.asarray(tau, dtype=x.dtype)
    return (np.sign(x) * np.maximum(np.abs(x) - tau, 0)).astype(x.dtype)


def svt(X: np.ndarray, tau: float) -> np.ndarray:
    """SVT main_LRS_PnP.py:118-124: thin SVD, soft-threshold the singular
    values, recompose (float32 LAPACK like the reference)."""
    X = np.asarray(X, dtype=F32)
    U, S, Vt = np.linalg.svd(X, full_matrices=False)
    return (U @ soft(np.diag(S), F32(tau)) @ Vt).astype(F32)


# --------------------------------------------------------------------------
# ISTA  (ista.m:13-24 ; main_LRS_PnP.py:131-149 ; main_LRS_PnP_DIP_pro.py:188-201)
# --------------------------------------------------------------------------
def step_constant(H: np.ndarray, mode: str) -> float:
    """'spectral': ``np.linalg.norm(H,2)**2`` (main_LRS_PnP.py:134, ista.m:15).
    'frob4': ``2*(tr(HᵀH)+tr(HᵀH))`` = 4‖H‖_F² (main_LRS_PnP_DIP_pro.py:190)."""
    if H.shape[0] == 0:
        return 0.0
    if mode == "spectral":
        return float(np.linalg.norm(H, 2) ** 2)
    if mode == "frob4":
        return float(4.0 * np.sum(H.astype(np.float64) ** 2)) if H.dtype == np.float64 else float(
            F32(2) * (np.trace(H.T @ H) + np.trace(H.T @ H)))
    raise ValueError(mode)


def ista_soft(y: np.ndarray, H: np.ndarray, lambda_ista: float, Nit: int, mode: str = "spectral",
              a: Optional[float] = None) -> np.ndarray:
    """One patch, literal restatement of ista.m:13-24 (x0 = 0; a = norm(H)^2;
    T = lambda/(2a); repeat: g = x + Hᵀ(y − Hx)/a ; x = soft(g, T))."""
    dt = H.dtype
    x = np.zeros((H.shape[1], 1), dtype=dt)
    if a is None:
        a = step_constant(H, mode)
    a = dt.type(a)
    T = dt.type(lambda_ista) / (dt.type(2) * a)
    y = y.reshape(-1, 1).astype(dt)
    for _ in range(Nit):
        g = x + (H.T @ (y - H @ x)) / a
        x = soft(g, T)
    return x


def patch_masks(blocks_copy: np.ndarray) -> np.ndarray:
    """Per-patch validity, main_LRS_PnP.py:276-280: an entry is missing iff the
    OBSERVED value is exactly 0.0."""
    return np.asarray(blocks_copy) != 0


def step_constants_batched(D: np.ndarray, mask: np.ndarray, mode: str) -> np.ndarray:
    """a[p] for every patch; distinct mask columns are de-duplicated because
    the spectral norm is an SVD (main_LRS_PnP.py:134 pays it per patch)."""
    n, P = mask.shape
    packed = np.packbits(mask, axis=0)
    uniq, inv = np.unique(packed.T, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    vals = np.zeros(len(uniq), dtype=np.float64)
    for u in range(len(uniq)):
        m = np.unpackbits(uniq[u], count=n).astype(bool)
        vals[u] = step_constant(D[m], mode)
    return vals[inv].astype(D.dtype)


def ista_soft_batched(blocks: np.ndarray, mask: np.ndarray, D: np.ndarray, a: np.ndarray,
                      lambda_ista: float, Nit: int) -> np.ndarray:
    """All patches at once in masked form.  Deleting the missing rows of y and D
    (delete_element, main_LRS_PnP.py:152-155,288-289) equals multiplying the
    residual by the 0/1 mask: α ← soft(α + Dᵀ(M⊙(y − Dα))/a, λ/2a).
    Patches with no valid entry are undefined in the reference (norm of a 0×K
    matrix raises, or a = 0 → NaN); here their coefficients stay 0."""
    dt = D.dtype
    n, P = blocks.shape
    K = D.shape[1]
    m = mask.astype(dt)
    ok = a > 0
    inv_a = np.where(ok, 1.0 / np.where(ok, a, 1), 0).astype(dt)
    T = (dt.type(lambda_ista) * inv_a / dt.type(2)).astype(dt)
    Y = blocks.astype(dt)
    A = np.zeros((K, P), dtype=dt)
    for _ in range(Nit):
        Rm = m * (Y - D @ A)
        G = A + (D.T @ Rm) * inv_a[None, :]
        A = soft(G, T[None, :])
    A[:, ~ok] = 0
    return A


# --------------------------------------------------------------------------
# the ADMM outer iteration (main_LRS_PnP.py:250-362)
# --------------------------------------------------------------------------
@dataclass
class Params:
    """Appendix C of SURVEY.md.  Defaults = main_LRS_PnP.py:218-238."""
    gamma: float = 0.5
    mu_1: float = 0.15
    mu_2: float = 0.15 * 6
    lambda_ista: float = 0.1
    Nit: int = 80
    bb: int = 36
    slidingDis: int = 36
    step: str = "spectral"      # 'frob4' for the DIP variants


@dataclass
class State:
    X: np.ndarray
    lambda_1: np.ndarray
    lambda_2: np.ndarray
    history: List[Dict[str, float]] = field(default_factory=list)


def sparse_step(X: np.ndarray, lambda_1: np.ndarray, Y_observed: np.ndarray, D: np.ndarray, prm: Params,
                a: Optional[np.ndarray] = None, dtype=F32) -> Tuple[np.ndarray, np.ndarray]:
    """main_LRS_PnP.py:259-303 → Phi_z [n,P] (full-dictionary reconstruction of
    every patch, :294/:302) and the step constants used."""
    blocks_copy, _, _, _ = get_image_block(Y_observed, prm.bb, prm.slidingDis)   # :244
    mask = patch_masks(blocks_copy)
    V = (X.astype(F32) + lambda_1.astype(F32) / F32(prm.mu_1)).astype(F32)       # :259
    blocks, _, _, _ = get_image_block(V, prm.bb, prm.slidingDis)
    Dd = D.astype(dtype)
    if a is None:
        a = step_constants_batched(Dd, mask, prm.step)
    A = ista_soft_batched(blocks.astype(dtype), mask, Dd, a.astype(dtype), prm.lambda_ista, prm.Nit)
    return (Dd @ A).astype(F32), a


def lambda1_summation(lambda_1: np.ndarray, R: int, C: int, bb: int, s: int) -> np.ndarray:
    """main_LRS_PnP.py:328,343: every covering patch adds λ1 once more, as
    sequential float32 adds (not Weight*λ1)."""
    W = coverage_weight(R, C, bb, s).astype(np.int64)
    out = np.zeros((R, C), dtype=F32)
    l1 = lambda_1.astype(F32)
    for t in range(int(W.max())):
        out = np.where(W > t, out + l1, out).astype(F32)
    return out


def admm_update(X, lambda_1, lambda_2, Y_observed, MtM, IMout, Weight, U, lam1sum, prm: Params):
    """main_LRS_PnP.py:346 and :361-362, float32, same operation order."""
    g, m1, m2 = F32(prm.gamma), F32(prm.mu_1), F32(prm.mu_2)
    num = (((g * Y_observed + m1 * IMout) + m2 * U) - lam1sum) - lambda_2
    den = (g * MtM + m1 * Weight) + m2
    Xn = (num / den).astype(F32)
    l1 = (lambda_1 + m1 * (Xn - IMout)).astype(F32)
    l2 = (lambda_2 + m2 * (Xn - U)).astype(F32)
    return Xn, l1, l2


def outer_iteration(st: State, Y_observed: np.ndarray, MtM: np.ndarray, D: np.ndarray, prm: Params,
                    a: Optional[np.ndarray] = None, low_rank=None, dtype=F32) -> State:
    """One pass of main_LRS_PnP.py:250-362 (metrics excluded)."""
    R, C = Y_observed.shape
    Phi_z, a = sparse_step(st.X, st.lambda_1, Y_observed, D, prm, a=a, dtype=dtype)
    Z = (st.X + F32(1.0 / prm.mu_2) * st.lambda_2).astype(F32)                    # :315
    U = svt(Z, 1.0 / prm.mu_2) if low_rank is None else low_rank(Z)
    IMout = col2im_accumulate(Phi_z, R, C, prm.bb, prm.slidingDis)               # :332-339
    Weight = coverage_weight(R, C, prm.bb, prm.slidingDis)
    l1s = lambda1_summation(st.lambda_1, R, C, prm.bb, prm.slidingDis)
    X, l1, l2 = admm_update(st.X, st.lambda_1, st.lambda_2, Y_observed.astype(F32), MtM.astype(F32), IMout,
                            Weight, U, l1s, prm)
    return State(X=X, lambda_1=l1, lambda_2=l2, history=st.history)


def run(Y_observed: np.ndarray, MtM: np.ndarray, D: np.ndarray, prm: Params, iteration_num: int = 2,
        dtype=F32) -> State:
    """main_LRS_PnP.py:218-229 init (X = Y_observed, λ = 0) + the loop."""
    Y = Y_observed.astype(F32)
    st = State(X=Y.copy(), lambda_1=np.zeros_like(Y), lambda_2=np.zeros_like(Y))
    blocks_copy, _, _, _ = get_image_block(Y, prm.bb, prm.slidingDis)
    a = step_constants_batched(D.astype(dtype), patch_masks(blocks_copy), prm.step)
    for _ in range(iteration_num):
        st = outer_iteration(st, Y, MtM, D, prm, a=a, dtype=dtype)
    return st


# --------------------------------------------------------------------------
# metrics (main_LRS_PnP.py:40-58, 379-384)
# --------------------------------------------------------------------------
def psnr_ref(a: np.ndarray, b: np.ndarray) -> float:
    """The reference's non-standard 10*log10(255/sqrt(mse)) (:46)."""
    mse = float(np.mean((a.astype(F32) - b.astype(F32)) ** 2))
    if mse < 1.0e-10:
        return 100.0
    return 10 * math.log10(255 / math.sqrt(mse))


def mpsnr_ref(clean: np.ndarray, pred: np.ndarray) -> float:
    """bach_mpsnr (:48-58) on ``[1,B,h,w]`` tensors."""
    return float(np.mean([psnr_ref(clean[0, k], pred[0, k]) for k in range(clean.shape[1])]))


# --------------------------------------------------------------------------
# PnP denoiser: NLmeansfilter.m on a K x 1 column  (pnp_ista.m:30, NLmeansfilter.m:1-91)
# PARITY UNPINNED: no MATLAB/Octave here and the Python scripts call skimage's denoise_nl_means instead
# (absent, version unpinned).  This restates the in-repo MATLAB file; nothing executes the original.
# --------------------------------------------------------------------------
def nlm_kernel_rowsums(f: int = 3) -> np.ndarray:
    """Row sums of make_kernel(f)/sum (NLmeansfilter.m:80-91,27).  On a single-column input the symmetric padding
    (:24) makes all 2f+1 window columns identical, so the 2-D weighted distance (:63) collapses to these taps."""
    k = np.zeros((2 * f + 1, 2 * f + 1))
    for d in range(1, f + 1):
        k[f - d:f + d + 1, f - d:f + d + 1] += 1.0 / (2 * d + 1) ** 2
    k /= f
    k /= k.sum()
    return k.sum(axis=1)


def nlm_column(x: np.ndarray, t: int, f: int, h: float) -> np.ndarray:
    """NLmeansfilter(x, t, f, h) for x of shape (K,) or (K,1): search radius t along the column only (the search
    window is clipped to the real column, :46-49), patch radius f, weights exp(-d/h^2) (:29,65), centre pixel
    weighted by the maximum weight (:77-78), input returned where the weights underflow to zero (:80-84)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    K = x.shape[0]
    xp = np.pad(x, f, mode="symmetric")
    kw = nlm_kernel_rowsums(f)
    h2 = float(h) * float(h)
    out = np.empty(K)
    for i in range(K):
        w1 = xp[i:i + 2 * f + 1]
        wmax = avg = sw = 0.0
        for r in range(max(i - t, 0), min(i + t, K - 1) + 1):
            if r == i:
                continue
            w2 = xp[r:r + 2 * f + 1]
            d = float(np.sum(kw * (w1 - w2) ** 2))
            w = math.exp(-d / h2) if h2 > 0 else (1.0 if d == 0 else 0.0)
            wmax = max(wmax, w)
            sw += w
            avg += w * x[r]
        avg += wmax * x[i]
        sw += wmax
        out[i] = avg / sw if sw > 0 else x[i]
    return out


def pnp_ista_nlm(y: np.ndarray, H: np.ndarray, lambda_ista: float, Nit: int, a: float, h_scale: float = 0.1,
                 t: int = 3, f: int = 3) -> np.ndarray:
    """pnp_ista.m:14-32: x0 = 0; T = lambda/(2 a); x <- NLmeansfilter(x + H'(y - Hx)/a, 3, 3, 0.1 T)."""
    H = np.asarray(H, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    x = np.zeros((H.shape[1], 1))
    T = lambda_ista / (2 * a)
    for _ in range(Nit):
        g = x + (H.T @ (y - H @ x)) / a
        x = nlm_column(g, t, f, h_scale * T).reshape(-1, 1)
    return x
