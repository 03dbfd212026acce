"""Generate the golden fixtures in this directory by executing the reference's
OWN functions (AST-extracted from /root/reference, see oracle/ref_extract.py).

Run in the build container only (``/root/reference`` must exist):

    python tests/golden/make_golden.py

The outer-loop drivers below re-type the *script body* of main_LRS_PnP.py
(lines 244-362), which is top-level code and cannot be imported, around the
reference's literal ``get_image_block`` / ``delete_element`` / ``ista`` / ``SVT``.
The NLM denoiser (skimage, absent) is replaced by the soft threshold of the
MATLAB twin (ista.m:23) through the ``denoise_nl_means`` global.

Outputs (all small, committed):
  index_kat.npz        x_index / y_index / blocks of get_image_block on a grid of geometries
  ista_kat.npz         literal ista() on random problems (spectral & frob4 step, soft & identity denoiser)
  prox_kat.npz         SVT / Shrinkage_Operator / soft_thresh / l1_prox
  bundled_inputs.npz   unfolded Y_observed / clean / pixel masks of two bundled cubes
  e2e_bundled.npz      2 outer iterations of the literal LRS-PnP loop, base cube, synthetic D (K=324)
  metrics_kat.npz      bach_mpsnr / pytorch_ssim.ssim / state_convergence of the reference on two bundled cubes
  e2e_small.npz        2 outer iterations, synthetic 12x12x20 cube, bb=8 stride 1, K=128 (spectral and frob4)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_extract as rx  # noqa: E402
from lrs_pnp_dip_b200 import matio, synth  # noqa: E402

torch.set_num_threads(8)


def literal_outer_loop(ns, Y_observed, mask, D, *, gamma, mu_1, mu_2, lambda_ista, Nit, bb, slidingDis,
                       iteration_num, low_rank="svt"):
    """main_LRS_PnP.py:218-362 with the reference's own functions (``ns``)."""
    get_image_block, delete_element, ista, SVT = (ns["get_image_block"], ns["delete_element"], ns["ista"],
                                                  ns.get("SVT"))
    Full_Dictionary = D
    M_transpose_Y = Y_observed
    M_transpose_M = mask
    lambda_1 = torch.zeros(Y_observed.size())
    lambda_2 = torch.zeros(Y_observed.size())
    X = Y_observed
    blocks_copy, rows, cols, idx_Mat = get_image_block(Y_observed, bb, slidingDis)          # :244
    per_iter = []
    for itr in range(iteration_num):
        blocks, rows, cols, idx_Mat = get_image_block((X + lambda_1 / mu_1), bb, slidingDis)   # :259
        Phi_z = torch.zeros(blocks.size())
        for jj in range(Phi_z.size()[1]):                                                    # :270
            pruned_Dictionary = Full_Dictionary
            valid_pixel = blocks[:, jj].view((bb ** 2, 1))
            tempt = blocks_copy[:, jj].view((bb ** 2, 1))
            missing_index = np.where(tempt.flatten() == 0)[0]
            if len(missing_index) > 0:
                valid_pixel = delete_element(valid_pixel, torch.Tensor(missing_index).tolist())
                pruned_Dictionary = delete_element(pruned_Dictionary, torch.Tensor(missing_index).tolist())
            Coefs = ista(valid_pixel, pruned_Dictionary, lambda_ista, 0, Nit)
            Phi_z[:, jj] = torch.mm(Full_Dictionary, Coefs).flatten()
        Z = X + (1 / mu_2) * lambda_2
        if low_rank == "svt":
            U = SVT(Z, 1 / mu_2)                                                             # :315
        else:                                                                                # DIP stand-in
            U = Z.clone()
        Weight = torch.zeros(Y_observed.size())
        IMout = torch.zeros(Y_observed.size())
        blocks_lamdba_1, rows, cols, idxMat = get_image_block(lambda_1, bb, slidingDis)      # :328
        lambda1_summation = torch.zeros(Y_observed.size())
        count = 0
        for i in range(len(cols)):                                                           # :332-344
            row = rows[i]
            col = cols[i]
            block = torch.Tensor(Phi_z[:, count].view((bb, bb)).numpy().transpose(1, 0))
            blocks_lambda = torch.Tensor(blocks_lamdba_1[:, count].view((bb, bb)).numpy().transpose(1, 0))
            IMout[row:row + bb, col:col + bb] = IMout[row:row + bb, col:col + bb] + block
            Weight[row:row + bb, col:col + bb] = Weight[row:row + bb, col:col + bb] + torch.ones(bb)
            lambda1_summation[row:row + bb, col:col + bb] = lambda1_summation[row:row + bb, col:col + bb] + blocks_lambda
            count = count + 1
        X = (gamma * M_transpose_Y + mu_1 * IMout + mu_2 * U - lambda1_summation - lambda_2) / (
            gamma * M_transpose_M + mu_1 * Weight + mu_2)                                    # :346
        lambda_1 = lambda_1 + mu_1 * (X - IMout)                                             # :361
        lambda_2 = lambda_2 + mu_2 * (X - U)                                                 # :362
        per_iter.append(dict(X=X.numpy().copy(), lambda_1=lambda_1.numpy().copy(), lambda_2=lambda_2.numpy().copy(),
                             Phi_z=Phi_z.numpy().copy(), IMout=IMout.numpy().copy(), Weight=Weight.numpy().copy(),
                             U=U.numpy().copy(), lam1sum=lambda1_summation.numpy().copy()))
    return per_iter


def make_index_kat():
    ns = rx.extract("main_LRS_PnP.py")
    gib = ns["get_image_block"]
    rng = np.random.default_rng(11)
    out = {}
    cases = [(1296, 128, 36, 36), (64, 41, 8, 3), (64, 40, 8, 8), (50, 23, 8, 1), (37, 19, 4, 3), (30, 30, 5, 2),
             (17, 9, 8, 1), (8, 8, 8, 1), (100, 12, 6, 4), (41, 64, 8, 5), (20, 20, 3, 7)]
    for ci, (R, C, bb, s) in enumerate(cases):
        X = torch.tensor(rng.standard_normal((R, C)).astype(np.float32))
        blocks, x, y, idx = gib(X, bb, s)
        out[f"c{ci}_geom"] = np.array([R, C, bb, s], dtype=np.int64)
        out[f"c{ci}_x"] = np.asarray(x, dtype=np.int64)
        out[f"c{ci}_y"] = np.asarray(y, dtype=np.int64)
        out[f"c{ci}_idxsum"] = np.array([float(idx.sum())])
        if blocks.numel() <= 200000:
            out[f"c{ci}_X"] = X.numpy()
            out[f"c{ci}_blocks"] = blocks.numpy()
        else:
            out[f"c{ci}_X"] = X.numpy()
            out[f"c{ci}_blocks_sample"] = blocks.numpy()[:, ::7].copy()
    out["ncases"] = np.array([len(cases)])
    np.savez_compressed(os.path.join(HERE, "index_kat.npz"), **out)
    print("index_kat:", len(cases), "cases")


def make_ista_kat():
    rng = np.random.default_rng(5)
    out = {}
    ns_lrs = rx.extract("main_LRS_PnP.py")
    ns_dip = rx.extract("main_LRS_PnP_DIP_pro.py")
    probs = [(64, 96, 40), (48, 128, 80), (30, 50, 25), (64, 256, 80)]
    for i, (n, K, Nit) in enumerate(probs):
        H = torch.tensor(rng.standard_normal((n, K)).astype(np.float32) / np.sqrt(n).astype(np.float32))
        x_true = np.zeros((K, 1), np.float32)
        sup = rng.choice(K, 5, replace=False)
        x_true[sup, 0] = rng.standard_normal(5).astype(np.float32)
        y = torch.tensor(H.numpy() @ x_true + 0.05 * rng.standard_normal((n, 1)).astype(np.float32))
        out[f"p{i}_H"] = H.numpy()
        out[f"p{i}_y"] = y.numpy()
        out[f"p{i}_Nit"] = np.array([Nit])
        ns_lrs["denoise_nl_means"] = rx.soft_shim(10.0)      # h = 0.1*T  (main_LRS_PnP.py:146)
        out[f"p{i}_x_spectral_soft"] = ns_lrs["ista"](y, H, 0.1, 0, Nit).numpy()
        ns_lrs["denoise_nl_means"] = rx.identity_shim
        out[f"p{i}_x_spectral_identity"] = ns_lrs["ista"](y, H, 0.1, 0, Nit).numpy()
        ns_dip["denoise_nl_means"] = rx.soft_shim(1.0)       # h = T      (main_LRS_PnP_DIP_pro.py:199)
        out[f"p{i}_x_frob4_soft"] = ns_dip["ista"](y, H, 0.1, 0, Nit).numpy()
        out[f"p{i}_a_spectral"] = np.array([np.linalg.norm(H, 2) ** 2], dtype=np.float64)
        out[f"p{i}_a_frob4"] = np.array([2 * (np.trace(torch.mm(H.T, H).numpy()) + np.trace(torch.mm(H.T, H).numpy()))],
                                        dtype=np.float64)
    out["nprobs"] = np.array([len(probs)])
    np.savez_compressed(os.path.join(HERE, "ista_kat.npz"), **out)
    print("ista_kat:", len(probs), "problems")


def make_prox_kat():
    rng = np.random.default_rng(7)
    ns = rx.extract("main_LRS_PnP.py")
    l1 = rx.extract("admm_utils.py")["l1_prox"]
    Z = (rng.standard_normal((96, 8)) @ rng.standard_normal((8, 24)) + 0.05 * rng.standard_normal((96, 24))).astype(np.float32)
    v = rng.standard_normal((257,)).astype(np.float32)
    out = dict(Z=Z, tau=np.array([1.0 / 0.9]), svt=ns["SVT"](torch.tensor(Z), 1.0 / 0.9).numpy(),
               v=v, thr=np.array([0.3]),
               shrink=ns["Shrinkage_Operator"](v.copy(), np.float32(0.3)),
               soft_thresh=ns["soft_thresh"](v.copy(), np.float32(0.3)),
               l1_prox=l1(torch.tensor(v), 0.3).numpy())
    np.savez_compressed(os.path.join(HERE, "prox_kat.npz"), **out)
    print("prox_kat done")


def bundled(name_noisy, name_clean, name_mask):
    d = os.path.join(rx.REFERENCE_ROOT, "data")
    noisy = matio.load_cube(os.path.join(d, name_noisy))
    clean = matio.load_cube(os.path.join(d, name_clean))
    msk = matio.loadmat_any(os.path.join(d, name_mask))["msk"]
    return matio.unfold_cube(noisy), matio.unfold_cube(clean), np.asarray(msk, np.uint8).transpose(0, 1, 3, 2).reshape(-1)


def make_bundled_inputs():
    out = {}
    for tag, files in dict(base=("low_rank_sparsity_noisy.mat", "low_rank_sparsity_clean.mat", "low_rank_sparsity_mask.mat"),
                           img5=("low_rank_sparsity_noisy_img5.mat", "low_rank_sparsity_clean_img5.mat", "fourth_mask.mat"),
                           img2=("low_rank_sparsity_noisy_img2.mat", "low_rank_sparsity_clean_img2.mat", "second_mask.mat"),
                           ).items():
        Y, clean, pm = bundled(*files)
        out[f"{tag}_Y"] = Y
        out[f"{tag}_clean"] = clean.astype(np.float16) if tag == "img2" else clean
        out[f"{tag}_pixmask"] = pm
    np.savez_compressed(os.path.join(HERE, "bundled_inputs.npz"), **out)
    print("bundled_inputs done")


def make_e2e_bundled():
    ns = rx.extract("main_LRS_PnP.py")
    ns["denoise_nl_means"] = rx.soft_shim(10.0)
    Y, clean, pm = bundled("low_rank_sparsity_noisy.mat", "low_rank_sparsity_clean.mat", "low_rank_sparsity_mask.mat")
    mask = np.repeat(pm.astype(np.float32)[:, None], 128, axis=1)
    K = 324
    D = synth.synthetic_dictionary(1296, K, seed=0)
    it = literal_outer_loop(ns, torch.tensor(Y), torch.tensor(mask), torch.tensor(D), gamma=0.5, mu_1=0.15,
                            mu_2=0.15 * 6, lambda_ista=0.1, Nit=80, bb=36, slidingDis=36, iteration_num=2)
    out = dict(K=np.array([K]), X1=it[0]["X"], X2=it[1]["X"], lambda_1_2=it[1]["lambda_1"].astype(np.float16),
               lambda_2_2=it[1]["lambda_2"].astype(np.float16),
               Phi_z1_norm=np.array([np.linalg.norm(it[0]["Phi_z"].astype(np.float64))]),
               Phi_z1_sample=it[0]["Phi_z"][::9, ::5].copy(),
               U1_sample=it[0]["U"][::5, ::3].copy())
    np.savez_compressed(os.path.join(HERE, "e2e_bundled.npz"), **out)
    print("e2e_bundled done")


def make_e2e_small():
    out = {}
    H_, W_, B_ = 12, 12, 20
    clean, noisy = synth.synthetic_cube(H_, W_, B_, rank=4, seed=21)
    pm = synth.pixel_mask(H_, W_, "bernoulli", keep=0.6, seed=23)
    Y = synth.observe(noisy, pm)
    mask = np.repeat(pm.astype(np.float32)[:, None], B_, axis=1)
    D = synth.synthetic_dictionary(64, 128, seed=0)
    out.update(Y=Y, clean=clean, pixmask=pm, D=D)
    for variant, script, scale, mu1, mu2 in (("spectral", "main_LRS_PnP.py", 10.0, 0.15, 0.9),
                                            ("frob4", "main_LRS_PnP_DIP_pro.py", 1.0, 0.1, 0.1)):
        ns = rx.extract(script)
        if "SVT" not in ns:
            ns["SVT"] = rx.extract("main_LRS_PnP.py")["SVT"]
        ns["denoise_nl_means"] = rx.soft_shim(scale)
        it = literal_outer_loop(ns, torch.tensor(Y), torch.tensor(mask), torch.tensor(D), gamma=0.5, mu_1=mu1, mu_2=mu2,
                                lambda_ista=0.1, Nit=80, bb=8, slidingDis=1, iteration_num=2,
                                low_rank="svt" if variant == "spectral" else "identity")
        for k in ("X", "lambda_1", "lambda_2", "IMout", "Weight", "U", "lam1sum"):
            out[f"{variant}_{k}_1"] = it[0][k]
            out[f"{variant}_{k}_2"] = it[1][k]
        out[f"{variant}_Phi_z_1"] = it[0]["Phi_z"]
    np.savez_compressed(os.path.join(HERE, "e2e_small.npz"), **out)
    print("e2e_small done")


def make_metrics_kat():
    """Reference metrics on the bundled cubes: bach_mpsnr (main_LRS_PnP.py:48-58) and pytorch_ssim.ssim
    (pytorch_ssim/__init__.py:65-73), both imported/extracted from the reference."""
    sys.path.insert(0, rx.REFERENCE_ROOT)
    import pytorch_ssim  # the reference's own module

    ns = rx.extract("main_LRS_PnP.py")
    out = {}
    for tag, files in dict(base=("low_rank_sparsity_noisy.mat", "low_rank_sparsity_clean.mat"),
                           img5=("low_rank_sparsity_noisy_img5.mat", "low_rank_sparsity_clean_img5.mat")).items():
        d = os.path.join(rx.REFERENCE_ROOT, "data")
        noisy = torch.tensor(matio.load_cube(os.path.join(d, files[0])))
        clean = torch.tensor(matio.load_cube(os.path.join(d, files[1])))
        out[f"{tag}_mpsnr_in"] = np.array([ns["bach_mpsnr"](clean, noisy)])
        out[f"{tag}_mssim_in"] = np.array([float(pytorch_ssim.ssim(clean, noisy))])
        out[f"{tag}_state"] = np.array([float(ns["state_convergence"](clean, noisy))])
    np.savez_compressed(os.path.join(HERE, "metrics_kat.npz"), **out)
    print("metrics_kat:", {k: float(v[0]) for k, v in out.items()})


if __name__ == "__main__":
    if not rx.reference_available():
        sys.exit("reference checkout not found; fixtures can only be generated in the build container")
    make_index_kat()
    make_ista_kat()
    make_prox_kat()
    make_bundled_inputs()
    make_e2e_small()
    make_e2e_bundled()
    make_metrics_kat()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
