"""Round-2 golden fixtures: the BASELINE.json configurations as shipped, through the reference's OWN functions.

Run in the build container only (``/root/reference`` must exist):

    python tests/golden/make_golden_r2.py

Like make_golden.py, the outer loop is the re-typed script body (``literal_outer_loop``) around the AST-extracted
reference functions, with the skimage NLM call replaced by the MATLAB twin's soft threshold.

Outputs:
  bundled_inputs_r2.npz  unfolded Y_observed / clean (fp16) / pixel masks of img3 and img4 (img2, img5, base are in
                         bundled_inputs.npz)
  e2e_configs.npz
    cfg1_*   main_LRS_PnP.py AS SHIPPED: noisy_img5 + fourth_mask (main_LRS_PnP.py:170,183), gamma 0.5, mu1 0.15,
             mu2 0.9, Nit 80, spectral step, SVT, 2 outer iterations, synthetic D (K = 324): X1, X2 and the reference's own
             bach_mpsnr / pytorch_ssim.ssim of X2 against the clean cube
    cfg23_<img>_*  the DIP variants' parameters (main_LRS_PnP_DIP_pro.py:324-341 = main_LRS_PnP_DIP_1-LiP.py:316-333:
             mu1 = mu2 = 0.1, Nit = 100, a = 4||H||_F^2, h = T) on img2..img5 with second/third/fourth masks
             (main_LRS_PnP_DIP_1-LiP.py:270-294), identity low-rank stand-in for the network, 2 outer iterations
    cfg5s_*  a reduced cfg-5 cube (4 x 9 pixels x 224 bands, every third image column dropped U Bernoulli(0.75),
             8x8 patches stride 1, K = 256, spectral step, SVT), 2 outer iterations of the literal loop

NB the literal ``get_image_block`` cannot be trusted beyond 8192 patches in THIS container: under numpy 2.3.5
``np.unravel_index`` on the (P,1)-shaped output of ``np.argwhere`` (main_LRS_PnP.py:94-96) repeats the 8192nd result for
every later element (a buffered-iterator bug of that numpy build, absent from the numpy 1.x the scripts were written
for).  Literal fixtures therefore keep P <= 8192 (the reduced cfg-5 cube has 29 x 217 = 6293 patches); larger geometries
are covered by the oracle, which is pinned to the literal functions at P <= 8192.

    python tests/golden/make_golden_r2.py cfg5s     # regenerate only the cfg5s_* entries
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import bundled, literal_outer_loop  # noqa: E402
from oracle import ref_extract as rx  # noqa: E402
from lrs_pnp_dip_b200 import matio, synth  # noqa: E402

torch.set_num_threads(8)

PAIRS = dict(img2=("low_rank_sparsity_noisy_img2.mat", "low_rank_sparsity_clean_img2.mat", "second_mask.mat"),
             img3=("low_rank_sparsity_noisy_img3.mat", "low_rank_sparsity_clean_img3.mat", "third_mask.mat"),
             img4=("low_rank_sparsity_noisy_img4.mat", "low_rank_sparsity_clean_img4.mat", "fourth_mask.mat"),
             img5=("low_rank_sparsity_noisy_img5.mat", "low_rank_sparsity_clean_img5.mat", "fourth_mask.mat"))
K_BUNDLED = 324


def ref_metrics(clean_unf, X_unf):
    sys.path.insert(0, rx.REFERENCE_ROOT)
    import pytorch_ssim  # the reference's own module

    ns = rx.extract("main_LRS_PnP.py")
    c = torch.tensor(matio.fold_cube(clean_unf.astype(np.float32), 36, 36))
    x = torch.tensor(matio.fold_cube(X_unf.astype(np.float32), 36, 36))
    return float(ns["bach_mpsnr"](c, x)), float(pytorch_ssim.ssim(c, x))


def main():
    if not rx.reference_available():
        sys.exit("reference checkout not found; fixtures can only be generated in the build container")
    inputs, out = {}, {}
    only5 = "cfg5s" in sys.argv[1:]
    if only5:
        out = dict(np.load(os.path.join(HERE, "e2e_configs.npz")))
        out = {k: v for k, v in out.items() if not k.startswith("cfg5s_")}
    else:
        bundled_configs(inputs, out)
    reduced_cfg5(out)
    if not only5:
        np.savez_compressed(os.path.join(HERE, "bundled_inputs_r2.npz"), **inputs)
    np.savez_compressed(os.path.join(HERE, "e2e_configs.npz"), **out)
    for f in ("bundled_inputs_r2.npz", "e2e_configs.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


def bundled_configs(inputs, out):
    D = synth.synthetic_dictionary(1296, K_BUNDLED, seed=0)
    out["K_bundled"] = np.array([K_BUNDLED])

    # ---- cfg 1 as shipped
    ns = rx.extract("main_LRS_PnP.py")
    ns["denoise_nl_means"] = rx.soft_shim(10.0)
    Y, clean, pm = bundled(*PAIRS["img5"])
    mask = np.repeat(pm.astype(np.float32)[:, None], 128, axis=1)
    it = literal_outer_loop(ns, torch.tensor(Y), torch.tensor(mask), torch.tensor(D), gamma=0.5, mu_1=0.15, mu_2=0.15 * 6,
                            lambda_ista=0.1, Nit=80, bb=36, slidingDis=36, iteration_num=2)
    out["cfg1_X1"], out["cfg1_X2"] = it[0]["X"], it[1]["X"]
    out["cfg1_mpsnr_mssim_2"] = np.array(ref_metrics(clean, it[1]["X"]))
    print("cfg1 (img5 + fourth_mask): MPSNR/MSSIM after 2 iterations", out["cfg1_mpsnr_mssim_2"])

    # ---- cfg 2/3 parameters on img2..img5
    ns_pro = rx.extract("main_LRS_PnP_DIP_pro.py")
    ns_lip = rx.extract("main_LRS_PnP_DIP_1-LiP.py")
    ns_pro["denoise_nl_means"] = ns_lip["denoise_nl_means"] = rx.soft_shim(1.0)
    # the two DIP scripts define the same ista: check on one problem, then use _pro's
    rng = np.random.default_rng(3)
    Ht = torch.tensor(rng.standard_normal((50, 70)).astype(np.float32))
    yt = torch.tensor(rng.standard_normal((50, 1)).astype(np.float32))
    assert torch.equal(ns_pro["ista"](yt, Ht, 0.1, 0, 20), ns_lip["ista"](yt, Ht, 0.1, 0, 20))
    for tag, files in PAIRS.items():
        Y, clean, pm = bundled(*files)
        if tag in ("img3", "img4"):
            inputs[f"{tag}_Y"], inputs[f"{tag}_clean"], inputs[f"{tag}_pixmask"] = Y, clean.astype(np.float16), pm
        mask = np.repeat(pm.astype(np.float32)[:, None], 128, axis=1)
        it = literal_outer_loop(ns_pro, torch.tensor(Y), torch.tensor(mask), torch.tensor(D), gamma=0.5, mu_1=0.1, mu_2=0.1,
                                lambda_ista=0.1, Nit=100, bb=36, slidingDis=36, iteration_num=2, low_rank="identity")
        out[f"cfg23_{tag}_X2"] = it[1]["X"]
        out[f"cfg23_{tag}_Phi_z1_sample"] = it[0]["Phi_z"][::9, ::5].copy()
        out[f"cfg23_{tag}_mpsnr_mssim_2"] = np.array(ref_metrics(clean, it[1]["X"]))
        print(f"cfg2/3 {tag}: MPSNR/MSSIM", out[f"cfg23_{tag}_mpsnr_mssim_2"])



def reduced_cfg5(out):
    ns = rx.extract("main_LRS_PnP.py")
    H_, W_, B_ = 4, 9, 224
    clean5, noisy5 = synth.synthetic_cube(H_, W_, B_, rank=4, seed=31)
    pm5 = synth.pixel_mask(H_, W_, "stripe+bernoulli", keep=0.75, seed=33)
    Y5 = synth.observe(noisy5, pm5)
    mask5 = np.repeat(pm5.astype(np.float32)[:, None], B_, axis=1)
    D5 = synth.synthetic_dictionary(64, 256, seed=0)
    ns["denoise_nl_means"] = rx.soft_shim(10.0)
    it = literal_outer_loop(ns, torch.tensor(Y5), torch.tensor(mask5), torch.tensor(D5), gamma=0.5, mu_1=0.15, mu_2=0.15 * 6,
                            lambda_ista=0.1, Nit=80, bb=8, slidingDis=1, iteration_num=2)
    out.update(cfg5s_geom=np.array([H_, W_, B_]), cfg5s_Y=Y5, cfg5s_pixmask=pm5, cfg5s_X1=it[0]["X"], cfg5s_X2=it[1]["X"],
               cfg5s_Phi_z1_cols3=it[0]["Phi_z"][:, ::3].copy(), cfg5s_IMout1=it[0]["IMout"])
    assert it[0]["Phi_z"].shape[1] <= 8192, "np.unravel_index of this numpy build breaks the literal indices beyond 8192 patches"
    print("cfg5 reduced: P =", it[0]["Phi_z"].shape[1], "observed pixels", int(pm5.sum()), "of", pm5.size)


if __name__ == "__main__":
    main()
