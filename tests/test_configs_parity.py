"""Parity on the BASELINE.json configurations AS SHIPPED, against fixtures produced by the reference's own functions in the
literal script loop (tests/golden/make_golden_r2.py → e2e_configs.npz):

  cfg 1    main_LRS_PnP.py: noisy_img5 + fourth_mask (main_LRS_PnP.py:170,183), 2 outer iterations
  cfg 2/3  the DIP variants' parameters (main_LRS_PnP_DIP_pro.py:324-341: mu1 = mu2 = 0.1, Nit = 100, a = 4||H||_F^2) on
           img2..img5 with second / third / fourth masks (main_LRS_PnP_DIP_1-LiP.py:270-294), identity low-rank stand-in
  cfg 5    a reduced cube with the stripe U Bernoulli mask at 224 bands (literal loop), and the FULL 1024 x 1024 x 224 cube
           sampled against the oracle

CPU tests pin the oracle to the fixtures; GPU tests check the CUDA path against the same fixtures.  Tolerances are the
north_star's: recovered cube 1e-4 rel-L2, MPSNR 0.01 dB, MSSIM 1e-4.
"""
import numpy as np
import pytest
import torch

from lrs_pnp_dip_b200 import matio, metrics, synth
from oracle import lrs_oracle as orc

IMGS = ("img2", "img3", "img4", "img5")
ZEROS = dict(img2=300, img3=330, img4=432, img5=432)          # mask zeros, SURVEY §8c


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def bundled(golden, tag):
    g = golden("bundled_inputs_r2" if tag in ("img3", "img4") else "bundled_inputs")
    Y, pm, clean = g[f"{tag}_Y"], g[f"{tag}_pixmask"], g[f"{tag}_clean"].astype(np.float32)
    assert int((pm == 0).sum()) == ZEROS[tag] and ((Y == 0).all(axis=1) == (pm == 0)).all()
    return Y, np.repeat(pm.astype(np.float32)[:, None], 128, axis=1), clean


def ref_metrics(clean, X):
    c = torch.from_numpy(matio.fold_cube(clean, 36, 36))
    x = torch.from_numpy(matio.fold_cube(np.asarray(X, np.float32), 36, 36))
    return metrics.mpsnr(c, x), metrics.ssim(c, x)


DIP = dict(gamma=0.5, mu_1=0.1, mu_2=0.1, lambda_ista=0.1, Nit=100, bb=36, slidingDis=36, step="frob4")


# ------------------------------------------------------------------------------------------------ CPU: oracle pinned
def test_oracle_cfg1_as_shipped(golden):
    ge = golden("e2e_configs")
    Y, MtM, clean = bundled(golden, "img5")
    D = synth.synthetic_dictionary(1296, int(ge["K_bundled"][0]), seed=0)
    st = orc.run(Y, MtM, D, orc.Params(), iteration_num=1)
    assert rel(st.X, ge["cfg1_X1"]) < 3e-5
    st = orc.outer_iteration(st, Y, MtM, D, orc.Params())
    assert rel(st.X, ge["cfg1_X2"]) < 5e-5
    # the repo's metric restatements reproduce the reference's own bach_mpsnr / pytorch_ssim.ssim on the fixture
    mp, ms = ref_metrics(clean, ge["cfg1_X2"])
    assert abs(mp - ge["cfg1_mpsnr_mssim_2"][0]) < 1e-3 and abs(ms - ge["cfg1_mpsnr_mssim_2"][1]) < 1e-5


@pytest.mark.parametrize("tag", IMGS)
def test_oracle_cfg23_parameters(golden, tag):
    ge = golden("e2e_configs")
    Y, MtM, _ = bundled(golden, tag)
    D = synth.synthetic_dictionary(1296, int(ge["K_bundled"][0]), seed=0)
    prm = orc.Params(**DIP)
    st = orc.State(X=Y.copy(), lambda_1=np.zeros_like(Y), lambda_2=np.zeros_like(Y))
    for _ in range(2):
        st = orc.outer_iteration(st, Y, MtM, D, prm, low_rank=lambda Z: Z.copy())
    assert rel(st.X, ge[f"cfg23_{tag}_X2"]) < 5e-5


def test_oracle_cfg5_reduced(golden):
    ge = golden("e2e_configs")
    H, W, B = (int(v) for v in ge["cfg5s_geom"])
    Y, pm = ge["cfg5s_Y"], ge["cfg5s_pixmask"]
    assert pm.reshape(H, W)[:, 2::3].sum() == 0                       # every third image column dropped
    MtM = np.repeat(pm.astype(np.float32)[:, None], B, axis=1)
    D = synth.synthetic_dictionary(64, 256, seed=0)
    prm = orc.Params(Nit=80, bb=8, slidingDis=1, step="spectral")
    phi, _ = orc.sparse_step(Y, np.zeros_like(Y), Y, D, prm)
    assert rel(phi[:, ::3], ge["cfg5s_Phi_z1_cols3"]) < 3e-5
    assert rel(orc.col2im_accumulate(phi, Y.shape[0], B, 8, 1), ge["cfg5s_IMout1"]) < 3e-5
    st = orc.run(Y, MtM, D, prm, iteration_num=2)
    assert rel(st.X, ge["cfg5s_X2"]) < 1e-4


# ------------------------------------------------------------------------------------------------ GPU: the CUDA path
@pytest.fixture(scope="module")
def lrs():
    import lrs_pnp_dip_b200 as m

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    m._lib.lib()
    return m


@pytest.mark.gpu
def test_gpu_cfg1_as_shipped(lrs, golden):
    ge = golden("e2e_configs")
    Y, MtM, clean = bundled(golden, "img5")
    D = synth.synthetic_dictionary(1296, int(ge["K_bundled"][0]), seed=0)
    sol = lrs.LRSPnP(Y, MtM, D, lrs.Params())
    sol.step()
    assert rel(sol.X.cpu().numpy(), ge["cfg1_X1"]) < 1e-4
    sol.step()
    X2 = sol.X.cpu().numpy()
    assert rel(X2, ge["cfg1_X2"]) < 1e-4
    mp, ms = ref_metrics(clean, X2)
    assert abs(mp - ge["cfg1_mpsnr_mssim_2"][0]) < 0.01 and abs(ms - ge["cfg1_mpsnr_mssim_2"][1]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("tag", IMGS)
def test_gpu_cfg23_parameters(lrs, golden, tag):
    ge = golden("e2e_configs")
    Y, MtM, clean = bundled(golden, tag)
    D = synth.synthetic_dictionary(1296, int(ge["K_bundled"][0]), seed=0)
    prm = lrs.Params(**DIP)
    sol = lrs.LRSPnP(Y, MtM, D, prm, low_rank=lambda Z: Z.clone())
    phi1 = sol.be.coder.phi_z(sol.X, sol.lambda_1).cpu().numpy()
    assert rel(phi1[::9, ::5], ge[f"cfg23_{tag}_Phi_z1_sample"]) < 1e-4
    sol.run(2)
    X2 = sol.X.cpu().numpy()
    assert rel(X2, ge[f"cfg23_{tag}_X2"]) < 1e-4
    mp, ms = ref_metrics(clean, X2)
    assert abs(mp - ge[f"cfg23_{tag}_mpsnr_mssim_2"][0]) < 0.01 and abs(ms - ge[f"cfg23_{tag}_mpsnr_mssim_2"][1]) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["tc", "simt"])
def test_gpu_cfg5_reduced(lrs, golden, engine):
    ge = golden("e2e_configs")
    H, W, B = (int(v) for v in ge["cfg5s_geom"])
    Y, pm = ge["cfg5s_Y"], ge["cfg5s_pixmask"]
    MtM = np.repeat(pm.astype(np.float32)[:, None], B, axis=1)
    D = synth.synthetic_dictionary(64, 256, seed=0)
    prm = lrs.Params(Nit=80, bb=8, slidingDis=1, step="spectral")
    sol = lrs.LRSPnP(Y, MtM, D, prm, engine=engine)
    phi = sol.be.coder.phi_z(sol.X, sol.lambda_1)
    assert rel(phi[:, ::3].cpu().numpy(), ge["cfg5s_Phi_z1_cols3"]) < 1e-4
    assert rel(lrs.col2im(phi, Y.shape[0], B, 8, 1).cpu().numpy(), ge["cfg5s_IMout1"]) < 1e-4
    sol.step()
    assert rel(sol.X.cpu().numpy(), ge["cfg5s_X1"]) < 1e-4
    sol.step()
    assert rel(sol.X.cpu().numpy(), ge["cfg5s_X2"]) < 1e-4


@pytest.mark.gpu
def test_gpu_cfg5_full_size_sampled_vs_oracle(lrs):
    """BASELINE.json cfg 5 at FULL size (1024 x 1024 x 224 cube, stripe U Bernoulli(0.75) mask, 227.5 M patches, K = 256,
    80 iterations) through the shipped engine in patch sub-ranges: for three groups of column starts the sparse step runs
    on all 1 048 569 row starts, and three blocks of 9 row starts are recomputed by the oracle and compared patch by patch."""
    import bench

    Y, pm, D = bench.make_inputs("cfg5")
    R, C = Y.shape
    assert (R, C) == (1048576, 224) and pm.reshape(1024, 1024)[:, 2::3].sum() == 0
    rng = np.random.default_rng(6)
    Lam = (rng.standard_normal((R, C)) * 0.01).astype(np.float32)
    prm = lrs.Params(Nit=80, bb=8, slidingDis=1, step="spectral")
    sc = lrs.SparseCoder(torch.from_numpy(Y).cuda(), torch.from_numpy(D).cuda(), prm)
    assert sc.fused and sc.P == 227539473
    Xd, Ld = sc.Y, torch.from_numpy(Lam).cuda()
    nR = R - 7
    oprm = orc.Params(mu_1=prm.mu_1, Nit=80, bb=8, slidingDis=1, step="spectral")
    for c0 in (0, 100, C - 8 - 3):                                  # 4 column starts each: 4.2 M patches per launch
        phi = sc.phi_z_range(Xd, Ld, c0 * nR, (c0 + 4) * nR)
        assert bool(torch.isfinite(phi).all())
        for r0 in (0, 524288 - 5, R - 16):
            rows = slice(r0, r0 + 16)
            ref, _ = orc.sparse_step(Y[rows, c0:c0 + 11], Lam[rows, c0:c0 + 11], Y[rows, c0:c0 + 11], D, oprm)   # 9 x 4 patches
            cols = (np.arange(4)[:, None] * nR + (r0 + np.arange(9))[None, :]).reshape(-1)
            got = phi[:, torch.as_tensor(cols, device="cuda")].cpu().numpy()
            assert rel(got, ref) < 1e-4, (c0, r0)


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["tc", "simt"])
def test_gpu_adversarial_dictionary_and_dynamic_range(lrs, engine):
    """Stress for the 3-pass fp16 split of the tensor-core engine: a COHERENT dictionary (atoms = one common direction +
    10 % perturbation: mutual coherence ~0.99, spectral norm^2 ~ K), patch values spanning five decades inside every
    8x8 window (unfolded rows scaled by 10^-(0..5)), Nit = 100.  Compared with the fp64 oracle and the fp32 oracle: the
    device result must be as close to fp64 as the reference's own fp32 arithmetic is (within 1e-4 rel-L2)."""
    rng = np.random.default_rng(12)
    R, C, K = 160, 24, 256
    base = rng.standard_normal((64, 1))
    D = base + 0.1 * rng.standard_normal((64, K))
    D = (D / np.linalg.norm(D, axis=0, keepdims=True)).astype(np.float32)
    coh = np.abs(D.T @ D - np.eye(K)).max()
    assert coh > 0.95
    rowscale = 10.0 ** (-(np.arange(R) % 6).astype(np.float64))     # six consecutive rows span 1 .. 1e-5
    X = (rng.standard_normal((R, C)) * rowscale[:, None]).astype(np.float32)
    pm = rng.random(R) < 0.7
    Y = X.copy()
    Y[Y == 0] = 1e-12
    Y[~pm] = 0
    Lam = (0.01 * rng.standard_normal((R, C)) * rowscale[:, None]).astype(np.float32)
    prm = lrs.Params(mu_1=0.15, lambda_ista=0.01, Nit=100, bb=8, slidingDis=1, step="spectral")
    oprm = orc.Params(mu_1=0.15, lambda_ista=0.01, Nit=100, bb=8, slidingDis=1, step="spectral")
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    phi = lrs.SparseCoder(cu(Y), cu(D), prm, engine=engine).phi_z(cu(X), cu(Lam)).cpu().numpy()
    ref32, a = orc.sparse_step(X, Lam, Y, D, oprm)
    ref64, _ = orc.sparse_step(X, Lam, Y, D, oprm, a=a.astype(np.float64), dtype=np.float64)
    assert np.isfinite(phi).all()
    assert rel(phi, ref64) < 1e-4 and rel(phi, ref32) < 1e-4
    # per patch too: no single window (whatever its scale) is off by more than 1e-3 of its own norm
    num = np.linalg.norm(phi.astype(np.float64) - ref64, axis=0)
    den = np.maximum(np.linalg.norm(ref64, axis=0), 1e-30)
    assert float((num / den).max()) < 1e-3
