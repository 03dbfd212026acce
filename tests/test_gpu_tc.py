"""tcgen05 bring-up: operand layouts / descriptors of the tensor-core engine, pinned against a NumPy
model of a TF32 GEMM (inputs with the 13 low mantissa bits dropped, fp32/64 accumulation)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def tf32_trunc(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def tf32_round(x):  # round to nearest, ties away (cvt.rna)
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = ((u + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return u.view(np.float32)


def probe(A, B, a_in_tmem, b_mn):
    from lrs_pnp_dip_b200 import _lib

    L = _lib.lib()
    N, Kd = B.shape
    Ad, Bd = torch.tensor(A).cuda(), torch.tensor(B).cuda()
    C = torch.zeros((128, N), dtype=torch.float32, device="cuda")
    _lib.check(L.lrs_tc_probe_f32(Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), N, Kd, a_in_tmem, b_mn,
                                  torch.cuda.current_stream().cuda_stream), "lrs_tc_probe_f32")
    torch.cuda.synchronize()
    return C.cpu().numpy()


@pytest.mark.parametrize("a_in_tmem", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("shape", [(64, 64), (256, 64), (64, 256), (128, 32)])
def test_probe_layouts_and_tf32_semantics(a_in_tmem, b_mn, shape):
    N, Kd = shape
    if not a_in_tmem and (N * Kd + 128 * Kd) * 4 > 200 * 1024:
        pytest.skip("operands exceed shared memory")
    rng = np.random.default_rng(N + Kd + 2 * a_in_tmem + b_mn)
    A = rng.standard_normal((128, Kd)).astype(np.float32)
    B = rng.standard_normal((N, Kd)).astype(np.float32)
    C = probe(A, B, a_in_tmem, b_mn)
    models = {
        "trunc/trunc": tf32_trunc(A).astype(np.float64) @ tf32_trunc(B).astype(np.float64).T,
        "round/round": tf32_round(A).astype(np.float64) @ tf32_round(B).astype(np.float64).T,
        "exact": A.astype(np.float64) @ B.astype(np.float64).T,
    }
    errs = {k: float(np.linalg.norm(C - v) / np.linalg.norm(v)) for k, v in models.items()}
    print(f"a_in_tmem={a_in_tmem} b_mn={b_mn} N={N} Kd={Kd}: {errs}")
    assert errs["exact"] < 2e-3, errs                      # layouts / descriptors are right
    assert errs["trunc/trunc"] < 2e-6, errs                # the tensor core drops the 13 low mantissa bits


def test_three_pass_split_recovers_fp32():
    """hi/lo split (3xTF32): A_hi B_hi + A_lo B_hi + A_hi B_lo ≈ fp32 product to ~1e-6."""
    rng = np.random.default_rng(0)
    A = rng.standard_normal((128, 64)).astype(np.float32)
    B = rng.standard_normal((64, 64)).astype(np.float32)
    Ah, Bh = tf32_trunc(A), tf32_trunc(B)
    Al, Bl = (A - Ah).astype(np.float32), (B - Bh).astype(np.float32)
    C = probe(Ah, Bh, 1, 0) + probe(Al, Bh, 1, 0) + probe(Ah, Bl, 1, 0)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    assert np.linalg.norm(C - ref) / np.linalg.norm(ref) < 2e-6
