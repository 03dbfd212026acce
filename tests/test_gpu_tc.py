"""tcgen05 bring-up: operand layouts / descriptors of the tensor-core engine, pinned against NumPy
models of the hardware's operand rounding (tf32: 13 low mantissa bits dropped; f16: inputs rounded to
fp16 by the kernel), fp32 accumulation."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def tf32_trunc(x):
    return (x.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def probe(A, B, a_in_tmem, b_mn, f16):
    from lrs_pnp_dip_b200 import _lib

    L = _lib.diag_lib()          # probes live in the diagnostics build (include/lrs_pnp_diag.h)
    N, Kd = B.shape
    Ad, Bd = torch.tensor(A).cuda(), torch.tensor(B).cuda()
    C = torch.zeros((128, N), dtype=torch.float32, device="cuda")
    _lib.check(L.lrs_tc_probe_f32(Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), N, Kd, a_in_tmem, b_mn, f16,
                                  torch.cuda.current_stream().cuda_stream), "lrs_tc_probe_f32", L)
    torch.cuda.synchronize()
    return C.cpu().numpy()


def relerr(C, ref):
    return float(np.linalg.norm(C - ref) / np.linalg.norm(ref))


@pytest.mark.parametrize("a_in_tmem", [0, 1])
@pytest.mark.parametrize("shape", [(64, 64), (256, 64), (64, 256), (128, 32)])
def test_tf32_kmajor_layout_and_truncation(a_in_tmem, shape):
    N, Kd = shape
    rng = np.random.default_rng(N + Kd + a_in_tmem)
    A = rng.standard_normal((128, Kd)).astype(np.float32)
    B = rng.standard_normal((N, Kd)).astype(np.float32)
    C = probe(A, B, a_in_tmem, 0, 0)
    trunc = tf32_trunc(A).astype(np.float64) @ tf32_trunc(B).astype(np.float64).T
    exact = A.astype(np.float64) @ B.astype(np.float64).T
    assert relerr(C, exact) < 2e-3          # layouts / descriptors are right
    assert relerr(C, trunc) < 2e-6          # kind::tf32 drops the 13 low mantissa bits of both operands


def test_tf32_mn_major_needs_swizzled_layout():
    """Documented hardware behaviour the design rests on: an MN-major tf32 operand in the SWIZZLE_NONE
    canonical layout is not usable (CUTLASS: 'for mn-major tf32 operands, SW128_32B is the only
    available smem layout'); the same layout works for 16-bit operands (next test)."""
    rng = np.random.default_rng(1)
    A = rng.standard_normal((128, 64)).astype(np.float32)
    B = rng.standard_normal((64, 64)).astype(np.float32)
    C = probe(A, B, 1, 1, 0)
    exact = A.astype(np.float64) @ B.astype(np.float64).T
    assert relerr(C, exact) > 0.5


@pytest.mark.parametrize("a_in_tmem", [0, 1])
@pytest.mark.parametrize("b_mn", [0, 1])
@pytest.mark.parametrize("shape", [(64, 64), (256, 64), (128, 256), (64, 256), (128, 32)])
def test_f16_layouts(a_in_tmem, b_mn, shape):
    N, Kd = shape
    rng = np.random.default_rng(N + Kd + 2 * a_in_tmem + b_mn)
    A = rng.standard_normal((128, Kd)).astype(np.float32)
    B = rng.standard_normal((N, Kd)).astype(np.float32)
    C = probe(A, B, a_in_tmem, b_mn, 1)
    ref = A.astype(np.float16).astype(np.float64) @ B.astype(np.float16).astype(np.float64).T
    assert relerr(C, ref) < 2e-6, (a_in_tmem, b_mn, shape)


def split16(x, scale):
    xs = (x * np.float32(scale)).astype(np.float32)
    h1 = xs.astype(np.float16).astype(np.float32)
    h2 = (xs - h1).astype(np.float16).astype(np.float32)
    return h1 / np.float32(scale), h2 / np.float32(scale)


def test_three_pass_fp16_split_recovers_fp32():
    """a = a1 + a2, b = b1 + b2 (fp16 pieces): a1 b1 + a2 b1 + a1 b2 reproduces the fp32 product to ~1e-7,
    with A in TMEM and B MN-major / K-major exactly as the fused kernel issues them."""
    rng = np.random.default_rng(0)
    A = (rng.standard_normal((128, 64)) * 0.7).astype(np.float32)
    B = (rng.standard_normal((256, 64)) * 0.125).astype(np.float32)
    a1, a2 = split16(A, 8.0)
    b1, b2 = split16(B, 4.0)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    for b_mn in (0, 1):
        C = probe(a1, b1, 1, b_mn, 1) + probe(a2, b1, 1, b_mn, 1) + probe(a1, b2, 1, b_mn, 1)
        assert relerr(C, ref) < 1e-6, b_mn
