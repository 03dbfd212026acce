"""CPU-only checks of the boundary and the host logic: the shared object loads and exports every
symbol include/lrs_pnp.h declares, the host-side geometry entry points agree with the oracle, the
stripe partition is consistent, and compute entry points refuse to run without a GPU (no fallback)."""
import os
import re

import numpy as np
import pytest
import torch

import lrs_pnp_dip_b200 as lrs
from lrs_pnp_dip_b200 import _lib
from oracle import lrs_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as ge

        ge.build()


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "lrs_pnp.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lrs_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.lrs_version() >= 100


def test_diagnostics_live_in_their_own_header_and_library():
    """Probes, the tile-walk replay and the timing counters are declared in include/lrs_pnp_diag.h and exported by
    liblrs_pnp_diag.so only: the product header / library carry none of them."""
    import ctypes

    hdr = open(os.path.join(ROOT, "include", "lrs_pnp_diag.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lrs_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.DIAG_SIGNATURES), declared ^ set(_lib.DIAG_SIGNATURES)
    prod = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert not hasattr(prod, name), f"{name} leaked into the product library"
    D = _lib.diag_lib()
    for name in declared | set(_lib.SIGNATURES):
        assert hasattr(D, name), name


@pytest.mark.parametrize("geom", [(1296, 128, 36, 36), (64, 41, 8, 3), (64, 40, 8, 8), (50, 23, 8, 1), (37, 19, 4, 3),
                                  (262144, 191, 8, 1), (20, 20, 3, 7), (8, 8, 8, 1)])
def test_host_geometry_matches_oracle(geom):
    R, C, bb, s = geom
    rs, cs = lrs.patch_grid(R, C, bb, s)
    assert np.array_equal(rs, orc.axis_starts(R, bb, s))
    assert np.array_equal(cs, orc.axis_starts(C, bb, s))
    assert lrs.patch_count(R, C, bb, s) == len(rs) * len(cs)


def test_baseline_patch_counts():
    assert lrs.patch_count(1296, 128, 36, 36) == 144
    assert lrs.patch_count(262144, 191, 8, 1) == 48_233_208          # SURVEY §8 cfg 4
    assert lrs.patch_count(1048576, 224, 8, 1) == 227_539_473        # cfg 5


def test_bad_geometry_is_an_error():
    with pytest.raises(lrs.LrsError):
        lrs.patch_count(4, 100, 8, 1)
    L = _lib.lib()
    assert L.lrs_axis_count(10, 0, 1) < 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    x = torch.randn(40, 23)
    with pytest.raises(lrs.LrsError):
        lrs.get_image_block(x, 8, 1)
    with pytest.raises(lrs.LrsError):
        lrs.soft_thresh(np.ones(4, np.float32), 0.1)
    with pytest.raises(lrs.LrsError):
        lrs.LRSPnP(np.zeros((40, 23), np.float32), np.ones((40, 23), np.float32), np.zeros((64, 64), np.float32),
                   lrs.Params(bb=8, slidingDis=1))
    # the raw ABI reports a CUDA error rather than computing on the host
    L = _lib.lib()
    rc = L.lrs_soft_f32(1, 0.1, 1, 4, None)
    assert rc == -2 and b"CUDA" in L.lrs_last_error()


def test_delete_element_is_reference_row_deletion():
    t = torch.arange(20.0).view(5, 4)
    out = lrs.delete_element(t, [1, 3])
    assert torch.equal(out, t[[0, 2, 4]])


def test_stripe_partition():
    for R, bb, world in [(262144, 8, 8), (1048576, 8, 8), (100, 8, 3), (64, 8, 2)]:
        bd = lrs.stripe_bounds(R, bb, world)
        assert bd[0] == 0 and bd[-1] == R - bb + 1 and np.all(np.diff(bd) > 0)
        assert np.diff(bd).max() - np.diff(bd).min() <= 1
        owned = 0
        for g in range(world):
            st = lrs.make_stripe(R, bb, g, world)
            assert st.rows_local == st.b - st.a + bb - 1
            owned += st.rows_owned
            # local stride-1 geometry has exactly the owned patch row-starts
            assert st.rows_local - bb + 1 == st.b - st.a
        assert owned == R


def test_ista_workspace_covers_both_engines():
    """lrs_ista_workspace_bytes is pure host arithmetic (no GPU): it must cover the FFMA engine everywhere and the fp16
    operand pieces of the tensor-core engine on the shapes that engine takes (n, K >= 128, P >= 8)."""
    from lrs_pnp_dip_b200 import _lib

    L = _lib.lib()
    al = lambda b: (b + 255) // 256 * 256
    for n, K, P in ((64, 256, 5000), (1296, 2592, 144), (1296, 2592, 2304), (36, 80, 33), (128, 128, 8)):
        need = L.lrs_ista_workspace_bytes(n, K, P)
        assert need >= 2 * al(K * P * 4) + al(n * P * 4) + 2 * al(P * 4)
        if n >= 128 and K >= 128 and P >= 8:
            pieces = 2 * 2 * n * K * 2                                         # hi/lo pieces of D and of its transpose
            assert need >= pieces + 2 * (K + n) * P * 2                        # + hi/lo pieces of alpha and r
    assert L.lrs_ista_workspace_bytes(0, 5, 5) == 0 and L.lrs_ista_workspace_bytes(5, 5, -1) == 0


def test_fused_kernel_tile_walk_covers_every_patch_once():
    """The persistent tcgen05 kernel walks (row block, column-start chunk) work items; lrs_debug_tile_walk replays that
    integer logic on the host.  Every patch of the requested range must be owned by exactly one valid lane, for whole
    ranges, sub-ranges cutting through columns and tiles, strides with appended starts and any SM count."""
    from lrs_pnp_dip_b200 import _lib

    L = _lib.diag_lib()
    rng = np.random.default_rng(0)
    cases = [(300, 20, 1, None, 148), (40, 23, 1, None, 148), (64, 41, 3, None, 148), (1500, 30, 1, None, 7),
             (262144 // 64, 191, 1, None, 148), (9, 8, 5, None, 3), (8, 8, 1, None, 148), (50, 9, 20, None, 148)]
    for _ in range(40):
        R, Cc, s = int(rng.integers(8, 700)), int(rng.integers(8, 60)), int(rng.integers(1, 6))
        cases.append((R, Cc, s, "random", int(rng.integers(1, 200))))
    for R, Cc, s, mode, sms in cases:
        P = lrs.ops.patch_count(R, Cc, 8, s)
        ranges = [(0, P)]
        if mode == "random" or P > 300:
            a = int(rng.integers(0, P))
            ranges += [(a, int(rng.integers(a + 1, P + 1))), (P - 1, P), (0, 1)]
        for b, e in ranges:
            visits = np.zeros(e - b, dtype=np.int32)
            tiles = np.zeros(1, dtype=np.int64)
            rc = L.lrs_debug_tile_walk(R, Cc, 8, s, b, e, sms, visits.ctypes.data, tiles.ctypes.data)
            assert rc == 0, (L.lrs_last_error(), R, Cc, s, b, e, sms)
            assert visits.min() == 1 and visits.max() == 1, (R, Cc, s, b, e, sms)
            assert (e - b + 127) // 128 <= tiles[0] <= (e - b + 127) // 128 + 2 * (Cc - 7) + 2


def test_jacobi_eigensolver_schedule_pairs_every_two_columns_once():
    """lrs_sym_eig_jacobi_f64 orders the column pairs of a sweep as a round-robin tournament of 16 column blocks over the 8
    CTAs of a cluster (cross pairs every round, pairs inside a block in round 0).  lrs_debug_jacobi_schedule replays that
    index arithmetic on the host: for every order 1..256 each pair of columns meets exactly once per sweep, no column is
    used twice inside a sub-round, and the sweep has the minimum number of sub-rounds of a parallel order."""
    from lrs_pnp_dip_b200 import _lib

    L = _lib.diag_lib()
    for C in range(1, 257):
        meets = np.zeros((C, C), dtype=np.int32)
        conflicts, subrounds = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
        rc = L.lrs_debug_jacobi_schedule(C, meets.ctypes.data, conflicts.ctypes.data, subrounds.ctypes.data)
        assert rc == 0, (C, L.lrs_last_error())
        assert conflicts[0] == 0, C
        iu = np.triu_indices(C, 1)
        assert (meets[iu] == 1).all() and np.tril(meets).sum() == 0, C
        w = -(-C // 16)
        w += (w * C) & 1
        assert subrounds[0] == 15 * w + ((w + (w & 1)) - 1 if w > 1 else 0), (C, subrounds[0])
    meets = np.zeros(4, dtype=np.int32)
    assert L.lrs_debug_jacobi_schedule(0, meets.ctypes.data, meets.ctypes.data, meets.ctypes.data) != 0
    assert L.lrs_debug_jacobi_schedule(257, meets.ctypes.data, meets.ctypes.data, meets.ctypes.data) != 0


def test_sparse_step_column_ranges_are_contiguous_and_start_short_when_shared():
    """SparseCoder._ranges (pure host logic): ascending, contiguous column-start ranges that cover every column start once;
    with shared_start the first range is the short one that runs on the dynamically dealt kernel instance (about 2.5 ms of
    work, never longer than a regular range), the rest are regular."""
    from lrs_pnp_dip_b200.solver import Params, SparseCoder

    class Fake(SparseCoder):
        def __init__(self, R, C, Nit):
            self.R, self.C, self.n, self.prm = R, C, 64, Params(Nit=Nit, bb=8, slidingDis=1)

    for R, C, Nit in ((262144, 191, 80), (32775, 191, 80), (1048576, 224, 80), (700, 29, 10), (131079, 191, 100), (4096, 9, 1)):
        f = Fake(R, C, Nit)
        nC, cpc = C - 7, f._chunk_cols()
        for shared in (False, True):
            r = f._ranges(shared)
            assert r[0][0] == 0 and r[-1][1] == nC and all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(0 < c1 - c0 <= cpc for c0, c1 in r)
            if not shared:
                assert len(r) == -(-nC // cpc)
        first = f._ranges(True)[0]
        work_s = (first[1] - first[0]) * (R - 7) * Nit / SparseCoder.FUSED_PATCH_ITERS_PER_S
        assert first[1] - first[0] == min(cpc, nC) or work_s >= SparseCoder.SHARED_START_SECONDS * 0.999
        assert first[1] - first[0] == 1 or (first[1] - first[0] - 1) * (R - 7) * Nit / SparseCoder.FUSED_PATCH_ITERS_PER_S \
            < SparseCoder.SHARED_START_SECONDS
