"""Property tests of the patch geometry (tier 1): for random (R, C, bb, stride) the host-side C entry points
and the oracle agree with each other and — when the reference checkout is present — with the reference's literal
get_image_block (index set, order, idx_Mat)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import lrs_pnp_dip_b200 as lrs
from oracle import lrs_oracle as orc
from oracle import ref_extract as rx

geoms = st.tuples(st.integers(1, 12), st.integers(1, 40), st.integers(0, 60), st.integers(0, 60)).map(
    lambda t: (t[0] + t[2], t[0] + t[3], t[0], t[1]))          # (R >= bb, C >= bb, bb, stride)


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(geoms)
def test_host_axis_starts_equal_oracle(g):
    R, C, bb, s = g
    rs, cs = lrs.patch_grid(R, C, bb, s)
    assert np.array_equal(rs, orc.axis_starts(R, bb, s)) and np.array_equal(cs, orc.axis_starts(C, bb, s))
    x, y = orc.patch_index(R, C, bb, s)
    assert len(x) == lrs.patch_count(R, C, bb, s) == len(rs) * len(cs)
    # coverage weight = number of selected windows over each element; windows stay inside the matrix
    W = orc.coverage_weight(R, C, bb, s)
    assert W.sum() == len(x) * bb * bb and x.max() + bb <= R and y.max() + bb <= C


@pytest.mark.skipif(not rx.reference_available(), reason="reference checkout not present")
@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(geoms)
def test_oracle_equals_literal_get_image_block(g):
    import torch

    R, C, bb, s = g
    gib = _literal()
    X = np.arange(R * C, dtype=np.float32).reshape(R, C)
    b_ref, x_ref, y_ref, idx_ref = gib(torch.tensor(X), bb, s)
    b, x, y, idx = orc.get_image_block(X, bb, s)
    assert np.array_equal(x, x_ref) and np.array_equal(y, y_ref)
    assert np.array_equal(b, b_ref.numpy()) and np.array_equal(idx, idx_ref.numpy())


_cache = {}


def _literal():
    if "gib" not in _cache:
        _cache["gib"] = rx.extract("main_LRS_PnP.py")["get_image_block"]
    return _cache["gib"]
