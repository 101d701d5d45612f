"""GPU parity of the unsupervised seed induction (src/data.py:367-402, src/utils.py:437-443): the
never-materialised global top-K of the similarity matrix against the oracle (bit-exact: entries, order, similarities)
and the drop-in visual_pivot_induction against the links the reference itself returned (tests/golden/seeds_*.npz)."""
from __future__ import annotations

import logging
import types

import numpy as np
import pytest
import torch

from oracle import oracle
from snag_b200 import seeds
from tests.conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names("seeds_"))
def test_seeds_golden(cuda_device, name):
    fx = load_golden(name)
    left, right, k = fx["left"].tolist(), fx["right"].tolist(), int(fx["unsup_k"])
    feats = torch.from_numpy(fx["feats"])
    rows, cols, sims = seeds.topk_similarity_entries(feats[left].to(cuda_device), feats[right].to(cuda_device), 100 * k)
    np.testing.assert_array_equal(rows.cpu().numpy(), fx["top_rows"])
    np.testing.assert_array_equal(cols.cpu().numpy(), fx["top_cols"])
    np.testing.assert_array_equal(sims.cpu().numpy(), fx["top_sims"])
    args = types.SimpleNamespace(unsup_k=k)
    ills = [tuple(t) for t in fx["ills"].tolist()]
    links = seeds.visual_pivot_induction(args, left, right, feats, ills, logging.getLogger("test_seeds"))
    assert links.dtype == np.int32
    np.testing.assert_array_equal(links, fx["links"])


@pytest.mark.parametrize("n_l,n_r,d,K", [(2000, 1500, 300, 5000), (700, 5000, 2048, 20000), (50, 40, 64, 2000), (33, 1, 8, 5)])
def test_topk_entries_vs_oracle(cuda_device, n_l, n_r, d, K):
    """K = 2000 of a 50 x 40 matrix asks for every entry (more than the 16-per-row pools hold); hubs (a few right-hand
    rows similar to many left-hand ones) put far more than 16 of the top K in single columns / rows."""
    rng = np.random.RandomState(n_l + n_r)
    base = rng.randn(max(n_l, n_r), d).astype(np.float32)
    x = base[:n_l] + 0.7 * rng.randn(n_l, d).astype(np.float32)
    y = base[:n_r] + 0.7 * rng.randn(n_r, d).astype(np.float32)
    if n_r > 3:
        x[: n_l // 2] += 3.0 * y[2]                      # a hub
    x = oracle.bf16_round(oracle.normalize_rows(x))
    y = oracle.bf16_round(oracle.normalize_rows(y))
    rows, cols, sims = seeds.topk_similarity_entries(torch.from_numpy(x).to(cuda_device), torch.from_numpy(y).to(cuda_device), K)
    orow, ocol, osim = oracle.topk_similarity_entries(x, y, K)
    np.testing.assert_array_equal(sims.cpu().numpy(), osim)
    np.testing.assert_array_equal(rows.cpu().numpy(), orow)
    np.testing.assert_array_equal(cols.cpu().numpy(), ocol)
