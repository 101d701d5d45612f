"""INTEGRATION.md's table "inputs the reference accepts and what happens to them here", as tests: every input this build
refuses is refused LOUDLY with the documented exception before anything reaches the device — there is no fallback to
fall into. Runs on CPU (argument validation precedes every kernel call; a CPU tensor that does get as far as the C ABI
raises SnagError)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from snag_b200 import SnagError, evaluate
from snag_b200 import loss as sloss


def _batch(n=64, d=16, b=8):
    g = torch.Generator().manual_seed(0)
    emb = torch.randn((n, d), generator=g)
    links = np.stack([np.arange(b), n // 2 + np.arange(b)], 1).astype(np.int32)
    return emb, links


def test_icl_loss_explicit_negatives_and_inversion_raise():
    emb, links = _batch()
    with pytest.raises(NotImplementedError):                       # MEAformer replay only (MEAformer.py:126)
        sloss.icl_loss(tau=0.1)(emb, links, neg_l=np.arange(4), neg_r=np.arange(4))
    with pytest.raises(NotImplementedError):
        sloss.icl_loss(tau=0.1, inversion=True)(emb, links)
    with pytest.raises(RuntimeError):                              # the reference raises as well (model/SNAG_loss.py:84-89)
        sloss.icl_loss(tau=0.1, n_view=3)(emb, links)


@pytest.mark.parametrize("tau", [0.01, 1e-3, -0.1])
def test_icl_loss_refuses_temperatures_outside_the_kernels_range(tau):
    emb, links = _batch()
    with pytest.raises(ValueError, match="1/tau"):
        sloss.icl_loss(tau=tau)(emb, links)


def test_icl_loss_rejects_malformed_links():
    emb, _ = _batch()
    with pytest.raises(ValueError, match=r"\[B, 2\]"):
        sloss.icl_loss(tau=0.1)(emb, np.arange(12).reshape(4, 3))


def test_ial_loss_unsupported_modes_raise():
    emb, links = _batch()
    with pytest.raises(NotImplementedError):
        sloss.ial_loss(tau=4.0, ab_weight=0.5, zoom=0.1, reduction="mean")(emb, emb, links, norm=False)
    with pytest.raises(NotImplementedError):
        sloss.ial_loss(tau=4.0, ab_weight=0.5, zoom=0.1, reduction="batchmean")(emb, emb, links)
    with pytest.raises(NotImplementedError):
        sloss.ial_loss(tau=4.0, ab_weight=0.5, zoom=0.1, inversion=True)(emb, emb, links)


def test_anchor_shard_rejects_unknown_exchange():
    with pytest.raises(ValueError):
        sloss.AnchorShard(grads="reduce")


def test_evaluation_argument_errors():
    emb = torch.randn((32, 8))
    left, right = torch.arange(10), 16 + torch.arange(10)
    with pytest.raises(ValueError, match="pair up"):
        evaluate.evaluate_alignment(emb, left, right[:9])
    with pytest.raises(ValueError, match="csls_k"):
        evaluate.evaluate_alignment(emb, left, right, csls=True, csls_k=0)
    with pytest.raises(ValueError, match="exceeds"):              # torch.topk raises in the reference (src/utils.py:431)
        evaluate.evaluate_alignment(emb, left, right, csls=True, csls_k=11)
    with pytest.raises(ValueError, match="no test pairs"):
        evaluate.metrics_from_ranks(np.zeros((0,), dtype=np.int32))


def test_cpu_tensors_never_fall_back():
    """A call that passes validation with host tensors must fail at the C ABI, not compute something on the CPU."""
    emb, links = _batch()
    with pytest.raises(SnagError):
        sloss.icl_loss(tau=0.1)(emb, links)
    with pytest.raises(SnagError):
        evaluate.evaluate_alignment(emb, torch.arange(10), 32 + torch.arange(10), csls=True, csls_k=3)
