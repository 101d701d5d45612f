"""Import the UNMODIFIED reference (zjukg/SNAG, SNAG_MMEA) from /root/reference for golden-vector generation.

The reference is imported in place (baseline/harness.py locates it: the byte-identical copy baseline/install_ref.py
puts under the git-ignored baseline/_ref/, or /root/reference in the build container) with three shims so that it
also runs on CPU tensors:
  - `easydict` and `unidecode` (absent from this image, needed only by config.py / torchlight) are stubbed;
  - torch.Tensor.cuda / nn.Module.cuda become the identity when no GPU is present (the losses hard-code
    `.cuda()`, model/SNAG_loss.py:90,96,165);
  - sys.dont_write_bytecode, because /root/reference is read-only.
"""
from __future__ import annotations

import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)
from baseline import harness  # noqa: E402

REF_ROOT = harness.ref_root()      # baseline/_ref/SNAG_MMEA (travels to the GPU box) or /root/reference/SNAG_MMEA


def load_reference():
    harness.load_reference()       # stubs for easydict / unidecode, .cuda() -> identity without a GPU, no bytecode
    import model.SNAG_loss as ref_loss   # noqa: E402
    import src.utils as ref_utils        # noqa: E402
    return types.SimpleNamespace(loss=ref_loss, utils=ref_utils)


def reference_test_loops(distance, top_k=(1, 10, 50)):
    """The two ranking loops of Runner._test (main.py:400-411, 422-429) run verbatim on a distance matrix, with
    torch.sort(stable=True) so that ties are defined. Returns the per-pair ranks and the top-3 ids per row."""
    import torch
    n = distance.shape[0]
    l2r, r2l, top3 = [], [], []
    for idx in range(n):
        values, indices = torch.sort(distance[idx, :], descending=False, stable=True)
        rank = (indices == idx).nonzero(as_tuple=False).squeeze().item()
        l2r.append(rank)
        top3.append(indices[:3].tolist())
    for idx in range(n):
        _, indices = torch.sort(distance[:, idx], descending=False, stable=True)
        rank = (indices == idx).nonzero(as_tuple=False).squeeze().item()
        r2l.append(rank)
    return l2r, r2l, top3
