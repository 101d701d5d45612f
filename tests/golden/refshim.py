"""Import the UNMODIFIED reference (zjukg/SNAG, SNAG_MMEA) from /root/reference for golden-vector generation.

Only usable in the build container (the reference checkout does not travel to the GPU box). Nothing is
copied: the reference modules are imported in place with three shims so that they run on CPU tensors:
  - `easydict` and `unidecode` (absent from this image, needed only by config.py / torchlight) are stubbed;
  - torch.Tensor.cuda / nn.Module.cuda become the identity when no GPU is present (the losses hard-code
    `.cuda()`, model/SNAG_loss.py:90,96,165);
  - sys.dont_write_bytecode, because /root/reference is read-only.
"""
from __future__ import annotations

import sys
import types

REF_ROOT = "/root/reference/SNAG_MMEA"


def load_reference():
    import torch

    sys.dont_write_bytecode = True
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")

        class EasyDict(dict):
            __getattr__ = dict.get
            __setattr__ = dict.__setitem__

        m.EasyDict = EasyDict
        sys.modules["easydict"] = m
    if "unidecode" not in sys.modules:
        m = types.ModuleType("unidecode")
        m.unidecode = lambda s: s
        sys.modules["unidecode"] = m
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
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
