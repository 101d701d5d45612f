"""Golden vectors for the fusion-output (joint) embeddings: runs the REFERENCE's MformerFusion
(SNAG_MMEA/model/SNAG_tools.py:22-51, unmodified, imported in place) on seeded modality embeddings and stores
what it returns plus the autograd gradients of a fixed scalar functional of (joint_emb, joint_emb_fz).

    python tests/golden/gen_golden_fusion.py        (build container only: needs /root/reference)

The attention weights `weight_norm` the reference's transformer layers produce are stored too: the drop-in tail
(snag_b200.fusion.joint_embeddings) is checked on exactly the weights the reference used, and
tests/test_patch.py runs the patched MformerFusion.forward against these outputs on CPU-shimmed modules.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from refshim import load_reference  # noqa: E402


def main():
    load_reference()
    import importlib
    tools = importlib.import_module("model.SNAG_tools")
    for name, N, M, D, heads, seed in (("fusion_m4_d64", 257, 4, 64, 4, 3408), ("fusion_m6_d96", 130, 6, 96, 2, 3409)):
        torch.manual_seed(seed)
        args = types.SimpleNamespace(num_hidden_layers=1, num_attention_heads=heads, hidden_size=D, intermediate_size=2 * D,
                                     use_intermediate=1)
        mod = tools.MformerFusion(args, modal_num=M, with_weight=1)
        mod.eval()
        with torch.no_grad():
            mod.weight_raw.copy_(torch.randn(6))
        embs = [(torch.randn(N, D) * (0.5 + m)).requires_grad_(True) for m in range(M)]
        embs[1].data[3].zero_()                                   # a zero row: F.normalize's eps clamp
        slots = list(embs) + [None] * (6 - M)                     # absent modalities are None (model/SNAG_tools.py:33)
        joint, joint_fz, hidden, weight_norm = mod(slots)
        c1 = torch.randn(joint.shape)
        c2 = torch.randn(joint_fz.shape)
        loss = (joint * c1).sum() + (joint_fz * c2).sum()
        grads = torch.autograd.grad(loss, embs + [mod.weight_raw], retain_graph=True)
        # gradient w.r.t. weight_norm as an independent leaf (what the fused backward returns for it)
        wn = weight_norm.detach().clone().requires_grad_(True)
        e2 = [e.detach().clone().requires_grad_(True) for e in embs]
        j2 = torch.cat([wn[:, m].unsqueeze(1) * torch.nn.functional.normalize(e2[m]) for m in range(M)], dim=1)
        wfz = torch.nn.functional.softmax(mod.weight_raw.detach(), dim=0).clone().requires_grad_(True)
        jf2 = torch.cat([wfz[m] * torch.nn.functional.normalize(e2[m]) for m in range(M)], dim=1)
        assert torch.equal(j2, joint.detach()) and torch.equal(jf2, joint_fz.detach())
        g2 = torch.autograd.grad((j2 * c1).sum() + (jf2 * c2).sum(), e2 + [wn, wfz])
        out = dict(N=N, M=M, D=D, heads=heads, seed=seed, weight_raw=mod.weight_raw.detach().numpy(),
                   weight_norm=weight_norm.detach().numpy(), weight_norm_fz=wfz.detach().numpy(),
                   joint=joint.detach().numpy(), joint_fz=joint_fz.detach().numpy(), hidden=hidden.detach().numpy(),
                   c1=c1.numpy(), c2=c2.numpy(), d_weight_norm=g2[M].numpy(), d_weight_norm_fz=g2[M + 1].numpy(),
                   d_weight_raw_full=grads[M].numpy())
        for m in range(M):
            out[f"emb{m}"] = embs[m].detach().numpy()
            out[f"d_emb{m}_tail"] = g2[m].numpy()                 # gradient through the tail only (weights held fixed)
        out["state"] = np.frombuffer(_state_bytes(mod), dtype=np.uint8)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "joint", tuple(joint.shape), "weight_norm row0", weight_norm[0].detach().numpy().round(3))


def _state_bytes(mod):
    import io
    buf = io.BytesIO()
    torch.save(mod.state_dict(), buf)
    return buf.getvalue()


if __name__ == "__main__":
    main()
