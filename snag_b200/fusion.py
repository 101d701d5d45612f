"""Fusion-output (joint) embeddings — SURVEY 8(f) rank 2: the tail of MformerFusion.forward
(SNAG_MMEA/model/SNAG_tools.py:44-49), which builds `joint_emb` (per-entity attention weights) and `joint_emb_fz`
(softmax of the six raw modality weights) as weighted concatenations of the L2-normalised modality embeddings.
The reference spends 2M F.normalize + 2M scalings + 2 torch.cat on it (and as many again backward); here one
bandwidth kernel reads every modality table once and writes both outputs, and one kernel does the backward.

`MformerFusion_forward` mirrors the reference method (same argument, same four return values); the transformer
layers it calls are the reference's own modules — only the tail is replaced. snag_b200.patch installs it.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import ops


class _JointFuse(torch.autograd.Function):
    """(w_ent [N, >=M], w_glob [>=M], e_0 .. e_{M-1} [N, d_m]) -> (joint, joint_fz), differentiable in all inputs."""

    @staticmethod
    def forward(ctx, w_ent, w_glob, *embs):
        embs = tuple(e.contiguous() for e in embs)
        w_ent, w_glob = w_ent.contiguous(), w_glob.contiguous()
        joint, fz = ops.joint_fuse_fwd(list(embs), w_ent, w_glob)
        ctx.save_for_backward(w_ent, w_glob, *embs)
        return joint, fz

    @staticmethod
    def backward(ctx, d_joint, d_fz):
        w_ent, w_glob, *embs = ctx.saved_tensors
        d_joint = None if d_joint is None else d_joint.contiguous().float()
        d_fz = None if d_fz is None else d_fz.contiguous().float()
        if d_joint is None and d_fz is None:
            return (None, None) + (None,) * len(embs)
        d_embs, d_w_ent, d_w_glob = ops.joint_fuse_bwd(list(embs), w_ent, w_glob, d_joint, d_fz)
        return (d_w_ent, d_w_glob, *d_embs)


def joint_embeddings(embs, weight_norm: torch.Tensor, weight_norm_fz: torch.Tensor):
    """model/SNAG_tools.py:44-49 for the present modalities `embs` (list of [N, d_m]):
    joint_emb = cat_m(weight_norm[:, m, None] * normalize(e_m)), joint_emb_fz = cat_m(weight_norm_fz[m] * normalize(e_m))."""
    return _JointFuse.apply(weight_norm.float(), weight_norm_fz.float(), *[e.float() for e in embs])


def MformerFusion_forward(self, embs):
    """Drop-in for MformerFusion.forward (model/SNAG_tools.py:32-51)."""
    embs = [embs[idx] for idx in range(len(embs)) if embs[idx] is not None]
    modal_num = len(embs)
    hidden_states = torch.stack(embs, dim=1)
    for layer_module in self.fusion_layer:                                    # the reference's own BertLayer stack
        layer_outputs = layer_module(hidden_states, output_attentions=True)
        hidden_states = layer_outputs[0]
    attention_pro = torch.sum(layer_outputs[1], dim=-3)
    attention_pro_comb = torch.sum(attention_pro, dim=-2) / math.sqrt(modal_num * self.args.num_attention_heads)
    weight_norm = F.softmax(attention_pro_comb, dim=-1)
    weight_norm_fz = F.softmax(self.weight_raw, dim=0)
    joint_emb, joint_emb_fz = joint_embeddings(embs, weight_norm, weight_norm_fz)
    return joint_emb, joint_emb_fz, hidden_states, weight_norm
