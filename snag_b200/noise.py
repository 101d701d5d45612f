"""Gauss modality noise masking — host-side mirror of the SNAG model's noise methods
(SNAG_MMEA/model/SNAG.py:66-98) and of the in-forward entity blend (SNAG_MMEA/model/SNAG_tools.py:122-129).

The functions below have the signatures of the reference's methods and are installed on the reference's SNAG
class by snag_b200.patch (they only use attributes the reference's __init__ creates). Differences from the
reference, none visible to its callers:
  - the per-row Bernoulli selection and the Gaussians are drawn in-kernel with a counter-based Philox generator
    (seeded from torch's CPU generator, so torch.manual_seed keeps runs reproducible) instead of a CPU
    torch.rand(N) + a data-dependent-shape CUDA randn; the output is a function of (seed, row, column) only and is
    therefore identical for any sharding of the rows;
  - update_noise writes the noisy copies out of place in ONE pass (read x, write x') instead of clone + bool-index
    gather + scatter; column statistics are single-pass fp64 reductions.
Bit parity with the reference's arithmetic is checked with the reference's own draws injected (ops.noise_mask with
mask= / zsel=, tests/test_noise_gpu.py)."""
from __future__ import annotations

import torch

from . import ops


def _next_seed() -> int:
    """62-bit seed taken from torch's default CPU generator — the generator the reference draws its row masks from."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def add_noise_to_embeddings(self, embeddings, mean, std, noise_ratio=0.1):
    """model/SNAG.py:66-75. Modifies `embeddings` in place (the reference passes a clone) and returns it."""
    x = embeddings if embeddings.is_contiguous() else embeddings.contiguous()
    out = ops.noise_mask(x.float() if x.dtype != torch.float32 else x, mean.float().contiguous(), std.float().contiguous(),
                         float(noise_ratio), float(self.args.mask_ratio), out=x if x.dtype == torch.float32 else None,
                         seed=_next_seed())
    if out.data_ptr() != embeddings.data_ptr():
        embeddings.copy_(out)
    return embeddings


def get_mean_std(self):
    """model/SNAG.py:77-84: column mean / unbiased std of the relation, attribute and image features; the image
    statistics only use entities that have an image (ent_wo_img excluded)."""
    n = self.img_features.size(0)
    valid = torch.ones((n,), dtype=torch.uint8, device=self.img_features.device)
    if self.ent_wo_img.numel() > 0:
        valid[self.ent_wo_img.long()] = 0
    self.img_mean, self.img_std = ops.col_mean_std(self.img_features.contiguous(), valid)
    self.rel_mean, self.rel_std = ops.col_mean_std(self.rel_features.contiguous())
    self.att_mean, self.att_std = ops.col_mean_std(self.att_features.contiguous())


def update_noise(self):
    """model/SNAG.py:86-98, once per epoch (main.py:253-254)."""
    r, rho = float(self.args.noise_ratio), float(self.args.mask_ratio)
    self.rel_noisy_features = ops.noise_mask(self.rel_features.contiguous(), self.rel_mean, self.rel_std, r, rho, seed=_next_seed())
    self.att_noisy_features = ops.noise_mask(self.att_features.contiguous(), self.att_mean, self.att_std, r, rho, seed=_next_seed())
    self.img_noisy_features = ops.noise_mask(self.img_features.contiguous(), self.img_mean, self.img_std, r, rho, seed=_next_seed())
    w = self.multimodal_encoder.entity_emb.weight.data
    self.ent_mean, self.ent_std = ops.col_mean_std(w.contiguous())
    self.entity_noise = ops.gauss_fill(self.ent_mean, self.ent_std, w.shape[0], _next_seed())
    self.entity_noise_mask = ops.philox_rowmask(w.shape[0], r * 0.5, _next_seed(), w.device).bool()


class _RowBlend(torch.autograd.Function):
    """e[mask] <- a*e[mask] + c*noise[mask] with the gradient of the reference's in-place indexed blend."""

    @staticmethod
    def forward(ctx, e, noise, mask_u8, a, c):
        ctx.save_for_backward(mask_u8)
        ctx.a = a
        return ops.rowblend_fwd(e.contiguous(), noise.contiguous(), mask_u8, a, c)

    @staticmethod
    def backward(ctx, g):
        (mask_u8,) = ctx.saved_tensors
        return ops.rowblend_bwd(g.contiguous(), mask_u8, ctx.a), None, None, None, None


def blend_entity_noise(entity_emb, entity_noise, entity_noise_mask, mask_ratio):
    """model/SNAG_tools.py:127-128 as a differentiable op."""
    import numpy as np
    a = float(np.float32(1.0 - mask_ratio * 0.5))
    c = float(np.float32(mask_ratio * 0.5))
    return _RowBlend.apply(entity_emb, entity_noise, entity_noise_mask.to(torch.uint8).contiguous(), a, c)


def encoder_forward(self, input_idx, adj, img_features=None, rel_features=None, att_features=None, name_features=None,
                    char_features=None, ent_wo_img=None, entity_noise=None, entity_noise_mask=None, _test=False):
    """MultiModalEncoder.forward (model/SNAG_tools.py:108-156) with the entity blend routed through the fused kernel;
    everything else calls the reference module's own sub-layers in the reference's order."""
    a = self.args
    gph_emb = img_emb = rel_emb = att_emb = name_emb = char_emb = None
    if a.w_gcn:
        ent = self.entity_emb(input_idx)
        if entity_noise is not None or entity_noise_mask is not None:
            ent = blend_entity_noise(ent, entity_noise, entity_noise_mask, a.mask_ratio)
        gph_emb = self.cross_graph_model(ent, adj)
    if a.w_img:
        img_emb = self.img_fc(img_features)
    if a.w_rel:
        rel_emb = self.rel_fc(rel_features)
    if a.w_attr:
        att_emb = self.att_fc(att_features)
    if a.w_name and name_features is not None:
        name_emb = self.name_fc(name_features)
    if a.w_char and char_features is not None:
        char_emb = self.char_fc(char_features)
    joint_emb, joint_emb_fz, hidden_states, weight_norm = self.fusion([img_emb, att_emb, rel_emb, gph_emb, name_emb, char_emb])
    return gph_emb, img_emb, rel_emb, att_emb, name_emb, char_emb, joint_emb, joint_emb_fz, hidden_states, weight_norm
