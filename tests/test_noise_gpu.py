"""GPU parity: Gauss modality noise masking, column statistics, entity-row blend — through the C ABI against the
reference's golden vectors (bit-exact with the reference's own draws injected) and the oracle."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from oracle import oracle
from snag_b200 import ops
from tests.conftest import golden_names, load_golden

pytestmark = pytest.mark.gpu


def _t(a, dev, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize("name", golden_names("noise_"))
def test_noise_mask_golden_bitwise(cuda_device, name):
    fx = load_golden(name)
    out = ops.noise_mask(_t(fx["x"], cuda_device), _t(fx["mean"], cuda_device), _t(fx["std"], cuda_device),
                         float(fx["noise_ratio"]), float(fx["mask_ratio"]),
                         mask=_t(fx["mask"].astype(np.uint8), cuda_device), zsel=_t(fx["z"], cuda_device))
    np.testing.assert_array_equal(out.cpu().numpy(), fx["out"])


def test_rowblend_golden_bitwise_and_gradient(cuda_device):
    fx = load_golden("rowblend")
    rho = float(fx["mask_ratio"])
    a, c = float(np.float32(1.0 - rho * 0.5)), float(np.float32(rho * 0.5))
    mask = _t(fx["mask"].astype(np.uint8), cuda_device)
    out = ops.rowblend_fwd(_t(fx["e"], cuda_device), _t(fx["noise"], cuda_device), mask, a, c)
    np.testing.assert_array_equal(out.cpu().numpy(), fx["out"])
    g = torch.randn(fx["e"].shape, device=cuda_device)
    gin = ops.rowblend_bwd(g, mask, a)
    ref = g.clone()
    ref[mask.bool()] = np.float32(a) * g[mask.bool()]
    assert torch.equal(gin, ref)


def test_noise_identities(cuda_device):
    rng = np.random.RandomState(0)
    x = _t(rng.randn(777, 1000).astype(np.float32), cuda_device)
    mean, std = ops.col_mean_std(x)
    none = torch.zeros(777, dtype=torch.uint8, device=cuda_device)
    allm = torch.ones(777, dtype=torch.uint8, device=cuda_device)
    z = torch.randn((777, 1000), device=cuda_device)
    assert torch.equal(ops.noise_mask(x, mean, std, 0.2, 0.7, mask=none, zsel=z[:1]), x)          # no row selected
    assert torch.equal(ops.noise_mask(x, mean, std, 0.2, 0.0, mask=allm, zsel=z), x)              # rho = 0
    out = ops.noise_mask(x, mean, torch.zeros_like(std), 0.2, 1.0, mask=allm, zsel=z)             # rho = 1, std = 0
    assert torch.equal(out, mean.expand_as(x).contiguous())
    assert torch.equal(ops.noise_mask(x, mean, std, 0.0, 0.7, seed=1), x)                         # ratio 0 (Philox)


@pytest.mark.parametrize("N,F", [(5000, 1000), (39594, 2048), (333, 300)])
def test_col_mean_std_vs_oracle(cuda_device, N, F):
    rng = np.random.RandomState(1)
    x = (rng.randn(N, F) * rng.rand(F) * 3 + rng.randn(F)).astype(np.float32)
    mean, std = ops.col_mean_std(_t(x, cuda_device))
    om, os_ = oracle.col_mean_std(x)
    np.testing.assert_allclose(mean.cpu().numpy(), om, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(std.cpu().numpy(), os_, rtol=1e-6, atol=1e-7)
    valid = rng.rand(N) < 0.85                                   # image statistics skip image-less entities
    mean, std = ops.col_mean_std(_t(x, cuda_device), _t(valid.astype(np.uint8), cuda_device))
    om, os_ = oracle.col_mean_std(x, valid)
    np.testing.assert_allclose(mean.cpu().numpy(), om, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(std.cpu().numpy(), os_, rtol=1e-6, atol=1e-7)


def test_philox_selection_matches_oracle_bitwise(cuda_device):
    for seed, ratio in ((3408, 0.2), (1 << 40 | 17, 0.1), (0, 0.8)):
        m = ops.philox_rowmask(39594, ratio, seed, cuda_device, row0=5)
        np.testing.assert_array_equal(m.cpu().numpy().astype(bool), oracle.philox_row_mask(seed, 39594, ratio, row0=5))


def test_philox_noise_statistics_and_sharding(cuda_device):
    """Production path (in-kernel Philox): selected-row fraction, standard-normal z, untouched rows bit-identical,
    and the output of a row shard equals the same rows of the unsharded call (counter = global element index)."""
    N, F, r, rho = 39594, 1000, 0.2, 0.7
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn((N, F), generator=g, device=cuda_device)
    mean, std = ops.col_mean_std(x)
    out = ops.noise_mask(x, mean, std, r, rho, seed=3408)
    sel = ops.philox_rowmask(N, r, 3408, cuda_device).bool()
    changed = (out != x).any(1)
    assert torch.equal(changed, sel)
    assert abs(sel.float().mean().item() - r) < 0.01
    assert torch.equal(out[~sel], x[~sel])
    z = ((out[sel] - np.float32(1.0 - rho) * x[sel]) / np.float32(rho) - mean) / std
    assert abs(z.mean().item()) < 2e-3 and abs(z.std().item() - 1) < 2e-3
    assert abs((z ** 3).mean().item()) < 1e-2 and abs((z ** 4).mean().item() - 3) < 3e-2
    assert abs(torch.corrcoef(torch.stack([z[:, 0], z[:, 1]]))[0, 1].item()) < 0.05
    r0, r1 = 12345, 23456
    part = ops.noise_mask(x[r0:r1].contiguous(), mean, std, r, rho, seed=3408, row0=r0)
    assert torch.equal(part, out[r0:r1])
    en = ops.gauss_fill(mean, std, N, 77)
    zz = (en - mean) / std
    assert abs(zz.mean().item()) < 2e-3 and abs(zz.std().item() - 1) < 2e-3
    assert torch.equal(ops.gauss_fill(mean, std, 1000, 77, row0=500), en[500:1500])


def test_odd_feature_width(cuda_device):
    """Feature widths that are not a multiple of 4 (attribute matrices of small datasets, src/data.py:507) go through the
    same float4 kernels, zero padded for the call: bit-exact against the oracle with injected draws."""
    rng = np.random.RandomState(4)
    N, F = 501, 997
    x = rng.randn(N, F).astype(np.float32)
    mean, std = oracle.col_mean_std(x)
    mask = rng.rand(N) < 0.3
    z = rng.randn(int(mask.sum()), F).astype(np.float32)
    out = ops.noise_mask(_t(x, cuda_device), _t(mean, cuda_device), _t(std, cuda_device), 0.3, 0.7,
                         mask=_t(mask.astype(np.uint8), cuda_device), zsel=_t(z, cuda_device))
    np.testing.assert_array_equal(out.cpu().numpy(), oracle.add_noise_to_embeddings(x, mean, std, mask, z, 0.7))
    prod = ops.noise_mask(_t(x, cuda_device), _t(mean, cuda_device), _t(std, cuda_device), 0.3, 0.7, seed=9)
    assert prod.shape == (N, F) and prod.is_contiguous()
    sel = (prod != _t(x, cuda_device)).any(1)
    assert abs(sel.float().mean().item() - 0.3) < 0.06
    en = ops.gauss_fill(_t(mean, cuda_device), _t(std, cuda_device), 2000, 5)
    assert en.shape == (2000, F) and abs(((en - _t(mean, cuda_device)) / _t(std, cuda_device)).std().item() - 1) < 0.01


def test_method_mirrors_on_a_model_shaped_object(cuda_device):
    """snag_b200.noise's drop-ins for SNAG.add_noise_to_embeddings / get_mean_std / update_noise (model/SNAG.py:66-98)
    and the differentiable entity blend (model/SNAG_tools.py:127-128), called the way the reference's class calls them
    (as methods, on the attributes its __init__ creates)."""
    import types
    from snag_b200 import noise
    g = torch.Generator(device="cuda").manual_seed(8)
    N = 6000
    rel = torch.poisson(torch.full((N, 1000), 0.05, device=cuda_device), generator=g)
    att = (torch.rand((N, 1000), generator=g, device=cuda_device) < 0.01).float()
    img = torch.nn.functional.normalize(torch.randn((N, 2048), generator=g, device=cuda_device))
    wo = torch.arange(0, N, 7, device=cuda_device)
    emb = torch.nn.Embedding(N, 300).to(cuda_device)
    torch.nn.init.normal_(emb.weight, std=1.0 / np.sqrt(N))
    me = types.SimpleNamespace(args=types.SimpleNamespace(noise_ratio=0.2, mask_ratio=0.7), rel_features=rel, att_features=att,
                               img_features=img, ent_wo_img=wo, multimodal_encoder=types.SimpleNamespace(entity_emb=emb))
    noise.get_mean_std(me)
    valid = torch.ones(N, dtype=torch.bool, device=cuda_device)
    valid[wo] = False
    np.testing.assert_allclose(me.img_mean.cpu().numpy(), img[valid].mean(0).cpu().numpy(), atol=1e-7)
    np.testing.assert_allclose(me.img_std.cpu().numpy(), img[valid].std(0).cpu().numpy(), rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(me.rel_std.cpu().numpy(), rel.std(0).cpu().numpy(), rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(me.att_mean.cpu().numpy(), att.mean(0).cpu().numpy(), atol=1e-7)
    torch.manual_seed(11)
    noise.update_noise(me)
    for name, feat in (("rel", rel), ("att", att), ("img", img)):
        noisy = getattr(me, f"{name}_noisy_features")
        assert noisy.data_ptr() != feat.data_ptr() and noisy.shape == feat.shape
        changed = (noisy != feat).any(1)
        assert abs(changed.float().mean().item() - 0.2) < 0.02
        assert torch.equal(noisy[~changed], feat[~changed])
    assert me.entity_noise.shape == (N, 300) and me.entity_noise_mask.dtype == torch.bool
    assert abs(me.entity_noise_mask.float().mean().item() - 0.1) < 0.015
    zn = (me.entity_noise - me.ent_mean) / me.ent_std
    assert abs(zn.mean().item()) < 5e-3 and abs(zn.std().item() - 1) < 5e-3
    torch.manual_seed(11)
    first = me.rel_noisy_features.clone()
    noise.update_noise(me)                                         # seeded from torch's CPU generator: reproducible
    assert torch.equal(first, me.rel_noisy_features)
    # add_noise_to_embeddings works in place on the clone it is handed and returns it (model/SNAG.py:74-75, :89)
    c = rel.clone()
    r = noise.add_noise_to_embeddings(me, c, me.rel_mean, me.rel_std, noise_ratio=0.5)
    assert r is c and abs((c != rel).any(1).float().mean().item() - 0.5) < 0.03
    # entity blend with autograd: forward and gradient equal the reference's in-place indexed blend
    e = emb(torch.arange(N, device=cuda_device))
    out = noise.blend_entity_noise(e, me.entity_noise, me.entity_noise_mask, 0.7)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    got = emb.weight.grad.clone()
    emb.weight.grad = None
    e2 = emb(torch.arange(N, device=cuda_device))
    m = me.entity_noise_mask
    e2[m] = (1.0 - 0.7 * 0.5) * e2[m] + 0.7 * 0.5 * me.entity_noise[m]
    (e2 * w).sum().backward()
    assert torch.equal(out, e2.detach()) and torch.equal(got, emb.weight.grad)
