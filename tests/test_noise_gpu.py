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
