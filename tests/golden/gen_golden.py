"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own functions.

    python tests/golden/gen_golden.py        (build container only: needs /root/reference)

Every .npz holds seeded inputs plus what the unmodified reference code returned for them:
  eval_*.npz   pairwise_distances / csls_sim (src/utils.py:202-218, 417-435) and the Runner._test ranking
               loops (main.py:400-429, run verbatim with a stable sort) -> distance stats, nv1/nv2, ranks,
               top-3 ids, Hits@k / MR / MRR accumulators
  icl_*.npz    icl_loss.forward (model/SNAG_loss.py:58-128) loss + autograd gradients
  ial_*.npz    ial_loss.forward (model/SNAG_loss.py:148-202) loss + gradient
  noise_*.npz  SNAG.add_noise_to_embeddings (model/SNAG.py:66-75) with its RNG draws replayed
  mll_*.npz    CustomMultiLossLayer.forward (model/SNAG_loss.py:22-29)
  mining_*.npz SNAG.Iter_new_links (model/SNAG.py:192-208): argmin vectors of the distance matrix and the returned links
               for a refresh epoch and a filter epoch
Inputs of the evaluation fixtures are pre-rounded to bf16 so that the reference (fp32) and the tensor-core
path see identical values; fixtures whose ground-truth margins are within 2e-5 of a competitor are
rejected and re-seeded, so that ranks do not depend on the accumulation order of the dot products.
"""
from __future__ import annotations

import argparse
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from refshim import load_reference, reference_test_loops  # noqa: E402


def bf16_round_t(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def clustered(n, d, sigma, seed, n_centres=16):
    g = torch.Generator().manual_seed(seed)
    centres = torch.randn((n_centres, d), generator=g)
    assign = torch.randint(0, n_centres, (n,), generator=g)
    x = torch.randn((n, d), generator=g) + centres[assign]
    y = x + sigma * torch.randn((n, d), generator=g)
    x = bf16_round_t(torch.nn.functional.normalize(x))
    y = bf16_round_t(torch.nn.functional.normalize(y))
    return x, y


def ref_metrics(ranks, n, top_k=(1, 10, 50)):
    """main.py:380-384, 402-408, 430-436 verbatim accumulators."""
    acc = np.zeros((len(top_k)), dtype=np.float32)
    mean, mrr = 0.0, 0.0
    for rank in ranks:
        mean += (rank + 1)
        mrr += 1.0 / (rank + 1)
        for i in range(len(top_k)):
            if rank < top_k[i]:
                acc[i] += 1
    mean /= n
    mrr /= n
    for i in range(len(top_k)):
        acc[i] = round(acc[i] / n, 4)
    return acc, mean, mrr


def eval_fixture(ref, x, y, k, csls, require_margin=True):
    n = x.shape[0]
    distance = ref.utils.pairwise_distances(x, y)                                    # main.py:386
    out = {"x": x.numpy(), "y": y.numpy(), "k": np.int32(k), "csls": np.int32(csls)}
    out["d_diag"] = distance.diagonal().numpy().copy()
    out["d_checksum"] = np.float64(distance.double().sum().item())
    if csls:
        sim = 1 - distance
        out["nv1"] = torch.mean(torch.topk(sim, k)[0], 1).numpy()                    # src/utils.py:431
        out["nv2"] = torch.mean(torch.topk(sim.t(), k)[0], 1).numpy()                # src/utils.py:432
        distance = 1 - ref.utils.csls_sim(sim, k)                                    # main.py:393
    g = distance.diagonal()
    # margin between the ground truth and its nearest competitor in value, both directions
    off = distance.clone()
    off.fill_diagonal_(float("inf"))
    margin = min((off - g[:, None]).abs().min().item(), (off - g[None, :]).abs().min().item())
    out["margin"] = np.float64(margin)
    if require_margin and margin < 2e-5:
        return None
    l2r, r2l, top3 = reference_test_loops(distance)
    out["rank_l2r"] = np.asarray(l2r, np.int32)
    out["rank_r2l"] = np.asarray(r2l, np.int32)
    out["top3"] = np.asarray(top3, np.int32)
    out["g"] = g.numpy().copy()
    for name, ranks in (("l2r", l2r), ("r2l", r2l)):
        acc, mean, mrr = ref_metrics(ranks, n)
        out[f"acc_{name}"] = acc
        out[f"mr_{name}"] = np.float64(mean)
        out[f"mrr_{name}"] = np.float64(mrr)
    return out


def gen_eval(ref, outdir):
    specs = [("eval_n384_d96_k10", 384, 96, 10, True, 2.0), ("eval_n384_d96_k3", 384, 96, 3, True, 2.0),
             ("eval_n384_d96_nocsls", 384, 96, 10, False, 2.0), ("eval_n700_d320_k10", 700, 320, 10, True, 3.0),
             ("eval_n257_d64_k16", 257, 64, 16, True, 1.5)]
    for name, n, d, k, csls, sigma in specs:
        seed = 3408
        while True:
            x, y = clustered(n, d, sigma, seed)
            fx = eval_fixture(ref, x, y, k, csls)
            if fx is not None:
                break
            seed += 1
        fx["seed"] = np.int32(seed)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **fx)
        print(name, "seed", seed, "margin %.2e" % fx["margin"], "hits@1", fx["acc_l2r"][0], "mrr %.4f" % fx["mrr_l2r"])

    # forced ties: dyadic-rational rows (every product and partial sum exact in fp32), duplicated targets.
    g = torch.Generator().manual_seed(7)
    n, d = 96, 32
    x = torch.randint(-4, 5, (n, d), generator=g).float() / 8.0
    y = x.clone()
    y[1::3] = y[0:-1:3][: len(y[1::3])]          # every third target duplicates its predecessor -> exact ties
    y[5] = x[40]
    y[40] = x[5]
    for k, csls, name in ((4, True, "eval_ties_dyadic_k4"), (4, False, "eval_ties_dyadic_nocsls")):
        fx = eval_fixture(ref, x, y, k, csls, require_margin=False)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), **fx)
        print(name, "margin %.2e" % fx["margin"], "ties present:", bool(fx["margin"] == 0.0))


def gen_losses(ref, outdir):
    torch.manual_seed(3408)
    N, D, B = 300, 48, 64
    emb = torch.randn(N, D, requires_grad=True)
    links = np.stack([np.random.RandomState(1).permutation(N // 2)[:B],
                      N // 2 + np.random.RandomState(2).permutation(N // 2)[:B]], 1).astype(np.int32)
    wn = (torch.softmax(torch.randn(N, 4), 1) * 4)[:, 2].clone().requires_grad_(True)
    for tau in (0.1, 0.05):
        crit = ref.loss.icl_loss(tau=tau, ab_weight=0.5, n_view=2)
        for weighted in (False, True):
            emb.grad = None
            wn.grad = None
            loss = crit(emb, links, weight_norm=wn if weighted else None)
            loss.backward()
            name = f"icl_tau{tau}_{'w' if weighted else 'nw'}"
            np.savez_compressed(os.path.join(outdir, name + ".npz"), emb=emb.detach().numpy(), links=links,
                                weight_norm=wn.detach().numpy(), tau=np.float64(tau), ab_weight=np.float64(0.5),
                                weighted=np.int32(weighted), loss=np.float32(loss.item()),
                                grad_emb=emb.grad.numpy().copy(),
                                grad_w=(wn.grad.numpy().copy() if weighted else np.zeros(N, np.float32)))
            print(name, "loss %.6f" % loss.item())
    # a batch that is not a multiple of anything, alpha != 0.5
    crit = ref.loss.icl_loss(tau=0.1, ab_weight=0.3, n_view=2)
    emb2 = torch.randn(91, 20, requires_grad=True)
    links2 = np.stack([np.arange(0, 37), np.arange(40, 77)], 1).astype(np.int32)
    loss = crit(emb2, links2)
    loss.backward()
    np.savez_compressed(os.path.join(outdir, "icl_odd.npz"), emb=emb2.detach().numpy(), links=links2,
                        weight_norm=np.zeros(91, np.float32), tau=np.float64(0.1), ab_weight=np.float64(0.3),
                        weighted=np.int32(0), loss=np.float32(loss.item()), grad_emb=emb2.grad.numpy().copy(),
                        grad_w=np.zeros(91, np.float32))
    print("icl_odd loss %.6f" % loss.item())

    # IAL: structured inputs (clustered) and a small tau so that the KL is not ~1e-8
    g = torch.Generator().manual_seed(11)
    centres = torch.randn(8, D, generator=g)
    src = (centres[torch.randint(0, 8, (N,), generator=g)] + 0.7 * torch.randn(N, D, generator=g)).requires_grad_(True)
    tar = (src.detach() + 0.8 * torch.randn(N, D, generator=g))
    for tau, red in ((4.0, "mean"), (0.5, "mean"), (0.5, "sum")):
        crit = ref.loss.ial_loss(tau=tau, ab_weight=0.5, zoom=0.1, reduction=red)
        src.grad = None
        loss = crit(src, tar, links)
        loss.backward()
        name = f"ial_tau{tau}_{red}"
        np.savez_compressed(os.path.join(outdir, name + ".npz"), src=src.detach().numpy(), tar=tar.numpy(), links=links,
                            tau=np.float64(tau), ab_weight=np.float64(0.5), zoom=np.float64(0.1), reduction=red,
                            loss=np.float32(loss.item()), grad_src=src.grad.numpy().copy())
        print(name, "loss %.6e" % loss.item())

    mll = ref.loss.CustomMultiLossLayer(loss_num=6)
    with torch.no_grad():
        mll.log_vars.copy_(torch.tensor([0.1, -0.2, 0.0, 0.3, 0.05, -0.1]))
    ls = [torch.tensor(0.7), torch.tensor(1.3), 0, torch.tensor(0.2)]
    out = mll(ls)
    np.savez_compressed(os.path.join(outdir, "mll.npz"), log_vars=mll.log_vars.detach().numpy(),
                        losses=np.asarray([0.7, 1.3, 0.0, 0.2], np.float32), out=np.float32(out.item()))
    print("mll %.6f" % out.item())


def gen_noise(ref, outdir):
    import importlib
    ref_snag = importlib.import_module("model.SNAG")   # reference module, imported in place (model/__init__ shadows the name)
    N, F = 96, 40
    torch.manual_seed(5)
    x = torch.randn(N, F)
    mean, std = x.mean(0), x.std(0)
    for ratio, rho, name in ((0.2, 0.7, "noise_r0.2_m0.7"), (0.8, 0.2, "noise_r0.8_m0.2")):
        fake_self = types.SimpleNamespace(args=types.SimpleNamespace(mask_ratio=rho))
        torch.manual_seed(3408)
        out = ref_snag.SNAG.add_noise_to_embeddings(fake_self, x.clone(), mean, std, noise_ratio=ratio)
        torch.manual_seed(3408)                      # replay the reference's draws (SNAG.py:68,70)
        mask = torch.rand(N) < ratio
        z = torch.randn_like(x[mask])
        np.savez_compressed(os.path.join(outdir, name + ".npz"), x=x.numpy(), mean=mean.numpy(), std=std.numpy(),
                            mask=mask.numpy(), z=z.numpy(), noise_ratio=np.float64(ratio), mask_ratio=np.float64(rho),
                            out=out.numpy())
        print(name, "selected", int(mask.sum()))
    # entity blend (SNAG_tools.py:127-128) on plain tensors
    e = torch.randn(N, 24)
    noise = torch.randn(N, 24)
    mask = torch.rand(N) < 0.35
    rho = 0.7
    blended = e.clone()
    blended[mask] = (1.0 - rho * 0.5) * blended[mask] + rho * 0.5 * noise[mask]
    np.savez_compressed(os.path.join(outdir, "rowblend.npz"), e=e.numpy(), noise=noise.numpy(), mask=mask.numpy(),
                        mask_ratio=np.float64(rho), out=blended.numpy())
    print("rowblend selected", int(mask.sum()))


def gen_mining(ref, outdir):
    """SNAG.Iter_new_links called unbound on a stand-in `self` (it only reads self.args.semi_learn_step)."""
    import importlib
    ref_snag = importlib.import_module("model.SNAG")
    fake_self = types.SimpleNamespace(args=types.SimpleNamespace(semi_learn_step=5))

    def one(name, emb, left, right, require_margin=True):
        d = ref.utils.pairwise_distances(emb[left], emb[right])
        if require_margin:          # nearest and second nearest further apart than the accumulation-order noise
            s1 = torch.sort(d, 1)[0]
            s0 = torch.sort(d, 0)[0]
            if min((s1[:, 1] - s1[:, 0]).min().item(), (s0[1] - s0[0]).min().item()) < 2e-5:
                return False
        preds_l = torch.argmin(d, dim=1).numpy()
        preds_r = torch.argmin(d.t(), dim=1).numpy()
        links_a = ref_snag.SNAG.Iter_new_links(fake_self, 4, left, emb, right, new_links=[])        # (4+1) % 25 == 5: refresh
        prev = links_a[::2] + [(left[0], right[-1])]
        links_b = ref_snag.SNAG.Iter_new_links(fake_self, 9, left, emb, right, new_links=prev)      # filter epoch
        np.savez_compressed(os.path.join(outdir, name + ".npz"), emb=emb.numpy(), left=np.asarray(left, np.int64),
                            right=np.asarray(right, np.int64), preds_l=preds_l.astype(np.int64), preds_r=preds_r.astype(np.int64),
                            dmin_l=d.min(1)[0].numpy(), dmin_r=d.min(0)[0].numpy(),
                            links_refresh=np.asarray(links_a, np.int64).reshape(-1, 2), prev=np.asarray(prev, np.int64).reshape(-1, 2),
                            links_filter=np.asarray(links_b, np.int64).reshape(-1, 2))
        print(name, "mutual pairs", len(links_a), "after filter", len(links_b))
        return True

    for name, n_l, n_r, dim, sigma in (("mining_n300x280_d96", 300, 280, 96, 1.5), ("mining_n1100x900_d320", 1100, 900, 320, 2.5)):
        seed = 3408
        while True:
            g = torch.Generator().manual_seed(seed)
            n_ent = n_l + n_r + 57
            x, y = clustered(max(n_l, n_r), dim, sigma, seed)
            emb = bf16_round_t(torch.nn.functional.normalize(torch.randn((n_ent, dim), generator=g)))
            perm = torch.randperm(n_ent, generator=g)
            left, right = perm[:n_l].tolist(), perm[n_l:n_l + n_r].tolist()
            emb[left] = x[:n_l]
            emb[right] = y[:n_r]
            if one(name, emb, left, right):
                break
            seed += 1
    # exact ties: dyadic rows with duplicates on both sides -> argmin must return the first index
    g = torch.Generator().manual_seed(9)
    emb = torch.randint(-4, 5, (120, 32), generator=g).float() / 8.0
    left, right = list(range(0, 50)), list(range(60, 120))
    emb[61] = emb[60]
    emb[70] = emb[3]
    emb[71] = emb[3]
    emb[4] = emb[3]
    one("mining_ties_dyadic", emb, left, right, require_margin=False)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=HERE)
    ap.add_argument("--only", default="", help="comma separated subset of: eval,losses,noise,mining")
    args = ap.parse_args()
    torch.set_num_threads(8)
    ref = load_reference()
    only = [s for s in args.only.split(",") if s]
    for nm, fn in (("eval", gen_eval), ("losses", gen_losses), ("noise", gen_noise), ("mining", gen_mining)):
        if not only or nm in only:
            fn(ref, args.out)
