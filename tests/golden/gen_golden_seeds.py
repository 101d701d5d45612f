"""Golden vectors for the unsupervised seed induction: runs the REFERENCE's visual_pivot_induction
(SNAG_MMEA/src/data.py:367-402, unmodified, imported in place) on seeded, bf16-rounded, L2-normalised features.

    python tests/golden/gen_golden_seeds.py        (build container only: needs /root/reference)

The reference's torch.topk over the fp32 torch.mm leaves the order of (near-)equal entries to the library; a seed is
kept only if the reference's links equal those of the oracle's canonical ordering (fp64-accumulated similarities, ties
by flat index), so the fixture does not depend on that.
"""
from __future__ import annotations

import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from refshim import load_reference  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    load_reference()
    import importlib
    data = importlib.import_module("src.data")
    log = logging.getLogger("gen_golden_seeds")
    for name, n_ent, n_l, n_r, dim, k, noise in (("seeds_n300x280_d64_k20", 640, 300, 280, 64, 20, 0.6),
                                                  ("seeds_n900x1000_d128_k40", 2000, 900, 1000, 128, 40, 0.8)):
        seed = 3408
        while True:
            g = torch.Generator().manual_seed(seed)
            base = torch.randn((max(n_l, n_r), dim), generator=g)
            feats = torch.randn((n_ent, dim), generator=g)
            perm = torch.randperm(n_ent, generator=g)
            left, right = perm[:n_l].tolist(), perm[n_l:n_l + n_r].tolist()
            m = min(n_l, n_r)
            feats[left[:m]] = base[:m] + noise * torch.randn((m, dim), generator=g)
            feats[right[:m]] = base[:m] + noise * torch.randn((m, dim), generator=g)
            feats = torch.nn.functional.normalize(feats).to(torch.bfloat16).float()
            ills = [(left[i], right[i]) for i in range(m)]
            args = types.SimpleNamespace(unsup_k=k)
            links = data.visual_pivot_induction(args, left, right, feats, ills, log)
            ours = oracle.visual_pivot_induction(left, right, feats.numpy(), k)
            if links.shape == ours.shape and np.array_equal(links, ours):
                break
            seed += 1
        rows, cols, sims = oracle.topk_similarity_entries(feats.numpy()[left], feats.numpy()[right], k * 100)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), feats=feats.numpy(), left=np.asarray(left, np.int64),
                            right=np.asarray(right, np.int64), unsup_k=k, links=links, ills=np.asarray(ills, np.int64),
                            top_rows=rows.astype(np.int64), top_cols=cols.astype(np.int64), top_sims=sims, seed=seed)
        print(name, "seed", seed, "links", links.shape, "true", sum(1 for a, b in links.tolist() if (a, b) in set(ills)))


if __name__ == "__main__":
    main()
