"""Golden vectors for --distance 1 (main.py:387-429), produced by the reference's own calls: scipy's
cdist(metric="cityblock") on the fp32 rows, torch.FloatTensor, the unmodified csls_sim, and the verbatim ranking loops
(stable sort). Run in the build container:  python tests/golden/gen_golden_l1.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import scipy.spatial.distance
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from refshim import load_reference, reference_test_loops  # noqa: E402


def main():
    ref = load_reference()
    rng = np.random.RandomState(3408)
    n, d = 220, 97
    centres = rng.randn(8, d).astype(np.float32)
    x = rng.randn(n, d).astype(np.float32) + centres[rng.randint(0, 8, n)]
    y = x + 0.6 * rng.randn(n, d).astype(np.float32)
    x[7], y[9] = x[6], y[8]                                   # duplicated rows: exact ties
    x = torch.nn.functional.normalize(torch.from_numpy(x)).numpy()
    y = torch.nn.functional.normalize(torch.from_numpy(y)).numpy()
    distance = torch.FloatTensor(scipy.spatial.distance.cdist(x, y, metric="cityblock"))      # main.py:388-390
    for k, csls in ((10, True), (3, False)):
        dist = 1 - ref.utils.csls_sim(1 - distance, k) if csls else distance                   # main.py:392-393
        l2r, r2l, top3 = reference_test_loops(dist)
        np.savez_compressed(os.path.join(HERE, f"l1_n{n}_d{d}_{'k%d' % k if csls else 'nocsls'}.npz"), x=x, y=y,
                            k=np.int32(k), csls=np.int32(csls), distance=distance.numpy(), dist=dist.numpy(),
                            rank_l2r=np.asarray(l2r, np.int32), rank_r2l=np.asarray(r2l, np.int32),
                            top3=np.asarray(top3, np.int32))


if __name__ == "__main__":
    main()
